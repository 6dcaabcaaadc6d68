#!/bin/bash
# r2s: compute-sanitizer memcheck / racecheck on small cases of the round-2 kernels
O=gpurun_out/r2s; mkdir -p $O
CS="compute-sanitizer --error-exitcode 9 --print-limit 5"
$CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.log 2>&1; echo "smoke memcheck rc=$?"
$CS --tool memcheck python -m pytest tests/test_gpu_cutmaps.py tests/test_gpu_topology.py tests/test_gpu_visibility.py -m gpu -q -x \
    -k "3-160-120 or expand and 333 or without_sgm_parity and 256 or create_subdivide_fill and 333 or remove_isolated and 333 or visibility_and_cut_parity and 320-240-2-4" > $O/memcheck_tests.log 2>&1; echo "tests memcheck rc=$?"
$CS --tool racecheck python -c "
import numpy as np, os
from smvs_b200 import api
S = np.load('tests/golden/sgm.npz')
r = api.sgm(S['main'], S['neigh'], S['M'], S['t'], float(S['min_depth']), float(S['max_depth']), int(S['D']), volumes=True)
assert np.array_equal(r['sgm'], S['sgm'])
print('sgm ok')
" > $O/racecheck_sgm.log 2>&1; echo "sgm racecheck rc=$?"
tail -4 $O/memcheck_smoke.log; tail -4 $O/memcheck_tests.log; tail -4 $O/racecheck_sgm.log
