#!/bin/bash
# r2ab: PCG grid barrier with release reduction / one acquire, partial sums in one round trip
O=gpurun_out/r2ab; mkdir -p $O
python benchmarks/cg_probe.py new > $O/cg_probe.log 2>&1; grep 'cg:' $O/cg_probe.log | tail -1
python bench.py --no-configs --no-cpu-baseline --steps 8 > $O/bench.json 2> $O/bench.err
python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['details']['ms_construct_solve_update'], d['e2e']['value'], d['e2e']['two_host_threads_per_gpu']['value'], d['roofline']['frac'])"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_integration.py -m gpu -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
