#!/bin/bash
# r2ak: drop-in optimize(): maps held for a stand-in Surface (default) vs the host Surface rebuilt
O=gpurun_out/r2ak; mkdir -p $O
timeout 1200 python -m pytest tests/test_integration.py tests/test_gpu_topology.py -m gpu -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/optimize_e2e.py time > $O/held.json 2> $O/held.err
SMVSB_REBUILD_SURFACE=1 python benchmarks/optimize_e2e.py time > $O/rebuilt.json 2> $O/rebuilt.err
tail -4 $O/pytest.log; cat $O/held.json $O/rebuilt.json; tail -2 $O/held.err
