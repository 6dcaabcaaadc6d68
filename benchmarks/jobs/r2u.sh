#!/bin/bash
# r2u: K1 with Markstein division: parity + bench
O=gpurun_out/r2u; mkdir -p $O
python bench.py --no-configs --no-cpu-baseline --steps 8 > $O/bench.json 2> $O/bench.err
python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['details']['ms_construct_solve_update'], d['e2e']['value'])"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_topology.py tests/test_integration.py -m gpu -q -s -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3; grep -h '"job"' $O/pytest.log | cut -c1-400
