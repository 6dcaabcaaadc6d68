#!/bin/bash
# r2w: e2e overheads (overlapped staging in set_views_u8, set_surface, pinned get_nodes); two views in flight at 1 CTA/SM
O=gpurun_out/r2w; mkdir -p $O
for k in 2 1; do
SMVSB_CG_CTAS_PER_SM=$k python bench.py --no-configs --no-cpu-baseline --steps 8 > $O/bench_c$k.json 2> $O/bench_c$k.err
python -c "
import json; d=json.loads(open('$O/bench_c$k.json').read().strip().splitlines()[-1]); print($k, d['value'], d['details']['ms_construct_solve_update'], d['e2e']['value'], d['e2e']['ms_set_views_set_surface_loop_get_nodes'], d['e2e']['two_host_threads_per_gpu']['value'])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_topology.py tests/test_gpu_visibility.py -m gpu -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
