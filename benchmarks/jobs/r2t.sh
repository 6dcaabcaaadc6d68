#!/bin/bash
# r2t: K1 occupancy experiment (launch bounds 3 / 4 / 5 / 6 CTAs per SM)
O=gpurun_out/r2t; mkdir -p $O
for m in 4 3 5 6; do
  SMVSB_K1_MINB=$m python bench.py --no-configs --no-cpu-baseline --steps 8 > $O/bench_minb$m.json 2> $O/bench_minb$m.err
  python -c "
import json; d=json.loads(open('$O/bench_minb$m.json').read().strip().splitlines()[-1]); print($m, d['value'], d['details']['ms_construct_solve_update'])"
done
