#!/bin/bash
python benchmarks/cg_probe.py new v1 new > gpurun_out/r2e_probe.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -s -k "not optimize" > gpurun_out/r2e_tests.log 2>&1
SMVSB_CG_TIMING=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_timing.json 2> gpurun_out/r2e_timing.err
for k in sgm_cost_kernel sgm_paths_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_${k}_r2e python benchmarks/sgm_bench.py > /dev/null 2>&1
done
tail -3 gpurun_out/r2e_tests.log
