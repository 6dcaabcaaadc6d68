#!/bin/bash
python benchmarks/cg_probe.py new v1 new > gpurun_out/r2f_probe.log 2>&1
python -m pytest tests -m gpu -q -s --durations=10 > gpurun_out/r2f_tests.log 2>&1
SMVSB_CG_TIMING=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_timing.json 2> gpurun_out/r2f_timing.err
timeout 300 compute-sanitizer --tool memcheck python benchmarks/tma_probe.py > gpurun_out/r2f_tma.log 2>&1
tail -3 gpurun_out/r2f_tests.log
