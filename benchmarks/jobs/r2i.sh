#!/bin/bash
python benchmarks/cg_probe.py new v1 > gpurun_out/r2i_probe.log 2>&1
python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/r2i_tests.log 2>&1
SMVSB_CG_TIMING=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_timing.json 2> gpurun_out/r2i_timing.err
python bench.py --steps 8 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -3 gpurun_out/r2i_tests.log; grep -v "^    iters" gpurun_out/r2i_probe.log | head -8
