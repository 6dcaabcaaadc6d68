#!/bin/bash
python benchmarks/cg_probe.py new m1 m3 v1 > gpurun_out/r2k_probe.log 2>&1
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python -m pytest tests/test_gpu_topology.py tests/test_integration.py -m gpu -q > gpurun_out/r2k_tests.log 2>&1
grep -v "^    iters" gpurun_out/r2k_probe.log | head -14; tail -3 gpurun_out/r2k_tests.log
