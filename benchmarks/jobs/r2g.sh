#!/bin/bash
python benchmarks/cg_probe.py new m1 m2 m3 v1 > gpurun_out/r2g_probe.log 2>&1
python benchmarks/optimize_probe.py > gpurun_out/r2g_opt.log 2>&1
python benchmarks/tma_probe.py > gpurun_out/r2g_tma.log 2>&1
SMVSB_TMA=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "set_scale" > gpurun_out/r2g_tma_tests.log 2>&1
python -m pytest tests/test_integration.py tests/test_gpu_topology.py -m gpu -q > gpurun_out/r2g_tests.log 2>&1
python benchmarks/sgm_bench.py > gpurun_out/r2g_sgm.json 2>&1
tail -3 gpurun_out/r2g_tests.log; tail -3 gpurun_out/r2g_tma.log
