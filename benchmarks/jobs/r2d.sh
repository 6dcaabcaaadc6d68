#!/bin/bash
# one-off GPU job: PCG kernel probe, SGM tests, fp64 peak, SGM timing
python benchmarks/cg_probe.py > gpurun_out/r2d_probe.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sgm or batch" > gpurun_out/r2d_sgm.log 2>&1
python -m pytest tests/test_gpu_fullsize.py tests/test_integration.py -m gpu -q -k "sgm" >> gpurun_out/r2d_sgm.log 2>&1
python - > gpurun_out/r2d_misc.log 2>&1 <<'PY'
from smvs_b200 import api
print("fp64 peak TFLOP/s", api.measure_fp64_peak(0))
import bench, argparse, json
a = argparse.Namespace(steps=6, no_cpu_baseline=True)
print(json.dumps(bench.sgm_config(a, api, 6567.7, "x")))
PY
tail -5 gpurun_out/r2d_sgm.log
