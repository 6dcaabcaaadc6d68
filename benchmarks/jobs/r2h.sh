#!/bin/bash
python benchmarks/integration_probe.py > gpurun_out/r2h_int.log 2>&1
for v in new v1; do
  ncu --set full --clock-control none --import-source on -k regex:cg_kernel -s 1 -c 1 -o gpurun_out/prof_cg_${v}_r2h python benchmarks/cg_probe.py $v > /dev/null 2>&1
done
python -m pytest tests/test_gpu_topology.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2h_tests.log 2>&1
tail -3 gpurun_out/r2h_tests.log; cat gpurun_out/r2h_int.log
