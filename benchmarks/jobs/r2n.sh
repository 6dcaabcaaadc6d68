#!/bin/bash
# r2n: expand / no-SGM visibility parity, drop-in member, warp-per-item kernels at coarse scales
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_topology.py tests/test_gpu_visibility.py tests/test_integration.py tests/test_gpu_fullsize.py -m gpu -q -k "not newton_loop and not sgm_bit" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/optimize_resident.py > $O/optimize_resident.json 2> $O/optimize_resident.err
python benchmarks/optimize_e2e.py gpu > $O/optimize_e2e.json 2> $O/optimize_e2e.err
tail -25 $O/pytest.log | cut -c1-300; cat $O/optimize_resident.json $O/optimize_e2e.json
