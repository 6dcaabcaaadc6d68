#!/bin/bash
# r2n: expand / no-SGM visibility parity, drop-in member
O=gpurun_out/r2n; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_topology.py tests/test_gpu_visibility.py tests/test_integration.py -m gpu -q -x -k "expand or without_sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
