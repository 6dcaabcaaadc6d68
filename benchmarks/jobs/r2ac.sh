#!/bin/bash
# r2ac: colour views through the device path (set_scale, resident optimize, drop-ins)
O=gpurun_out/r2ac; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_topology.py tests/test_integration.py tests/test_gpu_visibility.py -m gpu -q -x -k "colour or set_scale or resident or optimize or without_sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
