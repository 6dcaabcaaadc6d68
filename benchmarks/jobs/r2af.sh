#!/bin/bash
# r2af: DepthOptimizer::optimize() with use_sgm = false resident on the device
O=gpurun_out/r2af; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_topology.py tests/test_integration.py -m gpu -q -s -k "without_sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -E "rel_median|normal_median|passed|failed|Error"  $O/pytest.log
