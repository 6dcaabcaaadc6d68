#!/bin/bash
# r2am: the drop-in suite once more with the colour + shading case
O=gpurun_out/r2am; mkdir -p $O
timeout 1200 python -m pytest tests/test_integration.py -m gpu -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
