#!/bin/bash
# r2ai: SGM sum + WTA with 16-byte loads (8 lanes per pixel); resident optimize() grey vs colour at 1080p
O=gpurun_out/r2ai; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_integration.py -m gpu -q -x -k "sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/sgm_bench.py > $O/sgm.json 2> $O/sgm.err
SMVSB_SGM_WTA_BYTES=1 python benchmarks/sgm_bench.py > $O/sgm_bytes.json 2> $O/sgm_bytes.err
python benchmarks/optimize_resident.py --reps 3 > $O/grey.json 2> $O/grey.err
python benchmarks/optimize_resident.py --reps 3 --colour > $O/colour.json 2> $O/colour.err
tail -3 $O/pytest.log; cut -c1-600 $O/sgm.json; cut -c1-600 $O/sgm_bytes.json; cut -c1-500 $O/grey.json; cut -c1-500 $O/colour.json; tail -2 $O/colour.err
