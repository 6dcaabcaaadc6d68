#!/bin/bash
# r2m: drop-in build tests, cut maps, A/B benches, ncu evidence
O=gpurun_out/r2m; mkdir -p $O
timeout 1200 python -m pytest tests/test_integration.py tests/test_gpu_cutmaps.py tests/test_gpu_fullsize.py -m gpu -q -s -k "not newton_loop and not sgm_bit" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/set_scale_bench.py > $O/set_scale.json 2> $O/set_scale.err
python benchmarks/optimize_resident.py > $O/optimize_resident.json 2> $O/optimize_resident.err
python benchmarks/optimize_resident.py --shading > $O/optimize_resident_S.json 2>> $O/optimize_resident.err
python benchmarks/optimize_e2e.py gpu > $O/optimize_e2e.json 2> $O/optimize_e2e.err
SMVSB_MEMBERWISE=1 python benchmarks/optimize_e2e.py gpu > $O/optimize_e2e_memberwise.json 2>> $O/optimize_e2e.err
python benchmarks/sgm_bench.py > $O/sgm_bench.json 2> $O/sgm_bench.err
bash profiles/capture.sh r2 > $O/capture.log 2>&1
tail -4 $O/pytest.log; cat $O/set_scale.json $O/optimize_resident.json $O/optimize_resident_S.json $O/optimize_e2e.json $O/optimize_e2e_memberwise.json; tail -30 $O/capture.log
