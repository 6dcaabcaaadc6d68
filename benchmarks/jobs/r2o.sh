#!/bin/bash
# r2o: SGM with the warped-volume split and two lines per warp: parity + timing + ncu
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_integration.py -m gpu -q -k "sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/sgm_bench.py > $O/sgm_bench.json 2> $O/sgm_bench.err
SMVSB_SGM_PATHS_1LINE=1 python benchmarks/sgm_bench.py > $O/sgm_bench_1line.json 2>> $O/sgm_bench.err
for k in sgm_warp_volume_kernel sgm_cost_kernel sgm_paths128_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 \
        -o gpurun_out/prof_${k}_r2o python benchmarks/sgm_bench.py > /dev/null 2>&1
done
tail -5 $O/pytest.log | cut -c1-300; cat $O/sgm_bench.json $O/sgm_bench_1line.json
