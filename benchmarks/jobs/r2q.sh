#!/bin/bash
# r2q: SGM census with comparison bits: parity + timing + ncu
O=gpurun_out/r2q; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_integration.py -m gpu -q -k "sgm" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python benchmarks/sgm_bench.py > $O/sgm_bench.json 2> $O/sgm_bench.err
SMVSB_SGM_COST_SUMS=1 python benchmarks/sgm_bench.py > $O/sgm_bench_sums.json 2>> $O/sgm_bench.err
ncu --set full --clock-control none --import-source on -k regex:sgm_cost_bits_kernel -s 1 -c 1 \
    -o gpurun_out/prof_sgm_cost_bits_kernel_r2q python benchmarks/sgm_bench.py > /dev/null 2>&1
tail -5 $O/pytest.log | cut -c1-300; cat $O/sgm_bench.json $O/sgm_bench_sums.json
