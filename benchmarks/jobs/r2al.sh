#!/bin/bash
# r2al: ncu evidence for the SGM kernels of the final library (launch list + full capture of the new sum / WTA kernel)
O=gpurun_out/r2al; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sgm_r2f.csv python benchmarks/sgm_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sgm_sum_wta128_kernel -s 1 -c 1 -o $O/prof_sgm_sum_wta128_r2f python benchmarks/sgm_bench.py > /dev/null 2>&1
ls -la $O
