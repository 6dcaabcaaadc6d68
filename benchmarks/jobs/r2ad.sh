#!/bin/bash
# r2ad: final evidence of round 2 -- launch list and full captures of the two dominant kernels with the
# shipped library, then both bench arms exactly as the driver runs them
O=gpurun_out/r2ad; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/launches_r2f.csv python bench.py --steps 2 --warmup 3 \
    --no-cpu-baseline --no-configs > /dev/null 2>&1
for k in cg_kernel gn_patch_kernel; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 \
    -o $O/prof_${k}_r2f $B > /dev/null 2>&1
done
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
ls -la $O; cut -c1-400 $O/bench_ref.json; cut -c1-1200 $O/bench.json
