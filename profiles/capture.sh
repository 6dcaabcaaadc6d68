#!/bin/bash
# Captures the ncu evidence kept under profiles/ (run on the GPU box through gpurun):
#   bash profiles/capture.sh r1c
# writes gpurun_out/launches_<tag>.csv (launch list of the bench command) and one
# `--set full` report per hot kernel; summarise here with profiles/summarize.py.
tag=${1:-rX}
only=${2:-all}        # "vis": only the visibility / cutting kernels
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
if [ "$only" = all ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for k in cg_kernel gn_patch_kernel gn_assemble_kernel reproj_kernel grad_hess_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 \
        -o gpurun_out/prof_${k}_$tag $B > /dev/null 2>&1
done
for k in sgm_cost_kernel sgm_paths_kernel sgm_sum_wta_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 \
        -o gpurun_out/prof_${k}_$tag python benchmarks/sgm_bench.py > /dev/null 2>&1
done
fi
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_vis_$tag.csv python benchmarks/visibility_bench.py --reps 1 > /dev/null 2>&1
for k in zbuf_scatter_kernel vis_patch_kernel cut_border_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 \
        -o gpurun_out/prof_${k}_$tag python benchmarks/visibility_bench.py --reps 1 > /dev/null 2>&1
done
ls -la gpurun_out | grep $tag
