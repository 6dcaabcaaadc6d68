#!/bin/bash
# Captures the ncu evidence kept under profiles/ (run on the GPU box through gpurun):
#   bash profiles/capture.sh r2
# writes gpurun_out/launches_<tag>.csv (launch list of the bench command) and one
# `--set full` report per hot kernel; summarise here with profiles/summarize.py.
tag=${1:-rX}
only=${2:-all}        # "vis": only the visibility / cutting kernels
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
mkdir -p gpurun_out
full() {   # full <kernel regex> <skip> <out name> <command...>
    k=$1; s=$2; o=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 \
        -o gpurun_out/prof_${o}_$tag "$@" > /dev/null 2>&1
}
if [ "$only" = all ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 \
    --no-cpu-baseline --no-configs > /dev/null 2>&1
for k in cg_kernel gn_patch_kernel gn_assemble_kernel reproj_kernel apply_delta_kernel; do
    full $k 0 $k $B
done
# StereoView::set_scale: the TMA-staged fused kernel and the three kernels it replaces
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_set_scale_$tag.csv python benchmarks/set_scale_bench.py --reps 1 > /dev/null 2>&1
full set_scale_tma_kernel 8 set_scale_tma python benchmarks/set_scale_bench.py --reps 1
SMVSB_NO_TMA=1 full blur_x_kernel 8 blur_x $B
SMVSB_NO_TMA=1 full grad_hess_kernel 8 grad_hess $B
for k in sgm_cost_kernel sgm_paths_kernel sgm_sum_wta_kernel sgm_consistency_kernel; do
    full $k 1 $k python benchmarks/sgm_bench.py
done
# the resident optimize(): surface topology kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_optimize_$tag.csv python benchmarks/optimize_resident.py --reps 1 > /dev/null 2>&1
fi
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_vis_$tag.csv python benchmarks/visibility_bench.py --reps 1 > /dev/null 2>&1
if [ "$only" = vis ]; then
for k in zbuf_scatter_kernel vis_patch_kernel cut_border_kernel; do
    full $k 1 $k python benchmarks/visibility_bench.py --reps 1
done
fi
ls -la gpurun_out | grep $tag
