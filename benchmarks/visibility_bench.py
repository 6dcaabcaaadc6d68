"""Visibility + boundary cutting at BASELINE.json configs[1] size: 1920x1080,
6 neighbours, scale 2 (128 104 patches). Not part of the bench.py contract;
prints one JSON line: device time through the C ABI (host surface in, flags
and lists out), the kernels alone, and -- with --reference -- the reference's
create_subview_surfaces + cut_boundaries loop on one host core.

  python benchmarks/visibility_bench.py [--reference] [--scale 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smvs_b200 import api  # noqa: E402
from smvs_b200.workload import build_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--neighbours", type=int, default=6)
    ap.add_argument("--scale", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    wl = build_workload(a.width, a.height, a.neighbours, a.scale, shading=False, seed_index=0)
    sc = wl.scene
    sgm = np.ascontiguousarray(sc.init_depth, dtype=np.float32)
    # inverse calibration of the main view as fill_inverse_calibration builds it
    # (square pixels, principal point in the centre)
    ax = np.float32(sc.flen[0]) * np.float32(max(a.width, a.height))
    K = np.array([1 / ax, 0, -np.float32(a.width) * np.float32(0.5) / ax,
                  0, 1 / ax, -np.float32(a.height) * np.float32(0.5) / ax,
                  0, 0, 1], dtype=np.float32)
    out = {"workload": f"{a.width}x{a.height}, {a.neighbours} neighbours, scale {a.scale}",
           "patches": int(wl.patch_valid.sum())}
    with api.Context(0) as ctx:
        wl.push_views(ctx)
        t_vis, t_cut, n_cut = [], [], []
        for _ in range(a.reps + 1):
            t0 = time.perf_counter()
            ctx.set_surface(a.scale, wl.npx, wl.npy, wl.start_x, wl.start_y, wl.nodes,
                            wl.node_valid, wl.patch_valid, None, None)
            removed = ctx.visibility(sgm)
            nv, pv, off, ids = ctx.surface_state()
            t1 = time.perf_counter()
            cuts = []
            while True:
                d = ctx.cut_boundaries(K)
                cuts.append(d)
                if d <= 10:
                    break
            nv, pv, _, _ = ctx.surface_state()
            t2 = time.perf_counter()
            t_vis.append(t1 - t0)
            t_cut.append(t2 - t1)
            n_cut.append(len(cuts))
        out.update({"removed_by_visibility": int(removed), "cut_rounds": n_cut[-1],
                    "patches_left": int(pv.sum()),
                    "visibility_entries": int(off[-1]),
                    "ms_visibility_abi": 1e3 * float(np.median(t_vis[1:])),
                    "ms_cut_loop_abi": 1e3 * float(np.median(t_cut[1:]))})
        # the two image-side members of the same drop-in: joint bilateral filter
        # of the SGM init and StereoView::set_scale of one view (host in / out)
        from smvs_b200 import stereo_view
        guide = stereo_view.byte_to_float(sc.images[0])
        t_bil, t_set = [], []
        for _ in range(a.reps + 1):
            t0 = time.perf_counter()
            filtered = ctx.bilateral_filter(guide, sgm)
            t1 = time.perf_counter()
            ctx.view_set_scale(guide, a.scale)
            t2 = time.perf_counter()
            t_bil.append(t1 - t0)
            t_set.append(t2 - t1)
        out.update({"ms_bilateral_filter_abi": 1e3 * float(np.median(t_bil[1:])),
                    "ms_view_set_scale_abi": 1e3 * float(np.median(t_set[1:]))})
    if a.reference:
        from oracle import ref as oref
        R = oref.RefScene(sc)
        R.set_scale(a.scale)
        R.surface_create(a.scale, sc.init_depth)
        R.surface_set(wl.nodes, wl.node_valid, wl.patch_valid)
        R.set_sgm_depth(sgm)
        out["same_inverse_calibration"] = bool(np.array_equal(R.inverse_calibration(), K))
        t0 = time.perf_counter()
        left = R.create_subview_surfaces(True)
        t1 = time.perf_counter()
        rounds = 0
        while True:
            rounds += 1
            if R.cut_boundaries() <= 10:
                break
        t2 = time.perf_counter()
        _, nv_r, pv_r = R.surface_get()
        out.update({"ref_ms_visibility": 1e3 * (t1 - t0), "ref_ms_cut_loop": 1e3 * (t2 - t1),
                    "ref_cut_rounds": rounds, "ref_cores": 1,
                    "same_patches_as_reference": bool(np.array_equal(pv_r, pv)),
                    "same_nodes_as_reference": bool(np.array_equal(nv_r, nv))})
        t0 = time.perf_counter()
        ref_filtered = R.bilateral_filter(sgm)
        t1 = time.perf_counter()
        R.set_scale(a.scale)
        t2 = time.perf_counter()
        out.update({"ref_ms_bilateral_filter": 1e3 * (t1 - t0),
                    "ref_ms_set_scale_per_view": 1e3 * (t2 - t1) / (1 + a.neighbours),
                    "same_filtered_depth_as_reference": bool(np.array_equal(ref_filtered, filtered))})
        R.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
