"""DepthOptimizer::optimize() resident on the device (smvsb_optimize) at
BASELINE.json configs[1] size: 1920x1080, 6 neighbours, ladder 5 -> 2
(-o2), byte images and the SGM depth in, depth + normals out. Host wall
time of the C-ABI call (uploads, every kernel, downloads). Not part of the
bench.py contract; prints one JSON line.

  python benchmarks/optimize_resident.py [--reps 3] [--shading] [--colour]

--colour: the same scene as three-channel views (float RGB as
StereoView::get_image() holds it) through smvsb_optimize_rgb_f32.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smvs_b200 import api  # noqa: E402
from smvs_b200.workload import build_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--shading", action="store_true")
    ap.add_argument("--min-scale", type=int, default=2)
    ap.add_argument("--colour", action="store_true")
    a = ap.parse_args()
    w, h = 1920, 1080
    wl = build_workload(w, h, 6, a.min_scale, shading=a.shading, seed_index=0)
    sc = wl.scene
    ax = np.float32(sc.flen[0]) * np.float32(max(w, h))
    K = np.array([1 / ax, 0, -np.float32(w) * np.float32(0.5) / ax,
                  0, 1 / ax, -np.float32(h) * np.float32(0.5) / ax, 0, 0, 1], dtype=np.float32)
    sgm = np.ascontiguousarray(sc.init_depth, dtype=np.float32)
    images = sc.images
    if a.colour:
        rng = np.random.default_rng(0)
        images = []
        for im in sc.images:
            f = im.astype(np.float32)
            chans = [np.clip(f * g + o + rng.normal(0, 2.0, f.shape), 0, 255).astype(np.uint8)
                     for g, o in ((1.0, 0.0), (0.8, 20.0), (1.1, -10.0))]
            u8 = np.stack(chans, axis=2)
            images.append(np.clip(u8.astype(np.float32) / np.float32(255), 0, 1))
    ts, st = [], None
    with api.Context(0) as ctx:
        for _ in range(a.reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d, n, light, st = api.optimize(ctx, images[0], images[1:], wl.Mi, wl.ti,
                                           wl.flen_px, wl.inv_flen, K, sgm, num_iterations=5,
                                           min_scale=a.min_scale, shading=wl.shading,
                                           shading_grad=wl.shading_grad)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"workload": f"optimize() {w}x{h}, 6 neighbours, -o{a.min_scale}"
                                  + (" -S" if a.shading else "")
                                  + (", colour views" if a.colour else ""),
                      "ms_wall": ts, "ms_wall_median_warm": float(np.median(ts[1:])),
                      "valid_fraction": float((d > 0).mean()), "stats": st}))


if __name__ == "__main__":
    main()
