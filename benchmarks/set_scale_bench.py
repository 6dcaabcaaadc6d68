"""StereoView::set_scale of 7 views at 1920x1080 (byte images up, scale 2 / 3
/ 4) through smvsb_set_views_u8: the TMA-engine staged fused kernel
(cp.async.bulk + mbarrier, one pass over the image) against the three-kernel
path (SMVSB_NO_TMA=1: blur_x, blur_y, grad_hess, each a round trip through
L2/HBM). Prints one JSON line; outputs of the two paths are compared bitwise.

  python benchmarks/set_scale_bench.py [--reps 10]
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
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    out = {"workload": "set_scale, 7 views 1920x1080 u8", "ms_per_call": {}, "equal": {}}
    for scale in (2, 3, 4):
        wl = build_workload(1920, 1080, 6, scale, shading=False, seed_index=0)
        got = {}
        for mode in ("tma", "three_kernels"):
            if mode == "tma":
                os.environ.pop("SMVSB_NO_TMA", None)
            else:
                os.environ["SMVSB_NO_TMA"] = "1"
            with api.Context(0) as ctx:
                ts = []
                for _ in range(a.reps + 2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    wl.push_views_u8(ctx)
                    torch.cuda.synchronize()
                    ts.append((time.perf_counter() - t0) * 1e3)
                got[mode] = ctx.debug_get_view(1)
                out["ms_per_call"].setdefault(f"scale{scale}", {})[mode] = float(np.median(ts[2:]))
        out["equal"][f"scale{scale}"] = bool(all(np.array_equal(x, y) for x, y in
                                                 zip(got["tma"], got["three_kernels"])))
    os.environ.pop("SMVSB_NO_TMA", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
