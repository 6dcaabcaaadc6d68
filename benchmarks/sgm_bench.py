"""SGM timing at BASELINE.json configs[3]: 1920x1080, 128 planes, 8 paths,
one main / neighbour pair at full resolution (scale 0). Not part of the
bench.py contract; prints one JSON line with the device times of the three
kernels and their algorithmic HBM rates.

  python benchmarks/sgm_bench.py [--reference]   (--reference: time oracle/_ref too)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smvs_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--planes", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    sc = synth.make_scene(a.width, a.height, 1)
    M, t = synth.reprojection(sc, 0)
    M, t = M.astype(np.float32), t.astype(np.float32)
    dmin, dmax = float(sc.true_depth.min() * 0.7), float(sc.true_depth.max() * 1.3)
    ms = []
    for _ in range(a.reps + 2):
        r = api.sgm(sc.images[0], sc.images[1], M, t, dmin, dmax, a.planes)
        ms.append(r["ms"])
    ms = np.median(np.array(ms[2:]), axis=0)
    nvox = a.width * a.height * a.planes
    peak = 6567.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {
        "workload": f"SGM {a.width}x{a.height}x{a.planes}, 8 paths, P1=6 P2=96",
        "ms_cost_volume": float(ms[0]), "ms_aggregation": float(ms[1]),
        "ms_sum_wta": float(ms[2]), "ms_total": float(ms.sum()),
        "voxels": nvox,
        "cost_volume": {"bound": "integer ALU / shared memory", "bytes_per_voxel": 1,
                        "gb_s": nvox * 1 / ms[0] / 1e6},
        "aggregation": {"bound": "hbm", "bytes_per_voxel": 16,
                        "gb_s": nvox * 16 / ms[1] / 1e6, "frac": nvox * 16 / ms[1] / 1e6 / peak},
        "sum_wta": {"bound": "hbm", "bytes_per_voxel": 9,
                    "gb_s": nvox * 9 / ms[2] / 1e6, "frac": nvox * 9 / ms[2] / 1e6 / peak},
        "valid_fraction": float((r["depth"] > 0).mean()),
    }
    if a.reference:
        from oracle import ref as oref
        R = oref.RefScene(sc)
        rr = R.sgm_run(0, 1, 0, a.planes, dmin, dmax)
        out["reference_s_cost_agg_wta"] = [float(x) for x in rr["times"]]
        out["depth_equal_to_reference"] = bool(np.array_equal(rr["depth"], r["depth"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
