"""The reference's whole DepthOptimizer::optimize() (coarse-to-fine ladder 5->2,
visibility, cutting, subdivision -- reference host code) at BASELINE.json
configs[1] size (1920x1080, 6 neighbours, -o2): the pure-CPU build against the
build whose members of INTEGRATION.md run on the GPU. Not part of the bench.py
contract.

  python benchmarks/optimize_e2e.py cpu [S]   here or anywhere (one host core,
                                              ~2.5 min) -> benchmarks/_cache/
  python benchmarks/optimize_e2e.py gpu [S]   on the GPU box; compares with the
                                              cached CPU result, prints JSON
  python benchmarks/optimize_e2e.py time [S]  on the GPU box: wall time of the
                                              drop-in optimize() only (no
                                              CPU result needed); with
                                              SMVSB_REBUILD_SURFACE=1 the host
                                              Surface is rebuilt as well
(S = with shading.) SMVSB_TIMING=1 adds the per-scale split of the patched
run_newton_iterations.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smvs_b200 import synth  # noqa: E402
from oracle import ref as oref  # noqa: E402

CACHE = os.path.join(ROOT, "benchmarks", "_cache")


def main():
    mode = sys.argv[1]
    shading = len(sys.argv) > 2 and sys.argv[2] == "S"
    tag = "S" if shading else "n"
    os.makedirs(CACHE, exist_ok=True)
    sc = synth.make_scene(1920, 1080, 6, seed_index=31, shading=shading)
    path = None if mode == "cpu" else oref.INTEGRATION_LIB_PATH
    all_secs = []
    for _ in range(1 if mode == "cpu" else 3):   # gpu: first call pays CUDA start-up
        R = oref.RefScene(sc, init_linear=shading, lib_path=path)
        t0 = time.time()
        depth, normals, _ = R.optimize(sc.init_depth, regularization=0.01, num_iterations=5,
                                       min_scale=2, use_shading=shading,
                                       debug_lvl=int(os.environ.get("SMVS_DEBUG", "0")))
        all_secs.append(time.time() - t0)
        R.close()
    secs = all_secs[-1]
    if mode == "time":
        print(json.dumps({"config": "1920x1080, 6 neighbours, -o2" + (" -S" if shading else ""),
                          "rebuild_surface": os.environ.get("SMVSB_REBUILD_SURFACE", "0") == "1",
                          "seconds_each_call": all_secs,
                          "valid_fraction": float((depth > 0).mean())}))
        return
    cache = os.path.join(CACHE, f"optimize_cpu_{tag}.npz")
    if mode == "cpu":
        np.savez_compressed(cache, depth=depth, normals=normals, secs=secs)
        print(json.dumps({"mode": "cpu", "seconds": secs}))
        return
    ref = np.load(cache)
    d_cpu = ref["depth"]
    m = (d_cpu > 0) & (depth > 0)
    rel = np.abs(depth[m] - d_cpu[m]) / d_cpu[m]
    out = {"config": "1920x1080, 6 neighbours, -o2" + (" -S" if shading else ""),
           "cpu_seconds": float(ref["secs"]), "gpu_build_seconds": secs,
           "gpu_build_seconds_each_call": all_secs,
           "members": os.environ.get("SMVSB_MEMBERWISE", "0") == "1" and "member-wise drop-ins"
           or "resident optimize() drop-in",
           "same_valid_mask": bool(np.array_equal(d_cpu > 0, depth > 0)),
           "valid_fraction": float(m.mean()),
           "depth_rel_linf": float(rel.max()),
           "normals_linf": float(np.abs(normals - ref["normals"])[m].max()),
           "mean_rel_error_vs_truth": float((np.abs(depth[m] - sc.true_depth[m])
                                             / sc.true_depth[m]).mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
