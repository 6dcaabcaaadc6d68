"""CPU sides of tests/test_gpu_fullsize.py: the compiled-verbatim reference
(oracle/_ref) on BASELINE.json's full-size configurations, one host core per
job, results cached under benchmarks/_cache/ (git-ignored; a cache made in
the dev container travels to the GPU box with the gpurun snapshot, otherwise
the test starts the jobs itself).

  python benchmarks/fullsize_cpu.py JOB [JOB ...]   |  all
  JOB: loop_n loop_S   inner Newton loop, 1920x1080, 6 neighbours, scale 2
       opt_n  opt_S    DepthOptimizer::optimize(), 1920x1080, 6 neighbours, -o2
       sgm             SGMStereo::run_sgm, 1920x1080, 128 planes

TEST INFRASTRUCTURE: this is the checker, never the product path.
"""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from smvs_b200 import synth  # noqa: E402
from smvs_b200.workload import build_workload  # noqa: E402

CACHE = os.path.join(ROOT, "benchmarks", "_cache")
JOBS = ("loop_n", "loop_S", "opt_n", "opt_S", "sgm")
VERSION = 3           # bump when inputs change: old caches are then ignored


def cache_path(job):
    return os.path.join(CACHE, f"fullsize_v{VERSION}_{job}.npz")


def loop_workload(shading):
    return build_workload(1920, 1080, 6, scale=2, shading=shading, seed_index=3)


def optimize_scene(shading):
    return synth.make_scene(1920, 1080, 6, seed_index=31, shading=shading)


def sgm_scene():
    sc = synth.make_scene(1920, 1080, 1, seed_index=9)
    dmin, dmax = float(sc.true_depth.min() * 0.7), float(sc.true_depth.max() * 1.3)
    return sc, dmin, dmax


def sgm_inputs():
    from oracle import ref as oref
    sc, dmin, dmax = sgm_scene()
    R = oref.RefScene(sc)
    M, t = R.reprojection(0, 1, 1920, 1080, 1920, 1080)
    R.close()
    return sc, dmin, dmax, M, t


def lighting_condition_number(wl, normals_i16):
    """cond_2 of the 16x16 normal matrix of LightOptimizer::fit_lighting_to_image
    (lib/light_optimizer.cc:22-55), from the reference's normal map."""
    from smvs_b200.synth import sh_basis
    n = normals_i16.astype(np.float64).reshape(-1, 3) / 32767.0
    keep = (np.abs(np.linalg.norm(n, axis=1) - 1.0) < 1e-3) \
        & (wl.shading.reshape(-1) >= 0.05)
    sh = sh_basis(n[keep])
    return float(np.linalg.cond(sh.T @ sh))


def volume_digest(vol):
    """(crc32 of the bytes, sum of the entries) of a uint16 volume."""
    v = np.ascontiguousarray(vol, dtype=np.uint16)
    return (int(zlib.crc32(memoryview(v).cast("B"))), int(v.sum(dtype=np.uint64)))


def run(job):
    from oracle import ref as oref
    os.makedirs(CACHE, exist_ok=True)
    t0 = time.time()
    if job in ("loop_n", "loop_S"):
        sys.path.insert(0, ROOT)
        from bench import _ref_scene_for
        shading = job == "loop_S"
        wl = loop_workload(shading)
        R = _ref_scene_for(wl)
        light = R.fit_lighting() if shading else None
        # the normal map the lighting was fitted to, as int16 (for cond(A))
        normals0 = np.round(R.surface_normals() * 32767.0).astype(np.int16) \
            if shading else np.zeros(1, np.int16)
        st = R.newton_loop(light, 0.01, 0.0)
        out = dict(depth=R.surface_depth(), nodes=R.surface_get()[0], normals0=normals0,
                   light=np.zeros(16) if light is None else light,
                   newton_steps=st["newton_steps"], cg_iterations=st["cg_iterations"],
                   pixel_iterations=st["pixel_iterations"], n_active=st["n_active"])
        R.close()
    elif job in ("opt_n", "opt_S"):
        shading = job == "opt_S"
        sc = optimize_scene(shading)
        R = oref.RefScene(sc, init_linear=shading)
        depth, normals, light = R.optimize(sc.init_depth, regularization=0.01,
                                           num_iterations=5, min_scale=2,
                                           use_shading=shading)
        R.close()
        out = dict(depth=depth, normals=normals, light=light)
    elif job == "sgm":
        sc, dmin, dmax = sgm_scene()
        R = oref.RefScene(sc)
        r = R.sgm_run(0, 1, 0, 128, dmin, dmax, volumes=True)
        R.close()
        out = dict(depth=r["depth"], cost_digest=np.array(volume_digest(r["cost"])),
                   sgm_digest=np.array(volume_digest(r["sgm"])))
    else:
        raise SystemExit(f"unknown job {job}")
    out["seconds"] = time.time() - t0
    tmp = cache_path(job) + ".tmp.npz"
    np.savez_compressed(tmp, **out)
    os.replace(tmp, cache_path(job))
    print(f"{job}: {out['seconds']:.1f} s -> {cache_path(job)}")


if __name__ == "__main__":
    jobs = sys.argv[1:]
    if jobs == ["all"] or not jobs:
        jobs = list(JOBS)
    for j in jobs:
        run(j)
