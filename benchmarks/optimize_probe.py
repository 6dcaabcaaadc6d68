"""One-off: where does the resident optimize() leave the reference's path?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from smvs_b200 import api, synth
from oracle import ref as oref

shading = len(sys.argv) > 1 and sys.argv[1] == "S"
sc = synth.make_scene(640, 480, 3, seed_index=90, shading=shading)
for min_scale in (4, 3, 2):
    for iters in (1, 5):
        R = oref.RefScene(sc, init_linear=shading)
        sgm = R.sgm_roundtrip(sc.init_depth)
        d_cpu, n_cpu, l_cpu = R.optimize(sc.init_depth, regularization=0.01,
                                         num_iterations=iters, min_scale=min_scale,
                                         use_shading=shading)
        Mi, ti = R.Mt()
        sh, shg = R.shading() if shading else (None, None)
        with api.Context(0) as ctx:
            d, n, light, st = api.optimize(ctx, sc.images[0], sc.images[1:], Mi, ti, R.flen(0),
                                           R.inverse_flen(0), R.inverse_calibration(), sgm,
                                           num_iterations=iters, min_scale=min_scale,
                                           shading=sh, shading_grad=shg)
        R.close()
        m = (d_cpu > 0) & (d > 0)
        rel = np.abs(d[m] - d_cpu[m]) / d_cpu[m]
        print(f"min_scale {min_scale} iters {iters}: mask equal {np.array_equal(d_cpu > 0, d > 0)} "
              f"(cpu {int((d_cpu > 0).sum())} gpu {int((d > 0).sum())}) depth rel max {rel.max():.3e} "
              f"normals max {np.abs(n - n_cpu)[m].max():.3e} stats {st}", flush=True)
