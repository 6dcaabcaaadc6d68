"""One-off: the drop-in optimize() (integration build) against the CPU build
and against smvsb_optimize called directly."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from smvs_b200 import api, synth
from oracle import ref as oref

sc = synth.make_scene(640, 480, 2, seed_index=21)
out = {}
for name, path in (("cpu", None), ("int", oref.INTEGRATION_LIB_PATH)):
    R = oref.RefScene(sc, lib_path=path)
    d, n, _ = R.optimize(sc.init_depth, regularization=0.01, num_iterations=5, min_scale=2)
    if name == "cpu":
        sgm = R.sgm_roundtrip(sc.init_depth)
        Mi, ti = R.Mt()
        args = (Mi, ti, R.flen(0), R.inverse_flen(0), R.inverse_calibration())
    R.close()
    out[name] = d
with api.Context(0) as ctx:
    d, n, light, st = api.optimize(ctx, sc.images[0], sc.images[1:], args[0], args[1], args[2],
                                   args[3], args[4], sgm)
out["dev"] = d
print(st)
for a, b in (("cpu", "int"), ("cpu", "dev"), ("int", "dev")):
    ma, mb = out[a] > 0, out[b] > 0
    both = ma & mb
    rel = np.abs(out[a][both] - out[b][both]) / out[a][both]
    diff = ma != mb
    ys, xs = np.nonzero(diff)
    box = (int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())) if diff.any() else None
    print(a, b, "valid", int(ma.sum()), int(mb.sum()), "mask diff px", int(diff.sum()), "bbox", box,
          "rel max", float(rel.max()))
