"""One-off: per-phase timing of the PCG kernel variants on the full 2 MP system."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from smvs_b200 import api
from smvs_b200.workload import build_workload

wl = build_workload(1920, 1080, 6, 2, shading=False, seed_index=0)
os.environ["SMVSB_CG_TIMING"] = "1"
with api.Context(0) as ctx:
    wl.push_views_u8(ctx)
    wl.push_surface(ctx)
    ctx.gn_construct(None, None, 0.01, 0.0)
    xs = {}
    for variant in sys.argv[1:] or ["new", "new"]:
        os.environ.pop("SMVSB_CG_VARIANT", None)
        os.environ.pop("SMVSB_CG_MODE", None)
        if variant.startswith("m"):
            os.environ["SMVSB_CG_MODE"] = variant[1:]
        elif variant != "new":
            os.environ["SMVSB_CG_VARIANT"] = variant
        sys.stderr.write(f"--- {variant}\n")
        sys.stderr.flush()
        for _ in range(3):
            t0 = time.perf_counter()
            it, info = ctx.cg_solve()
            dt = time.perf_counter() - t0
            sys.stderr.write(f"    iters {it} info {info} wall {dt*1e3:.2f} ms\n")
        xs[variant] = ctx.get_delta()
    if "new" in xs and "v1" in xs:
        a, b = xs["new"], xs["v1"]
        print("max |x_new - x_v1| / max|x|:", np.max(np.abs(a - b)) / np.max(np.abs(b)))
