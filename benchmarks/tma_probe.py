"""One-off: the TMA-staged set_scale kernel against the three-kernel path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from smvs_b200 import api, synth, stereo_view

sc = synth.make_scene(352, 208, 1, seed_index=4)
img = sc.images[0]
for mode in ("", "1"):
    if mode:
        os.environ["SMVSB_TMA"] = "1"
    with api.Context(0) as ctx:
        f = stereo_view.byte_to_float(img)
        for scale in (2, 3):
            blur, grad, hess = ctx.view_set_scale(f, scale)
            rb, rg, rh = stereo_view.set_scale(img, scale)
            print("tma" if mode else "ref", scale, np.array_equal(blur, rb), np.array_equal(grad, rg),
                  np.array_equal(hess, rh), flush=True)
