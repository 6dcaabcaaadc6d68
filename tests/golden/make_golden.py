"""Generates the committed golden fixtures from the compiled-verbatim
reference (oracle/_ref, built from /root/reference by `make -C oracle ref`).

Run in the dev container:  python tests/golden/make_golden.py
Inputs come from the seeded generator smvs_b200/synth.py and from the
reference's own StereoView::set_scale / Surface::create / visibility code;
every expected array is an output of the reference's unmodified functions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref as oref          # noqa: E402
from smvs_b200 import synth             # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gn_fixture(name, w, h, n_sub, scale, shading, seed_index):
    sc = synth.make_scene(w, h, n_sub, seed_index=seed_index, shading=shading)
    R = oref.RefScene(sc, init_linear=shading)
    R.set_scale(scale)
    R.surface_create(scale, sc.init_depth)
    R.compute_visibility()
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    vo, vi = R.get_visibility()
    Mi, ti = R.Mt()
    d = dict(w=w, h=h, n_sub=n_sub, scale=scale,
             npx=info["npx"], npy=info["npy"], start_x=info["start_x"],
             start_y=info["start_y"], flen=R.flen(0), inv_flen=R.inverse_flen(0),
             main_grad=R.gradients(0), Mi=Mi, ti=ti, nodes=nodes, node_valid=nv,
             patch_valid=pv, vis_off=vo, vis_ids=vi, regularization=0.01)
    for k in range(n_sub):
        d[f"sub_grad{k}"] = R.gradients(k + 1)
        d[f"sub_hess{k}"] = R.hessian(k + 1)
    light = None
    if shading:
        img, grad = R.shading()
        d["shading"], d["shading_grad"] = img, grad
        light = R.fit_lighting()
        d["light"] = light

    rng = np.random.default_rng(seed_index + 7)
    active_full = nv.copy()
    active_part = (nv & (rng.random(nv.shape) < 0.4)).astype(np.uint8)
    variants = [("full", active_full, None, 0.0), ("part", active_part, None, 0.0)]
    if shading:
        variants += [("lit", active_full, light, 0.0), ("litR", active_full, light, 5.0)]
    d["variants"] = np.array([v[0] for v in variants])
    for tag, act, lt, lreg in variants:
        R.gn_construct(act, lt, 0.01, lreg)
        s = R.get_system()
        d[f"{tag}_active"] = act
        d[f"{tag}_lreg"] = lreg
        for k in ("g", "Hvals", "Houter", "Hinner", "Pvals", "Pouter", "Pinner"):
            d[f"{tag}_{k}"] = s[k]
        x, it, inf = R.cg_solve()
        d[f"{tag}_x"], d[f"{tag}_cg_iters"], d[f"{tag}_cg_info"] = x, it, inf

    # update step + full Newton loop from the initial surface (variant "full")
    R.gn_construct(active_full, None, 0.01, 0.0)
    x, _, _ = R.cg_solve()
    act, n_act, shift = R.update_nodes(x, active_full)
    d["upd_active"], d["upd_n_active"], d["upd_mean_shift"] = act, n_act, shift
    d["upd_nodes"] = R.surface_get()[0]
    R.surface_set(nodes, nv, pv)
    st = R.newton_loop(light, 0.01, 0.0)
    d["loop_newton_steps"] = st["newton_steps"]
    d["loop_cg_iterations"] = st["cg_iterations"]
    d["loop_pixel_iterations"] = st["pixel_iterations"]
    d["loop_n_active"] = st["n_active"]
    d["loop_nodes"] = R.surface_get()[0]
    d["loop_depth"] = R.surface_depth()
    d["loop_normals"] = R.surface_normals()
    np.savez_compressed(os.path.join(OUT, name), **d)
    R.close()
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items()
                 if k.startswith("loop_") and not hasattr(v, "shape")})


def sgm_fixture(name, w, h, D, seed_index):
    sc = synth.make_scene(w, h, 1, seed_index=seed_index)
    R = oref.RefScene(sc)
    dmin = float(sc.true_depth.min() * 0.7)
    dmax = float(sc.true_depth.max() * 1.3)
    r = R.sgm_run(0, 1, 0, D, dmin, dmax, volumes=True)
    M, t = R.reprojection(0, 1, w, h, w, h)
    np.savez_compressed(os.path.join(OUT, name), w=w, h=h, D=D, main=sc.images[0],
                        neigh=sc.images[1], M=M, t=t, min_depth=dmin, max_depth=dmax,
                        depth=r["depth"], cost=r["cost"].astype(np.uint8),
                        sgm=r["sgm"])
    R.close()
    print(name, "valid", float((r["depth"] > 0).mean()))


def vis_fixture(name, w, h, n_sub, scale, seed_index):
    """create_subview_surfaces (use_sgm) + the cut_boundaries loop + the joint
    bilateral filter + one set_scale, on a scene with occluders and a depth
    step (same recipe as tests/test_gpu_visibility.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_visibility import occluded_scene
    sc, init, sgm = occluded_scene(w, h, n_sub, seed_index)
    R = oref.RefScene(sc)
    R.set_scale(scale)
    R.surface_create(scale, init)
    R.set_sgm_depth(sgm)
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    Mi, ti = R.Mt()
    d = dict(w=w, h=h, n_sub=n_sub, scale=scale, npx=info["npx"], npy=info["npy"],
             start_x=info["start_x"], start_y=info["start_y"], flen=R.flen(0),
             inv_flen=R.inverse_flen(0), main_grad=R.gradients(0), Mi=Mi, ti=ti,
             nodes=nodes, node_valid=nv, patch_valid=pv, sgm=sgm,
             inv_calib=R.inverse_calibration(), depth_map=R.surface_depth(),
             image=R.image(0), scaleimage=R.scaleimage(0))
    for k in range(n_sub):
        d[f"sub_grad{k}"] = R.gradients(k + 1)
        d[f"sub_hess{k}"] = R.hessian(k + 1)
    d["filtered"] = R.bilateral_filter(sgm)
    d["vis_left"] = R.create_subview_surfaces(True)
    _, d["vis_node_valid"], d["vis_patch_valid"] = R.surface_get()
    d["vis_off"], d["vis_ids"] = R.get_visibility()
    cuts, states = [], []
    for _ in range(12):
        cuts.append(R.cut_boundaries())
        _, cnv, cpv = R.surface_get()
        states.append(np.concatenate([cnv, cpv]))
        if cuts[-1] <= 10:
            break
    d["cuts"] = np.array(cuts, dtype=np.int32)
    d["cut_states"] = np.stack(states)
    np.savez_compressed(os.path.join(OUT, name), **d)
    R.close()
    print(name, "patches", int(pv.sum()), "left", d["vis_left"], "cuts", cuts)


if __name__ == "__main__":
    gn_fixture("gn_s2.npz", 128, 96, 2, 2, True, 11)
    gn_fixture("gn_s4.npz", 256, 192, 2, 4, False, 12)
    sgm_fixture("sgm.npz", 96, 72, 64, 13)
    vis_fixture("vis_s2.npz", 160, 120, 2, 2, 14)
