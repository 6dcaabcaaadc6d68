"""Pins the plain restatement (oracle/oracle_port.cc) against the committed
golden fixtures, which are outputs of the compiled-verbatim reference
(tests/golden/make_golden.py), and against oracle/_ref run live."""
import os

import numpy as np
import pytest

from oracle import port as oport
from oracle import ref as oref
from smvs_b200 import synth

pytestmark = pytest.mark.skipif(not oport.available(), reason="oracle port not built")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def scene_from_golden(G):
    n = int(G["n_sub"])
    P = oport.PortScene(G["main_grad"], [G[f"sub_grad{k}"] for k in range(n)],
                        [G[f"sub_hess{k}"] for k in range(n)], G["Mi"], G["ti"],
                        float(G["flen"]), float(G["inv_flen"]),
                        G["shading"] if "shading" in G else None,
                        G["shading_grad"] if "shading_grad" in G else None)
    P.set_surface(int(G["scale"]), int(G["npx"]), int(G["npy"]), int(G["start_x"]),
                  int(G["start_y"]), G["nodes"], G["node_valid"], G["patch_valid"],
                  G["vis_off"], G["vis_ids"])
    return P


@pytest.mark.parametrize("fixture", ["gn_s2.npz", "gn_s4.npz"])
def test_port_matches_reference_fixtures(fixture):
    G = np.load(os.path.join(GOLD, fixture))
    P = scene_from_golden(G)
    for tag in G["variants"]:
        light = G["light"] if tag in ("lit", "litR") else None
        P.gn_construct(G[f"{tag}_active"], light, float(G["regularization"]),
                       float(G[f"{tag}_lreg"]))
        s = P.get_system()
        assert np.array_equal(s["Houter"], G[f"{tag}_Houter"])
        assert np.array_equal(s["Hinner"], G[f"{tag}_Hinner"])
        assert np.array_equal(s["Pinner"], G[f"{tag}_Pinner"])
        assert rel(s["g"], G[f"{tag}_g"]) < 1e-10
        assert rel(s["Hvals"], G[f"{tag}_Hvals"]) < 1e-10
        assert rel(s["Pvals"], G[f"{tag}_Pvals"]) < 1e-10
        x, it, info = P.cg_solve()
        assert it == int(G[f"{tag}_cg_iters"]) and info == int(G[f"{tag}_cg_info"])
        assert rel(x, G[f"{tag}_x"]) < 1e-6   # CG amplifies 1e-13 input noise
    P.gn_construct(G["full_active"], None, float(G["regularization"]), 0.0)
    x, _, _ = P.cg_solve()
    act, n_act, shift = P.update_nodes(x, G["full_active"])
    assert np.array_equal(act, G["upd_active"]) and n_act == int(G["upd_n_active"])
    assert abs(shift - float(G["upd_mean_shift"])) < 1e-9
    assert rel(P.get_nodes(), G["upd_nodes"]) < 1e-10
    P.close()


def test_port_newton_loop_fixture():
    G = np.load(os.path.join(GOLD, "gn_s4.npz"))
    P = scene_from_golden(G)
    st = P.newton_loop(None, float(G["regularization"]), 0.0)
    assert st["newton_steps"] == int(G["loop_newton_steps"])
    assert st["cg_iterations"] == int(G["loop_cg_iterations"])
    assert st["pixel_iterations"] == float(G["loop_pixel_iterations"])
    assert rel(P.get_nodes(), G["loop_nodes"]) < 1e-8
    P.close()


def test_port_sgm_fixture_bit_exact():
    G = np.load(os.path.join(GOLD, "sgm.npz"))
    r = oport.sgm(G["main"], G["neigh"], G["M"], G["t"], float(G["min_depth"]),
                  float(G["max_depth"]), int(G["D"]))
    assert np.array_equal(r["cost"], G["cost"].astype(np.uint16))
    assert np.array_equal(r["sgm"], G["sgm"])
    assert np.array_equal(r["depth"], G["depth"])


@pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")
def test_port_sgm_live_dark_regions():
    sc = synth.make_scene(120, 90, 1, seed_index=30)
    sc.images[0][20:40, 30:70] = 0
    sc.images[1][50:70, 10:60] = 10
    R = oref.RefScene(sc)
    r = R.sgm_run(0, 1, 0, 32, 0.05, 30.0, volumes=True)
    M, t = R.reprojection(0, 1, 120, 90, 120, 90)
    p = oport.sgm(sc.images[0], sc.images[1], M, t, 0.05, 30.0, 32)
    assert np.array_equal(p["cost"], r["cost"])
    assert np.array_equal(p["sgm"], r["sgm"])
    assert np.array_equal(p["depth"], r["depth"])
    R.close()


# ---------------------------------------------------------------------------
# visibility lists, boundary cutting, bilateral filter: the restatement in
# oracle_port.cc against the committed fixture (always) and the compiled
# reference (when built)
# ---------------------------------------------------------------------------

def _port_from_vis_fixture(G):
    n = int(G["n_sub"])
    P = oport.PortScene(G["main_grad"], [G[f"sub_grad{k}"] for k in range(n)],
                        [G[f"sub_hess{k}"] for k in range(n)], G["Mi"], G["ti"],
                        float(G["flen"]), float(G["inv_flen"]))
    P.set_surface(int(G["scale"]), int(G["npx"]), int(G["npy"]), int(G["start_x"]),
                  int(G["start_y"]), G["nodes"], G["node_valid"], G["patch_valid"],
                  None, None)
    return P


def test_port_visibility_and_cutting_match_the_fixture():
    G = np.load(os.path.join(GOLD, "vis_s2.npz"))
    P = _port_from_vis_fixture(G)
    assert np.array_equal(P.get_depth(int(G["h"]), int(G["w"])), G["depth_map"])
    removed = P.visibility(G["sgm"])
    nv, pv, off, ids = P.surface_state()
    assert int(G["patch_valid"].sum()) - removed == int(G["vis_left"])
    assert np.array_equal(pv, G["vis_patch_valid"]) and np.array_equal(nv, G["vis_node_valid"])
    assert np.array_equal(off, G["vis_off"]) and np.array_equal(ids, G["vis_ids"])
    for k, want in enumerate(G["cuts"]):
        assert P.cut_boundaries(G["inv_calib"]) == int(want)
        nv, pv, _, _ = P.surface_state()
        assert np.array_equal(np.concatenate([nv, pv]), G["cut_states"][k])
    assert np.array_equal(oport.bilateral_filter(G["image"], G["sgm"]), G["filtered"])


@pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("scale,seed", [(3, 61), (4, 62)])
def test_port_visibility_and_cutting_match_the_reference(scale, seed):
    from test_gpu_visibility import occluded_scene
    sc, init, sgm = occluded_scene(256, 192, 3, seed)
    R = oref.RefScene(sc)
    R.set_scale(scale)
    R.surface_create(scale, init)
    R.set_sgm_depth(sgm)
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    Mi, ti = R.Mt()
    P = oport.PortScene(R.gradients(0), [R.gradients(k + 1) for k in range(3)],
                        [R.hessian(k + 1) for k in range(3)], Mi, ti, R.flen(0),
                        R.inverse_flen(0))
    P.set_surface(info["scale"], info["npx"], info["npy"], info["start_x"],
                  info["start_y"], nodes, nv, pv, None, None)
    left = R.create_subview_surfaces(True)
    removed = P.visibility(sgm)
    assert int(pv.sum()) - removed == left
    _, nv_r, pv_r = R.surface_get()
    off_r, ids_r = R.get_visibility()
    nv_p, pv_p, off_p, ids_p = P.surface_state()
    assert np.array_equal(pv_p, pv_r) and np.array_equal(nv_p, nv_r)
    assert np.array_equal(off_p, off_r) and np.array_equal(ids_p, ids_r)
    K = R.inverse_calibration()
    for _ in range(12):
        d = R.cut_boundaries()
        assert P.cut_boundaries(K) == d
        _, nv_r, pv_r = R.surface_get()
        nv_p, pv_p, _, _ = P.surface_state()
        assert np.array_equal(pv_p, pv_r) and np.array_equal(nv_p, nv_r)
        if d <= 10:
            break
    R.close()


@pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")
def test_port_visibility_without_sgm_matches_the_reference():
    """use_sgm = false: depth-map pixels only in the z-buffer, NCC occlusion
    filter on the colour images (ncc_for_patch)."""
    import copy
    sc = synth.make_scene(256, 192, 3, seed_index=63)
    col = copy.copy(sc)
    col.images = [np.repeat(im[:, :, None], 3, axis=2) for im in sc.images]
    init = sc.init_depth.copy()
    init[60:100, 80:140] *= 0.8
    R = oref.RefScene(col)
    R.set_scale(3)
    R.surface_create(3, init)
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    Mi, ti = R.Mt()
    P = oport.PortScene(R.gradients(0), [R.gradients(k + 1) for k in range(3)],
                        [R.hessian(k + 1) for k in range(3)], Mi, ti, R.flen(0),
                        R.inverse_flen(0))
    P.set_surface(info["scale"], info["npx"], info["npy"], info["start_x"],
                  info["start_y"], nodes, nv, pv, None, None)
    P.set_images(R.image(0), [R.image(k + 1) for k in range(3)])
    left = R.create_subview_surfaces(False)
    removed = P.visibility(None)
    assert int(pv.sum()) - removed == left
    _, nv_r, pv_r = R.surface_get()
    off_r, ids_r = R.get_visibility()
    nv_p, pv_p, off_p, ids_p = P.surface_state()
    assert np.array_equal(pv_p, pv_r) and np.array_equal(nv_p, nv_r)
    assert np.array_equal(off_p, off_r) and np.array_equal(ids_p, ids_r)
    assert 0 < left < int(pv.sum())
    R.close()


@pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h,scale", [(256, 192, 4), (333, 207, 3), (160, 120, 2)])
def test_port_surface_topology_matches_the_reference(w, h, scale):
    """Surface::create from a depth map with holes, subdivide_patches,
    fill_patches_from_depth, remove_isolated_patches: same grids, nodes and
    flags as the reference after every operation."""
    sc = synth.make_scene(w, h, 1, seed_index=64)
    init = sc.init_depth.astype(np.float32).copy()
    yy, xx = np.mgrid[0:h, 0:w]
    init[(xx - 0.3 * w) ** 2 + (yy - 0.4 * h) ** 2 < (0.15 * h) ** 2] = 0.0
    init[:, int(0.7 * w):int(0.7 * w) + 3 * (1 << scale)] = 0.0
    R = oref.RefScene(sc)
    R.set_scale(scale)
    Mi, ti = R.Mt()
    P = oport.PortScene(R.gradients(0), [R.gradients(1)], [R.hessian(1)], Mi, ti,
                        R.flen(0), R.inverse_flen(0))

    def same():
        info = R.surface_info()
        assert {k: info[k] for k in P.info} == P.info
        nodes_r, nv_r, pv_r = R.surface_get()
        nv_p, pv_p, _, _ = P.surface_state()
        assert np.array_equal(pv_p, pv_r) and np.array_equal(nv_p, nv_r)
        m = np.repeat(nv_r.astype(bool), 4)
        assert np.array_equal(P.get_nodes().reshape(-1)[m], nodes_r.reshape(-1)[m])
        return int(pv_r.sum())

    R.surface_create(scale, init)
    P.surface_create(scale, init)
    n0 = same()
    assert 0 < n0 < P.n_patches
    R.surface_remove_isolated()
    P.surface_remove_isolated()
    same()
    assert R.surface_expand() == P.surface_expand() > 0
    same()
    for _ in range(2 if scale > 2 else 1):
        R.surface_subdivide()
        P.surface_subdivide()
        same()
        R.surface_fill_from_depth()
        P.surface_fill_from_depth()
        same()
        R.surface_remove_isolated()
        P.surface_remove_isolated()
        same()
    R.close()
