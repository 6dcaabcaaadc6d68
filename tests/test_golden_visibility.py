"""tests/golden/vis_s2.npz (made by tests/golden/make_golden.py from the
compiled-verbatim reference): visibility lists, patch / node deletions of the
cutting loop, the depth map, the joint bilateral filter and set_scale of one
small occluded scene.

  not gpu:  the oracle build still reproduces the fixture (pins oracle/_ref and
            its MVE shim against drift);
  gpu:      the CUDA path reproduces it through the C ABI, with no oracle in
            the loop -- everything for EQUALITY."""
import os
import sys

import numpy as np
import pytest

from smvs_b200 import api
from oracle import ref as oref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vis_s2.npz")


def lists_of(off, ids, valid):
    return [tuple(ids[off[p]:off[p + 1]]) if valid[p] else () for p in range(len(valid))]


@pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")
def test_oracle_reproduces_the_fixture():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_visibility import occluded_scene
    G = np.load(GOLD)
    sc, init, sgm = occluded_scene(int(G["w"]), int(G["h"]), int(G["n_sub"]), 14)
    assert np.array_equal(sgm, G["sgm"])
    R = oref.RefScene(sc)
    R.set_scale(int(G["scale"]))
    R.surface_create(int(G["scale"]), init)
    R.set_sgm_depth(sgm)
    nodes, nv, pv = R.surface_get()
    assert np.array_equal(nodes, G["nodes"]) and np.array_equal(pv, G["patch_valid"])
    assert np.array_equal(R.surface_depth(), G["depth_map"])
    assert np.array_equal(R.bilateral_filter(sgm), G["filtered"])
    assert R.create_subview_surfaces(True) == int(G["vis_left"])
    off, ids = R.get_visibility()
    assert np.array_equal(off, G["vis_off"]) and np.array_equal(ids, G["vis_ids"])
    for k, want in enumerate(G["cuts"]):
        assert R.cut_boundaries() == int(want)
        _, cnv, cpv = R.surface_get()
        assert np.array_equal(np.concatenate([cnv, cpv]), G["cut_states"][k])
    R.close()


@pytest.mark.gpu
def test_cuda_reproduces_the_fixture():
    G = np.load(GOLD)
    n = int(G["n_sub"])
    with api.Context(0) as ctx:
        ctx.set_views(G["main_grad"], [G[f"sub_grad{k}"] for k in range(n)],
                      [G[f"sub_hess{k}"] for k in range(n)], G["Mi"], G["ti"],
                      float(G["flen"]), float(G["inv_flen"]))
        ctx.set_surface(int(G["scale"]), int(G["npx"]), int(G["npy"]), int(G["start_x"]),
                        int(G["start_y"]), G["nodes"], G["node_valid"], G["patch_valid"],
                        None, None)
        assert np.array_equal(ctx.get_depth(), G["depth_map"])
        removed = ctx.visibility(G["sgm"])
        nv, pv, off, ids = ctx.surface_state()
        assert int(G["patch_valid"].sum()) - removed == int(G["vis_left"])
        assert np.array_equal(pv, G["vis_patch_valid"])
        assert np.array_equal(nv, G["vis_node_valid"])
        assert lists_of(off, ids, pv) == lists_of(G["vis_off"], G["vis_ids"],
                                                 G["vis_patch_valid"])
        for k, want in enumerate(G["cuts"]):
            assert ctx.cut_boundaries(G["inv_calib"]) == int(want)
            nv, pv, _, _ = ctx.surface_state()
            assert np.array_equal(np.concatenate([nv, pv]), G["cut_states"][k])
        assert np.array_equal(ctx.bilateral_filter(G["image"], G["sgm"]), G["filtered"])
        blur, grad, _ = ctx.view_set_scale(G["image"], int(G["scale"]))
        assert np.array_equal(blur, G["scaleimage"])
        assert np.array_equal(grad, G["main_grad"])
