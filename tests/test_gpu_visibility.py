"""Visibility lists and boundary cutting on the device against the reference's
DepthOptimizer::create_subview_surfaces / cut_boundaries (compiled verbatim,
oracle/_ref): yes/no decisions, so everything is compared for EQUALITY."""
import numpy as np
import pytest

from smvs_b200 import api, synth
from oracle import ref as oref

from util_scene import colour_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built")]


def occluded_scene(width, height, n_sub, seed_index):
    """A scene whose initial depth has a raised block (depth discontinuities,
    wrong photometry at its rim) and whose SGM depth has a foreground disc
    and holes (occlusions in the neighbours' z-buffers)."""
    sc = synth.make_scene(width, height, n_sub, seed_index=seed_index)
    init = sc.init_depth.copy()
    init[height // 3:height // 2, width // 3:width // 2] *= 0.8
    yy, xx = np.mgrid[0:height, 0:width]
    sgm = sc.init_depth.copy()
    disc = (xx - 0.7 * width) ** 2 + (yy - 0.6 * height) ** 2 < (0.12 * height) ** 2
    sgm[disc] *= 0.6
    sgm[(xx + 2 * yy) % 17 == 0] = 0.0
    return sc, init.astype(np.float32), sgm.astype(np.float32)


def lists_of(off, ids, valid):
    return [tuple(ids[off[p]:off[p + 1]]) if valid[p] else () for p in range(len(valid))]


def run_pair(width, height, n_sub, scale, seed_index):
    sc, init, sgm = occluded_scene(width, height, n_sub, seed_index)
    R = oref.RefScene(sc)
    R.set_scale(scale)
    R.surface_create(scale, init)
    R.set_sgm_depth(sgm)
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    Mi, ti = R.Mt()
    with api.Context(0) as ctx:
        ctx.set_views(R.gradients(0), [R.gradients(k + 1) for k in range(n_sub)],
                      [R.hessian(k + 1) for k in range(n_sub)], Mi, ti,
                      R.flen(0), R.inverse_flen(0))
        ctx.set_surface(info["scale"], info["npx"], info["npy"], info["start_x"],
                        info["start_y"], nodes, nv, pv, None, None)
        # the depth map the z-buffer is filled from: bit-exact
        assert np.array_equal(ctx.get_depth(), R.surface_depth())

        left = R.create_subview_surfaces(True)
        removed = ctx.visibility(sgm)
        _, nv_r, pv_r = R.surface_get()
        off_r, ids_r = R.get_visibility()
        nv_g, pv_g, off_g, ids_g = ctx.surface_state()
        assert np.array_equal(pv_g, pv_r)
        assert np.array_equal(nv_g, nv_r)
        assert int(pv.sum()) - removed == left == int(pv_g.sum())
        lr, lg = lists_of(off_r, ids_r, pv_r), lists_of(off_g, ids_g, pv_g)
        assert lr == lg
        stats = dict(patches=int(pv.sum()), removed=int(removed),
                     partial=sum(1 for l in lg if 0 < len(l) < n_sub), cuts=[])

        K = R.inverse_calibration()
        for _ in range(12):
            d_r = R.cut_boundaries()
            d_g = ctx.cut_boundaries(K)
            _, nv_r, pv_r = R.surface_get()
            nv_g, pv_g, _, _ = ctx.surface_state()
            assert d_g == d_r
            assert np.array_equal(pv_g, pv_r)
            assert np.array_equal(nv_g, nv_r)
            stats["cuts"].append(d_r)
            if d_r <= 10:
                break
    R.close()
    return stats


@pytest.mark.parametrize("width,height,n_sub,scale,seed", [
    (320, 240, 3, 2, 41), (320, 240, 3, 3, 42), (320, 240, 2, 4, 43),
    (640, 480, 4, 5, 44), (320, 240, 3, 1, 45)])
def test_visibility_and_cut_parity(width, height, n_sub, scale, seed):
    st = run_pair(width, height, n_sub, scale, seed)
    # the scene must actually exercise the rules
    assert st["removed"] > 0 or st["partial"] > 0
    assert sum(st["cuts"]) > 0


@pytest.mark.parametrize("width,height,n_sub,scale,seed", [
    (256, 192, 3, 3, 63), (320, 240, 4, 2, 64), (333, 207, 3, 4, 65), (640, 480, 5, 2, 66)])
def test_visibility_without_sgm_parity(width, height, n_sub, scale, seed):
    """use_sgm = false: z-buffers from the surface's depth map only, and the
    NCC occlusion filter ncc_for_patch (lib/depth_optimizer.cc:795-912) with
    the rim carried over to the next neighbour's tests -- deleted patches,
    nodes and visibility lists EQUAL to the reference's."""
    col = colour_scene(width, height, n_sub, seed)
    init = col.init_depth.astype(np.float32).copy()
    init[height // 3:height // 2, width // 3:width // 2 + 20] *= 0.8
    init[height // 2 + 10:height // 2 + 50, width // 8:width // 4] *= 1.15
    R = oref.RefScene(col)
    R.set_scale(scale)
    R.surface_create(scale, init)
    info = R.surface_info()
    nodes, nv, pv = R.surface_get()
    Mi, ti = R.Mt()
    with api.Context(0) as ctx:
        ctx.set_views(R.gradients(0), [R.gradients(k + 1) for k in range(n_sub)],
                      [R.hessian(k + 1) for k in range(n_sub)], Mi, ti,
                      R.flen(0), R.inverse_flen(0))
        ctx.set_surface(info["scale"], info["npx"], info["npy"], info["start_x"],
                        info["start_y"], nodes, nv, pv, None, None)
        with pytest.raises(api.SmvsbError):
            ctx.visibility(None)                 # no colour images yet
        ctx.set_color_images(R.image(0), [R.image(k + 1) for k in range(n_sub)])
        left = R.create_subview_surfaces(False)
        removed = ctx.visibility(None)
        _, nv_r, pv_r = R.surface_get()
        off_r, ids_r = R.get_visibility()
        nv_g, pv_g, off_g, ids_g = ctx.surface_state()
        assert np.array_equal(pv_g, pv_r) and np.array_equal(nv_g, nv_r)
        assert int(pv.sum()) - removed == left == int(pv_g.sum())
        lr, lg = lists_of(off_r, ids_r, pv_r), lists_of(off_g, ids_g, pv_g)
        assert lr == lg
        # the filter must have mattered: with the SGM-mode tests alone (and the
        # same z-buffer) more neighbours would be listed
        ctx.set_surface(info["scale"], info["npx"], info["npy"], info["start_x"],
                        info["start_y"], nodes, nv, pv, None, None)
        ctx.visibility(np.zeros((height, width), np.float32))
        _, pv_s, off_s, ids_s = ctx.surface_state()
        n_ncc = sum(len(l) for l in lg)
        n_plain = sum(len(l) for l in lists_of(off_s, ids_s, pv_s))
        assert 0 < n_ncc < n_plain
    R.close()


def test_full_size_visibility_and_cut_parity():
    """BASELINE.json configs[1] size: 1920x1080, 6 neighbours, scale 2
    (128 104 patches, 12.3 M pixel x neighbour warps)."""
    from smvs_b200.workload import build_workload
    wl = build_workload(1920, 1080, 6, 2, shading=False, seed_index=3)
    sc = wl.scene
    yy, xx = np.mgrid[0:1080, 0:1920]
    sgm = sc.init_depth.astype(np.float32).copy()
    sgm[(xx - 1300) ** 2 + (yy - 600) ** 2 < 150 ** 2] *= 0.6
    R = oref.RefScene(sc)
    R.set_scale(2)
    R.surface_create(2, sc.init_depth)
    nodes = wl.nodes.copy().reshape(-1, 4)
    # a slightly raised block: its rim crosses the depth-discontinuity
    # threshold without tripping the anisotropy test first
    ns = wl.npx + 1
    for iy in range(80, 120):
        nodes[iy * ns + 150:iy * ns + 210, 0] *= 0.97
    R.surface_set(nodes.reshape(-1), wl.node_valid, wl.patch_valid)
    R.set_sgm_depth(sgm)
    with api.Context(0) as ctx:
        wl.push_views(ctx)
        ctx.set_surface(2, wl.npx, wl.npy, wl.start_x, wl.start_y, nodes.reshape(-1),
                        wl.node_valid, wl.patch_valid, None, None)
        assert np.array_equal(ctx.get_depth(), R.surface_depth())
        left = R.create_subview_surfaces(True)
        removed = ctx.visibility(sgm)
        _, nv_r, pv_r = R.surface_get()
        off_r, ids_r = R.get_visibility()
        nv_g, pv_g, off_g, ids_g = ctx.surface_state()
        assert removed > 100 and int(pv_g.sum()) == left
        assert np.array_equal(pv_g, pv_r) and np.array_equal(nv_g, nv_r)
        m = pv_r.astype(bool)
        assert np.array_equal(np.diff(off_g)[m], np.diff(off_r)[m])
        assert lists_of(off_r, ids_r, pv_r) == lists_of(off_g, ids_g, pv_g)
        K = R.inverse_calibration()
        total = 0
        for _ in range(12):
            d_r, d_g = R.cut_boundaries(), ctx.cut_boundaries(K)
            _, nv_r, pv_r = R.surface_get()
            nv_g, pv_g, _, _ = ctx.surface_state()
            assert d_g == d_r
            assert np.array_equal(pv_g, pv_r) and np.array_equal(nv_g, nv_r)
            total += d_r
            if d_r <= 10:
                break
        assert total > 50
    R.close()


def test_error_paths_of_the_new_entry_points():
    import ctypes as C
    L = api.lib()
    sc = synth.make_scene(96, 64, 1, seed_index=49)
    with api.Context(0) as ctx:
        # nothing set yet
        with pytest.raises(api.SmvsbError):
            ctx.visibility(sc.init_depth)
        with pytest.raises(api.SmvsbError):
            ctx.cut_boundaries(np.eye(3, dtype=np.float32))
        # images: too small / too many channels / kernel too large
        with pytest.raises(api.SmvsbError):
            ctx.view_set_scale(np.zeros((2, 2), dtype=np.float32), 2)
        with pytest.raises(api.SmvsbError):
            ctx.bilateral_filter(np.zeros((8, 8, 5), dtype=np.float32),
                                 np.ones((8, 8), dtype=np.float32))
        with pytest.raises(api.SmvsbError):
            ctx.bilateral_filter(np.zeros((8, 8), dtype=np.float32),
                                 np.ones((8, 8), dtype=np.float32), kernel_size=9)
        assert b"kernel_size" in L.smvsb_last_error(ctx._h)
        # colour entry points: 2 channels, no_sgm without colour / with a
        # depth image of the wrong size
        with pytest.raises(api.SmvsbError):
            ctx.view_set_scale(np.zeros((8, 8, 2), dtype=np.float32), 2)
        grey = [np.zeros((64, 96), dtype=np.uint8)] * 2
        rgb = [np.zeros((64, 96, 3), dtype=np.float32)] * 2
        args = (np.eye(3).reshape(1, 9), np.zeros((1, 3)), 96.0, 1.0 / 96.0,
                np.eye(3, dtype=np.float32))
        with pytest.raises(api.SmvsbError):
            api.optimize(ctx, grey[0], grey[1:], *args, np.ones((64, 96), np.float32),
                         use_sgm=False)
        assert b"three-channel" in L.smvsb_last_error(ctx._h)
        with pytest.raises(api.SmvsbError):
            api.optimize(ctx, rgb[0], rgb[1:], *args, np.ones((32, 48), np.float32),
                         use_sgm=False)
        # a filter over an empty depth map gives an empty depth map
        out = ctx.bilateral_filter(np.full((8, 8), 0.5, dtype=np.float32),
                                   np.zeros((8, 8), dtype=np.float32))
        assert not out.any()
