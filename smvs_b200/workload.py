"""Synthetic Gauss-Newton workloads: everything smvsb_set_views /
smvsb_set_surface need for one reference view at one scale, built without
the oracle (numpy host mirrors of the reference's input producers), so that
the benchmark's product arm, its reference arm and the tests consume the
same bytes."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import stereo_view, synth


@dataclasses.dataclass
class Workload:
    scene: synth.Scene
    scale: int
    ps: int
    npx: int
    npy: int
    start_x: int
    start_y: int
    flen_px: float
    inv_flen: float
    main_grad: np.ndarray
    sub_grads: list
    sub_hess: list
    shading: np.ndarray | None
    shading_grad: np.ndarray | None
    Mi: np.ndarray
    ti: np.ndarray
    nodes: np.ndarray
    node_valid: np.ndarray
    patch_valid: np.ndarray
    vis_off: np.ndarray
    vis_ids: np.ndarray

    @property
    def n_patches_valid(self):
        return int(self.patch_valid.sum())

    @property
    def samples_per_patch(self):
        sampling = 4 if self.scale >= 5 else (2 if self.scale >= 3 else 1)
        return (self.ps // sampling) ** 2

    def h2d_bytes(self):
        n = self.main_grad.nbytes + sum(a.nbytes for a in self.sub_grads) \
            + sum(a.nbytes for a in self.sub_hess) + self.Mi.nbytes + self.ti.nbytes
        if self.shading is not None:
            n += self.shading.nbytes + self.shading_grad.nbytes
        n += self.nodes.nbytes + self.node_valid.nbytes + self.patch_valid.nbytes \
            + self.vis_off.nbytes + self.vis_ids.nbytes
        return int(n)

    def push_views(self, ctx):
        ctx.set_views(self.main_grad, self.sub_grads, self.sub_hess, self.Mi, self.ti,
                      self.flen_px, self.inv_flen, self.shading, self.shading_grad)

    def push_views_u8(self, ctx):
        """Byte images up, StereoView::set_scale on the device."""
        ctx.set_views_u8(self.scale, self.scene.images[0], self.scene.images[1:],
                         self.Mi, self.ti, self.flen_px, self.inv_flen,
                         with_shading=self.shading is not None)

    def h2d_bytes_u8(self):
        n = sum(im.nbytes for im in self.scene.images) + self.Mi.nbytes + self.ti.nbytes
        n += self.nodes.nbytes + self.node_valid.nbytes + self.patch_valid.nbytes \
            + self.vis_off.nbytes + self.vis_ids.nbytes
        return int(n)

    def push_surface(self, ctx):
        ctx.set_surface(self.scale, self.npx, self.npy, self.start_x, self.start_y,
                        self.nodes, self.node_valid, self.patch_valid, self.vis_off,
                        self.vis_ids)

    def push(self, ctx):
        self.push_views(ctx)
        self.push_surface(ctx)

    def restrict(self, x0, y0, nx, ny):
        """Copy with only the patches of the window [x0, x0+nx) x [y0, y0+ny)
        valid (bounded samples of the workload for the CPU arms)."""
        pv = np.zeros((self.npy, self.npx), dtype=np.uint8)
        pv[y0:y0 + ny, x0:x0 + nx] = self.patch_valid.reshape(self.npy, self.npx)[
            y0:y0 + ny, x0:x0 + nx]
        nv = _nodes_of_patches(pv)
        cnt = np.diff(self.vis_off.astype(np.int64))
        keep = np.repeat(pv.reshape(-1) != 0, cnt)
        cnt2 = np.where(pv.reshape(-1) != 0, cnt, 0)
        off = np.concatenate([[0], np.cumsum(cnt2)]).astype(np.uint32)
        return dataclasses.replace(self, patch_valid=pv.reshape(-1), node_valid=nv.reshape(-1),
                                   vis_off=off, vis_ids=self.vis_ids[keep].copy())


def _nodes_of_patches(pv2):
    npy, npx = pv2.shape
    nv = np.zeros((npy + 1, npx + 1), dtype=np.uint8)
    for dy in (0, 1):
        for dx in (0, 1):
            nv[dy:dy + npy, dx:dx + npx] |= pv2
    return nv


def _visibility(scene, Mi, ti, ps, npx, npy, sx, sy):
    """Patch p sees neighbour k if its corner and centre pixels (at the true
    depth) land inside the 3 % border the reference keeps
    (lib/depth_optimizer.cc:512-519). A simplified stand-in for
    create_subview_surfaces, used identically by both arms."""
    depth = scene._depth_fn
    iy, ix = np.mgrid[0:npy, 0:npx]
    x0 = (sx + ix * ps).astype(np.float64)
    y0 = (sy + iy * ps).astype(np.float64)
    vis = np.zeros((npy, npx, scene.n_sub), dtype=bool)
    for k in range(scene.n_sub):
        M, t = Mi[k].reshape(3, 3), ti[k]
        cut = 0.03 * max(scene.width, scene.height)
        ok = np.ones((npy, npx), dtype=bool)
        for (ox, oy) in ((0, 0), (ps - 1, 0), (0, ps - 1), (ps - 1, ps - 1),
                         (ps // 2, ps // 2)):
            u, v = x0 + ox + 0.5, y0 + oy + 0.5
            w = depth(u, v)
            p = M[0, 0] * u + M[0, 1] * v + M[0, 2]
            q = M[1, 0] * u + M[1, 1] * v + M[1, 2]
            r = M[2, 0] * u + M[2, 1] * v + M[2, 2]
            d = w * r + t[2]
            px = (w * p + t[0]) / d - 0.5
            py = (w * q + t[1]) / d - 0.5
            ok &= (px >= cut) & (px < scene.width - cut) & (py >= cut) \
                & (py < scene.height - cut)
        vis[:, :, k] = ok
    return vis


def build_workload(width, height, n_sub, scale=2, shading=False, seed_index=0,
                   init_noise=0.02, scene=None) -> Workload:
    sc = scene if scene is not None else synth.make_scene(
        width, height, n_sub, seed_index=seed_index, shading=shading,
        init_noise=init_noise)
    ps, npx, npy, sx, sy = synth.surface_grid(width, height, scale)
    _, main_grad, _ = stereo_view.set_scale(sc.images[0], scale)
    sub_grads, sub_hess = [], []
    for k in range(n_sub):
        _, g, h = stereo_view.set_scale(sc.images[k + 1], scale)
        sub_grads.append(g)
        sub_hess.append(h)
    sh = shg = None
    if shading:
        sh, shg = stereo_view.shading_inputs(sc.images[0])
    Mt = [synth.reprojection(sc, k) for k in range(n_sub)]
    Mi = np.array([m for m, _ in Mt], dtype=np.float64).reshape(n_sub, 9)
    ti = np.array([t for _, t in Mt], dtype=np.float64).reshape(n_sub, 3)
    ax = np.float32(sc.flen[0]) * np.float32(max(width, height))
    flen_px = float(ax)
    inv_flen = float(np.float32(1.0) / ax)

    vis = _visibility(sc, Mi, ti, ps, npx, npy, sx, sy)
    pv = vis.any(axis=2).astype(np.uint8)
    nv = _nodes_of_patches(pv)
    cnt = vis.sum(axis=2).reshape(-1)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    ids = np.nonzero(vis.reshape(-1, n_sub))[1].astype(np.uint8)
    nodes = synth.analytic_nodes(sc, scale, perturbed=True)
    return Workload(sc, scale, ps, npx, npy, sx, sy, flen_px, inv_flen, main_grad,
                    sub_grads, sub_hess, sh, shg, Mi, ti, nodes, nv.reshape(-1),
                    pv.reshape(-1), off, ids)
