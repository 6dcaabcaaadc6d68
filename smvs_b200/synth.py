"""Seeded synthetic MVE scenes.

One reference view with identity pose plus n neighbours on a circle around
its optical centre (the recipe of the reference's tests/test_optimization.cc:
44-63, generalised), a smooth analytic depth field, a band-limited texture
living on the surface, optional Lambertian shading under a fixed 16-term SH
light. Every neighbour image is rendered by inverting the exact warp, so the
true depth is a fixed point of the photometric energy.

This module is input generation only (numpy); it is shared by the tests, the
benchmark's product arm and its reference arm, so all of them see identical
bytes. It does not depend on oracle/ or on the CUDA library.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

BASE_SEED = 20160916

SH_LIGHT = np.array([1.0, 0.3, 0.5, -0.2, 0.1, 0.05, 0.1, -0.05, 0.02,
                     0, 0, 0, 0, 0, 0, 0], dtype=np.float64)


@dataclasses.dataclass
class Scene:
    width: int
    height: int
    n_sub: int
    flen: np.ndarray          # (1+n,) float32 normalised focal lengths
    rot: np.ndarray           # (1+n, 9) float32 world-to-camera rotations
    trans: np.ndarray         # (1+n, 3) float32 translations
    images: list              # 1+n uint8 arrays (H, W)
    true_depth: np.ndarray    # (H, W) float32 z-depth of the main view
    init_depth: np.ndarray    # (H, W) float32 perturbed depth
    seed: int
    shading: bool

    @property
    def flen_px(self) -> float:
        return float(np.float32(self.flen[0]) * np.float32(max(self.width, self.height)))


def _rodrigues(k, theta):
    k = np.asarray(k, dtype=np.float64)
    k = k / np.linalg.norm(k)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(theta) * K + (1 - math.cos(theta)) * (K @ K)


def make_cameras(n_sub, radius=0.3, inward_deg=3.0, phase=0.35):
    """Main camera = identity; neighbour k sits at radius*(cos a, sin a, 0)
    and is rotated inward_deg towards the main optical axis."""
    rot = [np.eye(3)]
    trans = [np.zeros(3)]
    for k in range(n_sub):
        a = phase + 2.0 * math.pi * k / max(n_sub, 1)
        c = radius * np.array([math.cos(a), math.sin(a), 0.0])
        chat = c / np.linalg.norm(c)
        axis = np.array([chat[1], -chat[0], 0.0])
        Rt = _rodrigues(axis, math.radians(inward_deg))   # camera-to-world
        R = Rt.T
        rot.append(R)
        trans.append(-R @ c)
    return (np.asarray(rot, dtype=np.float32).reshape(-1, 9),
            np.asarray(trans, dtype=np.float32))


def calibration(flen, w, h):
    """mve::CameraInfo::fill_calibration for paspect = 1, ppoint = .5."""
    ax = flen * max(w, h)
    return np.array([[ax, 0, 0.5 * w], [0, ax, 0.5 * h], [0, 0, 1.0]])


class DepthField:
    """w(u,v) = 5 + 0.6 sin(2 pi u / W * 1.5) cos(2 pi v / H) + 0.002 u * (1920 / W)
    on continuous pixel coordinates (pixel centres at integer + 0.5)."""

    def __init__(self, w, h):
        self.W, self.H = float(w), float(h)
        self.ax = 2 * math.pi * 1.5 / self.W
        self.ay = 2 * math.pi / self.H
        self.slope = 0.002 * 1920.0 / self.W

    def __call__(self, u, v):
        return 5.0 + 0.6 * np.sin(self.ax * u) * np.cos(self.ay * v) + self.slope * u

    def du(self, u, v):
        return 0.6 * self.ax * np.cos(self.ax * u) * np.cos(self.ay * v) + self.slope

    def dv(self, u, v):
        return -0.6 * self.ay * np.sin(self.ax * u) * np.sin(self.ay * v)

    def duv(self, u, v):
        return -0.6 * self.ax * self.ay * np.cos(self.ax * u) * np.sin(self.ay * v)


class Perturbation:
    """Smooth multiplicative noise 1 + amp * n(u,v), |n| <= 1."""

    def __init__(self, w, h, rng, amp=0.02):
        self.amp = amp
        self.k = rng.uniform(0.5, 3.0, size=(4, 2)) * 2 * math.pi / np.array([w, h])
        self.ph = rng.uniform(0, 2 * math.pi, size=4)

    def __call__(self, u, v):
        n = sum(np.sin(self.k[i, 0] * u + self.k[i, 1] * v + self.ph[i])
                for i in range(4)) / 4.0
        return 1.0 + self.amp * n


class Texture:
    """Band-limited pseudo-noise: plane waves with wavelengths around
    4/8/16/32 px (scaled with the image width), mapped into [0.2, 0.8]."""

    def __init__(self, w, rng, waves_per_octave=6):
        s = max(w / 1920.0, 0.5)
        ks, ps, am = [], [], []
        for lam in (6.0, 10.0, 18.0, 34.0):
            for _ in range(waves_per_octave):
                ang = rng.uniform(0, math.pi)
                l = lam * s * rng.uniform(0.85, 1.15)
                ks.append((2 * math.pi / l * math.cos(ang), 2 * math.pi / l * math.sin(ang)))
                ps.append(rng.uniform(0, 2 * math.pi))
                am.append(math.sqrt(lam))
        self.k = np.array(ks)
        self.ph = np.array(ps)
        self.amp = np.array(am)
        self.norm = 0.3 / (0.55 * np.sum(self.amp))

    def __call__(self, u, v):
        acc = np.zeros_like(u, dtype=np.float64)
        for (kx, ky), p, a in zip(self.k, self.ph, self.amp):
            acc += a * np.sin(kx * u + ky * v + p)
        return np.clip(0.5 + 1.8 * self.norm * acc, 0.2, 0.8)


def sh_basis(n):
    """Rescaled real SH basis, lib/spherical_harmonics.h:62-151 (inputs only:
    used to paint the synthetic images, not by any test as a checker)."""
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    x2, y2, z2 = x * x, y * y, z * z
    return np.stack([
        np.ones_like(x), y, z, x, x * y, y * z, -x2 - y2 + 2 * z2, x * z, x2 - y2,
        (3 * x2 - y2) * y, x * y * z, (4 * z2 - x2 - y2) * y,
        (2 * z2 - 3 * x2 - 3 * y2) * z, (4 * z2 - x2 - y2) * x, (x2 - y2) * z,
        (x2 - 3 * y2) * x], axis=-1)


def _surface_normal(u, v, depth: DepthField, w, h, f_px):
    """Camera-space normal in the reference's convention,
    lib/surface_derivative.cc:17-28."""
    x = u - 0.5 * w
    y = v - 0.5 * h
    wz, wx, wy = depth(u, v), depth.du(u, v), depth.dv(u, v)
    n = np.stack([wx, -wy, (x * wx + y * wy + wz) / f_px], axis=-1)
    return n / np.linalg.norm(n, axis=-1, keepdims=True)


def _invert_warp(M, t, depth: DepthField, pu, pv, iters=8):
    """Find main-view (u,v) with warp(u,v,depth(u,v)) = (pu,pv)."""
    u = pu.copy()
    v = pv.copy()
    for _ in range(iters):
        w = depth(u, v)
        wu, wv = depth.du(u, v), depth.dv(u, v)
        p = M[0, 0] * u + M[0, 1] * v + M[0, 2]
        q = M[1, 0] * u + M[1, 1] * v + M[1, 2]
        r = M[2, 0] * u + M[2, 1] * v + M[2, 2]
        a, b, d = w * p + t[0], w * q + t[1], w * r + t[2]
        fu, fv = a / d - pu, b / d - pv
        dau, dav = wu * p + w * M[0, 0], wv * p + w * M[0, 1]
        dbu, dbv = wu * q + w * M[1, 0], wv * q + w * M[1, 1]
        ddu, ddv = wu * r + w * M[2, 0], wv * r + w * M[2, 1]
        j00 = (dau * d - a * ddu) / (d * d)
        j01 = (dav * d - a * ddv) / (d * d)
        j10 = (dbu * d - b * ddu) / (d * d)
        j11 = (dbv * d - b * ddv) / (d * d)
        det = j00 * j11 - j01 * j10
        du = (j11 * fu - j01 * fv) / det
        dv = (-j10 * fu + j00 * fv) / det
        u -= du
        v -= dv
    return u, v


def make_scene(width, height, n_sub, seed_index=0, shading=False,
               init_noise=0.02) -> Scene:
    seed = BASE_SEED + int(seed_index)
    rng = np.random.default_rng(seed)
    rot, trans = make_cameras(n_sub)
    flen = np.ones(1 + n_sub, dtype=np.float32)
    K = calibration(1.0, width, height)
    Kinv = np.linalg.inv(K)
    depth = DepthField(width, height)
    tex = Texture(width, rng)
    pert = Perturbation(width, height, rng, amp=init_noise)
    f_px = float(max(width, height))

    def radiance(u, v):
        val = tex(u, v)
        if shading:
            n = _surface_normal(u, v, depth, width, height, f_px)
            s = sh_basis(n) @ SH_LIGHT
            val = val * s / 1.9
        return val

    ys, xs = np.mgrid[0:height, 0:width]
    pu = xs.astype(np.float64) + 0.5
    pv = ys.astype(np.float64) + 0.5
    def render(k):
        if k == 0:
            return np.clip(np.rint(255.0 * radiance(pu, pv)), 0, 255).astype(np.uint8)
        R = rot[k].astype(np.float64).reshape(3, 3)
        M = K @ R @ Kinv
        t = K @ trans[k].astype(np.float64)
        u, v = _invert_warp(M, t, depth, pu, pv)
        return np.clip(np.rint(255.0 * radiance(u, v)), 0, 255).astype(np.uint8)

    # numpy releases the GIL inside its loops: render the views concurrently
    # (the result does not depend on the scheduling)
    import concurrent.futures
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, 1 + n_sub)) as ex:
        images = list(ex.map(render, range(1 + n_sub)))

    true_depth = depth(pu, pv).astype(np.float32)
    init_depth = (depth(pu, pv) * pert(pu, pv)).astype(np.float32)
    sc = Scene(width, height, n_sub, flen, rot, trans, images, true_depth,
               init_depth, seed, shading)
    sc._depth_fn = depth
    sc._pert_fn = pert
    return sc


def reprojection(scene: Scene, k: int):
    """(Mi, ti) main -> neighbour k (0-based) computed in fp32 and widened,
    as DepthOptimizer::prepare_correspondences does
    (lib/depth_optimizer.cc:679-699; mve::CameraInfo::fill_reprojection)."""
    f32 = np.float32
    w, h = f32(scene.width), f32(scene.height)

    def calib(flen):
        ax = f32(flen) * f32(max(scene.width, scene.height))
        return np.array([[ax, 0, w * f32(0.5)], [0, ax, h * f32(0.5)], [0, 0, 1]], dtype=f32)

    def inv_calib(flen):
        ax = f32(flen) * f32(max(scene.width, scene.height))
        return np.array([[f32(1) / ax, 0, -w * f32(0.5) / ax],
                         [0, f32(1) / ax, -h * f32(0.5) / ax], [0, 0, 1]], dtype=f32)

    def mm(a, b):   # fp32, left-to-right accumulation from zero, no FMA
        out = np.zeros((a.shape[0], b.shape[1]), dtype=f32)
        for i in range(a.shape[0]):
            for j in range(b.shape[1]):
                s = f32(0)
                for q in range(a.shape[1]):
                    s = f32(s + f32(a[i, q] * b[q, j]))
                out[i, j] = s
        return out

    dst_K = calib(scene.flen[k + 1])
    dst_R = scene.rot[k + 1].reshape(3, 3).astype(f32)
    src_Ri = scene.rot[0].reshape(3, 3).astype(f32).T.copy()
    src_Ki = inv_calib(scene.flen[0])
    dst_t = scene.trans[k + 1].astype(f32).reshape(3, 1)
    src_t = scene.trans[0].astype(f32).reshape(3, 1)
    M = mm(mm(mm(dst_K, dst_R), src_Ri), src_Ki)
    tt = mm(dst_K, (dst_t - mm(mm(dst_R, src_Ri), src_t)).astype(f32))
    return M.astype(np.float64).reshape(9), tt.astype(np.float64).reshape(3)


def surface_grid(width, height, scale):
    """Grid geometry of Surface::Surface, lib/surface.cc:28-37."""
    ps = 1 << scale
    npx = (width - 2) // ps - 1
    npy = (height - 2) // ps - 1
    sx = (width - npx * ps) // 2
    sy = (height - npy * ps) // 2
    return ps, npx, npy, sx, sy


def analytic_nodes(scene: Scene, scale: int, perturbed=True):
    """Node parameters (f, dx, dy, dxy in patch units) sampled from the
    analytic (perturbed) depth field on the reference's grid. All nodes and
    patches valid."""
    ps, npx, npy, sx, sy = surface_grid(scene.width, scene.height, scale)
    iy, ix = np.mgrid[0:npy + 1, 0:npx + 1]
    u = (sx + ix * ps).astype(np.float64)
    v = (sy + iy * ps).astype(np.float64)
    d = scene._depth_fn
    if perturbed:
        p = scene._pert_fn
        eps = 1e-3

        def fn(a, b):
            return d(a, b) * p(a, b)
        f = fn(u, v)
        fx = (fn(u + eps, v) - fn(u - eps, v)) / (2 * eps)
        fy = (fn(u, v + eps) - fn(u, v - eps)) / (2 * eps)
        fxy = (fn(u + eps, v + eps) - fn(u + eps, v - eps)
               - fn(u - eps, v + eps) + fn(u - eps, v - eps)) / (4 * eps * eps)
    else:
        f, fx, fy, fxy = d(u, v), d.du(u, v), d.dv(u, v), d.duv(u, v)
    nodes = np.stack([f, fx * ps, fy * ps, fxy * ps * ps], axis=-1)
    return np.ascontiguousarray(nodes.reshape(-1, 4), dtype=np.float64)
