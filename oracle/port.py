"""ctypes binding of oracle/liboracle.so -- the plain C++ restatement of the
hot path (oracle/oracle_port.cc). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/liboracle.so not built (make -C oracle port)")
        _lib = C.CDLL(LIB_PATH)
        _lib.port_create.restype = C.c_void_p
        _lib.port_gn_construct.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PortScene:
    """Same call surface as smvs_b200.api.Context, on the CPU restatement."""

    def __init__(self, main_grad, sub_grads, sub_hess, Mi, ti, flen_px, inv_flen,
                 shading=None, shading_grad=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)  # noqa: E731
        self._keep = [f32(main_grad), [f32(a) for a in sub_grads],
                      [f32(a) for a in sub_hess], f32(shading), f32(shading_grad)]
        mg, sg, sh, s, sgr = self._keep
        n = len(sg)
        h, w = mg.shape[:2]
        sw = (C.c_int * max(n, 1))(*[a.shape[1] for a in sg])
        shh = (C.c_int * max(n, 1))(*[a.shape[0] for a in sg])
        gp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in sg])
        hp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in sh])
        Mi = np.ascontiguousarray(Mi, np.float64)
        ti = np.ascontiguousarray(ti, np.float64)
        self.h_ = C.c_void_p(lib().port_create(w, h, C.c_double(flen_px),
                                               C.c_double(inv_flen), _p(mg), _p(s),
                                               _p(sgr), n, sw, shh, gp, hp, _p(Mi), _p(ti)))
        self.w, self.h = w, h
        self.n_nodes = 0

    def close(self):
        if self.h_:
            lib().port_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_surface(self, scale, npx, npy, sx, sy, nodes, node_valid, patch_valid,
                    vis_off, vis_ids):
        nodes = np.ascontiguousarray(nodes, np.float64)
        nv = np.ascontiguousarray(node_valid, np.uint8)
        pv = np.ascontiguousarray(patch_valid, np.uint8)
        if vis_off is None:          # no lists yet: visibility() makes them
            vis_off, vis_ids = np.zeros(npx * npy + 1, np.uint32), np.zeros(1, np.uint8)
        vo = np.ascontiguousarray(vis_off, np.uint32)
        vi = np.ascontiguousarray(vis_ids, np.uint8)
        if vi.size == 0:
            vi = np.zeros(1, np.uint8)
        self.n_patches = npx * npy
        self._wh = None
        lib().port_set_surface(self.h_, int(scale), int(npx), int(npy), int(sx), int(sy),
                               _p(nodes), _p(nv), _p(pv), _p(vo), _p(vi))
        self.n_nodes = (npx + 1) * (npy + 1)

    def gn_construct(self, active, light16=None, regularization=0.01,
                     light_surf_regularization=0.0):
        a = np.ascontiguousarray(active, np.uint8)
        l = None if light16 is None else np.ascontiguousarray(light16, np.float64)
        return int(lib().port_gn_construct(self.h_, _p(a), _p(l), C.c_double(regularization),
                                           C.c_double(light_surf_regularization)))

    def get_system(self):
        sz = (C.c_uint64 * 3)()
        lib().port_get_system_sizes(self.h_, sz)
        n, nh, npc = int(sz[0]), int(sz[1]), int(sz[2])
        g = np.empty(n); Hv = np.empty((nh, 16)); Pv = np.empty((npc, 16))
        Ho = np.empty(n // 4 + 1, np.uint64); Hi = np.empty(nh, np.uint64)
        Po = np.empty(n // 4 + 1, np.uint64); Pi = np.empty(npc, np.uint64)
        lib().port_get_system(self.h_, _p(g), _p(Hv), _p(Ho), _p(Hi), _p(Pv), _p(Po), _p(Pi))
        return dict(g=g, Hvals=Hv, Houter=Ho, Hinner=Hi, Pvals=Pv, Pouter=Po, Pinner=Pi)

    def cg_solve(self, max_iter=200, err_tol=-1.0, q_tol=1e-3):
        x = np.empty(self.n_nodes * 4)
        it, info = C.c_int(0), C.c_int(0)
        lib().port_cg_solve(self.h_, int(max_iter), C.c_double(err_tol), C.c_double(q_tol),
                            _p(x), C.byref(it), C.byref(info))
        return x, it.value, info.value

    def update_nodes(self, delta, active, reproj_thresh=0.15, full_opt=False):
        d = np.ascontiguousarray(delta, np.float64)
        act = np.array(active, dtype=np.uint8, copy=True)
        n_active, shift = C.c_uint64(0), C.c_double(0)
        lib().port_update_nodes(self.h_, _p(d), C.c_double(reproj_thresh), int(full_opt),
                                _p(act), C.byref(n_active), C.byref(shift))
        return act, int(n_active.value), float(shift.value)

    # -- surface topology -----------------------------------------------------
    def _after_topology(self):
        info = (C.c_int * 6)()
        lib().port_surface_info(self.h_, info)
        self.info = dict(scale=info[0], npx=info[1], npy=info[2], start_x=info[3],
                         start_y=info[4], patchsize=info[5])
        self.n_nodes = (info[1] + 1) * (info[2] + 1)
        self.n_patches = info[1] * info[2]

    def surface_create(self, scale, init_depth):
        d = np.ascontiguousarray(init_depth, np.float32)
        lib().port_surface_create(self.h_, int(scale), _p(d))
        self._after_topology()

    def surface_subdivide(self):
        lib().port_surface_subdivide(self.h_)
        self._after_topology()

    def surface_fill_from_depth(self):
        lib().port_surface_fill_from_depth(self.h_)

    def surface_remove_isolated(self):
        lib().port_surface_remove_isolated(self.h_)

    def surface_expand(self):
        return int(lib().port_surface_expand(self.h_))

    # -- visibility / cutting (the callers' side of the loop) ----------------
    def set_images(self, main_image, sub_images):
        """Unblurred float images (h, w, 3) of the main view and the neighbours:
        what the use_sgm = false mode's NCC filter reads."""
        m = np.ascontiguousarray(main_image, np.float32)
        subs = [np.ascontiguousarray(a, np.float32) for a in sub_images]
        self._keep_images = [m, subs]
        ptrs = (C.c_void_p * max(len(subs), 1))(*[a.ctypes.data for a in subs])
        lib().port_set_images(self.h_, _p(m), ptrs)

    def visibility(self, sgm_depth):
        """create_subview_surfaces; sgm_depth None = the use_sgm = false mode."""
        d = None if sgm_depth is None else np.ascontiguousarray(sgm_depth, np.float32)
        return int(lib().port_visibility(self.h_, _p(d)))

    def cut_boundaries(self, inv_calib9):
        k = np.ascontiguousarray(inv_calib9, np.float32).reshape(9)
        return int(lib().port_cut_boundaries(self.h_, _p(k)))

    def surface_state(self):
        nv = np.empty(self.n_nodes, np.uint8)
        pv = np.empty(self.n_patches, np.uint8)
        vo = np.empty(self.n_patches + 1, np.uint32)
        vi = np.empty(self.n_patches * 32 + 1, np.uint8)
        lib().port_get_surface_state(self.h_, _p(nv), _p(pv), _p(vo), _p(vi))
        return nv, pv, vo, vi[:vo[-1]].copy()

    def get_depth(self, h, w):
        out = np.empty((h, w), np.float32)
        lib().port_depth_map(self.h_, _p(out))
        return out

    def get_nodes(self):
        out = np.empty((self.n_nodes, 4))
        lib().port_get_nodes(self.h_, _p(out))
        return out

    def newton_loop(self, light16=None, regularization=0.01,
                    light_surf_regularization=0.0, max_steps=200):
        l = None if light16 is None else np.ascontiguousarray(light16, np.float64)
        st = np.zeros(8)
        lib().port_newton_loop(self.h_, _p(l), C.c_double(regularization),
                               C.c_double(light_surf_regularization), int(max_steps), _p(st))
        return dict(newton_steps=int(st[0]), cg_iterations=int(st[1]),
                    pixel_iterations=float(st[2]), n_active=int(st[6]), nan=bool(st[7]))


def sgm(main_lum, neigh_lum, M, t, min_depth, max_depth, num_steps=128, penalty1=6,
        penalty2=96):
    m = np.ascontiguousarray(main_lum, np.uint8)
    n = np.ascontiguousarray(neigh_lum, np.uint8)
    h, w = m.shape
    nh, nw = n.shape
    M = np.ascontiguousarray(M, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    depth = np.empty((h, w), np.float32)
    cost = np.empty((h, w, num_steps), np.uint16)
    S = np.empty((h, w, num_steps), np.uint16)
    lib().port_sgm(w, h, _p(m), nw, nh, _p(n), _p(M), _p(t), C.c_float(min_depth),
                   C.c_float(max_depth), int(num_steps), int(penalty1), int(penalty2),
                   _p(depth), _p(cost), _p(S))
    return dict(depth=depth, cost=cost, sgm=S)


class Units:
    name = "restatement (oracle/oracle_port.cc)"

    @staticmethod
    def bicubic_eval(nodes16, x, y):
        n = np.ascontiguousarray(nodes16, np.float64).reshape(16)
        out = np.empty(6)
        lib().port_bicubic_eval(_p(n), C.c_double(x), C.c_double(y), _p(out))
        return out

    @staticmethod
    def node_derivatives(x, y, patchsize=0.0):
        out = np.empty(96)
        lib().port_bicubic_node_derivatives(C.c_double(x), C.c_double(y),
                                            C.c_double(patchsize), _p(out))
        return out

    @staticmethod
    def correspondence(M, t, u, v, w, wx=0.0, wy=0.0, grad=(0.0, 0.0), dn=None):
        M = np.ascontiguousarray(M, np.float64).reshape(9)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        g = np.ascontiguousarray(grad, np.float64).reshape(2)
        dn = np.zeros(96) if dn is None else np.ascontiguousarray(dn, np.float64)
        proj = np.empty(2); jac = np.empty(4); c_dn = np.empty((16, 2)); j_dn = np.empty((16, 2))
        depth = C.c_double(0)
        lib().port_correspondence(_p(M), _p(t), C.c_double(u), C.c_double(v), C.c_double(w),
                                  C.c_double(wx), C.c_double(wy), _p(g), _p(dn), _p(proj),
                                  _p(jac), _p(c_dn), _p(j_dn), C.byref(depth))
        return dict(proj=proj, jac=jac, c_dn=c_dn, jac_dn=j_dn, depth=depth.value)

    @staticmethod
    def surface_derivatives(dn, x, y, f, w, dx, dy, dxy, dxx, dyy):
        dn = np.ascontiguousarray(dn, np.float64)
        normal = np.empty(3); div = np.empty(6); dd = np.empty(96); nd = np.empty(48)
        lib().port_surface_derivatives(_p(dn), *[C.c_double(a) for a in
                                                 (x, y, f, w, dx, dy, dxy, dxx, dyy)],
                                       _p(normal), _p(div), _p(dd), _p(nd))
        return dict(normal=normal, div=div, div_deriv=dd, normal_deriv=nd)

    @staticmethod
    def sh_4band(normal):
        n = np.ascontiguousarray(normal, np.float64)
        sh = np.empty(16); d = np.empty(48)
        lib().port_sh_4band(_p(n), _p(sh), _p(d))
        return sh, d

    @staticmethod
    def ldl_inverse(A):
        A = np.array(A, dtype=np.float64, copy=True)
        lib().port_ldl_inverse(_p(A), A.shape[0])
        return A


def bilateral_filter(guide, depth, sigma=5.0, kernel_size=5):
    """DepthOptimizer::depthmap_bilateral_filter, restated."""
    g = np.ascontiguousarray(guide, np.float32)
    d = np.ascontiguousarray(depth, np.float32)
    h, w = g.shape[:2]
    ch = 1 if g.ndim == 2 else g.shape[2]
    out = np.empty((h, w), np.float32)
    lib().port_bilateral_filter(w, h, ch, _p(g), d.shape[1], d.shape[0], _p(d),
                                C.c_float(sigma), int(kernel_size), _p(out))
    return out
