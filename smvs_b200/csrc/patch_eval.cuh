/*
 * patch_eval.cuh -- BicubicPatch::compute_coefficients / evaluate_f / _dx /
 * _dy (lib/bicubic_patch.cc:56-86, 121-160) and Correspondence::update /
 * fill / fill_jacobian (lib/correspondence.cc:20-51, 88-100) restated so that
 * every value is BITWISE the CPU's: the reference's expression order, one
 * IEEE operation at a time, no FMA contraction (xd, gn_math.cuh). The
 * visibility and boundary-cutting kernels take yes/no decisions on these
 * values (depth tests, border tests, thresholds), so "close" is not enough.
 */
#ifndef SMVSB_PATCH_EVAL_CUH
#define SMVSB_PATCH_EVAL_CUH

#include "gn_math.cuh"

namespace smvsb {

/* Hermite interpolation matrix: 16 polynomial coefficients from the 16 node
 * values ordered (f x4, dx x4, dy x4, dxy x4), lib/bicubic_patch.cc:20-38.
 * (A table of small integers that any bicubic Hermite patch implies.) */
static __constant__ double c_hermite[256] = {
    1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    -3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0,
    -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0,
    0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0,
    9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1,
    -6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1,
    2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0,
    -6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1,
    4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1
};

/* One row of BicubicPatch::compute_coefficients: a = A x with
 * x[4 * c + node] = theta[node * 4 + c], summed in index order. Row r of the
 * product lands at coeffs[i][j] = coef[i * 4 + j] with i = r & 3, j = r >> 2;
 * coefficient_slot(r) is that index. theta: node-major (n00, n10, n01, n11)
 * x (f, dx, dy, dxy). */
__device__ __forceinline__ double
coefficient_row (double const* theta, int r)
{
    xd sum(0.0);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sum += xd(c_hermite[r * 16 + k]) * xd(theta[(k & 3) * 4 + (k >> 2)]);
    return sum.v;
}

__device__ __forceinline__ int
coefficient_slot (int r)
{
    return (r & 3) * 4 + (r >> 2);
}

/* The four node values of a patch, node-major, from the surface's node array
 * (lib/surface.cc:283-307: n00, n10, n01, n11). */
__device__ __forceinline__ void
load_patch_theta (double const* __restrict__ nodes, int npx, int idx, int idy,
    double* theta)
{
#pragma unroll
    for (int nd = 0; nd < 4; ++nd)
    {
        int const node = (idy + (nd >> 1)) * (npx + 1) + idx + (nd & 1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            theta[nd * 4 + c] = nodes[static_cast<size_t>(node) * 4 + c];
    }
}

__device__ __forceinline__ void
patch_coefficients (double const* theta, double* coef)
{
    for (int r = 0; r < 16; ++r)
        coef[coefficient_slot(r)] = coefficient_row(theta, r);
}

/* evaluate_f / evaluate_dx / evaluate_dy at pixel (i, j) of a patch of size
 * ps, with the 1/ps scaling of lib/surface_patch.cc:85-99. */
struct PatchSample
{
    double w, wx, wy;
};

template <bool DERIV>
__device__ __forceinline__ PatchSample
patch_sample (double const* cf, int i, int j, int ps)
{
    xd const size(static_cast<double>(ps));
    xd const sx = (xd(static_cast<double>(i)) + xd(0.5)) / size;
    xd const sy = (xd(static_cast<double>(j)) + xd(0.5)) / size;
    xd ex[4], ey[4];
    ex[0] = xd(1.0); ex[1] = sx; ex[2] = sx * sx; ex[3] = ex[2] * sx;
    ey[0] = xd(1.0); ey[1] = sy; ey[2] = sy * sy; ey[3] = ey[2] * sy;
    xd f(0.0);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            f += xd(cf[a * 4 + b]) * ex[a] * ey[b];
    PatchSample out;
    out.w = f.v;
    out.wx = 0.0;
    out.wy = 0.0;
    if (DERIV)
    {
        xd fx(0.0), fy(0.0);
#pragma unroll
        for (int a = 1; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                fx += xd(cf[a * 4 + b]) * xd(double(a)) * ex[a - 1] * ey[b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 1; b < 4; ++b)
                fy += xd(cf[a * 4 + b]) * ex[a] * xd(double(b)) * ey[b - 1];
        out.wx = (fx / size).v;
        out.wy = (fy / size).v;
    }
    return out;
}

/* Correspondence of pixel-centre (u, v) at depth w into one neighbour. */
struct Warp
{
    double projx, projy;    /* fill() - 0.5 */
    double depth;           /* get_depth(): w * r + t2 */
    double jac[4];          /* fill_jacobian */
};

template <bool JAC>
__device__ __forceinline__ Warp
warp_pixel (double const* __restrict__ Mt, double u, double v, double w,
    double wx, double wy)
{
    xd const m0(Mt[0]), m1(Mt[1]), m2(Mt[2]), m3(Mt[3]), m4(Mt[4]),
        m5(Mt[5]), m6(Mt[6]), m7(Mt[7]), m8(Mt[8]), T0(Mt[9]), T1(Mt[10]),
        T2(Mt[11]), W(w), U(u), V(v);
    xd const p = m0 * U + m1 * V + m2;
    xd const q = m3 * U + m4 * V + m5;
    xd const r = m6 * U + m7 * V + m8;
    xd const a = W * p + T0;
    xd const b = W * q + T1;
    xd const d = W * r + T2;
    Warp out;
    out.projx = (a / d - xd(0.5)).v;
    out.projy = (b / d - xd(0.5)).v;
    out.depth = d.v;
    if (JAC)
    {
        xd const WX(wx), WY(wy);
        xd const d2 = d * d;
        xd j0 = (WX * p + W * m0) / d;
        xd j2 = (WY * p + W * m1) / d;
        j0 -= a * (WX * r + W * m6) / d2;
        j2 -= a * (WY * r + W * m7) / d2;
        xd j1 = (WX * q + W * m3) / d;
        xd j3 = (WY * q + W * m4) / d;
        j1 -= b * (WX * r + W * m6) / d2;
        j3 -= b * (WY * r + W * m7) / d2;
        out.jac[0] = j0.v; out.jac[1] = j1.v; out.jac[2] = j2.v;
        out.jac[3] = j3.v;
    }
    return out;
}

} /* namespace smvsb */

#endif
