/*
 * oracle/oracle_port.cc -- plain C++ restatement of the SMVS hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). No MVE, no reference
 * headers, no SIMD: scalar loops that follow the reference function by
 * function (every function cites the file:line it restates). It is pinned
 * (tests/test_known_answers.py, tests/test_oracle_port.py) against
 *   - the reference's own unit-test assertions, and
 *   - the outputs of oracle/_ref (the reference's sources compiled verbatim),
 * and exists so that a checker can be built where /root/reference is absent.
 * Parity caveat shared with oracle/_ref: mve::Image::linear_at semantics are
 * restated from memory of MVE (clamped coordinates, fp32 weights, +0.5
 * rounding for bytes) -- "unpinned" at that seam.
 *
 * Deliberately written in the reference's formulation (16-column Jacobian
 * rows, rank-1 updates, std::map block assembly, O(D^2) SGM minimum), NOT in
 * the restructured form the CUDA kernels use, so it is an independent check.
 * The same holds for the callers' side of the loop restated here: sequential
 * z-buffer and per-patch tests of create_subview_surfaces, cut_boundaries with
 * mse_for_patch, the depth-map render and the joint bilateral filter (libm's
 * expf) -- all pinned for EQUALITY against oracle/_ref and tests/golden.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

double const R_FACTOR = 1e-4;   /* lib/gauss_newton_step.cc:17 */

/* lib/bicubic_patch.cc:20-38 */
double const coefficient_matrix[256] = {
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

/* ------------------------------------------------------------------ */
/* BicubicPatch                                                       */
/* ------------------------------------------------------------------ */

/* nodes16: 4 nodes (n00, n10, n01, n11) x (f, dx, dy, dxy).
 * compute_coefficients, lib/bicubic_patch.cc:56-86. */
void
patch_coefficients (double const* nodes16, double coeffs[4][4])
{
    double x[16];
    for (int c = 0; c < 4; ++c)
        for (int n = 0; n < 4; ++n)
            x[4 * c + n] = nodes16[n * 4 + c];
    double a[16];
    for (int r = 0; r < 16; ++r)
    {
        double s = 0.0;
        for (int k = 0; k < 16; ++k)
            s += coefficient_matrix[r * 16 + k] * x[k];
        a[r] = s;
    }
    for (int k = 0, j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i, ++k)
            coeffs[i][j] = a[k];
}

/* evaluate_f/dx/dy/dxy/dxx/dyy, lib/bicubic_patch.cc:121-187 */
void
patch_evaluate (double const coeffs[4][4], double px, double py, double* out6)
{
    double ex[4] = { 1.0, px, px * px, px * px * px };
    double ey[4] = { 1.0, py, py * py, py * py * py };
    double f = 0, dx = 0, dy = 0, dxy = 0, dxx = 0, dyy = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            f += coeffs[i][j] * ex[i] * ey[j];
    for (int i = 1; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            dx += coeffs[i][j] * i * ex[i - 1] * ey[j];
    for (int i = 2; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            dxx += coeffs[i][j] * i * (i - 1) * ex[i - 2] * ey[j];
    for (int i = 0; i < 4; ++i)
        for (int j = 1; j < 4; ++j)
            dy += coeffs[i][j] * ex[i] * j * ey[j - 1];
    for (int i = 0; i < 4; ++i)
        for (int j = 2; j < 4; ++j)
            dyy += coeffs[i][j] * ex[i] * j * (j - 1) * ey[j - 2];
    for (int i = 1; i < 4; ++i)
        for (int j = 1; j < 4; ++j)
            dxy += coeffs[i][j] * i * ex[i - 1] * j * ey[j - 1];
    out6[0] = f; out6[1] = dx; out6[2] = dy; out6[3] = dxy;
    out6[4] = dxx; out6[5] = dyy;
}

/* node_deriv + node_derivatives, lib/bicubic_patch.cc:258-316; out[96] =
 * 4 nodes x [d_f(4) d_dx(4) d_dy(4) d_dxy(4) d_dxx(4) d_dyy(4)]. patchsize
 * > 0 applies the 1/ps, 1/ps^2 scaling of :318-340 / lib/surface.cc:929-955. */
void
node_derivatives (double px, double py, double patchsize, double* out)
{
    double x[4] = { 1.0, px, px * px, px * px * px };
    double y[4] = { 1.0, py, py * py, py * py * py };
    std::fill(out, out + 96, 0.0);
    for (int node = 0; node < 4; ++node)
    {
        double* d = out + 24 * node;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 4; ++c)
                {
                    double const m = coefficient_matrix[16 * (j * 4 + i)
                        + 4 * c + node];
                    d[0 + c] += m * x[i] * y[j];
                    if (i > 0) d[4 + c] += m * i * x[i - 1] * y[j];
                    if (j > 0) d[8 + c] += m * x[i] * j * y[j - 1];
                    if (i > 0 && j > 0)
                        d[12 + c] += m * i * x[i - 1] * j * y[j - 1];
                    if (i > 1) d[16 + c] += m * i * (i - 1) * x[i - 2] * y[j];
                    if (j > 1) d[20 + c] += m * x[i] * j * (j - 1) * y[j - 2];
                }
    }
    if (patchsize > 0.0)
    {
        double const p2p = 1.0 / patchsize;
        for (int node = 0; node < 4; ++node)
        {
            for (int i = 4; i < 24; ++i) out[24 * node + i] *= p2p;
            for (int i = 12; i < 24; ++i) out[24 * node + i] *= p2p;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Correspondence, lib/correspondence.cc                              */
/* ------------------------------------------------------------------ */

struct Corr
{
    double p, q, r, t[3], w, wp[2], a, b, d, d2;
    double pp[2], qp[2], rp[2];

    /* update, :20-44 */
    void update (double const* M, double const* tt, double u, double v,
        double w_, double wdx, double wdy)
    {
        t[0] = tt[0]; t[1] = tt[1]; t[2] = tt[2];
        w = w_;
        pp[0] = M[0]; pp[1] = M[1];
        qp[0] = M[3]; qp[1] = M[4];
        rp[0] = M[6]; rp[1] = M[7];
        wp[0] = wdx; wp[1] = wdy;
        p = M[0] * u + M[1] * v + M[2];
        q = M[3] * u + M[4] * v + M[5];
        r = M[6] * u + M[7] * v + M[8];
        a = w * p + t[0];
        b = w * q + t[1];
        d = w * r + t[2];
        d2 = d * d;
    }
    /* fill, :46-51 */
    void fill (double* c) const { c[0] = a / d; c[1] = b / d; }
    /* fill_jacobian, :88-100 */
    void fill_jacobian (double* jac) const
    {
        jac[0] = (wp[0] * p + w * pp[0]) / d;
        jac[2] = (wp[1] * p + w * pp[1]) / d;
        jac[0] -= a * (wp[0] * r + w * rp[0]) / d2;
        jac[2] -= a * (wp[1] * r + w * rp[1]) / d2;
        jac[1] = (wp[0] * q + w * qp[0]) / d;
        jac[3] = (wp[1] * q + w * qp[1]) / d;
        jac[1] -= b * (wp[0] * r + w * rp[0]) / d2;
        jac[3] -= b * (wp[1] * r + w * rp[1]) / d2;
    }
    /* fill_derivative, :74-86 (c_dn[16][2]) */
    void fill_derivative (double const* dn, double* c_dn) const
    {
        double const du_w = (p * d - r * a) / d2;
        double const dv_w = (q * d - r * b) / d2;
        for (int n = 0; n < 4; ++n)
            for (int i = 0; i < 4; ++i)
            {
                c_dn[(n * 4 + i) * 2 + 0] = du_w * dn[n * 24 + i];
                c_dn[(n * 4 + i) * 2 + 1] = dv_w * dn[n * 24 + i];
            }
    }
    /* fill_jacobian_derivative_grad, :102-187 (jac_dn[16][2]) */
    void fill_jacobian_derivative_grad (double const* grad, double const* dn,
        double* jac_dn) const
    {
        double const d4 = d2 * d2;
        double const d_prime = 2.0 * d * r;
        double du_a_temp[2], du_a_prime[2], du_b_prime[2], du_c[2];
        double dv_a_temp[2], dv_a_prime[2], dv_b_prime[2], dv_c[2];
        double const du_c_prime = p * t[2] - r * t[0];
        double const dv_c_prime = q * t[2] - r * t[1];
        double du_x[2], dv_x[2];
        for (int k = 0; k < 2; ++k)
        {
            du_a_temp[k] = w * (pp[k] * r - p * rp[k]);
            du_a_prime[k] = 2.0 * du_a_temp[k];
            du_b_prime[k] = pp[k] * t[2] - rp[k] * t[0];
            du_c[k] = wp[k] * du_c_prime;
            dv_a_temp[k] = w * (qp[k] * r - q * rp[k]);
            dv_a_prime[k] = 2.0 * dv_a_temp[k];
            dv_b_prime[k] = qp[k] * t[2] - rp[k] * t[1];
            dv_c[k] = wp[k] * dv_c_prime;
            double const du_a_b_c = w * (du_a_temp[k] + du_b_prime[k]) + du_c[k];
            double const dv_a_b_c = w * (dv_a_temp[k] + dv_b_prime[k]) + dv_c[k];
            du_x[k] = (du_a_prime[k] + du_b_prime[k]) / d2
                - du_a_b_c * d_prime / d4;
            dv_x[k] = (dv_a_prime[k] + dv_b_prime[k]) / d2
                - dv_a_b_c * d_prime / d4;
        }
        double const du_c_prime_d = du_c_prime / d2;
        double const dv_c_prime_d = dv_c_prime / d2;
        for (int n = 0; n < 4; ++n)
            for (int i = 0; i < 4; ++i)
            {
                int const o = n * 24;
                double const du0 = du_x[0] * dn[o + i], du1 = du_x[1] * dn[o + i];
                double const dv0 = dv_x[0] * dn[o + i], dv1 = dv_x[1] * dn[o + i];
                jac_dn[(n * 4 + i) * 2 + 0] =
                    (du0 + du_c_prime_d * dn[o + 4 + i]) * grad[0]
                    + (dv0 + dv_c_prime_d * dn[o + 4 + i]) * grad[1];
                jac_dn[(n * 4 + i) * 2 + 1] =
                    (du1 + du_c_prime_d * dn[o + 8 + i]) * grad[0]
                    + (dv1 + dv_c_prime_d * dn[o + 8 + i]) * grad[1];
            }
    }
};

/* ------------------------------------------------------------------ */
/* surfderiv, lib/surface_derivative.cc                               */
/* ------------------------------------------------------------------ */

/* fill_normal, :17-28 */
void
fill_normal (double x, double y, double inv_flen, double w, double dx,
    double dy, double* n)
{
    double v[3] = { dx, -dy, (x * dx + y * dy + w) * inv_flen };
    double const len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    n[0] = v[0] / len; n[1] = v[1] / len; n[2] = v[2] / len;
}

/* normal_derivative, :31-65 (deriv[48]) */
void
normal_derivative (double const* d_node, double x, double y, double f,
    double w, double dx, double dy, double* deriv)
{
    double const f_sqr_inv = 1.0 / (f * f);
    double const a = w + x * dx + y * dy;
    double const t = dx * dx + dy * dy + a * a * f_sqr_inv;
    double const n = std::sqrt(t);
    for (int node = 0; node < 4; ++node)
        for (int i = 0; i < 4; ++i)
        {
            double const* dn = d_node + 24 * node;
            double const w_p = dn[i], dx_p = dn[4 + i], dy_p = dn[8 + i];
            double const a_p = w_p + x * dx_p + y * dy_p;
            double const t_p2 = dx * dx_p + dy * dy_p + f_sqr_inv * a * a_p;
            double const n_p = t_p2 / n;
            deriv[0 + node * 4 + i] = (dx_p * n - dx * n_p) / t;
            deriv[16 + node * 4 + i] = (-dy_p * n + dy * n_p) / t;
            deriv[32 + node * 4 + i] = (a_p * n - a * n_p) / (t * f);
        }
}

/* normal_divergence, :69-107 (div[6]) */
void
normal_divergence (double x, double y, double f, double w, double dx,
    double dy, double dxy, double dxx, double dyy, double* div)
{
    double const a = w + x * dx + y * dy;
    double const ax = 2.0 * dx + x * dxx + y * dxy;
    double const ay = 2.0 * dy + y * dyy + x * dxy;
    double t = a / f;
    t = t * t;
    t += dx * dx + dy * dy;
    double const n = std::sqrt(t);
    double nx = dx * dxx + dy * dxy;
    nx += (1.0 / (f * f)) * (w + x * dx + y * dy)
        * (dx + dx + x * dxx + y * dxy);
    nx /= n;
    double ny = dx * dxy + dy * dyy;
    ny += (1.0 / (f * f)) * (w + x * dx + y * dy)
        * (dy + dy + x * dxy + y * dyy);
    ny /= n;
    div[0] = (dxx * n - dx * nx) / t;
    div[1] = -((dxy * n - dy * nx) / t);
    div[2] = (ax * n - a * nx) / (t * f);
    div[3] = (dxy * n - dx * ny) / t;
    div[4] = -((dyy * n - dy * ny) / t);
    div[5] = (ay * n - a * ny) / (t * f);
}

/* normal_divergence_deriv, :109-190 (full_deriv[96]) */
void
normal_divergence_deriv (double const* d_node, double x, double y, double f,
    double w, double dx, double dy, double dxy, double dxx, double dyy,
    double* full)
{
    double const f_sqr_inv = 1.0 / (f * f);
    double const a = w + x * dx + y * dy;
    double const ax = 2.0 * dx + x * dxx + y * dxy;
    double const ay = 2.0 * dy + y * dyy + x * dxy;
    double const a_f2 = a * f_sqr_inv;
    double const t = dx * dx + dy * dy + a * a_f2;
    double const n = std::sqrt(t);
    double const b = dx * dxx + dy * dxy + a_f2 * ax;
    double const c = dx * dxy + dy * dyy + a_f2 * ay;
    double const nx = b / n, ny = c / n;
    for (int node = 0; node < 4; ++node)
        for (int i = 0; i < 4; ++i)
        {
            double const* dn = d_node + 24 * node;
            double const w_p = dn[i], dx_p = dn[4 + i], dy_p = dn[8 + i];
            double const dxy_p = dn[12 + i], dxx_p = dn[16 + i];
            double const dyy_p = dn[20 + i];
            double const a_p = w_p + x * dx_p + y * dy_p;
            double const ax_p = 2.0 * dx_p + x * dxx_p + y * dxy_p;
            double const ay_p = 2.0 * dy_p + y * dyy_p + x * dxy_p;
            double const t_p2 = dx * dx_p + dy * dy_p + f_sqr_inv * a * a_p;
            double const n_p = t_p2 / n;
            double const b_p = (dx_p * dxx + dx * dxx_p)
                + (dy_p * dxy + dy * dxy_p)
                + f_sqr_inv * (a_p * ax + a * ax_p);
            double const c_p = (dx_p * dxy + dx * dxy_p)
                + (dy_p * dyy + dy * dyy_p)
                + f_sqr_inv * (a_p * ay + a * ay_p);
            double const nx_p = (b_p * n - b * n_p) / t;
            double const ny_p = (c_p * n - c * n_p) / t;
            double const tt = t * t;
            double const xx_p = ((dxx_p * n + dxx * n_p - dx_p * nx
                - dx * nx_p) * t - (dxx * n - dx * nx) * t_p2 * 2.0) / tt;
            double const yy_p = ((dyy_p * n + dyy * n_p - dy_p * ny
                - dy * ny_p) * t - (dyy * n - dy * ny) * t_p2 * 2.0) / tt;
            double const xy_p = ((dxy_p * n + dxy * n_p - dx_p * ny
                - dx * ny_p) * t - (dxy * n - dx * ny) * t_p2 * 2.0) / tt;
            double const yx_p = ((dxy_p * n + dxy * n_p - dy_p * nx
                - dy * nx_p) * t - (dxy * n - dy * nx) * t_p2 * 2.0) / tt;
            double const zx_p = ((ax_p * n + ax * n_p - a_p * nx - a * nx_p)
                * t - (ax * n - a * nx) * t_p2 * 2.0) / (tt * f);
            double const zy_p = ((ay_p * n + ay * n_p - a_p * ny - a * ny_p)
                * t - (ay * n - a * ny) * t_p2 * 2.0) / (tt * f);
            int const o = node * 4 + i;
            full[0 + o] = xx_p;
            full[16 + o] = -yx_p;
            full[32 + o] = zx_p;
            full[48 + o] = xy_p;
            full[64 + o] = -yy_p;
            full[80 + o] = zy_p;
        }
}

/* ------------------------------------------------------------------ */
/* spherical harmonics, lib/spherical_harmonics.h:62-201              */
/* ------------------------------------------------------------------ */

void
sh_evaluate_4_band (double const* n, double* sh)
{
    double const x = n[0], y = n[1], z = n[2];
    double const x2 = x * x, y2 = y * y, z2 = z * z;
    sh[0] = 1.0; sh[1] = y; sh[2] = z; sh[3] = x;
    sh[4] = x * y; sh[5] = y * z;
    sh[6] = -x2 - y2 + 2.0 * z2;
    sh[7] = x * z; sh[8] = x * x - y * y;
    sh[9] = (3.0 * x2 - y2) * y;
    sh[10] = x * y * z;
    sh[11] = (4.0 * z2 - x2 - y2) * y;
    sh[12] = (2.0 * z2 - 3.0 * x2 - 3.0 * y2) * z;
    sh[13] = (4.0 * z2 - x2 - y2) * x;
    sh[14] = (x2 - y2) * z;
    sh[15] = (x2 - 3.0 * y2) * x;
}

void
sh_derivative_4_band (double const* n, double* d)
{
    double const x = n[0], y = n[1], z = n[2];
    double const x2 = x * x, y2 = y * y, z2 = z * z;
    double const v[48] = {
        0, 0, 0,   0, 1, 0,   0, 0, 1,   1, 0, 0,
        y, x, 0,   0, z, y,   -2 * x, -2 * y, 4 * z,
        z, 0, x,   2 * x, -2 * y, 0,
        6 * x * y, 3 * (x2 - y2), 0,
        y * z, x * z, x * y,
        -2 * x * y, 4 * z2 - x2 - 3 * y2, 8 * y * z,
        -6 * x * z, -6 * y * z, 6 * z2 - 3 * (x2 + y2),
        4 * z2 - 3 * x2 - y2, -2 * x * y, 8 * x * z,
        2 * x * z, -2 * y * z, x2 - y2,
        3 * (x2 - y2), -6 * x * y, 0 };
    std::copy(v, v + 48, d);
}

/* ------------------------------------------------------------------ */
/* ldl_inverse, lib/ldl_decomposition.h:43-92                         */
/* ------------------------------------------------------------------ */

void
ldl_inverse (double* A, int const size)
{
    std::vector<double> L(size * size, 0.0), D(size, 0.0);
    for (int j = 0; j < size; ++j)
    {
        D[j] = A[j * size + j];
        L[j * size + j] = 1.0;
        for (int k = 0; k < j; ++k)
            D[j] -= (L[j * size + k] * L[j * size + k]) * D[k];
        if (D[j] == 0.0)
            return;
        for (int i = j + 1; i < size; ++i)
        {
            L[i * size + j] = A[i * size + j];
            for (int k = 0; k < j; ++k)
                L[i * size + j] -= L[i * size + k] * D[k] * L[j * size + k];
            L[i * size + j] /= D[j];
        }
    }
    for (int i = 0; i < size; ++i)
        for (int j = i + 1; j < size; ++j)
        {
            double sum = 0.0;
            for (int k = i; k < j; ++k)
                sum -= L[j * size + k] * L[k * size + i];
            L[j * size + i] = sum;
        }
    for (int i = 0; i < size; ++i)
        D[i] = 1.0 / D[i];
    std::fill(A, A + size * size, 0.0);
    for (int r = 0; r < size; ++r)
        for (int c1 = 0; c1 < size; ++c1)
            for (int c2 = 0; c2 < size; ++c2)
                A[c1 * size + c2] += L[r * size + c2] * L[r * size + c1] * D[r];
}

/* ------------------------------------------------------------------ */
/* images                                                             */
/* ------------------------------------------------------------------ */

/* mve::Image<float>::linear_at (restated, see header) */
float
linear_at_f (float const* img, int w, int h, int c, float x, float y, int ch)
{
    x = std::max(0.0f, std::min(static_cast<float>(w - 1), x));
    y = std::max(0.0f, std::min(static_cast<float>(h - 1), y));
    int const fx = static_cast<int>(x), fy = static_cast<int>(y);
    int const fx1 = std::min(fx + 1, w - 1), fy1 = std::min(fy + 1, h - 1);
    float const w1 = x - static_cast<float>(fx), w0 = 1.0f - w1;
    float const w3 = y - static_cast<float>(fy), w2 = 1.0f - w3;
    float const v00 = img[(fy * w + fx) * c + ch];
    float const v10 = img[(fy * w + fx1) * c + ch];
    float const v01 = img[(fy1 * w + fx) * c + ch];
    float const v11 = img[(fy1 * w + fx1) * c + ch];
    return v00 * (w0 * w2) + v10 * (w1 * w2) + v01 * (w0 * w3)
        + v11 * (w1 * w3);
}

uint8_t
linear_at_u8 (uint8_t const* img, int w, int h, float x, float y)
{
    x = std::max(0.0f, std::min(static_cast<float>(w - 1), x));
    y = std::max(0.0f, std::min(static_cast<float>(h - 1), y));
    int const fx = static_cast<int>(x), fy = static_cast<int>(y);
    int const fx1 = std::min(fx + 1, w - 1), fy1 = std::min(fy + 1, h - 1);
    float const w1 = x - static_cast<float>(fx), w0 = 1.0f - w1;
    float const w3 = y - static_cast<float>(fy), w2 = 1.0f - w3;
    float const v = (float)img[fy * w + fx] * (w0 * w2)
        + (float)img[fy * w + fx1] * (w1 * w2)
        + (float)img[fy1 * w + fx] * (w0 * w3)
        + (float)img[fy1 * w + fx1] * (w1 * w3) + 0.5f;
    return static_cast<uint8_t>(v);
}

/* ------------------------------------------------------------------ */
/* scene state                                                        */
/* ------------------------------------------------------------------ */

struct Block { double v[16]; Block() { std::fill(v, v + 16, 0.0); } };

struct Port
{
    int w, h, n_sub;
    double flen, inv_flen;
    std::vector<float> main_grad, shading, shading_grad;
    std::vector<int> sub_w, sub_h;
    std::vector<std::vector<float> > sub_grad, sub_hess;
    std::vector<double> Mi, ti;
    /* StereoView::get_image(): unblurred float images, 3 channels (only the
     * use_sgm = false mode of create_subview_surfaces reads them) */
    std::vector<float> main_image;
    std::vector<std::vector<float> > sub_image;
    /* Surface::depth: the initialisation depth the nodes are made from */
    std::vector<float> init_depth;

    int scale, ps, npx, npy, sx, sy;
    std::vector<double> nodes;
    std::vector<uint8_t> node_valid, patch_valid, vis_ids;
    std::vector<uint32_t> vis_off;

    /* system in the reference's BSC layout */
    std::vector<double> g, Hvals, Pvals;
    std::vector<uint64_t> Houter, Hinner, Pouter, Pinner;

    int n_nodes () const { return (npx + 1) * (npy + 1); }
    void node_ids (int patch, int* ids) const
    {
        /* Surface::fill_node_ids_for_patch, lib/surface.cc:286-295 */
        int const idx = patch % npx, idy = patch / npx, ns = npx + 1;
        ids[0] = idy * ns + idx; ids[1] = ids[0] + 1;
        ids[2] = (idy + 1) * ns + idx; ids[3] = ids[2] + 1;
    }
    void patch_nodes16 (int patch, double* n16) const
    {
        int ids[4];
        node_ids(patch, ids);
        for (int n = 0; n < 4; ++n)
            for (int c = 0; c < 4; ++c)
                n16[n * 4 + c] = nodes[ids[n] * 4 + c];
    }
};

int
sampling_for_scale (int scale)
{
    int s = 4;
    if (scale < 5) s = 2;
    if (scale < 3) s = 1;
    return s;
}

/* SurfacePatch::fill_values_at_pixels sample order, lib/surface_patch.cc:
 * 86-119: pid steps by `sub`, rows that are not multiples of `sub` skipped. */
void
sample_pids (int size, int sub, std::vector<int>* pids)
{
    pids->clear();
    for (int pid = 0; pid < size * size;)
    {
        pids->push_back(pid);
        if (sub > 1)
        {
            pid += sub;
            if ((pid / size) % sub == 1)
                pid += size * (sub - 1);
        }
        else
            pid += 1;
    }
}

/* GaussNewtonStep::construct, lib/gauss_newton_step.cc:33-143, with
 * jacobian_entries_for_patch (:145-244) and
 * fill_gradient_and_hessian_entries (:246-518, scalar branch :334-383). */
void
gn_construct (Port& P, uint8_t const* active, double const* light,
    double regularization, double light_surf_reg)
{
    int const nn = P.n_nodes();
    std::size_t const num_params = static_cast<std::size_t>(nn) * 4;
    int const ps = P.ps;
    int const sampling = sampling_for_scale(P.scale);
    std::vector<double> table(static_cast<std::size_t>(ps) * ps * 96);
    for (int i = 0; i < ps * ps; ++i)
        node_derivatives(((i % ps) + 0.5) / ps, ((i / ps) + 0.5) / ps,
            static_cast<double>(ps), &table[i * 96]);
    std::vector<int> pids;
    sample_pids(ps, sampling, &pids);

    P.g.assign(num_params, 0.0);
    std::map<std::size_t, Block> blocks;
    std::vector<double> jac_entries, j_grad;

    for (int patch = 0; patch < P.npx * P.npy; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        int ids[4];
        P.node_ids(patch, ids);
        if (!active[ids[0]] && !active[ids[1]] && !active[ids[2]]
            && !active[ids[3]])
            continue;

        double sub_g[16], sub_h[256];
        std::fill(sub_g, sub_g + 16, 0.0);
        std::fill(sub_h, sub_h + 256, 0.0);
        double n16[16], coeffs[4][4];
        P.patch_nodes16(patch, n16);
        patch_coefficients(n16, coeffs);
        uint32_t const v0 = P.vis_off[patch];
        int const num_subs = static_cast<int>(P.vis_off[patch + 1] - v0);
        jac_entries.assign(static_cast<std::size_t>(num_subs) * 32, 0.0);
        j_grad.assign(static_cast<std::size_t>(num_subs) * 2, 0.0);
        int const px0 = P.sx + (patch % P.npx) * ps;
        int const py0 = P.sy + (patch / P.npx) * ps;

        for (std::size_t s = 0; s < pids.size(); ++s)
        {
            int const pid = pids[s];
            int const i = pid % ps, j = pid / ps;
            double const pixx = i + px0, pixy = j + py0;
            double e[6];
            patch_evaluate(coeffs, (i + 0.5) / ps, (j + 0.5) / ps, e);
            double const depth = e[0];
            double const ddx = e[1] / ps, ddy = e[2] / ps;
            double const d2xy = e[3] / (ps * ps), d2xx = e[4] / (ps * ps);
            double const d2yy = e[5] / (ps * ps);
            std::size_t const pix = static_cast<std::size_t>(pixy) * P.w
                + static_cast<std::size_t>(pixx);
            double const gm[2] = { P.main_grad[pix * 2], P.main_grad[pix * 2 + 1] };
            double const* dn00 = &table[pid * 96];

            for (int jn = 0; jn < num_subs; ++jn)
            {
                int const sub = P.vis_ids[v0 + jn];
                Corr C;
                C.update(&P.Mi[sub * 9], &P.ti[sub * 3], pixx + 0.5,
                    pixy + 0.5, depth, ddx, ddy);
                double proj[2], jac[4];
                C.fill(proj);
                C.fill_jacobian(jac);
                proj[0] -= 0.5; proj[1] -= 0.5;
                float const* sg = P.sub_grad[sub].data();
                float const* sh = P.sub_hess[sub].data();
                int const sw = P.sub_w[sub], shh = P.sub_h[sub];
                double gs[2], hs[4];
                gs[0] = linear_at_f(sg, sw, shh, 2, (float)proj[0], (float)proj[1], 0);
                gs[1] = linear_at_f(sg, sw, shh, 2, (float)proj[0], (float)proj[1], 1);
                hs[0] = linear_at_f(sh, sw, shh, 3, (float)proj[0], (float)proj[1], 0);
                hs[1] = linear_at_f(sh, sw, shh, 3, (float)proj[0], (float)proj[1], 1);
                hs[2] = hs[1];
                hs[3] = linear_at_f(sh, sw, shh, 3, (float)proj[0], (float)proj[1], 2);
                j_grad[jn * 2 + 0] = jac[0] * gs[0] + jac[1] * gs[1];
                j_grad[jn * 2 + 1] = jac[2] * gs[0] + jac[3] * gs[1];
                double c_dn[32], jac_dn[32];
                C.fill_derivative(dn00, c_dn);
                C.fill_jacobian_derivative_grad(gs, dn00, jac_dn);
                double const jh[4] = {
                    jac[0] * hs[0] + jac[1] * hs[2], jac[0] * hs[1] + jac[1] * hs[3],
                    jac[2] * hs[0] + jac[3] * hs[2], jac[2] * hs[1] + jac[3] * hs[3] };
                for (int col = 0; col < 16; ++col)
                {
                    jac_entries[(jn * 16 + col) * 2 + 0] = jac_dn[col * 2]
                        + jh[0] * c_dn[col * 2] + jh[1] * c_dn[col * 2 + 1];
                    jac_entries[(jn * 16 + col) * 2 + 1] = jac_dn[col * 2 + 1]
                        + jh[2] * c_dn[col * 2] + jh[3] * c_dn[col * 2 + 1];
                }
            }

            double div[6], div_deriv[96], normal_deriv[48];
            double basic = 0.0;
            double const x = pixx + 0.5 - static_cast<double>(P.w) / 2.0;
            double const y = pixy + 0.5 - static_cast<double>(P.h) / 2.0;
            if (regularization > 0.0)
            {
                basic = regularization * 0.005
                    / std::max(0.03, std::abs(gm[0]) + std::abs(gm[1]));
                normal_divergence(x, y, P.flen, depth, ddx, ddy, d2xy, d2xx,
                    d2yy, div);
                normal_divergence_deriv(dn00, x, y, P.flen, depth, ddx, ddy,
                    d2xy, d2xx, d2yy, div_deriv);
                normal_derivative(dn00, x, y, P.flen, depth, ddx, ddy,
                    normal_deriv);
            }

            /* photometric terms (scalar branch) */
            for (int jn = 0; jn < num_subs; ++jn)
            {
                double diff[2], wt[2];
                for (int k = 0; k < 2; ++k)
                {
                    diff[k] = j_grad[jn * 2 + k] - gm[k];
                    wt[k] = 1.0 / (R_FACTOR + std::abs(diff[k]));
                }
                double const* je = &jac_entries[jn * 32];
                for (int col = 0; col < 16; ++col)
                {
                    sub_g[col] += diff[0] * wt[0] * je[col * 2]
                        + diff[1] * wt[1] * je[col * 2 + 1];
                    for (int col2 = col; col2 < 16; ++col2)
                        sub_h[col * 16 + col2] +=
                            je[col * 2] * wt[0] * je[col2 * 2]
                            + je[col * 2 + 1] * wt[1] * je[col2 * 2 + 1];
                }
                for (int j2 = jn + 1; j2 < num_subs; ++j2)
                {
                    double sd[2], sw2[2];
                    for (int k = 0; k < 2; ++k)
                    {
                        sd[k] = j_grad[jn * 2 + k] - j_grad[j2 * 2 + k];
                        sw2[k] = 1.0 / (R_FACTOR + std::abs(sd[k]));
                    }
                    double const* j2e = &jac_entries[j2 * 32];
                    for (int col = 0; col < 16; ++col)
                    {
                        double const jace0 = (je[col * 2] - j2e[col * 2]) * sw2[0];
                        double const jace1 = (je[col * 2 + 1] - j2e[col * 2 + 1])
                            * sw2[1];
                        sub_g[col] += jace0 * sd[0] + jace1 * sd[1];
                        for (int col2 = col; col2 < 16; ++col2)
                            sub_h[col * 16 + col2] +=
                                jace0 * (je[col2 * 2] - j2e[col2 * 2])
                                + jace1 * (je[col2 * 2 + 1] - j2e[col2 * 2 + 1]);
                    }
                }
            }
            if (regularization <= 0.0)
                continue;

            double const num_diffs = (num_subs * (num_subs + 1)) / 2;
            basic *= num_diffs;
            if (light == nullptr || light_surf_reg > 0.0)
            {
                double geom_weight = 1.0;
                if (light != nullptr)
                    geom_weight *= light_surf_reg / 100;
                for (int v = 0; v < 6; ++v)
                {
                    double const weight = geom_weight
                        / (R_FACTOR + std::abs(div[v]));
                    for (int col = 0; col < 16; ++col)
                    {
                        sub_g[col] += div_deriv[16 * v + col] * div[v] * basic
                            * weight;
                        for (int col2 = col; col2 < 16; ++col2)
                            sub_h[col * 16 + col2] += div_deriv[v * 16 + col]
                                * div_deriv[v * 16 + col2] * basic * weight;
                    }
                }
                if (light == nullptr)
                    continue;
            }

            /* shading term, :419-515 */
            double normal[3], sh_d[48], sh[16];
            fill_normal(x, y, P.inv_flen, depth, ddx, ddy, normal);
            sh_derivative_4_band(normal, sh_d);
            sh_evaluate_4_band(normal, sh);
            double shading = 0.0;
            for (int l = 0; l < 16; ++l) shading += light[l] * sh[l];
            double lig[2] = { P.shading_grad[pix * 2], P.shading_grad[pix * 2 + 1] };
            double const liv = P.shading[pix];
            double const shading_weight = 0.001 * num_diffs
                / (R_FACTOR + std::abs(lig[0]) + std::abs(lig[1]));
            if (std::sqrt(lig[0] * lig[0] + lig[1] * lig[1]) < 1e-10)
                continue;
            if (shading * shading < 1e-10 || liv * liv < 1e-10)
                continue;
            double sgrad[2] = { 0.0, 0.0 };
            for (int l = 1; l < 16; ++l)
            {
                sgrad[0] += light[l] * (sh_d[l * 3] * div[0]
                    + sh_d[l * 3 + 1] * div[1] + sh_d[l * 3 + 2] * div[2]);
                sgrad[1] += light[l] * (sh_d[l * 3] * div[3]
                    + sh_d[l * 3 + 1] * div[4] + sh_d[l * 3 + 2] * div[5]);
            }
            double const render[2] = { sgrad[0] / shading, sgrad[1] / shading };
            lig[0] *= 1.0 / liv; lig[1] *= 1.0 / liv;
            double const err[2] = { render[0] - lig[0], render[1] - lig[1] };
            double rd[32];
            for (int col = 0; col < 16; ++col)
            {
                double sd = 0.0, gd0 = 0.0, gd1 = 0.0;
                for (int l = 1; l < 16; ++l)
                {
                    sd += light[l] * (sh_d[l * 3] * normal_deriv[col]
                        + sh_d[l * 3 + 1] * normal_deriv[16 + col]
                        + sh_d[l * 3 + 2] * normal_deriv[32 + col]);
                    gd0 += light[l] * (sh_d[l * 3] * div_deriv[col]
                        + sh_d[l * 3 + 1] * div_deriv[16 + col]
                        + sh_d[l * 3 + 2] * div_deriv[32 + col]);
                    gd1 += light[l] * (sh_d[l * 3] * div_deriv[48 + col]
                        + sh_d[l * 3 + 1] * div_deriv[64 + col]
                        + sh_d[l * 3 + 2] * div_deriv[80 + col]);
                }
                rd[col * 2] = (gd0 * shading - sgrad[0] * sd)
                    / (shading * shading);
                rd[col * 2 + 1] = (gd1 * shading - sgrad[1] * sd)
                    / (shading * shading);
            }
            double const w0 = shading_weight / (R_FACTOR + std::abs(err[0]));
            double const w1 = shading_weight / (R_FACTOR + std::abs(err[1]));
            for (int col = 0; col < 16; ++col)
            {
                sub_g[col] += err[0] * rd[col * 2] * w0
                    + err[1] * rd[col * 2 + 1] * w1;
                for (int col2 = col; col2 < 16; ++col2)
                    sub_h[col * 16 + col2] += rd[col * 2] * rd[col2 * 2] * w0
                        + rd[col * 2 + 1] * rd[col2 * 2 + 1] * w1;
            }
        }

        /* scatter, :88-121 */
        for (int node = 0; node < 4; ++node)
        {
            if (!active[ids[node]]) continue;
            for (int v = 0; v < 4; ++v)
                P.g[ids[node] * 4 + v] += sub_g[node * 4 + v];
        }
        for (std::size_t n1 = 0; n1 < 16; ++n1)
        {
            if (!active[ids[n1 / 4]]) continue;
            for (std::size_t n2 = n1; n2 < 16; ++n2)
            {
                if (!active[ids[n2 / 4]]) continue;
                std::size_t const id1 = ids[n1 / 4] * num_params + ids[n2 / 4];
                std::size_t const id2 = ids[n1 / 4] + num_params * ids[n2 / 4];
                std::size_t const ox = n1 % 4, oy = n2 % 4;
                blocks[id1].v[ox + 4 * oy] += sub_h[n1 * 16 + n2];
                if (n1 != n2)
                    blocks[id2].v[ox * 4 + oy] += sub_h[n1 * 16 + n2];
            }
        }
    }

    /* set_from_blocks + transpose = column-major block order, :123-142,
     * lib/block_sparse_matrix.h:155-190,241-274 */
    struct E { std::size_t row, col; Block const* b; };
    std::vector<E> es;
    for (auto const& kv : blocks)
        es.push_back(E{ 4 * kv.first % num_params, 4 * kv.first / num_params,
            &kv.second });
    std::stable_sort(es.begin(), es.end(), [](E const& a, E const& b)
        { return a.col != b.col ? a.col < b.col : a.row < b.row; });
    P.Hvals.clear(); P.Hinner.clear(); P.Houter.assign(nn + 1, 0);
    P.Pvals.clear(); P.Pinner.clear(); P.Pouter.assign(nn + 1, 0);
    for (E const& e : es)
    {
        P.Houter[e.col / 4 + 1] += 1;
        P.Hinner.push_back(e.row);
        P.Hvals.insert(P.Hvals.end(), e.b->v, e.b->v + 16);
        if (e.row == e.col)
        {
            P.Pouter[e.col / 4 + 1] += 1;
            P.Pinner.push_back(e.row);
            /* invert_blocks_inplace, block_sparse_matrix.h:300-316 */
            double b[16];
            std::copy(e.b->v, e.b->v + 16, b);
            ldl_inverse(b, 4);
            bool nan = false;
            for (int i = 0; i < 16; ++i) nan = nan || std::isnan(b[i]);
            if (nan) std::copy(e.b->v, e.b->v + 16, b);
            P.Pvals.insert(P.Pvals.end(), b, b + 16);
        }
    }
    for (int i = 0; i < nn; ++i)
    {
        P.Houter[i + 1] += P.Houter[i];
        P.Pouter[i + 1] += P.Pouter[i];
    }
}

/* BlockSparseMatrix::multiply, lib/block_sparse_matrix.h:276-298 */
void
bsm_multiply (std::vector<double> const& vals,
    std::vector<uint64_t> const& outer, std::vector<uint64_t> const& inner,
    std::vector<double> const& x, std::vector<double>* y)
{
    y->assign(x.size(), 0.0);
    for (std::size_t i = 0; i + 1 < outer.size(); ++i)
        for (uint64_t id = outer[i]; id < outer[i + 1]; ++id)
        {
            std::size_t ret_id = inner[id];
            int bid = 0;
            for (int br = 0; br < 4; ++br, ++ret_id)
                for (int bc = 0; bc < 4; ++bc)
                    (*y)[ret_id] += vals[id * 16 + bid++] * x[i * 4 + bc];
        }
}

double
dot (std::vector<double> const& a, std::vector<double> const& b)
{
    double s = 0.0;
    for (std::size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

/* ConjugateGradient::solve, lib/conjugate_gradient.h:72-202 (b = -g) */
void
cg_solve (Port const& P, int max_iter, double err_tol, double q_tol,
    std::vector<double>* x, int* iters, int* info)
{
    std::size_t const n = P.g.size();
    std::vector<double> b(n), r, d, z, Ad, tmp(n);
    for (std::size_t i = 0; i < n; ++i) b[i] = -P.g[i];
    if (err_tol < 0.0)
        err_tol = std::sqrt(dot(P.g, P.g)) * 0.01;
    x->assign(n, 0.0);
    r = b;
    bsm_multiply(P.Pvals, P.Pouter, P.Pinner, r, &z);
    double r_dot_r = dot(z, r);
    d = z;
    for (std::size_t i = 0; i < n; ++i) tmp[i] = b[i] + r[i];
    double Q0 = -1.0 * dot(*x, tmp);
    int it = 1;
    for (; it < max_iter; ++it)
    {
        bsm_multiply(P.Hvals, P.Houter, P.Hinner, d, &Ad);
        double const alpha = r_dot_r / dot(d, Ad);
        for (std::size_t i = 0; i < n; ++i) (*x)[i] += d[i] * alpha;
        for (std::size_t i = 0; i < n; ++i) r[i] -= Ad[i] * alpha;
        double new_r_dot_r = dot(r, r);
        if (new_r_dot_r < err_tol) { *iters = it; *info = 0; return; }
        for (std::size_t i = 0; i < n; ++i) tmp[i] = b[i] + r[i];
        double const Q1 = -1.0 * dot(*x, tmp);
        double const zeta = it * (Q1 - Q0) / Q1;
        if (zeta < q_tol) { *iters = it; *info = 0; return; }
        Q0 = Q1;
        bsm_multiply(P.Pvals, P.Pouter, P.Pinner, r, &z);
        new_r_dot_r = dot(z, r);
        double const beta = new_r_dot_r / r_dot_r;
        for (std::size_t i = 0; i < n; ++i) d[i] = z[i] + d[i] * beta;
        r_dot_r = new_r_dot_r;
    }
    *iters = it; *info = 1;
}

/* fill_node_reprojections, lib/depth_optimizer.cc:647-677 */
void
node_reprojections (Port const& P, uint8_t const* active,
    std::vector<int>* node_of, std::vector<double>* proj)
{
    node_of->clear(); proj->clear();
    for (int patch = 0; patch < P.npx * P.npy; ++patch)
    {
        if (!P.patch_valid[patch]) continue;
        int ids[4];
        P.node_ids(patch, ids);
        if (active[ids[0]] + active[ids[1]] + active[ids[2]] + active[ids[3]] == 0)
            continue;
        double n16[16], coeffs[4][4];
        P.patch_nodes16(patch, n16);
        patch_coefficients(n16, coeffs);
        int const px0 = P.sx + (patch % P.npx) * P.ps;
        int const py0 = P.sy + (patch / P.npx) * P.ps;
        for (uint32_t k = P.vis_off[patch]; k < P.vis_off[patch + 1]; ++k)
            for (int pid = 0; pid < P.ps * P.ps; ++pid)
            {
                int const i = pid % P.ps, j = pid / P.ps;
                double e[6];
                patch_evaluate(coeffs, (i + 0.5) / P.ps, (j + 0.5) / P.ps, e);
                int const sub = P.vis_ids[k];
                Corr C;
                C.update(&P.Mi[sub * 9], &P.ti[sub * 3], i + px0, j + py0,
                    e[0], 0.0, 0.0);
                double pr[2];
                C.fill(pr);
                for (int n = 0; n < 4; ++n)
                {
                    node_of->push_back(ids[n]);
                    proj->push_back(pr[0]);
                    proj->push_back(pr[1]);
                }
            }
    }
}

/* lib/depth_optimizer.cc:271-303 */
void
update_nodes (Port& P, std::vector<double> const& delta, double thresh,
    bool full_opt, std::vector<uint8_t>* active, uint64_t* n_active,
    double* mean_shift)
{
    std::vector<int> node_of, dummy;
    std::vector<double> p1, p2;
    node_reprojections(P, active->data(), &node_of, &p1);
    for (int i = 0; i < P.n_nodes(); ++i)       /* Surface::update_nodes */
        if (P.node_valid[i])
            for (int c = 0; c < 4; ++c)
                P.nodes[i * 4 + c] += delta[i * 4 + c];
    node_reprojections(P, active->data(), &dummy, &p2);
    double sum = 0.0;
    std::vector<double> diffs(node_of.size());
    for (std::size_t k = 0; k < node_of.size(); ++k)
    {
        double const ex = p1[2 * k] - p2[2 * k], ey = p1[2 * k + 1] - p2[2 * k + 1];
        diffs[k] = std::sqrt(ex * ex + ey * ey);
        sum += diffs[k];
    }
    if (mean_shift) *mean_shift = sum / static_cast<double>(node_of.size());
    if (!full_opt)
    {
        std::fill(active->begin(), active->end(), 0);
        for (std::size_t k = 0; k < node_of.size(); ++k)
            if (diffs[k] > thresh)
                (*active)[node_of[k]] = 1;
    }
    uint64_t cnt = 0;
    for (uint8_t a : *active) cnt += (a == 1);
    if (n_active) *n_active = cnt;
}

/* SGM, lib/sgm_stereo.cc (SSE branch semantics) -------------------- */

void
census_filter (uint8_t const* img, int w, int h, int c, uint64_t* out)
{
    /* :126-148 */
    std::fill(out, out + static_cast<std::size_t>(w) * h * c, 0);
    for (int x = 4; x < w - 5; ++x)
        for (int y = 3; y < h - 4; ++y)
            for (int d = 0; d < c; ++d)
            {
                uint8_t const thr = img[(static_cast<std::size_t>(y) * w + x) * c + d];
                if (thr == 0) continue;
                uint64_t census = 0;
                for (int i = x - 4; i < x + 5; ++i)
                    for (int j = y - 3; j < y + 4; ++j)
                    {
                        census *= 2;
                        if (thr < img[(static_cast<std::size_t>(j) * w + i) * c + d])
                            census += 1;
                    }
                out[(static_cast<std::size_t>(y) * w + x) * c + d] = census;
            }
}

} /* namespace */

/* ------------------------------------------------------------------ */
/* surface topology: creation from a depth map, subdivision, clean-up  */
/* ------------------------------------------------------------------ */

void remove_nodes_without_patch (Port& P);

/* Surface::fill_holes, lib/surface.cc:628-649: a patch wherever its four
 * nodes exist. */
int
fill_holes (Port& P)
{
    int filled = 0;
    for (int patch = 0; patch < P.npx * P.npy; ++patch)
    {
        if (P.patch_valid[patch])
            continue;
        int ids[4];
        P.node_ids(patch, ids);
        if (P.node_valid[ids[0]] && P.node_valid[ids[1]]
            && P.node_valid[ids[2]] && P.node_valid[ids[3]])
        {
            P.patch_valid[patch] = 1;
            filled += 1;
        }
    }
    return filled;
}

/* Surface::initialize_node_from_depth, lib/surface.cc:665-760 */
void
initialize_node_from_depth (Port& P, int idx, int idy)
{
    int const ns = P.npx + 1;
    int const x = idx * P.ps + P.sx, y = idy * P.ps + P.sy;
    if (P.node_valid[idy * ns + idx])
        return;
    int const win = P.ps / 2;
    std::vector<double> d[4];
    int const lo_i[4] = { -win, 0, -win, 0 }, hi_i[4] = { 0, win, 0, win };
    int const lo_j[4] = { -win, -win, 0, 0 }, hi_j[4] = { 0, 0, win, win };
    for (int q = 0; q < 4; ++q)
        for (int i = lo_i[q]; i < hi_i[q]; ++i)
            for (int j = lo_j[q]; j < hi_j[q]; ++j)
                if (x + i >= 0 && x + i < P.w && y + j >= 0 && y + j < P.h
                    && P.init_depth[static_cast<std::size_t>(y + j) * P.w
                        + x + i] > 0.0)
                    d[q].push_back(P.init_depth[static_cast<std::size_t>(
                        y + j) * P.w + x + i]);

    int num_non_zeros = 4;
    double avg[4];
    std::vector<double> all;
    for (int i = 0; i < 4; ++i)
    {
        all.insert(all.end(), d[i].begin(), d[i].end());
        if (d[i].empty())
        {
            avg[i] = 0.0;
            num_non_zeros -= 1;
            continue;
        }
        avg[i] = *std::min_element(d[i].begin(), d[i].end());
    }
    if (num_non_zeros == 0)
        return;
    if (all.size() < 2)
        return;
    std::nth_element(all.begin(), all.begin() + all.size() / 2, all.end());

    double f = all[all.size() / 2], dx = 0.0, dy = 0.0, dxy = 0.0;
    if (num_non_zeros == 4)
    {
        dx = ((avg[1] + avg[3]) - (avg[0] + avg[2])) / 2.0;
        dy = ((avg[2] + avg[3]) - (avg[0] + avg[1])) / 2.0;
        dxy = ((avg[3] - avg[2]) - (avg[1] - avg[0]));
    }
    else
    {
        if ((avg[1] == 0 || avg[0] == 0) && avg[3] != 0 && avg[2] != 0)
            dx = (avg[3] - avg[2]);
        else if ((avg[2] == 0 || avg[3] == 0) && avg[1] != 0 && avg[0] != 0)
            dx = (avg[1] - avg[0]);
        if ((avg[0] == 0 || avg[2] == 0) && avg[3] != 0 && avg[1] != 0)
            dy = (avg[3] - avg[1]);
        else if ((avg[1] == 0 || avg[2] == 0) && avg[0] != 0 && avg[2] != 0)
            dy = (avg[2] - avg[0]);
    }
    int const node = idy * ns + idx;
    P.node_valid[node] = 1;
    P.nodes[node * 4 + 0] = f; P.nodes[node * 4 + 1] = dx;
    P.nodes[node * 4 + 2] = dy; P.nodes[node * 4 + 3] = dxy;
}

/* Surface::fill_patches_from_depth, lib/surface.cc:141-152 */
void
fill_patches_from_depth (Port& P)
{
    for (int i = 0; i < P.npx + 1; ++i)
        for (int j = 0; j < P.npy + 1; ++j)
            initialize_node_from_depth(P, i, j);
    fill_holes(P);
    remove_nodes_without_patch(P);
}

/* Surface::Surface(bundle, view, scale, init_depth), lib/surface.cc:19-53 */
void
surface_create (Port& P, int scale, float const* init_depth)
{
    P.scale = scale;
    P.ps = 1 << scale;
    P.npx = (P.w - 2) / P.ps - 1;
    P.npy = (P.h - 2) / P.ps - 1;
    P.sx = (P.w - P.npx * P.ps) / 2;
    P.sy = (P.h - P.npy * P.ps) / 2;
    P.nodes.assign(static_cast<std::size_t>(P.n_nodes()) * 4, 0.0);
    P.node_valid.assign(P.n_nodes(), 0);
    P.patch_valid.assign(P.npx * P.npy, 0);
    P.vis_off.assign(P.npx * P.npy + 1, 0);
    P.vis_ids.clear();
    P.init_depth.assign(static_cast<std::size_t>(P.w) * P.h, 0.0f);
    for (std::size_t p = 0; p < P.init_depth.size(); ++p)
        if (init_depth[p] > 0.0)
            P.init_depth[p] = init_depth[p];
    fill_patches_from_depth(P);
}

/* Surface::subdivide_patches, lib/surface.cc:983-1107 */
void
subdivide_patches (Port& P)
{
    P.scale -= 1;
    P.ps = 1 << P.scale;
    int new_npx = (P.w - 2) / P.ps, new_npy = (P.h - 2) / P.ps;
    int offset_x = new_npx - P.npx * 2, offset_y = new_npy - P.npy * 2;
    if (offset_x >= 2)
    {
        new_npx = P.npx * 2 + 2;
        P.sx = (P.w - new_npx * P.ps) / 2;
        offset_x = 1;
    }
    else
    {
        offset_x = 0;
        new_npx = P.npx * 2;
    }
    if (offset_y >= 2)
    {
        new_npy = P.npy * 2 + 2;
        P.sy = (P.h - new_npy * P.ps) / 2;
        offset_y = 1;
    }
    else
    {
        offset_y = 0;
        new_npy = P.npy * 2;
    }
    int const new_ns = new_npx + 1;
    std::vector<double> new_nodes(static_cast<std::size_t>(new_ns)
        * (new_npy + 1) * 4, 0.0);
    std::vector<uint8_t> new_valid(static_cast<std::size_t>(new_ns)
        * (new_npy + 1), 0);

    /* five new nodes per patch; later patches overwrite shared edge nodes */
    double const at[5][2] = { {0.5, 0.0}, {0.0, 0.5}, {0.5, 0.5}, {1.0, 0.5},
        {0.5, 1.0} };
    int const off[5][2] = { {1, 0}, {0, 1}, {1, 1}, {2, 1}, {1, 2} };
    for (int patch = 0; patch < P.npx * P.npy; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        int const new_idx = 2 * (patch % P.npx) + offset_x;
        int const new_idy = 2 * (patch / P.npx) + offset_y;
        double n16[16], coeffs[4][4];
        P.patch_nodes16(patch, n16);
        patch_coefficients(n16, coeffs);
        for (int k = 0; k < 5; ++k)
        {
            double out6[6];
            patch_evaluate(coeffs, at[k][0], at[k][1], out6);
            int const node = new_idx + off[k][0]
                + new_ns * (new_idy + off[k][1]);
            new_valid[node] = 1;
            new_nodes[node * 4 + 0] = out6[0];
            new_nodes[node * 4 + 1] = out6[1] / 2;
            new_nodes[node * 4 + 2] = out6[2] / 2;
            new_nodes[node * 4 + 3] = out6[3] / 4;
        }
    }
    /* old nodes, derivatives rescaled to the new patch size */
    int const ns = P.npx + 1;
    for (int node = 0; node < P.n_nodes(); ++node)
    {
        if (!P.node_valid[node])
            continue;
        int const new_node = 2 * (node % ns) + offset_x
            + new_ns * (2 * (node / ns) + offset_y);
        new_valid[new_node] = 1;
        new_nodes[new_node * 4 + 0] = P.nodes[node * 4 + 0];
        new_nodes[new_node * 4 + 1] = P.nodes[node * 4 + 1] / 2;
        new_nodes[new_node * 4 + 2] = P.nodes[node * 4 + 2] / 2;
        new_nodes[new_node * 4 + 3] = P.nodes[node * 4 + 3] / 4;
    }
    P.npx = new_npx;
    P.npy = new_npy;
    P.nodes.swap(new_nodes);
    P.node_valid.swap(new_valid);
    P.patch_valid.assign(P.npx * P.npy, 0);
    P.vis_off.assign(P.npx * P.npy + 1, 0);
    P.vis_ids.clear();
    fill_holes(P);
    remove_nodes_without_patch(P);
}

/* Surface::expand, lib/surface.cc:482-628: two rounds of new nodes at the rim
 * (largest of up to eight extrapolations from neighbour triples, with the 0.9
 * hysteresis of check_swap_nodes :472-480), then fill_holes. Returns the
 * patches created. */
int
expand (Port& P)
{
    int const ns = P.npx + 1, nn = P.n_nodes();
    std::vector<uint8_t> has_new(nn, 0);
    std::vector<double> new_f(nn, 0.0);
    for (int iter = 0; iter < 2; ++iter)
    {
        for (int node = 0; node < nn; ++node)
        {
            if (P.node_valid[node] && !has_new[node])
                continue;
            int const ix = node % ns, iy = node / ns;
            /* fill_node_neighbors: 0..7 = NW N NE W E SW S SE, absent when
             * outside the grid or not valid */
            bool ok[8];
            double f[8], dx[8], dy[8];
            int k = 0;
            for (int oy = -1; oy < 2; ++oy)
                for (int ox = -1; ox < 2; ++ox)
                {
                    if (ox == 0 && oy == 0)
                        continue;
                    int const qx = ix + ox, qy = iy + oy;
                    ok[k] = qx >= 0 && qy >= 0 && qx <= P.npx && qy <= P.npy
                        && P.node_valid[qy * ns + qx];
                    if (ok[k])
                    {
                        double const* n = &P.nodes[(qy * ns + qx) * 4];
                        f[k] = n[0]; dx[k] = n[1]; dy[k] = n[2];
                    }
                    k += 1;
                }
            auto offer = [&] (double value)
            {
                /* check_swap_nodes */
                if (!has_new[node] || value * 0.9 > new_f[node])
                {
                    has_new[node] = 1;
                    new_f[node] = value;
                }
            };
            if (ok[0] && ok[1] && ok[3])
                offer(((f[3] + dx[3] / 2.0) + (f[1] + dy[1] / 2.0)) / 2.0);
            if (ok[1] && ok[2] && ok[4])
                offer(((f[4] - dx[4] / 2.0) + (f[1] + dy[1] / 2.0)) / 2.0);
            if (ok[3] && ok[5] && ok[6])
                offer(((f[3] + dx[3] / 2.0) + (f[6] - dy[6] / 2.0)) / 2.0);
            if (ok[4] && ok[6] && ok[7])
                offer(((f[4] - dx[4] / 2.0) + (f[6] - dy[6] / 2.0)) / 2.0);
            if (ok[0] && ok[1] && ok[2])
                offer(((f[0] + dy[0] / 2.0) + (f[1] + dy[1] / 2.0)
                    + (f[2] + dy[2] / 2.0)) / 3.0);
            if (ok[0] && ok[3] && ok[5])
                offer(((f[0] + dx[0] / 2.0) + (f[3] + dx[3] / 2.0)
                    + (f[5] + dx[5] / 2.0)) / 3.0);
            if (ok[5] && ok[6] && ok[7])
                offer(((f[5] - dy[5] / 2.0) + (f[6] - dy[6] / 2.0)
                    + (f[7] - dy[7] / 2.0)) / 3.0);
            if (ok[2] && ok[4] && ok[7])
                offer(((f[2] - dx[2] / 2.0) + (f[4] - dx[4] / 2.0)
                    + (f[7] - dx[7] / 2.0)) / 3.0);
        }
        for (int node = 0; node < nn; ++node)
            if (has_new[node])
            {
                P.node_valid[node] = 1;
                P.nodes[node * 4 + 0] = new_f[node];
                P.nodes[node * 4 + 1] = 0.0;
                P.nodes[node * 4 + 2] = 0.0;
                P.nodes[node * 4 + 3] = 0.0;
            }
    }
    int const filled = fill_holes(P);
    remove_nodes_without_patch(P);
    return filled;
}

/* Surface::remove_isolated_patches, lib/surface.cc:887-927: sequential, x
 * outer / y inner, deletions feed the counts of the patches visited later. */
void
remove_isolated_patches (Port& P)
{
    for (int x = 0; x < P.npx; ++x)
        for (int y = 0; y < P.npy; ++y)
        {
            if (!P.patch_valid[y * P.npx + x])
                continue;
            int valid_neighbors = 0;
            for (int dx = -1; dx < 2; ++dx)
                for (int dy = -1; dy < 2; ++dy)
                {
                    if (dx == 0 && dy == 0)
                        continue;
                    int const qx = x + dx, qy = y + dy;
                    if (qx < 0 || qy < 0 || qx > P.npx - 1 || qy > P.npy - 1)
                        continue;
                    valid_neighbors += P.patch_valid[qy * P.npx + qx] ? 1 : 0;
                }
            if (valid_neighbors < 3)
                P.patch_valid[y * P.npx + x] = 0;
        }
    remove_nodes_without_patch(P);
}

/* ------------------------------------------------------------------ */
/* visibility lists, boundary cutting, bilateral filter               */
/* ------------------------------------------------------------------ */

/* SurfacePatch::fill_values_at_pixels (lib/surface_patch.cc:57-120), all
 * pixels: depth and first derivatives (already / size) in pid order. */
void
patch_pixels (Port const& P, int patch, std::vector<double>* w,
    std::vector<double>* wx, std::vector<double>* wy)
{
    double n16[16], coeffs[4][4];
    P.patch_nodes16(patch, n16);
    patch_coefficients(n16, coeffs);
    int const size = P.ps;
    w->resize(size * size);
    if (wx) { wx->resize(size * size); wy->resize(size * size); }
    for (int pid = 0; pid < size * size; ++pid)
    {
        double x = static_cast<double>(pid % size);
        double y = static_cast<double>(pid / size);
        x += 0.5; y += 0.5;
        x /= size; y /= size;
        double out6[6];
        patch_evaluate(coeffs, x, y, out6);
        (*w)[pid] = out6[0];
        if (wx) { (*wx)[pid] = out6[1] / size; (*wy)[pid] = out6[2] / size; }
    }
}

/* Surface::get_depth_map (lib/surface.cc:155-168) */
void
depth_map (Port const& P, std::vector<float>* out)
{
    out->assign(static_cast<std::size_t>(P.w) * P.h, 0.0f);
    std::vector<double> w;
    for (int patch = 0; patch < P.npx * P.npy; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        patch_pixels(P, patch, &w, nullptr, nullptr);
        int const px0 = P.sx + (patch % P.npx) * P.ps;
        int const py0 = P.sy + (patch / P.npx) * P.ps;
        for (int pid = 0; pid < P.ps * P.ps; ++pid)
            (*out)[static_cast<std::size_t>(py0 + pid / P.ps) * P.w
                + px0 + pid % P.ps] = static_cast<float>(w[pid]);
    }
}

/* Surface::remove_nodes_without_patch (lib/surface.cc:762-867): a node goes
 * when none of the patches around it that lie inside the grid is left. */
void
remove_nodes_without_patch (Port& P)
{
    int const ns = P.npx + 1;
    for (int node = 0; node < P.n_nodes(); ++node)
    {
        if (!P.node_valid[node])
            continue;
        int const ix = node % ns, iy = node / ns;
        bool any = false;
        for (int dy = -1; dy <= 0; ++dy)
            for (int dx = -1; dx <= 0; ++dx)
            {
                int const px = ix + dx, py = iy + dy;
                if (px < 0 || py < 0 || px >= P.npx || py >= P.npy)
                    continue;
                any = any || P.patch_valid[py * P.npx + px];
            }
        if (!any)
            P.node_valid[node] = 0;
    }
}

/* DepthOptimizer::ncc_for_patch, lib/depth_optimizer.cc:795-912: NCC of the
 * patch (plus a two-pixel rim, where it fits) between the main colour image
 * and one neighbour's, -1 when a pixel warps outside the neighbour. */
double
ncc_for_patch (Port const& P, int patch, int sub, std::vector<double>* px_io,
    std::vector<double>* py_io, std::vector<double>* pd_io)
{
    /* The reference works on the optimizer's member vectors `pixels` and
     * `depths`: the rim pixels appended here stay in them after the return
     * (px_io / py_io / pd_io), and the caller's next neighbour tests them. */
    std::vector<double>& px = *px_io;
    std::vector<double>& py = *py_io;
    std::vector<double>& pd = *pd_io;
    int const ps = P.ps;
    int const px0 = P.sx + (patch % P.npx) * ps;
    int const py0 = P.sy + (patch / P.npx) * ps;
    int ids[4];
    P.node_ids(patch, ids);
    /* fill_values_at_nodes: corners and their depths */
    double const cx[4] = { double(px0), double(px0 + ps), double(px0),
        double(px0 + ps) };
    double const cy[4] = { double(py0), double(py0), double(py0 + ps),
        double(py0 + ps) };
    double cd[4];
    for (int i = 0; i < 4; ++i)
        cd[i] = P.nodes[ids[i] * 4];
    double const min0 = cx[0], min1 = cy[0], max0 = cx[3], max1 = cy[3];

    std::vector<double> w;
    patch_pixels(P, patch, &w, nullptr, nullptr);
    px.clear(); py.clear(); pd.clear();
    for (int i = 0; i < ps * ps; ++i)
    {
        px.push_back(px0 + i % ps);
        py.push_back(py0 + i / ps);
        pd.push_back(w[i]);
    }
    /* boundary, :813-859 (the loop runs over the list while it grows) */
    if (min0 > 1 && max0 < P.w - 2 && min1 > 1 && max1 < P.h - 2)
    {
        px.push_back(cx[0] - 1); py.push_back(cy[0] - 1); pd.push_back(cd[0]);
        px.push_back(cx[1] + 1); py.push_back(cy[1] - 1); pd.push_back(cd[1]);
        px.push_back(cx[2] - 1); py.push_back(cy[2] + 1); pd.push_back(cd[2]);
        px.push_back(cx[3] + 1); py.push_back(cy[3] + 1); pd.push_back(cd[3]);
    }
    for (std::size_t i = 0; i < px.size(); ++i)
    {
        if (min1 > 2 && py[i] == min1)
        {
            px.push_back(px[i]); py.push_back(py[i] - 2); pd.push_back(pd[i]);
            px.push_back(px[i]); py.push_back(py[i] - 1); pd.push_back(pd[i]);
        }
        if (max1 < P.h - 3 && py[i] == max1)
        {
            px.push_back(px[i]); py.push_back(py[i] + 2); pd.push_back(pd[i]);
            px.push_back(px[i]); py.push_back(py[i] + 1); pd.push_back(pd[i]);
        }
        if (min0 > 2 && px[i] == min0)
        {
            px.push_back(px[i] - 2); py.push_back(py[i]); pd.push_back(pd[i]);
            px.push_back(px[i] - 1); py.push_back(py[i]); pd.push_back(pd[i]);
        }
        if (max0 < P.w - 3 && px[i] == max0)
        {
            px.push_back(px[i] + 2); py.push_back(py[i]); pd.push_back(pd[i]);
            px.push_back(px[i] + 1); py.push_back(py[i]); pd.push_back(pd[i]);
        }
    }

    std::size_t const n = px.size();
    std::vector<double> v0(n * 3), v1(n * 3);
    double means0[3] = {0, 0, 0}, means1[3] = {0, 0, 0}, counter[3] = {0, 0, 0};
    int const sw = P.sub_w[sub], sh = P.sub_h[sub];
    for (std::size_t i = 0; i < n; ++i)
    {
        Corr c;
        c.update(&P.Mi[9 * sub], &P.ti[3 * sub], px[i] + 0.5, py[i] + 0.5,
            pd[i], 0.0, 0.0);
        double proj[2];
        c.fill(proj);
        proj[0] -= 0.5; proj[1] -= 0.5;
        if (proj[0] < 1 || proj[0] > sw - 2 || proj[1] < 1 || proj[1] > sh - 2)
            return -1;
        for (int ch = 0; ch < 3; ++ch)
        {
            double const cm = P.main_image[(static_cast<std::size_t>(
                static_cast<int>(py[i])) * P.w + static_cast<int>(px[i])) * 3 + ch];
            double const cs = linear_at_f(P.sub_image[sub].data(), sw, sh, 3,
                static_cast<float>(proj[0]), static_cast<float>(proj[1]), ch);
            counter[ch] += 1.0;
            means0[ch] += (cm - means0[ch]) / counter[ch];
            means1[ch] += (cs - means1[ch]) / counter[ch];
            v0[i * 3 + ch] = cm;
            v1[i * 3 + ch] = cs;
        }
    }
    for (std::size_t i = 0; i < n; ++i)
        for (int ch = 0; ch < 3; ++ch)
        {
            v0[i * 3 + ch] -= means0[ch];
            v1[i * 3 + ch] -= means1[ch];
        }
    double const norm0 = std::sqrt(dot(v0, v0)), norm1 = std::sqrt(dot(v1, v1));
    if (norm0 + norm1 < 0.001 * n)
        return 1;
    return dot(v0, v1) / (norm0 * norm1);
}

/* DepthOptimizer::create_subview_surfaces (lib/depth_optimizer.cc:433-604).
 * sgm_depth != NULL: the use_sgm mode; NULL: use_sgm = false, with the NCC
 * occlusion filter of :579-581 (needs the colour images). Returns the
 * patches deleted. */
int
create_subview_surfaces (Port& P, float const* sgm_depth)
{
    int const np = P.npx * P.npy;
    std::vector<std::vector<uint8_t> > subsurfaces(np);

    /* depth caches, :441-450 */
    std::vector<std::vector<float> > cache(P.n_sub);
    for (int s = 0; s < P.n_sub; ++s)
        cache[s].assign(static_cast<std::size_t>(P.sub_w[s] + 1)
            * (P.sub_h[s] + 1), 10000.0f);

    /* pixel list, :452-469 (x outer, y inner) */
    std::vector<float> depth;
    depth_map(P, &depth);
    std::vector<double> px, py, pd;
    for (int x = 0; x < P.w; ++x)
        for (int y = 0; y < P.h; ++y)
        {
            std::size_t const i = static_cast<std::size_t>(y) * P.w + x;
            if (depth[i] != 0)
            { px.push_back(x); py.push_back(y); pd.push_back(depth[i]); }
            if (sgm_depth != nullptr && sgm_depth[i] != 0)
            { px.push_back(x); py.push_back(y); pd.push_back(sgm_depth[i]); }
        }

    /* first pass: minimal depth, :471-500 */
    Corr C;
    for (int s = 0; s < P.n_sub; ++s)
    {
        double const sw = P.sub_w[s], sh = P.sub_h[s];
        int const cw = P.sub_w[s] + 1;
        for (std::size_t i = 0; i < px.size(); ++i)
        {
            C.update(&P.Mi[9 * s], &P.ti[3 * s], px[i] + 0.5, py[i] + 0.5,
                pd[i], 0.0, 0.0);
            double proj[2];
            C.fill(proj);
            proj[0] -= 0.5; proj[1] -= 0.5;
            double const cutoffset = 3.0;
            if (proj[0] < cutoffset || proj[0] >= sw - cutoffset
                || proj[1] < cutoffset || proj[1] >= sh - cutoffset)
                continue;
            int const cx = static_cast<int>(proj[0]);
            int const cy = static_cast<int>(proj[1]);
            for (int x = -1; x < 2; ++x)
                for (int y = -1; y < 2; ++y)
                    if (C.d < cache[s][(cy + y) * cw + cx + x])
                        cache[s][(cy + y) * cw + cx + x]
                            = static_cast<float>(C.d);
        }
    }

    /* second pass, :502-583. `lx, ly, ld` are the optimizer's member vectors
     * pixels / depths: filled once per patch (:508), refilled after a
     * neighbour passes the first test (:551), and -- in the use_sgm = false
     * mode -- left extended by ncc_for_patch for the next neighbour. */
    std::vector<double> w, wx, wy, lx, ly, ld;
    for (int patch = 0; patch < np; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        patch_pixels(P, patch, &w, &wx, &wy);
        int const px0 = P.sx + (patch % P.npx) * P.ps;
        int const py0 = P.sy + (patch / P.npx) * P.ps;
        lx.clear(); ly.clear(); ld.clear();
        for (int i = 0; i < P.ps * P.ps; ++i)
        {
            lx.push_back(px0 + i % P.ps);
            ly.push_back(py0 + i / P.ps);
            ld.push_back(w[i]);
        }
        for (int s = 0; s < P.n_sub; ++s)
        {
            double const sw = P.sub_w[s], sh = P.sub_h[s];
            int const cw = P.sub_w[s] + 1;
            bool success = true;
            for (std::size_t i = 0; i < lx.size() && success; ++i)
            {
                Corr c;
                c.update(&P.Mi[9 * s], &P.ti[3 * s], lx[i] + 0.5, ly[i] + 0.5,
                    ld[i], 0.0, 0.0);
                double proj[2];
                c.fill(proj);
                proj[0] -= 0.5; proj[1] -= 0.5;
                double const cutoffset = 0.03 * std::max(sw, sh);
                if (proj[0] < cutoffset || proj[0] >= sw - cutoffset
                    || proj[1] < cutoffset || proj[1] >= sh - cutoffset)
                {
                    success = false;
                    break;
                }
                int const cx = static_cast<int>(proj[0]);
                int const cy = static_cast<int>(proj[1]);
                for (int x = -1; x < 2; ++x)
                    for (int y = -1; y < 2; ++y)
                        if (c.d * 0.95 > cache[s][(cy + y) * cw + cx + x])
                            success = false;
            }
            if (!success)
                continue;

            /* :551: the member vectors are the patch's own pixels again */
            lx.resize(P.ps * P.ps); ly.resize(P.ps * P.ps);
            ld.resize(P.ps * P.ps);
            for (int i = 0; i < P.ps * P.ps; ++i)
            {
                lx[i] = px0 + i % P.ps; ly[i] = py0 + i / P.ps; ld[i] = w[i];
            }
            double max = 0.0;
            for (int i = 0; i < P.ps * P.ps; ++i)
            {
                Corr c;
                c.update(&P.Mi[9 * s], &P.ti[3 * s], px0 + i % P.ps + 0.5,
                    py0 + i / P.ps + 0.5, w[i], wx[i], wy[i]);
                double jac[4];
                c.fill_jacobian(jac);
                double S[2];
                S[0] = (std::sqrt((jac[0] - jac[3]) * (jac[0] - jac[3])
                    + (jac[1] + jac[2]) * (jac[1] + jac[2]))
                    + std::sqrt((jac[0] + jac[3]) * (jac[0] + jac[3])
                    + (jac[1] - jac[2]) * (jac[1] - jac[2]))) / 2.0;
                S[1] = std::fabs(S[0] - std::sqrt((jac[0] - jac[3])
                    * (jac[0] - jac[3]) + (jac[1] + jac[2]) * (jac[1] + jac[2])));
                double const hi = std::max(S[0], S[1]), lo = std::min(S[0], S[1]);
                double const sigma0 = hi * hi, sigma1 = lo * lo;
                max = std::max(max, sigma0 / sigma1);
            }
            if (max > 8.0)
                continue;
            /* filter possible occlusions from unreconstructed geometry */
            if (sgm_depth == nullptr
                && ncc_for_patch(P, patch, s, &lx, &ly, &ld) < 0)
                continue;
            subsurfaces[patch].push_back(static_cast<uint8_t>(s));
        }
    }

    /* :585-600 */
    int removed = 0;
    for (int patch = 0; patch < np; ++patch)
        if (P.patch_valid[patch] && subsurfaces[patch].empty())
        {
            P.patch_valid[patch] = 0;
            removed += 1;
        }
    if (removed > 0)
        remove_nodes_without_patch(P);
    P.vis_off.assign(np + 1, 0);
    P.vis_ids.clear();
    for (int patch = 0; patch < np; ++patch)
    {
        P.vis_off[patch] = static_cast<uint32_t>(P.vis_ids.size());
        P.vis_ids.insert(P.vis_ids.end(), subsurfaces[patch].begin(),
            subsurfaces[patch].end());
    }
    P.vis_off[np] = static_cast<uint32_t>(P.vis_ids.size());
    return removed;
}

/* DepthOptimizer::mse_for_patch, lib/depth_optimizer.cc:747-793 */
double
mse_for_patch (Port const& P, int patch)
{
    std::vector<double> w, wx, wy;
    patch_pixels(P, patch, &w, &wx, &wy);
    int const px0 = P.sx + (patch % P.npx) * P.ps;
    int const py0 = P.sy + (patch / P.npx) * P.ps;
    double error = 0.0, counter = 0.0;
    for (int i = 0; i < P.ps * P.ps; ++i)
    {
        int const x = px0 + i % P.ps, y = py0 + i / P.ps;
        double const gm0 = P.main_grad[(static_cast<std::size_t>(y) * P.w + x) * 2];
        double const gm1 = P.main_grad[(static_cast<std::size_t>(y) * P.w + x) * 2 + 1];
        for (uint32_t k = P.vis_off[patch]; k < P.vis_off[patch + 1]; ++k)
        {
            int const s = P.vis_ids[k];
            Corr c;
            c.update(&P.Mi[9 * s], &P.ti[3 * s], x + 0.5, y + 0.5, w[i],
                wx[i], wy[i]);
            double proj[2], jac[4];
            c.fill(proj);
            c.fill_jacobian(jac);
            proj[0] -= 0.5; proj[1] -= 0.5;
            double const gs0 = linear_at_f(P.sub_grad[s].data(), P.sub_w[s],
                P.sub_h[s], 2, static_cast<float>(proj[0]),
                static_cast<float>(proj[1]), 0);
            double const gs1 = linear_at_f(P.sub_grad[s].data(), P.sub_w[s],
                P.sub_h[s], 2, static_cast<float>(proj[0]),
                static_cast<float>(proj[1]), 1);
            double const d0 = gm0 - (0.0 + jac[0] * gs0 + jac[1] * gs1);
            double const d1 = gm1 - (0.0 + jac[2] * gs0 + jac[3] * gs1);
            error += std::sqrt(0.0 + d0 * d0 + d1 * d1);
            counter += 1.0;
        }
    }
    if (counter == 0.0)
        return 1.0;
    return error / counter;
}

/* DepthOptimizer::cut_boundaries, lib/depth_optimizer.cc:360-431.
 * inv: the main camera's inverse calibration (3x3, fp32). */
int
cut_boundaries (Port& P, float const* inv)
{
    int deleted = 0;
    int const np = P.npx * P.npy, ns = P.npx + 1;
    for (int patch = 0; patch < np; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        int ids[4];
        P.node_ids(patch, ids);
        double depths[4];
        for (int i = 0; i < 4; ++i)
            depths[i] = P.nodes[ids[i] * 4];
        /* std::multimap<double, size_t>: first of the minima, last of the maxima */
        int lo = 0, hi = 0;
        for (int i = 1; i < 4; ++i)
        {
            if (depths[i] < depths[lo]) lo = i;
            if (!(depths[i] < depths[hi])) hi = i;
        }
        double dd_factor = 5.0;
        if (lo + hi == 3)
            dd_factor *= 1.41421356237309504880168872420969808;
        float const fx = static_cast<float>(static_cast<double>(
            P.sx + (patch % P.npx) * P.ps)) + 0.5f;
        float const fy = static_cast<float>(static_cast<double>(
            P.sy + (patch / P.npx) * P.ps)) + 0.5f;
        float v[3];
        for (int r = 0; r < 3; ++r)
        {
            float sum = 0.0f;
            sum += inv[r * 3 + 0] * fx;
            sum += inv[r * 3 + 1] * fy;
            sum += inv[r * 3 + 2] * 1.0f;
            v[r] = sum;
        }
        float sq = 0.0f;
        for (int r = 0; r < 3; ++r)
            sq += v[r] * v[r];
        float const norm = std::sqrt(sq);
        double const threshold = dd_factor * depths[lo] * inv[0] * P.ps / norm;
        double const dist = depths[hi] - depths[lo];
        if (dist > threshold)
        {
            P.patch_valid[patch] = 0;
            deleted += 1;
        }
    }
    for (int patch = 0; patch < np; ++patch)
    {
        if (!P.patch_valid[patch])
            continue;
        int ids[4];
        P.node_ids(patch, ids);
        double const error = mse_for_patch(P, patch);
        for (int n = 0; n < 4; ++n)
        {
            /* Surface::fill_node_neighbors: the eight grid neighbours,
             * missing when outside the grid or not valid */
            int const ix = ids[n] % ns, iy = ids[n] / ns;
            int num_invalid = 0;
            for (int dy = -1; dy < 2; ++dy)
                for (int dx = -1; dx < 2; ++dx)
                {
                    if (dx == 0 && dy == 0)
                        continue;
                    int const qx = ix + dx, qy = iy + dy;
                    if (qx < 0 || qy < 0 || qx > P.npx || qy > P.npy
                        || !P.node_valid[qy * ns + qx])
                        num_invalid += 1;
                }
            if (num_invalid > 1 && error > 0.05)
            {
                P.patch_valid[patch] = 0;
                deleted += 1;
                break;
            }
        }
    }
    remove_nodes_without_patch(P);
    return deleted;
}

/* DepthOptimizer::depthmap_bilateral_filter, lib/depth_optimizer.cc:957-1004 */
void
bilateral_filter (int w, int h, int channels, float const* ci, int dm_w,
    int dm_h, float const* dm, float sigma, int kernel_size, float* out)
{
    float const scale_x = static_cast<float>(dm_w) / static_cast<float>(w);
    float const scale_y = static_cast<float>(dm_h) / static_cast<float>(h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            float acc_v = 0.0f, acc_w = 0.0f;
            for (int ky = -kernel_size; ky <= kernel_size; ++ky)
                for (int kx = -kernel_size; kx <= kernel_size; ++kx)
                {
                    int const ci_x = std::min(std::max(x + kx, 0), w - 1);
                    int const ci_y = std::min(std::max(y + ky, 0), h - 1);
                    float fx = scale_x * ci_x, fy = scale_y * ci_y;
                    float const mx = static_cast<float>(dm_w) - 1.f;
                    float const my = static_cast<float>(dm_h) - 1.f;
                    fx = fx < 0.f ? 0.f : (fx > mx ? mx : fx);
                    fy = fy < 0.f ? 0.f : (fy > my ? my : fy);
                    int const dm_x = static_cast<int>(fx);
                    int const dm_y = static_cast<int>(fy);
                    float const d = dm[dm_y * dm_w + dm_x];
                    if (d == 0.0f)
                        continue;
                    float weight = 1.0f;
                    float const gx = static_cast<float>(kx);
                    float const gy = static_cast<float>(ky);
                    weight *= std::exp(-(gx * gx) / (2.0f * sigma * sigma)
                        - (gy * gy) / (2.0f * sigma * sigma));
                    for (int c = 0; c < channels; ++c)
                    {
                        float const diff = ci[(ci_y * w + ci_x) * channels + c]
                            - ci[(y * w + x) * channels + c];
                        weight *= std::exp(-((diff * diff)
                            / (2.0f * 0.1f * 0.1f)));
                    }
                    acc_v += d * weight;
                    acc_w += weight;
                }
            out[y * w + x] = (acc_w > 0) ? acc_v / acc_w : 0.0f;
        }
}

extern "C" {

/* ---- unit-level entry points (same signatures as oracle/ref_driver.cc) -- */

void
port_bicubic_eval (double const* nodes16, double x, double y, double* out)
{
    double c[4][4];
    patch_coefficients(nodes16, c);
    patch_evaluate(c, x, y, out);
}

void
port_bicubic_node_derivatives (double x, double y, double patchsize, double* out)
{
    node_derivatives(x, y, patchsize, out);
}

void
port_correspondence (double const* M9, double const* t3, double u, double v,
    double w, double wx, double wy, double const* grad2, double const* dn96,
    double* proj, double* jac, double* c_dn, double* jac_dn, double* depth)
{
    Corr C;
    C.update(M9, t3, u, v, w, wx, wy);
    C.fill(proj);
    C.fill_jacobian(jac);
    C.fill_derivative(dn96, c_dn);
    C.fill_jacobian_derivative_grad(grad2, dn96, jac_dn);
    *depth = C.d;
}

void
port_surface_derivatives (double const* dn96, double x, double y, double f,
    double w, double dx, double dy, double dxy, double dxx, double dyy,
    double* normal, double* div, double* div_deriv, double* normal_deriv)
{
    fill_normal(x, y, 1.0 / f, w, dx, dy, normal);
    normal_divergence(x, y, f, w, dx, dy, dxy, dxx, dyy, div);
    normal_divergence_deriv(dn96, x, y, f, w, dx, dy, dxy, dxx, dyy, div_deriv);
    normal_derivative(dn96, x, y, f, w, dx, dy, normal_deriv);
}

void
port_sh_4band (double const* normal, double* sh16, double* deriv48)
{
    sh_evaluate_4_band(normal, sh16);
    sh_derivative_4_band(normal, deriv48);
}

void
port_ldl_inverse (double* A, int n)
{
    ldl_inverse(A, n);
}

/* ---- path-level entry points ------------------------------------------- */

void*
port_create (int w, int h, double flen, double inv_flen,
    float const* main_grad, float const* shading, float const* shading_grad,
    int n_sub, int const* sub_w, int const* sub_h,
    float const* const* sub_grad, float const* const* sub_hess,
    double const* Mi, double const* ti)
{
    Port* P = new Port();
    P->w = w; P->h = h; P->flen = flen; P->inv_flen = inv_flen;
    P->n_sub = n_sub;
    std::size_t const np = static_cast<std::size_t>(w) * h;
    P->main_grad.assign(main_grad, main_grad + np * 2);
    if (shading)
    {
        P->shading.assign(shading, shading + np);
        P->shading_grad.assign(shading_grad, shading_grad + np * 2);
    }
    for (int k = 0; k < n_sub; ++k)
    {
        std::size_t const n = static_cast<std::size_t>(sub_w[k]) * sub_h[k];
        P->sub_w.push_back(sub_w[k]); P->sub_h.push_back(sub_h[k]);
        P->sub_grad.emplace_back(sub_grad[k], sub_grad[k] + n * 2);
        P->sub_hess.emplace_back(sub_hess[k], sub_hess[k] + n * 3);
    }
    P->Mi.assign(Mi, Mi + 9 * n_sub);
    P->ti.assign(ti, ti + 3 * n_sub);
    P->npx = P->npy = 0;
    return P;
}

void
port_destroy (void* p)
{
    delete static_cast<Port*>(p);
}

void
port_set_surface (void* p, int scale, int npx, int npy, int sx, int sy,
    double const* nodes, uint8_t const* node_valid, uint8_t const* patch_valid,
    uint32_t const* vis_off, uint8_t const* vis_ids)
{
    Port* P = static_cast<Port*>(p);
    P->scale = scale; P->ps = 1 << scale; P->npx = npx; P->npy = npy;
    P->sx = sx; P->sy = sy;
    int const nn = (npx + 1) * (npy + 1), np = npx * npy;
    P->nodes.assign(nodes, nodes + nn * 4);
    P->node_valid.assign(node_valid, node_valid + nn);
    P->patch_valid.assign(patch_valid, patch_valid + np);
    P->vis_off.assign(vis_off, vis_off + np + 1);
    P->vis_ids.assign(vis_ids, vis_ids + vis_off[np]);
}

int64_t
port_gn_construct (void* p, uint8_t const* active, double const* light16,
    double regularization, double light_surf_regularization)
{
    Port* P = static_cast<Port*>(p);
    gn_construct(*P, active, light16, regularization, light_surf_regularization);
    return static_cast<int64_t>(P->Hinner.size());
}

void
port_get_system_sizes (void* p, uint64_t* sizes)
{
    Port* P = static_cast<Port*>(p);
    sizes[0] = P->g.size(); sizes[1] = P->Hinner.size();
    sizes[2] = P->Pinner.size();
}

void
port_get_system (void* p, double* g, double* Hvals, uint64_t* Houter,
    uint64_t* Hinner, double* Pvals, uint64_t* Pouter, uint64_t* Pinner)
{
    Port* P = static_cast<Port*>(p);
    if (g) std::copy(P->g.begin(), P->g.end(), g);
    if (Hvals) std::copy(P->Hvals.begin(), P->Hvals.end(), Hvals);
    if (Houter) std::copy(P->Houter.begin(), P->Houter.end(), Houter);
    if (Hinner) std::copy(P->Hinner.begin(), P->Hinner.end(), Hinner);
    if (Pvals) std::copy(P->Pvals.begin(), P->Pvals.end(), Pvals);
    if (Pouter) std::copy(P->Pouter.begin(), P->Pouter.end(), Pouter);
    if (Pinner) std::copy(P->Pinner.begin(), P->Pinner.end(), Pinner);
}

int
port_cg_solve (void* p, int max_iter, double err_tol, double q_tol,
    double* x_out, int* iters, int* info)
{
    Port* P = static_cast<Port*>(p);
    std::vector<double> x;
    cg_solve(*P, max_iter, err_tol, q_tol, &x, iters, info);
    std::copy(x.begin(), x.end(), x_out);
    return 0;
}

int
port_update_nodes (void* p, double const* delta, double thresh, int full_opt,
    uint8_t* active_inout, uint64_t* n_active, double* mean_shift)
{
    Port* P = static_cast<Port*>(p);
    std::vector<uint8_t> act(active_inout, active_inout + P->n_nodes());
    std::vector<double> d(delta, delta + P->n_nodes() * 4);
    update_nodes(*P, d, thresh, full_opt != 0, &act, n_active, mean_shift);
    std::copy(act.begin(), act.end(), active_inout);
    return 0;
}

void
port_get_nodes (void* p, double* nodes)
{
    Port* P = static_cast<Port*>(p);
    std::copy(P->nodes.begin(), P->nodes.end(), nodes);
}

/* lib/depth_optimizer.cc:204-304. stats[8] as ref_newton_loop; times 0. */
int
port_newton_loop (void* p, double const* light16, double regularization,
    double light_surf_regularization, int max_steps, double* stats)
{
    Port* P = static_cast<Port*>(p);
    int const nn = P->n_nodes();
    std::vector<uint8_t> active(P->node_valid.begin(), P->node_valid.end());
    uint64_t num_initial = 0;
    for (uint8_t a : active) num_initial += (a != 0);
    uint64_t num_active = num_initial;
    int const sampling = sampling_for_scale(P->scale);
    double const spp = double(P->ps * P->ps) / (sampling * sampling);
    double steps = 0, cg_iters = 0, pix = 0, nan = 0;
    for (; steps < max_steps && num_active > num_initial / 20;)
    {
        steps += 1;
        for (int patch = 0; patch < P->npx * P->npy; ++patch)
        {
            if (!P->patch_valid[patch]) continue;
            int ids[4];
            P->node_ids(patch, ids);
            if (active[ids[0]] || active[ids[1]] || active[ids[2]] || active[ids[3]])
                pix += spp;
        }
        gn_construct(*P, active.data(), light16, regularization,
            light_surf_regularization);
        std::vector<double> x;
        int it = 0, info = 0;
        cg_solve(*P, 200, -1.0, 1e-3, &x, &it, &info);
        cg_iters += it;
        if (std::isnan(x[0])) { nan = 1; break; }
        update_nodes(*P, x, 0.15, false, &active, &num_active, nullptr);
    }
    (void)nn;
    stats[0] = steps; stats[1] = cg_iters; stats[2] = pix;
    stats[3] = stats[4] = stats[5] = 0.0;
    stats[6] = (double)num_active; stats[7] = nan;
    return 0;
}

/* SGMStereo::run_sgm, lib/sgm_stereo.cc:98-124 with create_cost_volume
 * (:192-244), aggregate_sgm_costs (:429-667, SSE semantics) and
 * depth_from_sgm_volume (:274-306). cost_out / sgm_out: w*h*D uint16. */

/* Surface::create from an initial depth map (w*h); info[6] as ref_surface_info. */
void
port_surface_create (void* p, int scale, float const* init_depth)
{
    surface_create(*static_cast<Port*>(p), scale, init_depth);
}

void
port_surface_subdivide (void* p)
{
    subdivide_patches(*static_cast<Port*>(p));
}

void
port_surface_fill_from_depth (void* p)
{
    fill_patches_from_depth(*static_cast<Port*>(p));
}

int
port_surface_expand (void* p)
{
    return expand(*static_cast<Port*>(p));
}

void
port_surface_remove_isolated (void* p)
{
    remove_isolated_patches(*static_cast<Port*>(p));
}

void
port_surface_info (void* p, int* info)
{
    Port* P = static_cast<Port*>(p);
    info[0] = P->scale; info[1] = P->npx; info[2] = P->npy;
    info[3] = P->sx; info[4] = P->sy; info[5] = P->ps;
}

/* The views' unblurred float images, 3 channels each (StereoView::get_image). */
void
port_set_images (void* p, float const* main_image,
    float const* const* sub_images)
{
    Port* P = static_cast<Port*>(p);
    P->main_image.assign(main_image, main_image
        + static_cast<std::size_t>(P->w) * P->h * 3);
    P->sub_image.clear();
    for (int k = 0; k < P->n_sub; ++k)
        P->sub_image.emplace_back(sub_images[k], sub_images[k]
            + static_cast<std::size_t>(P->sub_w[k]) * P->sub_h[k] * 3);
}

/* create_subview_surfaces on the port's surface; sgm_depth: w*h, or NULL for
 * the use_sgm = false mode (port_set_images first). */
int
port_visibility (void* p, float const* sgm_depth)
{
    return create_subview_surfaces(*static_cast<Port*>(p), sgm_depth);
}

int
port_cut_boundaries (void* p, float const* inv_calib9)
{
    return cut_boundaries(*static_cast<Port*>(p), inv_calib9);
}

void
port_get_surface_state (void* p, uint8_t* node_valid, uint8_t* patch_valid,
    uint32_t* vis_off, uint8_t* vis_ids)
{
    Port* P = static_cast<Port*>(p);
    std::copy(P->node_valid.begin(), P->node_valid.end(), node_valid);
    std::copy(P->patch_valid.begin(), P->patch_valid.end(), patch_valid);
    std::copy(P->vis_off.begin(), P->vis_off.end(), vis_off);
    std::copy(P->vis_ids.begin(), P->vis_ids.end(), vis_ids);
}

void
port_depth_map (void* p, float* out)
{
    std::vector<float> d;
    depth_map(*static_cast<Port*>(p), &d);
    std::copy(d.begin(), d.end(), out);
}

void
port_bilateral_filter (int w, int h, int channels, float const* guide,
    int dm_w, int dm_h, float const* depth, float sigma, int kernel_size,
    float* out)
{
    bilateral_filter(w, h, channels, guide, dm_w, dm_h, depth, sigma,
        kernel_size, out);
}

int
port_sgm (int w, int h, uint8_t const* main_img, int nw, int nh,
    uint8_t const* neigh, float const* M, float const* t, float min_depth,
    float max_depth, int D, int P1, int P2, float* depth_out,
    uint16_t* cost_out, uint16_t* sgm_out)
{
    std::size_t const npix = static_cast<std::size_t>(w) * h;
    std::vector<float> depths(D);
    {
        float inv_depth = 1.0f / max_depth;
        float const inc = (1.0f / min_depth - inv_depth) / (D - 1);
        for (int i = 0; i < D; ++i) { depths[i] = 1.0f / inv_depth; inv_depth += inc; }
    }
    /* warped_neighbors_for_depth, :150-190 */
    std::vector<uint8_t> warped(npix * D, 0);
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y)
        {
            float const px = 0.5f + x, py = 0.5f + y;
            float tp[3];
            for (int r = 0; r < 3; ++r)
            {
                float s = 0.0f;
                s += M[3 * r] * px; s += M[3 * r + 1] * py; s += M[3 * r + 2] * 1.f;
                tp[r] = s;
            }
            for (int d = 0; d < D; ++d)
            {
                float q0 = tp[0] * depths[d] + t[0];
                float q1 = tp[1] * depths[d] + t[1];
                float const q2 = tp[2] * depths[d] + t[2];
                if (q2 < 0) continue;
                q0 /= q2; q1 /= q2; q0 -= 0.5f; q1 -= 0.5f;
                if (q0 < 0 || q1 < 0 || q0 > nw - 1 || q1 > nh - 1) continue;
                warped[(static_cast<std::size_t>(y) * w + x) * D + d] =
                    linear_at_u8(neigh, nw, nh, q0, q1);
            }
        }
    std::vector<uint64_t> main_census(npix), warped_census(npix * D);
    census_filter(main_img, w, h, 1, main_census.data());
    census_filter(warped.data(), w, h, D, warped_census.data());
    std::vector<uint16_t> C(npix * D, 255);
    for (std::size_t p = 0; p < npix; ++p)
        for (int i = 0; i < D; ++i)
        {
            if (warped[p * D + i] == 0) continue;
            C[p * D + i] = static_cast<uint8_t>(__builtin_popcountll(
                main_census[p] ^ warped_census[p * D + i]));
        }

    /* aggregation */
    std::vector<uint16_t> S(npix * D, 0);
    std::vector<uint16_t> L0(npix * D), L1(npix * D), L2(npix * D);
    auto copy_add = [&](std::vector<uint16_t>& L, std::size_t base) {
        for (int i = 0; i < D; ++i)
        {
            L[base + i] = C[base + i];
            S[base + i] = static_cast<uint16_t>(S[base + i] + C[base + i]);
        }
    };
    auto path = [&](std::vector<uint16_t>& L, std::size_t base, std::size_t pbase) {
        /* fill_path_cost_sse, :361-406, O(D^2) as in the reference */
        uint16_t min_prev = 0xffff;
        for (int i = 0; i < D; ++i) min_prev = std::min(min_prev, L[pbase + i]);
        std::vector<uint16_t> upd(D);
        for (int idx = 0; idx < D; ++idx)
        {
            uint16_t best = 0xffff;
            for (int j = 0; j < D; ++j)
            {
                uint16_t c = static_cast<uint16_t>(L[pbase + j] + P2);
                if (j == idx) c = L[pbase + j];
                else if (j == idx - 1 || j == idx + 1)
                    c = static_cast<uint16_t>(L[pbase + j] + P1);
                best = std::min(best, c);
            }
            upd[idx] = best;
        }
        for (int i = 0; i < D; ++i)
        {
            uint16_t v = static_cast<uint16_t>(C[base + i] + upd[i]);
            v = static_cast<uint16_t>(v - min_prev);
            L[base + i] = v;
            S[base + i] = static_cast<uint16_t>(S[base + i] + v);
        }
    };
    auto B = [&](int x, int y) { return (static_cast<std::size_t>(y) * w + x) * D; };
    /* left-to-right, right-to-left, :456-503 */
    std::fill(L0.begin(), L0.end(), 0);
    for (int y = 0; y < h; ++y) copy_add(L0, B(0, y));
    for (int x = 1; x < w; ++x)
        for (int y = 0; y < h; ++y) path(L0, B(x, y), B(x - 1, y));
    std::fill(L0.begin(), L0.end(), 0);
    for (int y = 0; y < h; ++y) copy_add(L0, B(w - 1, y));
    for (int x = w - 2; x >= 0; --x)
        for (int y = 0; y < h; ++y) path(L0, B(x, y), B(x + 1, y));
    /* top-to-bottom with both diagonals, :505-548 */
    for (int pass = 0; pass < 2; ++pass)
    {
        int const y0 = pass == 0 ? 0 : h - 1, dy = pass == 0 ? 1 : -1;
        std::fill(L0.begin(), L0.end(), 0);
        std::fill(L1.begin(), L1.end(), 0);
        std::fill(L2.begin(), L2.end(), 0);
        for (int x = 0; x < w; ++x)
        {
            copy_add(L0, B(x, y0)); copy_add(L1, B(x, y0)); copy_add(L2, B(x, y0));
        }
        for (int y = 0; y < h; ++y) copy_add(L1, B(0, y));
        for (int y = 0; y < h; ++y) copy_add(L2, B(w - 1, y));
        for (int y = y0 + dy; y >= 0 && y < h; y += dy)
            for (int x = 0; x < w; ++x)
            {
                if (x > 0) path(L1, B(x, y), B(x - 1, y - dy));
                if (x < w - 1) path(L2, B(x, y), B(x + 1, y - dy));
                path(L0, B(x, y), B(x, y - dy));
            }
    }
    /* depth_from_sgm_volume */
    for (std::size_t p = 0; p < npix; ++p)
    {
        uint16_t min_error = 0xffff;
        int min_index = 0;
        for (int i = 0; i < D; ++i)
            if (S[p * D + i] < min_error) { min_error = S[p * D + i]; min_index = i; }
        depth_out[p] = (min_index < 2 || main_img[p] < 25) ? 0.0f : depths[min_index];
    }
    if (cost_out) std::copy(C.begin(), C.end(), cost_out);
    if (sgm_out) std::copy(S.begin(), S.end(), sgm_out);
    return 0;
}

} /* extern "C" */
