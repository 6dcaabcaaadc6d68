/*
 * cg_v1_probe.cu -- EXPERIMENT ONLY (not part of the product): the round-1
 * single-view PCG kernel, kept callable through SMVSB_CG_VARIANT=v1 to bisect
 * a performance difference against the batched kernel. Removed after the
 * measurement.
 */
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace smvsb {
namespace v1 {

constexpr int CG_THREADS = 256;
constexpr int CG_MAX_BLOCKS = 1024;
constexpr int CG_UF = 4;

struct CgArgs
{
    int n_nodes, npx, npy;
    int max_iter;
    double err_tol;          /* < 0: 0.01 * ||g|| (lib/depth_optimizer.cc:247) */
    double q_tol;
    double const* H;
    double const* P;
    double const* g;         /* b = -g (lib/depth_optimizer.cc:251) */
    uint16_t const* rowmask; /* bit k: block k of the node's row exists */
    uint32_t const* rows;    /* nodes with a non-empty row, ascending */
    unsigned long long const* counts;   /* [0] blocks, [1] rows of the system */
    double* x;
    double* r;
    double* d;               /* search direction, double buffered */
    double* d2;
    double* Ad;
    double* z;
    double* partials;        /* [slot][CG_MAX_BLOCKS] */
    unsigned int* sync;      /* barrier counter */
    double* result;          /* [0] iterations, [1] info, [2] isnan(x[0]) */
};

__device__ __forceinline__ void
grid_barrier (unsigned int* counter, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        epoch += 1;
        unsigned int const target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        /* spin with relaxed loads (served by L2), one fence at the end: an
         * acquire load in the loop invalidates the SM's L1 on every poll
         * (CCTL.IVALL, ~40 polls per barrier) -- under the other CTA of the
         * SM, which may still be gathering vector entries through L1 */
        unsigned int v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];"
                : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long
now_ns (void)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

/* Streaming load of the four Hessian entries of one block row: one 256-bit
 * request per thread (LDG.E.NA.EFL2.256), not allocated in L1 -- L1 is left
 * to the vector entries the nine rows around a node share -- and marked
 * evict-first in L2 (H is 148 MB, read once per iteration). */
__device__ __forceinline__ void
ld_stream (double const* p, double2& h01, double2& h23)
{
    unsigned long long a, b, c, d;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.b64 "
        "{%0, %1, %2, %3}, [%4];"
        : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    h01.x = __longlong_as_double(a); h01.y = __longlong_as_double(b);
    h23.x = __longlong_as_double(c); h23.y = __longlong_as_double(d);
}

/* 16-byte load of a vector entry pair other rows re-use from L1. */
__device__ __forceinline__ double2
ld_vec (double const* p)
{
    return *reinterpret_cast<double2 const*>(p);
}

/* 16-byte load with an L2 eviction-priority hint (P: keep resident). */
__device__ __forceinline__ double2
ld_hint (double const* p, unsigned long long policy)
{
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
        : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
    return v;
}

__device__ __forceinline__ unsigned long long
policy_evict_last (void)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;"
        : "=l"(pol));
    return pol;
}

/* Sums of NV values over the block, each in a fixed order (warp shuffle
 * tree, then the warps' results left to right); valid in thread 0. */
template <int NV>
__device__ __forceinline__ void
block_sums (double (&v)[NV], double* s_red)
{
#pragma unroll
    for (int j = 0; j < NV; ++j)
        for (int off = 16; off > 0; off >>= 1)
            v[j] += __shfl_down_sync(0xffffffffu, v[j], off);
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j)
            s_red[j * (CG_THREADS / 32) + warp] = v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j)
        {
            double total = 0.0;
            for (int i = 0; i < CG_THREADS / 32; ++i)
                total += s_red[j * (CG_THREADS / 32) + i];
            v[j] = total;
        }
    }
}

/* Every block sums all per-block partials of slots first .. first+NV-1 in
 * the same order: warp j takes slot first+j, lane l adds partials l, l+32,
 * ... in sequence (loads issued in batches ahead of the adds), then the
 * shuffle tree. Results in s_bcast[0..NV-1], valid for all threads. */
template <int NV>
__device__ __forceinline__ void
all_sums (double const* partials, int first, double* s_bcast)
{
    __syncthreads();
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < NV)
    {
        double const* p = partials + (first + warp) * CG_MAX_BLOCKS;
        int const nb = gridDim.x;
        double v = 0.0;
        for (int base = lane; base < nb; base += 32 * 8)
        {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                t[u] = (base + 32 * u < nb) ? __ldcg(p + base + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + 32 * u < nb)
                    v += t[u];
        }
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0)
            s_bcast[warp] = v;
    }
    __syncthreads();
}

/*
 * Plain (weak, L1-cached) loads are correct for the vectors other CTAs wrote
 * in the previous phase: the grid barrier is a release (fence + atomic) /
 * acquire (ld.acquire.gpu + fence) pair extended to the CTA by bar.sync, so
 * causality order covers them, and the gpu-scope fence after the spin drops
 * the SM's L1 lines. Each vector entry is used by up to nine rows, most of
 * them in the same CTA pass: L1 serves the re-use instead of L2.
 *
 * VecOp: the vector the matrix is applied to. For CG it is the NEW search
 * direction z + beta * d_old, formed on the fly for the nine neighbours, so
 * the direction update (lib/conjugate_gradient.h:192-198) needs no pass and
 * no grid barrier of its own.
 */
struct PlainVec
{
    double const* v;
    __device__ __forceinline__ void load (int node, double* out) const
    {
        double2 const a = *reinterpret_cast<double2 const*>(
            v + static_cast<size_t>(node) * 4);
        double2 const b = *reinterpret_cast<double2 const*>(
            v + static_cast<size_t>(node) * 4 + 2);
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
    }
};

struct DirVec
{
    double const* z;
    double const* d_old;
    double beta;
    __device__ __forceinline__ void load (int node, double* out) const
    {
        double2 const z0 = ld_vec(z + static_cast<size_t>(node) * 4);
        double2 const z1 = ld_vec(z + static_cast<size_t>(node) * 4 + 2);
        double2 const d0 = ld_vec(d_old + static_cast<size_t>(node) * 4);
        double2 const d1 = ld_vec(d_old + static_cast<size_t>(node) * 4 + 2);
        out[0] = z0.x + d0.x * beta; out[1] = z0.y + d0.y * beta;
        out[2] = z1.x + d1.x * beta; out[3] = z1.y + d1.y * beta;
    }
};

/* (H v)[node, rp] for the thread's node and block row, blocks visited in
 * the reference's order (ascending column block,
 * lib/block_sparse_matrix.h:283-296). own[] receives v[node]. */
template <typename VecOp>
__device__ __forceinline__ double
spmv_row (CgArgs const& a, VecOp const& vec, int node, int rp,
    unsigned int mask, double* own)
{
    int const ns = a.npx + 1;
    int const ix = node % ns, iy = node / ns;
    double const* hrow = a.H + static_cast<size_t>(node) * 144 + rp * 4;
    double acc = 0.0;
    own[0] = 0.0; own[1] = 0.0; own[2] = 0.0; own[3] = 0.0;
    /* The reference drops the rows and columns of inactive nodes
     * (lib/gauss_newton_step.cc:91,101,105); here they are zero blocks, which
     * are neither fetched nor multiplied: as the active set shrinks from one
     * Newton step to the next, so does the Hessian traffic. */
    /* The mask (both nodes valid, active and inside the grid) is in a
     * register before the row starts, so the nine loads stay independent. */
    if (mask == 0)
        return 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
    {
        if (!((mask >> k) & 1u))
            continue;
        int const jx = ix + (k % 3) - 1, jy = iy + (k / 3) - 1;
        int const nj = jy * ns + jx;
        double2 h01, h23;
        ld_stream(hrow + k * 16, h01, h23);
        double v[4];
        vec.load(nj, v);
        if (k == 4)
        {
            own[0] = v[0]; own[1] = v[1]; own[2] = v[2]; own[3] = v[3];
        }
        acc += h01.x * v[0];
        acc += h01.y * v[1];
        acc += h23.x * v[2];
        acc += h23.y * v[3];
    }
    return acc;
}

/* 2 CTAs / SM: measured faster than 3 at 80 registers (fewer loads hoisted,
 * more barrier participants). */
__global__ void __launch_bounds__(CG_THREADS, 2)
cg_kernel (CgArgs const a)
{
    unsigned long long const keep = policy_evict_last();
    __shared__ double s_red[3 * CG_THREADS / 32];
    __shared__ double s_bcast[3];
    unsigned int epoch = 0;
    int const n = a.n_nodes * 4;
    int const stride = gridDim.x * CG_THREADS;
    int const t0 = blockIdx.x * CG_THREADS + threadIdx.x;
    int const quad = threadIdx.x & 28;      /* first lane of the node's quad */
    int const rp = threadIdx.x & 3;
    /* bound rounded up: whole warps iterate together (shuffles below) */
    int const n_round = ((n + CG_UF * stride - 1) / (CG_UF * stride))
        * (CG_UF * stride);

    /* r = b = -g; x = 0; z = P r; r_dot_r = z.r; ||g||^2
     * (lib/conjugate_gradient.h:85-117). d_old = 0 with beta = 0 makes the
     * first direction d = z. P is block diagonal: the four threads of a node
     * exchange their r entries by shuffle. */
    double p_zr = 0.0, p_gg = 0.0;
    for (int i = t0; i < n_round; i += stride)
    {
        bool const ok = i < n;
        double const gi = ok ? a.g[i] : 0.0;
        double const ri = -gi;
        double const r0 = __shfl_sync(0xffffffffu, ri, quad);
        double const r1 = __shfl_sync(0xffffffffu, ri, quad + 1);
        double const r2 = __shfl_sync(0xffffffffu, ri, quad + 2);
        double const r3 = __shfl_sync(0xffffffffu, ri, quad + 3);
        if (!ok)
            continue;
        double const* prow = a.P + static_cast<size_t>(i >> 2) * 16 + rp * 4;
        double2 const p01 = *reinterpret_cast<double2 const*>(prow);
        double2 const p23 = *reinterpret_cast<double2 const*>(prow + 2);
        double const zi = p01.x * r0 + p01.y * r1 + p23.x * r2 + p23.y * r3;
        a.r[i] = ri;
        a.x[i] = 0.0;
        a.z[i] = zi;
        a.d[i] = 0.0;
        a.d2[i] = 0.0;      /* rows outside the system are never written again */
        a.Ad[i] = 0.0;
        p_gg += gi * gi;
        p_zr += zi * ri;
    }
    {
        double v[2] = { p_zr, p_gg };
        block_sums<2>(v, s_red);
        if (threadIdx.x == 0)
        {
            a.partials[0 * CG_MAX_BLOCKS + blockIdx.x] = v[0];
            a.partials[1 * CG_MAX_BLOCKS + blockIdx.x] = v[1];
        }
    }
    grid_barrier(a.sync, epoch);
    all_sums<2>(a.partials, 0, s_bcast);
    double r_dot_r = s_bcast[0];
    double const gg = s_bcast[1];
    double const tol = (a.err_tol < 0.0) ? sqrt(gg) * 0.01 : a.err_tol;
    double Q0 = 0.0;     /* -x.(b + r) with x = 0 */
    double beta = 0.0;
    double* d_old = a.d;
    double* d_new = a.d2;

    /* the masks do not change during a solve: the first pass's is fetched
     * once, the others one pass ahead */
    int const n_rows = static_cast<int>(a.counts[1]);
    int const quad0 = t0 >> 2, quads = stride >> 2;
    int const node_first = (quad0 < n_rows) ? static_cast<int>(a.rows[quad0])
        : 0;
    unsigned int const mask_first = (quad0 < n_rows) ? a.rowmask[node_first]
        : 0u;
    int iter = 1;
    int info = SMVSB_CG_MAX_ITERATIONS;
    unsigned long long tm[4] = {0, 0, 0, 0};
    for (; iter < a.max_iter; ++iter)
    {
        unsigned long long const t_a = now_ns();
        /* d = z + beta d_old (:192-198 of the previous iteration);
         * Ad = A d; alpha = r_dot_r / d.Ad (:126-127) */
        DirVec dir;
        dir.z = a.z; dir.d_old = d_old; dir.beta = beta;
        double p_dAd = 0.0;
        unsigned int mask = mask_first;
        int node = node_first;
        for (int q = quad0; q < n_rows; q += quads)
        {
            /* next pass's row and mask travel while this pass streams */
            int const qn = q + quads;
            int const node_next = (qn < n_rows) ? static_cast<int>(a.rows[qn])
                : 0;
            unsigned int const mask_next = (qn < n_rows)
                ? a.rowmask[node_next] : 0u;
            double own[4];
            int const i = node * 4 + rp;
            double const v = spmv_row(a, dir, node, rp, mask, own);
            node = node_next;
            mask = mask_next;
            double const di = (rp == 0) ? own[0] : (rp == 1) ? own[1]
                : (rp == 2) ? own[2] : own[3];
            a.Ad[i] = v;
            d_new[i] = di;
            p_dAd += v * di;
        }
        int const slot = 2 + 4 * (iter & 1);
        {
            double v[1] = { p_dAd };
            block_sums<1>(v, s_red);
            if (threadIdx.x == 0)
                a.partials[slot * CG_MAX_BLOCKS + blockIdx.x] = v[0];
        }
        unsigned long long const t_b = now_ns();
        grid_barrier(a.sync, epoch);
        unsigned long long const t_c = now_ns();
        tm[0] += t_b - t_a; tm[1] += t_c - t_b;
        all_sums<1>(a.partials, slot, s_bcast);
        double const dAd = s_bcast[0];
        double const alpha = r_dot_r / dAd;

        /* x += alpha d; r -= alpha Ad; r.r; Q1 = -x.(b + r); z = P r; z.r
         * (:130-181) */
        double p_rr = 0.0, p_q = 0.0, p_zr2 = 0.0;
        /* CG_UF entries per thread in flight: the pass is latency bound */
        for (int i0 = t0; i0 < n_round; i0 += CG_UF * stride)
        {
            double xv[CG_UF], rv[CG_UF], gv[CG_UF];
            double2 p01[CG_UF], p23[CG_UF];
#pragma unroll
            for (int u = 0; u < CG_UF; ++u)
            {
                int const i = i0 + u * stride;
                xv[u] = 0.0; rv[u] = 0.0; gv[u] = 0.0;
                p01[u] = make_double2(0, 0); p23[u] = p01[u];
                if (i < n)
                {
                    double const dn = d_new[i], ad = a.Ad[i];
                    gv[u] = a.g[i];
                    xv[u] = a.x[i]; rv[u] = a.r[i];
                    double const* prow = a.P + static_cast<size_t>(i >> 2) * 16
                        + rp * 4;
                    p01[u] = ld_hint(prow, keep);
                    p23[u] = ld_hint(prow + 2, keep);
                    xv[u] += dn * alpha; rv[u] -= ad * alpha;
                }
            }
#pragma unroll
            for (int u = 0; u < CG_UF; ++u)
            {
                int const i = i0 + u * stride;
                double const q0 = __shfl_sync(0xffffffffu, rv[u], quad);
                double const q1 = __shfl_sync(0xffffffffu, rv[u], quad + 1);
                double const q2 = __shfl_sync(0xffffffffu, rv[u], quad + 2);
                double const q3 = __shfl_sync(0xffffffffu, rv[u], quad + 3);
                if (i < n)
                {
                    double const zi = p01[u].x * q0 + p01[u].y * q1
                        + p23[u].x * q2 + p23[u].y * q3;
                    a.x[i] = xv[u]; a.r[i] = rv[u];
                    a.z[i] = zi;
                    p_rr += rv[u] * rv[u];
                    p_q += xv[u] * (rv[u] - gv[u]);
                    p_zr2 += zi * rv[u];
                }
            }
        }
        {
            double v[3] = { p_rr, p_q, p_zr2 };
            block_sums<3>(v, s_red);
            if (threadIdx.x == 0)
            {
                a.partials[(slot + 1) * CG_MAX_BLOCKS + blockIdx.x] = v[0];
                a.partials[(slot + 2) * CG_MAX_BLOCKS + blockIdx.x] = v[1];
                a.partials[(slot + 3) * CG_MAX_BLOCKS + blockIdx.x] = v[2];
            }
        }
        unsigned long long const t_d = now_ns();
        grid_barrier(a.sync, epoch);
        all_sums<3>(a.partials, slot + 1, s_bcast);
        tm[2] += t_d - t_c; tm[3] += now_ns() - t_d;
        double const new_rr = s_bcast[0];
        double const xbr = s_bcast[1];
        double const new_zr = s_bcast[2];

        if (new_rr < tol)
        {
            info = SMVSB_CG_CONVERGENCE;
            break;
        }
        double const Q1 = -1.0 * xbr;
        double const zeta = iter * (Q1 - Q0) / Q1;
        if (zeta < a.q_tol)
        {
            info = SMVSB_CG_CONVERGENCE;
            break;
        }
        Q0 = Q1;
        beta = new_zr / r_dot_r;
        r_dot_r = new_zr;
        double* const tmp = d_old; d_old = d_new; d_new = tmp;
    }

    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        a.result[0] = iter;
        a.result[1] = info;
        for (int i = 0; i < 4; ++i)
            a.result[4 + i] = static_cast<double>(tm[i]);
    }
}


} /* namespace v1 */

void
cg_v1_launch (smvsb_ctx* c, int max_iter, double err_tol, double q_tol)
{
    v1::CgArgs a;
    a.n_nodes = c->n_nodes; a.npx = c->npx; a.npy = c->npy;
    a.max_iter = max_iter; a.err_tol = err_tol; a.q_tol = q_tol;
    a.H = c->H.p; a.P = c->P.p; a.g = c->g.p;
    a.rowmask = c->cg_rowmask.p; a.rows = c->cg_row_list.p;
    a.counts = c->cg_counts.p;
    a.x = c->x.p; a.r = c->r.p; a.d = c->d.p; a.d2 = c->d2.p;
    a.Ad = c->Ad.p; a.z = c->z.p;
    a.partials = c->cg_partials.p; a.sync = c->cg_sync.p;
    a.result = c->cg_result.p;
    size_t const n = static_cast<size_t>(c->n_nodes) * 4;
    int grid = c->num_sms * 2;
    int const need = static_cast<int>((n + v1::CG_THREADS - 1) / v1::CG_THREADS);
    grid = std::max(1, std::min(std::min(grid, need), v1::CG_MAX_BLOCKS));
    CUDA_CHECK(cudaMemsetAsync(c->cg_sync.p, 0, sizeof(unsigned int),
        c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->cg_result.p, 0, 16 * sizeof(double),
        c->stream));
    void* params[] = { &a };
    CUDA_CHECK(cudaLaunchCooperativeKernel((void const*)v1::cg_kernel,
        dim3(grid), dim3(v1::CG_THREADS), params, 0, c->stream));
    c->cg_grid = grid;
}

} /* namespace smvsb */
