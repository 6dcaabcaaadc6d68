/*
 * cg.cu -- ConjugateGradient::solve (lib/conjugate_gradient.h:72-202) with
 * BlockSparseMatrix<4>::multiply (lib/block_sparse_matrix.h:276-298) and the
 * SSEVector updates (lib/sse_vector.cc) as ONE persistent kernel.
 *
 * The Hessian lives in a fixed 3x3-stencil block row format
 * H[node][9][4][4] (a node couples only to its 8 grid neighbours), the
 * preconditioner as P[node][4][4]. Four threads own one node (one per block
 * row); a warp therefore streams 8 complete 1152-byte block rows per pass,
 * every 128-byte line fully used.
 *
 * The whole solve -- SpMV, the dot products, the reference's two stopping
 * tests (residual < tolerance and the Nash/Sofer quadratic-model test),
 * preconditioning and direction update -- runs on the device; grid-wide
 * reductions go through per-block partial sums that every block re-sums in
 * the same fixed order, so the result is deterministic run to run and the
 * stopping decision is taken identically by all blocks without a host
 * round trip. Two grid barriers per iteration: the direction update
 * d = z + beta d is folded into the next SpMV (formed on the fly for the nine
 * neighbours), the block-diagonal preconditioner into the residual update
 * (quad shuffles).
 */
#include <cstdlib>

#include "common.cuh"

namespace smvsb {

namespace {

constexpr int CG_THREADS = 256;
constexpr int CG_MAX_BLOCKS = 1024;

struct CgArgs
{
    int n_nodes, npx, npy;
    int max_iter;
    double err_tol;          /* < 0: 0.01 * ||g|| (lib/depth_optimizer.cc:247) */
    double q_tol;
    double const* H;
    double const* P;
    double const* g;         /* b = -g (lib/depth_optimizer.cc:251) */
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
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];"
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

/* Sum of `v` over the block in a fixed order; valid in thread 0. */
__device__ __forceinline__ double
block_sum (double v, double* s_red)
{
    for (int off = 16; off > 0; off >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, off);
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0)
        s_red[warp] = v;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < CG_THREADS / 32; ++i)
            total += s_red[i];
    return total;
}

/* Every block sums all per-block partials of `slot` in the same order. */
__device__ __forceinline__ double
all_sum (double const* partials, int slot, double* s_bcast)
{
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32)
            v += __ldcg(partials + slot * CG_MAX_BLOCKS + i);
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0)
            *s_bcast = v;
    }
    __syncthreads();
    return *s_bcast;
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
        double2 const z0 = *reinterpret_cast<double2 const*>(
            z + static_cast<size_t>(node) * 4);
        double2 const z1 = *reinterpret_cast<double2 const*>(
            z + static_cast<size_t>(node) * 4 + 2);
        double2 const d0 = *reinterpret_cast<double2 const*>(
            d_old + static_cast<size_t>(node) * 4);
        double2 const d1 = *reinterpret_cast<double2 const*>(
            d_old + static_cast<size_t>(node) * 4 + 2);
        out[0] = z0.x + d0.x * beta; out[1] = z0.y + d0.y * beta;
        out[2] = z1.x + d1.x * beta; out[3] = z1.y + d1.y * beta;
    }
};

/* (H v)[node, rp] for the thread's node and block row, blocks visited in
 * the reference's order (ascending column block,
 * lib/block_sparse_matrix.h:283-296). own[] receives v[node]. */
template <typename VecOp>
__device__ __forceinline__ double
spmv_row (CgArgs const& a, VecOp const& vec, int node, int rp, double* own)
{
    int const ns = a.npx + 1;
    int const ix = node % ns, iy = node / ns;
    double const* hrow = a.H + static_cast<size_t>(node) * 144 + rp * 4;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
    {
        int const jx = ix + (k % 3) - 1, jy = iy + (k / 3) - 1;
        if (jx < 0 || jx > a.npx || jy < 0 || jy > a.npy)
            continue;
        int const nj = jy * ns + jx;
        double2 const h01 = __ldcs(reinterpret_cast<double2 const*>(
            hrow + k * 16));
        double2 const h23 = __ldcs(reinterpret_cast<double2 const*>(
            hrow + k * 16 + 2));
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
    __shared__ double s_red[CG_THREADS / 32];
    __shared__ double s_bcast;
    unsigned int epoch = 0;
    int const n = a.n_nodes * 4;
    int const stride = gridDim.x * CG_THREADS;
    int const t0 = blockIdx.x * CG_THREADS + threadIdx.x;
    int const quad = threadIdx.x & 28;      /* first lane of the node's quad */
    int const rp = threadIdx.x & 3;
    /* bound rounded up: whole warps iterate together (shuffles below) */
    int const n_round = ((n + 2 * stride - 1) / (2 * stride)) * (2 * stride);

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
        double const zi = prow[0] * r0 + prow[1] * r1 + prow[2] * r2
            + prow[3] * r3;
        a.r[i] = ri;
        a.x[i] = 0.0;
        a.z[i] = zi;
        a.d[i] = 0.0;
        p_gg += gi * gi;
        p_zr += zi * ri;
    }
    double tot = block_sum(p_zr, s_red);
    if (threadIdx.x == 0) a.partials[0 * CG_MAX_BLOCKS + blockIdx.x] = tot;
    tot = block_sum(p_gg, s_red);
    if (threadIdx.x == 0) a.partials[1 * CG_MAX_BLOCKS + blockIdx.x] = tot;
    grid_barrier(a.sync, epoch);
    double r_dot_r = all_sum(a.partials, 0, &s_bcast);
    double const gg = all_sum(a.partials, 1, &s_bcast);
    double const tol = (a.err_tol < 0.0) ? sqrt(gg) * 0.01 : a.err_tol;
    double Q0 = 0.0;     /* -x.(b + r) with x = 0 */
    double beta = 0.0;
    double* d_old = a.d;
    double* d_new = a.d2;

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
        for (int i = t0; i < n; i += stride)
        {
            double own[4];
            double const v = spmv_row(a, dir, i >> 2, rp, own);
            double const di = (rp == 0) ? own[0] : (rp == 1) ? own[1]
                : (rp == 2) ? own[2] : own[3];
            a.Ad[i] = v;
            d_new[i] = di;
            p_dAd += v * di;
        }
        tot = block_sum(p_dAd, s_red);
        int const slot = 2 + 4 * (iter & 1);
        if (threadIdx.x == 0)
            a.partials[slot * CG_MAX_BLOCKS + blockIdx.x] = tot;
        unsigned long long const t_b = now_ns();
        grid_barrier(a.sync, epoch);
        unsigned long long const t_c = now_ns();
        tm[0] += t_b - t_a; tm[1] += t_c - t_b;
        double const dAd = all_sum(a.partials, slot, &s_bcast);
        double const alpha = r_dot_r / dAd;

        /* x += alpha d; r -= alpha Ad; r.r; Q1 = -x.(b + r); z = P r; z.r
         * (:130-181) */
        double p_rr = 0.0, p_q = 0.0, p_zr2 = 0.0;
        /* two elements per thread in flight: the pass is latency bound */
        for (int i0 = t0; i0 < n_round; i0 += 2 * stride)
        {
            int const i1 = i0 + stride;
            bool const ok0 = i0 < n, ok1 = i1 < n;
            double x0 = 0.0, ra = 0.0, x1 = 0.0, rb = 0.0;
            double g0 = 0.0, g1 = 0.0;
            double2 pa01 = make_double2(0, 0), pa23 = pa01, pb01 = pa01,
                pb23 = pa01;
            if (ok0)
            {
                x0 = a.x[i0]; ra = a.r[i0];
                double const dn = d_new[i0], ad = a.Ad[i0];
                g0 = a.g[i0];
                pa01 = __ldcs(reinterpret_cast<double2 const*>(
                    a.P + static_cast<size_t>(i0 >> 2) * 16 + rp * 4));
                pa23 = __ldcs(reinterpret_cast<double2 const*>(
                    a.P + static_cast<size_t>(i0 >> 2) * 16 + rp * 4 + 2));
                x0 += dn * alpha; ra -= ad * alpha;
            }
            if (ok1)
            {
                x1 = a.x[i1]; rb = a.r[i1];
                double const dn = d_new[i1], ad = a.Ad[i1];
                g1 = a.g[i1];
                pb01 = __ldcs(reinterpret_cast<double2 const*>(
                    a.P + static_cast<size_t>(i1 >> 2) * 16 + rp * 4));
                pb23 = __ldcs(reinterpret_cast<double2 const*>(
                    a.P + static_cast<size_t>(i1 >> 2) * 16 + rp * 4 + 2));
                x1 += dn * alpha; rb -= ad * alpha;
            }
            double const a0 = __shfl_sync(0xffffffffu, ra, quad);
            double const a1 = __shfl_sync(0xffffffffu, ra, quad + 1);
            double const a2 = __shfl_sync(0xffffffffu, ra, quad + 2);
            double const a3 = __shfl_sync(0xffffffffu, ra, quad + 3);
            double const b0 = __shfl_sync(0xffffffffu, rb, quad);
            double const b1 = __shfl_sync(0xffffffffu, rb, quad + 1);
            double const b2 = __shfl_sync(0xffffffffu, rb, quad + 2);
            double const b3 = __shfl_sync(0xffffffffu, rb, quad + 3);
            if (ok0)
            {
                double const zi = pa01.x * a0 + pa01.y * a1 + pa23.x * a2
                    + pa23.y * a3;
                a.x[i0] = x0; a.r[i0] = ra; a.z[i0] = zi;
                p_rr += ra * ra;
                p_q += x0 * (ra - g0);
                p_zr2 += zi * ra;
            }
            if (ok1)
            {
                double const zi = pb01.x * b0 + pb01.y * b1 + pb23.x * b2
                    + pb23.y * b3;
                a.x[i1] = x1; a.r[i1] = rb; a.z[i1] = zi;
                p_rr += rb * rb;
                p_q += x1 * (rb - g1);
                p_zr2 += zi * rb;
            }
        }
        tot = block_sum(p_rr, s_red);
        if (threadIdx.x == 0)
            a.partials[(slot + 1) * CG_MAX_BLOCKS + blockIdx.x] = tot;
        tot = block_sum(p_q, s_red);
        if (threadIdx.x == 0)
            a.partials[(slot + 2) * CG_MAX_BLOCKS + blockIdx.x] = tot;
        tot = block_sum(p_zr2, s_red);
        if (threadIdx.x == 0)
            a.partials[(slot + 3) * CG_MAX_BLOCKS + blockIdx.x] = tot;
        unsigned long long const t_d = now_ns();
        grid_barrier(a.sync, epoch);
        tm[2] += t_d - t_c; tm[3] += now_ns() - t_d;
        double const new_rr = all_sum(a.partials, slot + 1, &s_bcast);
        double const xbr = all_sum(a.partials, slot + 2, &s_bcast);
        double const new_zr = all_sum(a.partials, slot + 3, &s_bcast);

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

__global__ void
cg_finish_kernel (double const* x, double* result)
{
    result[2] = isnan(x[0]) ? 1.0 : 0.0;
}

__global__ void
spmv_kernel (CgArgs const a, double const* __restrict__ x,
    double* __restrict__ y)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_nodes * 4)
        return;
    PlainVec vec;
    vec.v = x;
    double own[4];
    y[i] = spmv_row(a, vec, i >> 2, i & 3, own);
}

CgArgs
make_args (smvsb_ctx* c)
{
    CgArgs a;
    a.n_nodes = c->n_nodes; a.npx = c->npx; a.npy = c->npy;
    a.max_iter = 0; a.err_tol = 0; a.q_tol = 0;
    a.H = c->H.p; a.P = c->P.p; a.g = c->g.p;
    a.x = c->x.p; a.r = c->r.p; a.d = c->d.p; a.d2 = c->d2.p;
    a.Ad = c->Ad.p; a.z = c->z.p;
    a.partials = c->cg_partials.p; a.sync = c->cg_sync.p;
    a.result = c->cg_result.p;
    return a;
}

} /* namespace */

void
launch_spmv (smvsb_ctx* c, double const* x, double* y)
{
    CgArgs a = make_args(c);
    int const n = c->n_nodes * 4;
    spmv_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a, x, y);
    smvsb::count_launches(c, 1);
    CUDA_CHECK(cudaGetLastError());
}

void
run_cg (smvsb_ctx* c, int max_iter, double err_tol, double q_tol, int* iters,
    int* info, bool* x0_nan)
{
    size_t const n = static_cast<size_t>(c->n_nodes) * 4;
    c->x.reserve(n); c->r.reserve(n); c->d.reserve(n); c->d2.reserve(n);
    c->Ad.reserve(n); c->z.reserve(n);
    c->cg_partials.reserve(10 * CG_MAX_BLOCKS);
    c->cg_sync.reserve(1);
    c->cg_result.reserve(16);

    CgArgs a = make_args(c);
    a.max_iter = max_iter; a.err_tol = err_tol; a.q_tol = q_tol;

    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
        cg_kernel, CG_THREADS, 0));
    if (per_sm < 1)
        throw Error(SMVSB_ERR_CUDA, "cg_kernel does not fit on an SM");
    int grid = c->num_sms * std::min(per_sm, 2);
    int const need = static_cast<int>((n + CG_THREADS - 1) / CG_THREADS);
    grid = std::max(1, std::min(std::min(grid, need), CG_MAX_BLOCKS));

    CUDA_CHECK(cudaMemsetAsync(c->cg_sync.p, 0, sizeof(unsigned int),
        c->stream));
    void* params[] = { &a };
    CUDA_CHECK(cudaLaunchCooperativeKernel((void const*)cg_kernel, dim3(grid),
        dim3(CG_THREADS), params, 0, c->stream));
    cg_finish_kernel<<<1, 1, 0, c->stream>>>(c->x.p, c->cg_result.p);
    smvsb::count_launches(c, 2);
    CUDA_CHECK(cudaGetLastError());

    double res[10];
    CUDA_CHECK(cudaMemcpyAsync(res, c->cg_result.p, sizeof(res),
        cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (getenv("SMVSB_CG_TIMING"))
        fprintf(stderr, "cg: iters %d grid %d | us/iter: spmv %.1f wait %.1f | "
            "update %.1f wait %.1f\n", (int)res[0], grid,
            res[4] / 1e3 / res[0], res[5] / 1e3 / res[0], res[6] / 1e3 / res[0],
            res[7] / 1e3 / res[0]);
    if (iters) *iters = static_cast<int>(res[0]);
    if (info) *info = static_cast<int>(res[1]);
    if (x0_nan) *x0_nan = (res[2] != 0.0);
}

} /* namespace smvsb */
