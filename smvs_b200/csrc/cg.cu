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
 * round trip.
 */
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
    double* d;
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

/* (H v)[node, rp] for the thread's node and block row. */
__device__ __forceinline__ double
spmv_row (CgArgs const& a, double const* __restrict__ v, int node, int rp)
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
        double2 const v01 = __ldcg(reinterpret_cast<double2 const*>(
            v + static_cast<size_t>(nj) * 4));
        double2 const v23 = __ldcg(reinterpret_cast<double2 const*>(
            v + static_cast<size_t>(nj) * 4 + 2));
        acc += h01.x * v01.x;
        acc += h01.y * v01.y;
        acc += h23.x * v23.x;
        acc += h23.y * v23.y;
    }
    return acc;
}

/* z[node, rp] = (P r)[node, rp] */
__device__ __forceinline__ double
precond_row (CgArgs const& a, double const* __restrict__ r, int node, int rp)
{
    double const* prow = a.P + static_cast<size_t>(node) * 16 + rp * 4;
    double const* rv = r + static_cast<size_t>(node) * 4;
    return prow[0] * rv[0] + prow[1] * rv[1] + prow[2] * rv[2]
        + prow[3] * rv[3];
}

__global__ void __launch_bounds__(CG_THREADS)
cg_kernel (CgArgs const a)
{
    __shared__ double s_red[CG_THREADS / 32];
    __shared__ double s_bcast;
    unsigned int epoch = 0;
    int const n = a.n_nodes * 4;
    int const stride = gridDim.x * CG_THREADS;
    int const t0 = blockIdx.x * CG_THREADS + threadIdx.x;

    /* r = b = -g; x = 0; z = P r; d = z; r_dot_r = z.r; ||g||^2
     * (lib/conjugate_gradient.h:85-117) */
    double p_zr = 0.0, p_gg = 0.0;
    for (int i = t0; i < n; i += stride)
    {
        double const gi = a.g[i];
        a.r[i] = -gi;
        a.x[i] = 0.0;
        p_gg += gi * gi;
    }
    grid_barrier(a.sync, epoch);
    for (int i = t0; i < n; i += stride)
    {
        double const zi = precond_row(a, a.r, i >> 2, i & 3);
        a.d[i] = zi;
        p_zr += zi * a.r[i];
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

    int iter = 1;
    int info = SMVSB_CG_MAX_ITERATIONS;
    for (; iter < a.max_iter; ++iter)
    {
        /* Ad = A d; alpha = r_dot_r / d.Ad  (:126-127) */
        double p_dAd = 0.0;
        for (int i = t0; i < n; i += stride)
        {
            double const v = spmv_row(a, a.d, i >> 2, i & 3);
            a.Ad[i] = v;
            p_dAd += v * a.d[i];
        }
        tot = block_sum(p_dAd, s_red);
        int const slot = 2 + 4 * (iter & 1);
        if (threadIdx.x == 0)
            a.partials[slot * CG_MAX_BLOCKS + blockIdx.x] = tot;
        grid_barrier(a.sync, epoch);
        double const dAd = all_sum(a.partials, slot, &s_bcast);
        double const alpha = r_dot_r / dAd;

        /* x += alpha d; r -= alpha Ad; r.r; Q1 = -x.(b + r); z = P r; z.r
         * (:130-181). The preconditioner is block diagonal, so z is local
         * to the node's four threads -- but they must see each other's
         * updated r, hence the two passes with a block-level barrier only
         * (a node's four entries always live in the same block). */
        double p_rr = 0.0, p_q = 0.0;
        for (int i = t0; i < n; i += stride)
        {
            double const xi = a.x[i] + a.d[i] * alpha;
            double const ri = a.r[i] - a.Ad[i] * alpha;
            a.x[i] = xi;
            a.r[i] = ri;
            p_rr += ri * ri;
            p_q += xi * (ri - a.g[i]);
        }
        __syncthreads();
        double p_zr2 = 0.0;
        for (int i = t0; i < n; i += stride)
        {
            double const zi = precond_row(a, a.r, i >> 2, i & 3);
            a.z[i] = zi;
            p_zr2 += zi * a.r[i];
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
        grid_barrier(a.sync, epoch);
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

        /* d = z + beta d (:192-198) */
        double const beta = new_zr / r_dot_r;
        for (int i = t0; i < n; i += stride)
            a.d[i] = a.z[i] + a.d[i] * beta;
        r_dot_r = new_zr;
        grid_barrier(a.sync, epoch);
    }

    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        a.result[0] = iter;
        a.result[1] = info;
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
    if (i < a.n_nodes * 4)
        y[i] = spmv_row(a, x, i >> 2, i & 3);
}

CgArgs
make_args (smvsb_ctx* c)
{
    CgArgs a;
    a.n_nodes = c->n_nodes; a.npx = c->npx; a.npy = c->npy;
    a.max_iter = 0; a.err_tol = 0; a.q_tol = 0;
    a.H = c->H.p; a.P = c->P.p; a.g = c->g.p;
    a.x = c->x.p; a.r = c->r.p; a.d = c->d.p; a.Ad = c->Ad.p; a.z = c->z.p;
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
    c->x.reserve(n); c->r.reserve(n); c->d.reserve(n); c->Ad.reserve(n);
    c->z.reserve(n);
    c->cg_partials.reserve(10 * CG_MAX_BLOCKS);
    c->cg_sync.reserve(1);
    c->cg_result.reserve(4);

    CgArgs a = make_args(c);
    a.max_iter = max_iter; a.err_tol = err_tol; a.q_tol = q_tol;

    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
        cg_kernel, CG_THREADS, 0));
    if (per_sm < 1)
        throw Error(SMVSB_ERR_CUDA, "cg_kernel does not fit on an SM");
    int grid = c->num_sms * per_sm;
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

    double res[3];
    CUDA_CHECK(cudaMemcpyAsync(res, c->cg_result.p, sizeof(res),
        cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (iters) *iters = static_cast<int>(res[0]);
    if (info) *info = static_cast<int>(res[1]);
    if (x0_nan) *x0_nan = (res[2] != 0.0);
}

} /* namespace smvsb */
