/*
 * cg.cu -- ConjugateGradient::solve (lib/conjugate_gradient.h:72-202) with
 * BlockSparseMatrix<4>::multiply (lib/block_sparse_matrix.h:276-298) and the
 * SSEVector updates (lib/sse_vector.cc) as ONE persistent kernel -- for one
 * view, or for several independent views (one system each) in the same
 * launch.
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
 *
 * All three phases (initialisation, SpMV, vector update) walk the compacted
 * list of the system's block rows (the nodes that are valid and active,
 * lib/gauss_newton_step.cc:91-105), so a system that has shrunk to 10 % of
 * the grid costs 10 % of the vector traffic as well.
 *
 * Several views per launch (smvsb_newton_loop_batch; the reference runs one
 * view per pool thread, app/smvsrecon.cc:658-733): what an iteration costs
 * besides the Hessian stream is two grid-wide synchronisations, a fixed
 * ~6 us whatever the system size. With V views in one launch every CTA
 * works through "its" rows of view 0, then of view 1, ... between two
 * barriers, so the fixed cost is paid once per V views. CTA b handles of
 * every view exactly the rows it would handle in a launch of its own, the
 * per-view partial sums are kept apart and re-summed in the same order, and
 * every view takes its stopping decisions for itself: the result of a view
 * is bitwise the one of a single-view launch. Views that have converged are
 * skipped.
 *
 * Measured and NOT kept (1920x1080 scale 2, B200): parking what only the
 * owning thread touches (x, r, A d, its row of P) in shared memory for the
 * whole solve. It shortens the vector-update phase (5.1 -> 3.9 us) but every
 * KB of shared memory is a KB less L1, and the SpMV needs L1 both for the
 * vector entries nine rows share and as landing space for ~150 KB of loads in
 * flight per SM: 22.3 -> 24.2 us with 86 KB of shared memory, 37 us with
 * 200 KB. Reading the blocks left of / above the diagonal as transposes of the
 * neighbours' mirror blocks (256-bit row load + 4x4 transpose across the quad,
 * slots 0..3 never fetched): the second use of a line does not hit L2 often
 * enough, SpMV 22.3 -> 43.6 us; and the assembled H is symmetric only to
 * rounding, so x changes. Likewise cp.async.bulk.prefetch.L2 of the rows a CTA reads first,
 * issued while HBM idles in the vector-update phase: the SpMV gains 1.3 us,
 * the update phase and the barriers lose more. Keeping part of H in L2 from
 * one iteration to the next (round 2: the CTA's first passes loaded with
 * L2::evict_last, the rest evict_first, 0 .. 148 MB kept, with and without a
 * persisting-L2 carve-out of 71 MB): SpMV 25.2 us with nothing kept, 24.8 with
 * 48 MB, 26.3 with 64 MB, 29.5 with everything, and the update phase grows from
 * 5.7 to 6.5 / 7.7 / 8.9 us as its vectors lose their place -- the SpMV phase is
 * not limited by what comes out of DRAM but by what the SMs can keep in flight
 * through L2 (H at 5.5 TB/s plus the gathered vector entries), so an L2 hit is
 * worth no more than an HBM access here (profiles/r2_cg_keep_probe.txt).
 * Fetching across the grid barriers -- the Hessian row of the first SpMV
 * pass and the thread's own operands of the first vector-update rows loaded
 * BEFORE the barrier in front of the phase, used after it: on a system of one
 * pass per CTA the SpMV phase drops from 3.3 to 1.9 us, but on the full system
 * it rises from 24.8 to 40.2 us (72 registers live across the barrier change
 * the streaming loop's schedule: fewer loads in flight), and the update-side
 * variant gains 0.7 us in its phase and loses 2.9 in the SpMV
 * (profiles/r2_cg_ahead_probe.txt). What a barrier costs CTA 0 between its
 * last store of a phase and the first instruction of the next is 2.3 - 2.8 us
 * whatever the system size. The compacted row list in 8 x 8 tiles of the node
 * grid instead of row-major strips (a CTA pass then gathers 10 x 10 instead of
 * 3 x 66 nodes' vector entries, 1.6 instead of 3.1 L2 fetches per entry): SpMV
 * 24.9 -> 25.5 us -- the eight warps of a CTA then stream eight separate
 * 9 KB pieces of H instead of one contiguous 72 KB piece.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace smvsb {

namespace {

constexpr int CG_THREADS = 256;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int CG_QUADS = CG_THREADS / 4;     /* block rows per CTA and pass */
constexpr int CG_MAX_BLOCKS = 1024;
constexpr int CG_UF = 4;          /* rows per thread in flight, update phase */
constexpr int CG_SLOTS = 10;      /* partial-sum slots per view */

/* One view's system and vectors. */
struct CgView
{
    int n_nodes, npx;
    int grid;                /* CTAs that work on this view: min(launch grid,
                                ceil(4 n_nodes / 256)) -- what a launch of its
                                own would use */
    int pad;
    double err_tol;          /* < 0: 0.01 * ||g|| (lib/depth_optimizer.cc:247) */
    double const* H;
    double const* P;
    double const* g;         /* b = -g (lib/depth_optimizer.cc:251) */
    uint16_t const* rowmask; /* bit k: block k of the node's row exists */
    uint32_t const* rows;    /* nodes with a non-empty row, ascending */
    unsigned long long const* counts;   /* [0] blocks, [1] rows of the system */
    double* x;               /* zeroed by the host before the launch */
    double* r;
    double* d;               /* search direction, double buffered */
    double* d2;
    double* Ad;
    double* z;
    double* partials;        /* [CG_SLOTS][CG_MAX_BLOCKS] */
    double* result;          /* [0] iterations, [1] info, [2] isnan(x[0]),
                                [4..7] phase times in ns (timing build) */
};

struct CgArgs
{
    int n_views;
    int max_iter;
    double q_tol;
    unsigned int* sync;      /* barrier counter */
    CgView v[SMVSB_MAX_BATCH];
};

/* Per-view scalars of the iteration, identical in every CTA. */
struct CgState
{
    double r_dot_r, Q0, beta, alpha, tol;
    double* partials;        /* copy of CgView::partials, see publish() */
    int n_rows, passes, done, iters, info;
    int grid;                /* copy of CgView::grid */
};

__device__ __forceinline__ void
grid_barrier (unsigned int* counter, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        epoch += 1;
        unsigned int const target = epoch * gridDim.x;
        /* arrive: one release reduction (the CTA's stores, ordered before it
         * by the bar.sync above, are visible to whoever observes the count);
         * 0.3 us per barrier less than fence + atomicAdd
         * (benchmarks/micro/barrier_probe.cu) */
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;"
            :: "l"(counter) : "memory");
        /* spin with relaxed loads (served by L2) and acquire once at the
         * end: an acquire load in the loop invalidates the SM's L1 on every
         * poll (CCTL.IVALL, ~40 polls per barrier) -- under the other CTA of
         * the SM, which may still be gathering vector entries through L1 */
        unsigned int v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];"
                : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];"
            : "=r"(v) : "l"(counter) : "memory");
    }
    __syncthreads();
}

template <bool TIMING>
__device__ __forceinline__ unsigned long long
now_ns (void)
{
    if (!TIMING)
        return 0;
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

/*
 * Block-wide sums in two stages, each in a fixed order. Stage 1 (flush, once
 * per view and phase): shuffle tree inside every warp, lane 0 parks the
 * warp's value in shared memory. Stage 2 (publish, once per phase): after a
 * __syncthreads one thread per (view, value) adds the warps' results left to
 * right and writes the CTA's partial sum for that view.
 */
template <int NV>
__device__ __forceinline__ void
warp_flush (double (&v)[NV], double* s_red, int view)
{
#pragma unroll
    for (int j = 0; j < NV; ++j)
        for (int off = 16; off > 0; off >>= 1)
            v[j] += __shfl_down_sync(0xffffffffu, v[j], off);
    if ((threadIdx.x & 31) == 0)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j)
            s_red[(view * 3 + j) * CG_WARPS + (threadIdx.x >> 5)] = v[j];
    }
}

/* (The view's grid size and partial-sum array are read from the shared copy:
 * indexing the kernel parameters with a per-thread view number is a divergent
 * constant-bank access -- the compiler hoists it above the branch, every
 * thread of the CTA fetches a different cache line, and the warp replays it
 * 32 times: measured 4.4 us per CG iteration for this one load.) */
template <int NV>
__device__ __forceinline__ void
publish (CgArgs const& a, CgState const* s_state, double const* s_red,
    int first_slot, bool init)
{
    __syncthreads();
    int const t = threadIdx.x;
    if (t < a.n_views * NV)
    {
        int const view = t / NV, j = t % NV;
        CgState const& S = s_state[view];
        if (static_cast<int>(blockIdx.x) < S.grid && (init || !S.done))
        {
            /* a view without rows was never flushed */
            double total = 0.0;
            if (S.passes > 0)
                for (int i = 0; i < CG_WARPS; ++i)
                    total += s_red[(view * 3 + j) * CG_WARPS + i];
            S.partials[(first_slot + j) * CG_MAX_BLOCKS + blockIdx.x] = total;
        }
    }
}

/* Every CTA sums, per view, the partials of slots first .. first+NV-1 of all
 * the view's CTAs in the same order: one warp per (view, slot), lane l adds
 * partials l, l+32, ... in sequence (loads issued in batches ahead of the
 * adds), then the shuffle tree. Results in s_bcast[view * 3 + j]. */
template <int NV>
__device__ __forceinline__ void
all_sums (CgArgs const& a, CgState const* s_state, int first_slot,
    double* s_bcast, bool init)
{
    __syncthreads();
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pair = warp; pair < a.n_views * NV; pair += CG_WARPS)
    {
        int const view = pair / NV, j = pair % NV;
        if (!init && s_state[view].done)
            continue;
        double const* p = s_state[view].partials
            + (first_slot + j) * CG_MAX_BLOCKS;
        int const nb = s_state[view].grid;
        double v = 0.0;
        /* 12 loads per lane in flight: one L2 round trip for up to 384 CTAs
         * (2 x 148 on B200) */
        for (int base = lane; base < nb; base += 32 * 12)
        {
            double t[12];
#pragma unroll
            for (int u = 0; u < 12; ++u)
                t[u] = (base + 32 * u < nb) ? __ldcg(p + base + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 12; ++u)
                if (base + 32 * u < nb)
                    v += t[u];
        }
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0)
            s_bcast[view * 3 + j] = v;
    }
    __syncthreads();
}

/*
 * Plain (weak, L1-cached) loads are correct for the vectors other CTAs wrote
 * in the previous phase: the grid barrier is a release (red.release.gpu) /
 * acquire (relaxed polls + one ld.acquire.gpu) pair extended to the CTA by
 * bar.sync, so causality order covers them, and the gpu-scope acquire after
 * the spin drops the SM's L1 lines. Each vector entry is used by up to nine rows, most of
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
spmv_row (double const* __restrict__ H, int ns, VecOp const& vec, int node,
    int rp, unsigned int mask, double* own)
{
    int const ix = node % ns, iy = node / ns;
    double const* hrow = H + static_cast<size_t>(node) * 144 + rp * 4;
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

/* Does CTA b work on view v between the next two barriers? */
__device__ __forceinline__ bool
view_on (CgArgs const& a, CgState const* s_state, int v, bool init)
{
    return static_cast<int>(blockIdx.x) < s_state[v].grid
        && (init || !s_state[v].done) && s_state[v].passes > 0;
}

/* The row this thread's quad handles in pass p of view V: whether there is
 * one is decided from the position alone, so that nothing branches on the
 * loaded index and the load can travel while other work is issued. */
__device__ __forceinline__ bool
pass_row (CgView const& V, int n_rows, int pass, int& node)
{
    int const q = pass * (V.grid * CG_QUADS) + blockIdx.x * CG_QUADS
        + (threadIdx.x >> 2);
    bool const ok = q < n_rows;
    node = ok ? static_cast<int>(V.rows[q]) : 0;
    return ok;
}

/* 2 CTAs / SM: measured faster than 3 at 80 registers (fewer loads hoisted,
 * more barrier participants).
 *
 * NV = compile-time bound on the number of views (1, 2, 4, 8): the loops over
 * the views are unrolled, so a view's pointers are kernel parameters at fixed
 * offsets (constant-bank operands of the instructions that use them), not
 * values fetched per pass. The kernel is bound by the number of loads a warp
 * has in flight, and every dependent fetch in front of a pass's Hessian loads
 * lengthens the time a warp spends per pass (measured with the view indexed
 * at run time: SpMV phase 32.6 us instead of 25.2 us on the full 2 MP
 * system). */
template <bool TIMING, int NV>
__global__ void __launch_bounds__(CG_THREADS, 2)
cg_kernel (CgArgs const a)
{
    unsigned long long const keep = policy_evict_last();
    __shared__ double s_red[NV * 3 * CG_WARPS];
    __shared__ double s_bcast[NV * 3];
    __shared__ CgState s_state[NV];
    /* z's address per view, read back from shared memory where it is needed:
     * a value the compiler cannot re-derive from the kernel parameters, so it
     * stays in a register through the SpMV loop. (As a plain parameter it is
     * re-fetched from the constant bank in front of every neighbour's load
     * once registers are tight -- ncu: short-scoreboard stalls 4.0 instead of
     * 0.7 per issue, SpMV phase +20 %.) */
    __shared__ double const* s_zptr[NV];
    unsigned int epoch = 0;
    int const quad = threadIdx.x & 28;      /* first lane of the node's quad */
    int const rp = threadIdx.x & 3;

    if (threadIdx.x < a.n_views)
    {
        CgView const& V = a.v[threadIdx.x];
        CgState& S = s_state[threadIdx.x];
        S.n_rows = static_cast<int>(V.counts[1]);
        int const per_pass = V.grid * CG_QUADS;
        int const passes = (S.n_rows + per_pass - 1) / per_pass;
        S.passes = ((passes + CG_UF - 1) / CG_UF) * CG_UF;
        S.done = 0; S.iters = 0; S.info = SMVSB_CG_MAX_ITERATIONS;
        S.r_dot_r = 0.0; S.Q0 = 0.0; S.beta = 0.0; S.alpha = 0.0; S.tol = 0.0;
        S.partials = V.partials;
        S.grid = V.grid;
        s_zptr[threadIdx.x] = V.z;
    }
    __syncthreads();

    /* r = b = -g; x = 0 (host memset); z = P r; r_dot_r = z.r; ||g||^2
     * (lib/conjugate_gradient.h:85-117). d_old = 0 with beta = 0 makes the
     * first direction d = z. P is block diagonal: the four threads of a node
     * exchange their r entries by shuffle. */
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
        if (v >= a.n_views || !view_on(a, s_state, v, true))
            continue;
        CgView const& V = a.v[v];
        int const n_rows = s_state[v].n_rows, passes = s_state[v].passes;
        double acc[2] = { 0.0, 0.0 };       /* z.r, g.g */
        for (int p = 0; p < passes; ++p)
        {
            int node;
            bool const ok = pass_row(V, n_rows, p, node);
            size_t const i = static_cast<size_t>(node) * 4 + rp;
            double const gi = ok ? V.g[i] : 0.0;
            double const ri = -gi;
            double const r0 = __shfl_sync(0xffffffffu, ri, quad);
            double const r1 = __shfl_sync(0xffffffffu, ri, quad + 1);
            double const r2 = __shfl_sync(0xffffffffu, ri, quad + 2);
            double const r3 = __shfl_sync(0xffffffffu, ri, quad + 3);
            if (ok)
            {
                double const* prow = V.P + static_cast<size_t>(node) * 16
                    + rp * 4;
                double2 const p01 = *reinterpret_cast<double2 const*>(prow);
                double2 const p23 = *reinterpret_cast<double2 const*>(prow + 2);
                double const zi = p01.x * r0 + p01.y * r1 + p23.x * r2
                    + p23.y * r3;
                V.r[i] = ri;
                V.z[i] = zi;
                V.d[i] = 0.0;
                V.d2[i] = 0.0;
                V.Ad[i] = 0.0;
                acc[1] += gi * gi;
                acc[0] += zi * ri;
            }
        }
        warp_flush<2>(acc, s_red, v);
    }
    publish<2>(a, s_state, s_red, 0, true);
    grid_barrier(a.sync, epoch);
    all_sums<2>(a, s_state, 0, s_bcast, true);
    if (threadIdx.x < a.n_views)
    {
        CgState& S = s_state[threadIdx.x];
        S.r_dot_r = s_bcast[threadIdx.x * 3 + 0];
        double const gg = s_bcast[threadIdx.x * 3 + 1];
        double const et = a.v[threadIdx.x].err_tol;
        S.tol = (et < 0.0) ? sqrt(gg) * 0.01 : et;
    }
    __syncthreads();

    int iter = 1;
    unsigned long long tm[4] = {0, 0, 0, 0};
    /* The rows and masks do not change during a solve: the first pass's of
     * every view are fetched once. */
    int first_node[NV];
    unsigned int first_mask[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
        first_node[v] = 0;
        first_mask[v] = 0u;
        if (v < a.n_views && static_cast<int>(blockIdx.x) < s_state[v].grid)
        {
            int const quad0 = blockIdx.x * CG_QUADS + (threadIdx.x >> 2);
            if (quad0 < s_state[v].n_rows)
            {
                first_node[v] = static_cast<int>(a.v[v].rows[quad0]);
                first_mask[v] = a.v[v].rowmask[first_node[v]];
            }
        }
    }
    for (; iter < a.max_iter; ++iter)
    {
        unsigned long long const t_a = now_ns<TIMING>();
        /* the direction is double buffered; all views swap in lock-step */
        bool const odd = (iter & 1) != 0;
        int const slot = 2 + 4 * (iter & 1);

        /* d = z + beta d_old (:192-198 of the previous iteration);
         * Ad = A d; alpha = r_dot_r / d.Ad (:126-127). Software pipeline over
         * the passes: the row index travels two passes ahead of the stream,
         * its mask one pass ahead, and the first two rows of the NEXT view are
         * fetched while this view streams, so no load of a pass waits for
         * another one. */
        {
            /* One pass = the CTA's 64 block rows; the next pass's row index
             * and mask are fetched while this pass streams. (A deeper
             * pipeline -- index two passes ahead, mask one -- measured 1.7 us
             * slower per iteration on the full 2 MP system.) */
#pragma unroll
            for (int v = 0; v < NV; ++v)
            {
                if (v >= a.n_views || !view_on(a, s_state, v, false))
                    continue;
                CgView const& V = a.v[v];
                int const n_rows = s_state[v].n_rows;
                int const quads = V.grid * CG_QUADS;
                int const quad0 = blockIdx.x * CG_QUADS + (threadIdx.x >> 2);
                DirVec dir;
                dir.z = s_zptr[v];
                dir.d_old = odd ? V.d : V.d2;
                dir.beta = s_state[v].beta;
                double* d_new = odd ? V.d2 : V.d;
                double acc[1] = { 0.0 };
                int node = first_node[v];
                unsigned int mask = first_mask[v];
                for (int q = quad0; q < n_rows; q += quads)
                {
                    int const qn = q + quads;
                    int const node_next = (qn < n_rows)
                        ? static_cast<int>(V.rows[qn]) : 0;
                    unsigned int const mask_next = (qn < n_rows)
                        ? V.rowmask[node_next] : 0u;
                    double own[4];
                    size_t const i = static_cast<size_t>(node) * 4 + rp;
                    double const val = spmv_row(V.H, V.npx + 1, dir, node, rp,
                        mask, own);
                    node = node_next;
                    mask = mask_next;
                    double const di = (rp == 0) ? own[0] : (rp == 1) ? own[1]
                        : (rp == 2) ? own[2] : own[3];
                    V.Ad[i] = val;
                    d_new[i] = di;
                    acc[0] += val * di;
                }
                warp_flush<1>(acc, s_red, v);
            }
            publish<1>(a, s_state, s_red, slot, false);
        }
        unsigned long long const t_b = now_ns<TIMING>();
        grid_barrier(a.sync, epoch);
        unsigned long long const t_c = now_ns<TIMING>();
        all_sums<1>(a, s_state, slot, s_bcast, false);
        if (threadIdx.x < a.n_views && !s_state[threadIdx.x].done)
            s_state[threadIdx.x].alpha = s_state[threadIdx.x].r_dot_r
                / s_bcast[threadIdx.x * 3];
        __syncthreads();

        /* x += alpha d; r -= alpha Ad; r.r; Q1 = -x.(b + r); z = P r; z.r
         * (:130-181) */
#pragma unroll
        for (int v = 0; v < NV; ++v)
        {
            if (v >= a.n_views || !view_on(a, s_state, v, false))
                continue;
            CgView const& V = a.v[v];
            int const n_rows = s_state[v].n_rows;
            int const passes = s_state[v].passes;
            double const alpha = s_state[v].alpha;
            double const* d_new = odd ? V.d2 : V.d;
            double acc[3] = { 0.0, 0.0, 0.0 };    /* r.r, x.(r - g), z.r */
            int nodes[CG_UF];
            bool oks[CG_UF];
#pragma unroll
            for (int u = 0; u < CG_UF; ++u)
                oks[u] = pass_row(V, n_rows, u, nodes[u]);
            /* CG_UF rows per thread in flight: the pass is latency bound */
            for (int p = 0; p < passes; p += CG_UF)
            {
                int nodes_next[CG_UF];
                bool oks_next[CG_UF];
#pragma unroll
                for (int u = 0; u < CG_UF; ++u)
                    oks_next[u] = pass_row(V, n_rows, p + CG_UF + u,
                        nodes_next[u]);

                double xv[CG_UF], rv[CG_UF], gv[CG_UF];
                double2 p01[CG_UF], p23[CG_UF];
#pragma unroll
                for (int u = 0; u < CG_UF; ++u)
                {
                    xv[u] = 0.0; rv[u] = 0.0; gv[u] = 0.0;
                    p01[u] = make_double2(0, 0); p23[u] = p01[u];
                    if (oks[u])
                    {
                        size_t const i = static_cast<size_t>(nodes[u]) * 4 + rp;
                        double const dn = d_new[i], ad = V.Ad[i];
                        gv[u] = V.g[i];
                        xv[u] = V.x[i]; rv[u] = V.r[i];
                        double const* prow = V.P
                            + static_cast<size_t>(nodes[u]) * 16 + rp * 4;
                        p01[u] = ld_hint(prow, keep);
                        p23[u] = ld_hint(prow + 2, keep);
                        xv[u] += dn * alpha; rv[u] -= ad * alpha;
                    }
                }
#pragma unroll
                for (int u = 0; u < CG_UF; ++u)
                {
                    double const q0 = __shfl_sync(0xffffffffu, rv[u], quad);
                    double const q1 = __shfl_sync(0xffffffffu, rv[u], quad + 1);
                    double const q2 = __shfl_sync(0xffffffffu, rv[u], quad + 2);
                    double const q3 = __shfl_sync(0xffffffffu, rv[u], quad + 3);
                    if (oks[u])
                    {
                        size_t const i = static_cast<size_t>(nodes[u]) * 4 + rp;
                        double const zi = p01[u].x * q0 + p01[u].y * q1
                            + p23[u].x * q2 + p23[u].y * q3;
                        V.x[i] = xv[u]; V.r[i] = rv[u];
                        V.z[i] = zi;
                        acc[0] += rv[u] * rv[u];
                        acc[1] += xv[u] * (rv[u] - gv[u]);
                        acc[2] += zi * rv[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < CG_UF; ++u)
                {
                    nodes[u] = nodes_next[u];
                    oks[u] = oks_next[u];
                }
            }
            warp_flush<3>(acc, s_red, v);
        }
        publish<3>(a, s_state, s_red, slot + 1, false);
        unsigned long long const t_d = now_ns<TIMING>();
        grid_barrier(a.sync, epoch);
        all_sums<3>(a, s_state, slot + 1, s_bcast, false);
        if (TIMING)
        {
            tm[0] += t_b - t_a; tm[1] += t_c - t_b;
            tm[2] += t_d - t_c; tm[3] += now_ns<TIMING>() - t_d;
        }

        /* the reference's two stopping tests, per view (:139, :170-176) */
        if (threadIdx.x < a.n_views && !s_state[threadIdx.x].done)
        {
            CgState& S = s_state[threadIdx.x];
            double const new_rr = s_bcast[threadIdx.x * 3 + 0];
            double const xbr = s_bcast[threadIdx.x * 3 + 1];
            double const new_zr = s_bcast[threadIdx.x * 3 + 2];
            bool stop = false;
            if (new_rr < S.tol)
                stop = true;
            else
            {
                double const Q1 = -1.0 * xbr;
                double const zeta = iter * (Q1 - S.Q0) / Q1;
                if (zeta < a.q_tol)
                    stop = true;
                else
                {
                    S.Q0 = Q1;
                    S.beta = new_zr / S.r_dot_r;
                    S.r_dot_r = new_zr;
                }
            }
            if (stop)
            {
                S.done = 1;
                S.iters = iter;
                S.info = SMVSB_CG_CONVERGENCE;
            }
        }
        __syncthreads();
        bool all_done = true;
        for (int v = 0; v < a.n_views; ++v)
            all_done = all_done && (s_state[v].done != 0);
        if (all_done)
            break;
    }

    if (blockIdx.x == 0 && threadIdx.x < a.n_views)
    {
        CgState const& S = s_state[threadIdx.x];
        double* res = a.v[threadIdx.x].result;
        res[0] = S.done ? S.iters : iter;
        res[1] = S.info;
        /* lib/depth_optimizer.cc:267 looks at the first entry of the solution;
         * the final barrier has made every CTA's x visible */
        res[2] = isnan(__ldcg(a.v[threadIdx.x].x)) ? 1.0 : 0.0;
        for (int i = 0; i < 4; ++i)
            res[4 + i] = static_cast<double>(tm[i]);
    }
}

/* bit k of rowmask[node]: block k of the node's 3x3 stencil row exists, i.e.
 * the node and its k-th grid neighbour are both valid and active. */
__global__ void __launch_bounds__(256)
cg_mark_kernel (int npx, int npy, uint8_t const* __restrict__ node_valid,
    uint8_t const* __restrict__ active, uint16_t* __restrict__ rowmask,
    uint32_t* __restrict__ block_rows, unsigned long long* __restrict__ counts)
{
    int const node = blockIdx.x * blockDim.x + threadIdx.x;
    int const ns = npx + 1;
    unsigned int m = 0;
    if (node < ns * (npy + 1) && node_valid[node] && active[node])
    {
        int const ix = node % ns, iy = node / ns;
        for (int k = 0; k < 9; ++k)
        {
            int const jx = ix + (k % 3) - 1, jy = iy + (k / 3) - 1;
            if (jx < 0 || jx > npx || jy < 0 || jy > npy)
                continue;
            int const nj = jy * ns + jx;
            if (node_valid[nj] && active[nj])
                m |= 1u << k;
        }
    }
    if (node < ns * (npy + 1))
        rowmask[node] = static_cast<uint16_t>(m);
    /* counts[0]: blocks of the system, counts[1]: its block rows */
    unsigned int const blocks = __reduce_add_sync(0xffffffffu, __popc(m));
    int const rows = __syncthreads_count(m != 0);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(counts, static_cast<unsigned long long>(blocks));
    if (threadIdx.x == 0)
    {
        block_rows[blockIdx.x] = rows;
        atomicAdd(counts + 1, static_cast<unsigned long long>(rows));
    }
}

/* exclusive prefix sum of the per-block row counts, one block */
__global__ void __launch_bounds__(1024)
cg_scan_kernel (uint32_t const* __restrict__ block_rows,
    uint32_t* __restrict__ block_off, int nb)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024)
    {
        int const i = base + threadIdx.x;
        uint32_t const v = (i < nb) ? block_rows[i] : 0;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += u;
        }
        if (lane == 31)
            s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t const u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o)
                    w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t const before = s_carry + (warp > 0 ? s_warp[warp - 1] : 0)
            + inc - v;
        if (i < nb)
            block_off[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023)
            s_carry = before + v;
        __syncthreads();
    }
}

/* rows[]: the nodes with a non-empty row in ascending order -- the solver
 * walks this list, so the work is spread evenly over the CTAs however the
 * active set is scattered over the image */
__global__ void __launch_bounds__(256)
cg_list_kernel (int n_nodes, uint16_t const* __restrict__ rowmask,
    uint32_t const* __restrict__ block_off, uint32_t* __restrict__ rows)
{
    __shared__ uint32_t s_warp[8];
    int const node = blockIdx.x * blockDim.x + threadIdx.x;
    bool const on = node < n_nodes && rowmask[node] != 0;
    unsigned int const ballot = __ballot_sync(0xffffffffu, on);
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        s_warp[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = block_off[blockIdx.x];
    for (int w = 0; w < warp; ++w)
        before += s_warp[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (on)
        rows[before] = node;
}

__global__ void
spmv_kernel (int n_nodes, int npx, double const* __restrict__ H,
    uint16_t const* __restrict__ rowmask, double const* __restrict__ x,
    double* __restrict__ y)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes * 4)
        return;
    PlainVec vec;
    vec.v = x;
    double own[4];
    y[i] = spmv_row(H, npx + 1, vec, i >> 2, i & 3, rowmask[i >> 2], own);
}

/* The system's row masks, counts and compacted row list (three small kernels
 * per solve). */
void
mark_system (smvsb_ctx* c)
{
    int const nb = (c->n_nodes + 255) / 256;
    c->cg_rowmask.reserve(c->n_nodes);
    c->cg_row_list.reserve(c->n_nodes);
    c->cg_block_rows.reserve(2 * static_cast<size_t>(nb));
    c->cg_counts.reserve(2);
    CUDA_CHECK(cudaMemsetAsync(c->cg_counts.p, 0,
        2 * sizeof(unsigned long long), c->stream));
    cg_mark_kernel<<<nb, 256, 0, c->stream>>>(c->npx, c->npy,
        c->node_valid.p, c->active.p, c->cg_rowmask.p, c->cg_block_rows.p,
        c->cg_counts.p);
    cg_scan_kernel<<<1, 1024, 0, c->stream>>>(c->cg_block_rows.p,
        c->cg_block_rows.p + nb, nb);
    cg_list_kernel<<<nb, 256, 0, c->stream>>>(c->n_nodes, c->cg_rowmask.p,
        c->cg_block_rows.p + nb, c->cg_row_list.p);
    smvsb::count_launches(c, 3);
    CUDA_CHECK(cudaGetLastError());
}

} /* namespace */

void
launch_spmv (smvsb_ctx* c, double const* x, double* y)
{
    mark_system(c);
    int const n = c->n_nodes * 4;
    spmv_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->n_nodes, c->npx,
        c->H.p, c->cg_rowmask.p, x, y);
    smvsb::count_launches(c, 1);
    CUDA_CHECK(cudaGetLastError());
}

/*
 * Enqueues one PCG launch for the systems of `n` contexts (same device; all
 * work goes to the stream of cs[0], which the caller has made the stream of
 * every context of the batch) and the copies of the results into the
 * contexts' pinned scalars. cg_collect() reads them after the caller has
 * synchronised the stream.
 */
void
cg_enqueue (smvsb_ctx* const* cs, int n, int max_iter, double err_tol,
    double q_tol)
{
    if (n < 1 || n > SMVSB_MAX_BATCH)
        throw Error(SMVSB_ERR_INVALID, "batch size out of range");
    smvsb_ctx* lead = cs[0];
    bool const timing = getenv("SMVSB_CG_TIMING") != nullptr;
    void const* kernel = nullptr;
    if (n == 1)
        kernel = timing ? (void const*)cg_kernel<true, 1>
            : (void const*)cg_kernel<false, 1>;
    else if (n == 2)
        kernel = timing ? (void const*)cg_kernel<true, 2>
            : (void const*)cg_kernel<false, 2>;
    else if (n <= 4)
        kernel = timing ? (void const*)cg_kernel<true, 4>
            : (void const*)cg_kernel<false, 4>;
    else
        kernel = timing ? (void const*)cg_kernel<true, 8>
            : (void const*)cg_kernel<false, 8>;

    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
        kernel, CG_THREADS, 0));
    if (per_sm < 1)
        throw Error(SMVSB_ERR_CUDA, "cg_kernel does not fit on an SM");
    /* (Two views in flight on one GPU with 1 CTA/SM each, so that one view's
     * SpMV runs under the other's barriers and vector update: measured 202
     * against 224 Mpix-iters/s end to end with 2 CTAs/SM and the two launches
     * taking turns, job r2w.) */
    int const grid_max = std::min(lead->num_sms * std::min(per_sm, 2),
        CG_MAX_BLOCKS);

    CgArgs a;
    a.n_views = n; a.max_iter = max_iter; a.q_tol = q_tol;
    lead->cg_sync.reserve(1);
    a.sync = lead->cg_sync.p;
    int grid = 1;
    for (int k = 0; k < n; ++k)
    {
        smvsb_ctx* c = cs[k];
        size_t const nn = static_cast<size_t>(c->n_nodes) * 4;
        c->x.reserve(nn); c->r.reserve(nn); c->d.reserve(nn);
        c->d2.reserve(nn); c->Ad.reserve(nn); c->z.reserve(nn);
        c->cg_partials.reserve(static_cast<size_t>(CG_SLOTS) * CG_MAX_BLOCKS);
        c->cg_result.reserve(16);
        mark_system(c);
        CUDA_CHECK(cudaMemsetAsync(c->x.p, 0, nn * sizeof(double),
            c->stream));
        CgView& V = a.v[k];
        V.n_nodes = c->n_nodes; V.npx = c->npx; V.pad = 0;
        int const need = static_cast<int>((nn + CG_THREADS - 1) / CG_THREADS);
        V.grid = std::max(1, std::min(grid_max, need));
        grid = std::max(grid, V.grid);
        V.err_tol = err_tol;
        V.H = c->H.p; V.P = c->P.p; V.g = c->g.p;
        V.rowmask = c->cg_rowmask.p; V.rows = c->cg_row_list.p;
        V.counts = c->cg_counts.p;
        V.x = c->x.p; V.r = c->r.p; V.d = c->d.p; V.d2 = c->d2.p;
        V.Ad = c->Ad.p; V.z = c->z.p;
        V.partials = c->cg_partials.p; V.result = c->cg_result.p;
    }
    CUDA_CHECK(cudaMemsetAsync(a.sync, 0, sizeof(unsigned int), lead->stream));
    void* params[] = { &a };
    CUDA_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid),
        dim3(CG_THREADS), params, 0, lead->stream));
    smvsb::count_launches(lead, 1);
    CUDA_CHECK(cudaGetLastError());
    for (int k = 0; k < n; ++k)
    {
        smvsb_ctx* c = cs[k];
        c->cg_grid = grid;
        CUDA_CHECK(cudaMemcpyAsync(c->h_scalars, c->cg_result.p,
            8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 8, c->cg_counts.p,
            2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
            c->stream));
    }
}

void
cg_collect (smvsb_ctx* c, int* iters, int* info, bool* x0_nan)
{
    double const* res = c->h_scalars;
    unsigned long long counts[2];
    std::memcpy(counts, c->h_scalars + 8, sizeof(counts));
    if (getenv("SMVSB_CG_TIMING"))
        fprintf(stderr, "cg: iters %d grid %d | us/iter: spmv %.1f wait %.1f | "
            "update %.1f wait %.1f\n", (int)res[0], c->cg_grid,
            res[4] / 1e3 / res[0], res[5] / 1e3 / res[0], res[6] / 1e3 / res[0],
            res[7] / 1e3 / res[0]);
    c->cg_blocks = counts[0];
    c->cg_rows = counts[1];
    if (iters) *iters = static_cast<int>(res[0]);
    if (info) *info = static_cast<int>(res[1]);
    if (x0_nan) *x0_nan = (res[2] != 0.0);
}

void
run_cg (smvsb_ctx* c, int max_iter, double err_tol, double q_tol, int* iters,
    int* info, bool* x0_nan)
{
    smvsb_ctx* cs[1] = { c };
    cg_enqueue(cs, 1, max_iter, err_tol, q_tol);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    cg_collect(c, iters, info, x0_nan);
}

} /* namespace smvsb */
