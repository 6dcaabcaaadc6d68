/*
 * sgm.cu -- SGMStereo::run_sgm on the GPU (reference: lib/sgm_stereo.cc).
 *
 *   K6  sgm_cost_kernel     create_cost_volume (:192-244): plane sweep of the
 *                           neighbour luminance at num_steps inverse-depth
 *                           planes (warped_neighbors_for_depth :150-190, fp32,
 *                           byte bilinear with +0.5 rounding), 9x7 census
 *                           (:126-148) of main image and of every warped
 *                           slice, Hamming distance; 255 where the warped
 *                           pixel is 0. The reference materialises the warped
 *                           volume (uint8) and its census volume (uint64,
 *                           2.1 GB at 2 MP x 128); here a block keeps the
 *                           warped tile (with halo) of four planes in shared
 *                           memory, compares two pixels per integer add
 *                           (16-bit fields) and only the uint8 cost leaves
 *                           the SM.
 *   K7  sgm_paths_kernel    aggregate_sgm_costs (:429-667), SSE branch
 *                           (constant P2, uint16 arithmetic): one warp per
 *                           scan line of a direction, all 8 directions in one
 *                           launch; disparities across the lanes, L_r carried
 *                           in registers, min over disparities by warp
 *                           shuffles. Diagonals follow the line with
 *                           wrap-around at the image border, where the
 *                           reference restarts the path (:515-534). Each
 *                           direction writes L_r - C (a byte) to its own volume.
 *   K8  sgm_sum_wta_kernel  S = 8 C + sum_r (L_r - C) (+ the reference's
 *                           corner extras) and depth_from_sgm_volume
 *                           (:274-306) in one pass; S is only materialised
 *                           when the caller asks for the volume.
 *
 * Layouts: cost C[pixel][disp] uint8, sum S[pixel][disp] uint16 (pixel-major,
 * disparity contiguous, like the reference's sse_*_volume), so a warp's
 * access to one pixel is one coalesced 128 B / 256 B segment.
 */
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace smvsb {

namespace {

struct SgmParams
{
    int w, h, nw, nh, D;
    float M[9];
    float t[3];
};

/*
 * Cost volume, SWAR formulation. A thread owns two horizontally adjacent
 * pixels and works on four depth planes at a time.
 *
 *  - Pixels are kept as 16-bit fields, two per register. For centre pair A
 *    and neighbour pair B, R = B + 0x00FF00FF - A has bit 8 of each field
 *    set iff A < B (field value B - A + 255 in [0, 510], no borrow between
 *    the fields): ONE integer add per two census comparisons.
 *  - The Hamming distance of two census words does not depend on the bit
 *    order, so no 63-bit word is ever assembled: per offset the comparison
 *    bits of the warped slice are XORed with the main image's bits for the
 *    same offset (precomputed once per block into shared memory, reused for
 *    all planes) and accumulated in the fields: 3 instructions per 2
 *    comparisons instead of ~8.
 *  - The warped slice of a 32 x 16 pixel tile with its 9x7 halo is computed
 *    once per plane into shared memory by the whole block (fp32 steps
 *    restated with explicit round-to-nearest ops, so bit-identical to the
 *    CPU); M * (x, y, 1) does not depend on the plane and is kept in
 *    registers.
 * Semantics restated from census_filter / create_cost_volume
 * (lib/sgm_stereo.cc:126-148, 192-244): census only for pixels with
 * 4 <= x < w-5, 3 <= y < h-4 and a non-zero centre; cost 255 where the
 * warped pixel is 0.
 */
constexpr int CT_W = 32, CT_H = 16;              /* pixel tile per block */
constexpr int CT_THREADS = (CT_W / 2) * CT_H;    /* 256: one pixel pair each */
constexpr int HALO_W = CT_W + 8, HALO_H = CT_H + 6;
constexpr int HALO_N = HALO_W * HALO_H;          /* 880 */
constexpr int HALO_PER_THREAD = (HALO_N + CT_THREADS - 1) / CT_THREADS;
constexpr int PLANES = 4;                        /* planes per iteration */

__device__ __forceinline__ uint8_t
warp_from_tp (SgmParams const& p, uint8_t const* __restrict__ neigh,
    float const* tp, float depth)
{
    float q0 = __fadd_rn(__fmul_rn(tp[0], depth), p.t[0]);
    float q1 = __fadd_rn(__fmul_rn(tp[1], depth), p.t[1]);
    float const q2 = __fadd_rn(__fmul_rn(tp[2], depth), p.t[2]);
    if (q2 < 0)
        return 0;
    q0 = __fsub_rn(__fdiv_rn(q0, q2), 0.5f);
    q1 = __fsub_rn(__fdiv_rn(q1, q2), 0.5f);
    if (q0 < 0 || q1 < 0 || q0 > static_cast<float>(p.nw - 1)
        || q1 > static_cast<float>(p.nh - 1))
        return 0;
    /* mve::Image<uint8_t>::linear_at */
    float const xx = fmaxf(0.0f, fminf(static_cast<float>(p.nw - 1), q0));
    float const yy = fmaxf(0.0f, fminf(static_cast<float>(p.nh - 1), q1));
    int const fx = static_cast<int>(xx), fy = static_cast<int>(yy);
    int const fx1 = min(fx + 1, p.nw - 1), fy1 = min(fy + 1, p.nh - 1);
    float const w1 = __fsub_rn(xx, static_cast<float>(fx));
    float const w0 = __fsub_rn(1.0f, w1);
    float const w3 = __fsub_rn(yy, static_cast<float>(fy));
    float const w2 = __fsub_rn(1.0f, w3);
    float const v00 = neigh[fy * p.nw + fx], v10 = neigh[fy * p.nw + fx1];
    float const v01 = neigh[fy1 * p.nw + fx], v11 = neigh[fy1 * p.nw + fx1];
    float s = __fmul_rn(v00, __fmul_rn(w0, w2));
    s = __fadd_rn(s, __fmul_rn(v10, __fmul_rn(w1, w2)));
    s = __fadd_rn(s, __fmul_rn(v01, __fmul_rn(w0, w3)));
    s = __fadd_rn(s, __fmul_rn(v11, __fmul_rn(w1, w3)));
    s = __fadd_rn(s, 0.5f);
    return static_cast<uint8_t>(s);
}

/* Pair of 16-bit fields at element offset e (0..8) of the five words
 * wd[0..4] that hold elements 0..9 of a tile row. */
#define SMVSB_WINDOW(wd, e) (((e) & 1) ? __funnelshift_r((wd)[(e) >> 1],   \
    (wd)[((e) >> 1) + 1], 16) : (wd)[(e) >> 1])

__global__ void __launch_bounds__(CT_THREADS)
sgm_cost_kernel (SgmParams const p, uint8_t const* __restrict__ main_img,
    uint8_t const* __restrict__ neigh, float const* __restrict__ depths,
    uint8_t* __restrict__ cost)
{
    /* tile of 16-bit pixels, HALO_W even; as words: HALO_W / 2 per row */
    __shared__ unsigned s_tile[PLANES][HALO_H][HALO_W / 2];
    /* main comparison bits, [63][CT_THREADS] words = 63 KB: dynamic */
    extern __shared__ unsigned s_mask_dyn[];
    unsigned (*s_mask)[CT_THREADS] =
        reinterpret_cast<unsigned (*)[CT_THREADS]>(s_mask_dyn);
    __shared__ float s_depths[256];

    int const tid = threadIdx.x;
    int const tx = tid % (CT_W / 2), ty = tid / (CT_W / 2);
    int const x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    int const px = x0 + 2 * tx, py = y0 + ty;        /* left pixel of the pair */

    for (int i = tid; i < p.D; i += CT_THREADS)
        s_depths[i] = depths[i];

    /* plane-independent part of the warp for this thread's halo pixels */
    float tp[HALO_PER_THREAD][3];
    bool in_img[HALO_PER_THREAD];
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k)
    {
        int const i = tid + k * CT_THREADS;
        int const gx = x0 - 4 + i % HALO_W, gy = y0 - 3 + i / HALO_W;
        in_img[k] = (i < HALO_N && gx >= 0 && gx < p.w && gy >= 0 && gy < p.h);
        float const fx = 0.5f + static_cast<float>(gx);
        float const fy = 0.5f + static_cast<float>(gy);
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            float s = __fmul_rn(p.M[3 * r], fx);
            s = __fadd_rn(s, __fmul_rn(p.M[3 * r + 1], fy));
            tp[k][r] = __fadd_rn(s, p.M[3 * r + 2]);      /* * 1.f */
        }
    }

    /* main image tile -> comparison bits per offset */
    unsigned short* tile16 = reinterpret_cast<unsigned short*>(&s_tile[0][0][0]);
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k)
    {
        int const i = tid + k * CT_THREADS;
        if (i < HALO_N)
        {
            int const gx = x0 - 4 + i % HALO_W, gy = y0 - 3 + i / HALO_W;
            tile16[i] = in_img[k] ? main_img[gy * p.w + gx] : 0;
        }
    }
    __syncthreads();
    bool const in0 = (px < p.w && py < p.h), in1 = (px + 1 < p.w && py < p.h);
    bool const int0 = in0 && px >= 4 && px < p.w - 5 && py >= 3 && py < p.h - 4;
    bool const int1 = in1 && px + 1 >= 4 && px + 1 < p.w - 5 && py >= 3
        && py < p.h - 4;
    {
        unsigned const A = s_tile[0][ty + 3][tx + 2];
        /* fields of pixels without a census (border, zero centre) stay 0 */
        unsigned keep = 0;
        if (int0 && (A & 0xffffu) != 0) keep |= 0x00000100u;
        if (int1 && (A >> 16) != 0) keep |= 0x01000000u;
#pragma unroll
        for (int j = 0; j < 7; ++j)
        {
            unsigned wd[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) wd[q] = s_tile[0][ty + j][tx + q];
#pragma unroll
            for (int e = 0; e < 9; ++e)
            {
                unsigned const B = SMVSB_WINDOW(wd, e);
                s_mask[j * 9 + e][tid] = (B + 0x00FF00FFu - A) & keep;
            }
        }
    }

    for (int d0 = 0; d0 < p.D; d0 += PLANES)
    {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < HALO_PER_THREAD; ++k)
        {
            int const i = tid + k * CT_THREADS;
            if (i < HALO_N)
            {
#pragma unroll
                for (int pl = 0; pl < PLANES; ++pl)
                {
                    unsigned short v = 0;
                    if (in_img[k])
                        v = warp_from_tp(p, neigh, tp[k], s_depths[d0 + pl]);
                    reinterpret_cast<unsigned short*>(
                        &s_tile[pl][0][0])[i] = v;
                }
            }
        }
        __syncthreads();

        unsigned A[PLANES], acc[PLANES];
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl)
        {
            A[pl] = s_tile[pl][ty + 3][tx + 2];
            acc[pl] = 0;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j)
        {
            unsigned wd[PLANES][5];
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl)
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    wd[pl][q] = s_tile[pl][ty + j][tx + q];
#pragma unroll
            for (int e = 0; e < 9; ++e)
            {
                unsigned const mm = s_mask[j * 9 + e][tid];
#pragma unroll
                for (int pl = 0; pl < PLANES; ++pl)
                {
                    unsigned const B = SMVSB_WINDOW(wd[pl], e);
                    acc[pl] += ((B + 0x00FF00FFu - A[pl]) ^ mm) & 0x01000100u;
                }
            }
        }
        /* cost bytes of the four planes for each of the two pixels */
        unsigned out0 = 0, out1 = 0;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl)
        {
            unsigned const w0v = A[pl] & 0xffffu, w1v = A[pl] >> 16;
            /* border pixels have no census on either side: distance 0 */
            unsigned c0 = int0 ? ((acc[pl] >> 8) & 0xffu) : 0u;
            unsigned c1 = int1 ? (acc[pl] >> 24) : 0u;
            if (w0v == 0) c0 = 255u;
            if (w1v == 0) c1 = 255u;
            out0 |= c0 << (8 * pl);
            out1 |= c1 << (8 * pl);
        }
        if (in0)
            *reinterpret_cast<unsigned*>(cost + (static_cast<size_t>(py) * p.w
                + px) * p.D + d0) = out0;
        if (in1)
            *reinterpret_cast<unsigned*>(cost + (static_cast<size_t>(py) * p.w
                + px + 1) * p.D + d0) = out1;
    }
}
#undef SMVSB_WINDOW

/* ------------------------------------------------------------------ */

enum PathKind
{
    PATH_L2R = 0, PATH_R2L, PATH_T2B, PATH_T2B_D1, PATH_T2B_D2,
    PATH_B2T, PATH_B2T_D1, PATH_B2T_D2
};

/*
 * All eight path directions in ONE launch: warp -> (direction, scan line),
 * 2h + 6w warps in flight (13 680 at 1920x1080) instead of h or w per
 * sequential launch. DPL = disparities per lane (D = 32 * DPL).
 * fill_path_cost_sse (:361-406):
 *   L(p,i) = C(p,i) + min(L(q,i), L(q,i-1)+P1, L(q,i+1)+P1, min_k L(q,k)+P2)
 *            - min_k L(q,k)            (all uint16, wrap-around)
 * and copy_cost_and_add_to_sgm (:408-426) where a path starts (L = C).
 * The directions cannot share one read-modify-write sum volume without
 * racing, so each writes its own byte volume of L - C, which lies in [0, P2]
 * (P2 <= 255): 1 B/voxel/direction. sgm_sum_wta_kernel adds them up.
 */
template <int DPL>
__global__ void __launch_bounds__(128)
sgm_paths_kernel (int w, int h, unsigned P1, unsigned P2,
    uint8_t const* __restrict__ cost, uint8_t* __restrict__ Dvol)
{
    int const D = 32 * DPL;
    int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int const lane = threadIdx.x & 31;
    int kind, line;
    if (gw < 2 * h)
    {
        kind = gw / h;                  /* PATH_L2R, PATH_R2L */
        line = gw % h;
    }
    else
    {
        gw -= 2 * h;
        if (gw >= 6 * w)
            return;
        kind = 2 + gw / w;              /* PATH_T2B .. PATH_B2T_D2 */
        line = gw % w;
    }
    bool const horizontal = (kind < 2);
    int const steps = horizontal ? w : h;
    size_t const nvox = static_cast<size_t>(w) * h * D;
    uint8_t* __restrict__ Dr = Dvol + static_cast<size_t>(kind) * nvox;

    /* start pixel and per-step increments; diagonals wrap around in x,
     * where the reference restarts the path (:515-534) */
    int x, y, dx, dy;
    switch (kind)
    {
    case PATH_L2R: x = 0; y = line; dx = 1; dy = 0; break;
    case PATH_R2L: x = w - 1; y = line; dx = -1; dy = 0; break;
    case PATH_T2B: x = line; y = 0; dx = 0; dy = 1; break;
    case PATH_T2B_D1: x = line; y = 0; dx = 1; dy = 1; break;
    case PATH_T2B_D2: x = line; y = 0; dx = -1; dy = 1; break;
    case PATH_B2T: x = line; y = h - 1; dx = 0; dy = -1; break;
    case PATH_B2T_D1: x = line; y = h - 1; dx = 1; dy = -1; break;
    default: x = line; y = h - 1; dx = -1; dy = -1; break;   /* B2T_D2 */
    }
    int const restart_x = (dx > 0) ? 0 : w - 1;   /* diagonals only */
    bool const diagonal = (!horizontal && dx != 0);

    auto load_cost = [&](size_t base, unsigned* C)
    {
        if (DPL == 4)
        {
            uchar4 const c4 = *reinterpret_cast<uchar4 const*>(cost + base);
            C[0] = c4.x; C[1] = c4.y; C[2] = c4.z; C[3] = c4.w;
        }
        else
        {
#pragma unroll
            for (int i = 0; i < DPL; ++i) C[i] = cost[base + i];
        }
    };

    if (DPL == 4)
    {
        /* Fast path for 128 planes: the four disparities of a lane live in
         * two registers as 16-bit pairs; the minima are DPX instructions
         * (VIMNMX3 / VIADDMNMX on 16x2). L <= C + P2 <= 510 never overflows a
         * half, so packed adds / subtracts are plain 32-bit ones. */
        unsigned const P1x2 = P1 | (P1 << 16), P2x2 = P2 | (P2 << 16);
        unsigned const BIG = 0x7000u;      /* "no neighbour" sentinel */
        unsigned P01 = 0, P23 = 0;
        size_t base = (static_cast<size_t>(y) * w + x) * D + lane * 4;
        unsigned c4 = *reinterpret_cast<unsigned const*>(cost + base);
        bool start = true;
        for (int s = 0; s < steps; ++s)
        {
            int xn = x + dx, yn = y + dy;
            if (xn < 0) xn = w - 1;
            if (xn >= w) xn = 0;
            bool const startn = diagonal && (xn == restart_x);
            size_t const basen = (static_cast<size_t>(yn) * w + xn) * D
                + lane * 4;
            unsigned c4n = 0;
            if (s + 1 < steps)
                c4n = *reinterpret_cast<unsigned const*>(cost + basen);

            unsigned const C01 = __byte_perm(c4, 0, 0x4140);
            unsigned const C23 = __byte_perm(c4, 0, 0x4342);
            unsigned D01 = 0, D23 = 0;
            if (start)
            {
                P01 = C01; P23 = C23;
            }
            else
            {
                unsigned const m2 = __vminu2(P01, P23);
                unsigned mn = min(m2 & 0xffffu, m2 >> 16);
                for (int off = 16; off > 0; off >>= 1)
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                unsigned below = __shfl_up_sync(0xffffffffu, P23, 1) >> 16;
                unsigned above = __shfl_down_sync(0xffffffffu, P01, 1)
                    & 0xffffu;
                if (lane == 0) below = BIG;
                if (lane == 31) above = BIG;
                unsigned const mid = (P01 >> 16) | (P23 << 16);   /* L1, L2 */
                unsigned const lo01 = below | (P01 << 16);        /* -, L0 */
                unsigned const hi23 = (P23 >> 16) | (above << 16);/* L3, - */
                unsigned const mn2 = mn * 0x10001u;
                unsigned const far2 = mn2 + P2x2;
                unsigned const b01 = __vimin3_u16x2(P01, far2,
                    __viaddmin_u16x2(mid, P1x2, lo01 + P1x2));
                unsigned const b23 = __vimin3_u16x2(P23, far2,
                    __viaddmin_u16x2(hi23, P1x2, mid + P1x2));
                D01 = b01 - mn2;          /* = L - C, in [0, P2] per half */
                D23 = b23 - mn2;
                P01 = C01 + D01;
                P23 = C23 + D23;
            }
            *reinterpret_cast<unsigned*>(Dr + base) =
                __byte_perm(D01, D23, 0x6420);
            c4 = c4n;
            x = xn; y = yn; base = basen; start = startn;
        }
        return;
    }

    unsigned Lp[DPL], C[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) Lp[i] = 0;
    size_t base = (static_cast<size_t>(y) * w + x) * D + lane * DPL;
    load_cost(base, C);
    bool start = true;

    for (int s = 0; s < steps; ++s)
    {
        /* next pixel: its cost is fetched while this one is computed */
        int xn = x + dx, yn = y + dy;
        if (xn < 0) xn = w - 1;
        if (xn >= w) xn = 0;
        bool const startn = diagonal && (xn == restart_x);
        size_t const basen = (static_cast<size_t>(yn) * w + xn) * D
            + lane * DPL;
        unsigned Cn[DPL];
        if (s + 1 < steps)
            load_cost(basen, Cn);

        unsigned Dv[DPL];
        if (start)
        {
#pragma unroll
            for (int i = 0; i < DPL; ++i) { Lp[i] = C[i]; Dv[i] = 0; }
        }
        else
        {
            unsigned mn = Lp[0];
#pragma unroll
            for (int i = 1; i < DPL; ++i) mn = min(mn, Lp[i]);
            for (int off = 16; off > 0; off >>= 1)
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
            unsigned const below = __shfl_up_sync(0xffffffffu, Lp[DPL - 1], 1);
            unsigned const above = __shfl_down_sync(0xffffffffu, Lp[0], 1);
            unsigned const far = (mn + P2) & 0xffffu;
            unsigned Ln[DPL];
#pragma unroll
            for (int i = 0; i < DPL; ++i)
            {
                unsigned best = min(Lp[i], far);
                bool const has_lo = (i > 0) || (lane > 0);
                bool const has_hi = (i < DPL - 1) || (lane < 31);
                unsigned const lo = (i > 0) ? Lp[i - 1] : below;
                unsigned const hi = (i < DPL - 1) ? Lp[i + 1] : above;
                if (has_lo) best = min(best, (lo + P1) & 0xffffu);
                if (has_hi) best = min(best, (hi + P1) & 0xffffu);
                Dv[i] = (best - mn) & 0xffffu;           /* = L - C, <= P2 */
                Ln[i] = (C[i] + Dv[i]) & 0xffffu;
            }
#pragma unroll
            for (int i = 0; i < DPL; ++i) Lp[i] = Ln[i];
        }
        if (DPL == 4)
            *reinterpret_cast<uchar4*>(Dr + base) = make_uchar4(
                (unsigned char)Dv[0], (unsigned char)Dv[1],
                (unsigned char)Dv[2], (unsigned char)Dv[3]);
        else
        {
#pragma unroll
            for (int i = 0; i < DPL; ++i)
                Dr[base + i] = static_cast<uint8_t>(Dv[i]);
        }
#pragma unroll
        for (int i = 0; i < DPL; ++i) C[i] = Cn[i];
        x = xn; y = yn; base = basen; start = startn;
    }
}

/*
 * S(p,i) = sum over the 8 directions of L_r(p,i) = 8 C + sum_r (L_r - C),
 * plus C once more at the four image corners: column 0 of the d1 volume and
 * column w-1 of the d2 volume are (re)initialised for ALL y after row 0 /
 * row h-1 already were (:521-534, :600-613). Then depth_from_sgm_volume
 * (:274-306): first minimum over the planes. One warp per pixel.
 */
template <int DPL>
__global__ void
sgm_sum_wta_kernel (int w, int h, uint8_t const* __restrict__ cost,
    uint8_t const* __restrict__ Dvol, uint8_t const* __restrict__ main_img,
    float const* __restrict__ depths, uint16_t* __restrict__ S_out,
    float* __restrict__ out)
{
    int const D = 32 * DPL;
    int const npix = w * h;
    int const p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int const lane = threadIdx.x & 31;
    if (p >= npix)
        return;
    size_t const nvox = static_cast<size_t>(npix) * D;
    size_t const base = static_cast<size_t>(p) * D + lane * DPL;
    int const px = p % w, py = p / w;
    unsigned const mult = 8u + (((px == 0 || px == w - 1)
        && (py == 0 || py == h - 1)) ? 1u : 0u);
    unsigned Sv[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i)
        Sv[i] = (mult * cost[base + i]) & 0xffffu;
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        uint8_t const* Dr = Dvol + r * nvox + base;
#pragma unroll
        for (int i = 0; i < DPL; ++i)
            Sv[i] = (Sv[i] + Dr[i]) & 0xffffu;
    }
    if (S_out != nullptr)
    {
#pragma unroll
        for (int i = 0; i < DPL; ++i)
            S_out[base + i] = static_cast<uint16_t>(Sv[i]);
    }
    unsigned best = 0xffffu;    /* numeric_limits<uint16_t>::max() */
    int best_i = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < DPL; ++i)
        if (Sv[i] < best)
        {
            best = Sv[i];
            best_i = lane * DPL + i;
            found = true;
        }
    /* key = value << 16 | index: min picks the lowest value, then index */
    unsigned key = found ? ((best << 16) | best_i) : 0xffffffffu;
    for (int off = 16; off > 0; off >>= 1)
        key = min(key, __shfl_xor_sync(0xffffffffu, key, off));
    if (lane == 0)
    {
        int const idx = (key == 0xffffffffu) ? 0 : static_cast<int>(
            key & 0xffffu);
        out[p] = (idx < 2 || main_img[p] < 25) ? 0.0f : depths[idx];
    }
}

__global__ void
u8_to_u16_kernel (size_t n, uint8_t const* __restrict__ in,
    uint16_t* __restrict__ out)
{
    size_t const i = static_cast<size_t>(blockIdx.x) * blockDim.x
        + threadIdx.x;
    if (i < n)
        out[i] = in[i];
}

template <int DPL>
void
run_paths (int w, int h, unsigned P1, unsigned P2, uint8_t const* cost,
    uint8_t* Dvol, cudaStream_t st)
{
    int const warps = 2 * h + 6 * w;
    sgm_paths_kernel<DPL><<<(warps * 32 + 127) / 128, 128, 0, st>>>(w, h, P1,
        P2, cost, Dvol);
    CUDA_CHECK(cudaGetLastError());
}

template <int DPL>
void
run_wta (int w, int h, uint8_t const* cost, uint8_t const* Dvol,
    uint8_t const* main_img, float const* depths, uint16_t* S_out, float* out,
    cudaStream_t st)
{
    int const blocks = (w * h * 32 + 255) / 256;
    sgm_sum_wta_kernel<DPL><<<blocks, 256, 0, st>>>(w, h, cost, Dvol, main_img,
        depths, S_out, out);
    CUDA_CHECK(cudaGetLastError());
}

thread_local std::string g_sgm_error;

} /* namespace */

std::string const&
sgm_last_error (void)
{
    return g_sgm_error;
}

int
sgm_run (int device, int w, int h, uint8_t const* main_lum, int nw, int nh,
    uint8_t const* neigh_lum, float const* M, float const* t,
    float min_depth, float max_depth, int num_steps, uint16_t penalty1,
    uint16_t penalty2, float* depth_out, uint16_t* cost_out,
    uint16_t* sgm_out, double* ms_out)
{
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    int rc = SMVSB_OK;
    try
    {
        if (!(w > 9 && h > 7 && nw > 1 && nh > 1 && main_lum && neigh_lum
            && M && t && depth_out))
            throw Error(SMVSB_ERR_INVALID, "smvsb_sgm: bad image arguments");
        if (num_steps < 32 || num_steps > 256 || num_steps % 32 != 0
            || (num_steps / 32 != 1 && num_steps / 32 != 2
                && num_steps / 32 != 4 && num_steps / 32 != 8))
            throw Error(SMVSB_ERR_INVALID,
                "smvsb_sgm: num_steps must be 32, 64, 128 or 256");
        /* the O(D) recurrence equals the reference's O(D^2) minimum only
         * for P1 <= P2 and without uint16 wrap-around */
        if (penalty1 > penalty2 || penalty2 > 255)
            throw Error(SMVSB_ERR_INVALID,
                "smvsb_sgm: need penalty1 <= penalty2 <= 255");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(SMVSB_ERR_CUDA, "no CUDA device (no CPU fallback)");
        if (device < 0 || device >= count)
            throw Error(SMVSB_ERR_INVALID, "device index out of range");
        CUDA_CHECK(cudaSetDevice(device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i)
            CUDA_CHECK(cudaEventCreate(&ev[i]));

        /* plane depths, lib/sgm_stereo.cc:195-203 (fp32 recurrence) */
        std::vector<float> depths(num_steps);
        {
            float inv_depth = 1.0f / max_depth;
            float const increment = (1.0f / min_depth - inv_depth)
                / (num_steps - 1);
            for (int i = 0; i < num_steps; ++i)
            {
                depths[i] = 1.0f / inv_depth;
                inv_depth += increment;
            }
        }

        size_t const npix = static_cast<size_t>(w) * h;
        size_t const nvox = npix * num_steps;
        DevBuf<uint8_t> d_main, d_neigh, d_cost, d_D;
        DevBuf<uint16_t> d_S;
        DevBuf<float> d_depths, d_out;
        d_main.reserve(npix);
        d_neigh.reserve(static_cast<size_t>(nw) * nh);
        d_cost.reserve(nvox);
        d_D.reserve(nvox * 8);                 /* L - C per direction */
        if (sgm_out || cost_out)
            d_S.reserve(nvox);
        d_depths.reserve(num_steps);
        d_out.reserve(npix);
        CUDA_CHECK(cudaMemcpyAsync(d_main.p, main_lum, npix,
            cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_neigh.p, neigh_lum,
            static_cast<size_t>(nw) * nh, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_depths.p, depths.data(),
            num_steps * sizeof(float), cudaMemcpyHostToDevice, st));

        SgmParams p;
        p.w = w; p.h = h; p.nw = nw; p.nh = nh; p.D = num_steps;
        std::copy(M, M + 9, p.M);
        std::copy(t, t + 3, p.t);

        CUDA_CHECK(cudaEventRecord(ev[0], st));
        dim3 const cb(CT_THREADS);
        dim3 const cg((w + CT_W - 1) / CT_W, (h + CT_H - 1) / CT_H);
        size_t const mask_bytes = 63 * CT_THREADS * sizeof(unsigned);
        CUDA_CHECK(cudaFuncSetAttribute(sgm_cost_kernel,
            cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(mask_bytes)));
        sgm_cost_kernel<<<cg, cb, mask_bytes, st>>>(p, d_main.p, d_neigh.p, d_depths.p,
            d_cost.p);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaEventRecord(ev[1], st));

        int const dpl = num_steps / 32;
        switch (dpl)
        {
        case 1: run_paths<1>(w, h, penalty1, penalty2, d_cost.p, d_D.p, st); break;
        case 2: run_paths<2>(w, h, penalty1, penalty2, d_cost.p, d_D.p, st); break;
        case 4: run_paths<4>(w, h, penalty1, penalty2, d_cost.p, d_D.p, st); break;
        default: run_paths<8>(w, h, penalty1, penalty2, d_cost.p, d_D.p, st); break;
        }
        CUDA_CHECK(cudaEventRecord(ev[2], st));
        uint16_t* const S_dev = sgm_out ? d_S.p : nullptr;
        switch (dpl)
        {
        case 1: run_wta<1>(w, h, d_cost.p, d_D.p, d_main.p, d_depths.p, S_dev, d_out.p, st); break;
        case 2: run_wta<2>(w, h, d_cost.p, d_D.p, d_main.p, d_depths.p, S_dev, d_out.p, st); break;
        case 4: run_wta<4>(w, h, d_cost.p, d_D.p, d_main.p, d_depths.p, S_dev, d_out.p, st); break;
        default: run_wta<8>(w, h, d_cost.p, d_D.p, d_main.p, d_depths.p, S_dev, d_out.p, st); break;
        }
        CUDA_CHECK(cudaEventRecord(ev[3], st));

        CUDA_CHECK(cudaMemcpyAsync(depth_out, d_out.p, npix * sizeof(float),
            cudaMemcpyDeviceToHost, st));
        if (sgm_out)
            CUDA_CHECK(cudaMemcpyAsync(sgm_out, d_S.p,
                nvox * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (cost_out)
        {
            /* widen through the (now free) S buffer */
            u8_to_u16_kernel<<<static_cast<unsigned>((nvox + 255) / 256), 256,
                0, st>>>(nvox, d_cost.p, d_S.p);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaMemcpyAsync(cost_out, d_S.p,
                nvox * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
        }
        if (ms_out)
        {
            float ms;
            for (int i = 0; i < 3; ++i)
            {
                CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
                ms_out[i] = ms;
            }
        }
    }
    catch (Error const& e)
    {
        g_sgm_error = e.msg;
        rc = e.code;
    }
    for (int i = 0; i < 4; ++i)
        if (ev[i]) cudaEventDestroy(ev[i]);
    if (st) cudaStreamDestroy(st);
    return rc;
}

} /* namespace smvsb */
