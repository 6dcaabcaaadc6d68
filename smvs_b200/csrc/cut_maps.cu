/*
 * cut_maps.cu -- MeshGenerator::cut_depth_maps (lib/mesh_generator.cc:25-158)
 * on the device: the cross-view consistency cut that follows the per-view
 * optimisation ("next" row 4). For every pixel of every
 * depth map the 3-D point is projected into all other views; the depth is
 * dropped when it faces away from its own camera, when a view that sees the
 * point much better disagrees, or when the views that agree with it do not
 * outweigh those it occludes.
 *
 * One thread per pixel, the loop over the other views inside the thread; all
 * maps of the scene stay resident (41 MB per 2 MP view). Everything is fp32
 * with the reference's operation order (math::Vector / Matrix operators,
 * mve::geom::pixel_3dpos, ViewProjection::get_proj / get_surface_power,
 * :302-344) and no contraction, the three comparisons the reference makes in
 * double (:118, :121, :123-128) are made in double: the decisions are the
 * CPU's. The camera matrices come from the host (the reference computes them
 * with MVE's CameraInfo; the kernel only consumes them).
 */
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace smvsb {

namespace {

struct CutView
{
    int w, h;
    float const* cut;        /* depth as given (MVE convention: along the ray) */
    float const* zdepth;     /* after depthmap_convert_conventions(.., false) */
    float const* normals;    /* world space, w*h*3 */
    float* out;
    float invproj[9], ctw[16], KR[9], t[3];
};

struct f3
{
    float x, y, z;
};

/* std::inner_product(a, a + 3, b, 0.f) */
__device__ __forceinline__ float
dot3 (float const* a, f3 const& b)
{
    float s = __fadd_rn(0.0f, __fmul_rn(a[0], b.x));
    s = __fadd_rn(s, __fmul_rn(a[1], b.y));
    return __fadd_rn(s, __fmul_rn(a[2], b.z));
}

__device__ __forceinline__ float
dot3 (f3 const& a, f3 const& b)
{
    float s = __fadd_rn(0.0f, __fmul_rn(a.x, b.x));
    s = __fadd_rn(s, __fmul_rn(a.y, b.y));
    return __fadd_rn(s, __fmul_rn(a.z, b.z));
}

/* mve::geom::pixel_3dpos, then Matrix4f::mult(pos, 1.0f) with cam-to-world */
__device__ __forceinline__ f3
world_pos (CutView const& v, int x, int y, float depth)
{
    f3 const px = { __fadd_rn(static_cast<float>(x), 0.5f),
        __fadd_rn(static_cast<float>(y), 0.5f), 1.0f };
    f3 ray = { dot3(v.invproj, px), dot3(v.invproj + 3, px),
        dot3(v.invproj + 6, px) };
    float const n = __fsqrt_rn(dot3(ray, ray));
    ray.x = __fmul_rn(__fdiv_rn(ray.x, n), depth);
    ray.y = __fmul_rn(__fdiv_rn(ray.y, n), depth);
    ray.z = __fmul_rn(__fdiv_rn(ray.z, n), depth);
    f3 out;
    out.x = __fadd_rn(dot3(v.ctw, ray), v.ctw[3]);
    out.y = __fadd_rn(dot3(v.ctw + 4, ray), v.ctw[7]);
    out.z = __fadd_rn(dot3(v.ctw + 8, ray), v.ctw[11]);
    return out;
}

/* ViewProjection::get_proj, :314-321 */
__device__ __forceinline__ f3
get_proj (CutView const& v, f3 const& pos)
{
    f3 out;
    out.x = __fsub_rn(dot3(v.KR, pos), v.t[0]);
    out.y = __fsub_rn(dot3(v.KR + 3, pos), v.t[1]);
    out.z = __fsub_rn(dot3(v.KR + 6, pos), v.t[2]);
    return out;
}

/* ViewProjection::get_surface_power, :323-344 */
__device__ __forceinline__ float
surface_power (CutView const& v, f3 const& pos, f3 const& normal)
{
    f3 const p = get_proj(v, pos);
    float const denom = __fmul_rn(p.z, p.z);
    float ud[3], vd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        ud[k] = __fdiv_rn(__fsub_rn(__fmul_rn(v.KR[k], p.z),
            __fmul_rn(v.KR[6 + k], p.x)), denom);
        vd[k] = __fdiv_rn(__fsub_rn(__fmul_rn(v.KR[3 + k], p.z),
            __fmul_rn(v.KR[6 + k], p.y)), denom);
    }
    f3 c;
    c.x = __fsub_rn(__fmul_rn(ud[1], vd[2]), __fmul_rn(ud[2], vd[1]));
    c.y = __fsub_rn(__fmul_rn(ud[2], vd[0]), __fmul_rn(ud[0], vd[2]));
    c.z = __fsub_rn(__fmul_rn(ud[0], vd[1]), __fmul_rn(ud[1], vd[0]));
    return -dot3(normal, c);
}

/* mve::image::depthmap_convert_conventions<float>(dm, invproj, false) */
__global__ void
to_zdepth_kernel (int w, int h, float const* __restrict__ in, float i0,
    float i1, float i2, float i3, float i4, float i5, float i6, float i7,
    float i8, float* __restrict__ out)
{
    int const x = blockIdx.x * blockDim.x + threadIdx.x;
    int const y = blockIdx.y;
    if (x >= w)
        return;
    float const m[9] = { i0, i1, i2, i3, i4, i5, i6, i7, i8 };
    f3 const px = { __fadd_rn(static_cast<float>(x), 0.5f),
        __fadd_rn(static_cast<float>(y), 0.5f), 1.0f };
    f3 const ray = { dot3(m, px), dot3(m + 3, px), dot3(m + 6, px) };
    double const len = static_cast<double>(__fsqrt_rn(dot3(ray, ray)));
    size_t const i = static_cast<size_t>(y) * w + x;
    out[i] = static_cast<float>(__dmul_rn(static_cast<double>(in[i]),
        __ddiv_rn(1.0, len)));
}

__global__ void __launch_bounds__(128)
cut_depth_maps_kernel (CutView const* __restrict__ views, int n_views, int vi)
{
    CutView const& V = views[vi];
    int const x = blockIdx.x * blockDim.x + threadIdx.x;
    int const y = blockIdx.y;
    if (x >= V.w)
        return;
    size_t const pix = static_cast<size_t>(y) * V.w + x;
    float const d = V.cut[pix];
    float result = d;
    if (d == 0.0f)
    {
        V.out[pix] = 0.0f;
        return;
    }
    f3 const pos = world_pos(V, x, y, d);
    f3 const normal = { V.normals[3 * pix], V.normals[3 * pix + 1],
        V.normals[3 * pix + 2] };
    float const sp = surface_power(V, pos, normal);
    if (sp < 0.0f)
        result = 0.0f;
    float consistency = 0.0f;
    bool cut_hard = false;
    for (int j = 0; j < n_views; ++j)
    {
        if (j == vi)
            continue;
        CutView const& J = views[j];
        f3 const proj = get_proj(J, pos);
        if (proj.z < 0.0f)
            continue;
        int const xj = static_cast<int>(__fdiv_rn(proj.x, proj.z));
        int const yj = static_cast<int>(__fdiv_rn(proj.y, proj.z));
        if (xj < 0 || xj >= J.w || yj < 0 || yj >= J.h)
            continue;
        size_t const pj = static_cast<size_t>(yj) * J.w + xj;
        float const dm_j = J.zdepth[pj];
        if (dm_j == 0.0f)
            continue;
        float const sp_j = surface_power(J, pos, normal);
        f3 const pos_j = world_pos(J, xj, yj, J.cut[pj]);
        f3 const normal_j = { J.normals[3 * pj], J.normals[3 * pj + 1],
            J.normals[3 * pj + 2] };
        float const sp_jj = surface_power(J, pos_j, normal_j);
        double const z = static_cast<double>(proj.z);
        if (__dmul_rn(static_cast<double>(dm_j), 1.01) < z)
            continue;
        if (__dmul_rn(static_cast<double>(dm_j), 0.997) > z)
        {
            if (static_cast<double>(sp_jj) > __dmul_rn(0.5,
                static_cast<double>(sp)))
                consistency = __fsub_rn(consistency, sp_jj);
            continue;
        }
        double const twice = __dmul_rn(2.0, static_cast<double>(sp));
        if (static_cast<double>(sp_jj) > twice
            || static_cast<double>(sp_j) > twice)
        {
            cut_hard = true;
            break;
        }
        consistency = __fadd_rn(consistency, sp_jj);
    }
    if (cut_hard || consistency <= 0.0f)
        result = 0.0f;
    V.out[pix] = result;
}

std::mutex g_cut_lock;
thread_local std::string g_cut_error;

} /* namespace */

std::string const&
cut_last_error (void)
{
    return g_cut_error;
}

int
cut_depth_maps (int device, int n_views, int const* w, int const* h,
    float const* const* depth, float const* const* normals,
    float const* invproj9, float const* cam_to_world16, float const* KR9,
    float const* t3, float* const* depth_out)
{
    int rc = SMVSB_OK;
    cudaStream_t st = nullptr;
    try
    {
        if (n_views < 1 || !w || !h || !depth || !normals || !invproj9
            || !cam_to_world16 || !KR9 || !t3 || !depth_out)
            throw Error(SMVSB_ERR_INVALID, "smvsb_cut_depth_maps: arguments");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(SMVSB_ERR_CUDA, "no CUDA device (no CPU fallback)");
        if (device < 0 || device >= count)
            throw Error(SMVSB_ERR_INVALID, "device index out of range");
        std::lock_guard<std::mutex> guard(g_cut_lock);
        CUDA_CHECK(cudaSetDevice(device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        std::vector<DevBuf<float>> d_cut(n_views), d_z(n_views),
            d_nrm(n_views), d_out(n_views);
        std::vector<CutView> hv(n_views);
        for (int i = 0; i < n_views; ++i)
        {
            if (w[i] < 1 || h[i] < 1 || !depth[i] || !normals[i]
                || !depth_out[i])
                throw Error(SMVSB_ERR_INVALID,
                    "smvsb_cut_depth_maps: view without maps");
            size_t const n = static_cast<size_t>(w[i]) * h[i];
            d_cut[i].reserve(n); d_z[i].reserve(n); d_nrm[i].reserve(n * 3);
            d_out[i].reserve(n);
            CUDA_CHECK(cudaMemcpyAsync(d_cut[i].p, depth[i],
                n * sizeof(float), cudaMemcpyHostToDevice, st));
            CUDA_CHECK(cudaMemcpyAsync(d_nrm[i].p, normals[i],
                n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
            CutView& v = hv[i];
            v.w = w[i]; v.h = h[i];
            v.cut = d_cut[i].p; v.zdepth = d_z[i].p; v.normals = d_nrm[i].p;
            v.out = d_out[i].p;
            std::copy(invproj9 + 9 * i, invproj9 + 9 * i + 9, v.invproj);
            std::copy(cam_to_world16 + 16 * i, cam_to_world16 + 16 * i + 16,
                v.ctw);
            std::copy(KR9 + 9 * i, KR9 + 9 * i + 9, v.KR);
            std::copy(t3 + 3 * i, t3 + 3 * i + 3, v.t);
            float const* m = v.invproj;
            dim3 const grid((w[i] + 127) / 128, h[i]);
            to_zdepth_kernel<<<grid, 128, 0, st>>>(w[i], h[i], d_cut[i].p,
                m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8],
                d_z[i].p);
            CUDA_CHECK(cudaGetLastError());
        }
        DevBuf<CutView> d_views;
        d_views.reserve(n_views);
        CUDA_CHECK(cudaMemcpyAsync(d_views.p, hv.data(),
            n_views * sizeof(CutView), cudaMemcpyHostToDevice, st));
        for (int i = 0; i < n_views; ++i)
        {
            dim3 const grid((w[i] + 127) / 128, h[i]);
            cut_depth_maps_kernel<<<grid, 128, 0, st>>>(d_views.p, n_views,
                i);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaMemcpyAsync(depth_out[i], d_out[i].p,
                static_cast<size_t>(w[i]) * h[i] * sizeof(float),
                cudaMemcpyDeviceToHost, st));
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        count_device_launches(device, 2 * n_views);
    }
    catch (Error const& e)
    {
        g_cut_error = e.msg;
        rc = e.code;
    }
    if (st) cudaStreamDestroy(st);
    return rc;
}

} /* namespace smvsb */
