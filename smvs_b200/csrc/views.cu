/*
 * views.cu -- StereoView::set_scale on the device (reference:
 * lib/stereo_view.cc:24-46, 97-188; mve::image::blur_gaussian,
 * byte_to_float_image): per-scale Gaussian blur of the byte image, then the
 * 3x3 quadratic-fit gradient (2 ch) and Hessian (3 ch). This is the producer
 * of the Gauss-Newton kernels' image inputs ("next"
 * row 1); doing it here shrinks the per-scale upload from 41.5 MB to 2 MB per
 * 2 MP view.
 *
 * Bit-compatible with the CPU: fp32 blur with the reference's operation
 * order (value * weight, then add; no FMA contraction; normalise by the
 * weight sum), kernel weights from the host's expf, fp64 stencil with the
 * 6x9 matrix applied term by term.
 */
#include <cmath>

#include "common.cuh"

namespace smvsb {

namespace {

constexpr int BLUR_MAX_KS = 64;

struct BlurKernel
{
    int ks;
    float w[BLUR_MAX_KS + 1];
    float wsum;
};

__device__ __forceinline__ float
pixel_value (uint8_t v)
{
    /* mve::image::byte_to_float_image: v / 255, clamped */
    float const f = __fdiv_rn(static_cast<float>(v), 255.0f);
    return fminf(1.0f, fmaxf(0.0f, f));
}

__device__ __forceinline__ float
pixel_value (float v)
{
    return v;
}

/* x pass: (byte -> float, v / 255, clamped ->) blur along x */
template <typename T>
__global__ void
blur_x_kernel (T const* __restrict__ img, int w, int h,
    BlurKernel const k, float* __restrict__ out)
{
    int const x = blockIdx.x * blockDim.x + threadIdx.x;
    int const y = blockIdx.y;
    if (x >= w)
        return;
    T const* row = img + static_cast<size_t>(y) * w;
    float acc = 0.0f;
    for (int i = -k.ks; i <= k.ks; ++i)
    {
        int const xi = min(max(x + i, 0), w - 1);
        float const v = pixel_value(row[xi]);
        acc = __fadd_rn(acc, __fmul_rn(v, k.w[abs(i)]));
    }
    out[static_cast<size_t>(y) * w + x] = __fdiv_rn(acc, k.wsum);
}

__global__ void
blur_y_kernel (float const* __restrict__ in, int w, int h,
    BlurKernel const k, float* __restrict__ out)
{
    int const x = blockIdx.x * blockDim.x + threadIdx.x;
    int const y = blockIdx.y;
    if (x >= w)
        return;
    float acc = 0.0f;
    for (int i = -k.ks; i <= k.ks; ++i)
    {
        int const yi = min(max(y + i, 0), h - 1);
        acc = __fadd_rn(acc, __fmul_rn(in[static_cast<size_t>(yi) * w + x],
            k.w[abs(i)]));
    }
    out[static_cast<size_t>(y) * w + x] = __fdiv_rn(acc, k.wsum);
}

__global__ void
byte_to_float_kernel (uint8_t const* __restrict__ img, int n,
    float* __restrict__ out)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float const v = __fdiv_rn(static_cast<float>(img[i]), 255.0f);
    out[i] = fminf(1.0f, fmaxf(0.0f, v));
}

/* compute_gradients_and_hessian, lib/stereo_view.cc:97-188. Row r of the 6x9
 * matrix times the 3x3 neighbourhood (x offset outer, y offset inner), all
 * nine terms in order like math::Matrix::mult.
 *   mode 0: write float2 gradient            (main view, shading gradient)
 *   mode 1: write packed texel gx gy hxx hxy hyy 0 0 0 (neighbours)        */
__global__ void
grad_hess_kernel (float const* __restrict__ in, int w, int h, int mode,
    float* __restrict__ out)
{
    int const x = blockIdx.x * blockDim.x + threadIdx.x;
    int const y = blockIdx.y;
    if (x >= w)
        return;
    size_t const pix = static_cast<size_t>(y) * w + x;
    double r[6] = {0, 0, 0, 0, 0, 0};
    if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1)
    {
        double const s6 = 1.0 / 6.0, s3 = -1.0 / 3.0, s4 = 1.0 / 4.0;
        double const M[6][9] = {
            { s6, s6, s6, s3, s3, s3, s6, s6, s6 },
            { s6, s3, s6, s6, s3, s6, s6, s3, s6 },
            { s4, 0.0, -s4, 0.0, 0.0, 0.0, -s4, 0.0, s4 },
            { -s6, -s6, -s6, 0.0, 0.0, 0.0, s6, s6, s6 },
            { -s6, 0.0, s6, -s6, 0.0, s6, -s6, 0.0, s6 },
            { -1.0 / 9.0, 2.0 / 9.0, -1.0 / 9.0, 2.0 / 9.0, 5.0 / 9.0,
              2.0 / 9.0, -1.0 / 9.0, 2.0 / 9.0, -1.0 / 9.0 } };
        double v[9];
        int c = 0;
#pragma unroll
        for (int a = -1; a < 2; ++a)
#pragma unroll
            for (int b = -1; b < 2; ++b)
                v[c++] = in[static_cast<size_t>(y + b) * w + (x + a)];
#pragma unroll
        for (int row = 0; row < 5; ++row)
        {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k)
                s = __dadd_rn(s, __dmul_rn(M[row][k], v[k]));
            r[row] = s;
        }
    }
    if (mode == 0)
    {
        reinterpret_cast<float2*>(out)[pix] = make_float2(
            static_cast<float>(r[3]), static_cast<float>(r[4]));
    }
    else
    {
        float4 a, b;
        a.x = static_cast<float>(r[3]);
        a.y = static_cast<float>(r[4]);
        a.z = static_cast<float>(__dmul_rn(2.0, r[0]));
        a.w = static_cast<float>(r[2]);
        b.x = static_cast<float>(__dmul_rn(2.0, r[1]));
        b.y = b.z = b.w = 0.0f;
        reinterpret_cast<float4*>(out)[2 * pix] = a;
        reinterpret_cast<float4*>(out)[2 * pix + 1] = b;
    }
}

__global__ void
unpack_texels_kernel (float const* __restrict__ texels, int n,
    float* __restrict__ grad, float* __restrict__ hess)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 const a = reinterpret_cast<float4 const*>(texels)[2 * i];
    float4 const b = reinterpret_cast<float4 const*>(texels)[2 * i + 1];
    grad[2 * i] = a.x; grad[2 * i + 1] = a.y;
    hess[3 * i] = a.z; hess[3 * i + 1] = a.w; hess[3 * i + 2] = b.x;
}

} /* namespace */

/* sigma = 0.12 * 2^scale + 0.2 (lib/stereo_view.cc:28), narrowed to float
 * at the call of blur_gaussian<float>(img, float sigma). */
static BlurKernel
make_blur_kernel (int scale)
{
    double const sigma_d = 0.12 * std::pow(2.0, scale) + 0.2;
    float const sigma = static_cast<float>(sigma_d);
    BlurKernel k;
    k.ks = static_cast<int>(std::ceil(sigma * 2.884f));
    if (k.ks > BLUR_MAX_KS)
        throw Error(SMVSB_ERR_INVALID, "blur radius too large for this scale");
    for (int i = 0; i <= k.ks; ++i)
    {
        float const x = static_cast<float>(i);
        k.w[i] = std::exp(-((x * x) / (2.0f * sigma * sigma)));
    }
    float wsum = 0.0f;
    for (int i = -k.ks; i <= k.ks; ++i)
        wsum += k.w[i < 0 ? -i : i];
    k.wsum = wsum;
    return k;
}

/* Blur + derivatives of one byte image that is already on the device.
 * tmp_a / tmp_b: w*h float scratch. */
void
device_set_scale (smvsb_ctx* c, uint8_t const* img_dev, int w, int h,
    int scale, float* tmp_a, float* tmp_b, int mode, float* out_dev)
{
    BlurKernel const k = make_blur_kernel(scale);
    dim3 const block(128, 1), grid((w + 127) / 128, h);
    blur_x_kernel<uint8_t><<<grid, block, 0, c->stream>>>(img_dev, w, h, k,
        tmp_a);
    CUDA_CHECK(cudaGetLastError());
    blur_y_kernel<<<grid, block, 0, c->stream>>>(tmp_a, w, h, k, tmp_b);
    CUDA_CHECK(cudaGetLastError());
    grad_hess_kernel<<<grid, block, 0, c->stream>>>(tmp_b, w, h, mode,
        out_dev);
    CUDA_CHECK(cudaGetLastError());
    count_launches(c, 3);
}

/* The same from a float image on the device (StereoView::image); the blurred
 * image is left in tmp_b, the packed gradient / Hessian texels in out_dev. */
void
device_set_scale_float (smvsb_ctx* c, float const* img_dev, int w, int h,
    int scale, float* tmp_a, float* tmp_b, float* out_dev)
{
    BlurKernel const k = make_blur_kernel(scale);
    dim3 const block(128, 1), grid((w + 127) / 128, h);
    blur_x_kernel<float><<<grid, block, 0, c->stream>>>(img_dev, w, h, k,
        tmp_a);
    CUDA_CHECK(cudaGetLastError());
    blur_y_kernel<<<grid, block, 0, c->stream>>>(tmp_a, w, h, k, tmp_b);
    CUDA_CHECK(cudaGetLastError());
    grad_hess_kernel<<<grid, block, 0, c->stream>>>(tmp_b, w, h, 1, out_dev);
    CUDA_CHECK(cudaGetLastError());
    count_launches(c, 3);
}

/* initialize_linear without gamma (lib/stereo_view.cc:64-84): shading image =
 * byte_to_float(image), plus its gradient. */
void
device_shading_inputs (smvsb_ctx* c, uint8_t const* img_dev, int w, int h,
    float* shading_dev, float* shading_grad_dev)
{
    int const n = w * h;
    byte_to_float_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(img_dev, n,
        shading_dev);
    CUDA_CHECK(cudaGetLastError());
    dim3 const block(128, 1), grid((w + 127) / 128, h);
    grad_hess_kernel<<<grid, block, 0, c->stream>>>(shading_dev, w, h, 0,
        shading_grad_dev);
    CUDA_CHECK(cudaGetLastError());
    count_launches(c, 2);
}

void
device_unpack_texels (smvsb_ctx* c, float const* texels, int n, float* grad,
    float* hess)
{
    unpack_texels_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(texels, n,
        grad, hess);
    CUDA_CHECK(cudaGetLastError());
    count_launches(c, 1);
}

} /* namespace smvsb */
