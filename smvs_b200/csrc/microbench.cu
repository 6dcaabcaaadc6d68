/*
 * microbench.cu -- the fp64 roof the Gauss-Newton construct kernel is held
 * against (SURVEY.md section 8d asks for a measured figure: MEASURED_PEAKS.json
 * carries HBM and bf16 tensor numbers only).
 *
 * dfma_peak_kernel: every thread runs 8 independent DFMA chains (latency
 * hidden by ILP and by 8 warps per scheduler), 2 flops per DFMA; the result is
 * stored so nothing is optimised away. Timed with CUDA events, best of 5.
 */
#include "common.cuh"

namespace smvsb {

namespace {

constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(256)
dfma_peak_kernel (double* __restrict__ out, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 4
    for (int i = 0; i < PEAK_ITERS; ++i)
    {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b);
        x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b);
        x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3))
        + ((x4 + x5) + (x6 + x7));
}

} /* namespace */

/* Measured dense fp64 FMA throughput of the device in TFLOP/s. */
double
measure_fp64_peak (int device)
{
    CUDA_CHECK(cudaSetDevice(device));
    int sms = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount,
        device));
    int const blocks = sms * 8, threads = 256;
    DevBuf<double> out;
    out.reserve(static_cast<size_t>(blocks) * threads);
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    double best = 0.0;
    try
    {
        CUDA_CHECK(cudaEventCreate(&e0));
        CUDA_CHECK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 6; ++rep)
        {
            CUDA_CHECK(cudaEventRecord(e0, st));
            dfma_peak_kernel<<<blocks, threads, 0, st>>>(out.p, 0.999999, 1e-7);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaEventRecord(e1, st));
            CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            double const flops = 2.0 * 8.0 * PEAK_ITERS
                * static_cast<double>(blocks) * threads;
            if (rep > 0)        /* the first launch warms up */
                best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
    }
    catch (...)
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        cudaStreamDestroy(st);
        throw;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    g_launches += 6;
    return best;
}

} /* namespace smvsb */
