// peak.cu -- register-only DFMA loop: the measured FP64 roofline denominator (MEASURED_PEAKS.json carries
// only HBM and bf16 figures).  Eight independent FMA chains per thread, no memory traffic in the loop.
#include "internal.cuh"

namespace nufi_b200
{

namespace
{

constexpr int kChains = 8;
constexpr int kInner = 64;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int outer, double a, double b)
{
    double x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // keep the chains alive
}

} // namespace

int measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return fail(nullptr, NUFI_B200_ERR_ARG, "tflops is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, NUFI_B200_ERR_CUDA, "no CUDA device");
    if (device < 0) cudaGetDevice(&device);
    NUFI_CUDA_CHECK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    NUFI_CUDA_CHECK(nullptr, cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double *out = nullptr;
    NUFI_CUDA_CHECK(nullptr, cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int outer = 2000;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, outer, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(out);
            return fail(nullptr, NUFI_B200_ERR_CUDA, cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * kChains * kInner * static_cast<double>(outer) * blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return NUFI_B200_OK;
}

} // namespace nufi_b200
