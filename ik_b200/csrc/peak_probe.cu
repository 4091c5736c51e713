// FMA-pipe peak probe: the roofline denominator for the IK kernels is the non-tensor FP64 / FP32 FMA rate,
// which MEASURED_PEAKS.json does not carry (it holds HBM copy bandwidth and bf16 tensor throughput).  This
// kernel issues long chains of independent DFMA / FFMA from every resident warp and reports TFLOP/s.
#include <cuda_runtime.h>

#include "../../include/ikb200.h"

namespace {

template <typename T, int CHAINS>
__global__ void __launch_bounds__(256) fma_probe_kernel(T *out, T a, T b, int iters) {
    T x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = T(threadIdx.x + i) * T(1e-3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    T s = T(0);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == T(-12345.678)) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains live
}

template <typename T> int run_probe(double *tflops) {
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return IKB_ERR_CUDA;
    constexpr int CHAINS = 8;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    T *out = nullptr;
    if (cudaMalloc(&out, sizeof(T) * blocks * threads) != cudaSuccess) return IKB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_probe_kernel<T, CHAINS><<<blocks, threads>>>(out, T(0.999), T(1e-3), iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return IKB_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * CHAINS * 16.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return IKB_OK;
}

}  // namespace

extern "C" int ikb_measure_fma_peak(int dtype, int device, double *tflops) {
    if (!tflops) return IKB_ERR_INVALID_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return IKB_ERR_NO_DEVICE;
    int rc = dtype == IKB_F64 ? run_probe<double>(tflops) : run_probe<float>(tflops);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}
