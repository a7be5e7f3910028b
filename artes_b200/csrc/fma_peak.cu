// fma_peak.cu -- FP64 / FP32 FMA issue-rate microbenchmark.
// MEASURED_PEAKS.json only holds an HBM copy and a bf16 GEMM figure; this path is bound by the FP64
// pipe, so bench.py measures its own denominator with these kernels.
#include <cuda_runtime.h>

namespace artes {

template <typename Tp, int ILP>
__global__ void __launch_bounds__(256) fma_kernel(Tp* out, int iters, Tp a, Tp b) {
    Tp acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = (Tp)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = acc[k] * a + b;
    }
    Tp s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == (Tp)123456789) out[0] = s;  // never true: keeps the loop alive
}

template <typename Tp>
static cudaError_t time_one(double* tflops, int sm_count, int iters, cudaStream_t stream) {
    constexpr int ILP = 8;
    Tp* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(Tp));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sm_count * 8, threads = 256;
    fma_kernel<Tp, ILP><<<blocks, threads, 0, stream>>>(d, iters / 8, (Tp)1.0000001, (Tp)1e-9);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, stream);
        fma_kernel<Tp, ILP><<<blocks, threads, 0, stream>>>(d, iters, (Tp)1.0000001, (Tp)1e-9);
        cudaEventRecord(e1, stream);
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * ILP * (double)iters * blocks * threads;
        best = fl / (ms * 1e-3) / 1e12 > best ? fl / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return e == cudaSuccess ? cudaGetLastError() : e;
}

cudaError_t fma_peak(double* fp64_tflops, double* fp32_tflops, int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaSuccess;
    if (fp64_tflops) { e = time_one<double>(fp64_tflops, sm_count, 1 << 15, stream); if (e != cudaSuccess) return e; }
    if (fp32_tflops) e = time_one<float>(fp32_tflops, sm_count, 1 << 17, stream);
    return e;
}

}  // namespace artes
