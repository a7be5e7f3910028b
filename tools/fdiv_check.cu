// Device check of the fast-mode division / square root (transport.cuh) against IEEE div.rn / sqrt.rn over 2^32 operands:
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fdivtest tools/fdiv_check.cu && tools/bin/fdivtest
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstring>
__device__ __forceinline__ double xdiv(double a, double b) {
    double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    x = fma(x, fma(-b, x, 1.0), x);
    const double q = a * x; return fma(fma(-b, q, a), x, q);
}
__device__ __forceinline__ double xsqrt(double a) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y; const double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h); g = fma(fma(-g, g, a), h, g); return (a > 0.0) ? g : 0.0;
}
__global__ void k(unsigned long long n, double* out) {
    double md = 0, ms = 0;
    for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        double a = ldexp((double)(h >> 11) * (1.0 / 9007199254740992.0) + 0.5, (int)(h & 127) - 64);
        unsigned long long g = h * 0x94D049BB133111EBull; g ^= g >> 31;
        double b = ldexp((double)(g >> 11) * (1.0 / 9007199254740992.0) + 0.5, (int)(g & 127) - 64);
        double q = xdiv(a, b), qe = a / b; double s = xsqrt(a), se = sqrt(a);
        md = fmax(md, fabs(q - qe) / (qe * 1.1102230246251565e-16)); ms = fmax(ms, fabs(s - se) / (se * 1.1102230246251565e-16));
    }
    for (int o = 16; o; o >>= 1) { md = fmax(md, __shfl_xor_sync(~0u, md, o)); ms = fmax(ms, __shfl_xor_sync(~0u, ms, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax((unsigned long long*)out, __double_as_longlong(md)); atomicMax((unsigned long long*)out + 1, __double_as_longlong(ms)); }
}
int main() { double* o; cudaMalloc(&o, 16); cudaMemset(o, 0, 16); k<<<592, 256>>>(1ull << 32, o); double h[2]; cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
  printf("max error vs IEEE over 2^32 operands (units of 2^-53 relative): fdiv %.3f fsqrt %.3f\n", h[0], h[1]); return 0; }
