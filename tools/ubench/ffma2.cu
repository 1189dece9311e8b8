// Microbenchmark: issue/pipe throughput of scalar FFMA vs packed FFMA2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int CH> __global__ void k_scalar(float* out, int n, float s) {
    float a[CH]; for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = __fmaf_rn(a[i], s, 0.5f + i);
    }
    float r = 0; for (int i = 0; i < CH; ++i) r += a[i]; out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int CH> __global__ void k_packed(float* out, int n, float s) {
    unsigned long long a[CH], b = pk(s, s), c[CH];
    for (int i = 0; i < CH; ++i) { a[i] = pk(threadIdx.x * 0.001f + i, 1.0f + i); c[i] = pk(0.5f + i, 0.25f + i); }
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma2(a[i], b, c[i]);
    }
    float r = 0; for (int i = 0; i < CH; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n = 20000; const int CH = 8;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_scalar<CH><<<148 * 4, 256>>>(d, n, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 4 * 256 * (double)n * CH;   // FMAs
        printf("scalar FFMA : %.3f ms  %.1f GFMA/s  (%.2f FMA/clk/SM at 1.965 GHz)\n", ms, fl / ms / 1e6, fl / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0); k_packed<CH><<<148 * 4, 256>>>(d, n, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 148.0 * 4 * 256 * (double)n * CH * 2;
        printf("packed FFMA2: %.3f ms  %.1f GFMA/s  (%.2f FMA/clk/SM)\n", ms, fl / ms / 1e6, fl / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
