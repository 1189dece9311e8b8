// Are FADD2 / FMUL2 / FFMA2 bit-identical to scalar add/mul/fma (rn)?  Random + edge inputs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__global__ void k(const float* A, const float* B, const float* Cc, int n, unsigned* bad) {
    int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 >= n) return;
    unsigned long long a = pk(A[i], A[i + 1]), b = pk(B[i], B[i + 1]), c = pk(Cc[i], Cc[i + 1]), d;
    float x, y;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); up(d, x, y);
    if (__float_as_uint(x) != __float_as_uint(__fmaf_rn(A[i], B[i], Cc[i])) || __float_as_uint(y) != __float_as_uint(__fmaf_rn(A[i+1], B[i+1], Cc[i+1]))) atomicAdd(&bad[0], 1);
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); up(d, x, y);
    if (__float_as_uint(x) != __float_as_uint(__fmul_rn(A[i], B[i])) || __float_as_uint(y) != __float_as_uint(__fmul_rn(A[i+1], B[i+1]))) atomicAdd(&bad[1], 1);
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(c)); up(d, x, y);
    if (__float_as_uint(x) != __float_as_uint(__fadd_rn(A[i], Cc[i])) || __float_as_uint(y) != __float_as_uint(__fadd_rn(A[i+1], Cc[i+1]))) atomicAdd(&bad[2], 1);
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(c)); up(d, x, y);
    if (__float_as_uint(x) != __float_as_uint(__fsub_rn(A[i], Cc[i])) || __float_as_uint(y) != __float_as_uint(__fsub_rn(A[i+1], Cc[i+1]))) atomicAdd(&bad[3], 1);
}
int main() {
    const int n = 1 << 22;
    std::vector<float> a(n), b(n), c(n);
    srand(1);
    for (int i = 0; i < n; ++i) {
        auto rnd = [&]() { unsigned u = ((unsigned)rand() << 16) ^ (unsigned)rand(); float f; if (i % 3 == 0) { u = (u & 0x807fffffu) | ((100u + (u >> 23) % 56u) << 23); } memcpy(&f, &u, 4); if (f != f) f = 1.0f; return f; };
        a[i] = rnd(); b[i] = rnd(); c[i] = rnd();
    }
    float *da, *db, *dc; unsigned* dbad;
    cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dc, n * 4); cudaMalloc(&dbad, 16); cudaMemset(dbad, 0, 16);
    cudaMemcpy(da, a.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dc, c.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 2 / 256, 256>>>(da, db, dc, n, dbad);
    unsigned bad[4]; cudaMemcpy(bad, dbad, 16, cudaMemcpyDeviceToHost);
    printf("mismatching pairs out of %d: fma %u mul %u add %u sub %u (%s)\n", n / 2, bad[0], bad[1], bad[2], bad[3], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
