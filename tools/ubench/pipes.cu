// Microbenchmark: per-sub-partition throughput of the instruction classes the fused
// sample+rollout kernel is made of (sm_100a): IMAD.WIDE.U32 / IMAD / IMAD.HI (Philox
// multiplies), LOP3, FFMA / FFMA2 (dynamics and cost), MUFU (Box-Muller), I2FP, and the
// Philox round itself -- as a function of the number of warps per sub-partition.
// Prints warp-instructions per clock per sub-partition (clock64 inside the kernel, so no
// assumption about the SM clock).  One CTA per SM, `W` warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8;       // independent chains per thread

struct OpWide {   // IMAD.WIDE.U32: 32x32 -> 64
    static constexpr const char *name = "IMAD.WIDE.U32";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            unsigned long long r;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a[i]), "r"(m));
            a[i] = (uint32_t)(r >> 32); b[i] = (uint32_t)r;
        }
    }
};
struct OpLo {     // IMAD (low 32 bits)
    static constexpr const char *name = "IMAD (lo)";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[i]), "r"(m), "r"(b[i]));
    }
};
struct OpHi {     // IMAD.HI.U32
    static constexpr const char *name = "IMAD.HI.U32";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
        (void)b;
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(m));
    }
};
struct OpLop {    // LOP3 with three register operands
    static constexpr const char *name = "LOP3";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(m));
    }
};
struct OpFfma {
    static constexpr const char *name = "FFMA";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float x = __uint_as_float(a[i]);
            asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x) : "f"(x), "f"(__uint_as_float(m)), "f"(__uint_as_float(b[i])));
            a[i] = __float_as_uint(x);
        }
    }
};
struct OpFfma2 {
    static constexpr const char *name = "FFMA2";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int i = 0; i < CH; i += 2) {
            unsigned long long x, y, z;
            asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "r"(a[i]), "r"(a[i + 1]));
            asm("mov.b64 %0, {%1,%2};" : "=l"(y) : "r"(m), "r"(m));
            asm("mov.b64 %0, {%1,%2};" : "=l"(z) : "r"(b[i]), "r"(b[i + 1]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(y), "l"(z));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(y), "l"(z));
            asm("mov.b64 {%0,%1}, %2;" : "=r"(a[i]), "=r"(a[i + 1]) : "l"(x));
        }
    }
};
struct OpMufu {
    static constexpr const char *name = "MUFU (ex2/lg2)";
    static constexpr int per = 2;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
        (void)b; (void)m;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float x = __uint_as_float(a[i]);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(x));
            asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(x));
            a[i] = __float_as_uint(x);
        }
    }
};
struct OpI2f {
    static constexpr const char *name = "I2FP.F32.U32";
    static constexpr int per = 1;
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
        (void)b; (void)m;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float x;
            asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x) : "r"(a[i]));
            a[i] = __float_as_uint(x);
        }
    }
};
// one Philox-4x32 round on CH/4 = 2 independent states (4 instructions per state)
struct OpPhilox {
    static constexpr const char *name = "Philox round (2 WIDE + 2 LOP3)";
    static constexpr int per = 2;      // CH/4 states x 4 instr = CH instr ... x per: two rounds per run
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
                unsigned long long p0, p1;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[i]), "r"(0xD2511F53u));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(a[i + 2]), "r"(0xCD9E8D57u));
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ a[i + 1] ^ m, n2 = (uint32_t)(p0 >> 32) ^ a[i + 3] ^ b[r];
                a[i] = n0; a[i + 1] = (uint32_t)p1; a[i + 2] = n2; a[i + 3] = (uint32_t)p0;
            }
        }
    }
};
// the fused kernel's mix per Philox call: 20 WIDE + 20 LOP3 + 8 MUFU + 25 FFMA2 + 12 scalar FP
struct OpMix {
    static constexpr const char *name = "mix: 2 Philox rounds x2 states + 2 MUFU + 5 FFMA2 + 2 FFMA  (=17 instr)";
    static constexpr int per = 17;     // counted per run() below; CH-independent, see main
    __device__ static void run(uint32_t (&a)[CH], uint32_t (&b)[CH], uint32_t m) {
        // 4 WIDE + 4 LOP3 : one round on two states
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
            unsigned long long p0, p1;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[i]), "r"(0xD2511F53u));
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(a[i + 2]), "r"(0xCD9E8D57u));
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ a[i + 1] ^ m, n2 = (uint32_t)(p0 >> 32) ^ a[i + 3] ^ m;
            a[i] = n0; a[i + 1] = (uint32_t)p1; a[i + 2] = n2; a[i + 3] = (uint32_t)p0;
        }
        // 2 MUFU on b[0], b[1]
        float x = __uint_as_float(b[0]), y = __uint_as_float(b[1]);
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(x));
        asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(y));
        b[0] = __float_as_uint(y); b[1] = __float_as_uint(x);
        // 5 FFMA2 on b[2..7] (three packed accumulators), 2 FFMA
        unsigned long long u, v, w, s;
        asm("mov.b64 %0, {%1,%2};" : "=l"(u) : "r"(b[2]), "r"(b[3]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "r"(b[4]), "r"(b[5]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(w) : "r"(b[6]), "r"(b[7]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(s) : "r"(m), "r"(m));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u) : "l"(u), "l"(s), "l"(v));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(s), "l"(w));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(w) : "l"(w), "l"(s), "l"(u));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u) : "l"(u), "l"(s), "l"(w));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(s), "l"(u));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(b[2]), "=r"(b[3]) : "l"(u));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(b[4]), "=r"(b[5]) : "l"(v));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(b[6]), "=r"(b[7]) : "l"(w));
        float f = __uint_as_float(b[0]), g = __uint_as_float(b[1]);
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(f), "f"(__uint_as_float(m)), "f"(g));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(g) : "f"(g), "f"(__uint_as_float(m)), "f"(f));
        b[0] = __float_as_uint(f); b[1] = __float_as_uint(g);
    }
};

template <class OP>
__global__ void bench_kernel(uint32_t *out, long long *cyc, int n, uint32_t m)
{
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 2654435761u + i * 40503u + 1u; b[i] = a[i] ^ 0x3f800000u; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) OP::run(a, b, m);
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class OP>
void run(uint32_t *d, long long *dc, int instr_per_run)
{
    printf("%-72s", OP::name);
    for (int w = 1; w <= 8; w *= 2) {
        const int n = 4000;
        bench_kernel<OP><<<148, 128 * w>>>(d, dc, n, 0x3f7fbe77u);
        bench_kernel<OP><<<148, 128 * w>>>(d, dc, n, 0x3f7fbe77u);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
        s /= 148;
        printf("  W=%d: %.3f", w, (double)w * n * instr_per_run / s);
    }
    printf("   [warp-instr / clk / sub-partition]\n");
}

int main()
{
    uint32_t *d; long long *dc;
    cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 148 * 8);
    run<OpWide>(d, dc, CH);
    run<OpLo>(d, dc, CH);
    run<OpHi>(d, dc, CH);
    run<OpLop>(d, dc, CH);
    run<OpFfma>(d, dc, CH);
    run<OpFfma2>(d, dc, CH);          // CH/2 pairs x 2 FFMA2
    run<OpMufu>(d, dc, CH * 2);
    run<OpI2f>(d, dc, CH);
    run<OpPhilox>(d, dc, 2 * (CH / 4) * 4);
    run<OpMix>(d, dc, (CH / 4) * 4 + 2 + 5 + 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
