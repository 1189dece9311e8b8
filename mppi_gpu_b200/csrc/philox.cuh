// philox.cuh -- counter-based Gaussian sampling for the control perturbations.
//
// Replaces the reference's per-sample cuRAND XORWOW state (48 B/sample, initialised by
// curand_init(tid,tid,tid) in set_data, src/point_mass.cu:780, and round-tripped through
// global memory every control step, :502-506) and its `0.025*curand_normal()` draws
// (src/point_mass_gpu.cu:85-90) with a stateless Philox-4x32-10 stream:
//
//     quad q = k_global / 4, row r = t*A + a, control step s
//     (x0,x1,x2,x3) = Philox4x32-10(counter = {q, r, s_lo, s_hi}, key = {seed_lo, seed_hi})
//     (n0,n1) = BoxMuller(x0,x1), (n2,n3) = BoxMuller(x2,x3)
//     eps[k = 4q+j, t, a] = sigma[a] * n_j
//
// One Philox call yields the float4 of four consecutive samples in one row of the K-minor
// eps layout, so the store is a single 16-byte vector store and the value of eps[k,t,a]
// depends only on (seed, step, k_global, t, a) -- never on the launch shape or on how K is
// sharded over GPUs.  The CPU restatement is oracle_sample_eps (oracle/mppi_oracle.c).
#pragma once

#include "common.cuh"

namespace mppi {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += kPhiloxW0;
        k.y += kPhiloxW1;
    }
    return c;
}

// uniform in (0,1]: x*2^-32 + 2^-33 as one fused op
__device__ __forceinline__ float u01(uint32_t x)
{
    return __fmaf_rn(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// two N(0,1) values from two 32-bit draws; MUFU.LG2 / MUFU.SQRT / MUFU.SIN / MUFU.COS
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float &n0, float &n1)
{
    const float u = u01(xa);
    const float v = u01(xb);
    // -2 ln(u) = (-2 ln 2) * log2(u)
    const float r = sqrt_approx(-1.3862943611198906f * __log2f(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    n0 = r * s;
    n1 = r * c;
}

// the four standard normals of (quad q, row r, step, seed)
__device__ __forceinline__ float4 normal4(uint32_t q, uint32_t r, unsigned long long step,
                                          unsigned long long seed)
{
    const uint4 x = philox4x32_10(make_uint4(q, r, (uint32_t)step, (uint32_t)(step >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float4 n;
    box_muller(x.x, x.y, n.x, n.y);
    box_muller(x.z, x.w, n.z, n.w);
    return n;
}

}  // namespace mppi
