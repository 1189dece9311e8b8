// philox.cuh -- counter-based Gaussian sampling for the control perturbations.
//
// Replaces the reference's per-sample cuRAND XORWOW state (48 B/sample, initialised by
// curand_init(tid,tid,tid) in set_data, src/point_mass.cu:780, and round-tripped through
// global memory every control step, :502-506) and its `0.025*curand_normal()` draws
// (src/point_mass_gpu.cu:85-90) with a stateless Philox-4x32-10 stream:
//
//     quad q = k_global / 4, row r = t*A + a, control step s
//     (x0,x1,x2,x3) = Philox4x32-10(counter = {q, r, s_lo, s_hi}, key = {seed_lo, seed_hi})
//     (e0,e1) = BoxMuller(x0,x1; sigma_a), (e2,e3) = BoxMuller(x2,x3; sigma_a)
//     eps[k = 4q+j, t, a] = e_j                    (a = r mod A)
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

// Host-precomputed sampler constants, passed to the kernels by value (constant bank):
// the ten Philox round keys of the seed and, per action dim, c[a] = -2 ln2 * sigma_a^2 so
// that sigma * sqrt(-2 ln u) = sqrt(c[a] * log2 u) costs one FMUL + one MUFU.
struct SamplerParams {
    uint32_t k0[10], k1[10];
    float c[kMaxAct];
};

inline SamplerParams make_sampler_params(unsigned long long seed, const float *sigma, int A)
{
    SamplerParams sp{};
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { sp.k0[i] = a; sp.k1[i] = b; a += kPhiloxW0; b += kPhiloxW1; }
    for (int i = 0; i < A; ++i) sp.c[i] = -1.3862943611198906f * (sigma[i] * sigma[i]);
    return sp;
}

#ifdef __CUDACC__
// ROUNDS = 10 is the generator of Random123 / cuRAND (the default); 7 is Random123's
// philox4x32_R<7>, the fewest rounds that pass BigCrush (Salmon et al., SC'11, table 2) -- 30 %
// fewer of the 32x32->64 multiplies that bound the sampling kernels.  Same key
// schedule, so sp.k0/k1 serve both.
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, const SamplerParams &sp)
{
    static_assert(ROUNDS == 7 || ROUNDS == 10, "Philox-4x32 with 7 or 10 rounds");
#pragma unroll
    for (int i = 0; i < ROUNDS; ++i) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ sp.k0[i], lo1, hi0 ^ c.w ^ sp.k1[i], lo0);
    }
    return c;
}

__device__ __forceinline__ float mufu_lg2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_sin(float x)
{
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_cos(float x)
{
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Box-Muller on two 32-bit draws, already scaled by sigma (through c = -2 ln2 sigma^2):
//   u     = xa*2^-32 + 2^-33            in (0,1]
//   theta = xb*2pi*2^-32 + 2pi*2^-33    in (0,2pi]
//   r     = sqrt(|c * log2 u|)          (|.|: MUFU.LG2 may return +tiny for u -> 1)
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float c, float &n0, float &n1)
{
    const float u  = __fmaf_rn(__uint2float_rn(xa), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float th = __fmaf_rn(__uint2float_rn(xb), 1.4629180792671596e-9f, 7.314590396335798e-10f);
    const float r  = mufu_sqrt(fabsf(__fmul_rn(c, mufu_lg2(u))));
    n0 = __fmul_rn(r, mufu_sin(th));
    n1 = __fmul_rn(r, mufu_cos(th));
}

// eps of (quad q, row r, step) for the four samples 4q..4q+3, c = sp.c[a] of that row
template <int ROUNDS = 10>
__device__ __forceinline__ float4 sample4(uint32_t q, uint32_t r, unsigned long long step,
                                          const SamplerParams &sp, float c)
{
    const uint4 x = philox4x32<ROUNDS>(make_uint4(q, r, (uint32_t)step, (uint32_t)(step >> 32)), sp);
    float4 n;
    box_muller(x.x, x.y, c, n.x, n.y);
    box_muller(x.z, x.w, c, n.z, n.w);
    return n;
}
#endif  // __CUDACC__

}  // namespace mppi
