// common.cuh -- shared device-side types and PTX helpers for the sm_100a MPPI core.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mppi_b200.h"

namespace mppi {

constexpr int kMaxAct = MPPI_MAX_ACT;

// ---------------------------------------------------------------------------------
// Device-resident problem description (one per handle).  Written by set_problem /
// set_state with small async copies; read (uniformly, L2-resident) by the kernels, so a
// captured CUDA graph never needs its kernel parameters patched.
// Replaces the per-sample PointMassModelGpu / Cost pointer bundles of the reference
// (include/point_mass_gpu.hpp:46-90, include/cost.hpp:32-46: 168 B + 48 B per sample).
// ---------------------------------------------------------------------------------
struct ProblemDev {
    float x0[2 * kMaxAct];
    float goal[2 * kMaxAct];
    float w[2 * kMaxAct];
    float inv_s[kMaxAct];
    float sigma[kMaxAct];
    float init_act[kMaxAct];
    float max_act[kMaxAct];
    float g[4];            // state gain {1, dt, 0, 1}       (src/point_mass.cu:48-51)
    float b[2];            // action gain {dt*dt/2, dt}      (src/point_mass.cu:46-47)
    float lambda;
    float neg_inv_lambda;  // -(1/lambda), float ops         (src/point_mass.cu:518)
    float wf[2 * kMaxAct]; // weights of the terminal cost (Cost::final_cost); == w unless
                           // mppi_set_terminal_weights gave the final state its own
};

// Per-step control block in device memory.
struct CtlDev {
    unsigned long long min_key;   // packed (ordered(S) << 32 | k_global), atomicMin target
    unsigned long long step;      // control-step counter == Philox counter words 2,3
    unsigned long long last_key;  // min_key of the last finished step (for get_step_info)
    float eta;                    // normaliser of the last finished step
    unsigned int done;            // CTA ticket counter of the merged average+finalize kernel
    unsigned int comm_error;      // set when a peer-mailbox wait timed out
    unsigned int pad_;
    unsigned long long t_xchg[4]; // %globaltimer (ns) of the last K-shard exchange: push begins,
                                  // own data + flags out, every peer's flag seen, merged
};

constexpr unsigned long long kMinKeyInit = ~0ull;

// monotone float -> uint32 map (total order equals float order, -0 < +0)
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o)
{
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}

// ---------------------------------------------------------------------------------
// Cross-CTA / cross-GPU sums are taken in 64-bit fixed point (scale 2^30, range +-8.6e9,
// resolution 9.3e-10): integer addition is associative, so atomics and all-reduces give
// the same bits whatever the arrival order.  Each contribution is a float partial that is
// already the sum of >= 256 terms; its conversion error (<= 4.7e-10) is far below the
// float32 rounding of the result.
// ---------------------------------------------------------------------------------
constexpr double kAccScale = 1073741824.0;           // 2^30
__device__ __forceinline__ void acc_add(long long *acc, float v)
{
    const long long q = __double2ll_rn((double)v * kAccScale);
    atomicAdd(reinterpret_cast<unsigned long long *>(acc), (unsigned long long)q);
}
__device__ __forceinline__ float acc_to_float(long long q)
{
    return (float)((double)q * (1.0 / kAccScale));
}

// ---------------------------------------------------------------------------------
// global memory access with cache hints
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_f4(float *p, float4 v)
{
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk[.tensor]) -- sm_90+/sm_100a
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tiled TMA load: box -> shared memory, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];"
        :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both sides 16 B aligned).
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}

// ---------------------------------------------------------------------------------
// intra-CTA signalling through shared memory (step kernel: rollout warp -> producer warp)
// and the generic -> async proxy fence a TMA load needs to see global data written by
// ordinary stores of the same kernel
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_cta_shared_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_cta_shared_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ float warp_min_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------
// system-scope accesses for the peer-mailbox exchange over NVLink (multi-GPU)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_v2_u64(unsigned long long *p, unsigned long long a,
                                                      unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#endif  // __CUDACC__

// Peer mailbox of one rank: two buffers (parity of the step number) of one slot per sender;
// slot s is written only by rank s (over NVLink for s != self).
//   [0] seq (single exchange) / key_seq  [1] key  [2] acc_seq  [3] reserved  [4 ...] acc[R+1]
constexpr int kMailboxHeaderWords = 4;
constexpr int kMaxWorld = 16;
constexpr int kMailboxBuffers = 2;
inline size_t mailbox_slot_words(int rows)
{
    return (size_t)((kMailboxHeaderWords + rows + 1 + 15) / 16 * 16);     // 128-byte multiple
}

}  // namespace mppi
