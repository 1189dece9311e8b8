// kernels.cu -- hand-written sm_100a kernels of one MPPI control step.
//
// Data layout in HBM (per shard, all float32):
//   eps      [R = T*A rows][k_pad]   K-minor: row r = t*A+a holds that perturbation component
//                                    of every sample; k_pad = K_local rounded up to 256.
//                                    (The reference keeps [K][T][A], src/point_mass.cu:784,
//                                    which makes every per-sample access a T*A-float stride.)
//   S, wt    [k_pad]                 rollout costs and unnormalised weights
//   acc      [R+1] int64             fixed-point (2^-30) accumulators: sum_k wt_k eps[r][k], eta
//   U, U_prev[T*A]
//
// Kernel chain of a control step (reference: PointMassModel::get_act, src/point_mass.cu:129-203):
//   sample -> rollout(+min) -> weights(+eta partials) -> average -> finalize
#include "kernels.cuh"
#include "philox.cuh"
#include "model.cuh"
#include "finalize.cuh"

namespace mppi {

// =================================================================================
// (1) sampling: replaces curand_normal in PointMassModelGpu::step
//     (src/point_mass_gpu.cu:85-90).  HBM-write bound: 4*K*T*A bytes.
//     One thread owns a quad of samples (one float4 column of eps) and a block of
//     consecutive time steps; per (t,a) one Philox call -> one 16-byte store, 512 B per
//     warp and row.  No per-row branches: the action index is the unrolled inner loop.
// =================================================================================
template <int A, int ROUNDS>
__global__ void __launch_bounds__(256)
sample_kernel(float *__restrict__ eps, size_t ld, int T, int t_per_cta,
              const CtlDev *__restrict__ ctl, unsigned long long k_offset,
              const __grid_constant__ SamplerParams sp, int use_step_override,
              unsigned long long step_override)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // local quad
    if (4 * q >= ld) return;
    const unsigned long long step = use_step_override ? step_override : ctl->step;
    const uint32_t qg = (uint32_t)((k_offset >> 2) + q);                 // global quad
    const int t0 = blockIdx.y * t_per_cta;
    const int t1 = min(T, t0 + t_per_cta);
    float *p = eps + (size_t)t0 * A * ld + 4 * q;
    uint32_t r = (uint32_t)(t0 * A);
#pragma unroll 2
    for (int t = t0; t < t1; ++t) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            stg_f4(p, sample4<ROUNDS>(qg, r, step, sp, sp.c[a]));
            p += ld;
            ++r;
        }
    }
}

// =================================================================================
// (2) fused rollout: PointMassModelGpu::run/step (src/point_mass_gpu.cu:82-121) with
//     Cost::step_cost / final_cost (src/cost.cu:42-64) accumulated in registers.
//     One thread integrates FOUR consecutive samples (one float4 of every eps row), U[t]
//     and U[t]*inv_s are staged once per CTA in shared memory, the trajectory is never
//     written.  Epilogue: warp-shuffle + CTA min of the packed (cost, index) key and one
//     atomicMin per CTA (replaces min_red, src/point_mass.cu:533-575).
//
//     Arithmetic.  The gains are structurally {1,dt,0,1}/{dt^2/2,dt} (src/point_mass.cu:46-51)
//     so g0*p == p and g2*p + g3*v == v exactly.  STRICT rounds every product and sum
//     separately in source order (== the reference's host build, bit-for-bit vs
//     oracle ORACLE_ARITH_STRICT); otherwise the fused operations are exactly those nvcc
//     -fmad=true forms for the reference's device build (== ORACLE_ARITH_FMA).
// =================================================================================
// One thread integrates SPT consecutive samples.  eps is consumed in chunks of CH time
// steps through a register double buffer: the loads of chunk c+1 are issued before the
// arithmetic of chunk c, so every thread keeps CH*A vector loads (96 B for A=3) in flight.
// SPT >= 2: the samples advance in pairs on the packed FP32x2 path (PointMass2).
template <int A, class MODEL, bool FUSED, int SPT, int ROUNDS = 10>
__global__ void __launch_bounds__(256, (SPT == 1 ? 3 : 2))
rollout_kernel(float *__restrict__ eps, size_t ld, long long k_local, int T,
               const float *__restrict__ U, const ProblemDev *__restrict__ prob,
               float *__restrict__ S, CtlDev *__restrict__ ctl, unsigned long long k_offset,
               const __grid_constant__ SamplerParams sp)
{
    static_assert(!FUSED || SPT == 4, "fused sampling works on Philox quads");
    constexpr int CH = 8 / SPT;
    constexpr bool PACKED = SPT >= 2;
    constexpr int NP = PACKED ? SPT / 2 : 1;                  // sample pairs per thread
    constexpr int UST = PACKED ? UStage2<A>::kStride : UStage<A>::kStride;
    extern __shared__ __align__(16) float smem_f[];          // [T][UST]
    __shared__ unsigned long long s_key[8];
    // the averaging kernel behind this one may be scheduled as soon as SMs have room for it; it
    // waits (griddepcontrol.wait) for this grid to finish before it reads anything
    asm volatile("griddepcontrol.launch_dependents;");

    for (int i = threadIdx.x; i < T * A; i += blockDim.x) {
        const float u = U[i];
        const float ui = __fmul_rn(u, prob->inv_s[i % A]);   // src/cost.cu:46
        const int t = i / A, a = i - t * A;
        if (PACKED) {
            float4 *rec = reinterpret_cast<float4 *>(smem_f + (size_t)t * UST) + a;
            *rec = make_float4(u, u, ui, ui);
        } else {
            smem_f[t * UST + 2 * a]     = u;
            smem_f[t * UST + 2 * a + 1] = ui;
        }
    }
    __syncthreads();

    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // group of SPT samples
    unsigned long long key = kMinKeyInit;

    if (SPT * g < ld) {
        // ---- per-thread state: scalar (SPT == 1) or packed pairs
        PointMass<A, MODEL> m1;
        PointMass2<A, MODEL> m2;
        float x1[2 * A], c1 = 0.0f;
        f2 x2[NP][2 * A], c2[NP];
        if constexpr (PACKED) {
            m2.load(prob);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                c2[j] = mk2(0.0f, 0.0f);
#pragma unroll
                for (int i = 0; i < 2 * A; ++i) x2[j][i] = mk2(prob->x0[i], prob->x0[i]);
            }
        } else {
            m1.load(prob);
#pragma unroll
            for (int i = 0; i < 2 * A; ++i) x1[i] = prob->x0[i];
        }
        float *ep = eps + SPT * g;

        // one time step for all SPT samples given the eps vectors e[a][j]
        auto advance = [&](int t, const float (&e)[A][SPT]) {
            if constexpr (PACKED) {
                f2 u[A], ui[A];
                UStage2<A>::fetch(smem_f, t, u, ui);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    f2 ej[A];
#pragma unroll
                    for (int a = 0; a < A; ++a) ej[a] = mk2(e[a][2 * j], e[a][(2 * j + 1) % SPT]);
                    m2.step(x2[j], c2[j], u, ui, ej);
                }
            } else {
                float u[A], ui[A], ej[A];
                UStage<A>::fetch(smem_f, t, u, ui);
#pragma unroll
                for (int a = 0; a < A; ++a) ej[a] = e[a][0];
                m1.step(x1, c1, u, ui, ej);
            }
        };

        if (FUSED) {
            const unsigned long long step = ctl->step;
            const uint32_t qg = (uint32_t)((k_offset >> 2) + g);
            uint32_t r = 0;
            float *pw = ep;
            // Software pipelined: the noise of step t+1 is drawn while step t is integrated (the
            // Philox rounds and Box-Muller do not depend on the state, only the dynamics do), so
            // the two dependent chains interleave inside one loop body.  -3 % on this kernel.
            float4 nn[A];
#pragma unroll
            for (int a = 0; a < A; ++a) nn[a] = sample4<ROUNDS>(qg, r++, step, sp, sp.c[a]);
#pragma unroll 2
            for (int t = 0; t < T; ++t) {
                float e[A][SPT];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const float4 n = nn[a];
                    stg_f4(pw, n);
                    pw += ld;
                    e[a][0] = n.x; e[a][1 % SPT] = n.y; e[a][2 % SPT] = n.z; e[a][3 % SPT] = n.w;
                }
#pragma unroll
                for (int a = 0; a < A; ++a) nn[a] = sample4<ROUNDS>(qg, r++, step, sp, sp.c[a]);   // one past T: unused
                advance(t, e);
            }
        } else {
            float bufA[CH][A][SPT], bufB[CH][A][SPT];
            const float *pl = ep;                      // next chunk to load
            auto load_full = [&](float (&buf)[CH][A][SPT]) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
#pragma unroll
                    for (int a = 0; a < A; ++a)
                        EpsVec<SPT>::load(pl + (size_t)(i * A + a) * ld, buf[i][a]);
                pl += (size_t)(CH * A) * ld;
            };
            auto load_guard = [&](float (&buf)[CH][A][SPT], int t0) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    if (t0 + i < T) {
#pragma unroll
                        for (int a = 0; a < A; ++a)
                            EpsVec<SPT>::load(pl + (size_t)(i * A + a) * ld, buf[i][a]);
                    }
                pl += (size_t)(CH * A) * ld;
            };
            auto run_full = [&](const float (&buf)[CH][A][SPT], int t0) {
#pragma unroll
                for (int i = 0; i < CH; ++i) advance(t0 + i, buf[i]);
            };
            auto run_guard = [&](const float (&buf)[CH][A][SPT], int t0) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    if (t0 + i < T) advance(t0 + i, buf[i]);
            };
            int t0 = 0;
            load_guard(bufA, 0);
            // steady state: chunks t0 (in bufA), t0+CH and t0+2CH are complete
#pragma unroll 1
            for (; t0 + 3 * CH <= T; t0 += 2 * CH) {
                load_full(bufB);
                run_full(bufA, t0);
                load_full(bufA);
                run_full(bufB, t0 + CH);
            }
#pragma unroll 1
            for (; t0 < T; t0 += 2 * CH) {
                load_guard(bufB, t0 + CH);
                run_guard(bufA, t0);
                load_guard(bufA, t0 + 2 * CH);
                run_guard(bufB, t0 + CH);
            }
        }
        // terminal cost on x[T] (charged on top of the last stage cost,
        // src/point_mass_gpu.cu:116)
        float c[SPT];
        if constexpr (PACKED) {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                c2[j] = add2(c2[j], m2.terminal_cost(x2[j], prob));
                un2(c2[j], c[2 * j], c[(2 * j + 1) % SPT]);
            }
        } else {
            c[0] = __fadd_rn(c1, m1.terminal_cost(x1, prob));
        }
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            S[SPT * g + j] = c[j];
            const long long k = (long long)(SPT * g) + j;
            if (k < k_local) {
                const unsigned long long kk =
                    ((unsigned long long)float_to_ordered(c[j]) << 32) |
                    (unsigned long long)(uint32_t)(k_offset + (unsigned long long)k);
                key = kk < key ? kk : key;
            }
        }
    }

    key = warp_min_u64(key);
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x < 32) {
        key = threadIdx.x < (blockDim.x >> 5) ? s_key[threadIdx.x] : kMinKeyInit;
        key = warp_min_u64(key);
        if (threadIdx.x == 0 && key != kMinKeyInit) atomicMin(&ctl->min_key, key);
    }
}

// ---------------------------------------------------------------------------------
// (2b) TMA-staged rollout.  Same arithmetic, different data path: a producer warp streams
//      [TT time steps x A rows x 256 samples] eps tiles into a shared-memory ring with
//      cp.async.bulk.tensor.2d; 256 consumer threads (one sample each) read their column
//      conflict-free.  No eps registers and no global-load latency on the dependent chain:
//      three CTAs fit per SM, and for small K (latency-bound) a step costs ~the FP32 chain.
// ---------------------------------------------------------------------------------
template <int A> struct RolloutTile {
    static constexpr int kSteps = (A == 1 ? 20 : A == 2 ? 10 : A == 3 ? 7 : 5);   // TT
    static constexpr int kRows  = kSteps * A;                                     // <= 21
};
constexpr int kRtStages = 3;

// W = samples per CTA slab (TMA box width): 256 for throughput, 128/64 when K is so small
// that 256-wide slabs would leave SMs idle (the step is then bound by the per-sample
// dependent chain, and fewer warps per SM sub-partition shorten it).
template <int A>
size_t rollout_tma_smem_bytes(int T, int W)
{
    return (size_t)kRtStages * RolloutTile<A>::kRows * W * sizeof(float) +
           (size_t)T * UStage<A>::kStride * sizeof(float) + 2 * kRtStages * sizeof(uint64_t) + 128;
}

template <int A, class MODEL, int W>
__global__ void __launch_bounds__(W + 32, (W == 256 ? 3 : W == 128 ? 6 : 8))
rollout_tma_kernel(const __grid_constant__ CUtensorMap tmap_eps, int nslab, long long k_local,
                   int T, const float *__restrict__ U, const ProblemDev *__restrict__ prob,
                   float *__restrict__ S, CtlDev *__restrict__ ctl, unsigned long long k_offset)
{
    constexpr int TT = RolloutTile<A>::kSteps;
    constexpr int ROWS = RolloutTile<A>::kRows;
    constexpr int UST = UStage<A>::kStride;
    constexpr int kRtConsumerWarps = W / 32;
    constexpr uint32_t kTileBytes = ROWS * W * sizeof(float);
    // declared aligned (TMA destinations need 128 B): no integer round-trip on the address, so
    // the accesses below stay shared-space LDS/STS instead of generic loads
    extern __shared__ __align__(1024) uint8_t base[];
    float *s_tile = reinterpret_cast<float *>(base);                          // [stage][ROWS][W]
    float *s_u = s_tile + (size_t)kRtStages * ROWS * W;                     // [T][UST]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(s_u + (size_t)T * UST);
    uint64_t *empty_bar = full_bar + kRtStages;
    __shared__ unsigned long long s_key[kRtConsumerWarps];
    asm volatile("griddepcontrol.launch_dependents;");        // see rollout_kernel

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < T * A; i += blockDim.x) {
        const float u = U[i];
        const int t = i / A, a = i - t * A;
        s_u[t * UST + 2 * a]     = u;
        s_u[t * UST + 2 * a + 1] = __fmul_rn(u, prob->inv_s[a]);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < kRtStages; ++st) {
            mbar_init(&full_bar[st], 1);
            mbar_init(&empty_bar[st], kRtConsumerWarps);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int ntile = (T + TT - 1) / TT;
    unsigned long long key = kMinKeyInit;

    if (warp == kRtConsumerWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_eps);
            int stage = 0;
            uint32_t phase = 0;
            for (int slab = blockIdx.x; slab < nslab; slab += gridDim.x)
                for (int tile = 0; tile < ntile; ++tile) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kTileBytes);
                    tma_load_2d(s_tile + (size_t)stage * ROWS * W, &tmap_eps, slab * W,
                                tile * ROWS, &full_bar[stage]);
                    if (++stage == kRtStages) { stage = 0; phase ^= 1; }
                }
        }
        __syncwarp();
    } else {
        PointMass<A, MODEL> m;
        m.load(prob);
        int stage = 0;
        uint32_t phase = 0;
        for (int slab = blockIdx.x; slab < nslab; slab += gridDim.x) {
            float x[2 * A], c = 0.0f;
#pragma unroll
            for (int i = 0; i < 2 * A; ++i) x[i] = prob->x0[i];
            for (int tile = 0; tile < ntile; ++tile) {
                mbar_wait(&full_bar[stage], phase);
                const float *col = s_tile + (size_t)stage * ROWS * W + threadIdx.x;
                const int t0 = tile * TT;
                if (t0 + TT <= T) {
#pragma unroll
                    for (int i = 0; i < TT; ++i) {
                        float u[A], ui[A], e[A];
                        UStage<A>::fetch(s_u, t0 + i, u, ui);
#pragma unroll
                        for (int a = 0; a < A; ++a) e[a] = col[(i * A + a) * W];
                        m.step(x, c, u, ui, e);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TT; ++i)
                        if (t0 + i < T) {
                            float u[A], ui[A], e[A];
                            UStage<A>::fetch(s_u, t0 + i, u, ui);
#pragma unroll
                            for (int a = 0; a < A; ++a) e[a] = col[(i * A + a) * W];
                            m.step(x, c, u, ui, e);
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == kRtStages) { stage = 0; phase ^= 1; }
            }
            c = __fadd_rn(c, m.terminal_cost(x, prob));        // src/point_mass_gpu.cu:116
            const long long k = (long long)slab * W + threadIdx.x;
            S[k] = c;
            if (k < k_local) {
                const unsigned long long kk =
                    ((unsigned long long)float_to_ordered(c) << 32) |
                    (unsigned long long)(uint32_t)(k_offset + (unsigned long long)k);
                key = kk < key ? kk : key;
            }
        }
        key = warp_min_u64(key);
        if (lane == 0) s_key[warp] = key;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        key = threadIdx.x < kRtConsumerWarps ? s_key[threadIdx.x] : kMinKeyInit;
        key = warp_min_u64(key);
        if (threadIdx.x == 0 && key != kMinKeyInit) atomicMin(&ctl->min_key, key);
    }
}

// =================================================================================
// (3) weights: exp_red + sum_red (src/point_mass.cu:510-531, :628-666) in one pass.
//     wt[k] = expf(-(1/lambda) * (S[k] - beta)); the CTA's eta partial is added to the
//     fixed-point accumulator acc[R] (integer atomics: exact, order independent).
// =================================================================================
__global__ void __launch_bounds__(256)
weights_kernel(const float *__restrict__ S, long long k_local, long long k_pad,
               const ProblemDev *__restrict__ prob, const CtlDev *__restrict__ ctl,
               float *__restrict__ wt, long long *__restrict__ acc_eta)
{
    __shared__ float s_sum[8];
    const float beta = ordered_to_float((uint32_t)(ctl->min_key >> 32));
    const float nil = prob->neg_inv_lambda;
    const long long k0 = 4ll * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 < k_pad) {
        const float4 s4 = *reinterpret_cast<const float4 *>(S + k0);
        w4.x = (k0 + 0 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(s4.x, beta))) : 0.0f;
        w4.y = (k0 + 1 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(s4.y, beta))) : 0.0f;
        w4.z = (k0 + 2 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(s4.z, beta))) : 0.0f;
        w4.w = (k0 + 3 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(s4.w, beta))) : 0.0f;
        *reinterpret_cast<float4 *>(wt + k0) = w4;
    }

    float s = (w4.x + w4.y) + (w4.z + w4.w);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 8 ? s_sum[threadIdx.x] : 0.0f;
        s = warp_sum(s);
        if (threadIdx.x == 0 && s != 0.0f) acc_add(acc_eta, s);
    }
}

// =================================================================================
// (4) weighted average  num[r] = sum_k wt[k] * eps[r][k]
//     (replaces the T x {update_act_kernel, sum_red_adim, copy_act} launches of
//     PointMassModel::update_act, src/point_mass.cu:384-480, :828-926).
//
//     GEMV-shaped and HBM-bound (4*K*T*A bytes read once, one FMA per 4 bytes), so no
//     tensor cores: a persistent, warp-specialised streaming reduction.  One CTA per SM;
//     warp 8 is the TMA producer: per tile one cp.async.bulk.tensor.2d of a
//     [kAvgTileR rows x 256 samples] eps box (40 KB) plus one 1 KB bulk copy of the matching
//     weights, landing on an mbarrier; kAvgStages tiles are in flight per SM.  Warps 0-7
//     consume: warp w owns the rows {w, w+8, ...} of every tile, each lane multiplies its 8
//     samples of the row by the weights held in registers, the warp shuffle-reduces and lane 0
//     accumulates the row sum in shared memory.  The tile -> CTA assignment is static and
//     every CTA's row sum is formed in a fixed order; the cross-CTA sum is taken in 64-bit
//     fixed point with integer atomics (acc[r] += round(partial * 2^30)), which is exact and
//     order independent: the result is bitwise reproducible run to run, and the same
//     accumulators are what a multi-GPU run all-reduces.
// =================================================================================
size_t average_smem_bytes(int R)
{
    size_t tiles = (size_t)kAvgStages * kAvgTileR * kAvgTileK * sizeof(float);
    size_t wts   = (size_t)kAvgStages * kAvgTileK * sizeof(float);
    size_t rows  = (size_t)((R + 31) / 32 * 32) * sizeof(float);
    size_t bars  = 2 * kAvgStages * sizeof(uint64_t);
    return tiles + wts + rows + bars + 128;   // + alignment slack
}

// MERGE_W  : src is S; the producer warp turns each slab of costs into weights
//            (w~ = expf(-(1/lambda)(S-beta)), part 3) while the TMA tile is in flight and adds
//            the slab's eta once (in the CTA that owns the slab's first tile).
//            Otherwise src is the weights array written by weights_kernel.
// MERGE_FIN: the last CTA to finish (ticket from an atomic counter) applies the U update
//            (part 5); with K-shards over NVLink peer memory (xa.world > 1) it first runs the
//            single exchange of xchg.cuh in line -- the shard averaged relative to its own
//            minimum.  NCCL shards run their all-reduces between the kernels instead.
template <bool MERGE_W, bool MERGE_FIN>
__global__ void __launch_bounds__(kAvgThreads, 1)
average_kernel(const __grid_constant__ CUtensorMap tmap_eps, const float *__restrict__ src,
               long long *__restrict__ acc, int rows, int nslab, int nchunk, long long k_local,
               const ProblemDev *__restrict__ prob, CtlDev *__restrict__ ctl, FinalizeArgs fin,
               const __grid_constant__ XchgArgs xa)
{
    // declared aligned (TMA destinations need 128 B): no integer round-trip on the address, so
    // the accesses below stay shared-space LDS/STS instead of generic loads
    extern __shared__ __align__(1024) uint8_t base[];
    float *s_tile = reinterpret_cast<float *>(base);                               // [stage][R][K]
    float *s_wt   = s_tile + (size_t)kAvgStages * kAvgTileR * kAvgTileK;           // [stage][K]
    float *s_row  = s_wt + (size_t)kAvgStages * kAvgTileK;                         // [rows32]
    const int rows32 = (rows + 31) / 32 * 32;
    uint64_t *full_bar  = reinterpret_cast<uint64_t *>(s_row + rows32);
    uint64_t *empty_bar = full_bar + kAvgStages;
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < rows32; i += blockDim.x) s_row[i] = 0.0f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kAvgStages; ++s) {
            mbar_init(&full_bar[s], MERGE_W ? 2 : 1);     // TMA tx (+ the weights arrival)
            mbar_init(&empty_bar[s], kAvgConsumerWarps);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    // Programmatic dependent launch: when launched with the attribute (launch_average, pdl), this
    // grid may start while the rollout kernel in front of it is still running -- barrier
    // initialisation and the zeroing above overlap its tail -- and everything the rollout wrote
    // (eps, S, the min key) is only touched behind this point.  Without the attribute: a no-op.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // static, balanced partition of the slab-major tile list
    const long long ntiles = (long long)nslab * nchunk;
    const long long t_begin = ntiles * blockIdx.x / gridDim.x;
    const long long t_end   = ntiles * (blockIdx.x + 1) / gridDim.x;

    constexpr uint32_t kTileBytes = kAvgTileR * kAvgTileK * sizeof(float);
    constexpr uint32_t kWtBytes   = kAvgTileK * sizeof(float);

    if (warp == kAvgConsumerWarps) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) tma_prefetch_desc(&tmap_eps);
        float beta = 0.f, nil = 0.f, eta_part = 0.f;
        float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
        int cur_slab = -1;
        if (MERGE_W) {
            beta = ordered_to_float((uint32_t)(ctl->min_key >> 32));
            nil = prob->neg_inv_lambda;
        }
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const int slab  = (int)(t / nchunk);
            const int chunk = (int)(t - (long long)slab * nchunk);
            if (MERGE_W && slab != cur_slab) {
                // exp_red (src/point_mass.cu:518) for this lane's 8 samples of the slab
                const long long k0 = (long long)slab * kAvgTileK + 4 * lane;
                const float4 sa = *reinterpret_cast<const float4 *>(src + k0);
                const float4 sb = *reinterpret_cast<const float4 *>(src + k0 + 128);
                wa.x = (k0 + 0 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sa.x, beta))) : 0.f;
                wa.y = (k0 + 1 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sa.y, beta))) : 0.f;
                wa.z = (k0 + 2 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sa.z, beta))) : 0.f;
                wa.w = (k0 + 3 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sa.w, beta))) : 0.f;
                wb.x = (k0 + 128 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sb.x, beta))) : 0.f;
                wb.y = (k0 + 129 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sb.y, beta))) : 0.f;
                wb.z = (k0 + 130 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sb.z, beta))) : 0.f;
                wb.w = (k0 + 131 < k_local) ? expf(__fmul_rn(nil, __fsub_rn(sb.w, beta))) : 0.f;
                cur_slab = slab;
            }
            if (MERGE_W && chunk == 0)       // the slab's eta is counted by exactly one CTA
                eta_part += ((wa.x + wa.y) + (wa.z + wa.w)) + ((wb.x + wb.y) + (wb.z + wb.w));
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full_bar[stage], MERGE_W ? kTileBytes : kTileBytes + kWtBytes);
                tma_load_2d(s_tile + (size_t)stage * kAvgTileR * kAvgTileK, &tmap_eps,
                            slab * kAvgTileK, chunk * kAvgTileR, &full_bar[stage]);
                if (!MERGE_W)
                    bulk_load_1d(s_wt + (size_t)stage * kAvgTileK, src + (size_t)slab * kAvgTileK,
                                 kWtBytes, &full_bar[stage]);
            }
            if (MERGE_W) {
                float *tw = s_wt + (size_t)stage * kAvgTileK;
                *reinterpret_cast<float4 *>(tw + 4 * lane) = wa;
                *reinterpret_cast<float4 *>(tw + 128 + 4 * lane) = wb;
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
            }
            if (++stage == kAvgStages) { stage = 0; phase ^= 1; }
        }
        if (MERGE_W) {
            eta_part = warp_sum(eta_part);
            if (lane == 0 && eta_part != 0.0f) acc_add(acc + rows, eta_part);
        }
        __syncwarp();
    } else {
        // ------------------------------ consumers ---------------------------------
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const int slab  = (int)(t / nchunk);
            const int chunk = (int)(t - (long long)slab * nchunk);
            (void)slab;
            mbar_wait(&full_bar[stage], phase);
            const float *tw = s_wt + (size_t)stage * kAvgTileK;
            const float4 w0 = *reinterpret_cast<const float4 *>(tw + 4 * lane);
            const float4 w1 = *reinterpret_cast<const float4 *>(tw + 128 + 4 * lane);
            const float *tile = s_tile + (size_t)stage * kAvgTileR * kAvgTileK;
            float a_[kAvgTileR / kAvgConsumerWarps];
#pragma unroll
            for (int rr = 0; rr < kAvgTileR / kAvgConsumerWarps; ++rr) {
                const float *row = tile + (size_t)(warp + kAvgConsumerWarps * rr) * kAvgTileK;
                const float4 e0 = *reinterpret_cast<const float4 *>(row + 4 * lane);
                const float4 e1 = *reinterpret_cast<const float4 *>(row + 128 + 4 * lane);
                float a = e0.x * w0.x;
                a = fmaf(e0.y, w0.y, a);
                a = fmaf(e0.z, w0.z, a);
                a = fmaf(e0.w, w0.w, a);
                a = fmaf(e1.x, w1.x, a);
                a = fmaf(e1.y, w1.y, a);
                a = fmaf(e1.z, w1.z, a);
                a = fmaf(e1.w, w1.w, a);
                a_[rr] = a;
            }
            // all shared-memory reads of this stage are done: hand the slot back early
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
#pragma unroll
            for (int rr = 0; rr < kAvgTileR / kAvgConsumerWarps; ++rr) a_[rr] = warp_sum(a_[rr]);
            if (lane == 0) {
#pragma unroll
                for (int rr = 0; rr < kAvgTileR / kAvgConsumerWarps; ++rr) {
                    const int r = chunk * kAvgTileR + warp + kAvgConsumerWarps * rr;
                    if (r < rows) s_row[r] += a_[rr];
                }
            }
            if (++stage == kAvgStages) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    if (t_end > t_begin) {
        // only the row chunks this CTA's tiles touched: with few tiles per CTA (small K) the
        // other rows would be 64-bit atomics that add zero, all contending for the same words
        const long long n = t_end - t_begin;
        const int c_first = (int)(t_begin % nchunk);
        for (int r = threadIdx.x; r < rows; r += blockDim.x) {
            const int chunk = r / kAvgTileR;
            if (n >= nchunk || (chunk - c_first + nchunk) % nchunk < n) acc_add(acc + r, s_row[r]);
        }
    }

    if (MERGE_FIN) {
        __threadfence();                       // this CTA's atomics before its ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(&ctl->done, 1u);
            s_last = (ticket == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (xa.world > 1) {
                // the tile ring is free now: [rows+1] int64 accumulators, the rescale factors and the
                // shards' keys (xchg_merge_body), U_new
                long long *s_acc = reinterpret_cast<long long *>(s_tile);
                double *s_f = reinterpret_cast<double *>(s_acc + (rows + 1));
                float *s_un = reinterpret_cast<float *>(s_f + 2 * (kMaxWorld + 1));   // s_f, then the keys
                const volatile long long *vacc = acc;      // other CTAs' atomics: read at L2
                for (int i = threadIdx.x; i <= rows; i += blockDim.x) {
                    s_acc[i] = vacc[i];
                    acc[i] = 0;                            // re-armed for the next step
                }
                __syncthreads();
                if (!xchg_merge_body(s_acc, rows, prob, ctl, xa, s_f, (int)blockDim.x, 0)) {
                    if (threadIdx.x == 0) publish_comm_error(ctl, fin.next_act);
                    return;
                }
                finalize_body(s_acc, fin.U, fin.U_prev, prob, ctl, fin.next_act, fin.T, fin.A,
                              fin.flags, s_un);
            } else {
                finalize_body(acc, fin.U, fin.U_prev, prob, ctl, fin.next_act, fin.T, fin.A,
                              fin.flags, s_tile);      // the tile ring is free now: reuse it for U_new
            }
        }
    }
}

// =================================================================================
// (5) finalize: U += num/eta (copy_act, src/point_mass.cu:756-761) from the fixed-point
//     accumulators acc[0..R-1] = sum_k w~_k eps_k[r], acc[R] = eta (replaces the per-t
//     sum_red_adim folds, :668-741), next action = U[0,:] (:195), receding-horizon shift with
//     repeat-last / init-act re-initialisation (shift_act, :805-824), step counter advance,
//     accumulators and min key re-armed for the next step.  One small CTA.
//     Single shard: executed by the last CTA of average_kernel (MERGE_FIN) instead.
//     Multi-shard: acc has been all-reduced (int64 sum, exact) before this kernel.
// =================================================================================
constexpr int kFinThreads = 256;

__global__ void __launch_bounds__(kFinThreads)
finalize_kernel(long long *__restrict__ acc, float *__restrict__ U, float *__restrict__ U_prev,
                const ProblemDev *__restrict__ prob, CtlDev *__restrict__ ctl,
                float *__restrict__ next_act, int T, int A, unsigned flags)
{
    extern __shared__ float s_u[];          // U_new [T*A]
    finalize_body(acc, U, U_prev, prob, ctl, next_act, T, A, flags, s_u);
}

// =================================================================================
// K-shard exchange over NVLink peer memory (MPPI_COMM_P2P) as kernels of their own: the
// two-exchange flow of MPPI_FLAG_SPLIT_KERNELS (beta first, then the sums).  The default
// single-exchange merge runs inside the last CTA of average_kernel / step_kernel / tile_kernel
// (xchg.cuh, where the protocol and its double buffering are described).  A bounded spin (about two seconds)
// turns a dead peer into an error code instead of a hang; a failed exchange leaves U untouched.
// =================================================================================
constexpr int kXchgThreads = 512;

// (3a) beta = min over all shards of the packed (cost, index) key -- MPPI_FLAG_SPLIT_KERNELS only
__global__ void __launch_bounds__(32)
xchg_min_kernel(CtlDev *__restrict__ ctl, const __grid_constant__ XchgArgs xa)
{
    const int lane = threadIdx.x, rank = xa.rank, world = xa.world;
    const size_t sw = (size_t)xa.slot_words;
    const unsigned long long seq = ctl->step + 1;
    const int par = (int)(seq & 1ull);
    const unsigned long long mine = ctl->min_key;
    unsigned long long key = kMinKeyInit;
    if (lane < world) {
        unsigned long long *slot = mb_slot(xa.peers.mb[lane], par, world, rank, sw);   // my slot at peer
        st_relaxed_sys_u64(slot + 1, mine);
        __threadfence_system();
        st_release_sys_u64(slot + 0, seq);
        const unsigned long long *in = mb_slot(xa.peers.mb[rank], par, world, lane, sw);  // lane's slot here
        if (wait_seq(in + 0, seq)) key = ld_relaxed_sys_u64(in + 1);
        else atomicExch(&ctl->comm_error, 1u);
    }
    key = warp_min_u64(key);
    if (lane == 0) ctl->min_key = key;
}

// (5') all-reduce(sum) of the fixed-point accumulators through the mailboxes, then finalize
//      (the second exchange of the MPPI_FLAG_SPLIT_KERNELS flow)
__global__ void __launch_bounds__(kXchgThreads)
xchg_sum_finalize_kernel(long long *__restrict__ acc, float *__restrict__ U,
                         float *__restrict__ U_prev, const ProblemDev *__restrict__ prob,
                         CtlDev *__restrict__ ctl, float *__restrict__ next_act, int T, int A,
                         unsigned flags, const __grid_constant__ XchgArgs xa)
{
    extern __shared__ float s_u[];
    const int R = T * A, rank = xa.rank, world = xa.world;
    const size_t sw = (size_t)xa.slot_words;
    const unsigned long long seq = ctl->step + 1;
    const int par = (int)(seq & 1ull);
    // push my accumulators into every peer's mailbox (including my own)
    for (int r = 0; r < world; ++r) {
        unsigned long long *slot = mb_slot(xa.peers.mb[r], par, world, rank, sw) + kMailboxHeaderWords;
        for (int i = threadIdx.x; i <= R; i += blockDim.x)
            st_relaxed_sys_u64(slot + i, (unsigned long long)acc[i]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        st_release_sys_u64(mb_slot(xa.peers.mb[threadIdx.x], par, world, rank, sw) + 2, seq);
        const unsigned long long *in = mb_slot(xa.peers.mb[rank], par, world, threadIdx.x, sw);
        if (!wait_seq(in + 2, seq)) atomicExch(&ctl->comm_error, 1u);
    }
    __syncthreads();
    if (*reinterpret_cast<volatile unsigned int *>(&ctl->comm_error)) {
        if (threadIdx.x == 0) publish_comm_error(ctl, next_act);
        return;
    }
    // integer sum in rank order (exact: any order gives the same bits)
    unsigned long long *my = xa.peers.mb[rank] + (size_t)par * world * sw;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) {
        long long v[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            v[r] = r < world ? (long long)ld_relaxed_sys_u64(my + (size_t)r * sw + kMailboxHeaderWords + i) : 0ll;
        long long sum = 0;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r) sum += v[r];
        acc[i] = sum;
    }
    __threadfence();
    __syncthreads();
    finalize_body(acc, U, U_prev, prob, ctl, next_act, T, A, flags, s_u);
}

// =================================================================================
// layout conversion (parity taps): reference [K][R]  <->  internal [R][k_pad]
// =================================================================================
__global__ void __launch_bounds__(256)
to_internal_kernel(const float *__restrict__ e_ref, float *__restrict__ eps, long long k_local,
                   size_t ld, int R)
{
    __shared__ float tile[32][33];
    const long long kb = (long long)blockIdx.x * 32;
    const int rb = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const long long k = kb + j;
        const int r = rb + tx;
        tile[j][tx] = (k < k_local && r < R) ? e_ref[(size_t)k * R + r] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int r = rb + j;
        const long long k = kb + tx;
        if (r < R && k < k_local) eps[(size_t)r * ld + k] = tile[tx][j];
    }
}

__global__ void __launch_bounds__(256)
to_reference_kernel(const float *__restrict__ eps, float *__restrict__ e_ref, long long k_local,
                    size_t ld, int R)
{
    __shared__ float tile[32][33];
    const long long kb = (long long)blockIdx.x * 32;
    const int rb = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = rb + j;
        const long long k = kb + tx;
        tile[j][tx] = (r < R && k < k_local) ? eps[(size_t)r * ld + k] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long long k = kb + j;
        const int r = rb + tx;
        if (k < k_local && r < R) e_ref[(size_t)k * R + r] = tile[tx][j];
    }
}

// =================================================================================
// debug taps for get_inf (src/point_mass.cu:236-262)
// =================================================================================
// weights_kernel of the reference with its double literals (src/point_mass.cu:751)
__global__ void __launch_bounds__(256)
norm_weights_kernel(const float *__restrict__ S, long long k_local, float lambda, float beta,
                    float eta, float *__restrict__ w_out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_local) return;
    const double inv_eta = 1.0 / (double)eta;
    const double nil = -(1.0 / (double)lambda);
    const float diff = __fsub_rn(S[k], beta);
    const float arg = (float)(nil * (double)diff);
    w_out[k] = (float)(inv_eta * (double)expf(arg));
}

template <int A, class MODEL>
__global__ void __launch_bounds__(128)
trajectories_kernel(const float *__restrict__ eps, size_t ld, long long k_local, int T,
                    const float *__restrict__ U, const ProblemDev *__restrict__ prob,
                    float *__restrict__ x_out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_local) return;
    PointMass<A, MODEL> m;
    m.load(prob);
    float x[2 * A], c = 0.0f;
#pragma unroll
    for (int i = 0; i < 2 * A; ++i) x[i] = prob->x0[i];
    float *xo = x_out + (size_t)k * (T + 1) * 2 * A;
#pragma unroll
    for (int i = 0; i < 2 * A; ++i) xo[i] = x[i];
    for (int t = 0; t < T; ++t) {
        float u[A], ui[A], e[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            u[a] = U[t * A + a];
            ui[a] = __fmul_rn(u[a], prob->inv_s[a]);
            e[a] = eps[(size_t)(t * A + a) * ld + k];
        }
        m.step(x, c, u, ui, e);
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) xo[(size_t)(t + 1) * 2 * A + i] = x[i];
    }
}

__global__ void clear_ctl_kernel(CtlDev *ctl)
{
    for (int i = 0; i < 4; ++i) ctl->t_xchg[i] = 0;
    ctl->min_key = kMinKeyInit;
    ctl->last_key = kMinKeyInit;
    ctl->step = 0;
    ctl->eta = 0.0f;
    ctl->done = 0;
    ctl->comm_error = 0;
}

// =================================================================================
// launch wrappers
// =================================================================================
#define MPPI_DISPATCH_A(A_, ...)                                        \
    switch (A_) {                                                       \
        case 1: { constexpr int kA = 1; __VA_ARGS__; break; }           \
        case 2: { constexpr int kA = 2; __VA_ARGS__; break; }           \
        case 3: { constexpr int kA = 3; __VA_ARGS__; break; }           \
        case 4: { constexpr int kA = 4; __VA_ARGS__; break; }           \
        default: return cudaErrorInvalidValue;                          \
    }

// every model a handle can select: {strict, fma} x {DoubleIntegrator, LinearAxis}
#define MPPI_DISPATCH_MODEL(c_, ...)                                                     \
    do {                                                                                 \
        if ((c_).general_gains) {                                                        \
            if ((c_).strict) { using kM = Model<true, LinearAxis>; __VA_ARGS__; }        \
            else             { using kM = Model<false, LinearAxis>; __VA_ARGS__; }       \
        } else {                                                                         \
            if ((c_).strict) { using kM = Model<true, DoubleIntegrator>; __VA_ARGS__; }  \
            else             { using kM = Model<false, DoubleIntegrator>; __VA_ARGS__; } \
        }                                                                                \
    } while (0)
#define MPPI_FOR_EACH_MODEL(...)                                                         \
    do {                                                                                 \
        { using kM = Model<true, DoubleIntegrator>; __VA_ARGS__; }                       \
        { using kM = Model<false, DoubleIntegrator>; __VA_ARGS__; }                      \
        { using kM = Model<true, LinearAxis>; __VA_ARGS__; }                             \
        { using kM = Model<false, LinearAxis>; __VA_ARGS__; }                            \
    } while (0)

// CTA size of rollout_kernel.  A small shard is a handful of warps per SM, and what its rollout
// costs is the number of warps on the fullest SM sub-partition (the kernel is bound by the FMA
// pipe, one per sub-partition): with 256-thread CTAs a shard of 163 CTAs puts two of them -- four
// warps per sub-partition -- on 15 SMs and one on the others (K = 166 667: 0.279 ms per step against
// 0.227 with 128-thread CTAs).  Take the largest CTA that reaches the smallest such count.
// k_pad4: samples at four per thread (the fused kernel; SPT < 4 launches scale it).
static void rollout_shape(long long k_pad4, int num_sms, int *threads, long long *per_sched)
{
    const long long warps = (k_pad4 / 4 + 31) / 32;
    *per_sched = -1;
    *threads = 256;
    for (int t = 256; t >= 64; t >>= 1) {
        const long long w = t / 32, ctas = (warps + w - 1) / w;
        const long long per_sm = (ctas + num_sms - 1) / num_sms;
        const long long ps = (per_sm * w + 3) / 4;
        if (*per_sched < 0 || ps < *per_sched) { *per_sched = ps; *threads = t; }
    }
}
int rollout_block_threads(long long k_pad4, int num_sms)
{
    int t; long long ps;
    rollout_shape(k_pad4, num_sms, &t, &ps);
    return t;
}
double rollout_warps_per_sched(long long k_pad, int num_sms)
{
    int t; long long ps;
    rollout_shape(k_pad, num_sms, &t, &ps);
    // far beyond one resident wave (16 warps per SM) the waves average out
    const double avg = 1.05 * (double)((k_pad / 4 + 31) / 32) / (4.0 * num_sms);
    return ps > 8 ? (avg > 8.0 ? avg : 8.0) : (double)ps;
}

cudaError_t launch_sample(const LaunchCtx &c, float *eps, const CtlDev *ctl,
                          bool use_step_override, unsigned long long step_override)
{
    const size_t quads = (size_t)c.k_pad / 4;
    const unsigned gx = (unsigned)((quads + 255) / 256);
    // enough CTAs for >= ~8 per SM, at most 16 time steps per thread
    int tpc = (int)(((long long)c.horizon * gx) / (8ll * c.num_sms));
    tpc = tpc < 1 ? 1 : (tpc > 16 ? 16 : tpc);
    dim3 grid(gx, (unsigned)((c.horizon + tpc - 1) / tpc));
    if (c.philox_rounds == 7) {
        MPPI_DISPATCH_A(c.act_dim,
            (sample_kernel<kA, 7><<<grid, 256, 0, c.stream>>>(eps, (size_t)c.k_pad, c.horizon, tpc, ctl,
                                                             (unsigned long long)c.k_offset, c.sampler,
                                                             use_step_override ? 1 : 0, step_override)));
    } else {
        MPPI_DISPATCH_A(c.act_dim,
            (sample_kernel<kA, 10><<<grid, 256, 0, c.stream>>>(eps, (size_t)c.k_pad, c.horizon, tpc, ctl,
                                                              (unsigned long long)c.k_offset, c.sampler,
                                                              use_step_override ? 1 : 0, step_override)));
    }
    return cudaGetLastError();
}

template <int A, class MODEL, bool FUSED, int SPT, int ROUNDS = 10>
static cudaError_t launch_rollout_t(const LaunchCtx &c, float *eps, const float *U,
                                    const ProblemDev *prob, float *S, CtlDev *ctl)
{
    const size_t groups = (size_t)c.k_pad / SPT;
    // CTA size: see rollout_block_threads()
    const int threads = rollout_block_threads(c.k_pad / SPT * 4, c.num_sms);
    const unsigned grid = (unsigned)((groups + threads - 1) / threads);
    const size_t smem = sizeof(float) * (size_t)c.horizon *
                        (SPT >= 2 ? UStage2<A>::kStride : UStage<A>::kStride);
    rollout_kernel<A, MODEL, FUSED, SPT, ROUNDS><<<grid, threads, smem, c.stream>>>(
        eps, (size_t)c.k_pad, (long long)c.k_local, c.horizon, U, prob, S, ctl,
        (unsigned long long)c.k_offset, c.sampler);
    return cudaGetLastError();
}

template <int A, class MODEL>
static cudaError_t launch_rollout_a(const LaunchCtx &c, float *eps, const float *U,
                                    const ProblemDev *prob, float *S, CtlDev *ctl, bool fused)
{
    if (fused)
        return c.philox_rounds == 7 ? launch_rollout_t<A, MODEL, true, 4, 7>(c, eps, U, prob, S, ctl)
                                    : launch_rollout_t<A, MODEL, true, 4>(c, eps, U, prob, S, ctl);
    switch (c.rollout_spt) {
        case 1: return launch_rollout_t<A, MODEL, false, 1>(c, eps, U, prob, S, ctl);
        case 2: return launch_rollout_t<A, MODEL, false, 2>(c, eps, U, prob, S, ctl);
        default: return launch_rollout_t<A, MODEL, false, 4>(c, eps, U, prob, S, ctl);
    }
}

cudaError_t launch_rollout(const LaunchCtx &c, float *eps, const float *U, const ProblemDev *prob,
                           float *S, CtlDev *ctl, bool fused)
{
    MPPI_DISPATCH_A(c.act_dim,
        MPPI_DISPATCH_MODEL(c, return launch_rollout_a<kA, kM>(c, eps, U, prob, S, ctl, fused)));
    return cudaSuccess;
}

template <int A, class MODEL, int W>
static cudaError_t launch_rollout_tma_w(const LaunchCtx &c, const CUtensorMap &tmap, const float *U,
                                        const ProblemDev *prob, float *S, CtlDev *ctl)
{
    const int nslab = (int)(c.k_pad / W);
    const int per_sm = (W == 256 ? 3 : W == 128 ? 6 : 8);
    const int grid = nslab < per_sm * c.num_sms ? nslab : per_sm * c.num_sms;
    rollout_tma_kernel<A, MODEL, W><<<grid, W + 32, rollout_tma_smem_bytes<A>(c.horizon, W), c.stream>>>(
        tmap, nslab, (long long)c.k_local, c.horizon, U, prob, S, ctl,
        (unsigned long long)c.k_offset);
    return cudaGetLastError();
}

template <int A, class MODEL>
static cudaError_t launch_rollout_tma_t(const LaunchCtx &c, const CUtensorMap &tmap, const float *U,
                                        const ProblemDev *prob, float *S, CtlDev *ctl)
{
    switch (c.rollout_tma_width) {
        case 64:  return launch_rollout_tma_w<A, MODEL, 64>(c, tmap, U, prob, S, ctl);
        case 128: return launch_rollout_tma_w<A, MODEL, 128>(c, tmap, U, prob, S, ctl);
        default:  return launch_rollout_tma_w<A, MODEL, 256>(c, tmap, U, prob, S, ctl);
    }
}

cudaError_t launch_rollout_tma(const LaunchCtx &c, const CUtensorMap &tmap, const float *U,
                               const ProblemDev *prob, float *S, CtlDev *ctl)
{
    MPPI_DISPATCH_A(c.act_dim,
        MPPI_DISPATCH_MODEL(c, return launch_rollout_tma_t<kA, kM>(c, tmap, U, prob, S, ctl)));
    return cudaSuccess;
}

int rollout_tma_rows(int A)
{
    switch (A) {
        case 1: return RolloutTile<1>::kRows;
        case 2: return RolloutTile<2>::kRows;
        case 3: return RolloutTile<3>::kRows;
        default: return RolloutTile<4>::kRows;
    }
}

cudaError_t launch_weights(const LaunchCtx &c, const float *S, const ProblemDev *prob,
                           const CtlDev *ctl, float *wt, long long *acc)
{
    weights_kernel<<<c.weights_blocks, 256, 0, c.stream>>>(S, (long long)c.k_local,
                                                           (long long)c.k_pad, prob, ctl, wt,
                                                           acc + c.rows);
    return cudaGetLastError();
}

cudaError_t launch_average(const LaunchCtx &c, const CUtensorMap &tmap_eps, const float *src,
                           long long *acc, bool merge_weights, bool merge_finalize,
                           const ProblemDev *prob, CtlDev *ctl, float *U, float *U_prev,
                           float *next_act, unsigned flags, const XchgArgs &xa, bool pdl)
{
    const size_t smem = average_smem_bytes(c.rows);
    const int nslab = (int)(c.k_pad / kAvgTileK);
    const int nchunk = (c.rows + kAvgTileR - 1) / kAvgTileR;
    FinalizeArgs fin{U, U_prev, next_act, c.horizon, c.act_dim, flags};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c.avg_grid);
    cfg.blockDim = dim3(kAvgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const long long k_local = (long long)c.k_local;
    const int rows = c.rows;
#define MPPI_AVG_LAUNCH(MW, MF)                                                                  \
    return cudaLaunchKernelEx(&cfg, average_kernel<MW, MF>, tmap_eps, src, acc, rows, nslab, nchunk, \
                              k_local, prob, ctl, fin, xa)
    if (merge_weights && merge_finalize) MPPI_AVG_LAUNCH(true, true);
    else if (merge_weights)              MPPI_AVG_LAUNCH(true, false);
    else if (merge_finalize)             MPPI_AVG_LAUNCH(false, true);
    else                                 MPPI_AVG_LAUNCH(false, false);
#undef MPPI_AVG_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_finalize(const LaunchCtx &c, long long *acc, float *U, float *U_prev,
                            const ProblemDev *prob, CtlDev *ctl, float *next_act, unsigned flags)
{
    const size_t smem = sizeof(float) * (size_t)c.rows;
    finalize_kernel<<<1, kFinThreads, smem, c.stream>>>(acc, U, U_prev, prob, ctl, next_act,
                                                        c.horizon, c.act_dim, flags);
    return cudaGetLastError();
}

XchgArgs make_xchg_args(unsigned long long *const *peer_mb, int rank, int world, int rows)
{
    XchgArgs xa{};
    xa.rank = rank;
    xa.world = world;
    xa.slot_words = (unsigned long long)mailbox_slot_words(rows);
    for (int r = 0; r < world && r < kMaxWorld; ++r) xa.peers.mb[r] = peer_mb ? peer_mb[r] : nullptr;
    return xa;
}

cudaError_t launch_xchg_min(const LaunchCtx &c, CtlDev *ctl, const XchgArgs &xa)
{
    xchg_min_kernel<<<1, 32, 0, c.stream>>>(ctl, xa);
    return cudaGetLastError();
}

cudaError_t launch_xchg_sum_finalize(const LaunchCtx &c, long long *acc, float *U, float *U_prev,
                                     const ProblemDev *prob, CtlDev *ctl, float *next_act,
                                     unsigned flags, const XchgArgs &xa)
{
    const size_t smem = sizeof(float) * (size_t)c.rows;
    xchg_sum_finalize_kernel<<<1, kXchgThreads, smem, c.stream>>>(acc, U, U_prev, prob, ctl, next_act,
                                                                  c.horizon, c.act_dim, flags, xa);
    return cudaGetLastError();
}

cudaError_t launch_to_internal(const LaunchCtx &c, const float *e_ref, float *eps)
{
    dim3 grid((unsigned)((c.k_local + 31) / 32), (unsigned)((c.rows + 31) / 32));
    to_internal_kernel<<<grid, 256, 0, c.stream>>>(e_ref, eps, (long long)c.k_local,
                                                   (size_t)c.k_pad, c.rows);
    return cudaGetLastError();
}

cudaError_t launch_to_reference(const LaunchCtx &c, const float *eps, float *e_ref)
{
    dim3 grid((unsigned)((c.k_local + 31) / 32), (unsigned)((c.rows + 31) / 32));
    to_reference_kernel<<<grid, 256, 0, c.stream>>>(eps, e_ref, (long long)c.k_local,
                                                    (size_t)c.k_pad, c.rows);
    return cudaGetLastError();
}

cudaError_t launch_norm_weights(const LaunchCtx &c, const float *S, float lambda, float beta,
                                float eta, float *w_out)
{
    const unsigned grid = (unsigned)((c.k_local + 255) / 256);
    norm_weights_kernel<<<grid, 256, 0, c.stream>>>(S, (long long)c.k_local, lambda, beta, eta,
                                                    w_out);
    return cudaGetLastError();
}

cudaError_t launch_trajectories(const LaunchCtx &c, const float *eps, const float *U_prev,
                                const ProblemDev *prob, float *x_out)
{
    const unsigned grid = (unsigned)((c.k_local + 127) / 128);
    MPPI_DISPATCH_A(c.act_dim,
        MPPI_DISPATCH_MODEL(c, (trajectories_kernel<kA, kM><<<grid, 128, 0, c.stream>>>(
            eps, (size_t)c.k_pad, (long long)c.k_local, c.horizon, U_prev, prob, x_out))));
    return cudaGetLastError();
}

// Opt-in to large dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is a
// property of the FUNCTION on the device, not of a handle: it is always set to the device's
// opt-in maximum, never to what one handle needs, so that controllers with different horizons can
// live side by side in one process (a smaller value set by a later handle would make an earlier
// handle's next launch fail).  What a handle needs is validated against that maximum
// (check_smem_requirements).
constexpr int kSmemOptIn = 227 * 1024;      // sm_100: 232448 bytes per block, static + dynamic
constexpr int kSmemStaticSlack = 256;       // the kernels' few static words (keys, flags)

template <class K>
static cudaError_t opt_in(K kernel)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemOptIn - (int)fa.sharedSizeBytes);
}

template <int A, class MODEL>
static cudaError_t configure_rollout_s()
{
    cudaError_t e;
    if ((e = opt_in(rollout_kernel<A, MODEL, true, 4>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_kernel<A, MODEL, true, 4, 7>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_kernel<A, MODEL, false, 4>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_kernel<A, MODEL, false, 2>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_kernel<A, MODEL, false, 1>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_tma_kernel<A, MODEL, 64>)) != cudaSuccess) return e;
    if ((e = opt_in(rollout_tma_kernel<A, MODEL, 128>)) != cudaSuccess) return e;
    return opt_in(rollout_tma_kernel<A, MODEL, 256>);
}
template <int A>
static cudaError_t configure_rollout()
{
    cudaError_t e = cudaSuccess;
    MPPI_FOR_EACH_MODEL(if ((e = configure_rollout_s<A, kM>()) != cudaSuccess) return e);
    return e;
}

cudaError_t configure_kernels(const LaunchCtx &c)
{
    cudaError_t e = cudaSuccess;
    if ((e = opt_in(average_kernel<true, true>)) != cudaSuccess) return e;
    if ((e = opt_in(average_kernel<true, false>)) != cudaSuccess) return e;
    if ((e = opt_in(average_kernel<false, true>)) != cudaSuccess) return e;
    if ((e = opt_in(average_kernel<false, false>)) != cudaSuccess) return e;
    if ((e = opt_in(finalize_kernel)) != cudaSuccess) return e;
    if ((e = opt_in(xchg_sum_finalize_kernel)) != cudaSuccess) return e;
    MPPI_DISPATCH_A(c.act_dim, e = configure_rollout<kA>());
    return e;
}

// what this shape needs of every kernel it may launch, against the opt-in maximum; returns the
// name of the first kernel that does not fit (nullptr: all fit)
const char *check_smem_requirements(const LaunchCtx &c, size_t *need, size_t *have)
{
    *have = (size_t)(kSmemOptIn - kSmemStaticSlack);
    struct { const char *name; size_t bytes; } req[] = {
        {"average_kernel", average_smem_bytes(c.rows)},
        {"finalize_kernel", sizeof(float) * (size_t)c.rows},
        {"rollout_kernel", sizeof(float) * (size_t)c.horizon * 16},
        {"rollout_tma_kernel",
         c.act_dim == 1 ? rollout_tma_smem_bytes<1>(c.horizon, 256) :
         c.act_dim == 2 ? rollout_tma_smem_bytes<2>(c.horizon, 256) :
         c.act_dim == 3 ? rollout_tma_smem_bytes<3>(c.horizon, 256) :
                          rollout_tma_smem_bytes<4>(c.horizon, 256)},
    };
    for (auto &r : req)
        if (r.bytes > *have) { *need = r.bytes; return r.name; }
    return nullptr;
}

cudaError_t launch_clear_ctl(const LaunchCtx &c, CtlDev *ctl)
{
    clear_ctl_kernel<<<1, 1, 0, c.stream>>>(ctl);
    return cudaGetLastError();
}

}  // namespace mppi
