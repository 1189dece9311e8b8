// step.cu -- the whole control step as ONE persistent, warp-specialised sm_100a kernel.
//
// Why.  The fused sample+rollout pass (parts 1+2) is bound by instruction issue (Philox
// integer rounds, Box-Muller on the MUFU pipe, the packed FP32x2 dynamics) and leaves HBM
// half idle; the weighted average (parts 3+4) is bound by HBM reads and leaves the issue slots
// three quarters idle.  Run back to back they cost the sum.  Here both run on every SM at the
// same time: the rollout warps of a CTA write their eps rows, and as soon as a warp has
// finished the horizon of its 128 samples -- so that their costs, hence their weights, are
// known -- the CTA's TMA producer streams those rows back into shared memory where four
// consumer warps fold them into the weighted sums.  The reference needs ~3T+8 launches for
// the same work (PointMassModel::get_act, src/point_mass.cu:129-203).
//
// Roles inside a CTA (one CTA per SM):
//   warps [0, NR)        rollout: per tile of 128 samples (one Philox quad per lane) sample
//                        eps, store it, integrate, store S, atomicMin the packed (S, k) key,
//                        then publish "my round n is complete" in shared memory.
//   warp  NR             producer: waits, in a FIXED order, for the tiles of its own CTA,
//                        turns the tile's costs into weights relative to the CTA's running
//                        minimum `ref` (online softmax: when a tile lowers ref, the weighted
//                        sums gathered so far are rescaled by exp(-(ref_old-ref_new)/lambda)),
//                        and issues one cp.async.bulk.tensor.2d per [40 rows x 128 samples] box.
//   warps [NR+1, NR+5)   consumers: warp c owns rows {c, c+4, ...} of every box; each lane
//                        keeps its own partial of every row it owns in shared memory
//                        (s_part[row][lane]) -- no shuffles in the steady state.
// Tile -> (CTA, warp, round) is a static function of the tile index and the consumption order
// is fixed, so every CTA's sums are formed in the same order on every run: results are
// bitwise reproducible.  At the end each CTA writes {ref, eta, row sums} to a record; the
// last CTA (ticket) merges the records in CTA order, rescaling each by exp(-(ref_c-beta)/lambda),
// converts to the fixed-point accumulators and applies the U update (part 5).
//
// Registers: the CTA is launched with 96 per thread (640 threads); the consumer warpgroup
// hands back all but 32 of its registers (setmaxnreg.dec) and the four warpgroups holding the
// rollout warps and the producer grow to 112 (setmaxnreg.inc) -- the packed FP32x2 rollout with
// three Philox streams in flight needs ~112 to stay free of spills.
//
// Nothing ever waits on a consumer or producer except those two roles themselves, and the
// rollout warps wait on nobody: the dependency graph has no cycle.
#include "finalize.cuh"
#include "kernels.cuh"
#include "model.cuh"
#include "philox.cuh"

namespace mppi {

constexpr int kStTileK     = 128;   // samples per rollout tile: one warp, four samples per lane
constexpr int kStTileR     = 40;    // eps rows per TMA box
constexpr int kStStages    = 6;     // boxes in flight per SM (6 x 20 KB)
constexpr int kStConsumers = 4;     // consumer warps
constexpr int kStRowsPerWarp = kStTileR / kStConsumers;

struct StepStageHdr {
    int   chunk;     // row box index of the tile in this stage; < 0: no more work
    float scale;     // != 1: rescale the partial sums before adding this tile (chunk 0 only)
    int   pad_[2];
};

struct StepSmemLayout {
    size_t tile, wt, part, u, hdr, bars, done, misc, total;
};

__host__ __device__ inline StepSmemLayout step_smem_layout(int T, int A, int nr)
{
    const int R = T * A;
    const int nchunk = (R + kStTileR - 1) / kStTileR;
    StepSmemLayout l;
    size_t o = 0;
    l.tile = o; o += (size_t)kStStages * kStTileR * kStTileK * sizeof(float);
    l.wt   = o; o += (size_t)kStStages * kStTileK * sizeof(float);
    l.part = o; o += (size_t)nchunk * kStTileR * 32 * sizeof(float);
    l.u    = o; o += (size_t)T * 4 * A * sizeof(float);
    l.hdr  = o; o += (size_t)kStStages * sizeof(StepStageHdr);
    l.bars = o; o += (size_t)2 * kStStages * sizeof(uint64_t);
    l.done = o; o += (size_t)((nr + 3) / 4 * 4) * sizeof(unsigned int);
    l.misc = o; o += 16;
    l.total = o;
    return l;
}

#ifdef MPPI_STEP_TRACE
// development aid: globaltimer stamps of CTA 0..3 -- [cta][0]: kernel start, [cta][1+w*8+round]:
// rollout warp w finished round, [cta][200]: producer issued its last box, [cta][201]: consumers
// done (after the CTA barrier), [cta][202]: record written, [cta][203]: finalize done
__device__ unsigned long long g_step_trace[4][256];
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STEP_TRACE(slot) do { if (blockIdx.x < 4) g_step_trace[blockIdx.x][slot] = gtimer(); } while (0)
#else
#define STEP_TRACE(slot) do { } while (0)
#endif

template <uint32_t N> __device__ __forceinline__ void setmaxnreg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N));
}
template <uint32_t N> __device__ __forceinline__ void setmaxnreg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N));
}
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

template <int A, bool STRICT, int NR>
__global__ void __launch_bounds__((NR + kStConsumers + 1) * 32, 1)
step_kernel(const __grid_constant__ CUtensorMap tmap_eps, float *__restrict__ eps, size_t ld,
            long long k_local, int T, const float *U,
            const ProblemDev *__restrict__ prob, float *__restrict__ S, CtlDev *__restrict__ ctl,
            unsigned long long k_offset, const __grid_constant__ SamplerParams sp,
            float *__restrict__ part, long long *__restrict__ acc, FinalizeArgs fin)
{
    static_assert((NR + 1) % 4 == 0, "rollout warps + producer must fill whole warpgroups");
    constexpr int kThreads = (NR + kStConsumers + 1) * 32;
    constexpr int kEpiThreads = (NR + 1) * 32;      // rollout + producer warps run the epilogue
    constexpr uint32_t kBoxBytes = kStTileR * kStTileK * sizeof(float);
    const int R = T * A;
    const int nchunk = (R + kStTileR - 1) / kStTileR;

    // declared 1024-byte aligned (TMA destinations need 128): no integer round-trip on the
    // address, so every access below stays a shared-space LDS/STS
    extern __shared__ __align__(1024) uint8_t base[];
    const StepSmemLayout L = step_smem_layout(T, A, NR);
    float *s_tile = reinterpret_cast<float *>(base + L.tile);          // [stage][40][128]
    float *s_wt   = reinterpret_cast<float *>(base + L.wt);            // [stage][128]
    float *s_part = reinterpret_cast<float *>(base + L.part);          // [nchunk*40][32]
    float *s_u    = reinterpret_cast<float *>(base + L.u);             // [T][4A]
    StepStageHdr *s_hdr = reinterpret_cast<StepStageHdr *>(base + L.hdr);
    uint64_t *full_bar  = reinterpret_cast<uint64_t *>(base + L.bars);
    uint64_t *empty_bar = full_bar + kStStages;
    unsigned int *s_done = reinterpret_cast<unsigned int *>(base + L.done);   // [NR] rounds finished
    float *s_misc = reinterpret_cast<float *>(base + L.misc);          // {ref, eta, last-CTA flag}
    int *s_last = reinterpret_cast<int *>(s_misc + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) STEP_TRACE(0);

    for (int i = threadIdx.x; i < R; i += kThreads) {
        const float u = U[i];
        const float ui = __fmul_rn(u, prob->inv_s[i % A]);            // src/cost.cu:46
        reinterpret_cast<float4 *>(s_u)[i] = make_float4(u, u, ui, ui);
    }
    for (int i = threadIdx.x; i < nchunk * kStTileR * 32; i += kThreads) s_part[i] = 0.0f;
    if (threadIdx.x < NR) s_done[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kStConsumers);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    const long long ntiles = (long long)(ld / kStTileK);
    const long long nslots = (long long)NR * gridDim.x;

    if (warp <= NR) setmaxnreg_inc<112>();
    else            setmaxnreg_dec<32>();

    if (warp < NR) {
        // ================================ rollout ====================================
        PointMass2<A, STRICT> m2;
        m2.load(prob);
        const unsigned long long step = ctl->step;
        unsigned long long key = kMinKeyInit;
        unsigned int round = 0;
        for (long long tile = (long long)warp * gridDim.x + blockIdx.x; tile < ntiles;
             tile += nslots) {
            const size_t g = (size_t)tile * 32 + lane;                 // this lane's Philox quad
            f2 x2[2][2 * A], c2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                c2[j] = mk2(0.0f, 0.0f);
#pragma unroll
                for (int i = 0; i < 2 * A; ++i) x2[j][i] = mk2(prob->x0[i], prob->x0[i]);
            }
            const uint32_t qg = (uint32_t)((k_offset >> 2) + g);
            float *pw = eps + 4 * g;
            uint32_t r = 0;
#pragma unroll 2
            for (int t = 0; t < T; ++t) {
                f2 e[2][A];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const float4 n = sample4(qg, r, step, sp, sp.c[a]);
                    stg_f4(pw, n);
                    pw += ld;
                    e[0][a] = mk2(n.x, n.y);
                    e[1][a] = mk2(n.z, n.w);
                    ++r;
                }
                f2 u[A], ui[A];
                UStage2<A>::fetch(s_u, t, u, ui);
                m2.step(x2[0], c2[0], u, ui, e[0]);
                m2.step(x2[1], c2[1], u, ui, e[1]);
            }
            // terminal cost on x[T], charged on top of the last stage cost
            // (src/point_mass_gpu.cu:116)
            float c[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                c2[j] = add2(c2[j], m2.state_cost(x2[j], mk2(0.0f, 0.0f)));
                un2(c2[j], c[2 * j], c[2 * j + 1]);
            }
            *reinterpret_cast<float4 *>(S + 4 * g) = make_float4(c[0], c[1], c[2], c[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long k = (long long)(4 * g) + j;
                if (k < k_local) {
                    const unsigned long long kk =
                        ((unsigned long long)float_to_ordered(c[j]) << 32) |
                        (unsigned long long)(uint32_t)(k_offset + (unsigned long long)k);
                    key = kk < key ? kk : key;
                }
            }
            // eps and S of this tile -> visible to the producer's TMA loads and its S read
            __threadfence();
            fence_proxy_async_all();
            __syncwarp();
            ++round;
            if (lane == 0) {
                st_release_cta_shared_u32(&s_done[warp], round);
                STEP_TRACE(1 + warp * 8 + (round - 1));
            }
        }
        key = warp_min_u64(key);
        if (lane == 0 && key != kMinKeyInit) atomicMin(&ctl->min_key, key);
    } else if (warp == NR) {
        // ================================ producer ===================================
        if (lane == 0) tma_prefetch_desc(&tmap_eps);
        const float nil = prob->neg_inv_lambda;
        float ref = __int_as_float(0x7f800000);        // +inf: no sample seen yet
        float eta_part = 0.0f;                          // this lane's share of eta, relative to ref
        int stage = 0;
        uint32_t phase = 0;
        for (unsigned int round = 0; (long long)round * nslots + blockIdx.x < ntiles; ++round) {
            for (int w = 0; w < NR; ++w) {
                const long long tile = (long long)round * nslots + (long long)w * gridDim.x + blockIdx.x;
                if (tile >= ntiles) break;
                if (lane == 0)
                    while (ld_acquire_cta_shared_u32(&s_done[w]) < round + 1) __nanosleep(256);
                __syncwarp();
                // exp_red (src/point_mass.cu:518) for this lane's four samples of the tile
                const long long k0 = tile * kStTileK + 4 * lane;
                const float4 s4 = __ldcg(reinterpret_cast<const float4 *>(S + k0));
                const bool v0 = k0 + 0 < k_local, v1 = k0 + 1 < k_local, v2 = k0 + 2 < k_local,
                           v3 = k0 + 3 < k_local;
                const float inf = __int_as_float(0x7f800000);
                float tmin = fminf(fminf(v0 ? s4.x : inf, v1 ? s4.y : inf),
                                   fminf(v2 ? s4.z : inf, v3 ? s4.w : inf));
                tmin = warp_min_f(tmin);
                float scale = 1.0f;
                if (tmin < ref) {
                    scale = expf(__fmul_rn(nil, __fsub_rn(ref, tmin)));   // ref = +inf -> 0
                    eta_part *= scale;
                    ref = tmin;
                }
                float4 w4;
                w4.x = v0 ? expf(__fmul_rn(nil, __fsub_rn(s4.x, ref))) : 0.0f;
                w4.y = v1 ? expf(__fmul_rn(nil, __fsub_rn(s4.y, ref))) : 0.0f;
                w4.z = v2 ? expf(__fmul_rn(nil, __fsub_rn(s4.z, ref))) : 0.0f;
                w4.w = v3 ? expf(__fmul_rn(nil, __fsub_rn(s4.w, ref))) : 0.0f;
                eta_part += (w4.x + w4.y) + (w4.z + w4.w);
                for (int chunk = 0; chunk < nchunk; ++chunk) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    *reinterpret_cast<float4 *>(s_wt + stage * kStTileK + 4 * lane) = w4;
                    if (lane == 0) {
                        s_hdr[stage].chunk = chunk;
                        s_hdr[stage].scale = chunk == 0 ? scale : 1.0f;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full_bar[stage], kBoxBytes);
                        tma_load_2d(s_tile + (size_t)stage * kStTileR * kStTileK, &tmap_eps,
                                    (int)(tile * kStTileK), chunk * kStTileR, &full_bar[stage]);
                    }
                    if (++stage == kStStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        // no more tiles: release the consumers
        if (lane == 0) STEP_TRACE(200);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
            s_hdr[stage].chunk = -1;
            s_hdr[stage].scale = 1.0f;
            mbar_arrive(&full_bar[stage]);
        }
        eta_part = warp_sum(eta_part);
        if (lane == 0) { s_misc[0] = ref; s_misc[1] = eta_part; }
    } else {
        // ================================ consumers ==================================
        const int cw = warp - (NR + 1);
        int stage = 0;
        uint32_t phase = 0;
        for (;;) {
            mbar_wait(&full_bar[stage], phase);
            const int chunk = s_hdr[stage].chunk;
            const float scale = s_hdr[stage].scale;
            if (chunk < 0) break;
            if (scale != 1.0f) {
                // a new minimum: everything gathered so far is relative to the old one
                for (int cc = 0; cc < nchunk; ++cc)
#pragma unroll
                    for (int rr = 0; rr < kStRowsPerWarp; ++rr) {
                        float *p = s_part + (size_t)(cc * kStTileR + cw + kStConsumers * rr) * 32 + lane;
                        *p = *p * scale;
                    }
            }
            const float4 w4 = *reinterpret_cast<const float4 *>(s_wt + stage * kStTileK + 4 * lane);
            const float *tile = s_tile + (size_t)stage * kStTileR * kStTileK + 4 * lane;
            float *pp = s_part + (size_t)(chunk * kStTileR + cw) * 32 + lane;
            // two batches of rows: the consumer warps live on 32 registers
            constexpr int kHalf = kStRowsPerWarp / 2;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                float a_[kHalf];
                float4 e_[kHalf];
#pragma unroll
                for (int i = 0; i < kHalf; ++i) {
                    const int rr = hb * kHalf + i;
                    e_[i] = *reinterpret_cast<const float4 *>(tile + (size_t)(cw + kStConsumers * rr) * kStTileK);
                    a_[i] = pp[(size_t)rr * kStConsumers * 32];
                }
                if (hb == 1) {
                    // every read of the stage's box is in registers: hand the slot back early
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                }
#pragma unroll
                for (int i = 0; i < kHalf; ++i) {
                    const int rr = hb * kHalf + i;
                    float a = a_[i];
                    a = fmaf(e_[i].x, w4.x, a);
                    a = fmaf(e_[i].y, w4.y, a);
                    a = fmaf(e_[i].z, w4.z, a);
                    a = fmaf(e_[i].w, w4.w, a);
                    pp[(size_t)rr * kStConsumers * 32] = a;
                }
            }
            if (++stage == kStStages) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    if (warp > NR) return;                 // consumers are done; the other warps finish the step
    if (threadIdx.x == 0) STEP_TRACE(201);

    // ---- this CTA's record: {ref, eta, row sums}, all relative to ref
    float *rec = part + (size_t)blockIdx.x * (R + 2);
    for (int r = warp; r < R; r += kEpiThreads / 32) {
        const float v = warp_sum(s_part[(size_t)r * 32 + lane]);
        if (lane == 0) rec[2 + r] = v;
    }
    if (threadIdx.x == 0) { rec[0] = s_misc[0]; rec[1] = s_misc[1]; }

    __threadfence();                       // the record and the min key before the ticket
    named_bar_sync(1, kEpiThreads);
    if (threadIdx.x == 0) {
        STEP_TRACE(202);
        const unsigned ticket = atomicAdd(&ctl->done, 1u);
        *s_last = (ticket == gridDim.x - 1);
    }
    named_bar_sync(1, kEpiThreads);
    if (*s_last) {
        // ---- merge the records in CTA order (deterministic), then part 5
        __threadfence();
        const float nil = prob->neg_inv_lambda;
        const unsigned long long mk = *reinterpret_cast<volatile unsigned long long *>(&ctl->min_key);
        const float beta = ordered_to_float((uint32_t)(mk >> 32));
        float *s_f = s_tile;                                           // [gridDim.x]
        for (int c = threadIdx.x; c < (int)gridDim.x; c += kEpiThreads) {
            const float ref_c = __ldcg(part + (size_t)c * (R + 2));
            s_f[c] = expf(__fmul_rn(nil, __fsub_rn(ref_c, beta)));     // ref_c = +inf -> 0
        }
        named_bar_sync(1, kEpiThreads);
        for (int i = threadIdx.x; i <= R; i += kEpiThreads) {
            const size_t off = i < R ? 2 + (size_t)i : 1;               // i == R: eta
            double s = 0.0;
            for (int c = 0; c < (int)gridDim.x; ++c)
                s += (double)__ldcg(part + (size_t)c * (R + 2) + off) * (double)s_f[c];
            acc[i] = __double2ll_rn(s * kAccScale);
        }
        __threadfence();
        named_bar_sync(1, kEpiThreads);
        finalize_body(acc, fin.U, fin.U_prev, prob, ctl, fin.next_act, fin.T, fin.A, fin.flags,
                      s_tile + 1024, kEpiThreads, 1);
#ifdef MPPI_STEP_TRACE
        if (threadIdx.x == 0) g_step_trace[0][203] = gtimer();
#endif
    }
}

#ifdef MPPI_STEP_TRACE
extern "C" int mppi_debug_read_step_trace(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, g_step_trace, sizeof(g_step_trace));
}
#endif

// =================================================================================
// launch wrappers
// =================================================================================
namespace {
constexpr int kStepNR = 15;   // rollout warps per CTA (+ 1 producer = 4 warpgroups)

template <int A, bool STRICT>
cudaError_t launch_step_t(const LaunchCtx &c, const CUtensorMap &tmap, float *eps, float *U,
                          const ProblemDev *prob, float *S, CtlDev *ctl, float *part,
                          long long *acc, const FinalizeArgs &fin)
{
    const size_t smem = step_smem_layout(c.horizon, c.act_dim, kStepNR).total;
    const long long ntiles = c.k_pad / kStTileK;
    const int grid = (int)(ntiles < c.num_sms ? ntiles : c.num_sms);
    step_kernel<A, STRICT, kStepNR><<<grid, (kStepNR + kStConsumers + 1) * 32, smem, c.stream>>>(
        tmap, eps, (size_t)c.k_pad, (long long)c.k_local, c.horizon, U, prob, S, ctl,
        (unsigned long long)c.k_offset, c.sampler, part, acc, fin);
    return cudaGetLastError();
}

template <int A>
cudaError_t configure_step_a(int smem)
{
    cudaError_t e = cudaFuncSetAttribute(step_kernel<A, false, kStepNR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(step_kernel<A, true, kStepNR>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}
}  // namespace

bool step_kernel_supported(int T, int A)
{
    return step_smem_layout(T, A, kStepNR).total <= 227 * 1024;
}

size_t step_part_floats(const LaunchCtx &c) { return (size_t)c.num_sms * ((size_t)c.rows + 2); }

cudaError_t configure_step(const LaunchCtx &c)
{
    if (!step_kernel_supported(c.horizon, c.act_dim)) return cudaSuccess;
    const int smem = (int)step_smem_layout(c.horizon, c.act_dim, kStepNR).total;
    switch (c.act_dim) {
        case 1: return configure_step_a<1>(smem);
        case 2: return configure_step_a<2>(smem);
        case 3: return configure_step_a<3>(smem);
        case 4: return configure_step_a<4>(smem);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_step(const LaunchCtx &c, const CUtensorMap &tmap, float *eps, float *U,
                        const ProblemDev *prob, float *S, CtlDev *ctl, float *part, long long *acc,
                        float *U_prev, float *next_act, unsigned flags)
{
    FinalizeArgs fin{U, U_prev, next_act, c.horizon, c.act_dim, flags};
#define MPPI_STEP_CASE(A_)                                                                       \
    case A_:                                                                                     \
        return c.strict ? launch_step_t<A_, true>(c, tmap, eps, U, prob, S, ctl, part, acc, fin) \
                        : launch_step_t<A_, false>(c, tmap, eps, U, prob, S, ctl, part, acc, fin)
    switch (c.act_dim) {
        MPPI_STEP_CASE(1);
        MPPI_STEP_CASE(2);
        MPPI_STEP_CASE(3);
        MPPI_STEP_CASE(4);
        default: return cudaErrorInvalidValue;
    }
#undef MPPI_STEP_CASE
}

}  // namespace mppi
