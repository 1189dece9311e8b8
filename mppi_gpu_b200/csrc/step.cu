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
// Roles inside a CTA (one CTA per SM, 20 warps):
//   warps [0, NR)        rollout: per tile of 128 samples (one Philox quad per lane) sample
//                        eps, store it, integrate, store S, atomicMin the packed (S, k) key,
//                        then flag the tile as complete in shared memory.  The warps pull
//                        tiles from the CTA's own list (an atomic counter in shared memory),
//                        so a sub-partition that also hosts the producer / a consumer warp
//                        simply takes fewer of them.
//   warp  NR             producer: waits, in list order, for the tiles of its own CTA, turns
//                        the tile's costs into weights relative to the CTA's running minimum
//                        `ref` (online softmax: when a tile lowers ref, the weighted sums
//                        gathered so far are rescaled by exp(-(ref_old-ref_new)/lambda)), and
//                        issues one cp.async.bulk.tensor.2d per [40 rows x 128 samples] box.
//   warps [NR+1, NR+5)   consumers: box (tile, chunk) belongs to warp chunk % 4, which has its
//                        own ring of stages -- one waiter and one releaser per mbarrier, in
//                        strict alternation with the producer (a shared ring would let a
//                        fast warp run two phases ahead of a barrier, which parity waits
//                        cannot tell apart) -- and the rows of a chunk are only ever touched
//                        by their owner.  Half-warp h takes the rows of parity h, lane
//                        q of it eight samples of the row; its partial of every row lives in
//                        shared memory (s_part[row][16]) -- no shuffles in the steady state.
// The tile list of a CTA is a static function of the tile index (tile = i * gridDim + cta) and
// the producer consumes it in list order whichever warp computed a tile, so every CTA's sums
// are formed in the same order on every run: results are bitwise reproducible.  At the end
// each CTA writes {ref, eta, row sums} to a record; the last CTA (ticket) merges the records
// in CTA order, rescaling each by exp(-(ref_c-beta)/lambda), converts to the fixed-point
// accumulators and applies the U update (part 5).
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
#include "merge.cuh"
#include "model.cuh"
#include "philox.cuh"

#include <stdlib.h>

namespace mppi {

constexpr int kStTileK     = 128;   // samples per rollout tile: one warp, four samples per lane
constexpr int kStTileR     = 40;    // eps rows per TMA box
constexpr int kStMaxStages = 8;     // boxes in flight per SM (8 x 20 KB): two per consumer warp;
                                    // 4 (one per consumer warp) when shared memory is short
constexpr int kStConsumers = 4;     // consumer warps

struct StepSmemLayout {
    size_t tile, wt, part, u, scale, bars, done, misc, total;
};

__host__ __device__ inline StepSmemLayout step_smem_layout(int T, int A, long long list_len,
                                                           int nstages)
{
    const int R = T * A;
    const int nchunk = (R + kStTileR - 1) / kStTileR;
    StepSmemLayout l;
    size_t o = 0;
    l.tile  = o; o += (size_t)nstages * kStTileR * kStTileK * sizeof(float);
    l.wt    = o; o += (size_t)nstages * kStTileK * sizeof(float);
    l.part  = o; o += (size_t)nchunk * kStTileR * 16 * sizeof(float);
    l.u     = o; o += (size_t)T * 4 * A * sizeof(float);
    l.scale = o; o += (size_t)kStMaxStages * sizeof(float);
    l.bars  = o; o += (size_t)2 * kStMaxStages * sizeof(uint64_t);
    l.done  = o; o += (size_t)((list_len + 3) / 4 * 4) * sizeof(unsigned int);
    l.misc  = o; o += 32;                            // {ref, eta, last, next} + the merge mbarrier
    l.total = o;
    return l;
}

// the merge of the per-CTA records reuses the ring / row-sum / U region: merged sums, U_new, the
// rescale factors and at least eight records per bulk-copy pass
__host__ __device__ inline size_t step_merge_bytes(int R, int grid)
{
    return merge_smem_bytes(R, grid);
}

#ifdef MPPI_STEP_TRACE
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// development aid (tools/step_trace.py): globaltimer stamps of CTA 0..3
//   g_step_trace[cta][0] kernel start, [200] producer issued its last box, [201] consumers done,
//   [202] record written, [0][203] finalize done
//   g_step_trace_li[cta][0][li] rollout of list entry li complete, [1][li] producer starts on it,
//   [2][li] producer has issued its last box
__device__ unsigned long long g_step_trace[4][256];
__device__ unsigned long long g_step_trace_li[4][3][128];
// every CTA: [0] start, [1] last rollout warp done, [2] producer done, [3] consumers done,
// [4] ticket taken; [0][5] merge begins (last CTA), [0][6] finalize done, [0][7] last CTA id
__device__ unsigned long long g_step_trace_all[160][8];
#define STEP_TRACE_ALL(slot) do { g_step_trace_all[blockIdx.x][slot] = gtimer(); } while (0)
#define STEP_TRACE(slot) do { if (blockIdx.x < 4) g_step_trace[blockIdx.x][slot] = gtimer(); } while (0)
#define STEP_TRACE_LI(j, li) \
    do { if (blockIdx.x < 4 && (li) < 128) g_step_trace_li[blockIdx.x][j][li] = gtimer(); } while (0)
#else
#define STEP_TRACE(slot) do { } while (0)
#define STEP_TRACE_ALL(slot) do { } while (0)
#define STEP_TRACE_LI(j, li) do { } while (0)
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

template <int A, class MODEL, int NR>
__global__ void __launch_bounds__((NR + kStConsumers + 1) * 32, 1)
step_kernel(const __grid_constant__ CUtensorMap tmap_eps, float *__restrict__ eps, size_t ld,
            long long k_local, int T, const float *U,
            const ProblemDev *__restrict__ prob, float *__restrict__ S, CtlDev *__restrict__ ctl,
            unsigned long long k_offset, const __grid_constant__ SamplerParams sp,
            float *__restrict__ part, FinalizeArgs fin, int nstages,
            const __grid_constant__ XchgArgs xa)
{
    static_assert((NR + 1) % 4 == 0, "rollout warps + producer must fill whole warpgroups");
    constexpr int kThreads = (NR + kStConsumers + 1) * 32;
    constexpr int kEpiThreads = (NR + 1) * 32;      // rollout + producer warps run the epilogue
    constexpr uint32_t kBoxBytes = kStTileR * kStTileK * sizeof(float);
    constexpr int kBoxFloats = kStTileR * kStTileK;
    const int R = T * A;
    const int nchunk = (R + kStTileR - 1) / kStTileR;

    // declared 1024-byte aligned (TMA destinations need 128): no integer round-trip on the
    // address, so every access below stays a shared-space LDS/STS
    extern __shared__ __align__(1024) uint8_t base[];
    const long long ntiles = (long long)(ld / kStTileK);
    const long long list_len = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // this CTA's tiles
    const StepSmemLayout L = step_smem_layout(T, A, (ntiles + gridDim.x - 1) / gridDim.x, nstages);
    float *s_tile = reinterpret_cast<float *>(base + L.tile);          // [stage][40][128]
    float *s_wt   = reinterpret_cast<float *>(base + L.wt);            // [stage][128]
    float *s_part = reinterpret_cast<float *>(base + L.part);          // [nchunk*40][16]
    float *s_u    = reinterpret_cast<float *>(base + L.u);             // [T][4A]
    float *s_scale = reinterpret_cast<float *>(base + L.scale);        // [stage] rescale of the box's tile
    uint64_t *full_bar  = reinterpret_cast<uint64_t *>(base + L.bars);
    uint64_t *empty_bar = full_bar + kStMaxStages;
    unsigned int *s_done = reinterpret_cast<unsigned int *>(base + L.done);   // [list] tile complete
    float *s_misc = reinterpret_cast<float *>(base + L.misc);          // {ref, eta, last-CTA flag, next}
    int *s_last = reinterpret_cast<int *>(s_misc + 2);
    unsigned int *s_next = reinterpret_cast<unsigned int *>(s_misc + 3);   // next list entry to roll out

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { STEP_TRACE(0); STEP_TRACE_ALL(0); }

    for (int i = threadIdx.x; i < R; i += kThreads) {
        const float u = U[i];
        const float ui = __fmul_rn(u, prob->inv_s[i % A]);            // src/cost.cu:46
        reinterpret_cast<float4 *>(s_u)[i] = make_float4(u, u, ui, ui);
    }
    for (int i = threadIdx.x; i < nchunk * kStTileR * 16; i += kThreads) s_part[i] = 0.0f;
    for (int i = threadIdx.x; i < (int)list_len; i += kThreads) s_done[i] = 0u;
    if (threadIdx.x == 0) {
        *s_next = (unsigned int)NR;
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    // The CTA is launched with 96 registers per thread.  setmaxnreg.inc can only draw on what the
    // CTA itself has handed back: the consumer warpgroup's 128 threads x (96 - 32) = 8192 registers
    // are exactly the 512 threads x (112 - 96) the four other warpgroups ask for.  Any pair that
    // frees less than it claims (e.g. 48 / 112) deadlocks the kernel in setmaxnreg.inc.
    // NR = 15: 640 threads launched with 96, rollout + producer grow to 112; NR = 11: 512 threads
    // launched with 128, rollout + producer grow to 160 (fewer warps, each with room to keep several
    // Philox streams in flight)
    constexpr uint32_t kRegLaunch = NR == 15 ? 96 : 128, kRegRoll = NR == 15 ? 112 : 160;
    static_assert(128 * (kRegLaunch - 32) == (NR + 1) * 32 * (kRegRoll - kRegLaunch),
                  "setmaxnreg: the pool must balance");
    if (warp <= NR) setmaxnreg_inc<kRegRoll>();
    else            setmaxnreg_dec<32>();

    if (warp < NR) {
        // ================================ rollout ====================================
        PointMass2<A, MODEL> m2;
        m2.load(prob);
        const unsigned long long step = ctl->step;
        unsigned long long key = kMinKeyInit;
        // First tile: list entry `warp` -- warp w sits on sub-partition w % 4, so a short list (a
        // small shard: fewer tiles than rollout warps) spreads evenly over the four sub-partitions
        // instead of landing wherever the race for the counter puts it (measured at 125k samples:
        // rollouts done after 117 us on CTAs that happened to balance, 165 us on those with three or
        // four tiles on one sub-partition).  Later tiles are pulled from the counter, which starts
        // behind the first round.
        for (unsigned int li = (unsigned int)warp; (long long)li < list_len;) {
            const long long tile = (long long)li * gridDim.x + blockIdx.x;
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
                c2[j] = add2(c2[j], m2.terminal_cost(x2[j], prob));
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
            if (lane == 0) {
                st_release_cta_shared_u32(&s_done[li], 1u);
                STEP_TRACE_LI(0, li);
                li = atomicAdd(s_next, 1u);
            }
            li = __shfl_sync(0xffffffffu, li, 0);
        }
        key = warp_min_u64(key);
        if (lane == 0 && key != kMinKeyInit) atomicMin(&ctl->min_key, key);
#ifdef MPPI_STEP_TRACE
        if (lane == 0) atomicMax(&g_step_trace_all[blockIdx.x][1], gtimer());
#endif
    } else if (warp == NR) {
        // ================================ producer ===================================
        if (lane == 0) tma_prefetch_desc(&tmap_eps);
        const float nil = prob->neg_inv_lambda;
        const float inf = __int_as_float(0x7f800000);
        float ref = inf;                                // no sample seen yet
        float eta_part = 0.0f;                          // this lane's share of eta, relative to ref
        const int depth = nstages / kStConsumers;       // ring depth of every consumer warp
        for (long long li = 0; li < list_len; ++li) {
            const long long tile = li * gridDim.x + blockIdx.x;
            if (lane == 0)
                while (ld_acquire_cta_shared_u32(&s_done[li]) == 0u) __nanosleep(256);
            __syncwarp();
            if (lane == 0) STEP_TRACE_LI(1, li);
            // exp_red (src/point_mass.cu:518) for this lane's four samples of the tile
            const long long k0 = tile * kStTileK + 4 * lane;
            const float4 s4 = __ldcg(reinterpret_cast<const float4 *>(S + k0));
            const bool v0 = k0 + 0 < k_local, v1 = k0 + 1 < k_local, v2 = k0 + 2 < k_local,
                       v3 = k0 + 3 < k_local;
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
            // newest rows first: the last time steps of the tile were written a few microseconds
            // ago and are still in L2, the first ones ~a tile time ago and are long evicted
            for (int chunk = nchunk - 1; chunk >= 0; --chunk) {
                // k-th box of its owner -> slot k % depth of the owner's ring
                const int cw = chunk & (kStConsumers - 1);
                const int own = (nchunk - cw + kStConsumers - 1) / kStConsumers;   // owner's boxes per tile
                const long long k = li * own + (own - 1 - (chunk >> 2));
                const int stage = cw + kStConsumers * (int)(k % depth);
                const uint32_t phase = (uint32_t)((k / depth) & 1);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                *reinterpret_cast<float4 *>(s_wt + stage * kStTileK + 4 * lane) = w4;
                if (lane == 0) s_scale[stage] = scale;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[stage], kBoxBytes);
                    tma_load_2d(s_tile + (size_t)stage * kBoxFloats, &tmap_eps,
                                (int)(tile * kStTileK), chunk * kStTileR, &full_bar[stage]);
                }
            }
            if (lane == 0) STEP_TRACE_LI(2, li);
        }
        if (lane == 0) { STEP_TRACE(200); STEP_TRACE_ALL(2); }
        eta_part = warp_sum(eta_part);
        if (lane == 0) { s_misc[0] = ref; s_misc[1] = eta_part; }
    } else {
        // ================================ consumers ==================================
        // this warp owns the boxes with chunk % 4 == cw; its k-th box sits in slot k % depth of
        // its own ring (stages cw, cw+4, ...)
        const int cw = warp - (NR + 1);
        const int h = lane >> 4, q = lane & 15;
        const int depth = nstages / kStConsumers;
        const int own = nchunk > cw ? (nchunk - cw + kStConsumers - 1) / kStConsumers : 0;
        long long k = 0;
        for (long long li = 0; li < list_len; ++li) {
            for (int chunk = cw + kStConsumers * (own - 1); chunk >= 0; chunk -= kStConsumers, ++k) {
                const int stage = cw + kStConsumers * (int)(k % depth);
                const uint32_t phase = (uint32_t)((k / depth) & 1);
                mbar_wait(&full_bar[stage], phase);
                if (chunk == cw + kStConsumers * (own - 1)) {
                    // first box of this tile for this warp: a new minimum rescales everything
                    // this warp has gathered so far (the rows of all its chunks)
                    const float scale = s_scale[stage];
                    if (scale != 1.0f) {
                        for (int cc = cw; cc < nchunk; cc += kStConsumers) {
                            float *p = s_part + (size_t)cc * kStTileR * 16 + lane;
#pragma unroll 4
                            for (int i = 0; i < kStTileR / 2; ++i) p[i * 32] *= scale;
                        }
                    }
                }
                const float *wt = s_wt + stage * kStTileK + 4 * q;
                const float4 w0 = *reinterpret_cast<const float4 *>(wt);
                const float4 w1 = *reinterpret_cast<const float4 *>(wt + 64);
                // half-warp h: rows h, h+2, ...; lane q: samples 4q..4q+3 and 64+4q..64+4q+3
                const float *tile = s_tile + (size_t)stage * kBoxFloats + (size_t)h * kStTileK + 4 * q;
                float *pp = s_part + (size_t)chunk * kStTileR * 16 + lane;     // [(2j+h)*16 + q]
#pragma unroll 2
                for (int j = 0; j < kStTileR / 2; ++j) {
                    const float4 e0 = *reinterpret_cast<const float4 *>(tile + (size_t)j * 2 * kStTileK);
                    const float4 e1 = *reinterpret_cast<const float4 *>(tile + (size_t)j * 2 * kStTileK + 64);
                    float a = pp[j * 32];
                    a = fmaf(e0.x, w0.x, a);
                    a = fmaf(e0.y, w0.y, a);
                    a = fmaf(e0.z, w0.z, a);
                    a = fmaf(e0.w, w0.w, a);
                    a = fmaf(e1.x, w1.x, a);
                    a = fmaf(e1.y, w1.y, a);
                    a = fmaf(e1.z, w1.z, a);
                    a = fmaf(e1.w, w1.w, a);
                    pp[j * 32] = a;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
        }
    }
    __syncthreads();
    if (warp > NR) return;                 // consumers are done; the other warps finish the step
    if (threadIdx.x == 0) { STEP_TRACE(201); STEP_TRACE_ALL(3); }

    // ---- this CTA's record: {row sums [R], eta, ref}, all relative to ref; the stride is a
    //      multiple of four floats so that the merge reads float4 columns
    const int rstride = record_stride(R);
    float *rec = part + (size_t)blockIdx.x * rstride;
    for (int rb = 2 * warp; rb < R; rb += 2 * (kEpiThreads / 32)) {      // two rows per warp pass
        const int r = rb + (lane >> 4);
        float v = r < R ? s_part[(size_t)r * 16 + (lane & 15)] : 0.0f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((lane & 15) == 0 && r < R) rec[r] = v;
    }
    if (threadIdx.x == 0) { rec[R] = s_misc[1]; rec[R + 1] = s_misc[0]; }

    __threadfence();                       // the record and the min key before the ticket
    fence_proxy_async_all();               // ... and before the last CTA's bulk copy of the records
    named_bar_sync(1, kEpiThreads);
    if (threadIdx.x == 0) {
        STEP_TRACE(202);
        STEP_TRACE_ALL(4);
        const unsigned ticket = atomicAdd(&ctl->done, 1u);
        *s_last = (ticket == gridDim.x - 1);
    }
    named_bar_sync(1, kEpiThreads);
    if (*s_last) {
        // ---- merge the records in CTA order (deterministic), then part 5 (merge.cuh).  The
        //      ring, the row sums and the U staging are free now and serve as its scratch.
#ifdef MPPI_STEP_TRACE
        if (threadIdx.x == 0) { g_step_trace_all[0][5] = gtimer(); g_step_trace_all[0][7] = blockIdx.x; }
#endif
        merge_records<4>(part, R, (int)gridDim.x, base, L.scale,
                         reinterpret_cast<uint64_t *>(s_misc + 4), kEpiThreads, 1, prob, ctl, fin, xa);
#ifdef MPPI_STEP_TRACE
        if (threadIdx.x == 0) { g_step_trace[0][203] = gtimer(); g_step_trace_all[0][6] = gtimer(); }
#endif
    }
}

#ifdef MPPI_STEP_TRACE
extern "C" int mppi_debug_read_step_trace(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, g_step_trace, sizeof(g_step_trace));
}
extern "C" int mppi_debug_read_step_trace_all(unsigned long long *out, int clear)
{
    int rc = (int)cudaMemcpyFromSymbol(out, g_step_trace_all, sizeof(g_step_trace_all));
    if (clear) {
        static unsigned long long zeros[160][8];
        rc |= (int)cudaMemcpyToSymbol(g_step_trace_all, zeros, sizeof(zeros));
    }
    return rc;
}
extern "C" int mppi_debug_read_step_trace_li(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, g_step_trace_li, sizeof(g_step_trace_li));
}
#endif

// =================================================================================
// launch wrappers
// =================================================================================
namespace {
#ifndef MPPI_STEP_NR
#define MPPI_STEP_NR 15
#endif
constexpr int kStepNR = MPPI_STEP_NR;   // rollout warps per CTA (+ 1 producer = whole warpgroups)
constexpr size_t kStepSmemMax = 227 * 1024;

struct StepGeom {
    long long ntiles, list_len;
    int grid, nstages;
    size_t smem;
};

// grid, tile-list length and the deepest TMA ring that fits next to the row sums
StepGeom step_geom(int T, int A, long long k_pad, int num_sms)
{
    StepGeom g{};
    g.ntiles = k_pad / kStTileK;
    g.grid = (int)(g.ntiles < num_sms ? g.ntiles : num_sms);
    g.list_len = (g.ntiles + g.grid - 1) / g.grid;
    g.nstages = 0;
    if (k_pad >= (1ll << 31)) return g;             // TMA box coordinates are int32 sample indices
    for (int ns = kStMaxStages; ns >= kStConsumers; ns -= kStConsumers) {
        const size_t b = step_smem_layout(T, A, g.list_len, ns).total;
        if (b <= kStepSmemMax && T * A + 1 <= 4 * (kStepNR + 1) * 32 &&
            step_merge_bytes(T * A, g.grid) <= step_smem_layout(T, A, g.list_len, ns).scale) {
            g.nstages = ns;
            g.smem = b;
            break;
        }
    }
    if (const char *env = getenv("MPPI_STEP_STAGES")) {
        const int v = atoi(env);
        if (v >= kStConsumers && v % kStConsumers == 0 && v <= g.nstages &&
            step_merge_bytes(T * A, g.grid) <= step_smem_layout(T, A, g.list_len, v).scale) {
            g.nstages = v;
            g.smem = step_smem_layout(T, A, g.list_len, v).total;
        }
    }
    return g;
}

template <int A, class MODEL>
cudaError_t launch_step_t(const LaunchCtx &c, const CUtensorMap &tmap, float *eps, float *U,
                          const ProblemDev *prob, float *S, CtlDev *ctl, float *part,
                          const FinalizeArgs &fin, const XchgArgs &xa)
{
    const StepGeom g = step_geom(c.horizon, c.act_dim, c.k_pad, c.num_sms);
    if (g.nstages == 0) return cudaErrorInvalidConfiguration;
    step_kernel<A, MODEL, kStepNR><<<g.grid, (kStepNR + kStConsumers + 1) * 32, g.smem, c.stream>>>(
        tmap, eps, (size_t)c.k_pad, (long long)c.k_local, c.horizon, U, prob, S, ctl,
        (unsigned long long)c.k_offset, c.sampler, part, fin, g.nstages, xa);
    return cudaGetLastError();
}

template <int A, class MODEL>
cudaError_t configure_step_m(int smem)
{
    return cudaFuncSetAttribute(step_kernel<A, MODEL, kStepNR>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <int A>
cudaError_t configure_step_a(int smem)
{
    cudaError_t e;
    if ((e = configure_step_m<A, Model<false, DoubleIntegrator>>(smem)) != cudaSuccess) return e;
    if ((e = configure_step_m<A, Model<true, DoubleIntegrator>>(smem)) != cudaSuccess) return e;
    if ((e = configure_step_m<A, Model<false, LinearAxis>>(smem)) != cudaSuccess) return e;
    return configure_step_m<A, Model<true, LinearAxis>>(smem);
}
}  // namespace

bool step_kernel_supported(int T, int A, long long k_pad, int num_sms)
{
    return step_geom(T, A, k_pad, num_sms).nstages >= kStConsumers;
}

size_t step_part_floats(const LaunchCtx &c)
{
    return (size_t)c.num_sms * (((size_t)c.rows + 2 + 3) & ~(size_t)3);
}

cudaError_t configure_step(const LaunchCtx &c)
{
    const StepGeom g = step_geom(c.horizon, c.act_dim, c.k_pad, c.num_sms);
    if (g.nstages == 0) return cudaSuccess;
    switch (c.act_dim) {
        case 1: return configure_step_a<1>((int)kStepSmemMax);
        case 2: return configure_step_a<2>((int)kStepSmemMax);
        case 3: return configure_step_a<3>((int)kStepSmemMax);
        case 4: return configure_step_a<4>((int)kStepSmemMax);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_step(const LaunchCtx &c, const CUtensorMap &tmap, float *eps, float *U,
                        const ProblemDev *prob, float *S, CtlDev *ctl, float *part,
                        float *U_prev, float *next_act, unsigned flags, const XchgArgs &xa)
{
    FinalizeArgs fin{U, U_prev, next_act, c.horizon, c.act_dim, flags};
#define MPPI_STEP_CASE(A_)                                                                       \
    case A_:                                                                                     \
        if (c.general_gains)                                                                     \
            return c.strict ? launch_step_t<A_, Model<true, LinearAxis>>(c, tmap, eps, U, prob, S, ctl, part, fin, xa) \
                            : launch_step_t<A_, Model<false, LinearAxis>>(c, tmap, eps, U, prob, S, ctl, part, fin, xa); \
        return c.strict ? launch_step_t<A_, Model<true, DoubleIntegrator>>(c, tmap, eps, U, prob, S, ctl, part, fin, xa) \
                        : launch_step_t<A_, Model<false, DoubleIntegrator>>(c, tmap, eps, U, prob, S, ctl, part, fin, xa)
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
