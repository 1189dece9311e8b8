// tile.cu -- the whole control step as ONE persistent kernel that keeps eps ON CHIP.
//
// Why.  Every other chain of this library moves the K x T x A perturbations through HBM at least
// twice (written by the sampler, read back by the weighted average): 8*A bytes per rollout-step,
// which is what bounds them (step.cu: 4.8 GB per step at K=1e6, 0.83 of the HBM peak).  But eps is
// only ever needed twice -- by the rollout of its own sample and, once that sample's cost (hence
// its weight) is known, by the weighted average -- and a rollout's eps is T*A*4 = 2.4 KB.  A tile
// of 64 samples is 154 KB: it fits in the 227 KB of shared memory of ONE SM.  So here the noise
// never leaves the SM that drew it:
//
//   tile = 64 consecutive samples; CTA c (one per SM) owns tiles c, c+grid, c+2*grid, ...
//
//   warps [0, NG)   generators, then averagers.  The T*A x 16 Philox calls of a tile (one call
//                   = the float4 of four consecutive samples in one eps row, philox.cuh) are
//                   cut into units of 4 rows x 16 quads = two calls per lane, pulled from a
//                   shared-memory counter; each result is ONE 16-byte STS into the tile
//                   [row][64+4].  Units complete passes of `upp` units; every pass has an
//                   mbarrier the integrator waits on, so it follows the generators a pass
//                   behind.  Once the integrator has published the tile's weights, thread r
//                   folds row r:  acc_r = acc_r*scale + sum_k w_k eps[r][k]  -- 16 LDS.128 of
//                   its row (pitch 68 floats: conflict free), the weights by broadcast, packed
//                   FP32x2 FMAs, the accumulator in a register for the whole kernel.
//   warp  NG        the integrator: lane l owns samples 2l, 2l+1 of the tile and advances them
//                   as one packed FP32x2 pair through the T steps (model.cuh: the operations
//                   of PointMassModelGpu::step / Cost::step_cost, src/point_mass_gpu.cu:82-109,
//                   src/cost.cu:42-64, bit for bit), reading eps[t] from shared memory one step
//                   ahead.  Then S -> global (the get_inf tap), the packed (S,k) min key, and
//                   the weights relative to the CTA's running minimum `ref` (online softmax: a
//                   tile that lowers ref carries the factor exp(-(ref_old-ref_new)/lambda) by
//                   which everything gathered so far is rescaled).
//
// HBM traffic of a step: 4 bytes per sample (S) plus the per-CTA records.  What bounds the
// kernel is instruction issue (the Philox rounds and Box-Muller on the MUFU pipe), which the
// generator warps keep busy without ever waiting for memory.
//
// The end is the step kernel's: each CTA writes {row sums, eta, ref} to a record, the last
// CTA (ticket) merges the records in CTA order and applies the U update (merge.cuh).  Tile ->
// CTA and row -> thread are static and every sum has a fixed order, so results are bitwise
// reproducible whichever warp generated which unit.
//
// The get_inf tap still returns eps: Philox is counter based, so the controller re-draws the
// step's noise into the HBM buffer on demand (same function, same bits).
#include "finalize.cuh"
#include "kernels.cuh"
#include "merge.cuh"
#include "model.cuh"
#include "philox.cuh"

#include <stdlib.h>

namespace mppi {

constexpr int kTlW        = 64;    // samples per tile: one integrator warp, a packed pair per lane
constexpr int kTlPitch    = kTlW + 4;   // floats per eps row in shared memory
constexpr int kTlQuads    = kTlW / 4;   // Philox calls per eps row
constexpr int kTlUnitRows = 4;     // rows per generator unit: 4 x 16 calls = two per lane
constexpr int kTlUnitsPerPass = 8; // units per pass: the integrator follows the generators 32 rows behind
constexpr int kTlMaxPass  = 32;    // mbarriers for "rows up to here are complete"
constexpr int kTlHeader   = 1024;  // bytes of barriers / flags / weights in front of the tile
constexpr int kTlRowsPerThread = 2;     // averaging: thread owns rows {gt, gt + NG*32}

// U staging: per time step {u_0..u_{A-1}, u_0*inv_s_0..}, padded to a multiple of four floats
__host__ __device__ constexpr int tile_u_stride(int A) { return (2 * A + 3) / 4 * 4; }
__host__ __device__ inline size_t tile_work_bytes(int T, int A)
{
    return (size_t)T * A * kTlPitch * sizeof(float) + (size_t)T * tile_u_stride(A) * sizeof(float);
}

struct TileGeom {
    long long ntiles, list_len;
    int grid, npass;
    size_t smem;
    bool ok;
};

template <int NG>
static TileGeom tile_geom(int T, int A, long long k_pad, int num_sms)
{
    TileGeom g{};
    const int R = T * A;
    g.ntiles = k_pad / kTlW;
    g.grid = (int)(g.ntiles < num_sms ? g.ntiles : num_sms);
    g.list_len = (g.ntiles + g.grid - 1) / g.grid;
    const int nunits = (R + kTlUnitRows - 1) / kTlUnitRows;
    g.npass = (nunits + kTlUnitsPerPass - 1) / kTlUnitsPerPass;
    const size_t work = tile_work_bytes(T, A);
    const size_t merge = merge_smem_bytes(R, g.grid);
    g.smem = kTlHeader + (work > merge ? work : merge);
    g.ok = g.smem <= 227 * 1024 && g.npass <= kTlMaxPass && R <= kTlRowsPerThread * NG * 32 &&
           R + 1 <= 4 * (NG + 1) * 32 &&
           k_pad < (1ll << 31) && g.list_len * (long long)nunits < (1ll << 31);
    return g;
}

__device__ __forceinline__ void named_bar(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// steps the integrator advances per branch-free block: the stage costs of the steps of a block
// are independent chains (only the state recurrence and the running sum are serial), so the
// scheduler can interleave them -- a warp that runs alone on its sub-partition has nothing but
// its own instruction-level parallelism to hide the 4-cycle dependent-issue latency
constexpr int kTlBlockSteps = 4;

template <int A, class MODEL, int NG>
__global__ void __launch_bounds__((NG + 1) * 32, 1)
tile_kernel(long long k_local, int T, const float *__restrict__ U,
            const ProblemDev *__restrict__ prob, float *__restrict__ S, CtlDev *__restrict__ ctl,
            unsigned long long k_offset, const __grid_constant__ SamplerParams sp,
            float *__restrict__ part, FinalizeArgs fin, long long ntiles,
            const __grid_constant__ XchgArgs xa)
{
    constexpr int kThreads = (NG + 1) * 32;
    constexpr int kGenThreads = NG * 32;
    constexpr int kUS = tile_u_stride(A);
    const int R = T * A;
    const int nunits = (R + kTlUnitRows - 1) / kTlUnitRows;
    const int npass = (nunits + kTlUnitsPerPass - 1) / kTlUnitsPerPass;

    // declared aligned and only ever addressed through typed pointers derived from it: every
    // access below is a shared-space LDS/STS (no generic-address round trip)
    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t *eps_full = reinterpret_cast<uint64_t *>(base);              // [kTlMaxPass]
    uint64_t *w_full   = eps_full + kTlMaxPass;                           // the tile's weights are out
    uint64_t *merge_bar = w_full + 1;
    unsigned int *s_next = reinterpret_cast<unsigned int *>(base + 288);  // next unit (never reset)
    int *s_last = reinterpret_cast<int *>(base + 292);
    float *s_scale = reinterpret_cast<float *>(base + 296);               // rescale carried by the tile
    float *s_w = reinterpret_cast<float *>(base + 512);                   // [64] weights of the tile
    uint8_t *region = base + kTlHeader;
    float *s_eps = reinterpret_cast<float *>(region);                     // [R][kTlPitch]
    float *s_u = s_eps + (size_t)R * kTlPitch;                            // [T][kUS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long list_len = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    for (int i = threadIdx.x; i < R; i += kThreads) {
        const int t = i / A, a = i - t * A;
        const float u = U[i];
        s_u[t * kUS + a] = u;
        s_u[t * kUS + A + a] = __fmul_rn(u, prob->inv_s[a]);             // src/cost.cu:46
    }
    if (threadIdx.x == 0) {
        *s_next = 0u;
        for (int p = 0; p < npass; ++p) {
            const int n = min(kTlUnitsPerPass, nunits - p * kTlUnitsPerPass);
            mbar_init(&eps_full[p], (uint32_t)n);                        // one arrival per unit
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned long long step = ctl->step;

    if (warp < NG) {
        // =========================== generators / averagers ===========================
        const int gt = threadIdx.x;                                       // < kGenThreads
        float accr[kTlRowsPerThread];
#pragma unroll
        for (int i = 0; i < kTlRowsPerThread; ++i) accr[i] = 0.0f;
        const int quad = lane & (kTlQuads - 1);
        const int rsub = lane >> 4;                                       // 0/1: which row of the pair
        float *wr = s_eps + 4 * quad;                                     // this lane's column of the tile

        unsigned int g_next = 0;
        if (lane == 0) g_next = atomicAdd(s_next, 1u);
        g_next = __shfl_sync(0xffffffffu, g_next, 0);

        for (long long it = 0; it < list_len; ++it) {
            const long long tile = it * gridDim.x + blockIdx.x;
            const uint32_t q = (uint32_t)((k_offset >> 2) + (unsigned long long)tile * kTlQuads) + quad;
            const unsigned int first = (unsigned int)it * (unsigned int)nunits;
            const unsigned int lim = first + (unsigned int)nunits;
            while (g_next < lim) {
                const int u = (int)(g_next - first);
                if (lane == 0) g_next = atomicAdd(s_next, 1u);           // the unit after this one
                const int ra = kTlUnitRows * u + rsub, rb = ra + 2;
                const int rac = min(ra, R - 1), rbc = min(rb, R - 1);
                const float4 na = sample4(q, (uint32_t)rac, step, sp, sp.c[rac % A]);
                const float4 nb = sample4(q, (uint32_t)rbc, step, sp, sp.c[rbc % A]);
                if (ra < R) *reinterpret_cast<float4 *>(wr + (size_t)ra * kTlPitch) = na;
                if (rb < R) *reinterpret_cast<float4 *>(wr + (size_t)rb * kTlPitch) = nb;
                g_next = __shfl_sync(0xffffffffu, g_next, 0);
                __syncwarp();                                             // every lane's STS ...
                if (lane == 0) mbar_arrive(&eps_full[u / kTlUnitsPerPass]);   // ... before the release
            }
            // ---- the tile's weights are known: fold my rows
            mbar_wait(w_full, (uint32_t)(it & 1));
            const float scale = *s_scale;
#pragma unroll
            for (int i = 0; i < kTlRowsPerThread; ++i) {
                const int r = gt + i * kGenThreads;
                if (r < R) {
                    const float *er = s_eps + (size_t)r * kTlPitch;
                    f2 p0 = mk2(0.0f, 0.0f), p1 = p0;
#pragma unroll
                    for (int j = 0; j < kTlQuads; ++j) {
                        const float4 e = *reinterpret_cast<const float4 *>(er + 4 * j);
                        const float4 w = *reinterpret_cast<const float4 *>(s_w + 4 * j);
                        p0 = fma2(mk2(e.x, e.y), mk2(w.x, w.y), p0);
                        p1 = fma2(mk2(e.z, e.w), mk2(w.z, w.w), p1);
                    }
                    float lo, hi;
                    un2(add2(p0, p1), lo, hi);
                    accr[i] = fmaf(accr[i], scale, lo + hi);
                }
            }
            // every averager is done with the tile before any generator overwrites it
            named_bar(1, kGenThreads);
        }
        // ---- this CTA's record, rows part
        float *rec = part + (size_t)blockIdx.x * record_stride(R);
#pragma unroll
        for (int i = 0; i < kTlRowsPerThread; ++i) {
            const int r = gt + i * kGenThreads;
            if (r < R) rec[r] = accr[i];
        }
    } else {
        // ================================ integrator ==================================
        PointMass2<A, MODEL> m2;
        m2.load(prob);
        const float nil = prob->neg_inv_lambda;
        const float inf = __int_as_float(0x7f800000);
        float ref = inf;                                // no sample seen yet
        float eta_part = 0.0f;                          // this lane's share of eta, relative to ref
        unsigned long long key = kMinKeyInit;
        const float *ecol = s_eps + 2 * lane;
        // one step from shared memory: eps[t] of my pair, {u, u*inv_s}[t] by broadcast
        auto one_step = [&](int t, f2 (&x2)[2 * A], f2 &c2) {
            f2 e[A], u[A], ui[A];
            const float *us = s_u + (size_t)t * kUS;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                e[a].r = *reinterpret_cast<const unsigned long long *>(ecol + (size_t)(t * A + a) * kTlPitch);
                u[a] = mk2(us[a], us[a]);
                ui[a] = mk2(us[A + a], us[A + a]);
            }
            m2.step(x2, c2, u, ui, e);
        };
        for (long long it = 0; it < list_len; ++it) {
            const long long tile = it * gridDim.x + blockIdx.x;
            const uint32_t par = (uint32_t)(it & 1);
            f2 x2[2 * A], c2 = mk2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < 2 * A; ++i) x2[i] = mk2(prob->x0[i], prob->x0[i]);
            int ready_rows = 0, ready_pass = 0;         // rows [0, ready_rows) of this tile are there
            auto need = [&](int row_end) {              // rows [0, row_end) wanted
                while (ready_rows < row_end) {
                    mbar_wait(&eps_full[ready_pass], par);
                    ++ready_pass;
                    ready_rows = ready_pass * kTlUnitsPerPass * kTlUnitRows;
                }
            };
            int t = 0;
#pragma unroll 1
            for (; t + kTlBlockSteps <= T; t += kTlBlockSteps) {
                need((t + kTlBlockSteps) * A);
#pragma unroll
                for (int i = 0; i < kTlBlockSteps; ++i) one_step(t + i, x2, c2);
            }
#pragma unroll 1
            for (; t < T; ++t) {
                need((t + 1) * A);
                one_step(t, x2, c2);
            }
            // terminal cost on x[T], charged on top of the last stage cost
            // (src/point_mass_gpu.cu:116)
            c2 = add2(c2, m2.terminal_cost(x2, prob));
            float s0, s1;
            un2(c2, s0, s1);
            const long long k0 = tile * kTlW + 2 * lane;
            *reinterpret_cast<float2 *>(S + k0) = make_float2(s0, s1);
            const bool v0 = k0 < k_local, v1 = k0 + 1 < k_local;
            if (v0) {
                const unsigned long long kk = ((unsigned long long)float_to_ordered(s0) << 32) |
                                              (unsigned long long)(uint32_t)(k_offset + (unsigned long long)k0);
                key = kk < key ? kk : key;
            }
            if (v1) {
                const unsigned long long kk = ((unsigned long long)float_to_ordered(s1) << 32) |
                                              (unsigned long long)(uint32_t)(k_offset + (unsigned long long)k0 + 1);
                key = kk < key ? kk : key;
            }
            // exp_red (src/point_mass.cu:518) relative to the running minimum
            float tmin = warp_min_f(fminf(v0 ? s0 : inf, v1 ? s1 : inf));
            float scale = 1.0f;
            if (tmin < ref) {
                scale = expf(__fmul_rn(nil, __fsub_rn(ref, tmin)));       // ref = +inf -> 0
                eta_part *= scale;
                ref = tmin;
            }
            const float w0 = v0 ? expf(__fmul_rn(nil, __fsub_rn(s0, ref))) : 0.0f;
            const float w1 = v1 ? expf(__fmul_rn(nil, __fsub_rn(s1, ref))) : 0.0f;
            eta_part += w0 + w1;
            *reinterpret_cast<float2 *>(s_w + 2 * lane) = make_float2(w0, w1);
            if (lane == 0) *s_scale = scale;
            __syncwarp();
            if (lane == 0) mbar_arrive(w_full);                           // release: s_w, s_scale
        }
        key = warp_min_u64(key);
        if (lane == 0 && key != kMinKeyInit) atomicMin(&ctl->min_key, key);
        eta_part = warp_sum(eta_part);
        if (lane == 0) {
            float *rec = part + (size_t)blockIdx.x * record_stride(R);
            rec[R] = eta_part;
            rec[R + 1] = ref;
        }
    }

    __threadfence();                       // the record and the min key before the ticket
    fence_proxy_async_all();               // ... and before the last CTA's bulk copy of the records
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&ctl->done, 1u);
        *s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (*s_last) {
        // the tile and the U staging are free now: scratch of the merge (the launcher sized the
        // region as the larger of the two uses)
        const size_t work = tile_work_bytes(T, A);
        const size_t mb = merge_smem_bytes(R, (int)gridDim.x);
        merge_records<4>(part, R, (int)gridDim.x, region, work > mb ? work : mb, merge_bar, kThreads,
                         0, prob, ctl, fin, xa);
    }
}

// =================================================================================
// launch wrappers
// =================================================================================
namespace {
constexpr int kTileNG = 19;   // generator warps per CTA (+ 1 integrator = 640 threads)

template <int A, class MODEL>
cudaError_t launch_tile_t(const LaunchCtx &c, float *U, const ProblemDev *prob, float *S, CtlDev *ctl,
                          float *part, const FinalizeArgs &fin, const XchgArgs &xa)
{
    const TileGeom g = tile_geom<kTileNG>(c.horizon, c.act_dim, c.k_pad, c.num_sms);
    if (!g.ok) return cudaErrorInvalidConfiguration;
    tile_kernel<A, MODEL, kTileNG><<<g.grid, (kTileNG + 1) * 32, g.smem, c.stream>>>(
        (long long)c.k_local, c.horizon, U, prob, S, ctl, (unsigned long long)c.k_offset, c.sampler,
        part, fin, g.ntiles, xa);
    return cudaGetLastError();
}

template <int A, class MODEL>
cudaError_t configure_tile_m()
{
    return cudaFuncSetAttribute(tile_kernel<A, MODEL, kTileNG>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

template <int A>
cudaError_t configure_tile_a()
{
    cudaError_t e;
    if ((e = configure_tile_m<A, Model<false, DoubleIntegrator>>()) != cudaSuccess) return e;
    if ((e = configure_tile_m<A, Model<true, DoubleIntegrator>>()) != cudaSuccess) return e;
    if ((e = configure_tile_m<A, Model<false, LinearAxis>>()) != cudaSuccess) return e;
    return configure_tile_m<A, Model<true, LinearAxis>>();
}
}  // namespace

bool tile_kernel_supported(int T, int A, long long k_pad, int num_sms)
{
    return tile_geom<kTileNG>(T, A, k_pad, num_sms).ok;
}

cudaError_t configure_tile(const LaunchCtx &c)
{
    switch (c.act_dim) {
        case 1: return configure_tile_a<1>();
        case 2: return configure_tile_a<2>();
        case 3: return configure_tile_a<3>();
        case 4: return configure_tile_a<4>();
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_tile(const LaunchCtx &c, float *U, const ProblemDev *prob, float *S, CtlDev *ctl,
                        float *part, float *U_prev, float *next_act, unsigned flags,
                        const XchgArgs &xa)
{
    FinalizeArgs fin{U, U_prev, next_act, c.horizon, c.act_dim, flags};
#define MPPI_TILE_CASE(A_)                                                                        \
    case A_:                                                                                      \
        if (c.general_gains)                                                                      \
            return c.strict ? launch_tile_t<A_, Model<true, LinearAxis>>(c, U, prob, S, ctl, part, fin, xa)  \
                            : launch_tile_t<A_, Model<false, LinearAxis>>(c, U, prob, S, ctl, part, fin, xa); \
        return c.strict ? launch_tile_t<A_, Model<true, DoubleIntegrator>>(c, U, prob, S, ctl, part, fin, xa) \
                        : launch_tile_t<A_, Model<false, DoubleIntegrator>>(c, U, prob, S, ctl, part, fin, xa)
    switch (c.act_dim) {
        MPPI_TILE_CASE(1);
        MPPI_TILE_CASE(2);
        MPPI_TILE_CASE(3);
        MPPI_TILE_CASE(4);
        default: return cudaErrorInvalidValue;
    }
#undef MPPI_TILE_CASE
}

}  // namespace mppi
