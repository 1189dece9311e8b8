// controller.cu -- host side of the B200 MPPI core and its C ABI (include/mppi_b200.h).
//
// Counterpart of the reference's host object PointMassModel (include/point_mass.hpp:23-116,
// src/point_mass.cu:19-491): owns every device buffer of one K-shard, builds the control
// step as a CUDA graph once and replays it per step.  Where the reference issues ~3T+8
// launches and as many cudaDeviceSynchronize() per get_act (src/point_mass.cu:129-203,
// :384-480), one step here is one cudaGraphLaunch of 1-3 kernels (step kernel / fused chain /
// unfused chain, plus one exchange kernel on K-shards); the finalizing kernel publishes the next
// action and the step counter in mapped host memory, so the graph holds no copy node.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mppi_b200.h"
#include "comm.hpp"
#include "kernels.cuh"
#include "finalize.cuh"

using namespace mppi;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(MPPI_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                        cudaGetErrorString(e__));                                         \
    } while (0)

constexpr int kStageSlots = 16;   // pinned staging ring for set_state

}  // namespace

struct mppi_handle {
    mppi_params p{};
    cudaStream_t stream = nullptr;
    LaunchCtx ctx{};
    int R = 0, S = 0;

    float *d_eps = nullptr, *d_S = nullptr, *d_wt = nullptr;
    long long *d_acc = nullptr;   // [R+1] fixed-point accumulators (rows, eta)
    float *d_U = nullptr, *d_Uprev = nullptr, *d_next = nullptr;
    ProblemDev *d_prob = nullptr;
    CtlDev *d_ctl = nullptr;

    ProblemDev h_prob{};
    float *h_stage = nullptr;     // pinned [kStageSlots][2*kMaxAct]
    int stage_slot = 0;
    bool stage_used[kStageSlots] = {};
    unsigned long long stage_tag[kStageSlots] = {};   // steps enqueued when the slot was staged
    float *h_next = nullptr;      // pinned + mapped [kNextFloats]: next action, error flag, step seq
    unsigned long long steps_enqueued = 0;   // == the seq the last enqueued step will publish

    CUtensorMap tmap{};        // average kernel: box {256, kAvgTileR}
    CUtensorMap tmap_ro{};     // TMA rollout: box {256, TT*A}
    CUtensorMap tmap_st{};     // step kernel: box {128, 40}
    float *d_part = nullptr;   // step kernel: one {ref, eta, row sums} record per CTA
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};   // [0] sampling, [1] injected noise

    // MPPI_FLAG_PIPELINED_SAMPLING: a second eps buffer with its tensor maps; d_eps / tmap /
    // tmap_ro are always the buffer the NEXT chain reads, the pairs are swapped after every
    // pipelined step.  graph_pipe[i]: the sampler-less chain on base buffer i.  The sampler of
    // the step after runs on stream2, behind the chain, ordered against it by two events.
    float *d_eps_alt = nullptr;
    float *eps_base[2] = {nullptr, nullptr};
    CUtensorMap tmap_alt{}, tmap_ro_alt{};
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_sampled = nullptr;      // stream2: the noise drawn ahead is complete
    cudaEvent_t ev_chain_done = nullptr;   // stream: the last chain has finished reading its buffer
    cudaGraphExec_t graph_pipe[2] = {nullptr, nullptr};
    bool presampled = false;   // d_eps holds (or is receiving) the noise of the next step
    float *eps_last = nullptr; // the noise the last step consumed (get_inf tap); == d_eps outside
                               // the pipeline
    NcclComm comm;
    unsigned long long *d_mailbox = nullptr;              // MPPI_COMM_P2P: this rank's mailbox
    unsigned long long *peer_mb[kMaxWorld] = {};          // every rank's mailbox mapped here
    bool p2p_connected = false;
    bool peers_are_ipc = false;                           // peer_mb[] came from cudaIpcOpenMemHandle
    std::vector<mppi_handle *> children;                  // single-process multi-device group

    bool step_ok = false, tile_ok = false;   // the one-kernel variants this shape supports
    // the last step kept eps on chip (tile kernel): the get_inf tap re-draws it for this Philox
    // step index before reading d_eps
    bool eps_on_chip = false;
    unsigned long long eps_step = 0;

    bool problem_set = false;
    bool terminal_set = false;   // mppi_set_terminal_weights gave the final state its own weights
    bool injected = false;
    bool profiling = false;
    bool pending = false;
    bool prof_pending = false, prof_sampled = false, prof_one_kernel = false;      // profiling times to collect at wait

    cudaEvent_t ev[MPPI_K_COUNT + 1] = {};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    double ms_sum[MPPI_K_COUNT] = {};
    int64_t launches[MPPI_K_COUNT] = {};
    int64_t total_launches = 0;
};

namespace {

bool multi(const mppi_handle *h) { return h->p.world_size > 1; }
bool p2p(const mppi_handle *h) { return multi(h) && h->p.comm == MPPI_COMM_P2P; }
bool fused(const mppi_handle *h) { return (h->p.flags & MPPI_FLAG_FUSED_SAMPLING) != 0; }
// the one-kernel step: sampled noise, single shard, row sums fit in shared memory
// the on-chip tile kernel (tile.cu): same conditions, takes precedence over the step kernel
bool tile_step(const mppi_handle *h, bool sample)
{
    return sample && (h->p.flags & MPPI_FLAG_TILE_KERNEL) && (!multi(h) || p2p(h)) &&
           !(h->p.flags & MPPI_FLAG_SPLIT_KERNELS) && h->d_part != nullptr && h->tile_ok;
}
bool one_kernel(const mppi_handle *h, bool sample)
{
    if (tile_step(h, sample)) return true;
    return sample && (h->p.flags & MPPI_FLAG_STEP_KERNEL) && (!multi(h) || p2p(h)) &&
           !(h->p.flags & MPPI_FLAG_SPLIT_KERNELS) && h->d_part != nullptr && h->step_ok;
}

// pipelined sampling applies to the sampled, unfused, graph-replayed chain only
bool pipelined(const mppi_handle *h, bool sample)
{
    return sample && h->d_eps_alt != nullptr && !h->profiling && !(h->p.flags & MPPI_FLAG_NO_GRAPH);
}
// the noise the last finished step consumed (get_inf tap)
float *last_eps(const mppi_handle *h) { return h->eps_last ? h->eps_last : h->d_eps; }

int encode_tmap(mppi_handle *h, CUtensorMap *out, int box_cols, int box_rows, float *base = nullptr)
{
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(MPPI_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)h->ctx.k_pad, (cuuint64_t)h->R};
    const cuuint64_t gstride[1] = {(cuuint64_t)h->ctx.k_pad * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiled>(fn)(
        out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base ? base : h->d_eps, gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MPPI_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
    return MPPI_OK;
}

int encode_eps_tmaps(mppi_handle *h)
{
    int rc;
    if ((rc = encode_tmap(h, &h->tmap, kAvgTileK, kAvgTileR)) != MPPI_OK) return rc;
    if ((rc = encode_tmap(h, &h->tmap_st, kStepTileK, kStepTileR)) != MPPI_OK) return rc;
    return encode_tmap(h, &h->tmap_ro, h->ctx.rollout_tma_width, rollout_tma_rows(h->p.act_dim));
}

// The eps buffer of a tile-kernel handle, on first need (see mppi_create).
int ensure_eps(mppi_handle *h)
{
    if (h->d_eps) return MPPI_OK;
    const size_t eps_bytes = sizeof(float) * (size_t)h->R * (size_t)h->ctx.k_pad;
    cudaError_t e = cudaMalloc(&h->d_eps, eps_bytes);
    if (e != cudaSuccess) {
        h->d_eps = nullptr;
        cudaGetLastError();
        return fail(MPPI_ERR_CUDA, "the eps buffer (%zu bytes: K_local x T x A floats) cannot be allocated: %s; "
                    "a tile-kernel handle steps without it, but injected noise and the eps / trajectory "
                    "taps need it", eps_bytes, cudaGetErrorString(e));
    }
    CK(cudaMemsetAsync(h->d_eps, 0, eps_bytes, h->stream));
    return encode_eps_tmaps(h);
}

// Enqueue the kernel chain of one control step on h->stream.  evs (optional) receives one
// event before the first kernel and one after each stage (profiling mode).
int enqueue_chain(mppi_handle *h, bool sample, cudaEvent_t *evs)
{
    const LaunchCtx &c = h->ctx;
    std::string err;
    int ei = 0;
    auto mark = [&](void) -> cudaError_t {
        return evs ? cudaEventRecord(evs[ei++], c.stream) : cudaSuccess;
    };
    CK(mark());
    if (one_kernel(h, sample)) {
        for (int i = 0; i < MPPI_K_AVERAGE; ++i) CK(mark());     // the time is booked on "average"
        // K-shards: the last CTA of the kernel also runs the NVLink exchange (xchg.cuh) -- one
        // kernel per step for any number of GPUs
        const XchgArgs xa = make_xchg_args(h->peer_mb, h->p.rank, p2p(h) ? h->p.world_size : 1, h->R);
        if (tile_step(h, sample))
            CK(launch_tile(c, h->d_U, h->d_prob, h->d_S, h->d_ctl, h->d_part, h->d_Uprev, h->d_next,
                           h->p.flags, xa));
        else
            CK(launch_step(c, h->tmap_st, h->d_eps, h->d_U, h->d_prob, h->d_S, h->d_ctl, h->d_part,
                           h->d_Uprev, h->d_next, h->p.flags, xa));
        CK(mark());
        CK(mark());
        CK(mark());
        return MPPI_OK;
    }
    if (sample && !fused(h)) CK(launch_sample(c, h->d_eps, h->d_ctl, false, 0));
    CK(mark());
    if (c.rollout_tma && !(sample && fused(h)))
        CK(launch_rollout_tma(c, h->tmap_ro, h->d_U, h->d_prob, h->d_S, h->d_ctl));
    else
        CK(launch_rollout(c, h->d_eps, h->d_U, h->d_prob, h->d_S, h->d_ctl, sample && fused(h)));
    CK(mark());
    // peer-mailbox shards average relative to their own minimum and merge afterwards (one
    // exchange); NCCL shards all-reduce beta first (two exchanges)
    const bool split = (h->p.flags & MPPI_FLAG_SPLIT_KERNELS) != 0;
    const bool one_xchg = p2p(h) && !split;
    const XchgArgs xa = make_xchg_args(h->peer_mb, h->p.rank, p2p(h) ? h->p.world_size : 1, h->R);
    if (p2p(h) && !one_xchg) {
        CK(launch_xchg_min(c, h->d_ctl, xa));
    } else if (multi(h) && !p2p(h)) {
        if (!h->comm.allreduce_min_u64(&h->d_ctl->min_key, 1, c.stream, err))
            return fail(MPPI_ERR_COMM, "%s", err.c_str());
    }
    CK(mark());
    if (split) CK(launch_weights(c, h->d_S, h->d_prob, h->d_ctl, h->d_wt, h->d_acc));
    CK(mark());
    // single shard, or peer-mailbox shards merging with one exchange: the last CTA of the
    // averaging kernel finishes the step (exchange included) -- no kernel behind it
    const bool merge_fin = !split && (!multi(h) || one_xchg);
    // the rollout kernel is directly in front of the average unless a collective or the weights
    // kernel sits between them: programmatic dependent launch of the average
    bool pdl = !split && !(multi(h) && !one_xchg);
    if (const char *env = getenv("MPPI_PDL")) pdl = pdl && atoi(env) != 0;
    CK(launch_average(c, h->tmap, split ? h->d_wt : h->d_S, h->d_acc, !split, merge_fin, h->d_prob,
                      h->d_ctl, h->d_U, h->d_Uprev, h->d_next, h->p.flags, xa, pdl));
    CK(mark());
    if (one_xchg) {
        // nothing: exchanged and finalized inside the averaging kernel
    } else if (p2p(h)) {
        // exchange + integer sum + U update in one kernel over peer memory
        CK(launch_xchg_sum_finalize(c, h->d_acc, h->d_U, h->d_Uprev, h->d_prob, h->d_ctl, h->d_next,
                                    h->p.flags, xa));
    } else if (multi(h)) {
        if (!h->comm.allreduce_sum_i64(h->d_acc, (size_t)h->R + 1, c.stream, err))
            return fail(MPPI_ERR_COMM, "%s", err.c_str());
    }
    CK(mark());
    if (!merge_fin && !p2p(h))
        CK(launch_finalize(c, h->d_acc, h->d_U, h->d_Uprev, h->d_prob, h->d_ctl, h->d_next,
                           h->p.flags));
    // next action (A floats), the exchange error flag at [kMaxAct] and the step counter are
    // stored by the finalizing kernel straight into mapped host memory (finalize.cuh)
    CK(mark());
    return MPPI_OK;
}

int kernels_per_step(const mppi_handle *h, bool sample)
{
    if (one_kernel(h, sample)) return 1;            // K-shards exchange inside the same kernel
    int n = 2;                                  // rollout, average(+weights,+finalize)
    if (sample && !fused(h)) n += 1;            // sampling
    if (h->p.flags & MPPI_FLAG_SPLIT_KERNELS) n += 1;      // separate weights kernel
    if (p2p(h)) n += (h->p.flags & MPPI_FLAG_SPLIT_KERNELS) ? 2 : 0;   // xchg_min, xchg_sum+finalize;
                                                                       // else inside the average
    else if (multi(h) || (h->p.flags & MPPI_FLAG_SPLIT_KERNELS)) n += 1;   // finalize kernel
    return n;
}

int build_graph(mppi_handle *h, int which)
{
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_chain(h, which == 0, nullptr);
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc != MPPI_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec[which], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(e));
    return MPPI_OK;
}

// The sampler-less chain (== the chain of the injected-noise mode) on the current buffer.
int build_pipe_graph(mppi_handle *h, int which)
{
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_chain(h, false, nullptr);
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc != MPPI_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_pipe[which], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(e));
    return MPPI_OK;
}

// Leave the pipeline before anything but a pipelined step touches eps: noise drawn ahead may
// still be landing in d_eps (wait for it, then it is void -- whoever draws for that step next
// gets the same values), and the plain graphs were captured on base buffer 0, so d_eps goes
// back to it.  eps_last keeps pointing at the noise the last step consumed.
int leave_pipeline(mppi_handle *h)
{
    if (!h->stream2) return MPPI_OK;
    if (h->presampled) CK(cudaStreamSynchronize(h->stream2));
    h->presampled = false;
    if (h->d_eps != h->eps_base[0]) {
        std::swap(h->d_eps, h->d_eps_alt);
        std::swap(h->tmap, h->tmap_alt);
        std::swap(h->tmap_ro, h->tmap_ro_alt);
    }
    return MPPI_OK;
}

// MPPI_FLAG_AUTO_CHAIN: the chain is chosen from an estimate of what each of the three costs on
// this shard (microseconds per control step), not from K alone.  The rollouts are bound by the
// FMA pipe of an SM sub-partition (Philox IMAD.WIDE and the packed FP32 dynamics share it,
// tools/ubench/pipes.cu), so what a sample+rollout pass costs is the number of 128-sample warps
// on the FULLEST sub-partition times the horizon -- a step function of the shard size -- while
// the eps passes are HBM streams with a fixed start-up cost.  Constants fitted on one B200
// (tools/chain_sweep.py: 7 (A, T) shapes x 16 shard sizes from 3e3 to 1e6, among them the 3-, 5-, 6-
// and 7-GPU shards of K = 1e6, profiles/r02_chain_sweep_v2.jsonl: the chosen family is within 3.6 %
// of the fastest on all 112); the choice is held to 5 % of the best chain by
// tests/test_gpu_step_kernel.py::test_auto_chain_is_within_5_percent_of_the_best.
//   fused    rollout: warps on the fullest sub-partition (rollout_warps_per_sched(); one warp
//            alone costs about as much as two: 1.9) x R x 0.086 us; average eps_bytes / 6.4 TB/s;
//            13.5 us of launches and tails; 7 us more once eps is twice the L2 (the average then
//            competes with the write-back of what the rollout left dirty)
//   unfused  sampler eps_bytes / 5.6 TB/s, rollout eps_bytes / 6.0 TB/s, average eps_bytes / 6.4 TB/s;
//            16.5 + 0.02 R us of launches, tails and one slab's serial chain
//   both     the streaming parts cost 3 % less than the sum of the kernels when replayed as a graph
//   step     15 rollout warps per SM, one tile each per round: max(list/4, ceil(min(list,15)/4))
//            x R x 0.0925 us, then what the consumers still have to read when the last rollout
//            ends -- one round of tiles at ~44 GB/s per SM -- and ~48 us of start-up, merge and
//            U update
struct ChainCost { double fused, unfused, step; };
ChainCost chain_cost(const LaunchCtx &c)
{
    const double R = c.rows, sms = c.num_sms;
    const double eps_bytes = 4.0 * (double)c.k_pad * R;
    const double warps = (double)((c.k_pad / 4 + 31) / 32);              // 128-sample warps
    ChainCost k;
    k.fused = 0.97 * (fmax(rollout_warps_per_sched(c.k_pad, c.num_sms), 1.9) * R * 0.086 +
                      eps_bytes / 6.4e6) + 13.5 + (eps_bytes > 2.5e8 ? 7.0 : 0.0);
    k.unfused = 0.97 * (eps_bytes / 5.6e6 + eps_bytes / 6.0e6 + eps_bytes / 6.4e6) + 16.5 + 0.02 * R;
    const double list = ceil(warps / sms);                               // tiles of the fullest CTA
    const double per_sched = fmax(warps / sms / 4.0, ceil(fmin(list, 15.0) / 4.0));
    const double drain = fmin(warps / sms, 15.0) * R * 512.0 / 44.0e3;
    k.step = per_sched * R * 0.0925 + drain + 48.0;
    return k;
}

uint32_t auto_chain(const LaunchCtx &c, bool can_step)
{
    const ChainCost k = chain_cost(c);
    if (can_step && k.step < k.fused && k.step < k.unfused) return MPPI_FLAG_STEP_KERNEL;
    if (k.fused <= k.unfused) return MPPI_FLAG_FUSED_SAMPLING;
    return MPPI_FLAG_PIPELINED_SAMPLING;
}

int upload_problem(mppi_handle *h)
{
    CK(cudaMemcpyAsync(h->d_prob, &h->h_prob, sizeof(ProblemDev), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int check_handle(mppi_handle *h)
{
    if (!h) return fail(MPPI_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->p.device));
    return MPPI_OK;
}

}  // namespace

// ===================================================================================
// C ABI
// ===================================================================================
extern "C" {

int mppi_abi_version(void) { return MPPI_ABI_VERSION; }

const char *mppi_last_error(void) { return g_last_error.c_str(); }

const char *mppi_kernel_name(int id)
{
    static const char *names[MPPI_K_COUNT] = {"sample", "rollout", "comm_min", "weights",
                                              "average", "comm_sum", "finalize"};
    return (id >= 0 && id < MPPI_K_COUNT) ? names[id] : "?";
}

int mppi_params_default(mppi_params *p)
{
    if (!p) return fail(MPPI_ERR_INVALID, "null params");
    memset(p, 0, sizeof *p);
    p->struct_size = sizeof *p;
    p->lambda = 1.0f;                               /* src/point_mass.cu:53-54 */
    for (int a = 0; a < MPPI_MAX_ACT; ++a) {
        p->sigma[a] = 0.025f;                       /* src/point_mass_gpu.cu:86 */
        p->inv_sigma[a] = 1.0f;                     /* src/point_mass_gpu.cu:58-61 */
        p->init_act[a] = 0.0f;
        p->max_act[a] = 1.0f;
    }
    p->world_size = 1;
    p->comm = MPPI_COMM_NONE;
    p->model = MPPI_MODEL_POINT_MASS;
    p->state_gain[0] = 1.0f; p->state_gain[3] = 1.0f;   /* identity until the caller fills them */
    return MPPI_OK;
}

int mppi_comm_unique_id(uint8_t id[MPPI_COMM_ID_BYTES])
{
    std::string err;
    if (!NcclComm::unique_id(id, err)) return fail(MPPI_ERR_COMM, "%s", err.c_str());
    return MPPI_OK;
}

int mppi_comm_p2p_handle(mppi_handle *h, uint8_t out[MPPI_P2P_HANDLE_BYTES])
{
    if (h && !h->children.empty())
        return fail(MPPI_ERR_INVALID, "a single-process device group connects its shards itself");
    int rc = check_handle(h);
    if (rc) return rc;
    if (!p2p(h) || !out) return fail(MPPI_ERR_INVALID, "handle was not created with MPPI_COMM_P2P");
    static_assert(sizeof(cudaIpcMemHandle_t) == MPPI_P2P_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t ipc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaIpcGetMemHandle(&ipc, h->d_mailbox));
    memcpy(out, &ipc, sizeof ipc);
    return MPPI_OK;
}

int mppi_comm_p2p_connect(mppi_handle *h, const uint8_t *handles)
{
    if (h && !h->children.empty())
        return fail(MPPI_ERR_INVALID, "a single-process device group connects its shards itself");
    int rc = check_handle(h);
    if (rc) return rc;
    if (!p2p(h) || !handles) return fail(MPPI_ERR_INVALID, "handle was not created with MPPI_COMM_P2P");
    for (int r = 0; r < h->p.world_size; ++r) {
        if (r == h->p.rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)r * MPPI_P2P_HANDLE_BYTES, sizeof ipc);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(MPPI_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
        h->peer_mb[r] = static_cast<unsigned long long *>(ptr);
    }
    h->peers_are_ipc = true;
    h->p2p_connected = true;
    return MPPI_OK;
}

int mppi_destroy(mppi_handle *h)
{
    if (!h) return MPPI_OK;
    if (!h->children.empty()) {
        for (mppi_handle *c : h->children)            // all shards idle before any memory goes
            if (c && c->stream) { cudaSetDevice(c->p.device); cudaStreamSynchronize(c->stream); }
        for (mppi_handle *c : h->children) mppi_destroy(c);
        delete h;
        return MPPI_OK;
    }
    cudaSetDevice(h->p.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < kMaxWorld; ++r)
        if (h->peers_are_ipc && h->peer_mb[r] && h->peer_mb[r] != h->d_mailbox)
            cudaIpcCloseMemHandle(h->peer_mb[r]);
    cudaFree(h->d_mailbox);
    for (auto &g : h->graph_exec) if (g) cudaGraphExecDestroy(g);
    for (auto &g : h->graph_pipe) if (g) cudaGraphExecDestroy(g);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    if (h->ev_sampled) cudaEventDestroy(h->ev_sampled);
    if (h->ev_chain_done) cudaEventDestroy(h->ev_chain_done);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    // d_eps / d_eps_alt swap roles every pipelined step: free by base pointer
    if (h->eps_base[1]) { cudaFree(h->eps_base[1]); h->d_eps = h->eps_base[0]; }
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    if (h->t0) cudaEventDestroy(h->t0);
    if (h->t1) cudaEventDestroy(h->t1);
    cudaFree(h->d_eps); cudaFree(h->d_S); cudaFree(h->d_wt); cudaFree(h->d_acc);
    cudaFree(h->d_U); cudaFree(h->d_Uprev); cudaFree(h->d_part);
    cudaFree(h->d_prob); cudaFree(h->d_ctl);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_next) cudaFreeHost(h->h_next);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MPPI_OK;
}

int mppi_create(const mppi_params *params, mppi_handle **out)
{
    if (!params || !out) return fail(MPPI_ERR_INVALID, "null argument");
    *out = nullptr;
    if (params->struct_size != sizeof(mppi_params))
        return fail(MPPI_ERR_INVALID, "mppi_params.struct_size %u != %zu (ABI mismatch)",
                    params->struct_size, sizeof(mppi_params));
    const mppi_params &p = *params;
    if (p.act_dim < 1 || p.act_dim > MPPI_MAX_ACT)
        return fail(MPPI_ERR_INVALID, "act_dim %d not in [1,%d]", p.act_dim, MPPI_MAX_ACT);
    if (p.state_dim != 2 * p.act_dim)
        return fail(MPPI_ERR_INVALID, "state_dim %d must be 2*act_dim (point mass: positions, velocities)",
                    p.state_dim);
    if (p.samples < 1 || p.samples > 0x7fffff00ll)       // TMA box coordinates are int32 sample indices
        return fail(MPPI_ERR_INVALID, "samples %lld out of range [1, 2^31-256]", (long long)p.samples);
    if (p.horizon < 1 || (long long)p.horizon * p.act_dim > 32768)
        return fail(MPPI_ERR_INVALID, "horizon %d out of range", p.horizon);
    if (!(p.lambda > 0.0f)) return fail(MPPI_ERR_INVALID, "lambda must be > 0");
    if (p.philox_rounds != 0 && p.philox_rounds != 7 && p.philox_rounds != 10)
        return fail(MPPI_ERR_INVALID, "philox_rounds %d: 10 (or 0) and 7 exist", p.philox_rounds);
    if (p.model != MPPI_MODEL_POINT_MASS && p.model != MPPI_MODEL_LINEAR_AXIS)
        return fail(MPPI_ERR_INVALID, "model %d unknown", p.model);
    if (p.world_size < 1 || p.rank < 0 || p.rank >= p.world_size)
        return fail(MPPI_ERR_INVALID, "rank %d / world_size %d invalid", p.rank, p.world_size);
    if (p.world_size > 1 && p.comm != MPPI_COMM_NCCL && p.comm != MPPI_COMM_P2P)
        return fail(MPPI_ERR_INVALID, "world_size > 1 needs comm = MPPI_COMM_NCCL or MPPI_COMM_P2P");
    if (p.world_size > kMaxWorld)
        return fail(MPPI_ERR_INVALID, "world_size %d > %d", p.world_size, kMaxWorld);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MPPI_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (p.device < 0 || p.device >= ndev)
        return fail(MPPI_ERR_INVALID, "device %d out of range (%d devices)", p.device, ndev);
    CK(cudaSetDevice(p.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p.device));
    if (prop.major != 10)
        return fail(MPPI_ERR_NO_DEVICE, "device %d is sm_%d%d; this build is sm_100a only",
                    p.device, prop.major, prop.minor);

    mppi_handle *h = new mppi_handle();
    h->p = p;
    // seven Philox rounds exist in the kernel chains only: a one-kernel request becomes the fused chain
    if (p.philox_rounds == 7 && (p.flags & (MPPI_FLAG_STEP_KERNEL | MPPI_FLAG_TILE_KERNEL)))
        h->p.flags = (p.flags & ~(uint32_t)(MPPI_FLAG_STEP_KERNEL | MPPI_FLAG_TILE_KERNEL)) | MPPI_FLAG_FUSED_SAMPLING;
    h->R = p.horizon * p.act_dim;
    h->S = p.state_dim;
    h->injected = (p.flags & MPPI_FLAG_INJECTED_NOISE) != 0;

    // K-shard geometry: whole Philox quads (4 samples) per shard so that eps[k] depends on
    // the global k only, whatever the number of shards.
    int64_t k0 = 0, k1 = 0;
    mppi_shard_range(p.samples, p.rank, p.world_size, &k0, &k1);
    if (k1 <= k0) { delete h; return fail(MPPI_ERR_INVALID, "shard %d of %d is empty", p.rank, p.world_size); }

    LaunchCtx &c = h->ctx;
    c.act_dim = p.act_dim;
    c.horizon = p.horizon;
    c.rows = h->R;
    c.k_local = k1 - k0;
    c.k_offset = k0;
    c.k_pad = (c.k_local + kKPad - 1) / kKPad * kKPad;
    c.seed = p.seed;
    c.strict = (p.flags & MPPI_FLAG_STRICT_ARITH) != 0;
    c.general_gains = p.model == MPPI_MODEL_LINEAR_AXIS;
    c.num_sms = prop.multiProcessorCount;
    const long long ntiles = (c.k_pad / kAvgTileK) * (long long)((h->R + kAvgTileR - 1) / kAvgTileR);
    c.avg_grid = (int)(ntiles < c.num_sms ? ntiles : c.num_sms);
    c.weights_blocks = (int)((c.k_pad + kWeightsBlockSamples - 1) / kWeightsBlockSamples);
    c.sampler = make_sampler_params(p.seed, p.sigma, p.act_dim);
    c.philox_rounds = p.philox_rounds == 7 ? 7 : 10;
    // samples per rollout thread: wide vectors once there are enough samples for many waves
    {
        const double waves4 = (double)c.k_pad / 4.0 / (512.0 * c.num_sms);
        c.rollout_spt = waves4 >= 8.0 ? 4 : (waves4 >= 2.0 ? 2 : 1);
        // below ~2 waves of the register-pipelined kernel the step is latency-bound: the
        // TMA-staged kernel (eps from shared memory, 1 sample per thread) is faster there
        c.rollout_tma = waves4 < 2.0;
        if (const char *env = getenv("MPPI_ROLLOUT_TMA")) c.rollout_tma = atoi(env) != 0;
        // slab width: keep >= 2 CTAs per SM worth of slabs where K allows
        c.rollout_tma_width = 256;
        if (c.k_pad / 256 < 2 * c.num_sms) c.rollout_tma_width = 128;
        if (c.k_pad / 128 < 2 * c.num_sms) c.rollout_tma_width = 64;
        if (const char *env = getenv("MPPI_ROLLOUT_TMA_W")) {
            const int v = atoi(env);
            if (v == 64 || v == 128 || v == 256) c.rollout_tma_width = v;
        }
        if (const char *env = getenv("MPPI_ROLLOUT_SPT")) {
            const int v = atoi(env);
            if (v == 1 || v == 2 || v == 4) c.rollout_spt = v;
        }
    }

    int rc = MPPI_OK;
#define CKH(call)                                                                         \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            rc = fail(MPPI_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,         \
                      cudaGetErrorString(e__));                                           \
            mppi_destroy(h);                                                              \
            return rc;                                                                    \
        }                                                                                 \
    } while (0)

    {
        // the step's critical path outranks the sampler that runs ahead on stream2 (kernel
        // nodes inherit the priority of the stream they were captured on)
        int least = 0, greatest = 0;
        CKH(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CKH(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
    }
    c.stream = h->stream;
    const size_t eps_bytes = sizeof(float) * (size_t)h->R * (size_t)c.k_pad;
    // A handle that runs the on-chip tile kernel never stores eps: its buffer (4*K*T*A bytes --
    // 240 GB at K=1e8, T=200, A=3, more than the GPU has) is only allocated if something asks
    // for the noise in memory (injected noise, the eps / trajectory taps, mppi_sample_only):
    // ensure_eps().  Every other chain needs it from the first step.
    // the one-kernel steps exist with ten Philox rounds only (philox_rounds = 7: kernel chains)
    const bool ten_rounds = p.philox_rounds != 7;
    h->tile_ok = ten_rounds && (p.flags & MPPI_FLAG_TILE_KERNEL) && !(p.flags & MPPI_FLAG_SPLIT_KERNELS) &&
                 !(p.flags & MPPI_FLAG_INJECTED_NOISE) && (p.world_size == 1 || p.comm == MPPI_COMM_P2P) &&
                 tile_kernel_supported(p.horizon, p.act_dim, c.k_pad, c.num_sms);
    if (!h->tile_ok) CKH(cudaMalloc(&h->d_eps, eps_bytes));
    CKH(cudaMalloc(&h->d_S, sizeof(float) * (size_t)c.k_pad));
    if (p.flags & MPPI_FLAG_SPLIT_KERNELS)     // the materialised weights of the split chain only
        CKH(cudaMalloc(&h->d_wt, sizeof(float) * (size_t)c.k_pad));
    CKH(cudaMalloc(&h->d_acc, sizeof(long long) * ((size_t)h->R + 1)));
    CKH(cudaMalloc(&h->d_U, sizeof(float) * (size_t)h->R));
    CKH(cudaMalloc(&h->d_Uprev, sizeof(float) * (size_t)h->R));
    if (p.world_size > 1 && p.comm == MPPI_COMM_P2P) {
        const size_t mb_bytes = sizeof(unsigned long long) * mailbox_slot_words(h->R) * p.world_size *
                                kMailboxBuffers;
        CKH(cudaMalloc(&h->d_mailbox, mb_bytes));
        CKH(cudaMemsetAsync(h->d_mailbox, 0, mb_bytes, h->stream));
        h->peer_mb[p.rank] = h->d_mailbox;
    }
    CKH(cudaMalloc(&h->d_prob, sizeof(ProblemDev)));
    CKH(cudaMalloc(&h->d_ctl, sizeof(CtlDev)));
    CKH(cudaMallocHost(&h->h_stage, sizeof(float) * kStageSlots * 2 * kMaxAct));
    CKH(cudaHostAlloc(&h->h_next, sizeof(float) * kNextFloats, cudaHostAllocMapped));
    memset(h->h_next, 0, sizeof(float) * kNextFloats);
    CKH(cudaHostGetDevicePointer(&h->d_next, h->h_next, 0));
    // eps zeroed like the reference's cudaMemset(_e, 0) (src/point_mass.cu:69): keeps the
    // pad columns finite and defines injected-noise mode before the first mppi_set_noise.
    if (h->d_eps) CKH(cudaMemsetAsync(h->d_eps, 0, eps_bytes, h->stream));
    CKH(cudaMemsetAsync(h->d_S, 0, sizeof(float) * (size_t)c.k_pad, h->stream));
    if (h->d_wt) CKH(cudaMemsetAsync(h->d_wt, 0, sizeof(float) * (size_t)c.k_pad, h->stream));
    CKH(cudaMemsetAsync(h->d_acc, 0, sizeof(long long) * ((size_t)h->R + 1), h->stream));
    CKH(cudaMemsetAsync(h->d_U, 0, sizeof(float) * (size_t)h->R, h->stream));
    CKH(cudaMemsetAsync(h->d_Uprev, 0, sizeof(float) * (size_t)h->R, h->stream));
    CKH(launch_clear_ctl(c, h->d_ctl));
    for (auto &e : h->ev) CKH(cudaEventCreate(&e));
    CKH(cudaEventCreate(&h->t0));
    CKH(cudaEventCreate(&h->t1));
    {
        size_t need = 0, have = 0;
        if (const char *k = check_smem_requirements(c, &need, &have)) {
            rc = fail(MPPI_ERR_INVALID, "horizon %d x act_dim %d: %s needs %zu bytes of shared memory, "
                      "the device offers %zu per block", p.horizon, p.act_dim, k, need, have);
            mppi_destroy(h);
            return rc;
        }
    }
    CKH(configure_kernels(c));
    if (p.flags & MPPI_FLAG_AUTO_CHAIN) {
        const bool can_step = ten_rounds && (p.world_size == 1 || p.comm == MPPI_COMM_P2P) &&
                              step_kernel_supported(p.horizon, p.act_dim, c.k_pad, c.num_sms);
        if (!(p.flags & MPPI_FLAG_SPLIT_KERNELS)) h->p.flags |= auto_chain(c, can_step);
        h->p.flags &= ~MPPI_FLAG_AUTO_CHAIN;
    }
    if ((h->p.flags & MPPI_FLAG_PIPELINED_SAMPLING) &&
        !(h->p.flags & (MPPI_FLAG_FUSED_SAMPLING | MPPI_FLAG_STEP_KERNEL | MPPI_FLAG_TILE_KERNEL))) {
        int least = 0, greatest = 0;
        CKH(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CKH(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, least));
        CKH(cudaEventCreateWithFlags(&h->ev_sampled, cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&h->ev_chain_done, cudaEventDisableTiming));
        CKH(cudaMalloc(&h->d_eps_alt, eps_bytes));
        h->eps_base[0] = h->d_eps;
        h->eps_base[1] = h->d_eps_alt;
        CKH(cudaMemsetAsync(h->d_eps_alt, 0, eps_bytes, h->stream));
    }
    h->step_ok = ten_rounds && (h->p.flags & MPPI_FLAG_STEP_KERNEL) &&
                 step_kernel_supported(p.horizon, p.act_dim, c.k_pad, c.num_sms);
    if (h->step_ok) CKH(configure_step(c));
    if (h->tile_ok) CKH(configure_tile(c));
    if (h->step_ok || h->tile_ok) {
        CKH(cudaMalloc(&h->d_part, sizeof(float) * step_part_floats(c)));
        CKH(cudaMemsetAsync(h->d_part, 0, sizeof(float) * step_part_floats(c), h->stream));
    }
#undef CKH

    // problem constants: gains as the reference forms them (src/point_mass.cu:46-51)
    ProblemDev &pd = h->h_prob;
    memset(&pd, 0, sizeof pd);
    const float dt = p.dt;
    const float dt2 = dt * dt;
    pd.g[0] = 1.0f; pd.g[1] = dt; pd.g[2] = 0.0f; pd.g[3] = 1.0f;
    pd.b[0] = (float)((double)dt2 / 2.0);
    pd.b[1] = dt;
    if (p.model == MPPI_MODEL_LINEAR_AXIS) {
        for (int i = 0; i < 4; ++i) pd.g[i] = p.state_gain[i];
        pd.b[0] = p.act_gain[0];
        pd.b[1] = p.act_gain[1];
    }
    pd.lambda = p.lambda;
    pd.neg_inv_lambda = -(1 / p.lambda);
    for (int a = 0; a < p.act_dim; ++a) {
        pd.inv_s[a] = p.inv_sigma[a];
        pd.sigma[a] = p.sigma[a];
        pd.init_act[a] = p.init_act[a];
        pd.max_act[a] = p.max_act[a];
    }
    if ((rc = upload_problem(h)) != MPPI_OK) { mppi_destroy(h); return rc; }
    if (h->d_eps && (rc = encode_eps_tmaps(h)) != MPPI_OK) { mppi_destroy(h); return rc; }
    if (h->d_eps_alt) {
        if ((rc = encode_tmap(h, &h->tmap_alt, kAvgTileK, kAvgTileR, h->d_eps_alt)) != MPPI_OK) { mppi_destroy(h); return rc; }
        if ((rc = encode_tmap(h, &h->tmap_ro_alt, c.rollout_tma_width, rollout_tma_rows(p.act_dim), h->d_eps_alt)) != MPPI_OK) { mppi_destroy(h); return rc; }
    }

    if (multi(h) && p.comm == MPPI_COMM_NCCL) {
        std::string err;
        if (!h->comm.init(p.rank, p.world_size, p.comm_id, err)) {
            rc = fail(MPPI_ERR_COMM, "%s", err.c_str());
            mppi_destroy(h);
            return rc;
        }
    }
    if (p.verbose)
        fprintf(stderr, "[mppi_b200] device %d (%s, %d SMs) shard %d/%d: K_local=%lld (offset %lld, pad %lld) "
                        "T=%d A=%d eps=%.1f MB avg_grid=%d\n",
                p.device, prop.name, c.num_sms, p.rank, p.world_size, (long long)c.k_local,
                (long long)c.k_offset, (long long)c.k_pad, p.horizon, p.act_dim,
                eps_bytes / 1048576.0, c.avg_grid);
    *out = h;
    return MPPI_OK;
}

int mppi_create_multi(const mppi_params *params, const int *devices, int num_devices,
                      mppi_handle **out)
{
    if (!params || !devices || !out) return fail(MPPI_ERR_INVALID, "null argument");
    *out = nullptr;
    if (num_devices < 1 || num_devices > kMaxWorld)
        return fail(MPPI_ERR_INVALID, "num_devices %d not in [1,%d]", num_devices, kMaxWorld);
    if (params->struct_size != sizeof(mppi_params))
        return fail(MPPI_ERR_INVALID, "mppi_params.struct_size mismatch (ABI)");
    if (num_devices == 1) {
        mppi_params p1 = *params;
        p1.device = devices[0]; p1.rank = 0; p1.world_size = 1; p1.comm = MPPI_COMM_NONE;
        return mppi_create(&p1, out);
    }
    for (int i = 0; i < num_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j])
                return fail(MPPI_ERR_INVALID, "device %d listed twice", devices[i]);
    mppi_handle *g = new mppi_handle();
    g->p = *params;
    g->p.world_size = num_devices;
    g->p.comm = MPPI_COMM_P2P;
    g->R = params->horizon * params->act_dim;
    g->S = params->state_dim;
    int rc = MPPI_OK;
    for (int i = 0; i < num_devices && rc == MPPI_OK; ++i) {
        mppi_params pc = *params;
        pc.device = devices[i]; pc.rank = i; pc.world_size = num_devices; pc.comm = MPPI_COMM_P2P;
        mppi_handle *c = nullptr;
        rc = mppi_create(&pc, &c);
        if (rc == MPPI_OK) g->children.push_back(c);
    }
    // direct peer mappings between all pairs (same process: no IPC needed)
    for (int i = 0; i < num_devices && rc == MPPI_OK; ++i) {
        cudaSetDevice(devices[i]);
        for (int j = 0; j < num_devices && rc == MPPI_OK; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (!can) { rc = fail(MPPI_ERR_COMM, "device %d cannot access device %d", devices[i], devices[j]); break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess)
                rc = fail(MPPI_ERR_COMM, "cudaDeviceEnablePeerAccess(%d->%d) -> %s", devices[i],
                          devices[j], cudaGetErrorString(e));
        }
    }
    if (rc != MPPI_OK) { mppi_destroy(g); return rc; }
    for (int i = 0; i < num_devices; ++i) {
        for (int j = 0; j < num_devices; ++j) g->children[i]->peer_mb[j] = g->children[j]->d_mailbox;
        g->children[i]->p2p_connected = true;
    }
    *out = g;
    return MPPI_OK;
}

int mppi_shard_range(int64_t samples, int rank, int world_size, int64_t *k_begin, int64_t *k_end)
{
    if (samples < 1 || world_size < 1 || rank < 0 || rank >= world_size || !k_begin || !k_end)
        return fail(MPPI_ERR_INVALID, "mppi_shard_range: bad arguments");
    const int64_t quads = (samples + 3) / 4;
    const int64_t q0 = quads * rank / world_size;
    const int64_t q1 = quads * (rank + 1) / world_size;
    *k_begin = 4 * q0;
    *k_end = (4 * q1 < samples) ? 4 * q1 : samples;
    return MPPI_OK;
}

int mppi_chain_estimate(int64_t samples_local, int horizon, int act_dim, int num_sms,
                        double est_us[3], uint32_t *choice)
{
    if (samples_local < 1 || horizon < 1 || act_dim < 1 || act_dim > MPPI_MAX_ACT || num_sms < 1)
        return fail(MPPI_ERR_INVALID, "mppi_chain_estimate: bad arguments");
    LaunchCtx c{};
    c.k_local = samples_local;
    c.k_pad = (samples_local + kKPad - 1) / kKPad * kKPad;
    c.horizon = horizon;
    c.act_dim = act_dim;
    c.rows = horizon * act_dim;
    c.num_sms = num_sms;
    const ChainCost k = chain_cost(c);
    if (est_us) { est_us[0] = k.unfused; est_us[1] = k.fused; est_us[2] = k.step; }
    if (choice) *choice = auto_chain(c, step_kernel_supported(horizon, act_dim, c.k_pad, num_sms));
    return MPPI_OK;
}

int mppi_local_samples(mppi_handle *h, int64_t *k_local, int64_t *k_offset)
{
    if (h && !h->children.empty()) {
        if (k_local) *k_local = h->p.samples;
        if (k_offset) *k_offset = 0;
        return MPPI_OK;
    }
    if (!h) return fail(MPPI_ERR_INVALID, "null handle");
    if (k_local) *k_local = h->ctx.k_local;
    if (k_offset) *k_offset = h->ctx.k_offset;
    return MPPI_OK;
}

int mppi_set_problem(mppi_handle *h, const float *x0, const float *u, const float *goal,
                     const float *w)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) {
            int rcc = mppi_set_problem(c, x0, u, goal, w);
            if (rcc) return rcc;
        }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (!x0 || !u || !goal || !w) return fail(MPPI_ERR_INVALID, "null argument");
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < h->S; ++i) {
        h->h_prob.x0[i] = x0[i];
        h->h_prob.goal[i] = goal[i];
        h->h_prob.w[i] = w[i];
        if (!h->terminal_set) h->h_prob.wf[i] = w[i];    // one Cost object: src/point_mass_gpu.cu:116
    }
    if ((rc = upload_problem(h)) != MPPI_OK) return rc;
    CK(cudaMemcpyAsync(h->d_U, u, sizeof(float) * h->R, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_Uprev, u, sizeof(float) * h->R, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->problem_set = true;
    return MPPI_OK;
}

int mppi_set_terminal_weights(mppi_handle *h, const float *w_final)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) {
            int rcc = mppi_set_terminal_weights(c, w_final);
            if (rcc) return rcc;
        }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->terminal_set = w_final != nullptr;
    for (int i = 0; i < h->S; ++i) h->h_prob.wf[i] = w_final ? w_final[i] : h->h_prob.w[i];
    return upload_problem(h);
}

int mppi_set_u(mppi_handle *h, const float *u)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) { int rcc = mppi_set_u(c, u); if (rcc) return rcc; }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (!u) return fail(MPPI_ERR_INVALID, "null argument");
    CK(cudaMemcpyAsync(h->d_U, u, sizeof(float) * h->R, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int mppi_set_state(mppi_handle *h, const float *x)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) { int rcc = mppi_set_state(c, x); if (rcc) return rcc; }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (!x) return fail(MPPI_ERR_INVALID, "null argument");
    // A staging slot may be rewritten once its async copy has run.  The copy of a slot staged
    // when `tag` steps had been enqueued precedes step tag+1 on the stream, so it is done as
    // soon as that step has published its sequence number (mapped host memory, no API call);
    // otherwise -- no step behind it yet, or still running -- settle it the slow way.  Only
    // reached after kStageSlots set_state calls without a finished step in between.
    if (h->stage_used[h->stage_slot]) {
        const unsigned long long seen = __atomic_load_n(
            reinterpret_cast<unsigned long long *>(h->h_next + kNextSeqOffset), __ATOMIC_ACQUIRE);
        if (seen < h->stage_tag[h->stage_slot] + 1) CK(cudaStreamSynchronize(h->stream));
    }
    h->stage_used[h->stage_slot] = true;
    h->stage_tag[h->stage_slot] = h->steps_enqueued;
    float *slot = h->h_stage + (size_t)h->stage_slot * 2 * kMaxAct;
    h->stage_slot = (h->stage_slot + 1) % kStageSlots;
    for (int i = 0; i < h->S; ++i) { slot[i] = x[i]; h->h_prob.x0[i] = x[i]; }
    CK(cudaMemcpyAsync(reinterpret_cast<char *>(h->d_prob) + offsetof(ProblemDev, x0), slot,
                       sizeof(float) * h->S, cudaMemcpyHostToDevice, h->stream));
    return MPPI_OK;
}

int mppi_step_enqueue(mppi_handle *h)
{
    if (h && !h->children.empty()) {
        // every shard's graph is enqueued before anything waits: the shards meet in the
        // peer-mailbox exchanges on the devices
        for (mppi_handle *c : h->children) { int rcc = mppi_step_enqueue(c); if (rcc) return rcc; }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (!h->problem_set) return fail(MPPI_ERR_STATE, "mppi_step before mppi_set_problem");
    if (p2p(h) && !h->p2p_connected)
        return fail(MPPI_ERR_STATE, "mppi_step before mppi_comm_p2p_connect");
    const bool sample = !h->injected;
    const int which = sample ? 0 : 1;
    if (!tile_step(h, sample) && (rc = ensure_eps(h)) != MPPI_OK) return rc;
    if (pipelined(h, sample)) {
        // step n: the chain reads d_eps, which the sampler filled after step n-1 had published
        // its action -- during the plant's turn; the sampler of step n+1 is queued behind this
        // step's chain on stream2 and fills d_eps_alt (two buffers so that get_inf still
        // returns the noise the last step consumed).  Queued BEHIND the chain, not beside it:
        // run side by side the sampler takes from the chain exactly what it saves (measured,
        // DESIGN.md section 4c), so the gain is latency, not throughput.
        if (!h->presampled) {
            // first pipelined step, or the first after a step of another kind: draw in line
            CK(launch_sample(h->ctx, h->d_eps, h->d_ctl, false, 0));
            h->total_launches += 1;
        } else {
            CK(cudaStreamWaitEvent(h->stream, h->ev_sampled, 0));
        }
        const int base = h->d_eps == h->eps_base[0] ? 0 : 1;
        if (!h->graph_pipe[base] && (rc = build_pipe_graph(h, base)) != MPPI_OK) return rc;
        CK(cudaGraphLaunch(h->graph_pipe[base], h->stream));
        CK(cudaEventRecord(h->ev_chain_done, h->stream));
        // off the critical path: the next step's noise.  The Philox step index is the host's
        // count of enqueued steps, which the device counter equals when that step runs.
        LaunchCtx side = h->ctx;
        side.stream = h->stream2;
        CK(cudaStreamWaitEvent(h->stream2, h->ev_chain_done, 0));
        CK(launch_sample(side, h->d_eps_alt, h->d_ctl, true, h->steps_enqueued + 1));
        CK(cudaEventRecord(h->ev_sampled, h->stream2));
        h->eps_last = h->d_eps;
        h->eps_on_chip = false;
        std::swap(h->d_eps, h->d_eps_alt);
        std::swap(h->tmap, h->tmap_alt);
        std::swap(h->tmap_ro, h->tmap_ro_alt);
        h->presampled = true;
        h->total_launches += kernels_per_step(h, true);
        h->steps_enqueued += 1;
        h->pending = true;
        return MPPI_OK;
    }
    // a step of any other kind advances the step counter: noise drawn ahead is void
    // (the plain chain draws the same values again)
    if ((rc = leave_pipeline(h)) != MPPI_OK) return rc;
    h->eps_last = h->d_eps;
    h->eps_on_chip = tile_step(h, sample);
    h->eps_step = h->steps_enqueued;          // the Philox step index this step draws with
    if (h->profiling || (h->p.flags & MPPI_FLAG_NO_GRAPH)) {
        rc = enqueue_chain(h, sample, h->profiling ? h->ev : nullptr);
        if (rc) return rc;
        if (h->profiling) { h->prof_pending = true; h->prof_sampled = sample; h->prof_one_kernel = one_kernel(h, sample); }
    } else {
        if (!h->graph_exec[which] && (rc = build_graph(h, which)) != MPPI_OK) return rc;
        CK(cudaGraphLaunch(h->graph_exec[which], h->stream));
    }
    h->total_launches += kernels_per_step(h, sample);
    h->steps_enqueued += 1;
    h->pending = true;
    return MPPI_OK;
}

int mppi_step_wait(mppi_handle *h, float *next_act)
{
    if (h && !h->children.empty()) {
        int first_err = MPPI_OK;
        for (size_t i = 0; i < h->children.size(); ++i) {
            int rcc = mppi_step_wait(h->children[i], i == 0 ? next_act : nullptr);
            if (rcc && !first_err) first_err = rcc;
        }
        return first_err;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (h->prof_pending || (h->p.flags & MPPI_FLAG_NO_GRAPH)) {
        CK(cudaStreamSynchronize(h->stream));
    } else {
        // spin on the step counter the finalizing kernel publishes in mapped host memory; look
        // at the stream now and then so that a failed launch cannot hang the caller
        volatile unsigned long long *seq =
            reinterpret_cast<volatile unsigned long long *>(h->h_next + kNextSeqOffset);
        for (unsigned spins = 0; *seq != h->steps_enqueued; ++spins) {
            if ((spins & 1023u) == 1023u) {
                cudaError_t q = cudaStreamQuery(h->stream);
                if (q != cudaErrorNotReady) {            // finished or failed: settle it the slow way
                    CK(cudaStreamSynchronize(h->stream));
                    break;
                }
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (*seq != h->steps_enqueued)
            return fail(MPPI_ERR_STATE, "step %llu finished without publishing its result (seq %llu)",
                        h->steps_enqueued, (unsigned long long)*seq);
    }
    h->pending = false;
    if (h->prof_pending) {
        // event i+1 closes stage i of {sample, rollout, comm_min, weights, average,
        // comm_sum, finalize(+D2H)}
        h->prof_pending = false;
        const bool sample = h->prof_sampled;
        for (int i = 0; i < MPPI_K_COUNT; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
            const bool split = (h->p.flags & MPPI_FLAG_SPLIT_KERNELS) != 0;
            const bool ran = h->prof_one_kernel ? (i == MPPI_K_AVERAGE)
                           : (i == MPPI_K_SAMPLE) ? (sample && !fused(h))
                           : (i == MPPI_K_COMM_MIN) ? (multi(h) && (!p2p(h) || split))
                           : (i == MPPI_K_COMM_SUM) ? (multi(h) && (!p2p(h) || split))
                           : (i == MPPI_K_WEIGHTS) ? split
                           : (i == MPPI_K_FINALIZE) ? ((split || multi(h)) && !p2p(h)) : true;
            if (ran) { h->ms_sum[i] += ms; h->launches[i] += 1; }
        }
    }
    // a failed exchange has not touched U and has published no action: nothing is copied to
    // the caller, and the handle (and its peers') is unusable from here on
    if (h->h_next[kMaxAct] != 0.0f)
        return fail(MPPI_ERR_COMM, "peer-mailbox exchange timed out (a rank did not arrive); "
                                   "U was left unchanged, the handle cannot be used any further");
    if (next_act)
        for (int a = 0; a < h->p.act_dim; ++a) next_act[a] = h->h_next[a];
    return MPPI_OK;
}

int mppi_step(mppi_handle *h, float *next_act)
{
    int rc = mppi_step_enqueue(h);
    if (rc) return rc;
    return mppi_step_wait(h, next_act);
}

int mppi_get_u(mppi_handle *h, float *u)
{
    if (h && !h->children.empty()) return mppi_get_u(h->children[0], u);
    int rc = check_handle(h);
    if (rc) return rc;
    if (!u) return fail(MPPI_ERR_INVALID, "null argument");
    CK(cudaMemcpyAsync(u, h->d_U, sizeof(float) * h->R, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int mppi_get_flags(mppi_handle *h, uint32_t *flags)
{
    if (h && !h->children.empty()) return mppi_get_flags(h->children[0], flags);
    if (!h || !flags) return fail(MPPI_ERR_INVALID, "null argument");
    *flags = h->p.flags;
    return MPPI_OK;
}

int mppi_get_step_info(mppi_handle *h, mppi_step_info *info)
{
    if (h && !h->children.empty()) return mppi_get_step_info(h->children[0], info);
    int rc = check_handle(h);
    if (rc) return rc;
    if (!info) return fail(MPPI_ERR_INVALID, "null argument");
    CtlDev ctl;
    CK(cudaMemcpyAsync(&ctl, h->d_ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    info->beta = ordered_to_float((uint32_t)(ctl.last_key >> 32));
    info->argmin = (int64_t)(ctl.last_key & 0xffffffffull);
    info->eta = ctl.eta;
    info->step = ctl.step;
    if (ctl.last_key == kMinKeyInit) { info->beta = NAN; info->argmin = -1; }
    return MPPI_OK;
}

int mppi_get_info(mppi_handle *h, float *x, float *u, float *e, float *cost, float *beta,
                  float *nabla, float *weight)
{
    if (h && !h->children.empty()) {
        // shards are contiguous in k: every per-sample array is filled shard by shard
        for (size_t i = 0; i < h->children.size(); ++i) {
            mppi_handle *c = h->children[i];
            const size_t k0 = (size_t)c->ctx.k_offset, R = (size_t)c->R, S = (size_t)c->S;
            const size_t T1 = (size_t)c->p.horizon + 1;
            int rcc = mppi_get_info(c, x ? x + k0 * T1 * S : nullptr, i == 0 ? u : nullptr,
                                    e ? e + k0 * R : nullptr, cost ? cost + k0 : nullptr,
                                    i == 0 ? beta : nullptr, i == 0 ? nabla : nullptr,
                                    weight ? weight + k0 : nullptr);
            if (rcc) return rcc;
        }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    const LaunchCtx &c = h->ctx;
    CK(cudaStreamSynchronize(h->stream));
    mppi_step_info info;
    if ((rc = mppi_get_step_info(h, &info)) != MPPI_OK) return rc;
    if (beta) *beta = info.beta;
    if (nabla) *nabla = info.eta;
    if (u) CK(cudaMemcpyAsync(u, h->d_U, sizeof(float) * h->R, cudaMemcpyDeviceToHost, h->stream));
    if (cost)
        CK(cudaMemcpyAsync(cost, h->d_S, sizeof(float) * (size_t)c.k_local, cudaMemcpyDeviceToHost,
                           h->stream));
    CK(cudaStreamSynchronize(h->stream));

    float *scratch = nullptr;
    auto release = [&]() { if (scratch) { cudaFree(scratch); scratch = nullptr; } };
#define CKS(call)                                                                         \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            release();                                                                    \
            return fail(MPPI_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                        cudaGetErrorString(e__));                                         \
        }                                                                                 \
    } while (0)
    if ((e || x) && (rc = ensure_eps(h)) != MPPI_OK) return rc;
    if ((e || x) && h->eps_on_chip) {
        // the last step kept its noise in shared memory: draw it again (counter based: same bits)
        CKS(launch_sample(c, h->d_eps, h->d_ctl, true, h->eps_step));
        h->total_launches += 1;
        h->eps_on_chip = false;               // d_eps now holds that step's noise
    }
    if (weight) {
        CKS(cudaMalloc(&scratch, sizeof(float) * (size_t)c.k_local));
        CKS(launch_norm_weights(c, h->d_S, h->p.lambda, info.beta, info.eta, scratch));
        CKS(cudaMemcpyAsync(weight, scratch, sizeof(float) * (size_t)c.k_local,
                            cudaMemcpyDeviceToHost, h->stream));
        CKS(cudaStreamSynchronize(h->stream));
        release();
    }
    if (e) {
        const size_t n = (size_t)c.k_local * h->R;
        CKS(cudaMalloc(&scratch, sizeof(float) * n));
        CKS(launch_to_reference(c, last_eps(h), scratch));
        CKS(cudaMemcpyAsync(e, scratch, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
        CKS(cudaStreamSynchronize(h->stream));
        release();
    }
    if (x) {
        const size_t n = (size_t)c.k_local * (h->p.horizon + 1) * h->S;
        CKS(cudaMalloc(&scratch, sizeof(float) * n));
        CKS(launch_trajectories(c, last_eps(h), h->d_Uprev, h->d_prob, scratch));
        CKS(cudaMemcpyAsync(x, scratch, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
        CKS(cudaStreamSynchronize(h->stream));
        release();
    }
#undef CKS
    return MPPI_OK;
}

int mppi_set_noise_mode(mppi_handle *h, int injected)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) mppi_set_noise_mode(c, injected);
        return MPPI_OK;
    }
    if (!h) return fail(MPPI_ERR_INVALID, "null handle");
    h->injected = injected != 0;
    return MPPI_OK;
}

int mppi_set_noise(mppi_handle *h, const float *e)
{
    if (h && !h->children.empty()) {
        if (!e) return fail(MPPI_ERR_INVALID, "null argument");
        for (mppi_handle *c : h->children) {
            int rcc = mppi_set_noise(c, e + (size_t)c->ctx.k_offset * c->R);
            if (rcc) return rcc;
        }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    if (!e) return fail(MPPI_ERR_INVALID, "null argument");
    const LaunchCtx &c = h->ctx;
    const size_t n = (size_t)c.k_local * h->R;
    if ((rc = leave_pipeline(h)) != MPPI_OK) return rc;   // d_eps may be receiving noise drawn ahead
    if ((rc = ensure_eps(h)) != MPPI_OK) return rc;
    h->eps_last = h->d_eps;
    h->eps_on_chip = false;
    float *scratch = nullptr;
    CK(cudaMalloc(&scratch, sizeof(float) * n));
    cudaError_t err = cudaMemcpyAsync(scratch, e, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream);
    if (err == cudaSuccess) err = launch_to_internal(c, scratch, h->d_eps);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    cudaFree(scratch);
    if (err != cudaSuccess)
        return fail(MPPI_ERR_CUDA, "mppi_set_noise -> %s", cudaGetErrorString(err));
    h->injected = true;
    return MPPI_OK;
}

int mppi_sample_only(mppi_handle *h, uint64_t step)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) { int rcc = mppi_sample_only(c, step); if (rcc) return rcc; }
        return MPPI_OK;
    }
    int rc = check_handle(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = leave_pipeline(h)) != MPPI_OK) return rc;
    if ((rc = ensure_eps(h)) != MPPI_OK) return rc;
    h->eps_last = h->d_eps;                              // what get_inf returns next
    h->eps_on_chip = false;
    CK(launch_sample(h->ctx, h->d_eps, h->d_ctl, true, step));
    CK(cudaStreamSynchronize(h->stream));
    h->total_launches += 1;
    return MPPI_OK;
}

int mppi_timer_start(mppi_handle *h)
{
    if (h && !h->children.empty()) return mppi_timer_start(h->children[0]);
    int rc = check_handle(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->t0, h->stream));
    return MPPI_OK;
}

int mppi_timer_stop(mppi_handle *h, float *elapsed_ms)
{
    if (h && !h->children.empty()) return mppi_timer_stop(h->children[0], elapsed_ms);
    int rc = check_handle(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->t1, h->stream));
    CK(cudaEventSynchronize(h->t1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->t0, h->t1));
    if (elapsed_ms) *elapsed_ms = ms;
    return MPPI_OK;
}

int mppi_set_profiling(mppi_handle *h, int enabled)
{
    if (h && !h->children.empty()) {
        for (mppi_handle *c : h->children) mppi_set_profiling(c, enabled);
        return MPPI_OK;
    }
    if (!h) return fail(MPPI_ERR_INVALID, "null handle");
    h->profiling = enabled != 0;
    return MPPI_OK;
}

int mppi_get_kernel_times(mppi_handle *h, double *ms_sum, int64_t *launches)
{
    if (h && !h->children.empty()) {
        for (size_t i = 1; i < h->children.size(); ++i) mppi_get_kernel_times(h->children[i], nullptr, nullptr);
        return mppi_get_kernel_times(h->children[0], ms_sum, launches);
    }
    if (!h) return fail(MPPI_ERR_INVALID, "null handle");
    for (int i = 0; i < MPPI_K_COUNT; ++i) {
        if (ms_sum) ms_sum[i] = h->ms_sum[i];
        if (launches) launches[i] = h->launches[i];
        h->ms_sum[i] = 0.0;
        h->launches[i] = 0;
    }
    return MPPI_OK;
}

int mppi_get_exchange_times(mppi_handle *h, double us[3])
{
    if (h && !h->children.empty()) return mppi_get_exchange_times(h->children[0], us);
    int rc = check_handle(h);
    if (rc) return rc;
    if (!us) return fail(MPPI_ERR_INVALID, "null argument");
    CtlDev ctl;
    CK(cudaMemcpyAsync(&ctl, h->d_ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    us[0] = us[1] = us[2] = 0.0;
    if (!p2p(h) || ctl.t_xchg[3] == 0) return MPPI_OK;
    us[0] = (double)(ctl.t_xchg[1] - ctl.t_xchg[0]) * 1e-3;   // push own data + flags
    us[1] = (double)(ctl.t_xchg[2] - ctl.t_xchg[1]) * 1e-3;   // wait for the slowest peer
    us[2] = (double)(ctl.t_xchg[3] - ctl.t_xchg[2]) * 1e-3;   // rescale + sum
    return MPPI_OK;
}

int mppi_get_launch_count(mppi_handle *h, int64_t *launches)
{
    if (h && launches && !h->children.empty()) {
        *launches = 0;
        for (mppi_handle *c : h->children) *launches += c->total_launches;
        return MPPI_OK;
    }
    if (!h || !launches) return fail(MPPI_ERR_INVALID, "null argument");
    *launches = h->total_launches;
    return MPPI_OK;
}

}  // extern "C"
