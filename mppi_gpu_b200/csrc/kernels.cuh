// kernels.cuh -- launch interface of the sm_100a kernels (definitions in kernels.cu).
#pragma once

#include "common.cuh"
#include "philox.cuh"
#include "xchg.cuh"

namespace mppi {

// Tile geometry of the weighted-average kernel (part 4).
constexpr int kAvgTileK    = 256;   // samples per tile (TMA box inner dim, 1 KB per row)
constexpr int kAvgTileR    = 40;    // eps rows per tile: divides T*A = 200/400/600
constexpr int kAvgStages   = 4;     // TMA pipeline depth: 4 x 41 KB in flight per SM
constexpr int kAvgConsumerWarps = 8;
constexpr int kAvgThreads  = (kAvgConsumerWarps + 1) * 32;   // + 1 producer warp
constexpr int kKPad        = 256;   // eps leading dimension is a multiple of this
constexpr int kWeightsBlockSamples = 1024;   // samples per CTA in the weights kernel

size_t average_smem_bytes(int R);

struct LaunchCtx {
    cudaStream_t stream;
    int act_dim;           // A
    int horizon;           // T
    int rows;              // R = T*A
    int64_t k_local;       // samples of this shard
    int64_t k_pad;         // leading dimension of eps / weights (multiple of kKPad)
    int64_t k_offset;      // global index of local sample 0 (multiple of 4)
    unsigned long long seed;
    bool strict;           // MPPI_FLAG_STRICT_ARITH
    bool general_gains;    // MPPI_MODEL_LINEAR_AXIS: LinearAxis dynamics instead of DoubleIntegrator
    int  num_sms;
    int  avg_grid;         // CTAs of the averaging kernel
    int  weights_blocks;   // CTAs of the weights kernel
    int  rollout_spt;      // samples per thread in the rollout kernel (1, 2 or 4)
    bool rollout_tma;      // use the TMA-staged rollout kernel (injected / unfused sampling)
    int  rollout_tma_width; // its slab width (TMA box inner dim): 64, 128 or 256 samples
    SamplerParams sampler; // Philox round keys of the seed, sigma-derived constants
    int  philox_rounds;    // 10 (default) or 7: sample_kernel and the fused rollout only
};

// (1) eps[r][k] = sigma[a] * N(0,1), Philox counter (k/4, r, step)
cudaError_t launch_sample(const LaunchCtx &c, float *eps, const CtlDev *ctl,
                          bool use_step_override, unsigned long long step_override);

// (2) S[k] = rollout cost; block min -> atomicMin(ctl->min_key).  fused: also samples eps.
cudaError_t launch_rollout(const LaunchCtx &c, float *eps, const float *U, const ProblemDev *prob,
                           float *S, CtlDev *ctl, bool fused_sampling);

// CTA size of rollout_kernel for a shard of k_pad4 samples at four per thread, and the number of
// its 128-sample warps on the fullest SM sub-partition (what the fused pass costs, in units of
// one warp's horizon)
int rollout_block_threads(long long k_pad4, int num_sms);
double rollout_warps_per_sched(long long k_pad, int num_sms);

// (2b) the same rollout fed by TMA tiles (tensor map with box {rollout_tma_width,
//      rollout_tma_rows(A)})
cudaError_t launch_rollout_tma(const LaunchCtx &c, const CUtensorMap &tmap, const float *U,
                               const ProblemDev *prob, float *S, CtlDev *ctl);
int rollout_tma_rows(int A);

// (3) wt[k] = expf(-(1/lambda)(S[k]-beta)), acc[R] += eta partial (fixed point)
cudaError_t launch_weights(const LaunchCtx &c, const float *S, const ProblemDev *prob,
                           const CtlDev *ctl, float *wt, long long *acc);

// (4) acc[r] += sum_{k in cta's tiles} wt[k] * eps[r][k]   (fixed point, r < R)
//     merge_weights : src = S, the weights (3) are formed inside (acc[R] += eta as well);
//                     otherwise src = wt from launch_weights
//     merge_finalize: the last CTA also runs (5); with xa.world > 1 (K-shards over NVLink peer
//                     memory) it runs the single exchange of xchg.cuh first, in line
cudaError_t launch_average(const LaunchCtx &c, const CUtensorMap &tmap_eps, const float *src,
                           long long *acc, bool merge_weights, bool merge_finalize,
                           const ProblemDev *prob, CtlDev *ctl, float *U, float *U_prev,
                           float *next_act, unsigned flags, const XchgArgs &xa, bool pdl = false);
//     pdl: the kernel launched in front of it on the same stream is one of the rollout kernels
//          (no collective, no weights kernel in between): programmatic dependent launch

// (5) U += acc[0..R-1]/acc[R]; shift, next_act, advance step, re-arm acc and min key
cudaError_t launch_finalize(const LaunchCtx &c, long long *acc, float *U, float *U_prev,
                            const ProblemDev *prob, CtlDev *ctl, float *next_act, unsigned flags);

// K-shard exchange over NVLink peer mailboxes (MPPI_COMM_P2P) as kernels of their own (the
// two-exchange flow of MPPI_FLAG_SPLIT_KERNELS; the single-exchange merge runs in line);
// xa from make_xchg_args (peer_mb[r] = mailbox of rank r as mapped in this process,
// cudaIpcOpenMemHandle or a raw peer pointer; peer_mb[rank] = the local one)
cudaError_t launch_xchg_min(const LaunchCtx &c, CtlDev *ctl, const XchgArgs &xa);
cudaError_t launch_xchg_sum_finalize(const LaunchCtx &c, long long *acc, float *U, float *U_prev,
                                     const ProblemDev *prob, CtlDev *ctl, float *next_act,
                                     unsigned flags, const XchgArgs &xa);

// layout conversion between the reference's [K][T*A] and the internal K-minor [T*A][k_pad]
cudaError_t launch_to_internal(const LaunchCtx &c, const float *e_ref, float *eps);
cudaError_t launch_to_reference(const LaunchCtx &c, const float *eps, float *e_ref);

// debug taps (get_inf): normalised weights with the reference's mixed-precision formula,
// and the trajectories x[K][T+1][S] recomputed from eps and the pre-update U.
cudaError_t launch_norm_weights(const LaunchCtx &c, const float *S, float lambda, float beta,
                                float eta, float *w_out);
cudaError_t launch_trajectories(const LaunchCtx &c, const float *eps, const float *U_prev,
                                const ProblemDev *prob, float *x_out);

// (1-5) the whole control step in one persistent warp-specialised kernel (step.cu): fused
//       sampling + rollout warps, TMA producer and consumer warps of the weighted average on
//       every SM at once, last CTA merges and applies the U update.  tmap: box {128, 40}.
//       part: step_part_floats() floats of scratch (one record per CTA).
bool step_kernel_supported(int T, int A, long long k_pad, int num_sms);
size_t step_part_floats(const LaunchCtx &c);
cudaError_t configure_step(const LaunchCtx &c);
//       xa.world > 1 (K-shards): the last CTA also exchanges the merged sums (relative to the
//       shard's own minimum) with the peers over NVLink before the U update (xchg.cuh).
cudaError_t launch_step(const LaunchCtx &c, const CUtensorMap &tmap, float *eps, float *U,
                        const ProblemDev *prob, float *S, CtlDev *ctl, float *part,
                        float *U_prev, float *next_act, unsigned flags, const XchgArgs &xa);
XchgArgs make_xchg_args(unsigned long long *const *peer_mb, int rank, int world, int rows);
constexpr int kStepTileK = 128, kStepTileR = 40;   // its TMA box

// (1-5) the whole control step in one persistent kernel that keeps eps in SHARED MEMORY
//       (tile.cu): per SM a tile of 64 samples is drawn by the generator warps, integrated by
//       one packed-FP32x2 warp and folded into per-thread row sums; eps never reaches HBM.
//       Same record / merge / finalize protocol as launch_step (part: step_part_floats()).
bool tile_kernel_supported(int T, int A, long long k_pad, int num_sms);
cudaError_t configure_tile(const LaunchCtx &c);
//       xa.world > 1 (K-shards): the last CTA also runs the NVLink exchange (xchg.cuh).
cudaError_t launch_tile(const LaunchCtx &c, float *U, const ProblemDev *prob, float *S, CtlDev *ctl,
                        float *part, float *U_prev, float *next_act, unsigned flags,
                        const XchgArgs &xa);

// reset the control block (min key armed, step 0)
cudaError_t launch_clear_ctl(const LaunchCtx &c, CtlDev *ctl);

// opt-in of every kernel to the device's maximum of dynamic shared memory (per function and
// device, never lowered), and the check of what the shape in c needs against that maximum
cudaError_t configure_kernels(const LaunchCtx &c);
const char *check_smem_requirements(const LaunchCtx &c, size_t *need, size_t *have);

}  // namespace mppi
