/*
 * mppi_b200.h -- C ABI of the B200-native MPPI controller core.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * NicolayP/mppi_gpu: one MPPI control step, PointMassModel::get_act
 * (reference src/point_mass.cu:129-203).  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree).  The C++ shim
 * include/mppi_b200/point_mass.hpp re-exposes the reference's class
 * `PointMassModel` (include/point_mass.hpp:23-44) on top of these calls.
 *
 * Conventions
 *   - every pointer argument is caller-owned HOST memory unless stated; the
 *     callee copies synchronously and retains nothing (reference behaviour,
 *     src/point_mass.cu:205-262);
 *   - every function returns 0 on success or a negative MPPI_ERR_* code;
 *     mppi_last_error() gives the message of the calling thread's last failure.
 *     (The reference's convention -- print "API error failed file:line" and
 *     exit(1), include/mppi_utils.hpp:19-25 -- is restored by the C++ shim.)
 *   - a handle is not thread-safe (neither is the reference object);
 *   - there is no CPU fallback: without a CUDA device mppi_create fails.
 *
 * Layouts (all float32): x0/goal/w [S], U [T][A], noise/eps [K][T][A],
 * trajectories [K][T+1][S], S = 2A (positions then velocities).
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_ABI_VERSION 3
#define MPPI_MAX_ACT 4            /* supported action dims: 1..4 (state dim = 2A) */
#define MPPI_COMM_ID_BYTES 128

/* status codes */
#define MPPI_OK                 0
#define MPPI_ERR_INVALID       -1   /* bad argument / unsupported shape          */
#define MPPI_ERR_CUDA          -2   /* a CUDA runtime/driver call failed         */
#define MPPI_ERR_NO_DEVICE     -3   /* no usable sm_100 device                   */
#define MPPI_ERR_COMM          -4   /* NCCL / peer-memory exchange failed        */
#define MPPI_ERR_STATE         -5   /* call sequence error (e.g. step before set_problem) */

/* mppi_params.flags */
#define MPPI_FLAG_STRICT_ARITH   (1u << 0)  /* rollout arithmetic without FMA contraction ==
                                               the reference's host build; default is the
                                               contraction of its device build (nvcc -fmad) */
#define MPPI_FLAG_INJECTED_NOISE (1u << 1)  /* step consumes the noise set by mppi_set_noise
                                               instead of sampling (validation mode)        */
#define MPPI_FLAG_CLAMP_ACTIONS  (1u << 2)  /* clamp the updated U to +-max_act (config key
                                               `max-a`; the reference parses and ignores it) */
#define MPPI_FLAG_REINIT_INIT_ACT (1u << 3) /* U[T-1] := init_act after the shift (config key
                                               `init-act`); default repeats the last row as
                                               the reference does, src/point_mass.cu:815-821 */
#define MPPI_FLAG_NO_GRAPH       (1u << 4)  /* launch the kernel chain directly instead of
                                               replaying the captured CUDA graph            */
#define MPPI_FLAG_FUSED_SAMPLING (1u << 5)  /* sample eps inside the rollout kernel (one pass
                                               writes eps and integrates)                   */
#define MPPI_FLAG_SPLIT_KERNELS  (1u << 6)  /* run weights (3) and finalize (5) as kernels of
                                               their own instead of inside the averaging
                                               kernel (4): per-part profiling                */

#define MPPI_FLAG_STEP_KERNEL     (1u << 7)  /* the whole control step as ONE persistent kernel:
                                               sampling + rollout warps and the TMA-fed
                                               weighted-average warps share every SM, so the
                                               arithmetic-bound and the HBM-bound halves overlap.
                                               Sampled noise, single shard, T*A small enough for
                                               the shared-memory row sums; otherwise the step
                                               silently uses the kernel chain                */

#define MPPI_FLAG_AUTO_CHAIN      (1u << 8)  /* let the library pick the kernel chain from an
                                               estimate of what each costs on this shard (a cost
                                               model fitted on B200 over 112 shapes: the rollouts
                                               are bound by the FMA pipe of the fullest SM
                                               sub-partition, the eps passes by HBM): the
                                               one-kernel step (single shard or MPPI_COMM_P2P),
                                               MPPI_FLAG_FUSED_SAMPLING, or the unfused chain
                                               with MPPI_FLAG_PIPELINED_SAMPLING.  At T=200, A=3:
                                               unfused up to ~1.1e5 samples per GPU, fused up to
                                               ~3.5e5, the one-kernel step above.
                                               mppi_get_flags returns the choice.             */

#define MPPI_FLAG_PIPELINED_SAMPLING (1u << 9) /* unfused chain only, a LATENCY option for closed
                                               loops: as soon as step n has published its action
                                               the noise of step n+1 is drawn (second stream,
                                               second eps buffer) while the caller's plant has
                                               its turn, so that mppi_step(n+1) is rollout +
                                               average only.  Philox is counter based: eps
                                               depends on (seed, step, k, t, a), not on U or x.
                                               Same noise, same results, bit for bit; throughput
                                               of back-to-back steps is unchanged.  Costs a
                                               second eps buffer; ignored with fused sampling,
                                               the step kernel, injected noise,
                                               MPPI_FLAG_NO_GRAPH and profiling (those steps run
                                               the plain chain)                               */

#define MPPI_FLAG_TILE_KERNEL     (1u << 10) /* the whole control step as ONE persistent kernel
                                               that keeps eps ON CHIP: per SM a tile of 64 samples
                                               is drawn into shared memory, integrated and folded
                                               into the weighted sums there; eps never reaches
                                               HBM (mppi_get_info re-draws it on demand: Philox
                                               is counter based).  Sampled noise, single shard
                                               or MPPI_COMM_P2P, T*A*(272+16) bytes must fit the
                                               227 KB of shared memory (T*A <= ~780); otherwise
                                               the step silently uses the kernel chain.
                                               Takes precedence over MPPI_FLAG_STEP_KERNEL.   */

/* mppi_params.comm */
#define MPPI_COMM_NONE  0   /* single shard                                          */
#define MPPI_COMM_NCCL  1   /* ncclAllReduce(min) for beta, ncclAllReduce(sum) for the
                               weighted-noise partials and eta                       */
#define MPPI_COMM_P2P   2   /* the same two exchanges as direct stores into peer mailboxes
                               over NVLink (CUDA IPC), fused with the U update: no NCCL
                               kernel on the step's critical path                    */
#define MPPI_P2P_HANDLE_BYTES 64

/* mppi_params.model: the dynamics functor the rollout kernels are instantiated on (the cost
 * is Cost::step_cost / final_cost for both) */
#define MPPI_MODEL_POINT_MASS   0   /* the reference's double integrator: gains {1,dt,0,1} /
                                       {dt^2/2,dt} formed from dt as in src/point_mass.cu:46-51 */
#define MPPI_MODEL_LINEAR_AXIS  1   /* per action dim  p' = g0 p + g1 v + b0 (u+e),
                                       v' = g2 p + g3 v + b1 (u+e)  with caller-given gains
                                       (the arguments of PointMassModelGpu::init,
                                       src/point_mass_gpu.cu:25-39): damped / geared point masses */

typedef struct mppi_handle mppi_handle;

/* Replaces the constructor arguments of PointMassModel (include/point_mass.hpp:25-30:
 * nb_sim, steps, dt, state_dim, act_dim, verbose) plus the quantities the reference
 * hard-codes: lambda = 1 (src/point_mass.cu:53-54), sigma = 0.025
 * (src/point_mass_gpu.cu:86), Sigma^-1 = 1 (src/point_mass_gpu.cu:58-61). */
typedef struct mppi_params {
    uint32_t struct_size;               /* sizeof(mppi_params), ABI check             */
    uint32_t flags;                     /* MPPI_FLAG_*                                */
    int64_t  samples;                   /* K, GLOBAL number of rollouts               */
    int32_t  horizon;                   /* T                                          */
    int32_t  state_dim;                 /* must equal 2*act_dim                       */
    int32_t  act_dim;                   /* A                                          */
    int32_t  verbose;
    float    dt;
    float    lambda;                    /* temperature                                */
    float    sigma[MPPI_MAX_ACT];       /* sampling std-dev per action dim            */
    float    inv_sigma[MPPI_MAX_ACT];   /* Sigma^-1 diagonal in the control cost      */
    float    init_act[MPPI_MAX_ACT];    /* used with MPPI_FLAG_REINIT_INIT_ACT        */
    float    max_act[MPPI_MAX_ACT];     /* used with MPPI_FLAG_CLAMP_ACTIONS          */
    uint64_t seed;                      /* Philox key                                 */
    int32_t  device;                    /* CUDA device ordinal                        */
    int32_t  rank;                      /* this shard, 0..world_size-1                */
    int32_t  world_size;                /* K is split into world_size contiguous shards */
    int32_t  comm;                      /* MPPI_COMM_*                                */
    uint8_t  comm_id[MPPI_COMM_ID_BYTES]; /* ncclUniqueId from mppi_comm_unique_id     */
    int32_t  model;                     /* MPPI_MODEL_*                               */
    float    state_gain[4];             /* {g0,g1,g2,g3}, MPPI_MODEL_LINEAR_AXIS only */
    float    act_gain[2];               /* {b0,b1},       MPPI_MODEL_LINEAR_AXIS only */
    int32_t  philox_rounds;             /* 0 or 10: Philox-4x32-10 (default); 7: Philox-4x32-7,
                                           Random123's crush-resistant minimum -- 30 % fewer of
                                           the multiplies that bound the sampling kernels (-13 % on
                                           the fused rollout).  The kernel chains only: with 7 the
                                           one-kernel steps are not used (MPPI_FLAG_STEP_KERNEL /
                                           MPPI_FLAG_TILE_KERNEL fall back to the fused chain)    */
} mppi_params;

/* per-step scalars, the reference's _beta / _nabla (src/point_mass.cu:250-257) */
typedef struct mppi_step_info {
    float    beta;          /* min_k S_k over ALL shards                              */
    float    eta;           /* sum_k exp(-(S_k-beta)/lambda) over all shards          */
    int64_t  argmin;        /* global index of the minimum (lowest index on ties)     */
    uint64_t step;          /* control steps completed so far                         */
} mppi_step_info;

/* kernel ids for mppi_get_kernel_times */
enum {
    MPPI_K_SAMPLE = 0,   /* Philox + Box-Muller, writes eps                           */
    MPPI_K_ROLLOUT,      /* dynamics + cost over T, block/atomic min                  */
    MPPI_K_COMM_MIN,     /* all-reduce(min) of beta (multi-shard only)                */
    MPPI_K_WEIGHTS,      /* w~ = exp(-(S-beta)/lambda), eta partials                  */
    MPPI_K_AVERAGE,      /* sum_k w~_k eps_k[t,a]  (TMA-staged, HBM-bound)            */
    MPPI_K_COMM_SUM,     /* all-reduce(sum) of the T*A+1 partials (multi-shard only)  */
    MPPI_K_FINALIZE,     /* fold partials, U update, shift, re-init, next action      */
    MPPI_K_COUNT
};

/* Fill *p with the reference-compatible preset: lambda 1, sigma 0.025,
 * inv_sigma 1, repeat-last re-init, no clamp, seed 0, device 0, one shard. */
int mppi_params_default(mppi_params *p);

/* == new PointMassModel(nb_sim, steps, dt, state_dim, act_dim, verbose)
 * (src/point_mass.cu:19-106): allocates all device state on params->device,
 * builds the CUDA graph of the control step. */
int mppi_create(const mppi_params *params, mppi_handle **out);

/* One controller over several GPUs of one process (the process model of a caller such as the
 * reference's main.cu, which is a single process): K is sharded over devices[0..n-1], the
 * shards exchange through peer memory (NVLink), every other call of this API works on the
 * returned handle unchanged -- per-sample arrays of mppi_get_info / mppi_set_noise are the
 * full [K] arrays.  params->device/rank/world_size/comm are ignored. */
int mppi_create_multi(const mppi_params *params, const int *devices, int num_devices,
                      mppi_handle **out);

/* == delete model (src/point_mass.cu:108-127) */
int mppi_destroy(mppi_handle *h);

/* == memcpy_set_data(x, u, goal, w) (src/point_mass.cu:205-228): x0 [S], U [T*A],
 * goal [S], cost weights w [S]. */
int mppi_set_problem(mppi_handle *h, const float *x0, const float *u, const float *goal,
                     const float *w);

/* Terminal weights: the final state x[T] is charged by a SECOND Cost object with weights
 * w_final[state_dim] (Cost::final_cost, src/cost.cu:57-64, unchanged).  The reference uses one
 * object for the stage and the final cost (src/point_mass_gpu.cu:107,116) although its Cost class
 * takes the weights as an argument (include/cost.hpp:8-14); that remains the default, and NULL
 * returns to it.  Sticky across mppi_set_problem.  No reference counterpart. */
int mppi_set_terminal_weights(mppi_handle *h, const float *w_final);

/* == set_x(x) (src/point_mass.cu:482-486): new initial state for the next step */
int mppi_set_state(mppi_handle *h, const float *x);

/* == get_act(next_act) (src/point_mass.cu:129-203): one control step; writes the
 * A floats of U[0,:] *before* the shift; U is left already shifted. Blocking: the
 * finalizing kernel stores the action and the step counter into pinned host memory
 * mapped into the device as soon as the new U[0,:] is known, and the call returns as
 * soon as that counter shows up; the shift itself (and, with
 * MPPI_FLAG_PIPELINED_SAMPLING, the next step's noise) completes behind it on the
 * handle's streams, which every other call on the handle is ordered after. */
int mppi_step(mppi_handle *h, float *next_act);

/* The two halves of mppi_step for callers that overlap host work: enqueue the
 * step on the handle's stream, then wait for it and fetch next_act. */
int mppi_step_enqueue(mppi_handle *h);
int mppi_step_wait(mppi_handle *h, float *next_act);

/* == get_u(u) (src/point_mass.cu:488-491): current U [T*A] */
int mppi_get_u(mppi_handle *h, float *u);

/* Overwrite U only (the reference does this through memcpy_set_data). */
int mppi_set_u(mppi_handle *h, const float *u);

/* == get_inf(x, u, e, cost, beta, nabla, weight) (src/point_mass.cu:236-262), the
 * parity tap.  Any pointer may be NULL.  Sizes: x [K_local*(T+1)*S] (recomputed by a
 * debug kernel from the last step's eps and pre-update U), u [T*A] (post-shift),
 * e [K_local*T*A] in the reference layout, cost [K_local], beta [1], nabla [1],
 * weight [K_local].  K_local = this shard's samples (== K for one shard). */
int mppi_get_info(mppi_handle *h, float *x, float *u, float *e, float *cost, float *beta,
                  float *nabla, float *weight);

/* the flags the handle runs with (MPPI_FLAG_AUTO_CHAIN resolved to the chosen chain) */
int mppi_get_flags(mppi_handle *h, uint32_t *flags);

/* beta / eta / argmin / step counter of the last step */
int mppi_get_step_info(mppi_handle *h, mppi_step_info *info);

/* NEW (no reference counterpart; the reference can only sample): inject the noise
 * of the next steps, e [K_local*T*A] in the reference layout [K][T][A]; implies
 * MPPI_FLAG_INJECTED_NOISE until mppi_set_noise_mode(h, 0). */
int mppi_set_noise(mppi_handle *h, const float *e);
int mppi_set_noise_mode(mppi_handle *h, int injected);

/* Run only the sampling kernel for control step `step` (fills eps; for tests). */
int mppi_sample_only(mppi_handle *h, uint64_t step);

/* shard geometry: rank r of world owns the global samples [*k_begin, *k_end); shards
 * are whole Philox quads (4 samples) so eps[k] never depends on the shard count.
 * Pure host arithmetic (no device needed). */
int mppi_shard_range(int64_t samples, int rank, int world_size, int64_t *k_begin,
                     int64_t *k_end);
int mppi_local_samples(mppi_handle *h, int64_t *k_local, int64_t *k_offset);

/* What MPPI_FLAG_AUTO_CHAIN would decide for a shard of samples_local rollouts on a GPU with
 * num_sms SMs, and why: the estimated microseconds per control step of the unfused chain, the fused
 * chain and the one-kernel step (est_us[0..2]; constants fitted on B200) and the flag of the chosen
 * chain (MPPI_FLAG_PIPELINED_SAMPLING = unfused).  Pure host arithmetic (no device needed). */
int mppi_chain_estimate(int64_t samples_local, int horizon, int act_dim, int num_sms,
                        double est_us[3], uint32_t *choice);

/* Timing helpers on the handle's stream (CUDA events). */
int mppi_timer_start(mppi_handle *h);
int mppi_timer_stop(mppi_handle *h, float *elapsed_ms);
/* When enabled, every step is launched kernel by kernel with CUDA events between
 * the kernels; times accumulate until read.  ms_sum [MPPI_K_COUNT], launches
 * [MPPI_K_COUNT] (either may be NULL); reading resets the accumulators. */
int mppi_set_profiling(mppi_handle *h, int enabled);
int mppi_get_kernel_times(mppi_handle *h, double *ms_sum, int64_t *launches);
/* MPPI_COMM_P2P: the phases of the LAST step's NVLink exchange on this rank, microseconds from
 * %globaltimer stamps taken inside the exchanging CTA: us[0] push of this shard's {key,
 * accumulators} into every peer's mailbox + flags, us[1] wait until every peer's flag has
 * arrived (= the skew to the slowest rank), us[2] rescale + sum.  Zeros without an exchange. */
int mppi_get_exchange_times(mppi_handle *h, double us[3]);
/* total kernels this handle has launched so far (graph replays included) */
int mppi_get_launch_count(mppi_handle *h, int64_t *launches);
const char *mppi_kernel_name(int kernel_id);

/* multi-shard setup, MPPI_COMM_NCCL: rank 0 creates the id, every rank passes it in
 * mppi_params */
int mppi_comm_unique_id(uint8_t id[MPPI_COMM_ID_BYTES]);
/* multi-shard setup, MPPI_COMM_P2P (one process per GPU of one NVLink domain): after
 * mppi_create every rank exports the IPC handle of its mailbox, the host all-gathers the
 * world_size x 64 bytes (rank order) and every rank connects. */
int mppi_comm_p2p_handle(mppi_handle *h, uint8_t out[MPPI_P2P_HANDLE_BYTES]);
int mppi_comm_p2p_connect(mppi_handle *h, const uint8_t *handles);

const char *mppi_last_error(void);
int mppi_abi_version(void);

/* Environment variables read at mppi_create -- kernel-variant overrides for tests and
 * tuning, never needed in production (every variant computes the same bits):
 *   MPPI_ROLLOUT_TMA=0|1      force the register-pipelined / the TMA-staged rollout kernel
 *   MPPI_ROLLOUT_TMA_W=64|128|256   slab width of the TMA-staged rollout
 *   MPPI_ROLLOUT_SPT=1|2|4    samples per thread of the register-pipelined rollout
 *   MPPI_STEP_STAGES=4|8      TMA ring depth of the one-kernel step                      */

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */
