/*
 * mppi_oracle.h -- CPU restatement of the reference MPPI control step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mppi_gpu_b200/,
 * include/, cpp/) may include, link or call this.  Allowed users: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the NicolayP/mppi_gpu tree).  Parity is pinned by
 *   - tests/golden/ (vectors produced by running the reference's own sources,
 *     see tests/golden/make_golden.py) and
 *   - oracle/_ref (the reference's point_mass_gpu.cu + cost.cu compiled for the
 *     host, see oracle/Makefile), compared bit-for-bit in tests/test_oracle.py.
 */
#ifndef MPPI_ORACLE_H_
#define MPPI_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_ACT 8

/* Arithmetic flavour of the per-sample rollout.
 *  STRICT : every multiply and add rounded separately, evaluation order as
 *           written in the reference source == the reference's host (CPU)
 *           build (x86-64 g++, no contraction).
 *  FMA    : the contraction nvcc applies to the same expressions with its
 *           default -fmad=true, written out with explicit fmaf() -- the
 *           arithmetic of the reference's *device* build. */
enum { ORACLE_ARITH_STRICT = 0, ORACLE_ARITH_FMA = 1 };

typedef struct oracle_problem {
    int32_t K;                       /* samples                      */
    int32_t T;                       /* horizon (steps)              */
    int32_t A;                       /* action dim, state dim = 2A   */
    int32_t arith;                   /* ORACLE_ARITH_*               */
    float   dt;
    float   lambda;                  /* reference hard-codes 1       */
    float   inv_s[ORACLE_MAX_ACT];   /* reference hard-codes 1       */
    float   goal[2 * ORACLE_MAX_ACT];
    float   w[2 * ORACLE_MAX_ACT];
    int32_t use_gains;               /* 0: gains from dt (reference constructor);
                                        1: g / b below (the arguments of
                                        PointMassModelGpu::init, src/point_mass_gpu.cu:25-39) */
    float   g[4];
    float   b[2];
    int32_t use_wf;                  /* 0: the terminal cost uses w (the reference: one Cost
                                        object for stage and final cost, src/point_mass_gpu.cu:107,116);
                                        1: a second Cost object with weights wf charges the final
                                        state (Cost::final_cost, src/cost.cu:57-64, unchanged) */
    float   wf[2 * ORACLE_MAX_ACT];
} oracle_problem;

/* gains of the double integrator: src/point_mass.cu:46-51 */
void  oracle_gains(float dt, float state_gain[4], float act_gain[2]);

/* one sample: PointMassModelGpu::run/step (src/point_mass_gpu.cu:82-121) with
 * Cost::step_cost/final_cost (src/cost.cu:42-64).  eps is [T*A] (one sample's
 * slice of the reference [K,T,A] layout); xtraj is [(T+1)*2A] or NULL. */
float oracle_rollout(const oracle_problem *p, const float *x0, const float *U,
                     const float *eps, float *xtraj);

/* all K samples, serial loop (== sim_gpu_kernel_, src/point_mass.cu:493-508).
 * eps [K,T,A]; S [K]; xtraj [K,(T+1),2A] or NULL. nthreads<=1 -> serial. */
void  oracle_rollout_all(const oracle_problem *p, const float *x0, const float *U,
                         const float *eps, float *S, float *xtraj, int nthreads);

/* beta = min_k S_k (src/point_mass.cu:273-322,533-575); argmin = lowest index */
float oracle_beta(const float *S, int64_t K, int64_t *argmin);

/* exp_red: out_k = expf(-(1/lambda)*(S_k-beta)), all float (src/point_mass.cu:518) */
void  oracle_exp(const float *S, int64_t K, float lambda, float beta, float *out);

/* eta = sum_k exp_k.  The reference's tree order (src/point_mass.cu:628-666) is
 * a property of its launch shape; the serial float sum and a double sum bracket
 * it.  Returns the serial float sum, *eta_f64 (if not NULL) the double one. */
float oracle_eta(const float *ex, int64_t K, double *eta_f64);

/* weights_kernel with its double literals (src/point_mass.cu:751) */
void  oracle_weights(const float *S, int64_t K, float lambda, float beta, float eta,
                     float *w);

/* update_act_cpu, verbatim loop order k,t,a (src/test.cu:97-105) */
void  oracle_update_act(float *u, const float *w, const float *e, int n, int t, int a);
/* the same loop split over column ranges (bit-identical result, for the threaded baseline) */
void  oracle_update_act_mt(float *u, const float *w, const float *e, int n, int t, int a,
                           int nthreads);
/* same sums accumulated in double then added to u (tolerance anchor) */
void  oracle_update_act_f64(float *u, const float *w, const float *e, int n, int t, int a);

/* shift_act + device-to-device copy back (src/point_mass.cu:198-199,805-824) */
void  oracle_shift(float *u, int T, int A);

/* one full control step == PointMassModel::get_act (src/point_mass.cu:129-203).
 * U is updated in place (post-shift, as get_u would return it afterwards).
 * Outputs may be NULL. */
void  oracle_step(const oracle_problem *p, const float *x0, float *U,
                  const float *eps, float *next_act, float *S, float *beta,
                  float *eta, float *weights, int64_t *argmin, int nthreads);

/* ---- sampling (new in the B200 build; the reference uses cuRAND XORWOW) ----
 * Philox-4x32-10 (Salmon et al., SC'11; Random123 v1.14 known-answer vectors;
 * same round/key constants as cuRAND's curand_philox4x32_x.h, CUDA 12.9). */
void  oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the same with 7 or 10 rounds (mppi_params.philox_rounds) */
void  oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);

/* The controller's noise stream: for quad q = k/4 and row r = t*A+a,
 *   (e0..e3) = BoxMuller(Philox(ctr = {q, r, step_lo, step_hi}, key = seed); sigma[a])
 *   eps[k=4q+j, t, a] = e_j
 * Writes the reference layout [K,T,A] for global samples k0 .. k0+K-1. */
void  oracle_sample_eps(uint64_t seed, uint64_t step, int64_t k0, int64_t K, int T, int A,
                        const float *sigma, float *eps);
void  oracle_sample_eps_rounds(uint64_t seed, uint64_t step, int64_t k0, int64_t K, int T, int A,
                               const float *sigma, int rounds, float *eps);

#ifdef __cplusplus
}
#endif
#endif
