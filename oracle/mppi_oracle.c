/*
 * mppi_oracle.c -- CPU restatement of the reference MPPI control step.
 *
 * TEST INFRASTRUCTURE ONLY (see mppi_oracle.h).  Not a fallback: the product
 * library never links this file.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off keeps every a*b+c as two rounded operations (the reference's
 * host build); fmaf() calls are the only fused operations and are explicit.
 *
 * Parity status: PINNED.  Bit-compared against the reference's own sources
 * compiled for the host (oracle/_ref, tests/test_oracle.py::test_oracle_matches_ref*)
 * and against the tests/golden npz fixtures produced by that build.
 */
#include "mppi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------
 * gains: src/point_mass.cu:46-51
 *   act[0] = _dt*_dt/2.0;  act[1] = _dt;  state = {1, _dt, 0, 1}
 * _dt is float, so _dt*_dt is a float product, widened to double for the
 * division by the double literal 2.0 and narrowed again on the store.
 * ------------------------------------------------------------------------- */
void oracle_gains(float dt, float state_gain[4], float act_gain[2])
{
    float dt2 = dt * dt;
    act_gain[0] = (float)((double)dt2 / 2.0);
    act_gain[1] = dt;
    state_gain[0] = 1.0f;
    state_gain[1] = dt;
    state_gain[2] = 0.0f;
    state_gain[3] = 1.0f;
}

/* Cost::step_cost, src/cost.cu:42-55.  x = state at t+1, u = U[t], e = eps[t]. */
static float step_cost_strict(const oracle_problem *p, const float *x, const float *u,
                              const float *e)
{
    float res = 0.0f;
    for (int i = 0; i < p->A; i++)
        res += u[i] * p->inv_s[i] * e[i];           /* (u*inv_s)*e, then += */
    res *= p->lambda;
    for (int i = 0; i < 2 * p->A; i++)
        res += (x[i] - p->goal[i]) * p->w[i] * (x[i] - p->goal[i]);
    return res;
}

/* same expressions as nvcc -fmad=true contracts them (checked in the PTX of the
 * reference's cost.cu: mul.f32 + fma.rn.f32 per term) */
static float step_cost_fma(const oracle_problem *p, const float *x, const float *u,
                           const float *e)
{
    float res = 0.0f;
    for (int i = 0; i < p->A; i++)
        res = fmaf(u[i] * p->inv_s[i], e[i], res);
    res = res * p->lambda;
    for (int i = 0; i < 2 * p->A; i++) {
        float d = x[i] - p->goal[i];
        res = fmaf(d * p->w[i], d, res);
    }
    return res;
}

/* Cost::final_cost, src/cost.cu:57-64.  The weights are those of the Cost object it is called
 * on: w (the reference's single object) or, with use_wf, the terminal weights wf. */
static float final_cost_strict(const oracle_problem *p, const float *x)
{
    const float *w = p->use_wf ? p->wf : p->w;
    float res = 0.0f;
    for (int i = 0; i < 2 * p->A; i++)
        res += (x[i] - p->goal[i]) * w[i] * (x[i] - p->goal[i]);
    return res;
}

static float final_cost_fma(const oracle_problem *p, const float *x)
{
    const float *w = p->use_wf ? p->wf : p->w;
    float res = 0.0f;
    for (int i = 0; i < 2 * p->A; i++) {
        float d = x[i] - p->goal[i];
        res = fmaf(d * w[i], d, res);
    }
    return res;
}

/* PointMassModelGpu::run + step, src/point_mass_gpu.cu:82-121 (host branch:
 * noise is the injected eps, nothing is sampled). */
float oracle_rollout(const oracle_problem *p, const float *x0, const float *U,
                     const float *eps, float *xtraj)
{
    const int A = p->A, S = 2 * p->A, T = p->T;
    float g[4], b[2];
    float xa[2 * ORACLE_MAX_ACT], xb[2 * ORACLE_MAX_ACT];
    float *x = xa, *xn = xb;
    float c = 0.0f;                                   /* run(): _c = 0 */

    if (p->use_gains) {
        for (int i = 0; i < 4; i++) g[i] = p->g[i];
        b[0] = p->b[0]; b[1] = p->b[1];
    } else {
        oracle_gains(p->dt, g, b);
    }
    for (int i = 0; i < S; i++) x[i] = x0[i];
    if (xtraj) memcpy(xtraj, x, sizeof(float) * S);

    for (int t = 0; t < T; t++) {
        const float *u = U + (size_t)t * A;
        const float *e = eps + (size_t)t * A;
        if (p->arith == ORACLE_ARITH_STRICT) {
            /* src/point_mass_gpu.cu:97-106, evaluation left to right */
            for (int i = 0; i < A; i++) {
                xn[i]     = g[0] * x[i] + g[1] * x[i + A] + b[0] * (u[i] + e[i]);
                xn[i + A] = g[2] * x[i] + g[3] * x[i + A] + b[1] * (u[i] + e[i]);
            }
            c += step_cost_strict(p, xn, u, e);       /* :107 */
        } else {
            /* contraction seen in the PTX of the reference's step():
             *   mul t=g1*v ; fma(g0,p,t) ; add ue=u+e ; fma(b0,ue,.) */
            for (int i = 0; i < A; i++) {
                float ue = u[i] + e[i];
                xn[i]     = fmaf(b[0], ue, fmaf(g[0], x[i], g[1] * x[i + A]));
                xn[i + A] = fmaf(b[1], ue, fmaf(g[2], x[i], g[3] * x[i + A]));
            }
            c = c + step_cost_fma(p, xn, u, e);
        }
        if (xtraj) memcpy(xtraj + (size_t)(t + 1) * S, xn, sizeof(float) * S);
        float *tmp = x; x = xn; xn = tmp;
    }
    /* terminal cost on x[T] -- charged a second time, :116 */
    if (p->arith == ORACLE_ARITH_STRICT) c += final_cost_strict(p, x);
    else                                 c = c + final_cost_fma(p, x);
    return c;
}

/* sim_gpu_kernel_, src/point_mass.cu:493-508: cost[k] = models[k].run() */
void oracle_rollout_all(const oracle_problem *p, const float *x0, const float *U,
                        const float *eps, float *S, float *xtraj, int nthreads)
{
    const size_t TA = (size_t)p->T * p->A;
    const size_t XS = (size_t)(p->T + 1) * 2 * p->A;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
#endif
    for (int64_t k = 0; k < p->K; k++)
        S[k] = oracle_rollout(p, x0, U, eps + (size_t)k * TA,
                              xtraj ? xtraj + (size_t)k * XS : NULL);
}

/* beta(): min over all costs, src/point_mass.cu:273-322 / min_red :533-575.
 * min is order independent; ties resolved to the lowest index. */
float oracle_beta(const float *S, int64_t K, int64_t *argmin)
{
    float m = INFINITY;
    int64_t idx = -1;
    for (int64_t k = 0; k < K; k++)
        if (S[k] < m) { m = S[k]; idx = k; }
    if (argmin) *argmin = idx;
    return m;
}

/* exp_red, src/point_mass.cu:518: expf(-(1/lambda[0]) * (cost[tid] - beta[0])) */
void oracle_exp(const float *S, int64_t K, float lambda, float beta, float *out)
{
    const float nil = -(1 / lambda);
    for (int64_t k = 0; k < K; k++)
        out[k] = expf(nil * (S[k] - beta));
}

float oracle_eta(const float *ex, int64_t K, double *eta_f64)
{
    float s = 0.0f;
    double d = 0.0;
    for (int64_t k = 0; k < K; k++) { s += ex[k]; d += (double)ex[k]; }
    if (eta_f64) *eta_f64 = d;
    return s;
}

/* weights_kernel, src/point_mass.cu:751:
 *   v_r[tid] = 1.0/nabla_1[0] * expf(-(1.0/lambda_1[0])*(v[tid] - beta[0]));
 * 1.0 is a double literal: 1.0/nabla and -(1.0/lambda)*(float diff) are formed in
 * double; the exponent argument narrows to float for expf; the product with
 * 1.0/nabla is double and narrows on the store. */
void oracle_weights(const float *S, int64_t K, float lambda, float beta, float eta,
                    float *w)
{
    const double inv_eta = 1.0 / (double)eta;
    const double nil = -(1.0 / (double)lambda);
    for (int64_t k = 0; k < K; k++) {
        float diff = S[k] - beta;
        float arg = (float)(nil * (double)diff);
        w[k] = (float)(inv_eta * (double)expf(arg));
    }
}

/* update_act_cpu, src/test.cu:97-105 (loop order and float accumulation kept) */
void oracle_update_act(float *u, const float *w, const float *e, int n, int t, int a)
{
    for (int k = 0; k < n; k++)
        for (int j = 0; j < t; j++)
            for (int i = 0; i < a; i++)
                u[j * a + i] += w[k] * e[(size_t)k * t * a + j * a + i];
}

/* update_act_cpu over disjoint column ranges in parallel: every u[j] still receives its K
 * contributions in the order k = 0..n-1, so the result is bit-identical to the serial loop. */
void oracle_update_act_mt(float *u, const float *w, const float *e, int n, int t, int a,
                          int nthreads)
{
    const int ta = t * a;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
    {
#ifdef _OPENMP
        const int nt = omp_get_num_threads(), id = omp_get_thread_num();
#else
        const int nt = 1, id = 0;
#endif
        const int j0 = (int)((long long)ta * id / nt), j1 = (int)((long long)ta * (id + 1) / nt);
        for (int k = 0; k < n; k++) {
            const float wk = w[k];
            const float *ek = e + (size_t)k * ta;
            for (int j = j0; j < j1; j++) u[j] += wk * ek[j];
        }
    }
}

void oracle_update_act_f64(float *u, const float *w, const float *e, int n, int t, int a)
{
    const int ta = t * a;
    double *acc = (double *)calloc((size_t)ta, sizeof(double));
    for (int k = 0; k < n; k++)
        for (int j = 0; j < ta; j++)
            acc[j] += (double)w[k] * (double)e[(size_t)k * ta + j];
    for (int j = 0; j < ta; j++) u[j] = (float)((double)u[j] + acc[j]);
    free(acc);
}

/* shift_act, src/point_mass.cu:805-824, followed by the D2D copy u <- u_swap
 * (:199): u'[t] = u[t+1] for t < T-1, u'[T-1] = u[T-1]. */
void oracle_shift(float *u, int T, int A)
{
    for (int t = 0; t < T - 1; t++)
        for (int j = 0; j < A; j++)
            u[t * A + j] = u[(t + 1) * A + j];
    /* last row keeps its value (repeat-last re-initialisation) */
}

/* PointMassModel::get_act, src/point_mass.cu:129-203 */
void oracle_step(const oracle_problem *p, const float *x0, float *U, const float *eps,
                 float *next_act, float *S_out, float *beta_out, float *eta_out,
                 float *weights_out, int64_t *argmin_out, int nthreads)
{
    const int64_t K = p->K;
    float *S  = S_out ? S_out : (float *)malloc(sizeof(float) * (size_t)K);
    float *ex = (float *)malloc(sizeof(float) * (size_t)K);
    float *w  = weights_out ? weights_out : (float *)malloc(sizeof(float) * (size_t)K);

    oracle_rollout_all(p, x0, U, eps, S, NULL, nthreads);          /* sim()     */
    float beta = oracle_beta(S, K, argmin_out);                     /* beta()    */
    oracle_exp(S, K, p->lambda, beta, ex);                          /* exp()     */
    float eta = oracle_eta(ex, K, NULL);                            /* nabla()   */
    oracle_weights(S, K, p->lambda, beta, eta, w);                  /* weights() */
    oracle_update_act(U, w, eps, (int)K, p->T, p->A);               /* update_act() */
    if (next_act)                                                   /* :195      */
        for (int j = 0; j < p->A; j++) next_act[j] = U[j];
    oracle_shift(U, p->T, p->A);                                    /* :198-199  */

    if (beta_out) *beta_out = beta;
    if (eta_out) *eta_out = eta;
    if (!S_out) free(S);
    if (!weights_out) free(w);
    free(ex);
}

/* ---------------------------------------------------------------------------
 * Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11).  Round: (hi0,lo0) = M0*c0, (hi1,lo1) = M1*c2,
 * c' = {hi1^c1^k0, lo1, hi0^c3^k1, lo0}; key bump k0 += W0, k1 += W1.
 * ------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    oracle_philox4x32(ctr, key, 10, out);
}

/* Philox-4x32 with `rounds` rounds (Random123's philox4x32_R<rounds>; 7 and 10 have known
 * answers in its kat_vectors) */
void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Box-Muller on a pair of 32-bit draws -> two N(0, sigma^2) values (sin first), the
 * same formulation as the device code (mppi_gpu_b200/csrc/philox.cuh):
 *   u     = xa*2^-32 + 2^-33           in (0,1]   (one fused op)
 *   theta = xb*2pi*2^-32 + 2pi*2^-33   in (0,2pi] (one fused op)
 *   r     = sqrt(|c * log2 u|),  c = -2 ln2 * sigma^2  (float ops)
 * libm log2f/sqrtf/sinf/cosf here, MUFU approximations on the device. */
static inline void box_muller(uint32_t xa, uint32_t xb, float c, float *n0, float *n1)
{
    float u  = fmaf((float)xa, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float th = fmaf((float)xb, 1.4629180792671596e-9f, 7.314590396335798e-10f);
    float r  = sqrtf(fabsf(c * log2f(u)));
    *n0 = r * sinf(th);
    *n1 = r * cosf(th);
}

void oracle_sample_eps(uint64_t seed, uint64_t step, int64_t k0, int64_t K, int T, int A,
                       const float *sigma, float *eps)
{
    oracle_sample_eps_rounds(seed, step, k0, K, T, A, sigma, 10, eps);
}

void oracle_sample_eps_rounds(uint64_t seed, uint64_t step, int64_t k0, int64_t K, int T, int A,
                              const float *sigma, int rounds, float *eps)
{
    const int R = T * A;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    float c[ORACLE_MAX_ACT];
    for (int a = 0; a < A; a++) c[a] = -1.3862943611198906f * (sigma[a] * sigma[a]);
    for (int64_t k = 0; k < K; k++) {
        const int64_t kg = k0 + k;
        const uint32_t q = (uint32_t)(kg >> 2);
        const int lane = (int)(kg & 3);
        for (int r = 0; r < R; r++) {
            uint32_t ctr[4] = { q, (uint32_t)r, (uint32_t)step, (uint32_t)(step >> 32) };
            uint32_t x[4];
            float n[4];
            oracle_philox4x32(ctr, key, rounds, x);
            box_muller(x[0], x[1], c[r % A], &n[0], &n[1]);
            box_muller(x[2], x[3], c[r % A], &n[2], &n[3]);
            eps[(size_t)k * R + r] = n[lane];
        }
    }
}
