/*
 * ref_harness.cu -- drives the REFERENCE's own per-sample code on the host.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ together
 * with the reference sources *where they lie* ($(REF)/src/point_mass_gpu.cu and
 * $(REF)/src/cost.cu, compiled by nvcc for the host; every method there is
 * __host__ __device__ and the host branch of step() consumes pre-loaded noise,
 * src/point_mass_gpu.cu:92-95).  Nothing is copied out of the reference tree;
 * this file only contains the loop that the reference runs on the device as
 * sim_gpu_kernel_ (src/point_mass.cu:493-508) and set_data (:763-795).
 *
 * The reference's init() allocates four private arrays per sample with malloc()
 * and never frees them (src/point_mass_gpu.cu:52-58).  The shared object is
 * linked with -Wl,--wrap=malloc so that, while a sample is being initialised,
 * those allocations land in a per-thread bump arena that this harness rewinds
 * after every sample (the reference sources themselves are untouched).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "point_mass_gpu.hpp"

namespace {
struct Arena {
    char  *base = nullptr;
    size_t cap = 0, off = 0;
    bool   active = false;
};
thread_local Arena g_arena;
}  // namespace

extern "C" void *__real_malloc(size_t n);

extern "C" void *__wrap_malloc(size_t n)
{
    Arena &a = g_arena;
    if (!a.active) return __real_malloc(n);
    n = (n + 63) & ~(size_t)63;          /* pad: step() touches _e[t*A+1] for A=1 */
    if (a.off + n > a.cap) {
        fprintf(stderr, "ref_harness: arena exhausted (%zu + %zu > %zu)\n", a.off, n, a.cap);
        abort();
    }
    void *p = a.base + a.off;
    a.off += n;
    return p;
}

static void arena_reserve(size_t bytes)
{
    Arena &a = g_arena;
    if (a.cap < bytes) {
        free(a.base);
        a.base = static_cast<char *>(__real_malloc(bytes));
        a.cap = bytes;
    }
    a.off = 0;
}

/* One sample through the reference's PointMassModelGpu::init + run. */
static float ref_one(int T, int A, float *state_gain, float *act_gain, float *x0, float *U,
                     float *goal, float *w, float lambda, float *eps_k, float *xbuf, int id)
{
    const int S = 2 * A;
    /* The reference builds its objects in cudaMalloc'd memory without running
     * a constructor; init() then calls set_x() before _x_size is assigned
     * (src/point_mass_gpu.cu:40-46).  Zeroed storage makes that first set_x a
     * no-op; x0 is written once _x_size is valid, which is what every control
     * step after the first does through set_x_kernel (src/point_mass.cu:797-803). */
    alignas(16) unsigned char storage[sizeof(PointMassModelGpu)];
    memset(storage, 0, sizeof storage);
    PointMassModelGpu *m = reinterpret_cast<PointMassModelGpu *>(storage);
    g_arena.off = 0;
    g_arena.active = true;
    m->init(xbuf, x0, U, eps_k, T, state_gain, S, act_gain, A, w, goal, lambda, id, false);
    g_arena.active = false;
    m->set_x(x0);
    return m->run(nullptr);
}

extern "C" {

/* K rollouts of the reference model.  eps is the reference layout [K,T,A] and is
 * left unchanged (run() copies it back onto itself through save_e()).
 * S_out [K]; xtraj [K,(T+1),2A] or NULL.  Returns 0. */
static int ref_rollout_costs_impl(int K, int T, int A, float *state, float *act, float lambda,
                                  const float *x0_in, const float *U_in, const float *goal_in,
                                  const float *w_in, float *eps, float *S_out, float *xtraj,
                                  int nthreads);

int ref_rollout_costs(int K, int T, int A, float dt, float lambda, const float *x0_in,
                      const float *U_in, const float *goal_in, const float *w_in,
                      float *eps, float *S_out, float *xtraj, int nthreads)
{
    /* gains exactly as the reference constructor forms them, src/point_mass.cu:46-51 */
    float state[4];
    float act[2];
    float _dt = dt;
    act[0] = _dt * _dt / 2.0;
    act[1] = _dt;
    state[0] = 1;
    state[1] = _dt;
    state[2] = 0;
    state[3] = 1;
    return ref_rollout_costs_impl(K, T, A, state, act, lambda, x0_in, U_in, goal_in, w_in, eps,
                                  S_out, xtraj, nthreads);
}

/* the same with caller-given gains: PointMassModelGpu::init takes state_gain / act_gain as
 * arguments (src/point_mass_gpu.cu:25-39); only the reference's constructor hard-codes them */
int ref_rollout_costs_gains(int K, int T, int A, const float *state_gain, const float *act_gain,
                            float lambda, const float *x0_in, const float *U_in,
                            const float *goal_in, const float *w_in, float *eps, float *S_out,
                            float *xtraj, int nthreads)
{
    float state[4] = {state_gain[0], state_gain[1], state_gain[2], state_gain[3]};
    float act[2] = {act_gain[0], act_gain[1]};
    return ref_rollout_costs_impl(K, T, A, state, act, lambda, x0_in, U_in, goal_in, w_in, eps,
                                  S_out, xtraj, nthreads);
}

static int ref_rollout_costs_impl(int K, int T, int A, float *state, float *act, float lambda,
                                  const float *x0_in, const float *U_in, const float *goal_in,
                                  const float *w_in, float *eps, float *S_out, float *xtraj,
                                  int nthreads)
{
    const int S = 2 * A;

    std::vector<float> x0(x0_in, x0_in + S), U(U_in, U_in + (size_t)T * A);
    std::vector<float> goal(goal_in, goal_in + S), w(w_in, w_in + S);
    const size_t XS = (size_t)(T + 1) * S;
    const size_t arena_bytes = 4 * 64 + sizeof(float) * ((size_t)T * A + 2 * S + A) + 4096;
    if (nthreads < 1) nthreads = 1;

#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        arena_reserve(arena_bytes);
        std::vector<float> xloc(XS);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int k = 0; k < K; k++) {
            float *xb = xtraj ? xtraj + (size_t)k * XS : xloc.data();
            memset(xb, 0, sizeof(float) * XS);   /* cudaMemset(_x, 0), src/point_mass.cu:65 */
            S_out[k] = ref_one(T, A, state, act, x0.data(), U.data(), goal.data(), w.data(),
                               lambda, eps + (size_t)k * T * A, xb, k);
        }
    }
    return 0;
}

int ref_sizeof_model(void) { return (int)sizeof(PointMassModelGpu); }

/* The reference's Cost object on its own (include/cost.hpp, src/cost.cu:42-64): step_cost and
 * final_cost of n independent (x [2A], u [A], e [A]) triples for the weights w.  Pins the
 * oracle's terminal-weight mode: a rollout charged with Cost(w).step_cost per step and
 * Cost(w_final).final_cost on the last state is recomposed from these two calls. */
int ref_cost_terms(int n, int A, const float *w_in, const float *goal_in, float lambda,
                   const float *inv_s_in, const float *x, const float *u, const float *e,
                   float *step_out, float *final_out)
{
    const int S = 2 * A;
    std::vector<float> w(w_in, w_in + S), goal(goal_in, goal_in + S), inv_s(inv_s_in, inv_s_in + A);
    Cost c(w.data(), S, goal.data(), S, lambda, inv_s.data(), A);
    for (int i = 0; i < n; i++) {
        std::vector<float> xi(x + (size_t)i * S, x + (size_t)(i + 1) * S);
        std::vector<float> ui(u + (size_t)i * A, u + (size_t)(i + 1) * A);
        std::vector<float> ei(e + (size_t)i * A, e + (size_t)(i + 1) * A);
        if (step_out) step_out[i] = c.step_cost(xi.data(), ui.data(), ei.data(), 0, 0);
        if (final_out) final_out[i] = c.final_cost(xi.data(), 0);
    }
    return 0;
}

}  /* extern "C" */
