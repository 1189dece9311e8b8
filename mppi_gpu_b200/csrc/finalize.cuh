// finalize.cuh -- the U update + receding-horizon shift executed by ONE CTA (part 5).
#pragma once

#include "common.cuh"

namespace mppi {

// U update + receding-horizon shift from the fixed-point accumulators, executed by ONE CTA
// (the finalize kernel, or the last CTA of the merged average kernel).  s_u: T*A floats.
__device__ __forceinline__ void finalize_body(long long *acc, float *__restrict__ U,
                                              float *__restrict__ U_prev,
                                              const ProblemDev *__restrict__ prob, CtlDev *ctl,
                                              float *__restrict__ next_act, int T, int A,
                                              unsigned flags, float *s_u, int nt = 0, int bar_id = 0)
{
    // nt threads (threadIdx.x < nt) take part, synchronised on named barrier bar_id;
    // default: the whole CTA on barrier 0 (== __syncthreads)
    const int R = T * A;
    if (nt == 0) nt = blockDim.x;
    const volatile long long *vacc = acc;       // written by other CTAs' atomics: read at L2
    const float eta = acc_to_float(vacc[R]);
    for (int i = threadIdx.x; i < R; i += nt) {
        const float u = U[i];
        float un = u + acc_to_float(vacc[i]) / eta;
        if (flags & MPPI_FLAG_CLAMP_ACTIONS) {
            const float m = prob->max_act[i % A];
            un = fminf(fmaxf(un, -m), m);
        }
        U_prev[i] = u;
        s_u[i] = un;
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nt) : "memory");
    for (int i = threadIdx.x; i < R; i += nt) {
        float v;
        if (i < R - A)                              v = s_u[i + A];
        else if (flags & MPPI_FLAG_REINIT_INIT_ACT) v = prob->init_act[i - (R - A)];
        else                                        v = s_u[i];
        U[i] = v;
        acc[i] = 0;
    }
    if (threadIdx.x < A) next_act[threadIdx.x] = s_u[threadIdx.x];
    if (threadIdx.x == 0) {
        next_act[kMaxAct] = ctl->comm_error ? 1.0f : 0.0f;    // read by the host with next_act
        acc[R] = 0;
        ctl->eta = eta;
        ctl->last_key = ctl->min_key;
        ctl->min_key = kMinKeyInit;
        ctl->step = ctl->step + 1;
        ctl->done = 0;
    }
}

struct FinalizeArgs {
    float *U, *U_prev, *next_act;
    int T, A;
    unsigned flags;
};

}  // namespace mppi
