// finalize.cuh -- the U update + receding-horizon shift executed by ONE CTA (part 5).
#pragma once

#include "common.cuh"

namespace mppi {

// U update + receding-horizon shift from the fixed-point accumulators, executed by ONE CTA
// (the finalize kernel, or the last CTA of the merged average kernel).  s_u: T*A floats.
//
// next_act points into PINNED HOST memory mapped into the device (zero copy): the kernel stores
// the A floats of the next action and the exchange-error flag straight into the caller's
// process, then publishes the step counter at next_act + kNextSeqOffset with a system-scope
// release -- as soon as the new U[0] is known, before the shift.  The host spins on that word
// instead of waiting for a D2H copy node and a stream synchronisation (mppi_step_wait) --
// several microseconds of a closed-loop control step.
constexpr int kNextSeqOffset = 2 * kMaxAct;     // in floats; 8-byte aligned
constexpr int kNextFloats = 2 * kMaxAct + 2;
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
    // Publish first: the action is all the caller is waiting for.  One thread stores the A
    // floats and the exchange-error flag and then the step counter with a system-scope release,
    // which orders its own earlier stores before it -- one round trip to host memory.  The
    // shift and the re-arming below finish while the host is already on its way; whatever it
    // does to this handle next is ordered behind this kernel by the stream.
    unsigned long long step = 0;
    if (threadIdx.x == 0) {
        step = ctl->step + 1;
        for (int a = 0; a < A; ++a) next_act[a] = s_u[a];
        next_act[kMaxAct] = ctl->comm_error ? 1.0f : 0.0f;               // read with next_act
        st_release_sys_u64(reinterpret_cast<unsigned long long *>(next_act + kNextSeqOffset), step);
    }
    for (int i = threadIdx.x; i < R; i += nt) {
        float v;
        if (i < R - A)                              v = s_u[i + A];
        else if (flags & MPPI_FLAG_REINIT_INIT_ACT) v = prob->init_act[i - (R - A)];
        else                                        v = s_u[i];
        U[i] = v;
        acc[i] = 0;
    }
    if (threadIdx.x == 0) {
        acc[R] = 0;
        ctl->eta = eta;
        ctl->last_key = ctl->min_key;
        ctl->min_key = kMinKeyInit;
        ctl->step = step;
        ctl->done = 0;
    }
}

struct FinalizeArgs {
    float *U, *U_prev, *next_act;
    int T, A;
    unsigned flags;
};

}  // namespace mppi
