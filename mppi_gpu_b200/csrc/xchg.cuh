// xchg.cuh -- the K-shard exchange over NVLink peer memory (MPPI_COMM_P2P), as device code that
// the finalizing CTA of ANY chain runs in line: the last CTA of tile_kernel / step_kernel or of
// average_kernel -- compute and collective in one kernel, no exchange launch on the step's
// critical path.
//
// Every rank owns a mailbox with one slot per sender and per parity of the step number; a
// sender stores its contribution straight into every peer's mailbox (st.relaxed.sys over
// NVLink), fences, then publishes the sequence number (st.release.sys); the owner polls its
// own memory (ld.acquire.sys).  Sequence = control step + 1, never reset, compared with >=.
//
// Why two buffers: a rank can enter exchange n+1 only after it has seen every peer's flag of
// exchange n, and a peer raises that flag before it reads the others' data; so a fast rank's
// writes of exchange n+1 may land while a slow rank still reads exchange n -- into the other
// buffer.  Writes of exchange n+2 need the slow rank's flag n+1, which it raises after it has
// consumed n: two buffers are enough, whatever the timing (no reliance on a step being longer
// than a mailbox read).
//
// Replaces the two latency-bound collectives of the K-sharded step:
// ncclAllReduce(min) for beta and ncclAllReduce(sum) for the weighted partials and eta.
#pragma once

#include "common.cuh"
#include "finalize.cuh"

namespace mppi {

struct PeerTable {
    unsigned long long *mb[kMaxWorld];     // mailbox base of every rank, as mapped in THIS process
};

// what the finalizing CTA needs to exchange; world <= 1: single shard, nothing to do
struct XchgArgs {
    PeerTable peers;
    int rank, world;
    unsigned long long slot_words;
};

// %globaltimer stamps of the last exchange, CtlDev::t_xchg: [0] push begins, [1] own data and
// flags are out, [2] every peer's flag has arrived, [3] merged
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq)
{
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(flag) < seq) {
        if (clock64() - t0 > 4000000000ll) return false;      // ~2 s at 2 GHz
        __nanosleep(32);
    }
    return true;
}

// slot of `sender` in buffer `par` of the mailbox at `mb`
__device__ __forceinline__ unsigned long long *mb_slot(unsigned long long *mb, int par, int world,
                                                       int sender, size_t slot_words)
{
    return mb + ((size_t)par * world + sender) * slot_words;
}

// A failed exchange must not touch U: publish the error and the step number (so that the host
// stops waiting) and leave everything else as it is.  The handle is unusable afterwards.
__device__ __forceinline__ void publish_comm_error(CtlDev *ctl, float *next_act)
{
    const unsigned long long step = ctl->step + 1;
    next_act[kMaxAct] = 1.0f;
    ctl->step = step;
    ctl->done = 0;
    st_release_sys_u64(reinterpret_cast<unsigned long long *>(next_act + kNextSeqOffset), step);
}

// ONE exchange per step.  Every shard has averaged with ITS OWN minimum beta_r as the softmax
// reference, so s_acc[0..R-1] = sum_k w~_k eps_k and s_acc[R] = eta_r are relative to beta_r
// (ctl->min_key).  Each rank pushes {key_r, acc_r} to every peer, takes the global minimum key,
// rescales every shard's accumulators by exp(-(beta_r - beta)/lambda) and sums them in rank
// order in double -- the same arithmetic on the same bits on every rank, so the replicated U
// stays bit-identical.  On return (true) s_acc holds the merged sums and ctl->min_key the global
// key.  Called by the threads [0, nthr) of one CTA; s_acc [R+1] and s_f [kMaxWorld + 1] are
// shared memory.  Returns false (for every thread) when a peer did not arrive.
__device__ __forceinline__ bool xchg_merge_body(long long *s_acc, int R,
                                                const ProblemDev *__restrict__ prob, CtlDev *ctl,
                                                const XchgArgs &xa, double *s_f, int nthr, int bar_id)
{
    const int rank = xa.rank, world = xa.world;
    const size_t sw = (size_t)xa.slot_words;
    const unsigned long long seq = ctl->step + 1;
    const int par = (int)(seq & 1ull);
    const unsigned long long mine = ctl->min_key;
    if (threadIdx.x == 0) ctl->t_xchg[0] = globaltimer_ns();
    // 16-byte stores (the accumulators start 32 bytes into a 128-byte aligned slot): half as many
    // NVLink writes to wait for at the fence
    const int npair = (R + 1) / 2;
    for (int r = 0; r < world; ++r) {
        unsigned long long *slot = mb_slot(xa.peers.mb[r], par, world, rank, sw);
        if (threadIdx.x == 0) st_relaxed_sys_u64(slot + 1, mine);
        for (int i = threadIdx.x; i < npair; i += nthr)
            st_relaxed_sys_v2_u64(slot + kMailboxHeaderWords + 2 * i, (unsigned long long)s_acc[2 * i],
                                  (unsigned long long)s_acc[2 * i + 1]);
        if (((R + 1) & 1) && threadIdx.x == 0)
            st_relaxed_sys_u64(slot + kMailboxHeaderWords + R, (unsigned long long)s_acc[R]);
    }
    __threadfence_system();
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (threadIdx.x < world) {
        st_release_sys_u64(mb_slot(xa.peers.mb[threadIdx.x], par, world, rank, sw) + 0, seq);
        if (threadIdx.x == 0) ctl->t_xchg[1] = globaltimer_ns();
        const unsigned long long *in = mb_slot(xa.peers.mb[rank], par, world, threadIdx.x, sw);
        if (!wait_seq(in + 0, seq)) atomicExch(&ctl->comm_error, 1u);
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (*reinterpret_cast<volatile unsigned int *>(&ctl->comm_error)) return false;
    unsigned long long *my = xa.peers.mb[rank] + (size_t)par * world * sw;      // local memory
    if (threadIdx.x == 0) ctl->t_xchg[2] = globaltimer_ns();
    if (threadIdx.x < world) {
        // thread r forms the factor of shard r; every one of them reads all keys (independent
        // loads, in flight together) and takes the same minimum
        unsigned long long keys[kMaxWorld];
        unsigned long long gkey = kMinKeyInit;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < world) {
                keys[r] = ld_relaxed_sys_u64(my + (size_t)r * sw + 1);
                gkey = keys[r] < gkey ? keys[r] : gkey;
            }
        const float beta = ordered_to_float((uint32_t)(gkey >> 32));
        const float nil = prob->neg_inv_lambda;
        unsigned long long kr = kMinKeyInit;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r == (int)threadIdx.x) kr = keys[r];
        const float beta_r = ordered_to_float((uint32_t)(kr >> 32));
        s_f[threadIdx.x] = kr == kMinKeyInit ? 0.0 : (double)expf(__fmul_rn(nil, __fsub_rn(beta_r, beta)));
        if (threadIdx.x == 0) ctl->min_key = gkey;         // beta / argmin of the whole step
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    for (int i = threadIdx.x; i <= R; i += nthr) {
        // all loads first (independent, in flight together), then the sum in rank order
        long long v[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            v[r] = r < world ? (long long)ld_relaxed_sys_u64(my + (size_t)r * sw + kMailboxHeaderWords + i) : 0ll;
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < world) sum += (double)v[r] * s_f[r];
        s_acc[i] = __double2ll_rn(sum);
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (threadIdx.x == 0) ctl->t_xchg[3] = globaltimer_ns();
    return true;
}

}  // namespace mppi
