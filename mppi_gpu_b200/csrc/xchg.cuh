// xchg.cuh -- the K-shard exchange over NVLink peer memory (MPPI_COMM_P2P), as device code that
// the finalizing CTA of ANY chain runs in line: the last CTA of tile_kernel / step_kernel or of
// average_kernel -- compute and collective in one kernel, no exchange launch on the step's
// critical path.
//
// Every rank owns a mailbox with one slot per sender and per parity of the step number; a
// sender stores its contribution straight into every peer's mailbox (st.relaxed.sys over
// NVLink) as 8-byte packets that carry the step's sequence number beside the payload; the owner
// polls its own memory until every packet shows that number.  One NVLink store latency per
// exchange: no fence, no flag, no second round trip.  (The two-exchange flow of
// MPPI_FLAG_SPLIT_KERNELS keeps data + fence + flag, compared with >=.)  Sequence = control
// step + 1, never reset.
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

// %globaltimer stamps of the last exchange, CtlDev::t_xchg: [0] push begins, [1] own packets
// are out, [2] every peer's key has arrived (the wait for the slowest rank), [3] merged
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

// One 8-byte packet of the single exchange: {32 bits of payload, 32 bits of sequence number}.
// An aligned 8-byte store is single-copy atomic, also over NVLink, so a packet that shows the
// step's sequence number carries the step's payload: no fence before a flag, no flag at all.
// (The same idea as the low-latency protocol of NCCL.)
__device__ __forceinline__ void put_packets(unsigned long long *pk, unsigned long long v, uint32_t seq32)
{
    const unsigned long long tag = (unsigned long long)seq32 << 32;
    st_relaxed_sys_u64(pk + 0, tag | (v & 0xffffffffull));
    st_relaxed_sys_u64(pk + 1, tag | (v >> 32));
}
// true when both halves carry seq32; *v is then the 64-bit payload
__device__ __forceinline__ bool get_packets(const unsigned long long *pk, uint32_t seq32,
                                            unsigned long long *v)
{
    const unsigned long long lo = ld_relaxed_sys_u64(pk + 0), hi = ld_relaxed_sys_u64(pk + 1);
    *v = (lo & 0xffffffffull) | (hi << 32);
    return (uint32_t)(lo >> 32) == seq32 && (uint32_t)(hi >> 32) == seq32;
}

// ONE exchange per step.  Every shard has averaged with ITS OWN minimum beta_r as the softmax
// reference, so s_acc[0..R-1] = sum_k w~_k eps_k and s_acc[R] = eta_r are relative to beta_r
// (ctl->min_key).  Each rank pushes {key_r, acc_r} to every peer as packets, takes the global
// minimum key, rescales every shard's accumulators by exp(-(beta_r - beta)/lambda) and sums them
// in rank order in double -- the same arithmetic on the same bits on every rank, so the
// replicated U stays bit-identical.  On return (true) s_acc holds the merged sums and
// ctl->min_key the global key.  Called by the threads [0, nthr) of one CTA; s_acc [R+1] and s_f
// [2 * (kMaxWorld + 1)] (rescale factors, then the shards' keys) are shared memory.  Returns false (for every thread) when a peer did not
// arrive within about two seconds.
__device__ __forceinline__ bool xchg_merge_body(long long *s_acc, int R,
                                                const ProblemDev *__restrict__ prob, CtlDev *ctl,
                                                const XchgArgs &xa, double *s_f, int nthr, int bar_id)
{
    const int rank = xa.rank, world = xa.world;
    const size_t sw = (size_t)xa.slot_words;
    const unsigned long long seq = ctl->step + 1;
    const uint32_t seq32 = (uint32_t)seq;
    const int par = (int)(seq & 1ull);
    const unsigned long long mine = ctl->min_key;
    if (threadIdx.x == 0) ctl->t_xchg[0] = globaltimer_ns();
    // ---- push: element 0 is the key, elements 1..R+1 the accumulators
    for (int r = 0; r < world; ++r) {
        unsigned long long *slot = mb_slot(xa.peers.mb[r], par, world, rank, sw);
        for (int j = threadIdx.x; j <= R + 1; j += nthr)
            put_packets(slot + 2 * j, j == 0 ? mine : (unsigned long long)s_acc[j - 1], seq32);
    }
    if (threadIdx.x == 0) ctl->t_xchg[1] = globaltimer_ns();
    // ---- every peer's key (this is where a slow rank is waited for)
    unsigned long long *my = xa.peers.mb[rank] + (size_t)par * world * sw;      // local memory
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_f + kMaxWorld + 1);
    if (threadIdx.x < world) {
        const unsigned long long *pk = my + (size_t)threadIdx.x * sw;
        unsigned long long k = kMinKeyInit;
        const long long t0 = clock64();
        while (!get_packets(pk, seq32, &k)) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(&ctl->comm_error, 1u); break; }
            __nanosleep(32);
        }
        s_key[threadIdx.x] = k;
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (*reinterpret_cast<volatile unsigned int *>(&ctl->comm_error)) return false;
    if (threadIdx.x == 0) {
        ctl->t_xchg[2] = globaltimer_ns();
        unsigned long long gkey = kMinKeyInit;
        for (int r = 0; r < world; ++r) gkey = s_key[r] < gkey ? s_key[r] : gkey;
        const float beta = ordered_to_float((uint32_t)(gkey >> 32));
        const float nil = prob->neg_inv_lambda;
        for (int r = 0; r < world; ++r) {
            const float beta_r = ordered_to_float((uint32_t)(s_key[r] >> 32));
            s_f[r] = s_key[r] == kMinKeyInit ? 0.0
                                             : (double)expf(__fmul_rn(nil, __fsub_rn(beta_r, beta)));
        }
        ctl->min_key = gkey;                      // beta / argmin of the whole step
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    // ---- the accumulators: all ranks' packets of an element are loaded together (independent
    //      loads in flight), re-polled until complete, then summed in rank order
    bool late = false;
    for (int i = threadIdx.x; i <= R; i += nthr) {
        unsigned long long v[kMaxWorld];
        const long long t0 = clock64();
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r)
                if (r < world) ok &= get_packets(my + (size_t)r * sw + 2 * (i + 1), seq32, &v[r]);
            if (ok) break;
            if (clock64() - t0 > 4000000000ll) { late = true; break; }
        }
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < world) sum += (double)(long long)v[r] * s_f[r];
        s_acc[i] = __double2ll_rn(sum);
    }
    if (late) atomicExch(&ctl->comm_error, 1u);
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (*reinterpret_cast<volatile unsigned int *>(&ctl->comm_error)) return false;
    if (threadIdx.x == 0) ctl->t_xchg[3] = globaltimer_ns();
    return true;
}

}  // namespace mppi
