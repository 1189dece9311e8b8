// merge.cuh -- the end of a one-kernel control step: the LAST CTA (ticket) merges the per-CTA
// records {row sums [R], eta, ref} and applies the U update (part 5).
//
// Every CTA of step_kernel (step.cu) / tile_kernel (tile.cu) has averaged relative to its own
// running minimum `ref` (online softmax).  The records are merged in CTA order, each rescaled by
// exp(-(ref_c - beta)/lambda), in double; the sums are converted to the 64-bit fixed-point
// accumulators the rest of the library (finalize_body, the K-shard exchange) works on.
// Replaces the reference's per-t sum_red_adim folds + copy_act + shift_act
// (src/point_mass.cu:439-471, :668-741, :756-761, :805-824).
#pragma once

#include "common.cuh"
#include "finalize.cuh"
#include "xchg.cuh"

namespace mppi {

// floats per record: a multiple of four so that records can be bulk-copied (16-byte granules)
__host__ __device__ inline int record_stride(int R) { return (R + 2 + 3) & ~3; }

// shared memory the merge needs: merged sums, U_new, the rescale factors and at least eight
// records per bulk-copy pass
__host__ __device__ inline size_t merge_smem_bytes(int R, int grid)
{
    return (size_t)(R + 1) * 8 + (size_t)(R + grid) * 4 + 128 + (size_t)8 * record_stride(R) * 4;
}

// Called by the threads [0, nthr) of the last CTA, all of them, after the ticket showed that
// every record is complete.  `region` (128-byte aligned, `cap` bytes) is shared memory nobody
// else uses any more; `mbar` an 8-byte aligned shared word; named barrier `bar_id` is theirs.
// Single shard: U update + shift + publish follow directly.  K-shards (xa.world > 1): the merged
// sums are relative to this shard's own minimum (ctl->min_key); the same CTA exchanges them with
// the peers over NVLink (xchg.cuh) before the U update -- compute and collective in one kernel.
template <int kMaxOut>
__device__ __noinline__ void merge_records(const float *__restrict__ part, int R, int nc,
                                              uint8_t *region, size_t cap, uint64_t *mbar, int nthr,
                                              int bar_id, const ProblemDev *__restrict__ prob,
                                              CtlDev *ctl, const FinalizeArgs &fin,
                                              const XchgArgs &xa)
{
    const int rstride = record_stride(R);
    __threadfence();
    const float nil = prob->neg_inv_lambda;
    const unsigned long long mk = *reinterpret_cast<volatile unsigned long long *>(&ctl->min_key);
    const float beta = ordered_to_float((uint32_t)(mk >> 32));
    // The records are pulled into shared memory with ONE bulk copy per pass (as many records as
    // fit) instead of strided L2 loads with a handful in flight per thread.
    long long *s_acc = reinterpret_cast<long long *>(region);       // [R+1] merged, fixed point
    float *s_unew = reinterpret_cast<float *>(s_acc + (R + 1));     // [R]
    float *s_f = s_unew + R;                                        // [per pass]
    const size_t rec_off = (((size_t)(R + 1) * 8 + (size_t)R * 4 + (size_t)nc * 4) + 127) & ~(size_t)127;
    float *s_rec = reinterpret_cast<float *>(region + rec_off);     // [per pass][rstride]
    const int per = (int)((cap - rec_off) / ((size_t)rstride * sizeof(float)));
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    double sum[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) sum[o] = 0.0;
    uint32_t parity = 0;
    for (int c0 = 0; c0 < nc; c0 += per, parity ^= 1) {
        const int n = min(per, nc - c0);
        asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");   // region free, barrier initialised
        if (threadIdx.x == 0) {
            fence_proxy_async_all();                                // generic accesses above -> async copy
            const uint32_t bytes = (uint32_t)((size_t)n * rstride * sizeof(float));
            mbar_arrive_expect_tx(mbar, bytes);
            bulk_load_1d(s_rec, part + (size_t)c0 * rstride, bytes, mbar);
        }
        mbar_wait(mbar, parity);
        for (int c = threadIdx.x; c < n; c += nthr)
            s_f[c] = expf(__fmul_rn(nil, __fsub_rn(s_rec[(size_t)c * rstride + R + 1], beta)));   // +inf -> 0
        asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            const int i = threadIdx.x + o * nthr;                   // i == R: eta
            if (i <= R) {
                double a = sum[o];
                for (int c = 0; c < n; ++c)                         // fixed order: CTA 0, 1, 2, ...
                    a += (double)s_rec[(size_t)c * rstride + i] * (double)s_f[c];
                sum[o] = a;
            }
        }
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
        const int i = threadIdx.x + o * nthr;
        if (i <= R) s_acc[i] = __double2ll_rn(sum[o] * kAccScale);   // stays on chip
    }
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
    if (xa.world > 1) {
        double *s_xf = reinterpret_cast<double *>(s_rec);           // the record buffer is free
        if (!xchg_merge_body(s_acc, R, prob, ctl, xa, s_xf, nthr, bar_id)) {
            if (threadIdx.x == 0) publish_comm_error(ctl, fin.next_act);
            return;
        }
    }
    finalize_body(s_acc, fin.U, fin.U_prev, prob, ctl, fin.next_act, fin.T, fin.A, fin.flags,
                  s_unew, nthr, bar_id);
}

}  // namespace mppi
