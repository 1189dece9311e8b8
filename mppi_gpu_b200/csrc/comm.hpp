// comm.hpp -- K-shard collectives: NCCL over NVLink, resolved at run time with dlopen so
// that a single-GPU controller has no NCCL dependency.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace mppi {

class NcclComm {
public:
    NcclComm() = default;
    ~NcclComm();
    NcclComm(const NcclComm &) = delete;
    NcclComm &operator=(const NcclComm &) = delete;

    // fills 128 bytes; returns false + err on failure
    static bool unique_id(uint8_t *id128, std::string &err);

    bool init(int rank, int world, const uint8_t *id128, std::string &err);
    // in-place all-reduce(min) of `count` uint64 values
    bool allreduce_min_u64(unsigned long long *buf, size_t count, cudaStream_t s, std::string &err);
    // in-place all-reduce(sum) of `count` int64 fixed-point accumulators (exact)
    bool allreduce_sum_i64(long long *buf, size_t count, cudaStream_t s, std::string &err);
    bool ready() const { return comm_ != nullptr; }

private:
    void *comm_ = nullptr;
};

}  // namespace mppi
