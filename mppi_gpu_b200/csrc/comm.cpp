// comm.cpp -- NCCL binding by dlopen (see comm.hpp).
#include "comm.hpp"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

namespace mppi {
namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string load_error;
};

NcclApi &api()
{
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            a.load_error = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return;
        }
#define MPPI_NCCL_SYM(field, name)                                                  \
        a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name));          \
        if (!a.field) { a.load_error = std::string("missing NCCL symbol ") + name; return; }
        MPPI_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        MPPI_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        MPPI_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        MPPI_NCCL_SYM(AllReduce, "ncclAllReduce")
        MPPI_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MPPI_NCCL_SYM
    });
    return a;
}

bool check(ncclResult_t r, const char *what, std::string &err)
{
    if (r == ncclSuccess) return true;
    err = std::string(what) + ": " + api().GetErrorString(r);
    return false;
}

}  // namespace

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

bool NcclComm::unique_id(uint8_t *id128, std::string &err)
{
    NcclApi &a = api();
    if (!a.load_error.empty()) { err = a.load_error; return false; }
    ncclUniqueId id;
    if (!check(a.GetUniqueId(&id), "ncclGetUniqueId", err)) return false;
    memcpy(id128, &id, sizeof id);
    return true;
}

bool NcclComm::init(int rank, int world, const uint8_t *id128, std::string &err)
{
    NcclApi &a = api();
    if (!a.load_error.empty()) { err = a.load_error; return false; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    if (!check(a.CommInitRank(&c, world, id, rank), "ncclCommInitRank", err)) return false;
    comm_ = c;
    return true;
}

NcclComm::~NcclComm()
{
    if (comm_) api().CommDestroy(static_cast<ncclComm_t>(comm_));
}

bool NcclComm::allreduce_min_u64(unsigned long long *buf, size_t count, cudaStream_t s,
                                 std::string &err)
{
    return check(api().AllReduce(buf, buf, count, ncclUint64, ncclMin,
                                 static_cast<ncclComm_t>(comm_), s),
                 "ncclAllReduce(min,u64)", err);
}

bool NcclComm::allreduce_sum_i64(long long *buf, size_t count, cudaStream_t s, std::string &err)
{
    return check(api().AllReduce(buf, buf, count, ncclInt64, ncclSum,
                                 static_cast<ncclComm_t>(comm_), s),
                 "ncclAllReduce(sum,i64)", err);
}

}  // namespace mppi
