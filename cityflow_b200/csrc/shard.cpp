#include "shard.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <stdexcept>

namespace cfb {
namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string &err) {
        if (lib) return true;
        // the copy torch already mapped (same soname) is reused when present
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p) err = std::string("libnccl lacks ") + n; return p; };
        GetUniqueId = (decltype(GetUniqueId)) sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank)) sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy)) sym("ncclCommDestroy");
        Send = (decltype(Send)) sym("ncclSend");
        Recv = (decltype(Recv)) sym("ncclRecv");
        GroupStart = (decltype(GroupStart)) sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd)) sym("ncclGroupEnd");
        AllGather = (decltype(AllGather)) sym("ncclAllGather");
        AllReduce = (decltype(AllReduce)) sym("ncclAllReduce");
        GetErrorString = (decltype(GetErrorString)) sym("ncclGetErrorString");
        return err.empty();
    }
};

NcclApi &api() {
    static NcclApi a;
    return a;
}

#define CFB_NCCL(x)                                                                                     \
    do {                                                                                                \
        ncclResult_t r_ = (x);                                                                          \
        if (r_ != ncclSuccess)                                                                          \
            throw std::runtime_error(std::string("cityflow_b200 NCCL error: ") + api().GetErrorString(r_)); \
    } while (0)

class NcclTransport : public ShardTransport {
public:
    NcclTransport(int rank, int world, const void *id, int device) : rank_(rank), world_(world) {
        cudaSetDevice(device);
        ncclUniqueId uid;
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
        memcpy(&uid, id, sizeof(uid));
        CFB_NCCL(api().CommInitRank(&comm_, world, uid, rank));
    }
    ~NcclTransport() override {
        if (comm_) api().CommDestroy(comm_);
    }
    int rank() const override { return rank_; }
    int world() const override { return world_; }
    void exchange(void *stream, const void *send, const std::vector<int> &sb, void *recv, const std::vector<int> &rb,
                  size_t bytes) override {
        cudaStream_t s = (cudaStream_t) stream;
        CFB_NCCL(api().GroupStart());
        for (int q = 0; q < world_; ++q) {
            if (q == rank_) continue;
            const size_t ns = (size_t) (sb[q + 1] - sb[q]) * bytes, nr = (size_t) (rb[q + 1] - rb[q]) * bytes;
            if (ns) CFB_NCCL(api().Send((const char *) send + (size_t) sb[q] * bytes, ns, ncclChar, q, comm_, s));
            if (nr) CFB_NCCL(api().Recv((char *) recv + (size_t) rb[q] * bytes, nr, ncclChar, q, comm_, s));
        }
        CFB_NCCL(api().GroupEnd());
    }
    void allGather(void *stream, const void *send, void *recvAll, size_t bytesPerRank) override {
        CFB_NCCL(api().AllGather(send, recvAll, bytesPerRank, ncclChar, comm_, (cudaStream_t) stream));
    }
    void exchangeAndGather(void *stream, const void *send, const std::vector<int> &sb, void *recv, const std::vector<int> &rb,
                           size_t bytes, const void *gSend, void *gRecvAll, size_t gBytes) override {
        cudaStream_t s = (cudaStream_t) stream;
        CFB_NCCL(api().GroupStart());
        for (int q = 0; q < world_; ++q) {
            if (q == rank_) continue;
            const size_t ns = (size_t) (sb[q + 1] - sb[q]) * bytes, nr = (size_t) (rb[q + 1] - rb[q]) * bytes;
            if (ns) CFB_NCCL(api().Send((const char *) send + (size_t) sb[q] * bytes, ns, ncclChar, q, comm_, s));
            if (nr) CFB_NCCL(api().Recv((char *) recv + (size_t) rb[q] * bytes, nr, ncclChar, q, comm_, s));
            CFB_NCCL(api().Send(gSend, gBytes, ncclChar, q, comm_, s));
            CFB_NCCL(api().Recv((char *) gRecvAll + (size_t) q * gBytes, gBytes, ncclChar, q, comm_, s));
        }
        CFB_NCCL(api().GroupEnd());
    }
    void allReduceSumInt(void *stream, int *devBuf, int n) override {
        CFB_NCCL(api().AllReduce(devBuf, devBuf, (size_t) n, ncclInt, ncclSum, comm_, (cudaStream_t) stream));
    }

private:
    int rank_, world_;
    ncclComm_t comm_ = nullptr;
};

}  // namespace

ShardTransport *createNcclTransport(int rank, int world, const void *uniqueId, int device, std::string &err) {
    if (!api().load(err)) return nullptr;
    try {
        return new NcclTransport(rank, world, uniqueId, device);
    } catch (const std::exception &e) {
        err = e.what();
        return nullptr;
    }
}

bool ncclUniqueIdBytes(unsigned char out[128], std::string &err) {
    if (!api().load(err)) return false;
    ncclUniqueId id;
    if (api().GetUniqueId(&id) != ncclSuccess) { err = "ncclGetUniqueId failed"; return false; }
    memcpy(out, &id, 128);
    return true;
}

}  // namespace cfb
