#include "shard.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <stdexcept>
#include <vector>

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
        for (void *p : opened_) cudaIpcCloseMemHandle(p);
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
    // cudaIpc handles travel through one NCCL all-gather (the only bootstrap channel this library has); whether
    // every rank could map every arena is agreed on with an all-reduce, so all ranks pick the same data plane.
    bool shareArena(void *localBase, size_t bytes, std::vector<void *> &peerBase, std::string &err) override {
        (void) bytes;
        peerBase.assign(world_, nullptr);
        peerBase[rank_] = localBase;
        cudaIpcMemHandle_t mine;
        int ok = cudaIpcGetMemHandle(&mine, localBase) == cudaSuccess ? 1 : 0;
        if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
        unsigned char *dSend = nullptr, *dAll = nullptr;
        int *dOk = nullptr;
        cudaStream_t s = nullptr;
        if (cudaMalloc(&dSend, 64) != cudaSuccess || cudaMalloc(&dAll, (size_t) 64 * world_) != cudaSuccess ||
            cudaMalloc(&dOk, sizeof(int)) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
            err = "shareArena: cudaMalloc failed";
            return false;
        }
        std::vector<cudaIpcMemHandle_t> all(world_);
        cudaMemcpyAsync(dSend, &mine, 64, cudaMemcpyHostToDevice, s);
        CFB_NCCL(api().AllGather(dSend, dAll, 64, ncclChar, comm_, s));
        cudaMemcpyAsync(all.data(), dAll, (size_t) 64 * world_, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        for (int q = 0; q < world_ && ok; ++q) {
            if (q == rank_) continue;
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            peerBase[q] = p;
            opened_.push_back(p);
        }
        cudaMemcpyAsync(dOk, &ok, sizeof(int), cudaMemcpyHostToDevice, s);
        CFB_NCCL(api().AllReduce(dOk, dOk, 1, ncclInt, ncclMin, comm_, s));
        int allOk = 0;
        cudaMemcpyAsync(&allOk, dOk, sizeof(int), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        cudaFree(dSend); cudaFree(dAll); cudaFree(dOk);
        cudaStreamDestroy(s);
        if (!allOk) {
            for (void *p : opened_) cudaIpcCloseMemHandle(p);
            opened_.clear();
            err = "peer access between the GPUs of this run is not available (cudaIpcOpenMemHandle failed on some rank)";
            return false;
        }
        return true;
    }

private:
    int rank_, world_;
    ncclComm_t comm_ = nullptr;
    std::vector<void *> opened_;   // peers' arenas mapped with cudaIpcOpenMemHandle
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
