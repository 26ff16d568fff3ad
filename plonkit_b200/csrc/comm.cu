// Transports of comm.cuh.
#include "comm.cuh"

#include <dlfcn.h>

#include <memory>
#include <nccl.h>  // types only: every NCCL function is resolved with dlsym (see NcclApi)

namespace pk {

// ---------------------------------------------------------------- in-process group (threads)
void LocalGroup::barrier() {
    std::unique_lock<std::mutex> lk(mu);
    PK_REQUIRE(!aborted, PK_ERR_INVALID, "a peer rank of the sharded prover failed");
    const uint64_t gen = generation;
    if (++arrived == world) {
        arrived = 0;
        ++generation;
        cv.notify_all();
        return;
    }
    cv.wait(lk, [&] { return generation != gen; });
    PK_REQUIRE(!aborted, PK_ERR_INVALID, "a peer rank of the sharded prover failed");
}
void LocalGroup::abort() {  // releases every rank waiting in barrier(); later arrivals of this call fail at its entry
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    arrived = 0;
    ++generation;
    cv.notify_all();
}
void LocalGroup::entry_barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t gen = entry_generation;
    if (++entry_arrived == world) {
        entry_arrived = 0;
        aborted = false;
        arrived = 0;
        ++entry_generation;
        cv.notify_all();
        return;
    }
    cv.wait(lk, [&] { return entry_generation != gen; });
}

struct LocalComm : Comm {
    LocalGroup* g;
    bool peers_done = false;
    LocalComm(LocalGroup* grp, int r, int device) : g(grp) {
        rank = r;
        world = grp->world;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->device[r] = device;
        }
    }
    void enable_peers() {  // after a barrier: every rank has attached, so every device is known
        if (peers_done) return;
        peers_done = true;
        for (int q = 0; q < world; ++q) {
            const int d = g->device[q];
            if (d < 0 || d == g->device[rank]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->device[rank], d);
            if (can && cudaDeviceEnablePeerAccess(d, 0) != cudaSuccess) cudaGetLastError();  // already enabled is fine
        }
    }
    // every rank: finish producing -> publish -> barrier -> copy out of the peers' buffers -> barrier (peers may now reuse them)
    void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        PK_CUDA(cudaStreamSynchronize(st));
        g->ptr[rank] = send;
        g->barrier();
        enable_peers();
        for (int q = 0; q < world; ++q) {
            uint8_t* dst = static_cast<uint8_t*>(recv) + (size_t)q * bytes;
            if (dst != g->ptr[q])  // in place (send == recv + rank * bytes): this rank's block is already there
                PK_CUDA(cudaMemcpyAsync(dst, g->ptr[q], bytes, cudaMemcpyDefault, st));
        }
        PK_CUDA(cudaStreamSynchronize(st));
        g->barrier();
    }
    void all_gather_multi(const void* const* send, void* const* recv, int count, size_t bytes, cudaStream_t st) override {
        PK_CUDA(cudaStreamSynchronize(st));
        g->ptrs[rank] = send;
        g->barrier();
        enable_peers();
        for (int c = 0; c < count; ++c)
            for (int q = 0; q < world; ++q)
                PK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(recv[c]) + (size_t)q * bytes, g->ptrs[q][c], bytes, cudaMemcpyDefault, st));
        PK_CUDA(cudaStreamSynchronize(st));
        g->barrier();
    }
    void all_to_all(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        PK_CUDA(cudaStreamSynchronize(st));
        g->ptr[rank] = send;
        g->barrier();
        enable_peers();
        for (int q = 0; q < world; ++q)  // block `rank` of peer q's send buffer is addressed to this rank
            PK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(recv) + (size_t)q * bytes,
                                    static_cast<const uint8_t*>(g->ptr[q]) + (size_t)rank * bytes, bytes, cudaMemcpyDefault, st));
        PK_CUDA(cudaStreamSynchronize(st));
        g->barrier();
    }
    bool map_peers(void* mine, size_t, void** out, cudaStream_t st) override {
        PK_CUDA(cudaStreamSynchronize(st));
        g->map_ptr[rank] = mine;
        g->barrier();
        enable_peers();
        bool ok = true;
        for (int q = 0; q < world; ++q) {   // same process: the peer's pointer is valid here once peer access is on
            out[q] = g->map_ptr[q];
            if (g->device[q] != g->device[rank]) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, g->device[rank], g->device[q]);
                ok = ok && can;
            }
        }
        g->barrier();
        return ok;  // the device topology is the same from every rank's side on an NVSwitch box
    }
    void stream_barrier(cudaStream_t st) override {
        PK_CUDA(cudaStreamSynchronize(st));
        g->barrier();
    }
    void begin() override { g->entry_barrier(); }
    void abort() override { g->abort(); }
};

Comm* make_local_comm(LocalGroup* g, int rank, int device) {
    PK_REQUIRE(g && rank >= 0 && rank < g->world, PK_ERR_INVALID, "rank outside the group");
    LocalComm* c = new LocalComm(g, rank, device);
    return c;
}

// ---------------------------------------------------------------- NCCL (one process per GPU)
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl_api() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (api.h) return api;
    // already-loaded copies (torch's bundled NCCL has the same SONAME) are reused by the dynamic loader
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    PK_REQUIRE(h != nullptr, PK_ERR_CUDA, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        PK_REQUIRE(p != nullptr, PK_ERR_CUDA, std::string("libnccl lacks ") + name);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.h = h;
    return api;
}
#define PK_NCCL(expr)                                                                                              \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess)                                                                                     \
            throw ::pk::PkError(PK_ERR_CUDA, std::string(#expr) + ": " + nccl_api().GetErrorString(_r));           \
    } while (0)

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    uint8_t* scratch = nullptr;          // device: [world][64] IPC handles / flags, and the barrier word
    std::vector<void*> opened;           // cudaIpcOpenMemHandle results to close
    ~NcclComm() override {
        for (void* p : opened) cudaIpcCloseMemHandle(p);
        if (scratch) cudaFree(scratch);
        if (comm) nccl_api().CommDestroy(comm);
    }
    // one process per GPU: the peers' buffers are mapped through CUDA IPC handles, exchanged with an all-gather
    bool map_peers(void* mine, size_t, void** out, cudaStream_t st) override {
        if (!scratch) PK_CUDA(cudaMalloc(&scratch, (size_t)(world + 1) * 64 + 64));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof(h));
        bool ok = cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
        if (!ok) cudaGetLastError();
        uint8_t* send = scratch + (size_t)world * 64;
        PK_CUDA(cudaMemcpyAsync(send, &h, 64, cudaMemcpyHostToDevice, st));
        PK_NCCL(nccl_api().AllGather(send, scratch, 64, ncclUint8, comm, st));
        std::vector<cudaIpcMemHandle_t> all(world);
        PK_CUDA(cudaMemcpyAsync(all.data(), scratch, (size_t)world * 64, cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
        for (int q = 0; q < world; ++q) {
            if (q == rank) { out[q] = mine; continue; }
            void* p = nullptr;
            if (ok && cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                opened.push_back(p);
                out[q] = p;
            } else {
                cudaGetLastError();
                ok = false;
                out[q] = nullptr;
            }
        }
        // agreement: every rank must take the same path
        int32_t flag = ok ? 1 : 0, result = 0;
        int32_t* word = reinterpret_cast<int32_t*>(scratch + (size_t)world * 64 + 64 - 8);
        PK_CUDA(cudaMemcpyAsync(word, &flag, 4, cudaMemcpyHostToDevice, st));
        PK_NCCL(nccl_api().AllReduce(word, word, 1, ncclInt32, ncclMin, comm, st));
        PK_CUDA(cudaMemcpyAsync(&result, word, 4, cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
        return result == 1;
    }
    void stream_barrier(cudaStream_t st) override {
        if (!scratch) PK_CUDA(cudaMalloc(&scratch, (size_t)(world + 1) * 64 + 64));
        int32_t* word = reinterpret_cast<int32_t*>(scratch + (size_t)world * 64 + 64 - 16);
        PK_NCCL(nccl_api().AllReduce(word, word, 1, ncclInt32, ncclSum, comm, st));  // stream-ordered on every rank
    }
    void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        PK_NCCL(nccl_api().AllGather(send, recv, bytes, ncclUint8, comm, st));
    }
    void all_gather_multi(const void* const* send, void* const* recv, int count, size_t bytes, cudaStream_t st) override {
        NcclApi& a = nccl_api();
        PK_NCCL(a.GroupStart());
        for (int c = 0; c < count; ++c) PK_NCCL(a.AllGather(send[c], recv[c], bytes, ncclUint8, comm, st));
        PK_NCCL(a.GroupEnd());
    }
    void all_to_all(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        NcclApi& a = nccl_api();
        PK_NCCL(a.GroupStart());
        for (int q = 0; q < world; ++q) {
            PK_NCCL(a.Send(static_cast<const uint8_t*>(send) + (size_t)q * bytes, bytes, ncclUint8, q, comm, st));
            PK_NCCL(a.Recv(static_cast<uint8_t*>(recv) + (size_t)q * bytes, bytes, ncclUint8, q, comm, st));
        }
        PK_NCCL(a.GroupEnd());
    }
};

void nccl_unique_id(uint8_t out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PK_NCCL(nccl_api().GetUniqueId(&id));
    memcpy(out, &id, 128);
}

Comm* make_nccl_comm(const uint8_t unique_id[128], int rank, int world) {
    PK_REQUIRE(world >= 1 && rank >= 0 && rank < world, PK_ERR_INVALID, "rank outside the communicator");
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    std::unique_ptr<NcclComm> c(new NcclComm());
    c->rank = rank;
    c->world = world;
    PK_NCCL(nccl_api().CommInitRank(&c->comm, world, id, rank));
    return c.release();
}

}  // namespace pk
