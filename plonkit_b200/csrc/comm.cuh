// Collectives of the sharded single-proof prover: the few exchange steps one proof split across the
// GPUs of a node needs — an all-gather of partial G1 sums (the "all-reduce" of the EC group: all-gather + local fold),
// one all-to-all inside the size-4n inverse NTT of the quotient, and an all-gather of the quotient coefficients.
//
// Two transports behind one interface:
//   * NcclComm  — one process per GPU (torchrun): NCCL over NVLink 5 / NVSwitch.  libnccl.so.2 is resolved at run time
//                 (dlopen), so inside a process that already imported torch the library's calls land in torch's own
//                 NCCL and no second copy is loaded.  Stream-ordered, no host synchronisation.
//   * LocalComm — `world` ranks as threads of ONE process, each with its own context (stream, scratch), on the same
//                 device or on several devices with peer access: the collective is a set of direct device-to-device
//                 copies out of the peers' buffers (NVLink peer copies between devices).  This is what the
//                 single-GPU parity tests drive with 2, 4 and 8 virtual ranks.
#pragma once
#include <condition_variable>
#include <mutex>

#include "common.cuh"

namespace pk {

struct Comm {
    int rank = 0, world = 1;
    virtual ~Comm() {}
    // recv[q * bytes ...] = the `bytes` at `send` on rank q, for every q; send is either disjoint from recv or exactly
    // recv + rank * bytes (in place)
    virtual void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st) = 0;
    // `count` all-gathers at once: recv[c] + q * bytes = send[c] of rank q
    virtual void all_gather_multi(const void* const* send, void* const* recv, int count, size_t bytes, cudaStream_t st) = 0;
    // block q (bytes each) of send goes to rank q; block q of recv comes from rank q
    virtual void all_to_all(const void* send, void* recv, size_t bytes, cudaStream_t st) = 0;
    // Direct access to the peers' memory (NVLink / NVSwitch peer mappings): out[q] = an address in THIS rank's address
    // space through which kernels of this rank can store into the `bytes`-long buffer `mine` of rank q (out[rank] = mine).
    // Collective; returns false on every rank if any pair of ranks cannot map each other (the caller then stays on the
    // collectives above).  The mappings live as long as the communicator.
    virtual bool map_peers(void* mine, size_t bytes, void** out, cudaStream_t st) = 0;
    // all ranks' work enqueued so far (peer stores included) is complete and visible before any rank's later work starts
    virtual void stream_barrier(cudaStream_t st) = 0;
    // first thing in every collective API call: all ranks meet; a failure of the previous call is forgotten here
    virtual void begin() {}
    // called by a rank that is about to fail, so that peers blocked in a collective fail too instead of hanging
    virtual void abort() {}
};

// shared state of an in-process group (threads)
struct LocalGroup {
    int world;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0, entry_arrived = 0;
    uint64_t generation = 0, entry_generation = 0;
    bool aborted = false;
    std::vector<void*> map_ptr;           // [world] published buffers (map_peers)
    std::vector<const void*> ptr;         // [world] published send pointers
    std::vector<const void* const*> ptrs; // [world] published pointer lists (all_gather_multi)
    std::vector<int> device;              // [world]
    explicit LocalGroup(int w) : world(w), map_ptr(w, nullptr), ptr(w, nullptr), ptrs(w, nullptr), device(w, -1) {}
    void barrier();  // throws PkError if the group was aborted
    void entry_barrier();  // all ranks meet at the start of a call; clears an abort left by the previous call
    void abort();
};

Comm* make_local_comm(LocalGroup* g, int rank, int device);
Comm* make_nccl_comm(const uint8_t unique_id[128], int rank, int world);
void nccl_unique_id(uint8_t out[128]);

}  // namespace pk
