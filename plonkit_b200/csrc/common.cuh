// Shared host-side plumbing for the CUDA library: context, error handling, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/plonkit_b200.h"
#include "ec.cuh"
#include "fp.cuh"

namespace pk {

struct PkError : std::runtime_error {
    int code;
    PkError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PK_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            throw ::pk::PkError(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                                 std::to_string(__LINE__) + ")");                            \
    } while (0)
#define PK_REQUIRE(cond, code, msg)                      \
    do {                                                 \
        if (!(cond)) throw ::pk::PkError((code), (msg)); \
    } while (0)

// RAII device buffer
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        if (count) PK_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    T* get() const { return p; }
    operator T*() const { return p; }
};

struct Profile {
    bool enabled = false;
    uint64_t kernel_launches = 0;
    uint64_t msm_accum_launches = 0, msm_accum_points = 0;
    double msm_accum_ms = 0;
    uint64_t ntt_launches = 0, ntt_elements = 0;
    double ntt_ms = 0;
    double phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double comm_ms[3] = {0, 0, 0};
    uint64_t comm_bytes[3] = {0, 0, 0};
    // pending event pairs (resolved lazily after a sync)
    struct Pending { cudaEvent_t a, b; int kind; uint64_t units; };
    std::vector<Pending> pending;
};

struct SrsTables;   // msm.cu
struct DomainCache; // ntt.cu
struct Comm;        // comm.cuh

}  // namespace pk

struct pk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // side stream: work of a proof that does not depend on the commitment in flight (the coset LDEs of polynomials that
    // are already in coefficient form) runs here, beside the MSM kernels on `stream`, and is joined before its first use
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false;
    int sm_count = 148;
    std::string last_error;
    pk::Profile prof;
    pk::SrsTables* srs = nullptr;
    pk::SrsTables* srs_lagrange = nullptr;   // optional Lagrange-form key (wire commitments from values)
    pk::DomainCache* domains = nullptr;
    pk::Comm* comm = nullptr;          // set by pk_comm_attach_*: this context is one rank of a sharded prover
    // small pinned staging area for results (commitments, scalars)
    uint8_t* pinned = nullptr;
    size_t pinned_bytes = 0;
};

namespace pk {

// event-pair timing of a kernel family on ctx->stream while profiling is enabled
struct ScopedKernelTimer {
    pk_ctx* ctx;
    int kind;
    uint64_t units;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t stream;
    ScopedKernelTimer(pk_ctx* c, int k, uint64_t u, cudaStream_t st = nullptr) : ctx(c), kind(k), units(u), stream(st ? st : c->stream) {
        if (ctx->prof.enabled) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, stream);
        }
    }
    ~ScopedKernelTimer() {
        if (a) {
            cudaEventRecord(b, stream);
            ctx->prof.pending.push_back({a, b, kind, units});
        }
    }
};
void profile_resolve(pk_ctx* ctx);  // capi.cu

// Scope in which the library's launches go to the side stream: it starts after everything enqueued on the main stream so
// far; side_join() makes the main stream wait for everything the side stream was given.
struct SideStreamScope {
    pk_ctx* ctx;
    cudaStream_t main;
    bool active;
    // while per-kernel profiling is on everything stays on the main stream: event pairs around a kernel family must not
    // include the time it spends waiting for SMs the other stream holds
    explicit SideStreamScope(pk_ctx* c) : ctx(c), main(c->stream), active(!c->prof.enabled) {
        if (!active) return;
        cudaEventRecord(ctx->ev_fork, main);
        cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
        ctx->stream = ctx->side;
    }
    ~SideStreamScope() {
        if (!active) return;
        cudaEventRecord(ctx->ev_join, ctx->side);
        ctx->stream = main;
        ctx->side_pending = true;
    }
};
static inline void side_join(pk_ctx* ctx) {
    if (!ctx->side_pending) return;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    ctx->side_pending = false;
}

static inline int ilog2(uint64_t x) { int l = 0; while ((uint64_t(1) << (l + 1)) <= x) ++l; return l; }

// host-side conversions between the ABI's u64[4] limbs and fr_t / fq_t
template <class F> static inline F host_load_canonical(const uint64_t* s) {
    F x;
    memcpy(x.v, s, 32);
    return x.to_mont();
}
template <class F> static inline void host_store_canonical(const F& x, uint64_t* d) {
    F c = x.from_mont();
    memcpy(d, c.v, 32);
}

}  // namespace pk
