// Radix-2 NTT over BN254 Fr for sm_100a.
//
// Replaces bellman_ce's Polynomial::{fft, ifft, coset_fft, icoset_fft_for_generator,
// bitreversed_lde_using_bitreversed_ntt} (call sites src/plonk.rs:104,140,152-159), which on
// the CPU are an in-place Cooley-Tukey radix-2 transform split over host threads.
//
// Device design: a size-2^L transform is cut into passes of up to 9 stages (see ntt_tile_log).  Each pass stages a tile of 2^k
// (strided) x C (contiguous) elements in shared memory — C >= 4 keeps every global access a >=128 B run of
// 128-bit loads/stores — runs its k butterfly stages there with one Montgomery multiplication per butterfly
// held in registers, and writes the tile back.  Forward transforms are decimation-in-frequency (natural ->
// bit-reversed), inverse ones decimation-in-time (bit-reversed -> natural), so the prover never needs a
// standalone bit-reversal pass: evaluations live in bit-reversed order between the two.  Coset shifts and the
// 1/n factor are fused into the first/last pass as an element-wise pre/post multiplication.
//
// Twiddles: one table w^e (e <= n_max/2) per context; smaller domains index it with a stride, and the inverse
// transform reads w^{-e} as -w^{n/2-e}, so no inverse table exists.
//
// HBM traffic per pass: 64 B per element (32 B in, 32 B out) + pre/post tables; a 2^20 transform is 3 passes.
#include "ntt.cuh"

namespace pk {

fr_t host_root_of_unity(int log_n) {
    fr_t g;
    for (int i = 0; i < 8; ++i) g.v[i] = FrRoots::root_2_28(i);
    for (int i = 0; i < 28 - log_n; ++i) g = g.sqr();
    return g;
}

// ---------------------------------------------------------------- table builders
__global__ void tw_build_kernel(fr_t* tw, fr_t root, size_t count) {
    size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e < count) st_fp(tw + e, root.pow_u64(e));
}

// out[s][j] = g_s^j  with g_s given per slot
__global__ void coset_scale_build_kernel(fr_t* out, fr_t g0, fr_t g1, fr_t g2, fr_t g3, size_t n) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const fr_t g = blockIdx.y == 0 ? g0 : blockIdx.y == 1 ? g1 : blockIdx.y == 2 ? g2 : g3;
    st_fp(out + blockIdx.y * n + j, g.pow_u64(j));
}
// out[k] = ginv^k * c
__global__ void geom_build_kernel(fr_t* out, fr_t ginv, fr_t c, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) st_fp(out + k, ginv.pow_u64(k) * c);
}
__global__ void fill_kernel(fr_t* out, fr_t c, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) st_fp(out + k, c);
}
__global__ void to_mont_kernel(fr_t* data, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) st_fp(data + k, ld_fp(data + k).to_mont());
}
__global__ void from_mont_kernel(const fr_t* src, fr_t* dst, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) st_fp(dst + k, ld_fp(src + k).from_mont());
}
__global__ void bitrev_kernel(const fr_t* src, fr_t* dst, int log_n) {  // blockIdx.y = row of a batch
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    size_t r = log_n ? (size_t)(__brev((unsigned)i) >> (32 - log_n)) : 0;
    const size_t row = (size_t)blockIdx.y << log_n;
    st_fp(dst + row + r, ld_fp(src + row + i));
}
// a[r][c] *= w^{(row0 + r) * c}  (w = primitive 2^log_total-th root, or its inverse): the four-step twiddle
__global__ void twiddle_rows_kernel(fr_t* a, size_t rows, size_t cols, fr_t w, size_t row0) {
    size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    const fr_t base = w.pow_u64(row0 + r);  // w^(row0 + r); then raised to c
    st_fp(a + r * cols + c, ld_fp(a + r * cols + c) * base.pow_u64(c));
}

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

void fr_to_mont(pk_ctx* ctx, fr_t* data, size_t n) {
    if (!n) return;
    to_mont_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(data, n);
    ctx->prof.kernel_launches++;
}
void fr_from_mont(pk_ctx* ctx, const fr_t* src, fr_t* dst, size_t n) {
    if (!n) return;
    from_mont_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(src, dst, n);
    ctx->prof.kernel_launches++;
}
void bitrev_permute(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, size_t rows) {
    size_t n = size_t(1) << log_n;
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "bitrev batch too large");
    bitrev_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)rows), 256, 0, ctx->stream>>>(src, dst, log_n);
    ctx->prof.kernel_launches++;
}
void ensure_twiddles(pk_ctx* ctx, int log_n) {
    PK_REQUIRE(log_n <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28 (Fr two-adicity)");
    if (!ctx->domains) ctx->domains = new DomainCache();
    DomainCache* dc = ctx->domains;
    if (dc->tw_log >= log_n && dc->tw.p) return;
    int lg = log_n < 12 ? 12 : log_n;
    size_t count = (size_t(1) << (lg - 1)) + 1;
    PK_CUDA(cudaStreamSynchronize(ctx->stream));  // nobody may still be reading the old table
    dc->tw.alloc(count);
    dc->tw_log = lg;
    tw_build_kernel<<<grid1d(count, 256), 256, 0, ctx->stream>>>(dc->tw.p, host_root_of_unity(lg), count);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- the pass kernel
struct NttPass {
    const fr_t* src;
    fr_t* dst;
    const fr_t* tw;
    const fr_t* pre;     // optional element-wise factor applied on load
    const fr_t* post;    // optional element-wise factor applied on store
    fr_t post_const;     // used when use_post_const
    int use_post_const;
    int L;               // log2 of the transform size
    int bl;              // lowest varying index bit of this pass's tile
    int k;               // stages in this pass
    int c_log;           // log2 of contiguous columns per tile (c_log <= bl)
    int tw_shift;        // tw_log - L
    size_t src_stride, dst_stride, pre_stride;
    // Fused all-to-all (sharded prover): the pass's results go straight into the peers' memory over NVLink instead of dst:
    // element g -> peer[g >> per_log] + (rank << per_log) + (g & (2^per_log - 1))
    int peer_mode, peer_rank, per_log;
    fr_t* peer[8];
};
struct PeerStore { int mode = 0, rank = 0, per_log = 0; fr_t* peer[8] = {}; };
__device__ __forceinline__ void ntt_store(const NttPass& p, fr_t* dst, size_t g, const fr_t& x) {
    if (p.peer_mode == 0) {
        st_fp(dst + g, x);
    } else {
        const size_t q = g >> p.per_log;
        st_fp(p.peer[q] + ((size_t)p.peer_rank << p.per_log) + (g & ((size_t(1) << p.per_log) - 1)), x);
    }
}

// (An index swizzle l ^ (bit3(l) ? 7 : 0), which removes the 1.6 M bank conflicts of the span-1/2/4 stages of the contiguous
// pass, was measured: 2^20 0.213 vs 0.204 ms, 2^24 3.74 vs 3.54 ms — its two extra ALU operations per access cost more
// issue slots than the conflicts cost shared-memory cycles.  Not used.)
__device__ __forceinline__ fr_t lds_fr(const uint4* s_lo, const uint4* s_hi, unsigned l) {
    uint4 a = s_lo[l], b = s_hi[l];
    fr_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4* s_lo, uint4* s_hi, unsigned l, const fr_t& x) {
    s_lo[l] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    s_hi[l] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

template <bool DIT> __global__ void __launch_bounds__(512) ntt_pass_kernel(NttPass p) {
    extern __shared__ uint4 sm[];
    const unsigned e_log = p.k + p.c_log;
    const unsigned E = 1u << e_log, half = E >> 1;
    uint4* s_lo = sm;
    uint4* s_hi = sm + E;
    const unsigned tid = threadIdx.x;
    const unsigned cmask = (1u << p.c_log) - 1;
    const fr_t* src = p.src + blockIdx.y * p.src_stride;
    fr_t* dst = p.dst + blockIdx.y * p.dst_stride;
    const fr_t* pre = p.pre ? p.pre + blockIdx.y * p.pre_stride : nullptr;

    const unsigned lo_groups_log = p.bl - p.c_log;
    const size_t blk = blockIdx.x;
    const size_t hi = blk >> lo_groups_log;
    const size_t lo0 = (blk & ((size_t(1) << lo_groups_log) - 1)) << p.c_log;
    const size_t base = (hi << (p.bl + p.k)) | lo0;

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned l = tid + r * half;
        if (l < E) {
            size_t g = base | ((size_t)(l >> p.c_log) << p.bl) | (l & cmask);
            fr_t x = ld_fp(src + g);
            if (pre) x = x * ldg_fp(pre + g);
            sts_fr(s_lo, s_hi, l, x);
        }
    }
    __syncthreads();

    const unsigned c = tid & cmask;
    const unsigned q = tid >> p.c_log;
    const size_t lo = lo0 | c;
    const size_t half_n = size_t(1) << (p.L - 1);
    for (int j = 0; j < p.k; ++j) {
        const int bitpos = DIT ? j : p.k - 1 - j;
        const unsigned lowmask = (1u << bitpos) - 1;
        const unsigned t0 = ((q >> bitpos) << (bitpos + 1)) | (q & lowmask);
        const unsigned l0 = (t0 << p.c_log) | c;
        const unsigned l1 = l0 | (1u << (bitpos + p.c_log));
        fr_t a = lds_fr(s_lo, s_hi, l0);
        fr_t b = lds_fr(s_lo, s_hi, l1);
        // exponent of the twiddle on the size-2^L domain
        const size_t low = ((size_t)(t0 & lowmask) << p.bl) | lo;
        if (!DIT) {
            const int s = p.L - p.bl - p.k + j;  // global stage
            const size_t e = low << s;
            fr_t u = a + b;
            fr_t d = a - b;
            if (e != 0) d = d * ldg_fp(p.tw + (e << p.tw_shift));
            sts_fr(s_lo, s_hi, l0, u);
            sts_fr(s_lo, s_hi, l1, d);
        } else {
            const int s = p.bl + j;
            const size_t e = low << (p.L - 1 - s);
            // w^{-e} = -w^{n/2 - e}:  t = b * w^{n/2-e};  a' = a - t, b' = a + t
            fr_t t = (e == 0) ? b.neg() : b * ldg_fp(p.tw + ((half_n - e) << p.tw_shift));
            sts_fr(s_lo, s_hi, l0, a - t);
            sts_fr(s_lo, s_hi, l1, a + t);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned l = tid + r * half;
        if (l < E) {
            size_t g = base | ((size_t)(l >> p.c_log) << p.bl) | (l & cmask);
            fr_t x = lds_fr(s_lo, s_hi, l);
            if (p.post) x = x * ldg_fp(p.post + g);
            else if (p.use_post_const) x = x * p.post_const;
            ntt_store(p, dst, g, x);
        }
    }
}

// ---------------------------------------------------------------- the same pass with TMA staging of the tile
// The coefficient tile comes in through the TMA unit: warp 0 arms an mbarrier with the tile's byte count and issues one
// bulk copy (cp.async.bulk ... mbarrier::complete_tx::bytes) per contiguous row of the tile — a single 16 KB copy for
// the contiguous pass, 2^k copies of C x 32 B for a strided pass — and the block waits on the barrier's phase instead of
// pushing every element through registers.  Bulk copies land the tile in its natural layout (32-byte elements), so the
// two 16-byte halves of element l are read in the order (l >> 2) & 1 selects: eight consecutive elements then cover all
// eight 16-byte bank groups (the split lo/hi arrays of the non-TMA kernel get the same effect by construction).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ fr_t lds_fr_nat(const uint4* sm, unsigned l) {
    const unsigned sw = (l >> 2) & 1;
    const uint4 t0 = sm[2 * l + sw], t1 = sm[2 * l + (sw ^ 1)];
    const uint4 a = sw ? t1 : t0, b = sw ? t0 : t1;
    fr_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr_nat(uint4* sm, unsigned l, const fr_t& x) {
    const unsigned sw = (l >> 2) & 1;
    const uint4 a = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]), b = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    sm[2 * l + sw] = sw ? b : a;
    sm[2 * l + (sw ^ 1)] = sw ? a : b;
}

template <bool DIT> __global__ void __launch_bounds__(512) ntt_pass_tma_kernel(NttPass p) {
    extern __shared__ __align__(128) uint4 sm[];
    const unsigned e_log = p.k + p.c_log;
    const unsigned E = 1u << e_log, half = E >> 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * E);
    const unsigned tid = threadIdx.x;
    const unsigned cmask = (1u << p.c_log) - 1;
    const fr_t* src = p.src + blockIdx.y * p.src_stride;
    fr_t* dst = p.dst + blockIdx.y * p.dst_stride;
    const fr_t* pre = p.pre ? p.pre + blockIdx.y * p.pre_stride : nullptr;

    const unsigned lo_groups_log = p.bl - p.c_log;
    const size_t blk = blockIdx.x;
    const size_t hi = blk >> lo_groups_log;
    const size_t lo0 = (blk & ((size_t(1) << lo_groups_log) - 1)) << p.c_log;
    const size_t base = (hi << (p.bl + p.k)) | lo0;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid < 32) {
        // rows of the tile that are contiguous in global memory: the whole tile when bl == c_log, else C elements
        const bool whole = p.bl == p.c_log;
        const unsigned nrows = whole ? 1u : (1u << p.k);
        const unsigned row_bytes = whole ? E * 32u : (32u << p.c_log);
        if (tid == 0) mbar_arrive_expect_tx(bar, E * 32u);
        __syncwarp();
        for (unsigned r = tid; r < nrows; r += 32)
            tma_load_1d(reinterpret_cast<uint8_t*>(sm) + (size_t)r * row_bytes, src + (base | ((size_t)r << p.bl)), row_bytes, bar);
    }
    mbar_wait(bar, 0);
    if (pre) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            unsigned l = tid + r * half;
            if (l < E) {
                size_t g = base | ((size_t)(l >> p.c_log) << p.bl) | (l & cmask);
                sts_fr_nat(sm, l, lds_fr_nat(sm, l) * ldg_fp(pre + g));
            }
        }
        __syncthreads();
    }

    const unsigned c = tid & cmask;
    const unsigned q = tid >> p.c_log;
    const size_t lo = lo0 | c;
    const size_t half_n = size_t(1) << (p.L - 1);
    for (int j = 0; j < p.k; ++j) {
        const int bitpos = DIT ? j : p.k - 1 - j;
        const unsigned lowmask = (1u << bitpos) - 1;
        const unsigned t0 = ((q >> bitpos) << (bitpos + 1)) | (q & lowmask);
        const unsigned l0 = (t0 << p.c_log) | c;
        const unsigned l1 = l0 | (1u << (bitpos + p.c_log));
        fr_t a = lds_fr_nat(sm, l0);
        fr_t b = lds_fr_nat(sm, l1);
        const size_t low = ((size_t)(t0 & lowmask) << p.bl) | lo;
        if (!DIT) {
            const int s = p.L - p.bl - p.k + j;
            const size_t e = low << s;
            fr_t u = a + b;
            fr_t d = a - b;
            if (e != 0) d = d * ldg_fp(p.tw + (e << p.tw_shift));
            sts_fr_nat(sm, l0, u);
            sts_fr_nat(sm, l1, d);
        } else {
            const int s = p.bl + j;
            const size_t e = low << (p.L - 1 - s);
            fr_t t = (e == 0) ? b.neg() : b * ldg_fp(p.tw + ((half_n - e) << p.tw_shift));
            sts_fr_nat(sm, l0, a - t);
            sts_fr_nat(sm, l1, a + t);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned l = tid + r * half;
        if (l < E) {
            size_t g = base | ((size_t)(l >> p.c_log) << p.bl) | (l & cmask);
            fr_t x = lds_fr_nat(sm, l);
            if (p.post) x = x * ldg_fp(p.post + g);
            else if (p.use_post_const) x = x * p.post_const;
            ntt_store(p, dst, g, x);
        }
    }
}
// PK_NTT_TMA = 1 / 0 selects the TMA-staged pass kernel (tiles of >= 16 elements only: bulk copies move >= 16 B rows and
// the kernel needs a full warp to issue them)
static int ntt_use_tma() {
    static const int v = [] { const char* e = getenv("PK_NTT_TMA"); return e ? atoi(e) : 0; }();
    return v;
}

// log2 of the shared-memory tile (elements); 2^(tile - 1) threads per block.  Measured on B200 (2^20 / 2^24 transform):
// tile 2^10 (512 threads, 2 blocks per SM) 0.237 / 4.20 ms, 2^9 0.206 / 3.54 ms, 2^8 0.198 / 3.50 ms — smaller blocks let the
// load, butterfly and store phases of different blocks overlap on an SM.  2^9 keeps >= 64-byte runs in the strided passes.
// PK_NTT_TILE_LOG overrides (8..10).
static int ntt_tile_log() {
    static const int v = [] { const char* e = getenv("PK_NTT_TILE_LOG"); int t = e ? atoi(e) : 9; return t < 8 ? 8 : (t > 10 ? 10 : t); }();
    return v;
}
#define NTT_TILE_LOG ntt_tile_log()
// pass plan: one contiguous pass of up to NTT_TILE_LOG stages + strided passes of up to 8 stages
struct PassPlan { int n; int k[8]; int bl[8]; int c_log[8]; };
static PassPlan plan_passes(int L, bool dit) {
    PassPlan pl;
    pl.n = 0;
    int kc = L < NTT_TILE_LOG ? L : NTT_TILE_LOG;
    int rem = L - kc;
    int ns = (rem + 7) / 8;
    int ks[8];
    for (int i = 0; i < ns; ++i) ks[i] = rem / ns + (i < rem % ns ? 1 : 0);
    if (dit) {
        // stages from bit 0 upward: contiguous first
        pl.k[0] = kc; pl.bl[0] = 0; pl.c_log[0] = 0; pl.n = 1;
        int s0 = kc;
        for (int i = 0; i < ns; ++i) {
            pl.k[pl.n] = ks[i]; pl.bl[pl.n] = s0;
            int cl = NTT_TILE_LOG - ks[i]; if (cl > s0) cl = s0;
            pl.c_log[pl.n] = cl;
            s0 += ks[i]; pl.n++;
        }
    } else {
        // stages from the top bit downward: strided passes first, contiguous last
        int s0 = 0;
        for (int i = 0; i < ns; ++i) {
            int bl = L - s0 - ks[i];
            pl.k[pl.n] = ks[i]; pl.bl[pl.n] = bl;
            int cl = NTT_TILE_LOG - ks[i]; if (cl > bl) cl = bl;
            pl.c_log[pl.n] = cl;
            s0 += ks[i]; pl.n++;
        }
        pl.k[pl.n] = kc; pl.bl[pl.n] = 0; pl.c_log[pl.n] = 0; pl.n++;
    }
    return pl;
}

// L_tw > L (inverse only): the 2^L elements are one aligned block of a size-2^L_tw transform in bit-reversed order and
// only the block-local stages (bits 0 .. L-1, twiddles of the big domain) are run — the first part of a transform
// whose upper stages cross GPUs (sharded prover).
template <bool DIT>
static void run_passes(pk_ctx* ctx, const fr_t* src, fr_t* dst, int L, const fr_t* pre, const fr_t* post, const fr_t* post_const,
                       int batch, size_t src_stride, size_t dst_stride, size_t pre_stride, int L_tw = 0,
                       const PeerStore* peers = nullptr) {
    PK_REQUIRE(L >= 0 && L <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28");
    size_t n = size_t(1) << L;
    if (L == 0) {  // size-1 transform is the identity (all scale factors are 1)
        for (int b = 0; b < batch; ++b)
            PK_CUDA(cudaMemcpyAsync(dst + b * dst_stride, src + b * src_stride, sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    if (L_tw < L) L_tw = L;
    PK_REQUIRE(L_tw == L || DIT, PK_ERR_INVALID, "block-local stages exist for the inverse transform only");
    ensure_twiddles(ctx, L_tw);
    DomainCache* dc = ctx->domains;
    PassPlan pl = plan_passes(L, DIT);
    ScopedKernelTimer timer(ctx, 1, (uint64_t)n * batch * pl.n);
    for (int i = 0; i < pl.n; ++i) {
        NttPass p;
        memset(&p, 0, sizeof(p));
        bool first = i == 0, last = i == pl.n - 1;
        p.src = first ? src : dst;
        p.dst = dst;
        p.src_stride = first ? src_stride : dst_stride;
        p.dst_stride = dst_stride;
        p.tw = dc->tw.p;
        p.tw_shift = dc->tw_log - L_tw;
        p.pre = first ? pre : nullptr;
        p.pre_stride = pre_stride;
        p.post = last ? post : nullptr;
        if (last && !post && post_const) { p.use_post_const = 1; p.post_const = *post_const; }
        p.L = L_tw; p.bl = pl.bl[i]; p.k = pl.k[i]; p.c_log = pl.c_log[i];
        if (last && peers && peers->mode) {  // the LAST pass writes into the peers' memory
            p.peer_mode = peers->mode; p.peer_rank = peers->rank; p.per_log = peers->per_log;
            for (int q = 0; q < 8; ++q) p.peer[q] = peers->peer[q];
        }
        unsigned E = 1u << (p.k + p.c_log);
        unsigned threads = E >> 1; if (threads < 1) threads = 1;
        dim3 grid((unsigned)(n / E), batch);
        size_t smem = (size_t)E * 32;
        if (ntt_use_tma() && threads >= 32 && (p.src != p.dst || true))
            ntt_pass_tma_kernel<DIT><<<grid, threads, smem + 16, ctx->stream>>>(p);
        else
            ntt_pass_kernel<DIT><<<grid, threads, smem, ctx->stream>>>(p);
        ctx->prof.kernel_launches++;
        ctx->prof.ntt_launches++;
    }
    PK_CUDA(cudaGetLastError());
}

void ntt_forward_bitrev(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, const fr_t* pre, int batch, size_t src_stride,
                        size_t dst_stride, size_t pre_stride) {
    run_passes<false>(ctx, src, dst, log_n, pre, nullptr, nullptr, batch, src_stride, dst_stride, pre_stride);
}
void ntt_inverse_from_bitrev(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, const fr_t* post) {
    fr_t ninv = fr_t::from_u32(2).inverse().pow_u64(log_n);  // 1/n
    run_passes<true>(ctx, src, dst, log_n, nullptr, post, post ? nullptr : &ninv, 1, 0, 0, 0);
}

CosetTables* get_coset_tables(pk_ctx* ctx, int log_n) {
    PK_REQUIRE(log_n + 2 <= 28, PK_ERR_DEGREE_TOO_LARGE, "4n coset domain larger than 2^28");
    ensure_twiddles(ctx, log_n + 2);
    DomainCache* dc = ctx->domains;
    auto it = dc->coset.find(log_n);
    if (it != dc->coset.end()) return it->second;
    CosetTables* ct = new CosetTables();
    ct->log_n = log_n;
    size_t n = size_t(1) << log_n;
    ct->scale4.alloc(4 * n);
    ct->iscale4n.alloc(4 * n);
    ct->l0.alloc(4 * n);
    fr_t g7, g7inv;
    for (int i = 0; i < 8; ++i) { g7.v[i] = FrRoots::gen7(i); g7inv.v[i] = FrRoots::gen7_inv(i); }
    fr_t w4 = host_root_of_unity(log_n + 2);
    fr_t gs[4];
    static const int brev2[4] = {0, 2, 1, 3};
    for (int s = 0; s < 4; ++s) gs[s] = g7 * w4.pow_u64(brev2[s]);
    coset_scale_build_kernel<<<dim3((unsigned)((n + 255) / 256), 4), 256, 0, ctx->stream>>>(ct->scale4.p, gs[0], gs[1], gs[2], gs[3], n);
    fr_t inv4n = fr_t::from_u32(2).inverse().pow_u64(log_n + 2);
    geom_build_kernel<<<grid1d(4 * n, 256), 256, 0, ctx->stream>>>(ct->iscale4n.p, g7inv, inv4n, 4 * n);
    ctx->prof.kernel_launches += 2;
    dc->coset[log_n] = ct;
    // L_0(X) = (1/N) sum_j X^j : LDE of the all-(1/N) coefficient vector
    DevBuf<fr_t> tmp(n);
    fr_t ninv = fr_t::from_u32(2).inverse().pow_u64(log_n);
    fill_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(tmp.p, ninv, n);
    ctx->prof.kernel_launches++;
    lde4_slots(ctx, tmp.p, ct->l0.p, log_n);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ct;
}

void lde4_slots(pk_ctx* ctx, const fr_t* coeffs, fr_t* out4n, int log_n) {
    CosetTables* ct = get_coset_tables(ctx, log_n);
    size_t n = size_t(1) << log_n;
    ntt_forward_bitrev(ctx, coeffs, out4n, log_n, ct->scale4.p, 4, 0, n, n);
}
void icoset4n_from_slots(pk_ctx* ctx, const fr_t* vals4n, fr_t* coeffs4n, int log_n) {
    CosetTables* ct = get_coset_tables(ctx, log_n);
    ntt_inverse_from_bitrev(ctx, vals4n, coeffs4n, log_n + 2, ct->iscale4n.p);
}

void ntt_inverse_local_stages(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_block, int log_total) {
    if (log_block == 0) {
        if (src != dst) PK_CUDA(cudaMemcpyAsync(dst, src, sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    run_passes<true>(ctx, src, dst, log_block, nullptr, nullptr, nullptr, 1, 0, 0, 0, log_total);
}
// the same stages with the all-to-all FUSED into the last pass: element k of this rank's block is stored straight into
// rank (k >> per_log)'s receive buffer at [rank][k mod 2^per_log] (peer[q] = that buffer as mapped into this process)
void ntt_inverse_local_stages_scatter(pk_ctx* ctx, const fr_t* src, fr_t* scratch, int log_block, int log_total, fr_t* const* peer,
                                      int world, int rank, int per_log) {
    PK_REQUIRE(log_block >= 1 && world <= 8, PK_ERR_INVALID, "bad fused all-to-all geometry");
    PeerStore ps;
    ps.mode = 1; ps.rank = rank; ps.per_log = per_log;
    for (int q = 0; q < world; ++q) ps.peer[q] = peer[q];
    run_passes<true>(ctx, src, scratch, log_block, nullptr, nullptr, nullptr, 1, 0, 0, 0, log_total, &ps);
}
// ---------------------------------------------------------------- pieces of the coset-sharded quotient (sharded prover)
// b[i] = c^i * sum_{u < F} a[i + u * n/F] * kappa^u  (i < n/F): the polynomial a (n coefficients) restricted to the coset
// c * H_{n/F}, on which X^(n/F) is the constant kappa = c^(n/F); cpow[i] = c^i.  A forward NTT of size n/F finishes the
// evaluation.  F = 1 is a plain pre-scale.
__global__ void coset_fold_kernel(const fr_t* a, const fr_t* cpow, fr_t kappa, int F, size_t nf, fr_t* b) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nf) return;
    fr_t acc = ld_fp(a + (size_t)(F - 1) * nf + i);
    for (int u = F - 2; u >= 0; --u) acc = acc * kappa + ld_fp(a + (size_t)u * nf + i);
    st_fp(b + i, acc * ldg_fp(cpow + i));
}
void coset_fold(pk_ctx* ctx, const fr_t* a, const fr_t* cpow, const fr_t& kappa, int F, size_t nf, fr_t* b) {
    coset_fold_kernel<<<grid1d(nf, 256), 256, 0, ctx->stream>>>(a, cpow, kappa, F, nf, b);
    ctx->prof.kernel_launches++;
}
// out[i] = a[i] * w^i with w the primitive 2^log_n-th root: the coefficients of a(w X)
__global__ void omega_scale_kernel(const fr_t* a, fr_t* out, const fr_t* tw, int tw_shift, int log_n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    const size_t half = size_t(1) << (log_n - 1);
    fr_t w = ldg_fp(tw + ((i & (half - 1)) << tw_shift));
    if (i & half) w = w.neg();
    st_fp(out + i, ld_fp(a + i) * w);
}
void omega_scale(pk_ctx* ctx, const fr_t* a, fr_t* out, int log_n) {
    PK_REQUIRE(log_n >= 1, PK_ERR_INVALID, "omega_scale needs a domain of at least 2");
    ensure_twiddles(ctx, log_n);
    DomainCache* dc = ctx->domains;
    omega_scale_kernel<<<grid1d(size_t(1) << log_n, 256), 256, 0, ctx->stream>>>(a, out, dc->tw.p, dc->tw_log - log_n, log_n);
    ctx->prof.kernel_launches++;
}
// The upper log2(G) stages of a size-2^L inverse transform (decimation in time) whose lower stages ran block-locally on
// G ranks: in[c][k'] = element k0 + k' of rank c's block after its local stages (what the all-to-all delivers), block
// length m = 2^L / G.  One thread owns the G values of one k, runs the stages in registers and writes
// out[c][k'] = coefficient (c * m + k0 + k') times g7inv^index / 2^L  (kscale[k'] = g7inv^(k0 + k') / 2^L, cscale[c] = g7inv^(c m)).
struct CrossArgs { fr_t cscale[8]; };
template <int G> __global__ void __launch_bounds__(128) ntt_cross_kernel(const fr_t* in, fr_t* out, const fr_t* kscale, CrossArgs ca,
                                                                         const fr_t* tw, int tw_shift, int L, size_t m, size_t k0,
                                                                         size_t per) {
    size_t kk = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (kk >= per) return;
    fr_t v[G];
#pragma unroll
    for (int c = 0; c < G; ++c) v[c] = ld_fp(in + (size_t)c * per + kk);
    const size_t k = k0 + kk;
    const size_t half_n = size_t(1) << (L - 1);
    int log_m = 0;
    while ((size_t(1) << log_m) < m) ++log_m;
#pragma unroll
    for (int u = 0; (1 << u) < G; ++u) {
        const int s = log_m + u;
#pragma unroll
        for (int c = 0; c < G; ++c) {
            if (c & (1 << u)) continue;
            const size_t low = k + (size_t)(c & ((1 << u) - 1)) * m;
            const size_t e = low << (L - 1 - s);
            const fr_t a = v[c], b = v[c | (1 << u)];
            const fr_t t = (e == 0) ? b.neg() : b * ldg_fp(tw + ((half_n - e) << tw_shift));
            v[c] = a - t;
            v[c | (1 << u)] = a + t;
        }
    }
    const fr_t ks = ldg_fp(kscale + kk);
#pragma unroll
    for (int c = 0; c < G; ++c) st_fp(out + (size_t)c * per + kk, v[c] * ks * ca.cscale[c]);
}
void ntt_inverse_cross_stages(pk_ctx* ctx, const fr_t* in, fr_t* out, const fr_t* kscale, const fr_t* cscale, int G, int log_total,
                              size_t k0) {
    PK_REQUIRE(G == 1 || G == 2 || G == 4 || G == 8, PK_ERR_INVALID, "the sharded prover runs on 1, 2, 4 or 8 ranks");
    ensure_twiddles(ctx, log_total);
    DomainCache* dc = ctx->domains;
    const size_t m = (size_t(1) << log_total) / G, per = m / G;
    CrossArgs ca;
    for (int c = 0; c < 8; ++c) ca.cscale[c] = c < G ? cscale[c] : fr_t::one();
    const int ts = dc->tw_log - log_total;
    dim3 grid = grid1d(per, 128);
    if (G == 1) ntt_cross_kernel<1><<<grid, 128, 0, ctx->stream>>>(in, out, kscale, ca, dc->tw.p, ts, log_total, m, k0, per);
    else if (G == 2) ntt_cross_kernel<2><<<grid, 128, 0, ctx->stream>>>(in, out, kscale, ca, dc->tw.p, ts, log_total, m, k0, per);
    else if (G == 4) ntt_cross_kernel<4><<<grid, 128, 0, ctx->stream>>>(in, out, kscale, ca, dc->tw.p, ts, log_total, m, k0, per);
    else ntt_cross_kernel<8><<<grid, 128, 0, ctx->stream>>>(in, out, kscale, ca, dc->tw.p, ts, log_total, m, k0, per);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}

void ntt_rows_natural(pk_ctx* ctx, fr_t* data, fr_t* tmp, int log_len, size_t rows, bool inverse) {
    const size_t len = size_t(1) << log_len;
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "row batch too large");
    if (!inverse) {
        ntt_forward_bitrev(ctx, data, tmp, log_len, nullptr, (int)rows, len, len, 0);
        bitrev_permute(ctx, tmp, data, log_len, rows);
    } else {
        bitrev_permute(ctx, data, tmp, log_len, rows);
        fr_t ninv = fr_t::from_u32(2).inverse().pow_u64(log_len);
        run_passes<true>(ctx, tmp, data, log_len, nullptr, nullptr, &ninv, (int)rows, len, len, 0);
    }
}
void twiddle_rows(pk_ctx* ctx, fr_t* a, size_t rows, size_t cols, int log_total, size_t row0, bool inverse) {
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "row batch too large");
    fr_t w = host_root_of_unity(log_total);
    if (inverse) w = w.inverse();
    twiddle_rows_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)rows), 256, 0, ctx->stream>>>(a, rows, cols, w, row0);
    ctx->prof.kernel_launches++;
}


}  // namespace pk
