// Multi-scalar multiplication over BN254 G1 for sm_100a (bucket method).
//
// Replaces bellman_ce's multiexp::dense_multiexp as reached through kate_commitment::commit_using_monomials
// (11 calls per proof under src/plonk.rs:140,152-159 and 11 under make_verification_key,
// src/plonk.rs:122-124).  The CPU algorithm walks c-bit windows, keeps 2^c-1 Jacobian buckets per thread and
// combines windows by c doublings.  The device design differs on purpose:
//
//   * FIXED-BASE WINDOW TABLES.  The SRS is loaded once and never changes, so at load time every base gets its
//     W = ceil(255/c) multiples 2^(c*w) * P_i precomputed in affine form (W x N x 64 B; 832 MB at N = 2^20, c = 20 —
//     HBM is 180 GB).  All windows then share ONE bucket space of 2^(c-1) signed-digit buckets: no per-window
//     bucket sets and no window-combine doubling chain at all.
//   * kernels (each its own launch, as named in BASELINE.json's north star):
//       window scan      msm_digits_hist_kernel   scalar -> signed base-2^c digits, bucket histogram
//       (offsets)        msm_offsets_kernel       exclusive scan of the histogram
//       scatter          msm_scatter_kernel       counting sort of (table index, sign) by bucket
//       bucket accum     msm_accum_kernel<true>   mixed XYZZ additions; load-balanced segmented reduction:
//                                                 every thread owns a fixed-length chunk of the sorted list,
//                                                 whole runs go straight to their bucket, runs cut by a chunk
//                                                 edge become partial entries for the next level
//                        msm_accum_kernel<false>  same over partial entries (full XYZZ adds) until one chunk remains
//       bucket reduce    msm_bucket_reduce_kernel sum (b+1) * B_b by segmented running sums
//       final fold       msm_final_reduce_kernel  -> one XYZZ point; the host normalises it to affine (1 inversion)
//   * work is independent of the scalar distribution (zero digits are skipped; repeated digits cannot unbalance it).
//
// Algorithmic HBM bytes: 96 B per (scalar, base) pair.  The kernels are bound by 32-bit integer
// multiply throughput, not HBM: one mixed addition is 10 Montgomery products of ~130 IMAD.WIDE each.
#include "msm.cuh"

namespace pk {

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

// ---------------------------------------------------------------- SRS upload + window tables
__global__ void bases_to_mont_kernel(g1_affine_t* pts, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p = ldg_affine(pts + i);
    if (p.is_inf()) return;
    p.x = p.x.to_mont();
    p.y = p.y.to_mont();
    st_affine(pts + i, p);
}
// table[w][i] = 2^c * table[w-1][i]
__global__ void srs_window_kernel(g1_affine_t* table, size_t n, int c, int W) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p = ldg_affine(table + i);
    for (int w = 1; w < W; ++w) {
        g1_xyzz_t a = g1_xyzz_t::dbl_affine(p);
        for (int k = 1; k < c; ++k) a = a.dbl();
        p = a.to_affine();
        st_affine(table + (size_t)w * n + i, p);
    }
}

static int pick_window_bits(uint64_t n) {
    int best = 8;
    double best_cost = 1e300;
    for (int c = 8; c <= 20; ++c) {
        int W = (255 + c - 1) / c;
        double table_bytes = (double)W * (double)n * 64.0;
        if (table_bytes > 64e9) continue;
        double cost = (double)n * W * 10.0 + (double)(1u << (c - 1)) * 40.0;
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

void srs_load(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits) {
    PK_REQUIRE(n >= 1, PK_ERR_INVALID, "empty SRS");
    PK_REQUIRE(n <= (uint64_t(1) << 26), PK_ERR_DEGREE_TOO_LARGE, "SRS larger than 2^26 bases (SETUP_MAX_POW2, src/plonk.rs:27)");
    int c = window_bits ? window_bits : pick_window_bits(n);
    PK_REQUIRE(c >= 2 && c <= 24, PK_ERR_INVALID, "window_bits out of range");
    int W = (255 + c - 1) / c;
    PK_REQUIRE((uint64_t)W * n < (uint64_t(1) << 31), PK_ERR_DEGREE_TOO_LARGE, "window table index does not fit 31 bits");
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    delete ctx->srs;
    ctx->srs = nullptr;
    SrsTables* s = new SrsTables();
    ctx->srs = s;
    s->n = n; s->c = c; s->W = W; s->B = 1u << (c - 1);
    s->table.alloc((size_t)W * n);
    PK_CUDA(cudaMemcpyAsync(s->table.p, bases_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    bases_to_mont_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(s->table.p, n);
    srs_window_kernel<<<grid1d(n, 128), 128, 0, ctx->stream>>>(s->table.p, n, c, W);
    ctx->prof.kernel_launches += 2;
    PK_CUDA(cudaGetLastError());
    size_t M = (size_t)n * W;
    s->hist.alloc(s->B + 1);
    s->offsets.alloc(s->B + 1);
    s->cursor.alloc(s->B);
    s->keys.alloc(M);
    s->items.alloc(M);
    s->buckets.alloc(s->B);
    // level-1 chunk: keep the chunk count <= 2^19 so the partial lists stay small
    uint32_t chunk = 64;
    while (M / chunk > (size_t(1) << 19)) chunk *= 2;
    s->chunk1 = chunk;
    size_t nchunks = (M + chunk - 1) / chunk;
    for (int k = 0; k < 2; ++k) {
        s->pkeys[k].alloc(2 * nchunks + 2);
        s->ppts[k].alloc(2 * nchunks + 2);
    }
    s->counts.alloc(16);
    s->red.alloc(4100);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---------------------------------------------------------------- window scan: signed digits
// digit w of canonical scalar k (8 x 32-bit LE limbs), before carry handling
__device__ __forceinline__ uint32_t raw_window(const uint32_t* k, int w, int c) {
    int pos = w * c;
    if (pos >= 256) return 0;
    int limb = pos >> 5, off = pos & 31;
    uint32_t v = k[limb] >> off;
    if (off && limb < 7) v |= k[limb + 1] << (32 - off);
    return v & ((1u << c) - 1);
}

// MODE 0: histogram; MODE 1: scatter
template <int MODE>
__global__ void msm_digits_kernel(const fr_t* scalars, uint32_t n, uint32_t table_n, uint32_t base_offset, int c, int W,
                                  uint32_t* hist_or_cursor, uint32_t* keys, uint32_t* items) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t k = ld_fp(scalars + i).from_mont();
    if (k.is_zero()) return;
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (int w = 0; w < W; ++w) {
        uint32_t d = raw_window(k.v, w, c) + carry;
        uint32_t neg = 0;
        if (d > half) { d = (1u << c) - d; neg = 1; carry = 1; } else carry = 0;
        if (d == 0) continue;
        uint32_t b = d - 1;
        if (MODE == 0) {
            atomicAdd(hist_or_cursor + b, 1u);
        } else {
            uint32_t pos = atomicAdd(hist_or_cursor + b, 1u);
            keys[pos] = b;
            items[pos] = ((uint32_t)w * table_n + base_offset + i) | (neg << 31);
        }
    }
}

// exclusive scan of hist[0..B) into offsets[0..B], cursor = copy; total -> counts[0].  Single block of 1024 threads.
__global__ void msm_offsets_kernel(const uint32_t* hist, uint32_t* offsets, uint32_t* cursor, uint32_t B, uint32_t* counts) {
    __shared__ uint32_t sh[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (B + 1023) / 1024;
    uint32_t lo = tid * per, hi = lo + per;
    if (hi > B) hi = B;
    uint32_t sum = 0;
    for (uint32_t b = lo; b < hi; ++b) sum += hist[b];
    sh[tid] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t v = tid >= d ? sh[tid - d] : 0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    uint32_t run = sh[tid] - sum;  // exclusive prefix of this thread's range
    for (uint32_t b = lo; b < hi; ++b) {
        offsets[b] = run;
        cursor[b] = run;
        run += hist[b];
    }
    if (tid == 1023) {
        offsets[B] = sh[1023];
        counts[0] = sh[1023];
    }
}

// ---------------------------------------------------------------- bucket accumulation (segmented reduction over the sorted list)
struct AccumParams {
    const uint32_t* keys;       // sorted bucket ids of this level's entries
    const uint32_t* items;      // level 1: table index | sign << 31
    const g1_xyzz_t* pts;       // level >= 2: partial sums
    const g1_affine_t* table;
    const uint32_t* count_in;   // number of entries of this level (device)
    uint32_t* count_out;        // number of entries of the next level
    uint32_t chunk;             // entries per thread
    g1_xyzz_t* buckets;
    uint32_t* out_keys;         // [2 * chunks]
    g1_xyzz_t* out_pts;
};

template <bool LEVEL1> __global__ void __launch_bounds__(128) msm_accum_kernel(AccumParams p) {
    const uint32_t count = *p.count_in;
    const uint32_t nchunks = (count + p.chunk - 1) / p.chunk;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *p.count_out = nchunks > 1 ? 2 * nchunks : 0;
    if (t >= nchunks) return;
    const uint32_t s = t * p.chunk;
    const uint32_t e = (s + p.chunk < count) ? s + p.chunk : count;
    const bool head_partial = s > 0 && p.keys[s - 1] == p.keys[s];
    const bool tail_partial = e < count && p.keys[e] == p.keys[e - 1];
    const uint32_t first_key = p.keys[s];
    uint32_t cur = first_key;
    bool in_head = true;
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    const g1_xyzz_t inf = g1_xyzz_t::infinity();
    for (uint32_t i = s; i < e; ++i) {
        const uint32_t k = p.keys[i];
        if (k != cur) {
            if (in_head && head_partial) st_xyzz(p.out_pts + 2 * (size_t)t, acc);
            else if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            in_head = false;
            acc = inf;
            cur = k;
        }
        if (LEVEL1) {
            const uint32_t it = p.items[i];
            g1_affine_t pt = ldg_affine(p.table + (it & 0x7fffffffu));
            if (it >> 31) pt.y = pt.y.neg();
            acc = acc.add_mixed(pt);
        } else {
            acc = acc.add(ld_xyzz(p.pts + i));
        }
    }
    // the last run: the tail run, or the only run of the chunk
    if (in_head) {
        const bool partial = head_partial || tail_partial;
        if (partial) st_xyzz(p.out_pts + 2 * (size_t)t, acc);
        else {
            if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            st_xyzz(p.out_pts + 2 * (size_t)t, inf);
        }
        st_xyzz(p.out_pts + 2 * (size_t)t + 1, inf);
        p.out_keys[2 * (size_t)t] = cur;
        p.out_keys[2 * (size_t)t + 1] = cur;
    } else {
        if (!head_partial) st_xyzz(p.out_pts + 2 * (size_t)t, inf);
        if (tail_partial) st_xyzz(p.out_pts + 2 * (size_t)t + 1, acc);
        else {
            if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            st_xyzz(p.out_pts + 2 * (size_t)t + 1, inf);
        }
        p.out_keys[2 * (size_t)t] = first_key;
        p.out_keys[2 * (size_t)t + 1] = cur;
    }
}

// ---------------------------------------------------------------- bucket reduction: sum_b (b + 1) * B_b
__device__ __forceinline__ void block_sum_xyzz(g1_xyzz_t& v, g1_xyzz_t* sh, g1_xyzz_t* out) {
    const unsigned tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (unsigned d = blockDim.x >> 1; d > 0; d >>= 1) {
        if (tid < d) sh[tid] = sh[tid].add(sh[tid + d]);
        __syncthreads();
    }
    if (tid == 0) st_xyzz(out, sh[0]);
}

__global__ void __launch_bounds__(128) msm_bucket_reduce_kernel(const g1_xyzz_t* buckets, uint32_t B, uint32_t seg, g1_xyzz_t* block_out) {
    __shared__ g1_xyzz_t sh[128];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = (uint64_t)t * seg;
    g1_xyzz_t R = g1_xyzz_t::infinity();
    if (lo < B) {
        uint32_t hi = (lo + seg < B) ? (uint32_t)(lo + seg) : B;
        g1_xyzz_t running = g1_xyzz_t::infinity(), sum = g1_xyzz_t::infinity();
        for (uint32_t b = hi; b-- > (uint32_t)lo;) {
            running = running.add(ld_xyzz(buckets + b));
            sum = sum.add(running);
        }
        // sum = sum_b (b - lo + 1) B_b ; add lo * running to get weights (b + 1)
        R = sum.add(running.mul_small((uint32_t)lo));
    }
    block_sum_xyzz(R, sh, block_out + blockIdx.x);
}
__global__ void __launch_bounds__(128) msm_final_reduce_kernel(const g1_xyzz_t* in, uint32_t n, g1_xyzz_t* out) {
    __shared__ g1_xyzz_t sh[128];
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc = acc.add(ld_xyzz(in + i));
    block_sum_xyzz(acc, sh, out);
}

void affine_to_abi(const g1_affine_t& p, uint64_t out[8]) {
    if (p.is_inf()) { memset(out, 0, 64); return; }
    fq_t x = p.x.from_mont(), y = p.y.from_mont();
    memcpy(out, x.v, 32);
    memcpy(out + 4, y.v, 32);
}

g1_affine_t msm_run(pk_ctx* ctx, const fr_t* scalars, uint64_t n, uint64_t base_offset) {
    SrsTables* s = ctx->srs;
    PK_REQUIRE(s != nullptr, PK_ERR_DEGREE_TOO_LARGE, "no SRS loaded (pk_srs_load_g1)");
    PK_REQUIRE(base_offset + n <= s->n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    if (n == 0) return g1_affine_t::infinity();
    cudaStream_t st = ctx->stream;
    const uint32_t B = s->B;
    PK_CUDA(cudaMemsetAsync(s->hist.p, 0, (B + 1) * sizeof(uint32_t), st));
    PK_CUDA(cudaMemsetAsync(s->buckets.p, 0, (size_t)B * sizeof(g1_xyzz_t), st));
    msm_digits_kernel<0><<<grid1d(n, 256), 256, 0, st>>>(scalars, (uint32_t)n, (uint32_t)s->n, (uint32_t)base_offset, s->c, s->W,
                                                         s->hist.p, nullptr, nullptr);
    msm_offsets_kernel<<<1, 1024, 0, st>>>(s->hist.p, s->offsets.p, s->cursor.p, B, s->counts.p);
    msm_digits_kernel<1><<<grid1d(n, 256), 256, 0, st>>>(scalars, (uint32_t)n, (uint32_t)s->n, (uint32_t)base_offset, s->c, s->W,
                                                         s->cursor.p, s->keys.p, s->items.p);
    ctx->prof.kernel_launches += 3;
    // accumulation levels (worst-case grids; the device-side counts bound the real work)
    size_t max_entries = (size_t)n * s->W;
    AccumParams p;
    memset(&p, 0, sizeof(p));
    p.table = s->table.p;
    p.buckets = s->buckets.p;
    int level = 0;
    uint32_t chunk = s->chunk1;
    const uint32_t* keys = s->keys.p;
    while (true) {
        size_t nchunks = (max_entries + chunk - 1) / chunk;
        p.keys = keys;
        p.items = s->items.p;
        p.pts = level == 0 ? nullptr : s->ppts[(level - 1) & 1].p;
        p.count_in = s->counts.p + level;
        p.count_out = s->counts.p + level + 1;
        p.chunk = chunk;
        p.out_keys = s->pkeys[level & 1].p;
        p.out_pts = s->ppts[level & 1].p;
        if (level == 0) {
            ScopedKernelTimer timer(ctx, 0, n);
            msm_accum_kernel<true><<<grid1d(nchunks, 128), 128, 0, st>>>(p);
            ctx->prof.msm_accum_launches++;
        } else {
            msm_accum_kernel<false><<<grid1d(nchunks, 128), 128, 0, st>>>(p);
        }
        ctx->prof.kernel_launches++;
        if (nchunks <= 1) break;
        keys = s->pkeys[level & 1].p;
        max_entries = 2 * nchunks;
        chunk = 16;
        ++level;
        PK_REQUIRE(level < 14, PK_ERR_INVALID, "MSM level overflow");
    }
    // bucket reduction
    uint32_t seg = 32;
    while ((B + seg - 1) / seg > 128u * 4096u) seg *= 2;
    uint32_t nthreads = (B + seg - 1) / seg;
    uint32_t nblocks = (nthreads + 127) / 128;
    msm_bucket_reduce_kernel<<<nblocks, 128, 0, st>>>(s->buckets.p, B, seg, s->red.p);
    msm_final_reduce_kernel<<<1, 128, 0, st>>>(s->red.p, nblocks, s->red.p + 4096);
    ctx->prof.kernel_launches += 2;
    PK_CUDA(cudaGetLastError());
    g1_xyzz_t* host_pt = reinterpret_cast<g1_xyzz_t*>(ctx->pinned);
    PK_CUDA(cudaMemcpyAsync(host_pt, s->red.p + 4096, sizeof(g1_xyzz_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    return host_pt->to_affine();
}

}  // namespace pk
