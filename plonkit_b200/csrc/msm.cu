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
//   * BATCHES.  Commitments that are independent inside a prover round (4 wires, 4 quotient chunks, 2 openings) go
//     through the kernels together: bucket ids are (set, bucket), so one sort, one accumulation and one reduction
//     serve up to 4 scalar sets and the latency-bound tail kernels are paid once per round instead of once per MSM.
//   * kernels (each its own launch, as named in BASELINE.json's north star):
//       window scan      msm_coarse_kernel<false> scalar -> signed base-2^c digits, coarse-bin histogram (shared memory)
//       offsets          u32_scan_*               exclusive scan of the coarse histogram
//       scatter          msm_coarse_kernel<true>  partition of (bucket, table index | sign) entries by coarse bin
//                        msm_fine_sort_kernel     per-bin counting sort by bucket, histogram in shared memory
//       bucket accum     msm_accum_kernel         mixed XYZZ additions; load-balanced segmented reduction:
//                                                 every thread owns a fixed-length chunk of the sorted list,
//                                                 whole runs go straight to their bucket, runs cut by a chunk
//                                                 edge become partial entries
//                        msm_accum_level/finish   the partial entries (full XYZZ adds), two grid levels + one block
//       bucket reduce    msm_super_hi/lo_kernel   bucket id = hi*L + lo: G1[hi] = sum_lo B, G0[lo] = sum_hi B (tree sums)
//                        msm_weighted_kernel      sum_hi (L*hi + 1) G1[hi] + sum_lo lo G0[lo]  ==  sum_b (b + 1) B_b
//                        msm_fold_kernel          -> one XYZZ point per scalar set; the host normalises it to affine
//   * work is independent of the scalar distribution (zero digits are skipped; repeated digits cannot unbalance it).
//
// Algorithmic HBM bytes: 96 B per (scalar, base) pair.  The kernels are bound by 32-bit integer
// multiply throughput, not HBM: one mixed addition is 10 Montgomery products of ~130 IMAD.WIDE each.
#include <memory>

#include "msm.cuh"

namespace pk {

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

struct ScalarSets { const fr_t* s[MSM_MAX_BATCH]; };
template <bool SCATTER>
__global__ void msm_coarse_kernel(ScalarSets sets, uint32_t n, uint32_t table_n, uint32_t base_offset, WindowPlan plan, int fine_bits,
                                  uint32_t NC, uint32_t* coarse, uint2* tmp);
__global__ void msm_fine_sort_kernel(const uint2* tmp, uint2* entries, const uint32_t* coarse_offset, int fine_bits);

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
// table[w][i] = 2^(width of window w-1) * table[w-1][i]  =  2^(start bit of window w) * base_i
__global__ void srs_window_kernel(g1_affine_t* table, size_t n, WindowPlan plan) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p = ldg_affine(table + i);
    for (int w = 1; w < plan.W; ++w) {
        g1_xyzz_t a = g1_xyzz_t::dbl_affine(p);
        for (int k = 1; k < plan.width[w - 1]; ++k) a = a.dbl();
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

// coarse bins hold 2^fine_bits buckets: small enough that one bin's entries (~6.6 K at N = 2^20) fit the fine sort's
// shared-memory staging area, and at most 2^14 bins so the per-block bin counters of the coarse kernels fit too (128 KB)
static int msm_fine_bits(size_t total_buckets) {
    const int lg = ilog2(total_buckets);
    int fb = lg < 8 ? lg : 8;
    while (((total_buckets + (size_t(1) << fb) - 1) >> fb) > (size_t(1) << 14)) ++fb;  // a 3-set group is not a power of two
    return fb;
}

static int max_batch_for(const SrsTables* s) {
    size_t M = (size_t)s->n * s->W;
    int nb = MSM_MAX_BATCH;
    while (nb > 1 && ((size_t)nb * M >= (size_t(1) << 31) || (size_t)nb * M * 8 > (size_t(6) << 30))) --nb;
    return nb;
}

static void alloc_scratch(const SrsTables* s, MsmScratch& sc, int nb) {
    const size_t M = (size_t)s->n * s->W;
    const size_t NB = (size_t)nb * s->B;
    size_t ncmax = 2;
    for (int k = 1; k <= nb; ++k) {  // every group size that may run on this scratch
        const size_t nbk = (size_t)k * s->B;
        const size_t nc = ((nbk + (size_t(1) << msm_fine_bits(nbk)) - 1) >> msm_fine_bits(nbk)) + 2;
        if (nc > ncmax) ncmax = nc;
    }
    sc.max_sets = nb;
    sc.coarse_count.alloc(ncmax);
    sc.coarse_offset.alloc(ncmax);
    sc.coarse_cursor.alloc(ncmax);
    sc.scan_sums.alloc(ncmax / 4096 + 2);
    sc.entries.alloc(nb * M);
    sc.tmp_entries.alloc(nb * M);
    sc.buckets.alloc(NB);
    const size_t nchunks = (nb * M + s->chunk1 - 1) / s->chunk1;
    for (int k = 0; k < 2; ++k) {  // partial-run lists: 2 entries per level-1 chunk
        sc.pkeys[k].alloc(2 * nchunks + 64);
        sc.ppts[k].alloc(2 * nchunks + 64);
    }
    sc.counts.alloc(16);
    sc.super.alloc((size_t)nb * ((size_t(1) << s->hi_bits) + (size_t(1) << s->lo_bits)));
    sc.red.alloc((size_t)MSM_MAX_BATCH * 1024 + MSM_MAX_BATCH);
}

void srs_load(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits, bool lagrange) {
    PK_REQUIRE(n >= 1, PK_ERR_INVALID, "empty SRS");
    PK_REQUIRE(n <= (uint64_t(1) << 26), PK_ERR_DEGREE_TOO_LARGE, "SRS larger than 2^26 bases (SETUP_MAX_POW2, src/plonk.rs:27)");
    int c = window_bits ? window_bits : pick_window_bits(n);
    PK_REQUIRE(c >= 4 && c <= 24, PK_ERR_INVALID, "window_bits out of range (4..24)");
    int W = (255 + c - 1) / c;
    PK_REQUIRE((uint64_t)W * n < (uint64_t(1) << 31), PK_ERR_DEGREE_TOO_LARGE, "window table index does not fit 31 bits");
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    // the old tables go first (two resident copies may not fit); the new ones are published only once every
    // allocation and kernel has succeeded, so a failed load leaves "no SRS", never a half-built one
    SrsTables*& slot = lagrange ? ctx->srs_lagrange : ctx->srs;
    if (!lagrange && ctx->srs_lagrange && ctx->srs_lagrange->scratch != &ctx->srs_lagrange->own_scratch) {
        delete ctx->srs_lagrange;  // it borrows the working set of the tables that are about to go
        ctx->srs_lagrange = nullptr;
    }
    delete slot;
    slot = nullptr;
    std::unique_ptr<SrsTables> holder(new SrsTables());
    SrsTables* s = holder.get();
    // Window widths: the 255 scalar bits (254 + one spare for the signed-digit carry) are split as evenly as possible
    // over W windows (e.g. c = 20: eight 20-bit and five 19-bit windows).  A plain "12 x 20 bits + 14 bits" split
    // would pile the short top window's 2^20 entries onto 2^14 buckets and unbalance the per-bin sort 4:1.
    WindowPlan plan;
    memset(&plan, 0, sizeof(plan));
    plan.W = W;
    {
        const int base = 255 / W, rem = 255 % W;
        int maxw = 0;
        for (int w = 0; w < W; ++w) {
            plan.width[w] = (uint8_t)(base + (w >= W - rem ? 1 : 0));  // the wider windows sit at the top
            if (plan.width[w] > maxw) maxw = plan.width[w];
        }
        c = maxw;
    }
    plan.c = c;
    s->plan = plan;
    s->n = n; s->c = c; s->W = W; s->B = 1u << (c - 1);
    s->hi_bits = (c - 1) / 2 < 7 ? (c - 1) / 2 : 7;
    s->lo_bits = (c - 1) - s->hi_bits;
    s->table.alloc((size_t)W * n);
    PK_CUDA(cudaMemcpyAsync(s->table.p, bases_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    bases_to_mont_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(s->table.p, n);
    srs_window_kernel<<<grid1d(n, 128), 128, 0, ctx->stream>>>(s->table.p, n, plan);
    ctx->prof.kernel_launches += 2;
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaFuncSetAttribute(msm_coarse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PK_CUDA(cudaFuncSetAttribute(msm_coarse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PK_CUDA(cudaFuncSetAttribute(msm_fine_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    const int nb = max_batch_for(s);
    size_t M = (size_t)n * W;
    // level-1 chunk: keep the chunk count of one scalar set <= 2^19 so the partial lists stay small
    uint32_t chunk = 64;
    while (M / chunk > (size_t(1) << 19)) chunk *= 2;
    s->chunk1 = chunk;
    if (lagrange && ctx->srs && ctx->srs->n == n && ctx->srs->c == s->c && ctx->srs->W == s->W && ctx->srs->chunk1 == s->chunk1)
        s->scratch = ctx->srs->scratch;  // same plan: one working set serves both keys (they never run concurrently)
    else
        alloc_scratch(s, s->own_scratch, nb);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
    slot = holder.release();
}

// ---------------------------------------------------------------- window scan + sort by bucket (two-level radix partition)
// Entries are (bucket id, table index | sign << 31).  Bucket ids are global: set * B + bucket.  They are sorted by a
// coarse partition on the high bits (block-local shared-memory histograms, one global atomic per (block, bin)) followed
// by a per-bin counting sort on the low `fine_bits` bits held entirely in shared memory.  Global atomics per batch drop
// from one per digit (54 M at N = 2^20 x 4 sets) to one per (block, coarse bin).

// calls f(w, global_bucket, neg) for every non-zero signed digit of the canonical scalar k under the window plan
template <class F> __device__ __forceinline__ void for_each_digit(const fr_t& k, const WindowPlan& plan, uint32_t set_base, F f) {
    uint32_t carry = 0;
    uint64_t acc = 0;  // sliding bit window over the limbs: registers only
    int nbits = 0, w = 0;
    auto emit = [&](uint32_t raw, int width) {
        uint32_t d = raw + carry;
        uint32_t neg = 0;
        if (d > (1u << (width - 1))) { d = (1u << width) - d; neg = 1; carry = 1; } else carry = 0;
        if (d != 0) f((uint32_t)w, set_base + d - 1, neg);
        ++w;
    };
#pragma unroll
    for (int limb = 0; limb < 8; ++limb) {
        acc |= (uint64_t)k.v[limb] << nbits;
        nbits += 32;
        while (w < plan.W && nbits >= (int)plan.width[w]) {
            const int width = plan.width[w];
            emit((uint32_t)acc & ((1u << width) - 1), width);
            acc >>= width;
            nbits -= width;
        }
    }
    if (w < plan.W) emit((uint32_t)acc, plan.width[w]);  // top window: the remaining bits (fewer than its width)
}

// (Warp-aggregated atomics — __match_any_sync on the bin, one shared-memory atomic per (warp, bin) — were measured for the
// histogram and scatter atomics below: they defuse a hot bucket, but on uniform scalars the match costs more than the
// uncontended atomics it saves: MSM 2^20 4.72 vs 3.92 ms, a proof 52.7 vs 45.5 ms.  Not used.)
#define COARSE_PER_THREAD 16
// SCATTER = false: coarse_counts[bin] += digits of this block in bin.
// SCATTER = true : reserves a range per (block, bin) in coarse_cursor and writes the entries there.
template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_coarse_kernel(ScalarSets sets, uint32_t n, uint32_t table_n, uint32_t base_offset, WindowPlan plan,
                                                         int fine_bits, uint32_t NC, uint32_t* coarse, uint2* tmp) {
    extern __shared__ uint32_t sh[];  // cnt[NC] (+ base[NC] when scattering)
    uint32_t* cnt = sh;
    uint32_t* base = sh + NC;
    for (uint32_t b = threadIdx.x; b < NC; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    // a block owns COARSE_PER_THREAD * blockDim consecutive scalars: many entries per (block, bin) keep the global
    // atomics rare and the scattered runs long
    const uint32_t i0 = blockIdx.x * (blockDim.x * COARSE_PER_THREAD) + threadIdx.x;
    const uint32_t set_base = blockIdx.y << (plan.c - 1);
    const fr_t* src = sets.s[blockIdx.y];
    for (int r = 0; r < COARSE_PER_THREAD; ++r) {
        const uint32_t i = i0 + r * blockDim.x;
        if (i >= n) break;
        fr_t k = ld_fp(src + i).from_mont();
        if (k.is_zero()) continue;
        for_each_digit(k, plan, set_base, [&](uint32_t, uint32_t g, uint32_t) { atomicAdd(&cnt[g >> fine_bits], 1u); });
    }
    __syncthreads();
    if (!SCATTER) {
        for (uint32_t b = threadIdx.x; b < NC; b += blockDim.x)
            if (cnt[b]) atomicAdd(coarse + b, cnt[b]);
        return;
    }
    for (uint32_t b = threadIdx.x; b < NC; b += blockDim.x) {
        if (cnt[b]) base[b] = atomicAdd(coarse + b, cnt[b]);
        cnt[b] = 0;
    }
    __syncthreads();
    for (int r = 0; r < COARSE_PER_THREAD; ++r) {
        const uint32_t i = i0 + r * blockDim.x;
        if (i >= n) break;
        fr_t k = ld_fp(src + i).from_mont();
        if (k.is_zero()) continue;
        for_each_digit(k, plan, set_base, [&](uint32_t w, uint32_t g, uint32_t neg) {
            const uint32_t bin = g >> fine_bits;
            const uint32_t slot = atomicAdd(&cnt[bin], 1u);
            tmp[base[bin] + slot] = make_uint2(g, (w * table_n + base_offset + i) | (neg << 31));
        });
    }
}

// Counting sort of one coarse bin by the low fine_bits of the bucket id.  The bin's entries are staged in shared
// memory (one global read per entry; bins hold ~6.6 K entries at N = 2^20), counted and ranked there, and written
// to their final place; entries beyond the staging capacity (skewed inputs) are re-read from global memory instead.
#define FS_THREADS 256
#define FS_CAP 9216  // staged entries per block (72 KB): the fullest bins of the balanced window plan hold ~8.7 K at N = 2^20
__global__ void __launch_bounds__(FS_THREADS) msm_fine_sort_kernel(const uint2* tmp, uint2* entries, const uint32_t* coarse_offset, int fine_bits) {
    extern __shared__ uint2 stage[];                                    // [FS_CAP] entries, then hist[2^fine_bits], part[FS_THREADS]
    const uint32_t F = 1u << fine_bits, fmask = F - 1;
    uint32_t* hist = reinterpret_cast<uint32_t*>(stage + FS_CAP);
    uint32_t* part = hist + F;
    const uint32_t lo = coarse_offset[blockIdx.x], hi = coarse_offset[blockIdx.x + 1];
    if (lo == hi) return;
    const uint32_t size = hi - lo;
    const uint32_t staged = size < FS_CAP ? size : FS_CAP;
    for (uint32_t b = threadIdx.x; b < F; b += FS_THREADS) hist[b] = 0;
    {
        // many loads in flight per thread before the first one is consumed (the kernel is latency-bound)
        constexpr int HALF = FS_CAP / FS_THREADS / 2;
        for (int h = 0; h < 2; ++h) {
            uint2 r[HALF];
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const uint32_t i = threadIdx.x + (h * HALF + k) * FS_THREADS;
                if (i < staged) r[k] = tmp[lo + i];
            }
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                const uint32_t i = threadIdx.x + (h * HALF + k) * FS_THREADS;
                if (i < staged) stage[i] = r[k];
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < size; i += FS_THREADS) {
        const uint32_t key = (i < staged ? stage[i].x : tmp[lo + i].x) & fmask;
        atomicAdd(&hist[key], 1u);
    }
    __syncthreads();
    // exclusive scan of hist[0..F): each thread owns F / FS_THREADS consecutive counters
    const uint32_t per = (F + FS_THREADS - 1) / FS_THREADS;
    const uint32_t b0 = threadIdx.x * per < F ? threadIdx.x * per : F, b1 = (b0 + per < F) ? b0 + per : F;
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += hist[b];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (unsigned d = 1; d < FS_THREADS; d <<= 1) {
        uint32_t t = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t b = b0; b < b1; ++b) { uint32_t x = hist[b]; hist[b] = run; run += x; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < size; i += FS_THREADS) {
        const uint2 ent = i < staged ? stage[i] : tmp[lo + i];
        const uint32_t pos = atomicAdd(&hist[ent.x & fmask], 1u);
        entries[lo + pos] = ent;
    }
}

// ---------------------------------------------------------------- exclusive scan of the histogram (tiles of 4096)
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* sh, uint32_t* total) {
    const unsigned tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {
        uint32_t t = tid >= d ? sh[tid - d] : 0;
        __syncthreads();
        sh[tid] += t;
        __syncthreads();
    }
    if (total) *total = sh[1023];
    return sh[tid] - v;
}
__global__ void __launch_bounds__(1024) u32_scan_tile_sums_kernel(const uint32_t* in, uint32_t* sums, uint32_t n) {
    __shared__ uint32_t sh[1024];
    const uint32_t base = blockIdx.x * 4096 + threadIdx.x * 4;
    uint32_t v = 0;
    for (int r = 0; r < 4; ++r) if (base + r < n) v += in[base + r];
    uint32_t tot;
    block_exclusive_scan_1024(v, sh, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
// in-place exclusive scan of m tile sums by one block; total -> *total_out and out_last
__global__ void __launch_bounds__(1024) u32_scan_spine_kernel(uint32_t* sums, uint32_t m, uint32_t* total_out, uint32_t* out_last) {
    __shared__ uint32_t sh[1024];
    const uint32_t per = (m + 1023) / 1024;
    uint32_t lo = threadIdx.x * per, hi = lo + per;
    if (hi > m) hi = m;
    uint32_t v = 0;
    for (uint32_t i = lo; i < hi; ++i) v += sums[i];
    uint32_t tot;
    uint32_t run = block_exclusive_scan_1024(v, sh, &tot);
    for (uint32_t i = lo; i < hi; ++i) { uint32_t x = sums[i]; sums[i] = run; run += x; }
    if (threadIdx.x == 0) { *total_out = tot; *out_last = tot; }
}
__global__ void __launch_bounds__(1024) u32_scan_apply_kernel(const uint32_t* in, const uint32_t* sums, uint32_t* out, uint32_t* out2,
                                                              uint32_t n) {
    __shared__ uint32_t sh[1024];
    const uint32_t base = blockIdx.x * 4096 + threadIdx.x * 4;
    uint32_t x[4], v = 0;
    for (int r = 0; r < 4; ++r) { x[r] = (base + r < n) ? in[base + r] : 0; v += x[r]; }
    uint32_t run = block_exclusive_scan_1024(v, sh, nullptr) + sums[blockIdx.x];
    for (int r = 0; r < 4; ++r) {
        if (base + r < n) { out[base + r] = run; out2[base + r] = run; }
        run += x[r];
    }
}

// ---------------------------------------------------------------- bucket accumulation (segmented reduction over the sorted list)
// Level 1 — msm_accum_kernel: every thread owns a fixed-length chunk of the sorted entry list (perfect balance for any
// scalar distribution) and adds its entries with mixed additions, the window-table gather of entry i + 1 in flight during
// the ten products of entry i.  Runs that lie inside a chunk go straight to their bucket; a run cut by a chunk edge
// leaves a partial (key, XYZZ) entry: slot 2t for the run entering chunk t from the left, slot 2t + 1 for the one leaving
// it on the right.  A slot without a partial sum is a key with bit 31 set and no point is stored or ever loaded.
// (Stitching the cut runs inside the block / warp through shared memory was built and measured: bit-exact, but the extra
// live state pushes the kernel over 128 registers and the hot loop slows by 6-10 %: 6.6 / 6.9 ms vs 6.2 ms per launch.)
// Levels 2, 3 — msm_accum_level_kernel: the same segmented reduction over those entries with full additions, 16 per
// thread; chunk edges are shifted by one entry so that the pair (tail of chunk t, head of chunk t + 1) — the common case,
// a run cut once — is never cut again: after level 2 almost every slot is a "no partial" marker.
// Rest — msm_accum_finish_kernel: ONE block loops over the remaining levels.
#define ACC_THREADS 128
#define KEY_INF 0x80000000u
struct AccumParams {
    const uint2* ent;           // sorted (bucket id, table index | sign << 31)
    const g1_affine_t* table;
    const uint32_t* count_in;   // number of entries (device)
    uint32_t* count_out;        // number of partial entries for level 2
    uint32_t chunk;             // entries per thread
    g1_xyzz_t* buckets;
    uint32_t* out_keys;         // [2 * blocks]
    g1_xyzz_t* out_pts;
};

__global__ void __launch_bounds__(ACC_THREADS) msm_accum_kernel(AccumParams p) {
    const uint32_t count = *p.count_in;
    const uint32_t nchunks = (count + p.chunk - 1) / p.chunk;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *p.count_out = nchunks > 1 ? 2 * nchunks : 0;
    if (t >= nchunks) return;
    const uint32_t s = t * p.chunk;
    const uint32_t e = (s + p.chunk < count) ? s + p.chunk : count;
    const bool head_partial = s > 0 && p.ent[s - 1].x == p.ent[s].x;
    const bool tail_partial = e < count && p.ent[e].x == p.ent[e - 1].x;
    // one entry ahead: the window-table gather of entry i + 1 (a random 64-byte read) is in flight while the ten
    // products of entry i execute
    uint2 en_next = p.ent[s];
    g1_affine_t pt_next = ldg_affine(p.table + (en_next.y & 0x7fffffffu));
    const uint32_t first_key = en_next.x;
    uint32_t cur = first_key;
    bool in_head = true, head_out = false;
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    const g1_xyzz_t inf = g1_xyzz_t::infinity();
    // a partial sum goes to its out slot; "no partial" is a key with KEY_INF set and no point store
    auto emit = [&](uint32_t slot, uint32_t key, const g1_xyzz_t& v) {
        if (v.is_inf()) p.out_keys[slot] = key | KEY_INF;
        else { p.out_keys[slot] = key; st_xyzz(p.out_pts + slot, v); }
    };
    for (uint32_t i = s; i < e; ++i) {
        const uint2 en = en_next;
        g1_affine_t pt = pt_next;
        if (i + 1 < e) {
            en_next = p.ent[i + 1];
            pt_next = ldg_affine(p.table + (en_next.y & 0x7fffffffu));
        }
        if (en.x != cur) {
            acc = acc.lnorm();
            if (in_head && head_partial) { emit(2 * t, first_key, acc); head_out = true; }
            else if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            in_head = false;
            acc = inf;
            cur = en.x;
        }
        if (en.y >> 31) pt.y = pt.y.neg();
        acc = acc.add_mixed_lazy(pt);  // coordinates stay in [0, 2p) inside a run
    }
    acc = acc.lnorm();
    if (in_head) {  // a single run
        if (head_partial || tail_partial) emit(2 * t, cur, acc);
        else {
            if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            p.out_keys[2 * t] = cur | KEY_INF;
        }
        p.out_keys[2 * t + 1] = cur | KEY_INF;
    } else {
        if (!head_out) p.out_keys[2 * t] = first_key | KEY_INF;
        if (tail_partial) emit(2 * t + 1, cur, acc);
        else {
            if (!acc.is_inf()) st_xyzz(p.buckets + cur, acc);
            p.out_keys[2 * t + 1] = cur | KEY_INF;
        }
    }
}

// one chunk [s, e) of a (key, point) level: complete runs -> buckets, the cut first / last run -> out entries 2t, 2t + 1
__device__ __forceinline__ void accum_level_chunk(const uint32_t* keys, const g1_xyzz_t* pts, uint32_t count, uint32_t s, uint32_t e,
                                                  uint32_t t, g1_xyzz_t* buckets, uint32_t* out_keys, g1_xyzz_t* out_pts) {
    auto key_at = [&](uint32_t i) -> uint32_t { return keys[i] & ~KEY_INF; };
    const bool head_partial = s > 0 && key_at(s - 1) == key_at(s);
    const bool tail_partial = e < count && key_at(e) == key_at(e - 1);
    const uint32_t first_key = key_at(s);
    uint32_t cur = first_key;
    bool in_head = true;
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    auto emit = [&](uint32_t slot, uint32_t key, const g1_xyzz_t& v) {
        if (v.is_inf()) out_keys[slot] = key | KEY_INF;
        else { out_keys[slot] = key; st_xyzz(out_pts + slot, v); }
    };
    bool head_done = false;
    for (uint32_t i = s; i < e; ++i) {
        const uint32_t kraw = keys[i];
        const uint32_t k = kraw & ~KEY_INF;
        if (k != cur) {
            if (in_head && head_partial) { emit(2 * t, first_key, acc); head_done = true; }
            else if (!acc.is_inf()) st_xyzz(buckets + cur, acc);
            in_head = false;
            acc = g1_xyzz_t::infinity();
            cur = k;
        }
        if (!(kraw & KEY_INF)) acc = acc.add(ld_xyzz(pts + i));
    }
    if (in_head) {  // a single run
        if (head_partial || tail_partial) emit(2 * t, cur, acc);
        else {
            if (!acc.is_inf()) st_xyzz(buckets + cur, acc);
            out_keys[2 * t] = cur | KEY_INF;
        }
        out_keys[2 * t + 1] = cur | KEY_INF;
    } else {
        if (!head_done) out_keys[2 * t] = first_key | KEY_INF;
        if (tail_partial) emit(2 * t + 1, cur, acc);
        else {
            if (!acc.is_inf()) st_xyzz(buckets + cur, acc);
            out_keys[2 * t + 1] = cur | KEY_INF;
        }
    }
}
#define LEVEL_CHUNK 16
// chunk t covers entries [t * LEVEL_CHUNK + 1, (t + 1) * LEVEL_CHUNK + 1), chunk 0 also entry 0
__device__ __forceinline__ uint32_t level_chunks(uint32_t count) { return count <= 1 ? (count ? 1u : 0u) : (count - 1 + LEVEL_CHUNK - 1) / LEVEL_CHUNK; }
struct LevelParams {
    const uint32_t* keys;
    const g1_xyzz_t* pts;
    const uint32_t* count_in;
    uint32_t* count_out;
    g1_xyzz_t* buckets;
    uint32_t* out_keys;
    g1_xyzz_t* out_pts;
    // finish kernel: ping-pong buffers
    uint32_t* keys2;
    g1_xyzz_t* pts2;
};
__global__ void __launch_bounds__(128) msm_accum_level_kernel(LevelParams p) {
    const uint32_t count = *p.count_in;
    const uint32_t nchunks = level_chunks(count);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *p.count_out = nchunks > 1 ? 2 * nchunks : 0;
    if (t >= nchunks) return;
    const uint32_t s = t ? t * LEVEL_CHUNK + 1 : 0;
    const uint32_t e = ((t + 1) * LEVEL_CHUNK + 1 < count) ? (t + 1) * LEVEL_CHUNK + 1 : count;
    accum_level_chunk(p.keys, p.pts, count, s, e, t, p.buckets, p.out_keys, p.out_pts);
}
__global__ void __launch_bounds__(256) msm_accum_finish_kernel(LevelParams p) {
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = *p.count_in;
    __syncthreads();
    const uint32_t* keys = p.keys;
    const g1_xyzz_t* pts = p.pts;
    uint32_t* okeys = p.keys2;
    g1_xyzz_t* opts = p.pts2;
    for (int level = 0; level < 16; ++level) {
        const uint32_t count = s_count;
        if (count == 0) break;
        const uint32_t nchunks = level_chunks(count);
        for (uint32_t t = threadIdx.x; t < nchunks; t += blockDim.x) {
            const uint32_t s = t ? t * LEVEL_CHUNK + 1 : 0;
            const uint32_t e = ((t + 1) * LEVEL_CHUNK + 1 < count) ? (t + 1) * LEVEL_CHUNK + 1 : count;
            accum_level_chunk(keys, pts, count, s, e, t, p.buckets, okeys, opts);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_count = nchunks > 1 ? 2 * nchunks : 0;
        __syncthreads();
        // next level reads what this one wrote
        const uint32_t* tk = keys; const g1_xyzz_t* tp = pts;
        keys = okeys; pts = opts;
        okeys = const_cast<uint32_t*>(tk); opts = const_cast<g1_xyzz_t*>(tp);
    }
}

// ---------------------------------------------------------------- bucket reduction: sum_b (b + 1) * B_b, b = hi * L + lo
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
// G1[set][hi] = sum_lo B[set][hi * L + lo]; one block per (hi, set)
__global__ void __launch_bounds__(128) msm_super_hi_kernel(const g1_xyzz_t* buckets, int lo_bits, int hi_bits, g1_xyzz_t* super) {
    __shared__ g1_xyzz_t sh[128];
    const uint32_t L = 1u << lo_bits, H = 1u << hi_bits;
    const g1_xyzz_t* src = buckets + ((size_t)blockIdx.y << (lo_bits + hi_bits)) + ((size_t)blockIdx.x << lo_bits);
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) acc = acc.add(ld_xyzz(src + i));
    block_sum_xyzz(acc, sh, super + (size_t)blockIdx.y * (H + L) + blockIdx.x);
}
// G0[set][lo] = sum_hi B[set][hi * L + lo]; block = 32 consecutive lo x 4 hi-groups
__global__ void __launch_bounds__(128) msm_super_lo_kernel(const g1_xyzz_t* buckets, int lo_bits, int hi_bits, g1_xyzz_t* super) {
    __shared__ g1_xyzz_t sh[128];
    const uint32_t L = 1u << lo_bits, H = 1u << hi_bits;
    const uint32_t lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const uint32_t lo = blockIdx.x * 32 + lane;
    const g1_xyzz_t* src = buckets + ((size_t)blockIdx.y << (lo_bits + hi_bits));
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    if (lo < L)
        for (uint32_t hi = grp; hi < H; hi += 4) acc = acc.add(ld_xyzz(src + ((size_t)hi << lo_bits) + lo));
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (grp < 2) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + 64]);
    __syncthreads();
    if (grp == 0 && lo < L) st_xyzz(super + (size_t)blockIdx.y * (H + L) + H + lo, sh[lane].add(sh[lane + 32]));
}
// sum_hi (L * hi + 1) G1[hi] + sum_lo lo * G0[lo], block partials
__global__ void __launch_bounds__(128) msm_weighted_kernel(const g1_xyzz_t* super, int lo_bits, int hi_bits, g1_xyzz_t* red, uint32_t red_stride) {
    __shared__ g1_xyzz_t sh[128];
    const uint32_t L = 1u << lo_bits, H = 1u << hi_bits;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    g1_xyzz_t R = g1_xyzz_t::infinity();
    if (j < H + L) {
        uint32_t weight = j < H ? (j << lo_bits) + 1 : j - H;
        R = ld_xyzz(super + (size_t)blockIdx.y * (H + L) + j).mul_small(weight);
    }
    block_sum_xyzz(R, sh, red + (size_t)blockIdx.y * red_stride + blockIdx.x);
}
__global__ void __launch_bounds__(128) msm_fold_kernel(const g1_xyzz_t* red, uint32_t n, uint32_t red_stride, g1_xyzz_t* out) {
    __shared__ g1_xyzz_t sh[128];
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc = acc.add(ld_xyzz(red + (size_t)blockIdx.x * red_stride + i));
    block_sum_xyzz(acc, sh, out + blockIdx.x);
}

void affine_to_abi(const g1_affine_t& p, uint64_t out[8]) {
    if (p.is_inf()) { memset(out, 0, 64); return; }
    fq_t x = p.x.from_mont(), y = p.y.from_mont();
    memcpy(out, x.v, 32);
    memcpy(out + 4, y.v, 32);
}

// enqueues one group of <= sc.max_sets scalar sets on stream st; the nb XYZZ results are copied to `dst` (pinned host
// memory or device memory) in stream order
static void msm_enqueue_group(pk_ctx* ctx, SrsTables* s, cudaStream_t st, const fr_t* const* scalars, int nb, uint64_t n,
                              uint64_t base_offset, g1_xyzz_t* dst) {
    MsmScratch& sc = *s->scratch;
    const uint32_t B = s->B;
    const uint32_t NB = (uint32_t)nb * B;
    ScalarSets sets;
    for (int k = 0; k < MSM_MAX_BATCH; ++k) sets.s[k] = scalars[k < nb ? k : 0];
    const int fine_bits = msm_fine_bits(NB);
    const uint32_t NC = (NB + (1u << fine_bits) - 1) >> fine_bits;
    PK_CUDA(cudaMemsetAsync(sc.coarse_count.p, 0, (size_t)(NC + 1) * sizeof(uint32_t), st));
    PK_CUDA(cudaMemsetAsync(sc.buckets.p, 0, (size_t)NB * sizeof(g1_xyzz_t), st));
    dim3 dgrid((unsigned)((n + 256 * COARSE_PER_THREAD - 1) / (256 * COARSE_PER_THREAD)), nb);
    msm_coarse_kernel<false><<<dgrid, 256, NC * sizeof(uint32_t), st>>>(sets, (uint32_t)n, (uint32_t)s->n, (uint32_t)base_offset, s->plan,
                                                                      fine_bits, NC, sc.coarse_count.p, nullptr);
    const uint32_t tiles = (NC + 4095) / 4096;
    u32_scan_tile_sums_kernel<<<tiles, 1024, 0, st>>>(sc.coarse_count.p, sc.scan_sums.p, NC);
    u32_scan_spine_kernel<<<1, 1024, 0, st>>>(sc.scan_sums.p, tiles, sc.counts.p, sc.coarse_offset.p + NC);
    u32_scan_apply_kernel<<<tiles, 1024, 0, st>>>(sc.coarse_count.p, sc.scan_sums.p, sc.coarse_offset.p, sc.coarse_cursor.p, NC);
    msm_coarse_kernel<true><<<dgrid, 256, 2 * NC * sizeof(uint32_t), st>>>(sets, (uint32_t)n, (uint32_t)s->n, (uint32_t)base_offset, s->plan,
                                                                         fine_bits, NC, sc.coarse_cursor.p, sc.tmp_entries.p);
    msm_fine_sort_kernel<<<NC, FS_THREADS, FS_CAP * sizeof(uint2) + ((size_t(1) << fine_bits) + FS_THREADS) * sizeof(uint32_t), st>>>(
        sc.tmp_entries.p, sc.entries.p, sc.coarse_offset.p, fine_bits);
    ctx->prof.kernel_launches += 6;
    // accumulation: level 1, two grid levels over the partial entries, one finishing block
    {
        const size_t max_entries = (size_t)nb * n * s->W;
        const size_t nchunks = (max_entries + s->chunk1 - 1) / s->chunk1;
        AccumParams p;
        memset(&p, 0, sizeof(p));
        p.ent = sc.entries.p;
        p.table = s->table.p;
        p.count_in = sc.counts.p;
        p.count_out = sc.counts.p + 1;
        p.chunk = s->chunk1;
        p.buckets = sc.buckets.p;
        p.out_keys = sc.pkeys[0].p;
        p.out_pts = sc.ppts[0].p;
        {
            ScopedKernelTimer timer(ctx, 0, (uint64_t)nb * n, st);
            msm_accum_kernel<<<grid1d(nchunks, ACC_THREADS), ACC_THREADS, 0, st>>>(p);
            ctx->prof.msm_accum_launches++;
        }
        LevelParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.buckets = sc.buckets.p;
        size_t entries = 2 * nchunks;
        int src = 0;
        for (int level = 0; level < 2; ++level) {
            const size_t chunks = (entries + LEVEL_CHUNK - 1) / LEVEL_CHUNK + 1;
            lp.keys = sc.pkeys[src].p; lp.pts = sc.ppts[src].p;
            lp.count_in = sc.counts.p + 1 + level; lp.count_out = sc.counts.p + 2 + level;
            lp.out_keys = sc.pkeys[src ^ 1].p; lp.out_pts = sc.ppts[src ^ 1].p;
            msm_accum_level_kernel<<<grid1d(chunks, 128), 128, 0, st>>>(lp);
            entries = 2 * chunks;
            src ^= 1;
        }
        lp.keys = sc.pkeys[src].p; lp.pts = sc.ppts[src].p;
        lp.count_in = sc.counts.p + 3; lp.count_out = nullptr;
        lp.keys2 = sc.pkeys[src ^ 1].p; lp.pts2 = sc.ppts[src ^ 1].p;
        msm_accum_finish_kernel<<<1, 256, 0, st>>>(lp);
        ctx->prof.kernel_launches += 4;
    }
    // bucket reduction
    const uint32_t H = 1u << s->hi_bits, L = 1u << s->lo_bits;
    msm_super_hi_kernel<<<dim3(H, nb), 128, 0, st>>>(sc.buckets.p, s->lo_bits, s->hi_bits, sc.super.p);
    msm_super_lo_kernel<<<dim3((L + 31) / 32, nb), 128, 0, st>>>(sc.buckets.p, s->lo_bits, s->hi_bits, sc.super.p);
    const uint32_t wblocks = (H + L + 127) / 128;
    PK_REQUIRE(wblocks <= 1024, PK_ERR_INVALID, "bucket reduction partial buffer too small");
    g1_xyzz_t* result = sc.red.p + (size_t)MSM_MAX_BATCH * 1024;
    msm_weighted_kernel<<<dim3(wblocks, nb), 128, 0, st>>>(sc.super.p, s->lo_bits, s->hi_bits, sc.red.p, 1024);
    msm_fold_kernel<<<nb, 128, 0, st>>>(sc.red.p, wblocks, 1024, result);
    ctx->prof.kernel_launches += 4;
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(dst, result, nb * sizeof(g1_xyzz_t), cudaMemcpyDefault, st));
}

void msm_run_batch(pk_ctx* ctx, const fr_t* const* scalars, int nb, uint64_t n, uint64_t base_offset, g1_affine_t* out,
                   SrsTables* tables) {
    SrsTables* s = tables ? tables : ctx->srs;
    PK_REQUIRE(s != nullptr, PK_ERR_DEGREE_TOO_LARGE, "no SRS loaded (pk_srs_load_g1)");
    PK_REQUIRE(base_offset + n <= s->n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    if (n == 0) {
        for (int k = 0; k < nb; ++k) out[k] = g1_affine_t::infinity();
        return;
    }
    const int group = s->scratch->max_sets;
    g1_xyzz_t* host_pt = reinterpret_cast<g1_xyzz_t*>(ctx->pinned);
    // (Splitting a group over two streams so that one half's sort overlaps the other half's accumulation was measured
    // at N = 2^20: 4 % slower, the halves pay the latency-bound tail kernels twice.  The overlap that pays is between
    // independent proofs, each on its own context: plonk.ProverPool.)
    for (int k = 0; k < nb; k += group) {
        const int g = nb - k < group ? nb - k : group;
        msm_enqueue_group(ctx, s, ctx->stream, scalars + k, g, n, base_offset, host_pt);
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < g; ++j) out[k + j] = host_pt[j].to_affine();
    }
}

void msm_run_batch_dev(pk_ctx* ctx, const fr_t* const* scalars, int nb, uint64_t n, uint64_t base_offset, g1_xyzz_t* out_dev) {
    SrsTables* s = ctx->srs;
    PK_REQUIRE(s != nullptr, PK_ERR_DEGREE_TOO_LARGE, "no SRS loaded (pk_srs_load_g1)");
    PK_REQUIRE(base_offset + n <= s->n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    if (n == 0) {
        PK_CUDA(cudaMemsetAsync(out_dev, 0, nb * sizeof(g1_xyzz_t), ctx->stream));  // ZZ = 0: infinity
        return;
    }
    const int group = s->scratch->max_sets;
    for (int k = 0; k < nb; k += group) {
        const int g = nb - k < group ? nb - k : group;
        msm_enqueue_group(ctx, s, ctx->stream, scalars + k, g, n, base_offset, out_dev + k);
    }
}

g1_affine_t msm_run(pk_ctx* ctx, const fr_t* scalars, uint64_t n, uint64_t base_offset) {
    g1_affine_t out;
    msm_run_batch(ctx, &scalars, 1, n, base_offset, &out);
    return out;
}

}  // namespace pk
