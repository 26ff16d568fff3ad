// MSM layer interface (implemented in msm.cu).
#pragma once
#include "common.cuh"

namespace pk {

static const int MSM_MAX_BATCH = 4;

// signed-digit window layout: window w covers `width[w]` bits starting at the sum of the lower widths; c = max width
struct WindowPlan {
    int W;
    int c;
    uint8_t width[64];
};

// working set of one MSM group
struct MsmScratch {
    int max_sets = 0;
    DevBuf<uint32_t> coarse_count, coarse_offset, coarse_cursor;  // [NC + 1] coarse-bin histogram / offsets / cursors
    DevBuf<uint32_t> scan_sums;          // block sums of the offsets scan
    DevBuf<uint2> entries, tmp_entries;  // [sets * n * W] (bucket id, table index | sign << 31): sorted / partitioned
    DevBuf<g1_xyzz_t> buckets;           // [sets * B]
    DevBuf<uint32_t> pkeys[2];           // partial-run lists (ping-pong between levels)
    DevBuf<g1_xyzz_t> ppts[2];
    DevBuf<uint32_t> counts;             // [16] per-level entry counts (device)
    DevBuf<g1_xyzz_t> super;             // [sets][2^hi_bits + 2^lo_bits] super-bucket sums
    DevBuf<g1_xyzz_t> red;               // reduction partials + [sets] results
};

struct SrsTables {
    uint64_t n = 0;          // resident bases
    int c = 0;               // window bits (signed digits in (-2^(c-1), 2^(c-1)])
    int W = 0;               // windows = ceil(255 / c)
    uint32_t B = 0;          // buckets per scalar set = 2^(c-1)
    WindowPlan plan;
    int lo_bits = 0, hi_bits = 0;  // bucket id = hi * 2^lo_bits + lo (two-level bucket reduction)
    DevBuf<g1_affine_t> table;   // [W][n]: table[w][i] = 2^(c*w) * base_i, affine, Montgomery form

    MsmScratch own_scratch;
    MsmScratch* scratch = &own_scratch;  // the Lagrange-form tables borrow the monomial tables' working set when the plans match
    uint32_t chunk1 = 64;                // entries per thread at level 1 (for a single scalar set)
};

// loads n affine bases (canonical limbs, host) and builds the window tables: the monomial-form key (ctx->srs) or, with
// lagrange = true, the Lagrange-form key of the same size (ctx->srs_lagrange; commit_using_values, src/plonk.rs:138-146)
void srs_load(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits, bool lagrange = false);
// out[k] = sum_i scalars[k][i] * base[base_offset + i] for k < nb <= MSM_MAX_BATCH; scalars on the device in
// Montgomery form; results affine (Montgomery) on the host.  One pass over the shared kernels for the whole batch.
// `tables`: which resident key (default ctx->srs)
void msm_run_batch(pk_ctx* ctx, const fr_t* const* scalars, int nb, uint64_t n, uint64_t base_offset, g1_affine_t* out,
                   SrsTables* tables = nullptr);
// same, but the nb XYZZ sums stay on the device (out_dev[k]) and nothing is synchronised: the partial sums of a sharded
// commitment, which are all-gathered and folded before they are normalised
void msm_run_batch_dev(pk_ctx* ctx, const fr_t* const* scalars, int nb, uint64_t n, uint64_t base_offset, g1_xyzz_t* out_dev);
g1_affine_t msm_run(pk_ctx* ctx, const fr_t* scalars, uint64_t n, uint64_t base_offset);
// host helper: affine Montgomery point -> canonical u64[8] ((0,0) for infinity)
void affine_to_abi(const g1_affine_t& p, uint64_t out[8]);

}  // namespace pk
