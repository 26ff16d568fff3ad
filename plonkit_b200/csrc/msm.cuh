// MSM layer interface (implemented in msm.cu).
#pragma once
#include "common.cuh"

namespace pk {

struct SrsTables {
    uint64_t n = 0;          // resident bases
    int c = 0;               // window bits (signed digits in (-2^(c-1), 2^(c-1)])
    int W = 0;               // windows = ceil(255 / c)
    uint32_t B = 0;          // buckets = 2^(c-1)
    DevBuf<g1_affine_t> table;   // [W][n]: table[w][i] = 2^(c*w) * base_i, affine, Montgomery form

    // scratch, sized for n pairs
    DevBuf<uint32_t> hist;       // [B + 1]
    DevBuf<uint32_t> offsets;    // [B + 1]
    DevBuf<uint32_t> cursor;     // [B]
    DevBuf<uint32_t> keys, items;        // [n * W]
    DevBuf<g1_xyzz_t> buckets;           // [B]
    DevBuf<uint32_t> pkeys[2];           // partial-run lists (ping-pong between levels)
    DevBuf<g1_xyzz_t> ppts[2];
    DevBuf<uint32_t> counts;             // [16] per-level entry counts (device)
    DevBuf<g1_xyzz_t> red;               // bucket-reduction partials
    uint32_t chunk1 = 64;                // entries per thread at level 1
};

// loads n affine bases (canonical limbs, host) and builds the window tables
void srs_load(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits);
// sum_i scalars[i] * base[base_offset + i]; scalars on the device in Montgomery form; result affine (Montgomery) on the host
g1_affine_t msm_run(pk_ctx* ctx, const fr_t* scalars, uint64_t n, uint64_t base_offset);
// host helper: affine Montgomery point -> canonical u64[8] ((0,0) for infinity)
void affine_to_abi(const g1_affine_t& p, uint64_t out[8]);

}  // namespace pk
