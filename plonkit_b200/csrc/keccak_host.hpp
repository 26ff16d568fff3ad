// Host-side Keccak-256 (original 0x01 padding = Ethereum keccak256) and the rolling transcript built on it.
// Replaces bellman's RollingKeccakTranscript (tiny-keccak 1.5.0, Cargo.lock:2047-2048), the Fiat-Shamir transcript
// plonkit selects with "keccak" (src/plonk.rs:139-159); byte-level behaviour follows contrib/template.sol:267-307.
// ~40 hashes of 100 bytes per proof: stays on the host.
#pragma once
#include <cstdint>
#include <cstring>

namespace pk {

struct Keccak256 {
    uint64_t a[25];
    uint8_t buf[136];
    size_t fill = 0;

    Keccak256() { memset(a, 0, sizeof(a)); }

    static uint64_t rol(uint64_t x, unsigned s) { return (x << s) | (x >> (64 - s)); }

    void permute() {
        static const uint64_t rc[24] = {0x1ULL, 0x8082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x808bULL, 0x80000001ULL,
                                        0x8000000080008081ULL, 0x8000000000008009ULL, 0x8aULL, 0x88ULL, 0x80008009ULL, 0x8000000aULL,
                                        0x8000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                        0x8000000000008002ULL, 0x8000000000000080ULL, 0x800aULL, 0x800000008000000aULL,
                                        0x8000000080008081ULL, 0x8000000000008080ULL, 0x80000001ULL, 0x8000000080008008ULL};
        // lane walk of rho+pi starting from lane 1: successive destination lanes and rotation amounts
        static const int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        static const int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        for (int r = 0; r < 24; ++r) {
            uint64_t bc[5];
            for (int i = 0; i < 5; ++i) bc[i] = a[i] ^ a[i + 5] ^ a[i + 10] ^ a[i + 15] ^ a[i + 20];
            for (int i = 0; i < 5; ++i) {
                uint64_t t = bc[(i + 4) % 5] ^ rol(bc[(i + 1) % 5], 1);
                for (int j = 0; j < 25; j += 5) a[j + i] ^= t;
            }
            uint64_t t = a[1];
            for (int i = 0; i < 24; ++i) {
                int j = piln[i];
                uint64_t tmp = a[j];
                a[j] = rol(t, rotc[i]);
                t = tmp;
            }
            for (int j = 0; j < 25; j += 5) {
                for (int i = 0; i < 5; ++i) bc[i] = a[j + i];
                for (int i = 0; i < 5; ++i) a[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
            }
            a[0] ^= rc[r];
        }
    }
    void absorb_block() {
        for (int i = 0; i < 17; ++i) {
            uint64_t w = 0;
            for (int k = 0; k < 8; ++k) w |= (uint64_t)buf[8 * i + k] << (8 * k);
            a[i] ^= w;
        }
        permute();
        fill = 0;
    }
    void update(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; ++i) {
            buf[fill++] = d[i];
            if (fill == 136) absorb_block();
        }
    }
    void finish(uint8_t out[32]) {
        memset(buf + fill, 0, 136 - fill);
        buf[fill] ^= 0x01;
        buf[135] ^= 0x80;
        fill = 136;
        absorb_block();
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 8; ++k) out[8 * i + k] = (uint8_t)(a[i] >> (8 * k));
    }
};

// state_0/state_1 rolling transcript with domain-separation tags 0, 1 (updates) and 2 (challenges)
struct RollingKeccakTranscript {
    uint8_t s0[32], s1[32];
    uint32_t counter = 0;
    RollingKeccakTranscript() { memset(s0, 0, 32); memset(s1, 0, 32); }

    static void put_be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

    void commit_be(const uint8_t value[32]) {  // update_with_u256
        uint8_t n0[32], n1[32], tag[4];
        for (uint32_t dst = 0; dst < 2; ++dst) {
            Keccak256 h;
            put_be32(tag, dst);
            h.update(tag, 4); h.update(s0, 32); h.update(s1, 32); h.update(value, 32);
            h.finish(dst ? n1 : n0);
        }
        memcpy(s0, n0, 32);
        memcpy(s1, n1, 32);
    }
    // canonical little-endian 32-bit limbs -> big-endian bytes
    void commit_limbs(const uint32_t v[8]) {
        uint8_t be[32];
        for (int i = 0; i < 8; ++i) put_be32(be + 4 * i, v[7 - i]);
        commit_be(be);
    }
    // 253-bit challenge as canonical little-endian limbs
    void challenge(uint32_t out[8]) {
        uint8_t tag[4], ctr[4], h32[32];
        Keccak256 h;
        put_be32(tag, 2);
        put_be32(ctr, counter++);
        h.update(tag, 4); h.update(s0, 32); h.update(s1, 32); h.update(ctr, 4);
        h.finish(h32);
        h32[0] &= 0x1f;
        for (int i = 0; i < 8; ++i)
            out[7 - i] = ((uint32_t)h32[4 * i] << 24) | ((uint32_t)h32[4 * i + 1] << 16) | ((uint32_t)h32[4 * i + 2] << 8) | h32[4 * i + 3];
    }
};

}  // namespace pk
