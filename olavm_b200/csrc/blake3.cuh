// BLAKE3 (default hash mode, 32-byte output) as a host + device function, and the three modes the reference's
// `Blake3GoldilocksConfig` (plonky2/plonky2/src/plonk/config.rs:153-161) builds on it:
//   Blake3_256::hash_no_pad   plonky2/plonky2/src/hash/blake3.rs:205-218  blake3 over the row's little-endian u64 image
//   Blake3_256::two_to_one    blake3.rs:220-233                          blake3(left || right): ONE compression
//   Blake3Permutation         blake3.rs:165-199                          the challenger's "hash onion" (host only)
// The hash itself is crate `blake3` 1.5.0 in the reference (Cargo.lock:220-221; not vendored); this is the published
// algorithm: 7 rounds of the G quarter-round over a 16-word state, message words permuted between rounds, 64-byte
// blocks chained inside 1024-byte chunks, chunk chaining values merged by a binary tree of parent compressions.
//
// One thread hashes one row: the message words never leave registers (the round schedule is resolved at compile time),
// a row of <= 128 columns is a single chunk (a chain of <= 16 compressions), wider rows (the Poseidon table's 134
// columns) go through the chunk tree with a 4-entry chaining-value stack (rows up to 2048 columns).  32-bit adds, xors
// and rotates only: ~900 ALU-pipe instructions per 64 bytes, so leaf hashing under this config is ~20x cheaper than the
// Poseidon sponge and the commitment moves towards the HBM roofline.
//
// Values are hashed in CANONICAL form (every LDE value in HBM is canonical).  The reference hashes whatever
// representative is in memory (blake3.rs:210-213), which differs from the canonical one with probability ~2^-32 per
// element -- and whenever it does, the reference's own verifier (which re-hashes canonical values read from the proof)
// rejects the path.  Documented in DESIGN.md; not reproducible bit-for-bit by any implementation with different
// arithmetic (including the reference's own AVX2 build against its scalar build).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OLA_B3_HD __host__ __device__ __forceinline__
#else
#define OLA_B3_HD inline
#endif

namespace ola {
namespace blake3 {

enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
constexpr size_t MAX_U64S = 2048;  // 16 chunks: the depth-4 chaining-value stack below

OLA_B3_HD uint32_t iv(int i) {
    // the SHA-256 initial values (first 32 bits of the fractional parts of the square roots of the first 8 primes)
    switch (i) {
        case 0: return 0x6A09E667u;
        case 1: return 0xBB67AE85u;
        case 2: return 0x3C6EF372u;
        case 3: return 0xA54FF53Au;
        case 4: return 0x510E527Fu;
        case 5: return 0x9B05688Cu;
        case 6: return 0x1F83D9ABu;
        default: return 0x5BE0CD19u;
    }
}

OLA_B3_HD uint32_t rotr(uint32_t x, int k) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, k);
#else
    return (x >> k) | (x << (32 - k));
#endif
}

#define OLA_B3_G(a, b, c, d, mx, my) \
    do {                             \
        a = a + b + (mx);            \
        d = rotr(d ^ a, 16);         \
        c = c + d;                   \
        b = rotr(b ^ c, 12);         \
        a = a + b + (my);            \
        d = rotr(d ^ a, 8);          \
        c = c + d;                   \
        b = rotr(b ^ c, 7);          \
    } while (0)

// one round with the message words in schedule order s0..s15
#define OLA_B3_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                   \
        OLA_B3_G(v0, v4, v8, v12, m[s0], m[s1]);                                           \
        OLA_B3_G(v1, v5, v9, v13, m[s2], m[s3]);                                           \
        OLA_B3_G(v2, v6, v10, v14, m[s4], m[s5]);                                          \
        OLA_B3_G(v3, v7, v11, v15, m[s6], m[s7]);                                          \
        OLA_B3_G(v0, v5, v10, v15, m[s8], m[s9]);                                          \
        OLA_B3_G(v1, v6, v11, v12, m[s10], m[s11]);                                        \
        OLA_B3_G(v2, v7, v8, v13, m[s12], m[s13]);                                         \
        OLA_B3_G(v3, v4, v9, v14, m[s14], m[s15]);                                         \
    } while (0)

// cv <- first 8 words of compress(cv, m, counter, block_len, flags).  The 7 schedules are the message permutation
// [2 6 3 10 7 0 4 13 1 11 12 5 9 14 15 8] applied 0..6 times (the table the reference repeats as MSG_SCHEDULE).
OLA_B3_HD void compress(uint32_t cv[8], const uint32_t m[16], uint32_t block_len, uint64_t counter, uint32_t flags) {
    uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
    uint32_t v8 = iv(0), v9 = iv(1), v10 = iv(2), v11 = iv(3);
    uint32_t v12 = (uint32_t)counter, v13 = (uint32_t)(counter >> 32), v14 = block_len, v15 = flags;
    OLA_B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    OLA_B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8);
    OLA_B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1);
    OLA_B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6);
    OLA_B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4);
    OLA_B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7);
    OLA_B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13);
    cv[0] = v0 ^ v8;
    cv[1] = v1 ^ v9;
    cv[2] = v2 ^ v10;
    cv[3] = v3 ^ v11;
    cv[4] = v4 ^ v12;
    cv[5] = v5 ^ v13;
    cv[6] = v6 ^ v14;
    cv[7] = v7 ^ v15;
}

OLA_B3_HD void set_iv(uint32_t cv[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cv[i] = iv(i);
}

// cur <- parent(left, cur)
OLA_B3_HD void parent(const uint32_t left[8], uint32_t cur[8], uint32_t flags) {
    uint32_t m[16], cv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m[i] = left[i];
        m[8 + i] = cur[i];
    }
    set_iv(cv);
    compress(cv, m, 64, 0, PARENT | flags);
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = cv[i];
}

// blake3 over the little-endian image of n u64 (ld(i) = the i-th), n <= MAX_U64S; out = the 32 digest bytes as 4 LE u64.
// MULTI = false promises n <= 128 (one chunk) and compiles the chunk tree away.
template <bool MULTI, class Load>
OLA_B3_HD void hash_u64s(size_t n, Load ld, uint64_t out[4]) {
    uint32_t cv[8];
    set_iv(cv);
    uint32_t stack[MULTI ? 4 : 1][8];
    int sp = 0;
    const size_t nblocks = n ? (n + 7) / 8 : 1;
    const size_t nchunks = (nblocks + 15) / 16;
    for (size_t b = 0; b < nblocks; ++b) {
        uint32_t m[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const size_t idx = 8 * b + k;
            const uint64_t v = idx < n ? ld(idx) : 0;
            m[2 * k] = (uint32_t)v;
            m[2 * k + 1] = (uint32_t)(v >> 32);
        }
        const size_t chunk = b >> 4;
        const bool first = (b & 15) == 0, last_block = b + 1 == nblocks, last_in_chunk = (b & 15) == 15 || last_block;
        uint32_t flags = (first ? (uint32_t)CHUNK_START : 0u) | (last_in_chunk ? (uint32_t)CHUNK_END : 0u);
        if (last_block && (!MULTI || nchunks == 1)) flags |= ROOT;
        compress(cv, m, last_block ? (uint32_t)((n - 8 * b) * 8) : 64u, MULTI ? chunk : 0, flags);
        if (MULTI && last_in_chunk && nchunks > 1) {
            if (!last_block) {
                // a completed chunk that is not the last: merge the subtrees it completes (one per trailing zero bit
                // of the number of chunks so far), push the result
                size_t total = chunk + 1;
                while ((total & 1) == 0) {
                    --sp;
                    parent(stack[sp], cv, 0);
                    total >>= 1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) stack[sp][i] = cv[i];
                ++sp;
                set_iv(cv);
            } else {
                // the last chunk: fold the stack from the top; the final parent is the root
                while (sp > 0) {
                    --sp;
                    parent(stack[sp], cv, sp == 0 ? (uint32_t)ROOT : 0u);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = (uint64_t)cv[2 * i] | ((uint64_t)cv[2 * i + 1] << 32);
}

// two_to_one: blake3 of the 64 bytes left || right -- one block, one chunk, root
OLA_B3_HD void two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint32_t m[16], cv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[2 * i] = (uint32_t)l[i];
        m[2 * i + 1] = (uint32_t)(l[i] >> 32);
        m[8 + 2 * i] = (uint32_t)r[i];
        m[8 + 2 * i + 1] = (uint32_t)(r[i] >> 32);
    }
    set_iv(cv);
    compress(cv, m, 64, 0, CHUNK_START | CHUNK_END | ROOT);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = (uint64_t)cv[2 * i] | ((uint64_t)cv[2 * i + 1] << 32);
}

// ---- host-only helpers (transcript and verifier) ----
struct PtrLoad {
    const uint64_t* p;
    OLA_B3_HD uint64_t operator()(size_t i) const { return p[i]; }
};
// hash_no_pad of canonical elements on the host (any n <= MAX_U64S)
inline void hash_host(const uint64_t* in, size_t n, uint64_t out[4]) { hash_u64s<true>(n, PtrLoad{in}, out); }

// Blake3Permutation::permute (blake3.rs:165-199): h1 = blake3(state as 96 canonical LE bytes), h2 = blake3(h1), ...;
// the u64 words of h1, h2, ... that are < p, in order, until 12 are collected (rejection sampling).
inline void permute_host(uint64_t state[12]) {
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    uint64_t buf[12], h[4];
    for (int i = 0; i < 12; ++i) buf[i] = state[i] >= P ? state[i] - P : state[i];
    size_t len = 12;
    int got = 0;
    while (got < 12) {
        hash_host(buf, len, h);
        for (int i = 0; i < 4; ++i) buf[i] = h[i];
        len = 4;
        for (int i = 0; i < 4 && got < 12; ++i)
            if (h[i] < P) state[got++] = h[i];
    }
}

// GenericHashOut::to_vec for BytesHash<32> (hash/hash_types.rs:142-152): 7-byte little-endian chunks -> 5 elements
inline void hash_to_fields(const uint64_t h[4], uint64_t out[5]) {
    // byte j of the digest is byte (j % 8) of word j / 8; element c takes bytes [7c, 7c + 7)
    for (int c = 0; c < 5; ++c) {
        uint64_t x = 0;
        for (int k = 0; k < 7 && 7 * c + k < 32; ++k) {
            const int j = 7 * c + k;
            x |= ((h[j >> 3] >> (8 * (j & 7))) & 0xFF) << (8 * k);
        }
        out[c] = x;
    }
}

}  // namespace blake3
}  // namespace ola
