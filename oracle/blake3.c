/* ORACLE (test infrastructure, NOT product code).
 *
 * BLAKE3 (default hash mode, 32-byte output) and the modes the reference's `Blake3GoldilocksConfig` builds on it.
 *
 * The algorithm itself lives OUTSIDE /root/reference: crate `blake3` 1.5.0 (Cargo.lock:220-221), called at
 *   plonky2/plonky2/src/hash/blake3.rs:176  (Blake3Permutation::permute: the "hash onion" with rejection sampling)
 *   plonky2/plonky2/src/hash/blake3.rs:216  (Blake3_256::hash_no_pad: blake3::hash over the row's raw u64 bytes)
 *   plonky2/plonky2/src/hash/blake3.rs:230  (Blake3_256::two_to_one: blake3::hash(left || right), 64 bytes)
 * This file restates the published BLAKE3 algorithm (the BLAKE3 paper, section 2: compression function G / rounds /
 * message permutation :2.2, chunk chaining values and flags :2.4-2.5, binary tree of chunks :2.1) -- the same constants
 * the reference repeats in its in-circuit description (blake3.rs:16-19 ROUND = 7, STATE_SIZE = 16, BLOCK_LEN = 64;
 * g_field :28-55; round_field :58-78; compress_pre_field :81-135).  PINNED against the official implementation: the
 * Python package `blake3` (bindings of the same Rust crate) generated tests/golden/blake3_kat.json
 * (tools/extract_blake3_golden.py); tests/test_oracle_blake3.py checks every vector.
 *
 * Deviation by design (as everywhere in this oracle, gl.h): the reference hashes the IN-MEMORY u64 of each field element,
 * which may be a non-canonical representative in [p, 2^64) (blake3.rs:210-213, goldilocks_field.rs:162-170); here the
 * canonical representative is hashed.  The two differ for an element with probability ~2^-32 (and when they do, the
 * reference's own verifier, which re-hashes the canonical values read from the proof, rejects its own prover's path).
 */
#include <string.h>

#include "oracle.h"

static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

static inline uint32_t rotr32(uint32_t x, int k) { return (x >> k) | (x << (32 - k)); }

static void b3_g(uint32_t *v, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    v[a] = v[a] + v[b] + mx;
    v[d] = rotr32(v[d] ^ v[a], 16);
    v[c] = v[c] + v[d];
    v[b] = rotr32(v[b] ^ v[c], 12);
    v[a] = v[a] + v[b] + my;
    v[d] = rotr32(v[d] ^ v[a], 8);
    v[c] = v[c] + v[d];
    v[b] = rotr32(v[b] ^ v[c], 7);
}

/* one compression; out[0..8) is the new chaining value (or the first 32 output bytes at the root) */
static void b3_compress(const uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint64_t counter, uint32_t flags,
                        uint32_t out[8]) {
    uint32_t m[16], t[16], v[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) |
               ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) v[i] = cv[i];
    for (int i = 0; i < 4; i++) v[8 + i] = B3_IV[i];
    v[12] = (uint32_t)counter;
    v[13] = (uint32_t)(counter >> 32);
    v[14] = block_len;
    v[15] = flags;
    for (int r = 0; r < 7; r++) {
        b3_g(v, 0, 4, 8, 12, m[0], m[1]);
        b3_g(v, 1, 5, 9, 13, m[2], m[3]);
        b3_g(v, 2, 6, 10, 14, m[4], m[5]);
        b3_g(v, 3, 7, 11, 15, m[6], m[7]);
        b3_g(v, 0, 5, 10, 15, m[8], m[9]);
        b3_g(v, 1, 6, 11, 12, m[10], m[11]);
        b3_g(v, 2, 7, 8, 13, m[12], m[13]);
        b3_g(v, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}

/* chaining value of chunk number `index` (len in [0, 1024]; 0 only for the empty input) */
static void b3_chunk_cv(const uint8_t *p, size_t len, uint64_t index, int is_root, uint32_t out[8]) {
    uint32_t cv[8];
    memcpy(cv, B3_IV, sizeof cv);
    size_t nblocks = len ? (len + 63) / 64 : 1;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64] = {0};
        size_t take = len - 64 * b < 64 ? len - 64 * b : 64;
        memcpy(block, p + 64 * b, take);
        uint32_t flags = 0;
        if (b == 0) flags |= B3_CHUNK_START;
        if (b + 1 == nblocks) flags |= B3_CHUNK_END | (is_root ? B3_ROOT : 0);
        uint32_t nx[8];
        b3_compress(cv, block, (uint32_t)take, index, flags, nx);
        memcpy(cv, nx, sizeof cv);
    }
    memcpy(out, cv, sizeof cv);
}

/* chaining value of the subtree over `len` bytes starting at chunk `index0`: the left subtree takes the largest power
 * of two of chunks that leaves at least one byte on the right */
static void b3_subtree_cv(const uint8_t *p, size_t len, uint64_t index0, int is_root, uint32_t out[8]) {
    if (len <= 1024) {
        b3_chunk_cv(p, len, index0, is_root, out);
        return;
    }
    size_t full = (len - 1) / 1024, left = 1;
    while (left * 2 <= full) left *= 2;
    uint32_t l[8], r[8];
    b3_subtree_cv(p, left * 1024, index0, 0, l);
    b3_subtree_cv(p + left * 1024, len - left * 1024, index0 + left, 0, r);
    uint8_t block[64];
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 4; k++) {
            block[4 * i + k] = (uint8_t)(l[i] >> (8 * k));
            block[32 + 4 * i + k] = (uint8_t)(r[i] >> (8 * k));
        }
    b3_compress(B3_IV, block, 64, 0, B3_PARENT | (is_root ? B3_ROOT : 0), out);
}

void orc_blake3(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t w[8];
    b3_subtree_cv(in, len, 0, 1, w);
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(w[i] >> (8 * k));
}

static void store_le64(uint8_t *p, uint64_t x) {
    for (int k = 0; k < 8; k++) p[k] = (uint8_t)(x >> (8 * k));
}
static uint64_t load_le64(const uint8_t *p) {
    uint64_t x = 0;
    for (int k = 0; k < 8; k++) x |= (uint64_t)p[k] << (8 * k);
    return x;
}

/* Blake3_256::hash_no_pad (blake3.rs:205-218): blake3 over the little-endian u64 image of the elements; the 32 output
 * bytes are kept as 4 little-endian u64 words (BytesHash<32>), which are NOT field elements (words may be >= p). */
void orc_blake3_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
    uint8_t stackbuf[2048] = {0}, *buf = stackbuf, digest[32];
    if (n * 8 > sizeof stackbuf) buf = (uint8_t *)__builtin_malloc(n * 8);
    for (size_t i = 0; i < n; i++) store_le64(buf + 8 * i, gl_canon(in[i]));
    orc_blake3(buf, n * 8, digest);
    if (buf != stackbuf) __builtin_free(buf);
    for (int i = 0; i < 4; i++) out[i] = load_le64(digest + 8 * i);
}

/* Blake3_256::two_to_one (blake3.rs:220-233): blake3(left || right) */
void orc_blake3_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint8_t buf[64], digest[32];
    for (int i = 0; i < 4; i++) {
        store_le64(buf + 8 * i, l[i]);
        store_le64(buf + 32 + 8 * i, r[i]);
    }
    orc_blake3(buf, 64, digest);
    for (int i = 0; i < 4; i++) out[i] = load_le64(digest + 8 * i);
}

/* Blake3Permutation::permute (blake3.rs:165-199): state -> canonical LE bytes (96) -> h1 = blake3(bytes),
 * h2 = blake3(h1), ...; the u64 words of h1, h2, ... that are < p, in order, until 12 are collected. */
void orc_blake3_permute(uint64_t state[12]) {
    uint8_t buf[96], digest[32];
    for (int i = 0; i < 12; i++) store_le64(buf + 8 * i, gl_canon(state[i]));
    size_t len = 96;
    int got = 0;
    while (got < 12) {
        orc_blake3(buf, len, digest);
        memcpy(buf, digest, 32);
        len = 32;
        for (int i = 0; i < 4 && got < 12; i++) {
            uint64_t w = load_le64(digest + 8 * i);
            if (w < GL_P) state[got++] = w;
        }
    }
}

/* GenericHashOut::to_vec for BytesHash<32> (hash_types.rs:142-152): chunks of 7 bytes, little-endian -> 5 elements
 * (7 + 7 + 7 + 7 + 4 bytes); this is what Challenger::observe_hash absorbs (challenger.rs:79-81). */
void orc_bytes_hash_to_fields(const uint64_t h[4], uint64_t out[5]) {
    uint8_t b[32];
    for (int i = 0; i < 4; i++) store_le64(b + 8 * i, h[i]);
    for (int c = 0; c < 5; c++) {
        uint64_t x = 0;
        for (int k = 0; k < 7 && 7 * c + k < 32; k++) x |= (uint64_t)b[7 * c + k] << (8 * k);
        out[c] = x;
    }
}
