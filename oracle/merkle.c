/* ORACLE (test infrastructure, NOT product code) -- see oracle/gl.h header.
 *
 * CPU restatement of the reference's Poseidon Merkle tree as used by the STARK prover.
 *
 * Follows:
 *   plonky2/plonky2/src/hash/merkle_tree/mod.rs   new_v2 :180-266 (leaf = hash_no_pad(row) :198 --
 *       NOT hash_or_noop; heap-ordered nodes; cap = nodes[2^h .. 2^(h+1)) :223-225; digests re-laid
 *       into the per-subtree interleaved sibling order :236-259), prove :273-308,
 *       build_merkle_nodes :311-337
 *   plonky2/plonky2/src/hash/merkle_proofs.rs      verify_merkle_proof_to_cap :52-77
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* merkle_tree/mod.rs:311-337: nodes[1] = root, children of i at 2i, 2i+1; nodes has `nleaves` entries,
 * the parents of the leaves live at nodes[nleaves/2 .. nleaves) */
void orc_build_merkle_nodes(const uint64_t *leaf_hashes, size_t nleaves, uint64_t *nodes) {
    size_t n = nleaves / 2;
    memset(nodes, 0, 4 * sizeof(uint64_t)); /* nodes[0] = zero hash */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++)
        orc_two_to_one(leaf_hashes + 8 * i, leaf_hashes + 8 * i + 4, nodes + 4 * (n + i));
    for (size_t lvl = n / 2; lvl >= 1; lvl /= 2) { /* level by level == the reference's reverse index loop */
#pragma omp parallel for schedule(static)
        for (size_t i = lvl; i < 2 * lvl; i++) orc_two_to_one(nodes + 8 * i, nodes + 8 * i + 4, nodes + 4 * i);
        if (lvl == 1) break;
    }
}

/* new_v2: rows is [nrows][ncols] row-major.  digests_out: 2*(nrows - 2^cap_height) hashes in the
 * reference's layout; cap_out: 2^cap_height hashes.  Returns 0, or -1 on bad arguments. */
int orc_merkle_new_v2(const uint64_t *rows, size_t nrows, size_t ncols, uint32_t cap_height,
                      uint64_t *digests_out, uint64_t *cap_out) {
    uint32_t tree_height_sub_1 = orc_log2_strict(nrows);
    if (((size_t)1 << tree_height_sub_1) != nrows || cap_height > tree_height_sub_1) return -1;
    size_t len_cap = (size_t)1 << cap_height;
    uint64_t *row_hashes = (uint64_t *)malloc(nrows * 32);
    orc_hash_rows(rows, nrows, ncols, row_hashes);
    uint64_t *nodes = (uint64_t *)malloc((nrows > 1 ? nrows : 2) * 32);
    if (nrows >= 2) orc_build_merkle_nodes(row_hashes, nrows, nodes);

    size_t num_digests = 2 * (nrows - len_cap);
    if (len_cap == nrows) {
        memcpy(cap_out, row_hashes, len_cap * 32);
    } else {
        memcpy(cap_out, nodes + 4 * len_cap, len_cap * 32);
    }
    uint32_t num_layers = tree_height_sub_1 - cap_height;
    size_t num_sub_tree_leaves = (size_t)1 << num_layers;
    size_t tree_len = num_digests >> cap_height;
    size_t num_trees = len_cap;
    if (num_digests > 0) {
        for (size_t i = 0; i < num_trees; i++) {
            for (size_t pair_idx = 0; pair_idx < num_sub_tree_leaves; pair_idx += 2) {
                size_t sibling_index = pair_idx << 1;
                memcpy(digests_out + 4 * (tree_len * i + sibling_index), row_hashes + 4 * (num_sub_tree_leaves * i + pair_idx), 32);
                memcpy(digests_out + 4 * (tree_len * i + sibling_index + 1),
                       row_hashes + 4 * (num_sub_tree_leaves * i + pair_idx + 1), 32);
            }
            for (uint32_t layer = 1; layer < num_layers; layer++) {
                size_t num_layer_nodes = num_sub_tree_leaves >> layer;
                for (size_t pair_idx = 0; pair_idx < num_layer_nodes; pair_idx += 2) {
                    size_t siblings_index = (pair_idx << layer) + ((size_t)1 << layer) - 1;
                    size_t sibling_index = siblings_index << 1;
                    size_t n_idx = ((size_t)1 << (tree_height_sub_1 - layer)) + num_layer_nodes * i + pair_idx;
                    memcpy(digests_out + 4 * (tree_len * i + sibling_index), nodes + 4 * n_idx, 32);
                    memcpy(digests_out + 4 * (tree_len * i + sibling_index + 1), nodes + 4 * (n_idx + 1), 32);
                }
            }
        }
    }
    free(nodes);
    free(row_hashes);
    return 0;
}

/* prove :273-308.  siblings_out receives num_layers hashes; returns num_layers. */
int orc_merkle_prove(const uint64_t *digests, size_t nrows, uint32_t cap_height, size_t leaf_index, uint64_t *siblings_out) {
    uint32_t num_layers = orc_log2_strict(nrows) - cap_height;
    size_t num_digests = 2 * (nrows - ((size_t)1 << cap_height));
    size_t tree_index = leaf_index >> num_layers;
    size_t tree_len = num_digests >> cap_height;
    const uint64_t *digest_tree = digests + 4 * tree_len * tree_index;
    size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
    for (uint32_t i = 0; i < num_layers; i++) {
        size_t parity = pair_index & 1;
        pair_index >>= 1;
        size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
        size_t sibling_index = 2 * siblings_index + (1 - parity);
        memcpy(siblings_out + 4 * i, digest_tree + 4 * sibling_index, 32);
    }
    return (int)num_layers;
}

/* merkle_proofs.rs:52-77; returns 1 if valid */
int orc_merkle_verify(const uint64_t *leaf, size_t ncols, size_t leaf_index, const uint64_t *cap,
                      const uint64_t *siblings, size_t nsib) {
    uint64_t cur[4];
    orc_hash_no_pad(leaf, ncols, cur);
    size_t index = leaf_index;
    for (size_t k = 0; k < nsib; k++) {
        size_t bit = index & 1;
        index >>= 1;
        uint64_t nxt[4];
        if (bit)
            orc_two_to_one(siblings + 4 * k, cur, nxt);
        else
            orc_two_to_one(cur, siblings + 4 * k, nxt);
        memcpy(cur, nxt, 32);
    }
    return memcmp(cur, cap + 4 * index, 32) == 0;
}
