"""Mirror of plonky2's hashing entry points used by the prover: `Hasher::hash_no_pad` / `MerkleTree::new_v2` under the
context's hasher (`ctx.hasher`: PoseidonHash, plonky2/plonky2/src/hash/poseidon.rs:640-652 + hashing.rs:66-108, or
Blake3_256<32>, hash/blake3.rs:201-234) and the raw Poseidon permutation (merkle_tree/mod.rs:180)."""
import numpy as np

from . import _lib


def poseidon(ctx, states):
    """Poseidon::poseidon (poseidon.rs:593) on a batch of width-12 states, shape [k, 12] or [12]."""
    a = np.array(states, dtype=np.uint64, copy=True, order="C")
    single = a.ndim == 1
    a2 = a.reshape(-1, 12)
    ctx.check(ctx._lib.ola_poseidon_permute(ctx.handle, _lib.hptr(a2), 0, a2.shape[0]))
    return a2[0] if single else a2


def hash_no_pad_rows(ctx, rows):
    """H::hash_no_pad of every row of a row-major [nrows, ncols] matrix -> [nrows, 4] (H = ctx.hasher; a BLAKE3 digest is its
    32 bytes as 4 little-endian u64, not field elements)."""
    a = np.ascontiguousarray(rows, dtype=np.uint64)
    out = np.empty((a.shape[0], 4), dtype=np.uint64)
    ctx.check(ctx._lib.ola_hash_rows(ctx.handle, _lib.hptr(a), _lib.hptr(out), 0, a.shape[0], a.shape[1]))
    return out


def merkle_tree(ctx, leaves, cap_height, want_nodes=False):
    """MerkleTree::new_v2(leaves, cap_height) (merkle_tree/mod.rs:180): returns cap [2^h, 4] and, optionally,
    the heap-ordered node array [2*nleaves, 4]."""
    a = np.ascontiguousarray(leaves, dtype=np.uint64)
    n = a.shape[0]
    if n == 0 or n & (n - 1):
        raise ValueError("number of leaves must be a power of 2")
    if (1 << cap_height) > n:
        raise ValueError("cap height should be at most log2(leaves.len())")
    cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
    nodes = np.empty((2 * n, 4), dtype=np.uint64) if want_nodes else None
    ctx.check(ctx._lib.ola_merkle_rows(ctx.handle, _lib.hptr(a), 0, n, a.shape[1], cap_height, _lib.hptr(cap), _lib.hptr(nodes)))
    return (cap, nodes) if want_nodes else cap
