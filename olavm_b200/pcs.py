"""Mirror of `plonky2::fri::oracle::PolynomialBatch` (plonky2/plonky2/src/fri/oracle.rs:31-242) and
`MerkleCap` (hash/merkle_tree/mod.rs:23).  The batch lives in HBM; accessors copy out on demand."""
import ctypes

import numpy as np

from . import _lib


class MerkleCap:
    """MerkleCap(pub Vec<H::Hash>) -- array [2^h, 4] of Goldilocks elements."""

    def __init__(self, hashes):
        self.hashes = np.ascontiguousarray(hashes, dtype=np.uint64).reshape(-1, 4)

    def __len__(self):
        return self.hashes.shape[0]

    def height(self):
        return len(self).bit_length() - 1

    def flatten(self):
        return self.hashes.reshape(-1)


def _reverse_bits(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


class PolynomialBatch:
    """A batch of polynomials committed with a Poseidon Merkle tree over their coset LDE."""

    def __init__(self, ctx, handle, cap, ncols, degree_log, rate_bits, cap_height, coset_first=0, coset_count=None):
        # a coset shard (one rank of a multi-GPU commit) holds leaves [coset_first*n, (coset_first+coset_count)*n)
        self.coset_first = coset_first
        self.coset_count = (1 << rate_bits) if coset_count is None else coset_count
        self.ctx = ctx
        self.handle = handle
        self.merkle_cap = MerkleCap(cap)
        self.ncols = ncols
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.cap_height = cap_height
        self.blinding = False

    # ---- constructors (oracle.rs:45-99)
    @classmethod
    def _commit(cls, ctx, cols, is_coeffs, rate_bits, cap_height, on_device=False, ncols=None, degree_log=None, coset_first=0,
                coset_count=None):
        if on_device:
            ptr = cols
        else:
            cols = np.ascontiguousarray(cols, dtype=np.uint64)
            if cols.ndim != 2:
                raise ValueError("expected [ncols, n] column-major batch")
            ncols, n = cols.shape
            if n == 0 or n & (n - 1):
                raise ValueError("polynomial length must be a power of 2")
            degree_log = n.bit_length() - 1
            ptr = _lib.hptr(cols)
        h = ctypes.c_void_p()
        if coset_count is None:
            cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
            ctx.check(ctx._lib.ola_commit(ctx.handle, ptr, 1 if on_device else 0, ncols, degree_log, 1 if is_coeffs else 0,
                                          rate_bits, cap_height, ctypes.byref(h), _lib.hptr(cap)))
            return cls(ctx, h, cap, ncols, degree_log, rate_bits, cap_height)
        # this rank's cap entries only; `merkle_cap` becomes the full cap after dist.allgather_cap
        cap = np.empty(((coset_count << cap_height) >> rate_bits, 4), dtype=np.uint64)
        ctx.check(ctx._lib.ola_commit_shard(ctx.handle, ptr, 1 if on_device else 0, ncols, degree_log, 1 if is_coeffs else 0,
                                            rate_bits, cap_height, coset_first, coset_count, ctypes.byref(h), _lib.hptr(cap)))
        return cls(ctx, h, cap, ncols, degree_log, rate_bits, cap_height, coset_first, coset_count)

    @classmethod
    def from_values(cls, ctx, values, rate_bits, blinding, cap_height, **kw):
        """PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, ...) (oracle.rs:45)."""
        if blinding:
            raise NotImplementedError("blinding is never used by the OlaVM STARK prover (prover.rs:119, :404, :485)")
        return cls._commit(ctx, values, False, rate_bits, cap_height, **kw)

    @classmethod
    def from_coeffs(cls, ctx, polynomials, rate_bits, blinding, cap_height, **kw):
        """PolynomialBatch::from_coeffs (oracle.rs:66)."""
        if blinding:
            raise NotImplementedError("blinding is never used by the OlaVM STARK prover")
        return cls._commit(ctx, polynomials, True, rate_bits, cap_height, **kw)

    # ---- accessors
    @property
    def polynomials(self):
        """PolynomialBatch.polynomials as [ncols, n] natural-order coefficients."""
        out = np.empty((self.ncols, 1 << self.degree_log), dtype=np.uint64)
        self.ctx.check(self.ctx._lib.ola_batch_get_coeffs(self.ctx.handle, self.handle, _lib.hptr(out)))
        return out

    def leaves(self, first=0, count=None):
        """merkle_tree.leaves[first:first+count] as row-major [count, ncols]."""
        L = self.coset_count << self.degree_log  # leaves held by this batch (all of them unless it is a coset shard)
        if count is None:
            count = L - first
        out = np.empty((count, self.ncols), dtype=np.uint64)
        self.ctx.check(self.ctx._lib.ola_batch_get_leaves(self.ctx.handle, self.handle, first, count, _lib.hptr(out)))
        return out

    def get_lde_values(self, index, step):
        """PolynomialBatch::get_lde_values(index, step) (oracle.rs:132-139)."""
        idx = _reverse_bits(index * step, self.degree_log + self.rate_bits)
        return self.leaves(idx, 1)[0]

    def prove(self, leaf_index):
        """MerkleTree::prove(leaf_index) (merkle_tree/mod.rs:273): sibling digests, bottom-up."""
        nsib = self.degree_log + self.rate_bits - self.cap_height  # paths stop at the cap: identical for a shard
        out = np.empty((max(nsib, 1), 4), dtype=np.uint64)
        k = self.ctx.check(self.ctx._lib.ola_batch_prove_leaf(self.ctx.handle, self.handle, leaf_index, _lib.hptr(out)))
        return out[:k]

    def free(self):
        if self.handle:
            self.ctx.check(self.ctx._lib.ola_batch_free(self.ctx.handle, self.handle))
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass
