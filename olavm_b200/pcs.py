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

    def __init__(self, ctx, handle, cap, ncols, degree_log, rate_bits, cap_height):
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
    def _commit(cls, ctx, cols, is_coeffs, rate_bits, cap_height, on_device=False, ncols=None, degree_log=None):
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
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        ctx.check(ctx._lib.ola_commit(ctx.handle, ptr, 1 if on_device else 0, ncols, degree_log, 1 if is_coeffs else 0,
                                      rate_bits, cap_height, ctypes.byref(h), _lib.hptr(cap)))
        return cls(ctx, h, cap, ncols, degree_log, rate_bits, cap_height)

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
        L = 1 << (self.degree_log + self.rate_bits)
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
        nsib = self.degree_log + self.rate_bits - self.cap_height
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
