"""Mirror of `plonky2_field::cfft` (plonky2/field/src/cfft/mod.rs) over batches of columns.

Same names and argument meaning as the reference; each function takes/returns host numpy uint64 arrays
of shape [ncols, n] (or [n] for a single polynomial) and runs on the GPU through the C ABI.
Error behaviour follows the reference's asserts (cfft/mod.rs:26-41, :76-97): non-power-of-two lengths,
sizes beyond the field's two-adicity and a zero domain offset raise.
"""
import numpy as np

from . import _lib


def _prep(a):
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    single = a.ndim == 1
    if single:
        a = a[None, :]
    n = a.shape[1]
    if n == 0 or n & (n - 1):
        raise ValueError("number of coefficients must be a power of 2")
    return a, single, n.bit_length() - 1


def evaluate_poly(ctx, p):
    """cfft::evaluate_poly (mod.rs:22): coefficients -> evaluations over the subgroup, natural order."""
    a, single, lg = _prep(p)
    ctx.check(ctx._lib.ola_ntt_forward(ctx.handle, _lib.hptr(a), 0, a.shape[0], lg))
    return a[0] if single else a


def interpolate_poly(ctx, evaluations):
    """cfft::interpolate_poly (mod.rs:128): evaluations over the subgroup -> coefficients."""
    a, single, lg = _prep(evaluations)
    ctx.check(ctx._lib.ola_ntt_inverse(ctx.handle, _lib.hptr(a), 0, a.shape[0], lg))
    return a[0] if single else a


def evaluate_poly_with_offset(ctx, p, domain_offset, blowup_factor, natural_order=True):
    """cfft::evaluate_poly_with_offset (mod.rs:65): evaluations over offset * <g_{n*blowup}>."""
    a, single, lg = _prep(p)
    if blowup_factor == 0 or blowup_factor & (blowup_factor - 1):
        raise ValueError("blowup factor must be a power of 2")
    if int(domain_offset) % 0xFFFFFFFF00000001 == 0:
        raise ValueError("domain offset cannot be zero")
    rb = blowup_factor.bit_length() - 1
    out = np.empty((a.shape[0], a.shape[1] << rb), dtype=np.uint64)
    ctx.check(ctx._lib.ola_coset_lde(ctx.handle, _lib.hptr(a), _lib.hptr(out), 0, a.shape[0], lg, rb, int(domain_offset),
                                     1 if natural_order else 0))
    return out[0] if single else out


def interpolate_poly_with_offset(ctx, evaluations, domain_offset):
    """cfft::interpolate_poly_with_offset (mod.rs:180): evaluations over offset * H -> coefficients."""
    a, single, lg = _prep(evaluations)
    if int(domain_offset) % 0xFFFFFFFF00000001 == 0:
        raise ValueError("domain offset cannot be zero")
    ctx.check(ctx._lib.ola_coset_intt(ctx.handle, _lib.hptr(a), 0, a.shape[0], lg, int(domain_offset)))
    return a[0] if single else a
