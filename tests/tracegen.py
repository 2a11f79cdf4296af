"""Synthetic VALID traces for the AIR tables (test infrastructure).

The Rust executor cannot run here (no cargo), so tests build small traces that satisfy each table's
constraints by construction, following the reference's own trace generators:
  circuits/src/generation/builtin.rs   generate_cmp_trace, generate_rc_trace :249-316
  circuits/src/stark/lookup.rs:68-131   permuted_cols (Halo2-style permuted input / table columns)
"""
import numpy as np

P = 0xFFFFFFFF00000001


def permuted_cols(inputs, table):
    """lookup.rs:68-131."""
    n = len(inputs)
    si = sorted(int(x) % P for x in inputs)
    st = sorted(int(x) % P for x in table)
    unused_inds, unused_vals = [], []
    perm = [0] * n
    i = j = 0
    while j < n and i < n:
        a, b = si[i], st[j]
        if a > b:
            unused_vals.append(st[j])
            j += 1
        elif a < b:
            if unused_vals:
                perm[i] = unused_vals.pop()
            else:
                unused_inds.append(i)
            i += 1
        else:
            perm[i] = st[j]
            i += 1
            j += 1
    unused_vals.extend(st[j:])
    unused_inds.extend(range(i, n))
    assert len(unused_inds) == len(unused_vals)
    for ind, val in zip(unused_inds, unused_vals):
        perm[ind] = val
    return np.array(si, dtype=np.uint64), np.array(perm, dtype=np.uint64)


def cmp_trace(pairs, log_n):
    """Cmp table (columns.rs:16-22): op0, op1, gte, abs_diff, abs_diff_inv, filter_looking_rc.
    Padding rows (0, 0, 1, 0, 0, 0) satisfy every constraint of cmp_stark.rs:36-44."""
    n = 1 << log_n
    t = np.zeros((6, n), dtype=np.uint64)
    t[2, :] = 1
    for i, (a, b) in enumerate(pairs):
        gte = 1 if a >= b else 0
        d = abs(a - b)
        t[:, i] = [a, b, gte, d, pow(d, P - 2, P) if d else 0, 1]
    return t


def rangecheck_trace(cmp_vals, log_n=16, cpu_vals=(), mem_sort_vals=(), mem_region_vals=()):
    """RangeCheck table (columns.rs:25-39), generate_rc_trace (builtin.rs:249-316)."""
    n = 1 << log_n
    assert n >= 1 << 16
    t = np.zeros((12, n), dtype=np.uint64)
    row = 0
    for col, vals in ((0, cpu_vals), (1, mem_sort_vals), (2, mem_region_vals), (3, cmp_vals)):
        for v in vals:
            t[col, row] = 1
            t[4, row] = v
            t[5, row] = v & 0xFFFF
            t[6, row] = v >> 16
            row += 1
    fix = np.minimum(np.arange(n, dtype=np.uint64), np.uint64(65535))
    t[9] = fix
    t[7], t[10] = permuted_cols(t[5], fix)
    t[8], t[11] = permuted_cols(t[6], fix)
    return t
