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


def cpu_padding_trace(log_n):
    """CPU table consisting of padding rows only (generate_cpu_trace with zero steps, generation/cpu.rs:180-208):
    opcode = inst = END (1 << 20), s_end = is_entry_sc = is_next_line_diff_inst = is_padding = 1, the rest 0."""
    n = 1 << log_n
    t = np.zeros((94, n), dtype=np.uint64)
    t[26] = 1 << 20  # COL_INST
    t[28] = 1 << 20  # COL_OPCODE
    t[74] = 1        # COL_S_END
    t[85] = 1        # COL_IS_ENTRY_SC
    t[86] = 1        # COL_IS_NEXT_LINE_DIFF_INST
    t[93] = 1        # COL_IS_PADDING
    return t


def random_binary_filter_trace(rng, ncols, log_n, binary_cols, zero_cols=()):
    """Random (non-satisfying) columns whose CTL filter columns are bits -- for pipeline-parity runs."""
    n = 1 << log_n
    t = rng.integers(0, P, size=(ncols, n), dtype=np.uint64)
    for c in binary_cols:
        t[c] = rng.integers(0, 2, size=n)
    for c in zero_cols:
        t[c] = 0
    return t


# CPU columns that appear in CTL filters (cpu_stark.rs ctl_filter_*): sums of these must stay in {0, 1}
CPU_FILTER_COLS = dict(s_mstore=73, s_mload=72, s_call=70, s_ret=71, tape_looking=88, sccall_ext=89, storage_ext=90, s_bitwise=76, s_gte=78,
                       s_rc=75, s_psdn=79, sccall_end=91, prog_imm=92, is_ext_line=14, is_padding=93)


def cpu_random_trace(rng, log_n):
    """Random CPU-table columns with every CTL filter binary: one-hot over the summed selector groups."""
    n = 1 << log_n
    t = rng.integers(0, P, size=(94, n), dtype=np.uint64)
    f = CPU_FILTER_COLS
    # {mstore, mload} and {call, ret} are summed by their filters: make each pair one-hot-or-zero
    for a, b in ((f["s_mstore"], f["s_mload"]), (f["s_call"], f["s_ret"])):
        pick = rng.integers(0, 3, size=n)
        t[a] = (pick == 1)
        t[b] = (pick == 2)
    for k in ("tape_looking", "sccall_ext", "storage_ext", "s_bitwise", "s_gte", "s_rc", "s_psdn", "sccall_end", "prog_imm"):
        t[f[k]] = rng.integers(0, 2, size=n)
    # filter 1 - is_ext_line - is_padding (ctl_filter_with_program_inst) must be binary too
    pick = rng.integers(0, 3, size=n)
    t[f["is_ext_line"]] = (pick == 1)
    t[f["is_padding"]] = (pick == 2)
    return t


def memory_random_trace(rng, log_n):
    """Random Memory-table columns (29) with binary CTL filters: the 11 op selectors one-hot-or-none (the looked filter
    is the sum of 9 of them, memory_stark.rs:44-57), s_poseidon / filter_looking_rc / filter_looking_rc_cond bits."""
    n = 1 << log_n
    t = rng.integers(0, P, size=(29, n), dtype=np.uint64)
    pick = rng.integers(0, 12, size=n)
    for k in range(11):
        t[6 + k] = (pick == k + 1)
    t[27] = rng.integers(0, 2, size=n)
    t[28] = rng.integers(0, 2, size=n)
    return t


def cmp_random_trace(rng, log_n):
    t = rng.integers(0, P, size=(6, 1 << log_n), dtype=np.uint64)
    t[5] = rng.integers(0, 2, size=1 << log_n)
    return t


def rangecheck_random_trace(rng, log_n=16):
    n = 1 << log_n
    t = rng.integers(0, P, size=(12, n), dtype=np.uint64)
    for c in range(4):
        t[c] = rng.integers(0, 2, size=n)
    return t
