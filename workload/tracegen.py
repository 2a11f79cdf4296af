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


def steps_to_records(steps):
    """The executed rows of cpu_vm_trace as the 66-u64 Step records generate_cpu_trace consumes (layout:
    include/ola_gpu.h ola_generate_cpu_trace; core/src/trace/trace.rs Step).  Single-contract runs: env_idx = call_sc_cnt = 0,
    zero storage / code addresses."""
    r = np.zeros((len(steps), 66), dtype=np.uint64)
    for i, s in enumerate(steps):
        r[i, 10], r[i, 11], r[i, 12] = s["tp"], s["clk"], s["pc"]
        r[i, 13], r[i, 14] = s.get("is_ext", 0), s.get("ext_cnt", 0)
        r[i, 15:25] = s["regs"]
        r[i, 25], r[i, 26], r[i, 27], r[i, 28] = s["inst"], s["op1_imm"], s["opcode"], s["imm"]
        r[i, 29], r[i, 30], r[i, 31], r[i, 32], r[i, 33] = s["op0"], s["op1"], s["dst"], s["aux0"], s["aux1"]
        r[i, 34] = s["idx_storage"]
        if s["s_op0"] is not None:
            r[i, 35 + s["s_op0"]] = 1
        if s["s_op1"] is not None:
            r[i, 45 + s["s_op1"]] = 1
        if s["s_dst"] is not None:
            r[i, 55 + s["s_dst"]] = 1
        if "s_op0_0" in s:
            r[i, 35] = s["s_op0_0"]
        for col, v in s.get("sel_raw", {}).items():   # table columns 36..65 -> record fields 35..64
            r[i, col - 1] = v
        r[i, 65] = s.get("filter_tape_looking", 0)
    return r


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


# ======================================================================================================================
# The remaining eight tables: random (pipeline-parity) traces with binary CTL filters, and VALID traces built by
# construction from each table's constraints (circuits/src/builtins/*/, circuits/src/program/*).
# ======================================================================================================================
OP_AND, OP_OR, OP_XOR, OP_POSEIDON = 1 << 18, 1 << 17, 1 << 16, 1 << 12  # core/src/program/binary_program.rs OlaOpcode masks
OP_TLOAD, OP_TSTORE, OP_SCCALL = 1 << 9, 1 << 8, 1 << 7

NCOLS = dict(bitwise=59, poseidon=134, poseidon_chunk=53, storage=48, tape=6, sccall=26, program=18, prog_chunk=40)


def _rand(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


def bitwise_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 59, log_n, [0])


def poseidon_random_trace(rng, log_n):
    n = 1 << log_n
    t = random_binary_filter_trace(rng, 134, log_n, [0, 1])
    pick = rng.integers(0, 3, size=n)  # filter_looked_storage_leaf + _branch is one CTL filter
    t[2] = (pick == 1)
    t[3] = (pick == 2)
    return t


def poseidon_chunk_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 53, log_n, [33, 42, 51] + list(range(43, 51)))


def storage_random_trace(rng, log_n):
    n = 1 << log_n
    t = random_binary_filter_trace(rng, 48, log_n, [44, 45])
    pick = rng.integers(0, 3, size=n)  # is_layer_256 - filter_is_for_prog must be a bit
    t[42] = (pick >= 1)
    t[46] = (pick == 2)
    return t


def tape_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 6, log_n, [5])


def sccall_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 26, log_n, [25])


def program_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 18, log_n, [16, 17])


def prog_chunk_random_trace(rng, log_n):
    return random_binary_filter_trace(rng, 40, log_n, [30, 39] + list(range(31, 39)))


# ---------------------------------------------------------------------------------------------------------- valid traces
def tape_valid_trace(rng, log_n):
    """Tape table (tape/columns.rs:3-9): tx 0 = an init segment (addr 0..3) then tstore / tload / sccall rows; the rest of
    the table is tx 1: one init row then tload repeats of it (every constraint of tape_stark.rs:59-137 holds)."""
    n = 1 << log_n
    assert n >= 16
    rows = []
    for a in range(4):
        rows.append((0, 1, 0, a, int(rng.integers(0, P, dtype=np.uint64)), 0))
    v4, v5 = int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64))
    rows += [(0, 0, OP_TSTORE, 4, v4, 1), (0, 0, OP_TLOAD, 4, v4, 0), (0, 0, OP_SCCALL, 5, v5, 1), (0, 0, OP_TLOAD, 5, v5, 1)]
    w = int(rng.integers(0, P, dtype=np.uint64))
    rows.append((1, 1, 0, 0, w, 0))
    while len(rows) < n:
        rows.append((1, 1, OP_TLOAD, 0, w, 0))
    return np.array(rows, dtype=np.uint64).T.copy()


def sccall_valid_trace(rng, log_n, used=None):
    n = 1 << log_n
    used = n // 2 if used is None else used
    t = np.zeros((26, n), dtype=np.uint64)
    t[:, :used] = _rand(rng, (26, used))
    t[10, :used] = rng.integers(0, 1 << 32, size=used)
    t[11, :used] = rng.integers(0, 1 << 32, size=used)
    t[12, :used] = t[11, :used] + t[10, :used]  # clk_caller_ret = clk_caller_call + op1_imm
    t[25, :used] = 0
    t[25, used:] = 1
    return t


def _horner(vals, beta):
    acc = 0
    for v in reversed(vals):
        acc = (acc * beta + int(v)) % P
    return acc


def program_valid_trace(rng, log_n, beta, prog_rows=None, n_exec=None, exec_rows=None):
    """Program table (program/columns.rs:3-16).  prog_rows: list of (addr0..3, pc, inst) program lines (default random);
    the executed lines are exec_rows when given (generate_prog_trace, generation/prog.rs:56-104: one row per executed
    instruction word and one per immediate), else drawn from the program lines.  comp = sum_i x_i beta^i
    (program_stark.rs:70-88); the permuted columns by lookup.rs permuted_cols."""
    n = 1 << log_n
    if prog_rows is None:
        prog_rows = [tuple(int(x) for x in _rand(rng, 6)) for _ in range(n // 2)]
    assert len(prog_rows) < n
    n_exec = n // 2 if n_exec is None else n_exec
    t = np.zeros((18, n), dtype=np.uint64)
    for i, r in enumerate(prog_rows):
        t[0:6, i] = r
        t[6, i] = _horner(r, beta)
        t[17, i] = 1
    if exec_rows is not None:
        assert len(exec_rows) <= n
        n_exec = len(exec_rows)
    for i in range(n_exec):
        r = exec_rows[i] if exec_rows is not None else prog_rows[int(rng.integers(0, len(prog_rows)))]
        t[8:14, i] = r
        t[14, i] = _horner(r, beta)
        t[16, i] = 1
    t[15], t[7] = permuted_cols(t[14], t[6])
    return t


def bitwise_valid_trace(rng, log_n, beta, n_ops=None, ops=None):
    """Bitwise table (bitwise/columns.rs:23-48): byte-limb decompositions, beta-compressed (tag, a, b, r) byte triples
    looked up in FIX_COMPRESS, byte range checks against FIX_RANGE_CHECK_U8.  ops = [(tag, a, b)] uses the operations of a
    VM run (insert_bitwise_combined, executor lib.rs:1095-1100) instead of random ones."""
    n = 1 << log_n
    assert n >= 256
    n_ops = (min(n // 8, 60) if n_ops is None else n_ops) if ops is None else len(ops)
    t = np.zeros((59, n), dtype=np.uint64)
    fixed = {(0, 0, 0, 0)}
    for i in range(n_ops):
        if ops is not None:
            tag, a, b = ops[i]
        else:
            tag = [OP_AND, OP_OR, OP_XOR][int(rng.integers(0, 3))]
            a, b = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
        r = a & b if tag == OP_AND else (a | b if tag == OP_OR else a ^ b)
        t[0:5, i] = [1, tag, a, b, r]
        for j in range(4):
            la, lb, lr = (a >> 8 * j) & 255, (b >> 8 * j) & 255, (r >> 8 * j) & 255
            t[5 + j, i], t[9 + j, i], t[13 + j, i] = la, lb, lr
            t[29 + j, i] = (tag + beta * la + beta * beta % P * lb + pow(beta, 3, P) * lr) % P
            fixed.add((tag, la, lb, lr))
    assert len(fixed) <= n
    fix = np.minimum(np.arange(n, dtype=np.uint64), np.uint64(255))
    t[37] = fix
    for j in range(4):
        t[17 + j], t[38 + j] = permuted_cols(t[5 + j], fix)
        t[21 + j], t[42 + j] = permuted_cols(t[9 + j], fix)
        t[25 + j], t[46 + j] = permuted_cols(t[13 + j], fix)
    for i, (tag, la, lb, lr) in enumerate(sorted(fixed)):
        t[50:54, i] = [tag, la, lb, lr]
        t[54, i] = (tag + beta * la + beta * beta % P * lb + pow(beta, 3, P) * lr) % P
    for j in range(4):
        t[33 + j], t[55 + j] = permuted_cols(t[29 + j], t[54])
    return t


def poseidon_valid_trace(orc, log_n, rows):
    """Poseidon table from (input[12], filters[4]) pairs; padding = the zero-input row (generation/poseidon.rs:83-126)."""
    n = 1 << log_n
    assert len(rows) <= n
    t = np.zeros((134, n), dtype=np.uint64)
    t[:, :] = orc.poseidon_table_row(np.zeros(12, dtype=np.uint64))[:, None]
    for i, (inp, filt) in enumerate(rows):
        r = orc.poseidon_table_row(np.array(inp, dtype=np.uint64))
        r[0:4] = filt
        t[:, i] = r
    return t


def prog_chunk_valid_trace(orc, rng, log_n, line_counts=(1, 3, 2), programs=None):
    """ProgChunk table (program/columns.rs:47-62): per program, lines of 8 instructions absorbed by a Poseidon sponge
    (cap = previous line's hash[8..12]).  Returns (trace, poseidon_rows, program_lines, prog_hashes):
    poseidon_rows = the (input, output) pairs the lines look up; program_lines = (addr0..3, pc, inst) with filter 1."""
    n = 1 << log_n
    t = np.zeros((40, n), dtype=np.uint64)
    t[39] = 1
    row = 0
    psdn, lines, roots = [], [], []
    # programs = [(code address [4], instruction words)]: real programs instead of random lines
    todo = [(None, None, L) for L in line_counts] if programs is None else [(a, w, (len(w) + 7) // 8) for a, w in programs]
    for paddr, pwords, L in todo:
        addr = [int(x) for x in _rand(rng, 4)] if paddr is None else [int(x) for x in paddr]
        cap = [0, 0, 0, 0]
        h = [0] * 12
        for j in range(L):
            last = j == L - 1
            if pwords is None:
                cnt = int(rng.integers(1, 9)) if last else 8
                inst = [int(x) for x in _rand(rng, 8)]
            else:
                cnt = min(8, len(pwords) - 8 * j)
                inst = [int(x) for x in pwords[8 * j:8 * j + cnt]] + [0] * (8 - cnt)
            for k in range(cnt, 8):
                inst[k] = h[k]  # overwrite-mode sponge: the unused slots keep the previous line's output (prog.rs:212-216)
            h = [int(x) for x in orc.poseidon(np.array(inst + cap, dtype=np.uint64))]
            t[0:4, row] = addr
            t[4, row] = 8 * j
            t[5:13, row] = inst
            t[13:17, row] = cap
            t[17:29, row] = h
            t[29, row] = 1 if j == 0 else 0
            t[30, row] = 1 if last else 0
            t[31:39, row] = [1 if k < cnt else 0 for k in range(8)]
            t[39, row] = 0
            psdn.append((inst + cap, h))
            for k in range(cnt):
                lines.append(tuple(addr) + (8 * j + k, inst[k]))
            cap = h[8:12]
            if last:
                roots.append((addr, h[0:4]))
            row += 1
    assert row <= n
    return t, psdn, lines, roots


def poseidon_chunk_valid_trace(orc, rng, log_n, lengths=(5, 8, 19)):
    """PoseidonChunk table (poseidon/columns.rs:44-71): one main line + ceil(len/8) ext lines per hash call.
    Returns (trace, poseidon_rows)."""
    n = 1 << log_n
    t = np.zeros((53, n), dtype=np.uint64)
    t[52] = 1
    row = 0
    psdn = []
    for ci, L in enumerate(lengths):
        tx, env, clk, op0, dst = 0, 0, 10 + 7 * ci, 1000 + 64 * ci, 5000 + 16 * ci
        t[0:8, row] = [tx, env, clk, OP_POSEIDON, op0, L, dst, 0]
        t[42, row] = 1
        t[52, row] = 0
        row += 1
        vals = [int(x) for x in _rand(rng, L)]
        cap, acc = [0, 0, 0, 0], 0
        n_ext = (L + 7) // 8
        for e in range(n_ext):
            cnt = min(8, L - 8 * e)
            v = vals[8 * e : 8 * e + cnt] + [0] * (8 - cnt)
            h = [int(x) for x in orc.poseidon(np.array(v + cap, dtype=np.uint64))]
            acc += cnt
            t[0:8, row] = [tx, env, clk, OP_POSEIDON, op0 + 8 * e, L, dst, acc]
            t[8:16, row] = v
            t[16:20, row] = cap
            t[20:32, row] = h
            t[32, row] = 1
            t[33, row] = 1 if e == n_ext - 1 else 0
            if cnt < 8:
                t[34 + cnt, row] = 1
            t[43:51, row] = [1 if k < cnt else 0 for k in range(8)]
            t[51, row] = 1
            t[52, row] = 0
            psdn.append((v + cap, h))
            cap = h[8:12]
            row += 1
    assert row <= n
    return t, psdn


def storage_valid_trace(orc, rng, log_n, accesses):
    """StorageAccess table (storage/columns.rs:3-33): 256 rows per access walking the sparse Merkle tree from layer 1
    (root) to layer 256 (leaf).  accesses: list of dict(addr_bits=[256 bits, layer 1 first], leaf=[4], pre_leaf=[4],
    is_write, for_prog).  Hashes are real Poseidon hashes of (path, sib | hash_type) in bit order, so the rows are
    consistent with ctl_storage_access_poseidon.  Returns (trace, poseidon_rows) with poseidon_rows =
    (input[12], output[12], is_leaf) for every looked-up hash (current and pre)."""
    n = 1 << log_n
    assert 256 * len(accesses) <= n
    t = np.zeros((48, n), dtype=np.uint64)
    psdn = []
    prev_root = None
    row0 = 0
    for ai, acc in enumerate(accesses):
        bits = acc["addr_bits"]
        sib = acc["sib"] if "sib" in acc else [[int(x) for x in _rand(rng, 4)] for _ in range(256)]
        # bottom-up: path_l = hash of the child at layer l + 1 (the leaf value at layer 256)
        path = [None] * 257
        pre_path = [None] * 257
        hsh = [None] * 257
        pre_hsh = [None] * 257
        path[256], pre_path[256] = list(acc["leaf"]), list(acc["pre_leaf"])
        for l in range(256, 0, -1):
            typ = 1 if l == 256 else 0
            for cur, pth, out in ((True, path, hsh), (False, pre_path, pre_hsh)):
                inp = (pth[l] + sib[l - 1] if bits[l - 1] == 0 else sib[l - 1] + pth[l]) + [typ, 0, 0, 0]
                o = [int(x) for x in orc.poseidon(np.array(inp, dtype=np.uint64))]
                out[l] = o[0:4]
                psdn.append((inp, o, l == 256))
            if l > 1:
                path[l - 1], pre_path[l - 1] = hsh[l], pre_hsh[l]
        root, pre_root = hsh[1], pre_hsh[1]
        if prev_root is not None:
            assert pre_root == prev_root, "accesses must chain: pre_root of an access is the previous access's root"
        prev_root = root
        limbs = []
        for k in range(4):
            v = 0
            for b in bits[64 * k : 64 * k + 64]:
                v = (2 * v + b) % P
            limbs.append(v)
        accv, marker = 0, 0
        for l in range(1, 257):
            r = row0 + l - 1
            b = bits[l - 1]
            accv = b if l % 64 == 1 else (2 * accv + b) % P
            marker += 1 if l in (1, 64, 128, 192, 256) else 0
            t[0, r] = acc.get("idx", ai + 1)
            t[1:5, r] = pre_root
            t[5:9, r] = root
            t[9, r] = acc["is_write"]
            t[10, r] = l
            t[11, r] = b
            t[12, r] = accv
            t[13:17, r] = limbs
            t[17:21, r] = pre_path[l]
            t[21:25, r] = path[l]
            t[25:29, r] = sib[l - 1]
            t[29, r] = 1 if l == 256 else 0
            t[30:34, r] = pre_hsh[l]
            t[34:38, r] = hsh[l]
            for k, ll in enumerate((1, 64, 128, 192, 256)):
                t[38 + k, r] = 1 if l == ll else 0
            t[43, r] = marker
            t[44, r] = 1 - b
            t[45, r] = b
            t[46, r] = 1 if (l == 256 and acc.get("for_prog")) else 0
        row0 += 256
    for r in range(row0, n):
        t[47, r] = 1
        t[5:9, r] = prev_root
    return t, psdn


class SparseMerkleTree:
    """The account tree the storage opcodes walk, as far as the StorageAccess AIR sees it: depth 256, a node is
    Poseidon(left | right | hash_type, 0, 0, 0)[0..4] with hash_type 1 when the children are leaves (layer 256) and 0
    above, an absent leaf is [0; 4] (tree_key_default).  Only the populated paths are materialised."""

    def __init__(self, orc):
        self.orc, self.leaves = orc, {}
        self.default = [None] * 257
        self.default[256] = [0, 0, 0, 0]
        for d in range(255, -1, -1):
            self.default[d] = self._h(self.default[d + 1], self.default[d + 1], d + 1 == 256)

    def _h(self, left, right, children_are_leaves):
        inp = list(left) + list(right) + [1 if children_are_leaves else 0, 0, 0, 0]
        return [int(x) for x in self.orc.poseidon(np.array(inp, dtype=np.uint64))[:4]]

    def _subtree(self, depth, prefix, keys):
        if not keys:
            return self.default[depth]
        if depth == 256:
            return self.leaves[keys[0]]
        zero = [k for k in keys if k[depth] == 0]
        one = [k for k in keys if k[depth] == 1]
        return self._h(self._subtree(depth + 1, prefix + (0,), zero), self._subtree(depth + 1, prefix + (1,), one), depth + 1 == 256)

    def root(self):
        return self._subtree(0, (), list(self.leaves))

    def siblings(self, bits):
        """sib[l - 1] = the sibling of the path node at layer l = 1..256 (layer 256 = the leaves)."""
        bits = tuple(bits)
        out = []
        for l in range(1, 257):
            pre = bits[: l - 1] + (1 - bits[l - 1],)
            out.append(self._subtree(l, pre, [k for k in self.leaves if k[:l] == pre]))
        return out

    def set(self, bits, leaf):
        self.leaves[tuple(bits)] = list(leaf)


def tree_key_bits(tree_key):
    """TreeKey (4 field elements) -> the 256 path bits, layer 1 first: limb k is bits [64k, 64k + 64), most significant first."""
    return [(int(tree_key[k]) >> (63 - j)) & 1 for k in range(4) for j in range(64)]


def storage_tables_from_log(orc, rng, st_log, extra_accesses=()):
    """StorageAccess table + its Poseidon rows for the sstore / sload accesses a VM run logged (cpu_vm_trace, st_log), in
    access order (storage_access_idx 1, 2, ...), walking ONE consistent sparse Merkle tree: every access's pre_root is the
    previous access's root.  extra_accesses (e.g. ProgChunk's code-root read, for_prog=1) follow the run's.  Returns
    (table, poseidon rows as (input, filters)) with the tree-key hashes (filter_looked_treekey) first."""
    tree = SparseMerkleTree(orc)
    accesses, rows = [], []
    for a in st_log:
        bits = tree_key_bits(a["tree_key"])
        assert tree.leaves.get(tuple(bits), [0, 0, 0, 0]) == list(a["pre_leaf"])
        sib = tree.siblings(bits)
        accesses.append(dict(idx=a["idx"], addr_bits=bits, leaf=a["leaf"], pre_leaf=a["pre_leaf"], is_write=a["is_write"], sib=sib))
        if a["is_write"]:
            tree.set(bits, a["leaf"])
        rows.append((a["hash_input"], [0, 1, 0, 0]))
    for k, a in enumerate(extra_accesses):
        bits = a["addr_bits"]
        accesses.append(dict(a, idx=len(st_log) + k + 1, sib=tree.siblings(bits)))
    log_n = max(8, (256 * len(accesses)).bit_length() - (1 if (256 * len(accesses)) & (256 * len(accesses) - 1) == 0 else 0))
    st, psdn_st = storage_valid_trace(orc, rng, log_n, accesses)
    rows += [(inp, [0, 0, 1, 0] if is_leaf else [0, 0, 0, 1]) for inp, _, is_leaf in psdn_st]
    return st, rows


def hash_system_valid(orc, rng, beta=0x1234567890ABCDEF % P):
    """A VALID five-table system [Poseidon, PoseidonChunk, StorageAccess, Program, ProgChunk] whose cross-table lookups
    ctl_chunk_poseidon, ctl_storage_access_poseidon, ctl_prog_chunk_prog and ctl_prog_chunk_storage (ola_stark.rs:358-379,
    :388-413, :530-563) are complete and consistent: one program of three lines is hashed by ProgChunk, its lines are the
    Program table, its root is read from the storage tree (256-layer Merkle walk), two Poseidon calls go through
    PoseidonChunk, and every sponge / Merkle hash is a row of the Poseidon table.
    Returns (table_ids, traces, compress_challenges)."""
    pc, psdn_prog, lines, roots = prog_chunk_valid_trace(orc, rng, 2, line_counts=(3,))
    pch, psdn_chunk = poseidon_chunk_valid_trace(orc, rng, 3, lengths=(5, 16))
    addr, leaf = roots[0]
    bits = []
    for limb in addr:
        bits += [(int(limb) >> (63 - j)) & 1 for j in range(64)]
    st, psdn_st = storage_valid_trace(orc, rng, 8, [dict(addr_bits=bits, leaf=leaf, pre_leaf=leaf, is_write=0, for_prog=1)])
    rows = [(inp, [1, 0, 0, 0]) for inp, _ in psdn_prog + psdn_chunk]
    rows += [(inp, [0, 0, 1, 0] if is_leaf else [0, 0, 0, 1]) for inp, _, is_leaf in psdn_st]
    ps = poseidon_valid_trace(orc, 10, rows)
    prog = program_valid_trace(rng, 5, beta, prog_rows=lines, n_exec=7)
    ids = [5, 6, 7, 10, 11]
    return ids, [ps, pch, st, prog, pc], [0, 0, 0, beta, 0]


# ---------------------------------------------------------------------------------------------------------------------
# A small Ola VM for the CPU table: real (non-padding) rows produced the way the reference produces them.
#   executor/src/lib.rs   Process::execute :2074-2310 and the per-opcode handlers :572-814 (mov/not, eq/neq, assert,
#                         cjmp, jmp, add/mul), execute_inst_end :1186-1262, Process::new :249-283
#   core/src/program/binary_program.rs:100-192  instruction word: opcode one-hot (bits 6..31), dst / op1 / op0 register
#                         one-hots at bits 32+i / 42+i / 52+i, bit 62 = "op1 is an immediate" (the immediate is the next word)
#   circuits/src/generation/cpu.rs:11-218       Step -> row, padding
# Only register-to-register opcodes are modelled (no memory, storage, tape or builtin lookups), which is what a CPU-only
# proof (or CPU + lookup-free Cmp / RangeCheck) can check; the rows do NOT come from the AIR transcription, so proving
# them with the quotient-degree check on tests the transcription against the reference's own trace semantics.
# ---------------------------------------------------------------------------------------------------------------------
OPCODE_SHIFT = {"add": 31, "mul": 30, "eq": 29, "assert": 28, "mov": 27, "jmp": 26, "cjmp": 25, "call": 24, "ret": 23, "mload": 22,
                "mstore": 21, "end": 20, "range": 19, "and": 18, "or": 17, "xor": 16, "not": 15, "neq": 14, "gte": 13, "poseidon": 12, "sload": 11,
                "sstore": 10, "tload": 9, "tstore": 8}
CPU_SELECTOR_COL = {"add": 66, "mul": 66, "eq": 66, "assert": 66, "neq": 66, "mov": 67, "jmp": 68, "cjmp": 69, "call": 70, "ret": 71,
                    "mload": 72, "mstore": 73, "end": 74, "range": 75, "and": 76, "or": 76, "xor": 76, "not": 77, "gte": 78, "poseidon": 79, "sload": 80, "sstore": 81,
                    "tload": 82, "tstore": 83}


def _finv(x):
    return pow(x, P - 2, P)


def _reg(s):
    assert isinstance(s, str) and s[0] == "r" and 0 <= int(s[1:]) < 10, s
    return int(s[1:])


def ola_encode(ins):
    """(op, operands...) -> [instruction word] or [instruction word, immediate].  Operand order as in the assembly text:
    add/mul/eq/neq/gte dst op0 op1 | mov/not dst op1 | cjmp op0 op1 | jmp/call/assert/range op1 | ret | end |
    mstore base offset value-reg, mload dst base offset: (anchor, offset, dst) = (op0 register, op1 immediate, dst register),
    assembler/src/encoder.rs:123-213."""
    op = ins[0]
    word = 1 << OPCODE_SHIFT[op]
    dst = op0 = op1 = None
    if op in ("add", "mul", "eq", "neq", "gte", "and", "or", "xor", "poseidon", "tload"):
        dst, op0, op1 = ins[1], ins[2], ins[3]
    elif op in ("tstore", "sstore", "sload"):
        op0, op1 = ins[1], ins[2]
    elif op in ("mov", "not"):
        dst, op1 = ins[1], ins[2]
    elif op == "cjmp":
        op0, op1 = ins[1], ins[2]
    elif op in ("jmp", "assert", "call", "range"):
        op1 = ins[1]
    elif op == "mstore":
        op0, op1, dst = ins[1], ins[2], ins[3]
    elif op == "mload":
        dst, op0, op1 = ins[1], ins[2], ins[3]
    else:
        assert op in ("end", "ret")
    imm = None
    if op in ("mstore", "mload") and isinstance(op1, tuple):
        # [anchor, offset register, factor]: op1 = the offset register, the immediate word carries the factor and the
        # op1_imm flag stays 0 (OlaOperand::RegisterWithFactor, core/src/program/binary_program.rs:150-153)
        reg, factor = op1
        word |= (1 << (32 + _reg(dst))) | (1 << (52 + _reg(op0))) | (1 << (42 + _reg(reg)))
        return [word, int(factor) % P]
    if op in ("mstore", "mload"):
        op1 = int(op1)
    if dst is not None:
        word |= 1 << (32 + _reg(dst))
    if op0 is not None:
        word |= 1 << (52 + _reg(op0))
    if op1 is not None:
        if op1 == "psp":
            pass  # the prophet stack pointer: neither an op1 register bit nor the immediate flag (binary_program.rs:286-290)
        elif isinstance(op1, str):
            word |= 1 << (42 + _reg(op1))
        else:
            word |= 1 << 62
            imm = int(op1) % P
    return [word] if imm is None else [word, imm]


PSP_START_ADDR = P - 0xFFFFFFFF        # core/src/vm/memory.rs:8-10: prophet (write-once) region [p - span, p), heap [p - 2 span, p - span)
HP_START_ADDR = P - 2 * 0xFFFFFFFF


def cpu_vm_trace(program, log_n, max_steps=1 << 20, want_side_tables=False, orc=None, init_tape=(), prophets=None):
    """Run `program` (a list of instruction tuples, jump targets = word addresses) and return the [94][2^log_n] CPU table
    and the executed steps; with want_side_tables also the (op0, op1) pairs of the gte rows (Cmp table), their
    |op0 - op1| (RangeCheck rows looked by Cmp) and the operands of the range rows (RangeCheck rows looked by the CPU)."""
    cmp_pairs, rc_cmp, rc_cpu, mem, mem_log, bit_ops, psdn_calls = [], [], [], {}, [], [], []
    tp, tape, tape_log = 0, {}, {}   # tape pointer, tape contents, per-address access log (gen_tape_table order)
    st_idx, st_cache, st_log = 0, {}, []   # storage_access_idx, tx storage cache (tree key -> value), access log
    psp = psp_start = PSP_START_ADDR       # Process::new (executor/src/lib.rs:267-269); hp starts one past the heap-pointer cell
    hp = HP_START_ADDR + 1
    mem[HP_START_ADDR] = HP_START_ADDR + 1   # "init heap ptr" (lib.rs:2088-2100); gen_memory_table drops that row (trace.rs:33-38), the
    prophets = prophets or {}               # Memory AIR pins the cell's first value instead (memory_stark.rs, ADDR_HEAP_PTR)
    for v in init_tape:                    # init_tape (executor/src/load_tx.rs:89-132): tx context, calldata, addresses; is_init cells
        tape[tp] = int(v) % P
        tape_log[tp] = [(1, 0, tape[tp], 0)]
        tp += 1
    words, at_pc = [], {}
    for ins in program:
        enc = ola_encode(ins)
        at_pc[len(words)] = (ins, enc)
        words += enc
    regs = [0] * 10
    pc, clk, steps = 0, 0, []
    while True:
        assert pc in at_pc, f"pc {pc} is not an instruction boundary"
        ins, enc = at_pc[pc]
        op, step = ins[0], len(enc)
        row = {"clk": clk, "pc": pc, "tp": tp, "regs": list(regs), "inst": enc[0], "imm": enc[1] if step == 2 else 0,
               "op1_imm": 1 if step == 2 else 0, "opcode": 1 << OPCODE_SHIFT[op], "op": op,
               "op0": 0, "op1": 0, "dst": 0, "aux0": 0, "aux1": 0, "s_op0": None, "s_op1": None, "s_dst": None, "idx_storage": st_idx}

        def val(x):  # get_index_value (lib.rs:297-320)
            if x == "psp":
                return psp_start
            if isinstance(x, str):
                row["s_op1"] = _reg(x)
                return regs[_reg(x)]
            return int(x) % P


        def run_prophet(at_pc_):  # Process::prophet (lib.rs:369-444) for the one built-in the prophet-using test programs share
            nonlocal psp, psp_start, hp
            spec = prophets.get(at_pc_)
            if spec is None:
                return
            if spec["fn"] == "printf":         # no outputs: the interpreter returns only the heap pointer; psp_start catches up
                psp_start = psp
                return
            if spec["fn"] in ("mod", "div", "split_hi", "split_lo"):   # one-line integer helpers written in the prophet language itself
                x, y = regs[1], regs[2]                                  # inputs: r1, r2 (read_prophet_input, lib.rs:322-367)
                v = {"mod": lambda: x % y, "div": lambda: x // y, "split_hi": lambda: x >> 32, "split_lo": lambda: x & 0xFFFFFFFF}[spec["fn"]]()
                psp_start = psp
                mem[psp] = v
                mem_log.append((psp, 0, 0, 1, v))
                psp += 1
                return
            assert spec["fn"] == "malloc" and spec["inputs"] == 1, "only malloc, printf and the integer helpers are modelled"
            ln = regs[1]                       # the first prophet input comes from r1 (PROPHET_INPUT_REG_START_INDEX)
            hp = (hp + ln) % P                 # travel_malloc (interpreter/src/interpreter/executor.rs:656-671): hp += len, returns the NEW hp
            psp_start = psp
            mem[psp] = hp                      # outputs go to the write-once region at psp, clk 0, opcode 0
            mem_log.append((psp, 0, 0, 1, hp))
            psp += 1

        if op == "end":
            steps.append(row)
            break
        if op in ("mov", "not"):
            v = val(ins[2])
            row["op1"] = v
            regs[_reg(ins[1])] = v if op == "mov" else (P - 1 - v) % P
            row["dst"], row["s_dst"] = regs[_reg(ins[1])], _reg(ins[1])
            pc += step
        elif op in ("add", "mul", "eq", "neq"):
            a = regs[_reg(ins[2])]
            row["op0"], row["s_op0"] = a, _reg(ins[2])
            b = val(ins[3])
            row["op1"] = b
            if op == "add":
                r = (a + b) % P
            elif op == "mul":
                r = (a * b) % P
            else:
                d = (a - b) % P
                row["aux0"] = _finv(d) if d else 0
                r = int(a == b) if op == "eq" else int(a != b)
            regs[_reg(ins[1])] = r
            row["dst"], row["s_dst"] = r, _reg(ins[1])
            pc += step
        elif op in ("sstore", "sload"):  # execute_inst_sstore / _sload, lib.rs:1263-1403, :1405-1530 (contract address 0)
            assert orc is not None, "storage opcodes hash the tree key: pass orc"
            mask = 1 << OPCODE_SHIFT[op]
            key_addr = regs[_reg(ins[1])]
            row["op0"], row["s_op0"] = key_addr, _reg(ins[1])
            val_addr = val(ins[2])
            row["op1"] = val_addr
            e = dict(row)                      # the ext line (aux_insert!, lib.rs:128-150): same clk / pc / instruction / registers,
            e["regs"] = list(row["regs"])      # a fresh RegisterSelector holding the operands, the 4 + 4 memory cells and the tree key
            e["is_ext"], e["ext_cnt"], e["ext_len"], e["s_op0"], e["s_op1"], e["s_dst"] = 1, 1, 1, None, None, None
            slot_key = [mem[(key_addr + i) % P] for i in range(4)]
            for i in range(4):                 # sstore interleaves key / value reads, sload reads the key first: same log per cell
                mem_log.append(((key_addr + i) % P, clk, mask, 0, slot_key[i]))
                if op == "sstore":
                    mem_log.append(((val_addr + i) % P, clk, mask, 0, mem[(val_addr + i) % P]))
            hin = [0, 0, 0, 0] + slot_key + [0, 0, 0, 0]   # StorageKey::raw_hashed_key (core/src/types/storage/mod.rs:37-46)
            tree_key = [int(x) for x in orc.poseidon(np.array(hin, dtype=np.uint64))[:4]]
            pre = st_cache.get(tuple(tree_key), [0, 0, 0, 0])
            if op == "sstore":
                value = [mem[(val_addr + i) % P] for i in range(4)]
                st_cache[tuple(tree_key)] = value
            else:
                value = pre
                for i in range(4):
                    mem[(val_addr + i) % P] = value[i]
                    mem_log.append(((val_addr + i) % P, clk, mask, 1, value[i]))
            st_idx += 1
            e["idx_storage"] = st_idx
            raw = {}
            for i in range(4):
                raw[36 + i], raw[36 + 4 + i] = (key_addr + i) % P, slot_key[i]
                raw[46 + i], raw[46 + 4 + i] = (val_addr + i) % P, value[i]
                raw[56 + i] = tree_key[i]
            e["sel_raw"] = raw
            st_log.append(dict(idx=st_idx, is_write=int(op == "sstore"), tree_key=tree_key, pre_leaf=pre, leaf=value, hash_input=hin))
            row["ext_len"] = 1
            steps.append(row)
            steps.append(e)
            pc += step
            clk += 1
            continue
        elif op in ("tstore", "tload"):  # execute_inst_tstore / _tload + tape_copy!, lib.rs:153-181, :1687-1846
            ext_rows = []
            if op == "tstore":   # copy `len` memory words at [op0] to the tape at tp, then tp += len
                base = regs[_reg(ins[1])]
                row["op0"], row["s_op0"] = base, _reg(ins[1])
                ln = val(ins[2])
                row["op1"] = ln
                tape_base, mask = tp, 1 << 8
            else:                # tload dst flag op1: flag 1 -> the last `op1` tape words, flag 0 -> the single word at tape address op1
                base = regs[_reg(ins[1])]
                row["dst"], row["s_dst"] = base, _reg(ins[1])
                flag = regs[_reg(ins[2])]
                row["aux1"], row["s_op0"] = flag, _reg(ins[2])
                v1 = val(ins[3])
                row["op1"] = v1
                assert flag in (0, 1), "TloadFlagInvalid"
                row["op0"] = flag
                tape_base, ln, mask = ((tp - v1) % P, v1, 1 << 9) if flag == 1 else (v1, 1, 1 << 9)
            for k in range(ln):
                maddr, taddr = (base + k) % P, tape_base + k
                e = dict(row)
                e["regs"] = list(row["regs"])
                e["is_ext"], e["ext_cnt"], e["filter_tape_looking"] = 1, k + 1, 1
                e["aux0"], e["s_op0_0"] = maddr, taddr
                if op == "tstore":
                    v = mem[maddr]
                    mem_log.append((maddr, clk, mask, 0, v))
                    tape[taddr] = v
                    tape_log.setdefault(taddr, []).append((0, mask, v, 1))
                else:
                    v = tape[taddr]
                    tape_log[taddr].append((tape_log[taddr][-1][0], mask, v, 1))
                    mem[maddr] = v
                    mem_log.append((maddr, clk, mask, 1, v))
                e["aux1"] = v
                ext_rows.append(e)
            row["ext_len"] = ln
            for e in ext_rows:
                e["ext_len"] = ln
            steps.append(row)
            steps.extend(ext_rows)
            if op == "tstore":
                tp += ln
            pc += step
            clk += 1
            continue
        elif op == "poseidon":  # execute_inst_poseidon, lib.rs:1547-1685: hash `len` memory words at [op0] into 4 words at [dst]
            assert orc is not None, "the poseidon opcode needs the oracle's permutation"
            src = regs[_reg(ins[2])]
            row["op0"], row["s_op0"] = src, _reg(ins[2])
            ln = val(ins[3])
            row["op1"] = ln
            dst_addr = regs[_reg(ins[1])]
            row["dst"], row["s_dst"] = dst_addr, _reg(ins[1])
            assert ln > 0
            chunk_rows = [dict(op0=src, acc=0, value=[0] * 8, cap=[0] * 4, hash=[0] * 12, ext=0)]  # the main line
            state, hash_pre, read_ptr, perm_rows = [0] * 12, [0] * 12, 0, []
            while True:
                cnt = min(8, ln - read_ptr)
                if cnt <= 0:
                    break
                for k in range(cnt):
                    state[k] = mem[(src + read_ptr + k) % P]
                    mem_log.append(((src + read_ptr + k) % P, clk, 1 << 12, 0, state[k]))
                out = [int(x) for x in orc.poseidon(np.array(state, dtype=np.uint64))]
                perm_rows.append((list(state), out))
                chunk_rows.append(dict(op0=(src + read_ptr) % P, acc=read_ptr + cnt, value=list(state[0:8]), cap=list(hash_pre[8:12]),
                                       hash=out, ext=1))
                hash_pre = out
                read_ptr += cnt
                if read_ptr + 8 > ln:   # the next chunk is the (possibly empty) tail: positions past it keep the previous output
                    tail = ln - read_ptr
                    state = list(out) if tail == 0 else state[:0] + [0] * tail + out[tail:]
                else:
                    state = [0] * 8 + out[8:12]
            for k in range(4):
                mem[(dst_addr + k) % P] = hash_pre[k]
                mem_log.append(((dst_addr + k) % P, clk, 1 << 12, 1, hash_pre[k]))
            psdn_calls.append(dict(clk=clk, dst=dst_addr, op0=src, op1=ln, rows=chunk_rows, perms=perm_rows))
            pc += step
        elif op in ("and", "or", "xor"):  # execute_inst_bitwise, lib.rs:1041-1105
            a = regs[_reg(ins[2])]
            row["op0"], row["s_op0"] = a, _reg(ins[2])
            b = val(ins[3])
            row["op1"] = b
            assert a < (1 << 32) and b < (1 << 32), "the Bitwise table works on u32 operands"
            r = a & b if op == "and" else (a | b if op == "or" else a ^ b)
            regs[_reg(ins[1])] = r
            row["dst"], row["s_dst"] = r, _reg(ins[1])
            bit_ops.append((1 << OPCODE_SHIFT[op], a, b))
            pc += step
        elif op == "gte":  # execute_inst_gte, lib.rs:1107-1185
            a = regs[_reg(ins[2])]
            row["op0"], row["s_op0"] = a, _reg(ins[2])
            b = val(ins[3])
            row["op1"] = b
            r = int(a >= b)
            d = (a - b) % P if r else (b - a) % P
            assert d <= 0xFFFFFFFF, "U32RangeCheckFail"
            regs[_reg(ins[1])] = r
            row["dst"], row["s_dst"] = r, _reg(ins[1])
            cmp_pairs.append((a, b))
            rc_cmp.append(d)
            pc += step
        elif op == "range":  # execute_inst_range, lib.rs:998-1039
            v = regs[_reg(ins[1])]
            assert v <= 0xFFFFFFFF, "U32RangeCheckFail"
            row["op1"], row["s_op1"] = v, _reg(ins[1])
            rc_cpu.append(v)
            pc += step
        elif op == "mstore":  # execute_inst_mstore, lib.rs:868-933
            base = regs[_reg(ins[1])]
            if isinstance(ins[2], tuple):  # [anchor, reg, factor] (ops.len() == 5): addr = anchor + factor * reg, op1_imm = 0
                reg, factor = ins[2]
                row["op1"], row["s_op1"], row["aux0"], row["op1_imm"] = regs[_reg(reg)], _reg(reg), int(factor) % P, 0
                off = row["aux0"] * row["op1"] % P
                row["op0"], row["s_op0"] = base, _reg(ins[1])
            else:                          # [anchor, offset] (ops.len() == 4): addr = anchor + offset, op1_imm = 1
                off = int(ins[2]) % P
                row["op0"], row["s_op0"], row["op1"] = base, _reg(ins[1]), off
            v = regs[_reg(ins[3])]
            row["dst"], row["s_dst"] = v, _reg(ins[3])
            row["aux1"] = (base + off) % P
            mem[row["aux1"]] = v
            mem_log.append((row["aux1"], clk, 1 << 21, 1, v))
            pc += step
        elif op == "mload":  # execute_inst_mload, lib.rs:935-996 (the two operand forms as for mstore)
            base = regs[_reg(ins[2])]
            if isinstance(ins[3], tuple):
                reg, factor = ins[3]
                row["op1"], row["s_op1"], row["aux0"], row["op1_imm"] = regs[_reg(reg)], _reg(reg), int(factor) % P, 0
                off = row["aux0"] * row["op1"] % P
                row["op0"], row["s_op0"] = base, _reg(ins[2])
            else:
                off = int(ins[3]) % P
                row["op0"], row["s_op0"], row["op1"] = base, _reg(ins[2]), off
            row["aux1"] = (base + off) % P
            regs[_reg(ins[1])] = mem[row["aux1"]]
            row["dst"], row["s_dst"] = regs[_reg(ins[1])], _reg(ins[1])
            mem_log.append((row["aux1"], clk, 1 << 22, 0, regs[_reg(ins[1])]))
            pc += step
        elif op == "call":  # execute_inst_call, lib.rs:816-849 (immediate target)
            assert not isinstance(ins[1], str)
            fp = regs[9]
            mem[(fp - 1) % P] = pc + step
            row["op0"], row["dst"], row["op1"] = (fp - 1) % P, pc + step, int(ins[1]) % P
            row["aux0"] = (fp - 2) % P
            row["aux1"] = mem[(fp - 2) % P]
            mem_log.append(((fp - 1) % P, clk, 1 << 24, 1, row["dst"]))
            mem_log.append(((fp - 2) % P, clk, 1 << 24, 0, row["aux1"]))
            pc = int(ins[1])
        elif op == "ret":  # execute_inst_ret, lib.rs:851-866
            fp = regs[9]
            row["op0"], row["aux0"] = (fp - 1) % P, (fp - 2) % P
            pc = mem[(fp - 1) % P]
            regs[9] = mem[(fp - 2) % P]
            row["dst"], row["aux1"] = pc, regs[9]
            mem_log.append(((fp - 1) % P, clk, 1 << 23, 0, pc))
            mem_log.append(((fp - 2) % P, clk, 1 << 23, 0, regs[9]))
        elif op == "assert":
            v = val(ins[1])
            assert v == 1, "assert failed in the VM"
            row["op1"] = v
            pc += step
        elif op == "cjmp":
            c = regs[_reg(ins[1])]
            row["op0"], row["s_op0"] = c, _reg(ins[1])
            t = val(ins[2])
            row["op1"] = t
            pc = t if c == 1 else pc + step
        elif op == "jmp":
            t = val(ins[1])
            row["op1"] = t
            pc = t
        else:
            raise ValueError(op)
        run_prophet(row["pc"])                 # lib.rs:2255-2257: the prophet labelled at this pc runs after its instruction
        steps.append(row)
        clk += 1
        assert len(steps) < max_steps, "program does not terminate"
    n = 1 << log_n
    assert len(steps) <= n, f"{len(steps)} steps do not fit 2^{log_n} rows"
    t = np.zeros((94, n), dtype=np.uint64)
    for i, s in enumerate(steps):  # generation/cpu.rs:62-178
        t[11, i], t[12, i], t[13, i] = s["tp"], s["clk"], s["pc"]
        t[14, i], t[15, i] = s.get("is_ext", 0), s.get("ext_cnt", 0)
        t[16:26, i] = s["regs"]
        t[26, i], t[27, i], t[28, i], t[29, i] = s["inst"], s["op1_imm"], s["opcode"], s["imm"]
        t[30, i], t[31, i], t[32, i], t[33, i], t[34, i] = s["op0"], s["op1"], s["dst"], s["aux0"], s["aux1"]
        t[35, i] = s["idx_storage"]
        if s["s_op0"] is not None:
            t[36 + s["s_op0"], i] = 1
        if s["s_op1"] is not None:
            t[46 + s["s_op1"], i] = 1
        if s["s_dst"] is not None:
            t[56 + s["s_dst"], i] = 1
        if "s_op0_0" in s:
            t[36, i] = s["s_op0_0"]                     # ext lines of tload / tstore keep the tape address in s_op0[0]
        for col, v in s.get("sel_raw", {}).items():     # ext lines of sstore / sload: cell addresses, cell values, tree key
            t[col, i] = v
        t[90, i] = 1 if (s.get("is_ext", 0) and s["op"] in ("sstore", "sload")) else 0   # is_storage_ext_line
        t[CPU_SELECTOR_COL[s["op"]], i] = 1
        t[85, i] = 1                                  # is_entry_sc: env_idx == 0
        t[86, i] = 1 if s.get("ext_len", 0) == s.get("ext_cnt", 0) else 0   # is_next_line_diff_inst: ext_length == ext_cnt
        t[87, i] = 0 if s["op"] == "end" else 1        # is_next_line_same_tx
        t[88, i] = s.get("filter_tape_looking", 0)
        # filter_looking_prog_imm (generation/cpu.rs:168-177): mload / mstore always fetch their second word, others when op1 is an immediate
        t[92, i] = 0 if s.get("is_ext", 0) else (1 if s["op"] in ("mload", "mstore") else s["op1_imm"])
    k = len(steps)
    if k != n:  # padding, generation/cpu.rs:180-208
        t[26, k:] = t[26, k - 1]
        t[35, k:] = t[35, k - 1]
        t[28, k:] = 1 << 20
        t[74, k:] = 1
        t[85, k:] = 1
        t[86, k:] = 1
        t[87, k:] = 0
        t[93, k:] = 1
    if want_side_tables == "all+storage":
        return t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log, bit_ops, psdn_calls, tape_log, st_log
    assert not st_log, "storage accesses are only returned with want_side_tables='all+storage'"
    if want_side_tables == "all+tape":
        return t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log, bit_ops, psdn_calls, tape_log
    assert not tape_log, "tape rows are only returned with want_side_tables='all+tape'"
    if want_side_tables == "all+poseidon":
        return t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log, bit_ops, psdn_calls
    assert not psdn_calls, "poseidon calls are only returned with want_side_tables='all+poseidon'"
    if want_side_tables == "all":
        return t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log, bit_ops
    assert not bit_ops or not want_side_tables, "bitwise rows are only returned with want_side_tables='all'"
    if want_side_tables == "memory":
        return t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log
    if want_side_tables:
        return t, steps, cmp_pairs, rc_cmp, rc_cpu
    return t, steps


def program_rows_of_run(program, steps):
    """(program lines, executed lines) of a VM run for the Program table: every word of the program at code address 0
    (generate_prog_trace, generation/prog.rs:106-131) and, per executed step, (pc, instruction) plus (pc + 1, immediate)
    when the instruction carries one (:56-104)."""
    words = []
    for ins in program:
        words += ola_encode(ins)
    prog_rows = [(0, 0, 0, 0, pc, w) for pc, w in enumerate(words)]
    exec_rows = []
    for s in steps:
        if s.get("is_ext", 0):
            continue  # ext lines fetch nothing (generate_prog_trace skips them, prog.rs:58-60)
        exec_rows.append((0, 0, 0, 0, s["pc"], s["inst"]))
        if s["op1_imm"] == 1 or s["op"] in ("mload", "mstore"):
            exec_rows.append((0, 0, 0, 0, s["pc"] + 1, s["imm"]))
    return prog_rows, exec_rows


MEM_OP_SELECTOR = {1 << 22: 6, 1 << 21: 7, 1 << 24: 8, 1 << 23: 9, 1 << 9: 10, 1 << 8: 11, 1 << 12: 13, 1 << 10: 14, 1 << 11: 15}  # mload, mstore, call, ret, tload, tstore, poseidon, sstore, sload (memory/columns.rs:16-25)


def tape_trace_from_log(tape_log, log_n):
    """Tape table of a VM run: gen_tape_table (executor/src/trace.rs:400-414: cells grouped by tape address in ascending
    order, access order inside an address) + generate_tape_trace (generation/tape.rs:10-73: padding repeats the last row
    as an unlooked tload).  Columns (tape/columns.rs:3-9): tx_idx, is_init_seg, opcode, addr, value, filter_looked."""
    n = 1 << log_n
    t = np.zeros((6, n), dtype=np.uint64)
    row = 0
    for addr in sorted(tape_log):
        for is_init, op, value, looked in tape_log[addr]:
            t[:, row] = [0, is_init, op, addr, value, looked]
            row += 1
    assert 2 <= row <= n
    if row != n:
        t[1, row:], t[2, row:], t[3, row:], t[4, row:] = t[1, row - 1], 1 << 9, t[3, row - 1], t[4, row - 1]
    return t


def poseidon_chunk_trace_from_calls(calls, log_n):
    """PoseidonChunk table of a VM run (insert_poseidon_chunk rows of execute_inst_poseidon + generate_poseidon_chunk_trace,
    generation/poseidon_chunk.rs:7-88): per call one main line and one ext line per absorbed chunk.  Returns (table [53][n],
    poseidon rows (input[12], output[12]))."""
    n = 1 << log_n
    t = np.zeros((53, n), dtype=np.uint64)
    row, psdn = 0, []
    for c in calls:
        for r in c["rows"]:
            t[0:8, row] = [0, 0, c["clk"], OP_POSEIDON, r["op0"], c["op1"], c["dst"], r["acc"]]
            t[8:16, row] = r["value"]
            t[16:20, row] = r["cap"]
            t[20:32, row] = r["hash"]
            t[32, row] = r["ext"]
            result = c["op1"] == r["acc"]
            t[33, row] = 1 if result else 0
            pad = c["op1"] % 8 if result else 0
            if pad:
                t[34 + pad, row] = 1
            t[42, row] = 1 - r["ext"]
            if r["ext"]:
                t[43:51, row] = [1 if (pad == 0 or k < pad) else 0 for k in range(8)]
            t[51, row] = r["ext"]
            row += 1
        psdn += c["perms"]
    assert 2 <= row <= n
    t[52, row:] = 1
    return t, psdn


def tape_records_from_log(tape_log):
    """TapeRow records [k, 5] = (is_init, opcode, addr, value, filter_looked) in gen_tape_table's order."""
    rows = [(is_init, op, addr, value, looked) for addr in sorted(tape_log) for is_init, op, value, looked in tape_log[addr]]
    return np.array(rows, dtype=np.uint64).reshape(-1, 5)


def poseidon_chunk_records_of_table(t):
    """PoseidonChunkRow records [k, 32] read back from the filled rows of a PoseidonChunk table (the record fields are table
    columns; the derived columns 33..52 are what the generator must reproduce)."""
    k = int((t[52] == 0).sum())
    r = np.zeros((k, 32), dtype=np.uint64)
    r[:, 0], r[:, 1], r[:, 2], r[:, 3], r[:, 4], r[:, 5], r[:, 6] = t[1, :k], t[2, :k], t[3, :k], t[6, :k], t[4, :k], t[5, :k], t[7, :k]
    r[:, 7:15], r[:, 15:19], r[:, 19:31], r[:, 31] = t[8:16, :k].T, t[16:20, :k].T, t[20:32, :k].T, t[32, :k]
    return r


def storage_records_of_table(t):
    """StorageHashRow records [k, 38] read back from the filled rows of a StorageAccess table, and how many of them are storage
    accesses (the program-hash reads, marked in column 46 on their layer-256 row, come last)."""
    k = int((t[47] == 0).sum())
    r = np.zeros((k, 38), dtype=np.uint64)
    r[:, 0], r[:, 1:5], r[:, 5:9] = t[0, :k], t[1:5, :k].T, t[5:9, :k].T
    r[:, 9], r[:, 10], r[:, 11], r[:, 12] = t[9, :k], t[10, :k], t[11, :k], t[12, :k]
    r[:, 13:17], r[:, 17:21], r[:, 21:25], r[:, 25] = t[13:17, :k].T, t[17:21, :k].T, t[21:25, :k].T, t[29, :k]
    r[:, 26:30], r[:, 30:34], r[:, 34:38] = t[30:34, :k].T, t[34:38, :k].T, t[25:29, :k].T
    prog = np.nonzero(t[46, :k])[0]
    n_access = k if len(prog) == 0 else int(prog[0]) - 255
    return r, n_access


def sccall_records_of_table(t):
    k = int((t[25] == 0).sum())
    return np.ascontiguousarray(t[1:25, :k].T)


def memory_cells_to_records(cells):
    """The MemoryTraceCell list of memory_trace_from_log as the 15-u64 records generate_memory_trace consumes (layout:
    include/ola_gpu.h ola_generate_memory_trace; core/src/trace/trace.rs MemoryTraceCell).  Single-contract runs: env_idx 0."""
    r = np.zeros((len(cells), 15), dtype=np.uint64)
    for i, c in enumerate(cells):
        r[i] = [0, c["is_rw"], c["addr"], c["clk"], c["op"], c["is_write"], c["value"], c["diff_addr"], c["diff_addr_inv"], c["diff_clk"],
                c["cond"], c["rw_addr_unchanged"], c["prophet"], c["heap"], c["rc_value"]]
    return r


def memory_trace_from_log(mem_log, log_n, want_cells=False):
    """Memory table of a VM run: gen_memory_table (executor/src/trace.rs:20-199: cells grouped by address in ascending
    order, access order inside an address; diff_addr / diff_clk / diff_addr_cond and the values each row sends to the
    RangeCheck table, by region: read-write stack below p - 2 span, read-write heap [p - 2 span, p - span), write-once
    prophet region [p - span, p)) + generate_memory_trace (circuits/src/generation/memory.rs:8-155: the two range-check
    filters, padding with write-once rows).  Log entries: (addr, clk, opcode mask or 0 for a prophet write, is_write,
    value).  Returns (table [29][2^log_n], mem_sort range-check values, mem_region range-check values) -- the second
    list only when a heap / prophet cell was touched (the two-value form is kept for runs that stay on the stack)."""
    SPAN = (1 << 32) - 1
    by_addr = {}
    for addr, clk, op, is_write, value in mem_log:
        by_addr.setdefault(addr, []).append((clk, op, is_write, value))
    cells = []
    origin_addr = origin_clk = 0
    first_row = first_heap_row = True
    for addr in sorted(by_addr):
        new_addr = True
        prophet, heap = int(addr >= P - SPAN), int(P - 2 * SPAN <= addr < P - SPAN)
        cond = (P - addr) if prophet else ((P - SPAN - addr) if heap else 0)
        for clk, op, is_write, value in by_addr[addr]:
            c = {"addr": addr, "clk": clk, "op": op, "is_write": is_write, "value": value, "is_rw": 1 - prophet, "prophet": prophet,
                 "heap": heap, "cond": cond, "diff_addr": 0, "diff_addr_inv": 0, "diff_clk": 0, "rw_addr_unchanged": 0, "rc_value": 0}
            if first_row:
                first_row = new_addr = False
                if heap:
                    first_heap_row = False
            elif new_addr:
                c["diff_addr"] = addr - origin_addr
                if prophet:                       # write-once region: the row is ordered by its distance to p, not by diff_addr
                    c["rc_value"] = cond
                elif heap and first_heap_row:     # the first heap cell: nothing to compare with
                    c["diff_addr"] = 0
                    first_heap_row = False
                else:
                    c["diff_addr_inv"] = _finv(c["diff_addr"])
                    c["rc_value"] = c["diff_addr"]
                new_addr = False
            else:
                c["diff_clk"] = clk - origin_clk
                if prophet:
                    c["rc_value"] = cond
                else:
                    c["rw_addr_unchanged"] = 1
                    c["rc_value"] = c["diff_clk"]
            assert c["rc_value"] <= 0xFFFFFFFF and cond <= 0xFFFFFFFF, "U32RangeCheckFail"
            cells.append(c)
            origin_clk = clk
        origin_addr = addr
    n = 1 << log_n
    k = len(cells)
    assert 2 <= k <= n
    t = np.zeros((29, n), dtype=np.uint64)
    rc_sort, rc_region = [], []
    for i, c in enumerate(cells):
        t[2, i], t[3, i], t[4, i], t[5, i] = c["is_rw"], c["addr"], c["clk"], c["op"]
        t[16 if c["op"] == 0 else MEM_OP_SELECTOR[c["op"]], i] = 1     # opcode 0 = a prophet write (COL_MEM_S_PROPHET)
        t[17, i], t[18, i] = c["is_write"], c["value"]
        t[19, i], t[20, i], t[21, i], t[22, i] = c["diff_addr"], c["diff_addr_inv"], c["diff_clk"], c["cond"]
        t[23, i], t[24, i], t[25, i], t[26, i] = c["rw_addr_unchanged"], c["prophet"], c["heap"], c["rc_value"]
        looking = not (i == 0 or c["prophet"] or (c["heap"] and not cells[i - 1]["heap"]))
        t[27, i] = int(looking)
        t[28, i] = int(c["heap"] or c["prophet"])
        if looking:
            rc_sort.append(c["rc_value"])
        if t[28, i]:
            rc_region.append(c["cond"])
    if k != n:  # memory.rs:113-146: padding continues the write-once region (from p - span when the last filled row is read-write)
        addr = P - SPAN if cells[-1]["is_rw"] else cells[-1]["addr"] + 1
        for i in range(k, n):
            assert addr < P, "the prophet region is full"
            t[16, i] = 1
            t[3, i] = addr
            t[17, i] = 1
            d = (addr - int(t[3, k - 1])) % P if i == k else 1
            t[19, i], t[20, i] = d, _finv(d)
            t[22, i] = (P - addr) % P
            t[24, i] = 1
            t[26, i] = t[22, i]
            addr += 1
    if want_cells:
        return t, cells
    if rc_region:
        return t, rc_sort, rc_region
    return t, rc_sort


def calls_program(n_iter, linear=False, bitwise=False, poseidon=False, tape=False):
    """Exercises memory and builtin opcodes on top of fib_program's set: a stack frame (mstore / mload relative to r9), a
    call / ret pair, gte comparisons in both directions and u32 range checks; with bitwise=True also and / or / xor in the
    callee.  linear=True replaces the Fibonacci step by r1 + r2 (the loop counter) so that long runs stay inside the u32
    range checks.  Jump targets are word addresses, resolved from labels below."""
    body = [
        ("mov", "r9", 100),               # frame pointer
        ("mov", "r0", 0),
        ("mov", "r1", 1),
        ("mov", "r2", 0),
        "loop",
        ("mstore", "r9", -2, "r9"),       # [fp-2] = fp   (what call / ret read back into r9)
        ("call", "step"),                 # return address stored at [fp-1]
        ("add", "r2", "r2", 1),
        ("gte", "r4", "r2", n_iter),      # r4 = (r2 >= n_iter)
        ("gte", "r5", "r1", "r0"),        # the pair is non-decreasing: r5 = 1
        ("assert", "r5"),
        ("not", "r6", "r4"),              # r6 = p - 1 - r4
        ("add", "r6", "r6", 2),           # r6 = 1 - r4  (+ p)
        ("cjmp", "r6", "loop"),           # loop while r2 < n_iter
        ("range", "r2"),
        ("mload", "r7", "r9", -3),        # last sum the callee spilled
        ("eq", "r8", "r7", "r1"),
        ("assert", "r8"),
    ] + ([
        # poseidon=True: spill 11 words at [40..51) and hash them (one full chunk + a tail of 3) into [60..64), then a second
        # call over exactly 8 words (a single full chunk) into [64..68) and a short one (5 words) into [68..72)
        ("mov", "r3", 40),
        ("mstore", "r3", 0, "r0"), ("mstore", "r3", 1, "r1"), ("mstore", "r3", 2, "r2"), ("mstore", "r3", 3, "r7"),
        ("mstore", "r3", 4, "r9"), ("mstore", "r3", 5, "r1"), ("mstore", "r3", 6, "r0"), ("mstore", "r3", 7, "r2"),
        ("mstore", "r3", 8, "r8"), ("mstore", "r3", 9, "r7"), ("mstore", "r3", 10, "r1"),
        ("mov", "r4", 60),
        ("poseidon", "r4", "r3", 11),
        ("mov", "r4", 64),
        ("mov", "r5", 8),
        ("poseidon", "r4", "r3", "r5"),
        ("mov", "r4", 68),
        ("poseidon", "r4", "r3", 5),
        ("mload", "r6", "r4", 0),         # read one digest word back
    ] if poseidon else []) + ([
        # tape=True: copy three stack words to the tape (tp 0 -> 3), read the last two back (flag 1) and then the word at
        # tape address 0 (flag 0) into another stack range, and check one of them
        ("mstore", "r9", -8, "r1"), ("mstore", "r9", -7, "r2"), ("mstore", "r9", -6, "r7"),
        ("add", "r4", "r9", -8),
        ("tstore", "r4", 3),
        ("add", "r5", "r9", -12),
        ("mov", "r6", 1),
        ("tload", "r5", "r6", 2),
        ("mov", "r6", 0),
        ("tload", "r5", "r6", 0),
        ("mload", "r3", "r9", -12),
        ("eq", "r8", "r3", "r1"),
        ("assert", "r8"),
    ] if tape else []) + [
        ("jmp", "done"),
        "step",                           # (r0, r1) <- (r1, r0 + r1); spills the sum to [fp-3]
        ("add", "r3", "r1", "r2") if linear else ("add", "r3", "r0", "r1"),
        ("mov", "r0", "r1"),
        ("mov", "r1", "r3"),
        ("mstore", "r9", -3, "r3"),
        ("range", "r3"),
    ]
    if bitwise:
        body += [
            ("and", "r7", "r3", 0xFF0F),
            ("or", "r8", "r7", "r2"),
            ("xor", "r7", "r8", "r3"),
        ]
    body += [("ret",), ("end",), "done", ("end",)]
    # resolve labels to word addresses
    labels, pc = {}, 0
    for x in body:
        if isinstance(x, str):
            labels[x] = pc
        else:
            pc += len(ola_encode(tuple(0 if (isinstance(a, str) and a in ("loop", "step", "done")) else a for a in x)))
    return [tuple(labels[a] if (isinstance(a, str) and a in labels) else a for a in x) for x in body if not isinstance(x, str)]


def storage_program():
    """sstore / sload over two slots of contract 0: write slot A, read it back, overwrite it (a repeated write), write slot
    B, read an absent slot C (all-zero value), read A again; every loaded word is then pulled into a register through
    mload so that the values the storage tree returned reach the register file.  Memory layout: keys at 200 / 210 / 220
    (4 words each), values at 300 / 310, read buffers at 400...; r9 stays 0 (no stack frame is needed)."""
    prog = [("mov", "r1", 200), ("mov", "r2", 300), ("mov", "r3", 400), ("mov", "r5", 210), ("mov", "r6", 310), ("mov", "r7", 220)]
    cells = {200: [1, 2, 3, 4], 210: [5, 6, 7, 8], 220: [9, 9, 9, 9], 300: [11, 12, 13, 14], 310: [21, 22, 23, 24]}
    for base, words in cells.items():
        prog.append(("mov", "r8", base))
        for i, w in enumerate(words):
            prog += [("mov", "r0", w), ("mstore", "r8", i, "r0")]
    prog += [("sstore", "r1", "r2"),            # A := (11, 12, 13, 14)      initial write
             ("sload", "r1", "r3"),             # [400..404) := A
             ("mload", "r4", "r3", 2),          # r4 = 13
             ("sstore", "r1", "r6"),            # A := (21, 22, 23, 24)      repeated write
             ("sstore", "r5", "r2"),            # B := (11, 12, 13, 14)
             ("sload", "r7", 410),              # [410..414) := C = 0        (immediate buffer address)
             ("sload", "r1", 420),              # [420..424) := A
             ("mov", "r8", 420), ("mload", "r0", "r8", 3),   # r0 = 24
             ("mov", "r8", 410), ("mload", "r2", "r8", 0),   # r2 = 0
             ("end",)]
    return prog


def fib_program(n_iter):
    """r0, r1 = fib pair; r2 = loop counter; loops n_iter times, checks the result bookkeeping with eq / assert / neq / not."""
    return [
        ("mov", "r0", 0),            # pc 0
        ("mov", "r1", 1),            # pc 2
        ("mov", "r2", 0),            # pc 4
        # loop (pc 6):
        ("add", "r3", "r0", "r1"),   # pc 6
        ("mov", "r0", "r1"),         # pc 7
        ("mov", "r1", "r3"),         # pc 8
        ("add", "r2", "r2", 1),      # pc 9
        ("neq", "r4", "r2", n_iter),  # pc 11
        ("cjmp", "r4", 6),           # pc 13
        ("eq", "r5", "r2", n_iter),  # pc 15
        ("assert", "r5"),            # pc 17
        ("mul", "r6", "r1", "r1"),   # pc 18
        ("not", "r7", "r6"),         # pc 19
        ("add", "r8", "r7", "r6"),   # pc 20   r8 = p - 1
        ("add", "r8", "r8", 1),      # pc 21   r8 = 0
        ("eq", "r9", "r8", 0),       # pc 23
        ("assert", "r9"),            # pc 25
        ("jmp", 28),                 # pc 26
        ("end",),                    # pc 28
    ]


def real_program_system(orc, rng, n_iter=12, linear=False, cpu_log=9, mem_log_n=7, cmp_log=6, prog_log=9, beta=0x1234567890ABCDEF % P,
                        bitwise=False, beta_bitwise=0x0FEDCBA987654321 % P, bitwise_log=9, poseidon=False, tape=False):
    """An eight-table system produced by RUNNING a program: [Cpu, Memory, Cmp, RangeCheck, Poseidon, StorageAccess, Program,
    ProgChunk]; with bitwise=True the program also executes and / or / xor and the Bitwise table (with its own compress
    challenge) joins as a ninth table behind the cpu->bitwise lookup; with poseidon=True the program also hashes memory
    ranges with the poseidon opcode and PoseidonChunk joins (cpu->poseidon_chunk, poseidon_chunk->memory x12,
    poseidon_chunk->poseidon); with tape=True also tstore / tload (CPU ext lines) and the Tape table: eleven of the twelve
    tables -- only SCCall, which needs a second contract, is not reached by a run.  The VM (cpu_vm_trace) fills the CPU table and logs memory accesses, comparisons and range checks; the
    Memory / Cmp / RangeCheck tables are generated from those logs the way the executor does; the Program table holds the
    program's words and one executed line per fetched word; ProgChunk hashes the program (Poseidon sponge over lines of
    8 words), its digest is read from the storage tree at code address 0, and every sponge / Merkle hash is a Poseidon
    row.  Lookups with real data: cpu->memory (x3), memory->rangecheck, cpu->cmp, cmp->rangecheck, cpu->rangecheck,
    cpu->program (instruction and immediate), prog_chunk->program, prog_chunk->poseidon, prog_chunk->storage,
    storage->poseidon.  Returns (table_ids, traces, compress_challenges)."""
    prog = calls_program(n_iter, linear=linear, bitwise=bitwise, poseidon=poseidon, tape=tape)
    cpu_t, steps, cmp_pairs, rc_cmp, rc_cpu, mlog, bit_ops, psdn_calls, tape_log = cpu_vm_trace(prog, cpu_log, want_side_tables="all+tape", orc=orc)
    mem_t, rc_sort = memory_trace_from_log(mlog, mem_log_n)
    cmp_t = cmp_trace(cmp_pairs, cmp_log)
    rc_t = rangecheck_trace(rc_cmp, cpu_vals=rc_cpu, mem_sort_vals=rc_sort)
    prog_rows, exec_rows = program_rows_of_run(prog, steps)
    words = [r[5] for r in prog_rows]
    chunk_log = max(3, ((len(words) + 7) // 8 - 1).bit_length())
    pc_t, psdn_prog, lines, roots = prog_chunk_valid_trace(orc, rng, chunk_log, programs=[([0, 0, 0, 0], words)])
    assert lines == prog_rows
    _, leaf = roots[0]
    st, psdn_st = storage_valid_trace(orc, rng, 8, [dict(addr_bits=[0] * 256, leaf=leaf, pre_leaf=leaf, is_write=0, for_prog=1)])
    rows = [(inp, [1, 0, 0, 0]) for inp, _ in psdn_prog]
    rows += [(inp, [0, 0, 1, 0] if is_leaf else [0, 0, 0, 1]) for inp, _, is_leaf in psdn_st]
    if poseidon:
        assert bitwise, "the ten-table system includes the Bitwise table"
        pch, psdn_chunk = poseidon_chunk_trace_from_calls(psdn_calls, 4)
        rows += [(inp, [1, 0, 0, 0]) for inp, _ in psdn_chunk]
    ps = poseidon_valid_trace(orc, 10, rows)
    pt = program_valid_trace(rng, prog_log, beta, prog_rows=prog_rows, exec_rows=exec_rows)
    if tape:
        assert poseidon and bitwise, "the eleven-table system includes Bitwise and PoseidonChunk"
        bw = bitwise_valid_trace(rng, bitwise_log, beta_bitwise, ops=bit_ops)
        tp_t = tape_trace_from_log(tape_log, 3)
        return ([0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11], [cpu_t, mem_t, bw, cmp_t, rc_t, ps, pch, st, tp_t, pt, pc_t],
                [0, 0, beta_bitwise, 0, 0, 0, 0, 0, 0, beta, 0])
    if poseidon:
        bw = bitwise_valid_trace(rng, bitwise_log, beta_bitwise, ops=bit_ops)
        return ([0, 1, 2, 3, 4, 5, 6, 7, 10, 11], [cpu_t, mem_t, bw, cmp_t, rc_t, ps, pch, st, pt, pc_t],
                [0, 0, beta_bitwise, 0, 0, 0, 0, 0, beta, 0])
    if bitwise:
        bw = bitwise_valid_trace(rng, bitwise_log, beta_bitwise, ops=bit_ops)
        return [0, 1, 2, 3, 4, 5, 7, 10, 11], [cpu_t, mem_t, bw, cmp_t, rc_t, ps, st, pt, pc_t], [0, 0, beta_bitwise, 0, 0, 0, 0, beta, 0]
    ids = [0, 1, 3, 4, 5, 7, 10, 11]
    return ids, [cpu_t, mem_t, cmp_t, rc_t, ps, st, pt, pc_t], [0, 0, 0, 0, 0, 0, beta, 0]


# ---------------------------------------------------------------------------------------------------------------------
# Running the reference's own assembly test programs (assembler/test_data/asm/*.json, committed as
# tests/golden/ola_programs.json by tools/extract_encoding_golden.py) through the VM above.
# ---------------------------------------------------------------------------------------------------------------------
# A transaction's initial tape as init_tape lays it out (executor/src/load_tx.rs:89-117): block_number, block_timestamp,
# sequencer_address[4], version, chain_id (address 7), caller_address[4], nonce, signature_r[4], signature_s[4], tx_hash[4],
# calldata (here one word: length 0), caller / callee / callee-code addresses.
CONTEXT_TAPE = [5, 1700000000, 1, 2, 3, 4, 3, 1027, 9, 9, 9, 9, 1, 11, 12, 13, 14, 21, 22, 23, 24, 31, 32, 33, 34] + [0] + [0, 0, 0, 1] + [0, 0, 0, 2] * 2


def reference_test_tape(calldata):
    """The initial tape executor/src/tests.rs::executor_run_test_program builds for a test with calldata: init_tape
    (load_tx.rs:89-117) over init_tx_context_mock (core/src/vm/transaction.rs:18-56), the calldata, then the caller /
    callee / callee-code addresses the test fixes (tests.rs:69-86)."""
    ctx = [3, 1692846754, 1, 2, 3, 4, 3, 1, 5, 6, 7, 8, 25, 129, 130, 131, 132, 133, 134, 135, 136, 137, 138, 139, 140]
    return ctx + [int(x) for x in calldata] + [17, 18, 19, 20] + [9, 10, 11, 12] + [13, 14, 15, 16]


# calldata of the reference's own executor tests (executor/src/tests.rs), by program
REFERENCE_CALLDATA = {"fibo_loop": [10, 1, 2, 1015130275], "ptr_call": [0, 2657046596], "sc_input": [10, 20, 2, 253268590],
                      "storage_u32": [0, 2364819430], "poseidon_hash": [0, 1239976900], "context_fetch": [0, 3458276513],
                      "printf": [5, 111, 108, 97, 118, 109, 11, 12, 8, 3238128773], "global": [0, 4171824493]}


def parse_ola_prophets(doc):
    """(program tuples, {pc: prophet spec}) of an assembler/test_data/asm/<name>.json document: a prophet is keyed by the word
    address of the instruction BEFORE its `.PROPHETn_m` label (relocate.rs:151-159) and runs after that instruction, so that the
    `mov rX psp` behind the label already sees its outputs.  Only the `malloc` built-in is
    modelled (cpu_vm_trace.run_prophet); anything else raises."""
    import re

    prog, labels = parse_ola_asm(doc["program"], want_labels="hosts")
    out = {}
    for p in doc.get("prophets", []):
        code = re.sub(r"\s+", "", p["code"])
        if code == "%{entry(){printf(cid.base,cid.flag);}%}":
            assert not p["outputs"]
            out[labels[p["label"]]] = {"fn": "printf", "inputs": 2}
            continue
        helpers = {"%{functionmod(feltx,felty)->felt{returnx%y;}entry(){cid.r=mod(cid.x,cid.y);}%}": ("mod", 2),
                   "%{functiondiv(feltx,felty)->felt{returnx/y;}entry(){cid.q=div(cid.x,cid.y);}%}": ("div", 2),
                   "%{functionsplit_hi(feltin)->felt{returnin/4294967296;}entry(){cid.out=split_hi(cid.in);}%}": ("split_hi", 1),
                   "%{functionsplit_lo(feltin)->felt{returnin%4294967296;}entry(){cid.out=split_lo(cid.in);}%}": ("split_lo", 1)}
        if code in helpers:
            assert len(p["inputs"]) == helpers[code][1] and len(p["outputs"]) == 1 and all(i["length"] == 1 and not i["is_ref"] for i in p["inputs"])
            out[labels[p["label"]]] = {"fn": helpers[code][0], "inputs": helpers[code][1]}
            continue
        if code != "%{entry(){cid.addr=malloc(cid.len);}%}":
            raise NotImplementedError("prophet: " + p["code"])
        assert len(p["inputs"]) == 1 and p["inputs"][0]["length"] == 1 and not p["inputs"][0]["is_ref"] and len(p["outputs"]) == 1
        out[labels[p["label"]]] = {"fn": "malloc", "inputs": 1}
    return prog, out


def parse_ola_asm(text, want_labels=False):
    """Assembly text -> the VM's instruction tuples.  As the reference assembler does (assembler/src/relocate.rs:21-86,
    encoder.rs): the scope labelled `main` moves to the front, an instruction occupies two words when its last operand is an
    immediate or a label or when it is mload / mstore, labels resolve to word addresses.  Memory operands [rN], [rN,off]."""
    import re

    lines = [l.strip() for l in text.split("\n") if l.strip()]
    scopes, cur = [], None
    for l in lines:
        if l.endswith(":") and not l.startswith("."):
            cur = [l]
            scopes.append(cur)
        else:
            assert cur is not None, "instruction before the first scope label"
            cur.append(l)
    scopes.sort(key=lambda sc: 0 if sc[0] == "main:" else 1)
    assert scopes[0][0] == "main:", "no main scope"
    is_reg = lambda a: re.fullmatch(r"r\d", a) is not None
    labels, pc, insts, last_pc, label_host = {}, 0, [], 0, {}
    for l in (x for sc in scopes for x in sc):
        if l.endswith(":"):
            labels[l[:-1]] = pc
            label_host[l[:-1]] = last_pc     # a prophet label's host is the instruction BEFORE it (relocate.rs:151-159, ori_counter)
            continue
        parts = l.replace(", ", ",").split()
        op, args = parts[0], parts[1:]
        insts.append((op, args))
        last_pc = pc
        two = op in ("mload", "mstore") or (bool(args) and not (is_reg(args[-1]) or args[-1] == "psp" or args[-1].startswith("[")))
        pc += 2 if two else 1
    out = []
    for op, args in insts:
        res = []
        for a in args:
            m = re.fullmatch(r"\[(r\d)(?:,([+-]?\d+))?\]", a)
            mf = re.fullmatch(r"\[(r\d),(r\d)(?:,([+-]?\d+))?\]", a)  # [anchor, offset reg(, factor = 1)]: operands.rs:80-114
            if mf:
                res.append((mf.group(1), (mf.group(2), int(mf.group(3) or 1))))
            elif m:
                res.append((m.group(1), int(m.group(2) or 0)))
            elif is_reg(a) or a == "psp":
                res.append(a)
            elif re.fullmatch(r"[+-]?\d+", a):
                res.append(int(a))
            else:
                res.append(labels[a])
        if op == "mstore":
            (base, off), v = res
            out.append(("mstore", base, off, v))
        elif op == "mload":
            dst, (base, off) = res
            out.append(("mload", dst, base, off))
        else:
            out.append((op, *res))
    return (out, label_host) if want_labels == "hosts" else ((out, labels) if want_labels else out)


def run_system(orc, rng, program, cpu_log=None, beta=0x1234567890ABCDEF % P, beta_bitwise=0x0FEDCBA987654321 % P, init_tape=(), prophets=None):
    """Run `program` (VM tuples) and build every table its run touches: always Cpu, Cmp, RangeCheck, Program; Memory, Bitwise,
    Tape, Poseidon + PoseidonChunk when the run produced rows for them.  Returns (table_ids, traces, compress_challenges,
    steps)."""
    nsteps = len(cpu_vm_trace(program, 20, want_side_tables="all+storage", orc=orc, init_tape=init_tape, prophets=prophets)[1]) if cpu_log is None else None
    if cpu_log is None:
        cpu_log = max(4, (nsteps - 1).bit_length())
    cpu_t, steps, cmp_pairs, rc_cmp, rc_cpu, mlog, bit_ops, psdn_calls, tape_log, st_log = cpu_vm_trace(program, cpu_log, want_side_tables="all+storage", orc=orc, init_tape=init_tape, prophets=prophets)
    lg = lambda k, lo: max(lo, (max(k, 1) - 1).bit_length())
    tabs = {0: cpu_t}
    cc = {}
    rc_sort, rc_region = [], []
    if len(mlog) >= 2:
        out = memory_trace_from_log(mlog, lg(len(mlog) + 1, 2))
        tabs[1], rc_sort = out[0], out[1]
        rc_region = out[2] if len(out) == 3 else []
    if bit_ops:
        tabs[2] = bitwise_valid_trace(rng, 9, beta_bitwise, ops=bit_ops)
        cc[2] = beta_bitwise
    tabs[3] = cmp_trace(cmp_pairs, lg(len(cmp_pairs) + 1, 4))
    tabs[4] = rangecheck_trace(rc_cmp, cpu_vals=rc_cpu, mem_sort_vals=rc_sort, mem_region_vals=rc_region)
    hash_rows = []
    if psdn_calls:
        tabs[6], psdn_rows = poseidon_chunk_trace_from_calls(psdn_calls, lg(sum(len(c["rows"]) for c in psdn_calls) + 1, 2))
        hash_rows += [(inp, [1, 0, 0, 0]) for inp, _ in psdn_rows]
    if st_log:  # sstore / sload: StorageAccess walks one consistent tree; tree-key, leaf and branch hashes join the Poseidon table
        tabs[7], st_rows = storage_tables_from_log(orc, rng, st_log)
        hash_rows += st_rows
    if hash_rows:
        tabs[5] = poseidon_valid_trace(orc, lg(len(hash_rows) + 1, 4), hash_rows)
    if tape_log:
        tabs[8] = tape_trace_from_log(tape_log, lg(sum(len(v) for v in tape_log.values()) + 1, 2))
    prog_rows, exec_rows = program_rows_of_run(program, steps)
    tabs[10] = program_valid_trace(rng, lg(max(len(prog_rows), len(exec_rows)) + 1, 2), beta, prog_rows=prog_rows, exec_rows=exec_rows)
    cc[10] = beta
    ids = sorted(tabs)
    return ids, [tabs[i] for i in ids], [cc.get(i, 0) for i in ids], steps
