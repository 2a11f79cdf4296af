"""Vectorised generator of a SATISFYING 12-table system for a loop-shaped Ola program (BASELINE configs[2]:
"full ola_prove of fibonacci-loop program, 2^22 CPU-table rows").

The spec of every table is `workload/tracegen.py` (a Python restatement of the reference executor and of
circuits/src/generation/*.rs); it runs ~20k steps/s, i.e. minutes for 4M rows.  This module produces the same
tables for ONE program shape -- straight-line prologue, a loop whose every iteration takes the same path, an exit --
by executing the loop body once over K lanes (numpy arrays, one lane per iteration):

  * `FIB_LOOP_ASM`: the inner loop of the reference's criterion benchmark program
    (assembler/test_data/asm/fibo_loop.json, scope `fib_non_recursive`, labels .LBL10_3 - .LBL10_6: stack-frame
    mload / mstore, the `gte` / `neq` / `and` / `cjmp` loop test, the fib update), without the printf prophets;
  * the loop-carried MEMORY state of lane k (a = F(k), b = F(k+1), i = k + 2) is supplied analytically and asserted
    consistent (the state lane k leaves equals the state lane k+1 starts from); the loop-carried REGISTER file is
    found by a dry pass and asserted to be a fixed point;
  * side tables come from the run's logs exactly as the executor produces them: Memory (gen_memory_table +
    generate_memory_trace), Cmp, RangeCheck, Bitwise, Program; the program's hash trio (ProgChunk, Poseidon,
    StorageAccess) and the untouched tables (PoseidonChunk, Tape, SCCall) are small and reuse tracegen.

tests/test_workload.py checks `fib_loop_system(n)` against tracegen's VM run of the same program column by column
at small n (the Halo2 permuted-table columns as multisets + their lookup property: the fill order of unused table
values is free, see permuted_cols_vec), and that every table satisfies its AIR.
"""
import numpy as np

from . import tracegen as tg

P = tg.P
SPAN = (1 << 32) - 1
U = np.uint64

FIB_LOOP_ASM = """
main:
  add r9 r9 5
  mov r0 {n}
  mstore [r9,-1] r0
  mov r0 0
  mstore [r9,-2] r0
  mov r0 1
  mstore [r9,-3] r0
  mov r0 2
  mstore [r9,-4] r0
  mov r0 2
  mstore [r9,-5] r0
  jmp .LBL10_4
.LBL10_4:
  mload r0 [r9,-5]
  mload r1 [r9,-1]
  gte r0 r1 r0
  gte r1 r0 0
  neq r0 r0 0
  and r1 r1 r0
  cjmp r1 .LBL10_5
  jmp .LBL10_6
.LBL10_5:
  mload r1 [r9,-2]
  mload r2 [r9,-3]
  add r0 r1 r2
  mstore [r9,-4] r0
  mload r0 [r9,-3]
  mstore [r9,-2] r0
  mload r0 [r9,-4]
  mstore [r9,-3] r0
  mload r0 [r9,-5]
  add r3 r0 1
  mstore [r9,-5] r3
  jmp .LBL10_4
.LBL10_6:
  mload r0 [r9,-4]
  add r9 r9 -5
  end
"""
PROLOGUE_STEPS, ITER_STEPS, EXIT_STEPS = 12, 19, 11  # steps before the first loop test / per iteration / final test + exit


def fib_loop_program(n):
    """VM tuples + labels of FIB_LOOP_ASM with loop bound n: the loop runs n - 1 times (i = 2 .. n)."""
    return tg.parse_ola_asm(FIB_LOOP_ASM.format(n=int(n)), want_labels=True)


def steps_of(n):
    return PROLOGUE_STEPS + ITER_STEPS * (n - 1) + EXIT_STEPS


def bound_for_rows(log_n_cpu):
    """The largest loop bound whose run fits 2^log_n_cpu CPU rows."""
    return ((1 << log_n_cpu) - PROLOGUE_STEPS - EXIT_STEPS) // ITER_STEPS + 1


# ---- field helpers on canonical uint64 arrays ---------------------------------------------------------------------
def addp(a, b):
    a, b = np.asarray(a, dtype=U), np.asarray(b, dtype=U)
    s = a + b
    wrap = (s < a) | (s >= U(P))
    return np.where(wrap, s - U(P), s)


def subp(a, b):
    a, b = np.asarray(a, dtype=U), np.asarray(b, dtype=U)
    return np.where(a >= b, a - b, a + (U(P) - b))


def inv_many(vals):
    """Field inverses of a uint64 array (0 -> 0): Montgomery's trick over the distinct values, Python integers."""
    vals = np.asarray(vals, dtype=U)
    uniq, back = np.unique(vals, return_inverse=True)
    xs = [int(x) for x in uniq]
    nz = [x for x in xs if x]
    pref, acc = [], 1
    for x in nz:
        pref.append(acc)
        acc = acc * x % P
    inv_acc = pow(acc, P - 2, P)
    inv_nz = [0] * len(nz)
    for i in range(len(nz) - 1, -1, -1):
        inv_nz[i] = inv_acc * pref[i] % P
        inv_acc = inv_acc * nz[i] % P
    it = iter(inv_nz)
    out = np.array([next(it) if x else 0 for x in xs], dtype=U)
    return out[back].reshape(vals.shape)


def fib_mod_p(count):
    f = np.zeros(count, dtype=U)
    a, b = 0, 1
    for i in range(count):
        f[i] = a
        a, b = b, (a + b) % P
    return f


def permuted_cols_vec(inputs, table):
    """Halo2-style permuted (input, table) columns (circuits/src/stark/lookup.rs:68-131): inputs sorted; the r-th
    occurrence of a value in the sorted inputs is paired with the r-th occurrence of that value in the sorted table when
    there is one; the remaining positions take the unused table values.  The reference hands the unused values out in
    stack order; here in ascending order -- any order satisfies eval_lookups (lookup.rs:13-35), the column is the same
    multiset."""
    si, st = np.sort(np.asarray(inputs, dtype=U)), np.sort(np.asarray(table, dtype=U))
    n = len(si)
    idx = np.arange(n)

    def rank(a):
        first = np.r_[True, a[1:] != a[:-1]]
        return idx - np.maximum.accumulate(np.where(first, idx, 0))

    cnt_t = np.searchsorted(st, si, "right") - np.searchsorted(st, si, "left")
    cnt_i = np.searchsorted(si, st, "right") - np.searchsorted(si, st, "left")
    matched = rank(si) < cnt_t
    used = rank(st) < cnt_i
    perm = si.copy()
    assert int((~matched).sum()) == int((~used).sum())
    perm[~matched] = st[~used]
    return si, perm


# ---- the lane-parallel VM --------------------------------------------------------------------------------------------
class _Run:
    """Executes instruction sequences over `lanes` lanes at once.  Registers and memory cells are uint64 arrays [lanes];
    control flow must be lane-uniform.  Rows go straight into the CPU table (strided slices), side effects into logs."""

    def __init__(self, program, cpu_table):
        self.t = cpu_table
        self.at_pc, words = {}, []
        for ins in program:
            enc = tg.ola_encode(ins)
            self.at_pc[len(words)] = (ins, enc)
            words += enc
        self.words = words
        self.mem_log, self.cmp_log, self.bit_log, self.exec_log = [], [], [], []

    def segment(self, pc, stop_pc, lanes, regs, mem, row0, stride, emit=True):
        """Run from pc until pc == stop_pc (after at least one step) or `end`.  Step s of lane k is CPU row / clk
        row0 + s + stride * k.  regs: list of 10 arrays [lanes]; mem: {addr: array [lanes]}.  Returns (pc, steps)."""
        t, s = self.t, 0
        lane_off = np.arange(lanes, dtype=U) * U(stride)
        full = lambda v: np.full(lanes, int(v) % P, dtype=U)
        while True:
            ins, enc = self.at_pc[pc]
            op, step = ins[0], len(enc)
            clk = U(row0 + s) + lane_off
            rows = slice(row0 + s, row0 + s + stride * (lanes - 1) + 1, stride) if lanes > 1 else slice(row0 + s, row0 + s + 1)
            pre = [r for r in regs]
            f = dict(op0=0, op1=0, dst=0, aux0=0, aux1=0)
            sel = {}
            reg = tg._reg

            def val(x):
                if isinstance(x, str):
                    sel["s_op1"] = reg(x)
                    return regs[reg(x)]
                return full(x)

            def memlog(addr, mask, is_write, v, sub):
                if emit:
                    self.mem_log.append((addr, clk, sub, mask, is_write, v))

            nxt = pc + step
            if op == "end":
                pass
            elif op in ("mov", "not"):
                v = val(ins[2])
                f["op1"] = v
                regs[reg(ins[1])] = v if op == "mov" else subp(full(P - 1), v)
                f["dst"], sel["s_dst"] = regs[reg(ins[1])], reg(ins[1])
            elif op in ("add", "eq", "neq"):
                a = regs[reg(ins[2])]
                f["op0"], sel["s_op0"] = a, reg(ins[2])
                b = val(ins[3])
                f["op1"] = b
                if op == "add":
                    r = addp(a, b)
                else:
                    d = subp(a, b)
                    f["aux0"] = inv_many(d) if emit else d
                    r = ((d == 0) if op == "eq" else (d != 0)).astype(U)
                regs[reg(ins[1])] = r
                f["dst"], sel["s_dst"] = r, reg(ins[1])
            elif op in ("and", "or", "xor"):
                a = regs[reg(ins[2])]
                f["op0"], sel["s_op0"] = a, reg(ins[2])
                b = val(ins[3])
                f["op1"] = b
                assert (a >> U(32)).max() == 0 and (b >> U(32)).max() == 0, "the Bitwise table works on u32 operands"
                r = a & b if op == "and" else (a | b if op == "or" else a ^ b)
                regs[reg(ins[1])] = r
                f["dst"], sel["s_dst"] = r, reg(ins[1])
                if emit:
                    self.bit_log.append((clk, 1 << tg.OPCODE_SHIFT[op], a, b))
            elif op == "gte":
                a = regs[reg(ins[2])]
                f["op0"], sel["s_op0"] = a, reg(ins[2])
                b = val(ins[3])
                f["op1"] = b
                r = (a >= b).astype(U)
                d = np.where(a >= b, a - b, b - a)
                assert d.max() <= 0xFFFFFFFF, "U32RangeCheckFail"
                regs[reg(ins[1])] = r
                f["dst"], sel["s_dst"] = r, reg(ins[1])
                if emit:
                    self.cmp_log.append((clk, a, b))
            elif op in ("mstore", "mload"):
                base_reg = ins[1] if op == "mstore" else ins[2]
                off = ins[2] if op == "mstore" else ins[3]
                assert not isinstance(off, tuple), "register-scaled memory operands are not vectorised"
                base = regs[reg(base_reg)]
                f["op0"], sel["s_op0"], f["op1"] = base, reg(base_reg), full(off)
                addr = addp(base, full(off))
                assert (addr == addr[0]).all(), "lane-dependent addresses are not vectorised"
                a0 = int(addr[0])
                f["aux1"] = addr
                if op == "mstore":
                    v = regs[reg(ins[3])]
                    f["dst"], sel["s_dst"] = v, reg(ins[3])
                    mem[a0] = v
                    memlog(a0, 1 << 21, 1, v, 0)
                else:
                    v = mem[a0]
                    regs[reg(ins[1])] = v
                    f["dst"], sel["s_dst"] = v, reg(ins[1])
                    memlog(a0, 1 << 22, 0, v, 0)
            elif op == "assert":
                v = val(ins[1])
                assert (v == 1).all(), "assert failed in the VM"
                f["op1"] = v
            elif op == "cjmp":
                c = regs[reg(ins[1])]
                f["op0"], sel["s_op0"] = c, reg(ins[1])
                tgt = val(ins[2])
                f["op1"] = tgt
                assert (c == c[0]).all() and (tgt == tgt[0]).all(), "control flow must be lane-uniform"
                nxt = int(tgt[0]) if int(c[0]) == 1 else pc + step
            elif op == "jmp":
                tgt = val(ins[1])
                f["op1"] = tgt
                assert (tgt == tgt[0]).all()
                nxt = int(tgt[0])
            else:
                raise NotImplementedError(op)
            if emit:  # generation/cpu.rs:62-178 (the column fill of tracegen.cpu_vm_trace)
                t[12, rows], t[13, rows] = clk, pc
                for i in range(10):
                    t[16 + i, rows] = pre[i]
                t[26, rows], t[27, rows], t[28, rows] = enc[0], int(step == 2), 1 << tg.OPCODE_SHIFT[op]
                t[29, rows] = enc[1] if step == 2 else 0
                t[30, rows], t[31, rows], t[32, rows], t[33, rows], t[34, rows] = f["op0"], f["op1"], f["dst"], f["aux0"], f["aux1"]
                for name, base_col in (("s_op0", 36), ("s_op1", 46), ("s_dst", 56)):
                    if name in sel:
                        t[base_col + sel[name], rows] = 1
                t[tg.CPU_SELECTOR_COL[op], rows] = 1
                t[85, rows] = 1
                t[86, rows] = 1
                t[87, rows] = 0 if op == "end" else 1
                t[92, rows] = 1 if op in ("mload", "mstore") else int(step == 2)
                self.exec_log.append((clk, 0, pc, enc[0]))
                if step == 2:
                    self.exec_log.append((clk, 1, pc + 1, enc[1]))
            s += 1
            if op == "end":
                return None, s
            pc = nxt
            if pc == stop_pc:
                return pc, s


def cpu_and_logs(n, log_n_cpu):
    """CPU table [94][2^log_n_cpu] of FIB_LOOP_ASM with bound n, plus the run's logs as flat arrays in execution order."""
    prog, labels = fib_loop_program(n)
    head = labels[".LBL10_4"]
    N = 1 << log_n_cpu
    K = n - 1
    nsteps = steps_of(n)
    assert n >= 2 and nsteps <= N, f"{nsteps} steps do not fit 2^{log_n_cpu} rows"
    t = np.zeros((94, N), dtype=U)
    run = _Run(prog, t)
    one = lambda v: np.array([int(v) % P], dtype=U)
    regs = [one(0) for _ in range(10)]
    mem = {}
    pc, s = run.segment(0, head, 1, regs, mem, 0, 0)
    assert s == PROLOGUE_STEPS and pc == head
    fp = int(regs[9][0])
    # loop-carried memory state at the head of iteration k: n, a = F(k), b = F(k+1), c = F(k+1) (the initial 2 for k = 0), i = k + 2
    F = fib_mod_p(K + 3)
    k_idx = np.arange(K, dtype=U)
    c0 = F[1:K + 1].copy()
    c0[0] = mem[fp - 4][0]
    lane_mem = {fp - 1: np.full(K, n, dtype=U), fp - 2: F[0:K].copy(), fp - 3: F[1:K + 1].copy(), fp - 4: c0, fp - 5: k_idx + U(2)}
    for a in lane_mem:
        assert lane_mem[a][0] == mem[a][0], "the analytic loop state disagrees with the prologue"
    # dry pass: registers an iteration leaves behind do not depend on the registers it starts with
    dry_regs = [np.zeros(K, dtype=U) for _ in range(10)]
    dry_regs[9] = np.full(K, fp, dtype=U)
    dry_mem = {a: v.copy() for a, v in lane_mem.items()}
    pc2, s2 = run.segment(head, head, K, dry_regs, dry_mem, PROLOGUE_STEPS, ITER_STEPS, emit=False)
    assert pc2 == head and s2 == ITER_STEPS
    lane_regs = [np.concatenate([regs[i], dry_regs[i][:-1]]) for i in range(10)]
    out_mem = {a: v.copy() for a, v in lane_mem.items()}
    pc3, s3 = run.segment(head, head, K, lane_regs, out_mem, PROLOGUE_STEPS, ITER_STEPS)
    assert pc3 == head and s3 == ITER_STEPS
    for i in range(10):
        assert (lane_regs[i] == dry_regs[i]).all(), "the register file is not a fixed point of the loop body"
    for a in lane_mem:
        assert (out_mem[a][:-1] == lane_mem[a][1:]).all(), "the analytic loop-carried memory state is inconsistent"
    regs = [r[-1:].copy() for r in lane_regs]
    mem = {a: v[-1:].copy() for a, v in out_mem.items()}
    pc4, s4 = run.segment(head, -1, 1, regs, mem, PROLOGUE_STEPS + ITER_STEPS * K, 0)
    assert pc4 is None and s4 == EXIT_STEPS
    k = nsteps
    if k != N:  # padding, generation/cpu.rs:180-208
        t[26, k:] = t[26, k - 1]
        t[35, k:] = t[35, k - 1]
        t[28, k:] = 1 << 20
        t[74, k:] = 1
        t[85, k:] = 1
        t[86, k:] = 1
        t[87, k:] = 0
        t[93, k:] = 1

    def flat(log, keycols, cols):
        """Concatenate per-slot lane arrays and sort them into execution order (clk, sub-order)."""
        parts = [[np.broadcast_to(np.asarray(e[c], dtype=U), np.asarray(e[0] if not isinstance(e[0], int) else e[1]).shape) for e in log] for c in cols]
        cat = [np.concatenate(p) if p else np.zeros(0, dtype=U) for p in parts]
        keys = [np.concatenate([np.broadcast_to(np.asarray(e[c], dtype=U), cat_shape(e)) for e in log]) if log else np.zeros(0, dtype=U) for c in keycols]
        order = np.lexsort(tuple(reversed(keys))) if log else np.zeros(0, dtype=np.int64)
        return [c[order] for c in cat]

    def cat_shape(e):
        for x in e:
            if isinstance(x, np.ndarray):
                return x.shape
        raise AssertionError

    mem_flat = flat(run.mem_log, (1, 2), (0, 1, 3, 4, 5))   # addr, clk, opmask, is_write, value  (execution order)
    cmp_flat = flat(run.cmp_log, (0,), (1, 2))               # op0, op1
    bit_flat = flat(run.bit_log, (0,), (1, 2, 3))            # opmask, a, b
    exe_flat = flat(run.exec_log, (0, 1), (2, 3))            # pc, word  (one entry per fetched word)
    return prog, t, nsteps, mem_flat, cmp_flat, bit_flat, exe_flat


# ---- side tables from the logs (vectorised restatements of the tracegen generators) --------------------------------------
def memory_trace_vec(addr, clk, op, is_write, value, log_n):
    """tracegen.memory_trace_from_log for accesses that stay in the read-write stack region.  Inputs in execution order.
    Returns (table [29][2^log_n], mem_sort range-check values)."""
    k, n = len(addr), 1 << log_n
    assert 2 <= k <= n and int(addr.max()) < P - 2 * SPAN, "stack accesses only"
    order = np.lexsort((np.arange(k), addr))  # by address, access order inside an address
    addr, clk, op, is_write, value = (x[order] for x in (addr, clk, op, is_write, value))
    t = np.zeros((29, n), dtype=U)
    new_addr = np.r_[False, addr[1:] != addr[:-1]]
    same = np.r_[False, addr[1:] == addr[:-1]]
    diff_addr = np.where(new_addr, addr - np.r_[addr[:1], addr[:-1]], U(0))
    diff_clk = np.where(same, clk - np.r_[clk[:1], clk[:-1]], U(0))
    rc_value = np.where(new_addr, diff_addr, diff_clk)
    assert int(rc_value.max()) <= 0xFFFFFFFF, "U32RangeCheckFail"
    t[2, :k], t[3, :k], t[4, :k], t[5, :k] = 1, addr, clk, op
    for mask, col in tg.MEM_OP_SELECTOR.items():
        t[col, :k] = (op == U(mask)).astype(U)
    t[17, :k], t[18, :k] = is_write, value
    t[19, :k], t[20, :k], t[21, :k] = diff_addr, inv_many(diff_addr), diff_clk
    t[23, :k] = same.astype(U)
    t[26, :k] = rc_value
    t[27, 1:k] = 1
    if k != n:  # generation/memory.rs:113-146: padding continues into the write-once region from p - span
        pad = np.arange(n - k, dtype=U) + U(P - SPAN)
        assert int(pad[-1]) < P, "the prophet region is full"
        t[16, k:], t[3, k:], t[17, k:] = 1, pad, 1
        d = np.ones(n - k, dtype=U)
        d[0] = U((int(pad[0]) - int(addr[-1])) % P)
        t[19, k:], t[20, k:] = d, inv_many(d)
        t[22, k:] = U(P) - pad
        t[24, k:] = 1
        t[26, k:] = t[22, k:]
    return t, rc_value[1:].copy()


def cmp_trace_vec(a, b, log_n):
    """tracegen.cmp_trace."""
    n, k = 1 << log_n, len(a)
    assert k <= n
    t = np.zeros((6, n), dtype=U)
    t[2, :] = 1
    d = np.where(a >= b, a - b, b - a)
    t[0, :k], t[1, :k], t[2, :k], t[3, :k], t[4, :k], t[5, :k] = a, b, (a >= b).astype(U), d, inv_many(d), 1
    return t, d


def rangecheck_trace_vec(cmp_vals, log_n, cpu_vals=(), mem_sort_vals=(), mem_region_vals=()):
    """tracegen.rangecheck_trace (generate_rc_trace, generation/builtin.rs:249-316)."""
    n = 1 << log_n
    assert n >= 1 << 16
    t = np.zeros((12, n), dtype=U)
    row = 0
    for col, vals in ((0, cpu_vals), (1, mem_sort_vals), (2, mem_region_vals), (3, cmp_vals)):
        vals = np.asarray(vals, dtype=U)
        m = len(vals)
        assert row + m <= n
        t[col, row:row + m] = 1
        t[4, row:row + m], t[5, row:row + m], t[6, row:row + m] = vals, vals & U(0xFFFF), vals >> U(16)
        row += m
    fix = np.minimum(np.arange(n, dtype=U), U(65535))
    t[9] = fix
    t[7], t[10] = permuted_cols_vec(t[5], fix)
    t[8], t[11] = permuted_cols_vec(t[6], fix)
    return t


def _compress_many(beta, cols):
    """sum_i cols[i] * beta^i per row, over the distinct rows (Python integers)."""
    stack = np.stack([np.asarray(c, dtype=U) for c in cols], axis=1)
    uniq, back = np.unique(stack, axis=0, return_inverse=True)
    pw = [pow(beta, i, P) for i in range(len(cols))]
    out = np.array([sum(int(x) * pw[i] for i, x in enumerate(r)) % P for r in uniq], dtype=U)
    return out[back.reshape(-1)]


def bitwise_trace_vec(tag, a, b, beta, log_n):
    """tracegen.bitwise_valid_trace with ops = zip(tag, a, b) (generate_bitwise_trace, generation/builtin.rs)."""
    n, k = 1 << log_n, len(tag)
    assert n >= 256 and k <= n
    t = np.zeros((59, n), dtype=U)
    AND, OR = U(1 << tg.OPCODE_SHIFT["and"]), U(1 << tg.OPCODE_SHIFT["or"])
    r = np.where(tag == AND, a & b, np.where(tag == OR, a | b, a ^ b))
    t[0, :k], t[1, :k], t[2, :k], t[3, :k], t[4, :k] = 1, tag, a, b, r
    fixed = [np.zeros((1, 4), dtype=U)]
    for j in range(4):
        sh = U(8 * j)
        la, lb, lr = (a >> sh) & U(255), (b >> sh) & U(255), (r >> sh) & U(255)
        t[5 + j, :k], t[9 + j, :k], t[13 + j, :k] = la, lb, lr
        t[29 + j, :k] = _compress_many(beta, (tag, la, lb, lr)) if k else 0
        fixed.append(np.stack([tag, la, lb, lr], axis=1))
    fixed = np.unique(np.concatenate(fixed), axis=0)  # lexicographic, as sorted(set of tuples)
    assert len(fixed) <= n
    fix = np.minimum(np.arange(n, dtype=U), U(255))
    t[37] = fix
    for j in range(4):
        t[17 + j], t[38 + j] = permuted_cols_vec(t[5 + j], fix)
        t[21 + j], t[42 + j] = permuted_cols_vec(t[9 + j], fix)
        t[25 + j], t[46 + j] = permuted_cols_vec(t[13 + j], fix)
    m = len(fixed)
    t[50:54, :m] = fixed.T
    t[54, :m] = _compress_many(beta, tuple(fixed.T))
    for j in range(4):
        t[33 + j], t[55 + j] = permuted_cols_vec(t[29 + j], t[54])
    return t


def program_trace_vec(prog_words, exec_pc, exec_word, beta, log_n):
    """tracegen.program_valid_trace for one program at code address 0: program lines (0,0,0,0,pc,word) and one executed
    line per fetched word (generate_prog_trace, generation/prog.rs:56-131)."""
    n, m, k = 1 << log_n, len(prog_words), len(exec_pc)
    assert m < n and k <= n
    t = np.zeros((18, n), dtype=U)
    z = np.zeros(m, dtype=U)
    pcs, words = np.arange(m, dtype=U), np.array([int(w) % P for w in prog_words], dtype=U)
    t[4, :m], t[5, :m], t[17, :m] = pcs, words, 1
    t[6, :m] = _compress_many(beta, (z, z, z, z, pcs, words))
    t[12, :k], t[13, :k], t[16, :k] = exec_pc, exec_word, 1
    # an executed line is a program line: its compressed value is the program line's
    assert (words[exec_pc.astype(np.int64)] == exec_word).all()
    t[14, :k] = t[6, :m][exec_pc.astype(np.int64)]
    t[15], t[7] = permuted_cols_vec(t[14], t[6])
    return t


def _log2_at_least(k, lo):
    return max(lo, (max(int(k), 1) - 1).bit_length())


def fib_loop_system(n, hasher, log_n_cpu=None, beta=0x1234567890ABCDEF % P, beta_bitwise=0x0FEDCBA987654321 % P, seed=0):
    """The full 12-table system of one run of FIB_LOOP_ASM with loop bound n.  `hasher` supplies `.poseidon(state[12])`
    and `.poseidon_table_row(input[12])` (the oracle in tests, the product's device entry points in bench.py).
    Returns (table_ids 0..11, traces, compress_challenges, info)."""
    rng = np.random.default_rng(seed)
    if log_n_cpu is None:
        log_n_cpu = _log2_at_least(steps_of(n), 4)
    prog, cpu_t, nsteps, (m_addr, m_clk, m_op, m_w, m_val), (c_a, c_b), (b_tag, b_a, b_b), (e_pc, e_word) = cpu_and_logs(n, log_n_cpu)
    mem_t, rc_sort = memory_trace_vec(m_addr, m_clk, m_op, m_w, m_val, _log2_at_least(len(m_addr) + 1, 2))
    cmp_t, rc_cmp = cmp_trace_vec(c_a, c_b, _log2_at_least(len(c_a) + 1, 4))
    rc_t = rangecheck_trace_vec(rc_cmp, _log2_at_least(len(rc_cmp) + len(rc_sort), 16), mem_sort_vals=rc_sort)
    bw_t = bitwise_trace_vec(b_tag, b_a, b_b, beta_bitwise, _log2_at_least(len(b_tag) + 1, 9))
    words = []
    for ins in prog:
        words += tg.ola_encode(ins)
    pt = program_trace_vec(words, e_pc, e_word, beta, _log2_at_least(max(len(words), len(e_pc)) + 1, 2))
    # the program's hash trio: ProgChunk absorbs the program's words, its digest is read from the storage tree at code
    # address 0, every sponge / Merkle hash is a Poseidon-table row
    chunk_log = max(3, ((len(words) + 7) // 8 - 1).bit_length())
    pc_t, psdn_prog, lines, roots = tg.prog_chunk_valid_trace(hasher, rng, chunk_log, programs=[([0, 0, 0, 0], words)])
    _, leaf = roots[0]
    st_t, psdn_st = tg.storage_valid_trace(hasher, rng, 8, [dict(addr_bits=[0] * 256, leaf=leaf, pre_leaf=leaf, is_write=0, for_prog=1)])
    rows = [(inp, [1, 0, 0, 0]) for inp, _ in psdn_prog]
    rows += [(inp, [0, 0, 1, 0] if is_leaf else [0, 0, 0, 1]) for inp, _, is_leaf in psdn_st]
    ps_t = tg.poseidon_valid_trace(hasher, _log2_at_least(len(rows) + 1, 4), rows)
    # tables the run does not touch: the generators' output for empty input
    pch_t = np.zeros((53, 4), dtype=U)   # generate_poseidon_chunk_trace (generation/poseidon_chunk.rs:7-88): is_padding rows
    pch_t[52] = 1
    tape_t = np.zeros((6, 4), dtype=U)   # generate_tape_trace (generation/tape.rs:10-73): unlooked tload rows
    tape_t[2] = 1 << 9
    sc_t = tg.sccall_valid_trace(rng, 2, used=0)   # generate_sccall_trace (generation/sccall.rs): padding rows
    ids = list(range(12))
    traces = [cpu_t, mem_t, bw_t, cmp_t, rc_t, ps_t, pch_t, st_t, tape_t, sc_t, pt, pc_t]
    cc = [0, 0, beta_bitwise, 0, 0, 0, 0, 0, 0, 0, beta, 0]
    info = dict(loop_bound=int(n), cpu_steps=int(nsteps), table_log_n=[int(t.shape[1]).bit_length() - 1 for t in traces],
                memory_accesses=int(len(m_addr)), cmp_rows=int(len(c_a)), bitwise_rows=int(len(b_tag)), fetched_words=int(len(e_pc)))
    return ids, traces, cc, info
