"""Workload side of the `ola prove` flow: the executor records of a real VM run as the flat arrays the
ola_generate_* entry points take, and the same records written as the JSON text `ola run` produces -- serde's layout of
core::trace::trace::Trace (core/src/trace/trace.rs:320-342; client/src/main.rs:166-169).  Inputs only: the product parses the
text (ola_trace_from_json), generates the tables on the GPU and proves them."""
import json

import numpy as np

from . import tracegen as tg

P = 0xFFFFFFFF00000001


def system_records(orc, rng, n_iter=6, cpu_log=9, mem_log_n=7):
    """One run of the calls / bitwise / poseidon / tape program (tracegen.calls_program) -> dict of flat records:
    steps [k,66], memory [k,15], rc_vals / rc_kinds [k], bw_tags / bw_op0 / bw_op1 / bw_res [k], cmp [k,6], psdn_inputs [k,12],
    psdn_filters [k,4], pchunk [k,32], storage [k,38] + n_storage_access, tape [k,5], sccall [0,24], prog_rows [m,6], roots [8].
    The program's digest is read from a sparse Merkle tree at code address 0, so the storage / poseidon lookups carry real data."""
    prog = tg.calls_program(n_iter, linear=False, bitwise=True, poseidon=True, tape=True)
    _, steps, cmp_pairs, rc_cmp, rc_cpu, mlog, bit_ops, psdn_calls, tape_log = tg.cpu_vm_trace(prog, cpu_log, want_side_tables="all+tape", orc=orc)
    _, cells = tg.memory_trace_from_log(mlog, mem_log_n, want_cells=True)
    out = tg.memory_trace_from_log(mlog, mem_log_n)
    rc_sort, rc_region = out[1], (out[2] if len(out) == 3 else [])
    prog_rows, _ = tg.program_rows_of_run(prog, steps)
    words = [r[5] for r in prog_rows]
    chunk_log = max(1, ((len(words) + 7) // 8 - 1).bit_length())
    _, psdn_prog, lines, roots = tg.prog_chunk_valid_trace(orc, rng, chunk_log, programs=[([0, 0, 0, 0], words)])
    assert lines == prog_rows
    _, leaf = roots[0]
    st, psdn_st = tg.storage_valid_trace(orc, rng, 8, [dict(addr_bits=[0] * 256, leaf=leaf, pre_leaf=leaf, is_write=0, for_prog=1)])
    pch, psdn_chunk = tg.poseidon_chunk_trace_from_calls(psdn_calls, max(1, (sum(len(c["rows"]) for c in psdn_calls) - 1).bit_length()))
    hashes = [(inp, [1, 0, 0, 0]) for inp, _ in psdn_prog]
    hashes += [(inp, [0, 0, 1, 0] if is_leaf else [0, 0, 0, 1]) for inp, _, is_leaf in psdn_st]
    hashes += [(inp, [1, 0, 0, 0]) for inp, _ in psdn_chunk]
    cmp_cells = []
    for a, b in cmp_pairs:
        d = abs(a - b)
        cmp_cells.append([a, b, 1 if a >= b else 0, d, pow(d, P - 2, P) if d else 0, 1])
    rc = [(v, 0) for v in rc_cpu] + [(v, 1) for v in rc_sort] + [(v, 2) for v in rc_region] + [(v, 3) for v in rc_cmp]
    storage, n_access = tg.storage_records_of_table(st)
    u = lambda x, shape: np.array(x, dtype=np.uint64).reshape(shape)
    return dict(steps=tg.steps_to_records(steps), memory=tg.memory_cells_to_records(cells),
                rc_vals=u([v for v, _ in rc], -1), rc_kinds=u([k for _, k in rc], -1),
                bw_tags=u([t for t, _, _ in bit_ops], -1), bw_op0=u([a for _, a, _ in bit_ops], -1), bw_op1=u([b for _, _, b in bit_ops], -1),
                bw_res=u([(a & b) if t == tg.OP_AND else ((a | b) if t == tg.OP_OR else (a ^ b)) for t, a, b in bit_ops], -1),
                cmp=u(cmp_cells, (-1, 6)), psdn_inputs=u([i for i, _ in hashes], (-1, 12)), psdn_filters=u([f for _, f in hashes], (-1, 4)),
                pchunk=tg.poseidon_chunk_records_of_table(pch), storage=storage, n_storage_access=n_access,
                tape=tg.tape_records_from_log(tape_log), sccall=np.zeros((0, 24), dtype=np.uint64), prog_rows=u(prog_rows, (-1, 6)),
                roots=u(list(st[5:9, 0]) * 2, 8))


def _ints(a):
    return [int(x) for x in a]


def records_to_json(rec, orc=None):
    """The records as serde_json writes a Trace: every field of every struct in declaration order, including the ones the
    generators do not read (PoseidonRow's round states -- real ones when `orc` is given --, RangeCheckRow's limbs, the
    instruction maps, builtin_storage, ret), so that the parser's skipping is exercised on the real shape of the file."""
    steps = []
    for s in rec["steps"]:
        s = _ints(s)
        steps.append(dict(env_idx=s[0], call_sc_cnt=s[1], clk=s[11], pc=s[12], tp=s[10], addr_storage=s[2:6], addr_code=s[6:10], instruction=s[25],
                          immediate_data=s[28], opcode=s[27], op1_imm=s[26], regs=s[15:25],
                          register_selector=dict(op0=s[29], op1=s[30], dst=s[31], aux0=s[32], aux1=s[33], op0_reg_sel=s[35:45], op1_reg_sel=s[45:55],
                                                 dst_reg_sel=s[55:65]),
                          is_ext_line=s[13], ext_cnt=s[14], filter_tape_looking=s[65], storage_access_idx=s[34]))
    memory = []
    for c in rec["memory"]:
        c = _ints(c)
        memory.append(dict(env_idx=c[0], addr=c[2], clk=c[3], is_rw=c[1], op=c[4], is_write=c[5], diff_addr=c[7], diff_addr_inv=c[8], diff_clk=c[9],
                           diff_addr_cond=c[10], filter_looked_for_main=1, rw_addr_unchanged=c[11], region_prophet=c[12], region_heap=c[13], value=c[6],
                           rc_value=c[14]))
    rcs = []
    for v, k in zip(_ints(rec["rc_vals"]), _ints(rec["rc_kinds"])):
        rcs.append(dict(val=v, limb_lo=v & 0xFFFF, limb_hi=v >> 16, filter_looked_for_mem_sort=int(k == 1), filter_looked_for_mem_region=int(k == 2),
                        filter_looked_for_cpu=int(k == 0), filter_looked_for_comparison=int(k == 3), filter_looked_for_storage=0))
    bws = []
    for t, a, b, r in zip(_ints(rec["bw_tags"]), _ints(rec["bw_op0"]), _ints(rec["bw_op1"]), _ints(rec["bw_res"])):
        row = dict(opcode=t, op0=a, op1=b, res=r)
        for name, v in (("op0", a), ("op1", b), ("res", r)):
            for j in range(4):
                row["%s_%d" % (name, j)] = (v >> (8 * j)) & 255
        bws.append(row)
    cmps = [dict(zip(("op0", "op1", "gte", "abs_diff", "abs_diff_inv", "filter_looking_rc"), _ints(c))) for c in rec["cmp"]]
    psdn = []
    for inp, f in zip(rec["psdn_inputs"], rec["psdn_filters"]):
        row = orc.poseidon_table_row(inp) if orc is not None else np.zeros(134, dtype=np.uint64)
        psdn.append(dict(input=_ints(inp), full_0_1=_ints(row[28:40]), full_0_2=_ints(row[40:52]), full_0_3=_ints(row[52:64]), partial=_ints(row[64:86]),
                         full_1_0=_ints(row[86:98]), full_1_1=_ints(row[98:110]), full_1_2=_ints(row[110:122]), full_1_3=_ints(row[122:134]),
                         output=_ints(row[16:28]), filter_looked_normal=bool(f[0]), filter_looked_treekey=bool(f[1]), filter_looked_storage=bool(f[2]),
                         filter_looked_storage_branch=bool(f[3])))
    pchunk = []
    for c in rec["pchunk"]:
        c = _ints(c)
        pchunk.append(dict(env_idx=c[0], clk=c[1], opcode=c[2], dst=c[3], op0=c[4], op1=c[5], acc_cnt=c[6], value=c[7:15], cap=c[15:19], hash=c[19:31],
                           is_ext_line=c[31]))

    def storage_row(c):
        c = _ints(c)
        return dict(storage_access_idx=c[0], pre_root=c[1:5], root=c[5:9], is_write=c[9], layer=c[10], layer_bit=c[11], addr_acc=c[12], addr=c[13:17],
                    pre_path=c[17:21], path=c[21:25], hash_type=c[25], pre_hash=c[26:30], hash=c[30:34], sibling=c[34:38])

    na = rec["n_storage_access"]
    tape = [dict(is_init=bool(c[0]), opcode=int(c[1]), addr=int(c[2]), value=int(c[3]), filter_looked=int(c[4])) for c in rec["tape"]]
    sccall = []
    for c in rec["sccall"]:
        c = _ints(c)
        sccall.append(dict(caller_env_idx=c[0], addr_storage=c[1:5], addr_code=c[5:9], caller_op1_imm=c[9], clk_caller_call=c[10], clk_caller_ret=c[11],
                           regs=c[12:22], callee_env_idx=c[22], clk_callee_end=c[23]))
    programs = {}
    for r in rec["prog_rows"]:
        r = _ints(r)
        key = "".join("%016x" % x for x in r[0:4])  # encode_addr (core/src/types/merkle_tree/mod.rs:176-185)
        programs.setdefault(key, []).append(r[5])
    words = [w for p in programs.values() for w in p]
    doc = dict(instructions={str(i): ["mov r0 \"%d\"" % i, 0, 1, w, 0] for i, w in enumerate(words[:4])},
               raw_instructions={str(i): "add r1 r2 r3" for i in range(min(4, len(words)))},
               raw_binary_instructions=["0x%016x" % w for w in words], addr_program_hash=programs,
               start_end_roots=[_ints(rec["roots"][0:4]), _ints(rec["roots"][4:8])], exec=steps, memory=memory, builtin_rangecheck=rcs,
               builtin_bitwise_combined=bws, builtin_cmp=cmps, builtin_poseidon=psdn, builtin_poseidon_chunk=pchunk,
               builtin_storage=[dict(env_idx=0, clk=7, diff_clk=0, opcode=1 << 11, root=[1, 2, 3, 4], addr=[0, 0, 0, 0], value=[5, 6, 7, 8])],
               builtin_storage_hash=[storage_row(c) for c in rec["storage"][:na]], builtin_program_hash=[storage_row(c) for c in rec["storage"][na:]],
               tape=tape, sc_call=sccall, ret=[1, 2, 3])
    return json.dumps(doc, separators=(",", ":"))


def records_of_fib_system(traces, info):
    """The executor records behind workload.fibloop.fib_loop_system's twelve tables, read back out of the tables (every record
    field is a table column: generate_cpu_trace copies Step field i to column i + 1, generate_memory_trace copies the cell into
    columns 1..5 and 17..26, and so on), so that the `ola prove` flow can be driven at BASELINE configs[2]'s scale: records in,
    tables generated on the GPU.  `info` = the dict fib_loop_system returns (filled row counts)."""
    cpu_t, mem_t, bw_t, cmp_t, rc_t, ps_t, pch_t, st_t, tape_t, sc_t, pt, pc_t = traces
    k = info["cpu_steps"]
    steps = np.empty((k, 66), dtype=np.uint64)
    steps[:, 0:65] = cpu_t[1:66, :k].T
    steps[:, 65] = cpu_t[88, :k]
    k = info["memory_accesses"]
    memory = np.empty((k, 15), dtype=np.uint64)
    memory[:, 0:5], memory[:, 5:15] = mem_t[1:6, :k].T, mem_t[17:27, :k].T
    k = info["bitwise_rows"]
    kc = info["cmp_rows"]
    live = rc_t[0:4].any(axis=0)
    kr = int(live.sum())
    assert live[:kr].all()
    kinds = np.argmax(rc_t[0:4, :kr] != 0, axis=0).astype(np.uint64)
    live = ps_t[0:4].any(axis=0)
    kp = int(live.sum())
    assert live[:kp].all()
    storage, n_access = tg.storage_records_of_table(st_t)
    lines = int((pc_t[39] == 0).sum())
    prog_rows = [tuple(int(pc_t[c, i]) for c in range(4)) + (int(pc_t[4, i]) + j, int(pc_t[5 + j, i])) for i in range(lines) for j in range(8) if pc_t[31 + j, i]]
    c = np.ascontiguousarray
    return dict(steps=steps, memory=memory, rc_vals=c(rc_t[4, :kr]), rc_kinds=kinds, bw_tags=c(bw_t[1, :k]), bw_op0=c(bw_t[2, :k]), bw_op1=c(bw_t[3, :k]),
                bw_res=c(bw_t[4, :k]), cmp=c(cmp_t[:, :kc].T), psdn_inputs=c(ps_t[4:16, :kp].T), psdn_filters=c(ps_t[0:4, :kp].T),
                pchunk=np.zeros((0, 32), dtype=np.uint64), storage=storage, n_storage_access=n_access, tape=np.zeros((0, 5), dtype=np.uint64),
                sccall=np.zeros((0, 24), dtype=np.uint64), prog_rows=np.array(prog_rows, dtype=np.uint64).reshape(-1, 6),
                roots=np.array(list(st_t[5:9, 0]) * 2, dtype=np.uint64))


REC_KIND_OF = dict(steps="step", memory="memory", rc_vals="rc_val", rc_kinds="rc_kind", bw_tags="bitwise_tag", bw_op0="bitwise_op0", bw_op1="bitwise_op1",
                   bw_res="bitwise_res", cmp="cmp", psdn_inputs="poseidon_input", psdn_filters="poseidon_filter", pchunk="poseidon_chunk",
                   storage="storage_hash", tape="tape", sccall="sccall", prog_rows="prog_row", roots="roots", n_storage_access="storage_access_count")
