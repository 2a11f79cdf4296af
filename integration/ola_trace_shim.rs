//! Seam F: `generate_traces` + `prove_with_traces` on the GPU from the executor's records -- the body of
//! `circuits::stark::prover::prove` (prover.rs:43-66) and of `ola prove` (client/src/main.rs:172-207).
//! Drop into the reference as circuits/src/stark/gpu_trace.rs (`pub mod gpu_trace;` in stark/mod.rs) next to gpu_prover.rs
//! (integration/ola_prover_shim.rs).  The host keeps serde and its own `Trace`; each `Vec<Row>` is flattened once into the
//! record layout documented in include/ola_gpu.h (struct field order, `.0` of every GoldilocksField) and handed to the
//! library WITHOUT a copy (`ola_trace_set_records` borrows).  All twelve tables are then generated in device memory and proved
//! there; what comes back is the proof in the wire format (`Buffer::read_all_proof`, serialization.rs:395-411).
//!
//! This file cannot be compiled in the build container of this repository (no cargo); it is the binding a maintainer adds.
//! The same flow is exercised here from C (tests/c/ola_prove_file.c) and through ctypes (tests/test_trace_json.py).
use std::ptr;

use anyhow::{ensure, Result};
use core::trace::trace::{
    BitwiseCombinedRow, CmpRow, MemoryTraceCell, PoseidonChunkRow, PoseidonRow, RangeCheckRow, SCCallRow, Step, StorageHashRow, TapeRow, Trace,
};
use core::types::merkle_tree::decode_addr;
use plonky2::fri::ola_gpu::*;
use plonky2::util::serialization::Buffer;

use crate::stark::gpu_prover::{ctx, last_error, C, D, F};
use crate::stark::proof::AllProof;

fn step(s: &Step, out: &mut Vec<u64>) {
    // 0 env_idx 1 call_sc_cnt 2..5 addr_storage 6..9 addr_code 10 tp 11 clk 12 pc 13 is_ext_line 14 ext_cnt 15..24 regs 25 instruction
    // 26 op1_imm 27 opcode 28 immediate_data 29..33 op0 op1 dst aux0 aux1 34 storage_access_idx 35..64 the three reg selectors 65 filter_tape_looking
    out.extend([s.env_idx.0, s.call_sc_cnt.0]);
    out.extend(s.addr_storage.iter().map(|x| x.0));
    out.extend(s.addr_code.iter().map(|x| x.0));
    out.extend([s.tp.0, s.clk as u64, s.pc, s.is_ext_line.0, s.ext_cnt.0]);
    out.extend(s.regs.iter().map(|x| x.0));
    out.extend([s.instruction.0, s.op1_imm.0, s.opcode.0, s.immediate_data.0]);
    let r = &s.register_selector;
    out.extend([r.op0.0, r.op1.0, r.dst.0, r.aux0.0, r.aux1.0, s.storage_access_idx.0]);
    out.extend(r.op0_reg_sel.iter().chain(&r.op1_reg_sel).chain(&r.dst_reg_sel).map(|x| x.0));
    out.push(s.filter_tape_looking.0);
}
fn memory(c: &MemoryTraceCell, out: &mut Vec<u64>) {
    out.extend([c.env_idx.0, c.is_rw.0, c.addr.0, c.clk.0, c.op.0, c.is_write.0, c.value.0, c.diff_addr.0, c.diff_addr_inv.0, c.diff_clk.0,
                c.diff_addr_cond.0, c.rw_addr_unchanged.0, c.region_prophet.0, c.region_heap.0, c.rc_value.0]);
}
fn rc_kind(c: &RangeCheckRow) -> u64 {
    // which table looks the value up: 0 cpu, 1 memory sort, 2 memory region, 3 comparison, 4 none (generate_rc_trace has no fifth filter column)
    if c.filter_looked_for_cpu.0 == 1 { 0 } else if c.filter_looked_for_mem_sort.0 == 1 { 1 } else if c.filter_looked_for_mem_region.0 == 1 { 2 }
    else if c.filter_looked_for_comparison.0 == 1 { 3 } else { 4 }
}
fn poseidon_chunk(c: &PoseidonChunkRow, out: &mut Vec<u64>) {
    out.extend([c.env_idx.0, c.clk as u64, c.opcode.0, c.dst.0, c.op0.0, c.op1.0, c.acc_cnt.0]);
    out.extend(c.value.iter().chain(&c.cap).chain(&c.hash).map(|x| x.0));
    out.push(c.is_ext_line.0);
}
fn storage_hash(c: &StorageHashRow, out: &mut Vec<u64>) {
    out.push(c.storage_access_idx);
    out.extend(c.pre_root.iter().chain(&c.root).map(|x| x.0));
    out.extend([c.is_write.0, c.layer, c.layer_bit, c.addr_acc.0]);
    out.extend(c.addr.iter().chain(&c.pre_path).chain(&c.path).map(|x| x.0));
    out.push(c.hash_type.0);
    out.extend(c.pre_hash.iter().chain(&c.hash).chain(&c.sibling).map(|x| x.0));
}
fn sccall(c: &SCCallRow, out: &mut Vec<u64>) {
    out.push(c.caller_env_idx.0);
    out.extend(c.addr_storage.iter().chain(&c.addr_code).map(|x| x.0));
    out.extend([c.caller_op1_imm.0, c.clk_caller_call.0, c.clk_caller_ret.0]);
    out.extend(c.regs.iter().map(|x| x.0));
    out.extend([c.callee_env_idx.0, c.clk_callee_end.0]);
}

fn flat<T>(rows: &[T], width: usize, f: impl Fn(&T, &mut Vec<u64>)) -> Vec<u64> {
    let mut out = Vec::with_capacity(rows.len() * width);
    rows.iter().for_each(|r| f(r, &mut out));
    debug_assert_eq!(out.len(), rows.len() * width);
    out
}

/// `prove` (prover.rs:43-66): `generate_traces(program, ...)` + `prove_with_traces(...)`, both on the GPU.
pub fn prove_trace_gpu(trace: &Trace) -> Result<AllProof<F, C, D>> {
    let guard = ctx()?.lock().unwrap();
    let c = guard.0;
    // every array below must outlive the ola_trace object: the library borrows them
    let steps = flat(&trace.exec, 66, step);
    let mem = flat(&trace.memory, 15, memory);
    let rc_val: Vec<u64> = trace.builtin_rangecheck.iter().map(|r| r.val.0).collect();
    let rc_kinds: Vec<u64> = trace.builtin_rangecheck.iter().map(rc_kind).collect();
    let bw = |f: fn(&BitwiseCombinedRow) -> u64| trace.builtin_bitwise_combined.iter().map(f).collect::<Vec<u64>>();
    let (bw_tag, bw_op0, bw_op1, bw_res) = (bw(|r| r.opcode), bw(|r| r.op0.0), bw(|r| r.op1.0), bw(|r| r.res.0));
    let cmp = flat(&trace.builtin_cmp, 6, |r: &CmpRow, o| o.extend([r.op0.0, r.op1.0, r.gte.0, r.abs_diff.0, r.abs_diff_inv.0, r.filter_looking_rc.0]));
    let ps_in = flat(&trace.builtin_poseidon, 12, |r: &PoseidonRow, o| o.extend(r.input.iter().map(|x| x.0)));
    let ps_f = flat(&trace.builtin_poseidon, 4, |r: &PoseidonRow, o| {
        o.extend([r.filter_looked_normal, r.filter_looked_treekey, r.filter_looked_storage, r.filter_looked_storage_branch].map(|b| b as u64))
    });
    let pchunk = flat(&trace.builtin_poseidon_chunk, 32, poseidon_chunk);
    let mut st = flat(&trace.builtin_storage_hash, 38, storage_hash); // the accesses, then the program-hash reads (storage.rs:23)
    st.extend(flat(&trace.builtin_program_hash, 38, storage_hash));
    let tape = flat(&trace.tape, 5, |r: &TapeRow, o| o.extend([r.is_init as u64, r.opcode.0, r.addr.0, r.value.0, r.filter_looked.0]));
    let sc = flat(&trace.sc_call, 24, sccall);
    let mut prog_rows: Vec<u64> = Vec::new(); // (code address 0..3, pc, word) in the order generate_traces walks addr_program_hash
    for (addr, words) in trace.addr_program_hash.iter() {
        let a = decode_addr(addr.clone());
        for (pc, w) in words.iter().enumerate() {
            prog_rows.extend([a[0].0, a[1].0, a[2].0, a[3].0, pc as u64, w.0]);
        }
    }
    let (start, end) = trace.start_end_roots;
    let roots: Vec<u64> = start.iter().chain(end.iter()).map(|x| x.0).collect();

    let mut t: *mut ola_trace = ptr::null_mut();
    ensure!(unsafe { ola_trace_new(&mut t) } == OLA_OK, "ola_trace_new failed");
    let set = |kind: i32, v: &Vec<u64>, n: usize| unsafe { ola_trace_set_records(t, kind, v.as_ptr(), n) };
    let mut rc = 0;
    rc |= set(OLA_REC_STEP, &steps, trace.exec.len());
    rc |= set(OLA_REC_MEMORY, &mem, trace.memory.len());
    rc |= set(OLA_REC_RC_VAL, &rc_val, rc_val.len());
    rc |= set(OLA_REC_RC_KIND, &rc_kinds, rc_kinds.len());
    rc |= set(OLA_REC_BITWISE_TAG, &bw_tag, bw_tag.len());
    rc |= set(OLA_REC_BITWISE_OP0, &bw_op0, bw_op0.len());
    rc |= set(OLA_REC_BITWISE_OP1, &bw_op1, bw_op1.len());
    rc |= set(OLA_REC_BITWISE_RES, &bw_res, bw_res.len());
    rc |= set(OLA_REC_CMP, &cmp, trace.builtin_cmp.len());
    rc |= set(OLA_REC_POSEIDON_INPUT, &ps_in, trace.builtin_poseidon.len());
    rc |= set(OLA_REC_POSEIDON_FILTER, &ps_f, trace.builtin_poseidon.len());
    rc |= set(OLA_REC_POSEIDON_CHUNK, &pchunk, trace.builtin_poseidon_chunk.len());
    rc |= set(OLA_REC_STORAGE_HASH, &st, trace.builtin_storage_hash.len() + trace.builtin_program_hash.len());
    rc |= unsafe { ola_trace_set_records(t, OLA_REC_STORAGE_ACCESS_COUNT, ptr::null(), trace.builtin_storage_hash.len()) };
    rc |= set(OLA_REC_TAPE, &tape, trace.tape.len());
    rc |= set(OLA_REC_SCCALL, &sc, trace.sc_call.len());
    rc |= set(OLA_REC_PROG_ROW, &prog_rows, prog_rows.len() / 6);
    rc |= set(OLA_REC_ROOTS, &roots, 1);
    if rc != OLA_OK {
        unsafe { ola_trace_free(t) };
        anyhow::bail!("ola_trace_set_records failed");
    }
    let mut bytes = vec![0u8; 1 << 24];
    let mut len = 0usize;
    let rc = unsafe { ola_prove_trace(c, t, bytes.as_mut_ptr(), bytes.len(), &mut len) };
    unsafe { ola_trace_free(t) };
    // OLA_ERR_QUOTIENT_DEGREE is the reference's panic "Quotient has failed, the vanishing polynomial is not divisible by Z_H"
    ensure!(rc == OLA_OK, "ola_prove_trace failed with {rc}: {}", last_error(c));
    bytes.truncate(len);
    let mut buffer = Buffer::new(bytes);
    buffer.read_all_proof::<F, C, D>().map_err(|e| anyhow::anyhow!("proof bytes do not parse: {e:?}"))
}
