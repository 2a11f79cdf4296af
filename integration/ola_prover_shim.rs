//! Seam A: `circuits::stark::prover::prove_with_traces` on the GPU.
//! Drop into the reference as circuits/src/stark/gpu_prover.rs (`pub mod gpu_prover;` in stark/mod.rs) next to
//! plonky2::fri::ola_gpu (integration/ola_gpu.rs).  Signature and error behaviour follow prover.rs:79-85:
//! the same `[Vec<PolynomialValues<F>>; NUM_TABLES]` in, an `AllProof<F, C, D>` out (read back from the wire format the
//! library writes, serialization.rs:395-411), `Err` where the reference returns / panics.
//!
//! This file cannot be compiled in the build container of this repository (no cargo); it is the binding a maintainer adds.
use std::ptr;

use anyhow::{anyhow, ensure, Result};
use once_cell::sync::OnceCell;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::fri::ola_gpu::*;
use plonky2::plonk::config::{GenericConfig, PoseidonGoldilocksConfig};
use plonky2::util::serialization::Buffer;

use crate::stark::ola_stark::{OlaStark, NUM_TABLES};
use crate::stark::proof::AllProof;

pub(crate) type F = GoldilocksField;
pub(crate) type C = PoseidonGoldilocksConfig;
pub(crate) const D: usize = 2;

/// One context per process and GPU: replaces `init_gpu()` / `free_gpu()` (plonky2/field/src/cfft/ntt/mod.rs:55-101).
pub(crate) struct Ctx(pub(crate) *mut ola_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
static CTX: OnceCell<std::sync::Mutex<Ctx>> = OnceCell::new();

pub(crate) fn ctx() -> Result<&'static std::sync::Mutex<Ctx>> {
    CTX.get_or_try_init(|| {
        let mut p: *mut ola_ctx = ptr::null_mut();
        let device = std::env::var("OLA_GPU_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let rc = unsafe { ola_gpu_init(device, &mut p) };
        ensure!(rc == OLA_OK, "ola_gpu_init failed with {rc}: a B200 (sm_100) GPU is required, there is no CPU fallback");
        Ok(std::sync::Mutex::new(Ctx(p)))
    })
}

pub(crate) fn last_error(c: *mut ola_ctx) -> String {
    unsafe { std::ffi::CStr::from_ptr(ola_gpu_last_error(c)).to_string_lossy().into_owned() }
}

/// `prove_with_traces` (prover.rs:79-327) + `Buffer::read_all_proof`.
pub fn prove_with_traces_gpu(
    ola_stark: &OlaStark<F, D>,
    trace_poly_values: &[Vec<PolynomialValues<F>>; NUM_TABLES],
) -> Result<AllProof<F, C, D>> {
    let guard = ctx()?.lock().unwrap(); // a context is not re-entrant (the reference serialises its GPU the same way, ntt/mod.rs:48-50)
    let c = guard.0;
    // Vec<PolynomialValues<F>> is column-major already; GoldilocksField is #[repr(transparent)] over u64.  The C ABI takes
    // one contiguous [columns][rows] block per table: gather the column Vecs (one memcpy per column).
    let mut blocks: Vec<Vec<u64>> = Vec::with_capacity(NUM_TABLES);
    let mut log_ns = [0u32; NUM_TABLES];
    for (t, cols) in trace_poly_values.iter().enumerate() {
        let n = cols[0].len();
        ensure!(n.is_power_of_two(), "trace length must be a power of two");
        log_ns[t] = n.trailing_zeros();
        let mut b = Vec::with_capacity(cols.len() * n);
        for col in cols {
            ensure!(col.len() == n, "ragged trace");
            b.extend(col.values.iter().map(|x| x.0));
        }
        blocks.push(b);
    }
    let ptrs: Vec<*const u64> = blocks.iter().map(|b| b.as_ptr()).collect();
    let ids: Vec<i32> = (0..NUM_TABLES as i32).collect();
    // generation/mod.rs:183-188 stored the Bitwise / Program betas in the starks
    let mut cc = [0u64; NUM_TABLES];
    cc[2] = ola_stark.bitwise_stark.get_compress_challenge().map(|x| x.0).unwrap_or(0);
    cc[10] = ola_stark.program_stark.get_compress_challenge().map(|x| x.0).unwrap_or(0);
    let mut out = vec![0u8; 64 << 20];
    let mut len = 0usize;
    let rc = unsafe {
        ola_prove(c, ids.as_ptr(), NUM_TABLES as u32, ptrs.as_ptr(), 0, log_ns.as_ptr(), cc.as_ptr(), 1, out.as_mut_ptr(), out.len(), &mut len)
    };
    match rc {
        OLA_OK => {}
        OLA_ERR_QUOTIENT_DEGREE => return Err(anyhow!("Quotient has failed, the vanishing polynomial is not divisible by Z_H ({})", last_error(c))),
        OLA_ERR_ZETA_IN_SUBGROUP => return Err(anyhow!("Opening point is in the subgroup.")),
        _ => return Err(anyhow!("ola_prove failed with {rc}: {}", last_error(c))),
    }
    out.truncate(len);
    let mut buf = Buffer::new(out);
    buf.read_all_proof::<F, C, D>().map_err(|e| anyhow!("malformed proof bytes from the GPU: {e:?}"))
}

/// The same with the transcript kept HERE: drives ola_prove_session_* with the host's own `Challenger`.
pub fn prove_with_traces_gpu_own_challenger(
    ola_stark: &OlaStark<F, D>,
    trace_poly_values: &[Vec<PolynomialValues<F>>; NUM_TABLES],
    challenger: &mut plonky2::iop::challenger::Challenger<F, <C as GenericConfig<D>>::Hasher>,
) -> Result<AllProof<F, C, D>> {
    use plonky2::field::types::Field;
    let guard = ctx()?.lock().unwrap();
    let c = guard.0;
    let blocks: Vec<Vec<u64>> = trace_poly_values.iter().map(|cols| cols.iter().flat_map(|p| p.values.iter().map(|x| x.0)).collect()).collect();
    let log_ns: Vec<u32> = trace_poly_values.iter().map(|cols| cols[0].len().trailing_zeros()).collect();
    let ptrs: Vec<*const u64> = blocks.iter().map(|b| b.as_ptr()).collect();
    let ids: Vec<i32> = (0..NUM_TABLES as i32).collect();
    let mut cc = [0u64; NUM_TABLES];
    cc[2] = ola_stark.bitwise_stark.get_compress_challenge().map(|x| x.0).unwrap_or(0);
    cc[10] = ola_stark.program_stark.get_compress_challenge().map(|x| x.0).unwrap_or(0);
    let mut s: *mut ola_session = ptr::null_mut();
    let rc = unsafe { ola_prove_session_begin(c, ids.as_ptr(), NUM_TABLES as u32, ptrs.as_ptr(), 0, log_ns.as_ptr(), cc.as_ptr(), 1, &mut s) };
    ensure!(rc == OLA_OK, "ola_prove_session_begin failed with {rc}: {}", last_error(c));
    let mut ev = ola_transcript_event { kind: 0, stage: 0, table: -1, elems: ptr::null(), count: 0 };
    loop {
        let rc = unsafe { ola_prove_session_next(s, &mut ev) };
        if rc != OLA_OK || ev.kind == OLA_EV_DONE || ev.kind == OLA_EV_FAILED {
            break;
        }
        match ev.kind {
            OLA_EV_OBSERVE => {
                let elems = unsafe { std::slice::from_raw_parts(ev.elems, ev.count) };
                for &x in elems {
                    challenger.observe_element(F::from_canonical_u64(x));
                }
            }
            OLA_EV_CHALLENGE => {
                let v: Vec<u64> = challenger.get_n_challenges(ev.count).iter().map(|x| x.0).collect();
                let rc = unsafe { ola_prove_session_supply(s, v.as_ptr(), v.len()) };
                ensure!(rc == OLA_OK, "ola_prove_session_supply failed with {rc}");
            }
            OLA_EV_COMPACT => challenger.compact(),
            _ => unreachable!(),
        }
    }
    let mut out = vec![0u8; 64 << 20];
    let mut len = 0usize;
    let rc = unsafe { ola_prove_session_finish(s, out.as_mut_ptr(), out.len(), &mut len) };
    ensure!(rc == OLA_OK, "proof failed with {rc}: {}", last_error(c));
    out.truncate(len);
    Buffer::new(out).read_all_proof::<F, C, D>().map_err(|e| anyhow!("malformed proof bytes from the GPU: {e:?}"))
}
