//! `prove()` next to `zkir_runtime::run()` (zkir-runtime/src/lib.rs:59-62): run the interpreter with trace recording,
//! pack `Vec<TraceRow>` (zkir-spec/src/trace.rs:24-50) into column-major BabyBear columns in pinned memory, and hand the
//! buffer to the CUDA prover through the C ABI.  Errors map to `RuntimeError::Other` (zkir-runtime/src/error.rs:35-36).
//!
//! NOT COMPILED in this repository's build image (no cargo/rustc).  The column packing below must stay in lock-step
//! with zkir_b200/csrc/host/pack.cc, which is the tested implementation of the same "converter".
pub mod ffi;

use std::ffi::CStr;
use zkir_runtime::{RuntimeError, VMConfig, VM};
use zkir_spec::Program;

#[derive(Clone, Copy, Debug)]
pub struct ProverConfig {
    pub log_blowup: u32,
    pub num_queries: u32,
    pub pow_bits: u32,
    pub max_cycles: u64,
    pub device: i32,
}

impl Default for ProverConfig {
    fn default() -> Self {
        Self { log_blowup: 1, num_queries: 100, pow_bits: 16, max_cycles: (1 << 24) + 16, device: 0 }
    }
}

pub struct Proof {
    pub bytes: Vec<u8>,
    pub public_values: [u32; 4],
    pub log_n: u32,
    pub cycles: u64,
    pub outputs: Vec<u64>,
}

/// Pinned host buffer owned by Rust, allocated by the library (cudaMallocHost) so the H2D copy is a single DMA.
struct Pinned {
    ptr: *mut u32,
    words: usize,
}
impl Pinned {
    fn new(words: usize) -> Result<Self, RuntimeError> {
        let p = unsafe { ffi::zkir_b200_alloc_pinned(words * 4) } as *mut u32;
        if p.is_null() {
            return Err(RuntimeError::Other("zkir_b200_alloc_pinned failed".into()));
        }
        Ok(Self { ptr: p, words })
    }
    fn as_mut_slice(&mut self) -> &mut [u32] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.words) }
    }
}
impl Drop for Pinned {
    fn drop(&mut self) {
        unsafe { ffi::zkir_b200_free_pinned(self.ptr as *mut _) }
    }
}

pub struct Prover {
    ctx: *mut ffi::zkir_ctx,
}
unsafe impl Send for Prover {} // a zkir_ctx is single-owner; moving it between threads is fine

impl Prover {
    pub fn new(device: i32) -> Result<Self, RuntimeError> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::zkir_b200_create(&mut ctx, device) };
        if rc != ffi::ZKIR_OK {
            return Err(last_error(std::ptr::null(), rc));
        }
        Ok(Self { ctx })
    }

    /// `cols`: `[112][1 << log_n]` column-major canonical values (see `pack_trace`).
    pub fn prove_columns(&mut self, cfg: &ProverConfig, cols: &[u32], log_n: u32, pv: &[u32; 4]) -> Result<Vec<u8>, RuntimeError> {
        assert_eq!(cols.len(), (ffi::ZKIR_AIR_V1_WIDTH as usize) << log_n);
        let params = ffi::zkir_params {
            log_blowup: cfg.log_blowup,
            num_queries: cfg.num_queries,
            pow_bits: cfg.pow_bits,
            width: ffi::ZKIR_AIR_V1_WIDTH,
            num_public: ffi::ZKIR_AIR_V1_NUM_PUBLIC,
        };
        let (mut p, mut len) = (std::ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe { ffi::zkir_b200_prove(self.ctx, &params, cols.as_ptr(), log_n, pv.as_ptr(), &mut p, &mut len) };
        if rc != ffi::ZKIR_OK {
            return Err(last_error(self.ctx, rc));
        }
        let out = unsafe { std::slice::from_raw_parts(p, len) }.to_vec();
        unsafe { ffi::zkir_b200_free_proof(p) };
        Ok(out)
    }
}
impl Drop for Prover {
    fn drop(&mut self) {
        unsafe { ffi::zkir_b200_destroy(self.ctx) }
    }
}

fn last_error(ctx: *const ffi::zkir_ctx, rc: i32) -> RuntimeError {
    let msg = unsafe { CStr::from_ptr(ffi::zkir_b200_last_error(ctx)) }.to_string_lossy().into_owned();
    RuntimeError::Other(format!("zkir_b200 error {rc}: {msg}"))
}

/// Program -> Proof.  The interpreter loop is untouched (north star: "zkir-spec, zkir-assembler and the interpreter
/// loop stay as-is"); only the consumer of `ExecutionResult.execution_trace` is new.
pub fn prove(program: &Program, inputs: &[u64], cfg: &ProverConfig) -> Result<Proof, RuntimeError> {
    let vm_cfg = VMConfig { max_cycles: cfg.max_cycles, enable_execution_trace: true, ..VMConfig::default() };
    let result = VM::new(program.clone(), inputs.to_vec(), vm_cfg).run()?;
    let rows = &result.execution_trace;
    let log_n = (rows.len().max(4)).next_power_of_two().trailing_zeros();
    let mut pinned = Pinned::new((ffi::ZKIR_AIR_V1_WIDTH as usize) << log_n)?;
    let pv = pack_trace(rows, program.header.entry_point, &result, log_n, pinned.as_mut_slice())?;
    let mut prover = Prover::new(cfg.device)?;
    let bytes = prover.prove_columns(cfg, pinned.as_mut_slice(), log_n, &pv)?;
    Ok(Proof { bytes, public_values: pv, log_n, cycles: result.cycles, outputs: result.outputs })
}

pub fn verify(proof: &[u8], cfg: &ProverConfig, pv: &[u32; 4]) -> bool {
    let params = ffi::zkir_params {
        log_blowup: cfg.log_blowup,
        num_queries: cfg.num_queries,
        pow_bits: cfg.pow_bits,
        width: ffi::ZKIR_AIR_V1_WIDTH,
        num_public: ffi::ZKIR_AIR_V1_NUM_PUBLIC,
    };
    unsafe { ffi::zkir_b200_verify(&params, proof.as_ptr(), proof.len(), pv.as_ptr()) == ffi::ZKIR_OK }
}

/// The "converter" the reference names but does not contain (zkir-spec/src/trace.rs:41, zkir-runtime/src/vm.rs:243-244):
/// one `TraceRow` -> one row of the 112 columns listed in zkir_b200/csrc/air_columns.h.  Mirrors pack.cc:
/// 40-bit values split into 2 x 20-bit limbs (zkir-spec/src/value.rs:592-601), decoded operand indices as one-hots,
/// ALU result / carries / branch helpers / I/O columns derived from the row and its successor, padding rows after halt.
fn pack_trace(
    rows: &[zkir_spec::TraceRow],
    entry_point: u32,
    result: &zkir_runtime::ExecutionResult,
    log_n: u32,
    cols: &mut [u32],
) -> Result<[u32; 4], RuntimeError> {
    let _ = (rows, entry_point, result, log_n, cols);
    // Intentionally a thin shim: link zkir_pack_trace()'s logic here, or call it through a second extern block fed with
    // a flat copy of the rows.  Kept unimplemented in this uncompiled source so it cannot silently drift from pack.cc.
    Err(RuntimeError::Other("pack_trace: port zkir_b200/csrc/host/pack.cc (see INTEGRATION.md, section 3)".into()))
}
