//! `prove()` next to `zkir_runtime::run()` (zkir-runtime/src/lib.rs:59-62): run the interpreter with trace recording,
//! pack `Vec<TraceRow>` (zkir-spec/src/trace.rs:24-50) into column-major BabyBear columns in pinned memory, and hand the
//! buffer to the CUDA prover through the C ABI.  Errors map to `RuntimeError::Other` (zkir-runtime/src/error.rs:35-36).
//!
//! NOT COMPILED in this repository's build image (no cargo/rustc).  It relies on one small addition to zkir-runtime:
//! `VM::run_keep_state`, i.e. `VM::run` (vm.rs:208-358) returning also the final registers and pc in `ExecutionResult`
//! (`final_regs: [u64; 16]`, `final_pc: u64`): padding rows and the last READ need the post-state of the last cycle.
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
struct Pinned<T: Copy> {
    ptr: *mut T,
    len: usize,
}
impl<T: Copy> Pinned<T> {
    fn new(len: usize) -> Result<Self, RuntimeError> {
        let p = unsafe { ffi::zkir_b200_alloc_pinned(len.max(1) * std::mem::size_of::<T>()) } as *mut T;
        if p.is_null() {
            return Err(RuntimeError::Other("zkir_b200_alloc_pinned failed".into()));
        }
        Ok(Self { ptr: p, len })
    }
    fn as_slice(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
    fn as_mut_slice(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl<T: Copy> Drop for Pinned<T> {
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

    /// `cols`: `[85][1 << log_n]` column-major canonical values (see `pack_trace`).
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
impl Prover {
    #[allow(clippy::too_many_arguments)]
    pub fn prove_rows(&mut self, cfg: &ProverConfig, pcs: &[u64], ins: &[u32], regs: &[u64], final_regs: &[u64; 16], final_pc: u64,
                      entry_point: u32, exit_code: u64, log_n: u32, pv_out: &mut [u32; 4]) -> Result<Vec<u8>, RuntimeError> {
        assert!(pcs.len() == ins.len() && regs.len() == 16 * pcs.len());
        let params = ffi::zkir_params {
            log_blowup: cfg.log_blowup,
            num_queries: cfg.num_queries,
            pow_bits: cfg.pow_bits,
            width: ffi::ZKIR_AIR_V1_WIDTH,
            num_public: ffi::ZKIR_AIR_V1_NUM_PUBLIC,
        };
        let (mut p, mut len) = (std::ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::zkir_b200_prove_rows(self.ctx, &params, pcs.as_ptr(), ins.as_ptr(), regs.as_ptr(), pcs.len() as u64, final_regs.as_ptr(),
                                      final_pc, entry_point, exit_code, log_n, pv_out.as_mut_ptr(), &mut p, &mut len)
        };
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
/// loop stay as-is"); only the consumer of `ExecutionResult.execution_trace` is new.  The rows are copied field by field
/// into pinned arrays and the device runs the converter (zkir_b200/csrc/trace_expand.cu), so no Rust port of it is needed.
pub fn prove(program: &Program, inputs: &[u64], cfg: &ProverConfig) -> Result<Proof, RuntimeError> {
    let vm_cfg = VMConfig { max_cycles: cfg.max_cycles, enable_execution_trace: true, ..VMConfig::default() };
    let mut vm = VM::new(program.clone(), inputs.to_vec(), vm_cfg);
    let result = vm.run_keep_state()?; // like VM::run (vm.rs:208-358) but also returns the final VMState (regs, pc)
    let rows = &result.execution_trace;
    let n = rows.len();
    let log_n = n.max(4).next_power_of_two().trailing_zeros();
    // SoA copy of TraceRow { pc, instruction, registers } into page-locked memory
    let mut pcs = Pinned::<u64>::new(n)?;
    let mut ins = Pinned::<u32>::new(n)?;
    let mut regs = Pinned::<u64>::new(16 * n)?;
    for (i, row) in rows.iter().enumerate() {
        pcs.as_mut_slice()[i] = row.pc;
        ins.as_mut_slice()[i] = row.instruction;
        regs.as_mut_slice()[16 * i..16 * i + 16].copy_from_slice(&row.registers);
    }
    let exit_code = match result.halt_reason { zkir_runtime::HaltReason::Exit(c) => c, _ => 0 };
    let mut pv = [0u32; 4];
    let mut prover = Prover::new(cfg.device)?;
    let bytes = prover.prove_rows(cfg, pcs.as_slice(), ins.as_slice(), regs.as_slice(), &result.final_regs, result.final_pc,
                                  program.header.entry_point, exit_code, log_n, &mut pv)?;
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
