//! `prove()` / `verify()` next to `zkir_runtime::run()` (zkir-runtime/src/lib.rs:59-62), backed by libzkir_b200.so.
//!
//! SOURCE ONLY: the build image has no Rust toolchain, so this file has never been compiled; it is the
//! binding a zkir-runtime maintainer would add.  The interpreter stays the upstream Rust one: the
//! only thing this crate does with an execution is copy `TraceRow.pc / instruction / registers` (zkir-spec/src/trace.rs:24-50)
//! into page-locked arrays and hand them to the C ABI of include/zkir_b200.h.  Nothing here computes a proof on the CPU.
pub mod ffi;

use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;

use zkir_runtime::{HaltReason, RuntimeError, VMConfig, VM};
use zkir_spec::Program;

/// Proving parameters (Plonky3's FriConfig fields; docs/PROVER_SPEC.md section 4).
#[derive(Clone, Copy, Debug)]
pub struct ProverConfig {
    pub log_blowup: u32,
    pub num_queries: u32,
    pub pow_bits: u32,
    /// `VMConfig.max_cycles` defaults to 1_000_000 (zkir-runtime/src/vm.rs:43): too small for a 2^20-cycle trace.
    pub max_cycles: u64,
    pub device: i32,
}

impl Default for ProverConfig {
    fn default() -> Self {
        Self { log_blowup: 1, num_queries: 100, pow_bits: 16, max_cycles: (1 << 24) + 16, device: 0 }
    }
}

impl ProverConfig {
    fn params(&self) -> ffi::zkir_params {
        ffi::zkir_params {
            log_blowup: self.log_blowup,
            num_queries: self.num_queries,
            pow_bits: self.pow_bits,
            width: ffi::ZKIR_AIR_V1_WIDTH,
            num_public: ffi::ZKIR_AIR_V1_NUM_PUBLIC,
        }
    }
}

/// Proof bytes (docs/PROVER_SPEC.md section 5) with the public values they bind: `[entry_pc, num_cycles, exit_lo, exit_hi]`.
#[derive(Clone, Debug)]
pub struct Proof {
    pub bytes: Vec<u8>,
    pub public_values: [u32; 4],
    pub log_n: u32,
    pub cycles: u64,
    pub outputs: Vec<u64>,
}

fn error_of(ctx: *const ffi::zkir_ctx, code: c_int) -> RuntimeError {
    // ZKIR_ERR_* -> RuntimeError::Other(String) (zkir-runtime/src/error.rs:35-36); the message comes from the library
    let msg = unsafe {
        let p = ffi::zkir_b200_last_error(ctx);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    RuntimeError::Other(format!("zkir_b200 error {code}: {msg}"))
}

/// Page-locked host array the recorder writes into (`zkir_b200_alloc_pinned`): the H2D copy is then one DMA.
pub struct Pinned<T: Copy> {
    ptr: *mut T,
    len: usize,
}

impl<T: Copy> Pinned<T> {
    pub fn new(len: usize) -> Result<Self, RuntimeError> {
        let bytes = len.max(1) * std::mem::size_of::<T>();
        let p = unsafe { ffi::zkir_b200_alloc_pinned(bytes) } as *mut T;
        if p.is_null() {
            return Err(RuntimeError::Other("zkir_b200_alloc_pinned failed (no CUDA device?)".to_string()));
        }
        Ok(Self { ptr: p, len })
    }
    pub fn as_mut_slice(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
    pub fn as_ptr(&self) -> *const T {
        self.ptr
    }
}

impl<T: Copy> Drop for Pinned<T> {
    fn drop(&mut self) {
        unsafe { ffi::zkir_b200_free_pinned(self.ptr as *mut _) }
    }
}

/// One `zkir_ctx` = one GPU.  Single-owner: `&mut self` on every proving call.
pub struct Prover {
    ctx: *mut ffi::zkir_ctx,
}

// the context may move between threads, but is used by one thread at a time (include/zkir_b200.h, threading note)
unsafe impl Send for Prover {}

impl Prover {
    pub fn new(device: i32) -> Result<Self, RuntimeError> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::zkir_b200_create(&mut ctx, device as c_int) };
        if rc != ffi::ZKIR_OK {
            return Err(error_of(ptr::null(), rc)); // no CPU fallback: without an sm_100 device this is the end
        }
        Ok(Self { ctx })
    }

    /// Link this context with the contexts of the other ranks for ONE proof over several GPUs (BASELINE config 5).
    /// `id` comes from `Prover::comm_unique_id()` on rank 0 and is handed to the other ranks by the host (pipe, MPI, file ...).
    /// Afterwards every `prove_*` call on the linked contexts is collective and returns the single-GPU proof bytes everywhere.
    pub fn comm_init(&mut self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), RuntimeError> {
        let rc = unsafe { ffi::zkir_b200_comm_init(self.ctx, id.as_ptr(), rank as c_int, world as c_int) };
        if rc != ffi::ZKIR_OK { Err(error_of(self.ctx, rc)) } else { Ok(()) }
    }

    pub fn comm_unique_id() -> Result<[u8; 128], RuntimeError> {
        let mut id = [0u8; 128];
        let rc = unsafe { ffi::zkir_b200_comm_unique_id(id.as_mut_ptr()) };
        if rc != ffi::ZKIR_OK { Err(error_of(ptr::null(), rc)) } else { Ok(id) }
    }

    /// Packed columns in (`[72][1 << log_n]`, column-major canonical values, host memory): `zkir_b200_prove`.
    pub fn prove_columns(&mut self, cfg: &ProverConfig, cols: &[u32], log_n: u32, pv: &[u32; 4]) -> Result<Vec<u8>, RuntimeError> {
        assert_eq!(cols.len(), (ffi::ZKIR_AIR_V1_WIDTH as usize) << log_n);
        let params = cfg.params();
        let (mut proof, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe { ffi::zkir_b200_prove(self.ctx, &params, cols.as_ptr(), log_n, pv.as_ptr(), &mut proof, &mut len) };
        if rc != ffi::ZKIR_OK {
            return Err(error_of(self.ctx, rc));
        }
        let bytes = unsafe { std::slice::from_raw_parts(proof, len) }.to_vec();
        unsafe { ffi::zkir_b200_free_proof(proof) };
        Ok(bytes)
    }

    /// Raw rows in, proof bytes out: the converter (zkir-spec/src/trace.rs:41, absent upstream) runs on the device.
    #[allow(clippy::too_many_arguments)]
    pub fn prove_rows(
        &mut self,
        cfg: &ProverConfig,
        pcs: &Pinned<u64>,
        instrs: &Pinned<u32>,
        regs: &Pinned<u64>, // [n_rows][16], PRE-state (zkir-runtime/src/vm.rs:245-253)
        n_rows: u64,
        final_regs: &[u64; 16],
        final_pc: u64,
        entry_point: u32,
        exit_code: u64,
        log_n: u32,
    ) -> Result<(Vec<u8>, [u32; 4]), RuntimeError> {
        let params = cfg.params();
        let mut pv = [0u32; 4];
        let (mut proof, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::zkir_b200_prove_rows(
                self.ctx, &params, pcs.as_ptr(), instrs.as_ptr(), regs.as_ptr(), n_rows, final_regs.as_ptr(), final_pc,
                entry_point, exit_code, log_n, pv.as_mut_ptr(), &mut proof, &mut len,
            )
        };
        if rc != ffi::ZKIR_OK {
            return Err(error_of(self.ctx, rc));
        }
        let bytes = unsafe { std::slice::from_raw_parts(proof, len) }.to_vec();
        unsafe { ffi::zkir_b200_free_proof(proof) }; // the library owns the buffer until here
        Ok((bytes, pv))
    }
}

impl Drop for Prover {
    fn drop(&mut self) {
        unsafe { ffi::zkir_b200_destroy(self.ctx) }
    }
}

/// The new public function, next to `run()`: interpret with the upstream VM (execution trace on), prove on the GPU.
///
/// Needs one upstream addition: `ExecutionResult` must carry the machine state after the last instruction
/// (`final_regs`, `final_pc`); padding rows and the value of a trailing READ are taken from it.
pub fn prove(program: &Program, inputs: &[u64], cfg: &ProverConfig) -> Result<Proof, RuntimeError> {
    let vm_cfg = VMConfig { max_cycles: cfg.max_cycles, enable_execution_trace: true, ..VMConfig::default() };
    let entry_point = program.header.entry_point;
    let result = VM::new(program.clone(), inputs.to_vec(), vm_cfg).run()?; // interpreter untouched
    let rows = &result.execution_trace; // Vec<TraceRow>, one per cycle, PRE-state registers
    let n = rows.len();
    let log_n = (n.max(4).next_power_of_two().trailing_zeros()).max(2);

    let (mut pcs, mut ins, mut regs) = (Pinned::<u64>::new(n)?, Pinned::<u32>::new(n)?, Pinned::<u64>::new(16 * n)?);
    {
        let (p, i, r) = (pcs.as_mut_slice(), ins.as_mut_slice(), regs.as_mut_slice());
        for (k, row) in rows.iter().enumerate() {
            p[k] = row.pc;
            i[k] = row.instruction;
            r[16 * k..16 * k + 16].copy_from_slice(&row.registers);
        }
    }
    let exit_code = match result.halt_reason {
        HaltReason::Exit(code) => code,
        _ => 0,
    };
    let mut prover = Prover::new(cfg.device)?;
    let (bytes, public_values) = prover.prove_rows(
        cfg, &pcs, &ins, &regs, n as u64, &result.final_regs, result.final_pc, entry_point, exit_code, log_n,
    )?;
    Ok(Proof { bytes, public_values, log_n, cycles: result.cycles, outputs: result.outputs.clone() })
}

/// CPU verifier (no GPU needed): accepts or rejects `proof` for these parameters and public values.
pub fn verify(proof: &Proof, cfg: &ProverConfig) -> bool {
    let params = cfg.params();
    let rc = unsafe { ffi::zkir_b200_verify(&params, proof.bytes.as_ptr(), proof.bytes.len(), proof.public_values.as_ptr()) };
    rc == ffi::ZKIR_OK
}
