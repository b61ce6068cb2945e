//! `prove()` / `verify()` next to `zkir_runtime::run()` (zkir-runtime/src/lib.rs:59-62), backed by libzkir_b200.so.
//!
//! SOURCE ONLY: the build image has no Rust toolchain, so this file has never been compiled; it is the
//! binding a zkir-runtime maintainer would add.  The interpreter stays the upstream Rust one: it
//! appends a register write log (pc, instruction word, (reg << 56) | value per cycle) to page-locked arrays through a two-line
//! hook in `VMState::write_reg` (zkir-runtime/src/state.rs:87-91), and this crate hands the arrays and the program to the C ABI
//! of include/zkir_b200.h.  Nothing here computes a proof on the CPU.
pub mod ffi;

use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;

use zkir_runtime::{ExecutionResult, HaltReason, RuntimeError, VMConfig, VM};
use zkir_spec::{Program, TraceRow};

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
        self.params_for(ffi::ZKIR_AIR_V2_WIDTH)
    }

    /// `width` selects the AIR profile of the proof: ZKIR_AIR_V2_WIDTH (core, 18 opcodes) or ZKIR_AIR_FULL_WIDTH (all 50)
    fn params_for(&self, width: u32) -> ffi::zkir_params {
        ffi::zkir_params {
            log_blowup: self.log_blowup,
            num_queries: self.num_queries,
            pow_bits: self.pow_bits,
            width,
            num_public: ffi::ZKIR_AIR_V2_NUM_PUBLIC,
        }
    }
}

/// Proof bytes (docs/PROVER_SPEC.md section 5) with the public values they bind: `[entry_pc, num_cycles, exit_lo, exit_hi, halted]`.
/// The program is part of the statement but not of the bytes: `verify` takes it.
#[derive(Clone, Debug)]
pub struct Proof {
    pub bytes: Vec<u8>,
    pub public_values: [u32; 5],
    pub log_n: u32,
    pub cycles: u64,
    pub outputs: Vec<u64>,
    /// public I/O transcript: (cycle, kind 0 READ / 1 WRITE, value lo20, value hi20) per READ / WRITE ecall; part of the statement
    pub io_events: Vec<[u32; 4]>,
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

    /// The program whose executions this context proves (`zkir_b200_set_program`): every executed (pc, instruction) is bound to it.
    pub fn set_program(&mut self, code: &[u32]) -> Result<(), RuntimeError> {
        let rc = unsafe { ffi::zkir_b200_set_program(self.ctx, code.as_ptr(), code.len()) };
        if rc != ffi::ZKIR_OK { Err(error_of(self.ctx, rc)) } else { Ok(()) }
    }

    /// The public I/O transcript of the execution about to be proven (`zkir_b200_set_io`).
    pub fn set_io(&mut self, events: &[[u32; 4]]) -> Result<(), RuntimeError> {
        let rc = unsafe { ffi::zkir_b200_set_io(self.ctx, events.as_ptr() as *const u32, events.len()) };
        if rc != ffi::ZKIR_OK { Err(error_of(self.ctx, rc)) } else { Ok(()) }
    }

    /// Register write log in (16 B per cycle, pinned), proof bytes out: the device rebuilds the pre-state registers of every row
    /// with a last-writer scan and runs the converter (zkir-spec/src/trace.rs:41, absent upstream).
    #[allow(clippy::too_many_arguments)]
    pub fn prove_writelog(
        &mut self,
        cfg: &ProverConfig,
        log: &WriteLog,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int,
        log_n: u32,
    ) -> Result<(Vec<u8>, [u32; 5]), RuntimeError> {
        let params = cfg.params();
        let mut pv = [0u32; 5];
        let (mut proof, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::zkir_b200_prove_writelog(
                self.ctx, &params, log.pcs.as_ptr(), log.instrs.as_ptr(), log.wlog.as_ptr(), log.len as u64, log.final_pc,
                entry_point, exit_code, halt_kind, log_n, pv.as_mut_ptr(), &mut proof, &mut len,
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

/// Memory log of a run, beside the register write log: what the full profile's offline memory checking needs from the interpreter.
/// A recorder hooked into `Memory::record_op` (zkir-runtime/src/memory.rs:243-253) fills it with one HashMap<u64, u32> of last
/// timestamps; rows without a load / store stay 0.
pub struct MemoryLog {
    pub old: Vec<u64>,   // [cycles] aligned 8-byte word before the access
    pub pts: Vec<u32>,   // [cycles] timestamp (cycle + 1) of the word's previous access, 0 = never
    pub widx: Vec<u64>,  // touched words (address / 8), strictly ascending
    pub word: Vec<u64>,  // their final contents
    pub ts: Vec<u32>,    // their last timestamps
}

impl Prover {
    /// Full-profile programs from the write log + memory log (28 B per cycle over PCIe, no host replay): the device rebuilds the
    /// registers, expands the 248-column table and proves it.
    #[allow(clippy::too_many_arguments)]
    pub fn prove_writelog_mem(
        &mut self,
        cfg: &ProverConfig,
        log: &WriteLog,
        mem: &MemoryLog,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int,
        log_n: u32,
    ) -> Result<(Vec<u8>, [u32; 5]), RuntimeError> {
        let params = cfg.params_for(ffi::ZKIR_AIR_FULL_WIDTH);
        let mut pv = [0u32; 5];
        let (mut proof, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::zkir_b200_prove_writelog_mem(
                self.ctx, &params, log.pcs.as_ptr(), log.instrs.as_ptr(), log.wlog.as_ptr(), mem.old.as_ptr(), mem.pts.as_ptr(),
                log.len as u64, mem.widx.as_ptr(), mem.word.as_ptr(), mem.ts.as_ptr(), mem.widx.len(), log.final_pc, entry_point,
                exit_code, halt_kind, log_n, pv.as_mut_ptr(), &mut proof, &mut len,
            )
        };
        if rc != ffi::ZKIR_OK {
            return Err(error_of(self.ctx, rc));
        }
        let bytes = unsafe { std::slice::from_raw_parts(proof, len) }.to_vec();
        unsafe { ffi::zkir_b200_free_proof(proof) };
        Ok((bytes, pv))
    }

    /// Many independent small proofs of ONE program (BASELINE config 4: add.zkasm x 4096): `traces[i]` = packed columns
    /// `[width][1 << log_ns[i]]` (zkir_pack_trace), `public_values[i]` its five public values, `ios[i]` its I/O transcript.
    /// The library overlaps the launch-latency-bound proofs on up to eight worker contexts of the same GPU.
    pub fn prove_batch(
        &mut self,
        cfg: &ProverConfig,
        traces: &[&[u32]],
        log_ns: &[u32],
        public_values: &[[u32; 5]],
        ios: &[Vec<[u32; 4]>],
    ) -> Result<Vec<Vec<u8>>, RuntimeError> {
        let n = traces.len();
        assert!(log_ns.len() == n && public_values.len() == n && ios.len() == n);
        let params = cfg.params();
        let t_ptrs: Vec<*const u32> = traces.iter().map(|t| t.as_ptr()).collect();
        let pv_ptrs: Vec<*const u32> = public_values.iter().map(|p| p.as_ptr()).collect();
        let io_ptrs: Vec<*const u32> = ios.iter().map(|e| e.as_ptr() as *const u32).collect();
        let io_lens: Vec<usize> = ios.iter().map(|e| e.len()).collect();
        let mut proofs = vec![ptr::null_mut::<u8>(); n];
        let mut lens = vec![0usize; n];
        let rc = unsafe {
            ffi::zkir_b200_prove_batch(
                self.ctx, &params, t_ptrs.as_ptr(), log_ns.as_ptr(), pv_ptrs.as_ptr(), io_ptrs.as_ptr(), io_lens.as_ptr(), n as u32,
                proofs.as_mut_ptr(), lens.as_mut_ptr(),
            )
        };
        let out = proofs
            .iter()
            .zip(&lens)
            .map(|(&p, &l)| {
                if p.is_null() {
                    return Vec::new();
                }
                let v = unsafe { std::slice::from_raw_parts(p, l) }.to_vec();
                unsafe { ffi::zkir_b200_free_proof(p) };
                v
            })
            .collect();
        if rc != ffi::ZKIR_OK {
            return Err(error_of(self.ctx, rc));
        }
        Ok(out)
    }

    /// Full-profile programs (MUL / DIV, bitwise, shifts, signed compares, loads / stores): the rows `VM::run` records with
    /// `enable_execution_trace` (zkir-spec/src/trace.rs:24-50) go in as three flat arrays; the library builds the 248-column table
    /// (zkir_pack_rows_full: multiplier block, lookup multiplicities, offline memory checking) and proves it on the GPU.
    #[allow(clippy::too_many_arguments)]
    pub fn prove_rows_full(
        &mut self,
        cfg: &ProverConfig,
        rows: &[TraceRow],
        final_regs: &[u64; 16],
        final_pc: u64,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int,
        log_n: u32,
    ) -> Result<(Vec<u8>, [u32; 5]), RuntimeError> {
        let params = cfg.params_for(ffi::ZKIR_AIR_FULL_WIDTH);
        let pcs: Vec<u64> = rows.iter().map(|r| r.pc).collect();
        let instrs: Vec<u32> = rows.iter().map(|r| r.instruction).collect();
        let regs: Vec<u64> = rows.iter().flat_map(|r| r.registers.iter().copied()).collect();
        let mut pv = [0u32; 5];
        let (mut proof, mut len) = (ptr::null_mut::<u8>(), 0usize);
        let rc = unsafe {
            ffi::zkir_b200_prove_rows(
                self.ctx, &params, pcs.as_ptr(), instrs.as_ptr(), regs.as_ptr(), rows.len() as u64, final_regs.as_ptr(), final_pc,
                entry_point, exit_code, halt_kind, log_n, pv.as_mut_ptr(), &mut proof, &mut len,
            )
        };
        if rc != ffi::ZKIR_OK {
            return Err(error_of(self.ctx, rc));
        }
        let bytes = unsafe { std::slice::from_raw_parts(proof, len) }.to_vec();
        unsafe { ffi::zkir_b200_free_proof(proof) };
        Ok((bytes, pv))
    }
}

/// The recorder the interpreter writes into while it executes: three page-locked arrays, one entry per cycle.
/// Upstream hook: `VMState::write_reg` stores `cur = (reg << 56) | value`; the cycle loop of `VM::run` (vm.rs:234-312) calls
/// `begin(pc, word)` before and `commit()` after executing an instruction, and sets `final_pc` where the VM halts.
pub struct WriteLog {
    pub pcs: Pinned<u32>,
    pub instrs: Pinned<u32>,
    pub wlog: Pinned<u64>,
    pub len: usize,
    pub cur: u64,
    pub final_pc: u64,
    /// appended by the READ / WRITE arms of `handle_syscall` (zkir-runtime/src/syscall.rs:104-119): (cycle, kind, value lo20, value hi20)
    pub io_events: Vec<[u32; 4]>,
}

impl WriteLog {
    pub fn pinned(capacity: usize) -> Result<Self, RuntimeError> {
        Ok(Self { pcs: Pinned::new(capacity)?, instrs: Pinned::new(capacity)?, wlog: Pinned::new(capacity)?, len: 0, cur: 0, final_pc: 0, io_events: Vec::new() })
    }
    #[inline]
    pub fn begin(&mut self, pc: u32, word: u32) {
        let k = self.len;
        self.pcs.as_mut_slice()[k] = pc;
        self.instrs.as_mut_slice()[k] = word;
        self.cur = 0;
    }
    #[inline]
    pub fn commit(&mut self) {
        let k = self.len;
        self.wlog.as_mut_slice()[k] = self.cur;
        self.len = k + 1;
    }
}

impl Drop for Prover {
    fn drop(&mut self) {
        unsafe { ffi::zkir_b200_destroy(self.ctx) }
    }
}

/// The new public function, next to `run()`: interpret with the upstream VM (write-log hook on), prove on the GPU.
pub fn prove(program: &Program, inputs: &[u64], cfg: &ProverConfig) -> Result<Proof, RuntimeError> {
    let vm_cfg = VMConfig { max_cycles: cfg.max_cycles, ..VMConfig::default() }; // no execution trace needed: the log replaces it
    let entry_point = program.header.entry_point;
    let mut log = WriteLog::pinned(cfg.max_cycles as usize)?;
    let mut vm = VM::new(program.clone(), inputs.to_vec(), vm_cfg);
    vm.set_write_log(&mut log); // the upstream hook described on `WriteLog`
    let result = vm.run()?; // interpreter otherwise untouched
    let (exit_code, halt_kind) = match result.halt_reason {
        HaltReason::Exit(code) => (code, ffi::ZKIR_HALT_EXIT),
        HaltReason::Ebreak => (0, ffi::ZKIR_HALT_EBREAK),
        _ => (0, ffi::ZKIR_HALT_CYCLE_LIMIT),
    };
    if unsafe { ffi::zkir_program_profile(program.code.as_ptr(), program.code.len()) } == 0 {
        return prove_full(program, inputs, cfg); // the program leaves the 18 core opcodes: full AIR profile
    }
    let rows = (log.len + 1).max(program.code.len()).max(1 << ffi::ZKIR_MIN_LOG_N); // range table, ROM and one padding row fit the trace
    let log_n = rows.next_power_of_two().trailing_zeros();
    let mut prover = Prover::new(cfg.device)?;
    prover.set_program(&program.code)?;
    prover.set_io(&log.io_events)?;
    let (bytes, public_values) = prover.prove_writelog(cfg, &log, entry_point, exit_code, halt_kind, log_n)?;
    Ok(Proof { bytes, public_values, log_n, cycles: result.cycles, outputs: result.outputs.clone(), io_events: log.io_events.clone() })
}

/// Full-profile flavour of `prove`: the interpreter records `Vec<TraceRow>` (`enable_execution_trace`), the library packs and proves.
/// Needs the machine state after the last instruction, which `ExecutionResult` does not carry upstream (vm.rs:54-78): two fields
/// (`final_regs`, `final_pc`) filled where `VM::run` builds the result (vm.rs:349-357).
fn prove_full(program: &Program, inputs: &[u64], cfg: &ProverConfig) -> Result<Proof, RuntimeError> {
    let vm_cfg = VMConfig { max_cycles: cfg.max_cycles, enable_execution_trace: true, ..VMConfig::default() };
    let result = VM::new(program.clone(), inputs.to_vec(), vm_cfg).run()?;
    let (exit_code, halt_kind) = match result.halt_reason {
        HaltReason::Exit(code) => (code, ffi::ZKIR_HALT_EXIT),
        HaltReason::Ebreak => (0, ffi::ZKIR_HALT_EBREAK),
        _ => (0, ffi::ZKIR_HALT_CYCLE_LIMIT),
    };
    let n = result.execution_trace.len();
    let log_n = (n + 1).max(program.code.len()).max(1 << ffi::ZKIR_MIN_LOG_N).next_power_of_two().trailing_zeros();
    let io_events = io_events_of(&result); // (cycle, kind, lo20, hi20) per READ / WRITE ecall, from the rows (r10 selects the syscall)
    let mut prover = Prover::new(cfg.device)?;
    prover.set_program(&program.code)?;
    prover.set_io(&io_events)?;
    let (bytes, public_values) = prover.prove_rows_full(
        cfg, &result.execution_trace, &result.final_regs, result.final_pc, program.header.entry_point, exit_code, halt_kind, log_n,
    )?;
    Ok(Proof { bytes, public_values, log_n, cycles: result.cycles, outputs: result.outputs.clone(), io_events })
}

/// The public I/O transcript of a recorded run: an ECALL row with r10 = 1 (READ: the value is the next row's r10) or 2 (WRITE: r11).
fn io_events_of(result: &ExecutionResult) -> Vec<[u32; 4]> {
    let rows = &result.execution_trace;
    let mut out = Vec::new();
    for (i, r) in rows.iter().enumerate() {
        if r.instruction & 0x7F != 0x50 {
            continue;
        }
        let v = match r.registers[10] {
            1 => rows.get(i + 1).map(|n| n.registers[10]).unwrap_or(result.final_regs[10]),
            2 => r.registers[11],
            _ => continue,
        };
        out.push([i as u32, (r.registers[10] == 2) as u32, (v & 0xF_FFFF) as u32, ((v >> 20) & 0xF_FFFF) as u32]);
    }
    out
}

/// CPU verifier (no GPU needed): accepts or rejects `proof` as a statement about `program` for these parameters and public values.
/// The AIR profile is read from the proof header (word 3 = trace width).
pub fn verify(proof: &Proof, program: &Program, cfg: &ProverConfig) -> bool {
    let width = proof.bytes.get(12..16).map(|b| u32::from_le_bytes([b[0], b[1], b[2], b[3]])).unwrap_or(ffi::ZKIR_AIR_V2_WIDTH);
    let params = cfg.params_for(width);
    let rc = unsafe {
        ffi::zkir_b200_verify(
            &params, proof.bytes.as_ptr(), proof.bytes.len(), proof.public_values.as_ptr(), program.code.as_ptr(), program.code.len(),
            proof.io_events.as_ptr() as *const u32, proof.io_events.len(),
        )
    };
    rc == ffi::ZKIR_OK
}
