//! Raw `extern "C"` view of include/zkir_b200.h (hot-path subset).  One line per C declaration.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct zkir_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct zkir_params {
    pub log_blowup: u32,
    pub num_queries: u32,
    pub pow_bits: u32,
    pub width: u32,
    pub num_public: u32,
}

pub const ZKIR_OK: c_int = 0;
pub const ZKIR_AIR_V2_WIDTH: u32 = 88;
/// trace width of the full AIR profile (all 50 opcodes: docs/PROVER_SPEC.md sections 3.6-3.8); include/zkir_b200.h ZKIR_AIR_FULL_WIDTH
pub const ZKIR_AIR_FULL_WIDTH: u32 = 248;
pub const ZKIR_AIR_V2_NUM_PUBLIC: u32 = 5; // entry_pc, num_cycles, exit_lo, exit_hi, halted
pub const ZKIR_MIN_LOG_N: u32 = 10;
pub const ZKIR_HALT_EXIT: c_int = 0;
pub const ZKIR_HALT_EBREAK: c_int = 1;
pub const ZKIR_HALT_CYCLE_LIMIT: c_int = 2;

extern "C" {
    pub fn zkir_b200_create(out: *mut *mut zkir_ctx, device_id: c_int) -> c_int;
    pub fn zkir_b200_destroy(ctx: *mut zkir_ctx);
    pub fn zkir_b200_last_error(ctx: *const zkir_ctx) -> *const c_char;
    pub fn zkir_b200_alloc_pinned(bytes: usize) -> *mut c_void;
    pub fn zkir_b200_free_pinned(p: *mut c_void);
    /// the program whose executions the context proves: `Program.code` (zkir-spec/src/program.rs:241-250); ROM of the lookup argument
    pub fn zkir_b200_set_program(ctx: *mut zkir_ctx, code: *const u32, n_code: usize) -> c_int;
    /// the public I/O transcript of the execution about to be proven: n events of {cycle, kind (0 READ, 1 WRITE), value lo20, value hi20}
    pub fn zkir_b200_set_io(ctx: *mut zkir_ctx, events: *const u32, n_events: usize) -> c_int;
    /// Program + inputs -> proof in one call (the library's own interpreter restatement records the write log into pinned memory)
    pub fn zkir_b200_prove_program(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        code: *const u32,
        n_code: usize,
        data: *const u8,
        n_data: usize,
        entry_point: u32,
        inputs: *const u64,
        n_inputs: usize,
        max_cycles: u64,
        public_values_out: *mut u32, // [5]
        out_cycles: *mut u64,
        out_log_n: *mut u32,
        proof: *mut *mut u8,
        proof_len: *mut usize,
    ) -> c_int;
    pub fn zkir_b200_prove(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        trace_cols: *const u32, // [width][1 << log_n], column-major, canonical BabyBear values, host memory
        log_n: u32,
        public_values: *const u32,
        proof: *mut *mut u8,
        proof_len: *mut usize,
    ) -> c_int;
    /// many independent small proofs of the context's program (BASELINE config 4): traces[i] = packed columns [width][1 << log_ns[i]],
    /// ios[i] / n_ios[i] = the public I/O transcript of execution i; up to 8 worker contexts overlap the launch-bound tiny proofs
    pub fn zkir_b200_prove_batch(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        traces: *const *const u32,
        log_ns: *const u32,
        public_values: *const *const u32,
        ios: *const *const u32,
        n_ios: *const usize,
        n_proofs: u32,
        proofs: *mut *mut u8,   // [n_proofs], each freed with zkir_b200_free_proof
        proof_lens: *mut usize, // [n_proofs]
    ) -> c_int;
    /// raw TraceRow data (zkir-spec/src/trace.rs:24-50): pcs[n], instrs[n], regs[n][16] pre-state; converter on the device
    pub fn zkir_b200_prove_rows(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        pcs: *const u64,
        instrs: *const u32,
        regs: *const u64,
        n_rows: u64,
        final_regs: *const u64, // [16] state after the last instruction
        final_pc: u64,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int, // ZKIR_HALT_*
        log_n: u32,
        public_values_out: *mut u32, // [5]
        proof: *mut *mut u8,
        proof_len: *mut usize,
    ) -> c_int;
    /// register write log: pcs[n] (u32), instrs[n], wlog[n] = (k << 56) | value when the row changed register k, else 0
    /// (logged where VMState::write_reg runs, zkir-runtime/src/state.rs:76-91); 16 B/row over PCIe, the fastest hand-off
    pub fn zkir_b200_prove_writelog(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        pcs: *const u32,
        instrs: *const u32,
        wlog: *const u64,
        n_rows: u64,
        final_pc: u64,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int, // ZKIR_HALT_*
        log_n: u32,
        public_values_out: *mut u32, // [5]
        proof: *mut *mut u8,
        proof_len: *mut usize,
    ) -> c_int;
    /// full profile from the write log + memory log (28 B/row): mem_old[n] / mem_pts[n] = on load / store rows the aligned 8-byte word
    /// before the access and the timestamp (cycle + 1, 0 = never) of its previous access, logged where Memory::record_op runs
    /// (zkir-runtime/src/memory.rs:243-253); mem_widx / mem_word / mem_ts [n_words] = the touched words (address / 8) in strictly
    /// ascending order with their final contents and last timestamps
    pub fn zkir_b200_prove_writelog_mem(
        ctx: *mut zkir_ctx,
        params: *const zkir_params,
        pcs: *const u32,
        instrs: *const u32,
        wlog: *const u64,
        mem_old: *const u64,
        mem_pts: *const u32,
        n_rows: u64,
        mem_widx: *const u64,
        mem_word: *const u64,
        mem_ts: *const u32,
        n_words: usize,
        final_pc: u64,
        entry_point: u32,
        exit_code: u64,
        halt_kind: c_int,
        log_n: u32,
        public_values_out: *mut u32, // [5]
        proof: *mut *mut u8,
        proof_len: *mut usize,
    ) -> c_int;
    /// one proof sharded over several GPUs (one context per GPU): rank 0 draws the id, the host hands it to the other ranks,
    /// every rank calls comm_init; afterwards the prove_* calls are collective and return the single-GPU proof bytes everywhere
    pub fn zkir_b200_comm_unique_id(id: *mut u8 /* [128] */) -> c_int;
    pub fn zkir_b200_comm_init(ctx: *mut zkir_ctx, id: *const u8 /* [128] */, rank: c_int, world: c_int) -> c_int;
    pub fn zkir_b200_comm_shutdown(ctx: *mut zkir_ctx) -> c_int;
    /// which AIR profile a program needs: 1 = core (88 columns), 0 = full (ZKIR_AIR_FULL_WIDTH), -1 = undefined opcode
    pub fn zkir_program_profile(code: *const u32, n_code: usize) -> c_int;
    pub fn zkir_b200_free_proof(p: *mut u8);
    pub fn zkir_b200_verify(
        params: *const zkir_params, proof: *const u8, len: usize, public_values: *const u32, code: *const u32, n_code: usize,
        io_events: *const u32, n_io: usize,
    ) -> c_int;
}
