/* zkir_b200.h -- C ABI of the B200-native STARK proving backend for the ZKIR v3.4 VM.
 *
 * The reference (seceq/zkir, /root/reference) has NO plugin / FFI seam and NO prover: its public runtime API
 * ends at `zkir_runtime::run()` / `VM::run` returning `ExecutionResult` (zkir-runtime/src/lib.rs:29-62,
 * zkir-runtime/src/vm.rs:54-78,208-358).  This header therefore DEFINES the seam a Rust `prove()` placed next to
 * `run()` would bind (see INTEGRATION.md for the `extern "C"` block and the safe wrapper).  Each entry point
 * below names the reference item it sits behind or replaces.
 *
 * Conventions: all integers little-endian; field elements are canonical BabyBear values (uint32_t < p,
 * p = 2^31 - 2^27 + 1) on this boundary; matrices are column-major `[width][1 << log_n]`.
 * Return 0 = OK, negative = error (ZKIR_ERR_*); nothing throws or exits across the ABI.  A `zkir_ctx` is
 * single-owner (one host thread at a time); independent contexts may live on different threads/devices.
 * There is NO CPU fallback: every proving entry point fails with ZKIR_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef ZKIR_B200_H
#define ZKIR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZKIR_OK 0
#define ZKIR_ERR_ARG (-1)   /* bad argument / unsupported shape        -> RuntimeError::Other (zkir-runtime/src/error.rs:35-36) */
#define ZKIR_ERR_CUDA (-2)  /* CUDA runtime failure or no device */
#define ZKIR_ERR_NCCL (-3)
#define ZKIR_ERR_OOM (-4)
#define ZKIR_ERR_VM (-5)    /* guest fault: DivisionByZero / MisalignedAccess / InvalidSyscall ... (error.rs:6-39) */
#define ZKIR_ERR_AIR (-6)   /* trace uses an opcode / value the AIR v2 does not constrain */
#define ZKIR_ERR_VERIFY (-7)

#define ZKIR_BABYBEAR_P 2013265921u
#define ZKIR_AIR_V2_WIDTH 88u      /* main trace columns of the CORE profile (docs/PROVER_SPEC.md section 3); 16 aux + 4 public columns are internal */
#define ZKIR_AIR_FULL_WIDTH 248u   /* main trace columns of the FULL profile (sections 3.7, 3.8: MUL/DIV family, bitwise, shifts, signed compares, loads / stores) */
#define ZKIR_AIR_V2_NUM_PUBLIC 5u  /* entry_pc, num_cycles, exit_lo, exit_hi, halted */
#define ZKIR_MIN_LOG_N 10u         /* the 1024-entry range table (zkir-spec/src/config.rs:76-80) occupies trace rows */

/* ---- proving parameters (the reference has none; Plonky3's FriConfig fields) */
typedef struct {
  uint32_t log_blowup;  /* >= 1 */
  uint32_t num_queries;
  uint32_t pow_bits;
  uint32_t width;       /* ZKIR_AIR_V2_WIDTH or ZKIR_AIR_FULL_WIDTH: selects the AIR profile of the proof (zkir_program_profile) */
  uint32_t num_public;  /* must equal ZKIR_AIR_V2_NUM_PUBLIC */
} zkir_params;

typedef struct zkir_ctx zkir_ctx;

/* ---- context: one per GPU.  (No reference counterpart.) */
int zkir_b200_create(zkir_ctx** out, int device_id);
void zkir_b200_destroy(zkir_ctx*);
const char* zkir_b200_last_error(const zkir_ctx*); /* ctx may be NULL: last create/VM error of this thread */
/* pinned host buffers the interpreter records into (north_star: "records the execution trace into pinned memory") */
void* zkir_b200_alloc_pinned(size_t bytes);
void zkir_b200_free_pinned(void*);

/* ---- the program whose executions this context proves: `Program.code` (zkir-spec/src/program.rs:241-250), loaded at 0x1000 by
 * VM::new (vm.rs:138-205).  The AIR binds every executed (pc, instruction) to this ROM through a lookup argument, and the
 * transcript absorbs its digest, so a proof is a statement about THIS program.  Must be called before the first prove call and
 * whenever the program changes (the decoded ROM columns and their LDE are cached per trace size).  Copies the words. */
int zkir_b200_set_program(zkir_ctx*, const uint32_t* code, size_t n_code);

/* ---- the public I/O transcript of the execution about to be proven: n_events x {cycle, kind (0 READ / 1 WRITE), value lo20, value hi20}
 * (zkir_vm_io).  The AIR sends every READ / WRITE row to a lookup bus whose other side is this list, and the transcript absorbs its
 * digest: a proof says "this program consumed these inputs and produced these outputs at these cycles".  Defaults to the empty
 * transcript; zkir_b200_prove_program sets it itself.  A list that does not match the trace gives ZKIR_ERR_AIR. */
int zkir_b200_set_io(zkir_ctx*, const uint32_t* events, size_t n_events);

/* ---- the hot path: trace columns -> proof.  Sits where `zkir_runtime::prove()` would call into Plonky3
 * (the call does not exist in the reference: zkir-runtime/src/lib.rs:29-62).  `trace_cols` is HOST memory
 * (pinned recommended), `[width][1 << log_n]`; `public_values[num_public]`.  The proof buffer is owned by the
 * library until zkir_b200_free_proof.  H2D of the trace and D2H of the proof happen inside this call. */
int zkir_b200_prove(zkir_ctx*, const zkir_params*, const uint32_t* trace_cols, uint32_t log_n,
                    const uint32_t* public_values, uint8_t** proof, size_t* proof_len);
/* same, trace already resident in device memory (canonical values); used by bench `value` timing */
int zkir_b200_prove_device(zkir_ctx*, const zkir_params*, const uint32_t* d_trace_cols, uint32_t log_n,
                           const uint32_t* public_values, uint8_t** proof, size_t* proof_len);
/* same, from the RAW interpreter rows instead of packed columns: what `ExecutionResult.execution_trace` holds
 * (zkir-spec/src/trace.rs:24-50: pc, instruction word, PRE-state registers[16]; cycle = row index), host memory (pinned
 * recommended).  The converter (trace.rs:41) runs on the device, so 140 B/row cross PCIe instead of 340 B/row.
 * final_regs / final_pc: machine state after the last instruction; exit_code: HaltReason::Exit code (0 otherwise).
 * halt_kind: ZKIR_HALT_* of the run.  Writes the 5 public values {entry_pc, num_cycles, exit_lo, exit_hi, halted} it proves to
 * public_values_out.  Rows the AIR v2 cannot constrain give ZKIR_ERR_AIR (same rules as zkir_pack_trace).
 * params.width = ZKIR_AIR_FULL_WIDTH (a program that needs the full profile: zkir_program_profile): the host replays the run's memory in
 * program order (12 B per row more over PCIe: the word and the previous timestamp each load / store sees), the device expands the
 * 248-column table; the proof is the same path. */
int zkir_b200_prove_rows(zkir_ctx*, const zkir_params*, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs,
                         uint64_t n_rows, const uint64_t* final_regs, uint64_t final_pc, uint32_t entry_point, uint64_t exit_code,
                         int halt_kind, uint32_t log_n, uint32_t* public_values_out, uint8_t** proof, size_t* proof_len);
/* same from the interpreter's REGISTER WRITE LOG, the most compact hand-off: per row pc (u32), instruction word and
 * wlog = (k << 56) | value if the row changed register k, else 0 (zkir_vm_trace_writelog builds it from recorded rows; a
 * Rust recorder logs it where VMState writes a register, zkir-runtime/src/state.rs:76-91).  The device rebuilds the
 * pre-state registers of every row with a last-writer scan (registers start at 0: vm.rs:177-181), then runs the same
 * converter.  16 B/row cross PCIe. */
int zkir_b200_prove_writelog(zkir_ctx*, const zkir_params*, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog,
                             uint64_t n_rows, uint64_t final_pc, uint32_t entry_point, uint64_t exit_code, int halt_kind, uint32_t log_n,
                             uint32_t* public_values_out, uint8_t** proof, size_t* proof_len);
/* FULL profile (params.width = ZKIR_AIR_FULL_WIDTH) from the write log + MEMORY log of zkir_vm_run_writelog_mem_cb (or a Rust recorder hooked
 * into Memory::record_op, memory.rs:243-253): mem_old / mem_pts [n_rows] = on load / store rows the aligned 8-byte word before the access and
 * the timestamp (cycle + 1, 0 = never) of its previous access; mem_widx / mem_word / mem_ts [n_words] = the touched words in strictly
 * ascending order, their final contents and last timestamps.  28 B/row cross PCIe; the device rebuilds the registers and expands the 248
 * columns (no host replay); rows the AIR cannot constrain return ZKIR_ERR_AIR as in zkir_pack_trace_full.  Needs zkir_b200_set_program (and
 * zkir_b200_set_io) first.  A log inconsistent with the program's loads and stores gives a proof the verifier rejects, never a wrong one. */
int zkir_b200_prove_writelog_mem(zkir_ctx*, const zkir_params*, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog,
                                 const uint64_t* mem_old, const uint32_t* mem_pts, uint64_t n_rows, const uint64_t* mem_widx,
                                 const uint64_t* mem_word, const uint32_t* mem_ts, size_t n_words, uint64_t final_pc, uint32_t entry_point,
                                 uint64_t exit_code, int halt_kind, uint32_t log_n, uint32_t* public_values_out, uint8_t** proof, size_t* proof_len);
/* Program -> Proof in one call, the drop-in for `zkir_runtime::prove(program, inputs)` (absent upstream: lib.rs:29-62): runs the
 * interpreter (vm.cc) with the register write log recorded straight into pinned memory, uploads the log in chunks WHILE the
 * interpreter is still running, then proves.  Sets the context's program itself.  out_cycles / out_log_n may be NULL.
 * params.width selects the profile the program needs (zkir_program_profile); with ZKIR_AIR_FULL_WIDTH the interpreter also records the
 * memory log (zkir_vm_run_writelog_mem_cb, 28 B/cycle) and the call continues as zkir_b200_prove_writelog_mem (on a sharded context: full
 * rows and zkir_b200_prove_rows). */
int zkir_b200_prove_program(zkir_ctx*, const zkir_params*, const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data,
                            uint32_t entry_point, const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles,
                            uint32_t* public_values_out, uint64_t* out_cycles, uint32_t* out_log_n, uint8_t** proof, size_t* proof_len);
int zkir_b200_expand_writelog(zkir_ctx*, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog, uint64_t n_rows,
                              uint64_t final_pc, uint32_t log_n, uint32_t* d_cols);
/* the device converter alone (parity tests): rows -> d_cols [width][1 << log_n] canonical, device memory */
int zkir_b200_expand_rows(zkir_ctx*, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows,
                          const uint64_t* final_regs, uint64_t final_pc, uint32_t log_n, uint32_t* d_cols);
/* ... and of the full profile: d_cols [ZKIR_AIR_FULL_WIDTH][1 << log_n]; the host replays the run's memory (which word and previous
 * timestamp every load / store sees), the device expands the rows; bit-identical to zkir_pack_rows_full */
int zkir_b200_expand_rows_full(zkir_ctx*, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows,
                               const uint64_t* final_regs, uint64_t final_pc, uint32_t log_n, uint32_t* d_cols);
/* many independent small proofs of the context's program (BASELINE config 4); traces[i] is host memory [width][1<<log_ns[i]],
 * ios[i] / n_ios[i] the public I/O transcript of execution i */
int zkir_b200_prove_batch(zkir_ctx*, const zkir_params*, const uint32_t* const* traces, const uint32_t* log_ns,
                          const uint32_t* const* public_values, const uint32_t* const* ios, const size_t* n_ios, uint32_t n_proofs,
                          uint8_t** proofs, size_t* proof_lens);
void zkir_b200_free_proof(uint8_t*);
size_t zkir_b200_proof_size(const zkir_params*, uint32_t log_n); /* bytes; depends only on the shape */
/* ---- ONE proof sharded over several GPUs of a box (BASELINE config 5), one context per GPU, normally one
 * process per GPU.  (No reference counterpart.)  Rank 0 draws an id, the host distributes it (torch.distributed / MPI / a
 * pipe), every rank calls comm_init: NCCL (bound with dlopen at this point, ZKIR_NCCL_LIB overrides the library) links the
 * contexts.  From then on every zkir_b200_prove* call on these contexts is COLLECTIVE: all ranks call it with the same trace
 * and parameters.  The LDE matrix is replicated; the Merkle leaf ranges of the trace, quotient and large FRI-layer trees are
 * cut into `world` contiguous segments (rank g hashes leaves [g*M/world, (g+1)*M/world) and builds that subtree), the
 * segment roots are exchanged with one all-gather of world*8 words per tree, and the authentication-path pieces each rank owns
 * are merged with one all-reduce at the end.  Every rank returns the same proof, bit-identical to the single-GPU proof. */
#define ZKIR_COMM_ID_LEN 128
int zkir_b200_comm_unique_id(uint8_t id[ZKIR_COMM_ID_LEN]);
int zkir_b200_comm_init(zkir_ctx*, const uint8_t id[ZKIR_COMM_ID_LEN], int rank, int world); /* world: power of two <= 64 */
int zkir_b200_comm_shutdown(zkir_ctx*);
/* test hook: run the sharded code path for `shards` segments on ONE GPU without NCCL (the context computes every segment in
 * turn); min_segment_leaves != 0 lowers the size below which a tree is not sharded (default 4096 leaves per segment). */
int zkir_b200_emulate_shards(zkir_ctx*, uint32_t shards, uint64_t min_segment_leaves);
/* the partition a sharded proof uses (host arithmetic only, no GPU needed): out = {sharded?, col_lo, col_hi, row_j0, row_count,
 * plane_lo, plane_hi, leaves_per_segment}: rank transforms trace columns [col_lo, col_hi) and quotient planes [plane_lo, plane_hi),
 * and owns the points [row_j0, row_j0 + row_count) of every coset.  min_segment_leaves = 0: the default threshold (4096). */
int zkir_b200_shard_plan(uint32_t world, uint32_t rank, const zkir_params*, uint32_t log_n, uint64_t min_segment_leaves, uint64_t out[8]);

/* CPU verifier (host code, no GPU needed).  0 = accept, ZKIR_ERR_VERIFY = reject (reason via last_error(NULL)).  `code` = the
 * program the proof is about (the verifier evaluates the ROM polynomials itself and absorbs the program digest). */
int zkir_b200_verify(const zkir_params*, const uint8_t* proof, size_t len, const uint32_t* public_values, const uint32_t* code,
                     size_t n_code, const uint32_t* io_events, size_t n_io);
/* host helpers shared by the interpreter, the prover's ROM builder and guest-side tooling */
void zkir_host_poseidon2_permute(uint32_t state16[16]);          /* width-16 Poseidon2 of docs/PROVER_SPEC.md section 2, canonical */
void zkir_rom_entry(uint32_t word, uint32_t* dec, uint32_t* imm); /* decoded ROM row of one instruction word (spec section 3.3) */
/* public columns of the AIR profile `width` (spec sections 3.3, 3.7): one row / the number of rows that can differ from the default
 * row (i = ~0) / all columns, column-major [pub width][2^log_n] */
void zkir_public_row(uint32_t width, uint64_t i, const uint32_t* code, size_t n_code, uint32_t* out);
uint64_t zkir_public_rows(uint32_t width, size_t n_code);
uint64_t zkir_image_words(size_t n_code); /* aligned 8-byte words of the initial memory image [0, 0x1000 + 4 n_code); RAM starts here */
void zkir_public_columns(uint32_t width, uint32_t log_n, const uint32_t* code, size_t n_code, uint32_t* cols);
void zkir_program_digest(const uint32_t* code, size_t n_code, uint32_t digest8[8]); /* what the transcript absorbs for the program */
void zkir_io_digest(const uint32_t* io_events, size_t n_io, uint32_t digest8[8]);   /* ... and for the public I/O transcript */

/* ---- per-kernel entry points (parity tests, ncu captures, roofline harness).  Device pointers, canonical values. */
/* batched NTT over `n_cols` contiguous columns of length 1<<log_n, in place, natural order in and out.
 * coset_shift != 0 (forward only): evaluates on shift*H instead of H. */
int zkir_b200_ntt(zkir_ctx*, uint32_t* d_cols, uint32_t n_cols, uint32_t log_n, int inverse, uint32_t coset_shift);
/* low-degree extension: d_in [n_cols][1<<log_n] evaluations on H -> d_out [n_cols][1<<(log_n+log_blowup)] on 31*H */
int zkir_b200_lde(zkir_ctx*, const uint32_t* d_in, uint32_t* d_out, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup);
int zkir_b200_poseidon2_permute(zkir_ctx*, uint32_t* d_states /* [n][16] */, uint64_t n);
/* leaf = sponge over a row of the column-major matrix; d_tree receives (2*rows-1)*8 words (leaves first) */
int zkir_b200_merkle_commit(zkir_ctx*, const uint32_t* d_matrix, uint32_t n_cols, uint32_t log_rows, uint32_t* d_tree,
                            uint32_t root[8]);
/* quotient values of the AIR on the LDE coset: d_lde [width + 16][M] (main then aux columns), d_publde [4][M] (public columns),
 * lookup = {z[4], theta[4]}, out d_q [4][M], M = 1<<(log_n+log_blowup); all natural order */
int zkir_b200_quotient(zkir_ctx*, const zkir_params*, const uint32_t* d_lde, const uint32_t* d_publde, uint32_t log_n,
                       const uint32_t* public_values, const uint32_t lookup[8], const uint32_t alpha[4], uint32_t* d_q); /* uses the context's I/O transcript */
/* the LogUp aux columns of a trace for given lookup challenges: d_trace [width][N] canonical -> d_aux [16][N] canonical; uses the
 * context's program for the ROM columns.  (The prover draws the challenges from the transcript; this entry point is for parity tests.) */
int zkir_b200_aux_columns(zkir_ctx*, const uint32_t* d_trace, uint32_t log_n, const uint32_t lookup[8], uint32_t* d_aux);
/* one FRI fold: d_in [1<<log_n][4] (ext4, AoS) on shift*H -> d_out [1<<(log_n-1)][4] */
int zkir_b200_fri_fold(zkir_ctx*, const uint32_t* d_in, uint32_t* d_out, uint32_t log_n, uint32_t shift, const uint32_t beta[4]);
/* device memory helpers so ctypes callers need no CUDA binding of their own */
int zkir_b200_dev_alloc(zkir_ctx*, void** d_ptr, size_t bytes);
int zkir_b200_dev_free(zkir_ctx*, void* d_ptr);
int zkir_b200_h2d(zkir_ctx*, void* d_dst, const void* h_src, size_t bytes);
int zkir_b200_d2h(zkir_ctx*, void* h_dst, const void* d_src, size_t bytes);
int zkir_b200_sync(zkir_ctx*);
/* stage timings (ms, CUDA events on the ctx stream) of the last zkir_b200_prove*: see ZKIR_STAGE_* */
#define ZKIR_STAGE_H2D 0
#define ZKIR_STAGE_LDE 1
#define ZKIR_STAGE_TRACE_COMMIT 2
#define ZKIR_STAGE_AUX 3 /* LogUp aux columns: generation, LDE, commitment */
#define ZKIR_STAGE_QUOTIENT 4
#define ZKIR_STAGE_QUOTIENT_COMMIT 5
#define ZKIR_STAGE_OPENINGS 6
#define ZKIR_STAGE_FRI 7
#define ZKIR_STAGE_QUERIES_D2H 8
#define ZKIR_STAGE_COUNT 9
int zkir_b200_last_stage_ms(zkir_ctx*, float out[ZKIR_STAGE_COUNT]);
uint64_t zkir_b200_kernel_launches(const zkir_ctx*); /* kernels launched by this ctx so far */
/* CUDA-event stopwatch on the context's own stream (the per-kernel entry points launch there, so an outside event on
 * another stream would not see them): start records an event, stop records a second one, waits for it and returns ms. */
int zkir_b200_timer_start(zkir_ctx*);
int zkir_b200_timer_stop(zkir_ctx*, float* ms);

/* ---- host side above the ABI: the interpreter that produces the trace.  Restates, in C++ (no Rust toolchain in
 * this image), zkir-runtime's VM: VM::new/VM::run (vm.rs:138-358), execute (execute.rs:35-673), Memory
 * (memory.rs:243-489), syscalls (syscall.rs:94-177), encode/decode (encoder.rs:18-151, decoder.rs:20-192). */
#define ZKIR_HALT_EXIT 0        /* HaltReason::Exit(code) */
#define ZKIR_HALT_EBREAK 1      /* HaltReason::Ebreak */
#define ZKIR_HALT_CYCLE_LIMIT 2 /* HaltReason::CycleLimit */
typedef struct { /* zkir-spec/src/trace.rs:149-167 (MemoryOp; the bound field is host bookkeeping and omitted) */
  uint64_t address, value, timestamp;
  uint8_t is_write, width;
} zkir_mem_op;
typedef struct zkir_vm_result zkir_vm_result;

uint32_t zkir_encode(uint32_t opcode, uint32_t r_a, uint32_t r_b, uint32_t r_c, int32_t imm);
int zkir_decode(uint32_t word, uint32_t out5[5]);
int zkir_vm_run(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, int record_trace, zkir_vm_result** out);
/* SYS_POSEIDON2 (syscall 4) errors by default like the reference's stub (crypto.rs:299-315); on != 0 enables this build's
 * semantics for the calling thread: 16 LE u32 words at R11 (mod p) -> width-16 Poseidon2 -> 16 words at R13, R10 <- 0 */
void zkir_vm_enable_poseidon2(int on);
void zkir_vm_free(zkir_vm_result*);
const char* zkir_vm_last_error(void);
uint64_t zkir_vm_cycles(const zkir_vm_result*);
int zkir_vm_halt_kind(const zkir_vm_result*);
uint64_t zkir_vm_exit_code(const zkir_vm_result*);
size_t zkir_vm_num_outputs(const zkir_vm_result*);
const uint64_t* zkir_vm_outputs(const zkir_vm_result*);
size_t zkir_vm_trace_len(const zkir_vm_result*);
const uint64_t* zkir_vm_trace_pc(const zkir_vm_result*);
const uint32_t* zkir_vm_trace_instr(const zkir_vm_result*);
const uint64_t* zkir_vm_trace_regs(const zkir_vm_result*); /* [cycle][16], PRE-state (vm.rs:245-253) */
const uint64_t* zkir_vm_trace_aux(const zkir_vm_result*);
const uint64_t* zkir_vm_trace_memop_begin(const zkir_vm_result*);
const zkir_mem_op* zkir_vm_trace_memops(const zkir_vm_result*);
int zkir_vm_trace_writelog(const zkir_vm_result*, uint32_t* pcs32 /*[trace_len]*/, uint64_t* wlog /*[trace_len]*/);
/* the recorder the north star describes: the interpreter appends (pc, word, (reg << 56) | value) for every cycle STRAIGHT into the
 * caller's arrays (pinned memory, zkir_b200_alloc_pinned) while it executes -- where VMState::write_reg is called upstream
 * (zkir-runtime/src/state.rs:76-91).  No per-cycle register snapshot is taken.  zkir_vm_logged_rows = cycles. */
int zkir_vm_run_writelog(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                         const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs,
                         uint64_t* wlog, uint64_t capacity, zkir_vm_result** out);
/* same with a progress callback every chunk_rows cycles (rows before rows_done are final); zkir_b200_prove_program uses it to overlap
 * the host->device copy of the log with the execution */
int zkir_vm_run_writelog_cb(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                            const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs,
                            uint64_t* wlog, uint64_t capacity, void (*on_chunk)(void* user, uint64_t rows_done), void* user,
                            uint64_t chunk_rows, zkir_vm_result** out);
uint64_t zkir_vm_logged_rows(const zkir_vm_result*);
/* public I/O transcript of the run: zkir_vm_io_len events of 4 words {cycle, kind (0 READ, 1 WRITE), value lo20, value hi20}, one per
 * READ / WRITE ecall in execution order (syscall.rs:104-119); part of the statement a proof makes (zkir_b200_set_io, zkir_b200_verify) */
size_t zkir_vm_io_len(const zkir_vm_result*);
const uint32_t* zkir_vm_io(const zkir_vm_result*);
/* write log + MEMORY log, for programs that need the full AIR profile (docs/PROVER_SPEC.md 3.8): like zkir_vm_run_writelog_cb, and on
 * every load / store row mem_old / mem_pts [capacity] receive the aligned 8-byte word before the access and the timestamp (cycle + 1;
 * 0 = never) of that word's previous access -- what the memory argument needs and only the interpreter has for free (upstream:
 * Memory::record_op, memory.rs:243-253).  zkir_vm_memlog_*: the touched words in ascending order, their final contents and last timestamps. */
int zkir_vm_run_writelog_mem_cb(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                                const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs, uint64_t* wlog,
                                uint64_t* mem_old, uint32_t* mem_pts, uint64_t capacity, void (*on_chunk)(void*, uint64_t), void* user,
                                uint64_t chunk_rows, zkir_vm_result** out);
size_t zkir_vm_memlog_count(const zkir_vm_result*);
const uint64_t* zkir_vm_memlog_widx(const zkir_vm_result*);
const uint64_t* zkir_vm_memlog_word(const zkir_vm_result*);
const uint32_t* zkir_vm_memlog_ts(const zkir_vm_result*);
/* Poseidon2Witness records of a traced run (zkir-spec/src/trace.rs:287-304): one per SYS_POSEIDON2 call, 34 words each = timestamp
 * (cycle, lo / hi word), input_state[16], output_state[16] (canonical field elements).  The permutation itself is not part of the AIR yet
 * (docs/PROVER_SPEC.md 3.5); a batch of these is what zkir_b200_poseidon2_permute re-computes on the device. */
size_t zkir_vm_poseidon2_count(const zkir_vm_result*);
const uint32_t* zkir_vm_poseidon2_witness(const zkir_vm_result*);
size_t zkir_vm_code_len(const zkir_vm_result*);
const uint32_t* zkir_vm_code(const zkir_vm_result*);
uint64_t zkir_vm_final_pc(const zkir_vm_result*);
const uint64_t* zkir_vm_final_regs(const zkir_vm_result*);

/* "converter" TraceRow -> field columns (named but absent in the reference: zkir-spec/src/trace.rs:41,
 * zkir-runtime/src/vm.rs:243-244).  Writes cols[width][1<<log_n] (host, pinned recommended) and
 * public_values[5] = {entry_pc, num_cycles, exit_lo, exit_hi, halted}.  min log_n via zkir_pack_min_log_n (>= 10: the range
 * table and the program ROM occupy trace rows). */
uint32_t zkir_pack_min_log_n(const zkir_vm_result*);
int zkir_pack_trace(const zkir_vm_result*, uint32_t entry_point, uint32_t log_n, uint32_t* cols, uint32_t* public_values);      /* core: 88 columns */
int zkir_pack_trace_full(const zkir_vm_result*, uint32_t entry_point, uint32_t log_n, uint32_t* cols, uint32_t* public_values); /* full: ZKIR_AIR_FULL_WIDTH columns */
/* the same from plain arrays (what an upstream `ExecutionResult.execution_trace` holds, vm.rs:54-78; arguments as zkir_b200_prove_rows) */
int zkir_pack_rows(const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows, const uint64_t* final_regs, uint64_t final_pc,
                   const uint32_t* code, size_t n_code, uint32_t entry_point, uint64_t exit_code, int halt_kind, uint32_t log_n, uint32_t* cols,
                   uint32_t* public_values);
int zkir_pack_rows_full(const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows, const uint64_t* final_regs,
                        uint64_t final_pc, const uint32_t* code, size_t n_code, uint32_t entry_point, uint64_t exit_code, int halt_kind,
                        uint32_t log_n, uint32_t* cols, uint32_t* public_values);
/* which AIR profile a program needs (execute.rs:35-673 by opcode, zkir-spec/src/opcode.rs:24-144): 1 = core, 0 = full,
 * -1 = it contains an undefined opcode.  The ROM is public: prover and verifier agree. */
int zkir_program_profile(const uint32_t* code, size_t n_code);

#ifdef __cplusplus
}
#endif
#endif
