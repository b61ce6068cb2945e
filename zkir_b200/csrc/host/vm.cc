// Host side above the C ABI: C++ restatement of the ZKIR v3.4 interpreter, with a linear-time SoA trace
// recorder.  The reference's host code is Rust and no Rust toolchain exists in this image, so this file plays
// the role of zkir-runtime's VM for the drop-in `prove()` path; it mirrors the reference's observable behaviour
// (cycle counts, outputs, halt reasons, trace rows, memory-op placement) and is pinned by the reference's own
// test expectations in tests/test_vm_reference.py.
//
// Follows (file:line in /root/reference):
//   encode               zkir-assembler/src/encoder.rs:18-151
//   decode               zkir-disassembler/src/decoder.rs:20-192
//   VM::new / VM::run    zkir-runtime/src/vm.rs:138-205, 208-358, 362-379
//   execute              zkir-runtime/src/execute.rs:35-673   (40-bit semantics: zkir-spec/src/value.rs:592-697)
//   VMState              zkir-runtime/src/state.rs:55-133
//   Memory               zkir-runtime/src/memory.rs:243-489
//   syscalls             zkir-runtime/src/syscall.rs:18-24, 30-78, 94-177
//   Program bytes        zkir-spec/src/program.rs:170-213, 300-346
// Deliberate difference: the reference collects a row's memory ops by scanning the whole memory trace every
// cycle (vm.rs:287-298, O(n^2)); here the recorder remembers where the cycle's ops start.  The filter it applies
// (timestamp == cycle && address != fetch pc) is reproduced exactly.
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <unordered_set>
#include <memory>
#include <algorithm>
#include "../../../include/zkir_b200.h"

namespace {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

const u64 MASK40 = (1ull << 40) - 1;
const u64 CODE_BASE = 0x1000;

enum Op : u32 {
  ADD = 0x00, SUB, MUL, MULH, DIVU, REMU, DIV, REM, ADDI,
  AND = 0x10, OR, XOR, ANDI, ORI, XORI,
  SLL = 0x18, SRL, SRA, SLLI, SRLI, SRAI,
  SLTU = 0x20, SGEU, SLT, SGE, SEQ, SNE, CMOV, CMOVZ, CMOVNZ,
  LB = 0x30, LBU, LH, LHU, LW, LD,
  SB = 0x38, SH, SW, SD,
  BEQ = 0x40, BNE, BLT, BGE, BLTU, BGEU,
  JAL = 0x48, JALR,
  ECALL = 0x50, EBREAK,
};

bool valid_opcode(u32 op) {
  return op <= 0x08 || (op >= 0x10 && op <= 0x15) || (op >= 0x18 && op <= 0x1D) || (op >= 0x20 && op <= 0x28) ||
         (op >= 0x30 && op <= 0x35) || (op >= 0x38 && op <= 0x3B) || (op >= 0x40 && op <= 0x45) || op == 0x48 ||
         op == 0x49 || op == 0x50 || op == 0x51;
}
// format classes (encoder.rs:98-151)
enum Fmt { FR, FI, FSHIFT, FS, FB, FJ, FSYS };
Fmt format_of(u32 op) {
  if (op == ADDI || (op >= ANDI && op <= XORI) || (op >= LB && op <= LD) || op == JALR) return FI;
  if (op >= SLLI && op <= SRAI) return FSHIFT;
  if (op >= SB && op <= SD) return FS;
  if (op >= BEQ && op <= BGEU) return FB;
  if (op == JAL) return FJ;
  if (op == ECALL || op == EBREAK) return FSYS;
  return FR;
}
inline int32_t sext(u32 v, int bits) { int sh = 32 - bits; return ((int32_t)(v << sh)) >> sh; }

struct Memory {  // sparse 4 KiB pages, little endian, uninitialised reads 0 (memory.rs:297-309)
  std::unordered_map<u64, std::unique_ptr<uint8_t[]>> pages;
  u64 last_page = ~0ull; uint8_t* last_ptr = nullptr;
  bool trace_enabled = false;
  u64 timestamp = 0;
  std::vector<zkir_mem_op> trace;
  // decoded-fetch mirror of the code segment [CODE_BASE, code_end): the interpreter fetches from here instead of four hash-map
  // byte reads per cycle (memory.rs:297-309); a store into the segment refreshes the word, so self-modifying code still works
  std::vector<u32> mirror; u64 code_end = 0;
  uint8_t* page(u64 pn, bool create) {
    if (pn == last_page && last_ptr) return last_ptr;
    auto it = pages.find(pn);
    if (it == pages.end()) {
      if (!create) return nullptr;
      std::unique_ptr<uint8_t[]> p(new uint8_t[4096]());
      it = pages.emplace(pn, std::move(p)).first;
    }
    last_page = pn; last_ptr = it->second.get();
    return last_ptr;
  }
  uint8_t rd8(u64 a) { uint8_t* p = page(a >> 12, false); return p ? p[a & 4095] : 0; }
  void wr8(u64 a, uint8_t v) { page(a >> 12, true)[a & 4095] = v; }
  void rec(u64 addr, u64 value, bool w, uint8_t width) {
    if (!trace_enabled) return;
    zkir_mem_op op; op.address = addr; op.value = value; op.timestamp = timestamp; op.is_write = w; op.width = width;
    trace.push_back(op);
  }
  u64 read(u64 a, int width) {
    u64 v = 0;
    for (int i = 0; i < width; i++) v |= (u64)rd8(a + i) << (8 * i);
    rec(a, v, false, (uint8_t)width);
    return v;
  }
  void write(u64 a, u64 v, int width) {
    for (int i = 0; i < width; i++) wr8(a + i, (uint8_t)(v >> (8 * i)));
    rec(a, v, true, (uint8_t)width);
    if (a < code_end && a + width > CODE_BASE) {
      for (u64 wa = (a > CODE_BASE ? a : CODE_BASE) & ~3ull; wa < a + width && wa < code_end; wa += 4) {
        u32 x = 0;
        for (int i = 0; i < 4; i++) x |= (u32)rd8(wa + i) << (8 * i);
        mirror[(wa - CODE_BASE) / 4] = x;
      }
    }
  }
};

}  // namespace

struct zkir_vm_result {
  u64 cycles = 0;
  int halt_kind = ZKIR_HALT_EBREAK;
  u64 exit_code = 0;
  std::vector<u64> outputs;
  u64 final_pc = 0;
  u64 final_regs[16] = {0};
  // SoA trace (one entry per cycle), PRE-state registers (vm.rs:245-253,302-312)
  std::vector<u64> pc;
  std::vector<u32> instr;
  std::vector<u64> regs;       // [cycle][16]
  std::vector<u64> aux;        // READ: value read; WRITE: value written; else 0
  std::vector<u64> memop_begin;  // CSR offsets into memops, size cycles+1
  std::vector<zkir_mem_op> memops;  // data memory ops per row (fetch excluded)
  std::vector<u32> code;            // the program's code words (the ROM the proof is bound to)
  std::vector<u32> io;              // public I/O transcript: 4 words per READ / WRITE ecall (cycle, kind 0/1, value lo20, value hi20)
  std::vector<u64> ml_widx, ml_word;   // memory-log mode: the aligned 8-byte words the run touched (ascending), their final contents
  std::vector<u32> ml_ts;              // ... and the timestamp (cycle + 1) of their last access
  std::vector<u32> pos2;            // Poseidon2Witness records (zkir-spec/src/trace.rs:287-304): 34 words per SYS_POSEIDON2 call --
                                    // timestamp lo / hi, input_state[16], output_state[16] (traced runs only)
  u64 logged = 0;                   // write-log mode: rows written to the caller's arrays
  std::string error;
};

std::string& zkir_host_error() { static thread_local std::string e; return e; }
#define g_vm_error zkir_host_error()

extern "C" {

uint32_t zkir_encode(uint32_t opcode, uint32_t r_a, uint32_t r_b, uint32_t r_c, int32_t imm) {
  // R: rd=r_a rs1=r_b rs2=r_c | I/shift: rd=r_a rs1=r_b imm | S,B: rs1=r_a rs2=r_b imm | J: rd=r_a imm
  u32 w = opcode & 0x7F;
  switch (format_of(opcode & 0x7F)) {
    case FR: w |= (r_a & 0xF) << 7 | (r_b & 0xF) << 11 | (r_c & 0xF) << 15; break;
    case FI: case FSHIFT: w |= (r_a & 0xF) << 7 | (r_b & 0xF) << 11 | ((u32)imm & 0x1FFFF) << 15; break;
    case FS: case FB: w |= (r_a & 0xF) << 7 | (r_b & 0xF) << 11 | ((u32)imm & 0x1FFFF) << 15; break;
    case FJ: w |= (r_a & 0xF) << 7 | ((u32)imm & 0x1FFFFF) << 11; break;
    case FSYS: break;
  }
  return w;
}

// out = {opcode, r_a, r_b, r_c, imm(as u32 two's complement)} with the same field meaning as zkir_encode
int zkir_decode(uint32_t word, uint32_t* out5) {
  u32 op = word & 0x7F;
  if (!valid_opcode(op)) return -1;
  out5[0] = op; out5[1] = (word >> 7) & 0xF; out5[2] = (word >> 11) & 0xF; out5[3] = 0; out5[4] = 0;
  switch (format_of(op)) {
    case FR: out5[3] = (word >> 15) & 0xF; break;
    case FI: case FS: case FB: out5[4] = (u32)sext((word >> 15) & 0x1FFFF, 17); break;
    case FSHIFT: out5[4] = (word >> 15) & 0xFF; break;
    case FJ: out5[2] = 0; out5[4] = (u32)sext((word >> 11) & 0x1FFFFF, 21); break;
    case FSYS: out5[1] = out5[2] = 0; break;
  }
  return 0;
}

const char* zkir_vm_last_error(void) { return g_vm_error.c_str(); }

void zkir_vm_free(zkir_vm_result* r) { delete r; }

static thread_local int g_enable_poseidon2 = 0;
// SYS_POSEIDON2 is a stub that errors upstream (crypto.rs:299-315, pinned by syscall_integration.rs:400-422), and that is the default
// here too; BASELINE config 3 needs the syscall, so it is an opt-in of the calling thread (VMConfig.enable_poseidon2_syscall)
void zkir_vm_enable_poseidon2(int on) { g_enable_poseidon2 = on; }
void zkir_host_poseidon2_permute(uint32_t* state16);   // verify.cc: the Poseidon2 permutation of docs/PROVER_SPEC.md section 2

// One interpreter for the three recording modes: none, full rows (ExecutionResult.execution_trace, trace.rs:24-50) and the
// register write log written STRAIGHT into caller-provided arrays while the program runs (pinned memory: north_star "the CPU
// interpreter records the execution trace into pinned memory"): wl_pcs/wl_ins/wl_log[capacity], 16 B per cycle.
static int vm_run_impl(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                       const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, int record_trace,
                       uint32_t* wl_pcs, uint32_t* wl_ins, uint64_t* wl_log, uint64_t wl_capacity, zkir_vm_result** out,
                       void (*on_chunk)(void*, uint64_t) = nullptr, void* cb_user = nullptr, uint64_t chunk_rows = 0,
                       uint64_t* ml_old = nullptr, uint32_t* ml_pts = nullptr) {
  *out = nullptr;
  if (entry_point < 0x1000) {  // vm.rs:141-147 (the reference panics)
    g_vm_error = "Program appears to be in debug format (entry_point < 0x1000)";
    return ZKIR_ERR_ARG;
  }
  std::unique_ptr<zkir_vm_result> res(new zkir_vm_result());
  res->code.assign(code, code + n_code);
  Memory mem;
  for (size_t i = 0; i < n_code; i++) mem.write(CODE_BASE + 4 * i, code[i], 4);
  for (size_t i = 0; i < n_data; i++) mem.wr8(CODE_BASE + 4 * n_code + i, data[i]);
  mem.mirror.assign(code, code + n_code);
  mem.code_end = CODE_BASE + 4 * n_code;
  mem.trace.clear();
  mem.trace_enabled = record_trace != 0;
  const bool wl = wl_log != nullptr;
  u64 regs[16] = {0};
  u64 pc = entry_point, cycles = 0;
  size_t input_pos = 0;
  bool halted = false;
  u64 cur_log = 0;          // write-log word of the current cycle
  u64 next_chunk = on_chunk ? chunk_rows : ~0ull;   // cycle count at which the next progress callback is due (no per-cycle division)
  u64 wide_row = ~0ull;     // first cycle that wrote a value above 40 bits (the log format cannot hold it)
  auto R = [&](u32 i) -> u64 { return i == 0 ? 0 : regs[i]; };
  auto W = [&](u32 i, u64 v) {
    if (!i) return;
    regs[i] = v;
    if (wl) { if ((v >> 40) && wide_row == ~0ull) wide_row = cycles; cur_log = ((u64)i << 56) | (v & MASK40); }
  };
  // memory log (full AIR profile, docs/PROVER_SPEC.md 3.8): what the memory argument needs to know about a load / store and only the
  // interpreter has for free -- the aligned 8-byte word before the access and the timestamp of that word's previous access
  // The argument starts from the PUBLIC image (the code words; zero elsewhere) and sees loads and stores only: a word whose first
  // access finds anything else (a data segment), or that SYS_POSEIDON2 wrote, is outside it and the run is refused (PACK_ERR_MEMVAL).
  std::unordered_map<u64, u32> ml_last;
  std::unordered_set<u64> ml_untracked;   // words written by a syscall
  std::unordered_map<u64, u64> ml_shadow; // ... and what the argument still holds for those it had seen before
  u64 ml_bad_row = ~0ull;
  auto ML = [&](u64 a) {
    if (!ml_old) return;
    const u64 w = a >> 3;
    const u64 cur = mem.read(w << 3, 8);
    ml_old[cycles] = cur;
    u32& ts = ml_last[w];
    if (ts == 0) {
      u64 init = 0;
      for (int h = 0; h < 2; h++) {
        const u64 wa = (w << 3) + 4 * h;
        if (wa >= CODE_BASE && (wa - CODE_BASE) / 4 < n_code) init |= (u64)code[(wa - CODE_BASE) / 4] << (32 * h);
      }
      if (cur != init && ml_bad_row == ~0ull) ml_bad_row = cycles;
    }
    if (!ml_untracked.empty() && ml_untracked.count(w) && ml_bad_row == ~0ull) ml_bad_row = cycles;
    ml_pts[cycles] = ts;
    ts = (u32)(cycles + 1);
  };
  auto fail = [&](const std::string& m) { g_vm_error = m; return ZKIR_ERR_VM; };
  auto slt40 = [](u64 a, u64 b) { u64 s = 1ull << 39; return ((a & MASK40) ^ s) < ((b & MASK40) ^ s); };
  auto sra40 = [](u64 v, u32 sh) -> u64 {
    bool neg = v & (1ull << 39);
    if (sh >= 40) return neg ? MASK40 : 0;
    u64 r = v >> sh;
    if (neg) r |= ((1ull << sh) - 1) << (40 - sh);
    return r & MASK40;
  };
  if (record_trace) res->memop_begin.push_back(0);

  while (!halted) {
    if (cycles >= max_cycles) { res->halt_kind = ZKIR_HALT_CYCLE_LIMIT; break; }
    mem.timestamp = cycles;
    const u64 fetch_pc = pc;
    if (pc % 4 != 0) { char b[64]; snprintf(b, sizeof b, "Misaligned PC: %#llx", (unsigned long long)pc); return fail(b); }
    const size_t ops_before = mem.trace.size();
    // fetch (vm.rs:362-379): from the mirror of the code segment; anywhere else through memory (reads 0 when unmapped).  The fetch
    // is a memory op of the cycle upstream, but the row filter drops it (address == fetch pc), so it is not recorded here.
    u32 word;
    if (pc >= CODE_BASE && pc < mem.code_end) word = mem.mirror[(pc - CODE_BASE) >> 2];
    else { const bool te = mem.trace_enabled; mem.trace_enabled = false; word = (u32)mem.read(pc, 4); mem.trace_enabled = te; }
    u32 op = word & 0x7F;
    if (!valid_opcode(op)) { char b[64]; snprintf(b, sizeof b, "Decode error: unknown opcode %#x", op); return fail(b); }
    if (record_trace) {
      res->pc.push_back(fetch_pc); res->instr.push_back(word);
      res->regs.insert(res->regs.end(), regs, regs + 16);
      res->aux.push_back(0);
    }
    if (wl) {
      if (cycles >= wl_capacity) return fail("write log is full: raise the capacity (max_cycles) of the pinned arrays");
      if (pc >> 32) return fail("write log: pc above 32 bits");
      wl_pcs[cycles] = (u32)pc; wl_ins[cycles] = word; cur_log = 0;
    }
    const u32 fa = (word >> 7) & 0xF, fb = (word >> 11) & 0xF, fc = (word >> 15) & 0xF;
    const i64 imm17 = sext((word >> 15) & 0x1FFFF, 17);
    const u32 shamt = (word >> 15) & 0xFF;
    u64 next_pc = pc + 4;
    bool is_ecall = false;
    auto misaligned = [&](u64 a, int al) { char b[96]; snprintf(b, sizeof b, "Misaligned access at %#llx (alignment %d)", (unsigned long long)a, al); return fail(b); };
    switch (op) {
      case ADD: W(fa, ((R(fb) & MASK40) + (R(fc) & MASK40)) & MASK40); break;
      case SUB: W(fa, ((R(fb) & MASK40) - (R(fc) & MASK40)) & MASK40); break;
      case MUL: W(fa, ((R(fb) & MASK40) * (R(fc) & MASK40)) & MASK40); break;
      case MULH: { unsigned __int128 p = (unsigned __int128)R(fb) * R(fc); W(fa, (u64)(p >> 40) & MASK40); break; }
      case DIVU: case REMU: case DIV: case REM: {
        u64 a = R(fb), b = R(fc);
        if (b == 0) { char m[64]; snprintf(m, sizeof m, "Division by zero at PC %#llx", (unsigned long long)pc); return fail(m); }
        u64 r;
        if (op == DIVU) r = a / b;
        else if (op == REMU) r = a % b;
        else if (op == DIV) r = ((i64)a == INT64_MIN && (i64)b == -1) ? a : (u64)((i64)a / (i64)b);
        else r = ((i64)a == INT64_MIN && (i64)b == -1) ? 0 : (u64)((i64)a % (i64)b);
        W(fa, r);
        break;
      }
      case ADDI: W(fa, ((R(fb) & MASK40) + ((u64)imm17 & MASK40)) & MASK40); break;
      case AND: W(fa, R(fb) & R(fc) & MASK40); break;
      case OR: W(fa, (R(fb) | R(fc)) & MASK40); break;
      case XOR: W(fa, (R(fb) ^ R(fc)) & MASK40); break;
      case ANDI: W(fa, R(fb) & (u64)imm17 & MASK40); break;
      case ORI: W(fa, (R(fb) | (u64)imm17) & MASK40); break;
      case XORI: W(fa, (R(fb) ^ (u64)imm17) & MASK40); break;
      case SLL: case SLLI: { u32 sh = op == SLL ? (u32)(R(fc) & 0x3F) : shamt; W(fa, sh >= 40 ? 0 : ((R(fb) & MASK40) << sh) & MASK40); break; }
      case SRL: case SRLI: { u32 sh = op == SRL ? (u32)(R(fc) & 0x3F) : shamt; W(fa, sh >= 40 ? 0 : (R(fb) & MASK40) >> sh); break; }
      case SRA: case SRAI: { u32 sh = op == SRA ? (u32)(R(fc) & 0x3F) : shamt; W(fa, sra40(R(fb) & MASK40, sh)); break; }
      case SLTU: W(fa, (R(fb) & MASK40) < (R(fc) & MASK40)); break;
      case SGEU: W(fa, !((R(fb) & MASK40) < (R(fc) & MASK40))); break;
      case SLT: W(fa, slt40(R(fb), R(fc))); break;
      case SGE: W(fa, !slt40(R(fb), R(fc))); break;
      case SEQ: W(fa, R(fb) == R(fc)); break;
      case SNE: W(fa, R(fb) != R(fc)); break;
      case CMOV: case CMOVNZ: if (R(fc) != 0) W(fa, R(fb)); break;
      case CMOVZ: if (R(fc) == 0) W(fa, R(fb)); break;
      case LB: ML(R(fb) + (u64)imm17); W(fa, (u64)(i64)(int8_t)mem.read(R(fb) + (u64)imm17, 1)); break;
      case LBU: ML(R(fb) + (u64)imm17); W(fa, mem.read(R(fb) + (u64)imm17, 1)); break;
      case LH: case LHU: {
        u64 a = R(fb) + (u64)imm17;
        if (a % 2) return misaligned(a, 2);
        ML(a);
        u64 v = mem.read(a, 2);
        W(fa, op == LH ? (u64)(i64)(int16_t)v : v);
        break;
      }
      case LW: { u64 a = R(fb) + (u64)imm17; if (a % 4) return misaligned(a, 4); ML(a); W(fa, mem.read(a, 4)); break; }
      case LD: { u64 a = R(fb) + (u64)imm17; if (a % 8) return misaligned(a, 8); ML(a); W(fa, mem.read(a, 8)); break; }
      // stores: S-type has rs1 (base) in bits 10:7 and rs2 (value) in bits 14:11 (encoder.rs:122-130)
      case SB: ML(R(fa) + (u64)imm17); mem.write(R(fa) + (u64)imm17, R(fb) & 0xFF, 1); break;
      case SH: { u64 a = R(fa) + (u64)imm17; if (a % 2) return misaligned(a, 2); ML(a); mem.write(a, R(fb) & 0xFFFF, 2); break; }
      case SW: { u64 a = R(fa) + (u64)imm17; if (a % 4) return misaligned(a, 4); ML(a); mem.write(a, R(fb) & 0xFFFFFFFFull, 4); break; }
      case SD: { u64 a = R(fa) + (u64)imm17; if (a % 8) return misaligned(a, 8); ML(a); mem.write(a, R(fb), 8); break; }
      case BEQ: if (R(fa) == R(fb)) next_pc = pc + (u64)imm17; break;
      case BNE: if (R(fa) != R(fb)) next_pc = pc + (u64)imm17; break;
      case BLT: if (slt40(R(fa), R(fb))) next_pc = pc + (u64)imm17; break;
      case BGE: if (!slt40(R(fa), R(fb))) next_pc = pc + (u64)imm17; break;
      case BLTU: if ((R(fa) & MASK40) < (R(fb) & MASK40)) next_pc = pc + (u64)imm17; break;
      case BGEU: if (!((R(fa) & MASK40) < (R(fb) & MASK40))) next_pc = pc + (u64)imm17; break;
      case JAL: { i64 off = sext((word >> 11) & 0x1FFFFF, 21); W(fa, pc + 4); next_pc = pc + (u64)off; break; }
      case JALR: { u64 t = R(fb) + (u64)imm17; W(fa, pc + 4); next_pc = t & ~1ull; break; }
      case ECALL: is_ecall = true; break;
      case EBREAK: halted = true; res->halt_kind = ZKIR_HALT_EBREAK; next_pc = pc; break;
    }
    pc = next_pc;
    if (is_ecall) {  // syscall.rs:94-177: number in R10
      u64 num = regs[10];
      switch (num) {
        case 0: halted = true; res->halt_kind = ZKIR_HALT_EXIT; res->exit_code = regs[11]; break;
        case 1: {
          u64 v = input_pos < n_inputs ? inputs[input_pos++] : 0;
          W(10, v);
          if (record_trace) res->aux.back() = v;
          const u32 ev[4] = {(u32)cycles, 0u, (u32)(v & 0xFFFFF), (u32)((v >> 20) & 0xFFFFF)};
          res->io.insert(res->io.end(), ev, ev + 4);
          break;
        }
        case 2: {
          res->outputs.push_back(regs[11]);
          if (record_trace) res->aux.back() = regs[11];
          const u32 ev[4] = {(u32)cycles, 1u, (u32)(regs[11] & 0xFFFFF), (u32)((regs[11] >> 20) & 0xFFFFF)};
          res->io.insert(res->io.end(), ev, ev + 4);
          break;
        }
        case 4: {
          // SYS_POSEIDON2 (syscall.rs:22,140-149).  The reference's implementation is a stub that errors (crypto.rs:299-315), so the
          // semantics are this build's: read 16 little-endian u32 words at R11 (reduced mod p),
          // apply the width-16 Poseidon2 permutation of docs/PROVER_SPEC.md section 2, write the 16 words at R13, R10 <- 0.
          if (!g_enable_poseidon2) return fail("Poseidon2 not yet implemented");  // crypto.rs:306-315 (the reference errors)
          const u64 src = regs[11], dst = regs[13];
          if ((src | dst) % 4) { char b[96]; snprintf(b, sizeof b, "Misaligned access at %#llx (alignment 4)", (unsigned long long)((src % 4) ? src : dst)); return fail(b); }
          uint32_t st[16];
          for (int k = 0; k < 16; k++) st[k] = (uint32_t)(mem.read(src + 4 * k, 4) % ZKIR_BABYBEAR_P);
          if (record_trace) { res->pos2.push_back((u32)cycles); res->pos2.push_back((u32)(cycles >> 32)); res->pos2.insert(res->pos2.end(), st, st + 16); }
          zkir_host_poseidon2_permute(st);
          if (record_trace) res->pos2.insert(res->pos2.end(), st, st + 16);
          if (ml_old) for (int k = 0; k < 16; k++) {   // the argument keeps the value its own loads / stores left in these words
            const u64 w = (dst + 4 * k) >> 3;
            if (ml_untracked.insert(w).second && ml_last.count(w)) ml_shadow[w] = mem.read(w << 3, 8);
          }
          for (int k = 0; k < 16; k++) mem.write(dst + 4 * k, st[k], 4);
          W(10, 0);
          break;
        }
        case 3: case 5: case 6: return fail("hash syscalls (SHA-256/Keccak-256/Blake3) are outside the proving path and not restated here");
        default: { char m[64]; snprintf(m, sizeof m, "Invalid syscall: %llu", (unsigned long long)num); return fail(m); }
      }
    }
    if (wl) {
      wl_log[cycles] = cur_log;
      if (cycles + 1 == next_chunk) { on_chunk(cb_user, next_chunk); next_chunk += chunk_rows; }   // rows [0, next_chunk) are final: upload may start
    }
    if (record_trace) {
      // data memory ops of this cycle: timestamp == cycle && address != fetch pc (vm.rs:291-298)
      for (size_t i = ops_before; i < mem.trace.size(); i++)
        if (mem.trace[i].address != fetch_pc) res->memops.push_back(mem.trace[i]);
      res->memop_begin.push_back(res->memops.size());
      mem.trace.resize(ops_before);  // keep the scratch list bounded
    }
    cycles++;
  }
  res->cycles = cycles;
  res->final_pc = pc;
  memcpy(res->final_regs, regs, sizeof regs);
  if (wl) {
    res->logged = cycles;
    if (wide_row != ~0ull) { g_vm_error = "write log: value above 40 bits at row " + std::to_string(wide_row); return ZKIR_ERR_AIR; }
  }
  if (ml_bad_row != ~0ull) {
    g_vm_error = "AIR v2 cannot constrain row " + std::to_string(ml_bad_row) + ": loaded value differs from the memory the AIR tracks (written by a syscall, or a data segment)";
    return ZKIR_ERR_AIR;
  }
  if (ml_old) {   // the touched words in ascending order with their final contents: the boundary of the memory argument
    std::vector<u64> ws;
    ws.reserve(ml_last.size());
    for (const auto& kv : ml_last) ws.push_back(kv.first);
    std::sort(ws.begin(), ws.end());
    for (u64 w : ws) { res->ml_widx.push_back(w); auto sh = ml_shadow.find(w); res->ml_word.push_back(sh != ml_shadow.end() ? sh->second : mem.read(w << 3, 8)); res->ml_ts.push_back(ml_last[w]); }
  }
  *out = res.release();
  return 0;
}

int zkir_vm_run(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, int record_trace, zkir_vm_result** out) {
  return vm_run_impl(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, record_trace, nullptr, nullptr, nullptr, 0, out);
}

int zkir_vm_run_writelog(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                         const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs, uint64_t* wlog,
                         uint64_t capacity, zkir_vm_result** out) {
  if (!pcs32 || !instrs || !wlog) { g_vm_error = "null write-log arrays"; return ZKIR_ERR_ARG; }
  return vm_run_impl(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, 0, pcs32, instrs, wlog, capacity, out);
}

// same, reporting progress: on_chunk(user, rows_done) is called every chunk_rows cycles from the interpreter's thread; the rows
// before rows_done are final (the prover starts their host->device copy while the interpreter keeps running)
int zkir_vm_run_writelog_cb(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                            const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs, uint64_t* wlog,
                            uint64_t capacity, void (*on_chunk)(void*, uint64_t), void* user, uint64_t chunk_rows, zkir_vm_result** out) {
  if (!pcs32 || !instrs || !wlog || (on_chunk && !chunk_rows)) { g_vm_error = "null write-log arrays"; return ZKIR_ERR_ARG; }
  return vm_run_impl(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, 0, pcs32, instrs, wlog, capacity, out, on_chunk, user, chunk_rows);
}

// write log + memory log (full AIR profile): additionally mem_old[capacity] / mem_pts[capacity] receive, on every load / store row, the
// aligned 8-byte word before the access and the timestamp (cycle + 1, 0 = never) of that word's previous access
int zkir_vm_run_writelog_mem_cb(const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data, uint32_t entry_point,
                                const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pcs32, uint32_t* instrs, uint64_t* wlog,
                                uint64_t* mem_old, uint32_t* mem_pts, uint64_t capacity, void (*on_chunk)(void*, uint64_t), void* user,
                                uint64_t chunk_rows, zkir_vm_result** out) {
  if (!pcs32 || !instrs || !wlog || !mem_old || !mem_pts || (on_chunk && !chunk_rows)) { g_vm_error = "null write-log arrays"; return ZKIR_ERR_ARG; }
  return vm_run_impl(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, 0, pcs32, instrs, wlog, capacity, out, on_chunk, user, chunk_rows,
                     mem_old, mem_pts);
}
size_t zkir_vm_memlog_count(const zkir_vm_result* r) { return r->ml_widx.size(); }
const uint64_t* zkir_vm_memlog_widx(const zkir_vm_result* r) { return r->ml_widx.data(); }
const uint64_t* zkir_vm_memlog_word(const zkir_vm_result* r) { return r->ml_word.data(); }
const uint32_t* zkir_vm_memlog_ts(const zkir_vm_result* r) { return r->ml_ts.data(); }

size_t zkir_vm_io_len(const zkir_vm_result* r) { return r->io.size() / 4; }
const uint32_t* zkir_vm_io(const zkir_vm_result* r) { return r->io.data(); }
size_t zkir_vm_poseidon2_count(const zkir_vm_result* r) { return r->pos2.size() / 34; }
const uint32_t* zkir_vm_poseidon2_witness(const zkir_vm_result* r) { return r->pos2.data(); }
size_t zkir_vm_code_len(const zkir_vm_result* r) { return r->code.size(); }
const uint32_t* zkir_vm_code(const zkir_vm_result* r) { return r->code.data(); }
uint64_t zkir_vm_logged_rows(const zkir_vm_result* r) { return r->logged; }

uint64_t zkir_vm_cycles(const zkir_vm_result* r) { return r->cycles; }
int zkir_vm_halt_kind(const zkir_vm_result* r) { return r->halt_kind; }
uint64_t zkir_vm_exit_code(const zkir_vm_result* r) { return r->exit_code; }
size_t zkir_vm_num_outputs(const zkir_vm_result* r) { return r->outputs.size(); }
const uint64_t* zkir_vm_outputs(const zkir_vm_result* r) { return r->outputs.data(); }
size_t zkir_vm_trace_len(const zkir_vm_result* r) { return r->pc.size(); }
const uint64_t* zkir_vm_trace_pc(const zkir_vm_result* r) { return r->pc.data(); }
const uint32_t* zkir_vm_trace_instr(const zkir_vm_result* r) { return r->instr.data(); }
const uint64_t* zkir_vm_trace_regs(const zkir_vm_result* r) { return r->regs.data(); }
const uint64_t* zkir_vm_trace_memop_begin(const zkir_vm_result* r) { return r->memop_begin.data(); }
const zkir_mem_op* zkir_vm_trace_memops(const zkir_vm_result* r) { return r->memops.data(); }
uint64_t zkir_vm_final_pc(const zkir_vm_result* r) { return r->final_pc; }
const uint64_t* zkir_vm_final_regs(const zkir_vm_result* r) { return r->final_regs; }
const uint64_t* zkir_vm_trace_aux(const zkir_vm_result* r) { return r->aux.data(); }

// Register write log of the recorded rows: pcs32[i] = pc of row i, wlog[i] = (k << 56) | value when row i changed
// register k (pre-state of row i+1 differs from row i; r0 never changes), 0 when it changed nothing.  Returns
// ZKIR_ERR_AIR if a row changed more than one register or wrote a value above 40 bits (the compact format cannot
// express it; use the full-row entry point).  16 B/row instead of 140 B/row cross PCIe.
int zkir_vm_trace_writelog(const zkir_vm_result* r, uint32_t* pcs32, uint64_t* wlog) {
  const size_t T = r->pc.size();
  for (size_t i = 0; i < T; i++) {
    const u64* cur = r->regs.data() + 16 * i;
    const u64* nxt = i + 1 < T ? cur + 16 : r->final_regs;
    u64 w = 0;
    int changed = 0;
    for (int k = 1; k < 16; k++) {
      if (nxt[k] != cur[k]) {
        changed++;
        if (nxt[k] >> 40) { g_vm_error = "write log: value above 40 bits at row " + std::to_string(i); return ZKIR_ERR_AIR; }
        w = ((u64)k << 56) | nxt[k];
      }
    }
    if (changed > 1) { g_vm_error = "write log: more than one register changed at row " + std::to_string(i); return ZKIR_ERR_AIR; }
    if (r->pc[i] >> 32) { g_vm_error = "write log: pc above 32 bits at row " + std::to_string(i); return ZKIR_ERR_AIR; }
    pcs32[i] = (u32)r->pc[i];
    wlog[i] = w;
  }
  return 0;
}

}  // extern "C"
