// Device-side trace converter: raw interpreter rows -> the 72 BabyBear columns of the core AIR v1.
//
// The reference's hand-off type is `Vec<TraceRow>` -- cycle, pc, instruction word and the PRE-state registers
// (zkir-spec/src/trace.rs:24-50, recorded at zkir-runtime/src/vm.rs:245-253,302-312); the "converter" that turns rows
// into field columns is named there (trace.rs:41, vm.rs:243-244) but absent.  zkir_b200/csrc/host/pack.cc is the host
// restatement; this kernel is the same function with one thread per row, so that only the raw rows (140 B/row instead
// of 288 B/row of columns) cross PCIe.  The two must agree bit for bit (tests/test_gpu_parity.py::test_expand_*).
//
// Bound: HBM writes (288 B/row) -- every store of a warp is one 128 B segment of one column.
#include <cuda_runtime.h>
#include "bb.cuh"
#include "kernels.h"
#include "air_columns.h"

namespace zkir {

__device__ __forceinline__ int sext_dev(u32 v, int bits) { const int sh = 32 - bits; return ((int)(v << sh)) >> sh; }

// One row: rg = PRE-state registers, w = instruction word, read_val = post-state r10 (only used by READ rows).
__device__ __forceinline__ void expand_row(u64 i, u64 N, u64 T, const u64 (&rg)[16], u64 pc, u32 w, u64 read_val, u32* cols, u64* errp,
                                           u32 col_lo, u32 col_hi) {
  const u64 LIMB = (1u << 20) - 1, M40 = (1ull << 40) - 1;
  const bool live = i < T;
  u32 err = 0;
  if (pc + 4 >= (1u << 30)) err = 1;
#pragma unroll
  for (int k = 1; k < 16; k++) if (rg[k] >> 40) err = 2;
  u32* col = cols + i;
  auto W = [&](int c, u32 v) { if ((u32)c >= col_lo && (u32)c < col_hi) col[(u64)c * N] = v; };

  W(ZKIR_COL_CLK, (u32)((live ? i : T) % BB_P));
  W(ZKIR_COL_PC, (u32)pc);
#pragma unroll
  for (int k = 1; k < 16; k++) { W(ZKIR_COL_R1_LO + 2 * (k - 1), (u32)(rg[k] & LIMB)); W(ZKIR_COL_R1_LO + 2 * (k - 1) + 1, (u32)(rg[k] >> 20)); }

  u32 rd = 0, rs1 = 0, rs2 = 0, writes = 0;  // writes: ADD, SUB, ADDI, JAL, READ
  u32 s_add = 0, s_sub = 0, s_addi = 0, s_beq = 0, s_bne = 0, s_jal = 0, s_pad = 0;
  u32 is_exit = 0, is_read = 0, is_write = 0;
  u64 av = 0, bv = 0, cv = 0;
  long long imm = 0;
  bool has_imm = false;
  u32 carry0 = 0, carry1 = 0, inv_lo = 0, inv_hi = 0, taken = 0;
  bool is_branch = false;
  if (!live) {
    s_pad = 1;
  } else {
    const u32 op = w & 0x7F;
    const u32 fa = (w >> 7) & 0xF, fb = (w >> 11) & 0xF, fc = (w >> 15) & 0xF;
    auto R = [&](u32 k) -> u64 {  // register read without dynamic indexing of the local array
      u64 v = 0;
#pragma unroll
      for (int j = 1; j < 16; j++) if (k == (u32)j) v = rg[j];
      return v;
    };
    if (op == 0x00 || op == 0x01) {          // ADD / SUB (execute.rs:43-78)
      rd = fa; rs1 = fb; rs2 = fc; writes = 1;
      av = R(rs1); bv = R(rs2);
      if (op == 0) s_add = 1; else s_sub = 1;
    } else if (op == 0x08) {                 // ADDI (execute.rs:185-197)
      rd = fa; rs1 = fb; imm = sext_dev((w >> 15) & 0x1FFFF, 17); has_imm = true; writes = 1;
      av = R(rs1); bv = (u64)imm & M40;
      s_addi = 1;
    } else if (op == 0x40 || op == 0x41) {   // BEQ / BNE, B-type: rs1 bits 10:7, rs2 bits 14:11 (encoder.rs:132-140)
      rs1 = fa; rs2 = fb; imm = sext_dev((w >> 15) & 0x1FFFF, 17); has_imm = true;
      av = R(rs1); bv = R(rs2);
      if (op == 0x40) s_beq = 1; else s_bne = 1;
    } else if (op == 0x48) {                 // JAL (execute.rs:639-647)
      rd = fa; imm = sext_dev((w >> 11) & 0x1FFFFF, 21); has_imm = true; writes = 1;
      cv = pc + 4;
      s_jal = 1;
    } else if (op == 0x50) {                 // ECALL (syscall.rs:94-119)
      const u64 num = rg[10];            // an ECALL row is is_exit + is_read + is_write
      if (num == 0) is_exit = 1;
      else if (num == 1) {                   // READ: the value is the post-state r10
        is_read = 1; rd = 10; writes = 1;
        cv = read_val;
        if (cv >> 40) err = 3;
      } else if (num == 2) is_write = 1;
      else err = 4;
    } else {
      err = 5;
    }
    const u64 a_lo = av & LIMB, a_hi = av >> 20, b_lo = bv & LIMB, b_hi = bv >> 20;
    if (op == 0x00 || op == 0x08) {
      cv = (av + bv) & M40;
      const u64 k0 = (a_lo + b_lo) >> 20;
      carry0 = (u32)k0; carry1 = (u32)((a_hi + b_hi + k0) >> 20);
    } else if (op == 0x01) {
      cv = (av - bv) & M40;
      const u64 k0 = a_lo < b_lo;
      carry0 = (u32)k0; carry1 = (u32)(a_hi < b_hi + k0);
    }
    if (op == 0x40 || op == 0x41) {
      // a branch row has no result, no carries and no destination register: the inverses live in the c cells, the "limb differs"
      // flags in the carry cells and `taken` in rd_l[1] (tools/gen_air.py, column layout)
      const u32 d_lo = bb_sub((u32)a_lo, (u32)b_lo), d_hi = bb_sub((u32)a_hi, (u32)b_hi);
      const u32 ne_lo = d_lo != 0, ne_hi = d_hi != 0;
      inv_lo = d_lo ? bb_from_mont(bb_inv(bb_to_mont(d_lo))) : 0;
      inv_hi = d_hi ? bb_from_mont(bb_inv(bb_to_mont(d_hi))) : 0;
      is_branch = true; carry0 = ne_lo; carry1 = ne_hi;
      const u32 ne = ne_lo | ne_hi;
      taken = op == 0x41 ? ne : !ne;
    }
  }
  u32 imm_lo = 0, imm_sign = 0;  // the high limb is the sign extension (0 or 2^20-1): not a column
  if (has_imm) {
    imm_lo = (u32)(((u64)imm & M40) & LIMB);
    imm_sign = imm < 0;
  }
  W(ZKIR_COL_IMM_LO, imm_lo); W(ZKIR_COL_IMM_SIGN, imm_sign);
  W(ZKIR_COL_S_ADD, s_add); W(ZKIR_COL_S_SUB, s_sub); W(ZKIR_COL_S_ADDI, s_addi); W(ZKIR_COL_S_BEQ, s_beq);
  W(ZKIR_COL_S_BNE, s_bne); W(ZKIR_COL_S_JAL, s_jal);   // s_pad = 1 - (the others) is not a column
#pragma unroll
  for (int k = 0; k < 3; k++) {  // register index = 4*h + l, two 4-way one-hots each (entry 3 implied); rdw[h] = rd_h[h] * writes
    W(ZKIR_COL_RD_H0 + k, (rd >> 2) == (u32)k); W(ZKIR_COL_RD_L0 + k, ((taken ? 1u : (rd & 3u))) == (u32)k);
    W(ZKIR_COL_RDW0 + k, ((rd >> 2) == (u32)k) ? writes : 0u);
    W(ZKIR_COL_RS1_H0 + k, (rs1 >> 2) == (u32)k); W(ZKIR_COL_RS1_L0 + k, (rs1 & 3u) == (u32)k);
    W(ZKIR_COL_RS2_H0 + k, (rs2 >> 2) == (u32)k); W(ZKIR_COL_RS2_L0 + k, (rs2 & 3u) == (u32)k);
  }
  W(ZKIR_COL_A_LO, (u32)(av & LIMB)); W(ZKIR_COL_A_HI, (u32)(av >> 20));
  W(ZKIR_COL_B_LO, (u32)(bv & LIMB)); W(ZKIR_COL_B_HI, (u32)(bv >> 20));
  W(ZKIR_COL_C_LO, is_branch ? inv_lo : (u32)(cv & LIMB)); W(ZKIR_COL_C_HI, is_branch ? inv_hi : (u32)(cv >> 20));
  W(ZKIR_COL_CARRY0, carry0); W(ZKIR_COL_CARRY1, carry1);
  W(ZKIR_COL_IS_EXIT, is_exit); W(ZKIR_COL_IS_READ, is_read); W(ZKIR_COL_IS_WRITE, is_write);
  if (err) {  // first offending row wins; the host reports it after the stream is drained
    const unsigned long long packed = (i << 8) | err;
    atomicMin(reinterpret_cast<unsigned long long*>(errp), packed);
  }
}


__global__ void __launch_bounds__(128) trace_expand_kernel(ExpandArgs a) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const bool live = i < a.T;
  u64 rg[16];
#pragma unroll
  for (int k = 0; k < 16; k++) rg[k] = live ? a.regs[16 * i + k] : a.final_regs[k];
  const u64 read_val = (i + 1 < a.T) ? a.regs[16 * (i + 1) + 10] : a.final_regs[10];
  expand_row(i, a.N, a.T, rg, live ? a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, a.cols, a.err, a.col_lo, a.col_hi);
}

// ---- write-log input: rebuild the pre-state registers with a last-writer scan.
// wlog[i] = (k << 56) | value if row i changed register k.  Chunk = WL_CHUNK rows = one block, one row per thread (so
// every column store of a warp is still one 128 B segment).
//   pass 1: per chunk and register, the last row of the chunk that wrote it (shared-memory atomicMax)
//   pass 2: one block turns that into "last writer before the chunk starts" (running maximum over chunks)
//   pass 3: per chunk, an exclusive prefix maximum over the rows (warp shuffles + one cross-warp step) gives every row
//           its last writer of each register; the register value is gathered from wlog (recent rows: L2 hits).
#define WL_CHUNK 256
#define WL_SCAN_THREADS 512
__global__ void __launch_bounds__(WL_CHUNK) wl_chunk_last_kernel(const u64* __restrict__ wlog, u64 T, int* __restrict__ chunk_last) {
  __shared__ int last[16];
  if (threadIdx.x < 16) last[threadIdx.x] = -1;
  __syncthreads();
  const u64 i = (u64)blockIdx.x * WL_CHUNK + threadIdx.x;
  if (i < T) {
    const u32 k = (u32)(wlog[i] >> 56) & 15u;
    if (k) atomicMax(&last[k], (int)i);
  }
  __syncthreads();
  if (threadIdx.x < 16) chunk_last[blockIdx.x * 16 + threadIdx.x] = last[threadIdx.x];
}
// in place: chunk_last[c][k] becomes the last writer of k BEFORE chunk c.  One block; each thread owns a contiguous range.
__global__ void __launch_bounds__(WL_SCAN_THREADS) wl_chunk_scan_kernel(int* cl, u32 n_chunks) {
  __shared__ int tot[WL_SCAN_THREADS][16];
  const u32 t = threadIdx.x, per = (n_chunks + WL_SCAN_THREADS - 1) / WL_SCAN_THREADS;
  const u32 c0 = t * per, c1 = c0 + per < n_chunks ? c0 + per : n_chunks;
  int run[16];
#pragma unroll
  for (int k = 0; k < 16; k++) run[k] = -1;
  for (u32 c = c0; c < c1; c++) {
#pragma unroll
    for (int k = 0; k < 16; k++) { const int v = cl[c * 16 + k]; run[k] = v > run[k] ? v : run[k]; }
  }
#pragma unroll
  for (int k = 0; k < 16; k++) tot[t][k] = run[k];
  __syncthreads();
  if (t < 16) {
    int r = -1;
    for (u32 u = 0; u < WL_SCAN_THREADS; u++) { const int v = tot[u][t]; tot[u][t] = r; r = v > r ? v : r; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; k++) run[k] = tot[t][k];
  for (u32 c = c0; c < c1; c++) {
#pragma unroll
    for (int k = 0; k < 16; k++) { const int v = cl[c * 16 + k]; cl[c * 16 + k] = run[k]; run[k] = v > run[k] ? v : run[k]; }
  }
}
__global__ void __launch_bounds__(WL_CHUNK) trace_expand_wl_kernel(WlArgs a) {
  __shared__ int warp_tot[WL_CHUNK / 32][16];
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 i = (u64)blockIdx.x * WL_CHUNK + tid;
  const u64 M40 = (1ull << 40) - 1;
  const u64 wl = i < a.T ? a.wlog[i] : 0;
  const u32 kw = (u32)(wl >> 56) & 15u;
  int before[16];
#pragma unroll
  for (int k = 1; k < 16; k++) {
    int v = kw == (u32)k ? (int)i : -1;  // inclusive prefix maximum inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v = t > v ? t : v; }
    if (lane == 31) warp_tot[warp][k] = v;
    const int ex = __shfl_up_sync(0xffffffffu, v, 1);
    before[k] = lane ? ex : -1;
  }
  __syncthreads();
  const int* prev = a.chunk_prev + (u64)blockIdx.x * 16;
  u64 rg[16];
  rg[0] = 0;
#pragma unroll
  for (int k = 1; k < 16; k++) {
    int v = prev[k];
    for (u32 w = 0; w < warp; w++) { const int t = warp_tot[w][k]; v = t > v ? t : v; }
    v = before[k] > v ? before[k] : v;
    rg[k] = v >= 0 ? (a.wlog[v] & M40) : 0;
  }
  if (i >= a.N) return;
  const bool live = i < a.T;
  // The log word is (k << 56) | value with value < 2^40 and bits 40..55 zero: a caller that logged an unmasked u64 (the reference's
  // write_reg takes any u64, state.rs:76-91) must get ZKIR_ERR_AIR like zkir_pack_trace / prove_rows give, not a proof of a truncated
  // execution.  Checked BEFORE the masks below; a payload without a register index (k = 0 is "nothing written") is rejected too.
  if (live && (((wl >> 40) & 0xFFFFull) || (wl >> 60) || (kw == 0 && wl != 0))) {
    const bool read_row = (a.ins[i] & 0x7F) == 0x50 && rg[10] == 1;
    atomicMin(reinterpret_cast<unsigned long long*>(a.err), (unsigned long long)((i << 8) | (read_row ? 3u : 2u)));
  }
  // READ rows need the post-state r10: the logged value if the row changed r10, else the unchanged pre-state
  const u64 read_val = kw == 10u ? (wl & M40) : rg[10];
  expand_row(i, a.N, a.T, rg, live ? (u64)a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, a.cols, a.err, a.col_lo, a.col_hi);
}

int launch_trace_expand(const ExpandArgs& a, cudaStream_t st, u64* launches) {
  trace_expand_kernel<<<(unsigned)((a.N + 127) / 128), 128, 0, st>>>(a);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
u64 trace_expand_wl_scratch_ints(u64 N) { return ((N + WL_CHUNK - 1) / WL_CHUNK) * 16; }
int launch_trace_expand_wl(const WlArgs& a, cudaStream_t st, u64* launches) {
  const u32 n_chunks = (u32)((a.N + WL_CHUNK - 1) / WL_CHUNK);
  wl_chunk_last_kernel<<<n_chunks, WL_CHUNK, 0, st>>>(a.wlog, a.T, a.chunk_prev);
  wl_chunk_scan_kernel<<<1, WL_SCAN_THREADS, 0, st>>>(a.chunk_prev, n_chunks);
  trace_expand_wl_kernel<<<n_chunks, WL_CHUNK, 0, st>>>(a);
  (*launches) += 3;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace zkir
