// Device-side trace converter: raw interpreter rows -> the 88 BabyBear main columns of the AIR v2.
//
// The reference's hand-off type is `Vec<TraceRow>` -- cycle, pc, instruction word and the PRE-state registers
// (zkir-spec/src/trace.rs:24-50, recorded at zkir-runtime/src/vm.rs:245-253,302-312); the "converter" that turns rows
// into field columns is named there (trace.rs:41, vm.rs:243-244) but absent.  The per-row function is air_pack.h, shared with
// the host packer (host/pack.cc); here one thread converts one row, so that only the raw rows (140 B/row, or the 16 B/row
// register write log) cross PCIe instead of 352 B/row of columns.  The LogUp multiplicity columns are histograms: the range-table
// counts go through a shared-memory histogram per block, the ROM counts through warp-aggregated atomics.
//
// Bound: HBM writes (352 B/row) -- every store of a warp is one 128 B segment of one column.
#include <cuda_runtime.h>
#include <stdio.h>
#include "bb.cuh"
#include "kernels.h"
#include "air_pack.h"

namespace zkir {

// ---- write-log input (both profiles): rebuild the pre-state registers of row i with a last-writer scan.
// wlog[i] = (k << 56) | value if row i changed register k.  Chunk = WL_CHUNK rows = one block, one row per thread (so every column store
// of a warp is still one 128 B segment).
//   pass 1 (wl_chunk_last_kernel): per chunk and register, the last row of the chunk that wrote it (shared-memory atomicMax)
//   pass 2 (wl_chunk_scan_kernel): one block turns that into "last writer before the chunk starts" (running maximum over chunks)
//   pass 3 (this function, inside the converter kernels): an exclusive prefix maximum over the rows of the chunk (warp shuffles + one
//           cross-warp step) gives every row its last writer of each register; the value is gathered from wlog (recent rows: L2 hits).
#define WL_CHUNK 256
#define WL_SCAN_THREADS 512
__device__ __forceinline__ void wl_rebuild_registers(const u64* __restrict__ wlog, u64 T, const int* __restrict__ chunk_prev, int (*warp_tot)[16],
                                                     u64 (&rg)[16], u64& wl, u32& kw) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 i = (u64)blockIdx.x * WL_CHUNK + tid;
  const u64 M40 = (1ull << 40) - 1;
  wl = i < T ? wlog[i] : 0;
  kw = (u32)(wl >> 56) & 15u;
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
  const int* prev = chunk_prev + (u64)blockIdx.x * 16;
  rg[0] = 0;
#pragma unroll
  for (int k = 1; k < 16; k++) {
    int v = prev[k];
    for (u32 w = 0; w < warp; w++) { const int t = warp_tot[w][k]; v = t > v ? t : v; }
    v = before[k] > v ? before[k] : v;
    rg[k] = v >= 0 ? (wlog[v] & M40) : 0;
  }
}
// the log word is (k << 56) | value with value < 2^40 and bits 40..55 zero: a caller that logged an unmasked u64 (the reference's write_reg
// takes any u64, state.rs:76-91) must get ZKIR_ERR_AIR like zkir_pack_trace / prove_rows give, not a proof of a truncated execution.
// Checked BEFORE the masks; a payload without a register index (k = 0 is "nothing written") is rejected too.
__device__ __forceinline__ bool wl_word_malformed(u64 wl, u32 kw) { return ((wl >> 40) & 0xFFFFull) || (wl >> 60) || (kw == 0 && wl != 0); }

#ifdef ZKIR_PROFILE_FULL
// ---- full profile (docs/PROVER_SPEC.md 3.6-3.8): one thread per row writes all 248 columns.  The per-row functions are air_pack.h's
// (shared with the host packer, bit-identical); the six lookup multiplicity columns are per-block shared-memory histograms flushed with
// atomics.  Bound: HBM writes (992 B/row).
struct FullWriter {
  u32* col; u64 N;
  __device__ __forceinline__ void operator()(int c, u32 v) { col[(u64)c * N] = v; }
};
#define FH_RNG 0
#define FH_AND 1024
#define FH_POW 2048
#define FH_B8 2176
#define FH_B4 2432
#define FH_B7 2448
#define FH_SIZE 2576
__global__ void __launch_bounds__(128) trace_expand_full_kernel(ExpandFullArgs fa);
// one row of the full table + its histogram contributions: shared by the rows converter and the write-log converter
__device__ __forceinline__ void expand_row_full(u64 i, u64 N, u64 T, const u64 (&rg)[16], u64 pc, u32 w, u64 read_val, u64 old_word, u32 prev_ts,
                                                u32 pre_err, u32* cols, u64* errp, u32 n_code, u32* hist, bool check_log = false, u64 wl = 0) {
  const bool live = i < T;
  FullWriter W = {cols + i, N};
  u32 err = pre_err;
  MemAccess ma;
  if (live && mem_decode(w, rg, ma)) {
    u64 loaded, new_word;
    const u32 merr = expand_mem_cells(i, w, rg, ma, old_word, (u64)prev_ts, W, &loaded, &new_word);
    if (!err) err = merr;
    if (ma.is_ld) {
      read_val = loaded;
      // write-log path: the register rebuild of the rows below used the LOGGED value of this load; it must be the one the memory log gives
      const u32 rd = (w >> 7) & 15;
      if (check_log && !err && rd && wl != (((u64)rd << 56) | loaded)) err = PACK_ERR_MEMVAL;
    }
  }
  const u32 rerr = expand_row_v2(i, T, rg, pc, w, read_val, W);
  if (!err) err = rerr;
  const u32 slot = (u32)((pc - 0x1000) >> 2);
  const bool in_rom = live && pc >= 0x1000 && !(pc & 3) && slot < n_code;
  if (live && !in_rom && !err) err = PACK_ERR_ROM;
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, in_rom ? slot : 0xffffffffu);
  if (in_rom && (u32)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(cols + (u64)ZKIR_COL_M_ROM * N + slot, (u32)__popc(peers));
  if (live && !err) {
    const u32* mine = cols + i;   // this thread's own stores above: visible to itself
    auto cell = [&](int c) { return mine[(u64)c * N]; };
    if (row_range_checked(w, rg[10])) for (int k = 0; k < 4; k++) atomicAdd(&hist[FH_RNG + (cell(ZKIR_COL_CH0 + k) & 1023u)], 1u);
    full_row_multiplicities(w, cell, [&](int t, u32 v) {
      const u32 base = t == MT_RNG ? FH_RNG : t == MT_AND ? FH_AND : t == MT_POW ? FH_POW : t == MT_B8 ? FH_B8 : t == MT_B4 ? FH_B4 : FH_B7;
      const u32 mask = t == MT_RNG || t == MT_AND ? 1023u : t == MT_POW || t == MT_B7 ? 127u : t == MT_B8 ? 255u : 15u;
      atomicAdd(&hist[base + (v & mask)], 1u);   // a value outside its table fails the lookup balance later; the mask only keeps the bin in range
    });
  }
  if (err) atomicMin(reinterpret_cast<unsigned long long*>(errp), (unsigned long long)((i << 8) | err));
}
__device__ __forceinline__ void full_hist_flush(const u32* hist, u32* cols, u64 N) {
  __syncthreads();
  const int tcol[6] = {ZKIR_COL_M_RNG, ZKIR_COL_M_AND, ZKIR_COL_M_POW, ZKIR_COL_M_B8, ZKIR_COL_M_B4, ZKIR_COL_M_B7};
  const u32 tbase[7] = {FH_RNG, FH_AND, FH_POW, FH_B8, FH_B4, FH_B7, FH_SIZE};
  for (u32 k = threadIdx.x; k < FH_SIZE; k += blockDim.x) {
    const u32 v = hist[k];
    if (!v) continue;
    int t = 0;
    while (k >= tbase[t + 1]) t++;
    atomicAdd(cols + (u64)tcol[t] * N + (k - tbase[t]), v);
  }
}
__global__ void __launch_bounds__(WL_CHUNK) trace_expand_wl_full_kernel(WlFullArgs a) {
  __shared__ int warp_tot[WL_CHUNK / 32][16];
  __shared__ u32 hist[FH_SIZE];
  for (u32 k = threadIdx.x; k < FH_SIZE; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  const u64 i = (u64)blockIdx.x * WL_CHUNK + threadIdx.x;
  const u64 M40 = (1ull << 40) - 1;
  u64 rg[16], wl; u32 kw;
  wl_rebuild_registers(a.wlog, a.T, a.chunk_prev, warp_tot, rg, wl, kw);
  if (i < a.N) {
    const bool live = i < a.T;
    u32 pre_err = PACK_OK;
    if (live && wl_word_malformed(wl, kw)) pre_err = ((a.ins[i] & 0x7F) == 0x50 && rg[10] == 1) ? PACK_ERR_TAPE40 : PACK_ERR_REG40;
    const u64 read_val = kw == 10u ? (wl & M40) : rg[10];   // READ rows: the post-state r10 (loads take theirs from the logged word)
    expand_row_full(i, a.N, a.T, rg, live ? (u64)a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, live ? a.old_word[i] : 0ull, live ? a.prev_ts[i] : 0u,
                    pre_err, a.cols, a.err, a.n_code, hist, true, wl);
  }
  full_hist_flush(hist, a.cols, a.N);
}
__global__ void __launch_bounds__(128) trace_expand_full_kernel(ExpandFullArgs fa) {
  __shared__ u32 hist[FH_SIZE];
  const ExpandArgs& a = fa.rows;
  for (u32 k = threadIdx.x; k < FH_SIZE; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < a.N) {
    const bool live = i < a.T;
    u64 rg[16];
#pragma unroll
    for (int k = 0; k < 16; k++) rg[k] = live ? a.regs[16 * i + k] : a.final_regs[k];
    const u64 read_val = (i + 1 < a.T) ? a.regs[16 * (i + 1) + 10] : a.final_regs[10];   // READ rows: post-state r10
    expand_row_full(i, a.N, a.T, rg, live ? a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, live ? fa.old_word[i] : 0ull, live ? fa.prev_ts[i] : 0u,
                    PACK_OK, a.cols, a.err, a.n_code, hist);
  }
  full_hist_flush(hist, a.cols, a.N);
}
static int clear_full_columns(u32* cols, u64 N, cudaStream_t st) {
  // columns no row writes in full (histograms, memory cells of non-memory rows, boundary cells, padding columns): zero first
  if (cudaMemsetAsync(cols + (u64)ZKIR_COL_M_RNG * N, 0, 2 * N * sizeof(u32), st) != cudaSuccess) return -2;
  if (cudaMemsetAsync(cols + (u64)ZKIR_COL_M_AND * N, 0, (u64)(ZKIR_COL_COUNT - ZKIR_COL_M_AND) * N * sizeof(u32), st) != cudaSuccess) return -2;
  return 0;
}
int launch_trace_expand_wl_full(const WlFullArgs& a, cudaStream_t st, u64* launches) {
  if (clear_full_columns(a.cols, a.N, st)) return -2;
  if (launch_wl_prefix(a.wlog, a.T, a.N, a.chunk_prev, st, launches)) return -2;
  trace_expand_wl_full_kernel<<<(unsigned)((a.N + WL_CHUNK - 1) / WL_CHUNK), WL_CHUNK, 0, st>>>(a);
  (*launches)++;
  { const cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "[zkir_b200] full write-log converter launch: %s\n", cudaGetErrorString(e)); return -2; } }
  return 0;
}
__global__ void add_u32_kernel(u32* dst, const u32* __restrict__ src, u64 n) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < n && src[i]) dst[i] += src[i];
}
int launch_add_u32(u32* dst, const u32* src, u64 n, cudaStream_t st, u64* launches) {
  add_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, src, n);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int launch_trace_expand_full(const ExpandFullArgs& fa, cudaStream_t st, u64* launches) {
  const ExpandArgs& a = fa.rows;
  if (clear_full_columns(a.cols, a.N, st)) return -2;
  trace_expand_full_kernel<<<(unsigned)((a.N + 127) / 128), 128, 0, st>>>(fa);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else   // core profile
struct ColWriter {
  u32* col; u64 N; u32 lo, hi;
  u32 ch[4];   // the four chunk values of the row, kept for the range histogram
  __device__ __forceinline__ void operator()(int c, u32 v) {
    if ((u32)c >= lo && (u32)c < hi) col[(u64)c * N] = v;
    if (c >= ZKIR_COL_CH0 && c < ZKIR_COL_CH0 + 4) ch[c - ZKIR_COL_CH0] = v;
  }
};

// One row + its contribution to the two multiplicity histograms.  hist = this block's shared range histogram (1024 bins).
__device__ __forceinline__ void expand_row(u64 i, u64 N, u64 T, const u64 (&rg)[16], u64 pc, u32 w, u64 read_val, u32* cols, u64* errp,
                                           u32 col_lo, u32 col_hi, u32 n_code, u32* hist) {
  ColWriter W = {cols + i, N, col_lo, col_hi, {0, 0, 0, 0}};
  u32 err = expand_row_v2(i, T, rg, pc, w, read_val, W);
  const bool live = i < T;
  // ROM multiplicity: one count for the program row of this pc; lanes of a warp that share a pc add once
  const u32 slot = (u32)((pc - 0x1000) >> 2);
  const bool in_rom = live && pc >= 0x1000 && !(pc & 3) && slot < n_code;
  if (live && !in_rom && !err) err = PACK_ERR_ROM;
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, in_rom ? slot : 0xffffffffu);
  if (in_rom && (u32)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(cols + (u64)ZKIR_COL_M_ROM * N + slot, (u32)__popc(peers));
  if (live && !err && row_range_checked(w, rg[10])) {
#pragma unroll
    for (int k = 0; k < 4; k++) atomicAdd(&hist[W.ch[k] & 1023u], 1u);
  }
  if (err) {  // first offending row wins; the host reports it after the stream is drained
    const unsigned long long packed = (i << 8) | err;
    atomicMin(reinterpret_cast<unsigned long long*>(errp), packed);
  }
}
__device__ __forceinline__ void hist_begin(u32* hist) {
  for (u32 k = threadIdx.x; k < 1024; k += blockDim.x) hist[k] = 0;
  __syncthreads();
}
__device__ __forceinline__ void hist_flush(const u32* hist, u32* m_rng) {
  __syncthreads();
  for (u32 k = threadIdx.x; k < 1024; k += blockDim.x) { const u32 v = hist[k]; if (v) atomicAdd(m_rng + k, v); }
}


__global__ void __launch_bounds__(128) trace_expand_kernel(ExpandArgs a) {
  __shared__ u32 hist[1024];
  hist_begin(hist);
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < a.N) {
    const bool live = i < a.T;
    u64 rg[16];
#pragma unroll
    for (int k = 0; k < 16; k++) rg[k] = live ? a.regs[16 * i + k] : a.final_regs[k];
    const u64 read_val = (i + 1 < a.T) ? a.regs[16 * (i + 1) + 10] : a.final_regs[10];
    expand_row(i, a.N, a.T, rg, live ? a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, a.cols, a.err, a.col_lo, a.col_hi, a.n_code, hist);
  }
  hist_flush(hist, a.cols + (u64)ZKIR_COL_M_RNG * a.N);
}

// ---- write-log input, core profile: passes 1 and 2 of the last-writer scan (see wl_rebuild_registers above), then the converter
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
  __shared__ u32 hist[1024];
  hist_begin(hist);
  const u64 i = (u64)blockIdx.x * WL_CHUNK + threadIdx.x;
  const u64 M40 = (1ull << 40) - 1;
  u64 rg[16], wl; u32 kw;
  wl_rebuild_registers(a.wlog, a.T, a.chunk_prev, warp_tot, rg, wl, kw);
  if (i < a.N) {
    const bool live = i < a.T;
    if (live && wl_word_malformed(wl, kw)) {
      const bool read_row = (a.ins[i] & 0x7F) == 0x50 && rg[10] == 1;
      atomicMin(reinterpret_cast<unsigned long long*>(a.err), (unsigned long long)((i << 8) | (read_row ? PACK_ERR_TAPE40 : PACK_ERR_REG40)));
    }
    // READ rows need the post-state r10: the logged value if the row changed r10, else the unchanged pre-state
    const u64 read_val = kw == 10u ? (wl & M40) : rg[10];
    expand_row(i, a.N, a.T, rg, live ? (u64)a.pcs[i] : a.final_pc, live ? a.ins[i] : 0u, read_val, a.cols, a.err, a.col_lo, a.col_hi, a.n_code, hist);
  }
  hist_flush(hist, a.cols + (u64)ZKIR_COL_M_RNG * a.N);
}

// the multiplicity columns are accumulated with atomics: clear them first (all other columns are fully written by the rows)
static int clear_multiplicities(u32* cols, u64 N, cudaStream_t st) {
  return cudaMemsetAsync(cols + (u64)ZKIR_COL_M_RNG * N, 0, 2 * N * sizeof(u32), st) == cudaSuccess ? 0 : -2;
}
int launch_trace_expand(const ExpandArgs& a, cudaStream_t st, u64* launches) {
  static_assert(ZKIR_COL_M_ROM == ZKIR_COL_M_RNG + 1, "multiplicity columns are adjacent");
  if (clear_multiplicities(a.cols, a.N, st)) return -2;
  trace_expand_kernel<<<(unsigned)((a.N + 127) / 128), 128, 0, st>>>(a);
  (*launches)++;
  { const cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "[zkir_b200] row converter launch (N=%llu T=%llu cols=%p): %s\n", (unsigned long long)a.N, (unsigned long long)a.T, (void*)a.cols, cudaGetErrorString(e)); return -2; } }
  return 0;
}
u64 trace_expand_wl_scratch_ints(u64 N) { return ((N + WL_CHUNK - 1) / WL_CHUNK) * 16; }
// passes 1 and 2 of the last-writer scan alone (the full profile's converter runs pass 3 in its own kernel)
int launch_wl_prefix(const u64* wlog, u64 T, u64 N, int* chunk_prev, cudaStream_t st, u64* launches) {
  const u32 n_chunks = (u32)((N + WL_CHUNK - 1) / WL_CHUNK);
  wl_chunk_last_kernel<<<n_chunks, WL_CHUNK, 0, st>>>(wlog, T, chunk_prev);
  wl_chunk_scan_kernel<<<1, WL_SCAN_THREADS, 0, st>>>(chunk_prev, n_chunks);
  (*launches) += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int launch_trace_expand_wl(const WlArgs& a, cudaStream_t st, u64* launches) {
  const u32 n_chunks = (u32)((a.N + WL_CHUNK - 1) / WL_CHUNK);
  if (clear_multiplicities(a.cols, a.N, st)) return -2;
  wl_chunk_last_kernel<<<n_chunks, WL_CHUNK, 0, st>>>(a.wlog, a.T, a.chunk_prev);
  wl_chunk_scan_kernel<<<1, WL_SCAN_THREADS, 0, st>>>(a.chunk_prev, n_chunks);
  trace_expand_wl_kernel<<<n_chunks, WL_CHUNK, 0, st>>>(a);
  (*launches) += 3;
  { const cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "[zkir_b200] write-log converter launch (N=%llu T=%llu chunks=%u): %s\n", (unsigned long long)a.N, (unsigned long long)a.T, n_chunks, cudaGetErrorString(e)); return -2; } }
  return 0;
}

#endif  // ZKIR_PROFILE_FULL

}  // namespace zkir
