// LogUp aux columns of the AIR v2 (docs/PROVER_SPEC.md section 3.4): for every trace row the seven lookup fractions
// n_j / (z - fingerprint_j) of air_generated.h (zkir_air_fractions: four 10-bit range lookups of the row's chunks, the range-table
// row, the ROM lookup of (pc, decoded word, imm), the ROM-table row), summed pairwise into three ext4 helper columns, and the
// running sum phi of all fractions of the earlier rows.  The reference only counts range-check witnesses on the host
// (zkir-runtime/src/range_check.rs:111-192); the argument that they are all in the table is this build's.
//
// Kernel 1: one thread per row, 7 ext4 inversions (tower inversion, bb.cuh).  Bound: integer-multiply pipe (about 900 modular
// multiplies per row); reads 44 of the 92 canonical input columns once, writes 12 columns + one ext4 per row.
// Kernels 2-4: exclusive prefix sum of the per-row ext4 totals (4 independent modular scans): block totals, one-block scan of
// the totals, block-local scan + offset.  Bound: HBM, 16 B read + 16 B written per row.
#include <cuda_runtime.h>
#include "bb.cuh"
#include "kernels.h"
#include "air_profile.h"

namespace zkir {

#ifdef ZKIR_PROFILE_FULL
__device__ __noinline__ E4 e4_inv_call(E4 a) { return e4_inv(a); }   // 81 inlined tower inversions would be most of the kernel's code
// Full profile (81 fractions, 41 helpers): the generated code builds one fraction at a time inside its own block (tools/gen_air.py:
// build_fractions), and this context consumes it at once -- invert, add to the helper it belongs to, store a helper as soon as the next
// one begins (the fractions of a helper are consecutive) -- so neither the fractions nor the helpers are ever all live.
struct FracCtxDev {
  typedef Fm F; typedef Xm X;
  const u32 *trace, *pub; u64 N, row;
  const E4* lc;   // shared: z, theta .. theta^NUM_THETA
  u32* aux;       // out: helper columns
  E4 cur, tot;    // the helper being summed; sum of everything (helpers and the fractions phi adds itself)
  int cur_k;      // index of that helper (the next one to store)
  __device__ __forceinline__ Fm L(int i) const { return Fm(bb_to_mont(__ldg(trace + (u64)i * N + row))); }
  __device__ __forceinline__ Fm P(int i) const { return Fm(bb_to_mont(__ldg(pub + (u64)i * N + row))); }
  __device__ __forceinline__ Fm K(u32 k) const { return Fm(bb_to_mont_c(k)); }
  __device__ __forceinline__ Xm z() const { Xm r; r.v = lc[0]; return r; }
  __device__ __forceinline__ Xm th(int k) const { Xm r; r.v = lc[k]; return r; }
  __device__ __forceinline__ Xm xf(Fm a) const { Xm r; r.v = e4_from_base(a.v); return r; }
  __device__ __forceinline__ void store_upto(int k) {   // helpers cur_k .. k-1 are complete: cur, then zeros
    for (; cur_k < k; cur_k++) {
#pragma unroll
      for (int q = 0; q < 4; q++) aux[(u64)(4 * cur_k + q) * N + row] = bb_from_mont(cur.c[q]);
      tot = e4_add(tot, cur);
      cur = e4_zero();
    }
  }
  __device__ __forceinline__ void frac(int j, Fm n, Xm d) {
    const int helper_of[ZKIR_AIR_NUM_FRACTIONS] = ZKIR_AIR_FRAC_HELPER_INIT;
    const int k = helper_of[j];   // j is a literal in the generated code
    if (k != ZKIR_AIR_NUM_HELPERS) store_upto(k);
    if (n.v == 0) return;         // lookup switched off on this row, table row never hit: no inversion
    const E4 v = e4_mulb(e4_inv_call(d.v), n.v);
    if (k == ZKIR_AIR_NUM_HELPERS) tot = e4_add(tot, v); else cur = e4_add(cur, v);
  }
};
#else
struct FracCtxDev {
  typedef Fm F; typedef Xm X;
  const u32 *trace, *pub; u64 N, row;
  const E4* lc;   // shared: z, theta .. theta^NUM_THETA
  u32 num[ZKIR_AIR_NUM_FRACTIONS]; E4 den[ZKIR_AIR_NUM_FRACTIONS];
  __device__ __forceinline__ Fm L(int i) const { return Fm(bb_to_mont(__ldg(trace + (u64)i * N + row))); }
  __device__ __forceinline__ Fm P(int i) const { return Fm(bb_to_mont(__ldg(pub + (u64)i * N + row))); }
  __device__ __forceinline__ Fm K(u32 k) const { return Fm(bb_to_mont_c(k)); }
  __device__ __forceinline__ Xm z() const { Xm r; r.v = lc[0]; return r; }
  __device__ __forceinline__ Xm th(int k) const { Xm r; r.v = lc[k]; return r; }
  __device__ __forceinline__ Xm xf(Fm a) const { Xm r; r.v = e4_from_base(a.v); return r; }
  __device__ __forceinline__ void frac(int j, Fm n, Xm d) { num[j] = n.v; den[j] = d.v; }
};
#endif

#define AUX_ROWS_THREADS 128
__global__ void __launch_bounds__(AUX_ROWS_THREADS) ZKIR_PF(aux_rows_kernel)(AuxArgs a) {
  __shared__ E4 lc[ZKIR_AIR_NUM_THETA + 1];
  if (threadIdx.x == 0) {
    E4 z, th;
    for (int k = 0; k < 4; k++) { z.c[k] = a.lookup[k]; th.c[k] = a.lookup[4 + k]; }
    lc[0] = z; lc[1] = th;
    for (int k = 2; k <= ZKIR_AIR_NUM_THETA; k++) lc[k] = e4_mul(lc[k - 1], th);
  }
  __syncthreads();
  const u64 N = 1ull << a.log_n;
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= N) return;
  FracCtxDev c;
  c.trace = a.trace; c.pub = a.pub; c.N = N; c.row = i; c.lc = lc;
#ifdef ZKIR_PROFILE_FULL
  c.aux = a.aux; c.cur = e4_zero(); c.tot = e4_zero(); c.cur_k = 0;
  zkir_air_fractions(c);
  c.store_upto(ZKIR_AIR_NUM_HELPERS);
  a.row_tot[i] = c.tot;
#else
  zkir_air_fractions(c);
  const int helper_of[ZKIR_AIR_NUM_FRACTIONS] = ZKIR_AIR_FRAC_HELPER_INIT;
  E4 h[ZKIR_AIR_NUM_HELPERS + 1];   // the helpers, then the fractions the running sum adds itself
#pragma unroll
  for (int k = 0; k <= ZKIR_AIR_NUM_HELPERS; k++) h[k] = e4_zero();
#pragma unroll
  for (int j = 0; j < ZKIR_AIR_NUM_FRACTIONS; j++) {
    // a zero numerator (lookup switched off on this row, table row never hit) needs no inversion
    if (c.num[j] != 0) h[helper_of[j]] = e4_add(h[helper_of[j]], e4_mulb(e4_inv(c.den[j]), c.num[j]));
  }
  E4 tot = h[ZKIR_AIR_NUM_HELPERS];
#pragma unroll
  for (int k = 0; k < ZKIR_AIR_NUM_HELPERS; k++) {
#pragma unroll
    for (int q = 0; q < 4; q++) a.aux[(u64)(4 * k + q) * N + i] = bb_from_mont(h[k].c[q]);
    tot = e4_add(tot, h[k]);
  }
  a.row_tot[i] = tot;
#endif
}

// ---- exclusive prefix sum of row_tot (ext4 = 4 independent sums mod p)
#define SCAN_THREADS 256
#define SCAN_PER 4
#define SCAN_BLOCK (SCAN_THREADS * SCAN_PER)
#ifndef ZKIR_PROFILE_FULL   // the scan and the I/O sum do not depend on the profile: compiled with the core build only
u64 aux_gen_blocks(u64 N) { return (N + SCAN_BLOCK - 1) / SCAN_BLOCK; }

__device__ __forceinline__ E4 block_reduce(E4 v, E4* sh) {   // sum over the block, result valid in thread 0
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; k++) for (int o = 16; o > 0; o >>= 1) v.c[k] = bb_add(v.c[k], __shfl_xor_sync(0xffffffffu, v.c[k], o));
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  E4 r = e4_zero();
  if (threadIdx.x == 0) for (u32 w = 0; w < blockDim.x / 32; w++) r = e4_add(r, sh[w]);
  return r;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(const E4* __restrict__ row_tot, E4* __restrict__ blk_tot, u64 N) {
  __shared__ E4 sh[SCAN_THREADS / 32];
  const u64 base = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * SCAN_PER;
  E4 s = e4_zero();
#pragma unroll
  for (int k = 0; k < SCAN_PER; k++) if (base + k < N) s = e4_add(s, row_tot[base + k]);
  s = block_reduce(s, sh);
  if (threadIdx.x == 0) blk_tot[blockIdx.x] = s;
}
// one block: blk_tot[b] becomes the sum of the totals of the blocks before b; flags an unbalanced grand total
__global__ void __launch_bounds__(1024) scan_blocks_kernel(E4* blk_tot, u64 n_blocks, u64 N, const u32* sio, u64* err) {
  __shared__ E4 part[1024];
  const u32 t = threadIdx.x;
  const u64 per = (n_blocks + 1023) / 1024, b0 = t * per, b1 = b0 + per < n_blocks ? b0 + per : n_blocks;
  E4 s = e4_zero();
  for (u64 b = b0; b < b1; b++) s = e4_add(s, blk_tot[b]);
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    E4 run = e4_zero();
    for (u32 u = 0; u < 1024; u++) { const E4 v = part[u]; part[u] = run; run = e4_add(run, v); }
    // the range and ROM fractions of all rows must cancel and the I/O rows must add up to the public transcript's sum: otherwise the
    // witness is not a valid lookup (a chunk outside the table, an instruction outside the program, another I/O transcript)
    if (err && ((run.c[0] ^ sio[0]) | (run.c[1] ^ sio[1]) | (run.c[2] ^ sio[2]) | (run.c[3] ^ sio[3])))
      atomicMin(reinterpret_cast<unsigned long long*>(err), (unsigned long long)(((N - 1) << 8) | 8u));
  }
  __syncthreads();
  E4 run = part[t];
  for (u64 b = b0; b < b1; b++) { const E4 v = blk_tot[b]; blk_tot[b] = run; run = e4_add(run, v); }
}
// aux = the four base columns of phi (the last aux columns of the profile)
__global__ void __launch_bounds__(SCAN_THREADS) scan_write_kernel(const E4* __restrict__ row_tot, const E4* __restrict__ blk_tot, u32* __restrict__ aux, u64 N) {
  __shared__ E4 wsum[SCAN_THREADS / 32];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 base = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * SCAN_PER;
  E4 v[SCAN_PER], s = e4_zero();
#pragma unroll
  for (int k = 0; k < SCAN_PER; k++) { v[k] = base + k < N ? row_tot[base + k] : e4_zero(); s = e4_add(s, v[k]); }
  // inclusive scan of the per-thread sums inside the warp, then across warps
  E4 inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 4; k++) { const u32 up = __shfl_up_sync(0xffffffffu, inc.c[k], o); if (lane >= (u32)o) inc.c[k] = bb_add(inc.c[k], up); }
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  E4 off = blk_tot[blockIdx.x];
  for (u32 w = 0; w < warp; w++) off = e4_add(off, wsum[w]);
  E4 run = e4_add(off, e4_sub(inc, s));   // exclusive prefix of this thread's first row
#pragma unroll
  for (int k = 0; k < SCAN_PER; k++) {
    if (base + k < N) {
#pragma unroll
      for (int q = 0; q < 4; q++) aux[(u64)q * N + base + k] = bb_from_mont(run.c[q]);
    }
    run = e4_add(run, v[k]);
  }
}

// S_io of the public I/O transcript: one block, a thread per event (strided), ext4 inversion each, block reduction
__global__ void __launch_bounds__(256) io_sum_kernel(const u32* __restrict__ ev, u32 n, u32* lookup) {
  __shared__ E4 sh[8];
  E4 z, th[5];
  for (int k = 0; k < 4; k++) { z.c[k] = lookup[k]; th[1].c[k] = lookup[4 + k]; }
  for (int k = 2; k < 5; k++) th[k] = e4_mul(th[k - 1], th[1]);
  E4 s = e4_zero();
  for (u32 e = threadIdx.x; e < n; e += blockDim.x) {
    E4 fp = e4_from_base(bb_to_mont_c(3u));
#pragma unroll
    for (int k = 0; k < 4; k++) fp = e4_add(fp, e4_mulb(th[k + 1], bb_to_mont(ev[4 * e + k] % BB_P)));
    s = e4_add(s, e4_inv(e4_sub(z, fp)));
  }
  s = block_reduce(s, sh);
  if (threadIdx.x == 0) for (int k = 0; k < 4; k++) lookup[8 + k] = s.c[k];
}
int launch_io_sum(const u32* events, u32 n_events, u32* lookup, cudaStream_t st, u64* launches) {
  io_sum_kernel<<<1, 256, 0, st>>>(events, n_events, lookup);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// running sum phi of the per-row totals into the four columns at phi_cols
int launch_aux_scan(const AuxArgs& a, u32* phi_cols, cudaStream_t st, u64* launches) {
  const u64 N = 1ull << a.log_n, nb = aux_gen_blocks(N);
  scan_totals_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(a.row_tot, a.blk_tot, N);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(a.blk_tot, nb, N, a.lookup + 8, a.err);
  scan_write_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(a.row_tot, a.blk_tot, phi_cols, N);
  (*launches) += 3;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#endif  // !ZKIR_PROFILE_FULL

int ZKIR_PF(launch_aux_gen)(const AuxArgs& a, cudaStream_t st, u64* launches) {
  const u64 N = 1ull << a.log_n;
  ZKIR_PF(aux_rows_kernel)<<<(unsigned)((N + AUX_ROWS_THREADS - 1) / AUX_ROWS_THREADS), AUX_ROWS_THREADS, 0, st>>>(a);
  (*launches)++;
  if (cudaGetLastError() != cudaSuccess) return -2;
  return launch_aux_scan(a, a.aux + (u64)4 * ZKIR_AIR_NUM_HELPERS * N, st, launches);
}

}  // namespace zkir
