// AIR quotient sweep: one thread per row of the LDE coset evaluates every constraint of the AIR v2 (main, aux and public columns)
// (air_generated.h, emitted by tools/gen_air.py -- the same list the CPU oracle and the verifier instantiate),
// folds them with powers of alpha in ext4 and divides by the vanishing polynomial.  Column-major LDE so adjacent
// threads read adjacent addresses; the LDE is coset-major, so "next row" (g*x) is simply the next memory row of the same coset.
//
// The reference has no constraint system; the transition semantics encoded are those of
// zkir-runtime/src/execute.rs and syscall.rs as cited in tools/gen_air.py.
// Bound: HBM (reads 2 x 4*M*W bytes unless the +blowup row hits L2, writes 16*M).
#include <cuda_runtime.h>
#include <stdlib.h>
#include "bb.cuh"
#include "kernels.h"
#include "air_profile.h"
#include "constants_generated.h"

namespace zkir {

struct QCtx {
  typedef Fm F; typedef Xm X;
  const u32 *lde, *aux, *pub; u64 M, row, nxt;
  const u32* pv;       // shared
  const E4* apow;      // shared, apow[i] = alpha^(K-1-i)
  const E4* lc;        // shared: lookup challenges z, theta .. theta^NUM_THETA, then the public I/O transcript's sum
  Fm is_first, is_last, is_trans;
  Acc4 acc;  // lazy 64-bit accumulator of sum_i alpha^(K-1-i) * C_i over the base-field constraints (bb.cuh)
  E4 accx;   // the ext4-valued constraints (LogUp) are few: plain ext4 products
  __device__ __forceinline__ Fm L(int i) const { return Fm(__ldg(lde + (u64)i * M + row)); }
  __device__ __forceinline__ Fm N(int i) const { return Fm(__ldg(lde + (u64)i * M + nxt)); }
  __device__ __forceinline__ Fm A(int i) const { return Fm(__ldg(aux + (u64)i * M + row)); }
  __device__ __forceinline__ Fm AN(int i) const { return Fm(__ldg(aux + (u64)i * M + nxt)); }
  __device__ __forceinline__ Fm P(int i) const { return Fm(__ldg(pub + (u64)i * M + row)); }
  __device__ __forceinline__ Fm PV(int i) const { return Fm(pv[i]); }
  __device__ __forceinline__ Fm K(u32 k) const { return Fm(bb_to_mont_c(k)); }
  __device__ __forceinline__ Xm z() const { Xm r; r.v = lc[0]; return r; }
  __device__ __forceinline__ Xm th(int k) const { Xm r; r.v = lc[k]; return r; }
  __device__ __forceinline__ Xm xf(Fm a) const { Xm r; r.v = e4_from_base(a.v); return r; }
  __device__ __forceinline__ Xm sio() const { Xm r; r.v = lc[ZKIR_AIR_NUM_THETA + 1]; return r; }
  __device__ __forceinline__ Xm x4(Fm a, Fm b, Fm c, Fm d) const { Xm r; r.v.c[0] = a.v; r.v.c[1] = b.v; r.v.c[2] = c.v; r.v.c[3] = d.v; return r; }
  // Full profile only: the constraint list is ~38 000 straight-line instructions (600 KB of code, far beyond the instruction cache).
  // The block re-converges at every fence, so its warps walk the code together and one instruction fetch serves all of them.
  __device__ __forceinline__ void fence() const { __syncthreads(); }
  __device__ __forceinline__ void emit(int idx, Fm v) {
    acc4_mac(acc, apow[idx], v.v);
    if (idx & 1) acc4_fix(acc);  // idx is a literal in the generated code: at most two products between fixes
  }
  __device__ __forceinline__ void emit_x(int idx, Xm v) { accx = e4_add(accx, e4_mul(apow[idx], v.v)); }
};

__global__ void ZKIR_PF(alpha_powers_kernel)(const u32* alpha, E4* apow) {  // apow[i] = alpha^(K-1-i); thread per power
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ZKIR_AIR_NUM_CONSTRAINTS) return;
  E4 a; for (int k = 0; k < 4; k++) a.c[k] = alpha[k];
  E4 r = e4_one();
  for (u32 e = ZKIR_AIR_NUM_CONSTRAINTS - 1 - i; e; e >>= 1) { if (e & 1) r = e4_mul(r, a); a = e4_mul(a, a); }
  apow[i] = r;
}

#ifdef ZKIR_PROFILE_FULL
#define QUOTIENT_THREADS 256
#else
#define QUOTIENT_THREADS 128
#endif
template <int MINB, int THREADS = QUOTIENT_THREADS>
__global__ void __launch_bounds__(THREADS, MINB) ZKIR_PF(quotient_kernel)(QuotientArgs a, const E4* apow_g, u32 g_inv, u32 g, u32 snn /* shift^N */, u32 wb /* w_B */) {
  __shared__ E4 apow[ZKIR_AIR_NUM_CONSTRAINTS];
  __shared__ E4 lc[ZKIR_AIR_NUM_THETA + 2];
  __shared__ u32 pv[ZKIR_AIR_NUM_PUBLIC];
  for (int i = threadIdx.x; i < ZKIR_AIR_NUM_CONSTRAINTS; i += blockDim.x) apow[i] = apow_g[i];
  if (threadIdx.x < ZKIR_AIR_NUM_PUBLIC) pv[threadIdx.x] = a.pv[threadIdx.x];
  // Z_H(x) = x^N - 1 = shift^N * w_B^z - 1 takes only B = 2^log_blowup (<= 16) values on the coset: one inversion per block and coset
  __shared__ u32 zh_s[16], zhi_s[16];
  if (threadIdx.x >= 64 && threadIdx.x < 64 + (1u << a.log_blowup)) {
    const u32 zz = threadIdx.x - 64;
    const u32 v = bb_sub(bb_mul(snn, bb_pow(wb, zz)), BB_ONE);
    zh_s[zz] = v; zhi_s[zz] = bb_inv(v);
  }
  if (threadIdx.x == 32) {
    E4 z, th, so;
    for (int k = 0; k < 4; k++) { z.c[k] = a.lookup[k]; th.c[k] = a.lookup[4 + k]; so.c[k] = a.lookup[8 + k]; }
    lc[0] = z; lc[1] = th;
    for (int k = 2; k <= ZKIR_AIR_NUM_THETA; k++) lc[k] = e4_mul(lc[k - 1], th);
    lc[ZKIR_AIR_NUM_THETA + 1] = so;
  }
  __syncthreads();
  const u64 M = 1ull << (a.log_n + a.log_blowup), N = 1ull << a.log_n;
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;  // memory row: coset z = i / N, point j = i % N, natural index j*B + z
  const u32 log_nj = a.seg_log_nj == 0xffffffffu ? a.log_n : a.seg_log_nj;
  if (t >= (1ull << (log_nj + a.log_blowup))) return;   // never inside a block: the row count is a power of two >= 2^11
  const u64 z = t >> log_nj, j = a.seg_j0 + (t & ((1ull << log_nj) - 1));
  const u64 i = (z << a.log_n) | j;
  QCtx c;
  c.lde = a.lde; c.aux = a.lde + (u64)ZKIR_AIR_WIDTH * M; c.pub = a.publde; c.M = M; c.row = i; c.nxt = (z << a.log_n) | ((j + 1) & (N - 1));
  c.pv = pv; c.apow = apow; c.lc = lc;
  const u32 x = a.xs[i];
  const u32 zh = zh_s[z];
  c.is_first = Fm(bb_mul(zh, a.dinv[i]));                         // Z_H/(x-1)
  c.is_last = Fm(bb_mul(zh, bb_mul(g, a.dinv[c.nxt])));  // Z_H/(x-g^-1) = Z_H*g/(g x-1), g*x_i = x_{i+B}
  c.is_trans = Fm(bb_sub(x, g_inv));
  c.acc = acc4_zero(); c.accx = e4_zero();
  zkir_air_eval(c);
  const u32 zi = zhi_s[z];
  const u64 nat = (j << a.log_blowup) | z;
  const E4 accv = e4_add(acc4_finish(c.acc), c.accx);
#pragma unroll
  for (int k = 0; k < 4; k++) (a.q_plane[k] ? a.q_plane[k] : a.q)[(u64)k * M + nat] = bb_mul(accv.c[k], zi);
}

static u32 hpow(u32 a, u64 e) { u64 r = 1, b = a; while (e) { if (e & 1) r = r * b % BB_P; b = b * b % BB_P; e >>= 1; } return (u32)r; }

#ifndef ZKIR_PROFILE_FULL   // profile-independent: compiled with the core build only
__global__ void domain_tables_kernel(u32* xs, u32* dinv, u64 M, u32 log_n, u32 log_b, u32 shift, u32 w) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const u64 nat = ((i & ((1ull << log_n) - 1)) << log_b) | (i >> log_n);
  u32 x = bb_mul(shift, bb_pow(w, nat));
  xs[i] = x;
  dinv[i] = bb_inv(bb_sub(x, BB_ONE));
}


int launch_domain_tables(u32* xs, u32* dinv, u32 log_n, u32 log_b, u32 shift_canon, cudaStream_t st, u64* launches) {
  const u32 log_m = log_n + log_b;
  u64 M = 1ull << log_m;
  domain_tables_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(xs, dinv, M, log_n, log_b, bb_to_mont_c(shift_canon), bb_to_mont_c(ZKIR_BB_ROOTS[log_m]));
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

#endif  // !ZKIR_PROFILE_FULL

int ZKIR_PF(launch_quotient)(const QuotientArgs& a, cudaStream_t st, u64* launches) {
  const u64 M = 1ull << (a.log_n + a.log_blowup);
  ZKIR_PF(alpha_powers_kernel)<<<(ZKIR_AIR_NUM_CONSTRAINTS + 63) / 64, 64, 0, st>>>(a.alpha, reinterpret_cast<E4*>(a.apow_scratch));
  const u32 g = ZKIR_BB_ROOTS[a.log_n];
  const u32 g_inv = hpow(g, BB_P - 2);
  const u32 snn = hpow(ZKIR_BB_GEN, 1ull << a.log_n);
  const u32 wb = ZKIR_BB_ROOTS[a.log_blowup];
  static int variant = -1;
  // AIR v2 (169 constraints, ext4 LogUp terms): 168 registers / 3 CTAs per SM measured fastest: 1.05 / 1.09 / 1.14 ms for variants 3 / 0 / 1 at 2^20 rows
  if (variant < 0) { const char* e = getenv("ZKIR_QUOTIENT_VARIANT"); variant = e ? atoi(e) : 3; }
  const u64 n_threads = a.seg_log_nj == 0xffffffffu ? M : (1ull << (a.seg_log_nj + a.log_blowup));
  const unsigned grid = (unsigned)((n_threads + QUOTIENT_THREADS - 1) / QUOTIENT_THREADS);
  const E4* ap = reinterpret_cast<const E4*>(a.apow_scratch);
  const u32 m_ginv = bb_to_mont_c(g_inv), m_g = bb_to_mont_c(g), m_snn = bb_to_mont_c(snn), m_wb = bb_to_mont_c(wb);
#ifdef ZKIR_PROFILE_FULL
  // full profile: blocks that re-converge at every fence (see QCtx::fence).  Measured at 2^18 rows (2^19 LDE rows), 256-thread blocks:
  // 1 / 2 / 3 resident blocks per SM (255 / 128 / 80 registers) = 3.56 / 2.16 / 1.88 ms: occupancy wins over spills.  Splitting the
  // constraint list over 2 / 4 / 8 launches of consecutive index ranges (no spills left in most parts) measured 3.67 / 2.81 / 2.39 ms:
  // every part repeats the operand fetch and selector sums that most constraints share, so one launch stays
  auto launch = [&](auto kernel, unsigned threads) { kernel<<<(unsigned)(n_threads / threads), threads, 0, st>>>(a, ap, m_ginv, m_g, m_snn, m_wb); };
  if (variant == 0) launch(ZKIR_PF(quotient_kernel)<2, 256>, 256);
  else if (variant == 1) launch(ZKIR_PF(quotient_kernel)<1, 256>, 256);
  else if (variant == 2) launch(ZKIR_PF(quotient_kernel)<4, 256>, 256);
  else if (variant == 4) launch(ZKIR_PF(quotient_kernel)<1, 512>, 512);
  else if (variant == 5) launch(ZKIR_PF(quotient_kernel)<1, 1024>, 1024);
  else if (variant == 6) launch(ZKIR_PF(quotient_kernel)<2, 512>, 512);
  else launch(ZKIR_PF(quotient_kernel)<3, 256>, 256);
#else
  if (variant == 3) ZKIR_PF(quotient_kernel)<3><<<grid, QUOTIENT_THREADS, 0, st>>>(a, ap, m_ginv, m_g, m_snn, m_wb);
  else if (variant == 1) ZKIR_PF(quotient_kernel)<6><<<grid, QUOTIENT_THREADS, 0, st>>>(a, ap, m_ginv, m_g, m_snn, m_wb);
  else if (variant == 2) ZKIR_PF(quotient_kernel)<8><<<grid, QUOTIENT_THREADS, 0, st>>>(a, ap, m_ginv, m_g, m_snn, m_wb);
  else ZKIR_PF(quotient_kernel)<4><<<grid, QUOTIENT_THREADS, 0, st>>>(a, ap, m_ginv, m_g, m_snn, m_wb);
#endif
  (*launches) += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace zkir
