// Fast batched BabyBear NTT / LDE for sm_100a: register radix-32 tiles, Shoup twiddles, no transposition passes.
//
// No reference counterpart exists (the reference has no NTT; its only field is Mersenne31,
// zkir-spec/src/field.rs:15-23).  Maths: docs/PROVER_SPEC.md section 4 step 1; design: DESIGN.md "NTT / LDE".
//
// A length-N transform (N = 2^log_n) is split into D "digits" of <= 10 bits (d_1 .. d_D, top to bottom in the
// memory position).  One kernel launch ("pass") transforms ONE digit for a batch of columns:
//   * a tile is R = 2^d points of that digit x 16 lanes along the contiguous axis; every thread keeps 2^ceil(d/2)
//     points in registers, runs a radix-2^a DIF with COMPILE-TIME twiddles (Shoup form: mul.hi + 2 mul.lo, the
//     integer-multiply pipe is the bound, see profiles/r01_microbench_b200.txt), multiplies by the inter-round
//     twiddle, exchanges once through shared memory (bank-conflict-free swizzles) and runs the second radix-2^b DIF;
//   * global loads/stores are 64 B (strided digits) or 128 B (contiguous digit) segments, tables are read with the
//     same addressing as the data, so they coalesce too.
// The inverse transform walks the digits top-down and leaves the coefficients in DIGIT-REVERSED order
// (coefficient k = x_1 + 2^d1 x_2 + ... sits at position ((x_1 2^d2 + x_2) 2^d3 + x_3 ...)); the forward transform
// walks bottom-up from that layout and produces natural order.  An LDE therefore needs no transposing pass and
// every pass runs in place on its destination buffer.  The LDE output is COSET-MAJOR: lde[col][z][i] is the value at
// x = shift * w_M^(i*B + z)  (natural LDE index idx = i*B + z), z < B = 2^log_blowup.
//
// Algorithmic bytes: 8*N*C per transform, 4*N*C*(1+B) per LDE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <map>
#include <utility>
#include <vector>
#include "bb.cuh"
#include "kernels.h"
#include "constants_generated.h"

namespace zkir {

// ---------------------------------------------------------------- compile-time field helpers (canonical values)
constexpr u32 C_G27 = 0x1a427a41u;  // generator of the 2^27 subgroup, 31^15
constexpr u32 c_root(int k) { return c_pow(C_G27, 1ull << (27 - k)); }  // == ZKIR_BB_ROOTS[k]
static_assert(c_root(1) == BB_P - 1 && c_root(2) == 1728404513u, "root chain must match ZKIR_BB_ROOTS");

// ---------------------------------------------------------------- register DIF, compile-time twiddles
template <int A, bool INV, int S, int I>
__device__ __forceinline__ void bfly(u32 (&x)[1 << A]) {
  constexpr int half = 1 << (A - 1 - S);
  if constexpr ((I & half) == 0) {
    constexpr int e = (I & (half - 1)) << S;  // exponent of w_{2^A}
    const u32 a = x[I], b = x[I + half];
    x[I] = bb_add(a, b);
    if constexpr (e == 0) {
      x[I + half] = bb_sub(a, b);
    } else {
      constexpr u32 root = INV ? c_pow(c_root(A), (1u << A) - 1) : c_root(A);
      constexpr u32 w = c_pow(root, e);
      constexpr u32 wq = c_shoup(w);
      x[I + half] = shoup_mul(a - b + BB_P, w, wq);
    }
  }
}
template <int A, bool INV, int S, int... I>
__device__ __forceinline__ void dif_stage(u32 (&x)[1 << A], std::integer_sequence<int, I...>) {
  (bfly<A, INV, S, I>(x), ...);
}
template <int A, bool INV, int... S>
__device__ __forceinline__ void dif_stages(u32 (&x)[1 << A], std::integer_sequence<int, S...>) {
  (dif_stage<A, INV, S>(x, std::make_integer_sequence<int, (1 << A)>{}), ...);
}
// in: x[j]; out: x[i] = X[bitrev_A(i)],  X[k] = sum_j x[j] w^(jk),  w = w_{2^A}^(+-1)
template <int A, bool INV>
__device__ __forceinline__ void dft_regs(u32 (&x)[1 << A]) {
  if constexpr (A > 0) dif_stages<A, INV>(x, std::make_integer_sequence<int, A>{});
}
template <int A>
__device__ __forceinline__ constexpr u32 brev(u32 i) {
  u32 r = 0;
  for (int b = 0; b < A; b++) r |= ((i >> b) & 1u) << (A - 1 - b);
  return r;
}

// ---------------------------------------------------------------- the tile pass
struct TileParams {
  const u32* in;
  u32* out;
  u32 tiles_per_col, tiles_b;  // tile id -> (a = id / tiles_b, b = id % tiles_b)
  u32 nz;                      // cosets: blockIdx.x = (col * tiles_per_col + tile) * nz + z, so the nz readers of a tile are co-scheduled (L2)
  u64 in_col, out_col;         // column strides (elements)
  u64 in_z, out_z;             // blockIdx.y (coset) strides
  u32 in_a, in_b, in_r, in_t;  // offsets inside a column (elements)
  u32 out_a, out_b, out_k, out_t;
  u32 split_log, split_max, split_extra;  // row mode only: chunk c = k >> split_log is stored at +c*split_extra if c < split_max
  const uint2* in_tab;  u32 in_tab_z, in_tab_mask;    // (w, wq) at ((input offset inside the column) & mask) + z * in_tab_z
  const uint2* out_tab; u32 out_tab_z, out_tab_mask;  // same for the (pre-split) output offset
  const uint2* tw_mid;         // [2^a][2^b]: w_R^(+-ks*q)
  // MODE 0, last pass of a sharded LDE: output row r of a coset goes to peer.p[r >> peer_log_nj] (and, if r is the first row of a
  // segment, also to the previous rank as its halo row) at the same matrix offset; peer_shift = log2(rows per segment / digit
  // stride), 0xffffffff = plain store into `out`
  u32 peer_shift, peer_mask;   // peer_mask = G - 1
  PeerPtrs peer;               // already offset like `out` relative to the own matrix
};

// MODE 0: lanes contiguous on both sides (in_t = out_t = 1, in_r = out_k = digit stride): a strided digit.
// MODE 1: transform axis contiguous on both sides (in_r = out_k = 1), tile = 16 rows: the lowest digit.
// MODE 2: contiguous on input (in_r = 1), lanes contiguous on output (out_t = 1): transposing pass of a natural-order NTT.
// LOG_S >= 0 fixes the digit stride of MODE 0 at compile time so that every global access is [thread base + immediate].
template <int LOG_R, int MODE>
struct TileShape {
  static constexpr int B_ = LOG_R / 2, A = LOG_R - B_, EA = 1 << A, XB = 1 << B_, R = 1 << LOG_R, NT = 16 * XB;
  static constexpr int ROW_PITCH = EA * (XB + 1);  // MODE 1: groups of XB words padded by one
  static constexpr int SMEM_WORDS = B_ == 0 ? 0 : (MODE == 2 ? 16 * (R + 1) : (MODE == 1 ? 16 * ROW_PITCH : 16 * R));
};

// PEER: the fused row redistribution of a sharded LDE (TileParams::peer); a separate instantiation so that the ordinary kernels
// keep their code and register allocation.
// PROBE (tools/ntt_probe.py, ZKIR_NTT_PROBE): measurement-only instantiations of the 2^10 tiles that split the pass into its two
// floors -- 1 = memory only (same loads, tables and stores, no arithmetic, no exchange), 2 = arithmetic only (no global loads or
// table reads; the store is predicated on a value that never occurs).  Their results are garbage by construction.
template <int LOG_R, int MODE, bool INV, int LOG_S, bool PEER = false, int PROBE = 0>
__global__ void __launch_bounds__(TileShape<LOG_R, MODE>::NT, (LOG_R >= 10 ? 2 : (LOG_R >= 8 ? 4 : 8)))
dft_tile_kernel(const TileParams p) {
  typedef TileShape<LOG_R, MODE> SH;
  constexpr int A = SH::A, B_ = SH::B_, EA = SH::EA, XB = SH::XB;
  extern __shared__ u32 sm[];
  const u32 tid = threadIdx.x;
  const u32 z = blockIdx.x % p.nz, bid = blockIdx.x / p.nz;
  const u32 tile = bid % p.tiles_per_col, col = bid / p.tiles_per_col;
  const u32 ta = tile / p.tiles_b, tb = tile - ta * p.tiles_b;
  const u32 in_off0 = ta * p.in_a + tb * p.in_b, out_off0 = ta * p.out_a + tb * p.out_b;
  // no __restrict__ / ld.global.nc on the data: most passes run in place (a tile reads all of its points before the
  // barrier and writes the same positions after it)
  const u32* in = p.in + (u64)col * p.in_col + (u64)z * p.in_z;
  u32* out = p.out + (u64)col * p.out_col + (u64)z * p.out_z;

  u32 q, l;
  if constexpr (MODE == 0) { l = tid & 15u; q = tid >> 4; } else { q = tid & (XB - 1); l = tid >> B_; }

  // ---- load: thread base pointer + i * (XB * stride)
  const u32 in_r = MODE == 0 ? (LOG_S >= 0 ? (1u << (LOG_S >= 0 ? LOG_S : 0)) : p.in_r) : 1u;
  const u32 in_thr = in_off0 + q * in_r + l * (MODE == 0 ? 1u : p.in_t);
  const size_t in_step = (size_t)XB * in_r;
  u32 x[EA];
  {
    const u32* pin = in + in_thr;
    if constexpr (PROBE == 2) {
#pragma unroll
      for (int i = 0; i < EA; i++) x[i] = tid * 2654435761u + i * 40503u + p.tiles_b;
    } else {
#pragma unroll
      for (int i = 0; i < EA; i++) x[i] = pin[i * in_step];
    }
    if (p.in_tab) {
      const uint2* __restrict__ pt = p.in_tab + (u64)z * p.in_tab_z + (in_thr & p.in_tab_mask);
#pragma unroll
      for (int i = 0; i < EA; i++) {
        if constexpr (PROBE == 2) { x[i] = shoup_mul(x[i], 0x12345u + i, c_shoup(0x12345u) + tid); continue; }
        const uint2 tw = __ldg(pt + i * in_step);
        if constexpr (PROBE == 1) x[i] ^= tw.x + tw.y; else x[i] = shoup_mul(x[i], tw.x, tw.y);
      }
    }
  }
  if constexpr (PROBE != 1) dft_regs<A, INV>(x);

  // ---- store of output index k = ks + EA * kq (kq compile-time), lanes per mode
  const u32 out_k = MODE == 1 ? 1u : (MODE == 0 ? in_r : p.out_k);
  auto store_group = [&](u32 ks_thr, u32 lane, auto& y, auto log_cnt) {
    constexpr int LC = decltype(log_cnt)::value, CNT = 1 << LC;  // y holds CNT values, y[j] = X[ks + EA_eff * brev(j)]
    constexpr int KSTEP = (B_ == 0) ? 1 : EA;                    // single-round tiles: k = brev(i)
    const u32 thr = ks_thr * out_k + lane * (MODE == 1 ? p.out_t : 1u);
    if (MODE == 1 && p.split_log < 32) {  // chunk-splitting store (quotient coefficients); uniform branch, slow path
      const uint2* __restrict__ pt = p.out_tab ? p.out_tab + (u64)z * p.out_tab_z : nullptr;
#pragma unroll
      for (int j = 0; j < CNT; j++) {
        const u32 k = ks_thr + KSTEP * brev<LC>(j);
        u32 v = y[j];
        if (pt) {
          const uint2 tw = __ldg(pt + ((in_off0 + k + lane * p.in_t) & p.out_tab_mask));
          v = shoup_mul(v, tw.x, tw.y);
        }
        const u32 c = k >> p.split_log;
        if (c < p.split_max) out[out_off0 + (k & ((1u << p.split_log) - 1)) + lane * p.out_t + c * p.split_extra] = v;
      }
      return;
    }
    u32* po = out + (out_off0 + thr);
    const size_t kstep = (size_t)KSTEP * out_k;
    if constexpr (PROBE == 2) {   // arithmetic only: keep the values alive, never store
      u32 acc = 0;
#pragma unroll
      for (int j = 0; j < CNT; j++) acc ^= p.out_tab ? shoup_mul(y[j], 0x54321u + j, c_shoup(0x54321u) + lane) : y[j];
      if (acc == 0xdeadbeefu && p.tiles_per_col == 0xffffffffu) po[0] = acc;
      return;
    }
    if constexpr (PEER) {
      // fused redistribution: digit index k = ks + KSTEP*brev(j) decides the owner of the row; (out_off0 + thr) < stride never carries
      const size_t moff = (size_t)(po - p.out);
      const bool low0 = (out_off0 + lane) == 0;   // this thread's rows are multiples of the digit stride
      const u32 seg_mask = (1u << p.peer_shift) - 1;
#pragma unroll
      for (int j = 0; j < CNT; j++) {
        const u32 k = ks_thr + KSTEP * brev<LC>(j);
        const u32 dest = k >> p.peer_shift;
        p.peer.p[dest][moff + brev<LC>(j) * kstep] = y[j];
        if (low0 && (k & seg_mask) == 0) {
          const u32 prev = (dest - 1) & p.peer_mask;
          if (prev != dest) p.peer.p[prev][moff + brev<LC>(j) * kstep] = y[j];
        }
      }
      return;
    }
    if (p.out_tab) {
      // row mode: the table follows the input layout of the row, otherwise the output layout
      const u32 toff = MODE == 1 ? (in_off0 + lane * p.in_t + ks_thr) : (out_off0 + thr);
      const uint2* __restrict__ pt = p.out_tab + (u64)z * p.out_tab_z + (toff & p.out_tab_mask);
#pragma unroll
      for (int j = 0; j < CNT; j++) {
        const uint2 tw = __ldg(pt + brev<LC>(j) * kstep);
        if constexpr (PROBE == 1) po[brev<LC>(j) * kstep] = y[j] ^ (tw.x + tw.y); else po[brev<LC>(j) * kstep] = shoup_mul(y[j], tw.x, tw.y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; j++) po[brev<LC>(j) * kstep] = y[j];
    }
  };

  if constexpr (B_ == 0) {
    store_group(0u, l, x, std::integral_constant<int, A>{});
  } else if constexpr (PROBE == 1) {
    // memory only: every thread stores XB-sized groups of the values it loaded at the addresses the real kernel writes
    u32 u, l2;
    if constexpr (MODE == 1) { u = tid & (XB - 1); l2 = tid >> B_; } else { l2 = tid & 15u; u = tid >> 4; }
#pragma unroll
    for (int g = 0; g < EA / XB; g++) {
      u32 y[XB];
#pragma unroll
      for (int j = 0; j < XB; j++) y[j] = x[(g * XB + j) % EA];
      store_group(u + g * XB, l2, y, std::integral_constant<int, B_>{});
    }
  } else {
    // ---- inter-round twiddle + exchange.  Shared-memory word of (ks, q, lane):
    //  MODE 0: line (q*EA/2 + ks/2) of 32 words, half ((ks^q)&1), lane           -> two thread bases + immediates
    //  MODE 1: lane*ROW_PITCH + ks*(XB+1) + q                                     -> one thread base + immediates
    //  MODE 2: lane*(R+1) + ks*XB + ((q + XB/2*(ks&1)) mod XB)                    (natural-order API pass only)
    const uint2* __restrict__ twq = p.tw_mid + q;
    u32 w0, w1;
    if constexpr (MODE == 0) {
      w0 = q * (EA / 2) * 32 + l + ((q & 1u) << 4);
      w1 = q * (EA / 2) * 32 + l + (((q & 1u) ^ 1u) << 4);
    } else if constexpr (MODE == 1) {
      w0 = w1 = l * SH::ROW_PITCH + q;
    } else {
      w0 = w1 = 0;
    }
#pragma unroll
    for (int i = 0; i < EA; i++) {
      constexpr int dummy = 0; (void)dummy;
      const u32 ks = brev<A>(i);
      u32 v = x[i];
      if (ks != 0) {
        if constexpr (PROBE == 2) v = shoup_mul(v, 0x777u + ks, c_shoup(0x777u) + q);
        else { const uint2 tw = __ldg(twq + ks * XB); v = shoup_mul(v, tw.x, tw.y); }
      }
      if constexpr (MODE == 0) sm[((ks & 1u) ? w1 : w0) + (ks >> 1) * 32] = v;
      else if constexpr (MODE == 1) sm[w0 + ks * (XB + 1)] = v;
      else sm[l * (SH::R + 1) + ks * XB + ((q + (XB / 2) * (ks & 1u)) & (XB - 1))] = v;
    }
    __syncthreads();
    u32 u, l2;
    if constexpr (MODE == 1) { u = tid & (XB - 1); l2 = tid >> B_; } else { l2 = tid & 15u; u = tid >> 4; }
#pragma unroll
    for (int g = 0; g < EA / XB; g++) {
      const u32 ks = u + g * XB;
      u32 y[XB];
      if constexpr (MODE == 0) {
        const u32 r0 = (ks >> 1) * 32 + l2 + ((ks & 1u) << 4), r1 = (ks >> 1) * 32 + l2 + (((ks & 1u) ^ 1u) << 4);
#pragma unroll
        for (int j = 0; j < XB; j++) y[j] = sm[((j & 1) ? r1 : r0) + j * (EA / 2) * 32];
      } else if constexpr (MODE == 1) {
        const u32 r0 = l2 * SH::ROW_PITCH + ks * (XB + 1);
#pragma unroll
        for (int j = 0; j < XB; j++) y[j] = sm[r0 + j];
      } else {
#pragma unroll
        for (int j = 0; j < XB; j++) y[j] = sm[l2 * (SH::R + 1) + ks * XB + ((j + (XB / 2) * (ks & 1u)) & (XB - 1))];
      }
      dft_regs<B_, INV>(y);
      store_group(ks, l2, y, std::integral_constant<int, B_>{});
    }
  }
}

template <int LOG_R, int MODE, bool INV, int LOG_S = -1, bool PEER = false, int PROBE = 0>
static cudaError_t launch_tile_t(const TileParams& p, u32 blocks, u32 z, cudaStream_t st) {
  typedef TileShape<LOG_R, MODE> SH;
  const size_t smem = (size_t)SH::SMEM_WORDS * sizeof(u32);
  auto kern = dft_tile_kernel<LOG_R, MODE, INV, LOG_S, PEER, PROBE>;
  if (smem > 48 * 1024) {  // per device and cheap: set it on every launch rather than caching a process-wide flag
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  TileParams q = p;
  q.nz = z;
  kern<<<blocks * z, SH::NT, smem, st>>>(q);
  return cudaGetLastError();
}
template <int MODE, bool INV>
static cudaError_t launch_tile_m(int log_r, const TileParams& p, u32 blocks, u32 z, cudaStream_t st) {
  if constexpr (MODE == 0 && !INV) {
    if (p.peer_shift != 0xffffffffu) {   // last pass of a sharded LDE: top digits are 4..10 bits wide
      if (log_r == 10 && p.in_r == 1024u) return launch_tile_t<10, 0, false, 10, true>(p, blocks, z, st);
      switch (log_r) {
        case 4: return launch_tile_t<4, 0, false, -1, true>(p, blocks, z, st);
        case 5: return launch_tile_t<5, 0, false, -1, true>(p, blocks, z, st);
        case 6: return launch_tile_t<6, 0, false, -1, true>(p, blocks, z, st);
        case 7: return launch_tile_t<7, 0, false, -1, true>(p, blocks, z, st);
        case 8: return launch_tile_t<8, 0, false, -1, true>(p, blocks, z, st);
        case 9: return launch_tile_t<9, 0, false, -1, true>(p, blocks, z, st);
        case 10: return launch_tile_t<10, 0, false, -1, true>(p, blocks, z, st);
      }
      return cudaErrorInvalidValue;
    }
  }
  // hot shape of a 2^20-row trace (digits 10+10): strided digit with a compile-time stride of 2^10
  static const int probe = getenv("ZKIR_NTT_PROBE") ? atoi(getenv("ZKIR_NTT_PROBE")) : 0;   // measurement only, see dft_tile_kernel
  if (probe && log_r == 10 && MODE != 2) {
    if (MODE == 0 && p.in_r == 1024u)
      return probe == 1 ? launch_tile_t<10, MODE, INV, (MODE == 0 ? 10 : -1), false, 1>(p, blocks, z, st) : launch_tile_t<10, MODE, INV, (MODE == 0 ? 10 : -1), false, 2>(p, blocks, z, st);
    if (MODE == 1) return probe == 1 ? launch_tile_t<10, MODE, INV, -1, false, 1>(p, blocks, z, st) : launch_tile_t<10, MODE, INV, -1, false, 2>(p, blocks, z, st);
  }
  if (MODE == 0 && log_r == 10 && p.in_r == 1024u) return launch_tile_t<10, MODE, INV, (MODE == 0 ? 10 : -1)>(p, blocks, z, st);
  switch (log_r) {
    case 1: return launch_tile_t<1, MODE, INV>(p, blocks, z, st);
    case 2: return launch_tile_t<2, MODE, INV>(p, blocks, z, st);
    case 3: return launch_tile_t<3, MODE, INV>(p, blocks, z, st);
    case 4: return launch_tile_t<4, MODE, INV>(p, blocks, z, st);
    case 5: return launch_tile_t<5, MODE, INV>(p, blocks, z, st);
    case 6: return launch_tile_t<6, MODE, INV>(p, blocks, z, st);
    case 7: return launch_tile_t<7, MODE, INV>(p, blocks, z, st);
    case 8: return launch_tile_t<8, MODE, INV>(p, blocks, z, st);
    case 9: return launch_tile_t<9, MODE, INV>(p, blocks, z, st);
    case 10: return launch_tile_t<10, MODE, INV>(p, blocks, z, st);
  }
  return cudaErrorInvalidValue;
}
static cudaError_t launch_tile(int log_r, int mode, bool inv, const TileParams& p, u32 blocks, u32 z, cudaStream_t st) {
  if (mode == 0) return inv ? launch_tile_m<0, true>(log_r, p, blocks, z, st) : launch_tile_m<0, false>(log_r, p, blocks, z, st);
  if (mode == 1) return inv ? launch_tile_m<1, true>(log_r, p, blocks, z, st) : launch_tile_m<1, false>(log_r, p, blocks, z, st);
  return inv ? launch_tile_m<2, true>(log_r, p, blocks, z, st) : launch_tile_m<2, false>(log_r, p, blocks, z, st);
}

// ---------------------------------------------------------------- table generation (one-off per shape, cached)
__device__ __forceinline__ uint2 make_pair_from_mont(u32 v_mont) {
  const u32 w = bb_from_mont(v_mont);
  return make_uint2(w, (u32)((((u64)w) << 32) / BB_P));
}
// out[ks * XB + q] = root^(ks*q)
__global__ void gen_tw_mid_kernel(uint2* out, u32 n, u32 log_xb, u32 root_mont) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 ks = i >> log_xb, q = i & ((1u << log_xb) - 1);
  out[i] = make_pair_from_mont(bb_pow(root_mont, (u64)ks * q));
}
// out[o] = c0 * root^((o >> log_rest) * (o & (rest - 1))),  o < S
__global__ void gen_level_tw_kernel(uint2* out, u64 S, u32 log_rest, u32 root_mont, u32 c0_mont) {
  const u64 o = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (o >= S) return;
  const u64 k = o >> log_rest, rest = o & ((1ull << log_rest) - 1);
  out[o] = make_pair_from_mont(bb_mul(c0_mont, bb_pow(root_mont, (k * rest) & (S - 1))));
}
struct DigitPlan {
  int nd;
  int d[4];
};
__host__ __device__ inline u64 plan_coef_index(const DigitPlan& pl, u64 pos) {
  // position digits top->bottom (x_1 .. x_D); coefficient k = x_1 + 2^d1 x_2 + 2^(d1+d2) x_3 ...
  int below = 0;
  for (int j = 0; j < pl.nd; j++) below += pl.d[j];
  u64 k = 0;
  int wshift = 0;
  for (int j = 0; j < pl.nd; j++) {
    below -= pl.d[j];
    const u64 xj = (pos >> below) & ((1ull << pl.d[j]) - 1);
    k |= xj << wshift;
    wshift += pl.d[j];
  }
  return k;
}
// out[z * n + pos] = c0 * (base * zroot^z)^(k(pos))
__global__ void gen_scale_kernel(uint2* out, u64 n, DigitPlan pl, u32 base_mont, u32 zroot_mont, u32 c0_mont) {
  const u64 pos = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const u32 z = blockIdx.y;
  const u32 b = bb_mul(base_mont, bb_pow(zroot_mont, z));
  out[(u64)z * n + pos] = make_pair_from_mont(bb_mul(c0_mont, bb_pow(b, plan_coef_index(pl, pos))));
}

static u32 h_pow(u32 a, u64 e) { u64 r = 1, b = a; while (e) { if (e & 1) r = r * b % BB_P; b = b * b % BB_P; e >>= 1; } return (u32)r; }
static u32 h_inv(u32 a) { return h_pow(a, BB_P - 2); }

struct FastNtt {
  cudaStream_t stream;
  u64* launches;
  std::map<std::vector<u64>, uint2*> cache;
  std::vector<void*> owned;

  uint2* alloc_pairs(u64 n) {
    void* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(uint2)) != cudaSuccess) return nullptr;
    owned.push_back(d);
    return (uint2*)d;
  }
  const uint2* tw_mid(int log_r, bool inv) {
    std::vector<u64> key = {1, (u64)log_r, (u64)inv};
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const int b = log_r / 2;
    const u32 n = 1u << log_r;
    uint2* d = alloc_pairs(n);
    if (!d) return nullptr;
    u32 root = ZKIR_BB_ROOTS[log_r];
    if (inv) root = h_inv(root);
    gen_tw_mid_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d, n, b, bb_to_mont_c(root));
    (*launches)++;
    cache[key] = d;
    return d;
  }
  // twiddles of the level whose block size is 2^log_s and whose sub-block ("rest") size is 2^log_rest, times c0
  const uint2* level_tw(int log_s, int log_rest, bool inv, u32 c0) {
    std::vector<u64> key = {2, (u64)log_s, (u64)log_rest, (u64)inv, c0};
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const u64 S = 1ull << log_s;
    uint2* d = alloc_pairs(S);
    if (!d) return nullptr;
    u32 root = ZKIR_BB_ROOTS[log_s];
    if (inv) root = h_inv(root);
    gen_level_tw_kernel<<<(unsigned)((S + 255) / 256), 256, 0, stream>>>(d, S, log_rest, bb_to_mont_c(root), bb_to_mont_c(c0));
    (*launches)++;
    cache[key] = d;
    return d;
  }
  // [nz][n]: c0 * (base * zroot^z)^(k(pos))
  const uint2* scale(const DigitPlan& pl, u32 base, u32 zroot, u32 nz, u32 c0) {
    std::vector<u64> key = {3, (u64)pl.nd, (u64)pl.d[0], (u64)pl.d[1], (u64)pl.d[2], (u64)pl.d[3], base, zroot, nz, c0};
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int log_n = 0;
    for (int j = 0; j < pl.nd; j++) log_n += pl.d[j];
    const u64 n = 1ull << log_n;
    uint2* d = alloc_pairs(n * nz);
    if (!d) return nullptr;
    gen_scale_kernel<<<dim3((unsigned)((n + 255) / 256), nz), 256, 0, stream>>>(d, n, pl, bb_to_mont_c(base), bb_to_mont_c(zroot), bb_to_mont_c(c0));
    (*launches)++;
    cache[key] = d;
    return d;
  }
  void free_all() { for (void* p : owned) cudaFree(p); owned.clear(); cache.clear(); }
};

FastNtt* fast_ntt_create(cudaStream_t st, u64* launch_counter) {
  FastNtt* f = new FastNtt();
  f->stream = st; f->launches = launch_counter;
  return f;
}
void fast_ntt_destroy(FastNtt* f) { if (f) { f->free_all(); delete f; } }

// digits for 2^log_n: at least two digits, each <= 10 bits, as even as possible; false if log_n < 8
bool fast_plan(int log_n, FastPlan* out) {
  if (log_n < 8 || log_n > 30) return false;
  int nd = (log_n + 9) / 10;
  if (nd < 2) nd = 2;
  out->log_n = log_n; out->nd = nd;
  const int base = log_n / nd, rem = log_n % nd;
  for (int i = 0; i < 4; i++) out->d[i] = 0;
  for (int i = 0; i < nd; i++) out->d[i] = base + (i >= nd - rem ? 1 : 0);  // larger digits last: the lowest one is split by the blowup
  return true;
}
static DigitPlan to_digit_plan(const FastPlan& pl) {
  DigitPlan d;
  d.nd = pl.nd;
  for (int i = 0; i < 4; i++) d.d[i] = pl.d[i];
  return d;
}
u64 fast_plan_coef_index(const FastPlan& pl, u64 pos) { return plan_coef_index(to_digit_plan(pl), pos); }

#define FCHECK(e) do { if ((e) != cudaSuccess) return -2; } while (0)

static void no_split(TileParams& p) { p.split_log = 32; p.split_max = 0; p.split_extra = 0; p.peer_shift = 0xffffffffu; p.peer_mask = 0; }

// strided pass over digit j (0-based) of `pl`, in place or in -> out with identical layouts
static int strided_pass(FastNtt* f, const FastPlan& pl, int j, bool inv, const u32* in, u64 in_col, u64 in_z, u32* out, u64 out_col, u64 out_z,
                        u32 n_cols, u32 nz, const uint2* out_tab, u32 out_tab_mask, const uint2* in_tab, u32 in_tab_mask, cudaStream_t st,
                        const PeerPtrs* peers = nullptr, u32 me = 0, u32 G = 1) {
  int log_stride = 0;
  for (int i = j + 1; i < pl.nd; i++) log_stride += pl.d[i];
  const int log_block = log_stride + pl.d[j];
  TileParams p = {};
  p.in = in; p.out = out;
  p.in_col = in_col; p.out_col = out_col; p.in_z = in_z; p.out_z = out_z;
  p.tiles_b = (1u << log_stride) / 16;
  p.tiles_per_col = (u32)((1ull << pl.log_n) >> log_block) * p.tiles_b;
  p.in_a = p.out_a = 1u << log_block;
  p.in_b = p.out_b = 16;
  p.in_r = p.out_k = 1u << log_stride;
  p.in_t = p.out_t = 1;
  no_split(p);
  p.in_tab = in_tab; p.in_tab_mask = in_tab_mask;
  p.out_tab = out_tab; p.out_tab_mask = out_tab_mask;
  p.tw_mid = f->tw_mid(pl.d[j], inv);
  if (!p.tw_mid) return -4;
  if (peers) {   // only the top digit spans all row segments: rows per segment = 2^(log_n) / G = 2^peer_shift digit strides
    if (j != 0 || out_tab) return -1;
    int log_g = 0;
    while ((1u << log_g) < G) log_g++;
    if (log_g > pl.d[0]) return -1;
    p.peer_shift = (u32)(pl.d[0] - log_g);
    p.peer_mask = G - 1;
    const ptrdiff_t rel = out - peers->p[me];
    for (u32 g = 0; g < G; g++) p.peer.p[g] = peers->p[g] + rel;
  }
  FCHECK(launch_tile(pl.d[j], 0, inv, p, n_cols * p.tiles_per_col, nz, st));
  (*f->launches)++;
  return 0;
}

// row pass over the lowest digit of `pl`
static int row_pass(FastNtt* f, const FastPlan& pl, bool inv, const u32* in, u64 in_col, u64 in_z, u32* out, u64 out_col, u64 out_z,
                    u32 n_cols, u32 nz, const uint2* in_tab, u32 in_tab_z, const uint2* out_tab, u32 out_tab_mask,
                    u32 split_log, u32 split_max, u32 split_extra, cudaStream_t st) {
  const int dl = pl.d[pl.nd - 1];
  const u32 m = 1u << dl, rows = (u32)((1ull << pl.log_n) >> dl);
  TileParams p = {};
  p.in = in; p.out = out;
  p.in_col = in_col; p.out_col = out_col; p.in_z = in_z; p.out_z = out_z;
  p.tiles_b = rows / 16; p.tiles_per_col = rows / 16;
  p.in_b = 16 * m; p.in_t = m; p.in_r = 1;
  p.out_k = 1;
  if (split_log < 32) { p.out_b = 16u << split_log; p.out_t = 1u << split_log; }
  else { p.out_b = 16 * m; p.out_t = m; }
  p.split_log = split_log; p.split_max = split_max; p.split_extra = split_extra; p.peer_shift = 0xffffffffu; p.peer_mask = 0;
  p.in_tab = in_tab; p.in_tab_z = in_tab_z; p.in_tab_mask = 0xffffffffu;
  p.out_tab = out_tab; p.out_tab_mask = out_tab_mask;
  p.tw_mid = f->tw_mid(dl, inv);
  if (!p.tw_mid) return -4;
  FCHECK(launch_tile(dl, 1, inv, p, n_cols * p.tiles_per_col, nz, st));
  (*f->launches)++;
  return 0;
}

// Inverse transform of n_cols columns: in [n_cols][N] (natural order, canonical or Montgomery values) ->
// coef [n_cols][N] in digit-reversed order, every coefficient multiplied by c0 (canonical constant, e.g. R/N).
// Optional per-coefficient scale `coef_tab` ([N] pairs indexed by position) applied by the last pass, and an optional
// split of the lowest digit into chunks (see TileParams).
int fast_intt(FastNtt* f, const FastPlan& pl, const u32* in, u64 in_col, u32* coef, u64 coef_col, u32 n_cols, u32 c0,
              const uint2* coef_tab, u32 split_log, u32 split_max, u32 split_extra, u32* split_out, u64 split_out_col, cudaStream_t st) {
  int log_block = pl.log_n;
  const u32* src = in;
  u64 src_col = in_col;
  for (int j = 0; j + 1 < pl.nd; j++) {
    const int log_rest = log_block - pl.d[j];
    const uint2* tw = f->level_tw(log_block, log_rest, true, j == 0 ? c0 : 1u);
    if (!tw) return -4;
    int rc = strided_pass(f, pl, j, true, src, src_col, 0, coef, coef_col, 0, n_cols, 1, tw, (u32)((1ull << log_block) - 1), nullptr, 0, st);
    if (rc) return rc;
    src = coef; src_col = coef_col;
    log_block = log_rest;
  }
  if (split_log < 32)
    return row_pass(f, pl, true, src, src_col, 0, split_out, split_out_col, 0, n_cols, 1, nullptr, 0, coef_tab, 0xffffffffu, split_log, split_max, split_extra, st);
  return row_pass(f, pl, true, src, src_col, 0, coef, coef_col, 0, n_cols, 1, nullptr, 0, coef_tab, 0xffffffffu, 32, 0, 0, st);
}

// Forward transforms on nz cosets from digit-reversed coefficients: out[col][z][i] = sum_k coef[k] (base*zroot^z)^k w_N^(ik) * c0
// coef layout [n_cols][N] (column stride coef_col), out column stride out_col, coset stride N.
bool fast_coset_ntt_can_fuse(const FastPlan& pl, u32 G) { return G >= 2 && (G & (G - 1)) == 0 && G <= (1u << pl.d[0]) && G <= ZKIR_MAX_SHARDS; }

int fast_coset_ntt(FastNtt* f, const FastPlan& pl, const u32* coef, u64 coef_col, u32* out, u64 out_col, u32 n_cols, u32 nz,
                   u32 base, u32 zroot, u32 c0, cudaStream_t st, const PeerPtrs* peers, u32 me, u32 G) {
  const u64 N = 1ull << pl.log_n;
  const uint2* sc = f->scale(to_digit_plan(pl), base, zroot, nz, c0);
  if (!sc) return -4;
  // lowest digit: rows, with the level twiddle of the block above it
  {
    const int dl = pl.d[pl.nd - 1];
    const int log_block = dl + pl.d[pl.nd - 2];
    const uint2* tw = f->level_tw(log_block, dl, false, 1u);
    if (!tw) return -4;
    int rc = row_pass(f, pl, false, coef, coef_col, 0, out, out_col, N, n_cols, nz, sc, (u32)N, tw, (u32)((1ull << log_block) - 1), 32, 0, 0, st);
    if (rc) return rc;
  }
  int log_rest = pl.d[pl.nd - 1];
  for (int j = pl.nd - 2; j >= 0; j--) {
    const int log_block = log_rest + pl.d[j];  // block transformed after this pass
    const uint2* tw = nullptr;
    u32 mask = 0;
    if (j > 0) {
      const int log_up = log_block + pl.d[j - 1];
      tw = f->level_tw(log_up, log_block, false, 1u);
      if (!tw) return -4;
      mask = (u32)((1ull << log_up) - 1);
    }
    int rc = strided_pass(f, pl, j, false, out, out_col, N, out, out_col, N, n_cols, nz, tw, mask, nullptr, 0, st, j == 0 ? peers : nullptr, me, G);
    if (rc) return rc;
    log_rest = log_block;
  }
  return 0;
}

// Natural-order in, natural-order out transform of n_cols columns (two digits only): pass A strided over the top
// digit into tmp, pass B contiguous over the low digit with a transposing store.  Canonical values.
int fast_ntt_natural(FastNtt* f, int log_n, bool inverse, u32 coset_shift, const u32* in, u64 in_col, u32* tmp, u32* out, u64 out_col,
                     u32 n_cols, cudaStream_t st) {
  FastPlan pl;
  if (!fast_plan(log_n, &pl) || pl.nd != 2) return -1;
  const u64 N = 1ull << log_n;
  const u32 c0 = inverse ? h_inv((u32)(N % BB_P)) : 1u;
  const uint2* tw = f->level_tw(log_n, pl.d[1], inverse, c0);
  if (!tw) return -4;
  const uint2* in_tab = nullptr;
  if (coset_shift) {
    DigitPlan nat; nat.nd = 1; nat.d[0] = log_n; nat.d[1] = nat.d[2] = nat.d[3] = 0;
    in_tab = f->scale(nat, coset_shift, 1u, 1, 1u);
    if (!in_tab) return -4;
  }
  int rc = strided_pass(f, pl, 0, inverse, in, in_col, 0, tmp, N, 0, n_cols, 1, tw, (u32)(N - 1), in_tab, 0xffffffffu, st);
  if (rc) return rc;
  const u32 m = 1u << pl.d[1], R1 = 1u << pl.d[0];
  TileParams p = {};
  p.in = tmp; p.out = out;
  p.in_col = N; p.out_col = out_col;
  p.tiles_b = R1 / 16; p.tiles_per_col = R1 / 16;
  p.in_b = 16 * m; p.in_t = m; p.in_r = 1;
  p.out_b = 16; p.out_t = 1; p.out_k = R1;
  no_split(p);
  p.tw_mid = f->tw_mid(pl.d[1], inverse);
  if (!p.tw_mid) return -4;
  FCHECK(launch_tile(pl.d[1], 2, inverse, p, n_cols * p.tiles_per_col, 1, st));
  (*f->launches)++;
  return 0;
}

// natural <-> coset-major reordering of an LDE matrix (API / test paths only): nat[col][i*B + z] <-> cm[col][z][i]
__global__ void coset_reorder_kernel(const u32* __restrict__ in, u32* __restrict__ out, u32 log_n, u32 log_b, u64 total, int to_natural) {
  const u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g >= total) return;
  const u32 log_m = log_n + log_b;
  const u64 col = g >> log_m, r = g & ((1ull << log_m) - 1);
  // g enumerates the DESTINATION linearly
  if (to_natural) {
    const u64 z = r & ((1ull << log_b) - 1), i = r >> log_b;
    out[g] = in[(col << log_m) + (z << log_n) + i];
  } else {
    const u64 z = r >> log_n, i = r & ((1ull << log_n) - 1);
    out[g] = in[(col << log_m) + (i << log_b) + z];
  }
}
int launch_coset_reorder(const u32* in, u32* out, u32 n_cols, u32 log_n, u32 log_b, int to_natural, cudaStream_t st, u64* launches) {
  const u64 total = (u64)n_cols << (log_n + log_b);
  coset_reorder_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, log_n, log_b, total, to_natural);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

const uint2* fast_scale_table(FastNtt* f, const FastPlan& pl, u32 base, u32 c0) { return f->scale(to_digit_plan(pl), base, 1u, 1, c0); }

}  // namespace zkir
