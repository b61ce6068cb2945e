// Poseidon2-BabyBear (width 16, x^7, 8 external + 13 internal rounds) kernels: permutation, row sponge (Merkle
// leaves), 2-to-1 compression (Merkle levels), the duplex challenger and the proof-of-work grinder.
//
// The reference's Poseidon2 is a stub that returns Err (zkir-runtime/src/crypto.rs:299-315, pinned by
// zkir-runtime/tests/syscall_integration.rs:400-422), so there is nothing upstream to match; parameters and
// round constants are frozen by docs/PROVER_SPEC.md / tools/gen_constants.py.  All values are Montgomery form.
// Bound: integer ALU (about 770 modular multiplies per permutation), not HBM.
#include <cuda_runtime.h>
#include <stdlib.h>
#include "bb.cuh"
#include "kernels.h"
#include "constants_generated.h"

namespace zkir {

__constant__ u32 c_rc_ext[128];
__constant__ u32 c_rc_int[ZKIR_P2_RP];
__constant__ u32 c_diag[16];

int poseidon2_init_constants() {
  u32 h[128];
  for (int i = 0; i < 128; i++) h[i] = bb_to_mont_c(ZKIR_P2_RC_EXT[i]);
  if (cudaMemcpyToSymbol(c_rc_ext, h, sizeof(u32) * 128) != cudaSuccess) return -2;
  for (int i = 0; i < ZKIR_P2_RP; i++) h[i] = bb_to_mont_c(ZKIR_P2_RC_INT[i]);
  if (cudaMemcpyToSymbol(c_rc_int, h, sizeof(u32) * ZKIR_P2_RP) != cudaSuccess) return -2;
  for (int i = 0; i < 16; i++) h[i] = bb_to_mont_c(ZKIR_P2_DIAG[i]);
  if (cudaMemcpyToSymbol(c_diag, h, sizeof(u32) * 16) != cudaSuccess) return -2;
  // internal_linear() hard-codes the structure of the diagonal: refuse to run if the generated constants ever change
  const u32 half = (BB_P + 1) / 2;
  auto hp = [](u32 a, u32 e) { u64 r = 1, b = a; while (e) { if (e & 1) r = r * b % BB_P; b = b * b % BB_P; e >>= 1; } return (u32)r; };
  const u32 want[16] = {BB_P - 2, 1, 2, half, 3, 4, BB_P - half, BB_P - 3, BB_P - 4, hp(half, 8), hp(half, 2), hp(half, 3), hp(half, 27),
                        BB_P - hp(half, 8), BB_P - hp(half, 4), BB_P - hp(half, 27)};
  for (int i = 0; i < 16; i++) if (ZKIR_P2_DIAG[i] != want[i]) return -2;
  return 0;
}

// Montgomery product with the `hi - t` of the reduction written as __viaddmin_u32 (an ALU-pipe instruction; see the note on pipe placement
// below): measured on the 2^20-row proof, 13.57 -> 13.18 ms with all four S-box products in this form (one / two of them: 13.53 / 13.39);
// computing lo * p^-1 by shifts and adds on top of it: 13.94 (slower).
__device__ __forceinline__ u32 p2_mul(u32 a, u32 b) {
  const u64 x = (u64)a * b;
  const u32 lo = (u32)x, hi = (u32)(x >> 32);
  const u32 t = __umulhi(lo * BB_PINV, BB_P);
  const u32 r = __viaddmin_u32(hi, 0u - t, 0xffffffffu);   // hi - t mod 2^32
  return __viaddmin_u32(r, BB_P, r);                        // min(r + p, r): the canonical representative
}
__device__ __forceinline__ u32 sbox7(u32 x) {
  const u32 x2 = p2_mul(x, x), x3 = p2_mul(x2, x), x4 = p2_mul(x2, x2);
  return p2_mul(x4, x3);
}
// Pipe placement of the plain adds of the thread-per-permutation kernels (leaves, compressions).  A modular add is `s = a + b` followed by
// VIADDMNMX(s, -p, s); ptxas issues about half of the `a + b` as IMAD.IADD on the integer-multiply pipe and half as IADD3 on the ALU pipe.
// Writing an add as __viaddmin_u32(a, b, ~0u) = min(a + b, 2^32 - 1) = a + b pins it to the ALU pipe (VIADDMNMX exists only there) and makes
// ptxas re-balance the rest.  Measured on the 2^20-row proof (ms per proof / trace commitment; ZKIR_P2_PIN bit 0 = the external linear
// layer, bit 1 = the internal one, bit 2 = the round-constant adds): 0: 13.89 / 6.69, 1: 14.56 / 7.18, 2: 14.01 / 6.78, 3: 14.40 / 7.06,
// **4: 13.57 / 6.47**, 5: 14.15 / 6.89, 6: 13.77 / 6.61; unrolling the round loops on top of 4: 14.4 - 15.5 (instruction cache).
// Only the round-constant adds gain (their second operand comes from constant memory); pinning the layers' adds overloads the ALU pipe.
#ifndef ZKIR_P2_PIN
#define ZKIR_P2_PIN 4
#endif
template <int ON>
__device__ __forceinline__ u32 p2_add(u32 a, u32 b) {
  if (ON) { const u32 t = __viaddmin_u32(a, b, 0xffffffffu); return __viaddmin_u32(t, 0u - BB_P, t); }
  return bb_add(a, b);
}
template <int ON>
__device__ __forceinline__ u32 p2_sub(u32 a, u32 b) {
  if (ON) { const u32 d = __viaddmin_u32(a, 0u - b, 0xffffffffu); return __viaddmin_u32(d, BB_P, d); }   // a - b wraps mod 2^32; min(d + p, d)
  return bb_sub(a, b);
}
#define EA p2_add<(ZKIR_P2_PIN & 1)>
#define IA p2_add<((ZKIR_P2_PIN >> 1) & 1)>
#define IS p2_sub<((ZKIR_P2_PIN >> 1) & 1)>
#define RA p2_add<((ZKIR_P2_PIN >> 2) & 1)>

// circ(2,3,1,1) on 4 values
__device__ __forceinline__ void m4(u32& a, u32& b, u32& c, u32& d) {
  u32 t01 = EA(a, b), t23 = EA(c, d);
  u32 t0123 = EA(t01, t23);
  u32 t01123 = EA(t0123, b), t01233 = EA(t0123, d);
  u32 nd = EA(t01233, EA(a, a));  // 3a + b + c + 2d
  u32 nb = EA(t01123, EA(c, c));  // a + 2b + 3c + d
  u32 na = EA(t01123, t01);       // 2a + 3b + c + d
  u32 nc = EA(t01233, t23);       // a + b + 2c + 3d
  a = na; b = nb; c = nc; d = nd;
}
__device__ __forceinline__ void external_linear(u32* s) {
#pragma unroll
  for (int c = 0; c < 4; c++) m4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]);
  u32 sums[4];
#pragma unroll
  for (int k = 0; k < 4; k++) sums[k] = EA(EA(s[k], s[4 + k]), EA(s[8 + k], s[12 + k]));
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = EA(s[i], sums[i & 3]);
}
// s[i] <- V[i]*s[i] + sum(s) with the frozen diagonal V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27,
// -2^-8, -1/16, -2^-27] (docs/PROVER_SPEC.md section 2; poseidon2_init_constants() checks the generated header against it).
// The nine small entries are additions / halvings (no integer-multiply pipe work), the other seven are Shoup products
// by compile-time constants; the values stay Montgomery representatives throughout (the map is linear).
constexpr u32 C_HALF = (BB_P + 1) / 2;
constexpr u32 C_DIAG_TAIL[7] = {c_pow(C_HALF, 8), c_pow(C_HALF, 2), c_pow(C_HALF, 3), c_pow(C_HALF, 27),
                                BB_P - c_pow(C_HALF, 8), BB_P - c_pow(C_HALF, 4), BB_P - c_pow(C_HALF, 27)};
template <int I>
__device__ __forceinline__ u32 diag_tail_mul(u32 x) {
  constexpr u32 w = C_DIAG_TAIL[I], wq = c_shoup(w);
  return shoup_mul(x, w, wq);
}
__device__ __forceinline__ void internal_linear(u32* s) {
  u32 sum = s[0];
#pragma unroll
  for (int i = 1; i < 16; i++) sum = IA(sum, s[i]);
  s[0] = IS(sum, IA(s[0], s[0]));
  s[1] = IA(s[1], sum);
  s[2] = IA(IA(s[2], s[2]), sum);
  s[3] = IA(bb_halve(s[3]), sum);
  s[4] = IA(IA(IA(s[4], s[4]), s[4]), sum);
  { const u32 d = IA(s[5], s[5]); s[5] = IA(IA(d, d), sum); }
  s[6] = IS(sum, bb_halve(s[6]));
  s[7] = IS(sum, IA(IA(s[7], s[7]), s[7]));
  { const u32 d = IA(s[8], s[8]); s[8] = IS(sum, IA(d, d)); }
  s[9] = IA(diag_tail_mul<0>(s[9]), sum);
  s[10] = IA(diag_tail_mul<1>(s[10]), sum);
  s[11] = IA(diag_tail_mul<2>(s[11]), sum);
  s[12] = IA(diag_tail_mul<3>(s[12]), sum);
  s[13] = IA(diag_tail_mul<4>(s[13]), sum);
  s[14] = IA(diag_tail_mul<5>(s[14]), sum);
  s[15] = IA(diag_tail_mul<6>(s[15]), sum);
}
__device__ __forceinline__ void poseidon2_permute(u32* s) {
  external_linear(s);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = sbox7(RA(s[i], c_rc_ext[r * 16 + i]));
    external_linear(s);
  }
#pragma unroll 1
  for (int r = 0; r < ZKIR_P2_RP; r++) {
    s[0] = sbox7(RA(s[0], c_rc_int[r]));
    internal_linear(s);
  }
#pragma unroll 1
  for (int r = 4; r < 8; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = sbox7(RA(s[i], c_rc_ext[r * 16 + i]));
    external_linear(s);
  }
}

__global__ void permute_kernel(u32* states, u64 n, int canonical_io) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 s[16];
  uint4* p = reinterpret_cast<uint4*>(states + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; k++) { uint4 v = p[k]; s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
  if (canonical_io) for (int k = 0; k < 16; k++) s[k] = bb_to_mont(s[k]);
  poseidon2_permute(s);
  if (canonical_io) for (int k = 0; k < 16; k++) s[k] = bb_from_mont(s[k]);
#pragma unroll
  for (int k = 0; k < 4; k++) p[k] = make_uint4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
}

// leaf = overwrite-mode sponge (rate 8) over one memory row of a column-major matrix [n_cols][col_stride].
// The matrix is coset-major (row r = z * 2^log_nc + i holds the LDE point of natural index i * 2^log_b + z), the tree is
// in natural order: the digest of memory row r goes to leaf ((r mod 2^log_nc) << log_b) | (r >> log_nc).  log_b = 0: identity.
// Segment form (one proof sharded over several GPUs): only the leaves [i0 << log_b, (i0 + 2^log_ni) << log_b) are hashed, i.e. the
// memory rows z * 2^log_nc + i with i in [i0, i0 + 2^log_ni) of every coset z; thread t -> z = t >> log_ni, i = i0 + (t mod 2^log_ni).
__global__ void __launch_bounds__(128) leaf_hash_kernel(const u32* __restrict__ mat, u64 col_stride, u32 n_cols, u64 n_threads, u32 log_nc, u32 log_b,
                                                        u32 log_ni, u64 i0, u32* __restrict__ digests) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  const u64 row = ((t >> log_ni) << log_nc) + i0 + (t & ((1ull << log_ni) - 1));
  u32 s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = 0;
  const u32* p = mat + row;
  u32 c = 0;
  for (; c + 8 <= n_cols; c += 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = __ldg(p + (u64)(c + k) * col_stride);
    poseidon2_permute(s);
  }
  if (c < n_cols) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (c + k < n_cols) s[k] = __ldg(p + (u64)(c + k) * col_stride);
    poseidon2_permute(s);
  }
  const u64 leaf = ((row & ((1ull << log_nc) - 1)) << log_b) | (row >> log_nc);
  uint4* d = reinterpret_cast<uint4*>(digests + 8 * leaf);
  d[0] = make_uint4(s[0], s[1], s[2], s[3]);
  d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// Prover leaves (docs/PROVER_SPEC.md section 4.1): leaf m = hash of the 2*B consecutive natural-order rows m*2B .. m*2B + 2B - 1 of a
// coset-major matrix, i.e. the points j = 2m, 2m + 1 of the trace domain on all B cosets, concatenated (n_cols is a multiple of 8, so a
// sponge block never straddles two rows).  One thread per leaf: 2*B*n_cols/8 dependent permutations, a quarter of the tree above it.
__global__ void __launch_bounds__(128) leaf_hash_rows_kernel(const u32* __restrict__ mat, u64 col_stride, u32 n_cols, u32 log_n, u32 log_b,
                                                             u64 first, u64 count, u32* __restrict__ digests) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (t >= count) return;
  const u64 m = first + t;
  u32 s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = 0;
  const u32 rows = 2u << log_b;
  for (u32 r = 0; r < rows; r++) {
    const u64 j = 2 * m + (r >> log_b), z = r & ((1u << log_b) - 1);
    const u32* p = mat + (z << log_n) + j;
    for (u32 c = 0; c < n_cols; c += 8) {
#pragma unroll
      for (int k = 0; k < 8; k++) s[k] = __ldg(p + (u64)(c + k) * col_stride);
      poseidon2_permute(s);
    }
  }
  uint4* d = reinterpret_cast<uint4*>(digests + 8 * m);
  d[0] = make_uint4(s[0], s[1], s[2], s[3]);
  d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// FRI layer leaves: leaf i = hash(F[i] || F[i+q] || ... || F[i+(arity-1)q]) over the q leaves of a layer of arity*q ext4 values:
// arity/2 absorptions of the overwrite-mode sponge (8 words each)
__global__ void __launch_bounds__(128) leaf_hash_pairs_kernel(const uint4* __restrict__ layer, u64 q, u64 first, u64 count, u32 arity,
                                                              u32* __restrict__ digests) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= count) return;
  i += first;
  u32 s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = 0;
  for (u32 m = 0; m < arity; m += 2) {
    const uint4 a = layer[i + (u64)m * q], b = layer[i + (u64)(m + 1) * q];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    poseidon2_permute(s);
  }
  uint4* d = reinterpret_cast<uint4*>(digests + 8 * i);
  d[0] = make_uint4(s[0], s[1], s[2], s[3]);
  d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// leaves of hash_tree (docs/PROVER_SPEC.md section 2): leaf c = hash(words[8c .. 8c+8)) (the last chunk may be short), leaves beyond the
// message are all-zero digests
__global__ void __launch_bounds__(128) words_leaf_hash_kernel(const u32* __restrict__ words, u32 n_words, u32 n_leaves, u32* __restrict__ digests) {
  const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_leaves) return;
  u32 s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = 0;
  if (8 * c < n_words) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (8 * c + k < n_words) s[k] = words[8 * c + k];
    poseidon2_permute(s);
  }
  uint4* d = reinterpret_cast<uint4*>(digests + 8 * c);
  d[0] = make_uint4(s[0], s[1], s[2], s[3]);
  d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// one Merkle level: out[i] = compress(in[2i], in[2i+1])
__global__ void __launch_bounds__(128) compress_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, u64 n_out) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  u32 s[16];
#pragma unroll
  for (int k = 0; k < 4; k++) { uint4 v = in[4 * i + k]; s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
  poseidon2_permute(s);
  out[2 * i] = make_uint4(s[0], s[1], s[2], s[3]);
  out[2 * i + 1] = make_uint4(s[4], s[5], s[6], s[7]);
}
static inline unsigned nblk(u64 n, unsigned t) { return (unsigned)((n + t - 1) / t); }

int launch_permute(u32* d_states, u64 n, bool canonical_io, cudaStream_t st, u64* launches) {
  if (!n) return 0;
  permute_kernel<<<nblk(n, 128), 128, 0, st>>>(d_states, n, canonical_io);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int launch_leaf_hash(const u32* mat, u64 col_stride, u32 n_cols, u64 n_rows, u32 log_b, u32* digests, cudaStream_t st, u64* launches,
                     u64 first_leaf, u64 seg_leaves) {
  u32 log_rows = 0;
  while ((1ull << log_rows) < n_rows) log_rows++;
  const u32 log_nc = log_rows - log_b;
  u64 n_threads = n_rows, i0 = 0;
  u32 log_ni = log_nc;
  if (seg_leaves && seg_leaves < n_rows) {  // leaf range of one shard: a multiple of 2^log_b leaves, power-of-two long
    if ((seg_leaves & (seg_leaves - 1)) || (seg_leaves >> log_b) == 0 || first_leaf % seg_leaves) return -1;
    n_threads = seg_leaves; i0 = first_leaf >> log_b;
    log_ni = 0;
    while ((1ull << log_ni) < (seg_leaves >> log_b)) log_ni++;
  }
  leaf_hash_kernel<<<nblk(n_threads, 128), 128, 0, st>>>(mat, col_stride, n_cols, n_threads, log_nc, log_b, log_ni, i0, digests);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int launch_leaf_hash_rows(const u32* mat, u64 col_stride, u32 n_cols, u32 log_n, u32 log_b, u32* digests, cudaStream_t st, u64* launches,
                          u64 first_leaf, u64 n_leaves) {
  if (n_cols % 8 || log_n < 1) return -1;
  leaf_hash_rows_kernel<<<nblk(n_leaves, 128), 128, 0, st>>>(mat, col_stride, n_cols, log_n, log_b, first_leaf, n_leaves, digests);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
__global__ void merkle_coop_kernel(u32* level, u64 n_in, u64 first, u32 chunk, const u32* __restrict__ pair_layer, u32 leaf_arity, ChalState* chal,
                                   u32* root_dst, u32* sample_out, u32 n_sample);
__global__ void challenger_kernel(ChalState* st, const u32* in, u32 n_in, u32* out, u32 n_out, u32 bits);
// tree: level 0 = n_leaves digests (already in place, or computed here from the FRI layer `pair_layer` of 2*n_leaves ext4
// values); builds the upper levels behind it.  Wide levels: one thread per compression (throughput); from 16384 nodes (ZKIR_COOP_MAX) down
// the cooperative kernel (latency): 32-node blocks climbing 5 levels while the level is wide, one block at the end.
// If `chal` is given, the launch that produces the root also copies it to root_dst, observes it and samples n_sample
// elements into sample_out.
// Segment form (first, seg != 0): only the subtree above the leaves [first, first + seg) is built (seg a power of two dividing
// first); every level keeps its place in the global tree layout, so after an all-gather of the level that holds the segment
// roots the rest is an ordinary tree over n_leaves / seg nodes.  No Fiat-Shamir step in that case.
int launch_merkle_levels(u32* tree, u64 n_leaves, cudaStream_t st, u64* launches, ChalState* chal, u32* root_dst, u32* sample_out, u32 n_sample,
                         const u32* pair_layer, u64 first, u64 seg, u32 leaf_arity) {
  u32* lvl = tree;
  u64 n_tot = n_leaves, f = 0, n = n_leaves;
  static u64 coop_max = 0;   // levels wider than this use one thread per compression (ZKIR_COOP_MAX, power of two >= 256)
  if (!coop_max) { const char* e = getenv("ZKIR_COOP_MAX"); coop_max = e && atoll(e) >= 256 ? (u64)atoll(e) : 16384; }   // measured best of 1024..65536
  if (seg && seg < n_leaves) {
    if ((seg & (seg - 1)) || first % seg || first + seg > n_leaves) return -1;
    f = first; n = seg; chal = nullptr;
  }
  const bool whole = n == n_tot;
  if (pair_layer && n > coop_max) {
    leaf_hash_pairs_kernel<<<nblk(n, 128), 128, 0, st>>>(reinterpret_cast<const uint4*>(pair_layer), n_tot, f, n, leaf_arity, tree);
    (*launches)++;
    pair_layer = nullptr;
  }
  while (n > coop_max) {
    u32* nxt = lvl + n_tot * 8;
    compress_kernel<<<nblk(n / 2, 128), 128, 0, st>>>(reinterpret_cast<const uint4*>(lvl + f * 8), reinterpret_cast<uint4*>(nxt + (f / 2) * 8), n / 2);
    (*launches)++;
    lvl = nxt; n >>= 1; n_tot >>= 1; f >>= 1;
  }
  bool chal_done = false;
  while (n > 1 || pair_layer) {
    const u32 chunk = n > 128 ? 32u : (u32)n;
    const u32 blocks = (u32)(n / chunk);
    const bool last = blocks == 1;
    u32 threads = 16 * (chunk / 2);
    if (threads < 32) threads = 32;
    merkle_coop_kernel<<<blocks, threads, 0, st>>>(lvl, n_tot, f, chunk, pair_layer, leaf_arity, last && whole ? chal : nullptr, root_dst, sample_out,
                                                   n_sample);
    (*launches)++;
    pair_layer = nullptr;
    if (last && chal) chal_done = true;
    for (u32 c = chunk; c > 1; c >>= 1) { lvl += n_tot * 8; n >>= 1; n_tot >>= 1; f >>= 1; }
    if (last) break;
  }
  if (chal && !chal_done) {  // not reached: every path above ends in a single-block launch
    if (root_dst) cudaMemcpyAsync(root_dst, lvl, 32, cudaMemcpyDeviceToDevice, st);
    challenger_kernel<<<1, 32, 0, st>>>(chal, lvl, 8, sample_out, n_sample, 0);
    (*launches)++;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

u64 hash_tree_scratch_words(u32 n_words) {
  u64 chunks = (n_words + 7) / 8, leaves = 1;
  while (leaves < chunks) leaves <<= 1;
  return (2 * leaves - 1) * 8 + 8;
}
// transcript step on a long message: observe hash_tree(words) and sample; log-depth instead of n_words / 8 serial permutations
int launch_observe_hash_tree(const u32* words, u32 n_words, u32* tree_scratch, ChalState* chal, u32* sample_out, u32 n_sample, cudaStream_t st, u64* launches);
// ---------------------------------------------------------------- duplex challenger, one warp, lane i < 16 owns state[i]
__device__ __forceinline__ u32 permute_warp(u32 x, int lane) {
  const unsigned FULL = 0xffffffffu;
  const int l16 = lane & 15, grp = lane & ~3, j = lane & 3;
  auto ext_lin = [&](u32 v) {
    u32 b = __shfl_sync(FULL, v, grp | ((j + 1) & 3));
    u32 c = __shfl_sync(FULL, v, grp | ((j + 2) & 3));
    u32 d = __shfl_sync(FULL, v, grp | ((j + 3) & 3));
    u32 y = bb_add(bb_add(bb_dbl(v), bb_add(bb_dbl(b), b)), bb_add(c, d));  // 2v + 3b + c + d
    u32 s1 = bb_add(y, __shfl_xor_sync(FULL, y, 4));
    u32 s2 = bb_add(s1, __shfl_xor_sync(FULL, s1, 8));
    return bb_add(y, s2);
  };
  x = ext_lin(x);
  for (int r = 0; r < 4; r++) { x = sbox7(bb_add(x, c_rc_ext[r * 16 + l16])); x = ext_lin(x); }
  const u32 dg = c_diag[l16];
  for (int r = 0; r < ZKIR_P2_RP; r++) {
    // latency matters here, not throughput: the sum of the OLD state and the diagonal products of lanes 1..15 do not
    // depend on the S-box, so they overlap it; afterwards only (new s0 - old s0) has to be broadcast.
    u32 s = x;
    s = bb_add(s, __shfl_xor_sync(FULL, s, 1));
    s = bb_add(s, __shfl_xor_sync(FULL, s, 2));
    s = bb_add(s, __shfl_xor_sync(FULL, s, 4));
    s = bb_add(s, __shfl_xor_sync(FULL, s, 8));
    const u32 x0 = sbox7(bb_add(x, c_rc_int[r]));                       // meaningful in lane 0 of each state
    const u32 delta = __shfl_sync(FULL, bb_sub(x0, x), lane & ~15);
    const u32 xi = l16 == 0 ? x0 : x;
    x = bb_add(bb_mul(xi, dg), bb_add(s, delta));
  }
  for (int r = 4; r < 8; r++) { x = sbox7(bb_add(x, c_rc_ext[r * 16 + l16])); x = ext_lin(x); }
  return x;
}

// observe n_in field elements (Montgomery) then sample n_out;  bits > 0: outputs are canonical integers masked to
// `bits` bits (query indices / PoW check), else Montgomery field elements.  One full warp must call this.
__device__ __forceinline__ void challenger_step(ChalState* st, const u32* in, u32 n_in, u32* out, u32 n_out, u32 bits, int lane) {
  u32 x = lane < 16 ? st->sponge[lane] : 0;
  u32 nin = st->n_in, nout = st->n_out;
  u32 inb = lane < 8 ? st->inbuf[lane] : 0;    // lane k < 8 holds input buffer slot k
  u32 outb = lane < 8 ? st->outbuf[lane] : 0;  // lane k < 8 holds output buffer slot k
  auto duplex = [&]() {
    if (lane < (int)nin) x = inb;
    nin = 0;
    x = permute_warp(x, lane);
    outb = x;  // lanes 0..7 matter
    nout = 8;
  };
  for (u32 i = 0; i < n_in; i++) {
    u32 v = in[i];
    nout = 0;
    if (lane == (int)nin) inb = v;
    nin++;
    if (nin == 8) duplex();
  }
  for (u32 i = 0; i < n_out; i++) {
    if (nin > 0 || nout == 0) duplex();
    nout--;
    u32 v = __shfl_sync(0xffffffffu, outb, nout);
    if (bits) { v = bb_from_mont(v); if (bits < 32) v &= (1u << bits) - 1; }
    if (lane == 0) out[i] = v;
  }
  __syncwarp();
  if (lane < 16) st->sponge[lane] = x;
  if (lane < 8) { st->inbuf[lane] = inb; st->outbuf[lane] = outb; }
  if (lane == 0) { st->n_in = nin; st->n_out = nout; }
}
__global__ void __launch_bounds__(32) challenger_kernel(ChalState* st, const u32* in, u32 n_in, u32* out, u32 n_out, u32 bits) {
  challenger_step(st, in, n_in, out, n_out, bits, threadIdx.x);
}

// Upper part of a Merkle tree, latency-oriented: 16 lanes cooperate on one compression (permute_warp); a block of
// 16 * chunk / 2 threads owns `chunk` consecutive nodes of the input level and climbs log2(chunk) levels through shared
// memory, writing every level to the global tree (authentication paths need them).  A thread-per-permutation level costs
// one full permutation latency (~9 us) however small it is; this kernel costs ~3 us per level, and wide levels are cut
// into many 32-node blocks so that all SMs share them.
// Leaf mode (pair_layer != nullptr): the input level does not exist yet; node i is first computed as the FRI leaf
// hash(f[i] || f[i + h] || ... || f[i + (arity-1)h]) of the ext4 layer and stored as level 0.
// When the block reaches the root (single block), warp 0 also runs the Fiat-Shamir step that always follows:
// copy the root into the proof, observe it, sample `n_sample` field elements.
#define COOP_MAX_CHUNK 128
__global__ void __launch_bounds__(16 * COOP_MAX_CHUNK / 2) merkle_coop_kernel(u32* level, u64 n_in, u64 first, u32 chunk,
                                                                             const u32* __restrict__ pair_layer, u32 leaf_arity, ChalState* chal,
                                                                             u32* root_dst, u32* sample_out, u32 n_sample) {
  __shared__ u32 buf[2][COOP_MAX_CHUNK * 8];
  const u32 tid = threadIdx.x, lane = tid & 31, l16 = tid & 15, slot = tid >> 4, slots = blockDim.x >> 4;
  // `level` is the start of a whole tree level of n_in nodes; the grid covers the nodes [first, first + gridDim.x * chunk)
  const u64 node0 = first + (u64)blockIdx.x * chunk;
  if (pair_layer) {
    const u64 h = n_in;  // leaves of this layer = its length / arity; the values of leaf i sit at i, i + h, ..., i + (arity-1)h
    for (u32 s0 = 0; s0 < chunk; s0 += slots) {
      if (s0 + (tid >> 5) * 2 >= chunk) break;
      const u32 sidx = s0 + slot;
      const bool active = sidx < chunk;
      u32 x = 0;
      for (u32 m = 0; m < leaf_arity; m += 2) {   // one absorption of the overwrite-mode sponge per pair of values
        if (active && l16 < 8) x = pair_layer[4 * (node0 + sidx + (u64)(m + (l16 >= 4 ? 1u : 0u)) * h) + (l16 & 3)];
        x = permute_warp(x, lane);
      }
      if (active && l16 < 8) { buf[0][8 * sidx + l16] = x; level[(node0 + sidx) * 8 + l16] = x; }
    }
  } else {
    const u32* src = level + node0 * 8;
    for (u32 i = tid; i < chunk * 8; i += blockDim.x) buf[0][i] = src[i];
  }
  __syncthreads();
  u32 cur = 0, n = chunk;
  u64 level_n = n_in, out0 = node0;
  u32* out_base = level + n_in * 8;
  while (n > 1) {
    const u32 n_out = n >> 1;
    out0 >>= 1;                                   // first node of this block in the level being written
    for (u32 s0 = 0; s0 < n_out; s0 += slots) {   // warp-uniform trip count
      const u32 sidx = s0 + slot;
      if (s0 + (tid >> 5) * 2 >= n_out) break;  // neither half of this warp has a node: warp-uniform exit
      const bool active = sidx < n_out;
      u32 x = active ? buf[cur][16 * sidx + l16] : 0u;
      x = permute_warp(x, lane);
      if (active && l16 < 8) {
        buf[cur ^ 1][8 * sidx + l16] = x;
        out_base[(out0 + sidx) * 8 + l16] = x;
      }
    }
    __syncthreads();
    cur ^= 1; n = n_out;
    level_n >>= 1;
    out_base += level_n * 8;
  }
  if (chal != nullptr && gridDim.x == 1 && tid < 32) {
    if (root_dst && lane < 8) root_dst[lane] = buf[cur][lane];
    challenger_step(chal, buf[cur], 8, sample_out, n_sample, 0, lane);
  }
}

// proof of work: smallest canonical w such that, after observe(w), sample_bits(bits) == 0.
__global__ void pow_grind_kernel(const ChalState* st, u32 bits, u32* result /* init 0xffffffff */) {
  __shared__ u32 base[16];
  __shared__ u32 s_nin;
  if (threadIdx.x < 16) base[threadIdx.x] = st->sponge[threadIdx.x];
  if (threadIdx.x == 0) s_nin = st->n_in;
  __syncthreads();
  const u32 nin = s_nin;  // < 8 pending inputs
  if (threadIdx.x < nin) base[threadIdx.x] = st->inbuf[threadIdx.x];
  __syncthreads();
  const u64 G = (u64)gridDim.x * blockDim.x;
  const u32 mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1);
  for (u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x; w < BB_P; w += G) {
    const u64 sweep_end = (w / G + 1) * G;
    u32 s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = base[k];
    s[nin] = bb_to_mont((u32)w);
    poseidon2_permute(s);
    // observe(w) fills slot nin; if that makes 8 the duplex already ran and sample pops out[7]; otherwise sample
    // triggers the duplex.  Either way exactly one permutation and the popped element is state[7].
    if ((bb_from_mont(s[7]) & mask) == 0) atomicMin(result, (u32)w);
    if (*(volatile u32*)result < sweep_end) break;  // every smaller candidate lives in a sweep that is complete
  }
}

int launch_observe_hash_tree(const u32* words, u32 n_words, u32* tree_scratch, ChalState* chal, u32* sample_out, u32 n_sample, cudaStream_t st, u64* launches) {
  u32 chunks = (n_words + 7) / 8, leaves = 1;
  while (leaves < chunks) leaves <<= 1;
  words_leaf_hash_kernel<<<nblk(leaves, 128), 128, 0, st>>>(words, n_words, leaves, tree_scratch);
  (*launches)++;
  if (leaves == 1) {   // the single leaf digest is the root
    challenger_kernel<<<1, 32, 0, st>>>(chal, tree_scratch, 8, sample_out, n_sample, 0);
    (*launches)++;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  return launch_merkle_levels(tree_scratch, leaves, st, launches, chal, nullptr, sample_out, n_sample);
}
int launch_challenger(ChalState* st_dev, const u32* in, u32 n_in, u32* out, u32 n_out, u32 bits, cudaStream_t st, u64* launches) {
  challenger_kernel<<<1, 32, 0, st>>>(st_dev, in, n_in, out, n_out, bits);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int launch_pow_grind(const ChalState* st_dev, u32 bits, u32* result, cudaStream_t st, u64* launches) {
  cudaMemsetAsync(result, 0xff, 4, st);
  pow_grind_kernel<<<148 * 4, 256, 0, st>>>(st_dev, bits, result);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace zkir
