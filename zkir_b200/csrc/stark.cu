// The remaining trace->proof kernels: out-of-domain openings, DEEP combination (the FRI input codeword), FRI
// folding, query gathering and representation changes.  All field data is Montgomery form on the device; ext4
// elements are stored AoS (16 B, uint4 loads).  No reference counterpart; the maths is
// docs/PROVER_SPEC.md sections "Openings", "FRI" and "Queries".
#include <cuda_runtime.h>
#include <stdlib.h>
#include "bb.cuh"
#include "kernels.h"

namespace zkir {

static inline unsigned nblk(u64 n, unsigned t) { return (unsigned)((n + t - 1) / t); }
#define CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? 0 : -2)

__global__ void map_kernel(u32* dst, const u32* src, u64 n, int to_mont) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = to_mont ? bb_to_mont(src[i]) : bb_from_mont(src[i]);
}
int launch_map(u32* dst, const u32* src, u64 n, int to_mont, cudaStream_t st, u64* launches) {
  if (!n) return 0;
  map_kernel<<<nblk(n, 256), 256, 0, st>>>(dst, src, n, to_mont);
  (*launches)++;
  return CHECK_LAUNCH();
}

__device__ __forceinline__ E4 ld_e4(const E4* p) { uint4 v = *reinterpret_cast<const uint4*>(p); E4 r; r.c[0] = v.x; r.c[1] = v.y; r.c[2] = v.z; r.c[3] = v.w; return r; }
__device__ __forceinline__ void st_e4(E4* p, E4 v) { *reinterpret_cast<uint4*>(p) = make_uint4(v.c[0], v.c[1], v.c[2], v.c[3]); }
__device__ E4 e4_pow_dev(E4 a, u64 e) { E4 r = e4_one(); while (e) { if (e & 1) r = e4_mul(r, a); a = e4_mul(a, a); e >>= 1; } return r; }

// out[pos] = (base * mul)^(k(pos)): k = coefficient index stored at memory position pos under the digit plan (the
// coefficient vectors of ntt_fast.cu are digit-reversed; nd = 1 means natural order).  Each thread owns a run of
// consecutive positions inside one row of the lowest digit, where k advances by 2^(sum of the upper digits).
#define EP_CHUNK 64
__global__ void ext_powers_kernel(const u32* base_ext, u32 mul_const, FastPlan plan, u32 chunk, E4* out, u64 n) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u64 p0 = t * chunk;
  if (p0 >= n) return;
  E4 u; for (int k = 0; k < 4; k++) u.c[k] = bb_mul(base_ext[k], mul_const);
  // k(pos): digits top->bottom x_1..x_D, k = x_1 + 2^d1 x_2 + ...
  int below = 0, upper = 0;
  for (int j = 0; j < plan.nd; j++) below += plan.d[j];
  for (int j = 0; j + 1 < plan.nd; j++) upper += plan.d[j];
  u64 k0 = 0; int wshift = 0;
  for (int j = 0; j < plan.nd; j++) {
    below -= plan.d[j];
    k0 |= ((p0 >> below) & ((1ull << plan.d[j]) - 1)) << wshift;
    wshift += plan.d[j];
  }
  E4 cur = e4_pow_dev(u, k0);
  const E4 step = e4_pow_dev(u, 1ull << upper);
  for (u32 j = 0; j < chunk; j++) { st_e4(out + p0 + j, cur); cur = e4_mul(cur, step); }
}
int launch_ext_powers(const u32* base_ext, u32 mul_const, const FastPlan& plan, E4* out, cudaStream_t st, u64* launches) {
  const u64 n = 1ull << plan.log_n;
  const u64 m = 1ull << plan.d[plan.nd - 1];
  const u32 chunk = (u32)(m < EP_CHUNK ? m : EP_CHUNK);
  ext_powers_kernel<<<nblk(n / chunk, 128), 128, 0, st>>>(base_ext, mul_const, plan, chunk, out, n);
  (*launches)++;
  return CHECK_LAUNCH();
}

// ---- openings: out1[k] = sum_j coef[k][j]*U1[j], out2[k] = sum_j coef[k][j]*U2[j]
// Block = 8 warps x OPEN_COLS columns over OPEN_ROWS rows: lane = row (128 B coalesced column reads), warp = column
// group, so the eight warps read the SAME 32 B/row of U1/U2 and seven of them hit L1.  Lazy 64-bit accumulators (bb.cuh).
// The kernel is latency-bound (few warps, long dependent accumulation chains): one column per warp at 64 registers and
// 4 CTAs/SM measured faster than two columns per warp at 80 registers and 2 CTAs/SM (ZKIR_OPEN_VARIANT=1).
#define OPEN_WARPS 8
#define OPEN_THREADS (32 * OPEN_WARPS)
#define OPEN_ROWS_MIN 512   // smallest row chunk of any variant: sizes the partial-sum scratch
template <int OPEN_COLS, int OPEN_ROWS, int UNROLL, int MINB>
__global__ void __launch_bounds__(OPEN_THREADS, MINB) open_partial_kernel(const u32* __restrict__ coef, u64 col_stride, u32 n_cols, u64 n,
                                                                      const E4* __restrict__ U1, const E4* __restrict__ U2,
                                                                      E4* partial, u32 n_chunks) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, chunk = blockIdx.x;
  const u32 k0 = (blockIdx.y * OPEN_WARPS + warp) * OPEN_COLS;
  if (k0 >= n_cols) return;
  Acc4 l1[OPEN_COLS], l2[OPEN_COLS];
#pragma unroll
  for (int c = 0; c < OPEN_COLS; c++) { l1[c] = acc4_zero(); l2[c] = acc4_zero(); }
  const u64 j0 = (u64)chunk * OPEN_ROWS + lane;
#pragma unroll UNROLL
  for (int r = 0; r < OPEN_ROWS / 32; r++) {
    const u64 j = j0 + (u64)r * 32;
    if (j < n) {
      const E4 u1 = ld_e4(U1 + j), u2 = ld_e4(U2 + j);
#pragma unroll
      for (int c = 0; c < OPEN_COLS; c++) {
        if (k0 + c < n_cols) {
          const u32 v = __ldg(coef + (u64)(k0 + c) * col_stride + j);
          acc4_mac(l1[c], u1, v);
          acc4_mac(l2[c], u2, v);
        }
      }
    }
    if (r & 1) {
#pragma unroll
      for (int c = 0; c < OPEN_COLS; c++) { acc4_fix(l1[c]); acc4_fix(l2[c]); }
    }
  }
#pragma unroll
  for (int c = 0; c < OPEN_COLS; c++) {
    const E4 a1 = acc4_finish(l1[c]), a2 = acc4_finish(l2[c]);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      u32 x = a1.c[k], y = a2.c[k];
      for (int o = 16; o > 0; o >>= 1) { x = bb_add(x, __shfl_xor_sync(0xffffffffu, x, o)); y = bb_add(y, __shfl_xor_sync(0xffffffffu, y, o)); }
      if (lane == 0 && k0 + c < n_cols) {
        partial[((u64)(k0 + c)) * n_chunks + chunk].c[k] = x;
        partial[((u64)n_cols + k0 + c) * n_chunks + chunk].c[k] = y;
      }
    }
  }
}
__global__ void open_final_kernel(const E4* partial, u32 n_cols, u32 n_chunks, E4* out1, E4* out2) {
  // one warp per (which, column)
  const u32 id = blockIdx.x;  // which * n_cols + k
  E4 acc = e4_zero();
  for (u32 c = threadIdx.x; c < n_chunks; c += 32) acc = e4_add(acc, ld_e4(partial + (u64)id * n_chunks + c));
  for (int k = 0; k < 4; k++) {
    u32 x = acc.c[k];
    for (int o = 16; o > 0; o >>= 1) x = bb_add(x, __shfl_xor_sync(0xffffffffu, x, o));
    acc.c[k] = x;
  }
  if (threadIdx.x == 0) { if (id < n_cols) st_e4(out1 + id, acc); else st_e4(out2 + (id - n_cols), acc); }
}
template <int COLS, int ROWS, int UNROLL, int MINB>
static void launch_open_t(const u32* coef, u64 col_stride, u32 n_cols, u64 n, const E4* U1, const E4* U2, E4* partial, u32* n_chunks_out, cudaStream_t st) {
  const u32 n_chunks = (u32)((n + ROWS - 1) / ROWS);
  dim3 grid(n_chunks, (n_cols + COLS * OPEN_WARPS - 1) / (COLS * OPEN_WARPS));
  open_partial_kernel<COLS, ROWS, UNROLL, MINB><<<grid, OPEN_THREADS, 0, st>>>(coef, col_stride, n_cols, n, U1, U2, partial, n_chunks);
  *n_chunks_out = n_chunks;
}
int launch_open(const u32* coef, u64 col_stride, u32 n_cols, u64 n, const E4* U1, const E4* U2, E4* out1, E4* out2,
                E4* partial_scratch, cudaStream_t st, u64* launches) {
  static int variant = -1;   // ZKIR_OPEN_VARIANT: experiments, 0 = default
  if (variant < 0) { const char* e = getenv("ZKIR_OPEN_VARIANT"); variant = e ? atoi(e) : 0; }
  u32 n_chunks = 0;
  switch (variant) {   // measured on the 2^20-row proof (openings stage): 0.80 / 0.98 / 0.81 ms
    case 1: launch_open_t<2, 2048, 4, 2>(coef, col_stride, n_cols, n, U1, U2, partial_scratch, &n_chunks, st); break;   // two columns per warp, 80 registers
    case 2: launch_open_t<1, 1024, 4, 4>(coef, col_stride, n_cols, n, U1, U2, partial_scratch, &n_chunks, st); break;
    default: launch_open_t<1, 2048, 4, 4>(coef, col_stride, n_cols, n, U1, U2, partial_scratch, &n_chunks, st); break;  // one column per warp, 64 registers, 4 CTAs/SM
  }
  open_final_kernel<<<2 * n_cols, 32, 0, st>>>(partial_scratch, n_cols, n_chunks, out1, out2);
  (*launches) += 2;
  return CHECK_LAUNCH();
}
u64 open_scratch_elems(u32 n_cols, u64 n) { return 2ull * n_cols * ((n + OPEN_ROWS_MIN - 1) / OPEN_ROWS_MIN); }

// ---- DEEP combination
// scratch layout: afp[0..width) = alpha^k, then [width+0]=A1, +1=A2, +2=A3, +3=alpha^W, +4=alpha^2W
__global__ void __launch_bounds__(128) deep_prep_kernel(DeepArgs a) {
  // one block: thread k owns alpha^k (square-and-multiply) and its three products; exact field sums, any order
  __shared__ E4 r1[128], r2[128], r3[128];
  const u32 W = a.width, t = threadIdx.x;
  E4 al; for (int k = 0; k < 4; k++) al.c[k] = a.alpha_fri[k];
  E4 s1 = e4_zero(), s2 = e4_zero(), s3 = e4_zero();
  for (u32 k = t; k < W; k += blockDim.x) {
    const E4 pw = e4_pow_dev(al, k);
    a.afp_scratch[k] = pw;
    s1 = e4_add(s1, e4_mul(pw, a.open_t[k]));
    s2 = e4_add(s2, e4_mul(pw, a.open_tg[k]));
    if (k < a.qwidth) s3 = e4_add(s3, e4_mul(pw, a.open_q[k]));
  }
  r1[t] = s1; r2[t] = s2; r3[t] = s3;
  __syncthreads();
  for (u32 h = blockDim.x / 2; h > 0; h >>= 1) {
    if (t < h) { r1[t] = e4_add(r1[t], r1[t + h]); r2[t] = e4_add(r2[t], r2[t + h]); r3[t] = e4_add(r3[t], r3[t + h]); }
    __syncthreads();
  }
  if (t == 0) { a.afp_scratch[W] = r1[0]; a.afp_scratch[W + 1] = r2[0]; a.afp_scratch[W + 2] = r3[0]; }
  if (t == 1) a.afp_scratch[W + 3] = e4_pow_dev(al, W);
  if (t == 2) a.afp_scratch[W + 4] = e4_pow_dev(al, 2ull * W);
}
__global__ void __launch_bounds__(128) deep_kernel(DeepArgs a) {
  extern __shared__ E4 afp[];  // width + 5
  for (u32 i = threadIdx.x; i < a.width + 5; i += blockDim.x) afp[i] = a.afp_scratch[i];
  __syncthreads();
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  const u32 log_nj = a.seg_log_nj == 0xffffffffu ? a.log_n : a.seg_log_nj;
  if (t >= (1ull << (log_nj + a.log_b))) return;
  const u64 i = ((t >> log_nj) << a.log_n) | (a.seg_j0 + (t & ((1ull << log_nj) - 1)));  // memory row (coset-major)
  const u64 nat = ((i & ((1ull << a.log_n) - 1)) << a.log_b) | (i >> a.log_n);
  Acc4 lt = acc4_zero(), lq = acc4_zero();  // lazy accumulators, fixed every second term
#pragma unroll 8
  for (u32 k = 0; k < a.width; k++) {
    acc4_mac(lt, afp[k], __ldg(a.lde + (u64)k * a.M + i));
    if (k & 1) acc4_fix(lt);
  }
#pragma unroll 4
  for (u32 k = 0; k < a.qwidth; k++) {
    acc4_mac(lq, afp[k], __ldg(a.qlde + (u64)k * a.M + i));
    if (k & 1) acc4_fix(lq);
  }
  const E4 rt = acc4_finish(lt), rq = acc4_finish(lq);
  const u32 W = a.width;
  const u32 x = a.xs[i];
  E4 z; for (int k = 0; k < 4; k++) z.c[k] = a.zeta[k];
  E4 gz = e4_mulb(z, a.g_mont);
  E4 dz = e4_sub(e4_from_base(x), z), dgz = e4_sub(e4_from_base(x), gz);
  E4 iz = e4_inv(dz), igz = e4_inv(dgz);
  E4 f = e4_mul(e4_sub(rt, afp[W]), iz);
  f = e4_add(f, e4_mul(afp[W + 3], e4_mul(e4_sub(rt, afp[W + 1]), igz)));
  f = e4_add(f, e4_mul(afp[W + 4], e4_mul(e4_sub(rq, afp[W + 2]), iz)));
  st_e4(a.out + nat, f);
}
int launch_deep(const DeepArgs& a, cudaStream_t st, u64* launches) {
  deep_prep_kernel<<<1, 128, 0, st>>>(a);
  const u64 n_threads = a.seg_log_nj == 0xffffffffu ? a.M : (1ull << (a.seg_log_nj + a.log_b));
  deep_kernel<<<nblk(n_threads, 128), 128, (a.width + 5) * sizeof(E4), st>>>(a);
  (*launches) += 2;
  return CHECK_LAUNCH();
}

// ---- FRI fold: out[i] = (f[i]+f[i+h])/2 + beta * (f[i]-f[i+h]) * c * w^-i,  c = 1/(2*shift_r)
__global__ void __launch_bounds__(256) fri_fold_kernel(const E4* __restrict__ in, E4* __restrict__ out, u64 h, const u32* beta_dev,
                                                      const u32* __restrict__ inv_w, u32 tw_stride, u32 c_mont, int square_beta) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= h) return;
  E4 beta; for (int k = 0; k < 4; k++) beta.c[k] = beta_dev[k];
  for (int k = 0; k < square_beta; k++) beta = e4_mul(beta, beta);   // beta^(2^square_beta): later half-folds of a fold-by-4/8 round
  const E4 a = ld_e4(in + i), b = ld_e4(in + i + h);
  const u32 half = bb_to_mont_c((BB_P + 1) / 2);
  E4 s = e4_mulb(e4_add(a, b), half);
  E4 d = e4_mulb(e4_sub(a, b), bb_mul(c_mont, inv_w[i * tw_stride]));
  st_e4(out + i, e4_add(s, e4_mul(beta, d)));
}
int launch_fri_fold(const E4* in, E4* out, u64 h, const u32* beta_dev, const u32* inv_w_table, u32 tw_stride, u32 c_mont,
                    cudaStream_t st, u64* launches, int square_beta) {
  fri_fold_kernel<<<nblk(h, 256), 256, 0, st>>>(in, out, h, beta_dev, inv_w_table, tw_stride, c_mont, square_beta);
  (*launches)++;
  return CHECK_LAUNCH();
}

// One committed FRI round = LA half-folds (fold by 2^LA) in ONE launch: thread i of the qn = n >> LA outputs loads the 2^LA values
// f[i + k*qn] -- exactly the values one Merkle leaf of the round holds -- and applies the half-folds of fri_fold_kernel level by level in
// registers (same operations in the same order: bit-identical), so the intermediate layers are never written and a round costs one
// launch instead of LA.  Level s pairs position j with j + (n >> (s+1)), challenge beta^(2^s), twiddle inv_w[j * (tw_stride << s)].
struct FoldConsts { u32 c_mont[3]; };
template <int LA>
__global__ void __launch_bounds__(256) fri_fold_multi_kernel(const E4* __restrict__ in, E4* __restrict__ out, u64 qn, const u32* beta_dev,
                                                            const u32* __restrict__ inv_w, u32 tw_stride, FoldConsts fc) {
  const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= qn) return;
  E4 beta; for (int k = 0; k < 4; k++) beta.c[k] = beta_dev[k];
  E4 v[1 << LA];
#pragma unroll
  for (int k = 0; k < (1 << LA); k++) v[k] = ld_e4(in + i + (u64)k * qn);
  const u32 half = bb_to_mont_c((BB_P + 1) / 2);
#pragma unroll
  for (int s = 0; s < LA; s++) {
    const int hk = 1 << (LA - s - 1);   // pairs (k, k + hk) of this thread's values
#pragma unroll
    for (int k = 0; k < hk; k++) {
      const E4 a = v[k], b = v[k + hk];
      const u64 j = i + (u64)k * qn;
      const E4 sm = e4_mulb(e4_add(a, b), half);
      const E4 d = e4_mulb(e4_sub(a, b), bb_mul(fc.c_mont[s], inv_w[j * ((u64)tw_stride << s)]));
      v[k] = e4_add(sm, e4_mul(beta, d));
    }
    beta = e4_mul(beta, beta);
  }
  st_e4(out + i, v[0]);
}
int launch_fri_fold_multi(const E4* in, E4* out, u64 n, u32 la, const u32* beta_dev, const u32* inv_w_table, u32 tw_stride, const u32 c_mont[3],
                          cudaStream_t st, u64* launches) {
  FoldConsts fc;
  for (int k = 0; k < 3; k++) fc.c_mont[k] = c_mont[k];
  const u64 qn = n >> la;
  if (la == 3) fri_fold_multi_kernel<3><<<nblk(qn, 256), 256, 0, st>>>(in, out, qn, beta_dev, inv_w_table, tw_stride, fc);
  else if (la == 2) fri_fold_multi_kernel<2><<<nblk(qn, 256), 256, 0, st>>>(in, out, qn, beta_dev, inv_w_table, tw_stride, fc);
  else if (la == 1) fri_fold_multi_kernel<1><<<nblk(qn, 256), 256, 0, st>>>(in, out, qn, beta_dev, inv_w_table, tw_stride, fc);
  else return -1;
  (*launches)++;
  return CHECK_LAUNCH();
}

// ---- query gather: one block per query, everything copied into the proof at fixed offsets
__device__ __forceinline__ void copy_path(const u32* tree, u64 n_leaves, u32 log_leaves, u64 idx, u32* out, u32 sl, bool own_leaf, bool own_top) {
  for (u32 t = threadIdx.x; t < log_leaves * 8; t += blockDim.x) {
    const u32 lvl = t >> 3, w = t & 7;
    const u64 off = 2 * n_leaves - 2 * (n_leaves >> lvl);  // node offset of level lvl
    const bool mine = lvl < sl ? own_leaf : own_top;
    out[t] = mine ? tree[(off + ((idx >> lvl) ^ 1)) * 8 + w] : 0u;
  }
}
__global__ void __launch_bounds__(128) query_kernel(QueryArgs a) {
  const u32 qi = blockIdx.x;
  const u64 M = 1ull << a.log_m;
  const u64 q = a.indices[qi];
  const u32 log_b = a.log_m - a.log_n;
  const bool top = a.shard_lo == 0;  // replicated data is contributed by the context that owns segment 0
  auto owns = [&](u64 leaf, u32 sl) { const u64 o = leaf >> sl; return sl == 0 ? top : (o >= a.shard_lo && o < a.shard_hi); };
  u32* out = a.out + (u64)qi * a.words_per_query;
  // matrix trees: leaf m = q >> log_lr holds 2^log_lr consecutive natural rows, all opened in natural order
  const u64 m = q >> a.log_lr, leaves = M >> a.log_lr;
  const u32 lr = 1u << a.log_lr, depth = a.log_m - a.log_lr;
  auto mrow = [&](u32 r) { const u64 nat = (m << a.log_lr) + r; return ((nat & ((1ull << log_b) - 1)) << a.log_n) | (nat >> log_b); };
  const bool row_mine = a.lde_sl ? owns(m, a.lde_sl) : top;
  for (u32 k = threadIdx.x; k < lr * a.width; k += blockDim.x) out[k] = row_mine ? a.lde[(u64)(k % a.width) * M + mrow(k / a.width)] : 0u;
  out += lr * a.width;
  copy_path(a.ttree, leaves, depth, m, out, a.ttree_sl, owns(m, a.ttree_sl), top); out += depth * 8;
  // aux columns (LogUp helpers and running sum): columns [width, width + aux_width) of the same matrix, committed in their own tree
  for (u32 k = threadIdx.x; k < lr * a.aux_width; k += blockDim.x) out[k] = row_mine ? a.lde[(u64)(a.width + k % a.aux_width) * M + mrow(k / a.aux_width)] : 0u;
  out += lr * a.aux_width;
  if (a.aux_width) { copy_path(a.atree, leaves, depth, m, out, a.atree_sl, owns(m, a.atree_sl), top); out += depth * 8; }
  const bool qrow_mine = a.qlde_sl ? owns(m, a.qlde_sl) : top;
  for (u32 k = threadIdx.x; k < lr * 8; k += blockDim.x) out[k] = qrow_mine ? a.qlde[(u64)(k & 7) * M + mrow(k >> 3)] : 0u;
  out += lr * 8;
  copy_path(a.qtree, leaves, depth, m, out, a.qtree_sl, owns(m, a.qtree_sl), top); out += depth * 8;
  u32 level = 0;
  for (u32 t = 0; t < a.fri_rounds; t++) {
    const u32 la = t < a.fold8_rounds ? 3 : a.last_log_arity;  // log2 of the fold arity of this round
    const u64 qn = (M >> level) >> la, i = q & (qn - 1);       // leaves of the layer; opened values at i + k*qn
    const u32* lay = reinterpret_cast<const u32*>(a.layers[level]);
    if (threadIdx.x < (4u << la)) out[threadIdx.x] = top ? lay[4 * (i + (threadIdx.x >> 2) * qn) + (threadIdx.x & 3)] : 0u;
    out += 4u << la;
    const u32 depth = a.log_m - level - la;
    copy_path(a.ltrees[level], qn, depth, i, out, a.layer_sl[t], owns(i, a.layer_sl[t]), top); out += depth * 8;
    level += la;
  }
}
int launch_queries(const QueryArgs& a, cudaStream_t st, u64* launches) {
  if (!a.num_queries) return 0;
  query_kernel<<<a.num_queries, 128, 0, st>>>(a);
  (*launches)++;
  return CHECK_LAUNCH();
}

}  // namespace zkir
