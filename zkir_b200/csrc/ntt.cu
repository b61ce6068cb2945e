// Batched BabyBear NTT / iNTT / LDE for sm_100a: multi-pass "four-step" decomposition, each pass a shared-memory
// tile of R points x T lanes with register radix-8 butterflies.
//
// No reference counterpart exists (the reference has no NTT; its only field is Mersenne31,
// zkir-spec/src/field.rs:15-23).  This implements docs/PROVER_SPEC.md section "LDE": natural-order input,
// natural-order output, forward root w_n = ZKIR_BB_ROOTS[log n], coset shift 31.
//
// Decomposition of a length-n transform, n = R1*m1:
//   pass A (strided):   for every l in [0,m1): size-R1 DFT over x[j1*m1 + l]; multiply by w_n^(k1*l); in place.
//   then every row k1 (m1 contiguous elements) needs a size-m1 DFT whose output index k' lands at X[k1 + R1*k'];
//   either recurse once more (pass B, in place inside the row) or finish with the transposing pass C.
// Data layout in HBM: column-major [n_cols][n]; lanes of a tile are 32 (16) consecutive addresses, so every global
// access is a full 128 B (64 B) segment.  Algorithmic bytes: 8*n*C per transform.
#include <cuda_runtime.h>
#include <stdio.h>
#include <map>
#include <vector>
#include "bb.cuh"
#include "kernels.h"
#include "constants_generated.h"

namespace zkir {

struct PassParams {
  const u32* in;
  u32* out;
  u32 tiles_a, tiles_b;         // tiles per column = tiles_a * tiles_b
  u32 lanes_valid;              // number of valid lanes along the tile axis that carries lanes (for ragged column counts)
  u64 in_col, out_col;          // column strides (elements)
  u64 in_a, in_b, in_t, in_r;   // input address = col*in_col + a*in_a + b*in_b + t*in_t + r*in_r
  u64 out_a, out_b, out_t, out_k;
  const u32* tw_small;          // R/2 entries: w_R^e (or inverse), Montgomery
  const u32* tw_big;            // optional, index a*tw_a + b*tw_b + t*tw_t + k*tw_k
  u64 tw_a, tw_b, tw_t, tw_k;
  const u32* in_scale;          // optional table indexed like the input address within the column
  const u32* out_scale;         // optional table indexed like the output address within the column
  u32 out_const;                // Montgomery constant multiplied into every output (BB_ONE = none)
  u32 use_out_const;
  u32 r_nonzero;                // input rows r >= r_nonzero are implicit zeros (zero padding of an LDE)
  u32 lanes_are_cols;           // single-pass mode: lane t is column a*T+t
};

template <int KK>
__device__ __forceinline__ void radix_group(u32* x, const u32* tws, u32 q, u32 stride, int tw_shift) {
  // DIF butterflies on 2^KK register values; x[i] sits at row (base + i*stride) of a block of size (stride << KK)
#pragma unroll
  for (int s = 0; s < KK; s++) {
    const int half = 1 << (KK - 1 - s);
#pragma unroll
    for (int i = 0; i < (1 << KK); i++) {
      if ((i & half) == 0) {
        const u32 o = q + (u32)(i & (half - 1)) * stride;  // offset inside the current sub-block
        const u32 tw = tws[o << (tw_shift + s)];
        const u32 a = x[i], b = x[i + half];
        x[i] = bb_add(a, b);
        x[i + half] = bb_mul(bb_sub(a, b), tw);
      }
    }
  }
}

template <int LOG_R, int LOG_T, int LOG_B, int NT>
__device__ __forceinline__ void dif_rounds(u32* tile, const u32* tws) {
  constexpr int T = 1 << LOG_T, PITCH = T + 1, R = 1 << LOG_R;
  if constexpr (LOG_B > 0) {
    constexpr int KK = LOG_B >= 3 ? 3 : LOG_B;
    constexpr u32 STRIDE = 1u << (LOG_B - KK);
    constexpr int ITEMS = (R >> KK) * T;
    for (int item = threadIdx.x; item < ITEMS; item += NT) {
      const u32 t = item & (T - 1), g = item >> LOG_T;
      const u32 q = g & (STRIDE - 1), blk = g >> (LOG_B - KK);
      const u32 base = (blk << LOG_B) + q;
      u32 x[1 << KK];
#pragma unroll
      for (int i = 0; i < (1 << KK); i++) x[i] = tile[(base + i * STRIDE) * PITCH + t];
      radix_group<KK>(x, tws, q, STRIDE, LOG_R - LOG_B);
#pragma unroll
      for (int i = 0; i < (1 << KK); i++) tile[(base + i * STRIDE) * PITCH + t] = x[i];
    }
    __syncthreads();
    dif_rounds<LOG_R, LOG_T, LOG_B - KK, NT>(tile, tws);
  }
}

template <int LOG_R, int LOG_T, int NT>
__global__ void __launch_bounds__(NT) ntt_pass_kernel(PassParams p) {
  constexpr int R = 1 << LOG_R, T = 1 << LOG_T, PITCH = T + 1;
  extern __shared__ u32 smem[];
  u32* tile = smem;                 // R * PITCH
  u32* tws = smem + R * PITCH;      // R/2 (at least 1)
  const u32 tiles_per_col = p.tiles_a * p.tiles_b;
  const u32 col = p.lanes_are_cols ? 0 : blockIdx.x / tiles_per_col;
  const u32 tile_id = p.lanes_are_cols ? blockIdx.x : blockIdx.x % tiles_per_col;
  const u32 ta = tile_id / p.tiles_b, tb = tile_id % p.tiles_b;
  const u32 lane0 = p.lanes_are_cols ? ta * T : 0;   // first lane's global lane number (for validity)
  const u32* in = p.in + col * p.in_col + ta * p.in_a + tb * p.in_b;
  u32* out = p.out + col * p.out_col + ta * p.out_a + tb * p.out_b;
  const u64 in_off0 = ta * p.in_a + tb * p.in_b;     // offset inside the column, for in_scale
  const u64 out_off0 = ta * p.out_a + tb * p.out_b;

  for (int i = threadIdx.x; i < (R / 2 > 0 ? R / 2 : 1); i += NT) tws[i] = p.tw_small[i];
  // ---- load
  if (p.in_t == 1) {  // lanes contiguous in memory
    for (int idx = threadIdx.x; idx < R * T; idx += NT) {
      const u32 t = idx & (T - 1), r = idx >> LOG_T;
      u32 v = 0;
      if (r < p.r_nonzero && lane0 + t < p.lanes_valid) {
        const u64 off = r * p.in_r + t;
        v = in[off];
        if (p.in_scale) v = bb_mul(v, p.in_scale[in_off0 + off]);
      }
      tile[r * PITCH + t] = v;
    }
  } else {  // transform axis contiguous in memory (in_r == 1)
    for (int idx = threadIdx.x; idx < R * T; idx += NT) {
      const u32 r = idx & (R - 1), t = idx >> LOG_R;
      u32 v = 0;
      if (r < p.r_nonzero && lane0 + t < p.lanes_valid) {
        const u64 off = t * p.in_t + r;
        v = in[off];
        if (p.in_scale) v = bb_mul(v, p.in_scale[p.lanes_are_cols ? (u64)r : in_off0 + off]);
      }
      tile[r * PITCH + t] = v;
    }
  }
  __syncthreads();
  dif_rounds<LOG_R, LOG_T, LOG_R, NT>(tile, tws);
  // ---- store: tile row rho holds output index bitrev(rho)
  if (p.out_t == 1) {
    for (int idx = threadIdx.x; idx < R * T; idx += NT) {
      const u32 t = idx & (T - 1), k = idx >> LOG_T;
      if (lane0 + t >= p.lanes_valid) continue;
      const u32 rho = LOG_R ? (__brev(k) >> (32 - LOG_R)) : 0;
      u32 v = tile[rho * PITCH + t];
      if (p.tw_big) v = bb_mul(v, p.tw_big[ta * p.tw_a + tb * p.tw_b + t * p.tw_t + k * p.tw_k]);
      const u64 off = k * p.out_k + t;
      if (p.out_scale) v = bb_mul(v, p.out_scale[out_off0 + off]);
      if (p.use_out_const) v = bb_mul(v, p.out_const);
      out[off] = v;
    }
  } else {  // out_k == 1
    for (int idx = threadIdx.x; idx < R * T; idx += NT) {
      const u32 k = idx & (R - 1), t = idx >> LOG_R;
      if (lane0 + t >= p.lanes_valid) continue;
      const u32 rho = LOG_R ? (__brev(k) >> (32 - LOG_R)) : 0;
      u32 v = tile[rho * PITCH + t];
      if (p.tw_big) v = bb_mul(v, p.tw_big[ta * p.tw_a + tb * p.tw_b + t * p.tw_t + k * p.tw_k]);
      const u64 off = t * p.out_t + k;
      if (p.out_scale) v = bb_mul(v, p.out_scale[p.lanes_are_cols ? (u64)k : out_off0 + off]);
      if (p.use_out_const) v = bb_mul(v, p.out_const);
      out[off] = v;
    }
  }
}

template <int LOG_R>
static cudaError_t launch_pass_r(const PassParams& p, u32 blocks, cudaStream_t st) {
  constexpr int LOG_T = LOG_R >= 11 ? 4 : 5;
  constexpr int R = 1 << LOG_R, T = 1 << LOG_T;
  constexpr int ITEMS = (R * T) / 8;
  constexpr int NT = ITEMS >= 512 ? 512 : (ITEMS >= 256 ? 256 : (ITEMS >= 128 ? 128 : 64));
  const size_t smem = (size_t)(R * (T + 1) + (R / 2 > 0 ? R / 2 : 1)) * sizeof(u32);
  auto kern = ntt_pass_kernel<LOG_R, LOG_T, NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, NT, smem, st>>>(p);
  return cudaGetLastError();
}

static cudaError_t launch_pass(int log_r, const PassParams& p, u32 blocks, cudaStream_t st) {
  switch (log_r) {
    case 1: return launch_pass_r<1>(p, blocks, st);
    case 2: return launch_pass_r<2>(p, blocks, st);
    case 3: return launch_pass_r<3>(p, blocks, st);
    case 4: return launch_pass_r<4>(p, blocks, st);
    case 5: return launch_pass_r<5>(p, blocks, st);
    case 6: return launch_pass_r<6>(p, blocks, st);
    case 7: return launch_pass_r<7>(p, blocks, st);
    case 8: return launch_pass_r<8>(p, blocks, st);
    case 9: return launch_pass_r<9>(p, blocks, st);
    case 10: return launch_pass_r<10>(p, blocks, st);
    case 11: return launch_pass_r<11>(p, blocks, st);
  }
  return cudaErrorInvalidValue;
}
static int lanes_of(int log_r) { return log_r >= 11 ? 16 : 32; }

// ---------------------------------------------------------------- table generation
__global__ void powers_kernel(u32* out, u64 n, u32 base, u32 c0) {  // out[i] = c0 * base^i  (all Montgomery)
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i < n) out[i] = bb_mul(c0, bb_pow(base, i));
}
__global__ void twiddle_matrix_kernel(u32* out, u32 log_s, u32 log_m, u32 root) {  // out[k*m + l] = root^(k*l mod S)
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  u64 S = 1ull << log_s;
  if (i >= S) return;
  u64 k = i >> log_m, l = i & ((1ull << log_m) - 1);
  out[i] = bb_pow(root, (k * l) & (S - 1));
}

static u32 host_pow_canon(u32 a, u64 e) { u64 r = 1, b = a; while (e) { if (e & 1) r = r * b % BB_P; b = b * b % BB_P; e >>= 1; } return (u32)r; }
static u32 host_inv_canon(u32 a) { return host_pow_canon(a, BB_P - 2); }

struct NttTables {
  cudaStream_t stream;
  std::map<u64, u32*> cache;  // key -> device table
  std::vector<void*> owned;
  u64* launches;

  u32* get_small(int log_r, bool inv) {
    u64 key = (1ull << 60) | ((u64)log_r << 1) | inv;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    u32 root = ZKIR_BB_ROOTS[log_r];
    if (inv) root = host_inv_canon(root);
    u64 n = log_r ? (1ull << (log_r - 1)) : 1;
    u32* d = nullptr;
    if (cudaMalloc(&d, n * 4) != cudaSuccess) return nullptr;
    powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, n, bb_to_mont_c(root), BB_ONE);
    (*launches)++;
    owned.push_back(d); cache[key] = d;
    return d;
  }
  // twiddle matrix for a sub-transform of size S = 2^log_s split as R = 2^log_r rows by m columns
  u32* get_big(int log_s, int log_r, bool inv) {
    u64 key = (2ull << 60) | ((u64)log_s << 9) | ((u64)log_r << 1) | inv;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    u32 root = ZKIR_BB_ROOTS[log_s];
    if (inv) root = host_inv_canon(root);
    u64 n = 1ull << log_s;
    u32* d = nullptr;
    if (cudaMalloc(&d, n * 4) != cudaSuccess) return nullptr;
    twiddle_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, log_s, log_s - log_r, bb_to_mont_c(root));
    (*launches)++;
    owned.push_back(d); cache[key] = d;
    return d;
  }
  // c0 * base^i table (canonical arguments), i < n
  u32* get_powers(u32 base, u32 c0, u64 n) {
    u64 key = (3ull << 60) ^ ((u64)base << 28) ^ ((u64)c0 * 0x9E3779B97F4A7C15ull) ^ n;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    u32* d = nullptr;
    if (cudaMalloc(&d, n * 4) != cudaSuccess) return nullptr;
    powers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, n, bb_to_mont_c(base), bb_to_mont_c(c0));
    (*launches)++;
    owned.push_back(d); cache[key] = d;
    return d;
  }
  void free_all() { for (void* p : owned) cudaFree(p); owned.clear(); cache.clear(); }
};

NttTables* ntt_tables_create(cudaStream_t st, u64* launch_counter) {
  NttTables* t = new NttTables();
  t->stream = st; t->launches = launch_counter;
  return t;
}
void ntt_tables_destroy(NttTables* t) { if (t) { t->free_all(); delete t; } }
u32* ntt_powers_table(NttTables* t, u32 base, u32 c0, u64 n) { return t->get_powers(base, c0, n); }

// split log_n into pass radices, each <= 11 (two passes up to 2^22, three above)
static void plan_radices(int log_n, int* lr, int* npass) {
  if (log_n <= 10) { lr[0] = log_n; *npass = 1; return; }
  if (log_n <= 22) { lr[0] = (log_n + 1) / 2; lr[1] = log_n - lr[0]; *npass = 2; return; }
  lr[0] = (log_n + 2) / 3; lr[1] = (log_n - lr[0] + 1) / 2; lr[2] = log_n - lr[0] - lr[1]; *npass = 3;
}

// Natural-order NTT of n_cols columns.  in: [n_cols][n >> log_pad] with the upper part implicit zeros (LDE zero
// padding), out: [n_cols][n].  Multi-pass transforms run in column batches through `tmp` (tmp_words capacity) so the
// intermediate of a batch is still L2-resident when the next pass reads it; in == out is allowed (in place).
int ntt_run(NttTables* tb, const u32* d_in, u64 in_col_stride, u32* d_out, u64 out_col_stride, u32* tmp, u64 tmp_words,
            u32 n_cols, int log_n, bool inverse, int log_pad, const u32* in_scale, const u32* out_scale,
            u32 out_const_mont, bool use_out_const, cudaStream_t st) {
  if (log_n < 1 || log_n > 30 || log_pad < 0 || log_pad >= log_n) return -1;
  int lr[3], npass;
  plan_radices(log_n, lr, &npass);
  const u64 n = 1ull << log_n;
  cudaError_t e;
  if (npass == 1) {
    const int T = lanes_of(lr[0]);
    PassParams p = {};
    p.in = d_in; p.out = d_out;
    p.tiles_a = (n_cols + T - 1) / T; p.tiles_b = 1;
    p.lanes_valid = n_cols; p.lanes_are_cols = 1;
    p.in_a = (u64)T * in_col_stride; p.in_t = in_col_stride; p.in_r = 1;
    p.out_a = (u64)T * out_col_stride; p.out_t = out_col_stride; p.out_k = 1;
    p.tw_small = tb->get_small(lr[0], inverse);
    p.in_scale = in_scale; p.out_scale = out_scale; p.out_const = out_const_mont; p.use_out_const = use_out_const;
    p.r_nonzero = (u32)(n >> log_pad);
    if (!p.tw_small) return -4;
    e = launch_pass(lr[0], p, p.tiles_a, st);
    (*tb->launches)++;
    return e == cudaSuccess ? 0 : -2;
  }
  const int L1 = lr[0], L2 = lr[1], L3 = npass == 3 ? lr[2] : 0;
  const u64 R1 = 1ull << L1, m1 = n >> L1, R2 = 1ull << L2, m2 = m1 >> L2;
  if (log_pad > L1) return -1;
  if (tmp_words < n) return -4;
  u32 batch = (u32)(tmp_words / n);
  if (batch > n_cols) batch = n_cols;
  const u32* tws1 = tb->get_small(L1, inverse);
  const u32* tws2 = tb->get_small(L2, inverse);
  const u32* tws3 = npass == 3 ? tb->get_small(L3, inverse) : nullptr;
  const u32* twb1 = tb->get_big(log_n, L1, inverse);
  const u32* twb2 = npass == 3 ? tb->get_big(log_n - L1, L2, inverse) : nullptr;
  if (!tws1 || !tws2 || !twb1 || (npass == 3 && (!tws3 || !twb2))) return -4;
  for (u32 c0 = 0; c0 < n_cols; c0 += batch) {
    const u32 nc = n_cols - c0 < batch ? n_cols - c0 : batch;
    {  // ---- pass A: strided size-R1 transforms, lanes along l in [0, m1); in -> tmp
      const int T = lanes_of(L1);
      PassParams p = {};
      p.in = d_in + (u64)c0 * in_col_stride; p.out = tmp;
      p.tiles_a = 1; p.tiles_b = (u32)(m1 / T);
      p.lanes_valid = T;
      p.in_col = in_col_stride; p.out_col = n;
      p.in_b = T; p.in_t = 1; p.in_r = m1;
      p.out_b = T; p.out_t = 1; p.out_k = m1;
      p.tw_small = tws1; p.tw_big = twb1;
      p.tw_b = T; p.tw_t = 1; p.tw_k = m1;
      p.in_scale = in_scale;
      p.r_nonzero = (u32)(R1 >> log_pad);
      e = launch_pass(L1, p, nc * p.tiles_b, st);
      (*tb->launches)++;
      if (e != cudaSuccess) return -2;
    }
    if (npass == 3) {  // ---- pass B: inside every row k1 (m1 contiguous), strided size-R2 transforms; in place on tmp
      const int T = lanes_of(L2);
      PassParams p = {};
      p.in = tmp; p.out = tmp;
      p.tiles_a = (u32)R1; p.tiles_b = (u32)(m2 / T);
      p.lanes_valid = T;
      p.in_col = n; p.out_col = n;
      p.in_a = m1; p.in_b = T; p.in_t = 1; p.in_r = m2;
      p.out_a = m1; p.out_b = T; p.out_t = 1; p.out_k = m2;
      p.tw_small = tws2; p.tw_big = twb2;
      p.tw_a = 0; p.tw_b = T; p.tw_t = 1; p.tw_k = m2;
      p.r_nonzero = (u32)R2;
      e = launch_pass(L2, p, nc * p.tiles_a * p.tiles_b, st);
      (*tb->launches)++;
      if (e != cudaSuccess) return -2;
    }
    {  // ---- pass C: contiguous transforms, transposing store; tmp -> out
      const int LC = npass == 3 ? L3 : L2;
      const int T = lanes_of(LC);
      PassParams p = {};
      p.in = tmp; p.out = d_out + (u64)c0 * out_col_stride;
      p.lanes_valid = T;
      p.in_col = n; p.out_col = out_col_stride;
      if (npass == 2) {  // tile = T consecutive k1; X[k1 + R1*k2]
        p.tiles_a = (u32)(R1 / T); p.tiles_b = 1;
        p.in_a = (u64)T * m1; p.in_t = m1; p.in_r = 1;
        p.out_a = T; p.out_t = 1; p.out_k = R1;
      } else {           // tile = T consecutive k1, one k2; X[k1 + R1*(k2 + R2*k3)]
        p.tiles_a = (u32)(R1 / T); p.tiles_b = (u32)R2;
        p.in_a = (u64)T * m1; p.in_b = m2; p.in_t = m1; p.in_r = 1;
        p.out_a = T; p.out_b = R1; p.out_t = 1; p.out_k = R1 * R2;
      }
      p.tw_small = npass == 3 ? tws3 : tws2;
      p.out_scale = out_scale; p.out_const = out_const_mont; p.use_out_const = use_out_const;
      p.r_nonzero = 1u << LC;
      e = launch_pass(LC, p, nc * p.tiles_a * p.tiles_b, st);
      (*tb->launches)++;
      if (e != cudaSuccess) return -2;
    }
  }
  return 0;
}

}  // namespace zkir
