// Measurement-only program (not part of the library): the forward ROW pass of the trace LDE (dft_tile_kernel<10, 1, false>, the slowest of
// the four LDE passes) against a TMA-staged variant of the same pass, same arithmetic, same tables, same output bytes.
//
//   shipped kernel : one CTA per (tile, coset), 512 threads, 2 CTAs/SM; every thread loads its 32 points straight from global memory
//                    (the two cosets of a tile are adjacent in blockIdx, so the second read is an L2 hit)
//   TMA variant    : persistent CTAs (1 per SM, 512 threads), the 64 KB tile (16 rows x 1024 contiguous points) arrives with ONE
//                    cp.async.bulk into shared memory, completion on an mbarrier, double-buffered: the copy of tile t+1 flies while tile t
//                    is transformed; both cosets are computed from the one staged tile (no second read at all)
//
// build: nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/ntt_tma_probe tools/ntt_tma_probe.cu -lcuda
// run  : tools/bin/ntt_tma_probe [n_cols=88] [reps=20]
#include <cuda.h>
#include "../zkir_b200/csrc/ntt_fast.cu"

namespace zkir {

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, u32 phase) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// forward / inverse row pass, LOG_R = 10 (A = B_ = 5), plain store (no split, no peers)
template <bool INV>
__global__ void __launch_bounds__(512, 1) dft_row_tma_kernel(const TileParams p, u32 n_tiles) {
  typedef TileShape<10, 1> SH;
  constexpr int A = SH::A, B_ = SH::B_, EA = SH::EA, XB = SH::XB, R = SH::R;
  extern __shared__ __align__(128) u32 sm[];
  u32* ex = sm + 32 * R;                       // exchange buffer, 16 * ROW_PITCH words
  __shared__ __align__(8) uint64_t bar[2];
  const u32 tid = threadIdx.x;
  const u32 q = tid & (XB - 1), l = tid >> B_;
  if (tid == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](u32 t, int b) {   // one thread: the tile's 16 rows are one contiguous 64 KB block of the column
    const u32 tile = t % p.tiles_per_col, col = t / p.tiles_per_col;
    const u32* src = p.in + (u64)col * p.in_col + (u64)tile * p.in_b;
    mbar_expect_tx(&bar[b], 16 * R * 4);
    bulk_g2s(sm + b * 16 * R, src, 16 * R * 4, &bar[b]);
  };
  if (tid == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  u32 it = 0;
  for (u32 t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
    const int b = it & 1;
    if (tid == 0 && t + gridDim.x < n_tiles) issue(t + gridDim.x, b ^ 1);   // stage[b^1] was last read before the trailing barrier of it-1
    mbar_wait(&bar[b], (it >> 1) & 1);
    const u32 tile = t % p.tiles_per_col, col = t / p.tiles_per_col;
    const u32 in_off0 = tile * p.in_b, out_off0 = tile * p.out_b;
    const u32 in_thr = in_off0 + q + l * p.in_t;
    for (u32 z = 0; z < p.nz; z++) {
      u32* out = p.out + (u64)col * p.out_col + (u64)z * p.out_z;
      u32 x[EA];
      const u32* ps = sm + b * 16 * R + l * R + q;
#pragma unroll
      for (int i = 0; i < EA; i++) x[i] = ps[i * XB];
      if (p.in_tab) {
        const uint2* __restrict__ pt = p.in_tab + (u64)z * p.in_tab_z + (in_thr & p.in_tab_mask);
#pragma unroll
        for (int i = 0; i < EA; i++) { const uint2 tw = __ldg(pt + i * XB); x[i] = shoup_mul(x[i], tw.x, tw.y); }
      }
      dft_regs<A, INV>(x);
      const uint2* __restrict__ twq = p.tw_mid + q;
      const u32 w0 = l * SH::ROW_PITCH + q;
#pragma unroll
      for (int i = 0; i < EA; i++) {
        const u32 ks = brev<A>(i);
        u32 v = x[i];
        if (ks != 0) { const uint2 tw = __ldg(twq + ks * XB); v = shoup_mul(v, tw.x, tw.y); }
        ex[w0 + ks * (XB + 1)] = v;
      }
      __syncthreads();
      const u32 u = tid & (XB - 1), l2 = tid >> B_;
#pragma unroll
      for (int g = 0; g < EA / XB; g++) {
        const u32 ks = u + g * XB;
        u32 y[XB];
        const u32 r0 = l2 * SH::ROW_PITCH + ks * (XB + 1);
#pragma unroll
        for (int j = 0; j < XB; j++) y[j] = ex[r0 + j];
        dft_regs<B_, INV>(y);
        u32* po = out + (out_off0 + ks + l2 * p.out_t);
        if (p.out_tab) {
          const u32 toff = in_off0 + l2 * p.in_t + ks;
          const uint2* __restrict__ pt = p.out_tab + (u64)z * p.out_tab_z + (toff & p.out_tab_mask);
#pragma unroll
          for (int j = 0; j < XB; j++) { const uint2 tw = __ldg(pt + brev<B_>(j) * EA); po[brev<B_>(j) * EA] = shoup_mul(y[j], tw.x, tw.y); }
        } else {
#pragma unroll
          for (int j = 0; j < XB; j++) po[brev<B_>(j) * EA] = y[j];
        }
      }
      __syncthreads();   // the exchange buffer (and, after the last coset, this stage buffer) may be overwritten
    }
  }
}

// Variant 2: the shipped launch shape (one CTA per (tile, coset), 2 CTAs/SM, no persistence, no extra shared memory): the tile arrives
// with one cp.async.bulk INTO THE EXCHANGE BUFFER, the threads pick their 32 points from shared memory, a barrier frees the buffer
// for the exchange.  TMA as the load mechanism only.
template <bool INV>
__global__ void __launch_bounds__(512, 2) dft_row_tma1_kernel(const TileParams p) {
  typedef TileShape<10, 1> SH;
  constexpr int A = SH::A, B_ = SH::B_, EA = SH::EA, XB = SH::XB, R = SH::R;
  extern __shared__ __align__(128) u32 sm[];
  __shared__ __align__(8) uint64_t bar;
  const u32 tid = threadIdx.x;
  const u32 q = tid & (XB - 1), l = tid >> B_;
  const u32 z = blockIdx.x % p.nz, bid = blockIdx.x / p.nz;
  const u32 tile = bid % p.tiles_per_col, col = bid / p.tiles_per_col;
  const u32 in_off0 = tile * p.in_b, out_off0 = tile * p.out_b;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, 16 * R * 4);
    bulk_g2s(sm, p.in + (u64)col * p.in_col + (u64)z * p.in_z + in_off0, 16 * R * 4, &bar);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  u32* out = p.out + (u64)col * p.out_col + (u64)z * p.out_z;
  const u32 in_thr = in_off0 + q + l * p.in_t;
  u32 x[EA];
  const u32* ps = sm + l * R + q;
#pragma unroll
  for (int i = 0; i < EA; i++) x[i] = ps[i * XB];
  __syncthreads();   // everybody has its points: the buffer becomes the exchange buffer
  if (p.in_tab) {
    const uint2* __restrict__ pt = p.in_tab + (u64)z * p.in_tab_z + (in_thr & p.in_tab_mask);
#pragma unroll
    for (int i = 0; i < EA; i++) { const uint2 tw = __ldg(pt + i * XB); x[i] = shoup_mul(x[i], tw.x, tw.y); }
  }
  dft_regs<A, INV>(x);
  const uint2* __restrict__ twq = p.tw_mid + q;
  const u32 w0 = l * SH::ROW_PITCH + q;
#pragma unroll
  for (int i = 0; i < EA; i++) {
    const u32 ks = brev<A>(i);
    u32 v = x[i];
    if (ks != 0) { const uint2 tw = __ldg(twq + ks * XB); v = shoup_mul(v, tw.x, tw.y); }
    sm[w0 + ks * (XB + 1)] = v;
  }
  __syncthreads();
  const u32 u = tid & (XB - 1), l2 = tid >> B_;
#pragma unroll
  for (int g = 0; g < EA / XB; g++) {
    const u32 ks = u + g * XB;
    u32 y[XB];
    const u32 r0 = l2 * SH::ROW_PITCH + ks * (XB + 1);
#pragma unroll
    for (int j = 0; j < XB; j++) y[j] = sm[r0 + j];
    dft_regs<B_, INV>(y);
    u32* po = out + (out_off0 + ks + l2 * p.out_t);
    const u32 toff = in_off0 + l2 * p.in_t + ks;
    const uint2* __restrict__ pt = p.out_tab + (u64)z * p.out_tab_z + (toff & p.out_tab_mask);
#pragma unroll
    for (int j = 0; j < XB; j++) { const uint2 tw = __ldg(pt + brev<B_>(j) * EA); po[brev<B_>(j) * EA] = shoup_mul(y[j], tw.x, tw.y); }
  }
}

// Variant 3: the STRIDED digit (dft_tile_kernel<10, 0, fwd, stride 2^10>: tile = 1024 rows x 16 contiguous lanes, 64-byte global segments,
// in place, no tables but the inter-round twiddles) with the tile fetched by cp.async.bulk.tensor from a 3-D tensor map
// {1024 inner, 1024 rows, planes}: four 16 x 256 boxes land in the exchange buffer, the threads pick their 32 points with LDS.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
template <bool INV>
__global__ void __launch_bounds__(512, 2) dft_strided_tma_kernel(const TileParams p, const __grid_constant__ CUtensorMap tm) {
  typedef TileShape<10, 0> SH;
  constexpr int A = SH::A, B_ = SH::B_, EA = SH::EA, XB = SH::XB, R = SH::R;
  extern __shared__ __align__(128) u32 sm[];
  __shared__ __align__(8) uint64_t bar;
  const u32 tid = threadIdx.x;
  const u32 l = tid & 15u, q = tid >> 4;
  const u32 z = blockIdx.x % p.nz, bid = blockIdx.x / p.nz;
  const u32 tile = bid % p.tiles_per_col, col = bid / p.tiles_per_col;   // one block of 2^20 per column: tile = lane group tb
  const u32 out_off0 = tile * p.out_b;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, 16 * R * 4);
    for (int k = 0; k < 4; k++) tma_load_3d(sm + k * 256 * 16, &tm, (int)(tile * 16), k * 256, (int)(col * p.nz + z), &bar);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  u32 x[EA];
#pragma unroll
  for (int i = 0; i < EA; i++) x[i] = sm[(q + i * XB) * 16 + l];
  __syncthreads();   // the staged tile has been read: the buffer becomes the exchange buffer
  dft_regs<A, INV>(x);
  u32* out = p.out + (u64)col * p.out_col + (u64)z * p.out_z;
  const uint2* __restrict__ twq = p.tw_mid + q;
  const u32 w0 = q * (EA / 2) * 32 + l + ((q & 1u) << 4), w1 = q * (EA / 2) * 32 + l + (((q & 1u) ^ 1u) << 4);
#pragma unroll
  for (int i = 0; i < EA; i++) {
    const u32 ks = brev<A>(i);
    u32 v = x[i];
    if (ks != 0) { const uint2 tw = __ldg(twq + ks * XB); v = shoup_mul(v, tw.x, tw.y); }
    sm[((ks & 1u) ? w1 : w0) + (ks >> 1) * 32] = v;
  }
  __syncthreads();
  const u32 l2 = tid & 15u, u = tid >> 4;
#pragma unroll
  for (int g = 0; g < EA / XB; g++) {
    const u32 ks = u + g * XB;
    u32 y[XB];
    const u32 r0 = (ks >> 1) * 32 + l2 + ((ks & 1u) << 4), r1 = (ks >> 1) * 32 + l2 + (((ks & 1u) ^ 1u) << 4);
#pragma unroll
    for (int j = 0; j < XB; j++) y[j] = sm[((j & 1) ? r1 : r0) + j * (EA / 2) * 32];
    dft_regs<B_, INV>(y);
    u32* po = out + (out_off0 + ks * 1024u + l2);
#pragma unroll
    for (int j = 0; j < XB; j++) po[(size_t)brev<B_>(j) * EA * 1024u] = y[j];
  }
}

}  // namespace zkir

using namespace zkir;

#define CK(e) do { cudaError_t e_ = (e); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const u32 n_cols = argc > 1 ? (u32)atoi(argv[1]) : 88u;
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  const int log_n = 20;
  const u64 N = 1ull << log_n, M = 2 * N;
  const u32 nz = 2;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  u64 launches = 0;
  FastNtt* f = fast_ntt_create(st, &launches);
  FastPlan pl;
  if (!fast_plan(log_n, &pl) || pl.nd != 2 || pl.d[1] != 10) { fprintf(stderr, "unexpected plan\n"); return 1; }
  // the tables of the forward coset transform (fast_coset_ntt): coset scale on the input, level twiddle on the output
  const uint2* sc = f->scale(to_digit_plan(pl), ZKIR_BB_GEN, ZKIR_BB_ROOTS[log_n + 1], nz, 1u);
  const uint2* tw = f->level_tw(pl.d[1] + pl.d[0], pl.d[1], false, 1u);
  const uint2* twm = f->tw_mid(pl.d[1], false);
  if (!sc || !tw || !twm) { fprintf(stderr, "tables\n"); return 1; }
  u32 *coef, *out_a, *out_b;
  CK(cudaMalloc(&coef, n_cols * N * 4)); CK(cudaMalloc(&out_a, n_cols * M * 4)); CK(cudaMalloc(&out_b, n_cols * M * 4));
  {
    std::vector<u32> h(n_cols * N);
    u64 s = 0x5EED;
    for (auto& v : h) { s = s * 6364136223846793005ull + 1442695040888963407ull; v = (u32)((s >> 33) % BB_P); }
    CK(cudaMemcpy(coef, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  // TileParams exactly as row_pass() builds them for this pass
  const u32 m = 1u << pl.d[1], rows = (u32)(N >> pl.d[1]);
  TileParams p = {};
  p.in = coef; p.in_col = N; p.in_z = 0; p.out_col = M; p.out_z = N;
  p.tiles_b = rows / 16; p.tiles_per_col = rows / 16;
  p.in_b = 16 * m; p.in_t = m; p.in_r = 1; p.out_k = 1; p.out_b = 16 * m; p.out_t = m;
  p.split_log = 32; p.split_max = 0; p.split_extra = 0; p.peer_shift = 0xffffffffu; p.peer_mask = 0;
  p.in_tab = sc; p.in_tab_z = (u32)N; p.in_tab_mask = 0xffffffffu;
  p.out_tab = tw; p.out_tab_z = 0; p.out_tab_mask = (u32)((1ull << (pl.d[1] + pl.d[0])) - 1);
  p.tw_mid = twm; p.nz = nz;
  const u32 n_tiles = n_cols * p.tiles_per_col;
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  const size_t smem_tma = (size_t)(32 * 1024 + 16 * TileShape<10, 1>::ROW_PITCH) * 4;
  CK(cudaFuncSetAttribute(dft_row_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_it = [&](auto launch, const char* name) -> int {
    for (int i = 0; i < 3; i++) launch();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; i++) launch();
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps, bytes = 4.0 * N * n_cols * (1 + nz);
    printf("%-44s %8.1f us per pass   %7.1f GB/s on the pass's algorithmic bytes (read N, write %u N)\n", name, us, bytes / us * 1e-3, nz);
    return 0;
  };
  p.out = out_a;
  TileParams pa = p;
  if (time_it([&] { launch_tile_t<10, 1, false>(pa, n_tiles, nz, st); }, "shipped dft_tile_kernel<10,1,fwd> (2 CTAs/SM)")) return 1;
  p.out = out_b;
  TileParams pb = p;
  for (int grid_mul = 1; grid_mul <= 1; grid_mul++) {
    const u32 grid = (u32)n_sm * grid_mul;
    if (time_it([&] { dft_row_tma_kernel<false><<<grid, 512, smem_tma, st>>>(pb, n_tiles); }, "TMA-staged persistent variant (1 CTA/SM)")) return 1;
  }
  CK(cudaGetLastError());
  {
    const size_t smem1 = (size_t)16 * TileShape<10, 1>::ROW_PITCH * 4;
    CK(cudaFuncSetAttribute(dft_row_tma1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    u32* out_c;
    CK(cudaMalloc(&out_c, n_cols * M * 4));
    TileParams pc = p;
    pc.out = out_c;
    if (time_it([&] { dft_row_tma1_kernel<false><<<n_tiles * nz, 512, smem1, st>>>(pc); }, "TMA as the load only (shipped shape, 2 CTAs/SM)")) return 1;
    CK(cudaGetLastError());
    std::vector<u32> ha(n_cols * M), hc(n_cols * M);
    CK(cudaMemcpy(ha.data(), out_a, ha.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.data(), out_c, hc.size() * 4, cudaMemcpyDeviceToHost));
    u64 d2 = 0;
    for (size_t i = 0; i < ha.size(); i++) d2 += ha[i] != hc[i];
    printf("load-only variant: outputs %s (%llu words differ)\n", d2 ? "DIFFER" : "identical", (unsigned long long)d2);
    cudaFree(out_c);
    if (d2) return 2;
  }
  // ---- the strided (top) digit of the forward transform, in place on a copy of the row pass's output
  {
    u32 *buf_a, *buf_b;
    CK(cudaMalloc(&buf_a, n_cols * M * 4)); CK(cudaMalloc(&buf_b, n_cols * M * 4));
    TileParams ps = {};
    ps.in_col = ps.out_col = M; ps.in_z = ps.out_z = N;
    ps.tiles_b = (1u << pl.d[1]) / 16; ps.tiles_per_col = ps.tiles_b;                 // one block of 2^20 per (column, coset)
    ps.in_a = ps.out_a = 1u << log_n; ps.in_b = ps.out_b = 16; ps.in_r = ps.out_k = 1u << pl.d[1]; ps.in_t = ps.out_t = 1;
    ps.split_log = 32; ps.peer_shift = 0xffffffffu; ps.nz = nz;
    ps.tw_mid = f->tw_mid(pl.d[0], false);
    if (!ps.tw_mid) return 1;
    const u32 nt = n_cols * ps.tiles_per_col;
    TileParams pa2 = ps, pb2 = ps;
    pa2.in = buf_a; pa2.out = buf_a; pb2.in = buf_b; pb2.out = buf_b;
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {1024, 1024, (cuuint64_t)n_cols * nz};
    const cuuint64_t gstr[2] = {4096, (cuuint64_t)N * 4};
    const cuuint32_t box[3] = {16, 256, 1}, estr[3] = {1, 1, 1};
    CUresult cr = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, buf_b, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
    const size_t smem0 = (size_t)16 * 1024 * 4;
    CK(cudaFuncSetAttribute(dft_strided_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    // correctness first: one pass each from the same input
    CK(cudaMemcpy(buf_a, out_a, n_cols * M * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(buf_b, out_a, n_cols * M * 4, cudaMemcpyDeviceToDevice));
    CK((launch_tile_t<10, 0, false, 10>(pa2, nt, nz, st)));
    dft_strided_tma_kernel<false><<<nt * nz, 512, smem0, st>>>(pb2, tm);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    std::vector<u32> h1(n_cols * M), h2(n_cols * M);
    CK(cudaMemcpy(h1.data(), buf_a, h1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), buf_b, h2.size() * 4, cudaMemcpyDeviceToHost));
    u64 d3 = 0;
    for (size_t i = 0; i < h1.size(); i++) d3 += h1[i] != h2[i];
    printf("strided digit: outputs %s (%llu words differ)\n", d3 ? "DIFFER" : "identical", (unsigned long long)d3);
    // timing (the passes run in place on whatever the buffers hold: the arithmetic is data-independent)
    if (time_it([&] { launch_tile_t<10, 0, false, 10>(pa2, nt, nz, st); }, "shipped dft_tile_kernel<10,0,fwd,2^10> strided digit")) return 1;
    if (time_it([&] { dft_strided_tma_kernel<false><<<nt * nz, 512, smem0, st>>>(pb2, tm); }, "strided digit, cp.async.bulk.tensor as the load")) return 1;
    CK(cudaGetLastError());
    cudaFree(buf_a); cudaFree(buf_b);
    if (d3) return 2;
  }
  // same bytes?
  std::vector<u32> ha(n_cols * M), hb(n_cols * M);
  CK(cudaMemcpy(ha.data(), out_a, ha.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), out_b, hb.size() * 4, cudaMemcpyDeviceToHost));
  u64 diff = 0;
  for (size_t i = 0; i < ha.size(); i++) diff += ha[i] != hb[i];
  printf("outputs %s (%llu of %zu words differ)\n", diff ? "DIFFER" : "identical", (unsigned long long)diff, ha.size());
  fast_ntt_destroy(f);
  return diff ? 2 : 0;
}
