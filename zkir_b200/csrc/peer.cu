// Row redistribution of the column-sharded LDE over NVLink peer memory (one proof on several GPUs).
//
// After its share of the LDE a rank holds whole columns [c_lo, c_hi) of the coset-major matrix [W][B][N].  Every other rank t
// needs, of each of those columns, the points j in [t*nj, (t+1)*nj] of every coset (its leaf segment plus the halo row the
// quotient reads as "next row").  One kernel reads the local columns once and stores every element straight into the same
// place of the owning peer's matrix (pointers obtained through CUDA IPC, or plain peer access inside one process): no staging
// buffers, no pack/unpack passes, (G-1)/G of W/G columns leave the rank = about 1/G of the matrix per rank crosses NVLink.
// Bound: NVLink store bandwidth.  The caller separates this kernel from the consumers with a barrier collective.
#include <cuda_runtime.h>
#include "bb.cuh"
#include "kernels.h"

namespace zkir {

__global__ void __launch_bounds__(256) lde_scatter_kernel(const u32* __restrict__ lde, PeerPtrs peers, u32 me, u32 G, u32 c_lo, u32 n_cols,
                                                          u64 N, u32 B, u64 nj, u32 vec) {
  // grid.y = column * B + z; grid.x covers the N points of that coset, `vec` elements per thread
  const u32 c = c_lo + blockIdx.y / B, z = blockIdx.y % B;
  const u64 j = (blockIdx.x * (u64)blockDim.x + threadIdx.x) * vec;
  if (j >= N) return;
  const u64 off = ((u64)c * B + z) * N + j;
  const u32 dest = (u32)(j / nj);                      // vec divides nj: the whole vector has one owner
  if (vec == 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(lde + off);
    if (dest != me) *reinterpret_cast<uint4*>(peers.p[dest] + off) = v;
    if (j % nj == 0) {                                 // first point of a segment = halo of the previous one (wraps)
      const u32 prev = (dest + G - 1) % G;
      if (prev != me) peers.p[prev][off] = v.x;
    }
  } else {
    const u32 v = lde[off];
    if (dest != me) peers.p[dest][off] = v;
    if (j % nj == 0) {
      const u32 prev = (dest + G - 1) % G;
      if (prev != me) peers.p[prev][off] = v;
    }
  }
}

int launch_lde_scatter(const u32* lde, const PeerPtrs& peers, u32 me, u32 G, u32 c_lo, u32 n_cols, u64 N, u32 B, u64 nj, cudaStream_t st,
                       u64* launches) {
  if (!n_cols) return 0;
  const u32 vec = (nj % 4 == 0 && N % 4 == 0) ? 4u : 1u;
  const u64 threads = N / vec;
  dim3 grid((unsigned)((threads + 255) / 256), n_cols * B);
  lde_scatter_kernel<<<grid, 256, 0, st>>>(lde, peers, me, G, c_lo, n_cols, N, B, nj, vec);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace zkir
