// Microbenchmarks that set the denominators DESIGN.md quotes: HBM copy/read bandwidth, L2-resident read bandwidth,
// and the integer-pipe ceiling for BabyBear modular multiplies (Montgomery: IMAD.WIDE + IMAD + IMAD.HI + IADD + VIADDMNMX).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../zkir_b200/csrc/bb.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void read_kernel(const uint4* __restrict__ in, size_t n, int reps, uint32_t* sink) {
  uint32_t acc = 0;
  for (int r = 0; r < reps; r++)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      uint4 v = in[i];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void write_kernel(uint4* __restrict__ out, size_t n, int reps) {
  for (int r = 0; r < reps; r++)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_uint4(r, r, r, r);
}
// ILP independent chains of dependent Montgomery multiplies
template <int ILP>
__global__ void mul_kernel(uint32_t* out, int iters) {
  uint32_t x[ILP], w = threadIdx.x * 2654435761u % BB_P;
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = (threadIdx.x + 77u * k + blockIdx.x) % BB_P;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = bb_mul(x[k], w);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void add_kernel(uint32_t* out, int iters) {
  uint32_t x[ILP], w = threadIdx.x * 2654435761u % BB_P;
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = (threadIdx.x + 77u * k + blockIdx.x) % BB_P;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = bb_add(x[k], w);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Shoup multiply by a constant with precomputed quotient: 3 IMAD-class + 1 VIADDMNMX
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wq) {
  uint32_t q = __umulhi(x, wq);
  uint32_t r = x * w - q * BB_P;
  uint32_t t = r - BB_P;
  return r < t ? r : t;
}
template <int ILP>
__global__ void shoup_kernel(uint32_t* out, int iters) {
  uint32_t x[ILP], w = threadIdx.x * 2654435761u % BB_P;
  uint32_t wq = (uint32_t)((((uint64_t)w) << 32) / BB_P);
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = (threadIdx.x + 77u * k + blockIdx.x) % BB_P;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = shoup_mul(x[k], w, wq);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs=%d L2=%d MB clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, prop.clockRate);
  const int SM = prop.multiProcessorCount;
  size_t big = 1ull << 30;
  uint4 *a, *b; uint32_t* sink;
  CK(cudaMalloc(&a, big)); CK(cudaMalloc(&b, big)); CK(cudaMalloc(&sink, 1 << 24));
  CK(cudaMemset(a, 1, big)); CK(cudaMemset(b, 2, big));
  for (int bpsm : {4, 8, 16}) {
    float ms = time_ms([&] { copy_kernel<<<SM * bpsm, 256>>>(a, b, big / 16); }, 10);
    printf("copy 1GiB->1GiB grid=%dxSM: %.3f ms  %.0f GB/s (read+write)\n", bpsm, ms, 2.0 * big / ms / 1e6);
  }
  {
    float ms = time_ms([&] { cudaMemcpyAsync(b, a, big, cudaMemcpyDeviceToDevice, 0); }, 10);
    printf("cudaMemcpy D2D 1GiB: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * big / ms / 1e6);
  }
  {
    float ms = time_ms([&] { read_kernel<<<SM * 8, 256>>>(a, big / 16, 1, sink); }, 10);
    printf("read 1GiB: %.3f ms  %.0f GB/s\n", ms, 1.0 * big / ms / 1e6);
    ms = time_ms([&] { write_kernel<<<SM * 8, 256>>>(b, big / 16, 1); }, 10);
    printf("write 1GiB: %.3f ms  %.0f GB/s\n", ms, 1.0 * big / ms / 1e6);
  }
  for (size_t mb : {8, 16, 32, 48, 64, 96}) {
    size_t bytes = mb << 20;
    const int reps = 20;
    float ms = time_ms([&] { read_kernel<<<SM * 8, 256>>>(a, bytes / 16, reps, sink); }, 5);
    printf("L2-resident read %zu MiB x%d: %.3f ms  %.0f GB/s\n", mb, reps, ms, 1.0 * bytes * reps / ms / 1e6);
  }
  for (size_t mb : {8, 16, 32, 48}) {
    size_t bytes = mb << 20;
    float ms = time_ms([&] { for (int r = 0; r < 10; r++) copy_kernel<<<SM * 8, 256>>>(a, b, bytes / 16); }, 5);
    printf("L2-resident copy %zu MiB (x10 launches): %.3f ms  %.0f GB/s (read+write)\n", mb, ms, 2.0 * bytes * 10 / ms / 1e6);
  }
  const int iters = 4096;
  const double thr = (double)SM * 8 * 256;
  {
    float ms = time_ms([&] { mul_kernel<8><<<SM * 8, 256>>>(sink, iters); }, 5);
    printf("montgomery mul ILP8: %.3f ms  %.1f Gmul/s  (%.2f mul/clk/SM at 1.965 GHz)\n", ms, thr * 8 * iters / ms / 1e6, thr * 8 * iters / ms / 1e6 / SM / 1.965);
    ms = time_ms([&] { mul_kernel<16><<<SM * 8, 256>>>(sink, iters); }, 5);
    printf("montgomery mul ILP16: %.3f ms  %.1f Gmul/s  (%.2f mul/clk/SM)\n", ms, thr * 16 * iters / ms / 1e6, thr * 16 * iters / ms / 1e6 / SM / 1.965);
    ms = time_ms([&] { shoup_kernel<8><<<SM * 8, 256>>>(sink, iters); }, 5);
    printf("shoup mul ILP8: %.3f ms  %.1f Gmul/s  (%.2f mul/clk/SM)\n", ms, thr * 8 * iters / ms / 1e6, thr * 8 * iters / ms / 1e6 / SM / 1.965);
    ms = time_ms([&] { add_kernel<8><<<SM * 8, 256>>>(sink, iters); }, 5);
    printf("modular add ILP8: %.3f ms  %.1f Gadd/s  (%.2f add/clk/SM)\n", ms, thr * 8 * iters / ms / 1e6, thr * 8 * iters / ms / 1e6 / SM / 1.965);
  }
  return 0;
}
