// Microbenchmark: row-per-warp gathers of 256-byte rows with 64-bit loads (32 lanes x float2),
// DEPTH independent loads in flight per lane, codes broadcast by shuffle from one coalesced load.
// Compare with the half-warp-per-row 128-bit loads of tma_gather.cu (k_ldg).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldg_rpw ldg_rpw.cu && ./ldg_rpw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int DEPTH, int W>
__global__ void __launch_bounds__(W * 32, 1) k_rpw(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                   float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* my = ids + ((size_t)blockIdx.x * W + warp) * iters * 32;
  float2 acc = make_float2(0, 0);
  int c_next = my[lane];
  for (int it = 0; it < iters; ++it) {      // 32 rows per iteration
    const int c = c_next;
    if (it + 1 < iters) c_next = my[(it + 1) * 32 + lane];
#pragma unroll
    for (int u0 = 0; u0 < 32; u0 += DEPTH) {
      float2 v[DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u0 + u);
        v[u] = __ldg(reinterpret_cast<const float2*>(tab + (size_t)cc * 64) + lane);
      }
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}

template <int DEPTH, int W>
void run(const float* tab, int rows, int sms, float* out) {
  const int iters = 200 * 32 / W / 2;
  size_t n = (size_t)sms * W * iters * 32;
  int* h = new int[n];
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
  int* ids; cudaMalloc(&ids, n * 4); cudaMemcpy(ids, h, n * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a); k_rpw<DEPTH, W><<<sms, W * 32>>>(tab, ids, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf("rpw ldg.64 depth=%2d warps=%2d: %.3f ms  %.2f Grows/s  %.2f TB/s  (%s)\n", DEPTH, W, best, n / best / 1e6,
         n * 256.0 / best / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(ids); delete[] h;
}

int main() {
  const int rows = 100000, sms = 148;
  float* tab; float* out;
  cudaMalloc(&tab, (size_t)rows * 256); cudaMemset(tab, 0, (size_t)rows * 256);
  cudaMalloc(&out, 4);
  run<4, 32>(tab, rows, sms, out);
  run<8, 32>(tab, rows, sms, out);
  run<16, 32>(tab, rows, sms, out);
  run<8, 16>(tab, rows, sms, out);
  run<16, 16>(tab, rows, sms, out);
  run<32, 16>(tab, rows, sms, out);
  run<8, 24>(tab, rows, sms, out);
  run<16, 24>(tab, rows, sms, out);
  return 0;
}
