// Microbenchmark: how fast can an SM gather 256-byte rows with per-row TMA bulk copies
// (cp.async.bulk global->shared, mbarrier completion) vs plain 128-bit loads?  One CTA per SM,
// W warps; every warp repeatedly gathers batches of B random rows of a table and sums them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather tma_gather.cu && ./tma_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int B>
__global__ void __launch_bounds__(1024, 1) k_tma(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                 float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem) + warp * B * 64;
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  const int* my = ids + ((size_t)blockIdx.x * 32 + warp) * iters * B;
  float4 acc = make_float4(0, 0, 0, 0);
  unsigned phase = 0;
  for (int it = 0; it < iters; ++it) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(B * 256));
    __syncwarp();
    if (lane < B) {
      const int c = my[it * B + lane];
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                   ::"r"(s32(ring + lane * 64)), "l"(tab + (size_t)c * 64), "r"(s32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(bar)), "r"(phase) : "memory");
    phase ^= 1;
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const float4 x = *reinterpret_cast<const float4*>(ring + u * 64 + (lane & 15) * 4);
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();
  }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

template <int B>
__global__ void __launch_bounds__(1024, 1) k_ldg(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                 float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* my = ids + ((size_t)blockIdx.x * 32 + warp) * iters * B;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
    const int c = lane < B ? my[it * B + lane] : 0;
#pragma unroll
    for (int u0 = 0; u0 < B; u0 += 8) {          // half-warp per row, two rows per instruction like the product kernel
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u0 + 2 * u + (lane >> 4));
        v[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)cc * 64 + (lane & 15) * 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

int main() {
  const int rows = 100000, iters = 200, sms = 148;   // 25.6 MB table: L2 resident like the Gowalla tables
  constexpr int B = 16;
  float* tab; int* ids; float* out;
  cudaMalloc(&tab, (size_t)rows * 256); cudaMemset(tab, 0, (size_t)rows * 256);
  size_t n = (size_t)sms * 32 * iters * B;
  int* h = new int[n];
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
  cudaMalloc(&ids, n * 4); cudaMemcpy(ids, h, n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 4);
  cudaFuncSetAttribute(k_tma<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * B * 256);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a); k_tma<B><<<sms, 1024, 32 * B * 256>>>(tab, ids, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("tma  bulk 256B rows: %.3f ms  %.2f Grows/s  %.2f TB/s  (%s)\n", ms, n / ms / 1e6, n * 256.0 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaEventRecord(a); k_ldg<B><<<sms, 1024>>>(tab, ids, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("ldg  128-bit  rows: %.3f ms  %.2f Grows/s  %.2f TB/s  (%s)\n", ms, n / ms / 1e6, n * 256.0 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
