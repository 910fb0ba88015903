// Microbenchmark: per-row TMA bulk gathers through a multi-stage per-warp ring (S stages of B rows,
// one mbarrier per stage, the copy for stage s+S-1 is issued before stage s is consumed) --
// is the single-stage 45 G rows/s of tma_gather.cu a latency or a throughput limit?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ring tma_ring.cu && ./tma_ring
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D%=;\nbra W%=;\nD%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}

// W warps per CTA, each with its own ring of S stages x B rows (256 B rows); lanes < B issue one copy each
template <int B, int S, int W, int ROWB>
__global__ void __launch_bounds__(W * 32, 1) k_ring(const char* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                    float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[W * S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = smem + (size_t)warp * S * B * ROWB;
  if (lane < S) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[warp * S + lane])));
  asm volatile("fence.mbarrier_init.release.cluster;");
  __syncwarp();
  const int* my = ids + ((size_t)blockIdx.x * W + warp) * iters * B;
  float2 acc = make_float2(0, 0);
  int c_next = lane < B ? my[lane] : 0;          // ids of the next stage to issue: always one issue ahead
  auto issue = [&](int it) {
    const int s = it % S;
    const int c = c_next;
    if (it + 1 < iters && lane < B) c_next = my[(it + 1) * B + lane];
    const uint32_t bar = s32(&bars[warp * S + s]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(B * ROWB));
    __syncwarp();
    if (lane < B) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(ring + ((size_t)s * B + lane) * ROWB)), "l"(tab + (size_t)c * ROWB), "n"(ROWB), "r"(bar) : "memory");
    }
  };
  for (int it = 0; it < S - 1 && it < iters; ++it) issue(it);
  for (int it = 0; it < iters; ++it) {
    if (it + S - 1 < iters) issue(it + S - 1);
    const int s = it % S;
    mbar_wait(s32(&bars[warp * S + s]), (it / S) & 1);
    const unsigned char* st = ring + (size_t)s * B * ROWB;
#pragma unroll
    for (int u = 0; u < B; ++u) {
#pragma unroll
      for (int q = 0; q < ROWB / 256; ++q) {
        const float2 x = *reinterpret_cast<const float2*>(st + u * ROWB + q * 256 + lane * 8);
        acc.x += x.x; acc.y += x.y;
      }
    }
    __syncwarp();
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}

template <int B, int S, int W, int ROWB>
void run(const char* tab, int rows, int sms, float* out) {
  const int iters = 6400 / B * 32 / W;       // same rows per SM for every shape
  size_t n = (size_t)sms * W * iters * B;
  int* h = new int[n];
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
  int* ids; cudaMalloc(&ids, n * 4); cudaMemcpy(ids, h, n * 4, cudaMemcpyHostToDevice);
  const int smem = W * S * B * ROWB;
  cudaFuncSetAttribute(k_ring<B, S, W, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a); k_ring<B, S, W, ROWB><<<sms, W * 32, smem>>>(tab, ids, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf("ring B=%2d S=%d W=%2d row=%dB smem=%3dKB: %.3f ms  %.2f Grows/s  %.2f TB/s  (%s)\n", B, S, W, ROWB, smem / 1024, best,
         n / best / 1e6, n * (double)ROWB / best / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(ids); delete[] h;
}

int main() {
  const int rows = 100000, sms = 148;   // 25.6 MB table: L2 resident like the Gowalla tables
  char* tab; float* out;
  cudaMalloc(&tab, (size_t)rows * 512); cudaMemset(tab, 0, (size_t)rows * 512);
  cudaMalloc(&out, 4);
  run<16, 1, 32, 256>(tab, rows, sms, out);
  run<16, 2, 16, 256>(tab, rows, sms, out);
  run<16, 2, 24, 256>(tab, rows, sms, out);
  run<16, 3, 16, 256>(tab, rows, sms, out);
  run<16, 4, 8, 256>(tab, rows, sms, out);
  run<16, 4, 12, 256>(tab, rows, sms, out);
  run<32, 2, 8, 256>(tab, rows, sms, out);
  run<32, 2, 12, 256>(tab, rows, sms, out);
  run<32, 3, 8, 256>(tab, rows, sms, out);
  run<32, 4, 4, 256>(tab, rows, sms, out);
  run<32, 4, 6, 256>(tab, rows, sms, out);
  run<32, 6, 4, 256>(tab, rows, sms, out);
  run<8, 4, 16, 512>(tab, rows, sms, out);
  run<16, 3, 8, 512>(tab, rows, sms, out);
  run<16, 2, 12, 512>(tab, rows, sms, out);
  return 0;
}
