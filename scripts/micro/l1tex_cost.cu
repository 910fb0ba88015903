// Microbenchmark (round 2): what does one 256-byte row cost an SM on each path?
// One persistent CTA per SM; every warp sums batches of random rows.  Reported per variant:
// G rows/s chip-wide and SM cycles per row (at the SM clock read from the device attribute).
//   ldg64      : warp-per-row LDG.64 gathers out of an L2-resident table (25.6 MB)
//   ldg64_l1   : the same loads from a 96 KB table (L1-resident): cost of an L1 hit
//   ldg32x2    : the row as two LDG.32 (one 128-byte line per instruction)
//   ldg128h    : half-warp per row, LDG.128 (two rows per instruction)
//   lds64      : rows read from a staged copy in shared memory (768 rows)
//   mix        : hot-first: HOT of every 16 rows from shared memory, the rest LDG.64
//   stg64      : row stores to random rows
//   bulk       : per-row cp.async.bulk (TMA 1-D) into a per-warp ring, double buffered, then LDS.64
//   gather4    : cp.async.bulk.tensor.2d tile::gather4 (4 rows per instruction), same ring
//   bulk+ldg   : every warp fetches half of its rows by TMA and half by LDG.64
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1tex_cost l1tex_cost.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ int g_abort;

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, unsigned parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  g_abort = 1;
  return false;
}

constexpr int B = 16;   // rows per batch (LSU kernels)

enum { V_LDG64 = 0, V_LDG32X2, V_LDG128H, V_LDS64, V_MIX, V_STG64 };

template <int V, int HOT>
__global__ void __launch_bounds__(1024, 1) k_lsu(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                 float* out, int hot_rows);

// Same row paths, but the codes are read with warp-uniform LDS.128 (4 codes per instruction, the way the
// product kernel does) from a 1 KB per-warp block that is re-used every 16 batches: no shuffles.
// V: 0 = LDG.64 gathers, 3 = LDS.64 from the staged rows, 4 = HOT staged + (16-HOT) LDG.64, 6 = codes only
template <int V, int HOT>
__global__ void __launch_bounds__(1024, 1) k_nosh(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                  float* out, int hot_rows) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) int codes[32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (V == 3 || V == 4) {
    for (int i = threadIdx.x; i < hot_rows * 16; i += blockDim.x)
      reinterpret_cast<float4*>(smem)[i] = reinterpret_cast<const float4*>(tab)[i];
  }
  const int* my = ids + ((size_t)blockIdx.x * 32 + warp) * iters * B;
  for (int i = lane; i < 256; i += 32) codes[warp][i] = (V == 3) ? my[i] % 512 : my[i];
  __syncthreads();
  const char* tl = reinterpret_cast<const char*>(tab) + lane * 8;
  const uint32_t sl = s32(smem) + lane * 8;
  float2 acc = make_float2(0, 0);
  for (int it = 0; it < iters; ++it) {
    const int4* cp = reinterpret_cast<const int4*>(&codes[warp][(it & 15) * 16]);
    int c[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) { const int4 t = cp[g]; c[4 * g] = t.x; c[4 * g + 1] = t.y; c[4 * g + 2] = t.z; c[4 * g + 3] = t.w; }
    if (V == 0) {
      float2 v[B];
#pragma unroll
      for (int u = 0; u < B; ++u) v[u] = __ldg(reinterpret_cast<const float2*>(tl + (size_t)(uint32_t)c[u] * 256));
#pragma unroll
      for (int u = 0; u < B; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    } else if (V == 3) {
#pragma unroll
      for (int u = 0; u < B; ++u) {
        float2 x;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x.x), "=f"(x.y) : "r"(sl + (uint32_t)c[u] * 256));
        acc.x += x.x; acc.y += x.y;
      }
    } else if (V == 4) {
      float2 v[B - HOT > 0 ? B - HOT : 1];
#pragma unroll
      for (int u = HOT; u < B; ++u) v[u - HOT] = __ldg(reinterpret_cast<const float2*>(tl + (size_t)(uint32_t)c[u] * 256));
#pragma unroll
      for (int u = 0; u < HOT; ++u) {
        float2 x;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x.x), "=f"(x.y) : "r"(sl + ((uint32_t)c[u] & 511u) * 256));
        acc.x += x.x; acc.y += x.y;
      }
#pragma unroll
      for (int u = HOT; u < B; ++u) { acc.x += v[u - HOT].x; acc.y += v[u - HOT].y; }
    } else {
#pragma unroll
      for (int u = 0; u < B; ++u) acc.x += __int_as_float(c[u]);
    }
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}

template <int V, int HOT>
__global__ void __launch_bounds__(1024, 1) k_lsu(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                 float* out, int hot_rows) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (V == V_LDS64 || V == V_MIX) {
    for (int i = threadIdx.x; i < hot_rows * 16; i += blockDim.x)
      reinterpret_cast<float4*>(smem)[i] = reinterpret_cast<const float4*>(tab)[i];
    __syncthreads();
  }
  const int* my = ids + ((size_t)blockIdx.x * 32 + warp) * iters * B;
  float2 acc = make_float2(0, 0);
  float4 acc4 = make_float4(0, 0, 0, 0);
  int c_next = lane < B ? my[lane] : 0;
  for (int it = 0; it < iters; ++it) {
    const int c = c_next;
    if (it + 1 < iters && lane < B) c_next = my[(it + 1) * B + lane];
    if (V == V_LDG64) {
      float2 v[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u);
        v[u] = __ldg(reinterpret_cast<const float2*>(tab + (size_t)cc * 64) + lane);
      }
#pragma unroll
      for (int u = 0; u < B; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    } else if (V == V_LDG32X2) {
      float v[B][2];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u);
        v[u][0] = __ldg(tab + (size_t)cc * 64 + lane);
        v[u][1] = __ldg(tab + (size_t)cc * 64 + 32 + lane);
      }
#pragma unroll
      for (int u = 0; u < B; ++u) { acc.x += v[u][0]; acc.y += v[u][1]; }
    } else if (V == V_LDG128H) {
      float4 v[B / 2];
#pragma unroll
      for (int u = 0; u < B / 2; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, 2 * u + (lane >> 4));
        v[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)cc * 64) + (lane & 15));
      }
#pragma unroll
      for (int u = 0; u < B / 2; ++u) { acc4.x += v[u].x; acc4.y += v[u].y; acc4.z += v[u].z; acc4.w += v[u].w; }
    } else if (V == V_LDS64) {
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u) % 768;
        const float2 x = *(reinterpret_cast<const float2*>(smem + (size_t)cc * 256) + lane);
        acc.x += x.x; acc.y += x.y;
      }
    } else if (V == V_MIX) {
      float2 v[B - HOT > 0 ? B - HOT : 1];
#pragma unroll
      for (int u = HOT; u < B; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u);
        v[u - HOT] = __ldg(reinterpret_cast<const float2*>(tab + (size_t)cc * 64) + lane);
      }
#pragma unroll
      for (int u = 0; u < HOT; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u) % 768;
        const float2 x = *(reinterpret_cast<const float2*>(smem + (size_t)cc * 256) + lane);
        acc.x += x.x; acc.y += x.y;
      }
#pragma unroll
      for (int u = HOT; u < B; ++u) { acc.x += v[u - HOT].x; acc.y += v[u - HOT].y; }
    } else if (V == V_STG64) {
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, u);
        *(reinterpret_cast<float2*>(const_cast<float*>(tab) + (size_t)cc * 64) + lane) = make_float2((float)it, (float)u);
      }
    }
  }
  if (acc.x + acc4.x == 123.456f) out[0] = acc.y + acc4.y + acc4.z + acc4.w;
}


// "task mix": what one task of the layer kernel moves, with no bookkeeping at all -- 12 gathered rows
// (Zipf ids), NLD own rows loaded and NST rows stored at unique, streaming addresses (row = task id).
template <int NLD, int NST, bool SCATTER = false>
__global__ void __launch_bounds__(1024, 1) k_task(const float* __restrict__ tab, const int* __restrict__ ids, int iters,
                                                  const float* __restrict__ own, float* __restrict__ dst, float* out) {
  __shared__ __align__(16) int codes[32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* my = ids + ((size_t)blockIdx.x * 32 + warp) * iters * 16;
  for (int i = lane; i < 256; i += 32) codes[warp][i] = my[i];
  __syncthreads();
  const char* tl = reinterpret_cast<const char*>(tab) + lane * 8;
  float2 acc = make_float2(0, 0);
  const size_t nwarps = (size_t)gridDim.x * 32;
  for (int it = 0; it < iters; ++it) {
    size_t task = (size_t)it * nwarps + (size_t)blockIdx.x * 32 + warp;           // unique row per task
    if (SCATTER) task = (task * 100003ull) % ((size_t)iters * nwarps);            // ... in scattered order (degree-sorted schedule)
    const int4* cp = reinterpret_cast<const int4*>(&codes[warp][(it & 15) * 16]);
    int c[12];
#pragma unroll
    for (int g = 0; g < 3; ++g) { const int4 t = cp[g]; c[4 * g] = t.x; c[4 * g + 1] = t.y; c[4 * g + 2] = t.z; c[4 * g + 3] = t.w; }
    float2 v[12], o[NLD > 0 ? NLD : 1];
#pragma unroll
    for (int u = 0; u < 12; ++u) v[u] = __ldg(reinterpret_cast<const float2*>(tl + (size_t)(uint32_t)c[u] * 256));
#pragma unroll
    for (int q = 0; q < NLD; ++q) o[q] = __ldg(reinterpret_cast<const float2*>(own + ((size_t)q * iters * nwarps + task) * 64) + lane);
    float2 a = make_float2(0, 0);
#pragma unroll
    for (int u = 0; u < 12; ++u) { a.x += v[u].x; a.y += v[u].y; }
#pragma unroll
    for (int q = 0; q < NLD; ++q) { a.x += o[q].x; a.y += o[q].y; }
#pragma unroll
    for (int q = 0; q < NST; ++q) *(reinterpret_cast<float2*>(dst + ((size_t)q * iters * nwarps + task) * 64) + lane) = a;
    acc.x += a.x; acc.y += a.y;
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}

// TMA variants: per-warp ring of S stages x B rows.  MODE 0: per-row bulk copies issued by lanes 0..B-1;
// MODE 1: gather4 issued by lanes 0..B/4-1; MODE 2: bulk for the first B/2 rows, LDG.64 for the rest.
template <int MODE, int W, int S, int B>
__global__ void __launch_bounds__(W * 32, 1) k_tma(const float* __restrict__ tab, const __grid_constant__ CUtensorMap tmap,
                                                   const int* __restrict__ ids, int iters, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[W * S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = smem + (size_t)warp * S * B * 256;
  if (lane < S) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[warp * S + lane])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  const int* my = ids + ((size_t)blockIdx.x * W + warp) * iters * B;
  constexpr int NT = MODE == 2 ? B / 2 : B;            // rows that travel by TMA
  auto issue = [&](int it) {
    const int st = it % S;
    const uint32_t bar = s32(&bars[warp * S + st]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NT * 256) : "memory");
    __syncwarp();
    if (MODE == 1) {
      if (lane < B / 4) {
        const int4 c = *reinterpret_cast<const int4*>(my + (size_t)it * B + lane * 4);
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(s32(ring + ((size_t)st * B + lane * 4) * 256)), "l"(&tmap), "r"(0), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"(bar)
                     : "memory");
      }
    } else {
      if (lane < NT) {
        const int c = my[(size_t)it * B + lane];
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                     ::"r"(s32(ring + ((size_t)st * B + lane) * 256)), "l"(tab + (size_t)c * 64), "r"(bar) : "memory");
      }
    }
  };
  float2 acc = make_float2(0, 0);
  for (int it = 0; it < S - 1 && it < iters; ++it) issue(it);
  for (int it = 0; it < iters; ++it) {
    if (it + S - 1 < iters) issue(it + S - 1);
    float2 v[MODE == 2 ? B / 2 : 1];
    if (MODE == 2) {
      const int c = lane < B ? my[(size_t)it * B + lane] : 0;
#pragma unroll
      for (int u = 0; u < B / 2; ++u) {
        const int cc = __shfl_sync(0xffffffffu, c, B / 2 + u);
        v[u] = __ldg(reinterpret_cast<const float2*>(tab + (size_t)cc * 64) + lane);
      }
    }
    const int st = it % S;
    if (!mbar_wait_bounded(s32(&bars[warp * S + st]), (it / S) & 1)) return;
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const float2 x = *(reinterpret_cast<const float2*>(ring + ((size_t)st * B + u) * 256) + lane);
      acc.x += x.x; acc.y += x.y;
    }
    if (MODE == 2) {
#pragma unroll
      for (int u = 0; u < B / 2; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
    __syncwarp();
  }
  if (acc.x == 123.456f) out[0] = acc.y;
}

static double g_clk_ghz = 1.965;
static void report(const char* name, float ms, size_t n) {
  int ab = 0;
  cudaMemcpyFromSymbol(&ab, g_abort, 4);
  printf("%-28s %8.3f ms  %7.2f Grows/s  %6.2f TB/s  %5.2f SM-cycles/row  %s%s\n", name, ms, n / ms / 1e6, n * 256.0 / ms / 1e9,
         ms * 1e-3 * g_clk_ghz * 1e9 * 148 / n, cudaGetErrorString(cudaGetLastError()), ab ? "  [ABORTED: mbarrier timeout]" : "");
  fflush(stdout);
}

template <typename F>
static float best_of(F f, int reps = 4) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(2); }
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 100000, sms = 148, iters = 200;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  if (clk_khz > 0) g_clk_ghz = clk_khz * 1e-6;
  printf("SM clock (attribute) %.3f GHz, batch %d rows\n", g_clk_ghz, B);
  float* tab; float* out;
  CK(cudaMalloc(&tab, (size_t)rows * 256)); CK(cudaMemset(tab, 0, (size_t)rows * 256));
  CK(cudaMalloc(&out, 4));
  const size_t n = (size_t)sms * 32 * iters * B;
  int* h = new int[n]; int* h_small = new int[n];
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); h_small[i] = h[i] % 384; }
  // Zipf-distributed ids over the table (popularity rank scattered by a fixed permutation), like the item / user
  // popularity of the synthetic interval graphs (alpha 1.0 / 0.8): do popular rows serialise at their L2 slice?
  auto zipf_ids = [&](double alpha, int nrows) {
    double* cdf = new double[nrows]; double acc = 0;
    for (int r = 0; r < nrows; ++r) { acc += pow(r + 1.0, -alpha); cdf[r] = acc; }
    int* perm = new int[nrows];
    for (int r = 0; r < nrows; ++r) perm[r] = r;
    uint64_t t = 0x9E3779B97F4A7C15ull;
    for (int r = nrows - 1; r > 0; --r) { t ^= t << 13; t ^= t >> 7; t ^= t << 17; int q = (int)(t % (r + 1)); int x = perm[r]; perm[r] = perm[q]; perm[q] = x; }
    int* out = new int[n];
    for (size_t i = 0; i < n; ++i) {
      t ^= t << 13; t ^= t >> 7; t ^= t << 17;
      const double u = (t >> 11) * (1.0 / 9007199254740992.0) * acc;
      int lo = 0, hi = nrows - 1;
      while (lo < hi) { int mid = (lo + hi) / 2; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
      out[i] = perm[lo];
    }
    int* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, out, n * 4, cudaMemcpyHostToDevice));
    delete[] cdf; delete[] perm; delete[] out;
    return d;
  };
  int* ids_z10 = zipf_ids(1.0, 52621);
  int* ids_z08 = zipf_ids(0.8, 48653);
  int* ids; int* ids_small;
  CK(cudaMalloc(&ids, n * 4)); CK(cudaMemcpy(ids, h, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&ids_small, n * 4)); CK(cudaMemcpy(ids_small, h_small, n * 4, cudaMemcpyHostToDevice));
  const int hot_rows = 768; const size_t hot_bytes = (size_t)hot_rows * 256;
  CK(cudaFuncSetAttribute(k_lsu<V_LDS64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));
  CK(cudaFuncSetAttribute(k_lsu<V_MIX, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));
  CK(cudaFuncSetAttribute(k_lsu<V_MIX, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));
  CK(cudaFuncSetAttribute(k_lsu<V_MIX, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));

  report("ldg64 (L2-resident table)", best_of([&] { k_lsu<V_LDG64, 0><<<sms, 1024>>>(tab, ids, iters, out, hot_rows); }), n);
  report("ldg64_l1 (96 KB table)", best_of([&] { k_lsu<V_LDG64, 0><<<sms, 1024>>>(tab, ids_small, iters, out, hot_rows); }), n);
  report("ldg32x2", best_of([&] { k_lsu<V_LDG32X2, 0><<<sms, 1024>>>(tab, ids, iters, out, hot_rows); }), n);
  report("ldg32x2_l1", best_of([&] { k_lsu<V_LDG32X2, 0><<<sms, 1024>>>(tab, ids_small, iters, out, hot_rows); }), n);
  report("ldg128 half-warp", best_of([&] { k_lsu<V_LDG128H, 0><<<sms, 1024>>>(tab, ids, iters, out, hot_rows); }), n);
  report("ldg128 half-warp_l1", best_of([&] { k_lsu<V_LDG128H, 0><<<sms, 1024>>>(tab, ids_small, iters, out, hot_rows); }), n);
  report("lds64 (768 staged rows)", best_of([&] { k_lsu<V_LDS64, 0><<<sms, 1024, hot_bytes>>>(tab, ids, iters, out, hot_rows); }), n);
  report("mix 4/16 lds + 12/16 ldg", best_of([&] { k_lsu<V_MIX, 4><<<sms, 1024, hot_bytes>>>(tab, ids, iters, out, hot_rows); }), n);
  report("mix 8/16 lds + 8/16 ldg", best_of([&] { k_lsu<V_MIX, 8><<<sms, 1024, hot_bytes>>>(tab, ids, iters, out, hot_rows); }), n);
  report("mix 12/16 lds + 4/16 ldg", best_of([&] { k_lsu<V_MIX, 12><<<sms, 1024, hot_bytes>>>(tab, ids, iters, out, hot_rows); }), n);
  CK(cudaFuncSetAttribute(k_nosh<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(512 * 256)));
  CK(cudaFuncSetAttribute(k_nosh<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(512 * 256)));
  CK(cudaFuncSetAttribute(k_nosh<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(512 * 256)));
  CK(cudaFuncSetAttribute(k_nosh<4, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(512 * 256)));
  report("nosh codes only", best_of([&] { k_nosh<6, 0><<<sms, 1024>>>(tab, ids, iters, out, 512); }), n);
  report("nosh ldg64 (L2)", best_of([&] { k_nosh<0, 0><<<sms, 1024>>>(tab, ids, iters, out, 512); }), n);
  report("nosh ldg64 (L1: 96 KB)", best_of([&] { k_nosh<0, 0><<<sms, 1024>>>(tab, ids_small, iters, out, 512); }), n);
  CK(cudaFuncSetAttribute(k_nosh<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
  report("nosh ldg64 zipf1.0 (L1 190K)", best_of([&] { k_nosh<0, 0><<<sms, 1024>>>(tab, ids_z10, iters, out, 512); }), n);
  report("nosh ldg64 zipf0.8 (L1 190K)", best_of([&] { k_nosh<0, 0><<<sms, 1024>>>(tab, ids_z08, iters, out, 512); }), n);
  report("nosh ldg64 uniform (L1 90K)", best_of([&] { k_nosh<0, 0><<<sms, 1024, 100 * 1024>>>(tab, ids, iters, out, 512); }), n);
  report("nosh ldg64 zipf1.0 (L1 90K)", best_of([&] { k_nosh<0, 0><<<sms, 1024, 100 * 1024>>>(tab, ids_z10, iters, out, 512); }), n);
  report("nosh ldg64 zipf0.8 (L1 90K)", best_of([&] { k_nosh<0, 0><<<sms, 1024, 100 * 1024>>>(tab, ids_z08, iters, out, 512); }), n);
  report("nosh lds64 (512 staged)", best_of([&] { k_nosh<3, 0><<<sms, 1024, 512 * 256>>>(tab, ids, iters, out, 512); }), n);
  report("nosh mix 4 lds + 12 ldg", best_of([&] { k_nosh<4, 4><<<sms, 1024, 512 * 256>>>(tab, ids, iters, out, 512); }), n);
  report("nosh mix 8 lds + 8 ldg", best_of([&] { k_nosh<4, 8><<<sms, 1024, 512 * 256>>>(tab, ids, iters, out, 512); }), n);
  report("nosh mix 12 lds + 4 ldg", best_of([&] { k_nosh<4, 12><<<sms, 1024, 512 * 256>>>(tab, ids, iters, out, 512); }), n);
  report("stg64 (random rows)", best_of([&] { k_lsu<V_STG64, 0><<<sms, 1024>>>(tab, ids, iters, out, hot_rows); }), n);
  CK(cudaMemset(tab, 0, (size_t)rows * 256));

  {
    const int it_t = 64;                                   // 148 x 32 x 64 = 303,104 tasks: one layer launch of the Gowalla shape
    const size_t ntask = (size_t)sms * 32 * it_t;
    float* own; float* dst;
    CK(cudaMalloc(&own, ntask * 256 * 2)); CK(cudaMemset(own, 0, ntask * 256 * 2));
    CK(cudaMalloc(&dst, ntask * 256 * 2));
    float* flush; CK(cudaMalloc(&flush, 256u << 20));
    auto run = [&](const char* nm, auto f) {
      float best = 1e9;
      for (int r = 0; r < 4; ++r) {
        CK(cudaMemset(flush, r, 256u << 20));                // L2 flush, like bench.py
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("%-34s %8.3f ms  = %6.1f SM-cycles/task (12 gathers + own loads + stores)\n", nm, best, best * 1e-3 * g_clk_ghz * 1e9 * 148 / ntask);
    };
    run("task mix: gathers only", [&] { k_task<0, 0><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix: + 1 own load", [&] { k_task<1, 0><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix: + 1 own load + 1 store", [&] { k_task<1, 1><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix: + 2 own loads + 2 stores", [&] { k_task<2, 2><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix scattered rows: 1 own + 1 store", [&] { k_task<1, 1, true><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix scattered rows: 2 own + 2 stores", [&] { k_task<2, 2, true><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
    run("task mix: + 0 own loads + 2 stores", [&] { k_task<0, 2><<<sms, 1024>>>(tab, ids_z10, it_t, own, dst, out); });
  }
  // tensor map for gather4: 2-D [rows, 64] fp32, box {64, 1}
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  CUtensorMap tmap;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  }
  auto tma_set = [&](auto w_, auto s_, auto b_) {
    constexpr int W = decltype(w_)::value, S = decltype(s_)::value, BB = decltype(b_)::value;
    const int it2 = iters * 16 / BB;
    const size_t nn = (size_t)sms * W * it2 * BB;
    const size_t sm = (size_t)W * S * BB * 256;
    char nm[64];
    CK(cudaFuncSetAttribute(k_tma<0, W, S, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_tma<1, W, S, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_tma<2, W, S, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    snprintf(nm, sizeof nm, "tma bulk/row W%d S%d B%d", W, S, BB);
    report(nm, best_of([&] { k_tma<0, W, S, BB><<<sms, W * 32, sm>>>(tab, tmap, ids, it2, out); }), nn);
    snprintf(nm, sizeof nm, "tma gather4 W%d S%d B%d", W, S, BB);
    report(nm, best_of([&] { k_tma<1, W, S, BB><<<sms, W * 32, sm>>>(tab, tmap, ids, it2, out); }), nn);
    snprintf(nm, sizeof nm, "tma bulk+ldg64 W%d S%d B%d", W, S, BB);
    report(nm, best_of([&] { k_tma<2, W, S, BB><<<sms, W * 32, sm>>>(tab, tmap, ids, it2, out); }), nn);
  };
  tma_set(std::integral_constant<int, 32>(), std::integral_constant<int, 3>(), std::integral_constant<int, 8>());
  tma_set(std::integral_constant<int, 16>(), std::integral_constant<int, 3>(), std::integral_constant<int, 16>());
  tma_set(std::integral_constant<int, 8>(), std::integral_constant<int, 3>(), std::integral_constant<int, 32>());
  return 0;
}
