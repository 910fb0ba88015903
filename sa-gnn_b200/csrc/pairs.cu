// Sampled pair scores over the per-interval outputs: the second consumer of the
// propagation's outputs in LIU-YUXI/SA-GNN, and the producer of the SPARSE part of its upstream
// gradient.
//
//   model.py:196-198   preds_one[s] = sum_c lrelu(user_vector[k][suids[s], c] * item_vector[k][siids[s], c])
//   model.py:171-173   preds[s]     = sum_c        final_user[uids[s], c]      * final_item[iids[s], c]
//
// TF runs this as two embedding_lookup gathers ([n, d] each), a Mul, a Maximum and a reduce_sum, and in
// the backward two IndexedSlices -> dense conversions (unsorted_segment_sum).  Here: one warp per
// sample, 128-bit row loads, a shuffle-tree reduction (deterministic); the backward adds
// g[s] * sigma'(x) * other_row straight into the dense upstream-gradient tables that
// sagnn_propagate_bwd consumes (red.global.add.f32: samples repeat users, so rows collide).
// Tables are addressed as (base, row stride) so that both the [T,R,d] and the [R,T,d] layouts work.
#include <cub/cub.cuh>

#include "common.cuh"

namespace sagnn {

template <bool ACT>
__global__ void __launch_bounds__(256)
pair_scores_fwd_kernel(const float* __restrict__ u_rows, int64_t u_stride, const float* __restrict__ i_rows,
                       int64_t i_stride, const int32_t* __restrict__ uids, const int32_t* __restrict__ iids,
                       int64_t n, int q4, float leaky, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const float4* pu = reinterpret_cast<const float4*>(u_rows + (int64_t)__ldg(uids + s) * u_stride);
    const float4* pi = reinterpret_cast<const float4*>(i_rows + (int64_t)__ldg(iids + s) * i_stride);
    float acc = 0.f;
    for (int q = lane; q < q4; q += 32) {
      const float4 a = __ldg(pu + q), b = __ldg(pi + q);
      float x[4] = {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += ACT ? fmaxf(leaky * x[j], x[j]) : x[j];   // Utils/NNLayers.py:135-136
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) scores[s] = acc;
  }
}

template <bool ACT>
__global__ void __launch_bounds__(256)
pair_scores_bwd_kernel(const float* __restrict__ u_rows, int64_t u_stride, const float* __restrict__ i_rows,
                       int64_t i_stride, const int32_t* __restrict__ uids, const int32_t* __restrict__ iids,
                       int64_t n, int q4, float leaky, const float* __restrict__ g_scores, float* d_u, int64_t du_stride,
                       float* d_i, int64_t di_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const int64_t u = __ldg(uids + s), i = __ldg(iids + s);
    const float g = __ldg(g_scores + s);
    const float4* pu = reinterpret_cast<const float4*>(u_rows + u * u_stride);
    const float4* pi = reinterpret_cast<const float4*>(i_rows + i * i_stride);
    float* gu = d_u ? d_u + u * du_stride : nullptr;
    float* gi = d_i ? d_i + i * di_stride : nullptr;
    for (int q = lane; q < q4; q += 32) {
      const float4 a4 = __ldg(pu + q), b4 = __ldg(pi + q);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = a[j] * b[j];
        // TF MaximumGrad(leaky*x, x): the gradient takes the leaky branch where leaky*x >= x
        const float sp = ACT ? ((leaky * x >= x) ? leaky : 1.f) : 1.f;
        if (gu) atomicAdd(gu + 4 * q + j, g * sp * b[j]);
        if (gi) atomicAdd(gi + 4 * q + j, g * sp * a[j]);
      }
    }
  }
}

// Deterministic variant: the samples arrive sorted by the row they scatter into (stable radix sort of
// (row id, sample index)); the warp at the head of each run of equal ids adds the run's terms in sample order and
// read-modify-writes the gradient row once -- single writer per row, no atomics, same bits every run.
// SIDE 0: scatter into the user rows (other = item row), SIDE 1: into the item rows.
template <bool ACT>
__global__ void __launch_bounds__(256)
pair_scores_bwd_sorted_kernel(const float* __restrict__ u_rows, int64_t u_stride, const float* __restrict__ i_rows,
                              int64_t i_stride, const int32_t* __restrict__ uids, const int32_t* __restrict__ iids,
                              const int32_t* __restrict__ keys, const int32_t* __restrict__ order, int64_t n, int q4,
                              float leaky, const float* __restrict__ g_scores, float* d_rows, int64_t d_stride, int side) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const int32_t key = __ldg(keys + s);
    if (s > 0 && __ldg(keys + s - 1) == key) continue;      // not the head of its run
    float* out = d_rows + (int64_t)key * d_stride;
    for (int q = lane; q < q4; q += 32) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int64_t t = s; t < n && __ldg(keys + t) == key; ++t) {
        const int64_t smp = __ldg(order + t);
        const float g = __ldg(g_scores + smp);
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(u_rows + (int64_t)__ldg(uids + smp) * u_stride) + q);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(i_rows + (int64_t)__ldg(iids + smp) * i_stride) + q);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = a[j] * b[j];
          const float sp = ACT ? ((leaky * x >= x) ? leaky : 1.f) : 1.f;   // TF MaximumGrad tie rule, as above
          acc[j] += g * sp * (side ? a[j] : b[j]);
        }
      }
      float4* o4 = reinterpret_cast<float4*>(out) + q;
      float4 v = *o4;
      v.x += acc[0]; v.y += acc[1]; v.z += acc[2]; v.w += acc[3];
      *o4 = v;
    }
  }
}

__global__ void pair_iota_kernel(int32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

struct PairWs { size_t keys, order, iota, tmp, tmp_bytes, total; };
static int pair_ws_layout(int64_t n, PairWs& w) {
  size_t tb = 0;
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                             (int32_t*)nullptr, (int)n));
  const size_t arr = align_up(sizeof(int32_t) * (size_t)(n > 0 ? n : 1), 256);
  w.keys = 0; w.order = arr; w.iota = 2 * arr; w.tmp = 3 * arr; w.tmp_bytes = tb; w.total = 3 * arr + align_up(tb, 256);
  return SAGNN_OK;
}

static int check_pairs(const void* u_rows, const void* i_rows, const void* uids, const void* iids, int64_t n, int d,
                       int64_t u_stride, int64_t i_stride, int activation, const char* fn) {
  SAGNN_REQUIRE(n >= 0, SAGNN_INVALID_ARG, "%s: n=%lld", fn, (long long)n);
  SAGNN_REQUIRE(d > 0 && d % 4 == 0, SAGNN_INVALID_ARG, "%s: latdim d=%d must be a positive multiple of 4", fn, d);
  SAGNN_REQUIRE(u_stride % 4 == 0 && i_stride % 4 == 0 && u_stride >= d && i_stride >= d, SAGNN_INVALID_ARG,
                "%s: row strides (%lld, %lld floats) must be multiples of 4 and >= d", fn, (long long)u_stride,
                (long long)i_stride);
  SAGNN_REQUIRE(activation == 0 || activation == 1, SAGNN_INVALID_ARG, "%s: activation %d (0 = none, 1 = leakyRelu)", fn,
                activation);
  SAGNN_REQUIRE(n == 0 || (u_rows && i_rows && uids && iids), SAGNN_INVALID_ARG, "%s: NULL tensor", fn);
  SAGNN_REQUIRE(((uintptr_t)u_rows | (uintptr_t)i_rows) % 16 == 0, SAGNN_INVALID_ARG, "%s: tables must be 16-byte aligned", fn);
  return SAGNN_OK;
}

static unsigned pair_blocks(int64_t n) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + 7) / 8;                       // 8 warps (samples in flight) per block
  const int64_t cap = (int64_t)sms * 8;                   // a multiple of the SM count, grid-stride beyond it
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace sagnn

using namespace sagnn;

extern "C" int sagnn_pair_scores_fwd(const float* u_rows, int64_t u_stride, const float* i_rows, int64_t i_stride,
                                     const int32_t* uids, const int32_t* iids, int64_t n, int d, int activation,
                                     float leaky, float* scores, sagnn_stream_t stream) {
  if (int rc = check_pairs(u_rows, i_rows, uids, iids, n, d, u_stride, i_stride, activation, "pair_scores_fwd")) return rc;
  SAGNN_REQUIRE(n == 0 || scores, SAGNN_INVALID_ARG, "pair_scores_fwd: NULL scores");
  if (n == 0) return SAGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (activation)
    pair_scores_fwd_kernel<true><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, n, d / 4, leaky, scores);
  else
    pair_scores_fwd_kernel<false><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, n, d / 4, leaky, scores);
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

extern "C" int sagnn_pair_scores_bwd(const float* u_rows, int64_t u_stride, const float* i_rows, int64_t i_stride,
                                     const int32_t* uids, const int32_t* iids, int64_t n, int d, int activation,
                                     float leaky, const float* g_scores, float* d_u_rows, int64_t du_stride,
                                     float* d_i_rows, int64_t di_stride, sagnn_stream_t stream) {
  if (int rc = check_pairs(u_rows, i_rows, uids, iids, n, d, u_stride, i_stride, activation, "pair_scores_bwd")) return rc;
  SAGNN_REQUIRE(n == 0 || g_scores, SAGNN_INVALID_ARG, "pair_scores_bwd: NULL g_scores");
  SAGNN_REQUIRE((!d_u_rows || du_stride >= d) && (!d_i_rows || di_stride >= d), SAGNN_INVALID_ARG,
                "pair_scores_bwd: gradient row strides must be >= d");
  if (n == 0 || (!d_u_rows && !d_i_rows)) return SAGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (activation)
    pair_scores_bwd_kernel<true><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, n, d / 4, leaky,
                                                                 g_scores, d_u_rows, du_stride, d_i_rows, di_stride);
  else
    pair_scores_bwd_kernel<false><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, n, d / 4, leaky,
                                                                  g_scores, d_u_rows, du_stride, d_i_rows, di_stride);
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

extern "C" int sagnn_pair_scores_bwd_ws_bytes(int64_t n, size_t* bytes) {
  SAGNN_REQUIRE(bytes && n >= 0 && n < ((int64_t)1 << 31), SAGNN_INVALID_ARG, "pair_scores_bwd_ws_bytes: n=%lld", (long long)n);
  PairWs w;
  if (int rc = pair_ws_layout(n, w)) return rc;
  *bytes = w.total;
  return SAGNN_OK;
}

extern "C" int sagnn_pair_scores_bwd_det(const float* u_rows, int64_t u_stride, const float* i_rows, int64_t i_stride,
                                         const int32_t* uids, const int32_t* iids, int64_t n, int d, int activation,
                                         float leaky, const float* g_scores, float* d_u_rows, int64_t du_stride,
                                         float* d_i_rows, int64_t di_stride, void* ws, size_t ws_bytes,
                                         sagnn_stream_t stream) {
  if (int rc = check_pairs(u_rows, i_rows, uids, iids, n, d, u_stride, i_stride, activation, "pair_scores_bwd_det")) return rc;
  SAGNN_REQUIRE(n < ((int64_t)1 << 31), SAGNN_INVALID_ARG, "pair_scores_bwd_det: n=%lld", (long long)n);
  SAGNN_REQUIRE(n == 0 || g_scores, SAGNN_INVALID_ARG, "pair_scores_bwd_det: NULL g_scores");
  SAGNN_REQUIRE((!d_u_rows || (du_stride >= d && du_stride % 4 == 0)) && (!d_i_rows || (di_stride >= d && di_stride % 4 == 0)),
                SAGNN_INVALID_ARG, "pair_scores_bwd_det: gradient row strides must be multiples of 4 and >= d");
  SAGNN_REQUIRE(((uintptr_t)d_u_rows | (uintptr_t)d_i_rows) % 16 == 0, SAGNN_INVALID_ARG,
                "pair_scores_bwd_det: gradient tables must be 16-byte aligned");
  if (n == 0 || (!d_u_rows && !d_i_rows)) return SAGNN_OK;
  PairWs w;
  if (int rc = pair_ws_layout(n, w)) return rc;
  SAGNN_REQUIRE(ws && ws_bytes >= w.total, SAGNN_WORKSPACE_TOO_SMALL, "pair_scores_bwd_det: workspace %zu < %zu bytes",
                ws_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  int32_t* keys = (int32_t*)(base + w.keys); int32_t* order = (int32_t*)(base + w.order); int32_t* iota = (int32_t*)(base + w.iota);
  pair_iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(iota, n);
  for (int side = 0; side < 2; ++side) {
    float* dst = side ? d_i_rows : d_u_rows;
    if (!dst) continue;
    size_t tb = w.tmp_bytes;
    SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(base + w.tmp, tb, side ? iids : uids, keys, iota, order, (int)n, 0, 32, st));
    if (activation)
      pair_scores_bwd_sorted_kernel<true><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, keys, order,
                                                                          n, d / 4, leaky, g_scores, dst, side ? di_stride : du_stride, side);
    else
      pair_scores_bwd_sorted_kernel<false><<<pair_blocks(n), 256, 0, st>>>(u_rows, u_stride, i_rows, i_stride, uids, iids, keys, order,
                                                                           n, d / 4, leaky, g_scores, dst, side ? di_stride : du_stride, side);
  }
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}
