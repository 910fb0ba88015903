// Device-side construction of the interval-graph plan: A_k CSR, stable A_k^T CSR,
// degrees, optional edge weights, and the degree-binned schedule.
//
// Replaces (LIU-YUXI/SA-GNN) DataHandler.transpose (DataHandler.py:9-11) and the 2T
// DataHandler.transToLsts calls of model.py:227-237 (DataHandler.py:47-69): same index
// order (row-major COO of the canonical CSR; transposed rows list user ids ascending),
// same value-sum "degrees" (int64), same dead int32 normalisation when asked for.
// One-time setup; the sort / scan primitives come from CUB (toolkit headers).
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace sagnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file,
            line, what);
  return SAGNN_CUDA_ERROR;
}

const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// flags[0] |= 1 unsorted rows, |= 2 id out of range
__global__ void check_coo_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                                 int64_t n, int U, int I, int* flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = row[i], c = col[i];
  int f = 0;
  if (r < 0 || r >= U || c < 0 || c >= I) f |= 2;
  if (i + 1 < n && row[i + 1] < r) f |= 1;
  if (f) atomicOr(flags, f);
}

__global__ void degree_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                              int64_t n, int32_t* deg_u, int32_t* deg_i) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(deg_u + row[i], 1);
  atomicAdd(deg_i + col[i], 1);
}

__global__ void iota_kernel(int32_t* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)i;
}

__global__ void iota_u32_kernel(uint32_t* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ in, const int32_t* __restrict__ perm, int64_t n,
                              T* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}

// per-row int64 sums of the stored values (np.sum(mat, axis) of DataHandler.py:54-55)
__global__ void valsum_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ val,
                              int64_t row0, int64_t n_rows, int64_t* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int64_t s = 0;
  for (int64_t e = rowptr[row0 + r]; e < rowptr[row0 + r + 1]; ++e) s += val[e];
  out[row0 + r] = s;
}

__device__ __forceinline__ double ref_inv_sqrt(double s) {
  // DataHandler.py:54-55: 1 / (sqrt(sum + 1e-8) + 1e-8), IEEE double like numpy
  return __ddiv_rn(1.0, __dadd_rn(__dsqrt_rn(__dadd_rn(s, 1e-8)), 1e-8));
}

// One thread per edge of (interval k, side): looks its row up by binary search.
// mode 0: weights w = (float)(D_row[row] * D_col[col]) on structural degrees
// mode 1: data  = (int32)((double)val * D_row[row] * D_col[col]) on value sums (DataHandler.py:56-59)
template <int MODE>
__global__ void edge_norm_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx,
                                 const int32_t* __restrict__ val, const int32_t* __restrict__ deg,
                                 const int64_t* __restrict__ valsum, int64_t row0, int64_t n_rows_side,
                                 int64_t other_row0, int64_t e0, int64_t n_edges,
                                 float* __restrict__ w_out, int32_t* __restrict__ data_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  int64_t e = e0 + i;
  // largest r with rowptr[row0 + r] <= e
  int64_t lo = 0, hi = n_rows_side;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[row0 + mid] <= e) lo = mid; else hi = mid;
  }
  int64_t r = row0 + lo;
  int64_t c = other_row0 + idx[e];
  if (MODE == 0) {
    double dr = ref_inv_sqrt((double)deg[r]);
    double dc = ref_inv_sqrt((double)deg[c]);
    w_out[e] = (float)__dmul_rn(dr, dc);
  } else {
    double dr = ref_inv_sqrt((double)valsum[r]);
    double dc = ref_inv_sqrt((double)valsum[c]);
    double v = __dmul_rn(__dmul_rn((double)val[e], dr), dc);
    data_out[i] = (int32_t)v;   // C truncation == numpy store into an int32 array
  }
}

__global__ void rel_indptr_kernel(const int64_t* __restrict__ rowptr, int64_t row0, int64_t n,
                                  int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = (int32_t)(rowptr[row0 + i] - rowptr[row0]);
}

// sorted_deg is descending: count entries > kChunk (binary search, one thread), max degree
__global__ void count_long_kernel(const int32_t* __restrict__ sorted_deg, int64_t n, int64_t* out) {
  int64_t lo = 0, hi = n;   // first index with deg <= kChunk
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_deg[mid] > kChunk) lo = mid + 1; else hi = mid;
  }
  out[0] = lo;
  out[1] = n ? sorted_deg[0] : 0;
}

__global__ void chunk_count_kernel(const int32_t* __restrict__ sorted_deg, int64_t n_long,
                                   int64_t* __restrict__ nch) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_long) nch[i] = (sorted_deg[i] + kChunk - 1) / kChunk;
}

__global__ void chunk_fill_kernel(const int64_t* __restrict__ chunk_base, int64_t n_long,
                                  int64_t n_chunks, uint32_t* __restrict__ chunk_lr) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  int64_t lo = 0, hi = n_long;   // largest lr with chunk_base[lr] <= c
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (chunk_base[mid] <= c) lo = mid; else hi = mid;
  }
  chunk_lr[c] = (uint32_t)lo;
}

// task records: chunk c of long row lr covers edges [rs + deg*ci/nch, rs + deg*(ci+1)/nch)
__global__ void task_fill_kernel(const int64_t* __restrict__ rowptr, const uint32_t* __restrict__ order,
                                 const uint32_t* __restrict__ long_row, const int64_t* __restrict__ chunk_base,
                                 const uint32_t* __restrict__ chunk_lr, int64_t n_chunks, int64_t n_short,
                                 sagnn_task* __restrict__ tasks) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunks + n_short) return;
  sagnn_task k;
  if (t < n_chunks) {
    const uint32_t lr = chunk_lr[t];
    const int64_t cb = chunk_base[lr], nch = chunk_base[lr + 1] - cb, ci = t - cb;
    k.grow = long_row[lr];
    const int64_t rs = rowptr[k.grow], deg = rowptr[k.grow + 1] - rs;
    k.e0 = rs + deg * ci / nch;
    k.meta = (uint32_t)(rs + deg * (ci + 1) / nch - k.e0) | 0x80000000u;
  } else {
    k.grow = order[t - n_chunks];
    k.e0 = rowptr[k.grow];
    k.meta = (uint32_t)(rowptr[k.grow + 1] - k.e0);
  }
  tasks[t] = k;
}

struct CastI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

static inline unsigned blocks_for(int64_t n, int threads = 256) {
  return (unsigned)((n + threads - 1) / threads);
}

static int bits_for(int64_t n) {   // number of key bits needed for values in [0, n)
  int b = 1;
  while (b < 31 && ((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace sagnn

using namespace sagnn;

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" const char* sagnn_last_error(void) { return sagnn::last_error(); }
extern "C" const char* sagnn_version(void) { return "sagnn_b200 0.1 (sm_100a)"; }

extern "C" int sagnn_plan_create(int T, int U, int I, const int64_t* nnz_host, sagnn_plan** out) {
  SAGNN_REQUIRE(out != nullptr, SAGNN_INVALID_ARG, "plan_create: out is NULL");
  *out = nullptr;
  SAGNN_REQUIRE(T > 0 && U > 0 && I > 0 && nnz_host, SAGNN_INVALID_ARG,
                "plan_create: need T,U,I > 0 and nnz (got T=%d U=%d I=%d)", T, U, I);
  int64_t N = (int64_t)U + I;
  SAGNN_REQUIRE((int64_t)T * N < ((int64_t)1 << 31), SAGNN_INVALID_ARG,
                "plan_create: T*(U+I) = %lld rows exceeds 2^31", (long long)((int64_t)T * N));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("plan_create: no CUDA device (%s); sagnn_b200 has no CPU fallback",
              cudaGetErrorString(e));
    (void)cudaGetLastError();
    return SAGNN_CUDA_ERROR;
  }
  sagnn_plan* p = new sagnn_plan();
  p->T = T; p->U = U; p->I = I; p->N = N; p->n_rows = (int64_t)T * N;
  p->nnz.assign(nnz_host, nnz_host + T);
  p->base.resize(T + 1);
  p->base[0] = 0;
  for (int k = 0; k < T; ++k) {
    if (nnz_host[k] <= 0 || nnz_host[k] >= ((int64_t)1 << 31)) {
      set_error("plan_create: nnz[%d] = %lld must be in [1, 2^31)", k, (long long)nnz_host[k]);
      delete p;
      return SAGNN_INVALID_ARG;
    }
    p->base[k + 1] = p->base[k] + 2 * nnz_host[k];
    p->e_total += nnz_host[k];
  }
  p->is_set.assign(T, 0);
  SAGNN_CUDA(cudaGetDevice(&p->device));
  SAGNN_CUDA(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
  cudaError_t a1 = cudaMalloc(&p->deg, sizeof(int32_t) * p->n_rows);
  cudaError_t a2 = cudaMalloc(&p->rowptr, sizeof(int64_t) * (p->n_rows + 1));
  cudaError_t a3 = cudaMalloc(&p->idx, sizeof(int32_t) * 2 * p->e_total);
  if (a1 != cudaSuccess || a2 != cudaSuccess || a3 != cudaSuccess) {
    sagnn_plan_destroy(p);
    return cuda_fail(a1 != cudaSuccess ? a1 : (a2 != cudaSuccess ? a2 : a3), "cudaMalloc(plan)",
                     __FILE__, __LINE__);
  }
  SAGNN_CUDA(cudaMemset(p->deg, 0, sizeof(int32_t) * p->n_rows));
  *out = p;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_destroy(sagnn_plan* p) {
  if (!p) return SAGNN_OK;
  sagnn::free_host_cache(p);
  cudaFree(p->deg); cudaFree(p->rowptr); cudaFree(p->idx); cudaFree(p->val); cudaFree(p->w);
  cudaFree(p->valsum); cudaFree(p->order); cudaFree(p->long_row); cudaFree(p->chunk_base);
  cudaFree(p->chunk_lr); cudaFree(p->tasks);
  delete p;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_set_interval(sagnn_plan* p, int k, const int32_t* row, const int32_t* col,
                                       const int32_t* val, const float* w, int64_t nnz,
                                       sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(p && row && col, SAGNN_INVALID_ARG, "set_interval: NULL plan/row/col");
  SAGNN_REQUIRE(k >= 0 && k < p->T, SAGNN_INVALID_ARG, "set_interval: k=%d outside [0,%d)", k, p->T);
  SAGNN_REQUIRE(nnz == p->nnz[k], SAGNN_INVALID_ARG, "set_interval: nnz=%lld but plan was created with %lld",
                (long long)nnz, (long long)p->nnz[k]);
  SAGNN_REQUIRE(!p->finalized && !p->is_set[k], SAGNN_INVALID_ARG,
                "set_interval: interval %d already set or plan finalized", k);
  bool any_set = false;
  for (char c : p->is_set) any_set |= (c != 0);
  SAGNN_REQUIRE(!any_set || ((val != nullptr) == p->has_val && (w != nullptr) == p->has_custom_w),
                SAGNN_INVALID_ARG, "set_interval: val/w must be given for all intervals or none");
  if (val && !p->val) SAGNN_CUDA(cudaMalloc(&p->val, sizeof(int32_t) * 2 * p->e_total));
  if (w && !p->w) SAGNN_CUDA(cudaMalloc(&p->w, sizeof(float) * 2 * p->e_total));
  p->has_val = val != nullptr;
  p->has_custom_w = w != nullptr;

  const int U = p->U, I = p->I;
  int* flags = nullptr;
  SAGNN_CUDA(cudaMalloc(&flags, sizeof(int)));
  SAGNN_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));
  check_coo_kernel<<<blocks_for(nnz), 256, 0, st>>>(row, col, nnz, U, I, flags);
  int hflags = 0;
  SAGNN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  cudaFree(flags);
  SAGNN_REQUIRE(!(hflags & 2), SAGNN_OUT_OF_RANGE,
                "set_interval: interval %d has an edge outside [0,%d) x [0,%d)", k, U, I);
  SAGNN_REQUIRE(!(hflags & 1), SAGNN_UNSORTED_INPUT,
                "set_interval: interval %d: segment ids are not increasing", k);

  int32_t* deg_u = p->deg + (int64_t)k * p->N;
  int32_t* deg_i = deg_u + U;
  degree_kernel<<<blocks_for(nnz), 256, 0, st>>>(row, col, nnz, deg_u, deg_i);

  // A_k CSR: column ids in input order
  int32_t* idx_u = p->idx + p->base[k];
  int32_t* idx_i = idx_u + nnz;
  SAGNN_CUDA(cudaMemcpyAsync(idx_u, col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, st));
  if (val) SAGNN_CUDA(cudaMemcpyAsync(p->val + p->base[k], val, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, st));
  if (w) SAGNN_CUDA(cudaMemcpyAsync(p->w + p->base[k], w, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st));

  // A_k^T CSR: stable sort of the edges by item id; edges arrive ordered by user id, so
  // every item row lists its users ascending (== csr_matrix(coo.transpose()), DataHandler.py:9-11)
  int32_t *keys_out = nullptr, *perm_in = nullptr, *perm = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  SAGNN_CUDA(cudaMalloc(&keys_out, sizeof(int32_t) * nnz));
  SAGNN_CUDA(cudaMalloc(&perm_in, sizeof(int32_t) * nnz));
  SAGNN_CUDA(cudaMalloc(&perm, sizeof(int32_t) * nnz));
  iota_kernel<<<blocks_for(nnz), 256, 0, st>>>(perm_in, nnz);
  int end_bit = bits_for(I);
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, col, keys_out, perm_in, perm, nnz, 0, end_bit, st));
  SAGNN_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, col, keys_out, perm_in, perm, nnz, 0, end_bit, st));
  gather_kernel<int32_t><<<blocks_for(nnz), 256, 0, st>>>(row, perm, nnz, idx_i);
  if (val) gather_kernel<int32_t><<<blocks_for(nnz), 256, 0, st>>>(val, perm, nnz, p->val + p->base[k] + nnz);
  if (w) gather_kernel<float><<<blocks_for(nnz), 256, 0, st>>>(w, perm, nnz, p->w + p->base[k] + nnz);
  SAGNN_CUDA(cudaGetLastError());
  SAGNN_CUDA(cudaStreamSynchronize(st));
  cudaFree(tmp); cudaFree(keys_out); cudaFree(perm_in); cudaFree(perm);
  p->is_set[k] = 1;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_finalize(sagnn_plan* p, int weight_mode, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(p, SAGNN_INVALID_ARG, "finalize: NULL plan");
  SAGNN_REQUIRE(!p->finalized, SAGNN_INVALID_ARG, "finalize: already finalized");
  for (int k = 0; k < p->T; ++k)
    SAGNN_REQUIRE(p->is_set[k], SAGNN_INVALID_ARG, "finalize: interval %d was never set", k);
  SAGNN_REQUIRE(weight_mode >= 0 && weight_mode <= 2, SAGNN_INVALID_ARG, "finalize: bad weight_mode %d", weight_mode);
  SAGNN_REQUIRE(weight_mode != SAGNN_WEIGHTS_CUSTOM || p->has_custom_w, SAGNN_INVALID_ARG,
                "finalize: SAGNN_WEIGHTS_CUSTOM needs w_dev at set_interval");
  const int64_t R = p->n_rows;

  // row pointers: one exclusive scan over the degrees in global row order
  {
    auto in = thrust::make_transform_iterator((const int32_t*)p->deg, CastI64());
    void* tmp = nullptr; size_t tb = 0;
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, p->rowptr, R, st));
    SAGNN_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, p->rowptr, R, st));
    int64_t total = 2 * p->e_total;
    SAGNN_CUDA(cudaMemcpyAsync(p->rowptr + R, &total, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }

  if (p->has_val) {
    SAGNN_CUDA(cudaMalloc(&p->valsum, sizeof(int64_t) * R));
    valsum_kernel<<<blocks_for(R), 256, 0, st>>>(p->rowptr, p->val, 0, R, p->valsum);
  }

  p->weight_mode = weight_mode;
  if (weight_mode == SAGNN_WEIGHTS_LIGHTGCN) {
    if (!p->w) SAGNN_CUDA(cudaMalloc(&p->w, sizeof(float) * 2 * p->e_total));
    for (int k = 0; k < p->T; ++k) {
      int64_t ru = (int64_t)k * p->N, ri = ru + p->U, n = p->nnz[k];
      edge_norm_kernel<0><<<blocks_for(n), 256, 0, st>>>(p->rowptr, p->idx, nullptr, p->deg, nullptr, ru,
                                                         p->U, ri, p->base[k], n, p->w, nullptr);
      edge_norm_kernel<0><<<blocks_for(n), 256, 0, st>>>(p->rowptr, p->idx, nullptr, p->deg, nullptr, ri,
                                                         p->I, ru, p->base[k] + n, n, p->w, nullptr);
    }
  } else if (weight_mode == SAGNN_WEIGHTS_NONE && p->w) {
    cudaFree(p->w);   // custom weights supplied but not wanted
    p->w = nullptr;
  }

  // ---- degree-binned schedule ---------------------------------------------------------
  int32_t *sorted_deg = nullptr; uint32_t *rows_in = nullptr, *rows_sorted = nullptr;
  SAGNN_CUDA(cudaMalloc(&sorted_deg, sizeof(int32_t) * R));
  SAGNN_CUDA(cudaMalloc(&rows_in, sizeof(uint32_t) * R));
  SAGNN_CUDA(cudaMalloc(&rows_sorted, sizeof(uint32_t) * R));
  iota_u32_kernel<<<blocks_for(R), 256, 0, st>>>(rows_in, R);
  {
    void* tmp = nullptr; size_t tb = 0;
    SAGNN_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, p->deg, sorted_deg, rows_in, rows_sorted, R, 0, 32, st));
    SAGNN_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
    SAGNN_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, p->deg, sorted_deg, rows_in, rows_sorted, R, 0, 32, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  int64_t* cnt = nullptr;
  SAGNN_CUDA(cudaMalloc(&cnt, 2 * sizeof(int64_t)));
  count_long_kernel<<<1, 1, 0, st>>>(sorted_deg, R, cnt);
  int64_t hcnt[2] = {0, 0};
  SAGNN_CUDA(cudaMemcpyAsync(hcnt, cnt, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  cudaFree(cnt);
  p->n_long = hcnt[0];
  p->max_deg = (int32_t)hcnt[1];
  p->n_short = R - p->n_long;
  SAGNN_CUDA(cudaMalloc(&p->order, sizeof(uint32_t) * (p->n_short ? p->n_short : 1)));
  SAGNN_CUDA(cudaMemcpyAsync(p->order, rows_sorted + p->n_long, sizeof(uint32_t) * p->n_short, cudaMemcpyDeviceToDevice, st));
  SAGNN_CUDA(cudaMalloc(&p->long_row, sizeof(uint32_t) * (p->n_long ? p->n_long : 1)));
  SAGNN_CUDA(cudaMalloc(&p->chunk_base, sizeof(int64_t) * (p->n_long + 1)));
  p->n_chunks = 0;
  if (p->n_long) {
    SAGNN_CUDA(cudaMemcpyAsync(p->long_row, rows_sorted, sizeof(uint32_t) * p->n_long, cudaMemcpyDeviceToDevice, st));
    int64_t* nch = nullptr;
    SAGNN_CUDA(cudaMalloc(&nch, sizeof(int64_t) * (p->n_long + 1)));
    SAGNN_CUDA(cudaMemsetAsync(nch, 0, sizeof(int64_t) * (p->n_long + 1), st));
    chunk_count_kernel<<<blocks_for(p->n_long), 256, 0, st>>>(sorted_deg, p->n_long, nch);
    void* tmp = nullptr; size_t tb = 0;
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nch, p->chunk_base, p->n_long + 1, st));
    SAGNN_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, nch, p->chunk_base, p->n_long + 1, st));
    SAGNN_CUDA(cudaMemcpyAsync(&p->n_chunks, p->chunk_base + p->n_long, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp); cudaFree(nch);
    SAGNN_CUDA(cudaMalloc(&p->chunk_lr, sizeof(uint32_t) * p->n_chunks));
    chunk_fill_kernel<<<blocks_for(p->n_chunks), 256, 0, st>>>(p->chunk_base, p->n_long, p->n_chunks, p->chunk_lr);
  } else {
    SAGNN_CUDA(cudaMemsetAsync(p->chunk_base, 0, sizeof(int64_t), st));
  }
  {
    const int64_t nt = p->n_chunks + p->n_short;
    SAGNN_CUDA(cudaMalloc(&p->tasks, sizeof(sagnn_task) * (nt ? nt : 1)));
    task_fill_kernel<<<blocks_for(nt), 256, 0, st>>>(p->rowptr, p->order, p->long_row, p->chunk_base,
                                                     p->chunk_lr, p->n_chunks, p->n_short, p->tasks);
  }
  SAGNN_CUDA(cudaGetLastError());
  SAGNN_CUDA(cudaStreamSynchronize(st));
  cudaFree(sorted_deg); cudaFree(rows_in); cudaFree(rows_sorted);
  p->finalized = true;
  return SAGNN_OK;
}

static int check_ks(const sagnn_plan* p, int k, int side, const char* fn) {
  SAGNN_REQUIRE(p, SAGNN_INVALID_ARG, "%s: NULL plan", fn);
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "%s: plan not finalized", fn);
  SAGNN_REQUIRE(k >= 0 && k < p->T && (side == 0 || side == 1), SAGNN_INVALID_ARG,
                "%s: bad interval %d / side %d", fn, k, side);
  return SAGNN_OK;
}

extern "C" int sagnn_plan_get_csr(const sagnn_plan* p, int k, int side, int32_t* indptr,
                                  int32_t* indices, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_ks(p, k, side, "get_csr")) return rc;
  int64_t row0 = (int64_t)k * p->N + (side ? p->U : 0);
  int64_t R = side ? p->I : p->U;
  if (indptr) rel_indptr_kernel<<<blocks_for(R + 1), 256, 0, st>>>(p->rowptr, row0, R, indptr);
  if (indices)
    SAGNN_CUDA(cudaMemcpyAsync(indices, p->idx + p->base[k] + (side ? p->nnz[k] : 0),
                               sizeof(int32_t) * p->nnz[k], cudaMemcpyDeviceToDevice, st));
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

extern "C" int sagnn_plan_get_degrees(const sagnn_plan* p, int k, int side, int32_t* deg,
                                      int64_t* valsum, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_ks(p, k, side, "get_degrees")) return rc;
  int64_t row0 = (int64_t)k * p->N + (side ? p->U : 0);
  int64_t R = side ? p->I : p->U;
  if (deg) SAGNN_CUDA(cudaMemcpyAsync(deg, p->deg + row0, sizeof(int32_t) * R, cudaMemcpyDeviceToDevice, st));
  if (valsum) {
    SAGNN_REQUIRE(p->valsum, SAGNN_INVALID_ARG, "get_degrees: plan was built without stored values");
    SAGNN_CUDA(cudaMemcpyAsync(valsum, p->valsum + row0, sizeof(int64_t) * R, cudaMemcpyDeviceToDevice, st));
  }
  return SAGNN_OK;
}

extern "C" int sagnn_plan_norm_data(const sagnn_plan* p, int k, int side, int32_t* data,
                                    sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_ks(p, k, side, "norm_data")) return rc;
  SAGNN_REQUIRE(p->valsum && data, SAGNN_INVALID_ARG, "norm_data: plan was built without stored values");
  int64_t ru = (int64_t)k * p->N, ri = ru + p->U, n = p->nnz[k];
  if (side == 0)
    edge_norm_kernel<1><<<blocks_for(n), 256, 0, st>>>(p->rowptr, p->idx, p->val, p->deg, p->valsum, ru, p->U,
                                                       ri, p->base[k], n, nullptr, data);
  else
    edge_norm_kernel<1><<<blocks_for(n), 256, 0, st>>>(p->rowptr, p->idx, p->val, p->deg, p->valsum, ri, p->I,
                                                       ru, p->base[k] + n, n, nullptr, data);
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

extern "C" int sagnn_plan_get_weights(const sagnn_plan* p, int k, int side, float* w,
                                      sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_ks(p, k, side, "get_weights")) return rc;
  SAGNN_REQUIRE(p->w && w, SAGNN_INVALID_ARG, "get_weights: plan has no edge weights");
  SAGNN_CUDA(cudaMemcpyAsync(w, p->w + p->base[k] + (side ? p->nnz[k] : 0), sizeof(float) * p->nnz[k],
                             cudaMemcpyDeviceToDevice, st));
  return SAGNN_OK;
}
