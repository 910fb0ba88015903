// Device-side producer of the interval graphs: buckets raw (user, item, timestamp)
// events into T equal time intervals and builds every interval's adjacency list, ready for
// sagnn_plan_set_interval.  Replaces the Python triple loop of `trans_sub`
// (LIU-YUXI/SA-GNN preprocess_to_trnmat.ipynb cell 7): interval id = int((t - minn) / ((maxx - minn) / T))
// clamped to T-1; per (interval, user, item) only the FIRST event in visiting order is kept and its
// timestamp becomes the stored value; csr_matrix((vals,(rows,cols))) then orders each interval
// row-major.  Here: one packed 64-bit key (interval | user | item) per event, ONE stable radix sort
// over the used key bits (CUB), head flags, a scan and a scatter -- the sorted order is at once
// interval-major and row-major inside an interval, and stability makes the first event of a run
// the earliest one.  Bit-exact against the notebook's own output (tests/golden/trnmat_*.npz).
#include <cub/cub.cuh>

#include "common.cuh"

namespace sagnn {

__global__ void event_key_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ it,
                                 const int64_t* __restrict__ t, int64_t n, int U, int I, int T, int64_t minn,
                                 double interval, int bits_u, int bits_i, uint64_t* __restrict__ key,
                                 uint32_t* __restrict__ idx, int* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int64_t uu = u[e], ii = it[e], tt = t[e];
  if (uu < 0 || uu >= U || ii < 0 || ii >= I || tt < minn) {   // reported by the host; keep the key in range meanwhile
    atomicOr(bad, 1);
    uu = 0; ii = 0; tt = minn;
  }
  // int(((one_data - minn) / interval)) of the notebook: exact int -> double, IEEE division, truncation
  long long g = (long long)((double)(tt - minn) / interval);
  if (g >= T) g = T - 1;
  if (g < 0) g = 0;
  key[e] = ((uint64_t)g << (bits_u + bits_i)) | ((uint64_t)(uint32_t)uu << bits_i) | (uint64_t)(uint32_t)ii;
  idx[e] = (uint32_t)e;
}

__global__ void event_head_kernel(const uint64_t* __restrict__ key, int64_t n, int64_t* __restrict__ head) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  head[e] = (e == 0 || key[e] != key[e - 1]) ? 1 : 0;
}

__global__ void event_scatter_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ idx,
                                     const int64_t* __restrict__ head, const int64_t* __restrict__ pos,
                                     const int64_t* __restrict__ t, int64_t n, int bits_u, int bits_i,
                                     int32_t* __restrict__ row, int32_t* __restrict__ col, int32_t* __restrict__ val,
                                     unsigned long long* __restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !head[e]) return;
  const uint64_t k = key[e];
  const int64_t o = pos[e];
  row[o] = (int32_t)((k >> bits_i) & ((1ull << bits_u) - 1));
  col[o] = (int32_t)(k & ((1ull << bits_i) - 1));
  val[o] = (int32_t)t[idx[e]];                       // the first event of the run (stable sort) -- np.intc
  atomicAdd(cnt + (k >> (bits_u + bits_i)), 1ull);
}

template <typename T>
struct EvTmp {
  T* p = nullptr;
  ~EvTmp() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
  operator T*() const { return p; }
};

static int ev_bits(int64_t n) {   // bits needed for values in [0, n)
  int b = 1;
  while (b < 62 && ((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace sagnn

using namespace sagnn;

extern "C" int sagnn_bucket_events(const int32_t* users, const int32_t* items, const int64_t* times, int64_t n,
                                   int U, int I, int T, int64_t minn, int64_t maxx, int32_t* row_out,
                                   int32_t* col_out, int32_t* val_out, int64_t* nnz_host, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(U > 0 && I > 0 && T > 0 && nnz_host, SAGNN_INVALID_ARG, "bucket_events: need U, I, T > 0 and nnz_host");
  SAGNN_REQUIRE(n >= 0 && n < ((int64_t)1 << 32), SAGNN_INVALID_ARG, "bucket_events: n=%lld events (need < 2^32)", (long long)n);
  SAGNN_REQUIRE(maxx > minn, SAGNN_INVALID_ARG,
                "bucket_events: maxx (%lld) must exceed minn (%lld): the notebook's interval width would be 0",
                (long long)maxx, (long long)minn);
  const int bits_u = ev_bits(U), bits_i = ev_bits(I), bits_g = ev_bits(T);
  SAGNN_REQUIRE(bits_u + bits_i + bits_g <= 64, SAGNN_INVALID_ARG, "bucket_events: (T, U, I) do not fit a 64-bit key");
  for (int k = 0; k < T; ++k) nnz_host[k] = 0;
  if (n == 0) return SAGNN_OK;
  SAGNN_REQUIRE(users && items && times && row_out && col_out && val_out, SAGNN_INVALID_ARG, "bucket_events: NULL tensor");
  const double interval = (double)(maxx - minn) / (double)T;     // (maxx-minn)/gragh_num, a Python float
  EvTmp<uint64_t> key, key_s;
  EvTmp<uint32_t> idx, idx_s;
  EvTmp<int64_t> head, pos;
  EvTmp<unsigned long long> cnt;
  EvTmp<int> bad;
  EvTmp<char> tmp;
  SAGNN_CUDA(key.alloc(n)); SAGNN_CUDA(key_s.alloc(n)); SAGNN_CUDA(idx.alloc(n)); SAGNN_CUDA(idx_s.alloc(n));
  SAGNN_CUDA(head.alloc(n)); SAGNN_CUDA(pos.alloc(n)); SAGNN_CUDA(cnt.alloc(T)); SAGNN_CUDA(bad.alloc(1));
  SAGNN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * T, st));
  SAGNN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  event_key_kernel<<<blocks, 256, 0, st>>>(users, items, times, n, U, I, T, minn, interval, bits_u, bits_i, key, idx, bad);
  size_t tb = 0, tb2 = 0;
  const int end_bit = bits_u + bits_i + bits_g;
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key_s.p, idx.p, idx_s.p, n, 0, end_bit, st));
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, head.p, pos.p, n, st));
  SAGNN_CUDA(tmp.alloc(tb > tb2 ? tb : tb2));
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs((void*)tmp.p, tb, key.p, key_s.p, idx.p, idx_s.p, n, 0, end_bit, st));
  event_head_kernel<<<blocks, 256, 0, st>>>(key_s, n, head);
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb2, head.p, pos.p, n, st));
  event_scatter_kernel<<<blocks, 256, 0, st>>>(key_s, idx_s, head, pos, times, n, bits_u, bits_i, row_out, col_out,
                                               val_out, cnt);
  SAGNN_CUDA(cudaGetLastError());
  std::vector<unsigned long long> hc(T);
  int hbad = 0;
  SAGNN_CUDA(cudaMemcpyAsync(hc.data(), cnt, sizeof(unsigned long long) * T, cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  SAGNN_REQUIRE(!hbad, SAGNN_OUT_OF_RANGE, "bucket_events: an event has an id outside [0,%d) x [0,%d) or a timestamp below minn", U, I);
  for (int k = 0; k < T; ++k) nnz_host[k] = (int64_t)hc[k];
  return SAGNN_OK;
}
