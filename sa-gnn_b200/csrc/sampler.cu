// Device-side sampleSslBatch (LIU-YUXI/SA-GNN model.py:304-339).  The reference, per train
// step and per interval k, densifies `subMat[k][batIds].toarray()` (batch x item floats) only to read
// each batch user's positive items back with argwhere, then draws `np.random.choice(posset, 2*sslNum)`
// in a Python loop.  The plan already holds every A_k as CSR on the device, so the positives of user u
// in interval k are simply idx[rowptr[g] .. rowptr[g+1]) with g = k*(U+I) + u: one scan for the output
// offsets and one kernel that draws the samples with a counter-based generator.
//
// Same output contract as the reference: for batch position b (ascending) with
// s = min(sslNum, |posset| / 2) > 0, `all = choice(posset, 2*s)` (uniform, with replacement), and for
// j < s:  iLocs[cur] = all[j] (positive), iLocs[cur+1] = all[s+j] (negative),
//         uLocs[cur] = uLocs[cur+1] = batIds[b], uLocs_seq[cur] = uLocs_seq[cur+1] = b, cur += 2.
// Users with fewer than two positives in the interval emit nothing.  The random STREAM differs from
// numpy's Mersenne Twister by construction; parity is on the contract above (tests check membership,
// counts, layout, uniformity and seed determinism).
#include <cub/cub.cuh>

#include "common.cuh"

namespace sagnn {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {     // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

__global__ void ssl_count_kernel(const int32_t* __restrict__ deg, const int32_t* __restrict__ bat, int batch, int U,
                                 int64_t row0, int ssl_num, int64_t* __restrict__ cnt, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > batch) return;
  if (b == batch) { cnt[b] = 0; return; }                   // slot of the total after the exclusive scan
  const int u = bat[b];
  if (u < 0 || u >= U) { atomicOr(bad, 1); cnt[b] = 0; return; }
  const int half = deg[row0 + u] / 2;
  cnt[b] = 2 * (int64_t)(half < ssl_num ? half : ssl_num);
}

__global__ void ssl_draw_kernel(const int32_t* __restrict__ deg, const int64_t* __restrict__ rowptr,
                                const int32_t* __restrict__ idx, const int32_t* __restrict__ bat, int batch, int U,
                                int64_t row0, int ssl_num, uint64_t seed, const int64_t* __restrict__ off,
                                int32_t* __restrict__ u_locs, int32_t* __restrict__ i_locs, int32_t* __restrict__ u_seq) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(tid / (2 * ssl_num)), j = (int)(tid % (2 * ssl_num));   // j-th element of `all`
  if (b >= batch) return;
  const int u = bat[b];
  if (u < 0 || u >= U) return;
  const int d = deg[row0 + u];
  const int s = d / 2 < ssl_num ? d / 2 : ssl_num;
  if (j >= 2 * s) return;
  const uint64_t r = mix64(mix64(seed ^ ((uint64_t)b << 32 | (uint32_t)j)) + (uint64_t)row0);
  const int pick = (int)(((r >> 32) * (uint64_t)d) >> 32);                    // uniform in [0, d)
  const int64_t o = off[b] + (j < s ? 2 * j : 2 * (j - s) + 1);               // positives even, negatives odd
  i_locs[o] = idx[rowptr[row0 + u] + pick];
  u_locs[o] = u;
  u_seq[o] = b;
}

template <typename T>
struct SmpTmp {
  T* p = nullptr;
  ~SmpTmp() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
  operator T*() const { return p; }
};

// all T intervals at once: grid.y = interval.  cnt / off are [T][batch + 1]; after ONE exclusive scan over the whole
// array, off[k][b] - off[k][0] is user b's offset inside interval k's output and off[k][batch] - off[k][0] its total
__global__ void ssl_count_all_kernel(const int32_t* __restrict__ deg, const int32_t* __restrict__ bat, int batch, int U,
                                     int64_t N, int ssl_num, int64_t* __restrict__ cnt, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b > batch) return;
  int64_t* c = cnt + (int64_t)k * (batch + 1);
  if (b == batch) { c[b] = 0; return; }
  const int u = bat[b];
  if (u < 0 || u >= U) { atomicOr(bad, 1); c[b] = 0; return; }
  const int half = deg[(int64_t)k * N + u] / 2;
  c[b] = 2 * (int64_t)(half < ssl_num ? half : ssl_num);
}

__global__ void ssl_draw_all_kernel(const int32_t* __restrict__ deg, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ idx, const int32_t* __restrict__ bat, int batch, int U,
                                    int64_t N, int ssl_num, uint64_t seed, const int64_t* __restrict__ off, int64_t cap,
                                    int32_t* __restrict__ u_locs, int32_t* __restrict__ i_locs, int32_t* __restrict__ u_seq,
                                    int64_t* __restrict__ totals) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const int64_t* o_k = off + (int64_t)k * (batch + 1);
  if (tid == 0) totals[k] = o_k[batch] - o_k[0];
  const int b = (int)(tid / (2 * ssl_num)), j = (int)(tid % (2 * ssl_num));
  if (b >= batch) return;
  const int u = bat[b];
  if (u < 0 || u >= U) return;
  const int64_t row0 = (int64_t)k * N;
  const int d = deg[row0 + u];
  const int s = d / 2 < ssl_num ? d / 2 : ssl_num;
  if (j >= 2 * s) return;
  const uint64_t r = mix64(mix64(seed ^ ((uint64_t)b << 32 | (uint32_t)j)) + (uint64_t)row0);   // same draws as the per-interval call
  const int pick = (int)(((r >> 32) * (uint64_t)d) >> 32);
  const int64_t o = (int64_t)k * cap + (o_k[b] - o_k[0]) + (j < s ? 2 * j : 2 * (j - s) + 1);
  i_locs[o] = idx[rowptr[row0 + u] + pick];
  u_locs[o] = u;
  u_seq[o] = b;
}

// grow-only sampler scratch kept in the plan (one user at a time per plan, like the propagation workspace)
static int sampler_scratch(const sagnn_plan* cp, size_t bytes, char** out) {
  sagnn_plan* p = const_cast<sagnn_plan*>(cp);
  if (p->smp_bytes < bytes) {
    cudaFree(p->smp_scratch); p->smp_scratch = nullptr; p->smp_bytes = 0;
    SAGNN_CUDA(cudaMalloc(&p->smp_scratch, bytes));
    p->smp_bytes = bytes;
  }
  *out = (char*)p->smp_scratch;
  return SAGNN_OK;
}

}  // namespace sagnn

using namespace sagnn;

// Difference to the reference worth knowing: posset there is `temLabel != 0` (model.py:317), i.e. stored explicit
// zeros would not be sampled; here every stored edge of the interval counts.  The reference's interval matrices
// store raw timestamps (preprocess_to_trnmat.ipynb:379), never zeros, so the sets are the same on its data.
// The draws come from a counter-based generator: same output contract, not numpy's MT19937 stream.
extern "C" int sagnn_sample_ssl_batch(const sagnn_plan* p, int k, const int32_t* bat_ids, int batch, int ssl_num,
                                      uint64_t seed, int32_t* u_locs, int32_t* i_locs, int32_t* u_locs_seq,
                                      int64_t* n_out_host, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(p && n_out_host, SAGNN_INVALID_ARG, "sample_ssl_batch: NULL plan / n_out_host");
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "sample_ssl_batch: plan not finalized");
  SAGNN_REQUIRE(k >= 0 && k < p->T, SAGNN_INVALID_ARG, "sample_ssl_batch: interval %d outside [0,%d)", k, p->T);
  SAGNN_REQUIRE(batch >= 0 && ssl_num >= 1 && ssl_num <= (1 << 20), SAGNN_INVALID_ARG,
                "sample_ssl_batch: batch=%d sslNum=%d", batch, ssl_num);
  *n_out_host = 0;
  if (batch == 0) return SAGNN_OK;
  SAGNN_REQUIRE(bat_ids && u_locs && i_locs && u_locs_seq, SAGNN_INVALID_ARG, "sample_ssl_batch: NULL tensor");
  const int64_t row0 = (int64_t)k * p->N;                   // user rows of A_k in the global row space
  size_t tb = 0;
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)nullptr, (int64_t*)nullptr, batch + 1, st));
  const size_t arr = align_up(sizeof(int64_t) * (batch + 1), 256);
  char* base = nullptr;
  if (int rc = sampler_scratch(p, 2 * arr + 256 + tb, &base)) return rc;
  int64_t* cnt = (int64_t*)base; int64_t* off = (int64_t*)(base + arr);
  int* bad = (int*)(base + 2 * arr); void* tmp = base + 2 * arr + 256;
  SAGNN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  ssl_count_kernel<<<(batch + 1 + 255) / 256, 256, 0, st>>>(p->deg, bat_ids, batch, p->U, row0, ssl_num, cnt, bad);
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, batch + 1, st));
  const int64_t threads = (int64_t)batch * 2 * ssl_num;
  ssl_draw_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p->deg, p->rowptr, p->idx, bat_ids, batch, p->U, row0,
                                                                    ssl_num, seed, off, u_locs, i_locs, u_locs_seq);
  SAGNN_CUDA(cudaGetLastError());
  int hbad = 0;
  int64_t total = 0;
  SAGNN_CUDA(cudaMemcpyAsync(&total, off + batch, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  SAGNN_REQUIRE(!hbad, SAGNN_OUT_OF_RANGE, "sample_ssl_batch: a batch id is outside [0,%d)", p->U);
  *n_out_host = total;
  return SAGNN_OK;
}

// every interval of the plan in one call: outputs [T, cap] with cap = batch * 2 * ssl_num, n_out_host [T]; two kernels,
// one scan and ONE stream synchronisation per training step instead of T of each (and no allocation)
extern "C" int sagnn_sample_ssl_batch_all(const sagnn_plan* p, const int32_t* bat_ids, int batch, int ssl_num,
                                          uint64_t seed, int32_t* u_locs, int32_t* i_locs, int32_t* u_locs_seq,
                                          int64_t* n_out_host, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(p && n_out_host, SAGNN_INVALID_ARG, "sample_ssl_batch_all: NULL plan / n_out_host");
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "sample_ssl_batch_all: plan not finalized");
  SAGNN_REQUIRE(batch >= 0 && ssl_num >= 1 && ssl_num <= (1 << 20), SAGNN_INVALID_ARG,
                "sample_ssl_batch_all: batch=%d sslNum=%d", batch, ssl_num);
  const int T = p->T;
  for (int k = 0; k < T; ++k) n_out_host[k] = 0;
  if (batch == 0) return SAGNN_OK;
  SAGNN_REQUIRE(bat_ids && u_locs && i_locs && u_locs_seq, SAGNN_INVALID_ARG, "sample_ssl_batch_all: NULL tensor");
  SAGNN_REQUIRE(T <= 65535, SAGNN_INVALID_ARG, "sample_ssl_batch_all: T=%d", T);
  const int64_t n = (int64_t)T * (batch + 1);
  SAGNN_REQUIRE(n < ((int64_t)1 << 31), SAGNN_INVALID_ARG, "sample_ssl_batch_all: T*(batch+1) too large");
  size_t tb = 0;
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)nullptr, (int64_t*)nullptr, (int)n, st));
  const size_t arr = align_up(sizeof(int64_t) * n, 256), tot = align_up(sizeof(int64_t) * T, 256);
  char* base = nullptr;
  if (int rc = sampler_scratch(p, 2 * arr + tot + 256 + tb, &base)) return rc;
  int64_t* cnt = (int64_t*)base; int64_t* off = (int64_t*)(base + arr); int64_t* totals = (int64_t*)(base + 2 * arr);
  int* bad = (int*)(base + 2 * arr + tot); void* tmp = base + 2 * arr + tot + 256;
  SAGNN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  ssl_count_all_kernel<<<dim3((batch + 1 + 255) / 256, T), 256, 0, st>>>(p->deg, bat_ids, batch, p->U, p->N, ssl_num, cnt, bad);
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)n, st));
  const int64_t threads = (int64_t)batch * 2 * ssl_num, cap = threads;
  ssl_draw_all_kernel<<<dim3((unsigned)((threads + 255) / 256), T), 256, 0, st>>>(
      p->deg, p->rowptr, p->idx, bat_ids, batch, p->U, p->N, ssl_num, seed, off, cap, u_locs, i_locs, u_locs_seq, totals);
  SAGNN_CUDA(cudaGetLastError());
  int hbad = 0;
  SAGNN_CUDA(cudaMemcpyAsync(n_out_host, totals, sizeof(int64_t) * T, cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  if (hbad) for (int k = 0; k < T; ++k) n_out_host[k] = 0;
  SAGNN_REQUIRE(!hbad, SAGNN_OUT_OF_RANGE, "sample_ssl_batch_all: a batch id is outside [0,%d)", p->U);
  return SAGNN_OK;
}

namespace sagnn {
// ---------------------------------------------------------------------------------------------------
// Device-side sampleTrainBatch + negSamp (LIU-YUXI/SA-GNN model.py:252-302, DataHandler.py:28-41).
// The reference densifies labelMat[batIds].toarray() (batch x item floats) per step to test `temLabel[item] == 0`
// inside a Python rejection loop; here the same test is a binary search in the user's rows of the T interval CSRs
// the plan already holds (trnMat is the union of the interval matrices, preprocess_to_trnmat.ipynb cells 13-14).
// Contract, per batch position b with u = batIds[b], seq = handler.sequence[u]:
//   posset = seq[:-1]; sampNum = min(train_sample_num, len(posset)); if sampNum == 0 nothing is emitted, choose = 1;
//   else choose = randint(1, max(min(pred_num + 1, len(posset) - 3), 1)), the positive is posset[-choose] (sampNum
//   times) and the negatives are sampNum uniform items with no training interaction of u and not in {seq[-1], tstInt[u]}
//   (negSamp);  uLocs / uLocs_seq / iLocs: positives first (batch order), then the negatives in the same order;
//   hist = posset[:-choose]: sequence[b] = the last pos_length items of hist, right-aligned, mask = 1 on them.
// Rows b >= batch of sequence / mask (the reference pads to args.batch) are zero.  Counter-based generator: the
// contract is pinned, numpy's / random's streams are not reproducible by construction.
__device__ __forceinline__ bool has_item(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx, int T,
                                         int64_t N, int u, int item) {
  for (int k = 0; k < T; ++k) {
    int64_t lo = rowptr[(int64_t)k * N + u], hi = rowptr[(int64_t)k * N + u + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int c = idx[mid];
      if (c == item) return true;
      if (c < item) lo = mid + 1; else hi = mid;
    }
  }
  return false;
}

__global__ void train_count_kernel(const int64_t* __restrict__ seq_ptr, const int32_t* __restrict__ bat, int batch, int U,
                                   int tsn, int pred_num, uint64_t seed, int64_t* __restrict__ cnt,
                                   int32_t* __restrict__ choose, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > batch) return;
  if (b == batch) { cnt[b] = 0; return; }
  const int u = bat[b];
  if (u < 0 || u >= U) { atomicOr(bad, 1); cnt[b] = 0; choose[b] = 1; return; }
  const int64_t len = seq_ptr[u + 1] - seq_ptr[u];
  const int64_t np = len > 0 ? len - 1 : 0;                    // len(posset)
  const int64_t sn = np < tsn ? np : tsn;
  cnt[b] = sn;
  int ch = 1;
  if (sn > 0) {
    int64_t hi = pred_num + 1 < np - 3 ? pred_num + 1 : np - 3;
    hi = hi < 1 ? 1 : hi;                                      // randint(1, hi), both ends included
    const uint64_t r = mix64(mix64(seed ^ 0x7472616eull) + (uint64_t)b);
    ch = 1 + (int)(((r >> 32) * (uint64_t)hi) >> 32);
  }
  choose[b] = ch;
}

__global__ void train_draw_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx, int T, int64_t N,
                                  const int64_t* __restrict__ seq_ptr, const int32_t* __restrict__ seq_items,
                                  const int32_t* __restrict__ tst, const int32_t* __restrict__ bat, int batch, int n_item,
                                  int tsn, uint64_t seed, const int64_t* __restrict__ off, const int32_t* __restrict__ choose,
                                  int32_t* __restrict__ u_locs, int32_t* __restrict__ i_locs, int32_t* __restrict__ u_seq) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(tid / tsn), j = (int)(tid % tsn);
  if (b >= batch) return;
  const int64_t sn = off[b + 1] - off[b];
  if (j >= sn) return;
  const int u = bat[b];
  const int64_t s0 = seq_ptr[u], len = seq_ptr[u + 1] - s0;
  const int64_t half = off[batch];                             // number of positives = number of negatives
  const int64_t o = off[b] + j;
  const int last = seq_items[s0 + len - 1], held = tst ? tst[u] : -1;
  u_locs[o] = u_locs[o + half] = u;
  u_seq[o] = u_seq[o + half] = b;
  i_locs[o] = seq_items[s0 + (len - 1) - choose[b]];           // posset[-choose]
  int neg = 0;
  for (int tries = 0; tries < 4096; ++tries) {                 // negSamp: uniform rejection sampling
    const uint64_t r = mix64(mix64(seed ^ ((uint64_t)b << 32 | (uint32_t)j)) + (uint64_t)tries * 0x9e3779b97f4a7c15ull);
    neg = (int)(((r >> 32) * (uint64_t)n_item) >> 32);
    if (neg != last && neg != held && !has_item(rowptr, idx, T, N, u, neg)) break;
  }
  i_locs[o + half] = neg;
}

__global__ void train_seq_kernel(const int64_t* __restrict__ seq_ptr, const int32_t* __restrict__ seq_items,
                                 const int32_t* __restrict__ bat, int batch, int batch_pad, int pos_length,
                                 const int32_t* __restrict__ choose, int32_t* __restrict__ sequence, float* __restrict__ mask) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(tid / pos_length), c = (int)(tid % pos_length);
  if (b >= batch_pad) return;
  int item = 0;
  float m = 0.f;
  if (b < batch) {
    const int u = bat[b];
    const int64_t s0 = seq_ptr[u], len = seq_ptr[u + 1] - s0;
    int64_t nh = (len > 0 ? len - 1 : 0) - choose[b];          // len(posset[:-choose])
    nh = nh < 0 ? 0 : nh;
    const int64_t keep = nh < pos_length ? nh : pos_length;    // right-aligned: columns [pos_length - keep, pos_length)
    if (c >= pos_length - keep) {
      item = seq_items[s0 + nh - (pos_length - c)];
      m = 1.f;
    }
  }
  sequence[tid] = item;
  mask[tid] = m;
}

}  // namespace sagnn

extern "C" int sagnn_sample_train_batch(const sagnn_plan* p, const int64_t* seq_ptr, const int32_t* seq_items,
                                        const int32_t* tst_int, const int32_t* bat_ids, int batch, int batch_pad,
                                        int train_sample_num, int pred_num, int pos_length, uint64_t seed,
                                        int32_t* u_locs, int32_t* i_locs, int32_t* u_locs_seq, int32_t* sequence,
                                        float* mask, int32_t* choose_out, int64_t* n_out_host, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SAGNN_REQUIRE(p && n_out_host, SAGNN_INVALID_ARG, "sample_train_batch: NULL plan / n_out_host");
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "sample_train_batch: plan not finalized");
  SAGNN_REQUIRE(batch >= 0 && batch_pad >= batch && train_sample_num >= 1 && pred_num >= 0 && pos_length >= 1,
                SAGNN_INVALID_ARG, "sample_train_batch: batch=%d batch_pad=%d train_sample_num=%d pred_num=%d pos_length=%d",
                batch, batch_pad, train_sample_num, pred_num, pos_length);
  *n_out_host = 0;
  SAGNN_REQUIRE(seq_ptr && seq_items && sequence && mask && (batch == 0 || (bat_ids && u_locs && i_locs && u_locs_seq)),
                SAGNN_INVALID_ARG, "sample_train_batch: NULL tensor");
  SmpTmp<int64_t> cnt, off;
  SmpTmp<int32_t> choose;
  SmpTmp<int> bad;
  SmpTmp<char> tmp;
  SAGNN_CUDA(cnt.alloc(batch + 1)); SAGNN_CUDA(off.alloc(batch + 1)); SAGNN_CUDA(choose.alloc(batch + 1)); SAGNN_CUDA(bad.alloc(1));
  SAGNN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  train_count_kernel<<<(batch + 1 + 255) / 256, 256, 0, st>>>(seq_ptr, bat_ids, batch, p->U, train_sample_num, pred_num, seed,
                                                             cnt, choose, bad);
  size_t tb = 0;
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, batch + 1, st));
  SAGNN_CUDA(tmp.alloc(tb));
  SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, cnt.p, off.p, batch + 1, st));
  if (batch > 0) {
    const int64_t threads = (int64_t)batch * train_sample_num;
    train_draw_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p->rowptr, p->idx, p->T, p->N, seq_ptr, seq_items,
                                                                        tst_int, bat_ids, batch, p->I, train_sample_num, seed,
                                                                        off, choose, u_locs, i_locs, u_locs_seq);
  }
  const int64_t cells = (int64_t)batch_pad * pos_length;
  if (cells > 0)
    train_seq_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(seq_ptr, seq_items, bat_ids, batch, batch_pad, pos_length,
                                                                     choose, sequence, mask);
  SAGNN_CUDA(cudaGetLastError());
  if (choose_out && batch > 0)
    SAGNN_CUDA(cudaMemcpyAsync(choose_out, choose.p, sizeof(int32_t) * batch, cudaMemcpyDeviceToDevice, st));
  int hbad = 0;
  int64_t total = 0;
  SAGNN_CUDA(cudaMemcpyAsync(&total, off.p + batch, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
  SAGNN_REQUIRE(!hbad, SAGNN_OUT_OF_RANGE, "sample_train_batch: a batch id is outside [0,%d)", p->U);
  *n_out_host = 2 * total;
  return SAGNN_OK;
}
