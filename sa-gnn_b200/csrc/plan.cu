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

#include <cstdlib>

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

// ---- schedule construction ---------------------------------------------------------------
// table id (= segment id) of a global row / sorted position: 2*k + side
__device__ __forceinline__ uint32_t table_of(int64_t g, int64_t N, int U) {
  const int64_t k = g / N;
  return (uint32_t)(2 * k + ((g - k * N) >= U ? 1 : 0));
}
__device__ __forceinline__ int64_t table_row0(uint32_t t, int64_t N, int U) {
  return (int64_t)(t >> 1) * N + ((t & 1) ? U : 0);
}

// sort key: rows grouped by table, long rows (deg > kChunk) first by descending degree, then the short
// rows by descending degree CLASS = number of `bucket`-edge gather blocks (stable => ascending row id
// inside a class).  Lock-step lane groups only need equal block counts, and a class that keeps its rows
// in id order reads its own rows and writes its outputs as a forward-moving window over the tables
// instead of scattered 256-byte rows (scripts/micro/l1tex_cost.cu "task mix": +20-30 % for scattered).
//
// window > 0 (SAGNN_SORT_WINDOW): the short rows are first cut into windows of `window` consecutive row ids and the
// class sort happens INSIDE each window, so that all classes of a window run close in time: the own-row reads and the
// output stores of the whole launch sweep every table ONCE, densely, instead of once per class and sparsely.  Lane
// groups still meet equal classes except at the ~16 class boundaries of a window.
__global__ void sched_key_kernel(const int32_t* __restrict__ deg, int64_t n_rows, int64_t N, int U, int bucket, int window,
                                 uint64_t* __restrict__ key, uint32_t* __restrict__ row) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_rows) return;
  const int d = deg[g];
  const int cls = d > kChunk ? d : (d + bucket - 1) / bucket;   // long rows keep their exact order and stay in front
  const uint32_t t = table_of(g, N, U);
  uint32_t sub = d > kChunk ? (uint32_t)(0x7fffffff - d) : (uint32_t)(0x7fffffff - cls);
  if (window > 0 && d <= kChunk) {
    const uint32_t win = (uint32_t)((g - table_row0(t, N, U)) / window);      // < 2^23: checked by the caller
    sub = 0x80000000u | (win << 8) | (uint32_t)(255 - cls);                   // cls <= kChunk = 64
  }
  key[g] = ((uint64_t)t << 32) | sub;
  row[g] = (uint32_t)g;
}

// hot slots = the kHotRows highest-degree rows of every table; per sorted position also the
// number of tasks / slices / long rows it contributes (inputs of the three scans)
__global__ void sched_slot_kernel(const uint32_t* __restrict__ srow, const int32_t* __restrict__ deg,
                                  int64_t n_rows, int64_t N, int U, int K, int u_begin, int u_end, int i_begin,
                                  int i_end, int32_t* __restrict__ slot_of,
                                  int32_t* __restrict__ hot_ids, int64_t* __restrict__ n_task,
                                  int64_t* __restrict__ n_chunk, int64_t* __restrict__ n_long) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  if (i == n_rows) { n_task[i] = 0; n_chunk[i] = 0; n_long[i] = 0; return; }
  const uint32_t g = srow[i];
  const uint32_t t = table_of(i, N, U);          // sorted positions keep the table layout
  const int64_t pos = i - table_row0(t, N, U);
  slot_of[g] = pos < K ? (int32_t)pos : -1;
  if (pos < K) hot_ids[(int64_t)t * kHotRows + pos] = (int32_t)(g - table_row0(t, N, U));
  const int d = deg[g];
  const bool lg = d > kChunk;
  // rows outside this rank's block (row sharding) contribute nothing to the schedule
  const int64_t r = (int64_t)g - table_row0(t, N, U);
  const bool own = (t & 1) ? (r >= i_begin && r < i_end) : (r >= u_begin && r < u_end);
  const int64_t nt = !own ? 0 : (lg ? (d + kChunk - 1) / kChunk : 1);
  n_task[i] = nt;
  n_chunk[i] = lg ? nt : 0;
  n_long[i] = (lg && own) ? 1 : 0;
}

// kernel-side edge codes: one warp per row, stable partition hot-first (hot = slot, cold = id)
__global__ void sched_encode_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx,
                                    const float* __restrict__ w, const int32_t* __restrict__ slot_of,
                                    int64_t n_rows, int64_t N, int U, int32_t* __restrict__ enc,
                                    float* __restrict__ w_enc, int32_t* __restrict__ nhot_row) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_rows) return;
  const uint32_t t = table_of(g, N, U);
  const int64_t src0 = table_row0(t ^ 1u, N, U);
  const int64_t rs = rowptr[g], re = rowptr[g + 1];
  int nh = 0;
  for (int64_t b0 = rs; b0 < re; b0 += 32) {            // warp-uniform trip count
    const int64_t e = b0 + lane;
    const bool hot = e < re && slot_of[src0 + idx[e]] >= 0;
    nh += __popc(__ballot_sync(0xffffffffu, hot));
  }
  int64_t hp = rs, cp = rs + nh;
  for (int64_t b0 = rs; b0 < re; b0 += 32) {
    const int64_t e = b0 + lane;
    const bool in = e < re;
    int c = 0, s = -1;
    float wv = 0.f;
    if (in) { c = idx[e]; s = slot_of[src0 + c]; if (w) wv = w[e]; }
    const bool hot = in && s >= 0;
    const unsigned hb = __ballot_sync(0xffffffffu, hot), cbal = __ballot_sync(0xffffffffu, in && !hot);
    const unsigned below = (1u << lane) - 1u;
    if (hot) { const int64_t o = hp + __popc(hb & below); enc[o] = s; if (w) w_enc[o] = wv; }
    else if (in) { const int64_t o = cp + __popc(cbal & below); enc[o] = c; if (w) w_enc[o] = wv; }
    hp += __popc(hb);
    cp += __popc(cbal);
  }
  if (lane == 0) nhot_row[g] = nh;
}

// one thread per task: the sorted position that owns it (binary search on the task scan), then
// the slice [rs + deg*ci/nt, rs + deg*(ci+1)/nt) of that row
__global__ void sched_task_kernel(const uint32_t* __restrict__ srow, const int32_t* __restrict__ deg,
                                  const int64_t* __restrict__ rowptr, const int32_t* __restrict__ nhot_row,
                                  const int64_t* __restrict__ task_off, const int64_t* __restrict__ chunk_off,
                                  const int64_t* __restrict__ long_off, const int32_t* __restrict__ enc,
                                  int64_t n_rows, int64_t n_tasks,
                                  int64_t N, int U, sagnn_task* __restrict__ tasks,
                                  int64_t* __restrict__ chunk_base, uint32_t* __restrict__ chunk_lr) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_tasks) return;
  int64_t lo = 0, hi = n_rows;    // largest i with task_off[i] <= j
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (task_off[mid] <= j) lo = mid; else hi = mid;
  }
  const int64_t i = lo;
  const uint32_t g = srow[i];
  const uint32_t t = table_of(i, N, U);
  const int64_t row0 = table_row0(t, N, U);
  const int64_t d = deg[g];
  const bool lg = d > kChunk;
  const int64_t nt = lg ? (d + kChunk - 1) / kChunk : 1;
  const int64_t ci = j - task_off[i];
  const int64_t rs = rowptr[g];
  const int64_t e0 = rs + d * ci / nt, e1 = rs + d * (ci + 1) / nt;
  int64_t nh = rs + nhot_row[g] - e0;
  nh = nh < 0 ? 0 : (nh > e1 - e0 ? e1 - e0 : nh);
  sagnn_task k;
  k.row = (uint32_t)(g - row0);
  k.meta = (uint32_t)(e1 - e0) | ((uint32_t)nh << 8) | (lg ? 0x80000000u : 0u);
  k.e_off = (uint32_t)(e0 - rowptr[row0]);
  k.aux = 0;
  for (int q = 0; q < 4; ++q) k.c[q] = (e0 + q < e1) ? enc[e0 + q] : 0;
  if (lg) {
    const int64_t c = chunk_off[i] + ci;
    k.aux = (uint32_t)c;
    chunk_lr[c] = (uint32_t)long_off[i];
    if (ci == 0) chunk_base[long_off[i]] = chunk_off[i];
  }
  tasks[j] = k;
}

// ---- packed task stream (packet-stream kernel) ----------------------------------------------
// bytes of edge codes (+ weights) task j contributes to its packet: n padded to a multiple of four
__global__ void pkt_size_kernel(const sagnn_task* __restrict__ tasks, int64_t n_tasks, int wmul,
                                int64_t* __restrict__ cbytes) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_tasks) return;
  if (j == n_tasks) { cbytes[j] = 0; return; }
  const uint32_t n = tasks[j].meta & 0x7fu, nh = (tasks[j].meta >> 8) & 0x7fu;
  cbytes[j] = (int64_t)(((nh + 3u) & ~3u) + ((n - nh + 3u) & ~3u)) * 4 * wmul;   // [hot slots | row ids], each padded to 4
}

// one warp per task: its record, codes (and weights) into the packet; packet p of the whole plan starts
// at byte 16*kPktTasks*p + coff[first task of p] (every packet opens with kPktTasks 16-byte records)
__global__ void pkt_fill_kernel(const sagnn_task* __restrict__ tasks, int64_t n_tasks,
                                const int64_t* __restrict__ coff, const sagnn_seg* __restrict__ seg, int S,
                                const int32_t* __restrict__ codes, const float* __restrict__ w,
                                unsigned char* __restrict__ stream, uint32_t* __restrict__ dir) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n_tasks) return;
  int lo = 0, hi = S;                     // segment that owns task j: largest s with task_begin <= j
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seg[mid].task_begin <= j) lo = mid; else hi = mid;
  }
  const sagnn_seg sg = seg[lo];
  const int64_t local = j - sg.task_begin;
  const int64_t pk = sg.pkt_begin + local / kPktTasks;
  const int t = (int)(local % kPktTasks);
  const int64_t f = j - t;                // first task of the packet
  const int64_t pkt_off = (int64_t)16 * kPktTasks * pk + coff[f];
  const uint32_t code_off = (uint32_t)(16 * kPktTasks + (coff[j] - coff[f]));
  const sagnn_task k = tasks[j];
  const uint32_t n = k.meta & 0x7fu, nh = (k.meta >> 8) & 0x7fu, nc = n - nh;
  const uint32_t nh4 = (nh + 3u) & ~3u, nc4 = (nc + 3u) & ~3u;
  unsigned char* pkt = stream + pkt_off;
  if (lane == 0) {
    reinterpret_cast<uint4*>(pkt)[t] = make_uint4(k.row, n | (nh << 8) | (k.meta & 0x80000000u), k.aux, code_off);
    if (t == 0) dir[pk] = (uint32_t)(pkt_off >> 4);
    if (j == sg.task_end - 1)             // the segment's last packet: pad with no-work records
      for (int q = t + 1; q < kPktTasks; ++q) reinterpret_cast<uint4*>(pkt)[q] = make_uint4(0u, 0x40000000u, 0u, 0u);
  }
  // the task's codes are hot-first: [hot slots, padded to 4][source-row ids, padded to 4][weights, same split]
  const int32_t* src = codes + sg.edge_base + k.e_off;
  int32_t* dst = reinterpret_cast<int32_t*>(pkt + code_off);
  for (uint32_t e = lane; e < nh4; e += 32) dst[e] = e < nh ? src[e] : 0;
  for (uint32_t e = lane; e < nc4; e += 32) dst[nh4 + e] = e < nc ? src[nh + e] : 0;
  if (w) {
    const float* ws = w + sg.edge_base + k.e_off;
    float* wd = reinterpret_cast<float*>(pkt + code_off) + nh4 + nc4;
    for (uint32_t e = lane; e < nh4; e += 32) wd[e] = e < nh ? ws[e] : 0.f;
    for (uint32_t e = lane; e < nc4; e += 32) wd[nh4 + e] = e < nc ? ws[nh + e] : 0.f;
  }
}

struct CastI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

// scratch device buffer released on every exit path (error returns included)
template <typename T>
struct DevTmp {
  T* p = nullptr;
  ~DevTmp() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
  operator T*() const { return p; }
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

namespace sagnn {
// persistent CTAs (one per SM) dealt to segments so that max(work / CTAs) is smallest; also the
// per-interval tables (all CTAs on the two segments of one interval)
// deal `n_cta` persistent CTAs to the segments [seg_lo, seg_hi) so that max(work / CTAs) is smallest
static void deal_ctas(const std::vector<double>& cost, int seg_lo, int seg_hi, int n_cta, sagnn_cta* out,
                      std::vector<int>* counts) {
  const int S = seg_hi - seg_lo;
  std::vector<int> n(S, 1);
  for (int left = n_cta - S; left > 0; --left) {   // next CTA -> segment with the largest work per CTA
    int best = 0;
    for (int t = 1; t < S; ++t)
      if (cost[seg_lo + t] / n[t] > cost[seg_lo + best] / n[best]) best = t;
    n[best]++;
  }
  int c = 0;
  for (int t = 0; t < S; ++t)
    for (int r = 0; r < n[t]; ++r) out[c++] = sagnn_cta{seg_lo + t, r, n[t], 0};
  if (counts)
    for (int t = 0; t < S; ++t) (*counts)[seg_lo + t] = n[t];
}

// CTA -> segment tables: one per wave (a wave = as many whole intervals as fit one CTA per segment;
// usually a single wave with all T intervals) and one per interval (pipelined host entry point)
int apply_cta_split(sagnn_plan* p, const std::vector<double>& cost, cudaStream_t st) {
  const int sms = p->num_sms;
  const int per_wave = sms / 2 < 1 ? 1 : sms / 2;          // intervals per wave
  p->n_waves = (p->T + per_wave - 1) / per_wave;
  p->seg_ctas.assign(2 * p->T, 0);
  std::vector<sagnn_cta> cta((size_t)p->n_waves * sms);
  for (int w = 0; w < p->n_waves; ++w) {
    const int k_lo = w * per_wave, k_hi = (w + 1) * per_wave < p->T ? (w + 1) * per_wave : p->T;
    deal_ctas(cost, 2 * k_lo, 2 * k_hi, sms, cta.data() + (size_t)w * sms, &p->seg_ctas);
  }
  if (!p->cta_dev) SAGNN_CUDA(cudaMalloc(&p->cta_dev, sizeof(sagnn_cta) * cta.size()));
  SAGNN_CUDA(cudaMemcpyAsync(p->cta_dev, cta.data(), sizeof(sagnn_cta) * cta.size(), cudaMemcpyHostToDevice, st));
  p->cta_host = cta;
  std::vector<sagnn_cta> cta_int((size_t)p->T * sms);
  for (int k = 0; k < p->T; ++k) deal_ctas(cost, 2 * k, 2 * k + 2, sms, cta_int.data() + (size_t)k * sms, nullptr);
  if (!p->cta_int_dev) SAGNN_CUDA(cudaMalloc(&p->cta_int_dev, sizeof(sagnn_cta) * cta_int.size()));
  SAGNN_CUDA(cudaMemcpyAsync(p->cta_int_dev, cta_int.data(), sizeof(sagnn_cta) * cta_int.size(),
                             cudaMemcpyHostToDevice, st));
  p->cta_int_host = cta_int;
  SAGNN_CUDA(cudaStreamSynchronize(st));
  return SAGNN_OK;
}
}  // namespace sagnn

using namespace sagnn;

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" const char* sagnn_last_error(void) { return sagnn::last_error(); }
extern "C" const char* sagnn_version(void) { return "sagnn_b200 0.2 (sm_100a; packet-stream kernel v10)"; }

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
  cudaError_t e0 = cudaGetDevice(&p->device);
  if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device);
  if (e0 != cudaSuccess) {
    delete p;
    return cuda_fail(e0, "cudaGetDevice / cudaDeviceGetAttribute", __FILE__, __LINE__);
  }
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
  cudaFree(p->valsum); cudaFree(p->chunk_base); cudaFree(p->chunk_lr); cudaFree(p->tasks);
  cudaFree(p->smp_scratch);
  cudaFree(p->enc); cudaFree(p->w_enc); cudaFree(p->pkt_stream); cudaFree(p->pkt_dir); cudaFree(p->hot_ids); cudaFree(p->seg_dev); cudaFree(p->cta_dev); cudaFree(p->cta_int_dev);
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
  DevTmp<int> flags;
  SAGNN_CUDA(flags.alloc(1));
  SAGNN_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));
  check_coo_kernel<<<blocks_for(nnz), 256, 0, st>>>(row, col, nnz, U, I, flags);
  int hflags = 0;
  SAGNN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  SAGNN_CUDA(cudaStreamSynchronize(st));
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
  DevTmp<int32_t> keys_out, perm_in, perm;
  DevTmp<char> tmp;
  size_t tmp_bytes = 0;
  SAGNN_CUDA(keys_out.alloc(nnz));
  SAGNN_CUDA(perm_in.alloc(nnz));
  SAGNN_CUDA(perm.alloc(nnz));
  iota_kernel<<<blocks_for(nnz), 256, 0, st>>>(perm_in, nnz);
  int end_bit = bits_for(I);
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, col, keys_out.p, perm_in.p, perm.p, nnz, 0, end_bit, st));
  SAGNN_CUDA(tmp.alloc(tmp_bytes));
  SAGNN_CUDA(cub::DeviceRadixSort::SortPairs((void*)tmp.p, tmp_bytes, col, keys_out.p, perm_in.p, perm.p, nnz, 0, end_bit, st));
  gather_kernel<int32_t><<<blocks_for(nnz), 256, 0, st>>>(row, perm, nnz, idx_i);
  if (val) gather_kernel<int32_t><<<blocks_for(nnz), 256, 0, st>>>(val, perm, nnz, p->val + p->base[k] + nnz);
  if (w) gather_kernel<float><<<blocks_for(nnz), 256, 0, st>>>(w, perm, nnz, p->w + p->base[k] + nnz);
  SAGNN_CUDA(cudaGetLastError());
  SAGNN_CUDA(cudaStreamSynchronize(st));
  p->is_set[k] = 1;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_rebalance(sagnn_plan* p, const double* seg_work, sagnn_stream_t stream) {
  SAGNN_REQUIRE(p && p->finalized && seg_work, SAGNN_INVALID_ARG, "rebalance: NULL or unfinalized plan");
  std::vector<double> w(seg_work, seg_work + 2 * p->T);
  for (double x : w) SAGNN_REQUIRE(x > 0, SAGNN_INVALID_ARG, "rebalance: work must be positive");
  p->seg_cost = w;
  return apply_cta_split(p, w, (cudaStream_t)stream);
}

extern "C" int sagnn_plan_get_split(const sagnn_plan* p, int* ctas_per_segment) {
  SAGNN_REQUIRE(p && p->finalized && ctas_per_segment, SAGNN_INVALID_ARG, "get_split: NULL or unfinalized plan");
  for (int t = 0; t < 2 * p->T; ++t) ctas_per_segment[t] = p->seg_ctas[t];
  return SAGNN_OK;
}

extern "C" int sagnn_plan_set_row_block(sagnn_plan* p, int u_begin, int u_end, int i_begin, int i_end) {
  SAGNN_REQUIRE(p, SAGNN_INVALID_ARG, "set_row_block: NULL plan");
  SAGNN_REQUIRE(!p->finalized, SAGNN_INVALID_ARG, "set_row_block: plan already finalized");
  SAGNN_REQUIRE(0 <= u_begin && u_begin <= u_end && u_end <= p->U && 0 <= i_begin && i_begin <= i_end && i_end <= p->I,
                SAGNN_INVALID_ARG, "set_row_block: need 0 <= u_begin <= u_end <= %d and 0 <= i_begin <= i_end <= %d "
                "(got [%d,%d) [%d,%d))", p->U, p->I, u_begin, u_end, i_begin, i_end);
  p->u_begin = u_begin; p->u_end = u_end; p->i_begin = i_begin; p->i_end = i_end;
  p->row_block = true;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_set_latdim_hint(sagnn_plan* p, int d) {
  SAGNN_REQUIRE(p && !p->finalized, SAGNN_INVALID_ARG, "set_latdim_hint: NULL or finalized plan");
  SAGNN_REQUIRE(d >= 4 && d % 4 == 0, SAGNN_INVALID_ARG, "set_latdim_hint: d=%d", d);
  p->latdim_hint = d;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_set_hot_rows(sagnn_plan* p, int n) {
  SAGNN_REQUIRE(p && !p->finalized, SAGNN_INVALID_ARG, "set_hot_rows: NULL or finalized plan");
  SAGNN_REQUIRE(n >= 0, SAGNN_INVALID_ARG, "set_hot_rows: n=%d", n);
  p->hot_rows_wanted = n;
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
    DevTmp<char> tmp; size_t tb = 0;
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, p->rowptr, R, st));
    SAGNN_CUDA(tmp.alloc(tb));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, in, p->rowptr, R, st));
    int64_t total = 2 * p->e_total;
    SAGNN_CUDA(cudaMemcpyAsync(p->rowptr + R, &total, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
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

  // ---- schedule: per-segment task lists, hot slots, hot-first edge codes ------------------
  p->pkt = sagnn::plan_uses_pkt(p->latdim_hint);
  // hot slots (sagnn_plan_set_hot_rows, or SAGNN_HOT_ROWS=N for A/B runs): the N highest-degree rows of every
  // table get slot numbers, every task's edge codes list its hot edges first, and the packet-stream kernel
  // stages those rows in shared memory by TMA bulk copies (capped by what fits next to the packet rings at the
  // hinted latdim; the round-1 kernel stages nothing).  OFF unless asked for: measured on B200 (Gowalla shape,
  // 472 staged rows) 0.61 vs 0.47 ms per step -- the staged copy takes the shared memory the L1 cache would
  // otherwise use for exactly those rows, an L1 hit costs the same L1TEX cycles as a shared-memory read, and the
  // extra lock-step phase costs more than the L2 round trips it saves.
  {
    int want = p->hot_rows_wanted;
    if (const char* e = getenv("SAGNN_HOT_ROWS")) want = atoi(e);
    const int cap = p->pkt ? sagnn::pkt_hot_capacity(p->latdim_hint, p->w != nullptr) : 0;
    p->hot_rows = want < 0 ? 0 : (want > cap ? cap : want);
  }
  SAGNN_REQUIRE(p->num_sms >= 2, SAGNN_INVALID_ARG, "finalize: need at least 2 SMs");
  SAGNN_REQUIRE(2 * p->e_total < ((int64_t)1 << 32), SAGNN_INVALID_ARG,
                "finalize: %lld edge entries exceed 2^32", (long long)(2 * p->e_total));
  const int64_t N = p->N;
  const int U = p->U;
  DevTmp<uint64_t> key_in, key_out;
  DevTmp<uint32_t> row_in, srow;
  DevTmp<int32_t> slot_of, nhot_row;
  DevTmp<int64_t> cnt3, off3;   // [3][R+1]: tasks, slices, long rows per sorted position
  SAGNN_CUDA(key_in.alloc(R));
  SAGNN_CUDA(key_out.alloc(R));
  SAGNN_CUDA(row_in.alloc(R));
  SAGNN_CUDA(srow.alloc(R));
  SAGNN_CUDA(slot_of.alloc(R));
  SAGNN_CUDA(nhot_row.alloc(R));
  SAGNN_CUDA(cnt3.alloc(3 * (R + 1)));
  SAGNN_CUDA(off3.alloc(3 * (R + 1)));
  SAGNN_CUDA(cudaMalloc(&p->hot_ids, sizeof(int32_t) * 2 * p->T * kHotRows));
  SAGNN_CUDA(cudaMemsetAsync(p->hot_ids, 0, sizeof(int32_t) * 2 * p->T * kHotRows, st));
  int bucket = p->pkt ? 4 : 1;   // the packet kernel gathers in blocks of four slots
  if (const char* e = getenv("SAGNN_SORT_BUCKET")) bucket = atoi(e) > 0 ? atoi(e) : bucket;
  int window = 0;                // 0: one class sort per table; W > 0: class sort inside windows of W row ids
  if (const char* e = getenv("SAGNN_SORT_WINDOW")) window = atoi(e) > 0 ? atoi(e) : 0;
  if (window > 0 && ((int64_t)(p->U > p->I ? p->U : p->I) / window) >= (1 << 23)) window = 0;
  sched_key_kernel<<<blocks_for(R), 256, 0, st>>>(p->deg, R, N, U, bucket, window, key_in, row_in);
  {
    DevTmp<char> tmp; size_t tb = 0;
    const int end_bit = 32 + bits_for(2 * p->T);
    SAGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_in.p, key_out.p, row_in.p, srow.p, R, 0, end_bit, st));
    SAGNN_CUDA(tmp.alloc(tb));
    SAGNN_CUDA(cub::DeviceRadixSort::SortPairs((void*)tmp.p, tb, key_in.p, key_out.p, row_in.p, srow.p, R, 0, end_bit, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
  }
  int64_t *n_task = cnt3.p, *n_chunk = cnt3.p + (R + 1), *n_long = cnt3.p + 2 * (R + 1);
  int64_t *task_off = off3.p, *chunk_off = off3.p + (R + 1), *long_off = off3.p + 2 * (R + 1);
  if (!p->row_block) { p->u_begin = 0; p->u_end = p->U; p->i_begin = 0; p->i_end = p->I; }
  sched_slot_kernel<<<blocks_for(R + 1), 256, 0, st>>>(srow, p->deg, R, N, U, p->hot_rows, p->u_begin, p->u_end,
                                                       p->i_begin, p->i_end, slot_of, p->hot_ids, n_task,
                                                       n_chunk, n_long);
  {
    DevTmp<char> tmp; size_t tb = 0;
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, n_task, task_off, R + 1, st));
    SAGNN_CUDA(tmp.alloc(tb));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, n_task, task_off, R + 1, st));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, n_chunk, chunk_off, R + 1, st));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, n_long, long_off, R + 1, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
  }
  SAGNN_CUDA(cudaMemcpy(&p->n_tasks, task_off + R, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SAGNN_CUDA(cudaMemcpy(&p->n_chunks, chunk_off + R, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SAGNN_CUDA(cudaMemcpy(&p->n_long, long_off + R, sizeof(int64_t), cudaMemcpyDeviceToHost));
  p->n_short = (p->row_block ? (int64_t)p->T * ((p->u_end - p->u_begin) + (p->i_end - p->i_begin)) : R) - p->n_long;
  {
    DevTmp<char> tmp; size_t tb = 0;
    DevTmp<int32_t> dmax;
    SAGNN_CUDA(dmax.alloc(1));
    SAGNN_CUDA(cub::DeviceReduce::Max(nullptr, tb, p->deg, dmax.p, R, st));
    SAGNN_CUDA(tmp.alloc(tb));
    SAGNN_CUDA(cub::DeviceReduce::Max((void*)tmp.p, tb, p->deg, dmax.p, R, st));
    SAGNN_CUDA(cudaMemcpyAsync(&p->max_deg, dmax, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
    
  }

  const bool pkt = p->pkt;
  if (p->hot_rows > 0) {   // hot-first edge codes (and weights) per row
    SAGNN_CUDA(cudaMalloc(&p->enc, sizeof(int32_t) * 2 * p->e_total));
    if (p->w) SAGNN_CUDA(cudaMalloc(&p->w_enc, sizeof(float) * 2 * p->e_total));
    sched_encode_kernel<<<blocks_for(R * 32), 256, 0, st>>>(p->rowptr, p->idx, p->w, slot_of, R, N, U, p->enc,
                                                            p->w_enc, nhot_row);
  } else {      // no hot slots (always so for the round-1 kernel): the codes ARE the canonical CSR, no second copy
    SAGNN_CUDA(cudaMemsetAsync(nhot_row, 0, sizeof(int32_t) * R, st));
  }

  SAGNN_CUDA(cudaMalloc(&p->tasks, sizeof(sagnn_task) * (p->n_tasks ? p->n_tasks : 1)));
  SAGNN_CUDA(cudaMalloc(&p->chunk_base, sizeof(int64_t) * (p->n_long + 1)));
  SAGNN_CUDA(cudaMalloc(&p->chunk_lr, sizeof(uint32_t) * (p->n_chunks ? p->n_chunks : 1)));
  SAGNN_CUDA(cudaMemcpyAsync(p->chunk_base + p->n_long, &p->n_chunks, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  sched_task_kernel<<<blocks_for(p->n_tasks), 256, 0, st>>>(srow, p->deg, p->rowptr, nhot_row, task_off, chunk_off,
                                                            long_off, p->enc ? p->enc : p->idx, R, p->n_tasks, N, U, p->tasks, p->chunk_base,
                                                            p->chunk_lr);
  SAGNN_CUDA(cudaGetLastError());

  // per-segment descriptors (task ranges come from the scan at the table boundaries)
  const int S = 2 * p->T;
  p->seg_host.resize(S);
  {
    std::vector<int64_t> tb(S + 1);
    for (int t = 0; t <= S; ++t) {
      const int64_t pos = t == S ? R : (int64_t)(t >> 1) * N + ((t & 1) ? U : 0);
      SAGNN_CUDA(cudaMemcpyAsync(&tb[t], task_off + pos, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
    SAGNN_CUDA(cudaStreamSynchronize(st));
    for (int t = 0; t < S; ++t) {
      p->seg_host[t].edge_base = p->base[t >> 1] + ((t & 1) ? p->nnz[t >> 1] : 0);
      p->seg_host[t].task_begin = tb[t];
      p->seg_host[t].task_end = tb[t + 1];
    }
    int64_t pk = 0;
    for (int t = 0; t < S; ++t) {
      p->seg_host[t].pkt_begin = pk;
      pk += (tb[t + 1] - tb[t] + kPktTasks - 1) / kPktTasks;
      p->seg_host[t].pkt_end = pk;
    }
    p->n_pkts = pk;
  }
  SAGNN_CUDA(cudaMalloc(&p->seg_dev, sizeof(sagnn_seg) * S));
  SAGNN_CUDA(cudaMemcpyAsync(p->seg_dev, p->seg_host.data(), sizeof(sagnn_seg) * S, cudaMemcpyHostToDevice, st));

  if (pkt) {    // packed task stream: one scan over the tasks' code bytes places every packet
    DevTmp<int64_t> cbytes, coff;
    DevTmp<char> tmp; size_t tb = 0;
    const int64_t nt = p->n_tasks;
    SAGNN_CUDA(cbytes.alloc(nt + 1));
    SAGNN_CUDA(coff.alloc(nt + 1));
    pkt_size_kernel<<<blocks_for(nt + 1), 256, 0, st>>>(p->tasks, nt, p->w ? 2 : 1, cbytes);
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cbytes.p, coff.p, nt + 1, st));
    SAGNN_CUDA(tmp.alloc(tb));
    SAGNN_CUDA(cub::DeviceScan::ExclusiveSum((void*)tmp.p, tb, cbytes.p, coff.p, nt + 1, st));
    int64_t code_bytes = 0;
    SAGNN_CUDA(cudaMemcpyAsync(&code_bytes, coff.p + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SAGNN_CUDA(cudaStreamSynchronize(st));
    const int64_t total = (int64_t)16 * kPktTasks * p->n_pkts + code_bytes;
    SAGNN_REQUIRE((total >> 4) < ((int64_t)1 << 32), SAGNN_INVALID_ARG, "finalize: packet stream of %lld bytes exceeds 2^36",
                  (long long)total);
    p->pkt_stream_bytes = (size_t)total;
    SAGNN_CUDA(cudaMalloc(&p->pkt_stream, total ? total : 16));
    SAGNN_CUDA(cudaMalloc(&p->pkt_dir, sizeof(uint32_t) * (p->n_pkts + 1)));
    const uint32_t end16 = (uint32_t)(total >> 4);
    SAGNN_CUDA(cudaMemcpyAsync(p->pkt_dir + p->n_pkts, &end16, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (nt)
      pkt_fill_kernel<<<blocks_for(nt * 32), 256, 0, st>>>(p->tasks, nt, coff, p->seg_dev, S, p->enc ? p->enc : p->idx,
                                                           p->enc ? p->w_enc : p->w, (unsigned char*)p->pkt_stream, p->pkt_dir);
    SAGNN_CUDA(cudaGetLastError());
    SAGNN_CUDA(cudaStreamSynchronize(st));
    cudaFree(p->tasks);     // the stream carries the records, codes and weights from here on
    p->tasks = nullptr;
    cudaFree(p->enc); p->enc = nullptr;
    cudaFree(p->w_enc); p->w_enc = nullptr;
  }

  // persistent CTAs (one per SM) are dealt to segments in proportion to their cost.  Measured on
  // B200 (scripts/trace_cta.py): a cold (global) edge costs ~4 units, a hot (staged) edge ~0.9,
  // a row ~30 (task bookkeeping + epilogue).
  {
    std::vector<double> cost(S);
    {
      DevTmp<int64_t> hsum;
      DevTmp<char> tmp; size_t tb = 0;
      SAGNN_CUDA(hsum.alloc(S));
      auto in = thrust::make_transform_iterator((const int32_t*)nhot_row, CastI64());
      SAGNN_CUDA(cub::DeviceReduce::Sum(nullptr, tb, in, hsum.p, R, st));
      SAGNN_CUDA(tmp.alloc(tb));
      for (int t = 0; t < S; ++t) {
        const int64_t row0 = (int64_t)(t >> 1) * N + ((t & 1) ? U : 0);
        const int64_t rows = (t & 1) ? p->I : p->U;
        SAGNN_CUDA(cub::DeviceReduce::Sum((void*)tmp.p, tb, in + row0, hsum.p + t, rows, st));
      }
      std::vector<int64_t> hot(S);
      SAGNN_CUDA(cudaMemcpyAsync(hot.data(), hsum, sizeof(int64_t) * S, cudaMemcpyDeviceToHost, st));
      SAGNN_CUDA(cudaStreamSynchronize(st));
      std::vector<int64_t> own_e(S);
      for (int t = 0; t < S; ++t) {   // edges of the rows this rank owns (all of them without row sharding)
        own_e[t] = p->nnz[t >> 1];
        if (p->row_block) {
          const int64_t row0 = (int64_t)(t >> 1) * N + ((t & 1) ? U : 0);
          int64_t a = 0, b = 0;
          SAGNN_CUDA(cudaMemcpy(&a, p->rowptr + row0 + ((t & 1) ? p->i_begin : p->u_begin), sizeof(int64_t), cudaMemcpyDeviceToHost));
          SAGNN_CUDA(cudaMemcpy(&b, p->rowptr + row0 + ((t & 1) ? p->i_end : p->u_end), sizeof(int64_t), cudaMemcpyDeviceToHost));
          own_e[t] = b - a;
        }
      }
      for (int t = 0; t < S; ++t) {
        const double e = (double)own_e[t], h = p->row_block ? 0.0 : (double)hot[t];
        const double rows = (t & 1) ? p->i_end - p->i_begin : p->u_end - p->u_begin;
        cost[t] = 4.0 * (e - h) + 0.9 * h + 30.0 * rows + 1.0;
      }
    }
    p->seg_cost = cost;
    if (int rc = sagnn::apply_cta_split(p, cost, st)) return rc;
  }
  SAGNN_CUDA(cudaGetLastError());
  SAGNN_CUDA(cudaStreamSynchronize(st));
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
