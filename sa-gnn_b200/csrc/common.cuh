// Shared definitions of the sagnn_b200 library (plan layout, error plumbing).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "sagnn_b200.h"

namespace sagnn {

constexpr int kChunk = 64;        // max edges one warp gathers for one task
constexpr int kPktTasks = 4;      // tasks per packet of the packed task stream (packet-stream kernel)
constexpr int kHotRows = 768;     // max hot slots per source table (slot = degree rank); array stride
constexpr int kSmemBudget = 226 * 1024; // dynamic shared memory a persistent CTA of the packet-stream kernel may use

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SAGNN_CUDA(call)                                                      \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return ::sagnn::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define SAGNN_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::sagnn::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- packet-stream kernel geometry shared by the plan builder and the kernel ----
#ifndef SAGNN_PKT_X
#define SAGNN_PKT_X 0           // timing experiments (wrong results): 1 = no gathers, 2 = no own rows / epilogue
#endif
#ifndef SAGNN_PKT_PF
#define SAGNN_PKT_PF 0          // 1: L1 prefetch of a task's next gather block while the current one is in flight
#endif
#ifndef SAGNN_PKT_THREADS
#define SAGNN_PKT_THREADS 1024
#endif
#ifndef SAGNN_PKT_SLOTS
#define SAGNN_PKT_SLOTS 3       // packets resident / in flight per warp
#endif
constexpr int kPktThreads = SAGNN_PKT_THREADS;
constexpr int kPktWarps = kPktThreads / 32;

// hot slots a plan may hand out for latdim d (what fits next to the packet ring)
inline int pkt_hot_capacity(int d, bool weighted) {
  const int ns = weighted ? 2 : SAGNN_PKT_SLOTS;
  const size_t pkt = (size_t)kPktWarps * ns * (kPktTasks * 16 + kPktTasks * (kChunk + 4) * 4 * (weighted ? 2 : 1));
  const long cap = ((long)kSmemBudget - (long)pkt) / (4L * d);
  return (int)(cap < 0 ? 0 : (cap > kHotRows ? kHotRows : cap));
}



}  // namespace sagnn

// One unit of work for a lane group: a whole short row or one <=kChunk-edge slice of a long
// row.  Edges of a row are stored hot-first in `enc`: the first n_hot codes are hot-slot
// numbers (rows staged in shared memory), the rest are source-row ids.
struct __align__(32) sagnn_task {
  uint32_t row;     // row id inside its own table (user id or item id)
  uint32_t meta;    // bits 0-6: edges n (<= kChunk); bits 8-14: hot edges; bit 31: slice of a long row
  uint32_t e_off;   // first edge, relative to the segment's first edge
  uint32_t aux;     // slices: global slice id (partial-sum slot)
  int32_t c[4];     // the first four edge codes, so short rows need no dependent code load
};

// A segment = one CSR: (interval k, side).  seg = 2*k + side; its rows are the rows of table
// `seg`, the rows it gathers from are the rows of table `seg ^ 1`.
struct sagnn_seg {
  int64_t edge_base;    // first edge of the segment in idx / enc
  int64_t task_begin;   // its slice of the task list: long-row slices first (longest rows first),
  int64_t task_end;     //   then short rows in descending-degree order
  int64_t pkt_begin;    // its slice of the packet directory (packet-stream kernel): packets of
  int64_t pkt_end;      //   kPktTasks consecutive tasks, the last one padded with no-work records
};

struct sagnn_cta {      // persistent CTA -> segment binding (CTAs are dealt out by segment cost)
  int seg, rank, count, pad;
};

// Global row space: row g = k*(U+I) + (side ? U + r : r); edges of interval k are stored
// as [A_k CSR column ids (item ids) | A_k^T CSR column ids (user ids)] starting at
// idx[2*sum_{j<k} nnz_j], so one exclusive scan of the degrees in global row order is
// the row-pointer array of every CSR at once.
struct sagnn_plan {
  int T = 0, U = 0, I = 0;
  int device = 0;
  int num_sms = 0;
  int64_t N = 0;                  // U + I
  int64_t n_rows = 0;             // T * N
  std::vector<int64_t> nnz;       // per interval
  std::vector<int64_t> base;      // 2 * prefix sum of nnz (edge offset of interval k)
  int64_t e_total = 0;            // sum nnz
  std::vector<char> is_set;
  bool has_val = false;
  bool has_custom_w = false;
  bool finalized = false;
  int weight_mode = 0;
  void* smp_scratch = nullptr;    // sampler scratch (counts, offsets, scan temp), grow-only
  size_t smp_bytes = 0;
  int hot_rows_wanted = 0;        // sagnn_plan_set_hot_rows (before finalize)
  int hot_rows = 0;               // hot slots per source table the edge codes use (packet-stream kernel: what fits at latdim_hint)
  int latdim_hint = 64;
  bool pkt = true;                // schedule built for the packet-stream kernel (else: v8 task records + edge codes)
  // row sharding (sagnn_plan_set_row_block): only user rows [u_begin,u_end) and item rows [i_begin,i_end)
  // of every interval get tasks; the others are some other rank's.  Default: all rows.
  int u_begin = 0, u_end = 0, i_begin = 0, i_end = 0;
  bool row_block = false;

  // canonical CSRs (what transToLsts / transpose produce; parity hooks read these)
  int32_t* deg = nullptr;         // [n_rows] structural degrees
  int64_t* rowptr = nullptr;      // [n_rows + 1]
  int32_t* idx = nullptr;         // [2 * e_total] source ids, canonical order
  int32_t* val = nullptr;         // [2 * e_total] stored values (optional)
  float* w = nullptr;             // [2 * e_total] edge weights, canonical order (optional)
  int64_t* valsum = nullptr;      // [n_rows] value-sum degrees (optional)

  // kernel-side schedule
  int32_t* enc = nullptr;         // [2 * e_total] edge codes, hot-first inside every row
  float* w_enc = nullptr;         // [2 * e_total] weights in enc order (optional)
  int32_t* hot_ids = nullptr;     // [2T, kHotRows] row ids (inside table t) of table t's hot slots
  sagnn_task* tasks = nullptr;    // [n_tasks] grouped by segment (v8 kernel only; the packet stream replaces it)
  // packed task stream (packet-stream kernel): what a warp needs for kPktTasks tasks in one contiguous,
  // 16-byte aligned block = one TMA bulk copy: kPktTasks records {row, n | flags, slice id, code offset}
  // then every task's edge codes (padded to 4) [each followed by its weights]
  uint4* pkt_stream = nullptr;
  uint32_t* pkt_dir = nullptr;    // [n_pkts + 1] packet offsets in 16-byte units
  int64_t n_pkts = 0;
  size_t pkt_stream_bytes = 0;
  sagnn_seg* seg_dev = nullptr;   // [2T]
  sagnn_cta* cta_dev = nullptr;   // [n_waves][num_sms] all intervals: one launch per wave (normally one wave)
  int n_waves = 1;
  sagnn_cta* cta_int_dev = nullptr;  // [T][num_sms] one interval per launch
  std::vector<sagnn_cta> cta_host, cta_int_host;   // host copies: the packet-stream kernel takes its table by value
  int64_t* chunk_base = nullptr;  // [n_long + 1] first slice of each long row
  uint32_t* chunk_lr = nullptr;   // [n_chunks] long-row rank of each slice
  std::vector<sagnn_seg> seg_host;
  std::vector<double> seg_cost;   // per-segment work estimate behind the CTA split
  std::vector<int> seg_ctas;      // CTAs per segment (all-interval launches)
  int64_t n_tasks = 0, n_short = 0, n_long = 0, n_chunks = 0;
  int32_t max_deg = 0;

  // diagnostics (sagnn_debug_trace): per-launch, per-CTA {segment, start, staged, end} timestamps
  unsigned long long* trace_dev = nullptr;
  int trace_capacity = 0;
  mutable int trace_launch = 0;

  // host-entry cache (sagnn_propagate_host)
  struct HostCache {
    int L = 0, d = 0;
    float *uE = nullptr, *iE = nullptr, *gU = nullptr, *gI = nullptr;
    float *uO = nullptr, *iO = nullptr, *dU = nullptr, *dI = nullptr;
    void *masks = nullptr, *ws = nullptr;
    size_t ws_bytes = 0;
    bool masks_valid = false;     // the masks of the last host forward are still in `masks`
    float leaky = 0.f;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> ev;
  } hc;
};

namespace sagnn {
void free_host_cache(sagnn_plan* p);
int apply_cta_split(sagnn_plan* p, const std::vector<double>& cost, cudaStream_t st);
bool use_pkt();   // packet-stream kernel allowed (default); SAGNN_KERNEL=v8 forces the cp.async-ring kernel for every plan
// which kernel a plan's schedule is built for: packet stream (v10) up to latdim 64, cp.async rings (v8) from 128 on --
// measured on the ML-10M shape (d=128, mean user degree 143): v8 4.6 ms, v10 5.5-6.4 ms per step
inline bool plan_uses_pkt(int latdim_hint) {
  const char* e = getenv("SAGNN_KERNEL");
  if (e && (e[0] == 'v' || e[0] == 'V') && e[1] == '1' && e[2] == '0') return true;   // SAGNN_KERNEL=v10: always
  return use_pkt() && latdim_hint < 128;
}
}
