// Shared definitions of the sagnn_b200 library (plan layout, error plumbing).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "sagnn_b200.h"

namespace sagnn {

constexpr int kChunk = 64;       // max edges one lane-group gathers for one task
constexpr int kThreads = 256;    // CTA size of the propagation kernel

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

}  // namespace sagnn

// One unit of work for a lane group: a whole short row or one <=kChunk-edge slice of a long row.
struct __align__(16) sagnn_task {
  uint32_t grow;   // global row id
  uint32_t meta;   // edge count (<= kChunk); bit 31 set for a slice of a long row
  int64_t e0;      // first edge (offset into idx)
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

  // device arrays
  int32_t* deg = nullptr;         // [n_rows] structural degrees
  int64_t* rowptr = nullptr;      // [n_rows + 1]
  int32_t* idx = nullptr;         // [2 * e_total] source ids
  int32_t* val = nullptr;         // [2 * e_total] stored values (optional)
  float* w = nullptr;             // [2 * e_total] edge weights (optional)
  int64_t* valsum = nullptr;      // [n_rows] value-sum degrees (optional)

  // schedule (degree-binned): long rows (deg > kChunk) are cut into chunks, listed
  // first and longest-first; short rows follow in descending-degree order.
  uint32_t* order = nullptr;      // [n_short] global row ids
  uint32_t* long_row = nullptr;   // [n_long] global row ids
  int64_t* chunk_base = nullptr;  // [n_long + 1] first chunk of each long row
  uint32_t* chunk_lr = nullptr;   // [n_chunks] long-row rank of each chunk
  sagnn_task* tasks = nullptr;    // [n_chunks + n_short] chunk tasks first (longest rows first), then short rows
  int64_t n_short = 0, n_long = 0, n_chunks = 0;
  int32_t max_deg = 0;

  // host-entry cache (sagnn_propagate_host)
  struct HostCache {
    int L = 0, d = 0;
    float *uE = nullptr, *iE = nullptr, *gU = nullptr, *gI = nullptr;
    float *uO = nullptr, *iO = nullptr, *dU = nullptr, *dI = nullptr;
    void *masks = nullptr, *ws = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> ev;
  } hc;
};

namespace sagnn {
void free_host_cache(sagnn_plan* p);
}
