// Fused interval-graph propagation kernels (sm_100a) and their C-ABI drivers.
//
// One launch = one GNN layer over ALL T intervals and BOTH orientations
// (LIU-YUXI/SA-GNN model.py:118-125: the 2T messagePropagate calls of one layer,
// model.py:80-92, plus the residual add and the layer sum of model.py:124-127).
// The backward is the same gather run on the same two CSRs with the sign-masked
// upstream as the source -- no scatter, no float atomics, deterministic.
//
// Work decomposition: ONE WARP OWNS ONE TASK (the 32 lanes span the latent dimension).  A task is
// either a whole short row (deg <= 64) or one <=64-edge slice of a long row; slice sums go through
// a workspace and the warp that finishes a group of 16 slices last (integer ticket) reduces them in
// fixed slice order, so results do not depend on scheduling.  Two kernels share this contract:
// spmm_pkt.cuh (v9, default: TMA-staged packed task stream, static packet interleave) and
// spmm_rpw.cuh (v8, SAGNN_KERNEL=v8: cp.async rings + per-segment queues, kept for A/B runs).
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace sagnn {

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_MSG = 2 };

constexpr int kMaxPeers = 16;
constexpr int kMaxCtas = 160;     // persistent CTAs (one per SM) the by-value CTA table of the packet-stream kernel holds

// what one persistent CTA of the packet-stream kernel works on, passed BY VALUE inside the kernel
// parameters (constant bank): everything derived from it is provably warp-uniform, so the compiler
// keeps segment bases, strides and stream positions in uniform registers
struct PktCta {
  uint32_t seg;        // segment (2*interval + side)
  uint32_t q0;         // first packet of warp 0 inside the segment's packet list
  uint32_t stride;     // packets between two consecutive packets of one warp (= 32 x CTAs of the segment)
  uint32_t pkt_begin;  // the segment's slice of the packet directory
  uint32_t n_pk;
};

struct SpmmParams {
  const uint32_t* pkt_dir;   // packet-stream kernel: packet offsets (16-byte units) ...
  const uint4* pkt_stream;   // ... into the packed records + edge codes (+ weights)
  const sagnn_task* tasks;   // v8 kernel: task records, edge codes, weights
  const int32_t* enc;
  const float* w;            // weights in enc order or NULL
  const int64_t* chunk_base;
  const uint32_t* chunk_lr;
  const int32_t* hot_ids;
  const sagnn_seg* seg;
  const sagnn_cta* cta;
  const sagnn_cta* cta_host; // host copy of `cta` (launch code only)
  int hot_rows;              // hot slots per table used by the plan's edge codes
  int pdl;                   // host only: launch as a programmatic dependent of the kernel before it on the stream
  int single_seg;            // >= 0: every CTA works on this segment (messagePropagate); -1: use cta[]
  int n_seg_total;           // 2T
  int U, I;
  // gather sources: user rows read item-table rows (src_i), item rows read user-table rows (src_u)
  const float* src_u;
  const float* src_i;
  const uint8_t* smask_u;    // BWD: sign masks of the source rows (one byte = 4 sign bits per float4)
  const uint8_t* smask_i;
  const float* a_u;          // FWD: E^l (residual)          BWD: G (dense upstream)
  const float* a_i;
  const float* b_u;          // FWD: sum_{j<l} E^j or NULL   BWD: running gradient g or NULL (== G)
  const float* b_i;
  float* o1_u;               // FWD: E^{l+1} or NULL         BWD / MSG: destination
  float* o1_i;
  float* o2_u;               // FWD: layer-sum output or NULL
  float* o2_i;
  uint8_t* mask_u;           // FWD: sign masks out or NULL
  uint8_t* mask_i;
  float* partials;           // [n_chunks, d]
  uint32_t* tickets;         // slice-reduction tickets, one region per tree level; zero on entry and on exit
  int64_t n_chunks;
  unsigned* ctrs;            // [2T] per-segment task-queue heads, zero on entry
  float leaky;
  int out_add_next;          // FWD: o2 = b + a (+ E^{l+1} when set)
  int T;                     // number of intervals (row stride of the [R,T,d] layout)
  // layout flags (row-per-warp kernel only): 0 = [T,R,d] (default), 1 = [R,T,d], the transposed
  // hand-off of model.py:133-134
  int a_rtd, b_rtd, o2_rtd, src_rtd;
  // row-per-warp backward: sign masks one level down (of the rows being written); the pre-masked
  // copy sigma'(Z^{l-1}) (.) n goes to o2 and is the gather source of the next level
  const uint8_t* pmask_u;
  const uint8_t* pmask_i;
  unsigned long long* trace; // diagnostics: per CTA {seg, t_start, t_staged, t_end} (ns) or NULL
  // fused hand-off (sagnn_propagate_fwd_scatter): the last layer writes row r of its layer sum straight
  // into the receive buffer of the rank that owns row block r / blk -- peer memory over NVLink
  int peer_n, peer_rank, peer_blk_u, peer_blk_i;
  float* peer_u[kMaxPeers];  // receive buffers [peer_n, blk_u, T, d] of every rank (this rank's own included)
  float* peer_i[kMaxPeers];
  PktCta ctad[kMaxCtas];     // packet-stream kernel: per-CTA work descriptor (filled per launch)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float4 ld_nc(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld_stream(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
// strong (L1-bypassing) load at GPU scope: used for data published by other SMs in this launch
__device__ __forceinline__ float4 ld_strong(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// ticket increment with release semantics (MEMBAR.ALL.GPU + atomic, no L1 invalidate)
__device__ __forceinline__ unsigned ticket_release_add(unsigned* p) {
  unsigned old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory");
  return old;
}
// acquire side of the ticket, executed by the LAST arriver of a group only (1 of up to 16 slices): its reads of the
// other slices' partials are ordered after the ticket values it observed.  GPU-scope acquire costs an invalidation
// of the SM's L1 (CCTL.IVALL) -- per group that is affordable, on every ticket (atom.acq_rel) it was not.
__device__ __forceinline__ void ticket_acquire_fence() {
#if !defined(SAGNN_NO_TICKET_ACQUIRE)
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---- TMA bulk copy (global -> shared, mbarrier-completed): stages the hot rows ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- cp.async (LDGSTS): register-free gathers into the per-warp ring ----------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// keep a CTA-lifetime value in a register: the compiler must not rematerialise it from the
// kernel parameters inside the gather loop (it does, under the 64-register cap)
template <typename T>
__device__ __forceinline__ void pin64(T*& p) { asm volatile("" : "+l"(p)); }
__device__ __forceinline__ void pin32(uint32_t& v) { asm volatile("" : "+r"(v)); }

}  // namespace sagnn
#include "spmm_rpw.cuh"
#include "spmm_pkt.cuh"
namespace sagnn {

// Backward, top level: the gather source is sigma'(Z^{L-1}) (.) G.  One streaming pass writes it
// (both tables, 128-bit accesses, one mask byte per float4) so that the gather kernel reads plain
// rows: cheaper than a mask load + selects per gathered edge (12 per row on the Gowalla shape).
// rtd_T > 0: the upstream is laid out [R,T,d] (rtd_T = T, k0 = first interval of the range, rows_u /
// rows_i = U / I); masks and output stay [T,R,d].
__global__ void __launch_bounds__(256)
premask_kernel(const float4* __restrict__ in_u, const float4* __restrict__ in_i, const uint8_t* __restrict__ m_u,
               const uint8_t* __restrict__ m_i, float4* __restrict__ out_u, float4* __restrict__ out_i, int64_t n4_u,
               int64_t n4_i, float leaky, int rtd_T, int k0, int64_t rows_u, int64_t rows_i, int q) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the level kernel after me starts up under my tail
  const int64_t step = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto masked = [&](float4 x, uint32_t b) {
    x.x = (b & 1u) ? x.x : leaky * x.x;
    x.y = (b & 2u) ? x.y : leaky * x.y;
    x.z = (b & 4u) ? x.z : leaky * x.z;
    x.w = (b & 8u) ? x.w : leaky * x.w;
    return x;
  };
  if (rtd_T > 0) {
    // [R,T,d] upstream: walk it in ITS order -- one thread per (row, float4 column), the row's intervals in an inner
    // loop, so a warp reads whole contiguous T*d-float rows (q = d/4 is a power of two: shifts, no divisions);
    // element (r, k0 + k, c) of the input goes to (k, r, c) of the [T,R,d] range the masks and the output use
    const int qs = 31 - __clz(q);
    const int nk_u = (int)(n4_u / (rows_u << qs)), nk_i = rows_i > 0 ? (int)(n4_i / (rows_i << qs)) : 0;
    const int64_t nu = rows_u << qs, n = nu + (rows_i << qs);
    for (int64_t i = i0; i < n; i += step) {
      const bool it = i >= nu;
      const int64_t j = it ? i - nu : i, r = j >> qs, c = j & (q - 1);
      const int64_t rows = it ? rows_i : rows_u;
      const float4* in = it ? in_i : in_u;
      const uint8_t* m = it ? m_i : m_u;
      float4* out = it ? out_i : out_u;
      const int nk = it ? nk_i : nk_u;
      for (int k = 0; k < nk; ++k) {
        const int64_t o = (((int64_t)k * rows + r) << qs) + c;
        out[o] = masked(in[((r * rtd_T + k0 + k) << qs) + c], m[o]);
      }
    }
    return;
  }
  const int64_t n = n4_u + n4_i;
  for (int64_t i = i0; i < n; i += step) {
    const bool it = i >= n4_u;
    const int64_t j = it ? i - n4_u : i;
    (it ? out_i : out_u)[j] = masked(it ? in_i[j] : in_u[j], it ? m_i[j] : m_u[j]);
  }
}

// SAGNN_BWD_PREMASK=0 keeps the per-edge masks of the top backward level (A/B runs)
static bool use_premask(const sagnn_plan* p) {
  static const bool v = [] { const char* e = getenv("SAGNN_BWD_PREMASK"); return !(e && e[0] == '0'); }();
  return v || p->pkt;   // the packet-stream kernel has no per-edge mask path
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
// SAGNN_KERNEL=v8 selects the cp.async-ring kernel (A/B runs); default: packet-stream kernel.
// Read once: the plan's schedule (task records + codes vs packed stream) is built for one of them.
bool use_pkt() {
  static const bool v = [] {
    const char* e = getenv("SAGNN_KERNEL");
    return !(e && (e[0] == 'v' || e[0] == 'V') && e[1] == '8');
  }();
  return v;
}

// SAGNN_PDL=0: plain stream order between the launches of a chain (A/B runs)
static bool use_pdl() {
  static const bool v = [] { const char* e = getenv("SAGNN_PDL"); return !(e && e[0] == '0'); }();
  return v;
}

template <int D, int MODE, bool WEIGHTED, bool RTD, bool HOT>
static int launch_pkt_h(const sagnn_plan* plan, const SpmmParams& prm_in, cudaStream_t st) {
  SpmmParams prm = prm_in;
  if (plan->trace_dev && plan->trace_launch < plan->trace_capacity)   // diagnostics only
    prm.trace = plan->trace_dev + (size_t)(plan->trace_launch++) * plan->num_sms * 4;
  using G = PktGeo<D, WEIGHTED>;
  static_assert(G::SMEM <= (size_t)kSmemBudget, "shared-memory budget exceeded");
  static std::atomic<uint64_t> configured{0};   // bit per device: the attribute is per device
  auto kern = spmm_pkt_kernel<D, MODE, WEIGHTED, RTD, HOT>;
  const uint64_t bit = 1ull << (plan->device & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    SAGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    configured.fetch_or(bit, std::memory_order_release);
  }
  if (plan->n_tasks == 0) return SAGNN_OK;
  SAGNN_REQUIRE(plan->num_sms <= kMaxCtas, SAGNN_INVALID_ARG, "device has %d SMs, the CTA table holds %d", plan->num_sms, kMaxCtas);
  for (int c = 0; c < plan->num_sms; ++c) {
    const int sgi = prm.single_seg >= 0 ? prm.single_seg : prm.cta_host[c].seg;
    const int rank = prm.single_seg >= 0 ? c : prm.cta_host[c].rank;
    const int count = prm.single_seg >= 0 ? plan->num_sms : prm.cta_host[c].count;
    const sagnn_seg& sh = plan->seg_host[sgi];
    prm.ctad[c] = PktCta{(uint32_t)sgi, (uint32_t)rank * kPktWarps, (uint32_t)count * kPktWarps,
                         (uint32_t)sh.pkt_begin, (uint32_t)(sh.pkt_end - sh.pkt_begin)};
  }
  // the staging area is only allocated when the plan has hot slots: without them the L1 keeps that capacity
  const int staged = plan->hot_rows < G::HOT_CAP ? plan->hot_rows : G::HOT_CAP;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->num_sms); cfg.blockDim = dim3(kPktThreads);
  cfg.dynamicSmemBytes = G::PKT_SMEM + (size_t)staged * G::ROWB; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (prm.pdl && use_pdl()) ? 1 : 0;
  SAGNN_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  return SAGNN_OK;
}

// plans with hot slots (SAGNN_HOT_ROWS at plan build) run the instance with the shared-memory staging path
template <int D, int MODE, bool WEIGHTED, bool RTD>
static int launch_pkt_t(const sagnn_plan* plan, const SpmmParams& prm, cudaStream_t st) {
  return plan->hot_rows > 0 ? launch_pkt_h<D, MODE, WEIGHTED, RTD, true>(plan, prm, st)
                            : launch_pkt_h<D, MODE, WEIGHTED, RTD, false>(plan, prm, st);
}

template <int D, int MODE>
static int launch_pkt_v(const sagnn_plan* plan, const SpmmParams& prm, cudaStream_t st) {
  const bool wt = prm.w != nullptr;
  if (MODE != MODE_MSG && (prm.a_rtd | prm.b_rtd | prm.o2_rtd | prm.src_rtd))
    return wt ? launch_pkt_t<D, MODE, true, MODE != MODE_MSG>(plan, prm, st)
              : launch_pkt_t<D, MODE, false, MODE != MODE_MSG>(plan, prm, st);
  return wt ? launch_pkt_t<D, MODE, true, false>(plan, prm, st) : launch_pkt_t<D, MODE, false, false>(plan, prm, st);
}

template <int MODE>
static int launch_pkt_mode(const sagnn_plan* plan, const SpmmParams& prm, int d, cudaStream_t st) {
  SAGNN_REQUIRE(plan->pkt_stream && !prm.smask_u, SAGNN_INVALID_ARG,
                "the packet-stream kernel needs a packed plan and pre-masked backward sources");
  switch (d) {
#ifndef SAGNN_ONLY_D64   // development builds: -DSAGNN_ONLY_D64 instantiates d = 64 only (compile time)
    case 32:  return launch_pkt_v<32, MODE>(plan, prm, st);
    case 128: return launch_pkt_v<128, MODE>(plan, prm, st);
    case 256: return launch_pkt_v<256, MODE>(plan, prm, st);
#endif
    case 64:  return launch_pkt_v<64, MODE>(plan, prm, st);
  }
  set_error("latdim d=%d unsupported (need 32, 64, 128 or 256)", d);
  return SAGNN_INVALID_ARG;
}

template <int VPL, int MODE, bool WEIGHTED, bool MASKED, bool RTD>
static int launch_rpw_t(const sagnn_plan* plan, const SpmmParams& prm_in, cudaStream_t st) {
  SpmmParams prm = prm_in;
  if (plan->trace_dev && plan->trace_launch < plan->trace_capacity)   // diagnostics only
    prm.trace = plan->trace_dev + (size_t)(plan->trace_launch++) * plan->num_sms * 4;
  using G = RowGeo<VPL, MASKED>;
  static_assert(G::SMEM <= 227 * 1024, "shared-memory budget exceeded");
  static std::atomic<uint64_t> configured{0};   // bit per device: the attribute is per device
  auto kern = spmm_rpw_kernel<VPL, MODE, WEIGHTED, MASKED, RTD>;
  const uint64_t bit = 1ull << (plan->device & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    SAGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    configured.fetch_or(bit, std::memory_order_release);
  }
  if (plan->n_tasks == 0) return SAGNN_OK;
  kern<<<plan->num_sms, kRpwThreads, G::SMEM, st>>>(prm);
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

template <int VPL, int MODE, bool WEIGHTED, bool MASKED>
static int launch_rpw_l(const sagnn_plan* plan, const SpmmParams& prm, cudaStream_t st) {
  if (MODE != MODE_MSG && (prm.a_rtd | prm.b_rtd | prm.o2_rtd | prm.src_rtd))
    return launch_rpw_t<VPL, MODE, WEIGHTED, MASKED, MODE != MODE_MSG>(plan, prm, st);
  return launch_rpw_t<VPL, MODE, WEIGHTED, MASKED, false>(plan, prm, st);
}

template <int VPL, int MODE>
static int launch_rpw_v(const sagnn_plan* plan, const SpmmParams& prm, cudaStream_t st) {
  const bool wt = prm.w != nullptr;
  if constexpr (MODE == MODE_BWD) {
    if (prm.smask_u)   // raw upstream as the source: sign masks per edge
      return wt ? launch_rpw_l<VPL, MODE, true, true>(plan, prm, st) : launch_rpw_l<VPL, MODE, false, true>(plan, prm, st);
  }
  return wt ? launch_rpw_l<VPL, MODE, true, false>(plan, prm, st) : launch_rpw_l<VPL, MODE, false, false>(plan, prm, st);
}

template <int MODE>
static int launch_rpw_mode(const sagnn_plan* plan, const SpmmParams& prm, int d, cudaStream_t st) {
  SAGNN_REQUIRE(plan->hot_rows == 0, SAGNN_INVALID_ARG, "the row-per-warp kernel needs a plan without hot slots");
  switch (d) {
#ifndef SAGNN_ONLY_D64
    case 32:  return launch_rpw_v<1, MODE>(plan, prm, st);
    case 128: return launch_rpw_v<4, MODE>(plan, prm, st);
    case 256: return launch_rpw_v<8, MODE>(plan, prm, st);
#endif
    case 64:  return launch_rpw_v<2, MODE>(plan, prm, st);
  }
  set_error("latdim d=%d unsupported (need 32, 64, 128 or 256)", d);
  return SAGNN_INVALID_ARG;
}

static int launch(const sagnn_plan* plan, const SpmmParams& prm, int d, int mode, cudaStream_t st) {
  if (plan->pkt) {
    switch (mode) {
      case MODE_FWD: return launch_pkt_mode<MODE_FWD>(plan, prm, d, st);
      case MODE_BWD: return launch_pkt_mode<MODE_BWD>(plan, prm, d, st);
      default:       return launch_pkt_mode<MODE_MSG>(plan, prm, d, st);
    }
  }
  switch (mode) {
    case MODE_FWD: return launch_rpw_mode<MODE_FWD>(plan, prm, d, st);
    case MODE_BWD: return launch_rpw_mode<MODE_BWD>(plan, prm, d, st);
    default:       return launch_rpw_mode<MODE_MSG>(plan, prm, d, st);
  }
}

static bool d_ok(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }

// workspace layout: [tickets | partials | table buffer 0 | table buffer 1]
struct WsLayout {
  size_t tickets_off, partials_off, buf_off[2], pm_off[2], total;
  size_t fwd_total;      // what a forward / messagePropagate call touches (no pre-masked backward tables)
  size_t ticket_words;   // slice-reduction tickets of all tree levels
  size_t zero_bytes;     // tickets + one set of per-segment queue heads per layer launch, zeroed per call
  size_t table_floats;   // T*(U+I)*d
  size_t user_floats;    // T*U*d  (user part comes first inside a table buffer)
};

static WsLayout ws_layout(const sagnn_plan* p, int n_layers, int d) {
  WsLayout w{};
  size_t off = 0;
  w.tickets_off = off;
  int levels = 1;        // fan-in-16 tree over the slices of the longest row
  for (int64_t n = ((int64_t)p->max_deg + kChunk - 1) / kChunk; n > 16; n = (n + 15) / 16) ++levels;
  w.ticket_words = (size_t)levels * (size_t)(p->n_chunks ? p->n_chunks : 1);
  w.zero_bytes = sizeof(uint32_t) * (w.ticket_words + 2 * (size_t)p->T * n_layers);
  off = align_up(off + w.zero_bytes, 256);
  w.partials_off = off;
  off = align_up(off + sizeof(float) * (size_t)p->n_chunks * d, 256);
  w.table_floats = (size_t)p->n_rows * d;
  w.user_floats = (size_t)p->T * p->U * d;
  int nbuf = n_layers - 1 < 2 ? n_layers - 1 : 2;
  for (int b = 0; b < 2; ++b) {
    w.buf_off[b] = off;
    if (b < nbuf) off = align_up(off + sizeof(float) * w.table_floats, 256);
  }
  w.fwd_total = off;
  const int npm = n_layers < 2 ? n_layers : 2;
  for (int b = 0; b < 2; ++b) {   // backward: pre-masked copies of the upstream / running gradient
    w.pm_off[b] = off;
    if (b < npm) off = align_up(off + sizeof(float) * w.table_floats, 256);
  }
  w.total = off;
  return w;
}

static size_t mask_layer_bytes(const sagnn_plan* p, int d) { return (size_t)p->n_rows * (d / 4); }

static void base_params(const sagnn_plan* p, SpmmParams& s) {
  s = SpmmParams{};
  s.pkt_dir = p->pkt_dir; s.pkt_stream = p->pkt_stream;
  s.tasks = p->tasks; s.enc = p->enc ? p->enc : p->idx;   // without hot slots the edge codes are the CSR's column ids
  s.w = (p->pkt || !p->w_enc) ? p->w : p->w_enc;   // packet stream: weights travel inside the packets, this is only the flag
  s.chunk_base = p->chunk_base; s.chunk_lr = p->chunk_lr;
  s.hot_ids = p->hot_ids; s.seg = p->seg_dev; s.cta = p->cta_dev; s.cta_host = p->cta_host.data(); s.single_seg = -1;
  s.hot_rows = p->hot_rows;
  s.T = p->T;
  s.n_chunks = p->n_chunks;
  s.trace = nullptr;
  s.n_seg_total = 2 * p->T;
  s.U = p->U; s.I = p->I;
}

static int check_common(const sagnn_plan* p, int n_layers, int d, const char* fn) {
  SAGNN_REQUIRE(p, SAGNN_INVALID_ARG, "%s: NULL plan", fn);
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "%s: plan not finalized", fn);
  SAGNN_REQUIRE(n_layers >= 1 && n_layers <= 64, SAGNN_INVALID_ARG, "%s: n_layers=%d", fn, n_layers);
  SAGNN_REQUIRE(d_ok(d), SAGNN_INVALID_ARG, "%s: latdim d=%d unsupported (need 32, 64, 128 or 256)", fn, d);
  return SAGNN_OK;
}

}  // namespace sagnn

using namespace sagnn;

extern "C" int sagnn_debug_trace(sagnn_plan* p, uint64_t* trace_dev, int capacity_launches) {
  SAGNN_REQUIRE(p, SAGNN_INVALID_ARG, "debug_trace: NULL plan");
  p->trace_dev = (unsigned long long*)trace_dev;
  p->trace_capacity = trace_dev ? capacity_launches : 0;
  p->trace_launch = 0;
  return SAGNN_OK;
}

extern "C" int sagnn_plan_stats(const sagnn_plan* p, int64_t* out8) {
  SAGNN_REQUIRE(p && out8, SAGNN_INVALID_ARG, "plan_stats: NULL argument");
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "plan_stats: plan not finalized");
  out8[0] = p->n_rows; out8[1] = p->n_short; out8[2] = p->n_long; out8[3] = p->n_chunks;
  out8[4] = p->max_deg; out8[5] = 2 * p->e_total; out8[6] = p->hot_rows; out8[7] = p->num_sms;
  return SAGNN_OK;
}

extern "C" int sagnn_workspace_bytes(const sagnn_plan* p, int n_layers, int d, size_t* fwd_bytes,
                                     size_t* mask_bytes, size_t* bwd_bytes) {
  if (int rc = check_common(p, n_layers, d, "workspace_bytes")) return rc;
  WsLayout w = ws_layout(p, n_layers, d);
  if (fwd_bytes) *fwd_bytes = w.fwd_total;
  if (bwd_bytes) *bwd_bytes = w.total;
  if (mask_bytes) *mask_bytes = mask_layer_bytes(p, d) * n_layers;
  return SAGNN_OK;
}

struct PeerScatter {   // host-side description of the fused hand-off
  int world, rank;
  const void* const* user_ptrs;
  const void* const* item_ptrs;
};

// interval < 0: all T intervals in one launch per layer; otherwise only that interval's two
// segments (all 148 CTAs work on them) -- lets a caller pipeline copies with compute
static int fwd_impl(const sagnn_plan* p, int interval, const float* uE, const float* iE, float* uOut,
                    float* iOut, int L, int d, float leaky, void* masks, void* ws, size_t ws_bytes,
                    cudaStream_t st, unsigned flags = 0, int l_begin = 0, int l_end = -1,
                    const PeerScatter* ps = nullptr) {
  if (l_end < 0) l_end = L;
  SAGNN_REQUIRE(0 <= l_begin && l_begin <= l_end && l_end <= L, SAGNN_INVALID_ARG,
                "propagate_fwd: layer range [%d,%d) outside [0,%d]", l_begin, l_end, L);
  SAGNN_REQUIRE(!(flags & ~(unsigned)SAGNN_LAYOUT_RTD), SAGNN_INVALID_ARG, "propagate_fwd: unknown flags 0x%x", flags);
  if (int rc = check_common(p, L, d, "propagate_fwd")) return rc;
  SAGNN_REQUIRE(interval < p->T, SAGNN_INVALID_ARG, "propagate_fwd: interval %d outside [0,%d)", interval, p->T);
  SAGNN_REQUIRE(uE && iE && uOut && iOut && ws, SAGNN_INVALID_ARG, "propagate_fwd: NULL tensor");
  WsLayout w = ws_layout(p, L, d);
  SAGNN_REQUIRE(ws_bytes >= w.fwd_total, SAGNN_WORKSPACE_TOO_SMALL,
                "propagate_fwd: workspace %zu < %zu bytes", ws_bytes, w.fwd_total);
  char* base = (char*)ws;
  SpmmParams s;
  base_params(p, s);
  s.tickets = (uint32_t*)(base + w.tickets_off);
  s.partials = (float*)(base + w.partials_off);
  s.leaky = leaky;
  if (interval >= 0) { s.cta = p->cta_int_dev + (size_t)interval * p->num_sms; s.cta_host = p->cta_int_host.data() + (size_t)interval * p->num_sms; }
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, w.zero_bytes, st));
  s.ctrs = s.tickets + w.ticket_words;
  float* buf[2] = {(float*)(base + w.buf_off[0]), (float*)(base + w.buf_off[1])};
  const size_t mlw = mask_layer_bytes(p, d);
  const size_t mu = (size_t)p->T * p->U * (d / 4);
  for (int l = l_begin; l < l_end; ++l) {
    s.ctrs = s.tickets + w.ticket_words + (size_t)l * 2 * p->T;
    const float* cur_u = l == 0 ? uE : buf[(l - 1) & 1];
    const float* cur_i = l == 0 ? iE : buf[(l - 1) & 1] + w.user_floats;
    const bool last = (l == L - 1);
    s.src_u = cur_u; s.src_i = cur_i;
    s.a_u = cur_u;   s.a_i = cur_i;
    s.o1_u = last ? nullptr : buf[l & 1];
    s.o1_i = last ? nullptr : buf[l & 1] + w.user_floats;
    // layer sum: before layer l the output holds sum_{j<l} E^j (j=0 is the input itself)
    if (l == 0) { s.b_u = nullptr; s.b_i = nullptr; }
    else if (l == 1) { s.b_u = uE; s.b_i = iE; }
    else { s.b_u = uOut; s.b_i = iOut; }
    // [R,T,d] hand-off (model.py:133-134): the layer-sum output (and its partial sums read back) are transposed
    s.o2_rtd = (flags & SAGNN_LAYOUT_RTD) ? 1 : 0;
    s.b_rtd = (l >= 2 && (flags & SAGNN_LAYOUT_RTD)) ? 1 : 0;
    const bool write_out = last || l >= 1;
    s.o2_u = write_out ? uOut : nullptr;
    s.o2_i = write_out ? iOut : nullptr;
    s.out_add_next = last ? 1 : 0;
    if (ps && last) {   // the finished layer sums go straight to the ranks that own their row blocks
      s.peer_n = ps->world; s.peer_rank = ps->rank;
      s.peer_blk_u = (p->U + ps->world - 1) / ps->world;
      s.peer_blk_i = (p->I + ps->world - 1) / ps->world;
      for (int r = 0; r < ps->world; ++r) { s.peer_u[r] = (float*)ps->user_ptrs[r]; s.peer_i[r] = (float*)ps->item_ptrs[r]; }
    }
    s.mask_u = masks ? (uint8_t*)masks + (size_t)l * mlw : nullptr;
    s.mask_i = masks ? (uint8_t*)masks + (size_t)l * mlw + mu : nullptr;
    for (int wv = 0; wv < (interval >= 0 ? 1 : p->n_waves); ++wv) {   // one launch unless 2T exceeds the SM count
      if (interval < 0) { s.cta = p->cta_dev + (size_t)wv * p->num_sms; s.cta_host = p->cta_host.data() + (size_t)wv * p->num_sms; }
      s.pdl = (l > l_begin || wv > 0) ? 1 : 0;   // the launch before it on the stream is a kernel of this call
      if (int rc = launch(p, s, d, MODE_FWD, st)) return rc;
    }
  }
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_fwd(const sagnn_plan* p, const float* uE, const float* iE, float* uOut,
                                   float* iOut, int L, int d, float leaky, void* masks, void* ws,
                                   size_t ws_bytes, sagnn_stream_t stream) {
  return fwd_impl(p, -1, uE, iE, uOut, iOut, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int sagnn_propagate_fwd_ex(const sagnn_plan* p, const float* uE, const float* iE, float* uOut,
                                      float* iOut, int L, int d, float leaky, void* masks, void* ws,
                                      size_t ws_bytes, unsigned flags, sagnn_stream_t stream) {
  return fwd_impl(p, -1, uE, iE, uOut, iOut, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream, flags);
}

extern "C" int sagnn_propagate_fwd_scatter(const sagnn_plan* p, const float* uE, const float* iE, float* uOut,
                                           float* iOut, int L, int d, float leaky, void* masks, void* ws,
                                           size_t ws_bytes, int world, int rank, const void* const* user_recv,
                                           const void* const* item_recv, sagnn_stream_t stream) {
  SAGNN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, SAGNN_INVALID_ARG,
                "propagate_fwd_scatter: world=%d rank=%d (need 1 <= world <= %d)", world, rank, kMaxPeers);
  SAGNN_REQUIRE(user_recv && item_recv, SAGNN_INVALID_ARG, "propagate_fwd_scatter: NULL pointer table");
  for (int r = 0; r < world; ++r)
    SAGNN_REQUIRE(user_recv[r] && item_recv[r], SAGNN_INVALID_ARG, "propagate_fwd_scatter: NULL receive buffer of rank %d", r);
  PeerScatter ps{world, rank, user_recv, item_recv};
  return fwd_impl(p, -1, uE, iE, uOut, iOut, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream,
                  SAGNN_LAYOUT_RTD, 0, -1, &ps);
}

extern "C" int sagnn_propagate_fwd_interval(const sagnn_plan* p, int k, const float* uE, const float* iE,
                                            float* uOut, float* iOut, int L, int d, float leaky, void* masks,
                                            void* ws, size_t ws_bytes, sagnn_stream_t stream) {
  SAGNN_REQUIRE(k >= 0, SAGNN_INVALID_ARG, "propagate_fwd_interval: k=%d", k);
  return fwd_impl(p, k, uE, iE, uOut, iOut, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream);
}

static int bwd_impl(const sagnn_plan* p, int interval, const float* gU, const float* gI, float* dU, float* dI,
                    int L, int d, float leaky, const void* masks, void* ws, size_t ws_bytes, cudaStream_t st,
                    unsigned flags = 0, int ph_begin = 0, int ph_end = -1) {
  // phases: 0 = streaming pre-mask of the upstream (row-per-warp kernel), j >= 1 = level kernel of step j-1
  if (ph_end < 0) ph_end = L + 1;
  SAGNN_REQUIRE(0 <= ph_begin && ph_begin <= ph_end && ph_end <= L + 1, SAGNN_INVALID_ARG,
                "propagate_bwd: phase range [%d,%d) outside [0,%d]", ph_begin, ph_end, L + 1);
  const int s_begin = ph_begin > 0 ? ph_begin - 1 : 0, s_end = ph_end - 1;
  if (ph_begin == ph_end) return SAGNN_OK;
  SAGNN_REQUIRE(!(flags & ~(unsigned)SAGNN_LAYOUT_RTD), SAGNN_INVALID_ARG, "propagate_bwd: unknown flags 0x%x", flags);
  if (int rc = check_common(p, L, d, "propagate_bwd")) return rc;
  SAGNN_REQUIRE(interval < p->T, SAGNN_INVALID_ARG, "propagate_bwd: interval %d outside [0,%d)", interval, p->T);
  SAGNN_REQUIRE(gU && gI && dU && dI && masks && ws, SAGNN_INVALID_ARG, "propagate_bwd: NULL tensor");
  WsLayout w = ws_layout(p, L, d);
  SAGNN_REQUIRE(ws_bytes >= w.total, SAGNN_WORKSPACE_TOO_SMALL,
                "propagate_bwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  char* base = (char*)ws;
  SpmmParams s;
  base_params(p, s);
  s.tickets = (uint32_t*)(base + w.tickets_off);
  s.partials = (float*)(base + w.partials_off);
  s.leaky = leaky;
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, w.zero_bytes, st));
  s.ctrs = s.tickets + w.ticket_words;
  float* buf[2] = {(float*)(base + w.buf_off[0]), (float*)(base + w.buf_off[1])};
  float* pmb[2] = {(float*)(base + w.pm_off[0]), (float*)(base + w.pm_off[1])};
  const size_t mlw = mask_layer_bytes(p, d);
  const size_t mu = (size_t)p->T * p->U * (d / 4);
  bool premasked = false;   // the pre-mask pass was launched by this call
  if (interval >= 0) { s.cta = p->cta_int_dev + (size_t)interval * p->num_sms; s.cta_host = p->cta_int_host.data() + (size_t)interval * p->num_sms; }
  for (int l = L - 1 - s_begin, step = s_begin; step < s_end || (step == 0 && ph_begin == 0); --l, ++step) {
    s.ctrs = s.tickets + w.ticket_words + (size_t)step * 2 * p->T;
    // g = total gradient w.r.t. E^{l+1}; at the top level it is the upstream itself
    const float* g_u = step == 0 ? gU : buf[(step - 1) & 1];
    const float* g_i = step == 0 ? gI : buf[(step - 1) & 1] + w.user_floats;
    s.src_u = g_u; s.src_i = g_i;
    s.smask_u = (const uint8_t*)masks + (size_t)l * mlw;        // sigma'(Z0^l): masks user-table rows
    s.smask_i = (const uint8_t*)masks + (size_t)l * mlw + mu;   // sigma'(Z1^l): masks item-table rows
    {
      // below the top level the source is the copy the level above already multiplied by sigma'(Z^l);
      // at the top level one streaming pass makes that copy of the upstream
      if (step == 0 && use_premask(p)) {
        float* pu = pmb[L >= 2 ? 1 : 0]; float* pi = pu + w.user_floats;
        const int64_t ru = interval >= 0 ? p->U : (int64_t)p->T * p->U, ri = interval >= 0 ? p->I : (int64_t)p->T * p->I;
        const int64_t ou = interval >= 0 ? (int64_t)interval * p->U : 0, oi = interval >= 0 ? (int64_t)interval * p->I : 0;
        const int64_t q = d / 4;
        const bool rtd = (flags & SAGNN_LAYOUT_RTD) != 0;   // whole-tensor base + interval offset inside the kernel
        if (ph_begin == 0) {
          premask_kernel<<<p->num_sms * 8, 256, 0, st>>>(
              (const float4*)gU + (rtd ? 0 : ou * q), (const float4*)gI + (rtd ? 0 : oi * q), s.smask_u + ou * q,
              s.smask_i + oi * q, (float4*)pu + ou * q, (float4*)pi + oi * q, ru * q, ri * q, leaky, rtd ? p->T : 0,
              interval >= 0 ? interval : 0, p->U, p->I, (int)q);
          SAGNN_CUDA(cudaGetLastError());
          premasked = true;
        }
        s.src_u = pu; s.src_i = pi;
        s.smask_u = nullptr; s.smask_i = nullptr;
      }
      if (step > 0) {
        s.src_u = pmb[(step - 1) & 1]; s.src_i = pmb[(step - 1) & 1] + w.user_floats;
        s.smask_u = nullptr; s.smask_i = nullptr;
      }
      if (l > 0) {   // hand the next level its pre-masked source: sigma'(Z^{l-1}) (.) n
        s.o2_u = pmb[step & 1]; s.o2_i = pmb[step & 1] + w.user_floats;
        s.pmask_u = (const uint8_t*)masks + (size_t)(l - 1) * mlw;
        s.pmask_i = (const uint8_t*)masks + (size_t)(l - 1) * mlw + mu;
      } else {
        s.o2_u = nullptr; s.o2_i = nullptr; s.pmask_u = nullptr; s.pmask_i = nullptr;
      }
    }
    s.a_u = gU; s.a_i = gI;
    s.a_rtd = (flags & SAGNN_LAYOUT_RTD) ? 1 : 0;            // the dense upstream is read in the caller's layout
    s.src_rtd = (s.a_rtd && s.src_u == gU) ? 1 : 0;          // ... also as the gather source when it is not pre-masked
    s.b_u = step == 0 ? nullptr : g_u;
    s.b_i = step == 0 ? nullptr : g_i;
    s.o1_u = l == 0 ? dU : buf[step & 1];
    s.o1_i = l == 0 ? dI : buf[step & 1] + w.user_floats;
    if (step >= s_end) break;                                  // phase 0 only: just the pre-mask pass
    for (int wv = 0; wv < (interval >= 0 ? 1 : p->n_waves); ++wv) {
      if (interval < 0) { s.cta = p->cta_dev + (size_t)wv * p->num_sms; s.cta_host = p->cta_host.data() + (size_t)wv * p->num_sms; }
      s.pdl = (step > s_begin || wv > 0 || premasked) ? 1 : 0;
      if (int rc = launch(p, s, d, MODE_BWD, st)) return rc;
    }
  }
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_bwd(const sagnn_plan* p, const float* gU, const float* gI, float* dU,
                                   float* dI, int L, int d, float leaky, const void* masks, void* ws,
                                   size_t ws_bytes, sagnn_stream_t stream) {
  return bwd_impl(p, -1, gU, gI, dU, dI, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int sagnn_propagate_bwd_ex(const sagnn_plan* p, const float* gU, const float* gI, float* dU,
                                      float* dI, int L, int d, float leaky, const void* masks, void* ws,
                                      size_t ws_bytes, unsigned flags, sagnn_stream_t stream) {
  return bwd_impl(p, -1, gU, gI, dU, dI, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream, flags);
}

// ---- row sharding: one layer (level) at a time, the caller all-gathers the table in between ----
extern "C" int sagnn_propagate_fwd_layers(const sagnn_plan* p, int l_begin, int l_end, const float* uE,
                                          const float* iE, float* uOut, float* iOut, int L, int d, float leaky,
                                          void* masks, void* ws, size_t ws_bytes, sagnn_stream_t stream) {
  return fwd_impl(p, -1, uE, iE, uOut, iOut, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream, 0, l_begin,
                  l_end);
}

extern "C" int sagnn_propagate_bwd_levels(const sagnn_plan* p, int ph_begin, int ph_end, const float* gU,
                                          const float* gI, float* dU, float* dI, int L, int d, float leaky,
                                          const void* masks, void* ws, size_t ws_bytes, sagnn_stream_t stream) {
  SAGNN_REQUIRE(use_premask(p), SAGNN_INVALID_ARG, "propagate_bwd_levels: needs pre-masked sources (SAGNN_BWD_PREMASK=0 is set)");
  return bwd_impl(p, -1, gU, gI, dU, dI, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream, 0, ph_begin, ph_end);
}

extern "C" int sagnn_workspace_table(const sagnn_plan* p, int L, int d, int which, int index, size_t* offset_bytes,
                                     size_t* user_bytes, size_t* item_bytes) {
  if (int rc = check_common(p, L, d, "workspace_table")) return rc;
  SAGNN_REQUIRE(offset_bytes && (which == 0 || which == 1) && index >= 0 && index < (which ? L : L - 1),
                SAGNN_INVALID_ARG, "workspace_table: which=%d index=%d (forward: 0 <= index < n_layers-1, "
                "backward: 0 <= index < n_layers; n_layers=%d)", which, index, L);
  WsLayout w = ws_layout(p, L, d);
  // forward: E^{index+1}, written by layer `index`; backward: the pre-masked gather source of level
  // step `index` (written by phase `index`: the pre-mask pass for 0, the level kernel of step index-1 after)
  *offset_bytes = which == 0 ? w.buf_off[index & 1]
                             : (index == 0 ? w.pm_off[L >= 2 ? 1 : 0] : w.pm_off[(index - 1) & 1]);
  if (user_bytes) *user_bytes = sizeof(float) * w.user_floats;
  if (item_bytes) *item_bytes = sizeof(float) * (w.table_floats - w.user_floats);
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_bwd_interval(const sagnn_plan* p, int k, const float* gU, const float* gI,
                                            float* dU, float* dI, int L, int d, float leaky, const void* masks,
                                            void* ws, size_t ws_bytes, sagnn_stream_t stream) {
  SAGNN_REQUIRE(k >= 0, SAGNN_INVALID_ARG, "propagate_bwd_interval: k=%d", k);
  return bwd_impl(p, k, gU, gI, dU, dI, L, d, leaky, masks, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int sagnn_message_propagate(const sagnn_plan* p, int k, int side, const float* src, float* out,
                                       int d, float leaky, void* ws, size_t ws_bytes,
                                       sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_common(p, 1, d, "message_propagate")) return rc;
  SAGNN_REQUIRE(k >= 0 && k < p->T && (side == 0 || side == 1), SAGNN_INVALID_ARG,
                "message_propagate: bad interval %d / side %d", k, side);
  SAGNN_REQUIRE(src && out && ws, SAGNN_INVALID_ARG, "message_propagate: NULL tensor");
  WsLayout w = ws_layout(p, 1, d);
  SAGNN_REQUIRE(ws_bytes >= w.fwd_total, SAGNN_WORKSPACE_TOO_SMALL,
                "message_propagate: workspace %zu < %zu bytes", ws_bytes, w.fwd_total);
  char* base = (char*)ws;
  SpmmParams s;
  base_params(p, s);
  s.tickets = (uint32_t*)(base + w.tickets_off);
  s.partials = (float*)(base + w.partials_off);
  s.leaky = leaky;
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, w.zero_bytes, st));
  s.ctrs = s.tickets + w.ticket_words;
  s.single_seg = 2 * k + side;
  // the kernel indexes tables as [T, rows, d]; shift the bases so that interval k lands on the
  // caller's single-interval tensors
  const intptr_t src_shift = (intptr_t)sizeof(float) * (intptr_t)k * (side ? p->U : p->I) * d;
  const intptr_t out_shift = (intptr_t)sizeof(float) * (intptr_t)k * (side ? p->I : p->U) * d;
  const float* vsrc = (const float*)((intptr_t)src - src_shift);
  float* vout = (float*)((intptr_t)out - out_shift);
  if (side) { s.src_u = vsrc; s.o1_i = vout; } else { s.src_i = vsrc; s.o1_u = vout; }
  return launch(p, s, d, MODE_MSG, st);
}

// ---------------------------------------------------------------------------------------
// host-buffer entry point
// ---------------------------------------------------------------------------------------
void sagnn::free_host_cache(sagnn_plan* p) {
  auto& h = p->hc;
  cudaFree(h.uE); cudaFree(h.iE); cudaFree(h.gU); cudaFree(h.gI);
  cudaFree(h.uO); cudaFree(h.iO); cudaFree(h.dU); cudaFree(h.dI);
  cudaFree(h.masks); cudaFree(h.ws);
  if (h.stream) cudaStreamDestroy(h.stream);
  if (h.copy_in) cudaStreamDestroy(h.copy_in);
  if (h.copy_out) cudaStreamDestroy(h.copy_out);
  for (auto e : h.ev) cudaEventDestroy(e);
  h = sagnn_plan::HostCache();
}

// device buffers / streams of the host-buffer entry points, (re)allocated when L or d change
static int host_ensure(sagnn_plan* p, int L, int d) {
  auto& h = p->hc;
  if (h.L == L && h.d == d) return SAGNN_OK;
  const size_t nu = sizeof(float) * (size_t)p->T * p->U * d;
  const size_t ni = sizeof(float) * (size_t)p->T * p->I * d;
  free_host_cache(p);
  size_t fb = 0, mb = 0, bb = 0;
  if (int rc = sagnn_workspace_bytes(p, L, d, &fb, &mb, &bb)) return rc;
  SAGNN_CUDA(cudaMalloc(&h.uE, nu)); SAGNN_CUDA(cudaMalloc(&h.iE, ni));
  SAGNN_CUDA(cudaMalloc(&h.gU, nu)); SAGNN_CUDA(cudaMalloc(&h.gI, ni));
  SAGNN_CUDA(cudaMalloc(&h.uO, nu)); SAGNN_CUDA(cudaMalloc(&h.iO, ni));
  SAGNN_CUDA(cudaMalloc(&h.dU, nu)); SAGNN_CUDA(cudaMalloc(&h.dI, ni));
  SAGNN_CUDA(cudaMalloc(&h.masks, mb ? mb : 1));
  h.ws_bytes = fb > bb ? fb : bb;
  SAGNN_CUDA(cudaMalloc(&h.ws, h.ws_bytes));
  SAGNN_CUDA(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
  SAGNN_CUDA(cudaStreamCreateWithFlags(&h.copy_in, cudaStreamNonBlocking));
  SAGNN_CUDA(cudaStreamCreateWithFlags(&h.copy_out, cudaStreamNonBlocking));
  h.ev.resize(4);
  for (auto& e : h.ev) SAGNN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h.L = L; h.d = d;
  h.masks_valid = false;
  return SAGNN_OK;
}

extern "C" int sagnn_host_forward(sagnn_plan* p, const float* uE, const float* iE, float* uO, float* iO, int L,
                                  int d, float leaky, int keep_masks) {
  if (int rc = check_common(p, L, d, "host_forward")) return rc;
  SAGNN_REQUIRE(uE && iE && uO && iO, SAGNN_INVALID_ARG, "host_forward: NULL embedding/output");
  if (int rc = host_ensure(p, L, d)) return rc;
  auto& h = p->hc;
  const size_t nu = sizeof(float) * (size_t)p->T * p->U * d;
  const size_t ni = sizeof(float) * (size_t)p->T * p->I * d;
  SAGNN_CUDA(cudaMemcpyAsync(h.uE, uE, nu, cudaMemcpyHostToDevice, h.stream));
  SAGNN_CUDA(cudaMemcpyAsync(h.iE, iE, ni, cudaMemcpyHostToDevice, h.stream));
  if (int rc = sagnn_propagate_fwd(p, h.uE, h.iE, h.uO, h.iO, L, d, leaky, keep_masks ? h.masks : nullptr, h.ws,
                                   h.ws_bytes, h.stream))
    return rc;
  SAGNN_CUDA(cudaMemcpyAsync(uO, h.uO, nu, cudaMemcpyDeviceToHost, h.stream));
  SAGNN_CUDA(cudaMemcpyAsync(iO, h.iO, ni, cudaMemcpyDeviceToHost, h.stream));
  SAGNN_CUDA(cudaStreamSynchronize(h.stream));
  h.masks_valid = keep_masks != 0;
  h.leaky = leaky;
  return SAGNN_OK;
}

extern "C" int sagnn_host_backward(sagnn_plan* p, const float* gU, const float* gI, float* dU, float* dI, int L,
                                   int d, float leaky) {
  if (int rc = check_common(p, L, d, "host_backward")) return rc;
  SAGNN_REQUIRE(gU && gI && dU && dI, SAGNN_INVALID_ARG, "host_backward: NULL gradient buffer");
  auto& h = p->hc;
  SAGNN_REQUIRE(h.L == L && h.d == d && h.masks_valid, SAGNN_INVALID_ARG,
                "host_backward: no matching sagnn_host_forward(keep_masks=1) on this plan");
  const size_t nu = sizeof(float) * (size_t)p->T * p->U * d;
  const size_t ni = sizeof(float) * (size_t)p->T * p->I * d;
  SAGNN_CUDA(cudaMemcpyAsync(h.gU, gU, nu, cudaMemcpyHostToDevice, h.stream));
  SAGNN_CUDA(cudaMemcpyAsync(h.gI, gI, ni, cudaMemcpyHostToDevice, h.stream));
  if (int rc = sagnn_propagate_bwd(p, h.gU, h.gI, h.dU, h.dI, L, d, leaky, h.masks, h.ws, h.ws_bytes, h.stream))
    return rc;
  SAGNN_CUDA(cudaMemcpyAsync(dU, h.dU, nu, cudaMemcpyDeviceToHost, h.stream));
  SAGNN_CUDA(cudaMemcpyAsync(dI, h.dI, ni, cudaMemcpyDeviceToHost, h.stream));
  SAGNN_CUDA(cudaStreamSynchronize(h.stream));
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_host(sagnn_plan* p, const float* uE, const float* iE, const float* gU,
                                    const float* gI, float* uO, float* iO, float* dU, float* dI, int L,
                                    int d, float leaky) {
  if (int rc = check_common(p, L, d, "propagate_host")) return rc;
  SAGNN_REQUIRE(uE && iE && uO && iO, SAGNN_INVALID_ARG, "propagate_host: NULL embedding/output");
  const bool bwd = gU != nullptr;
  SAGNN_REQUIRE(!bwd || (gI && dU && dI), SAGNN_INVALID_ARG, "propagate_host: backward needs gI, dU, dI");
  if (int rc = host_ensure(p, L, d)) return rc;
  auto& h = p->hc;
  h.masks_valid = bwd;
  h.leaky = leaky;
  // Pipelined per interval over three streams: embeddings (then upstream gradients) stream in on
  // copy_in, interval k is propagated as soon as its slice has landed, and its outputs
  // (then gradients) stream out on copy_out while later intervals are still computing.
  const int T = p->T;
  if ((int)h.ev.size() < 4 * T) {
    for (auto e : h.ev) cudaEventDestroy(e);
    h.ev.resize(4 * T);
    for (auto& e : h.ev) SAGNN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const size_t su = (size_t)p->U * d, si = (size_t)p->I * d;       // floats per interval slice
  for (int k = 0; k < T; ++k) {
    SAGNN_CUDA(cudaMemcpyAsync(h.uE + k * su, uE + k * su, su * 4, cudaMemcpyHostToDevice, h.copy_in));
    SAGNN_CUDA(cudaMemcpyAsync(h.iE + k * si, iE + k * si, si * 4, cudaMemcpyHostToDevice, h.copy_in));
    SAGNN_CUDA(cudaEventRecord(h.ev[k], h.copy_in));
  }
  if (bwd) {
    for (int k = 0; k < T; ++k) {
      SAGNN_CUDA(cudaMemcpyAsync(h.gU + k * su, gU + k * su, su * 4, cudaMemcpyHostToDevice, h.copy_in));
      SAGNN_CUDA(cudaMemcpyAsync(h.gI + k * si, gI + k * si, si * 4, cudaMemcpyHostToDevice, h.copy_in));
      SAGNN_CUDA(cudaEventRecord(h.ev[T + k], h.copy_in));
    }
  }
  for (int k = 0; k < T; ++k) {
    SAGNN_CUDA(cudaStreamWaitEvent(h.stream, h.ev[k], 0));
    if (int rc = fwd_impl(p, k, h.uE, h.iE, h.uO, h.iO, L, d, leaky, bwd ? h.masks : nullptr, h.ws, h.ws_bytes,
                          h.stream))
      return rc;
    SAGNN_CUDA(cudaEventRecord(h.ev[2 * T + k], h.stream));
    SAGNN_CUDA(cudaStreamWaitEvent(h.copy_out, h.ev[2 * T + k], 0));
    SAGNN_CUDA(cudaMemcpyAsync(uO + k * su, h.uO + k * su, su * 4, cudaMemcpyDeviceToHost, h.copy_out));
    SAGNN_CUDA(cudaMemcpyAsync(iO + k * si, h.iO + k * si, si * 4, cudaMemcpyDeviceToHost, h.copy_out));
  }
  if (bwd) {
    for (int k = 0; k < T; ++k) {
      SAGNN_CUDA(cudaStreamWaitEvent(h.stream, h.ev[T + k], 0));
      if (int rc = bwd_impl(p, k, h.gU, h.gI, h.dU, h.dI, L, d, leaky, h.masks, h.ws, h.ws_bytes, h.stream))
        return rc;
      SAGNN_CUDA(cudaEventRecord(h.ev[3 * T + k], h.stream));
      SAGNN_CUDA(cudaStreamWaitEvent(h.copy_out, h.ev[3 * T + k], 0));
      SAGNN_CUDA(cudaMemcpyAsync(dU + k * su, h.dU + k * su, su * 4, cudaMemcpyDeviceToHost, h.copy_out));
      SAGNN_CUDA(cudaMemcpyAsync(dI + k * si, h.dI + k * si, si * 4, cudaMemcpyDeviceToHost, h.copy_out));
    }
  }
  SAGNN_CUDA(cudaStreamSynchronize(h.copy_out));
  SAGNN_CUDA(cudaStreamSynchronize(h.stream));
  return SAGNN_OK;
}
