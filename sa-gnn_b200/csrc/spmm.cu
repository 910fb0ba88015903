// Fused interval-graph propagation kernels (sm_100a) and their C-ABI drivers.
//
// One launch = one GNN layer over ALL T intervals and BOTH orientations
// (LIU-YUXI/SA-GNN model.py:118-125: the 2T messagePropagate calls of one layer,
// model.py:80-92, plus the residual add and the layer sum of model.py:124-127).
// The backward is the same gather run on the same two CSRs with the sign-masked
// upstream as the source -- no scatter, no float atomics, deterministic.
//
// Work decomposition: a "group" of LPR = d/4 lanes owns one task; each lane keeps a
// float4 (128-bit loads) of the row.  A task is either a whole short row (deg <= 64)
// or one <=64-edge chunk of a long row; chunk partial sums go through a workspace and
// the group that finishes a row last (integer ticket) reduces them in fixed chunk
// order, so results do not depend on scheduling.  Tasks are ordered longest-first and
// dealt round-robin over a grid sized to the SM count x occupancy.
#include <cstdio>

#include "common.cuh"

#ifndef SAGNN_THREADS
#define SAGNN_THREADS 1024     // one persistent CTA per SM
#endif
#ifndef SAGNN_UNR
#define SAGNN_UNR 4            // independent 128-bit gathers in flight per lane
#endif
#ifndef SAGNN_HOT_BYTES
#define SAGNN_HOT_BYTES (192 * 1024)   // shared memory given to staged hot rows
#endif

namespace sagnn {

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_MSG = 2 };
constexpr int kThreads = SAGNN_THREADS;

struct SpmmParams {
  const sagnn_task* tasks;
  const int32_t* enc;
  const float* w;            // weights in enc order or NULL
  const int64_t* chunk_base;
  const uint32_t* chunk_lr;
  const int32_t* hot_ids;
  const sagnn_seg* seg;
  const sagnn_cta* cta;
  int single_seg;            // >= 0: every CTA works on this segment (messagePropagate); -1: use cta[]
  int U, I;
  // gather sources: user rows read item-table rows (src_i), item rows read user-table rows (src_u)
  const float* src_u;
  const float* src_i;
  const uint32_t* smask_u;   // BWD: sign masks of the source rows
  const uint32_t* smask_i;
  const float* a_u;          // FWD: E^l (residual)          BWD: G (dense upstream)
  const float* a_i;
  const float* b_u;          // FWD: sum_{j<l} E^j or NULL   BWD: running gradient g or NULL (== G)
  const float* b_i;
  float* o1_u;               // FWD: E^{l+1} or NULL         BWD / MSG: destination
  float* o1_i;
  float* o2_u;               // FWD: layer-sum output or NULL
  float* o2_i;
  uint32_t* mask_u;          // FWD: sign masks out or NULL
  uint32_t* mask_i;
  float* partials;           // [n_chunks, d]
  uint32_t* tickets;         // [n_long], zero on entry, zero again on exit
  float leaky;
  int out_add_next;          // FWD: o2 = b + a (+ E^{l+1} when set)
};

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
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---- TMA bulk copy (global -> shared, mbarrier-completed) ---------------------------------
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

__device__ __forceinline__ sagnn_task ld_task(const sagnn_task* tasks, int64_t t, int64_t t_end) {
  sagnn_task k;
  if (t < t_end) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(tasks + t));
    k.row = (uint32_t)raw.x; k.meta = (uint32_t)raw.y; k.e_off = (uint32_t)raw.z; k.aux = (uint32_t)raw.w;
  } else {
    k.row = 0; k.meta = 0x40000000u; k.e_off = 0; k.aux = 0;   // bit 30: no work
  }
  return k;
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// One gather slot: slot u of the current block is hot (shared memory) when u < nhb, cold
// (global, read-only path) when nhb <= u < nb, idle otherwise.  Both loads target the same
// registers, so no moves are needed to merge them.
template <int U_>
__device__ __forceinline__ void gather_slot(float4& v, uint32_t hot_addr, const void* gaddr, int nhb, int nb) {
  asm volatile(
      "{\n\t.reg .pred ph, pc;\n\t"
      "setp.gt.s32 ph, %6, %8;\n\t"
      "setp.gt.s32 pc, %7, %8;\n\t"
      "and.pred pc, pc, !ph;\n\t"
      "@ph ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "@pc ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "r"(hot_addr), "l"(gaddr), "r"(nhb), "r"(nb), "n"(U_));
}
// same, with the hot / cold decision already made (warm slots: hot ids that are not staged)
__device__ __forceinline__ void gather_slot_flags(float4& v, uint32_t hot_addr, const void* gaddr, int is_hot,
                                                  int is_cold) {
  asm volatile(
      "{\n\t.reg .pred ph, pc;\n\t"
      "setp.ne.s32 ph, %6, 0;\n\t"
      "setp.ne.s32 pc, %7, 0;\n\t"
      "@ph ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "@pc ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "r"(hot_addr), "l"(gaddr), "r"(is_hot), "r"(is_cold));
}

// keep a CTA-lifetime value in a register: the compiler must not rematerialise it from the
// kernel parameters inside the gather loop (it does, under the 64-register cap)
template <typename T>
__device__ __forceinline__ void pin64(T*& p) { asm volatile("" : "+l"(p)); }
__device__ __forceinline__ void pin32(uint32_t& v) { asm volatile("" : "+r"(v)); }

template <int LPR, int V, int MODE, bool WARM, int KST, int UNR>
__device__ __forceinline__ void gather_block(float4 (&val)[UNR][V], uint32_t (&mw)[UNR][V], bool (&cold)[UNR],
                                             const int (&cs)[UNR], uint32_t hot_lane, const char* src_lane,
                                             const char* smask_lane, const int* warm_ids, int nhb, int nb) {
  constexpr int D = LPR * V * 4;
  constexpr int WPR = D / 32;
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    int c = cs[u];
    bool is_hot = u < nhb;
    if (WARM) {
      if (is_hot && c >= KST) { c = warm_ids[c]; is_hot = false; }
    }
    cold[u] = (u < nb) && !is_hot;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t ha = hot_lane + (uint32_t)c * (D * 4) + v * LPR * 16;
      const char* ga = src_lane + (int64_t)c * (D * 4) + v * LPR * 16;
      if (WARM) gather_slot_flags(val[u][v], ha, ga, is_hot ? 1 : 0, cold[u] ? 1 : 0);
      else if (u == 0) gather_slot<0>(val[u][v], ha, ga, nhb, nb);
      else if (u == 1) gather_slot<1>(val[u][v], ha, ga, nhb, nb);
      else if (u == 2) gather_slot<2>(val[u][v], ha, ga, nhb, nb);
      else if (u == 3) gather_slot<3>(val[u][v], ha, ga, nhb, nb);
      else if (u == 4) gather_slot<4>(val[u][v], ha, ga, nhb, nb);
      else if (u == 5) gather_slot<5>(val[u][v], ha, ga, nhb, nb);
      else if (u == 6) gather_slot<6>(val[u][v], ha, ga, nhb, nb);
      else gather_slot<7>(val[u][v], ha, ga, nhb, nb);
      if (MODE == MODE_BWD) {
        if (cold[u])
          mw[u][v] = __ldg(reinterpret_cast<const uint32_t*>(smask_lane + (int64_t)c * (WPR * 4) + v * (LPR / 8) * 4));
      }
    }
  }
}

struct SegPtrs {               // segment-uniform table pointers, written once per CTA to shared memory
  const float* a;
  const float* b;
  float* o1;
  float* o2;
  uint32_t* mk;
};

// One persistent CTA per SM, bound to one segment (interval, orientation).  It first stages
// the segment's hottest source rows in shared memory with TMA bulk copies, then walks its
// share of the segment's task list.  The lane groups of a warp run in lock step (trip
// counts are warp maxima, loads are predicated), so every shuffle uses the full mask; tasks
// arrive sorted by degree, so the groups sharing a warp have near-equal rows.
template <int LPR, int V, int MODE, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads, 1)
spmm_layer_kernel(const __grid_constant__ SpmmParams p) {
  constexpr int D = LPR * V * 4;
  constexpr int WPR = D / 32;                         // mask words per row
  constexpr int GPW = 32 / LPR;                       // lane groups per warp
  constexpr int UNR = SAGNN_UNR;
  constexpr int KST = (SAGNN_HOT_BYTES / (4 * D)) < kHotRows ? (SAGNN_HOT_BYTES / (4 * D)) : kHotRows;
  constexpr bool WARM = KST < kHotRows;               // hot slots that do not fit: read via their ids
  static_assert(LPR % UNR == 0, "unroll must divide the group width");
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ __align__(128) float hot[];       // [KST][D]
  __shared__ __align__(8) uint64_t bar;
  __shared__ SegPtrs sp;
  __shared__ int warm_ids[WARM ? kHotRows : 1];

  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int gbase = lane - gl;                        // first lane of my group

  int seg, rank, count;
  if (p.single_seg >= 0) { seg = p.single_seg; rank = blockIdx.x; count = gridDim.x; }
  else { const sagnn_cta c = p.cta[blockIdx.x]; seg = c.seg; rank = c.rank; count = c.count; }
  const int k = seg >> 1;
  const bool item_side = seg & 1;
  const int r_own = item_side ? p.I : p.U, r_src = item_side ? p.U : p.I;
  const float* src = (item_side ? p.src_u : p.src_i) + (int64_t)k * r_src * D;
  const uint32_t* smask =
      (MODE == MODE_BWD) ? (item_side ? p.smask_u : p.smask_i) + (int64_t)k * r_src * WPR : nullptr;
  const sagnn_seg sg = p.seg[seg];
  const int32_t* enc = p.enc + sg.edge_base;
  const float* wts = WEIGHTED ? p.w + sg.edge_base : nullptr;

  // ---- stage the hot rows of the source table (TMA bulk copies, one per row) -------------
  {
    const int32_t* ids = p.hot_ids + (int64_t)(seg ^ 1) * kHotRows;
    const int n_stage = r_src < KST ? r_src : KST;
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      const int64_t own0 = (int64_t)k * r_own;
      const float* a = item_side ? p.a_i : p.a_u;
      const float* b = item_side ? p.b_i : p.b_u;
      float* o1 = item_side ? p.o1_i : p.o1_u;
      float* o2 = item_side ? p.o2_i : p.o2_u;
      uint32_t* mk = item_side ? p.mask_i : p.mask_u;
      sp.a = a ? a + own0 * D : nullptr;
      sp.b = b ? b + own0 * D : nullptr;
      sp.o1 = o1 ? o1 + own0 * D : nullptr;
      sp.o2 = o2 ? o2 + own0 * D : nullptr;
      sp.mk = mk ? mk + own0 * WPR : nullptr;
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, (unsigned)(n_stage * D * 4));
    for (int s = threadIdx.x; s < n_stage; s += kThreads)
      tma_bulk_g2s(hot + (size_t)s * D, src + (int64_t)__ldg(ids + s) * D, D * 4, &bar);
    if (WARM) {
      const int n_hot = r_src < kHotRows ? r_src : kHotRows;
      for (int s = threadIdx.x; s < n_hot; s += kThreads) warm_ids[s] = __ldg(ids + s);
    }
    mbar_wait(&bar, 0);
    if (MODE == MODE_BWD) {
      // the backward gathers sigma'(Z) (.) g: mask the staged rows once instead of per edge
      __syncthreads();
      for (int s = threadIdx.x / LPR; s < n_stage; s += kThreads / LPR) {
        const int id = __ldg(ids + s);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4* q = reinterpret_cast<float4*>(hot + (size_t)s * D + (v * LPR + gl) * 4);
          float4 x = *q;
          const uint32_t b = __ldg(smask + (int64_t)id * WPR + ((v * LPR + gl) >> 3)) >> ((gl & 7) * 4);
          x.x = (b & 1u) ? x.x : p.leaky * x.x;
          x.y = (b & 2u) ? x.y : p.leaky * x.y;
          x.z = (b & 4u) ? x.z : p.leaky * x.z;
          x.w = (b & 8u) ? x.w : p.leaky * x.w;
          *q = x;
        }
      }
    }
    __syncthreads();
  }

  // lane-specific bases of everything the gather loop touches, pinned in registers
  const char* src_lane = reinterpret_cast<const char*>(src) + gl * 16;
  const char* smask_lane = reinterpret_cast<const char*>(smask) + (gl >> 3) * 4;
  uint32_t hot_lane = smem_u32(hot) + gl * 16;
  const float leaky = p.leaky;
  pin64(src_lane);
  if (MODE == MODE_BWD) pin64(smask_lane);
  pin32(hot_lane);
  pin64(enc);

  // ---- my share of the segment's tasks ----------------------------------------------------
  const int64_t t_end = sg.task_end;
  const int64_t stride = (int64_t)count * (kThreads / 32) * GPW;
  const int64_t warp_t0 = sg.task_begin + ((int64_t)rank * (kThreads / 32) + threadIdx.x / 32) * GPW;
  int64_t t = warp_t0 + lane / LPR;

  // software pipeline: task records two rounds ahead, the first code batch one round ahead
  sagnn_task nxt = ld_task(p.tasks, t, t_end);
  sagnn_task nxt2 = ld_task(p.tasks, t + stride, t_end);
  int nxt_c = 0;
  float nxt_w = 0.f;
  if (gl < (int)(nxt.meta & 0x7fu)) {
    nxt_c = __ldg(enc + nxt.e_off + gl);
    if (WEIGHTED) nxt_w = __ldg(wts + nxt.e_off + gl);
  }

  for (int64_t tb = warp_t0; tb < t_end; tb += stride, t += stride) {   // warp-uniform trip count
    const sagnn_task cur = nxt;
    int myc = nxt_c;
    float myw = nxt_w;
    nxt = nxt2;
    nxt2 = ld_task(p.tasks, t + 2 * stride, t_end);
    nxt_c = 0;
    if (gl < (int)(nxt.meta & 0x7fu)) {
      nxt_c = __ldg(enc + nxt.e_off + gl);
      if (WEIGHTED) nxt_w = __ldg(wts + nxt.e_off + gl);
    }

    const bool valid = !(cur.meta & 0x40000000u);
    const bool multi = (cur.meta >> 31) != 0;          // slice of a long row
    const int n = (int)(cur.meta & 0x7fu);
    const int nh = (int)((cur.meta >> 8) & 0x7fu);
    const uint32_t own_off = cur.row * (uint32_t)(D * 4) + gl * 16;   // byte offset inside my table (< 4 GB)

    // the row's own dense operand is independent of the gather: issue it first
    float4 own_a[V];
    if (MODE != MODE_MSG) {
      const char* a = reinterpret_cast<const char*>(sp.a);
#pragma unroll
      for (int v = 0; v < V; ++v)
        own_a[v] = valid ? ld_nc(reinterpret_cast<const float*>(a + own_off + v * LPR * 16)) : f4_zero();
    }

    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = f4_zero();

    // ---- gather-reduce over this task's edges (lock step over the warp) -------------------
    int nmax = n;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
    for (int eb = 0; eb < nmax; eb += LPR) {
      // prefetch the next batch of codes while this one is gathered
      int c_next = 0;
      float w_next = 0.f;
      if (eb + LPR + gl < n) {
        c_next = __ldg(enc + cur.e_off + eb + LPR + gl);
        if (WEIGHTED) w_next = __ldg(wts + cur.e_off + eb + LPR + gl);
      }
      const int nbmax = min(LPR, nmax - eb);
      for (int j = 0; j < nbmax; j += UNR) {
        const int nb = n - eb - j;             // my group's edges left from slot j on (may be <= 0)
        const int nhb = nh - eb - j;           // ... of which hot (staged in shared memory)
        float4 val[UNR][V];
        uint32_t mw[UNR][V];
        float wv[UNR];
        bool cold[UNR];
        int cs[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          cs[u] = __shfl_sync(FULL, myc, gbase + j + u);
          if (WEIGHTED) wv[u] = __shfl_sync(FULL, myw, gbase + j + u);
        }
        gather_block<LPR, V, MODE, WARM, KST, UNR>(val, mw, cold, cs, hot_lane, src_lane, smask_lane, warm_ids, nhb, nb);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (u < nb) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              float4 x = val[u][v];
              if (MODE == MODE_BWD) {
                if (cold[u]) {   // cold source = sigma'(Z) (.) g : pass where Z > 0, else leaky
                  const uint32_t b = mw[u][v] >> ((gl & 7) * 4);
                  x.x = (b & 1u) ? x.x : leaky * x.x;
                  x.y = (b & 2u) ? x.y : leaky * x.y;
                  x.z = (b & 4u) ? x.z : leaky * x.z;
                  x.w = (b & 8u) ? x.w : leaky * x.w;
                }
              }
              if (WEIGHTED) {
                acc[v].x = fmaf(wv[u], x.x, acc[v].x);
                acc[v].y = fmaf(wv[u], x.y, acc[v].y);
                acc[v].z = fmaf(wv[u], x.z, acc[v].z);
                acc[v].w = fmaf(wv[u], x.w, acc[v].w);
              } else {
                acc[v] = f4_add(acc[v], x);
              }
            }
          }
        }
      }
      myc = c_next;
      myw = w_next;
    }

    // ---- long rows: publish the partial sum; the last slice to arrive reduces -------------
    // (release-only ticket; the reducer reads with strong loads that bypass L1, so no
    //  acquire fence / L1 invalidate is needed)
    bool finish = valid;
    if (__any_sync(FULL, multi)) {
      uint32_t lr = 0;
      if (multi) {
        lr = __ldg(p.chunk_lr + cur.aux);
        float* mine = p.partials + (int64_t)cur.aux * D;
#pragma unroll
        for (int v = 0; v < V; ++v) st_f4(mine + (v * LPR + gl) * 4, acc[v]);
      }
      __syncwarp();
      unsigned old = 0;
      if (multi && gl == 0) old = ticket_release_add(p.tickets + lr);
      old = __shfl_sync(FULL, old, gbase);
      if (multi) {
        const int64_t cb = __ldg(p.chunk_base + lr);
        const int nch = (int)(__ldg(p.chunk_base + lr + 1) - cb);
        finish = (old == (unsigned)(nch - 1));
        if (finish) {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = f4_zero();
          const float* part = p.partials + cb * D;
          for (int c0 = 0; c0 < nch; c0 += 4) {
            float4 val[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < V; ++v)
                val[u][v] = (c0 + u < nch) ? ld_strong(part + (int64_t)(c0 + u) * D + (v * LPR + gl) * 4)
                                           : f4_zero();
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < V; ++v) acc[v] = f4_add(acc[v], val[u][v]);
          }
          if (gl == 0) p.tickets[lr] = 0u;   // ready for the next launch
        }
      }
      __syncwarp();
    }

    // ---- fused epilogue (all lanes take part in the shuffles; stores are predicated) -------
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t off = own_off + v * LPR * 16;     // bytes
      if (MODE == MODE_BWD) {
        // n = G + g + A (sigma' . g_other)      
        if (finish) {
          const char* g = reinterpret_cast<const char*>(sp.b);
          char* dst = reinterpret_cast<char*>(sp.o1);
          const float4 gv = g ? ld_stream(reinterpret_cast<const float*>(g + off)) : own_a[v];
          st_f4(reinterpret_cast<float*>(dst + off), f4_add(f4_add(own_a[v], gv), acc[v]));
        }
      } else {
        const float4 z = acc[v];
        const float lzx = leaky * z.x, lzy = leaky * z.y, lzz = leaky * z.z, lzw = leaky * z.w;
        // LeakyReLU = max(leaky*z, z)  (Utils/NNLayers.py:135-136)
        const float4 act = make_float4(fmaxf(lzx, z.x), fmaxf(lzy, z.y), fmaxf(lzz, z.z), fmaxf(lzw, z.w));
        if (MODE == MODE_MSG) {
          if (finish) st_f4(reinterpret_cast<float*>(reinterpret_cast<char*>(sp.o1) + off), act);
        } else {
          const char* b = reinterpret_cast<const char*>(sp.b);
          char* o1 = reinterpret_cast<char*>(sp.o1);
          char* o2 = reinterpret_cast<char*>(sp.o2);
          uint32_t* mk = sp.mk;
          const float4 nxt_e = f4_add(own_a[v], act);          // E^{l+1} = E^l + lrelu(Z^l)
          if (finish && o1) st_f4(reinterpret_cast<float*>(o1 + off), nxt_e);
          if (finish && o2) {
            float4 o = own_a[v];
            if (b) o = f4_add(ld_stream(reinterpret_cast<const float*>(b + off)), own_a[v]);
            if (p.out_add_next) o = f4_add(o, nxt_e);
            st_stream(reinterpret_cast<float*>(o2 + off), o);
          }
          if (mk) {   // CTA-uniform
            // TF MaximumGrad sends the gradient to leaky*z where leaky*z >= z: bit = pass-through
            uint32_t word = ((!(lzx >= z.x) ? 1u : 0u) | (!(lzy >= z.y) ? 2u : 0u) |
                             (!(lzz >= z.z) ? 4u : 0u) | (!(lzw >= z.w) ? 8u : 0u)) << ((gl & 7) * 4);
            word |= __shfl_xor_sync(FULL, word, 1);
            word |= __shfl_xor_sync(FULL, word, 2);
            word |= __shfl_xor_sync(FULL, word, 4);
            if (finish && (gl & 7) == 0) mk[(size_t)cur.row * WPR + ((v * LPR + gl) >> 3)] = word;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
template <int LPR, int V, int MODE, bool WEIGHTED>
static int launch_t(const sagnn_plan* plan, const SpmmParams& prm, cudaStream_t st) {
  constexpr int D = LPR * V * 4;
  constexpr int KST = (SAGNN_HOT_BYTES / (4 * D)) < kHotRows ? (SAGNN_HOT_BYTES / (4 * D)) : kHotRows;
  constexpr size_t smem = (size_t)KST * D * 4;
  static bool configured = false;
  auto kern = spmm_layer_kernel<LPR, V, MODE, WEIGHTED>;
  if (!configured) {
    SAGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (plan->n_tasks == 0) return SAGNN_OK;
  kern<<<plan->num_sms, kThreads, smem, st>>>(prm);
  SAGNN_CUDA(cudaGetLastError());
  return SAGNN_OK;
}

template <int MODE>
static int launch_mode(const sagnn_plan* plan, const SpmmParams& prm, int d, cudaStream_t st) {
  const bool wt = prm.w != nullptr;
  switch (d) {
    case 32:  return wt ? launch_t<8, 1, MODE, true>(plan, prm, st)  : launch_t<8, 1, MODE, false>(plan, prm, st);
    case 64:  return wt ? launch_t<16, 1, MODE, true>(plan, prm, st) : launch_t<16, 1, MODE, false>(plan, prm, st);
    case 128: return wt ? launch_t<32, 1, MODE, true>(plan, prm, st) : launch_t<32, 1, MODE, false>(plan, prm, st);
    case 256: return wt ? launch_t<32, 2, MODE, true>(plan, prm, st) : launch_t<32, 2, MODE, false>(plan, prm, st);
  }
  set_error("latdim d=%d unsupported (need 32, 64, 128 or 256)", d);
  return SAGNN_INVALID_ARG;
}

static int launch(const sagnn_plan* plan, const SpmmParams& prm, int d, int mode, cudaStream_t st) {
  switch (mode) {
    case MODE_FWD: return launch_mode<MODE_FWD>(plan, prm, d, st);
    case MODE_BWD: return launch_mode<MODE_BWD>(plan, prm, d, st);
    default:       return launch_mode<MODE_MSG>(plan, prm, d, st);
  }
}

static bool d_ok(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }

// workspace layout: [tickets | partials | table buffer 0 | table buffer 1]
struct WsLayout {
  size_t tickets_off, partials_off, buf_off[2], total;
  size_t table_floats;   // T*(U+I)*d
  size_t user_floats;    // T*U*d  (user part comes first inside a table buffer)
};

static WsLayout ws_layout(const sagnn_plan* p, int n_layers, int d) {
  WsLayout w{};
  size_t off = 0;
  w.tickets_off = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)(p->n_long ? p->n_long : 1), 256);
  w.partials_off = off;
  off = align_up(off + sizeof(float) * (size_t)p->n_chunks * d, 256);
  w.table_floats = (size_t)p->n_rows * d;
  w.user_floats = (size_t)p->T * p->U * d;
  int nbuf = n_layers - 1 < 2 ? n_layers - 1 : 2;
  for (int b = 0; b < 2; ++b) {
    w.buf_off[b] = off;
    if (b < nbuf) off = align_up(off + sizeof(float) * w.table_floats, 256);
  }
  w.total = off;
  return w;
}

static size_t mask_layer_words(const sagnn_plan* p, int d) { return (size_t)p->n_rows * (d / 32); }

static void base_params(const sagnn_plan* p, SpmmParams& s) {
  s = SpmmParams{};
  s.tasks = p->tasks; s.enc = p->enc; s.w = p->w_enc;
  s.chunk_base = p->chunk_base; s.chunk_lr = p->chunk_lr;
  s.hot_ids = p->hot_ids; s.seg = p->seg_dev; s.cta = p->cta_dev; s.single_seg = -1;
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

extern "C" int sagnn_plan_stats(const sagnn_plan* p, int64_t* out8) {
  SAGNN_REQUIRE(p && out8, SAGNN_INVALID_ARG, "plan_stats: NULL argument");
  SAGNN_REQUIRE(p->finalized, SAGNN_NOT_FINALIZED, "plan_stats: plan not finalized");
  out8[0] = p->n_rows; out8[1] = p->n_short; out8[2] = p->n_long; out8[3] = p->n_chunks;
  out8[4] = p->max_deg; out8[5] = 2 * p->e_total; out8[6] = 0; out8[7] = p->num_sms;
  return SAGNN_OK;
}

extern "C" int sagnn_workspace_bytes(const sagnn_plan* p, int n_layers, int d, size_t* fwd_bytes,
                                     size_t* mask_bytes, size_t* bwd_bytes) {
  if (int rc = check_common(p, n_layers, d, "workspace_bytes")) return rc;
  WsLayout w = ws_layout(p, n_layers, d);
  if (fwd_bytes) *fwd_bytes = w.total;
  if (bwd_bytes) *bwd_bytes = w.total;
  if (mask_bytes) *mask_bytes = sizeof(uint32_t) * mask_layer_words(p, d) * n_layers;
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_fwd(const sagnn_plan* p, const float* uE, const float* iE, float* uOut,
                                   float* iOut, int L, int d, float leaky, void* masks, void* ws,
                                   size_t ws_bytes, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_common(p, L, d, "propagate_fwd")) return rc;
  SAGNN_REQUIRE(uE && iE && uOut && iOut && ws, SAGNN_INVALID_ARG, "propagate_fwd: NULL tensor");
  WsLayout w = ws_layout(p, L, d);
  SAGNN_REQUIRE(ws_bytes >= w.total, SAGNN_WORKSPACE_TOO_SMALL,
                "propagate_fwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  char* base = (char*)ws;
  SpmmParams s;
  base_params(p, s);
  s.tickets = (uint32_t*)(base + w.tickets_off);
  s.partials = (float*)(base + w.partials_off);
  s.leaky = leaky;
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, sizeof(uint32_t) * (size_t)(p->n_long ? p->n_long : 1), st));
  float* buf[2] = {(float*)(base + w.buf_off[0]), (float*)(base + w.buf_off[1])};
  const size_t mlw = mask_layer_words(p, d);
  const size_t mu = (size_t)p->T * p->U * (d / 32);
  for (int l = 0; l < L; ++l) {
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
    const bool write_out = last || l >= 1;
    s.o2_u = write_out ? uOut : nullptr;
    s.o2_i = write_out ? iOut : nullptr;
    s.out_add_next = last ? 1 : 0;
    s.mask_u = masks ? (uint32_t*)masks + (size_t)l * mlw : nullptr;
    s.mask_i = masks ? (uint32_t*)masks + (size_t)l * mlw + mu : nullptr;
    if (int rc = launch(p, s, d, MODE_FWD, st)) return rc;
  }
  return SAGNN_OK;
}

extern "C" int sagnn_propagate_bwd(const sagnn_plan* p, const float* gU, const float* gI, float* dU,
                                   float* dI, int L, int d, float leaky, const void* masks, void* ws,
                                   size_t ws_bytes, sagnn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int rc = check_common(p, L, d, "propagate_bwd")) return rc;
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
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, sizeof(uint32_t) * (size_t)(p->n_long ? p->n_long : 1), st));
  float* buf[2] = {(float*)(base + w.buf_off[0]), (float*)(base + w.buf_off[1])};
  const size_t mlw = mask_layer_words(p, d);
  const size_t mu = (size_t)p->T * p->U * (d / 32);
  for (int l = L - 1, step = 0; l >= 0; --l, ++step) {
    // g = total gradient w.r.t. E^{l+1}; at the top level it is the upstream itself
    const float* g_u = step == 0 ? gU : buf[(step - 1) & 1];
    const float* g_i = step == 0 ? gI : buf[(step - 1) & 1] + w.user_floats;
    s.src_u = g_u; s.src_i = g_i;
    s.smask_u = (const uint32_t*)masks + (size_t)l * mlw;        // sigma'(Z0^l): masks user-table rows
    s.smask_i = (const uint32_t*)masks + (size_t)l * mlw + mu;   // sigma'(Z1^l): masks item-table rows
    s.a_u = gU; s.a_i = gI;
    s.b_u = step == 0 ? nullptr : g_u;
    s.b_i = step == 0 ? nullptr : g_i;
    s.o1_u = l == 0 ? dU : buf[step & 1];
    s.o1_i = l == 0 ? dI : buf[step & 1] + w.user_floats;
    if (int rc = launch(p, s, d, MODE_BWD, st)) return rc;
  }
  return SAGNN_OK;
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
  SAGNN_REQUIRE(ws_bytes >= w.total, SAGNN_WORKSPACE_TOO_SMALL,
                "message_propagate: workspace %zu < %zu bytes", ws_bytes, w.total);
  char* base = (char*)ws;
  SpmmParams s;
  base_params(p, s);
  s.tickets = (uint32_t*)(base + w.tickets_off);
  s.partials = (float*)(base + w.partials_off);
  s.leaky = leaky;
  SAGNN_CUDA(cudaMemsetAsync(s.tickets, 0, sizeof(uint32_t) * (size_t)(p->n_long ? p->n_long : 1), st));
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

extern "C" int sagnn_propagate_host(sagnn_plan* p, const float* uE, const float* iE, const float* gU,
                                    const float* gI, float* uO, float* iO, float* dU, float* dI, int L,
                                    int d, float leaky) {
  if (int rc = check_common(p, L, d, "propagate_host")) return rc;
  SAGNN_REQUIRE(uE && iE && uO && iO, SAGNN_INVALID_ARG, "propagate_host: NULL embedding/output");
  const bool bwd = gU != nullptr;
  SAGNN_REQUIRE(!bwd || (gI && dU && dI), SAGNN_INVALID_ARG, "propagate_host: backward needs gI, dU, dI");
  auto& h = p->hc;
  const size_t nu = sizeof(float) * (size_t)p->T * p->U * d;
  const size_t ni = sizeof(float) * (size_t)p->T * p->I * d;
  if (h.L != L || h.d != d) {
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
  }
  // embeddings in on the compute stream; upstream gradients in on a second copy stream
  // (they are only needed by the backward); outputs leave on a third while backward runs.
  SAGNN_CUDA(cudaMemcpyAsync(h.uE, uE, nu, cudaMemcpyHostToDevice, h.stream));
  SAGNN_CUDA(cudaMemcpyAsync(h.iE, iE, ni, cudaMemcpyHostToDevice, h.stream));
  if (bwd) {
    SAGNN_CUDA(cudaMemcpyAsync(h.gU, gU, nu, cudaMemcpyHostToDevice, h.copy_in));
    SAGNN_CUDA(cudaMemcpyAsync(h.gI, gI, ni, cudaMemcpyHostToDevice, h.copy_in));
    SAGNN_CUDA(cudaEventRecord(h.ev[0], h.copy_in));
  }
  if (int rc = sagnn_propagate_fwd(p, h.uE, h.iE, h.uO, h.iO, L, d, leaky, bwd ? h.masks : nullptr, h.ws,
                                   h.ws_bytes, h.stream))
    return rc;
  SAGNN_CUDA(cudaEventRecord(h.ev[1], h.stream));
  SAGNN_CUDA(cudaStreamWaitEvent(h.copy_out, h.ev[1], 0));
  SAGNN_CUDA(cudaMemcpyAsync(uO, h.uO, nu, cudaMemcpyDeviceToHost, h.copy_out));
  SAGNN_CUDA(cudaMemcpyAsync(iO, h.iO, ni, cudaMemcpyDeviceToHost, h.copy_out));
  if (bwd) {
    SAGNN_CUDA(cudaStreamWaitEvent(h.stream, h.ev[0], 0));
    if (int rc = sagnn_propagate_bwd(p, h.gU, h.gI, h.dU, h.dI, L, d, leaky, h.masks, h.ws, h.ws_bytes,
                                     h.stream))
      return rc;
    SAGNN_CUDA(cudaMemcpyAsync(dU, h.dU, nu, cudaMemcpyDeviceToHost, h.stream));
    SAGNN_CUDA(cudaMemcpyAsync(dI, h.dI, ni, cudaMemcpyDeviceToHost, h.stream));
  }
  SAGNN_CUDA(cudaStreamSynchronize(h.copy_out));
  SAGNN_CUDA(cudaStreamSynchronize(h.stream));
  return SAGNN_OK;
}
