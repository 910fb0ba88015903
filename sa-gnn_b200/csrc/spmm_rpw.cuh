// Row-per-warp layer kernel (v8).  Included by spmm.cu after the shared helpers.
//
// Same contract as spmm_layer_kernel (one launch = one GNN layer over all T intervals and both
// orientations: LIU-YUXI/SA-GNN model.py:118-127; backward = SURVEY A.2), same plan, different
// mapping: ONE WARP OWNS ONE TASK.  The 32 lanes span the latent dimension (d/32 floats each:
// 64-bit loads at d=64, 128-bit at d>=128), so every control decision (trip counts, slice or
// whole row) is warp-uniform: no lock-step lane groups, no per-slot predicates.
//
//   * Nothing a task needs besides the gathered rows is loaded into registers ahead of time:
//     task records, the task's edge codes (and weights) and its own dense operands are brought
//     into a small per-warp ring in shared memory with cp.async, LA tasks ahead, and read back
//     with warp-uniform LDS (one LDS.128 = four edge codes for all lanes, no shuffles).
//   * The gather issues ALL loads of a row before the first add (binary decomposition of the
//     degree: blocks of 16, then 8, 4, 2, 1 slots, every load unpredicated), so a row of up to
//     15 edges costs one L2 round trip and instruction count is proportional to the degree.
//   * No staged hot rows: the 228 KB of L1 that the rings leave free cache the popular source
//     rows by themselves (a build of the half-warp kernel without staging ran within 1 %).
//   * Backward: only the top level applies sign masks per edge; every level writes, next to the
//     gradient it hands down, the copy already multiplied by the next level's sigma', so lower
//     levels gather plain rows.
// Measured ceiling of this access pattern (scripts/micro/ldg_rpw.cu): 77 G rows/s = 19.7 TB/s
// of 256-byte row gathers out of L2 on B200, identical to the half-warp 128-bit form.
#pragma once

#include <type_traits>

namespace sagnn {

#ifndef SAGNN_RPW_THREADS
#define SAGNN_RPW_THREADS 1024
#endif
#ifndef SAGNN_RPW_PREFETCH
#define SAGNN_RPW_PREFETCH 0    // 1: prefetch the next task's source rows into L1 while this task's gathers fly (measured: -7 %)
#endif
#ifndef SAGNN_RPW_LA
#define SAGNN_RPW_LA 2          // tasks of look-ahead of the cp.async request stage
#endif
constexpr int kRpwThreads = SAGNN_RPW_THREADS;
constexpr int kRpwGrab = 8;     // tasks per queue atomic (half of the 16-slot record ring)

template <int VPL, bool MASKED>
struct RowGeo {
  static constexpr int D = 32 * VPL;
  static constexpr int ROWB = D * 4;
  static constexpr int CW = VPL < 4 ? VPL : 4;        // floats per lane per chunk (one load instruction)
  static constexpr int NV = VPL / CW;                 // chunks per lane (2 at d=256)
  static constexpr int CHB = 32 * CW * 4;             // bytes one chunk spans across the warp
  static constexpr int MPR = D / 4;                   // mask bytes per row: byte q = float4 q, bit i = element 4q+i
  static constexpr int RPS = VPL + (MASKED ? 1 : 0);  // registers one in-flight gather slot holds
  // half of the gather super-block: the largest power of two with (2*HB-1) slots in <= 36 registers
#ifdef SAGNN_RPW_HB
  static constexpr int HB = SAGNN_RPW_HB;
#else
  static constexpr int HB = (15 * RPS <= 36) ? 8 : (7 * RPS <= 36) ? 4 : (3 * RPS <= 36) ? 2 : 1;
#endif
  static constexpr int LA = (VPL == 8 || SAGNN_RPW_LA < 2) ? 1 : (VPL == 4 && SAGNN_RPW_LA > 2) ? 2 : SAGNN_RPW_LA;
  static constexpr int RD = LA + 1;                   // ring depth
  // per-warp ring (bytes): 16 task records | RD x 64 codes | RD x 64 weights | RD x (a row, b row)
  static constexpr int REC_OFF = 0;
  static constexpr int CODE_OFF = 256;
  static constexpr int WT_OFF = CODE_OFF + RD * 256;
  static constexpr int OWN_OFF = WT_OFF + RD * 256;
  static constexpr int WARP_BYTES = OWN_OFF + RD * 2 * ROWB;
  static constexpr size_t SMEM = (size_t)(kRpwThreads / 32) * WARP_BYTES;
};

// ---- cp.async helpers --------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async_b(uint32_t dst_smem, const void* src) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
  else if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}

// same, issued only when `pred` is non-zero (keeps a run of copies free of branches)
template <int BYTES>
__device__ __forceinline__ void cp_async_p(uint32_t dst_smem, const void* src, int pred) {
  if constexpr (BYTES == 16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(dst_smem), "l"(src), "r"(pred) : "memory");
  else if constexpr (BYTES == 8)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}"
                 ::"r"(dst_smem), "l"(src), "r"(pred) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                 ::"r"(dst_smem), "l"(src), "r"(pred) : "memory");
}

// ---- per-lane chunk loads / stores (CW floats = 4, 8 or 16 bytes) ---------------------------
template <int CW>
__device__ __forceinline__ void ldg_chunk(float* v, const char* p) {            // read-only path
  if constexpr (CW == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (CW == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
  else { v[0] = __ldg(reinterpret_cast<const float*>(p)); }
}
template <int CW>
__device__ __forceinline__ void ld_strong_chunk(float* v, const char* p) {      // L1-bypassing, GPU scope
  if constexpr (CW == 4)
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p) : "memory");
  else if constexpr (CW == 2)
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p) : "memory");
  else
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v[0]) : "l"(p) : "memory");
}
template <int CW>
__device__ __forceinline__ void lds_chunk(float* v, uint32_t addr) {
  if constexpr (CW == 4)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
  else if constexpr (CW == 2)
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr) : "memory");
  else
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr) : "memory");
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int2 lds_i2(uint32_t addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int lds_i1(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_i4(uint32_t addr, int4 v) {
  asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int CW>
__device__ __forceinline__ void st_chunk(char* p, const float* v) {
  if constexpr (CW == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (CW == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else *reinterpret_cast<float*>(p) = v[0];
}
template <int CW>
__device__ __forceinline__ void stcs_chunk(char* p, const float* v) {
  if constexpr (CW == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  else if constexpr (CW == 2) __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  else __stcs(reinterpret_cast<float*>(p), v[0]);
}

// acc += x, optionally scaled by an edge weight and / or the sign mask of the source row
// (bit i of `bits` belongs to x[i]: pass where set, leaky where clear).  Packed FADD2 / FFMA2.
template <int VPL, bool WEIGHTED, bool MASKED>
__device__ __forceinline__ void rpw_accumulate(float* acc, const float* x, float w, uint32_t bits, float leaky) {
  if constexpr (VPL == 1) {
    if (!WEIGHTED && !MASKED) acc[0] += x[0];
    else {
      float s = WEIGHTED ? w : 1.f;
      if (MASKED) s = (bits & 1u) ? s : s * leaky;
      acc[0] = fmaf(x[0], s, acc[0]);
    }
  } else {
    const float s = WEIGHTED ? w : 1.f;
    const float sl = s * leaky;
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
      float2 a = make_float2(acc[i], acc[i + 1]);
      if (!WEIGHTED && !MASKED) {
        a = __fadd2_rn(a, make_float2(x[i], x[i + 1]));
      } else {
        const float s0 = MASKED ? (((bits >> i) & 1u) ? s : sl) : s;
        const float s1 = MASKED ? (((bits >> (i + 1)) & 1u) ? s : sl) : s;
        a = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(s0, s1), a);
      }
      acc[i] = a.x; acc[i + 1] = a.y;
    }
  }
}

// MODE_FWD / MODE_MSG: MASKED must be false.  MODE_BWD: MASKED = the source is the raw upstream
// (top level, sign masks applied per edge); !MASKED = the source was pre-multiplied by sigma'.
// RTD: some tensor of this launch uses the [R,T,d] layout (runtime row strides); otherwise every
// row stride is the compile-time d*4.
template <int VPL, int MODE, bool WEIGHTED, bool MASKED, bool RTD>
__global__ void __launch_bounds__(kRpwThreads, 1)
spmm_rpw_kernel(const __grid_constant__ SpmmParams p) {
  using G = RowGeo<VPL, MASKED>;
  constexpr int D = G::D, ROWB = G::ROWB, CW = G::CW, NV = G::NV, CHB = G::CHB, MPR = G::MPR, HB = G::HB;
  constexpr int LA = G::LA, RD = G::RD;
  constexpr bool BWD = MODE == MODE_BWD;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NOWORK = 0x40000000;                  // meta bit 30
  static_assert(!MASKED || BWD, "sign masks belong to the backward");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float leaky = p.leaky;
  // everything below is derived from a shuffled value, so the compiler keeps it in uniform registers
  const int seg = __shfl_sync(FULL, p.single_seg >= 0 ? p.single_seg : p.cta[blockIdx.x].seg, 0);
  // the first grab of every warp is static (CTA rank inside its segment), the queue serves the rest
  const unsigned seg_rank = p.single_seg >= 0 ? blockIdx.x : (unsigned)p.cta[blockIdx.x].rank;
  const unsigned seg_ctas = p.single_seg >= 0 ? gridDim.x : (unsigned)p.cta[blockIdx.x].count;
  if (p.trace && threadIdx.x == 0) {
    p.trace[blockIdx.x * 4 + 0] = (unsigned long long)seg;
    p.trace[blockIdx.x * 4 + 1] = globaltimer_ns();
    p.trace[blockIdx.x * 4 + 2] = p.trace[blockIdx.x * 4 + 1];
  }

  const sagnn_seg sg = p.seg[seg];
  const unsigned n_seg_tasks = (unsigned)(sg.task_end - sg.task_begin);
  unsigned* ctr = p.ctrs + seg;
  const int k = seg >> 1;
  const bool item_side = seg & 1;
  const int r_own = item_side ? p.I : p.U, r_src = item_side ? p.U : p.I;
  const int64_t own0 = (int64_t)k * r_own;
  // [T,R,d]: interval k starts at row k*R, rows are D floats apart;
  // [R,T,d]: interval k starts at float k*D of row 0, rows are T*D floats apart (model.py:133-134)
  const bool src_rtd = RTD && p.src_rtd, a_rtd = RTD && p.a_rtd, b_rtd = RTD && p.b_rtd, o2_rtd = RTD && p.o2_rtd;
  const char* src = reinterpret_cast<const char*>((item_side ? p.src_u : p.src_i) +
                                                  (src_rtd ? (int64_t)k * D : (int64_t)k * r_src * D));
  const uint32_t src_stride = (uint32_t)ROWB * (src_rtd ? p.T : 1);   // bytes between source rows
  const uint8_t* smask = MASKED ? (item_side ? p.smask_u : p.smask_i) + (int64_t)k * r_src * MPR : nullptr;
  const int32_t* enc = p.enc + sg.edge_base;
  const float* wts = WEIGHTED ? p.w + sg.edge_base : nullptr;
  const sagnn_task* tasks = p.tasks + sg.task_begin;
  const float* a_f = item_side ? p.a_i : p.a_u;
  const float* b_f = item_side ? p.b_i : p.b_u;
  float* o1_f = item_side ? p.o1_i : p.o1_u;
  float* o2_f = item_side ? p.o2_i : p.o2_u;
  uint8_t* mk_f = item_side ? p.mask_i : p.mask_u;
  const uint8_t* pm_f = item_side ? p.pmask_i : p.pmask_u;
  const char* a_base = a_f ? reinterpret_cast<const char*>(a_f + (a_rtd ? (int64_t)k * D : own0 * D)) : nullptr;
  const char* b_base = b_f ? reinterpret_cast<const char*>(b_f + (b_rtd ? (int64_t)k * D : own0 * D)) : nullptr;
  char* o1_base = o1_f ? reinterpret_cast<char*>(o1_f + own0 * D) : nullptr;
  char* o2_base = o2_f ? reinterpret_cast<char*>(o2_f + (o2_rtd ? (int64_t)k * D : own0 * D)) : nullptr;
  uint8_t* mk_base = mk_f ? mk_f + own0 * MPR : nullptr;
  const uint8_t* pm_base = pm_f ? pm_f + own0 * MPR : nullptr;
  const uint32_t a_stride = (uint32_t)ROWB * (a_rtd ? p.T : 1);
  const uint32_t b_stride = (uint32_t)ROWB * (b_rtd ? p.T : 1);
  const uint32_t o2_stride = (uint32_t)ROWB * (o2_rtd ? p.T : 1);
  // which optional tensors this launch has: one pinned register instead of pointer tests per task
  enum { F_B = 1, F_O1 = 2, F_O2 = 4, F_MK = 8, F_ADDNEXT = 16, F_PM = 32, F_PEER = 64 };
  uint32_t flags = (b_base ? F_B : 0) | (o1_base ? F_O1 : 0) | (o2_base ? F_O2 : 0) | (mk_base ? F_MK : 0) |
                   (p.out_add_next ? F_ADDNEXT : 0) | (pm_base ? F_PM : 0) |
                   ((RTD && MODE == MODE_FWD && p.peer_n > 0) ? F_PEER : 0);
  pin32(flags);

  // my bytes inside a chunk; my sign bits inside a row's mask bytes: byte (chunk v) = v*32 + mbyte
  uint32_t lane_off = (uint32_t)lane * (CW * 4);
  pin32(lane_off);
  const int mbyte = CW == 4 ? lane : (CW == 2 ? lane >> 1 : lane >> 2);
  const int mshift = CW == 4 ? 0 : (CW == 2 ? (lane & 1) * 2 : (lane & 3));
  auto mask_bits = [&](const uint8_t* base, uint32_t c) -> uint32_t {   // sign bits of row c that belong to my elements
    uint32_t bits = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      bits |= (((uint32_t)__ldg(base + (uint64_t)c * MPR + v * 32 + mbyte)) >> mshift) << (v * 4);
    return bits;
  };

  // ---- per-warp ring -----------------------------------------------------------------------
  // (pinned: the compiler must not rebuild these from %tid / the kernel parameters at every use)
  uint32_t rec_ring = smem_u32(smem_raw) + (uint32_t)warp * G::WARP_BYTES + G::REC_OFF;
  uint32_t code_lane = rec_ring + (G::CODE_OFF - G::REC_OFF) + lane * 4;      // my word of code-ring slot 0
  uint32_t own_lane = rec_ring + (G::OWN_OFF - G::REC_OFF) + lane_off;        // my bytes of own-ring slot 0
  pin32(rec_ring);
  pin32(code_lane);
  pin32(own_lane);
  const uint32_t code_ring = rec_ring + (G::CODE_OFF - G::REC_OFF), wt_ring = rec_ring + (G::WT_OFF - G::REC_OFF);
  const char* src_lane = src + lane_off;
  uint32_t src_stride_r = src_stride;
  pin64(src_lane);
  pin32(src_stride_r);

  const unsigned q_base = seg_ctas * (unsigned)(kRpwThreads / 32) * kRpwGrab;   // tasks handed out statically
  auto issue = [&]() -> unsigned { return lane == 0 ? atomicAdd(ctr, (unsigned)kRpwGrab) + q_base : 0u; };
  // records of the grab starting at task b -> half `h` of the record ring (16 bytes each)
  auto fetch_records = [&](unsigned b, int h) {
    if (lane < kRpwGrab) {
      const unsigned t = b + lane;
      const uint32_t dst = rec_ring + (uint32_t)(h * kRpwGrab + lane) * 16;
      if (t < n_seg_tasks) cp_async_b<16>(dst, tasks + t);
      else sts_i4(dst, make_int4(0, NOWORK, 0, 0));
    }
  };
  // everything task q needs besides the gathered rows -> ring slot `slot`
  auto request = [&](int q, int slot) {
    const int4 rq = lds_i4(rec_ring + (uint32_t)(q & 15) * 16);
    if (!(rq.y & NOWORK)) {
      const int nn = rq.y & 0x7f;
      const uint32_t ln = lane_off / (CW * 4);
      const uint32_t cdst = code_lane + (uint32_t)slot * 256;
      const int32_t* csrc = enc + ((uint32_t)rq.z + ln);
      cp_async_p<4>(cdst, csrc, ln < (uint32_t)nn);
      cp_async_p<4>(cdst + 128, csrc + 32, ln + 32 < (uint32_t)nn);
      if (WEIGHTED) {
        const uint32_t wdst = cdst + (G::WT_OFF - G::CODE_OFF);
        const float* wsrc = wts + ((uint32_t)rq.z + ln);
        cp_async_p<4>(wdst, wsrc, ln < (uint32_t)nn);
        cp_async_p<4>(wdst + 128, wsrc + 32, ln + 32 < (uint32_t)nn);
      }
      if (MODE != MODE_MSG) {
        const uint32_t odst = own_lane + (uint32_t)slot * (2 * ROWB);
        const uint64_t ra = (uint64_t)(uint32_t)rq.x * a_stride + lane_off;
        const uint64_t rb = RTD ? (uint64_t)(uint32_t)rq.x * b_stride + lane_off : ra;
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async_b<CW * 4>(odst + v * CHB, a_base + ra + v * CHB);
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async_p<CW * 4>(odst + ROWB + v * CHB, b_base + rb + v * CHB, flags & F_B);
      }
    }
  };

  // ---- start-up: first grab's records, then the first LA requests --------------------------
  const unsigned b_first = (seg_rank * (unsigned)(kRpwThreads / 32) + (unsigned)warp) * kRpwGrab;
  unsigned pend = issue();                            // base of the next grab, still in flight
  fetch_records(b_first, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
#pragma unroll
  for (int q = 0; q < LA; ++q) {
    request(q, q);
    cp_async_commit();
  }

  int slot_t = 0, slot_q = LA % RD;                   // ring slots of the current task / of the request
  for (int t = 0;; ++t) {
    cp_async_wait<LA - 1>();                          // task t's operands (requested LA iterations ago) have landed
    __syncwarp();                                     // ... for every lane; also fences ring-slot reuse
    const int4 rec = lds_i4(rec_ring + (uint32_t)(t & 15) * 16);   // {row, meta, e_off, aux}
    if (rec.y & NOWORK) break;                        // grabs only grow: the first empty record ends the stream
    if ((t & 7) == 0) {                               // the other half of the record ring is free: fetch the next grab
      const unsigned b_next = __shfl_sync(FULL, pend, 0);
      pend = issue();
      fetch_records(b_next, ((t >> 3) + 1) & 1);
    }
    request(t + LA, slot_q);
    cp_async_commit();

    const uint32_t row = (uint32_t)rec.x;
    const int n = rec.y & 0x7f;
    const bool multi = rec.y < 0;                     // bit 31: slice of a long row
    const uint32_t cb = code_ring + (uint32_t)slot_t * 256;
    const uint32_t wb = wt_ring + (uint32_t)slot_t * 256;
    uint32_t pbits = 0;                               // backward: my sign bits one level down (for the pre-masked copy)
    if (BWD && (flags & F_PM) && !multi) pbits = mask_bits(pm_base, row);

    float acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = 0.f;

    // ---- gather-reduce ---------------------------------------------------------------------
    float val[2 * HB][VPL];
    uint32_t mb[MASKED ? 2 * HB : 1];
    // K loads (K = 1, 2, 4, 8, 16) of the codes at position j into registers s0.., all unpredicated
    auto gather = [&](auto kc, int j, int s0) {
      constexpr int K = decltype(kc)::value;
      int c[K];
      if constexpr (K >= 4) {
#pragma unroll
        for (int g = 0; g < K / 4; ++g) {
          const int4 c4 = lds_i4(cb + (uint32_t)(j + 4 * g) * 4);
          c[4 * g] = c4.x; c[4 * g + 1] = c4.y; c[4 * g + 2] = c4.z; c[4 * g + 3] = c4.w;
        }
      } else if constexpr (K == 2) {
        const int2 c2 = lds_i2(cb + (uint32_t)j * 4);
        c[0] = c2.x; c[1] = c2.y;
      } else {
        c[0] = lds_i1(cb + (uint32_t)j * 4);
      }
#pragma unroll
      for (int u = 0; u < K; ++u) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          ldg_chunk<CW>(val[s0 + u] + v * CW, src_lane + (uint64_t)(uint32_t)c[u] * src_stride_r + v * CHB);
        if constexpr (MASKED) mb[s0 + u] = mask_bits(smask, (uint32_t)c[u]);
      }
    };
    auto reduce = [&](auto kc, int j, int s0) {
      constexpr int K = decltype(kc)::value;
#pragma unroll
      for (int u = 0; u < K; ++u) {
        float w = 1.f;
        if constexpr (WEIGHTED) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(wb + (uint32_t)(j + u) * 4) : "memory");
        rpw_accumulate<VPL, WEIGHTED, MASKED>(acc, val[s0 + u], w, MASKED ? mb[MASKED ? s0 + u : 0] : 0u, leaky);
      }
    };
    {
      int j = 0;
      for (; j + 2 * HB <= n; j += 2 * HB) {           // long tasks: 2*HB rows in flight per lane
        gather(std::integral_constant<int, 2 * HB>(), j, 0);
        reduce(std::integral_constant<int, 2 * HB>(), j, 0);
      }
      const int rem = n - j;                           // < 2*HB: one block per set bit, all issued before the first add
      constexpr int S8 = 0, S4 = HB >= 8 ? 8 : 0, S2 = S4 + (HB >= 4 ? 4 : 0), S1 = S2 + (HB >= 2 ? 2 : 0);
      int jj = j;
      if constexpr (HB >= 8) { if (rem & 8) { gather(std::integral_constant<int, 8>(), jj, S8); jj += 8; } }
      if constexpr (HB >= 4) { if (rem & 4) { gather(std::integral_constant<int, 4>(), jj, S4); jj += 4; } }
      if constexpr (HB >= 2) { if (rem & 2) { gather(std::integral_constant<int, 2>(), jj, S2); jj += 2; } }
      if (rem & 1) gather(std::integral_constant<int, 1>(), jj, S1);
#if SAGNN_RPW_PREFETCH
      if constexpr (LA >= 2) {
        // The next task's codes were requested a full iteration ago: make sure they have landed, then
        // every lane prefetches ONE of its source rows into L1 (two instructions per 128-byte line pair
        // for up to 32 rows), so that task's gathers find their rows next to the SM.
        cp_async_wait<LA - 1>();
        __syncwarp();
        const int4 rn = lds_i4(rec_ring + (uint32_t)((t + 1) & 15) * 16);
        const uint32_t ln = lane_off / (CW * 4);
        if (!(rn.y & NOWORK) && ln < (uint32_t)(rn.y & 0x7f)) {
          const uint32_t sn = slot_t + 1 == RD ? 0 : slot_t + 1;
          const uint32_t c = (uint32_t)lds_i1(code_lane + sn * 256);
          const char* a = src + (uint64_t)c * src_stride_r;
#pragma unroll
          for (int b = 0; b < ROWB; b += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + b));
        }
      }
#endif
      jj = j;
      if constexpr (HB >= 8) { if (rem & 8) { reduce(std::integral_constant<int, 8>(), jj, S8); jj += 8; } }
      if constexpr (HB >= 4) { if (rem & 4) { reduce(std::integral_constant<int, 4>(), jj, S4); jj += 4; } }
      if constexpr (HB >= 2) { if (rem & 2) { reduce(std::integral_constant<int, 2>(), jj, S2); jj += 2; } }
      if (rem & 1) reduce(std::integral_constant<int, 1>(), jj, S1);
    }

    // ---- long rows: publish the slice sum; reduce through a fan-in-16 ticket tree -------------
    // The last arriver of every group of 16 slices (then of 16 groups, ...) sums them in slice
    // order: deterministic, no float atomics.  Release increments on every ticket, one acquire fence in
    // the last arriver of a group before it reads the partials (L1-bypassing loads).
    bool finish = true;
    if (multi) {
      const uint32_t aux = (uint32_t)rec.w;            // global slice id
      const uint32_t lr = __ldg(p.chunk_lr + aux);
      const int64_t cbase = __ldg(p.chunk_base + lr);
      const int nch = (int)(__ldg(p.chunk_base + lr + 1) - cbase);
      int pos = (int)((int64_t)aux - cbase);           // my slice inside the row
      bool active = true;
      finish = false;
      unsigned* tk = p.tickets;                        // ticket region of the current level
      for (int sh = 0; active; sh += 4) {              // level stride = 16^level = 1 << sh
        const int stride = 1 << sh;
        const int gs = pos & ~((16 << sh) - 1);        // members: slots gs + j*stride, j < 16, below nch
        int members = (nch - gs + stride - 1) >> sh;
        members = members > 16 ? 16 : members;
        char* mine = reinterpret_cast<char*>(p.partials + (cbase + pos) * D) + lane_off;
#pragma unroll
        for (int v = 0; v < NV; ++v) st_chunk<CW>(mine + v * CHB, acc + v * CW);
        __syncwarp();
        unsigned* my_tk = tk + (cbase + gs);           // one ticket per group, named by its first slot
        unsigned old = 0;
        if (lane == 0) old = ticket_release_add(my_tk);
        old = __shfl_sync(FULL, old, 0);
        if (old != (unsigned)(members - 1)) {
          active = false;                              // someone else finishes this group
        } else {
          ticket_acquire_fence();                      // pairs with the other slices' release increments
          if (lane == 0) *my_tk = 0u;                  // ready for the next launch
#pragma unroll
          for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
          const char* part = reinterpret_cast<const char*>(p.partials + (cbase + gs) * D) + lane_off;
          for (int c0 = 0; c0 < members; c0 += 4) {
            float pv[4][VPL];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                if (c0 + u < members) {
                  ld_strong_chunk<CW>(pv[u] + v * CW, part + (int64_t)(c0 + u) * stride * ROWB + v * CHB);
                } else {
#pragma unroll
                  for (int i = 0; i < CW; ++i) pv[u][v * CW + i] = 0.f;
                }
              }
#pragma unroll
            for (int u = 0; u < 4; ++u) rpw_accumulate<VPL, false, false>(acc, pv[u], 1.f, 0u, leaky);
          }
          if (gs == 0 && 16 * stride >= nch) {         // that was the whole row
            finish = true;
            active = false;
          } else {
            pos = gs;                                  // my sum becomes slot gs of the next level
          }
        }
        tk += p.n_chunks;
      }
      __syncwarp();
      if (BWD && (flags & F_PM) && finish) pbits = mask_bits(pm_base, row);
    }

    // ---- fused epilogue ------------------------------------------------------------------------
    if (finish) {
      float own_a[VPL], own_b[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) { own_a[i] = 0.f; own_b[i] = 0.f; }
      if (MODE != MODE_MSG) {
        const uint32_t o = own_lane + (uint32_t)slot_t * (2 * ROWB);
#pragma unroll
        for (int v = 0; v < NV; ++v) lds_chunk<CW>(own_a + v * CW, o + v * CHB);
        if (flags & F_B) {
#pragma unroll
          for (int v = 0; v < NV; ++v) lds_chunk<CW>(own_b + v * CW, o + ROWB + v * CHB);
        }
      }
      const uint64_t off = (uint64_t)row * ROWB + lane_off;       // bytes inside a contiguous [rows, d] table
      if (BWD) {
        // n = G + g + A (sigma' . g_other)      (SURVEY A.2); at the top level g == G
        float o[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) o[i] = own_a[i] + ((flags & F_B) ? own_b[i] : own_a[i]) + acc[i];
#pragma unroll
        for (int v = 0; v < NV; ++v) st_chunk<CW>(o1_base + off + v * CHB, o + v * CW);
        if (flags & F_O2) {                                        // the source of the next level down: sigma'(Z^{l-1}) (.) n
          float om[VPL];
#pragma unroll
          for (int i = 0; i < VPL; ++i) om[i] = ((pbits >> i) & 1u) ? o[i] : leaky * o[i];
#pragma unroll
          for (int v = 0; v < NV; ++v) st_chunk<CW>(o2_base + off + v * CHB, om + v * CW);
        }
      } else {
        // LeakyReLU = max(leaky*z, z)  (Utils/NNLayers.py:135-136)
        float act[VPL];
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const float lz = leaky * acc[i];
          act[i] = fmaxf(lz, acc[i]);
          // TF MaximumGrad sends the gradient to leaky*z where leaky*z >= z: bit = pass-through
          bits |= (!(lz >= acc[i]) ? 1u : 0u) << i;
        }
        if (MODE == MODE_MSG) {
#pragma unroll
          for (int v = 0; v < NV; ++v) st_chunk<CW>(o1_base + off + v * CHB, act + v * CW);
        } else {
          float nxt_e[VPL];                                         // E^{l+1} = E^l + lrelu(Z^l)
#pragma unroll
          for (int i = 0; i < VPL; ++i) nxt_e[i] = own_a[i] + act[i];
          if (flags & F_O1) {
#pragma unroll
            for (int v = 0; v < NV; ++v) st_chunk<CW>(o1_base + off + v * CHB, nxt_e + v * CW);
          }
          if (flags & F_O2) {
            float o[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
              o[i] = (flags & F_B) ? own_b[i] + own_a[i] : own_a[i];
              if (flags & F_ADDNEXT) o[i] += nxt_e[i];
            }
            char* dst = o2_base + (uint64_t)row * o2_stride + lane_off;
            if constexpr (RTD && MODE == MODE_FWD) {
              if (flags & F_PEER) {
                // fused hand-off: row r belongs to the consumer rank r / blk; its receive buffer is
                // [source rank, blk, T, d], so this row lands at [my rank][r % blk][k] -- a peer-memory
                // store over NVLink (or a local one for my own block)
                const uint32_t blk = item_side ? (uint32_t)p.peer_blk_i : (uint32_t)p.peer_blk_u;
                const uint32_t pr = row / blk, lr = row - pr * blk;
                float* pb = item_side ? p.peer_i[pr] : p.peer_u[pr];
                dst = reinterpret_cast<char*>(pb) +
                      (((uint64_t)p.peer_rank * blk + lr) * (uint32_t)p.T + (uint32_t)k) * ROWB + lane_off;
              }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) stcs_chunk<CW>(dst + v * CHB, o + v * CW);
          }
          if (flags & F_MK) {
            uint8_t* mrow = mk_base + (uint64_t)row * MPR;
            if constexpr (CW == 4) {
#pragma unroll
              for (int v = 0; v < NV; ++v) mrow[v * 32 + lane] = (uint8_t)((bits >> (v * 4)) & 0xfu);
            } else if constexpr (CW == 2) {
              const uint32_t hi = __shfl_down_sync(FULL, bits, 1);
              if (!(lane & 1)) mrow[lane >> 1] = (uint8_t)(bits | (hi << 2));
            } else {
              const uint32_t b1_ = __shfl_down_sync(FULL, bits, 1), b2_ = __shfl_down_sync(FULL, bits, 2),
                             b3_ = __shfl_down_sync(FULL, bits, 3);
              if (!(lane & 3)) mrow[lane >> 2] = (uint8_t)(bits | (b1_ << 1) | (b2_ << 2) | (b3_ << 3));
            }
          }
        }
      }
    }
    slot_t = slot_t + 1 == RD ? 0 : slot_t + 1;
    slot_q = slot_q + 1 == RD ? 0 : slot_q + 1;
  }
  cp_async_wait<0>();
  if (p.trace) {
    __syncthreads();
    if (threadIdx.x == 0) p.trace[blockIdx.x * 4 + 3] = globaltimer_ns();
  }
}

}  // namespace sagnn
