// Packet-stream layer kernel (v9, the default).  Included by spmm.cu after spmm_rpw.cuh (shares its
// per-lane load / store / accumulate helpers).
//
// Same contract as the v8 kernel (one launch = one GNN layer over all T intervals and both
// orientations: LIU-YUXI/SA-GNN model.py:118-127; backward = SURVEY A.2), same row-per-warp mapping
// (32 lanes span the latent dimension, every control decision warp-uniform), but built around what
// the round-2 microbenchmark (scripts/micro/l1tex_cost.cu, profiles/r2_micro_l1tex_cost.txt) says
// bounds a gather kernel on B200: the SM's L1TEX / LSU pipe.  A 256-byte row costs that pipe
// 3.8 cycles as an LDG from L2 and 8+ as a cp.async, so everything that is not a gathered row must
// stay out of it:
//
//   * The schedule is a PACKED TASK STREAM built by the plan: packets of 4 tasks = 4 records
//     {row, n | flags, slice id, code offset} followed by the tasks' edge codes (and weights), in
//     schedule order.  A warp brings its next packets into shared memory with ONE TMA bulk copy
//     each (cp.async.bulk + mbarrier, issued by one lane, SLOTS-1 packets ahead): no per-task
//     record / code requests, no LSU work besides the warp-uniform LDS.128 that read them back.
//   * Packets are dealt statically (packet q of a segment belongs to warp q mod (32 x CTAs of the
//     segment)); tasks are sorted by descending degree, so the interleave balances itself and the
//     queue atomics of v8 are gone.
//   * TWO TASKS IN FLIGHT PER WARP (512 threads x 128 registers): the gathers and own-row loads of
//     task t+1 are issued before task t is reduced, so a warp always has a task's worth of rows in
//     flight while it adds, runs the epilogue and stores (ncu on the one-task form: issue slots 57 %
//     busy, the rest long-scoreboard stalls on the gathers; SAGNN_PKT_BANKS=1 builds that form).
//   * The task's own dense rows (E^l row, layer-sum row / G and g rows) are plain 64/128-bit loads
//     into registers, issued together with the task's gathers.
//   * Gather: as v8 -- all loads of a row issued before the first add (blocks of 16, 8, 4, 2, 1
//     unpredicated ld.global.nc), packed FADD2 / FFMA2 accumulation.
//   * Long rows (deg > 64): <= 64-edge slices, fan-in-16 ticket tree, deterministic (as v8).
#pragma once

namespace sagnn {

#ifndef SAGNN_PKT_BANKS
#define SAGNN_PKT_BANKS 1       // tasks in flight per warp: 1 = one at a time (1024 threads, default); 2 = software-pipelined (512 threads x 128 registers: measured 0.78 vs 0.50 ms, half the warps cost more than the overlap gains)
#endif
#ifndef SAGNN_PKT_THREADS
#define SAGNN_PKT_THREADS (SAGNN_PKT_BANKS == 2 ? 512 : 1024)
#endif
#ifndef SAGNN_PKT_SLOTS
#define SAGNN_PKT_SLOTS 3       // packets resident / in flight per warp
#endif
constexpr int kPktThreads = SAGNN_PKT_THREADS;
constexpr int kPktWarps = kPktThreads / 32;
constexpr uint32_t kPktNoWork = 0x40000000u;   // record meta bit 30: padding record (end of the segment)

template <int VPL, bool WEIGHTED>
struct PktGeo {
  static constexpr int D = 32 * VPL;
  static constexpr int ROWB = D * 4;
  static constexpr int CW = VPL < 4 ? VPL : 4;        // floats per lane per chunk (one load instruction)
  static constexpr int NV = VPL / CW;                 // chunks per lane (2 at d=256)
  static constexpr int CHB = 32 * CW * 4;             // bytes one chunk spans across the warp
  static constexpr int MPR = D / 4;                   // mask bytes per row
  // half of the gather super-block (2*HB-1 rows in flight per lane below a full block, 2*HB in one)
#ifdef SAGNN_PKT_HB
  static constexpr int HB = SAGNN_PKT_HB;
#else
  static constexpr int HB = (15 * VPL <= 36) ? 8 : (7 * VPL <= 36) ? 4 : (3 * VPL <= 36) ? 2 : 1;
#endif
  static constexpr int NS = WEIGHTED ? 2 : SAGNN_PKT_SLOTS;    // packets resident / in flight per warp
  static constexpr int SLOT_BYTES = kPktTasks * 16 + kPktTasks * kChunk * 4 * (WEIGHTED ? 2 : 1);
  static constexpr size_t SMEM = (size_t)kPktWarps * NS * SLOT_BYTES;
};

__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "PW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra PD_%=;\n\t"
      "bra PW_%=;\n\t"
      "PD_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_u32(uint32_t dst_smem, const void* src_gmem, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}

template <int VPL, int MODE, bool WEIGHTED, bool RTD>
__global__ void __launch_bounds__(kPktThreads, 1)
spmm_pkt_kernel(const __grid_constant__ SpmmParams p) {
  using G = PktGeo<VPL, WEIGHTED>;
  constexpr int D = G::D, ROWB = G::ROWB, CW = G::CW, NV = G::NV, CHB = G::CHB, MPR = G::MPR, HB = G::HB;
  constexpr int NS = G::NS;
  constexpr bool BWD = MODE == MODE_BWD;
  constexpr bool OWN = MODE != MODE_MSG;              // the task has dense rows of its own
  constexpr int BANKS = SAGNN_PKT_BANKS;
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[kPktWarps * NS];
  const int lane = threadIdx.x & 31;
  // shuffled: tells the compiler the value is warp-uniform (uniform registers for everything derived from it)
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
  const float leaky = p.leaky;
  // the CTA's work descriptor comes by value through the kernel parameters: warp-uniform by construction
  const PktCta cd = p.ctad[blockIdx.x];
  const int seg = (int)cd.seg;
  if (p.trace && threadIdx.x == 0) {
    p.trace[blockIdx.x * 4 + 0] = (unsigned long long)seg;
    p.trace[blockIdx.x * 4 + 1] = globaltimer_ns();
    p.trace[blockIdx.x * 4 + 2] = p.trace[blockIdx.x * 4 + 1];
  }

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

  const char* src_lane = src + lane_off;
  uint32_t src_stride_r = src_stride;
  pin64(src_lane);

  // ---- my packets: q0, q0 + stride, ... of the segment's packet list -------------------------
  const unsigned n_pk = cd.n_pk;
  const uint32_t* dir = p.pkt_dir + cd.pkt_begin;     // packet offsets, 16-byte units into the stream
  const unsigned q0 = cd.q0 + (unsigned)warp;
  const unsigned stride = cd.stride;
  const unsigned n_my = q0 < n_pk ? (n_pk - q0 + stride - 1) / stride : 0u;

  uint32_t slot0 = smem_u32(smem_raw) + (uint32_t)warp * (NS * G::SLOT_BYTES);
  uint32_t bar0 = smem_u32(&bars[warp * NS]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto load_dir = [&](unsigned j) -> uint2 {          // {first, end} 16-byte unit of my j-th packet
    const uint32_t* d = dir + (q0 + j * stride);
    return make_uint2(__ldg(d), __ldg(d + 1));
  };
  auto issue_pkt = [&](unsigned slot, uint2 dv) {      // one lane, one bulk copy: records + codes (+ weights)
    if (lane == 0) {
      const unsigned bytes = (dv.y - dv.x) * 16u;
      mbar_expect_tx_u32(bar0 + 8 * slot, bytes);
      tma_bulk_g2s_u32(slot0 + slot * G::SLOT_BYTES, reinterpret_cast<const char*>(p.pkt_stream) + (uint64_t)dv.x * 16u,
                       bytes, bar0 + 8 * slot);
    }
  };

  if (n_my > 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if ((unsigned)s < n_my) issue_pkt(s, load_dir(s));
  }
  uint2 dir_next = make_uint2(0u, 0u);
  if (n_my > (unsigned)NS) dir_next = load_dir(NS);

  // own rows of a task (row r): a = E^l / G row, b = layer-sum / running-gradient row, pm = sign bits one level down
  struct Own { float a[VPL]; float b[VPL]; uint32_t pbits; };
  auto load_own = [&](Own& o, uint32_t row) {
    if (!OWN) return;
    const char* pa = a_base + (uint64_t)row * a_stride + lane_off;
#pragma unroll
    for (int v = 0; v < NV; ++v) ldg_chunk<CW>(o.a + v * CW, pa + v * CHB);
    if (flags & F_B) {
      const char* pb = b_base + (uint64_t)row * (RTD ? b_stride : a_stride) + lane_off;
#pragma unroll
      for (int v = 0; v < NV; ++v) ldg_chunk<CW>(o.b + v * CW, pb + v * CHB);
    }
    if (BWD && (flags & F_PM)) o.pbits = mask_bits(pm_base, row);
  };

  // ---- task stream ----------------------------------------------------------------------------
  // A bank = the registers of one task in flight: gathered rows, own rows, record fields.  With two
  // banks (the default) the gathers of task t+1 are issued before task t is reduced, so a warp always
  // has a task's worth of rows in flight while it adds, runs the epilogue and stores.
  struct Bank {
    float val[2 * HB][VPL];
    float acc[VPL];
    Own own;
    uint32_t row, aux, cb;     // row id, slice id, shared-memory address of the task's edge codes
    int meta;                  // n | long-row-slice flag (bit 31)
    int jl;                    // first edge of the blocks still in flight
    unsigned refill;           // ring slot this task frees when it is done (NS: none)
  };
  unsigned j = 0;                                      // packet the cursor is in
  int t = 0;                                           // next record inside it
  unsigned slot = 0, parity = 0;                       // ring slot of packet j and its mbarrier phase
  unsigned j_issue = NS;                               // next packet to bring in

  auto gather = [&](Bank& b, auto kc, int jj, int s0) {   // K loads of the codes at position jj, all unpredicated
    constexpr int K = decltype(kc)::value;
    int c[K];
    if constexpr (K >= 4) {
#pragma unroll
      for (int g = 0; g < K / 4; ++g) {
        const int4 c4 = lds_i4(b.cb + (uint32_t)(jj + 4 * g) * 4);
        c[4 * g] = c4.x; c[4 * g + 1] = c4.y; c[4 * g + 2] = c4.z; c[4 * g + 3] = c4.w;
      }
    } else if constexpr (K == 2) {
      const int2 c2 = lds_i2(b.cb + (uint32_t)jj * 4);
      c[0] = c2.x; c[1] = c2.y;
    } else {
      c[0] = lds_i1(b.cb + (uint32_t)jj * 4);
    }
#pragma unroll
    for (int u = 0; u < K; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
        ldg_chunk<CW>(b.val[s0 + u] + v * CW, src_lane + (uint64_t)(uint32_t)c[u] * src_stride_r + v * CHB);
  };
  auto reduce = [&](Bank& b, auto kc, int jj, int s0) {
    constexpr int K = decltype(kc)::value;
    const uint32_t wb = b.cb + (uint32_t)(((b.meta & 0x7f) + 3) & ~3) * 4;   // the task's weights follow its codes
#pragma unroll
    for (int u = 0; u < K; ++u) {
      float w = 1.f;
      if constexpr (WEIGHTED) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(wb + (uint32_t)(jj + u) * 4) : "memory");
      rpw_accumulate<VPL, WEIGHTED, false>(b.acc, b.val[s0 + u], w, 0u, leaky);
    }
  };
  constexpr int S8 = 0, S4 = HB >= 8 ? 8 : 0, S2 = S4 + (HB >= 4 ? 4 : 0), S1 = S2 + (HB >= 2 ? 2 : 0);

  // next record of my stream -> bank; false at the end of the stream
  auto fetch = [&](Bank& b) -> bool {
    if (j >= n_my) return false;
    if (t == 0) mbar_wait_u32(bar0 + 8 * slot, parity);              // first touch of packet j
    const uint32_t sbase = slot0 + slot * G::SLOT_BYTES;
    const int4 rec = lds_i4(sbase + (uint32_t)t * 16);               // {row, meta, slice id, code offset}
    if (rec.y & (int)kPktNoWork) { j = n_my; return false; }         // padding: the segment's last packet ends here
    b.row = (uint32_t)rec.x; b.meta = rec.y; b.aux = (uint32_t)rec.z; b.cb = sbase + (uint32_t)rec.w;
    b.refill = NS;
    if (++t == kPktTasks) {
      b.refill = slot;
      t = 0;
      ++j;
      slot = slot + 1 == (unsigned)NS ? 0u : slot + 1;
      parity ^= (slot == 0u);
    }
    return true;
  };
  // issue everything the task loads: full blocks are gathered and added on the spot, the last (partial)
  // block and the own rows stay in flight
  auto start = [&](Bank& b) {
    const int n = b.meta & 0x7f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) b.acc[i] = 0.f;
    int jl = 0;
    for (; jl + 2 * HB <= n; jl += 2 * HB) {           // long tasks: 2*HB rows in flight per lane
      gather(b, std::integral_constant<int, 2 * HB>(), jl, 0);
      reduce(b, std::integral_constant<int, 2 * HB>(), jl, 0);
    }
    b.jl = jl;
    const int rem = n - jl;                            // < 2*HB: one block per set bit, all issued before the first add
    int jj = jl;
    if constexpr (HB >= 8) { if (rem & 8) { gather(b, std::integral_constant<int, 8>(), jj, S8); jj += 8; } }
    if constexpr (HB >= 4) { if (rem & 4) { gather(b, std::integral_constant<int, 4>(), jj, S4); jj += 4; } }
    if constexpr (HB >= 2) { if (rem & 2) { gather(b, std::integral_constant<int, 2>(), jj, S2); jj += 2; } }
    if (rem & 1) gather(b, std::integral_constant<int, 1>(), jj, S1);
    if (b.meta >= 0) load_own(b.own, b.row);           // slices of long rows: only the finisher needs the own rows
  };
  auto finish = [&](Bank& b) {
    {
      const int rem = (b.meta & 0x7f) - b.jl;
      int jj = b.jl;
      if constexpr (HB >= 8) { if (rem & 8) { reduce(b, std::integral_constant<int, 8>(), jj, S8); jj += 8; } }
      if constexpr (HB >= 4) { if (rem & 4) { reduce(b, std::integral_constant<int, 4>(), jj, S4); jj += 4; } }
      if constexpr (HB >= 2) { if (rem & 2) { reduce(b, std::integral_constant<int, 2>(), jj, S2); jj += 2; } }
      if (rem & 1) reduce(b, std::integral_constant<int, 1>(), jj, S1);
    }
    // the last task of a packet: its codes and weights (and those of the tasks before it) are consumed
    if (b.refill < (unsigned)NS) {
      __syncwarp();
      if (j_issue < n_my) {
        issue_pkt(b.refill, dir_next);
        ++j_issue;
        if (j_issue < n_my) dir_next = load_dir(j_issue);
      }
    }
    const uint32_t row = b.row;
    const bool multi = b.meta < 0;                     // bit 31: slice of a long row
    float (&acc)[VPL] = b.acc;
    Own& own = b.own;

    // ---- long rows: publish the slice sum; reduce through a fan-in-16 ticket tree -------------
    // The last arriver of every group of 16 slices (then of 16 groups, ...) sums them in slice
    // order: deterministic, no float atomics.  Release-only tickets: the reducer reads the partials with
    // L1-bypassing GPU-scope loads that depend on the ticket value.  (An acq_rel ticket or an acquire
    // fence in the last arriver would be the formally complete pattern, but ptxas implements GPU-scope
    // acquire as CCTL.IVALL: one invalidation of the SM's whole L1 per slice, which the gathers of
    // high-degree graphs pay for.)
    bool whole_row = true;
    if (multi) {
      const uint32_t aux = b.aux;                      // global slice id
      const uint32_t lr = __ldg(p.chunk_lr + aux);
      const int64_t cbase = __ldg(p.chunk_base + lr);
      const int nch = (int)(__ldg(p.chunk_base + lr + 1) - cbase);
      int pos = (int)((int64_t)aux - cbase);           // my slice inside the row
      bool active = true;
      whole_row = false;
      unsigned* tk = p.tickets;                        // ticket region of the current level
      for (int sh = 0; active; sh += 4) {              // level stride = 16^level = 1 << sh
        const int lstride = 1 << sh;
        const int gs = pos & ~((16 << sh) - 1);        // members: slots gs + j*lstride, j < 16, below nch
        int members = (nch - gs + lstride - 1) >> sh;
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
                  ld_strong_chunk<CW>(pv[u] + v * CW, part + (int64_t)(c0 + u) * lstride * ROWB + v * CHB);
                } else {
#pragma unroll
                  for (int i = 0; i < CW; ++i) pv[u][v * CW + i] = 0.f;
                }
              }
#pragma unroll
            for (int u = 0; u < 4; ++u) rpw_accumulate<VPL, false, false>(acc, pv[u], 1.f, 0u, leaky);
          }
          if (gs == 0 && 16 * lstride >= nch) {        // that was the whole row
            whole_row = true;
            active = false;
          } else {
            pos = gs;                                  // my sum becomes slot gs of the next level
          }
        }
        tk += p.n_chunks;
      }
      __syncwarp();
      if (whole_row) load_own(own, row);               // slices do not prefetch: only the finisher needs the own rows
    }

    // ---- fused epilogue ------------------------------------------------------------------------
    if (whole_row) {
      const uint64_t off = (uint64_t)row * ROWB + lane_off;       // bytes inside a contiguous [rows, d] table
      if (BWD) {
        // n = G + g + A (sigma' . g_other)      (SURVEY A.2); at the top level g == G
        float o[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) o[i] = own.a[i] + ((flags & F_B) ? own.b[i] : own.a[i]) + acc[i];
#pragma unroll
        for (int v = 0; v < NV; ++v) st_chunk<CW>(o1_base + off + v * CHB, o + v * CW);
        if (flags & F_O2) {                                        // the source of the next level down: sigma'(Z^{l-1}) (.) n
          float om[VPL];
#pragma unroll
          for (int i = 0; i < VPL; ++i) om[i] = ((own.pbits >> i) & 1u) ? o[i] : leaky * o[i];
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
          for (int i = 0; i < VPL; ++i) nxt_e[i] = own.a[i] + act[i];
          if (flags & F_O1) {
#pragma unroll
            for (int v = 0; v < NV; ++v) st_chunk<CW>(o1_base + off + v * CHB, nxt_e + v * CW);
          }
          if (flags & F_O2) {
            float o[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
              o[i] = (flags & F_B) ? own.b[i] + own.a[i] : own.a[i];
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
  };

  if constexpr (BANKS == 1) {
    Bank A;
    while (fetch(A)) {
      start(A);
      finish(A);
    }
  } else {
    Bank A, B;
    bool more = fetch(A);
    if (more) start(A);
    while (more) {
      const bool more_b = fetch(B);
      if (more_b) start(B);
      finish(A);
      if (!more_b) break;
      more = fetch(A);
      if (more) start(A);
      finish(B);
    }
  }
  if (p.trace) {
    __syncthreads();
    if (threadIdx.x == 0) p.trace[blockIdx.x * 4 + 3] = globaltimer_ns();
  }
}

}  // namespace sagnn
