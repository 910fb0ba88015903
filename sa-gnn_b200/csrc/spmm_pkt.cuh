// Packet-stream layer kernel (v10, the default).  Included by spmm.cu after spmm_rpw.cuh (shares its
// per-lane load / store / accumulate helpers).
//
// Same contract as the v8 kernel (one launch = one GNN layer over all T intervals and both
// orientations: LIU-YUXI/SA-GNN model.py:118-127; backward = SURVEY A.2).  What round 2 measured
// (profiles/r2_micro_l1tex_cost.txt, profiles/r2_ncu_*.md) and what this kernel does about it:
//
//   * A bookkeeping-free loop that moves exactly what a layer moves (12 gathered 256-byte rows, the
//     task's own rows, its stores) runs at 63-77 SM cycles per task; v8 and the first packet kernel
//     both sat at ~115 with the issue slots ~57 % busy on ~200 warp instructions per task, three
//     quarters of the tasks having <= 7 edges.  The cost is per TASK, not per edge.
//   * So a task is owned by a LANE GROUP of d/4 lanes (one float4 per lane: 128-bit loads), and a warp
//     runs G = 128/d tasks in lock step (4 at d=32, 2 at d=64, 1 at d>=128): every instruction of the
//     bookkeeping, the gathers and the epilogue serves G tasks.  Tasks arrive sorted by degree, so
//     the tasks of one iteration have (nearly) the same length; control flow follows the longest,
//     shorter ones predicate their loads off.
//   * The schedule is a PACKED TASK STREAM built by the plan: packets of 4 tasks = 4 records
//     {row, n | flags, slice id, code offset} followed by the tasks' edge codes (and weights), in
//     schedule order.  A warp brings its next packets into shared memory with ONE TMA bulk copy
//     each (cp.async.bulk + mbarrier, issued by one lane, SLOTS-1 packets ahead): no per-task
//     record / code requests, no LSU work besides the LDS.128 that read them back.
//   * Packets are dealt statically (packet q of a segment belongs to warp q mod (32 x CTAs of the
//     segment)): the degree-sorted interleave balances itself, no queue atomics.  The CTA's work
//     descriptor comes by value through the kernel parameters, so everything derived from it lives in
//     uniform registers.
//   * Gather: all loads of a (remainder) block are issued before the first add; full blocks of 8
//     slots, then one straight-line case per remainder length; packed FADD2 / FFMA2 accumulation.
//     The task's own dense rows (E^l row, layer-sum row / G and g rows) are loaded with its gathers.
//   * Long rows (deg > 64): <= 64-edge slices, fan-in-16 ticket tree, deterministic (as v8), run per
//     lane group.
#pragma once

namespace sagnn {

constexpr uint32_t kPktNoWork = 0x40000000u;   // record meta bit 30: padding record (end of the segment)

template <int D_, bool WEIGHTED>
struct PktGeo {
  static constexpr int D = D_;
  static constexpr int ROWB = D * 4;
#ifdef SAGNN_PKT_LPT64   // experiment build: lane-group width at d = 64 (8: four tasks per warp, two float4 per lane)
  static constexpr int LPT = D >= 128 ? 32 : (D == 64 ? SAGNN_PKT_LPT64 : D / 4);
#else
  static constexpr int LPT = D >= 128 ? 32 : D / 4;   // lanes per task: one float4 per lane and chunk
#endif
  static constexpr int G = 32 / LPT;                  // tasks a warp runs in lock step
  static constexpr int NV = D / (LPT * 4);            // float4 chunks per lane (2 at d=256)
  static constexpr int VPL = 4 * NV;                  // floats per lane
  static constexpr int CHB = LPT * 16;                // bytes one chunk spans across the lane group
  static constexpr int MPR = D / 4;                   // mask bytes per row: byte q = float4 q, bit i = element 4q+i
#ifdef SAGNN_PKT_GS
  static constexpr int GS = SAGNN_PKT_GS;
#else
  // gather slots per block: 4 (16 value registers; 8 spills at d = 64 under the 64-register cap and measured slower,
  // and at d = 128 it measured 6.4 vs 5.5 ms on the ML-10M shape), 2 at d = 256
  static constexpr int GS = NV == 2 ? 2 : 4;
#endif
  static constexpr int NS = WEIGHTED ? 2 : SAGNN_PKT_SLOTS;    // packets resident / in flight per warp
  // a task's codes: [hot slots, padded to 4][source-row ids, padded to 4] (<= kChunk + 4 words) [+ its weights, same split]
  static constexpr int SLOT_BYTES = kPktTasks * 16 + kPktTasks * (kChunk + 4) * 4 * (WEIGHTED ? 2 : 1);
  static constexpr size_t PKT_SMEM = (size_t)kPktWarps * NS * SLOT_BYTES;
  // the rest of the CTA's shared memory holds staged copies of the source table's most popular rows
  static constexpr int HOT_CAP = (int)((kSmemBudget - PKT_SMEM) / ROWB);
  static constexpr size_t SMEM = PKT_SMEM + (size_t)HOT_CAP * ROWB;
  static_assert(kPktTasks % G == 0, "a packet holds whole lock-step groups");
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

template <int D_, int MODE, bool WEIGHTED, bool RTD, bool HOT_>
__global__ void __launch_bounds__(kPktThreads, 1)
spmm_pkt_kernel(const __grid_constant__ SpmmParams p) {
  using Geo = PktGeo<D_, WEIGHTED>;
  constexpr int D = Geo::D, ROWB = Geo::ROWB, LPT = Geo::LPT, G = Geo::G, NV = Geo::NV, VPL = Geo::VPL, CHB = Geo::CHB,
                MPR = Geo::MPR, GS = Geo::GS, NS = Geo::NS;
  constexpr bool BWD = MODE == MODE_BWD;
  constexpr bool OWN = MODE != MODE_MSG;              // the task has dense rows of its own
  constexpr bool HOT = HOT_;                          // plan with hot slots: popular source rows staged in shared memory
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ __align__(128) unsigned char smem_raw[];      // [packet rings of all warps | staged hot rows]
  __shared__ __align__(8) uint64_t bars[kPktWarps * NS];
  __shared__ __align__(8) uint64_t hot_bar;
  // programmatic dependent launch: the next launch of the chain may take my SM as soon as I leave it and run its
  // own start-up (plan data only) under the tail of this grid; no-ops when launched without the attribute
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  // shuffled: tells the compiler the value is warp-uniform (uniform registers for everything derived from it)
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
  const int grp = lane / LPT;                          // my task inside the lock-step group
  const int piece = lane % LPT;                        // my float4 inside a row chunk
  const unsigned gmask = G == 1 ? FULL : (((1u << LPT) - 1u) << (grp * LPT));   // the lanes of my lane group
  const int leader = grp * LPT;
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
  // which optional tensors this launch has (uniform)
  enum { F_B = 1, F_O1 = 2, F_O2 = 4, F_MK = 8, F_ADDNEXT = 16, F_PM = 32, F_PEER = 64 };
  const uint32_t flags = (b_base ? F_B : 0) | (o1_base ? F_O1 : 0) | (o2_base ? F_O2 : 0) | (mk_base ? F_MK : 0) |
                         (p.out_add_next ? F_ADDNEXT : 0) | (pm_base ? F_PM : 0) |
                         ((RTD && MODE == MODE_FWD && p.peer_n > 0) ? F_PEER : 0);

  uint32_t lane_off = (uint32_t)piece * 16;            // my bytes inside a chunk
  pin32(lane_off);
  auto mask_bits = [&](const uint8_t* base, uint32_t c) -> uint32_t {   // sign bits of row c that belong to my elements
    uint32_t bits = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) bits |= ((uint32_t)__ldg(base + (uint64_t)c * MPR + v * LPT + piece)) << (v * 4);
    return bits;
  };
  const char* src_lane = src + lane_off;
  pin64(src_lane);

  // ---- hot rows: the plan's first `hot_rows` slots of my source table, staged by TMA bulk copies -------
  // Edge codes are hot-first inside every task: the first nh codes are slot numbers, the rest row ids.  Slots
  // that do not fit at this latdim (plan built for a smaller d) are read through their row ids.
  const int n_hot = !HOT ? 0 : (p.hot_rows < Geo::HOT_CAP ? p.hot_rows : Geo::HOT_CAP);   // slots staged in shared memory
  const int32_t* hot_ids = p.hot_ids + (size_t)(seg ^ 1) * kHotRows;
  const uint32_t hot0 = smem_u32(smem_raw) + (uint32_t)Geo::PKT_SMEM;
  if (HOT && n_hot > 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");   // staged rows are table data of the launch before
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&hot_bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx_u32(smem_u32(&hot_bar), (unsigned)n_hot * ROWB);
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < n_hot; sl += kPktThreads)
      tma_bulk_g2s_u32(hot0 + (uint32_t)sl * ROWB, src + (uint64_t)(uint32_t)__ldg(hot_ids + sl) * src_stride, ROWB,
                       smem_u32(&hot_bar));
  }
  const uint32_t hot_lane = hot0 + lane_off;
  const bool all_fit = p.hot_rows <= Geo::HOT_CAP;     // (uniform) every hot slot of the plan is staged

  // ---- my packets -----------------------------------------------------------------------------
  // The first NS + 1 packets of every warp are dealt statically (packet q0 + s*stride of the segment's
  // list), the rest come from the segment's queue head in batches of qb packets per atomic (1 for small
  // segments, up to 8 when a warp has hundreds of packets: all warps of a segment hit ONE counter, and
  // millions of single-packet grabs serialise in its L2 slice), fetched ahead: queue head -> directory
  // entry -> bulk copy are three dependent round trips, each hidden behind a packet's worth of work.  SMs do not run at the same speed (GPC size, die of the L2
  // slice): with a purely static deal the slowest CTA of a segment finished 15-20 % after the mean.
  const unsigned n_pk = cd.n_pk;
  const uint32_t* dir = p.pkt_dir + cd.pkt_begin;     // packet offsets, 16-byte units into the stream
  const unsigned q0 = cd.q0 + (unsigned)warp;
  const unsigned stride = cd.stride;
  const unsigned q_dyn0 = (unsigned)(NS + 1) * stride; // packets below this index are dealt statically
  unsigned* qhead = p.ctrs + seg;

  const uint32_t slot0 = smem_u32(smem_raw) + (uint32_t)warp * (NS * Geo::SLOT_BYTES);
  const uint32_t bar0 = smem_u32(&bars[warp * NS]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto load_dir = [&](unsigned q) -> uint2 {          // {first, end} 16-byte unit of packet q (0,0 past the end)
    if (q >= n_pk) return make_uint2(0u, 0u);
    const uint32_t* dp = dir + q;
    return make_uint2(__ldg(dp), __ldg(dp + 1));
  };
  // next packet of the segment's queue: the atomic is issued now, its result (lane 0's register) is only
  // broadcast when the packet index is needed, one refill later -- nobody waits for the round trip
  unsigned qb = n_pk / (stride * 8u);                  // >= 8 grabs per warp before batches grow
  qb = qb < 1u ? 1u : (qb > 8u ? 8u : qb);
  auto grab_issue = [&]() -> unsigned { return lane == 0 ? atomicAdd(qhead, qb) : 0u; };
  auto grab_value = [&](unsigned raw) -> unsigned { return __shfl_sync(FULL, raw, 0) + q_dyn0; };
  auto issue_pkt = [&](unsigned slot, uint2 dv) {      // one lane, one bulk copy: records + codes (+ weights)
    if (lane == 0) {
      const unsigned bytes = (dv.y - dv.x) * 16u;
      mbar_expect_tx_u32(bar0 + 8 * slot, bytes);
      tma_bulk_g2s_u32(slot0 + slot * Geo::SLOT_BYTES, reinterpret_cast<const char*>(p.pkt_stream) + (uint64_t)dv.x * 16u,
                       bytes, bar0 + 8 * slot);
    }
  };
  // start-up: the directory entries of my NS + 1 static packets are loaded together (independent loads, one DRAM
  // round trip), NS bulk copies go out at once, the (NS+1)-th entry waits for the first free slot; the queue's first
  // ticket is requested now and only read at that first refill -- nothing in the prologue waits for an atomic
  unsigned n_issued = 0;                               // packets of mine brought in (or on their way)
  uint2 dv0[NS + 1];
#pragma unroll
  for (int s = 0; s <= NS; ++s) dv0[s] = load_dir(q0 + (unsigned)s * stride);
#pragma unroll
  for (int s = 0; s < NS; ++s)
    if (dv0[s].y != 0u) { issue_pkt(s, dv0[s]); ++n_issued; }
  uint2 dir_next = dv0[NS];                            // packet of the next refill ((0,0): none)
  unsigned q_raw = 0;                                  // lane 0: ticket of the batch after the current one
  unsigned q_cur = n_pk, q_end = n_pk;                 // current batch: packets [q_cur, q_end) still to fetch
  bool more = dir_next.y != 0u;                        // the queue may still have packets for me
  auto next_packet = [&]() -> unsigned {               // warp-uniform; n_pk or more when the queue is drained
    if (q_cur == q_end && more) {                      // batch used up: move to the one grabbed a batch ago
      q_cur = grab_value(q_raw);
      q_end = q_cur + qb;
      more = q_cur < n_pk;
      if (more) q_raw = grab_issue();
    }
    return q_cur < q_end ? q_cur++ : n_pk;
  };
  if (more) q_raw = grab_issue();
  // everything above read the plan (packet stream, directory, this launch's own queue head); everything below
  // reads or writes tables of the launch before: wait until that grid has completed and its stores are visible
  asm volatile("griddepcontrol.wait;" ::: "memory");

  bool hot_ready = n_hot == 0;                         // the staged rows have landed (waited for on first use)
  unsigned j = 0;                                      // packet the cursor is in (count of mine)
  int t = 0;                                           // next record inside it
  unsigned slot = 0, parity = 0;                       // ring slot of packet j and its mbarrier phase

  while (j < n_issued) {
    if (t == 0) {
      mbar_wait_u32(bar0 + 8 * slot, parity);                                  // first touch of packet j
    }
    const uint32_t sbase = slot0 + slot * Geo::SLOT_BYTES;
    const int4 rec = lds_i4(sbase + (uint32_t)(t + grp) * 16);                // my task: {row, meta, slice id, code offset}
    if (__shfl_sync(FULL, rec.y, 0) & (int)kPktNoWork) break;                 // padding: the segment's last packet ends here
    const bool valid = !(rec.y & (int)kPktNoWork);
    const uint32_t row = (uint32_t)rec.x;
    const int n = valid ? (rec.y & 0x7f) : 0;
    const bool multi = valid && rec.y < 0;                                     // bit 31: slice of a long row
    const int nh = (HOT && valid) ? ((rec.y >> 8) & 0x7f) : 0;                 // my task's leading codes that are hot slots
    const int nc = n - nh;                                                     // ... the rest are source-row ids
    const uint32_t hb = sbase + (uint32_t)rec.w;                               // my task's hot slots ...
    const uint32_t cb = hb + (uint32_t)((nh + 3) & ~3) * 4;                    // ... its source-row ids ...
    const uint32_t hwb = cb + (uint32_t)((nc + 3) & ~3) * 4;                   // ... and the weights, same split
    const uint32_t wb = hwb + (uint32_t)((nh + 3) & ~3) * 4;
    // the lock-step group follows its longest task: cold edges (global gathers) and hot edges (shared memory)
#if SAGNN_PKT_X == 1 || SAGNN_PKT_X >= 3   // timing experiments: no gathers at all
    int nmax = 0, hmax = 0;
#else
    int nmax = n - nh, hmax = nh;
#endif
    if constexpr (G >= 2) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, LPT));
    if constexpr (G >= 4) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 2 * LPT));
    if constexpr (HOT && G >= 2) hmax = max(hmax, __shfl_xor_sync(FULL, hmax, LPT));
    if constexpr (HOT && G >= 4) hmax = max(hmax, __shfl_xor_sync(FULL, hmax, 2 * LPT));
    unsigned refill = NS;                                                      // ring slot these tasks free (NS: none)
    t += G;
    if (t == kPktTasks) {
      refill = slot;
      t = 0;
      ++j;
      slot = slot + 1 == (unsigned)NS ? 0u : slot + 1;
      parity ^= (slot == 0u);
    }

    float acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
    float val[GS][VPL];

    // K slots starting at edge jj: loads of the lanes whose task is that long (all lanes when G == 1)
    auto gather = [&](auto kc, int jj) {
      constexpr int K = decltype(kc)::value;
      int c[K];
      if constexpr (K >= 4) {
#pragma unroll
        for (int g = 0; g < K / 4; ++g) {
          const int4 c4 = lds_i4(cb + (uint32_t)(jj + 4 * g) * 4);
          c[4 * g] = c4.x; c[4 * g + 1] = c4.y; c[4 * g + 2] = c4.z; c[4 * g + 3] = c4.w;
        }
        if constexpr (K % 4 >= 2) { const int2 c2 = lds_i2(cb + (uint32_t)(jj + (K & ~3)) * 4); c[K & ~3] = c2.x; c[(K & ~3) + 1] = c2.y; }
        if constexpr (K % 2 == 1) c[K - 1] = lds_i1(cb + (uint32_t)(jj + K - 1) * 4);
      } else if constexpr (K == 3) {
        const int2 c2 = lds_i2(cb + (uint32_t)jj * 4);
        c[0] = c2.x; c[1] = c2.y; c[2] = lds_i1(cb + (uint32_t)(jj + 2) * 4);
      } else if constexpr (K == 2) {
        const int2 c2 = lds_i2(cb + (uint32_t)jj * 4);
        c[0] = c2.x; c[1] = c2.y;
      } else {
        c[0] = lds_i1(cb + (uint32_t)jj * 4);
      }
      const int left = nc - jj;                         // my task's cold edges from jj on (<= 0: none)
#pragma unroll
      for (int u = 0; u < K; ++u) {
        if (G == 1 || u < left) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            ldg_chunk<4>(val[u] + v * 4, src_lane + (uint64_t)(uint32_t)c[u] * src_stride + v * CHB);
        }
      }
    };
    auto reduce = [&](auto kc, int jj) {
      constexpr int K = decltype(kc)::value;
      const int left = nc - jj;
#pragma unroll
      for (int u = 0; u < K; ++u) {
        if (G == 1 || u < left) {
          float w = 1.f;
          if constexpr (WEIGHTED) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(wb + (uint32_t)(jj + u) * 4) : "memory");
          rpw_accumulate<VPL, WEIGHTED, false>(acc, val[u], w, 0u, leaky);
        }
      }
    };

    // ---- gather-reduce: full blocks on the spot, the remainder stays in flight under the own-row loads ----
#if SAGNN_PKT_PF
    // While a block's loads fly, the NEXT block's rows are pulled into L1 (no registers, no scoreboard): its loads
    // then find them next to the SM instead of paying another L2 round trip.  Lane 2*slot + line fetches one 128-byte line.
    auto prefetch_block = [&](int jj) {
      constexpr int LINES = ROWB / 128 > 0 ? ROWB / 128 : 1;
      if (piece < GS * LINES) {
        const int e = jj + piece / LINES;
        if (e < nc) {
          const uint32_t code = (uint32_t)lds_i1(cb + (uint32_t)e * 4);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (uint64_t)code * src_stride + (uint32_t)(piece % LINES) * 128));
        }
      }
    };
#endif
    int jl = 0;
    for (; jl + GS <= nmax; jl += GS) {
      gather(std::integral_constant<int, GS>(), jl);
#if SAGNN_PKT_PF
      if (jl + GS < nmax) prefetch_block(jl + GS);
#endif
      reduce(std::integral_constant<int, GS>(), jl);
    }
    const int rem = nmax - jl;                          // < GS, warp-uniform: one straight-line case per length
    switch (rem) {
      case 1: gather(std::integral_constant<int, 1>(), jl); break;
      case 2: gather(std::integral_constant<int, 2>(), jl); break;
      case 3: gather(std::integral_constant<int, 3>(), jl); break;
      case 4: if constexpr (GS > 4) gather(std::integral_constant<int, 4>(), jl); break;
      case 5: if constexpr (GS > 4) gather(std::integral_constant<int, 5>(), jl); break;
      case 6: if constexpr (GS > 4) gather(std::integral_constant<int, 6>(), jl); break;
      case 7: if constexpr (GS > 4) gather(std::integral_constant<int, 7>(), jl); break;
      default: break;
    }
    // own rows: a = E^l / G row, b = layer-sum / running-gradient row, pbits = sign bits one level down
    float own_a[VPL], own_b[VPL];                       // always loaded before the epilogue reads them
    uint32_t pbits = 0;
    auto load_own = [&]() {
      if (!OWN) return;
      const char* pa = a_base + (uint64_t)row * a_stride + lane_off;
#pragma unroll
      for (int v = 0; v < NV; ++v) ldg_chunk<4>(own_a + v * 4, pa + v * CHB);
      if (flags & F_B) {
        const char* pb = b_base + (uint64_t)row * (RTD ? b_stride : a_stride) + lane_off;
#pragma unroll
        for (int v = 0; v < NV; ++v) ldg_chunk<4>(own_b + v * 4, pb + v * CHB);
      }
      if (BWD && (flags & F_PM)) pbits = mask_bits(pm_base, row);
    };
#if SAGNN_PKT_X != 2 && SAGNN_PKT_X != 4   // (2, 4 = timing experiments without own-row loads)
    if (valid && !multi) load_own();                    // slices of long rows: only the finisher needs the own rows
#endif

    // ---- hot edges: staged rows out of shared memory, added while the cold remainder and the own rows fly ----
    if (HOT && hmax > 0) {
      if (!hot_ready) { mbar_wait_u32(smem_u32(&hot_bar), 0); hot_ready = true; }
      for (int hj = 0; hj < hmax; hj += 4) {
        const int4 hc = lds_i4(hb + (uint32_t)hj * 4);
        const int sl[4] = {hc.x, hc.y, hc.z, hc.w};
        const int left = nh - hj;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (u < left) {
            float hv[VPL];
            if (all_fit || sl[u] < n_hot) {
#pragma unroll
              for (int v = 0; v < NV; ++v) lds_chunk<4>(hv + v * 4, hot_lane + (uint32_t)sl[u] * ROWB + v * CHB);
            } else {                                    // slot beyond what fits at this latdim: through its row id
              const uint32_t r = (uint32_t)__ldg(hot_ids + sl[u]);
#pragma unroll
              for (int v = 0; v < NV; ++v) ldg_chunk<4>(hv + v * 4, src_lane + (uint64_t)r * src_stride + v * CHB);
            }
            float w = 1.f;
            if constexpr (WEIGHTED) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(hwb + (uint32_t)(hj + u) * 4) : "memory");
            rpw_accumulate<VPL, WEIGHTED, false>(acc, hv, w, 0u, leaky);
          }
        }
      }
    }
    switch (rem) {
      case 1: reduce(std::integral_constant<int, 1>(), jl); break;
      case 2: reduce(std::integral_constant<int, 2>(), jl); break;
      case 3: reduce(std::integral_constant<int, 3>(), jl); break;
      case 4: if constexpr (GS > 4) reduce(std::integral_constant<int, 4>(), jl); break;
      case 5: if constexpr (GS > 4) reduce(std::integral_constant<int, 5>(), jl); break;
      case 6: if constexpr (GS > 4) reduce(std::integral_constant<int, 6>(), jl); break;
      case 7: if constexpr (GS > 4) reduce(std::integral_constant<int, 7>(), jl); break;
      default: break;
    }

    // the last tasks of a packet: its codes and weights are consumed, bring the next packet into its slot
    if (refill < (unsigned)NS) {
      __syncwarp();
      if (dir_next.y != 0u) {                            // the queue had a packet left for me
        issue_pkt(refill, dir_next);
        ++n_issued;
        dir_next = load_dir(next_packet());
      }
    }

    // ---- long rows: publish the slice sum; reduce through a fan-in-16 ticket tree -------------
    // The last arriver of every group of 16 slices (then of 16 groups, ...) sums them in slice
    // order: deterministic, no float atomics.  Tickets are release increments; the last arriver of a group
    // (only it) issues the matching acquire fence before reading the partials with L1-bypassing GPU-scope
    // loads.  (An acq_rel ticket on EVERY slice was measured: ptxas implements GPU-scope acquire as
    // CCTL.IVALL, one invalidation of the SM's whole L1 per slice, which the gathers of high-degree graphs
    // pay for; one per group of 16 is not.)  Lane groups run this independently (group-masked shuffles).
    bool whole_row = valid;
#if SAGNN_PKT_X == 3        // timing experiment: no slice publication / ticket tree
    if (multi) whole_row = false;
    if (false) {
#else
    if (multi) {
#endif
      const uint32_t aux = (uint32_t)rec.z;            // global slice id
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
        for (int v = 0; v < NV; ++v) st_chunk<4>(mine + v * CHB, acc + v * 4);
        __syncwarp(gmask);
        unsigned* my_tk = tk + (cbase + gs);           // one ticket per group, named by its first slot
        unsigned old = 0;
        if (lane == leader) old = ticket_release_add(my_tk);
        old = __shfl_sync(gmask, old, leader);
        if (old != (unsigned)(members - 1)) {
          active = false;                              // someone else finishes this group
        } else {
          ticket_acquire_fence();
          if (lane == leader) *my_tk = 0u;             // ready for the next launch
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
                  ld_strong_chunk<4>(pv[u] + v * 4, part + (int64_t)(c0 + u) * lstride * ROWB + v * CHB);
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) pv[u][v * 4 + i] = 0.f;
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
      __syncwarp(gmask);
      if (whole_row) load_own();
    }

    // ---- fused epilogue ------------------------------------------------------------------------
#if SAGNN_PKT_X == 4
    own_a[0] = own_a[1] = own_a[2] = own_a[3] = 1.f; own_b[0] = own_b[1] = own_b[2] = own_b[3] = 1.f;
#endif
#if SAGNN_PKT_X == 2 || SAGNN_PKT_X == 5
    if (whole_row && acc[0] + own_a[0] + own_b[0] == 123.456f) {   // keeps the loads alive, never true
#else
    if (whole_row) {
#endif
      const uint64_t off = (uint64_t)row * ROWB + lane_off;       // bytes inside a contiguous [rows, d] table
      if (BWD) {
        // n = G + g + A (sigma' . g_other)      (SURVEY A.2); at the top level g == G
        float o[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) o[i] = own_a[i] + ((flags & F_B) ? own_b[i] : own_a[i]) + acc[i];
#pragma unroll
        for (int v = 0; v < NV; ++v) st_chunk<4>(o1_base + off + v * CHB, o + v * 4);
        if (flags & F_O2) {                                        // the source of the next level down: sigma'(Z^{l-1}) (.) n
          float om[VPL];
#pragma unroll
          for (int i = 0; i < VPL; ++i) om[i] = ((pbits >> i) & 1u) ? o[i] : leaky * o[i];
#pragma unroll
          for (int v = 0; v < NV; ++v) st_chunk<4>(o2_base + off + v * CHB, om + v * 4);
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
          for (int v = 0; v < NV; ++v) st_chunk<4>(o1_base + off + v * CHB, act + v * 4);
        } else {
          float nxt_e[VPL];                                         // E^{l+1} = E^l + lrelu(Z^l)
#pragma unroll
          for (int i = 0; i < VPL; ++i) nxt_e[i] = own_a[i] + act[i];
          if (flags & F_O1) {
#pragma unroll
            for (int v = 0; v < NV; ++v) st_chunk<4>(o1_base + off + v * CHB, nxt_e + v * 4);
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
            for (int v = 0; v < NV; ++v) stcs_chunk<4>(dst + v * CHB, o + v * 4);
          }
          if (flags & F_MK) {                                       // one mask byte per float4: no shuffles
            uint8_t* mrow = mk_base + (uint64_t)row * MPR;
#pragma unroll
            for (int v = 0; v < NV; ++v) mrow[v * LPT + piece] = (uint8_t)((bits >> (v * 4)) & 0xfu);
          }
        }
      }
    }
  }
  if (HOT && !hot_ready) mbar_wait_u32(smem_u32(&hot_bar), 0);   // never leave while the staging copies are still landing
  if (p.trace) {
    __syncthreads();
    if (threadIdx.x == 0) p.trace[blockIdx.x * 4 + 3] = globaltimer_ns();
  }
}

}  // namespace sagnn
