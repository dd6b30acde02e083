// k_tc_edge: fused edge model + scatter_mean of one CSPLayer on the tensor cores
// (CSPLayer.edge_model and the aggregation of node_model, cspnet.py:129-160).
//
// Work item = (variant, tile); a tile is 128 edge rows = whole (i, all j) segments of equal
// length n, generated on the fly from the per-tile (i, j) row tables -- no edge_index,
// nothing of size O(E) ever touches HBM.  TMEM lanes = output channels, TMEM columns = edges
// ("transposed" orientation), four 128-channel x 128-edge fp32 units U0..U3 = all 512 columns.
//
//   init   U_m    = P_i[i] + P_j[j]          tcgen05.st, thread = channel (coalesced gathers);
//                                            done for tile t+1 while GEMM2 of tile t runs
//   GEMM1  U_m   += W_fd[m] emb^T            A = weights (K-major image, bulk-copied, UBLKCP)
//                                            B = sinusoid embedding built in smem by a rotation
//                                                recurrence (thread = edge row), K-major
//   E1     a1     = SiLU(U) -> fp16          MN-major B operand of GEMM2 in smem
//   GEMM2  U_m'   = W2[m'] a1^T              one unit after the other, so that
//   E2     agg_i  = mean_j SiLU(U + b2)      E2 of unit m' overlaps GEMM2 of unit m'+1;
//                                            the segmented mean is an in-thread running sum
//
//   warps 0-15: worker warp w owns TMEM lane quarter w%4 of unit w/4 (init, E1, E2) and, as
//               member of embedding group w/4, the edge row (w%4)*32+lane of every 4th K chunk
//   warp 16   : MMA issue (one lane), TMEM alloc        warp 17: weight loader (bulk copies)
#include "cb2_tc_edge_epi.cuh"

namespace cb2 {

using namespace ptx;

constexpr int TE_KC = 32;                       // K per GEMM1 pipeline stage
constexpr int TE_WSTAGES = 3;                   // weight ring R1 (always available)
constexpr int TE_WEXTRA = 2;                    // two more stages inside the a1 region, usable during GEMM1 only
constexpr int TE_WEXTRA_OFF = 65536;            // ... placed after the embedding ring
constexpr int TE_W_BYTES = 32768;               // GEMM1: [4 k8][512 ch][16 B]; GEMM2: [16 k8][128 ch][16 B]
constexpr int TE_ASLOTS = 8;                    // two slots per embedding group
constexpr int TE_A_BYTES = 128 * TE_KC * 2;     // 8 KB: [4 k8][128 edges][16 B]
constexpr int TE_AREGION = 128 * H * 2;         // 128 KB: a1, MN-major [64 k8][16 e8][8 k][8 e]
constexpr int TE_W_OFF = TE_AREGION;
constexpr int TE_BAR_OFF = TE_W_OFF + TE_WSTAGES * TE_W_BYTES;
constexpr int TE_TAB_OFF = TE_BAR_OFF + 320;    // 2 buffers x (off_i[128], off_j[128]) uint32
constexpr int TE_SMEM = TE_TAB_OFF + 2 * 1024;
constexpr int TE_WORKERS = 512;
constexpr int TE_NISSUE = 3;                    // MMA-issuing threads (warps 16..18), loader = warp 19
constexpr int TE_THREADS = TE_WORKERS + 32 * (TE_NISSUE + 1);

constexpr int TE_NCH1 = DIS / TE_KC;            // 24 stages of K=32
constexpr int TE_NCH2 = 4;                      // per output unit: 4 stages of K=128
#ifndef CB2_TE_DEFER
#define CB2_TE_DEFER 4
#endif
constexpr int TE_DEFER = CB2_TE_DEFER;                     // GEMM1 chunks issued for units 0..2 before unit 3 is clear (<= 5 W stages)
static_assert(TE_SMEM <= 232448, "shared memory budget");
static_assert(TE_NCH2 >= TE_NISSUE, "every issuer must own a stage of every output unit");

// single-thread roles on the critical path poll their barriers (see mbar_wait_spin)
#define MBAR_WAIT_CRIT mbar_wait_spin

__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Note on mbarrier parities: a wait only compares one parity bit, so a thread that asks for phase
// k+1 of a barrier whose phase k is still pending sees "completed".  Every waiter here is therefore
// kept within one phase of its barrier: ring slots/stages that are waited for by DIFFERENT issuer
// threads from one use to the next are only reached after the previous use has been consumed
// (GEMM1 weight stage kc % 5: the issuers are gated by the embedding producers, which run in
// lockstep with the MMAs, and at the start of an item by the common wait on unit 3), and the
// embedding slots of the pair kernel, whose period is shorter than its weight run-ahead, are always
// waited for by the same thread.
#define TE_WORKER_BARRIER() asm volatile("bar.sync 1, 512;" ::: "memory")

// development aid: timeline of CTA 0's items 1..3 (clock64 stamps), read back by cb2_debug_edge_timeline()
__device__ long long g_edge_dbg[3 * 96];
#ifdef CB2_EDGE_TIMELINE
#define TE_STAMP(slot)                                                      \
  do {                                                                      \
    if (blockIdx.x == 0 && it >= 1 && it <= 3) g_edge_dbg[(it - 1) * 96 + (slot)] = clock64(); \
  } while (0)
#else
#define TE_STAMP(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(TE_THREADS, 1) k_tc_edge(TcEdgeArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + TE_BAR_OFF;
  auto a_full = [&](int s) { return bars + 8 * s; };            // 8
  auto a_empty = [&](int s) { return bars + 64 + 8 * s; };      // 8
  auto w_full = [&](int s) { return bars + 128 + 8 * s; };      // 5
  auto w_empty = [&](int s) { return bars + 168 + 8 * s; };     // 5
  const uint32_t acc1_full = bars + 208, a1_ready = bars + 216;
  auto acc2_full = [&](int u) { return bars + 224 + 8 * u; };
  auto acc_init = [&](int u) { return bars + 256 + 8 * u; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TE_BAR_OFF + 288);

  // weight stage s: 0..2 = ring R1, 3..4 = extra stages inside the a1 region (GEMM1 only)
  auto w_addr = [&](int s) {
    return s < TE_WSTAGES ? sbase + TE_W_OFF + s * TE_W_BYTES : sbase + TE_WEXTRA_OFF + (s - TE_WSTAGES) * TE_W_BYTES;
  };
  uint32_t *tab = reinterpret_cast<uint32_t *>(smem + TE_TAB_OFF);   // [buf][0: off_i, 1: off_j][128]

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  if (tid == 0) {
    for (int s = 0; s < TE_ASLOTS; s++) { mbar_init(a_full(s), 128); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < TE_WSTAGES + TE_WEXTRA; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    *reinterpret_cast<volatile uint32_t *>(smem + TE_BAR_OFF + 296) = 0u;
    mbar_init(acc1_full, TE_NISSUE);
    mbar_init(a1_ready, TE_WORKERS);
    for (int u = 0; u < 4; u++) { mbar_init(acc2_full(u), TE_NISSUE); mbar_init(acc_init(u), 128); }
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int n_items = g.n_tiles * g.V;

  if (warp == 16 + TE_NISSUE) {
    // ------------------------------ weight loader ------------------------------
    if (lane == 0) {
      uint32_t use[TE_WSTAGES + TE_WEXTRA] = {0, 0, 0, 0, 0};
      uint32_t it = 0;
      auto load_stage = [&](int st, const __half *src) {
        MBAR_WAIT_CRIT(w_empty(st), (use[st] & 1) ^ 1);
        mbar_arrive_expect_tx(w_full(st), TE_W_BYTES);
        bulk_g2s(w_addr(st), src, TE_W_BYTES, w_full(st));
        use[st]++;
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        bool a1_dead = (it == 0);
#pragma unroll
        for (int kc = 0; kc < TE_NCH1; kc++) {       // GEMM1: five stages
          const int st = kc % (TE_WSTAGES + TE_WEXTRA);
          if (st >= TE_WSTAGES && !a1_dead) {         // the extra stages alias a1 of the previous item
            MBAR_WAIT_CRIT(acc2_full(3), (it - 1) & 1);
            a1_dead = true;
          }
          load_stage(st, g.w_fd_t + (int64_t)kc * (TE_W_BYTES / 2));
        }
#pragma unroll
        for (int c2 = 0; c2 < 4 * TE_NCH2; c2++)     // GEMM2: ring R1 only
          load_stage(c2 % TE_WSTAGES, g.w2_t + (int64_t)c2 * (TE_W_BYTES / 2));
      }
    }
  } else if (warp >= 16 && warp < 16 + TE_NISSUE) {
    // ------------------------------ MMA issuers ------------------------------
    // Any synchronisation point in an issuing thread (an mbarrier wait, even a plain shared-memory
    // poll) idles the tensor pipe for ~170 cycles because only ~2 MMAs are queued ahead (measured
    // with scripts/mma_bench.cu: 8 MMAs + 2 waits = 976 cycles instead of 512).  Three threads
    // therefore issue alternate K chunks: while one waits, the others' MMAs keep the pipe busy.
    // All MMAs accumulate (the units are pre-loaded), so their relative order is irrelevant.
    if (lane == 0) {
      const int ii = warp - 16;
      constexpr uint32_t idesc_kk = idesc_f16_f32(128, 128);
      constexpr uint32_t idesc_kmn = idesc_b_mn(idesc_f16_f32(128, 128));
      constexpr int NS = TE_WSTAGES + TE_WEXTRA;
      uint32_t it = 0;
      // descriptors = base (LBO/SBO/version fields + smem base) + (byte offset >> 4) in the low word
      const uint64_t d_lbo2k = smem_desc_kmajor(sbase, 2048, 128);   // emb slots, a1, W2 stages
      const uint64_t d_lbo8k = smem_desc_kmajor(sbase, 8192, 128);   // W_fd stages
      auto w_off = [](int st) { return st < TE_WSTAGES ? TE_W_OFF + st * TE_W_BYTES : TE_WEXTRA_OFF + (st - TE_WSTAGES) * TE_W_BYTES; };
      // uses of weight stage st per item: GEMM1 chunk kc uses stage kc % 5, GEMM2 stage c2 uses c2 % 3
      auto uses_per_item = [](int st) { return st == 0 ? 11u : st < 3 ? 10u : st == 3 ? 5u : 4u; };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        if (ii == 0) TE_STAMP(0);
        // Units 0..2 are cleared long before unit 3 (the last one GEMM2 completes, its E2 is the
        // tail of the previous item): the first TE_DEFER chunks are issued for units 0..2 only, their
        // unit-3 MMAs follow once unit 3 is clear (the chunks' slots and stages are held until then;
        // the rings are deep enough: 8 slots, 5 stages).
        for (int u = 0; u < 3; u++) MBAR_WAIT_CRIT(acc_init(u), it & 1);
        tc_fence_after_sync();
        if (ii == 0) TE_STAMP(1);
        auto g1_mmas = [&](int as, int ws, int m_lo, int m_hi) {
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const uint64_t bd = d_lbo2k + (uint64_t)((as * TE_A_BYTES + 2 * j * 2048) >> 4);
#pragma unroll
            for (int m = 0; m < 4; m++) {
              if (m < m_lo || m >= m_hi) continue;
              const uint64_t ad = d_lbo8k + (uint64_t)((w_off(ws) + 2 * j * 8192 + m * 2048) >> 4);
              umma_f16(tmem + m * 128, ad, bd, idesc_kk, 1u);
            }
          }
        };
        // GEMM1: U_m += W_fd[m] emb^T
#pragma unroll
        for (int kc = 0; kc < TE_DEFER; kc++) {
          if (kc % TE_NISSUE != ii) continue;
          const int as = (kc & 3) + 4 * ((kc >> 2) & 1), ws = kc % NS;
          MBAR_WAIT_CRIT(a_full(as), (it * 3 + (kc >> 3)) & 1);
          MBAR_WAIT_CRIT(w_full(ws), (it * uses_per_item(ws) + kc / NS) & 1);
          tc_fence_after_sync();
          g1_mmas(as, ws, 0, 3);
        }
        MBAR_WAIT_CRIT(acc_init(3), it & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int kc = 0; kc < TE_DEFER; kc++) {
          if (kc % TE_NISSUE != ii) continue;
          const int as = (kc & 3) + 4 * ((kc >> 2) & 1), ws = kc % NS;
          g1_mmas(as, ws, 3, 4);
          umma_commit(a_empty(as));
          umma_commit(w_empty(ws));
        }
#pragma unroll
        for (int kc = TE_DEFER; kc < TE_NCH1; kc++) {
          if (kc % TE_NISSUE != ii) continue;
          const int as = (kc & 3) + 4 * ((kc >> 2) & 1), ws = kc % NS;
          MBAR_WAIT_CRIT(a_full(as), (it * 3 + (kc >> 3)) & 1);
          MBAR_WAIT_CRIT(w_full(ws), (it * uses_per_item(ws) + kc / NS) & 1);
          tc_fence_after_sync();
          g1_mmas(as, ws, 0, 4);
          umma_commit(a_empty(as));
          umma_commit(w_empty(ws));
        }
        umma_commit(acc1_full);          // count TE_NISSUE: complete when every issuer's GEMM1 MMAs are done
        if (ii == 0) TE_STAMP(2);
        // GEMM2: U_m' += W2[m'] a1^T (units pre-loaded with b2), unit after unit
        MBAR_WAIT_CRIT(a1_ready, it & 1);
        tc_fence_after_sync();
        if (ii == 0) TE_STAMP(3);
#pragma unroll
        for (int c2 = 0; c2 < 4 * TE_NCH2; c2++) {
          if (c2 % TE_NISSUE != ii) continue;
          const int u = c2 / TE_NCH2, kc = c2 % TE_NCH2, ws = c2 % TE_WSTAGES;
          MBAR_WAIT_CRIT(w_full(ws), (it * uses_per_item(ws) + 5 + c2 / TE_WSTAGES) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint64_t ad = d_lbo2k + (uint64_t)((w_off(ws) + 2 * j * 2048) >> 4);
            const uint64_t bd = d_lbo2k + (uint64_t)(((kc * 16 + 2 * j) * 2048) >> 4);
            umma_f16(tmem + u * 128, ad, bd, idesc_kmn, 1u);
          }
          umma_commit(w_empty(ws));
          // last stage of this issuer within unit u -> its share of "unit u complete"
          if (c2 + TE_NISSUE >= (u + 1) * TE_NCH2) {
            umma_commit(acc2_full(u));
            if (c2 == (u + 1) * TE_NCH2 - 1) TE_STAMP(4 + u);
          }
        }
      }
    }
  } else {
    // ------------------------------ workers (512 threads) ------------------------------
    const int q = warp % 4, u4 = warp / 4;       // TMEM lane quarter; unit / embedding group
    const int r = q * 32 + lane;                 // edge row owned while producing the embedding
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + u4 * 128;
    const int c = u4 * 128 + q * 32 + lane;      // output channel owned in the epilogues
    const __half *Pc = g.P + c;
    const float bias = 0.5f * __ldg(g.b2 + c);      // the units hold (W2 a1 + b2) / 2 (cb2_tc.cuh silu_of_half)
    const int oc = g.agg_col + c;
    __half *out = g.agg_kt > 0 ? g.agg16 + (int64_t)(oc >> 3) * 1024 + (oc & 7) : g.agg16 + oc;
    const AggStride agg_ld = g.agg_kt > 0 ? AggStride{(int64_t)128 * g.agg_kt, 8, 0u}
                                          : AggStride{(int64_t)128 * g.ld_agg, (int)g.ld_agg, 0u};
    const A1Dst a1_dst{sbase + (uint32_t)((c / 8) * 2048 + (c % 8) * 16), 0u};

    // The (i, j) node ids of this thread's edge row are fetched one item ahead (fetch_rows); the
    // tile tables (group 0 writes them) and the fractional-coordinate difference follow half an
    // item later (publish_rows), so that no global-load latency sits on the critical path.
    int ri_p = -1, rj_p = 0;
    auto fetch_rows = [&](int item) {
      const int tile = item % g.n_tiles;
      ri_p = g.row_i[(int64_t)tile * 128 + r];
      rj_p = g.row_j[(int64_t)tile * 128 + r];
    };
    auto publish_rows = [&](int item, int buf, float (&dl)[3]) {
      const int v = item / g.n_tiles;
      dl[0] = dl[1] = dl[2] = 0.f;
      if (ri_p >= 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) dl[d] = g.x[(int64_t)rj_p * 3 + d] - g.x[(int64_t)ri_p * 3 + d];
      }
      if (u4 == 0) {
        const uint32_t vbase = (uint32_t)v * (uint32_t)g.N;
        uint32_t oi = TE_PAD, oj = vbase * (uint32_t)H2 + (uint32_t)H;
        if (ri_p >= 0) {
          oi = (vbase + (uint32_t)ri_p) * (uint32_t)H2;
          oj = (vbase + (uint32_t)rj_p) * (uint32_t)H2 + (uint32_t)H;
        }
        tab[buf * 256 + r] = oi;
        tab[buf * 256 + 128 + r] = oj;
      }
    };
    // pull the tile's rows of P towards L2 long before E1 gathers them
    auto prefetch_rows = [&](int buf) {
      if (warp < 8) {
        const int e = (warp * 32 + lane) >> 1;
        const uint32_t o = tab[buf * 256 + ((lane & 1) ? 128 : 0) + e];
        if (o != TE_PAD) {
          const char *pp = reinterpret_cast<const char *>(g.P + o);
#pragma unroll
          for (int l = 0; l < 8; l++) prefetch_l2(pp + l * 128);
        }
      }
    };
    // GEMM1 accumulates onto zeros (the hoisted terms P_i + P_j are added in E1)
    auto clear_unit = [&]() {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; j++) z[j] = 0u;
#pragma unroll
      for (int cb = 0; cb < 4; cb++) tmem_st32(taddr + cb * 32, z);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(acc_init(u4));
    };

    float dlt[3] = {0.f, 0.f, 0.f}, dlt_next[3] = {0.f, 0.f, 0.f};
    uint32_t it = 0;
    // prologue: tables + accumulator init of the first item, row ids of the second
    fetch_rows(blockIdx.x);
    publish_rows(blockIdx.x, 0, dlt);
    TE_WORKER_BARRIER();
    clear_unit();
    if ((int)(blockIdx.x + gridDim.x) < n_items) fetch_rows(blockIdx.x + gridDim.x);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
      const int buf = it & 1;
      const int tile = item % g.n_tiles;
      const int next = item + gridDim.x;
      const bool has_next = next < n_items;
      // the embedding ring aliases a1 of the previous item: all of its GEMM2 must have completed
      if (it > 0) mbar_wait(acc2_full(3), (it - 1) & 1);
      const int n = g.seg_n[tile];
      const uint32_t *t_oi = tab + buf * 256;
      // ---- sinusoid embedding: group u4 builds the chunks with kc % 4 == u4 into slot u4 ----
      {
        const bool valid = t_oi[r] != TE_PAD;
#pragma unroll 1
        for (int d = 0; d < 3; d++) {
          float s1, c1, s48, c48, sk, ck;
          sincospif(2.0f * dlt[d], &s1, &c1);
          sincospif(96.0f * dlt[d], &s48, &c48);
          sincospif((float)(32 * u4) * dlt[d], &sk, &ck);      // frequency 16*u4
#pragma unroll 1
          for (int half = 0; half < 2; half++) {                  // chunks m = u4 and u4 + 4 of this dimension
            const int cidx = d * 2 + half;                        // this group's chunk number within the item
            const int as = u4 + 4 * (cidx & 1);                   // two slots per group, used alternately
            const uint32_t use = it * 3 + (cidx >> 1);            // how often slot `as` has been used
            mbar_wait(a_empty(as), (use & 1) ^ 1);
            if (tid == 0) TE_STAMP(32 + 2 * cidx);
            uint8_t *slot = smem + as * TE_A_BYTES + r * 16;
#pragma unroll
            for (int p = 0; p < 4; p++) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; e++) {
                w[e] = valid ? pack_half2(sk, ck) : 0u;
                const float sn = fmaf(sk, c1, ck * s1);
                const float cn = fmaf(ck, c1, -sk * s1);
                sk = sn; ck = cn;
              }
              *reinterpret_cast<uint4 *>(slot + p * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            {  // jump over the other three groups' 48 frequencies
              const float sn = fmaf(sk, c48, ck * s48);
              const float cn = fmaf(ck, c48, -sk * s48);
              sk = sn; ck = cn;
            }
            fence_proxy_async_smem();
            mbar_arrive(a_full(as));
            if (tid == 0) TE_STAMP(33 + 2 * cidx);
          }
        }
      }
      // every group has finished the previous item (its tables are dead): publish the next item's
      TE_WORKER_BARRIER();
      if (has_next) publish_rows(next, buf ^ 1, dlt_next);
      // ---- E1: a1 = SiLU(U + P_i + P_j), thread = channel, MN-major fp16 operand of GEMM2 ----
      if (tid == 0) TE_STAMP(8);
      const E1Cg cgk{g.cg ? g.cg + c : nullptr, g.node2graph, (uint32_t)(item / g.n_tiles) * (uint32_t)g.N, nullptr};
      e1_dispatch(n, taddr, Pc, cgk, t_oi, t_oi + 128, a1_dst, acc1_full, it & 1);
      if (tid == 0) TE_STAMP(9);
      {  // pre-load the unit with b2: every GEMM2 MMA accumulates, so the issuers need no ordering
        uint32_t bv[32];
#pragma unroll
        for (int j = 0; j < 32; j++) bv[j] = __float_as_uint(bias);
#pragma unroll
        for (int cb = 0; cb < 4; cb++) tmem_st32(taddr + cb * 32, bv);
        tmem_st_wait();
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(a1_ready);
      if (tid == 0) TE_STAMP(10);
      TE_WORKER_BARRIER();                         // the next item's tables are visible to everybody
      if (has_next) prefetch_rows(buf ^ 1);
      if (next + (int)gridDim.x < n_items) fetch_rows(next + gridDim.x);
      // ---- E2: agg_i = mean_j SiLU(U + b2); then the unit is re-initialised for the next item ----
      {
        MBAR_WAIT_WORKER(acc2_full(u4), it & 1);
        tc_fence_after_sync();
        if (lane == 0 && q == 0) TE_STAMP(11 + 4 * u4);
        e2_dispatch(n, taddr, bias, t_oi, out, agg_ld);
        if (lane == 0 && q == 0) TE_STAMP(12 + 4 * u4);
        tc_fence_before_sync();
        if (has_next) clear_unit();
        if (lane == 0 && q == 0) TE_STAMP(13 + 4 * u4);
      }
      dlt[0] = dlt_next[0]; dlt[1] = dlt_next[1]; dlt[2] = dlt_next[2];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

int debug_edge_timeline(long long *out96) {
  CB2_CUDA_OK(cudaMemcpyFromSymbol(out96, g_edge_dbg, sizeof(long long) * 3 * 96));
  return CB2_OK;
}

int launch_tc_edge2(const TcEdgeArgs &a, int n_sm, cudaStream_t st);   // cb2_tc_edge2.cu

int launch_tc_edge(const TcEdgeArgs &a, int n_sm, cudaStream_t st) {
  const int n_items = a.n_tiles * a.V;
  if (n_items == 0) return CB2_OK;
  // both CFG variants: the CTA-pair kernel shares the sinusoid GEMM between them (default)
  if (a.V == 2 && !a.single_cta && n_sm >= 2) return launch_tc_edge2(a, n_sm, st);
  if ((uint64_t)a.V * (uint64_t)a.N * (uint64_t)H2 >= (1ull << 32))
    return fail(CB2_ERR_UNSUPPORTED, "tensor-core edge kernel: V*N*1024 must fit 32 bits (shard the batch)");
  // per-device function attribute: set on every launch (cheap, legal during stream capture)
  CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_edge, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM));
  const int grid = n_items < n_sm ? n_items : n_sm;
  k_tc_edge<<<grid, TE_THREADS, TE_SMEM, st>>>(a);
  CB2_LAUNCH_OK("k_tc_edge");
  return CB2_OK;
}

}  // namespace cb2
