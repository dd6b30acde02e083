// k_tc_edge2: fused edge model + scatter_mean of one CSPLayer for BOTH classifier-free-guidance
// variants at once on a CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256)
// (CSPLayer.edge_model and the aggregation of node_model, cspnet.py:129-160; the two decoder calls
// of Chemeleon.model_predictions, chemeleon.py:260-285, share the state and hence the edge embedding).
//
//   e_ij^v = SiLU(W2 SiLU(P^v_i[i] + cg[g] + P^v_j[j] + W_fd emb(x_j - x_i)) + b2),  agg^v_i = mean_j e_ij^v
//
// The sinusoid term W_fd emb(x_j - x_i) -- 60 % of the edge FLOPs -- does not depend on the variant:
// it is computed ONCE per edge (the one-CTA kernel k_tc_edge computes it once per (edge, variant)).
// A work item is one tile of 128 edge rows (whole (i, all j) segments of equal n), both variants:
//
//   GEMM1  X_u  = W_fd[256u : 256u+256] emb^T     M 256 (channels: 128 per CTA) x N 128 edges, K 768
//                                                 A = W_fd, each CTA loads ITS 256 channels (tensor-map TMA)
//                                                 B = sinusoid embedding, each CTA builds 64 of the 128 rows
//   E1     a1^v = SiLU(X + P^v_i + cg + P^v_j)    thread = (channel, variant); written into the shared
//                                                 memory of CTA v (own, or the peer's through DSMEM)
//   GEMM2  O_o  = W2[256o : 256o+256] [a1^0 | a1^1]^T   M 256 x N 256 = (128 edges, 2 variants), K 512
//                                                 A = W2, each CTA loads its 128 rows per unit
//                                                 B = a1^rank: CTA 0 holds variant 0, CTA 1 variant 1
//   E2     agg^v_i = mean_j SiLU(O + b2)          thread = (channel, variant), in-thread running sums
//
// Per tile and CTA: 96 + 64 MMAs of 64 / 128 cycles = 14.3 k tensor cycles for 256 (edge, variant)
// rows per pair (k_tc_edge: 20.5 k per 128 rows and CTA) and 640 KB of weights streamed from L2
// (k_tc_edge: 1280 KB per 128 rows) -- the L2 -> SM feed was what bounded the one-CTA kernel.
//
// TMEM (512 columns per CTA; lanes = this CTA's channels): X0 = [0,128), X1 = [128,256) (channels
// 256 rank + 128 u + lane), O0 = [256,512), O1 = [0,256) re-uses X once E1 has read it; inside an O unit
// columns [0,128) are the edges of variant 0, [128,256) those of variant 1.
//
//   warps 0-15 : workers.  Embedding group m = warp/2 builds row (warp%2)*32+lane of the chunks with
//                kc % 8 == m; lane quarter q = warp%4 and (unit, variant) = (warp/8, (warp/4)%2) in E1 / E2
//   warps 16-17: MMA issue (lane 0, even CTA only; measured: 1 thread 3.84 ms, 2 threads 3.17 ms, 4 threads 3.21 ms
//                per launch at C3 -- and two fewer warps leave the epilogues 104 registers instead of 96),
//                TMEM alloc (warp 16 of both CTAs), relay of the odd CTA's a1_rx (warp 16, odd CTA)
//   warp 18    : weight loader (lane 0): cp.async.bulk.tensor with cta_group::2, completing on the even
//                CTA's barrier
#include <cuda.h>

#ifdef CB2_EDGE_TIMELINE
namespace cb2 {
__device__ long long g_edge2_dbg[3 * 96];
__device__ int g_edge2_it;
}
#define E1_STAMP(k)                                                                         \
  do {                                                                                      \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_edge2_it >= 1 && g_edge2_it <= 3)          \
      g_edge2_dbg[(g_edge2_it - 1) * 96 + 40 + (k)] = clock64();                            \
  } while (0)
#endif
#include "cb2_tc_edge_epi.cuh"
#include "cb2_tmap.cuh"

namespace cb2 {

using namespace ptx;

constexpr int T2_A1_BYTES = 128 * H * 2;                  // 128 KB: a1 of this CTA's variant, MN-major [64 k8][16 e8][8 k][8 e]
constexpr int T2_ESLOTS = 8;                              // one embedding slot per group
constexpr int T2_E_BYTES = 64 * 32 * 2;                   // 4 KB: [4 k8][64 rows][16 B]
constexpr int T2_WSTAGES = 4;                             // dedicated weight stages (GEMM1 and GEMM2)
constexpr int T2_WALIAS = 8;                              // + stages inside the a1 region, usable by GEMM1 only: a1 is dead
                                                          // from the end of GEMM2 of the previous tile until E1 of this one
constexpr int T2_WRING1 = T2_WSTAGES + T2_WALIAS;         // GEMM1 ring: 12 stages = 3 k tensor cycles of cover for the L2 round trip
constexpr int T2_W_BYTES = 16384;                         // GEMM1: [4 k8][256 ch][16 B]; GEMM2: [8 k8][128 ch][16 B]
constexpr int T2_E_OFF = T2_A1_BYTES;
constexpr int T2_W_OFF = T2_E_OFF + T2_ESLOTS * T2_E_BYTES;
constexpr int T2_BAR_OFF = T2_W_OFF + T2_WSTAGES * T2_W_BYTES;
constexpr int T2_TAB_OFF = T2_BAR_OFF + 512;              // 2 buffers x (off_i[128], off_j[128]) uint32
constexpr int T2_SEG_OFF = T2_TAB_OFF + 2 * 1024;          // 2 buffers x crystal of each segment [32] (tiles with n >= 4)
constexpr int T2_SMEM = T2_SEG_OFF + 2 * 128;
constexpr int T2_NISSUE = 2;                              // MMA-issuing threads; must divide T2_WSTAGES and T2_ESLOTS (see below)
constexpr int T2_WORKERS = 512;
constexpr int T2_THREADS = T2_WORKERS + 32 * (T2_NISSUE + 1);
constexpr int T2_NCH1 = DIS / 32;                         // 24 GEMM1 chunks of K = 32
constexpr int T2_NST2 = 16;                               // GEMM2: 2 output units x 8 stages of K = 64
constexpr int T2_LOADS = T2_NCH1 + T2_NST2;               // 40 weight stages per tile
static_assert(T2_SMEM <= 232448, "shared memory budget");
static_assert(T2_WSTAGES % T2_NISSUE == 0 && T2_WRING1 % T2_NISSUE == 0 && T2_ESLOTS % T2_NISSUE == 0 &&
                  T2_NCH1 % T2_WRING1 == 0 && T2_NST2 % T2_WSTAGES == 0 && T2_WALIAS * T2_W_BYTES <= T2_A1_BYTES,
              "a ring slot must always be waited for by the same issuing thread (mbarrier parity discipline)");
// weight stage st: 0..3 dedicated, 4..11 inside the a1 region
__host__ __device__ constexpr int t2_w_off(int st) {
  return st < T2_WSTAGES ? T2_W_OFF + st * T2_W_BYTES : (st - T2_WSTAGES) * T2_W_BYTES;
}
// how often stage st has been used before use number `u` of tile `it`.  Per tile a dedicated stage
// serves 2 GEMM1 chunks (u = 0, 1) and 4 GEMM2 stages (u = 2..5), an aliased one 2 GEMM1 chunks.
__host__ __device__ constexpr uint32_t t2_w_use(int st, uint32_t it, int u) {
  return st < T2_WSTAGES ? it * (T2_NCH1 / T2_WRING1 + T2_NST2 / T2_WSTAGES) + u : it * (T2_NCH1 / T2_WRING1) + u;
}

// Parity discipline: an mbarrier wait compares one parity bit, so a thread that asks for phase k+1 of
// a barrier whose phase k is still pending sees "completed".  Here every barrier is waited for, phase
// after phase, by the SAME threads: chunk / stage i is issued by thread i % T2_NISSUE, and because
// T2_NISSUE divides the ring sizes a slot never changes hands.

#define T2_WORKER_BARRIER() asm volatile("bar.sync 1, 512;" ::: "memory")

// sin / cos of 2 pi t: exact range reduction to [-1/2, 1/2] turns, then the MUFU approximations (absolute error
// 2^-21 there; the 16-step rotation recurrence that follows amplifies it linearly, far below the fp16 rounding
// of the embedding).  sincospif costs ~100 instructions per call and the producers sit on GEMM1's critical path.
__device__ __forceinline__ void sincos_turns(float t, float &s, float &c) {
  const float a = (t - rintf(t)) * 6.283185307179586f;
  s = __sinf(a);
  c = __cosf(a);
}

// development aid: timeline of cluster 0's even CTA, tiles 1..3 (clock64 stamps), read back by cb2_debug_edge_timeline()
#ifndef CB2_EDGE_TIMELINE
__device__ long long g_edge2_dbg[3 * 96];
#endif
#ifdef CB2_EDGE_TIMELINE
#define T2_STAMP(slot)                                                      \
  do {                                                                      \
    if (blockIdx.x == 0 && it >= 1 && it <= 3) g_edge2_dbg[(it - 1) * 96 + (slot)] = clock64(); \
  } while (0)
#else
#define T2_STAMP(slot) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
    k_tc_edge2(TcEdgeArgs g, const __grid_constant__ CUtensorMap tm_fd, const __grid_constant__ CUtensorMap tm_w2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + T2_BAR_OFF;
  auto a_full = [&](int s) { return bars + 8 * s; };             // 8, even CTA: 64 local + 64 remote arrivals
  auto a_empty = [&](int s) { return bars + 64 + 8 * s; };       // 8, both CTAs (multicast commit)
  auto w_full = [&](int s) { return bars + 128 + 8 * s; };       // 12, even CTA: expect_tx of both CTAs' boxes
  auto w_empty = [&](int s) { return bars + 224 + 8 * s; };      // 12, both CTAs (multicast commit)
  const uint32_t x_full = bars + 320;                            // both: GEMM1 of the tile has completed
  const uint32_t a1_rx = bars + 328;                             // both: this CTA's 512 workers are through E1 (a1 written, X read,
                                                                 //       O1 holds b2) + the 64 KB the peer's workers sent (st.async)
  const uint32_t a1_peer = bars + 336;                           // even: the odd CTA's a1_rx has completed (relayed by one thread)
  const uint32_t o0_ready = bars + 344;                          // even: 2 x 256: O0 drained and re-loaded with b2
  const uint32_t x_free = bars + 352;                            // even: 2 x 256: O1 (= X) drained and cleared
  auto o_full = [&](int o) { return bars + 360 + 8 * o; };       // both: GEMM2 unit o has completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + T2_BAR_OFF + 376);
  uint32_t *tab = reinterpret_cast<uint32_t *>(smem + T2_TAB_OFF);   // [buf][0: off_i, 1: off_j][128], variant-0 offsets
  uint32_t *tab_g = reinterpret_cast<uint32_t *>(smem + T2_SEG_OFF); // [buf][32] crystal of segment s

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1, n_cl = gridDim.x >> 1;
  if (tid == 0) {
    for (int s = 0; s < T2_ESLOTS; s++) { mbar_init(a_full(s), 128); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < T2_WRING1; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(x_full, T2_NISSUE);
    mbar_init(a1_rx, T2_WORKERS);
    mbar_init(a1_peer, 1);
    mbar_init(o0_ready, T2_WORKERS);
    mbar_init(x_free, T2_WORKERS);
    for (int o = 0; o < 2; o++) mbar_init(o_full(o), T2_NISSUE);
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc2(smem_u32(tmem_slot), 512);
    tmem_relinquish2();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // the peer's barriers are initialised before anybody arrives on them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int n_items = g.n_tiles;

  if (warp == 16 + T2_NISSUE) {
    // ------------------------------ weight loader (both CTAs) ------------------------------
    // A single thread needs ~400 cycles per stage (mbarrier test latency + expect_tx + TMA issue), more than
    // GEMM1 leaves it (256): lanes 0 and 1 run the same loop on alternate stages, so the fixed latencies
    // are shared (the ring sizes are even: a stage always belongs to the same lane).
    if (lane < 2) {
      prefetch_tensormap(&tm_fd);
      prefetch_tensormap(&tm_w2);
      const uint32_t wfull0 = mapa_shared(w_full(0), 0);
      uint32_t it = 0;
      for (int item = cl; item < n_items; item += n_cl, it++) {
#pragma unroll
        for (int p = 0; p < T2_NCH1 / 2; p++) {          // GEMM1: chunk i = 2 p + lane, stage i % 12
          const int i = 2 * p + lane;
          const int st = (2 * p) % T2_WRING1 + lane;
          // the aliased stages live in the a1 region: GEMM2 of the previous tile must have read it
          if (2 * p == T2_WSTAGES && it > 0) mbar_wait_spin(o_full(0), (it - 1) & 1);
          if (p == 8 && lane == 0) T2_STAMP(92);
          mbar_wait_spin(w_empty(st), (t2_w_use(2 * p % T2_WRING1, it, 2 * p / T2_WRING1) & 1) ^ 1);
          if (p == 8 && lane == 0) T2_STAMP(93);
          if (leader) mbar_arrive_expect_tx(w_full(st), 2 * T2_W_BYTES);
          tma_load_3d_pair(sbase + t2_w_off(2 * p % T2_WRING1) + lane * T2_W_BYTES, &tm_fd, 0, 8 * (int)rank, 4 * i,
                           wfull0 + 8 * st);
          if (p == 8 && lane == 0) T2_STAMP(94);
        }
#pragma unroll
        for (int p = 0; p < T2_NST2 / 2; p++) {          // GEMM2: stage s2 = 2 p + lane, ring stage s2 % 4; O1 first
          const int s2 = 2 * p + lane;
          const int st = (2 * p) % T2_WSTAGES + lane;
          const int o = 1 - (p >> 2), s = s2 & 7;
          mbar_wait_spin(w_empty(st), (t2_w_use(0, it, T2_NCH1 / T2_WRING1 + 2 * p / T2_WSTAGES) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(w_full(st), 2 * T2_W_BYTES);
          tma_load_3d_pair(sbase + t2_w_off(2 * p % T2_WSTAGES) + lane * T2_W_BYTES, &tm_w2, 0, 0,
                           (2 * (int)rank + o) * 64 + 8 * s, wfull0 + 8 * st);
        }
      }
    }
  } else if (warp >= 16 && warp < 16 + T2_NISSUE) {
    // ------------------------------ MMA issuers (even CTA only) ------------------------------
    if (lane == 0 && leader) {
      const int ii = warp - 16;
      constexpr uint32_t idesc1 = idesc_f16_f32(256, 128);
      constexpr uint32_t idesc2 = idesc_b_mn(idesc_f16_f32(256, 256));
      const uint64_t d_w1 = smem_desc_kmajor(sbase, 4096, 128);    // W_fd stage: [4 k8][256 ch]
      const uint64_t d_e = smem_desc_kmajor(sbase, 1024, 128);     // embedding slot: [4 k8][64 rows]
      const uint64_t d_2k = smem_desc_kmajor(sbase, 2048, 128);    // W2 stage [8 k8][128 ch]; a1 (MN-major)
      uint32_t it = 0;
      for (int item = cl; item < n_items; item += n_cl, it++) {
        if (ii == 0) T2_STAMP(0);
        mbar_wait_spin(x_free, it & 1);
        tc_fence_after_sync();
        if (ii == 0) T2_STAMP(1);
        // GEMM1: X_u += W_fd[u] emb^T  (the units were cleared by the workers).  The barrier probes of this
        // thread's NEXT chunk are fired before the MMAs of the current one (see mbar_test).
        auto g1_pa = [&](int kc) { return (it * 3 + kc / T2_ESLOTS) & 1; };
        auto g1_pw = [&](int kc) { return t2_w_use(kc % T2_WRING1, it, kc / T2_WRING1) & 1; };
        bool ra = mbar_test(a_full(ii % T2_ESLOTS), g1_pa(ii)), rw = mbar_test(w_full(ii % T2_WRING1), g1_pw(ii));
#pragma unroll
        for (int kc = 0; kc < T2_NCH1; kc++) {
          if (kc % T2_NISSUE != ii) continue;
          const int as = kc % T2_ESLOTS, ws = kc % T2_WRING1;
          if (kc == 12) T2_STAMP(88);
          if (!ra) mbar_wait_spin(a_full(as), g1_pa(kc));
          if (kc == 12) T2_STAMP(89);
          if (!rw) mbar_wait_spin(w_full(ws), g1_pw(kc));
          tc_fence_after_sync();
          T2_STAMP(48 + kc);
          const int kn = kc + T2_NISSUE;
          if (kn < T2_NCH1) {
            ra = mbar_test(a_full(kn % T2_ESLOTS), g1_pa(kn));
            rw = mbar_test(w_full(kn % T2_WRING1), g1_pw(kn));
          }
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const uint64_t bd = d_e + (uint64_t)((T2_E_OFF + as * T2_E_BYTES + 2 * j * 1024) >> 4);
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const uint64_t ad = d_w1 + (uint64_t)((t2_w_off(ws) + 2 * j * 4096 + u * 2048) >> 4);
              umma2_f16(tmem + u * 128, ad, bd, idesc1, 1u);
            }
          }
          if (kc == 12) T2_STAMP(90);
          umma2_commit_mc(a_empty(as), (uint16_t)3);
          umma2_commit_mc(w_empty(ws), (uint16_t)3);
          if (kc == 12) T2_STAMP(91);
        }
        umma2_commit_mc(x_full, (uint16_t)3);
        if (ii == 0) T2_STAMP(2);
        // GEMM2: O_o += W2[o] [a1^0 | a1^1]^T  (the units hold b2).  O1 -- the unit that aliases X -- goes first:
        // its E2 and re-initialisation then overlap GEMM2 of O0, and the next tile's GEMM1 starts at once.
        // (Tried: swapping the roles of the two TMEM halves from tile to tile with O0 first, so that GEMM1 never
        // waits for the epilogue of the last unit -- with and without moving the embedding production to the
        // warps that are free then: 3.42 / 3.47 ms per launch at C3 against 3.17 ms, faster only for n = 6.)
        mbar_wait_spin(a1_rx, it & 1);                 // this CTA: workers through E1, the peer's half of a1 has landed
        mbar_wait_spin(a1_peer, it & 1);               // ... and the same in the odd CTA
        if (ii == 0) T2_STAMP(3);
        mbar_wait_spin(o0_ready, it & 1);
        fence_proxy_async_smem();                      // a1 bytes written by st.async -> async proxy (the MMA reads them)
        tc_fence_after_sync();
        if (ii == 0) T2_STAMP(4);
        auto g2_pw = [&](int s2) { return t2_w_use(s2 % T2_WSTAGES, it, T2_NCH1 / T2_WRING1 + s2 / T2_WSTAGES) & 1; };
        rw = mbar_test(w_full(ii % T2_WSTAGES), g2_pw(ii));
#pragma unroll
        for (int s2 = 0; s2 < T2_NST2; s2++) {
          if (s2 % T2_NISSUE != ii) continue;
          const int o = 1 - (s2 >> 3), s = s2 & 7, ws = s2 % T2_WSTAGES;   // O1 (aliases X) first
          if (!rw) mbar_wait_spin(w_full(ws), g2_pw(s2));
          tc_fence_after_sync();
          T2_STAMP(72 + s2);
          if (s2 + T2_NISSUE < T2_NST2) rw = mbar_test(w_full((s2 + T2_NISSUE) % T2_WSTAGES), g2_pw(s2 + T2_NISSUE));
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint64_t ad = d_2k + (uint64_t)((t2_w_off(ws) + 2 * j * 2048) >> 4);
            const uint64_t bd = d_2k + (uint64_t)(((8 * s + 2 * j) * 2048) >> 4);
            umma2_f16(tmem + (o == 0 ? 256 : 0), ad, bd, idesc2, 1u);
          }
          umma2_commit_mc(w_empty(ws), (uint16_t)3);
          if (s >= 8 - T2_NISSUE) umma2_commit_mc(o_full(o), (uint16_t)3);   // this issuer's last stage of unit o
        }
      }
    } else if (lane == 0 && !leader && warp == 16) {
      // odd CTA: relay "my a1_rx has completed" to the even CTA, where the MMAs are issued
      const uint32_t dst = mapa_shared(a1_peer, 0);
      uint32_t it = 0;
      for (int item = cl; item < n_items; item += n_cl, it++) {
        mbar_wait_spin(a1_rx, it & 1);
        fence_proxy_async_smem();
        mbar_arrive_remote(dst);
      }
    }
  } else {
    // ------------------------------ workers (512 threads per CTA) ------------------------------
    const int q = warp & 3, grp = warp >> 2;
    const int e_u = grp >> 1, e_v = grp & 1;        // E1: (X unit, variant); E2: (O unit, variant)
    // embedding group: chunks with kc % 8 == m8.  The O1 warps (8..15) finish a tile first (GEMM2 computes
    // O1 first), so they take the groups whose chunks GEMM1 of the next tile asks for first
    const int m8 = (warp >> 1) ^ 4;
    const int prow = (warp & 1) * 32 + lane;        // row of this CTA's half of the tile built by this thread
    const int trow = 64 * (int)rank + prow;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    // E1
    const uint32_t taddr_x = tq + e_u * 128;
    const int c1 = 256 * (int)rank + 128 * e_u + q * 32 + lane;
    const __half *Pc = g.P + (size_t)e_v * (size_t)g.N * H2 + c1;

    // a1^v lives in CTA v: own shared memory, or the peer's through st.async (completing on the peer's a1_rx)
    const uint32_t a1_off = sbase + (uint32_t)((c1 / 8) * 2048 + (c1 % 8) * 16);
    const A1Dst a1_dst = e_v == (int)rank ? A1Dst{a1_off, 0u}
                                          : A1Dst{mapa_shared(a1_off, rank ^ 1u), mapa_shared(a1_rx, rank ^ 1u)};
    // the O units hold (W2 a1 + b2) / 2: image of W2 / 2, pre-loaded with b2 / 2 (cb2_tc.cuh silu_of_half)
    const float bias_o1 = 0.5f * __ldg(g.b2 + 256 * rank + 128 + q * 32 + lane);     // O1 re-uses the X columns
    // E2
    const uint32_t taddr_o = tq + (e_u == 0 ? 256 : 0) + 128 * e_v;
    const int co = 256 * (int)rank + 128 * e_u + q * 32 + lane;
    const float bias_o = 0.5f * __ldg(g.b2 + co);
    const int oc = g.agg_col + co;
    __half *out = g.agg_kt > 0 ? g.agg16 + (int64_t)(oc >> 3) * 1024 + (oc & 7) : g.agg16 + oc;
    const uint32_t vrow = (uint32_t)e_v * (uint32_t)g.N;
    const AggStride agg_ld = g.agg_kt > 0 ? AggStride{(int64_t)128 * g.agg_kt, 8, vrow}
                                          : AggStride{(int64_t)128 * g.ld_agg, (int)g.ld_agg, vrow};
    // barriers of the even CTA, as seen from this CTA
    const uint32_t a_full_dst = leader ? a_full(m8) : mapa_shared(a_full(m8), 0);
    const uint32_t refill_dst = leader ? (e_u == 0 ? o0_ready : x_free) : mapa_shared(e_u == 0 ? o0_ready : x_free, 0);
    // arrivals on the even CTA's barriers: ordered by the tcgen05 / proxy fences and a CTA-scope release
    // (a cluster-scope release costs ~1.4 k cycles per arrival: measured)
    auto arrive_even = [&](uint32_t dst) {
      if (leader) mbar_arrive(dst); else mbar_arrive_remote(dst);
    };

    int ri_p = -1, rj_p = 0, ri_t = -1, rj_t = 0, rj_prev = -1, n_f = 0;
    auto fetch_rows = [&](int item) {
      n_f = g.seg_n[item];
      ri_p = g.row_i[(int64_t)item * 128 + trow];
      rj_p = g.row_j[(int64_t)item * 128 + trow];
      if (warp < 4) {
        ri_t = g.row_i[(int64_t)item * 128 + tid];
        rj_t = g.row_j[(int64_t)item * 128 + tid];
        // node j of the same position in the previous segment: equal = same crystal, its P_j row is already asked for
        rj_prev = tid >= n_f ? g.row_j[(int64_t)item * 128 + tid - n_f] : -1;
      }
    };
    auto publish_rows = [&](int buf, float (&dl)[3], bool &valid) {
      valid = ri_p >= 0;
      dl[0] = dl[1] = dl[2] = 0.f;
      if (valid) {
#pragma unroll
        for (int d = 0; d < 3; d++) dl[d] = g.x[(int64_t)rj_p * 3 + d] - g.x[(int64_t)ri_p * 3 + d];
      }
      if (warp < 4) {
        uint32_t oi = TE_PAD, oj = (uint32_t)H;
        if (ri_t >= 0) {
          oi = (uint32_t)ri_t * (uint32_t)H2;
          oj = (uint32_t)rj_t * (uint32_t)H2 + (uint32_t)H;
        }
        tab[buf * 256 + tid] = oi;
        tab[buf * 256 + 128 + tid] = oj;
        // crystal of every segment (n >= 4: at most 32 segments): E1 then needs ONE global load for the lattice term
        if (n_f >= 4 && ri_t >= 0 && tid % n_f == 0) tab_g[buf * 32 + tid / n_f] = (uint32_t)g.node2graph[ri_t];
        // pull the tile's rows of P (both variants) towards L2 long before E1 gathers them -- only THIS CTA's
        // 256 channels (4 of the 8 lines of a half row), a row of P_i once per segment, a row of P_j once per
        // crystal in the tile
        const int ri_prev = __shfl_up_sync(0xffffffffu, ri_t, 1);
        if (ri_t >= 0) {
#pragma unroll
          for (int v = 0; v < 2; v++) {
            const char *pv = reinterpret_cast<const char *>(g.P + (size_t)v * (size_t)g.N * H2) + 512 * rank;
            if (lane == 0 || ri_prev != ri_t) {
#pragma unroll
              for (int l = 0; l < 4; l++) prefetch_l2(pv + (size_t)oi * 2 + l * 128);
            }
            if (rj_prev != rj_t) {
#pragma unroll
              for (int l = 0; l < 4; l++) prefetch_l2(pv + (size_t)oj * 2 + l * 128);
            }
          }
        }
      }
    };
    // sinusoid embedding chunk kc = 8 d + m8 of a tile: frequencies 16 m8 .. 16 m8 + 15 of dimension d,
    // columns (sin, cos) interleaved (weights.fd_column_order), by a rotation recurrence
    auto produce = [&](int d, const float (&dl)[3], bool valid, uint32_t tile_it) {
      const uint32_t use = tile_it * 3 + d;
      float s1, c1r, sk, ck;
      sincos_turns(dl[d], s1, c1r);                          // rotation by one frequency step: angle 2 pi dl
      sincos_turns((float)(16 * m8) * dl[d], sk, ck);        // first frequency of this chunk: 16 m8
      mbar_wait(a_empty(m8), (use & 1) ^ 1);
      uint8_t *slot = smem + T2_E_OFF + m8 * T2_E_BYTES + prow * 16;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          w[e] = valid ? pack_half2(sk, ck) : 0u;
          const float sn = fmaf(sk, c1r, ck * s1);
          const float cn = fmaf(ck, c1r, -sk * s1);
          sk = sn; ck = cn;
        }
        *reinterpret_cast<uint4 *>(slot + p * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
      arrive_even(a_full_dst);
    };
    // tcgen05.st of a constant into `ncol` columns, then hand the columns to the MMA side
    auto fill_cols = [&](uint32_t ta, int ncol, uint32_t val) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; j++) z[j] = val;
      for (int cb = 0; cb < ncol; cb += 32) tmem_st32(ta + cb, z);
      tmem_st_wait();
      tc_fence_before_sync();
    };
    auto refill = [&]() {       // O0 threads: b2 for the next GEMM2; O1 threads: zeros for the next GEMM1
      fill_cols(taddr_o, 128, e_u == 0 ? __float_as_uint(bias_o) : 0u);
      arrive_even(refill_dst);
    };

    float dlt[3], dlt_next[3] = {0.f, 0.f, 0.f};
    bool valid = false, valid_next = false;
    uint32_t it = 0;
    // prologue: tables + units of the first tile, its first embedding chunk, row ids of the second
    fetch_rows(cl);
    int n = n_f, n_next = 0;                        // atoms per crystal of this / the next tile (seg_n, fetched ahead)
    publish_rows(0, dlt, valid);
    refill();
    T2_WORKER_BARRIER();
    if (cl + n_cl < n_items) fetch_rows(cl + n_cl);
    produce(0, dlt, valid, 0);
    for (int item = cl; item < n_items; item += n_cl, it++) {
      const int buf = it & 1;
      const int next = item + n_cl;
      const bool has_next = next < n_items;
      if (lane == 0 && q == 0) T2_STAMP(8 + 8 * grp);
#ifdef CB2_EDGE_TIMELINE
      if (tid == 0 && blockIdx.x == 0) g_edge2_it = (int)it;
#endif
      produce(1, dlt, valid, it);
      produce(2, dlt, valid, it);
      if (lane == 0 && q == 0) T2_STAMP(9 + 8 * grp);
      const uint32_t *t_oi = tab + buf * 256;
      const E1Cg cgk{g.cg ? g.cg + c1 : nullptr, g.node2graph, 0u, n >= 4 ? tab_g + buf * 32 : nullptr};
      // ---- E1: a1^v = SiLU(X + P^v_i + cg + P^v_j) for (unit e_u, variant e_v), into CTA e_v ----
      // thread = (channel of unit e_u, variant e_v), all 128 columns; X_u is read by two warps per lane quarter.
      // (A both-variants-per-thread form that reads X once -- half the TMEM reads -- was measured 12 % SLOWER:
      // 3.61 vs 3.23 ms per launch at C3; every thread then feeds a local and a remote store stream.)
      e1_dispatch(n, taddr_x, Pc, cgk, t_oi, t_oi + 128, a1_dst, x_full, it & 1);
      if (lane == 0 && q == 0) T2_STAMP(10 + 8 * grp);
      tc_fence_before_sync();
      asm volatile("bar.sync %0, 256;" ::"r"(2 + e_u) : "memory");
      tc_fence_after_sync();
      // the X columns become O1 columns, pre-loaded with b2
      fill_cols(tq + e_u * 128 + e_v * 64, 64, __float_as_uint(bias_o1));
      fence_proxy_async_smem();                      // own a1 stores -> async proxy (st.async needs no producer fence)
      if (tid == 0) mbar_arrive_expect_tx(a1_rx, 64 * 1024); else mbar_arrive(a1_rx);   // + the peer's 64 KB
      if (lane == 0 && q == 0) T2_STAMP(11 + 8 * grp);
      // Tables of the next tile.  No block-wide barrier is needed: GEMM1 of THIS tile has completed (x_full),
      // hence every warp has produced its chunks of this tile, hence has finished E2 of the previous tile,
      // the last reader of that buffer; the readers of the new entries (E1 of the next tile) are ordered
      // behind them by the next x_full, which needs these warps' chunks of the next tile.
      if (has_next) publish_rows(buf ^ 1, dlt_next, valid_next);
      n_next = n_f;
      if (next + n_cl < n_items) fetch_rows(next + n_cl);
      // the first embedding chunk of the next tile goes out while GEMM2 runs (its slot is free: GEMM1 has completed)
      if (has_next) produce(0, dlt_next, valid_next, it + 1);
      // ---- E2: agg^v_i = mean_j SiLU(O_o); then the unit is re-initialised for the next tile ----
      if (lane == 0 && q == 0) T2_STAMP(12 + 8 * grp);
      mbar_wait(o_full(e_u), it & 1);
      tc_fence_after_sync();
      if (lane == 0 && q == 0) T2_STAMP(13 + 8 * grp);
      e2_dispatch(n, taddr_o, bias_o, t_oi, out, agg_ld);
      tc_fence_before_sync();
      if (lane == 0 && q == 0) T2_STAMP(14 + 8 * grp);
      if (has_next) refill();
      if (lane == 0 && q == 0) T2_STAMP(15 + 8 * grp);
      dlt[0] = dlt_next[0]; dlt[1] = dlt_next[1]; dlt[2] = dlt_next[2];
      valid = valid_next;
      n = n_next;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // nobody leaves while the peer may still write into this CTA
  if (warp == 16) tmem_dealloc2(tmem, 512);
}

int debug_edge2_timeline(long long *out96x3) {
  CB2_CUDA_OK(cudaMemcpyFromSymbol(out96x3, g_edge2_dbg, sizeof(long long) * 3 * 96));
  return CB2_OK;
}

// ---- host side: tensor maps over the weight images -------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_3d(CUtensorMap *tm, const void *base, const uint64_t (&dims_)[3], const uint64_t (&strides_)[2],
                         const uint32_t (&box_)[3]) {
  // resolved once; the driver entry point does not depend on the device (no per-process configuration)
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CB2_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || p == nullptr)
      return fail(CB2_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  const cuuint64_t dims[3] = {dims_[0], dims_[1], dims_[2]};
  const cuuint64_t strides[2] = {strides_[0], strides_[1]};
  const cuuint32_t box[3] = {box_[0], box_[1], box_[2]};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CB2_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return CB2_OK;
}

int encode_weight_map(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t k8, uint32_t box_rows,
                      uint32_t box_k8) {
  // fp16 image [k8][rows][8] seen as [k8][rows / 32][256 halves]: a 512-byte innermost dimension (the
  // largest a box allows) instead of the natural 16-byte one, which would cost one request per row
  return encode_tensor_map_3d(tm, base, {256, rows / 32, k8}, {512, rows * 16}, {256, box_rows / 32, box_k8});
}

int launch_tc_edge2(const TcEdgeArgs &a, int n_sm, cudaStream_t st) {
  if (a.n_tiles == 0) return CB2_OK;
  if (a.V != 2) return fail(CB2_ERR_BAD_ARG, "k_tc_edge2 handles the two CFG variants of a tile together (V == 2)");
  if ((uint64_t)a.V * (uint64_t)a.N * (uint64_t)H2 >= (1ull << 32))
    return fail(CB2_ERR_UNSUPPORTED, "tensor-core edge kernel: V*N*1024 must fit 32 bits (shard the batch)");
  CUtensorMap tm_fd, tm_w2;
  CB2_TRY(encode_weight_map(&tm_fd, a.w_fd_t, H, DIS / 8, 256, 4));     // W_fd image [96][512][8]: box = 256 ch x K 32
  CB2_TRY(encode_weight_map(&tm_w2, a.w2_t, 128, 4 * (H / 8), 128, 8)); // W2 image [4 blocks x 64][128][8]: 128 ch x K 64
  CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_edge2, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));   // per device
  int n_cl = n_sm / 2;
  if (a.n_tiles < n_cl) n_cl = a.n_tiles;
  k_tc_edge2<<<2 * n_cl, T2_THREADS, T2_SMEM, st>>>(a, tm_fd, tm_w2);
  CB2_LAUNCH_OK("k_tc_edge2");
  return CB2_OK;
}

}  // namespace cb2
