// k_tc_node2: the node-level chain BETWEEN two edge kernels as ONE kernel on a CTA pair (cluster of 2,
// tcgen05 cta_group::2, M = 256 rows: 128 per CTA).  Everything between the aggregation of layer l and the
// hoisted edge terms of layer l+1 is row-local (cspnet.py:154-163 node_model, :76-92 FilmLayer.forward,
// :151 layer_norm, :138-145 the h_i / h_j part of edge_mlp.0):
//
//   G1  z   = SiLU([hn | agg] Wn1^T + bn1)                    K 1024, A streamed from cat16 (row-panel)
//   G2  h   = h + SiLU(z Wn2^T + bn2)                         K 512,  A = z   (shared memory)
//   G3  y   = h16 Wp^T + bp ; f = SiLU(LN_film(y) scale[g] + shift[g]) ; h = h + f ; hn = LN_layer(h)
//                                                             K 512,  A = h16 (shared memory)
//   G4  P   = hn [W_hi ; W_hj]^T                              K 512, N 1024, A = hn (shared memory)
//
// The unfused path ran four kernels per layer (k_tc_film, k_tc_linear x 3): z16, h16 and the A operand of
// the hoist GEMM made a round trip through HBM, and every 128-row tile re-streamed its weight block from
// L2 -- both node GEMM kernels were L2-feed-bound at ~2x their floors.  Here the
// activations of a panel never leave the SM between the four GEMMs (one 128 KB buffer X in the canonical
// K-major operand layout, written by the epilogues) and a CTA pair shares every weight byte: each CTA loads
// half of the 256 output channels of an MMA (1.5 MB instead of 3 MB per 128 rows).
//
//   mode HEAD (first layer)  : G3, G4       X <- h16 of the embedding (TMA)
//   mode FULL (layer l | l+1): G1 .. G4
//   mode TAIL (last layer)   : G1, G2       (final LayerNorm and heads follow as separate kernels)
//
// TMEM (512 columns per CTA, lanes = this CTA's 128 rows): G1..G3 accumulate all 512 output columns and are
// drained by the 16 epilogue warps before the next GEMM starts; G4 runs as four units of 256 columns that
// ping-pong between the two halves, so its epilogue (the P store) overlaps its MMAs.
//
//   warp 0      : loaders -- lane 0 weight stages (two 16 KB boxes = 128 channels x K 64 each, 3-stage ring),
//                 lane 1 the A chunks of G1 (into the slots of X, which is dead while G1 runs);
//                 tensor-map TMA with cta_group::2: both CTAs' boxes complete on the even CTA's barrier
//   warp 1      : MMA issue (lane 0, even CTA only), TMEM alloc (both CTAs)
//   warps 2-17  : epilogue, thread = row (TMEM lane), warp = (lane quarter, 128-column group)
#include <cuda.h>

#include "cb2_tc.cuh"
#include "cb2_tmap.cuh"

namespace cb2 {

using namespace ptx;

constexpr int N2_CH_BYTES = 16384;                        // one operand block [8 k8][128 rows][16 B] (K = 64)
constexpr int N2_X_BYTES = 128 * H * 2;                   // 128 KB: [64 k8][128 rows][16 B] = 8 blocks
constexpr int N2_ASLOTS = N2_X_BYTES / N2_CH_BYTES;       // 8
constexpr int N2_WSTAGES = 3;
constexpr int N2_W_BYTES = 2 * N2_CH_BYTES;               // a weight stage = two boxes
constexpr int N2_W_OFF = N2_X_BYTES;
constexpr int N2_BAR_OFF = N2_W_OFF + N2_WSTAGES * N2_W_BYTES;
constexpr int N2_SMEM = N2_BAR_OFF + 512;
constexpr int N2_THREADS = 32 * 18;
static_assert(N2_SMEM <= 232448, "shared memory budget");

// development aid: clock64 timeline of cluster 0's even CTA (MMA thread and epilogue warp 2), panels 1..3 of a
// FULL launch; read back by cb2_debug_node2_timeline()
__device__ long long g_node2_dbg[3 * 64];
#ifdef CB2_NODE_TIMELINE
#define N2_STAMP(itv, slot)                                                                    \
  do {                                                                                         \
    if (blockIdx.x == 0 && g.do_mlp && g.do_film && (itv) >= 1 && (itv) <= 3)                  \
      g_node2_dbg[((itv) - 1) * 64 + (slot)] = clock64();                                      \
  } while (0)
#else
#define N2_STAMP(itv, slot) do { } while (0)
#endif

#ifdef N2_EXP_NO_COND
#define N2_COND(p) make_float4(1.f, 1.f, 1.f, 1.f)
#else
#define N2_COND(p) __ldg(reinterpret_cast<const float4 *>(p))
#endif

struct TcNodeArgs {
  int64_t M;               // V * N rows
  int n_pairs;             // pairs of 128-row panels
  int do_mlp, do_film;     // G1 + G2 | G3 + G4
  const float *bn1, *bn2;  // node MLP biases
  const float *bp;         // FiLM projection bias
  const float *g1, *b1;    // FilmLayer.norm
  const float *g2, *b2;    // CSPLayer.layer_norm (of the layer whose hoist GEMM follows)
  const float *cond;       // [V*B,1024] scale | shift
  const int32_t *node2graph;
  int N, B;
  float *h;                // residual stream, fp32 in the PANEL layout [panel][128 c4][128 rows][4 floats], in place: a
                           // thread owns a row, so its 16-byte pieces are coalesced across the warp (a row-major h
                           // costs one 128-byte line per thread and access: measured 28 k cycles for E2 instead of 8 k)
  float *h_rowmajor;       // mode TAIL: the new h goes HERE, [M,512] row-major, for the final LayerNorm (not to h)
  __half *cat16;           // row-panel, 1024 columns per panel: columns 0:512 <- LN_layer(h)
  __half *P;               // [M,1024] fp16 row-major hoisted edge terms
};

// 32-byte global stores for row-major outputs (P, the final h): a thread owns a row, so consecutive lanes are
// 1-2 KB apart -- a 256-bit access is one whole sector per thread
__device__ __forceinline__ float4 ld_f4_stream(const float *p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Per-column parameters (biases, LayerNorm affine terms: 14 KB, read by every thread for every row): kept in the
// ~24 KB of L1 that 227 KB of shared memory leave, by making them the ONLY global data that allocates there
// (ncu before: L1 hit rate 12 %, long-scoreboard stalls on these loads in every epilogue block)
__device__ __forceinline__ float4 ld_param4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// read-only data that must not displace them (the FiLM row of the thread's crystal: ~100 KB per panel)
__device__ __forceinline__ float4 ld_nc_f4_stream(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_f4_stream(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_u4_stream(void *p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_f8(float *p, const float (&v)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_u8(void *p, const uint32_t (&v)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N2_THREADS, 1)
    k_tc_node2(TcNodeArgs g, const __grid_constant__ CUtensorMap tm_cat, const __grid_constant__ CUtensorMap tm_h16,
               const __grid_constant__ CUtensorMap tm_wn1, const __grid_constant__ CUtensorMap tm_wn2,
               const __grid_constant__ CUtensorMap tm_wp, const __grid_constant__ CUtensorMap tm_whij) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + N2_BAR_OFF;
  auto w_full = [&](int s) { return bars + 8 * s; };              // even CTA: expect_tx of both CTAs' boxes
  auto w_empty = [&](int s) { return bars + 24 + 8 * s; };        // both CTAs (multicast commit)
  auto a_full = [&](int s) { return bars + 48 + 8 * s; };         // even CTA
  auto a_empty = [&](int s) { return bars + 112 + 8 * s; };       // both CTAs
  const uint32_t acc_all = bars + 176;                            // both: all 512 columns of G1 / G2 / G3 are complete
  auto acc_full = [&](int hf) { return bars + 184 + 8 * hf; };    // both: a G4 unit in TMEM half hf is complete
  auto acc_empty = [&](int hf) { return bars + 200 + 8 * hf; };   // even: 2 x 16 warps have drained that half
  const uint32_t x_ready = bars + 216;                            // even: 2 x 16 warps: TMEM drained, X holds the next A operand
  const uint32_t x_dead = bars + 224;                             // both: the last GEMM that reads X has completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + N2_BAR_OFF + 240);

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1, n_cl = gridDim.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < N2_WSTAGES; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < N2_ASLOTS; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    mbar_init(acc_all, 1);
    for (int hf = 0; hf < 2; hf++) { mbar_init(acc_full(hf), 1); mbar_init(acc_empty(hf), 32); }
    mbar_init(x_ready, 32);
    mbar_init(x_dead, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(tmem_slot), 512);
    tmem_relinquish2();
  }
#ifdef N2_STAGGER
  // experiment: start the clusters out of phase so that their memory-heavy and tensor-heavy phases interleave
  if (warp == 0) {
    const long long t0 = clock64(), d = (long long)(cl & 3) * N2_STAGGER;
    while (clock64() - t0 < d) { }
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // the peer's barriers are initialised before anybody arrives on them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ weight loader ------------------------------
      prefetch_tensormap(&tm_wn1);
      prefetch_tensormap(&tm_wn2);
      prefetch_tensormap(&tm_wp);
      prefetch_tensormap(&tm_whij);
      const uint32_t wfull0 = mapa_shared(w_full(0), 0);
      uint32_t cw = 0;
      // one stage: box a = rows [row_a, row_a + 128) x k8 [k8_a, k8_a + 8), box b likewise
      auto load_w = [&](const CUtensorMap *tm, int row_a, int k8_a, int row_b, int k8_b) {
        const int s = (int)(cw % N2_WSTAGES);
        mbar_wait_spin(w_empty(s), ((cw / N2_WSTAGES) & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(w_full(s), 2 * N2_W_BYTES);
        const uint32_t dst = sbase + N2_W_OFF + s * N2_W_BYTES;
        tma_load_3d_pair(dst, tm, 0, row_a / 32, k8_a, wfull0 + 8 * s);
        tma_load_3d_pair(dst + N2_CH_BYTES, tm, 0, row_b / 32, k8_b, wfull0 + 8 * s);
        cw++;
      };
      const int r128 = 128 * (int)rank;
      for (int pp = cl; pp < g.n_pairs; pp += n_cl) {
        if (g.do_mlp) {
          for (int kc = 0; kc < 16; kc++) load_w(&tm_wn1, r128, 8 * kc, 256 + r128, 8 * kc);
          for (int kc = 0; kc < 8; kc++) load_w(&tm_wn2, r128, 8 * kc, 256 + r128, 8 * kc);
        }
        if (g.do_film) {
          for (int kc = 0; kc < 8; kc++) load_w(&tm_wp, r128, 8 * kc, 256 + r128, 8 * kc);
          for (int u = 0; u < 4; u++)
            for (int p = 0; p < 4; p++) load_w(&tm_whij, 256 * u + r128, 16 * p, 256 * u + r128, 16 * p + 8);
        }
      }
    } else if (lane == 1) {
      // ------------------------------ A-chunk loader ------------------------------
      // G1 streams [hn | agg] (16 chunks of K = 64) through the 8 slots of X; without G1 (mode HEAD) the 8
      // chunks of h16 ARE the A operand of G3.  X is free once the last GEMM of the previous panel has read it.
      prefetch_tensormap(&tm_cat);
      prefetch_tensormap(&tm_h16);
      const uint32_t afull0 = mapa_shared(a_full(0), 0);
      const int n_a = g.do_mlp ? 16 : 8;
      const CUtensorMap *tm = g.do_mlp ? &tm_cat : &tm_h16;
      uint32_t ca = 0, it = 0;
      for (int pp = cl; pp < g.n_pairs; pp += n_cl, it++) {
        const int panel = 2 * pp + (int)rank;
        const int base = panel * n_a;                              // chunk index: 16 (cat16) or 8 (h16) per panel
        for (int kc = 0; kc < n_a; kc++, ca++) {
          const int s = (int)(ca % N2_ASLOTS);
          if (kc == 0 && it > 0) mbar_wait_spin(x_dead, (it - 1) & 1);
          mbar_wait_spin(a_empty(s), ((ca / N2_ASLOTS) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(a_full(s), 2 * N2_CH_BYTES);
          tma_load_3d_pair(sbase + s * N2_CH_BYTES, tm, 0, 0, base + kc, afull0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (even CTA only) ------------------------------
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = idesc_f16_f32(256, 256);
      const uint64_t d0 = smem_desc_kmajor(sbase, 2048, 128);      // [8 k8][128 rows][16 B] blocks: LBO 2 KB, SBO 128 B
      uint32_t cw = 0, ca = 0, n_xr = 0;
      // four K = 16 steps of one block pair: D[tcol .. tcol+256) (+)= A[a_off] W[w_off]^T
      auto issue_box = [&](uint32_t a_off, uint32_t w_off, uint32_t tcol, bool first) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          umma2_f16(tmem + tcol, d0 + (uint64_t)((a_off + j * 4096) >> 4), d0 + (uint64_t)((w_off + j * 4096) >> 4), idesc,
                    (first && j == 0) ? 0u : 1u);
      };
      auto wait_w = [&]() -> uint32_t {
        const int s = (int)(cw % N2_WSTAGES);
        mbar_wait_spin(w_full(s), (cw / N2_WSTAGES) & 1);
        return (uint32_t)(N2_W_OFF + s * N2_W_BYTES);
      };
      auto done_w = [&]() {
        umma2_commit_mc(w_empty((int)(cw % N2_WSTAGES)), (uint16_t)3);
        cw++;
      };
      auto wait_x = [&]() {
        mbar_wait_spin(x_ready, n_xr & 1);
        n_xr++;
      };
      // acc_empty(hf) completes twice per panel: units hf (phase parity 0) and hf + 2 (parity 1) have been drained
      auto wait_half = [&](int hf, uint32_t parity) { mbar_wait_spin(acc_empty(hf), parity); };
      // a GEMM with all 512 output columns, K = 64 nk: A chunk kc from X (streamed = through the A ring)
#ifdef CB2_NODE_TIMELINE
      long long tw_a = 0, tw_w = 0, tw_i = 0;      // G1 of a panel: cycles waiting for A, for W, issuing
#define N2_TCLK() clock64()
#else
#define N2_TCLK() 0
#endif
      auto gemm512 = [&](int nk, bool streamed) {
        for (int kc = 0; kc < nk; kc++) {
          uint32_t a_off = (uint32_t)kc * N2_CH_BYTES;
          [[maybe_unused]] const long long c0 = N2_TCLK();
          if (streamed) {
            const int s = (int)(ca % N2_ASLOTS);
            mbar_wait_spin(a_full(s), (ca / N2_ASLOTS) & 1);
            a_off = (uint32_t)s * N2_CH_BYTES;
          }
          [[maybe_unused]] const long long c1 = N2_TCLK();
          const uint32_t w_off = wait_w();
          [[maybe_unused]] const long long c2 = N2_TCLK();
          tc_fence_after_sync();
          issue_box(a_off, w_off, 0, kc == 0);
          issue_box(a_off, w_off + N2_CH_BYTES, 256, kc == 0);
          if (streamed) {
            umma2_commit_mc(a_empty((int)(ca % N2_ASLOTS)), (uint16_t)3);
            ca++;
          }
          done_w();
#ifdef CB2_NODE_TIMELINE
          if (streamed) { tw_a += c1 - c0; tw_w += c2 - c1; tw_i += clock64() - c2; }
#endif
        }
        umma2_commit_mc(acc_all, (uint16_t)3);
      };
      uint32_t it = 0;
      for (int pp = cl; pp < g.n_pairs; pp += n_cl, it++) {
        // TMEM is free once the epilogues of the previous panel are through
        if (it > 0) {
          if (g.do_film) { wait_half(0, 1); wait_half(1, 1); } else wait_x();
          tc_fence_after_sync();
        }
        N2_STAMP(it, 0);
        if (g.do_mlp) {
          gemm512(16, true);                    // G1
          N2_STAMP(it, 1);
#ifdef CB2_NODE_TIMELINE
          if (blockIdx.x == 0 && g.do_film && it >= 1 && it <= 3) {
            g_node2_dbg[(it - 1) * 64 + 40] = tw_a; g_node2_dbg[(it - 1) * 64 + 41] = tw_w; g_node2_dbg[(it - 1) * 64 + 42] = tw_i;
          }
          tw_a = tw_w = tw_i = 0;
#endif
          wait_x();                             // E1: X = z
          N2_STAMP(it, 2);
          tc_fence_after_sync();
          gemm512(8, false);                    // G2
          N2_STAMP(it, 3);
          if (!g.do_film) umma2_commit_mc(x_dead, (uint16_t)3);
        }
        if (g.do_film) {
          if (g.do_mlp) {
            wait_x();                           // E2: X = h16
            tc_fence_after_sync();
          }
          N2_STAMP(it, 4);
          gemm512(8, !g.do_mlp);                // G3 (mode HEAD: X arrives through the A ring)
          N2_STAMP(it, 5);
          wait_x();                             // E3: X = hn
          N2_STAMP(it, 6);
          tc_fence_after_sync();
          for (int u = 0; u < 4; u++) {         // G4: four units of 256 columns, alternating TMEM halves
            const int hf = u & 1;
            if (u >= 2) {
              wait_half(hf, 0);
              tc_fence_after_sync();
            }
            for (int p = 0; p < 4; p++) {
              const uint32_t w_off = wait_w();
              tc_fence_after_sync();
              issue_box((uint32_t)(2 * p) * N2_CH_BYTES, w_off, 256 * hf, p == 0);
              issue_box((uint32_t)(2 * p + 1) * N2_CH_BYTES, w_off + N2_CH_BYTES, 256 * hf, false);
              done_w();
            }
            umma2_commit_mc(acc_full(hf), (uint16_t)3);
            N2_STAMP(it, 7 + u);
          }
          umma2_commit_mc(x_dead, (uint16_t)3);
        }
      }
    }
  } else {
    // ------------------------------ epilogue (16 warps per CTA) ------------------------------
    const int q = warp & 3, cgp = (warp - 2) >> 2;
    const int c0 = cgp * 128;                                      // first of this warp's 128 columns (G1..G3)
    const int row = q * 32 + lane;                                 // row of the panel owned by this thread
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    const int bar_id = 1 + q;                                      // the four warps that share this quarter's rows
    uint8_t *xrow = smem + row * 16;                               // X[k8][row]: + k8 * 2048
    // exchange slot of the LayerNorm partial sums: the last 16-byte cell this thread itself writes in pass 3
    float *slot = reinterpret_cast<float *>(xrow + (cgp * 16 + 15) * 2048);
    const uint32_t x_ready_dst = leader ? x_ready : mapa_shared(x_ready, 0);
    auto signal = [&](uint32_t dst_even, bool wrote_x) {
      tc_fence_before_sync();
      if (wrote_x) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(dst_even); else mbar_arrive_remote(dst_even);
      }
    };
    uint32_t n_all = 0;
#ifdef CB2_NODE_TIMELINE
    uint32_t eit = 0;
#define N2_ESTAMP(slot) do { if (warp == 2 && lane == 0) N2_STAMP(eit, slot); } while (0)
#else
#define N2_ESTAMP(slot) do { } while (0)
#endif
    for (int pp = cl; pp < g.n_pairs; pp += n_cl) {
      const int panel = 2 * pp + (int)rank;
      const int64_t grow = (int64_t)panel * 128 + row;
      const bool valid = grow < g.M;
      const bool pvalid = (int64_t)panel * 128 < g.M;              // false: the phantom second panel of an odd count
      // panel layout: cell (c4, row) = 16 bytes at ((c4 * 128) + row) * 16; this thread: c4 = c0 / 4 + ...
      float *hrow = g.h + (int64_t)panel * (128 * H) + (c0 >> 2) * 512 + row * 4;
      const float *cs = g.cond;
      if (pvalid) {
        // pull this warp's piece of the residual stream (32 c4 blocks of 512 bytes) towards L2
        const char *hw = reinterpret_cast<const char *>(g.h + (int64_t)panel * (128 * H) + (c0 >> 2) * 512 + q * 128);
#pragma unroll
        for (int l = 0; l < 4; l++) prefetch_l2(hw + ((lane + 32 * l) >> 2) * 2048 + ((lane + 32 * l) & 3) * 128);
      }
      if (valid) {
        // ... and the FiLM row of this thread's crystal
        if (g.do_film) {
          cs = g.cond + ((grow / g.N) * g.B + g.node2graph[(int)(grow % g.N)]) * H2;
#pragma unroll
          for (int l = 0; l < 4; l++) {
            prefetch_l2(reinterpret_cast<const char *>(cs + c0) + l * 128);
            prefetch_l2(reinterpret_cast<const char *>(cs + H + c0) + l * 128);
          }
        }
      }
      const uint32_t taddr = tq + c0;
      if (g.do_mlp) {
        // ---- E1: z = SiLU(acc + bn1) -> fp16 -> X ----
        mbar_wait(acc_all, n_all & 1);
        n_all++;
        tc_fence_after_sync();
        N2_ESTAMP(16);
        {
          uint32_t accA[16], accB[16];
          tmem_ld16(taddr, accA);
#pragma unroll
          for (int hb = 0; hb < 8; hb++) {
            tmem_ld_wait();
            uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
            uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
            if (hb < 7) tmem_ld16(taddr + (hb + 1) * 16, nxt);
            const int c = c0 + hb * 16;
            uint32_t w[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 b = ld_param4(g.bn1 + c + 4 * j4);
              w[2 * j4] = pack_half2(silu_fast(__uint_as_float(acc[4 * j4]) + b.x), silu_fast(__uint_as_float(acc[4 * j4 + 1]) + b.y));
              w[2 * j4 + 1] = pack_half2(silu_fast(__uint_as_float(acc[4 * j4 + 2]) + b.z), silu_fast(__uint_as_float(acc[4 * j4 + 3]) + b.w));
            }
            *reinterpret_cast<uint4 *>(xrow + (c >> 3) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(xrow + ((c >> 3) + 1) * 2048) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        signal(x_ready_dst, true);
        N2_ESTAMP(17);
        // ---- E2: h = h + SiLU(acc + bn2) -> fp32 (global) and fp16 -> X ----
        float hv[16], hn_[16];
        auto load_h = [&](int blk, float (&dst)[16]) {
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
#ifdef N2_EXP_NO_E2_LOAD
            const float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#else
            const float4 v = pvalid ? ld_f4_stream(hrow + (blk * 4 + j4) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
            dst[4 * j4] = v.x; dst[4 * j4 + 1] = v.y; dst[4 * j4 + 2] = v.z; dst[4 * j4 + 3] = v.w;
          }
        };
        load_h(0, hn_);
        mbar_wait(acc_all, n_all & 1);
        n_all++;
        tc_fence_after_sync();
        N2_ESTAMP(18);
#pragma unroll 1
        for (int hb = 0; hb < 8; hb++) {
          uint32_t acc[16];
          tmem_ld16(taddr + hb * 16, acc);
#pragma unroll
          for (int j = 0; j < 16; j++) hv[j] = hn_[j];
          if (hb < 7) load_h(hb + 1, hn_);
          const int c = c0 + hb * 16;
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = ld_param4(g.bn2 + c + 4 * j4);
            o[4 * j4] = hv[4 * j4] + silu_fast(__uint_as_float(acc[4 * j4]) + b.x);
            o[4 * j4 + 1] = hv[4 * j4 + 1] + silu_fast(__uint_as_float(acc[4 * j4 + 1]) + b.y);
            o[4 * j4 + 2] = hv[4 * j4 + 2] + silu_fast(__uint_as_float(acc[4 * j4 + 2]) + b.z);
            o[4 * j4 + 3] = hv[4 * j4 + 3] + silu_fast(__uint_as_float(acc[4 * j4 + 3]) + b.w);
          }
          if (g.h_rowmajor != nullptr) {          // mode TAIL: row-major, for the final LayerNorm
            if (valid) {
              float a[8], b[8];
#pragma unroll
              for (int j = 0; j < 8; j++) { a[j] = o[j]; b[j] = o[8 + j]; }
              st_f8(g.h_rowmajor + grow * H + c + 0, a);
              st_f8(g.h_rowmajor + grow * H + c + 8, b);
            }
          } else if (pvalid) {
#ifndef N2_EXP_NO_E2_STORE
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++)
              st_f4_stream(hrow + (hb * 4 + j4) * 512, make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]));
#endif
          }
          if (g.do_film) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = pack_half2(o[2 * j], o[2 * j + 1]);
            *reinterpret_cast<uint4 *>(xrow + (c >> 3) * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(xrow + ((c >> 3) + 1) * 2048) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        signal(x_ready_dst, g.do_film != 0);
        N2_ESTAMP(19);
      }
      if (g.do_film) {
        // ---- E3: FiLM + residual + layer LayerNorm on y = acc + bp (three passes over TMEM) ----
        mbar_wait(acc_all, n_all & 1);
        n_all++;
        tc_fence_after_sync();
        N2_ESTAMP(20);
        float s = 0.f, ss = 0.f;
        {
          uint32_t accA[16], accB[16];
          tmem_ld16(taddr, accA);
#pragma unroll
          for (int hb = 0; hb < 8; hb++) {
            tmem_ld_wait();
            uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
            uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
            if (hb < 7) tmem_ld16(taddr + (hb + 1) * 16, nxt);
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 b = ld_param4(g.bp + c0 + hb * 16 + 4 * j4);
              const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const float y = __uint_as_float(acc[4 * j4 + k]) + bb[k];
                s += y;
                ss = fmaf(y, y, ss);
              }
            }
          }
        }
        // X is dead here (G3 has read it): the exchange cells live inside it
        slot[0] = s;
        slot[1] = ss;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        float mean1, rstd1;
        {
          float ts = 0.f, tss = 0.f;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float2 v = *reinterpret_cast<const float2 *>(xrow + (k * 16 + 15) * 2048);
            ts += v.x;
            tss += v.y;
          }
          mean1 = ts * (1.0f / H);
          rstd1 = rsqrtf(fmaxf(tss * (1.0f / H) - mean1 * mean1, 0.f) + 1e-5f);
        }
        N2_ESTAMP(21);
        // pass 2: f = SiLU(LN(y) scale + shift), h += f (global, and kept in TMEM), statistics of the new h
        s = 0.f;
        ss = 0.f;
        {
          float hv[16], hn_[16];
          auto load_h = [&](int blk, float (&dst)[16]) {
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 v = pvalid ? ld_f4_stream(hrow + (blk * 4 + j4) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
              dst[4 * j4] = v.x; dst[4 * j4 + 1] = v.y; dst[4 * j4 + 2] = v.z; dst[4 * j4 + 3] = v.w;
            }
          };
          load_h(0, hn_);
#pragma unroll 1
          for (int blk = 0; blk < 8; blk++) {
            const int col0 = c0 + blk * 16;
            uint32_t acc[16];
            tmem_ld16(taddr + blk * 16, acc);
#pragma unroll
            for (int j = 0; j < 16; j++) hv[j] = hn_[j];
            if (blk < 7) load_h(blk + 1, hn_);
            float4 sc[4], sh[4];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              sc[j4] = N2_COND(cs + col0 + 4 * j4);
              sh[j4] = N2_COND(cs + H + col0 + 4 * j4);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 pb = ld_param4(g.bp + col0 + 4 * j4);
              const float4 pg = ld_param4(g.g1 + col0 + 4 * j4);
              const float4 pbb = ld_param4(g.b1 + col0 + 4 * j4);
              const float bps[4] = {pb.x, pb.y, pb.z, pb.w}, g1s[4] = {pg.x, pg.y, pg.z, pg.w};
              const float b1s[4] = {pbb.x, pbb.y, pbb.z, pbb.w};
              const float scs[4] = {sc[j4].x, sc[j4].y, sc[j4].z, sc[j4].w};
              const float shs[4] = {sh[j4].x, sh[j4].y, sh[j4].z, sh[j4].w};
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int j = 4 * j4 + k;
                const float y = __uint_as_float(acc[j]) + bps[k];
                const float ln = fmaf((y - mean1) * rstd1, g1s[k], b1s[k]);
                const float hn = hv[j] + silu_fast(fmaf(ln, scs[k], shs[k]));
                s += hn;
                ss = fmaf(hn, hn, ss);
                acc[j] = __float_as_uint(hn);
              }
            }
            tmem_st16(taddr + blk * 16, acc);
#ifndef N2_EXP_NO_P2_STORE
            if (pvalid) {
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++)
                st_f4_stream(hrow + (blk * 4 + j4) * 512, make_float4(__uint_as_float(acc[4 * j4]), __uint_as_float(acc[4 * j4 + 1]),
                                                                      __uint_as_float(acc[4 * j4 + 2]), __uint_as_float(acc[4 * j4 + 3])));
            }
#endif
          }
        }
        tmem_st_wait();
        N2_ESTAMP(22);
        slot[2] = s;
        slot[3] = ss;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        float mean2, rstd2;
        {
          float ts = 0.f, tss = 0.f;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float2 v = *reinterpret_cast<const float2 *>(xrow + (k * 16 + 15) * 2048 + 8);
            ts += v.x;
            tss += v.y;
          }
          mean2 = ts * (1.0f / H);
          rstd2 = rsqrtf(fmaxf(tss * (1.0f / H) - mean2 * mean2, 0.f) + 1e-5f);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // all four warps have read the cells: pass 3 may overwrite them
        // pass 3: hn = LN_layer(h) -> fp16 -> X (A operand of G4) and the row-panel cat16 (A operand of the next G1)
        {
          __half *dst = g.cat16 + (int64_t)panel * 128 * H2 + (int64_t)(c0 >> 3) * 1024 + row * 8;
          uint32_t accA[16], accB[16];
          tmem_ld16(taddr, accA);
#pragma unroll
          for (int hb = 0; hb < 8; hb++) {
            tmem_ld_wait();
            uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
            uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
            if (hb < 7) tmem_ld16(taddr + (hb + 1) * 16, nxt);
            const int c = c0 + hb * 16;
            uint32_t w[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 pg = ld_param4(g.g2 + c + 4 * j4);
              const float4 pb = ld_param4(g.b2 + c + 4 * j4);
              const float a0 = fmaf((__uint_as_float(acc[4 * j4]) - mean2) * rstd2, pg.x, pb.x);
              const float a1 = fmaf((__uint_as_float(acc[4 * j4 + 1]) - mean2) * rstd2, pg.y, pb.y);
              const float a2 = fmaf((__uint_as_float(acc[4 * j4 + 2]) - mean2) * rstd2, pg.z, pb.z);
              const float a3 = fmaf((__uint_as_float(acc[4 * j4 + 3]) - mean2) * rstd2, pg.w, pb.w);
              w[2 * j4] = pack_half2(a0, a1);
              w[2 * j4 + 1] = pack_half2(a2, a3);
            }
            const uint4 lo = make_uint4(w[0], w[1], w[2], w[3]), hi = make_uint4(w[4], w[5], w[6], w[7]);
            *reinterpret_cast<uint4 *>(xrow + (c >> 3) * 2048) = lo;
            *reinterpret_cast<uint4 *>(xrow + ((c >> 3) + 1) * 2048) = hi;
            if (valid) {
              st_u4_stream(dst + (hb * 2) * 1024, lo);
              st_u4_stream(dst + (hb * 2 + 1) * 1024, hi);
            }
          }
        }
        signal(x_ready_dst, true);
        N2_ESTAMP(23);
        // ---- E4: the four units of P = hn [W_hi ; W_hj]^T -> fp16 row-major; this warp: 64 columns per unit ----
        __half *prow = g.P + grow * H2 + cgp * 64;
#pragma unroll 1
        for (int u = 0; u < 4; u++) {
          const int hf = u & 1;
          mbar_wait(acc_full(hf), (uint32_t)(u >> 1));       // two completions per panel and half: units hf, hf + 2
          tc_fence_after_sync();
          N2_ESTAMP(24 + 2 * u);
          const uint32_t ta = tq + 256 * hf + cgp * 64;
          uint32_t accA[16], accB[16];
          tmem_ld16(ta, accA);
#pragma unroll
          for (int hb = 0; hb < 4; hb++) {
            tmem_ld_wait();
            uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
            uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
            if (hb < 3) tmem_ld16(ta + (hb + 1) * 16, nxt);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = pack_half2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
            if (valid) st_u8(prow + u * 256 + hb * 16, w);
          }
          signal(leader ? acc_empty(hf) : mapa_shared(acc_empty(hf), 0), false);
          N2_ESTAMP(25 + 2 * u);
        }
      }
#ifdef CB2_NODE_TIMELINE
      eit++;
#endif
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // nobody leaves while the peer's MMAs / TMA may still touch this CTA
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

int debug_node2_timeline(long long *out64x3) {
  CB2_CUDA_OK(cudaMemcpyFromSymbol(out64x3, g_node2_dbg, sizeof(long long) * 3 * 64));
  return CB2_OK;
}

// a row-panel activation buffer as a sequence of 16 KB chunks (128 rows x K 64, already in operand order):
// [chunk][32 rows of 512 bytes]; coordinates (0, 0, chunk)
static int encode_chunk_map(CUtensorMap *tm, const void *base, uint64_t bytes) {
  return encode_tensor_map_3d(tm, base, {256, 32, bytes / N2_CH_BYTES}, {512, N2_CH_BYTES}, {256, 32, 1});
}

int launch_tc_node2(const cb2_model *m, const cb2_layer_weights *Lmlp, const cb2_layer_weights *Lfilm,
                    const cb2_batch *b, const float *film_cond, float *h, float *h_rowmajor, const __half *h16,
                    __half *cat16, __half *P, int n_sm, cudaStream_t st) {
  const int64_t VN = (int64_t)b->n_variants * b->n_nodes;
  if (VN == 0) return CB2_OK;
  if (!Lmlp && !Lfilm) return fail(CB2_ERR_BAD_ARG, "tc_node2: nothing to do");
  const int64_t n_panels = (VN + 127) / 128;
  TcNodeArgs a{};
  a.M = VN; a.n_pairs = (int)((n_panels + 1) / 2);
  a.do_mlp = Lmlp != nullptr; a.do_film = Lfilm != nullptr;
  a.N = b->n_nodes; a.B = b->n_graphs; a.node2graph = b->node2graph; a.cond = film_cond;
  a.h = h; a.h_rowmajor = Lfilm ? nullptr : h_rowmajor; a.cat16 = cat16; a.P = P;
  if (!Lfilm && !h_rowmajor) return fail(CB2_ERR_BAD_ARG, "tc_node2: mode TAIL needs the row-major output");
  // a map the launch does not use still has to be a valid descriptor: it points at the FiLM projection image
  CUtensorMap tm_cat, tm_h16, tm_wn1, tm_wn2, tm_wp, tm_whij;
  CB2_TRY(encode_weight_map(&tm_wp, m->film_wp_t, H, H / 8, 128, 8));
  tm_wn1 = tm_wn2 = tm_whij = tm_cat = tm_h16 = tm_wp;
  if (Lmlp) {
    a.bn1 = Lmlp->bn1; a.bn2 = Lmlp->bn2;
    CB2_TRY(encode_weight_map(&tm_wn1, Lmlp->wn1_t, H, H2 / 8, 128, 8));
    CB2_TRY(encode_weight_map(&tm_wn2, Lmlp->wn2_t, H, H / 8, 128, 8));
    CB2_TRY(encode_chunk_map(&tm_cat, cat16, (uint64_t)n_panels * 128 * H2 * 2));
  } else {
    CB2_TRY(encode_chunk_map(&tm_h16, h16, (uint64_t)n_panels * 128 * H * 2));
  }
  if (Lfilm) {
    if (!film_cond) return fail(CB2_ERR_BAD_ARG, "tc_node2: FiLM stage without film_cond");
    a.bp = m->film_bp; a.g1 = m->film_g; a.b1 = m->film_b; a.g2 = Lfilm->ln_g; a.b2 = Lfilm->ln_b;
    CB2_TRY(encode_weight_map(&tm_whij, Lfilm->w_hij_t, H2, H / 8, 128, 8));
  }
  CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_node2, cudaFuncAttributeMaxDynamicSharedMemorySize, N2_SMEM));   // per device
  int n_cl = n_sm / 2;
  if (a.n_pairs < n_cl) n_cl = a.n_pairs;
  k_tc_node2<<<2 * n_cl, N2_THREADS, N2_SMEM, st>>>(a, tm_cat, tm_h16, tm_wn1, tm_wn2, tm_wp, tm_whij);
  CB2_LAUNCH_OK("k_tc_node2");
  return CB2_OK;
}

}  // namespace cb2
