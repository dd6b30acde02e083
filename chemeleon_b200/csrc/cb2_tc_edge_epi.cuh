// Epilogue bodies shared by the two tensor-core edge kernels (k_tc_edge: one CTA per work item;
// k_tc_edge2: CTA pair, cta_group::2, sinusoid GEMM shared by the two CFG variants).  TMEM lanes =
// output channels, TMEM columns = edge rows of the tile, so a thread owns one channel and walks over
// the tile's edges: E1 turns the GEMM1 accumulator into the fp16 GEMM2 operand, E2 takes the
// segmented mean (scatter_mean of cspnet.py:155-160) as an in-thread running sum.
#pragma once

#include "cb2_tc.cuh"

namespace cb2 {

using namespace ptx;

constexpr uint32_t TE_PAD = 0xFFFFFFFFu;        // off_i of a padding row

#ifndef E1_STAMP
#define E1_STAMP(k) do { } while (0)
#endif

#ifdef CB2_SPIN_WORKERS
#define MBAR_WAIT_WORKER mbar_wait_spin
#else
#define MBAR_WAIT_WORKER mbar_wait
#endif

// ---- E2 body: mean over the n edges of each segment of SiLU(U + b2) for one TMEM unit ----
// Segment boundaries are compile-time for N > 0 (no branches in the running sum); N == 0 is
// the generic runtime-n version used for segment lengths without a specialisation.
// Output row r of channel c lives at out[(r / 128) * ld_agg.hi + (r % 128) * ld_agg.lo]: row-major
// (hi = 128 ld, lo = ld) or the row-panel layout of the GEMM A operands (hi = 128 kt, lo = 8).
struct AggStride { int64_t hi; int lo; uint32_t row0; };   // row0: added to the table's node row (variant base)
__device__ __forceinline__ void e2_store(const uint32_t *t_oi, int seg0, __half *out, AggStride ld_agg, float mean) {
  const uint32_t o = t_oi[seg0];
  if (o != TE_PAD) {
    const uint32_t r = (o >> 10) + ld_agg.row0;
    out[(int64_t)(r >> 7) * ld_agg.hi + (int)(r & 127) * ld_agg.lo] = __float2half_rn(fminf(fmaxf(mean, -65504.f), 65504.f));
  }
}

template <int N>
__device__ __noinline__ void e2_unit(int n_rt, uint32_t taddr, float bias, const uint32_t *t_oi, __half *out,
                                     AggStride ld_agg) {
  const int n = N > 0 ? N : n_rt;
  const float inv_n = 1.0f / (float)n;
  float sum = 0.f;
  int cnt = 0, seg0 = 0;
  uint32_t accA[32], accB[32];
  tmem_ld32(taddr, accA);
#pragma unroll
  for (int cb = 0; cb < 4; cb++) {
    tmem_ld_wait();
    uint32_t (&acc)[32] = (cb & 1) ? accB : accA;
    uint32_t (&nxt)[32] = (cb & 1) ? accA : accB;
    if (cb < 3) tmem_ld32(taddr + (cb + 1) * 32, nxt);   // next chunk's TMEM read overlaps this chunk's math
    float t[32];
#pragma unroll
    // fp32 tanh here (one MUFU op per element): the paired fp16 form (tanh.approx.f16x2, one MUFU op per two
    // SiLUs) was measured slower in both edge kernels (more issue slots: E2 is latency / issue bound) at 1.5x
    // the error, and the mean is better taken over unrounded values
    for (int j = 0; j < 32; j++) t[j] = silu_of_half(__uint_as_float(acc[j]));   // the accumulator holds (W2 a1 + b2) / 2
#pragma unroll
    for (int j = 0; j < 32; j++) {
      sum += t[j];
      if (N > 0) {
        if ((cb * 32 + j + 1) % (N > 0 ? N : 1) == 0) {
          e2_store(t_oi, cb * 32 + j + 1 - N, out, ld_agg, sum * inv_n);
          sum = 0.f;
        }
      } else if (++cnt == n) {
        e2_store(t_oi, seg0, out, ld_agg, sum * inv_n);
        sum = 0.f;
        cnt = 0;
        seg0 += n;
      }
    }
  }
}

__device__ __forceinline__ void e2_dispatch(int n, uint32_t taddr, float bias, const uint32_t *t_oi, __half *out,
                                            AggStride ld_agg) {
  switch (n) {
#define CB2_E2_CASE(N) case N: e2_unit<N>(n, taddr, bias, t_oi, out, ld_agg); break;
    CB2_E2_CASE(1) CB2_E2_CASE(2) CB2_E2_CASE(3) CB2_E2_CASE(4) CB2_E2_CASE(5) CB2_E2_CASE(6) CB2_E2_CASE(7)
    CB2_E2_CASE(8) CB2_E2_CASE(9) CB2_E2_CASE(10) CB2_E2_CASE(11) CB2_E2_CASE(12) CB2_E2_CASE(13) CB2_E2_CASE(14)
    CB2_E2_CASE(15) CB2_E2_CASE(16) CB2_E2_CASE(17) CB2_E2_CASE(18) CB2_E2_CASE(19) CB2_E2_CASE(20) CB2_E2_CASE(21)
    CB2_E2_CASE(22) CB2_E2_CASE(23) CB2_E2_CASE(24) CB2_E2_CASE(25) CB2_E2_CASE(26) CB2_E2_CASE(27) CB2_E2_CASE(28)
    CB2_E2_CASE(29) CB2_E2_CASE(30) CB2_E2_CASE(31) CB2_E2_CASE(32) CB2_E2_CASE(33) CB2_E2_CASE(34) CB2_E2_CASE(35)
    CB2_E2_CASE(36) CB2_E2_CASE(37) CB2_E2_CASE(38) CB2_E2_CASE(39) CB2_E2_CASE(40)
#undef CB2_E2_CASE
    default: e2_unit<0>(n, taddr, bias, t_oi, out, ld_agg); break;
  }
}

// ---- E1 body: a1[c][e] = SiLU(U[c][e] + P_i[i(e)][c] + P_j[j(e)][c]) for one lane quarter of a unit ----
// The hoisted node terms are gathered thread = channel (128 B per warp and node) BEFORE the wait
// on the GEMM1 accumulator, so their latency is off the critical path.  N > 0: the tile's segment
// structure is compile-time, each thread loads P_i once per segment and the N rows of P_j once
// per crystal (they repeat for every segment of the same crystal).  N == 0: generic version, two
// gathers per edge.
// The per-crystal lattice term cg[g][c] (fp32, O(10..100) for Angstrom-scale cells) is added here in
// fp32 -- it is never folded into the fp16 P rows, whose O(1) node signal it would swamp.
// cgc = cg + channel (NULL: no term).  The crystal of a segment comes from `seg_g` (per-segment table in
// shared memory, filled a tile ahead: one global load on the E1 path) or, without it, from node2graph.
struct E1Cg { const float *cgc; const int32_t *n2g; uint32_t vbase; const uint32_t *seg_g; };
__device__ __forceinline__ float e1_cg(const E1Cg &k, uint32_t oi) {
  return k.cgc ? __ldg(k.cgc + (size_t)__ldg(k.n2g + ((oi >> 10) - k.vbase)) * H) : 0.f;
}

// Destination of a thread's a1 values: a shared-window address, either in this CTA (rbar == 0: plain
// st.shared) or in the PEER CTA of a pair (rbar = the peer's receive mbarrier: st.async, whose bytes
// complete on that barrier -- the producer needs no fence and no cluster-scope release).
struct A1Dst { uint32_t addr; uint32_t rbar; };
__device__ __forceinline__ void e1_store(const A1Dst &d, uint32_t off, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  if (d.rbar != 0u) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     d.addr + off),
                 "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(d.rbar)
                 : "memory");
  } else {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d.addr + off), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                 : "memory");
  }
}

template <int N>
__device__ __noinline__ void e1_unit(uint32_t taddr, const __half *Pc, E1Cg cgk, const uint32_t *t_oi, const uint32_t *t_oj,
                                     A1Dst a1_dst, uint32_t acc1_full, uint32_t parity) {
  if constexpr (N == 0) {
    const uint4 *ti = reinterpret_cast<const uint4 *>(t_oi);
    const uint4 *tj = reinterpret_cast<const uint4 *>(t_oj);
    MBAR_WAIT_WORKER(acc1_full, parity);
    tc_fence_after_sync();
#pragma unroll 1
    for (int cb = 0; cb < 4; cb++) {
      uint32_t acc[32];
      tmem_ld32(taddr + cb * 32, acc);
      float pv[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const uint4 oi = ti[cb * 8 + j4], oj = tj[cb * 8 + j4];
        const uint32_t ois[4] = {oi.x, oi.y, oi.z, oi.w};
        const uint32_t ojs[4] = {oj.x, oj.y, oj.z, oj.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
          pv[j4 * 4 + k] = ois[k] == TE_PAD ? 0.f : __half2float(Pc[ois[k]]) + e1_cg(cgk, ois[k]) + __half2float(Pc[ojs[k]]);
      }
      tmem_ld_wait();
#pragma unroll
      for (int p = 0; p < 4; p++) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++)
          w[e] = silu2_half(__uint_as_float(acc[8 * p + 2 * e]) + pv[8 * p + 2 * e],
                            __uint_as_float(acc[8 * p + 2 * e + 1]) + pv[8 * p + 2 * e + 1]);
        e1_store(a1_dst, (cb * 4 + p) * 128, w[0], w[1], w[2], w[3]);
      }
    }
  } else {
    constexpr int S = 128 / N;
    // Stage 1: every gather of the tile is issued back to back as RAW values (no arithmetic that would
    // make the in-order issue wait for the first load); they are converted after the wait on GEMM1.
    uint32_t oiv[S];
    __half pi_raw[S], pj_raw[N];
    float cgv[S];
#pragma unroll
    for (int sgm = 0; sgm < S; sgm++) oiv[sgm] = t_oi[sgm * N];
    uint32_t cur = t_oj[0];
#pragma unroll
    for (int sgm = 0; sgm < S; sgm++) {
      const uint32_t oi = oiv[sgm];
      pi_raw[sgm] = Pc[oi == TE_PAD ? 0u : oi];
      cgv[sgm] = 0.f;
      if (cgk.cgc != nullptr && oi != TE_PAD) {
        const uint32_t gidx = cgk.seg_g ? cgk.seg_g[sgm] : (uint32_t)__ldg(cgk.n2g + ((oi >> 10) - cgk.vbase));
        cgv[sgm] = __ldg(cgk.cgc + (size_t)gidx * H);
      }
    }
#pragma unroll
    for (int k = 0; k < N; k++) pj_raw[k] = Pc[cur + (uint32_t)k * (uint32_t)H2];
    E1_STAMP(0);
    MBAR_WAIT_WORKER(acc1_full, parity);
    tc_fence_after_sync();
    E1_STAMP(1);
    float piv[S], pj[N];
#pragma unroll
    for (int sgm = 0; sgm < S; sgm++) piv[sgm] = oiv[sgm] == TE_PAD ? 0.f : __half2float(pi_raw[sgm]) + cgv[sgm];
#pragma unroll
    for (int k = 0; k < N; k++) pj[k] = __half2float(pj_raw[k]);
    // 16-column TMEM loads, double-buffered: the load of block hb+1 is in flight while block hb
    // goes through the SiLU (register budget: 2 x 16 accumulators + P_i / P_j values)
    uint32_t accA[16], accB[16];
    tmem_ld16(taddr, accA);
    E1_STAMP(4);
#pragma unroll
    for (int hb = 0; hb < 8; hb++) {
      tmem_ld_wait();
      if (hb == 0) E1_STAMP(5);
      uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
      uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
      if (hb < 7) tmem_ld16(taddr + (hb + 1) * 16, nxt);
      float x[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int e = hb * 16 + j;
        x[j] = __uint_as_float(acc[j]);
        if (e < S * N) {
          if (e % N == 0 && e > 0) {            // segment start: same crystal as before?
            const uint32_t oj0 = t_oj[e];
            if (oj0 != cur) {
              cur = oj0;
#pragma unroll
              for (int k = 0; k < N; k++) pj[k] = __half2float(Pc[cur + (uint32_t)k * (uint32_t)H2]);
            }
          }
          x[j] += piv[e / N] + pj[e % N];
        }
      }
#pragma unroll
      for (int p = 0; p < 2; p++) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) w[e] = silu2_half(x[8 * p + 2 * e], x[8 * p + 2 * e + 1]);
        e1_store(a1_dst, (hb * 2 + p) * 128, w[0], w[1], w[2], w[3]);
      }
      if (hb == 0) E1_STAMP(2);
      if (hb == 7) E1_STAMP(3);
    }
  }
}

__device__ __forceinline__ void e1_dispatch(int n, uint32_t taddr, const __half *Pc, E1Cg cgk, const uint32_t *t_oi,
                                            const uint32_t *t_oj, A1Dst a1_dst, uint32_t acc1_full,
                                            uint32_t parity) {
  switch (n) {
#define CB2_E1_CASE(N) case N: e1_unit<N>(taddr, Pc, cgk, t_oi, t_oj, a1_dst, acc1_full, parity); break;
    CB2_E1_CASE(4) CB2_E1_CASE(5) CB2_E1_CASE(6) CB2_E1_CASE(7)
    CB2_E1_CASE(8) CB2_E1_CASE(9) CB2_E1_CASE(10) CB2_E1_CASE(11) CB2_E1_CASE(12) CB2_E1_CASE(13)
    CB2_E1_CASE(14) CB2_E1_CASE(15) CB2_E1_CASE(16) CB2_E1_CASE(17) CB2_E1_CASE(18) CB2_E1_CASE(19)
    CB2_E1_CASE(20) CB2_E1_CASE(21) CB2_E1_CASE(22) CB2_E1_CASE(23) CB2_E1_CASE(24) CB2_E1_CASE(25)
    CB2_E1_CASE(26) CB2_E1_CASE(27) CB2_E1_CASE(28) CB2_E1_CASE(29) CB2_E1_CASE(30) CB2_E1_CASE(31)
    CB2_E1_CASE(32) CB2_E1_CASE(33) CB2_E1_CASE(34) CB2_E1_CASE(35) CB2_E1_CASE(36) CB2_E1_CASE(37)
    CB2_E1_CASE(38) CB2_E1_CASE(39) CB2_E1_CASE(40)
#undef CB2_E1_CASE
    default: e1_unit<0>(taddr, Pc, cgk, t_oi, t_oj, a1_dst, acc1_full, parity); break;
  }
}

}  // namespace cb2
