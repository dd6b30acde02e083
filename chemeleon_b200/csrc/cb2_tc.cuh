// Shared declarations of the tensor-core translation units.
#pragma once

#include "cb2_internal.cuh"
#include "cb2_ptx.cuh"

namespace cb2 {

// x sigmoid(x) = h + h tanh(h) with h = x/2: three instructions, ONE MUFU op (tanh.approx.f32,
// relative error 2^-11, the size of the fp16 rounding the result feeds).  The epilogues of the
// edge kernel are MUFU-bound (2 x 65536 SiLUs per 128-edge item at 16 MUFU/clk/SM), so this
// form beats x / (1 + 2^(-x log2 e)) (ex2 + rcp: two MUFU ops) by 6 % of the whole step, at an
// unchanged error against the exact path (measured: 3.2e-4 .. 5.1e-4 per edge layer for both).
// x -> -inf gives -0, x -> +inf gives x.
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ... of h = x / 2: the W2 operand image is that of W2 / 2 and the units are pre-loaded with b2 / 2 (both exact),
// so the GEMM2 accumulator already holds h and E2 saves one multiply per element
__device__ __forceinline__ float silu_of_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) { return pack_half2_sat(a, b); }

// Two SiLUs per MUFU op: (a, b) -> fp16x2 {silu(a), silu(b)} with tanh.approx.f16x2.  Used where
// the result is rounded to fp16 anyway (a1, the GEMM2 operand) or averaged (E2): the inputs are
// rounded to fp16 first (saturating), h + h tanh(h) is one HFMA2 with a single rounding.
__device__ __forceinline__ uint32_t silu2_half(float a, float b) {
  const uint32_t xh = pack_half2_sat(a, b);
  uint32_t hh, t, y;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(hh) : "r"(xh), "r"(0x38003800u));   // h = x/2
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hh));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(hh), "r"(t));
  return y;
}

struct TcEdgeArgs {
  const __half *P;         // [V*N,1024] fp16 row-major hoisted terms (P_i | P_j)
  const float *cg;         // [B,512] fp32 per-crystal term W_ip vec(L L^T) + b1 (NULL = 0); added in E1, never rounded to fp16
  const int32_t *node2graph;
  int single_cta;          // V == 2: force the one-CTA kernel (cb2_model.flags & CB2_MODEL_EDGE_SINGLE_CTA)
  const float *x;          // [N,3]
  const int32_t *row_i;    // [n_tiles*128] node i of the tile's edge row, -1 = padding
  const int32_t *row_j;
  const int32_t *seg_n;    // [n_tiles]
  const __half *w_fd_t;    // [96][512][8]      K-major image of W_fd (kernel column order)
  const __half *w2_t;      // [4][64][128][8]   K-major image of W2 / 2, one block of 128 output channels each
  const float *b2;
  __half *agg16;           // [V*N, ld_agg] row-major, or row-panel layout with agg_kt columns per
  int64_t ld_agg;          // panel when agg_kt > 0; written at column offset agg_col
  int agg_col;
  int agg_kt;
  int N, V, n_tiles;
};

int launch_tc_edge(const TcEdgeArgs &a, int n_sm, cudaStream_t st);

}  // namespace cb2
