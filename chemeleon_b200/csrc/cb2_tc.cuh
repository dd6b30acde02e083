// Shared declarations of the tensor-core translation units.
#pragma once

#include "cb2_internal.cuh"
#include "cb2_ptx.cuh"

namespace cb2 {

// x * sigmoid(x) = x / (1 + 2^(-x log2 e)) with the ftz ex2/rcp approximations: five
// instructions, two MUFU ops, relative error ~1e-6 (the result feeds an fp16 rounding).
// x -> -inf gives -0, x -> +inf gives x.  (A Newton reciprocal on the FMA pipe instead of
// MUFU.RCP was measured: 12 % slower when used everywhere, no gain when used in E1 only.)
__device__ __forceinline__ float silu_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) { return pack_half2_sat(a, b); }

struct TcEdgeArgs {
  const float *P;          // [V*N,1024] hoisted terms (P_i + lattice term + b1 | P_j)
  const float *x;          // [N,3]
  const int32_t *row_i;    // [n_tiles*128] node i of the tile's edge row, -1 = padding
  const int32_t *row_j;
  const int32_t *seg_n;    // [n_tiles]
  const __half *w_fd_t;    // [96][512][8]      K-major image of W_fd (kernel column order)
  const __half *w2_t;      // [4][64][128][8]   K-major image of W2, one block of 128 output channels each
  const float *b2;
  __half *agg16;           // [V*N, ld_agg] row-major, or row-panel layout with agg_kt columns per
  int64_t ld_agg;          // panel when agg_kt > 0; written at column offset agg_col
  int agg_col;
  int agg_kt;
  int N, V, n_tiles;
};

int launch_tc_edge(const TcEdgeArgs &a, int n_sm, cudaStream_t st);

}  // namespace cb2
