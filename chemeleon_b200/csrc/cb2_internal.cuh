// Internal helpers shared by the chemeleon_b200 CUDA sources (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/chemeleon_b200.h"

namespace cb2 {

constexpr int H = CB2_HIDDEN;          // 512
constexpr int H2 = 2 * CB2_HIDDEN;     // 1024
constexpr int NFREQ = CB2_NUM_FREQS;   // 128
constexpr int DIS = 6 * NFREQ;         // 768
constexpr int NTYPE = CB2_MAX_ATOM_TYPES;
constexpr int HEADC = CB2_HEAD_COLS;

// ---- error plumbing -------------------------------------------------------
void set_error(const std::string &msg);
int fail(cb2_status code, const std::string &msg);
void count_launch(int n = 1);

#define CB2_CUDA_OK(expr)                                                         \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess)                                                        \
      return cb2::fail(CB2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define CB2_LAUNCH_OK(name)                                                       \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess)                                                        \
      return cb2::fail(CB2_ERR_CUDA, std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
    cb2::count_launch();                                                          \
  } while (0)

#define CB2_TRY(expr)               \
  do {                              \
    int _s = (expr);                \
    if (_s != CB2_OK) return _s;    \
  } while (0)

// ---- tracing: NVTX ranges around the phases of a forward / timestep; header-only
// NVTX v3 costs a null-pointer check per call when no profiler is attached ----
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

// ---- workspace carving ------------------------------------------------------
struct Arena {
  char *base;
  size_t off;
  size_t cap;
  bool dry;  // true: only measure
  Arena(void *p, size_t c, bool d) : base((char *)p), off(0), cap(c), dry(d) {}
  template <typename T>
  T *take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T *p = dry ? nullptr : (T *)(base + off);
    off += bytes;
    return p;
  }
  bool ok() const { return dry || off <= cap; }
};

// Buffers of one decoder forward (all row-major fp32 unless noted).
struct ForwardWs {
  float *h;       // [VN,512] residual stream (tensor-core node-chain path: panel layout, whole panels of 128 rows)
  float *h_final; // where the trunk left the final h, row-major (NULL = h)
  float *y;       // [VN,512] FiLM projection / scratch
  float *cat;     // [VN,1024] cols 0:512 = LN(h) (hn), 512:1024 = aggregated edge features
  float *P;       // [VN,1024] hoisted edge-MLP terms P_i | P_j = hn [W_hi;W_hj]^T (fp16 on the tensor-core path);
                  //           the per-crystal term cg stays separate, in fp32
  float *z1;      // [VN,512] node-MLP hidden
  float *hf;      // [VN,512] final-LN features (when the caller does not want them)
  float *cg;      // [n_layers,B,512]  W_ip vec(L L^T) + b1 of every layer (one launch per forward)
  float *emb;     // [Ec,768] exact path: sinusoid embedding of an edge chunk
  float *a1;      // [Ec,512]
  float *e2;      // [Ec,512]
  // tensor-core path (fp16 row-major activations)
  __half *h16;    // [VN,512]
  __half *cat16;  // [VN,1024]
  __half *z16;    // [VN,512]
};

// Buffers the sampler step adds on top of ForwardWs.
struct StepWs {
  float *film_cond;  // [VB,1024]
  float *head_out;   // [VN,128]
  float *lat_out;    // [VB,9]
};

size_t carve_forward(Arena &a, const cb2_batch *b, int n_layers, int precision, ForwardWs &w);
size_t carve_step(Arena &a, const cb2_batch *b, StepWs &w);

// ---- launchers implemented across the .cu files ------------------------------
struct GemmEpilogue {
  const float *bias = nullptr;       // [N]
  // gathered per-row bias: C[r, c] += gbias[gidx[r % gmod] * gld + c] for c < gcols
  const float *gbias = nullptr;
  const int32_t *gidx = nullptr;
  int32_t gmod = 1, gcols = 0, gld = 0;
  // edge gather-add: C[r, c] += P[(prow_off + ei[r]) * 1024 + c] + P[(prow_off + ej[r]) * 1024 + 512 + c]
  const float *P = nullptr;
  const int32_t *ei = nullptr;
  const int32_t *ej = nullptr;
  int64_t prow_off = 0;
  // ... + cg[n2g[ei[r]] * 512 + c]: per-crystal lattice term of the edge MLP (only with P)
  const float *ecg = nullptr;
  const int32_t *n2g = nullptr;
  int silu = 0;
  const float *residual = nullptr;   // [M, ldr] added after the activation
  int64_t ldr = 0;
};

int launch_sgemm_nt(const float *A, int64_t lda, const float *W, float *C, int64_t ldc, int64_t M, int N,
                    int K, const GemmEpilogue &epi, cudaStream_t st);

// device helpers
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// two floats -> packed fp16x2 (a in the low half), saturating to +-65504 instead of inf
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

__device__ __forceinline__ float wrap01(float v) {
  // torch's float remainder(v, 1.0): fmod, then shift negatives up by 1.
  float r = fmodf(v, 1.0f);
  if (r != 0.0f && r < 0.0f) r += 1.0f;
  return r;
}

}  // namespace cb2
