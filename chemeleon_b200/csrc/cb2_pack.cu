// Host-side packing of the fp16 tcgen05 operand images (what weights.pack_weights produces in
// Python), so that a non-Python host can feed cb2_model from a plain fp32 state_dict.  Pure CPU code.
#include "cb2_internal.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace cb2 {

static inline uint16_t f2h(float v) {
  const __half h = __float2half_rn(v);      // host implementation of cuda_fp16.h: round to nearest even
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}
static inline uint16_t d2h(double v) { return f2h((float)v); }
static inline double h2d(uint16_t u) {
  __half h;
  std::memcpy(&h, &u, 2);
  return (double)__half2float(h);
}

// [rows,K] row-major -> K-major no-swizzle image [K/8][rows][8]
static void kmajor(const float *w, int rows, int K, int64_t ldw, uint16_t *out, const int *col_perm) {
  for (int k8 = 0; k8 < K / 8; k8++)
    for (int r = 0; r < rows; r++)
      for (int e = 0; e < 8; e++) {
        const int k = k8 * 8 + e;
        out[((int64_t)k8 * rows + r) * 8 + e] = f2h(w[(int64_t)r * ldw + (col_perm ? col_perm[k] : k)]);
      }
}

}  // namespace cb2

using namespace cb2;

extern "C" {

size_t cb2_pack_bytes(int32_t kind, int32_t rows, int32_t K) {
  if (rows <= 0 || K <= 0 || K % 8 != 0) return 0;
  switch (kind) {
    case CB2_PACK_KMAJOR:
    case CB2_PACK_FD:
    case CB2_PACK_ROW_BLOCKS: return (size_t)rows * K * 2;
    case CB2_PACK_HEAD_SPLIT: return 16 + (size_t)256 * 3 * K * 2;
    default: return 0;
  }
}

int cb2_pack_weights(int32_t kind, const float *w, int32_t rows, int32_t K, void *out, size_t out_bytes) {
  if (!w || !out) return fail(CB2_ERR_BAD_ARG, "pack_weights: null argument");
  const size_t need = cb2_pack_bytes(kind, rows, K);
  if (need == 0) return fail(CB2_ERR_BAD_ARG, "pack_weights: bad kind / shape (K must be a multiple of 8)");
  if (out_bytes < need) return fail(CB2_ERR_WORKSPACE, "pack_weights: output buffer too small: call cb2_pack_bytes()");
  uint16_t *o = reinterpret_cast<uint16_t *>(out);
  switch (kind) {
    case CB2_PACK_KMAJOR:
      kmajor(w, rows, K, K, o, nullptr);
      return CB2_OK;
    case CB2_PACK_FD: {
      // W_fd [rows, 6F]: reference column order = sin block [d F + k] then cos block [3F + d F + k]
      // (cspnet.py:49-51); kernel order K' = d 2F + 2k + {0: sin, 1: cos}
      if (K % 6 != 0) return fail(CB2_ERR_BAD_ARG, "pack_weights: W_fd needs K = 6 * num_freqs");
      const int F = K / 6;
      std::vector<int> perm(K);
      for (int d = 0; d < 3; d++)
        for (int k = 0; k < F; k++) {
          perm[d * 2 * F + 2 * k] = d * F + k;
          perm[d * 2 * F + 2 * k + 1] = 3 * F + d * F + k;
        }
      kmajor(w, rows, K, K, o, perm.data());
      return CB2_OK;
    }
    case CB2_PACK_ROW_BLOCKS:
      // one K-major image per block of 128 output rows: [rows/128][K/8][128][8]
      // ... of w / 2 (exact in fp16): the GEMM2 accumulator of the edge kernels then holds the x / 2 that the tanh
      // form of SiLU starts from (cb2_tc.cuh silu_of_half)
      if (rows % 128 != 0) return fail(CB2_ERR_BAD_ARG, "pack_weights: row blocks need rows % 128 == 0");
      {
        std::vector<float> half((size_t)rows * K);
        for (int64_t i = 0; i < (int64_t)rows * K; i++) half[i] = 0.5f * w[i];
        for (int b = 0; b < rows / 128; b++)
          kmajor(half.data() + (int64_t)b * 128 * K, 128, K, K, o + (int64_t)b * 128 * K, nullptr);
      }
      return CB2_OK;
    case CB2_PACK_HEAD_SPLIT: {
      // 16-byte header {float 1/s} + K-major image [3K/8][256][8] of s [w_hi | w_hi | w_lo] (rows >= `rows` zero);
      // s = 2^floor(log2(1024 / max|w|)) keeps w_lo = s w - w_hi out of the fp16 subnormals
      if (rows > 256) return fail(CB2_ERR_BAD_ARG, "pack_weights: head image holds at most 256 rows");
      double wmax = 0.0;
      for (int64_t i = 0; i < (int64_t)rows * K; i++) wmax = std::fmax(wmax, std::fabs((double)w[i]));
      const double s = wmax > 0 ? std::exp2(std::floor(std::log2(1024.0 / wmax))) : 1.0;
      std::memset(out, 0, need);
      const float inv = (float)(1.0 / s);
      std::memcpy(out, &inv, 4);
      uint16_t *img = o + 8;
      const int K3 = 3 * K;
      for (int r = 0; r < rows; r++)
        for (int k = 0; k < K; k++) {
          const double ws = (double)w[(int64_t)r * K + k] * s;
          const uint16_t hi = d2h(ws), lo = d2h(ws - h2d(hi));
          const int cols[3] = {k, K + k, 2 * K + k};
          const uint16_t vals[3] = {hi, hi, lo};
          for (int t = 0; t < 3; t++) img[((int64_t)(cols[t] / 8) * 256 + r) * 8 + cols[t] % 8] = vals[t];
        }
      (void)K3;
      return CB2_OK;
    }
    default: return fail(CB2_ERR_BAD_ARG, "pack_weights: unknown kind");
  }
}

}  // extern "C"
