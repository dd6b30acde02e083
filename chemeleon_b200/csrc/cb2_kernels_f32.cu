// Exact-mode (fp32, CUDA-core) kernels of the CSPNet decoder.
//
// These follow the reference's arithmetic closely (fp32 everywhere, accurate
// expf/sinf/cosf, deterministic summation order) and are the parity anchor for
// the tensor-core path.  Reference: chemeleon/modules/cspnet.py.
#include "cb2_internal.cuh"

namespace cb2 {

// ---------------------------------------------------------------------------
// node embedding: h[v*N + n] = emb[a[n]]                      (cspnet.py:357)
// ---------------------------------------------------------------------------
__global__ void k_embed(const int64_t *__restrict__ a, const float *__restrict__ emb, float *__restrict__ h,
                        int N, int V) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over V*N*128 float4
  int64_t total = (int64_t)V * N * (H / 4);
  if (idx >= total) return;
  int c4 = (int)(idx % (H / 4));
  int64_t row = idx / (H / 4);
  int n = (int)(row % N);
  long long t = a[n];
  if (t < 0 || t >= NTYPE) t = 0;
  reinterpret_cast<float4 *>(h)[idx] = reinterpret_cast<const float4 *>(emb)[(int64_t)t * (H / 4) + c4];
}

// ... into the panel layouts of the node-chain kernel (cb2_tc_node2.cu): fp32 [panel][128 c4][128 rows][4] and the
// fp16 operand copy [panel][64 k8][128 rows][8]; a thread = (row, k8), rows past V*N are zero-filled
__global__ void k_embed_panels(const int64_t *__restrict__ a, const float *__restrict__ emb, float *__restrict__ hp,
                               __half *__restrict__ h16, int N, int V) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // [panel][k8][row]
  const int64_t rows = (int64_t)V * N, rows_p = (rows + 127) / 128 * 128;
  if (idx >= rows_p * (H / 8)) return;
  const int r = (int)(idx % 128), k8 = (int)((idx / 128) % (H / 8));
  const int64_t panel = idx / (128 * (H / 8)), row = panel * 128 + r;
  float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
  if (row < rows) {
    long long t = a[row % N];
    if (t < 0 || t >= NTYPE) t = 0;
    lo = reinterpret_cast<const float4 *>(emb)[(int64_t)t * (H / 4) + 2 * k8];
    hi = reinterpret_cast<const float4 *>(emb)[(int64_t)t * (H / 4) + 2 * k8 + 1];
  }
  float4 *dst = reinterpret_cast<float4 *>(hp) + panel * (128 * 128) + (2 * k8) * 128 + r;
  dst[0] = lo;
  dst[128] = hi;
  reinterpret_cast<uint4 *>(h16)[idx] =
      make_uint4(pack_half2_sat(lo.x, lo.y), pack_half2_sat(lo.z, lo.w), pack_half2_sat(hi.x, hi.y), pack_half2_sat(hi.z, hi.w));
}

int launch_embed_panels(const int64_t *a, const float *emb, float *hp, __half *h16, int N, int V, cudaStream_t st) {
  const int64_t rows_p = ((int64_t)V * N + 127) / 128 * 128, total = rows_p * (H / 8);
  if (total == 0) return CB2_OK;
  k_embed_panels<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, emb, hp, h16, N, V);
  CB2_LAUNCH_OK("k_embed_panels");
  return CB2_OK;
}

int launch_embed(const int64_t *a, const float *emb, float *h, int N, int V, cudaStream_t st) {
  int64_t total = (int64_t)V * N * (H / 4);
  if (total == 0) return CB2_OK;
  k_embed<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, emb, h, N, V);
  CB2_LAUNCH_OK("k_embed");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// FiLM conditioning: cond[r] = SiLU(time_table[t] + text_part[r])   (cspnet.py:70-73,80-81)
// ---------------------------------------------------------------------------
// text_row (optional): crystal r uses row text_row[r] of text_part, so that crystals with the same
// prompt share one row (and all null rows are one row)
__global__ void k_film_cond(const float *__restrict__ time_table, const float *__restrict__ text_part,
                            const int32_t *__restrict__ text_row, const int32_t *__restrict__ t_dev,
                            float *__restrict__ out, int64_t rows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * H2) return;
  int c = (int)(idx % H2);
  float v = text_row ? text_part[(int64_t)text_row[idx / H2] * H2 + c] : text_part[idx];
  if (time_table != nullptr) v += time_table[(int64_t)(*t_dev) * H2 + c];
  out[idx] = silu_exact(v);
}

int launch_film_cond(const float *time_table, const float *text_part, const int32_t *text_row, const int32_t *t_dev,
                     float *out, int64_t rows, cudaStream_t st) {
  int64_t total = rows * H2;
  if (total == 0) return CB2_OK;
  k_film_cond<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(time_table, text_part, text_row, t_dev, out, rows);
  CB2_LAUNCH_OK("k_film_cond");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// SGEMM  C[M,N] = epi(A[M,K] W[N,K]^T)   (both operands K-contiguous: torch Linear)
// 128x128x16 tiles, 256 threads, 8x8 outputs per thread.
// ---------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;

__global__ void __launch_bounds__(256) k_sgemm_nt(const float *__restrict__ A, int64_t lda,
                                                  const float *__restrict__ W, float *__restrict__ C,
                                                  int64_t ldc, int64_t M, int N, int K, GemmEpilogue epi) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid / 16, tx = tid % 16;
  // loader mapping: 2 float4 of A and 2 of W per thread per k-tile
  const int lrow = tid / 4;  // 0..63
  const int lk = (tid % 4) * 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int64_t r = m0 + lrow + 64 * i;
      ra[i] = (r < M) ? *reinterpret_cast<const float4 *>(A + r * lda + k0 + lk) : make_float4(0, 0, 0, 0);
      int n = n0 + lrow + 64 * i;
      rb[i] = (n < N) ? *reinterpret_cast<const float4 *>(W + (int64_t)n * K + k0 + lk) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int r = lrow + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
      As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y;
      Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
    }
  };
  const int nk = K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8 + 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int64_t r = m0 + ty * 8 + i;
    if (r >= M) continue;
    const float *gb = nullptr;
    if (epi.gbias != nullptr) gb = epi.gbias + (int64_t)epi.gidx[r % epi.gmod] * epi.gld;
    const float *pi = nullptr, *pj = nullptr, *ecg = nullptr;
    if (epi.P != nullptr) {
      const int ni = epi.ei[r];
      pi = epi.P + (epi.prow_off + ni) * H2;
      pj = epi.P + (epi.prow_off + epi.ej[r]) * H2 + H;
      if (epi.ecg != nullptr) ecg = epi.ecg + (int64_t)epi.n2g[ni] * H;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int c = n0 + tx * 8 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (epi.bias != nullptr) v += epi.bias[c];
      if (gb != nullptr && c < epi.gcols) v += gb[c];
      if (pi != nullptr) v += pi[c] + pj[c];
      if (ecg != nullptr) v += ecg[c];
      if (epi.silu) v = silu_exact(v);
      if (epi.residual != nullptr) v += epi.residual[r * epi.ldr + c];
      acc[i][j] = v;
    }
    float *crow = C + r * ldc + n0 + tx * 8;
    if (n0 + tx * 8 + 7 < N) {
      *reinterpret_cast<float4 *>(crow) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4 *>(crow + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    } else {
      for (int j = 0; j < 8; j++)
        if (n0 + tx * 8 + j < N) crow[j] = acc[i][j];
    }
  }
}

int launch_sgemm_nt(const float *A, int64_t lda, const float *W, float *C, int64_t ldc, int64_t M, int N,
                    int K, const GemmEpilogue &epi, cudaStream_t st) {
  if (M == 0) return CB2_OK;
  if (K % BK != 0 || (lda % 4) != 0 || (ldc % 4) != 0)
    return fail(CB2_ERR_BAD_ARG, "sgemm: K must be a multiple of 16 and lda/ldc multiples of 4");
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  k_sgemm_nt<<<grid, 256, 0, st>>>(A, lda, W, C, ldc, M, N, K, epi);
  CB2_LAUNCH_OK("k_sgemm_nt");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// FiLM apply + CSP LayerNorm, one warp per node row.
//   h' = SiLU(LN_f(y) * scale + shift) + h          (cspnet.py:88-96)
//   hn = LN_c(h')                                   (cspnet.py:174-176)
// With cond == nullptr the FiLM block is skipped (t=None,text=None: cspnet.py:372).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ln_row(float (&v)[16], const float *__restrict__ g, const float *__restrict__ b,
                                       int lane) {
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 16; q++) s += v[q];
  float mean = warp_sum(s) * (1.0f / H);
  float s2 = 0.f;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    float d = v[q] - mean;
    s2 += d * d;
  }
  float var = warp_sum(s2) * (1.0f / H);
  float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    int c = (lane + 32 * q4) * 4;
    float4 gg = *reinterpret_cast<const float4 *>(g + c);
    float4 bb = *reinterpret_cast<const float4 *>(b + c);
    v[q4 * 4 + 0] = (v[q4 * 4 + 0] - mean) * rstd * gg.x + bb.x;
    v[q4 * 4 + 1] = (v[q4 * 4 + 1] - mean) * rstd * gg.y + bb.y;
    v[q4 * 4 + 2] = (v[q4 * 4 + 2] - mean) * rstd * gg.z + bb.z;
    v[q4 * 4 + 3] = (v[q4 * 4 + 3] - mean) * rstd * gg.w + bb.w;
  }
}

__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&v)[16], int lane) {
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    float4 t = *reinterpret_cast<const float4 *>(p + (lane + 32 * q4) * 4);
    v[q4 * 4 + 0] = t.x; v[q4 * 4 + 1] = t.y; v[q4 * 4 + 2] = t.z; v[q4 * 4 + 3] = t.w;
  }
}

__device__ __forceinline__ void store_row(float *__restrict__ p, const float (&v)[16], int lane) {
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++)
    *reinterpret_cast<float4 *>(p + (lane + 32 * q4) * 4) =
        make_float4(v[q4 * 4 + 0], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
}

__device__ __forceinline__ void store_row_half(__half *__restrict__ p, const float (&v)[16], int lane) {
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    uint2 u;
    u.x = pack_half2_sat(v[q4 * 4 + 0], v[q4 * 4 + 1]);
    u.y = pack_half2_sat(v[q4 * 4 + 2], v[q4 * 4 + 3]);
    *reinterpret_cast<uint2 *>(p + (lane + 32 * q4) * 4) = u;
  }
}

// fp16 copy of a normalised row in the row-panel layout [rows/128][kt/8][128][8] the tensor-core
// GEMMs read their A operand from (lane pairs write one 16-byte core-matrix row).
__device__ __forceinline__ void store_row_half_panel(__half *__restrict__ base, int64_t row, int kt,
                                                     const float (&v)[16], int lane) {
  __half *p = base + (row >> 7) * 128 * kt + (row & 127) * 8;
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    const int c = (lane + 32 * q4) * 4;
    uint2 u;
    u.x = pack_half2_sat(v[q4 * 4 + 0], v[q4 * 4 + 1]);
    u.y = pack_half2_sat(v[q4 * 4 + 2], v[q4 * 4 + 3]);
    *reinterpret_cast<uint2 *>(p + (c >> 3) * 1024 + (c & 7)) = u;
  }
}

__global__ void __launch_bounds__(256, 3) k_film_apply(const float *__restrict__ y, float *__restrict__ h,
                                                    const float *__restrict__ cond,
                                                    const int32_t *__restrict__ node2graph,
                                                    const float *__restrict__ fg, const float *__restrict__ fb,
                                                    const float *__restrict__ cg, const float *__restrict__ cb,
                                                    float *__restrict__ hn, int64_t ld_hn,
                                                    __half *__restrict__ hn16, int64_t ld_hn16, int hn16_kt,
                                                    int N, int B, int64_t rows) {
  int64_t row = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float v[16];
  if (cond != nullptr) {
    // every global load of the row is issued before the first reduction (the per-crystal FiLM
    // row hangs off the node2graph load: fetch that one first)
    const int gidx = node2graph[(int)(row % N)];
    float hv[16];
    load_row(y + row * H, v, lane);
    load_row(h + row * H, hv, lane);
    const float *cs = cond + ((row / N) * B + gidx) * H2;
    float4 sc[4], sh[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int c = (lane + 32 * q4) * 4;
      sc[q4] = *reinterpret_cast<const float4 *>(cs + c);
      sh[q4] = *reinterpret_cast<const float4 *>(cs + H + c);
    }
    ln_row(v, fg, fb, lane);
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      v[q4 * 4 + 0] = silu_exact(v[q4 * 4 + 0] * sc[q4].x + sh[q4].x) + hv[q4 * 4 + 0];
      v[q4 * 4 + 1] = silu_exact(v[q4 * 4 + 1] * sc[q4].y + sh[q4].y) + hv[q4 * 4 + 1];
      v[q4 * 4 + 2] = silu_exact(v[q4 * 4 + 2] * sc[q4].z + sh[q4].z) + hv[q4 * 4 + 2];
      v[q4 * 4 + 3] = silu_exact(v[q4 * 4 + 3] * sc[q4].w + sh[q4].w) + hv[q4 * 4 + 3];
    }
    store_row(h + row * H, v, lane);
  } else {
    load_row(h + row * H, v, lane);
  }
  ln_row(v, cg, cb, lane);
  if (hn != nullptr) store_row(hn + row * ld_hn, v, lane);
  if (hn16 != nullptr) {
    if (hn16_kt > 0) store_row_half_panel(hn16, row, hn16_kt, v, lane);
    else store_row_half(hn16 + row * ld_hn16, v, lane);
  }
}

int launch_film_apply(const float *y, float *h, const float *cond, const int32_t *node2graph, const float *fg,
                      const float *fb, const float *cg, const float *cb, float *hn, int64_t ld_hn, __half *hn16,
                      int64_t ld_hn16, int hn16_kt, int N, int B, int V, cudaStream_t st) {
  int64_t rows = (int64_t)V * N;
  if (rows == 0) return CB2_OK;
  k_film_apply<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(y, h, cond, node2graph, fg, fb, cg, cb, hn, ld_hn,
                                                          hn16, ld_hn16, hn16_kt, N, B, rows);
  CB2_LAUNCH_OK("k_film_apply");
  return CB2_OK;
}

// plain LayerNorm rows (final_layer_norm, cspnet.py:385-386).  split16 (tensor-core path): the
// normalised row is also written as two fp16 terms hi + lo (hi = fp16(x), lo = fp16(x - hi)) into
// a 1024-column row-panel buffer, hi at columns 0:512 and lo at 512:1024 -- the A operand of the
// split-precision head GEMM (three fp16 products reproduce the fp32 product to ~2^-22).
__global__ void __launch_bounds__(256) k_layernorm(const float *__restrict__ x, const float *__restrict__ g,
                                                   const float *__restrict__ b, float *__restrict__ out,
                                                   __half *__restrict__ split16, const float *__restrict__ w3,
                                                   float *__restrict__ out3, int64_t ld3, int64_t rows) {
  int64_t row = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float v[16];
  load_row(x + row * H, v, lane);
  ln_row(v, g, b, lane);
  if (out != nullptr) store_row(out + row * H, v, lane);
  if (split16 != nullptr) {
    float lo[16];
#pragma unroll
    for (int q = 0; q < 16; q++) lo[q] = v[q] - __half2float(__float2half_rn(fminf(fmaxf(v[q], -65504.f), 65504.f)));
    store_row_half_panel(split16, row, H2, v, lane);
    store_row_half_panel(split16 + (H / 8) * 1024, row, H2, lo, lane);
  }
  if (w3 != nullptr) {
    // three-row head on the normalised row, fp32 (coord_out of the corrector forward, cspnet.py:388:
    // the type logits and the lattice head of that forward are never used, chemeleon.py:440-450)
    float d[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      float s = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < 4; q4++) {
        const float4 w = *reinterpret_cast<const float4 *>(w3 + r * H + (lane + 32 * q4) * 4);
        s = fmaf(v[q4 * 4 + 0], w.x, s); s = fmaf(v[q4 * 4 + 1], w.y, s);
        s = fmaf(v[q4 * 4 + 2], w.z, s); s = fmaf(v[q4 * 4 + 3], w.w, s);
      }
      d[r] = warp_sum(s);
    }
    if (lane < 3) out3[row * ld3 + lane] = lane == 0 ? d[0] : lane == 1 ? d[1] : d[2];
  }
}

int launch_layernorm(const float *x, const float *g, const float *b, float *out, __half *split16, const float *w3,
                     float *out3, int64_t ld3, int64_t rows, cudaStream_t st) {
  if (rows == 0) return CB2_OK;
  k_layernorm<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, g, b, out, split16, w3, out3, ld3, rows);
  CB2_LAUNCH_OK("k_layernorm");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// lattice term of the edge MLP, once per crystal:
//   cg[g] = W_ip vec(L L^T) + b1                       (cspnet.py:143-149, hoisted)
// ---------------------------------------------------------------------------
struct LatticeIpLayers {
  const float *w_ip[CB2_MAX_LAYERS];
  const float *b1[CB2_MAX_LAYERS];
};

// grid = (ceil(B / LIP_G), n_layers): one launch per forward covers every layer's term, cg[l][g][:].  A thread owns a
// channel: its nine weights and bias stay in registers over LIP_G crystals (one block per crystal re-read them
// with a 36-byte stride 4096 x 6 times per forward: 60 us at C3 for 50 MB of output)
constexpr int LIP_G = 32;
__global__ void __launch_bounds__(512) k_lattice_ip(const float *__restrict__ lat, LatticeIpLayers lw,
                                                    float *__restrict__ cg, int B, int32_t *__restrict__ range_flags) {
  const int g0 = blockIdx.x * LIP_G, li = blockIdx.y;
  const int ng = min(LIP_G, B - g0);
  __shared__ float ip[LIP_G][9];
  if (threadIdx.x < 9 * ng) {
    const int gl = threadIdx.x / 9, e = threadIdx.x % 9;
    const int a = e / 3, b = e % 3;
    const float *L = lat + (int64_t)(g0 + gl) * 9;
    ip[gl][e] = L[a * 3 + 0] * L[b * 3 + 0] + L[a * 3 + 1] * L[b * 3 + 1] + L[a * 3 + 2] * L[b * 3 + 2];
  }
  __syncthreads();
  const int c = threadIdx.x;
  const float *w_ip = lw.w_ip[li] + c * 9;
  float w[9];
#pragma unroll
  for (int m = 0; m < 9; m++) w[m] = w_ip[m];
  const float bias = lw.b1[li][c];
  for (int gl = 0; gl < ng; gl++) {
    float s = bias;
#pragma unroll
    for (int m = 0; m < 9; m++) s = fmaf(w[m], ip[gl][m], s);
    cg[((int64_t)li * B + g0 + gl) * H + c] = s;
    // tensor-core mode: outside this range the fp16 GEMM2 operand saturates -- tell the caller
    if (range_flags != nullptr && !(fabsf(s) <= CB2_TC_RANGE_LIMIT)) atomicOr(range_flags + g0 + gl, CB2_FLAG_TC_RANGE);
  }
}

// cg: [n_layers, B, 512]
int launch_lattice_ip(const float *lat, const cb2_model *m, float *cg, int B, int32_t *range_flags, cudaStream_t st) {
  if (B == 0) return CB2_OK;
  LatticeIpLayers lw;
  for (int li = 0; li < m->n_layers; li++) {
    lw.w_ip[li] = m->layers[li].w_ip;
    lw.b1[li] = m->layers[li].b1;
  }
  k_lattice_ip<<<dim3((unsigned)((B + LIP_G - 1) / LIP_G), (unsigned)m->n_layers), 512, 0, st>>>(lat, lw, cg, B, range_flags);
  CB2_LAUNCH_OK("k_lattice_ip");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// exact path: sinusoid embedding of an edge chunk in the reference's layout
//   emb[e, d*F+k] = sin(fl32(2 pi k) * ((x_j - x_i) mod 1)_d), cos block at +3F   (cspnet.py:48-52,324)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_embed(const float *__restrict__ x, const int32_t *__restrict__ ei,
                                                    const int32_t *__restrict__ ej, float *__restrict__ emb,
                                                    int64_t n_rows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over rows * 384
  if (idx >= n_rows * (3 * NFREQ)) return;
  int64_t r = idx / (3 * NFREQ);
  int dk = (int)(idx % (3 * NFREQ));
  int d = dk / NFREQ, k = dk % NFREQ;
  float fd = wrap01(x[(int64_t)ej[r] * 3 + d] - x[(int64_t)ei[r] * 3 + d]);
  float freq = __fmul_rn(6.283185307179586f, (float)k);
  float arg = __fmul_rn(fd, freq);
  float s, c;
  sincosf(arg, &s, &c);
  emb[r * DIS + dk] = s;
  emb[r * DIS + 3 * NFREQ + dk] = c;
}

int launch_edge_embed(const float *x, const int32_t *ei, const int32_t *ej, float *emb, int64_t n_rows,
                      cudaStream_t st) {
  int64_t total = n_rows * 3 * NFREQ;
  if (total == 0) return CB2_OK;
  k_edge_embed<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ei, ej, emb, n_rows);
  CB2_LAUNCH_OK("k_edge_embed");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// exact path: agg[i] = mean_j e[(i,j)]  (scatter.py:88-112 with index = edge_index[0]);
// segments are contiguous (i outer, j inner), summed in j order (deterministic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_segment_mean(const float *__restrict__ e, const int64_t *__restrict__ node_eoff,
                                                      const int32_t *__restrict__ node_n, float *__restrict__ out,
                                                      int64_t ld_out, int node_lo, int64_t e0, int64_t out_row_off) {
  int i = node_lo + blockIdx.x;
  int n = node_n[i];
  const float4 *src = reinterpret_cast<const float4 *>(e + (node_eoff[i] - e0) * H) + threadIdx.x;
  float4 s = make_float4(0, 0, 0, 0);
  for (int j = 0; j < n; j++) {
    float4 v = src[(int64_t)j * (H / 4)];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float cnt = (float)(n < 1 ? 1 : n);
  s.x = s.x / cnt; s.y = s.y / cnt; s.z = s.z / cnt; s.w = s.w / cnt;
  *(reinterpret_cast<float4 *>(out + (out_row_off + i) * ld_out) + threadIdx.x) = s;
}

int launch_segment_mean(const float *e, const int64_t *node_eoff, const int32_t *node_n, float *out,
                        int64_t ld_out, int node_lo, int node_hi, int64_t e0, int64_t out_row_off,
                        cudaStream_t st) {
  if (node_hi <= node_lo) return CB2_OK;
  k_segment_mean<<<node_hi - node_lo, 128, 0, st>>>(e, node_eoff, node_n, out, ld_out, node_lo, e0, out_row_off);
  CB2_LAUNCH_OK("k_segment_mean");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// lattice head: mean over the crystal's nodes, Linear(512,9), @ L   (cspnet.py:390-394)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_lattice_head(const float *__restrict__ hf, const float *__restrict__ w_lat,
                                                      const float *__restrict__ lat,
                                                      const int32_t *__restrict__ graph_off,
                                                      float *__restrict__ out, int N, int B) {
  int g = blockIdx.x % B;
  int v = blockIdx.x / B;
  int lo = graph_off[g], hi = graph_off[g + 1];
  int c = threadIdx.x;
  float s = 0.f;
  for (int n = lo; n < hi; n++) s += hf[((int64_t)v * N + n) * H + c];
  int cnt = hi - lo;
  s = s / (float)(cnt < 1 ? 1 : cnt);
  __shared__ float red[9][16];
  __shared__ float M[9];
  int lane = c % 32, warp = c / 32;
#pragma unroll
  for (int m = 0; m < 9; m++) {
    float p = warp_sum(s * w_lat[m * H + c]);
    if (lane == 0) red[m][warp] = p;
  }
  __syncthreads();
  if (c < 9) {
    float t = 0.f;
    for (int w = 0; w < 16; w++) t += red[c][w];
    M[c] = t;
  }
  __syncthreads();
  if (c < 9) {
    int i = c / 3, k = c % 3;
    const float *L = lat + (int64_t)g * 9;
    out[((int64_t)v * B + g) * 9 + c] = M[i * 3 + 0] * L[0 * 3 + k] + M[i * 3 + 1] * L[1 * 3 + k] + M[i * 3 + 2] * L[2 * 3 + k];
  }
}

int launch_lattice_head(const float *hf, const float *w_lat, const float *lat, const int32_t *graph_off,
                        float *out, int N, int B, int V, cudaStream_t st) {
  if (B * V == 0) return CB2_OK;
  k_lattice_head<<<B * V, 512, 0, st>>>(hf, w_lat, lat, graph_off, out, N, B);
  CB2_LAUNCH_OK("k_lattice_head");
  return CB2_OK;
}

// fp32 -> fp16 row copy (tensor-core path inputs)
__global__ void k_to_half(const float *__restrict__ x, __half *__restrict__ y, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4 *>(x)[i];
  uint2 u;
  u.x = pack_half2_sat(v.x, v.y);
  u.y = pack_half2_sat(v.z, v.w);
  reinterpret_cast<uint2 *>(y)[i] = u;
}

int launch_to_half(const float *x, __half *y, int64_t n, cudaStream_t st) {
  if (n == 0) return CB2_OK;
  k_to_half<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(x, y, n / 4);
  CB2_LAUNCH_OK("k_to_half");
  return CB2_OK;
}

// ---------------------------------------------------------------------------
// Text-conditioning tail (TextEncoder.get_text_embeds after the language model,
// text_encoder/text_encoder.py:40-45,186-205): x -> Linear -> LayerNorm -> GELU (exact, erf) -> Linear.
// This kernel is the LayerNorm + GELU between the two fp32 GEMMs; one warp per row, any width % 32 == 0.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ln_gelu(float *__restrict__ x, const float *__restrict__ g,
                                                 const float *__restrict__ b, int width, int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float *p = x + row * width;
  float s = 0.f;
  for (int c = lane; c < width; c += 32) s += p[c];
  const float mean = warp_sum(s) / (float)width;
  float s2 = 0.f;
  for (int c = lane; c < width; c += 32) {
    const float d = p[c] - mean;
    s2 += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(s2) / (float)width + 1e-5f);
  for (int c = lane; c < width; c += 32) {
    const float v = (p[c] - mean) * rstd * g[c] + b[c];
    p[c] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  }
}

int launch_ln_gelu(float *x, const float *g, const float *b, int width, int64_t rows, cudaStream_t st) {
  if (rows == 0) return CB2_OK;
  if (width % 32 != 0) return fail(CB2_ERR_BAD_ARG, "ln_gelu: width must be a multiple of 32");
  k_ln_gelu<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, g, b, width, rows);
  CB2_LAUNCH_OK("k_ln_gelu");
  return CB2_OK;
}

}  // namespace cb2
