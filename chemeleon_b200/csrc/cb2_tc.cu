// Tensor-core path of the CSPNet decoder: tcgen05.mma (kind::f16, fp16 operands,
// fp32 accumulators in TMEM), weights streamed by bulk async copies (TMA unit) from
// pre-tiled operand images, A operands built in shared memory by the CUDA cores.
//
//   k_tc_linear : C = epi(A16 W16^T)       node-level GEMMs (FiLM proj, hoisted
//                                           [W_hi;W_hj], node MLP)        cspnet.py:86,161
//   k_tc_edge   : fused CSPLayer edge model + scatter_mean, edges generated on the
//                 fly from per-tile (i,j) rows                            cspnet.py:138-160
//
// Shared-memory operand layout everywhere: K-major, no swizzle, [K/8][rows][8 halves]
// (core matrix = 8 rows x 16 B; SBO = 128 B, LBO = rows*16 B).
#include "cb2_tc.cuh"

namespace cb2 {

using namespace ptx;

// =============================================================================================
// k_tc_linear: 128 x 256 output tile per CTA, K streamed in chunks of 32 through a 4-stage ring
// (96 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, so that one CTA's
// epilogue overlaps the other's main loop).
//   warps 0-3: load the A chunk (fp16 row-major global -> canonical smem), later the epilogue
//   warp 4   : TMEM alloc, MMA issue (one lane)
// =============================================================================================
constexpr int TL_BM = 128, TL_NB = 256, TL_KC = 32, TL_STAGES = 4;
constexpr int TL_A_BYTES = TL_BM * TL_KC * 2;   // 8 KB
constexpr int TL_W_BYTES = TL_NB * TL_KC * 2;   // 16 KB
constexpr int TL_STAGE_BYTES = TL_A_BYTES + TL_W_BYTES;
constexpr int TL_SMEM = TL_STAGES * TL_STAGE_BYTES + 1024;

struct TcLinearArgs {
  const __half *A;
  int64_t lda;
  int64_t M;
  int K;
  const __half *Wt;   // operand image [K/8][Nw][8]
  int Nw;
  float *C;
  int64_t ldc;
  __half *C16;
  int64_t ldc16;
  const float *bias;
  int silu;
  const float *residual;
  int64_t ldr;
  const float *gbias;
  const int32_t *gidx;
  int gmod, gcols, gld;
};

__global__ void __launch_bounds__(160, 2) k_tc_linear(TcLinearArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + TL_STAGES * TL_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 64 + 8 * s; };
  const uint32_t acc_bar = bar_base + 128;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TL_STAGES * TL_STAGE_BYTES + 192);

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int64_t m0 = (int64_t)blockIdx.x * TL_BM;
  const int n0 = blockIdx.y * TL_NB;
  const int nk = g.K / TL_KC;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; s++) {
      mbar_init(full_bar(s), 129);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), TL_NB);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ---------------- producer ----------------
    const int r = tid;  // row of the tile
    const int64_t grow = m0 + r;
    const bool valid = grow < g.M;
    const __half *arow = g.A + (valid ? grow : 0) * g.lda;
    // the A rows are prefetched three K chunks ahead into registers (the loads of one chunk are
    // only 64 B per thread: without the look-ahead this loop is latency-bound)
    constexpr int PF = 3;
    uint4 pre[PF + 1][TL_KC / 8];
    auto gload = [&](uint4 (&dst)[TL_KC / 8], int kc) {
#pragma unroll
      for (int k8 = 0; k8 < TL_KC / 8; k8++)
        dst[k8] = valid ? *reinterpret_cast<const uint4 *>(arow + kc * TL_KC + k8 * 8) : make_uint4(0, 0, 0, 0);
    };
#pragma unroll
    for (int u = 0; u < PF; u++)
      if (u < nk) gload(pre[u], u);
    for (int kc0 = 0; kc0 < nk; kc0 += PF + 1) {
#pragma unroll
      for (int u = 0; u <= PF; u++) {
        const int kc = kc0 + u;
        if (kc >= nk) break;
        if (kc + PF < nk) gload(pre[(u + PF) % (PF + 1)], kc + PF);
        const int s = kc % TL_STAGES;
        mbar_wait(empty_bar(s), ((kc / TL_STAGES) & 1) ^ 1);
        const uint32_t a_s = sbase + s * TL_STAGE_BYTES;
        const uint32_t w_s = a_s + TL_A_BYTES;
        if (tid == 0) {
          mbar_arrive_expect_tx(full_bar(s), TL_W_BYTES);
#pragma unroll
          for (int k8 = 0; k8 < TL_KC / 8; k8++) {
            const __half *src = g.Wt + ((int64_t)(kc * (TL_KC / 8) + k8) * g.Nw + n0) * 8;
            bulk_g2s(w_s + k8 * (TL_NB * 16), src, TL_NB * 16, full_bar(s));
          }
        }
#pragma unroll
        for (int k8 = 0; k8 < TL_KC / 8; k8++)
          *reinterpret_cast<uint4 *>(smem + s * TL_STAGE_BYTES + k8 * (TL_BM * 16) + r * 16) = pre[u][k8];
        fence_proxy_async_smem();
        mbar_arrive(full_bar(s));
      }
    }
    // ---------------- epilogue ----------------
    mbar_wait(acc_bar, 0);
    tc_fence_after_sync();
    const float *gb = nullptr;
    if (g.gbias != nullptr && valid) gb = g.gbias + (int64_t)g.gidx[grow % g.gmod] * g.gld;
#pragma unroll 1
    for (int c0 = 0; c0 < TL_NB; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
      tmem_ld_wait();
      if (!valid) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int c = n0 + c0 + j;
        float x = __uint_as_float(acc[j]);
        if (g.bias != nullptr) x += g.bias[c];
        if (gb != nullptr && c < g.gcols) x += gb[c];
        if (g.silu) x = silu_fast(x);
        v[j] = x;
      }
      if (g.residual != nullptr) {
        const float4 *rr = reinterpret_cast<const float4 *>(g.residual + grow * g.ldr + n0 + c0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float4 t = rr[j];
          v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
        }
      }
      if (g.C != nullptr) {
        float4 *cc = reinterpret_cast<float4 *>(g.C + grow * g.ldc + n0 + c0);
#pragma unroll
        for (int j = 0; j < 8; j++) cc[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (g.C16 != nullptr) {
        uint4 *cc = reinterpret_cast<uint4 *>(g.C16 + grow * g.ldc16 + n0 + c0);
#pragma unroll
        for (int j = 0; j < 4; j++)
          cc[j] = make_uint4(pack_half2(v[8 * j], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                             pack_half2(v[8 * j + 4], v[8 * j + 5]), pack_half2(v[8 * j + 6], v[8 * j + 7]));
      }
    }
    tc_fence_before_sync();
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16_f32(TL_BM, TL_NB);
      for (int kc = 0; kc < nk; kc++) {
        const int s = kc % TL_STAGES;
        mbar_wait(full_bar(s), (kc / TL_STAGES) & 1);
        tc_fence_after_sync();
        const uint32_t a_s = sbase + s * TL_STAGE_BYTES;
        const uint32_t w_s = a_s + TL_A_BYTES;
#pragma unroll
        for (int j = 0; j < TL_KC / 16; j++) {
          const uint64_t ad = smem_desc_kmajor(a_s + 2 * j * (TL_BM * 16), TL_BM * 16, 128);
          const uint64_t bd = smem_desc_kmajor(w_s + 2 * j * (TL_NB * 16), TL_NB * 16, 128);
          umma_f16(tmem, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(acc_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TL_NB);
}

int launch_tc_linear(const TcLinearArgs &a, cudaStream_t st) {
  if (a.M == 0) return CB2_OK;
  if (a.K % TL_KC != 0 || a.Nw % TL_NB != 0 || (a.lda % 8) != 0)
    return fail(CB2_ERR_BAD_ARG, "tc_linear: K%32, N%256, lda%8 must be 0");
  static bool attr_set = false;
  if (!attr_set) {
    CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)((a.M + TL_BM - 1) / TL_BM), (unsigned)(a.Nw / TL_NB));
  k_tc_linear<<<grid, 160, TL_SMEM, st>>>(a);
  CB2_LAUNCH_OK("k_tc_linear");
  return CB2_OK;
}

// launchers from cb2_kernels_f32.cu reused by the tensor-core path
int launch_film_apply(const float *y, float *h, const float *cond, const int32_t *node2graph, const float *fg,
                      const float *fb, const float *cg, const float *cb, float *hn, int64_t ld_hn, __half *hn16,
                      int64_t ld_hn16, int N, int B, int V, cudaStream_t st);
int launch_lattice_ip(const float *lat, const float *w_ip, const float *b1, float *cg, int B, cudaStream_t st);
int launch_to_half(const float *x, __half *y, int64_t n, cudaStream_t st);

static int g_num_sms = 0;

static int num_sms(int *out) {
  if (g_num_sms == 0) {
    int dev = 0;
    CB2_CUDA_OK(cudaGetDevice(&dev));
    CB2_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  *out = g_num_sms;
  return CB2_OK;
}

int tc_linear_simple(const void *A16, int64_t lda, const void *Wt, int Nw, const float *bias, float *C,
                     int64_t ldc, int64_t M, int K, int silu, cudaStream_t st) {
  TcLinearArgs a{};
  a.A = (const __half *)A16; a.lda = lda; a.M = M; a.K = K; a.Wt = (const __half *)Wt; a.Nw = Nw;
  a.C = C; a.ldc = ldc; a.bias = bias; a.silu = silu;
  return launch_tc_linear(a, st);
}

int tc_edge_layer(const cb2_layer_weights &L, const cb2_batch *b, const float *x, const float *P, __half *agg16,
                  int64_t ld_agg, int agg_col, cudaStream_t st) {
  int sms = 0;
  CB2_TRY(num_sms(&sms));
  TcEdgeArgs e{};
  e.P = P; e.x = x; e.row_i = b->tile_row_i; e.row_j = b->tile_row_j; e.seg_n = b->tile_seg_n;
  e.w_fd_t = (const __half *)L.w_fd_t; e.w2_t = (const __half *)L.w2_t; e.b2 = L.b2;
  e.agg16 = agg16; e.ld_agg = ld_agg; e.agg_col = agg_col; e.N = b->n_nodes; e.V = b->n_variants;
  e.n_tiles = b->n_tiles;
  return launch_tc_edge(e, sms, st);
}

int tc_forward_layers(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, ForwardWs &w,
                      cudaStream_t st) {
  const int N = b->n_nodes, B = b->n_graphs, V = b->n_variants;
  const int64_t VN = (int64_t)V * N;
  if (!m->film_wp_t) return fail(CB2_ERR_BAD_ARG, "tensor-core path needs the fp16 operand images (pack with tensor_core=True)");
  CB2_TRY(launch_to_half(w.h, w.h16, VN * H, st));
  for (int li = 0; li < m->n_layers; li++) {
    const cb2_layer_weights &L = m->layers[li];
    if (!L.w_hij_t || !L.w_fd_t || !L.w2_t || !L.wn1_t || !L.wn2_t)
      return fail(CB2_ERR_BAD_ARG, "tensor-core path: layer operand image missing");
    if (io->film_cond != nullptr) {
      TcLinearArgs a{};
      a.A = w.h16; a.lda = H; a.M = VN; a.K = H; a.Wt = (const __half *)m->film_wp_t; a.Nw = H;
      a.C = w.y; a.ldc = H; a.bias = m->film_bp;
      CB2_TRY(launch_tc_linear(a, st));
    }
    CB2_TRY(launch_film_apply(w.y, w.h, io->film_cond, b->node2graph, m->film_g, m->film_b, L.ln_g, L.ln_b,
                              nullptr, 0, w.cat16, H2, N, B, V, st));
    CB2_TRY(launch_lattice_ip(io->lattices, L.w_ip, L.b1, w.cg, B, st));
    {
      TcLinearArgs a{};   // P = hn [W_hi;W_hj]^T  (+ lattice term + b1 on the P_i half)
      a.A = w.cat16; a.lda = H2; a.M = VN; a.K = H; a.Wt = (const __half *)L.w_hij_t; a.Nw = H2;
      a.C = w.P; a.ldc = H2;
      a.gbias = w.cg; a.gidx = b->node2graph; a.gmod = N; a.gcols = H; a.gld = H;
      CB2_TRY(launch_tc_linear(a, st));
    }
    CB2_TRY(tc_edge_layer(L, b, io->frac_coords, w.P, w.cat16, H2, H, st));
    {
      TcLinearArgs a{};   // z = SiLU([hn|agg] Wn1^T + bn1)
      a.A = w.cat16; a.lda = H2; a.M = VN; a.K = H2; a.Wt = (const __half *)L.wn1_t; a.Nw = H;
      a.C16 = w.z16; a.ldc16 = H; a.bias = L.bn1; a.silu = 1;
      CB2_TRY(launch_tc_linear(a, st));
      TcLinearArgs a2{};  // h = h + SiLU(z Wn2^T + bn2)
      a2.A = w.z16; a2.lda = H; a2.M = VN; a2.K = H; a2.Wt = (const __half *)L.wn2_t; a2.Nw = H;
      a2.C = w.h; a2.ldc = H; a2.C16 = w.h16; a2.ldc16 = H; a2.bias = L.bn2; a2.silu = 1;
      a2.residual = w.h; a2.ldr = H;
      CB2_TRY(launch_tc_linear(a2, st));
    }
  }
  return CB2_OK;
}

}  // namespace cb2
