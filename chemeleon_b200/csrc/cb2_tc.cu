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
// k_tc_linear: persistent, warp-specialised GEMM  C = epi(A16 W16^T).
//   One CTA per SM walks over 128 x 256 output tiles (n fastest, so that the CTAs working on one
//   row panel at the same time share its A chunks through L2).  K is streamed in chunks of 64
//   through a 3-stage ring (48 KB per stage: 16 KB of A + 32 KB of W), two 256-column TMEM
//   accumulators let the epilogue of tile t overlap the main loop of tile t+1.
//     warp 0   : loader - bulk async copies only.  The fp16 activations live in HBM in the
//                "row-panel" layout [M/128][K/8][128 rows][8 halves], i.e. already in the canonical
//                K-major shared-memory order, so one chunk of A is ONE contiguous 16 KB copy
//     warp 1   : TMEM alloc, MMA issue (one lane)
//     warps 2-17: epilogue (a single warp per scheduler is issue-latency-bound: measured).  TMEM
//                hands every thread one row; 32 x 32 blocks are transposed through shared memory so
//                that bias / per-crystal bias / residual loads and the fp32 + fp16 stores are
//                coalesced (4 rows x 128 B per instruction)
// =============================================================================================
constexpr int TL_BM = 128, TL_NB = 256, TL_KC = 64, TL_STAGES = 3;
constexpr int TL_A_BYTES = TL_BM * TL_KC * 2;   // 16 KB
constexpr int TL_W_BYTES = TL_NB * TL_KC * 2;   // 32 KB
constexpr int TL_STAGE_BYTES = TL_A_BYTES + TL_W_BYTES;
constexpr int TL_PITCH = 36;                    // floats per staged row (16-byte aligned, +4 pad: conflict-free)
constexpr int TL_STG_BYTES = 16 * 32 * TL_PITCH * 4;   // one 32 x 32 fp32 block per epilogue warp
constexpr int TL_SMEM = TL_STAGES * TL_STAGE_BYTES + TL_STG_BYTES + 256;
constexpr int TL_THREADS = 32 * 18;

struct TcLinearArgs {
  const __half *A;    // row-panel layout, a_kt columns per panel; columns [0,K) are consumed
  int a_kt;
  int64_t M;
  int K;
  const __half *Wt;   // operand image [K/8][Nw][8]
  int Nw;
  float *C;           // fp32 row-major output (optional)
  int64_t ldc;
  __half *C16;        // fp16 row-panel output (optional): c16_kt columns per panel, written at [c16_k0, c16_k0+Nw)
  int c16_kt, c16_k0;
  __half *C16r;       // fp16 row-major output (optional), leading dimension ldc16r
  int64_t ldc16r;
  const float *bias;
  int silu;
  const float *residual;
  int64_t ldr;
  const float *gbias;     // per-crystal additive term, row r uses gbias[gidx[r % gmod]][c] for c < gcols
  const int32_t *gidx;
  int gmod, gcols, gld;
  int a_wrap;             // > 0: K chunk kc reads A chunk kc % a_wrap (A sections reused along K)
  int n_store;            // columns >= n_store are computed but not stored (0 = Nw)
  const float *out_scale; // device scalar multiplied into the accumulator before the bias (NULL = 1)
};

__global__ void __launch_bounds__(TL_THREADS, 1) k_tc_linear(TcLinearArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + TL_STAGES * TL_STAGE_BYTES + TL_STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 32 + 8 * s; };
  auto acc_full = [&](int b) { return bar_base + 64 + 8 * b; };
  auto acc_empty = [&](int b) { return bar_base + 80 + 8 * b; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TL_STAGES * TL_STAGE_BYTES + TL_STG_BYTES + 96);

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int n_nt = g.Nw / TL_NB;
  const int n_tiles = (int)((g.M + TL_BM - 1) / TL_BM) * n_nt;
  const int nk = g.K / TL_KC;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------- loader ----------------
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / n_nt, n0 = (t % n_nt) * TL_NB;
        const __half *ap = g.A + (int64_t)mt * TL_BM * g.a_kt;
        for (int kc = 0; kc < nk; kc++, it++) {
          const int s = it % TL_STAGES;
          mbar_wait_spin(empty_bar(s), ((it / TL_STAGES) & 1) ^ 1);
          const uint32_t a_s = sbase + s * TL_STAGE_BYTES, w_s = a_s + TL_A_BYTES;
          mbar_arrive_expect_tx(full_bar(s), TL_STAGE_BYTES);
          const int ka = g.a_wrap > 0 ? kc % g.a_wrap : kc;
          bulk_g2s(a_s, ap + (int64_t)ka * (TL_KC / 8) * (TL_BM * 8), TL_A_BYTES, full_bar(s));
#pragma unroll
          for (int k8 = 0; k8 < TL_KC / 8; k8++)
            bulk_g2s(w_s + k8 * (TL_NB * 16), g.Wt + ((int64_t)(kc * (TL_KC / 8) + k8) * g.Nw + n0) * 8, TL_NB * 16,
                     full_bar(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16_f32(TL_BM, TL_NB);
      int it = 0, tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
        const int buf = tl & 1;
        mbar_wait_spin(acc_empty(buf), ((tl >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int kc = 0; kc < nk; kc++, it++) {
          const int s = it % TL_STAGES;
          mbar_wait_spin(full_bar(s), (it / TL_STAGES) & 1);
          tc_fence_after_sync();
          const uint32_t a_s = sbase + s * TL_STAGE_BYTES, w_s = a_s + TL_A_BYTES;
#pragma unroll
          for (int j = 0; j < TL_KC / 16; j++) {
            const uint64_t ad = smem_desc_kmajor(a_s + 2 * j * (TL_BM * 16), TL_BM * 16, 128);
            const uint64_t bd = smem_desc_kmajor(w_s + 2 * j * (TL_NB * 16), TL_NB * 16, 128);
            umma_f16(tmem + buf * TL_NB, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full(buf));
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (16 warps) ----------------
    // A warp may only read the TMEM lane quarter warp % 4; the four warps of a quarter split the
    // 256 columns.  Each warp pulls its 32 x 64 block into registers in two halves (the accumulator
    // goes back to the MMA warp after the second) and writes them out through its staging rows.
    const int q = warp & 3, cgrp = (warp - 2) >> 2;
    float *stg = reinterpret_cast<float *>(smem + TL_STAGES * TL_STAGE_BYTES) + (warp - 2) * (32 * TL_PITCH);
    const int orow = lane >> 3, ocol = (lane & 7) * 4;             // output pass: 4 rows x 128 B per instruction
    const int n_store = g.n_store > 0 ? g.n_store : g.Nw;
    const float oscale = g.out_scale ? __ldg(g.out_scale) : 1.0f;
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      const int buf = tl & 1;
      const int mt = t / n_nt, n0 = (t % n_nt) * TL_NB + cgrp * 64;
      const int64_t r0 = (int64_t)mt * TL_BM + q * 32;             // first row of this warp
      // rows past M and column blocks past n_store are computed by the MMA but never touched here
      const int nvalid = n0 >= n_store ? 0 : (int)(g.M - r0 < 32 ? g.M - r0 : 32);
      int gi = 0;
      if (g.gbias != nullptr && lane < nvalid) gi = g.gidx[(r0 + lane) % g.gmod];
      mbar_wait(acc_full(buf), (tl >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * TL_NB + cgrp * 64;
      const int64_t crow = (r0 + orow) * g.ldc, rrow = (r0 + orow) * g.ldr;
      __half *c16p = g.C16 ? g.C16 + (int64_t)mt * TL_BM * g.c16_kt + (q * 32 + orow) * 8 : nullptr;
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        const int col = n0 + hb * 32 + ocol;
        const bool use_gb = g.gbias != nullptr && col < g.gcols;
        // the residual rows (or, without a residual, the per-crystal bias rows) of this pass are
        // fetched before the transposition, so that their latency is not paid store by store
        // (the residual aliases C: the compiler cannot move these loads above the stores itself)
        const bool pre_res = g.residual != nullptr;
        float4 pre[8];
#pragma unroll
        for (int itr = 0; itr < 8; itr++) {
          const int rr = itr * 4 + orow;
          int gsrc = 0;
          if (g.gbias != nullptr) gsrc = __shfl_sync(0xffffffffu, gi, rr);
          pre[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < nvalid) {
            if (pre_res) pre[itr] = *reinterpret_cast<const float4 *>(g.residual + rrow + (int64_t)itr * 4 * g.ldr + col);
            else if (use_gb) pre[itr] = *reinterpret_cast<const float4 *>(g.gbias + (int64_t)gsrc * g.gld + col);
          }
        }
        uint32_t acc[32];
        tmem_ld32(taddr + hb * 32, acc);
        tmem_ld_wait();
        if (hb == 1) {                                             // the accumulator goes back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(buf));
        }
        __syncwarp();                                              // previous output pass has drained the staging rows
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<uint4 *>(stg + lane * TL_PITCH + 4 * j) =
              make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias != nullptr && nvalid > 0) b4 = *reinterpret_cast<const float4 *>(g.bias + col);
        const int c16 = g.c16_k0 + col;
        __half *c16c = c16p ? c16p + (int64_t)(c16 >> 3) * (TL_BM * 8) + (c16 & 7) : nullptr;
#pragma unroll
        for (int itr = 0; itr < 8; itr++) {
          const int rr = itr * 4 + orow;
          int gsrc = 0;
          if (g.gbias != nullptr && pre_res) gsrc = __shfl_sync(0xffffffffu, gi, rr);
          if (rr < nvalid) {
            float4 x = *reinterpret_cast<const float4 *>(stg + rr * TL_PITCH + ocol);
            x.x = fmaf(x.x, oscale, b4.x); x.y = fmaf(x.y, oscale, b4.y);
            x.z = fmaf(x.z, oscale, b4.z); x.w = fmaf(x.w, oscale, b4.w);
            if (use_gb) {
              float4 t4 = pre[itr];
              if (pre_res) t4 = *reinterpret_cast<const float4 *>(g.gbias + (int64_t)gsrc * g.gld + col);
              x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
            }
            if (g.silu) { x.x = silu_fast(x.x); x.y = silu_fast(x.y); x.z = silu_fast(x.z); x.w = silu_fast(x.w); }
            if (pre_res) { x.x += pre[itr].x; x.y += pre[itr].y; x.z += pre[itr].z; x.w += pre[itr].w; }
            if (g.C != nullptr) *reinterpret_cast<float4 *>(g.C + crow + (int64_t)itr * 4 * g.ldc + col) = x;
            if (g.C16r != nullptr)
              *reinterpret_cast<uint2 *>(g.C16r + (r0 + rr) * g.ldc16r + col) = make_uint2(pack_half2(x.x, x.y), pack_half2(x.z, x.w));
            if (c16c != nullptr)
              *reinterpret_cast<uint2 *>(c16c + itr * 32) = make_uint2(pack_half2(x.x, x.y), pack_half2(x.z, x.w));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// SM count of the CURRENT device (queried per call: the library keeps no per-process device state,
// so engines on several GPUs of one process each get their own answer)
int num_sms(int *out) {
  int dev = 0;
  CB2_CUDA_OK(cudaGetDevice(&dev));
  CB2_CUDA_OK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return CB2_OK;
}

int launch_tc_linear(const TcLinearArgs &a, cudaStream_t st) {
  if (a.M == 0) return CB2_OK;
  if (a.K % TL_KC != 0 || a.K <= 0 || a.Nw % TL_NB != 0 || a.a_kt % 8 != 0 || (a.a_wrap == 0 && a.K > a.a_kt) ||
      a.a_wrap * TL_KC > a.a_kt || (a.n_store % 64) != 0)
    return fail(CB2_ERR_BAD_ARG, "tc_linear: K%64, N%256, a_kt%8, n_store%64 must be 0 and K <= a_kt");
  // per-device function attribute: set on every launch (cheap, legal during stream capture)
  CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM));
  int sms = 0;
  CB2_TRY(num_sms(&sms));
  const int64_t n_tiles = ((a.M + TL_BM - 1) / TL_BM) * (a.Nw / TL_NB);
  k_tc_linear<<<(unsigned)(n_tiles < sms ? n_tiles : sms), TL_THREADS, TL_SMEM, st>>>(a);
  CB2_LAUNCH_OK("k_tc_linear");
  return CB2_OK;
}

// Row-major -> row-panel layout of an fp16 / fp32 activation matrix (rows past M are zero-filled).
// A warp instruction covers 4 rows x 8 column groups: 256 B reads, 64 B writes per segment.
template <typename T>
__global__ void __launch_bounds__(256) k_to_panels(const T *__restrict__ x, int64_t ldx, __half *__restrict__ y, int64_t M,
                                                   int K, int kt, int k0) {
  const int k8n = K / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t grp = idx / 32;                 // one warp = 4 rows x 8 column groups
  const int lane = (int)(idx % 32);
  const int k8blocks = (k8n + 7) / 8;
  const int64_t r = (grp / k8blocks) * 4 + lane / 8;
  const int k8 = (int)(grp % k8blocks) * 8 + lane % 8;
  const int64_t Mp = (M + 127) / 128 * 128;
  if (r >= Mp || k8 >= k8n) return;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (r < M) {
    if constexpr (sizeof(T) == 2) {
      o = *reinterpret_cast<const uint4 *>(x + r * ldx + k8 * 8);
    } else {
      const float4 a = *reinterpret_cast<const float4 *>(x + r * ldx + k8 * 8);
      const float4 b = *reinterpret_cast<const float4 *>(x + r * ldx + k8 * 8 + 4);
      o = make_uint4(pack_half2(a.x, a.y), pack_half2(a.z, a.w), pack_half2(b.x, b.y), pack_half2(b.z, b.w));
    }
  }
  *reinterpret_cast<uint4 *>(y + (r / 128) * 128 * kt + (int64_t)((k0 >> 3) + k8) * 1024 + (r % 128) * 8) = o;
}

template <typename T>
static int launch_to_panels(const T *x, int64_t ldx, __half *y, int64_t M, int K, int kt, int k0, cudaStream_t st) {
  if (M == 0) return CB2_OK;
  const int64_t Mp = (M + 127) / 128 * 128;
  const int64_t warps = (Mp / 4) * ((K / 8 + 7) / 8);
  k_to_panels<T><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(x, ldx, y, M, K, kt, k0);
  CB2_LAUNCH_OK("k_to_panels");
  return CB2_OK;
}

// launchers from cb2_kernels_f32.cu reused by the tensor-core path
int launch_film_apply(const float *y, float *h, const float *cond, const int32_t *node2graph, const float *fg,
                      const float *fb, const float *cg, const float *cb, float *hn, int64_t ld_hn, __half *hn16,
                      int64_t ld_hn16, int hn16_kt, int N, int B, int V, cudaStream_t st);
int launch_lattice_ip(const float *lat, const cb2_model *m, float *cg, int B, int32_t *range_flags, cudaStream_t st);
int launch_tc_film(const cb2_model *m, const cb2_layer_weights &L, const cb2_batch *b, const float *film_cond,
                   const __half *h16, float *h, __half *cat16, int n_sm, cudaStream_t st);

// C-ABI unit entry: row-major fp16 A; the row-panel copy the kernel reads is made in the caller's workspace.
int tc_linear_simple(const void *A16, int64_t lda, const void *Wt, int Nw, const float *bias, float *C,
                     int64_t ldc, int64_t M, int K, int silu, void *workspace, size_t workspace_bytes, cudaStream_t st) {
  if (M == 0) return CB2_OK;
  if (K % TL_KC != 0 || lda % 8 != 0) return fail(CB2_ERR_BAD_ARG, "linear_tc: K%64 and lda%8 must be 0");
  const int64_t Mp = (M + 127) / 128 * 128;
  if (!workspace || workspace_bytes < (size_t)Mp * K * sizeof(__half))
    return fail(CB2_ERR_WORKSPACE, "linear_tc: workspace too small: call cb2_linear_tc_workspace_bytes()");
  __half *panels = reinterpret_cast<__half *>(workspace);
  CB2_TRY(launch_to_panels<__half>((const __half *)A16, lda, panels, M, K, K, 0, st));
  TcLinearArgs a{};
  a.A = panels; a.a_kt = K; a.M = M; a.K = K; a.Wt = (const __half *)Wt; a.Nw = Nw;
  a.C = C; a.ldc = ldc; a.bias = bias; a.silu = silu;
  return launch_tc_linear(a, st);
}

int tc_edge_layer(const cb2_model *m, const cb2_layer_weights &L, const cb2_batch *b, const float *x, const __half *P,
                  const float *cg, __half *agg16, int64_t ld_agg, int agg_col, int agg_kt, cudaStream_t st) {
  int sms = 0;
  CB2_TRY(num_sms(&sms));
  TcEdgeArgs e{};
  e.P = P; e.x = x; e.cg = cg; e.node2graph = b->node2graph; e.single_cta = (m->flags & CB2_MODEL_EDGE_SINGLE_CTA) != 0; e.row_i = b->tile_row_i; e.row_j = b->tile_row_j; e.seg_n = b->tile_seg_n;
  e.w_fd_t = (const __half *)L.w_fd_t; e.w2_t = (const __half *)L.w2_t; e.b2 = L.b2;
  e.agg16 = agg16; e.ld_agg = ld_agg; e.agg_col = agg_col; e.agg_kt = agg_kt; e.N = b->n_nodes; e.V = b->n_variants;
  e.n_tiles = b->n_tiles;
  return launch_tc_edge(e, sms, st);
}

// Output heads (type_out | coord_out, cspnet.py:388-401) in split precision: A = [hi | lo | hi] of
// the final LayerNorm (k_layernorm wrote hi | lo into cat16, the hi section is read twice),
// W image = s [w_hi | w_hi | w_lo] with a power-of-two scale s that keeps w_lo out of the fp16
// subnormals; the blob starts with 1/s.  hi w_hi + lo w_hi + hi w_lo = x w (1 + O(2^-22)).
int tc_head(const cb2_model *m, const __half *split16, int64_t VN, float *head_out, cudaStream_t st) {
  TcLinearArgs a{};
  a.A = split16; a.a_kt = H2; a.a_wrap = H2 / TL_KC; a.M = VN; a.K = 3 * H;
  a.Wt = (const __half *)m->w_head_t + 8; a.Nw = 256; a.out_scale = (const float *)m->w_head_t;
  a.C = head_out; a.ldc = HEADC; a.n_store = HEADC; a.bias = m->b_head;
  return launch_tc_linear(a, st);
}

int launch_tc_node2(const cb2_model *m, const cb2_layer_weights *Lmlp, const cb2_layer_weights *Lfilm,
                    const cb2_batch *b, const float *film_cond, float *h, float *h_rowmajor, const __half *h16,
                    __half *cat16, __half *P, int n_sm, cudaStream_t st);

bool tc_uses_node_chain(const cb2_model *m, const cb2_forward_io *io) {
  return io->film_cond != nullptr && !(m->flags & CB2_MODEL_NODE_UNFUSED);
}

// One CSPNet trunk pass on the tensor cores.  fp32 row-major: h (residual stream); fp16 row-major: P;
// fp16 row-panel: h16, cat16 = [LN(h) | agg], z16.
//
// Default (FiLM conditioning present): 2 launches per layer -- the edge kernel and ONE node-chain kernel
// k_tc_node2 per layer boundary (node MLP of layer l -> FiLM block, LayerNorm and hoist GEMM of layer l+1).
// cb2_model.flags & CB2_MODEL_NODE_UNFUSED (and forwards without FiLM) take round 1's four kernels per layer.
int tc_forward_layers(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, ForwardWs &w,
                      cudaStream_t st) {
  const int N = b->n_nodes, B = b->n_graphs, V = b->n_variants;
  const int64_t VN = (int64_t)V * N;
  if (!m->film_wp_t) return fail(CB2_ERR_BAD_ARG, "tensor-core path needs the fp16 operand images (pack with tensor_core=True)");
  for (int li = 0; li < m->n_layers; li++) {
    const cb2_layer_weights &L = m->layers[li];
    if (!L.w_hij_t || !L.w_fd_t || !L.w2_t || !L.wn1_t || !L.wn2_t)
      return fail(CB2_ERR_BAD_ARG, "tensor-core path: layer operand image missing");
  }
  const bool chain = tc_uses_node_chain(m, io);   // then h / h16 arrive in the panel layouts (k_embed_panels)
  if (io->film_cond != nullptr && !chain) CB2_TRY(launch_to_panels<float>(w.h, H, w.h16, VN, H, H, 0, st));
  int sms = 0;
  CB2_TRY(num_sms(&sms));
  CB2_TRY(launch_lattice_ip(io->lattices, m, w.cg, B, io->flags, st));      // all layers' lattice terms, one launch
  __half *P16 = reinterpret_cast<__half *>(w.P);
  if (chain) {
    {
      NvtxRange r("cb2:node_chain(head)");
      CB2_TRY(launch_tc_node2(m, nullptr, &m->layers[0], b, io->film_cond, w.h, nullptr, w.h16, w.cat16, P16, sms, st));
    }
    for (int li = 0; li < m->n_layers; li++) {
      const cb2_layer_weights &L = m->layers[li];
      NvtxRange r_layer("cb2:csp_layer");
      {
        NvtxRange r("cb2:edge");
        CB2_TRY(tc_edge_layer(m, L, b, io->frac_coords, P16, w.cg + (size_t)li * B * H, w.cat16, 0, H, H2, st));
      }
      NvtxRange r("cb2:node_chain");
      CB2_TRY(launch_tc_node2(m, &L, li + 1 < m->n_layers ? &m->layers[li + 1] : nullptr, b, io->film_cond, w.h,
                              w.P, w.h16, w.cat16, P16, sms, st));
    }
    w.h_final = w.P;          // mode TAIL wrote the final h row-major, fp32, over P (dead after the last edge kernel)
    return CB2_OK;
  }
  for (int li = 0; li < m->n_layers; li++) {
    const cb2_layer_weights &L = m->layers[li];
    NvtxRange r_layer("cb2:csp_layer");
    if (io->film_cond != nullptr) {
      // FiLM projection + LN + FiLM + SiLU + residual + layer LN in one kernel (y stays in TMEM)
      NvtxRange r("cb2:film");
      CB2_TRY(launch_tc_film(m, L, b, io->film_cond, w.h16, w.h, w.cat16, sms, st));
    } else {
      CB2_TRY(launch_film_apply(nullptr, w.h, nullptr, b->node2graph, m->film_g, m->film_b, L.ln_g, L.ln_b,
                                nullptr, 0, w.cat16, 0, H2, N, B, V, st));
    }
    {
      NvtxRange r("cb2:hoist");
      TcLinearArgs a{};   // P = hn [W_hi;W_hj]^T in fp16; the per-crystal lattice term stays in fp32 (cg)
      a.A = w.cat16; a.a_kt = H2; a.M = VN; a.K = H; a.Wt = (const __half *)L.w_hij_t; a.Nw = H2;
      a.C16r = P16; a.ldc16r = H2;
      CB2_TRY(launch_tc_linear(a, st));
    }
    {
      NvtxRange r("cb2:edge");
      CB2_TRY(tc_edge_layer(m, L, b, io->frac_coords, P16, w.cg + (size_t)li * B * H, w.cat16, 0, H, H2, st));
    }
    {
      NvtxRange r("cb2:node_mlp");
      TcLinearArgs a{};   // z = SiLU([hn|agg] Wn1^T + bn1)
      a.A = w.cat16; a.a_kt = H2; a.M = VN; a.K = H2; a.Wt = (const __half *)L.wn1_t; a.Nw = H;
      a.C16 = w.z16; a.c16_kt = H; a.bias = L.bn1; a.silu = 1;
      CB2_TRY(launch_tc_linear(a, st));
      TcLinearArgs a2{};  // h = h + SiLU(z Wn2^T + bn2)
      a2.A = w.z16; a2.a_kt = H; a2.M = VN; a2.K = H; a2.Wt = (const __half *)L.wn2_t; a2.Nw = H;
      a2.C = w.h; a2.ldc = H; a2.C16 = w.h16; a2.c16_kt = H; a2.bias = L.bn2; a2.silu = 1;
      a2.residual = w.h; a2.ldr = H;
      CB2_TRY(launch_tc_linear(a2, st));
    }
  }
  return CB2_OK;
}

}  // namespace cb2
