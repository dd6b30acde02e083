// k_tc_film: the FiLM block and the CSPLayer LayerNorm of one layer fused behind the FiLM
// projection GEMM (FilmLayer.forward + CSPLayer.layer_norm, cspnet.py:76-92,151):
//
//   y  = h16 Wp^T + bp                                  tcgen05, 128 rows x 512 columns in TMEM
//   f  = SiLU(LN_film(y) * scale[g] + shift[g])          per-crystal FiLM row from film_cond
//   h  = h + f                                           residual stream, fp32 row-major, in place
//   hn = LN_layer(h)  -> fp16, row-panel layout          A operand of the hoist GEMM / node MLP
//
// A 128 x 512 fp32 tile is exactly the TMEM of one SM, so a CTA owns whole rows and both
// LayerNorms are row-local: y never goes to HBM (the unfused pair wrote and re-read it:
// 0.67 GB of 1.67 GB per layer at C3).  Persistent, one CTA per SM:
//   warp 0   : loader (bulk copies: one 8 KB block of the row-panel A operand + one 32 KB
//              block of the weight image per K chunk of 32, 4-stage ring)
//   warp 1   : MMA issue (M128 N256 K16 x 2 column halves), TMEM alloc
//   warps 2-17: epilogue; warp = (lane quarter q, column group of 128).  TMEM hands every
//              thread one row, so the row statistics are in-thread sums plus one exchange
//              between the four warps of a quarter through shared memory:
//       pass 1  sum, sum of squares of y                 (TMEM read)
//       pass 2  f, h += f  (16-column blocks of h transposed through staging rows both ways so that
//               the global accesses are coalesced, the next block already in flight), h kept in
//               TMEM (tcgen05.st), its statistics
//       pass 3  LN_layer -> fp16 panel store             (thread = row is coalesced there)
// The accumulator is single-buffered (the tile fills TMEM), so the main loop of the next tile
// only overlaps the epilogue through the prefetched ring stages: measured 0.33 ms per launch at
// C3 against 0.12 + 0.26 ms for the unfused pair (ablation: 0.15 ms of it is the h / FiLM-row
// traffic of pass 2, the rest the serialised main loop + three TMEM passes).
#include "cb2_tc.cuh"

namespace cb2 {

using namespace ptx;

constexpr int TF_KC = 32, TF_STAGES = 4;
constexpr int TF_A_BYTES = 128 * TF_KC * 2;           // 8 KB
constexpr int TF_W_BYTES = H * TF_KC * 2;             // 32 KB: [4 k8][512 rows][16 B]
constexpr int TF_STAGE_BYTES = TF_A_BYTES + TF_W_BYTES;
constexpr int TF_PITCH = 20;                         // floats per staged row: 16 columns + 4 pad (conflict-free row reads)
constexpr int TF_STG_OFF = TF_STAGES * TF_STAGE_BYTES;
constexpr int TF_STG_BYTES = 16 * 32 * TF_PITCH * 4;  // one 32 x 16 fp32 block per epilogue warp
constexpr int TF_PAR_OFF = TF_STG_OFF + TF_STG_BYTES; // bias, g1, b1, g2, b2: 5 x 512 floats
constexpr int TF_PART_OFF = TF_PAR_OFF + 5 * H * 4;   // [2 passes][4 column groups][128 rows] float2
constexpr int TF_BAR_OFF = TF_PART_OFF + 2 * 4 * 128 * 8;
constexpr int TF_SMEM = TF_BAR_OFF + 128;
constexpr int TF_THREADS = 32 * 18;
static_assert(TF_SMEM <= 232448, "shared memory budget");

// streaming 16-byte load: the residual stream passes through once, keep L1 for the FiLM rows
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

struct TcFilmArgs {
  const __half *A;       // h16, row-panel layout with 512 columns
  int64_t M;             // V * N rows
  const __half *Wt;      // FiLM projection image [64 k8][512][8]
  const float *bias;     // film_bp
  const float *g1, *b1;  // FilmLayer.norm
  const float *g2, *b2;  // CSPLayer.layer_norm
  const float *cond;     // [V*B,1024]: scale | shift
  const int32_t *node2graph;
  int N, B;
  float *h;              // [M,512] residual stream, updated in place
  __half *hn16;          // row-panel output with hn_kt columns per panel, written at columns 0:512
  int hn_kt;
};

__global__ void __launch_bounds__(TF_THREADS, 1) k_tc_film(TcFilmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + TF_BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 32 + 8 * s; };
  const uint32_t acc_full = bar_base + 64, acc_free = bar_base + 72;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TF_BAR_OFF + 96);
  float *par = reinterpret_cast<float *>(smem + TF_PAR_OFF);
  float2 *part = reinterpret_cast<float2 *>(smem + TF_PART_OFF);

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int n_tiles = (int)((g.M + 127) / 128);
  constexpr int nk = H / TF_KC;   // 16 chunks

  if (tid == 0) {
    for (int s = 0; s < TF_STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);
    fence_barrier_init();
  }
  for (int i = tid; i < H; i += TF_THREADS) {
    par[i] = g.bias[i];
    par[H + i] = g.g1[i];
    par[2 * H + i] = g.b1[i];
    par[3 * H + i] = g.g2[i];
    par[4 * H + i] = g.b2[i];
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
        const __half *ap = g.A + (int64_t)t * 128 * H;
        for (int kc = 0; kc < nk; kc++, it++) {
          const int s = it % TF_STAGES;
          mbar_wait_spin(empty_bar(s), ((it / TF_STAGES) & 1) ^ 1);
          const uint32_t a_s = sbase + s * TF_STAGE_BYTES, w_s = a_s + TF_A_BYTES;
          mbar_arrive_expect_tx(full_bar(s), TF_STAGE_BYTES);
          bulk_g2s(a_s, ap + (int64_t)kc * (TF_KC / 8) * 1024, TF_A_BYTES, full_bar(s));
          bulk_g2s(w_s, g.Wt + (int64_t)kc * (TF_KC / 8) * H * 8, TF_W_BYTES, full_bar(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16_f32(128, 256);
      int it = 0, tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
        mbar_wait_spin(acc_free, (tl & 1) ^ 1);        // the epilogue has drained the previous tile
        tc_fence_after_sync();
        for (int kc = 0; kc < nk; kc++, it++) {
          const int s = it % TF_STAGES;
          mbar_wait_spin(full_bar(s), (it / TF_STAGES) & 1);
          tc_fence_after_sync();
          const uint32_t a_s = sbase + s * TF_STAGE_BYTES, w_s = a_s + TF_A_BYTES;
#pragma unroll
          for (int j = 0; j < TF_KC / 16; j++) {
            const uint64_t ad = smem_desc_kmajor(a_s + 2 * j * 2048, 2048, 128);
#pragma unroll
            for (int nh = 0; nh < 2; nh++) {
              const uint64_t bd = smem_desc_kmajor(w_s + 2 * j * (H * 16) + nh * 256 * 16, H * 16, 128);
              umma_f16(tmem + nh * 256, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (16 warps) ----------------
    const int q = warp & 3, cgp = (warp - 2) >> 2;
    const int c0 = cgp * 128;                                      // first column of this warp
    const int row = q * 32 + lane;                                 // row of the tile owned by this thread
    float *stg = reinterpret_cast<float *>(smem + TF_STG_OFF) + (warp - 2) * (32 * TF_PITCH);
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + c0;
    const float *p_bias = par, *p_g1 = par + H, *p_b1 = par + 2 * H, *p_g2 = par + 3 * H, *p_b2 = par + 4 * H;
    const int bar_id = 1 + q;                                      // the four warps that share this quarter's rows
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      const int64_t r0 = (int64_t)t * 128 + q * 32;                // first row of this warp
      const int nvalid = (int)(g.M - r0 < 32 ? (g.M - r0 < 0 ? 0 : g.M - r0) : 32);
      const int64_t grow = r0 + lane;
      const float *cs = g.cond;
      if (lane < nvalid) cs = g.cond + ((grow / g.N) * g.B + g.node2graph[(int)(grow % g.N)]) * H2;
      // while the main loop of this tile runs: pull this warp's block of h (32 rows x 512 B) and the
      // FiLM rows of its crystals towards L2, so that pass 2 does not wait on DRAM
      if (lane < nvalid) {
        const char *hp = reinterpret_cast<const char *>(g.h + grow * H + c0);
#pragma unroll
        for (int l = 0; l < 4; l++) prefetch_l2(hp + l * 128);
        const char *cp = reinterpret_cast<const char *>(cs + c0);
#pragma unroll
        for (int l = 0; l < 4; l++) {
          prefetch_l2(cp + l * 128);
          prefetch_l2(cp + H * 4 + l * 128);
        }
      }
      mbar_wait(acc_full, tl & 1);
      tc_fence_after_sync();

      // ---- pass 1: statistics of y = acc + bias over the row ----
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
          for (int j = 0; j < 16; j++) {
            const float y = __uint_as_float(acc[j]) + p_bias[c0 + hb * 16 + j];
            s += y;
            ss = fmaf(y, y, ss);
          }
        }
      }
      part[cgp * 128 + row] = make_float2(s, ss);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      float mean1, rstd1;
      {
        float ts = 0.f, tss = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float2 v = part[k * 128 + row];
          ts += v.x;
          tss += v.y;
        }
        mean1 = ts * (1.0f / H);
        rstd1 = rsqrtf(fmaxf(tss * (1.0f / H) - mean1 * mean1, 0.f) + 1e-5f);
      }

      // ---- pass 2: f = SiLU(LN(y) scale + shift), h += f (kept in TMEM), statistics of the new h ----
      // 16-column blocks; the h block of the next iteration is already in flight (registers) while
      // this one is processed: 8 rows x 64 B per load instruction, transposed through the staging rows
      s = 0.f;
      ss = 0.f;
      {
        const int prow = lane >> 2, pcol = (lane & 3) * 4;           // coalesced pass: 8 rows x 64 B per instruction
        const float *hsrc = g.h + (r0 + prow) * H + c0 + pcol;
        float *hdst = g.h + (r0 + prow) * H + c0 + pcol;
        float4 hnext[4];
        auto load_h = [&](int blk) {
#pragma unroll
          for (int itr = 0; itr < 4; itr++) {
            hnext[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (itr * 8 + prow < nvalid) hnext[itr] = ld_stream_f4(hsrc + (int64_t)itr * 8 * H + blk * 16);
          }
        };
        load_h(0);
#pragma unroll 1
        for (int blk = 0; blk < 8; blk++) {
          const int col0 = c0 + blk * 16;
          uint32_t acc[16];
          tmem_ld16(taddr + blk * 16, acc);
          __syncwarp();                                            // the previous block's stores have drained the staging rows
#pragma unroll
          for (int itr = 0; itr < 4; itr++)
            *reinterpret_cast<float4 *>(stg + (itr * 8 + prow) * TF_PITCH + pcol) = hnext[itr];
          __syncwarp();
          if (blk < 7) load_h(blk + 1);
          float4 sc[4], sh[4];                                      // FiLM row of this thread's crystal (L1-resident)
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            sc[j4] = __ldg(reinterpret_cast<const float4 *>(cs + col0 + 4 * j4));
            sh[j4] = __ldg(reinterpret_cast<const float4 *>(cs + H + col0 + 4 * j4));
          }
          float hv[16];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float4 v = *reinterpret_cast<const float4 *>(stg + lane * TF_PITCH + 4 * j);
            hv[4 * j] = v.x; hv[4 * j + 1] = v.y; hv[4 * j + 2] = v.z; hv[4 * j + 3] = v.w;
          }
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float scs[4] = {sc[j4].x, sc[j4].y, sc[j4].z, sc[j4].w}, shs[4] = {sh[j4].x, sh[j4].y, sh[j4].z, sh[j4].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int j = 4 * j4 + k, c = col0 + j;
              const float y = __uint_as_float(acc[j]) + p_bias[c];
              const float ln = fmaf((y - mean1) * rstd1, p_g1[c], p_b1[c]);
              const float hn = hv[j] + silu_fast(fmaf(ln, scs[k], shs[k]));
              s += hn;
              ss = fmaf(hn, hn, ss);
              acc[j] = __float_as_uint(hn);
            }
          }
          tmem_st16(taddr + blk * 16, acc);
          // new h block: one row per thread -> staging -> coalesced global stores
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint4 *>(stg + lane * TF_PITCH + 4 * j) = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int itr = 0; itr < 4; itr++)
            if (itr * 8 + prow < nvalid)
              *reinterpret_cast<float4 *>(hdst + (int64_t)itr * 8 * H + blk * 16) =
                  *reinterpret_cast<const float4 *>(stg + (itr * 8 + prow) * TF_PITCH + pcol);
        }
      }
      tmem_st_wait();
      part[512 + cgp * 128 + row] = make_float2(s, ss);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      float mean2, rstd2;
      {
        float ts = 0.f, tss = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float2 v = part[512 + k * 128 + row];
          ts += v.x;
          tss += v.y;
        }
        mean2 = ts * (1.0f / H);
        rstd2 = rsqrtf(fmaxf(tss * (1.0f / H) - mean2 * mean2, 0.f) + 1e-5f);
      }

      // ---- pass 3: hn = LN_layer(h) -> fp16, row-panel layout (16 B per thread and column group) ----
      {
        __half *dst = g.hn16 + (int64_t)t * 128 * g.hn_kt + (int64_t)(c0 >> 3) * 1024 + row * 8;
        uint32_t accA[16], accB[16];
        tmem_ld16(taddr, accA);
#pragma unroll
        for (int hb = 0; hb < 8; hb++) {
          tmem_ld_wait();
          uint32_t (&acc)[16] = (hb & 1) ? accB : accA;
          uint32_t (&nxt)[16] = (hb & 1) ? accA : accB;
          if (hb < 7) tmem_ld16(taddr + (hb + 1) * 16, nxt);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int c = c0 + hb * 16 + j;
            const float a = fmaf((__uint_as_float(acc[j]) - mean2) * rstd2, p_g2[c], p_b2[c]);
            const float b = fmaf((__uint_as_float(acc[j + 1]) - mean2) * rstd2, p_g2[c + 1], p_b2[c + 1]);
            w[j / 2] = pack_half2(a, b);
          }
          if (lane < nvalid) {
            *reinterpret_cast<uint4 *>(dst + (hb * 2) * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(dst + (hb * 2 + 1) * 1024) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int launch_tc_film(const cb2_model *m, const cb2_layer_weights &L, const cb2_batch *b, const float *film_cond,
                   const __half *h16, float *h, __half *cat16, int n_sm, cudaStream_t st) {
  const int64_t VN = (int64_t)b->n_variants * b->n_nodes;
  if (VN == 0) return CB2_OK;
  CB2_CUDA_OK(cudaFuncSetAttribute(k_tc_film, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));   // per device
  TcFilmArgs a{};
  a.A = h16; a.M = VN; a.Wt = (const __half *)m->film_wp_t; a.bias = m->film_bp;
  a.g1 = m->film_g; a.b1 = m->film_b; a.g2 = L.ln_g; a.b2 = L.ln_b;
  a.cond = film_cond; a.node2graph = b->node2graph; a.N = b->n_nodes; a.B = b->n_graphs;
  a.h = h; a.hn16 = cat16; a.hn_kt = H2;
  const int64_t n_tiles = (VN + 127) / 128;
  k_tc_film<<<(unsigned)(n_tiles < n_sm ? n_tiles : n_sm), TF_THREADS, TF_SMEM, st>>>(a);
  CB2_LAUNCH_OK("k_tc_film");
  return CB2_OK;
}

}  // namespace cb2
