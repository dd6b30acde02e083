// Per-timestep state update of the reverse diffusion (HBM-bound, fused elementwise).
//
//   predictor: CFG mix                         chemeleon.py:288-290
//              D3PM.p_logits (absorbing)       diff_utils.py:307-329, 258-286, 215-234
//              lattice DDPM ancestral step     chemeleon.py:413-425
//              coordinate predictor            chemeleon.py:427-437
//   corrector: Langevin step + mod-1 wrap      chemeleon.py:452-463
//
// The D3PM posterior uses the closed form of the absorbing-state matrices
// (Q_s = (1-b) I + b 1 e0^T, Qbar_s = diag I + off 1 e0^T): no [T,104,104] table.
// Noise is either read from injected tensors (parity mode; the reference's draw
// order) or generated in-kernel with Philox4x32-10 keyed by
// (seed, global sample id, atom, timestep, stream) so that results do not depend
// on how samples are sharded over GPUs.
#include "cb2_internal.cuh"

namespace cb2 {

// ---- Philox4x32-10 ----------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float normal_from(uint32_t a, uint32_t b) {
  float u1 = u01(a), u2 = u01(b);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ U4 noise_block(uint64_t seed, int64_t gid, int atom, int t, int stream, int sub) {
  uint64_t key = (uint64_t)gid * 1024ull + (uint64_t)atom;
  U4 c;
  c.x = (uint32_t)key;
  c.y = (uint32_t)(key >> 32);
  c.z = (uint32_t)t;
  c.w = (uint32_t)(stream * 64 + sub);
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

struct UpdateParams {
  int N, B, V, T;
  const int32_t *node2graph;
  const int32_t *node_base;
  int64_t *a;
  float *x;
  float *l;
  const int32_t *t_dev;
  int32_t *flags;
  const float *coef;
  float cs, one_minus_cs;
  int noise_mode, t_start;
  const float *rand_a, *rand_l, *rand_x, *rand_x2;
  uint64_t seed;
  const uint64_t *seed_dev;
  const int64_t *graph_gid;
  const float *head_out;
  const float *lat_out;
};

__device__ __forceinline__ float cfg_mix(const UpdateParams &p, float cond, float null_) {
  // (1 - s) * null + s * cond, each product rounded (torch does not contract)
  return __fadd_rn(__fmul_rn(p.one_minus_cs, null_), __fmul_rn(p.cs, cond));
}

// One warp per node (types + coordinate predictor), then one warp per crystal (lattice).
__global__ void __launch_bounds__(256) k_update_predictor(UpdateParams p) {
  const int64_t w = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int t = *p.t_dev;
  const float *cf = p.coef + (int64_t)t * CB2_COEF_COLS;
  const int s = p.t_start - t;
  if (w < p.N) {
    const int n = (int)w;
    const int g = p.node2graph[n];
    const int atom = n - p.node_base[n];
    // ---- atom types: lane owns classes 4*lane .. 4*lane+3 (lanes 0..25) ----
    const bool act = lane < NTYPE / 4;
    float lg[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (act) {
      float4 c4 = *reinterpret_cast<const float4 *>(p.head_out + (int64_t)n * HEADC + 4 * lane);
      if (p.V == 2) {
        float4 n4 = *reinterpret_cast<const float4 *>(p.head_out + ((int64_t)p.N + n) * HEADC + 4 * lane);
        lg[0] = cfg_mix(p, c4.x, n4.x); lg[1] = cfg_mix(p, c4.y, n4.y);
        lg[2] = cfg_mix(p, c4.z, n4.z); lg[3] = cfg_mix(p, c4.w, n4.w);
      } else {
        lg[0] = c4.x; lg[1] = c4.y; lg[2] = c4.z; lg[3] = c4.w;
      }
    }
    const long long at = p.a[n];
    float val[4];
    if (t == 1) {
#pragma unroll
      for (int q = 0; q < 4; q++) val[q] = lg[q];
    } else {
      float m = fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float pr[4], ssum = 0.f;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        pr[q] = act ? expf(lg[q] - m) : 0.f;
        ssum += pr[q];
      }
      ssum = warp_sum(ssum);
      float rest = 0.f;  // sum of p over classes != 0
#pragma unroll
      for (int q = 0; q < 4; q++) {
        pr[q] = pr[q] / ssum;
        if (!(lane == 0 && q == 0)) rest += pr[q];
      }
      rest = warp_sum(rest);
      const float beta = cf[8], omb = cf[9], q00 = cf[10], diag = cf[11], off = cf[12], qb00 = cf[13];
      float u4[4] = {1.f, 1.f, 1.f, 1.f};
      if (act) {
        if (p.noise_mode == 0) {
          float4 u = *reinterpret_cast<const float4 *>(p.rand_a + ((int64_t)s * p.N + n) * NTYPE + 4 * lane);
          u4[0] = u.x; u4[1] = u.y; u4[2] = u.z; u4[3] = u.w;
        } else {
          U4 r = noise_block(p.seed_dev ? *p.seed_dev : p.seed, p.graph_gid[g], atom, t, 0, lane);
          u4[0] = u01(r.x); u4[1] = u01(r.y); u4[2] = u01(r.z); u4[3] = u01(r.w);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        int c = 4 * lane + q;
        float f1 = (at == 0) ? (c == 0 ? q00 : beta) : (c == at ? omb : 0.f);
        float f2 = (c == 0) ? __fadd_rn(__fmul_rn(pr[q], qb00), __fmul_rn(rest, off)) : __fmul_rn(pr[q], diag);
        float o = logf(f1 + 1.0e-6f) + logf(f2 + 1.0e-6f);
        float u = fminf(fmaxf(u4[q], 1.0e-6f), 1.0f);
        float gum = -logf(-logf(u));
        val[q] = act ? (o + gum) : -INFINITY;
      }
    }
    // argmax, first maximum wins (torch.argmax)
    float best = val[0];
    int bi = 4 * lane;
#pragma unroll
    for (int q = 1; q < 4; q++)
      if (val[q] > best) { best = val[q]; bi = 4 * lane + q; }
    if (!act) { best = -INFINITY; bi = 1 << 20; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      // NaN-safe: a NaN never wins unless everything is NaN
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) p.a[n] = (bi < NTYPE) ? bi : 0;
    // ---- coordinate predictor: lanes 0..2 ----
    if (lane < 3) {
      float pc = p.head_out[(int64_t)n * HEADC + NTYPE + lane];
      float px = pc;
      if (p.V == 2) px = cfg_mix(p, pc, p.head_out[((int64_t)p.N + n) * HEADC + NTYPE + lane]);
      float z = 0.f;
      if (t > 1) {
        if (p.noise_mode == 0) z = p.rand_x[((int64_t)s * p.N + n) * 3 + lane];
        else { U4 r = noise_block(p.seed_dev ? *p.seed_dev : p.seed, p.graph_gid[g], atom, t, 2, lane); z = normal_from(r.x, r.y); }
      }
      float pxs = __fmul_rn(px, cf[5]);
      float xo = p.x[(int64_t)n * 3 + lane];
      float xn = __fadd_rn(__fsub_rn(xo, __fmul_rn(cf[3], pxs)), __fmul_rn(cf[4], z));
      p.x[(int64_t)n * 3 + lane] = xn;
      if (!isfinite(xn)) atomicOr(p.flags + g, CB2_FLAG_NONFINITE);
    }
  } else if (w < (int64_t)p.N + p.B) {
    // ---- lattice ancestral step: lanes 0..8 ----
    const int g = (int)(w - p.N);
    if (lane < 9) {
      const bool mask = !(lane == 1 || lane == 6 || lane == 7);  // [[1,0,1],[1,1,1],[0,0,1]]
      float plc = p.lat_out[(int64_t)g * 9 + lane];
      float pl = plc;
      if (p.V == 2) pl = cfg_mix(p, plc, p.lat_out[((int64_t)p.B + g) * 9 + lane]);
      float z = 0.f;
      if (t > 1) {
        if (p.noise_mode == 0) z = p.rand_l[((int64_t)s * p.B + g) * 9 + lane];
        else { U4 r = noise_block(p.seed_dev ? *p.seed_dev : p.seed, p.graph_gid[g], 1023, t, 1, lane); z = normal_from(r.x, r.y); }
      }
      z = mask ? z : 0.f;
      float lo = p.l[(int64_t)g * 9 + lane];
      float ln = __fadd_rn(__fmul_rn(cf[0], __fsub_rn(lo, __fmul_rn(cf[1], pl))), __fmul_rn(cf[2], z));
      ln = mask ? ln : 0.f;
      if (t == p.T) ln = fminf(fmaxf(ln, -6.0f), 6.0f);
      p.l[(int64_t)g * 9 + lane] = ln;
      if (!isfinite(ln)) atomicOr(p.flags + g, CB2_FLAG_NONFINITE);
    }
  }
}

__global__ void __launch_bounds__(256) k_update_corrector(UpdateParams p) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.N * 3) return;
  const int n = (int)(idx / 3), d = (int)(idx % 3);
  const int t = *p.t_dev;
  const float *cf = p.coef + (int64_t)t * CB2_COEF_COLS;
  const int s = p.t_start - t;
  float pc = p.head_out[(int64_t)n * HEADC + NTYPE + d];
  float px = pc;
  if (p.V == 2) px = cfg_mix(p, pc, p.head_out[((int64_t)p.N + n) * HEADC + NTYPE + d]);
  float z = 0.f;
  if (t > 1) {
    if (p.noise_mode == 0) z = p.rand_x2[((int64_t)s * p.N + n) * 3 + d];
    else {
      int g = p.node2graph[n];
      U4 r = noise_block(p.seed_dev ? *p.seed_dev : p.seed, p.graph_gid[g], n - p.node_base[n], t, 3, d);
      z = normal_from(r.x, r.y);
    }
  }
  float pxs = __fmul_rn(px, cf[5]);
  float xh = p.x[idx];
  float xn = __fadd_rn(__fsub_rn(xh, __fmul_rn(cf[6], pxs)), __fmul_rn(cf[7], z));
  xn = wrap01(xn);
  p.x[idx] = xn;
  if (!isfinite(xn)) atomicOr(p.flags + p.node2graph[n], CB2_FLAG_NONFINITE);
}

__global__ void k_advance_t(int32_t *t_dev) { *t_dev = *t_dev - 1; }

static UpdateParams make_params(const cb2_batch *b, cb2_state *s, const cb2_step_args *a) {
  const int T = a->timesteps;
  UpdateParams p;
  p.N = b->n_nodes; p.B = b->n_graphs; p.V = b->n_variants; p.T = T;
  p.node2graph = b->node2graph; p.node_base = b->node_base;
  p.a = s->atom_types; p.x = s->frac_coords; p.l = s->lattices; p.t_dev = s->t_dev; p.flags = s->flags;
  p.coef = a->coef;
  p.cs = a->cond_scale;
  p.one_minus_cs = (float)(1.0 - (double)a->cond_scale);
  p.noise_mode = a->noise_mode; p.t_start = a->t_start;
  p.rand_a = a->rand_a; p.rand_l = a->rand_l; p.rand_x = a->rand_x; p.rand_x2 = a->rand_x2;
  p.seed = a->seed; p.seed_dev = a->seed_dev; p.graph_gid = a->graph_gid;
  p.head_out = nullptr; p.lat_out = nullptr;
  return p;
}

static int check_update_args(const cb2_batch *b, const cb2_state *s, const cb2_step_args *a) {
  if (!b || !s || !a) return fail(CB2_ERR_BAD_ARG, "update: null argument");
  if (b->n_variants != 1 && b->n_variants != 2) return fail(CB2_ERR_BAD_ARG, "update: n_variants must be 1 or 2");
  if (!s->atom_types || !s->frac_coords || !s->lattices || !s->t_dev || !s->flags || !a->coef)
    return fail(CB2_ERR_BAD_ARG, "update: null state/coef pointer");
  if (a->noise_mode == 0 && (!a->rand_a || !a->rand_l || !a->rand_x || !a->rand_x2))
    return fail(CB2_ERR_BAD_ARG, "update: noise_mode=0 needs the four injected noise tensors");
  if (a->noise_mode == 1 && !a->graph_gid) return fail(CB2_ERR_BAD_ARG, "update: noise_mode=1 needs graph_gid");
  if (a->noise_mode != 0 && a->noise_mode != 1) return fail(CB2_ERR_BAD_ARG, "update: bad noise_mode");
  return CB2_OK;
}

int update_predictor(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                     const float *lat_out, cudaStream_t st) {
  CB2_TRY(check_update_args(b, s, a));
  if (!head_out || !lat_out) return fail(CB2_ERR_BAD_ARG, "update_predictor: null decoder outputs");
  UpdateParams p = make_params(b, s, a);
  p.head_out = head_out; p.lat_out = lat_out;
  int64_t warps = (int64_t)b->n_nodes + b->n_graphs;
  if (warps == 0) return CB2_OK;
  k_update_predictor<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(p);
  CB2_LAUNCH_OK("k_update_predictor");
  return CB2_OK;
}

int update_corrector(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                     cudaStream_t st) {
  CB2_TRY(check_update_args(b, s, a));
  if (!head_out) return fail(CB2_ERR_BAD_ARG, "update_corrector: null decoder outputs");
  UpdateParams p = make_params(b, s, a);
  p.head_out = head_out;
  int64_t total = (int64_t)b->n_nodes * 3;
  if (total > 0) {
    k_update_corrector<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    CB2_LAUNCH_OK("k_update_corrector");
  }
  k_advance_t<<<1, 1, 0, st>>>(s->t_dev);
  CB2_LAUNCH_OK("k_advance_t");
  return CB2_OK;
}

// ---- streaming frames (TrajectoryContainer.get_atoms input, schema.py:57-68, as a compact wire format) ----
// frame = { int32 t; int32 n_nodes; int32 n_graphs; int32 reserved;
//           uint8 types[N] (values > 103 -> 0, schema.py:60-62), zero-padded to a multiple of 4;
//           float coords[N][3]; float lattice[B][9] }
__global__ void __launch_bounds__(256) k_pack_frame(const int64_t *__restrict__ a, const float *__restrict__ x,
                                                    const float *__restrict__ l, const int32_t *__restrict__ t_dev,
                                                    int N, int B, uint8_t *__restrict__ frame) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Np = (N + 3) & ~3;
  int32_t *hdr = reinterpret_cast<int32_t *>(frame);
  uint8_t *types = frame + 16;
  float *coords = reinterpret_cast<float *>(frame + 16 + Np);
  float *lat = coords + (int64_t)N * 3;
  if (idx == 0) { hdr[0] = *t_dev; hdr[1] = N; hdr[2] = B; hdr[3] = 0; }
  if (idx < Np) {
    long long v = idx < N ? a[idx] : 0;
    types[idx] = (uint8_t)((v < 0 || v > 103) ? 0 : v);
  }
  if (idx < (int64_t)N * 3) coords[idx] = x[idx];
  if (idx < (int64_t)B * 9) lat[idx] = l[idx];
}

size_t frame_bytes(int N, int B) { return 16 + (size_t)((N + 3) & ~3) + (size_t)N * 12 + (size_t)B * 36; }

int pack_frame(const cb2_batch *b, const cb2_state *s, void *frame, size_t bytes, cudaStream_t st) {
  if (!b || !s || !frame || !s->atom_types || !s->frac_coords || !s->lattices || !s->t_dev)
    return fail(CB2_ERR_BAD_ARG, "pack_frame: null argument");
  if (bytes < frame_bytes(b->n_nodes, b->n_graphs)) return fail(CB2_ERR_WORKSPACE, "pack_frame: frame buffer too small");
  const int64_t work = (int64_t)b->n_nodes * 3 > (int64_t)b->n_graphs * 9 ? (int64_t)b->n_nodes * 3 : (int64_t)b->n_graphs * 9;
  k_pack_frame<<<(unsigned)((work + 4 + 255) / 256), 256, 0, st>>>(s->atom_types, s->frac_coords, s->lattices, s->t_dev,
                                                                  b->n_nodes, b->n_graphs, (uint8_t *)frame);
  CB2_LAUNCH_OK("k_pack_frame");
  return CB2_OK;
}

}  // namespace cb2
