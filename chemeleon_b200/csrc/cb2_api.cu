// extern "C" entry points of libchemeleon_b200.so and the native orchestration of
// one decoder forward / one sampler timestep (see include/chemeleon_b200.h).
#include "cb2_internal.cuh"

#include <atomic>
#include <mutex>

namespace cb2 {

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

void set_error(const std::string &msg) { g_err = msg; }
int fail(cb2_status code, const std::string &msg) {
  g_err = msg;
  return (int)code;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// launchers from the other translation units
int launch_embed(const int64_t *a, const float *emb, float *h, int N, int V, cudaStream_t st);
int launch_embed_panels(const int64_t *a, const float *emb, float *hp, __half *h16, int N, int V, cudaStream_t st);
bool tc_uses_node_chain(const cb2_model *m, const cb2_forward_io *io);
int launch_film_cond(const float *time_table, const float *text_part, const int32_t *text_row, const int32_t *t_dev,
                     float *out, int64_t rows, cudaStream_t st);
int launch_ln_gelu(float *x, const float *g, const float *b, int width, int64_t rows, cudaStream_t st);
int launch_film_apply(const float *y, float *h, const float *cond, const int32_t *node2graph, const float *fg,
                      const float *fb, const float *cg, const float *cb, float *hn, int64_t ld_hn, __half *hn16,
                      int64_t ld_hn16, int hn16_kt, int N, int B, int V, cudaStream_t st);
int launch_layernorm(const float *x, const float *g, const float *b, float *out, __half *split16, const float *w3,
                     float *out3, int64_t ld3, int64_t rows, cudaStream_t st);
int launch_validity(const int64_t *a, const float *x, const float *lat, const int32_t *graph_off, int B,
                    const int32_t *target, float max_len, float thr, int32_t *flags, float *min_dist,
                    float *max_abc, cudaStream_t st);
int tc_head(const cb2_model *m, const __half *split16, int64_t VN, float *head_out, cudaStream_t st);
int launch_lattice_ip(const float *lat, const cb2_model *m, float *cg, int B, int32_t *range_flags, cudaStream_t st);
int launch_edge_embed(const float *x, const int32_t *ei, const int32_t *ej, float *emb, int64_t n_rows,
                      cudaStream_t st);
int launch_segment_mean(const float *e, const int64_t *node_eoff, const int32_t *node_n, float *out,
                        int64_t ld_out, int node_lo, int node_hi, int64_t e0, int64_t out_row_off,
                        cudaStream_t st);
int launch_lattice_head(const float *hf, const float *w_lat, const float *lat, const int32_t *graph_off,
                        float *out, int N, int B, int V, cudaStream_t st);
int update_predictor(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                     const float *lat_out, cudaStream_t st);
int update_corrector(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                     cudaStream_t st);
size_t frame_bytes(int N, int B);
int pack_frame(const cb2_batch *b, const cb2_state *s, void *frame, size_t bytes, cudaStream_t st);
// tensor-core path (cb2_tc.cu)
int tc_forward_layers(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, ForwardWs &w,
                      cudaStream_t st);
int tc_linear_simple(const void *A16, int64_t lda, const void *Wt, int Nw, const float *bias, float *C,
                     int64_t ldc, int64_t M, int K, int silu, void *workspace, size_t workspace_bytes, cudaStream_t st);
int debug_edge_timeline(long long *out96);
int debug_edge2_timeline(long long *out96x3);
int debug_node2_timeline(long long *out64x3);
int tc_edge_layer(const cb2_model *m, const cb2_layer_weights &L, const cb2_batch *b, const float *x, const __half *P,
                  const float *cg, __half *agg16, int64_t ld_agg, int agg_col, int agg_kt, cudaStream_t st);

size_t carve_forward(Arena &a, const cb2_batch *b, int n_layers, int precision, ForwardWs &w) {
  const size_t VN = (size_t)b->n_variants * b->n_nodes;
  w = ForwardWs{};
  w.h = a.take<float>((precision == CB2_PRECISION_FP32 ? VN : (VN + 127) / 128 * 128) * H);
  w.hf = a.take<float>(VN * H);
  w.cg = a.take<float>((size_t)n_layers * b->n_graphs * H);
  if (precision == CB2_PRECISION_FP32) {
    w.y = a.take<float>(VN * H);
    w.P = a.take<float>(VN * H2);
    w.cat = a.take<float>(VN * H2);
    w.z1 = a.take<float>(VN * H);
    const size_t ec = (size_t)b->chunk_max_edges;
    w.emb = a.take<float>(ec * DIS);
    w.a1 = a.take<float>(ec * H);
    w.e2 = a.take<float>(ec * H);
  } else {
    // tensor-core path: the FiLM projection y never leaves TMEM and P is fp16
    const size_t VNp = (VN + 127) / 128 * 128;   // row-panel layout: whole panels of 128 rows
    w.P = reinterpret_cast<float *>(a.take<__half>(VN * H2));
    w.h16 = a.take<__half>(VNp * H);
    w.cat16 = a.take<__half>(VNp * H2);
    w.z16 = a.take<__half>(VNp * H);
  }
  return a.off;
}

size_t carve_step(Arena &a, const cb2_batch *b, StepWs &w) {
  const size_t VN = (size_t)b->n_variants * b->n_nodes;
  const size_t VB = (size_t)b->n_variants * b->n_graphs;
  w.film_cond = a.take<float>(VB * H2);
  w.head_out = a.take<float>(VN * HEADC);
  w.lat_out = a.take<float>(VB * 9);
  return a.off;
}

static int check_model(const cb2_model *m) {
  if (!m) return fail(CB2_ERR_BAD_ARG, "null model");
  if (m->abi_version != CB2_ABI_VERSION) return fail(CB2_ERR_BAD_ARG, "cb2_model.abi_version mismatch");
  if (m->hidden != H || m->n_atom_types != NTYPE || m->n_freqs != NFREQ)
    return fail(CB2_ERR_UNSUPPORTED,
                "kernels are built for hidden=512, 104 atom types, 128 frequencies only (no generic fallback)");
  if (m->n_layers < 1 || m->n_layers > CB2_MAX_LAYERS) return fail(CB2_ERR_UNSUPPORTED, "n_layers out of range");
  return CB2_OK;
}

static int check_batch(const cb2_batch *b, int precision) {
  if (!b) return fail(CB2_ERR_BAD_ARG, "null batch");
  if (b->n_variants != 1 && b->n_variants != 2) return fail(CB2_ERR_BAD_ARG, "n_variants must be 1 or 2");
  if (b->n_nodes < 0 || b->n_graphs < 0) return fail(CB2_ERR_BAD_ARG, "negative sizes");
  if (b->n_nodes > 0 && (!b->node2graph || !b->node_base || !b->node_n || !b->graph_off))
    return fail(CB2_ERR_BAD_ARG, "null topology pointer");
  if (precision == CB2_PRECISION_FP32) {
    if (b->n_nodes > 0 && (!b->edge_i || !b->edge_j || !b->node_eoff || !b->host_chunk_node_lo ||
                           !b->host_chunk_edge_lo || b->n_chunks < 1))
      return fail(CB2_ERR_BAD_ARG, "exact path needs edge_i/edge_j/node_eoff and the host chunk tables");
  } else if (precision == CB2_PRECISION_TC_F16) {
    if (b->n_nodes > 0 && (!b->tile_row_i || !b->tile_row_j || !b->tile_seg_n || b->n_tiles < 1))
      return fail(CB2_ERR_BAD_ARG, "tensor-core path needs the tile tables");
    if (b->max_n > CB2_TILE_ROWS)
      return fail(CB2_ERR_UNSUPPORTED, "tensor-core path supports crystals of at most 128 atoms");
  } else {
    return fail(CB2_ERR_BAD_ARG, "unknown precision");
  }
  return CB2_OK;
}

// ---- exact-mode edge model of one layer: emb -> GEMM1(+P gather, SiLU) -> GEMM2 -> segment mean ----
static int f32_edge_layer(const cb2_layer_weights &L, const cb2_batch *b, const float *x, const float *P,
                          const float *cg, float *agg, int64_t ld_agg, ForwardWs &w, cudaStream_t st) {
  const int N = b->n_nodes, V = b->n_variants;
  for (int c = 0; c < b->n_chunks; c++) {
    const int nlo = b->host_chunk_node_lo[c], nhi = b->host_chunk_node_lo[c + 1];
    const int64_t e0 = b->host_chunk_edge_lo[c], e1 = b->host_chunk_edge_lo[c + 1];
    const int64_t rows = e1 - e0;
    if (rows <= 0) continue;
    if (rows > b->chunk_max_edges) return fail(CB2_ERR_BAD_ARG, "chunk larger than chunk_max_edges");
    CB2_TRY(launch_edge_embed(x, b->edge_i + e0, b->edge_j + e0, w.emb, rows, st));
    for (int v = 0; v < V; v++) {
      GemmEpilogue e1p;
      e1p.P = P; e1p.ei = b->edge_i + e0; e1p.ej = b->edge_j + e0; e1p.prow_off = (int64_t)v * N;
      e1p.ecg = cg; e1p.n2g = b->node2graph;
      e1p.silu = 1;
      CB2_TRY(launch_sgemm_nt(w.emb, DIS, L.w_fd, w.a1, H, rows, H, DIS, e1p, st));
      GemmEpilogue e2p;
      e2p.bias = L.b2; e2p.silu = 1;
      CB2_TRY(launch_sgemm_nt(w.a1, H, L.w2, w.e2, H, rows, H, H, e2p, st));
      CB2_TRY(launch_segment_mean(w.e2, b->node_eoff, b->node_n, agg, ld_agg, nlo, nhi, e0, (int64_t)v * N, st));
    }
  }
  return CB2_OK;
}

// ---- exact-mode layers -------------------------------------------------------
static int f32_forward_layers(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, ForwardWs &w,
                              cudaStream_t st) {
  const int N = b->n_nodes, B = b->n_graphs, V = b->n_variants;
  const int64_t VN = (int64_t)V * N;
  CB2_TRY(launch_lattice_ip(io->lattices, m, w.cg, B, nullptr, st));      // all layers' lattice terms, one launch
  for (int li = 0; li < m->n_layers; li++) {
    const cb2_layer_weights &L = m->layers[li];
    if (io->film_cond != nullptr) {
      GemmEpilogue e;
      e.bias = m->film_bp;
      CB2_TRY(launch_sgemm_nt(w.h, H, m->film_wp, w.y, H, VN, H, H, e, st));
    }
    CB2_TRY(launch_film_apply(w.y, w.h, io->film_cond, b->node2graph, m->film_g, m->film_b, L.ln_g, L.ln_b,
                              w.cat, H2, nullptr, 0, 0, N, B, V, st));
    {
      GemmEpilogue e;  // P = hn [W_hi;W_hj]^T; the lattice term + b1 (cg) joins it per edge
      CB2_TRY(launch_sgemm_nt(w.cat, H2, L.w_hij, w.P, H2, VN, H2, H, e, st));
    }
    CB2_TRY(f32_edge_layer(L, b, io->frac_coords, w.P, w.cg + (size_t)li * B * H, w.cat + H, H2, w, st));
    {
      GemmEpilogue e;
      e.bias = L.bn1; e.silu = 1;
      CB2_TRY(launch_sgemm_nt(w.cat, H2, L.wn1, w.z1, H, VN, H, H2, e, st));
      GemmEpilogue e2;
      e2.bias = L.bn2; e2.silu = 1; e2.residual = w.h; e2.ldr = H;
      CB2_TRY(launch_sgemm_nt(w.z1, H, L.wn2, w.h, H, VN, H, H, e2, st));
    }
  }
  return CB2_OK;
}

static int decoder_forward(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, ForwardWs &w,
                           cudaStream_t st) {
  const int N = b->n_nodes, B = b->n_graphs, V = b->n_variants;
  const int64_t VN = (int64_t)V * N;
  if (N == 0) return CB2_OK;
  NvtxRange r_fwd(io->coords_only ? "cb2:forward(corrector)" : "cb2:forward(predictor)");
  w.h_final = nullptr;
  if (io->precision != CB2_PRECISION_FP32 && tc_uses_node_chain(m, io))
    CB2_TRY(launch_embed_panels(io->atom_types, m->emb, w.h, w.h16, N, V, st));   // h, h16 in the chain kernel's layouts
  else
    CB2_TRY(launch_embed(io->atom_types, m->emb, w.h, N, V, st));
  if (io->precision == CB2_PRECISION_FP32) {
    CB2_TRY(f32_forward_layers(m, b, io, w, st));
  } else {
    CB2_TRY(tc_forward_layers(m, b, io, w, st));
  }
  NvtxRange r_heads("cb2:heads");
  const float *h_last = w.h_final ? w.h_final : w.h;
  float *hf = io->node_features ? io->node_features : w.hf;
  const bool tc_heads = io->precision != CB2_PRECISION_FP32 && m->w_head_t != nullptr;
  if (io->coords_only && tc_heads) {
    // corrector forward: only pred_x is used (chemeleon.py:440-450).  The three coordinate rows are taken
    // in fp32 straight from the final LayerNorm's registers: no head GEMM over 104 unused type columns,
    // no split-precision copy of the features (and no feature store unless the caller wants them).
    CB2_TRY(launch_layernorm(h_last, m->final_g, m->final_b, io->node_features, nullptr, m->w_head + (size_t)NTYPE * H,
                             io->head_out + NTYPE, HEADC, VN, st));
    return CB2_OK;
  }
  CB2_TRY(launch_layernorm(h_last, m->final_g, m->final_b, hf, tc_heads ? w.cat16 : nullptr, nullptr, nullptr, 0, VN, st));
  if (tc_heads) {   // cat16 is dead after the last layer: it carries the hi | lo split of the features
    CB2_TRY(tc_head(m, w.cat16, VN, io->head_out, st));
  } else {
    GemmEpilogue e;
    e.bias = m->b_head;
    CB2_TRY(launch_sgemm_nt(hf, H, m->w_head, io->head_out, HEADC, VN, HEADC, H, e, st));
  }
  if (!io->coords_only) {
    if (!io->lattice_out) return fail(CB2_ERR_BAD_ARG, "lattice_out is NULL but coords_only is 0");
    CB2_TRY(launch_lattice_head(hf, m->w_lat, io->lattices, b->graph_off, io->lattice_out, N, B, V, st));
  }
  return CB2_OK;
}

}  // namespace cb2

using namespace cb2;

extern "C" {

int cb2_abi_version(void) { return CB2_ABI_VERSION; }

const char *cb2_last_error(void) { return g_err.c_str(); }

uint64_t cb2_launch_count(void) { return g_launches.load(); }

int cb2_validity_filter(const int64_t *atom_types, const float *frac_coords, const float *lattices,
                        const int32_t *graph_off, int32_t n_graphs, const int32_t *target_reduced_counts,
                        float max_length, float min_distance, int32_t *flags, float *min_dist, float *max_abc,
                        void *stream) {
  if (n_graphs < 0) return fail(CB2_ERR_BAD_ARG, "validity_filter: negative n_graphs");
  if (n_graphs > 0 && (!atom_types || !frac_coords || !lattices || !graph_off || !flags || !min_dist || !max_abc))
    return fail(CB2_ERR_BAD_ARG, "validity_filter: null argument");
  return launch_validity(atom_types, frac_coords, lattices, graph_off, n_graphs, target_reduced_counts, max_length,
                         min_distance, flags, min_dist, max_abc, (cudaStream_t)stream);
}

/* development aid (not part of the documented ABI): clock64 timeline of the edge kernel's CTA 0 */
int cb2_debug_edge_timeline(long long *out96) { return debug_edge_timeline(out96); }
int cb2_debug_edge2_timeline(long long *out96x3) { return debug_edge2_timeline(out96x3); }
int cb2_debug_node2_timeline(long long *out64x3) { return debug_node2_timeline(out64x3); }

int cb2_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(CB2_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(CB2_ERR_DEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                    std::to_string(prop.minor) + "; this library contains sm_100a code only");
  return CB2_OK;
}

size_t cb2_workspace_bytes(const cb2_model *m, const cb2_batch *batch, int precision) {
  if (!batch || !m) return 0;
  Arena a(nullptr, 0, true);
  ForwardWs fw;
  StepWs sw;
  carve_forward(a, batch, m->n_layers, precision, fw);
  carve_step(a, batch, sw);
  return a.off + 256;
}

int cb2_embed_nodes(const cb2_model *m, const cb2_batch *b, const int64_t *atom_types, float *h, void *stream) {
  CB2_TRY(check_model(m));
  if (!b || !atom_types || !h) return fail(CB2_ERR_BAD_ARG, "embed: null argument");
  return launch_embed(atom_types, m->emb, h, b->n_nodes, b->n_variants, (cudaStream_t)stream);
}

int cb2_film_cond(const cb2_model *m, const cb2_batch *b, const float *text_part, const int32_t *t_dev,
                  float *film_cond, void *stream) {
  CB2_TRY(check_model(m));
  if (!b || !text_part || !film_cond) return fail(CB2_ERR_BAD_ARG, "film_cond: null argument");
  if (t_dev != nullptr && m->film_time_table == nullptr)
    return fail(CB2_ERR_BAD_ARG, "film_cond: t_dev given but the model has no film_time_table");
  return launch_film_cond(t_dev ? m->film_time_table : nullptr, text_part, nullptr, t_dev, film_cond,
                          (int64_t)b->n_variants * b->n_graphs, (cudaStream_t)stream);
}

size_t cb2_text_condition_workspace_bytes(const cb2_text_tail *t, int32_t n_prompts) {
  if (!t || n_prompts < 0) return 0;
  const size_t R = (size_t)n_prompts + 1;
  return ((R * t->embed_dim * sizeof(float) + 255) & ~size_t(255)) * 2 + ((R * t->text_dim * sizeof(float) + 255) & ~size_t(255));
}

int cb2_text_condition(const cb2_text_tail *t, const float *enc, int32_t n_prompts, float *text_part, void *workspace,
                       size_t workspace_bytes, void *stream) {
  if (!t || !text_part || (n_prompts > 0 && !enc) || n_prompts < 0) return fail(CB2_ERR_BAD_ARG, "text_condition: null argument");
  if (!t->w1 || !t->b1 || !t->ln_g || !t->ln_b || !t->w2 || !t->b2 || !t->null_embeds || !t->w_text || !t->b_cond)
    return fail(CB2_ERR_BAD_ARG, "text_condition: cb2_text_tail has a null weight");
  const int E = t->embed_dim, D = t->text_dim;
  if (E % 32 != 0 || D % 16 != 0 || E <= 0 || D <= 0)
    return fail(CB2_ERR_UNSUPPORTED, "text_condition: embed_dim % 32 and text_dim % 16 must be 0");
  const int64_t R = (int64_t)n_prompts + 1;
  Arena a(workspace, workspace_bytes, false);
  float *x = a.take<float>((size_t)R * E);      // encoder rows, then the learned null embedding
  float *y = a.take<float>((size_t)R * E);
  float *z = a.take<float>((size_t)R * D);
  if (!workspace || !a.ok()) return fail(CB2_ERR_WORKSPACE, "text_condition: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_prompts > 0)
    CB2_CUDA_OK(cudaMemcpyAsync(x, enc, (size_t)n_prompts * E * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CB2_CUDA_OK(cudaMemcpyAsync(x + (size_t)n_prompts * E, t->null_embeds, (size_t)E * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  GemmEpilogue e1;
  e1.bias = t->b1;
  CB2_TRY(launch_sgemm_nt(x, E, t->w1, y, E, R, E, E, e1, st));            // text_emb.0
  CB2_TRY(launch_ln_gelu(y, t->ln_g, t->ln_b, E, R, st));                   // text_emb.1, text_emb.2
  GemmEpilogue e2;
  e2.bias = t->b2;
  CB2_TRY(launch_sgemm_nt(y, E, t->w2, z, D, R, D, E, e2, st));            // text_emb.3 -> [R, text_dim]
  GemmEpilogue e3;
  e3.bias = t->b_cond;
  CB2_TRY(launch_sgemm_nt(z, D, t->w_text, text_part, H2, R, H2, D, e3, st));   // text half of FilmLayer.mlp_cond
  return CB2_OK;
}

int cb2_linear_f32(const float *A, int64_t lda, const float *W, const float *bias, float *C, int64_t ldc,
                   int64_t M, int32_t N, int32_t K, int32_t silu, void *stream) {
  if (!A || !W || !C) return fail(CB2_ERR_BAD_ARG, "linear: null argument");
  GemmEpilogue e;
  e.bias = bias;
  e.silu = silu;
  return launch_sgemm_nt(A, lda, W, C, ldc, M, N, K, e, (cudaStream_t)stream);
}

size_t cb2_linear_tc_workspace_bytes(int64_t M, int32_t K) {
  if (M <= 0 || K <= 0) return 0;
  return (size_t)((M + 127) / 128 * 128) * (size_t)K * sizeof(__half);
}

int cb2_linear_tc(const void *A16, int64_t lda, const void *Wt, int32_t Nw, const float *bias, float *C,
                  int64_t ldc, int64_t M, int32_t K, int32_t silu, void *workspace, size_t workspace_bytes,
                  void *stream) {
  if (!A16 || !Wt || !C) return fail(CB2_ERR_BAD_ARG, "linear_tc: null argument");
  return tc_linear_simple(A16, lda, Wt, Nw, bias, C, ldc, M, K, silu, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int cb2_edge_layer(const cb2_model *m, int32_t layer, const cb2_batch *b, const float *frac_coords,
                   const void *P, const float *cg, void *agg, int64_t ld_agg, int32_t precision, void *workspace,
                   size_t workspace_bytes, void *stream) {
  CB2_TRY(check_model(m));
  CB2_TRY(check_batch(b, precision));
  if (layer < 0 || layer >= m->n_layers) return fail(CB2_ERR_BAD_ARG, "edge_layer: bad layer index");
  if (!frac_coords || !P || !agg) return fail(CB2_ERR_BAD_ARG, "edge_layer: null argument");
  const cb2_layer_weights &L = m->layers[layer];
  if (precision == CB2_PRECISION_FP32) {
    Arena a(workspace, workspace_bytes, false);
    ForwardWs fw;
    carve_forward(a, b, m->n_layers, precision, fw);
    if (!workspace || !a.ok()) return fail(CB2_ERR_WORKSPACE, "workspace too small: call cb2_workspace_bytes()");
    return f32_edge_layer(L, b, frac_coords, (const float *)P, cg, (float *)agg, ld_agg, fw, (cudaStream_t)stream);
  }
  if (!L.w_fd_t || !L.w2_t) return fail(CB2_ERR_BAD_ARG, "edge_layer: fp16 operand images missing");
  return tc_edge_layer(m, L, b, frac_coords, (const __half *)P, cg, (__half *)agg, ld_agg, 0, 0, (cudaStream_t)stream);
}

int cb2_decoder_forward(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io, void *workspace,
                        size_t workspace_bytes, void *stream) {
  CB2_TRY(check_model(m));
  if (!io) return fail(CB2_ERR_BAD_ARG, "null io");
  CB2_TRY(check_batch(b, io->precision));
  if (!io->atom_types || !io->frac_coords || !io->lattices || !io->head_out)
    return fail(CB2_ERR_BAD_ARG, "forward: null state/output pointer");
  Arena a(workspace, workspace_bytes, false);
  ForwardWs fw;
  carve_forward(a, b, m->n_layers, io->precision, fw);
  if (!workspace || !a.ok()) return fail(CB2_ERR_WORKSPACE, "workspace too small: call cb2_workspace_bytes()");
  return decoder_forward(m, b, io, fw, (cudaStream_t)stream);
}

int cb2_update_predictor(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                         const float *lattice_out, void *stream) {
  return update_predictor(b, s, a, head_out, lattice_out, (cudaStream_t)stream);
}

int cb2_update_corrector(const cb2_batch *b, cb2_state *s, const cb2_step_args *a, const float *head_out,
                         void *stream) {
  return update_corrector(b, s, a, head_out, (cudaStream_t)stream);
}

size_t cb2_frame_bytes(int32_t n_nodes, int32_t n_graphs) { return frame_bytes(n_nodes, n_graphs); }

int cb2_pack_frame(const cb2_batch *b, const cb2_state *s, void *frame, size_t frame_bytes_, void *stream) {
  return pack_frame(b, s, frame, frame_bytes_, (cudaStream_t)stream);
}

int cb2_sampler_step(const cb2_model *m, const cb2_batch *b, cb2_state *s, const cb2_step_args *a,
                     void *workspace, size_t workspace_bytes, void *stream) {
  CB2_TRY(check_model(m));
  if (!s || !a) return fail(CB2_ERR_BAD_ARG, "null state/args");
  CB2_TRY(check_batch(b, a->precision));
  if (!m->film_time_table) return fail(CB2_ERR_BAD_ARG, "sampler_step needs cb2_model.film_time_table");
  if (!a->text_part) return fail(CB2_ERR_BAD_ARG, "sampler_step needs text_part");
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes, false);
  ForwardWs fw;
  StepWs sw;
  carve_forward(ar, b, m->n_layers, a->precision, fw);
  carve_step(ar, b, sw);
  if (!workspace || !ar.ok()) return fail(CB2_ERR_WORKSPACE, "workspace too small: call cb2_workspace_bytes()");
  CB2_TRY(launch_film_cond(m->film_time_table, a->text_part, a->text_row, s->t_dev, sw.film_cond,
                           (int64_t)b->n_variants * b->n_graphs, st));
  cb2_forward_io io{};
  io.atom_types = s->atom_types;
  io.frac_coords = s->frac_coords;
  io.lattices = s->lattices;
  io.film_cond = sw.film_cond;
  io.head_out = sw.head_out;
  io.lattice_out = sw.lat_out;
  io.node_features = nullptr;
  io.coords_only = 0;
  io.precision = a->precision;
  io.flags = s->flags;
  NvtxRange r_step("cb2:sampler_step");
  CB2_TRY(decoder_forward(m, b, &io, fw, st));                        // predictor (cond | null)
  {
    NvtxRange r("cb2:update(predictor)");
    CB2_TRY(update_predictor(b, s, a, sw.head_out, sw.lat_out, st));
  }
  io.coords_only = 1;
  CB2_TRY(decoder_forward(m, b, &io, fw, st));                        // corrector
  NvtxRange r("cb2:update(corrector)");
  CB2_TRY(update_corrector(b, s, a, sw.head_out, st));
  return CB2_OK;
}

}  // extern "C"
