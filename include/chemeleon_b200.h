/*
 * chemeleon_b200.h -- C-ABI of libchemeleon_b200.so
 *
 * B200 (sm_100a) kernels for the reverse-diffusion sampling path of
 * ryannduma/chemeleon.  The reference is pure Python/PyTorch and has no FFI or
 * plugin layer of its own, so each entry point below cites the
 * reference *function* it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C: POD structs of raw pointers + sizes, no torch / C++ types;
 *  - every buffer (weights, state, outputs, workspace) is CALLER-OWNED; the
 *    library never allocates or frees device memory and keeps no per-process
 *    configuration (no environment switches; the only process-wide state is the
 *    thread-local error string and the launch counter);
 *  - pointers are DEVICE pointers unless the field name starts with `host_`;
 *  - kernels are enqueued on the given stream (a cudaStream_t passed as void*)
 *    and never synchronise, so every call is CUDA-graph capturable;
 *  - return 0 on success, a negative cb2_status otherwise; cb2_last_error()
 *    returns a thread-local description.  Unsupported shapes / a non-sm_100
 *    device are hard errors: there is no fallback path of any kind.
 */
#ifndef CHEMELEON_B200_H
#define CHEMELEON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB2_ABI_VERSION 2
#define CB2_HIDDEN 512
#define CB2_MAX_LAYERS 16
#define CB2_MAX_ATOM_TYPES 104
#define CB2_NUM_FREQS 128
#define CB2_HEAD_COLS 128       /* type logits 0..103, coord 104..106, zero pad */
#define CB2_COEF_COLS 16        /* columns of the per-timestep coefficient table */
#define CB2_TILE_ROWS 128       /* edge rows per tensor-core tile */

typedef enum {
  CB2_OK = 0,
  CB2_ERR_BAD_ARG = -1,
  CB2_ERR_UNSUPPORTED = -2,     /* shape / mode the kernels are not built for */
  CB2_ERR_CUDA = -3,
  CB2_ERR_DEVICE = -4,          /* not an sm_100 device */
  CB2_ERR_WORKSPACE = -5
} cb2_status;

typedef enum {
  CB2_PRECISION_FP32 = 0,       /* exact mode: CUDA-core fp32 kernels */
  CB2_PRECISION_TC_F16 = 1      /* tcgen05 kind::f16 (fp16 operands, fp32 TMEM accumulate) */
} cb2_precision;

/* ---- weights of one CSPLayer (cspnet.py:100-181), W1 split by input block ---- */
typedef struct {
  const float *w_hij;   /* [1024,512] rows 0:512 = W1[:,0:512] (h_i), 512:1024 = W1[:,512:1024] (h_j) */
  const float *w_ip;    /* [512,9]    W1[:,1024:1033]  (lattice inner products, row-major 3x3) */
  const float *b1;      /* [512] */
  const float *w_fd;    /* [512,768]  W1[:,1033:1801]  (sinusoid embedding, reference column order) */
  const float *w2;      /* [512,512]  edge_mlp.2 */
  const float *b2;
  const float *wn1;     /* [512,1024] node_mlp.0 */
  const float *bn1;
  const float *wn2;     /* [512,512]  node_mlp.2 */
  const float *bn2;
  const float *ln_g;    /* CSPLayer.layer_norm */
  const float *ln_b;
  /* fp16 tcgen05 operand images (K-major, no swizzle: [K/8][rows][8]); NULL when only exact mode is
   * used.  w_fd_t has its K columns permuted to d*256 + 2k + {sin,cos} (weights.fd_column_order);
   * w2_t is the image of edge_mlp.2.weight / 2, stored as four images of 128 output channels each: [4][64][128][8]. */
  const void *w_hij_t, *w_fd_t, *w2_t, *wn1_t, *wn2_t;
} cb2_layer_weights;

/* ---- the CSPNet decoder (cspnet.py:184-405) ---- */
typedef struct {
  int32_t abi_version;  /* CB2_ABI_VERSION */
  int32_t hidden;       /* must be 512 */
  int32_t n_layers;
  int32_t n_atom_types; /* must be 104 */
  int32_t n_freqs;      /* must be 128 */
  int32_t timesteps;    /* T */
  const float *emb;             /* [104,512] node_embedding (smooth=False) */
  const float *film_wp;         /* [512,512] FilmLayer.proj */
  const float *film_bp;
  const float *film_g;          /* FilmLayer.norm */
  const float *film_b;
  const void *film_wp_t;
  const float *film_time_table; /* [T+1,1024] W_cond[:, :128] @ time_emb(t)  (may be NULL for forward-only use) */
  cb2_layer_weights layers[CB2_MAX_LAYERS];
  const float *final_g;         /* final_layer_norm */
  const float *final_b;
  const float *w_head;          /* [128,512] rows 0:104 type_out.weight, 104:107 coord_out.weight */
  const float *b_head;          /* [128] */
  const void *w_head_t;         /* tensor-core heads (may be NULL: fp32 SIMT heads): 16-byte header {float 1/s}, then the
                                 * fp16 K-major image [192][256][8] of s*[w_hi | w_hi | w_lo] (split precision, rows >= 128 zero) */
  const float *w_lat;           /* [9,512] lattice_out.weight */
  int32_t flags;                /* CB2_MODEL_* bits, 0 = defaults */
} cb2_model;

/* cb2_model.flags */
#define CB2_MODEL_EDGE_SINGLE_CTA 1  /* V == 2: run the one-CTA edge kernel per variant instead of the CTA-pair
                                      * kernel that shares the sinusoid GEMM between the variants (A/B testing) */
#define CB2_MODEL_NODE_UNFUSED 2     /* run the node-level GEMMs of a layer as four kernels (FiLM block, hoist GEMM, node
                                      * MLP x 2) instead of the one CTA-pair chain kernel per layer boundary (A/B testing) */

/* ---- topology of one ragged batch; fixed for a whole sampling run ----
 * Replaces CSPNet.gen_edges (cspnet.py:319-324): edges are implied by the
 * per-crystal node ranges, no dense adjacency is ever built. */
typedef struct {
  int32_t n_nodes;      /* N  = sum of atoms, ONE variant */
  int32_t n_graphs;     /* B */
  int32_t n_variants;   /* V: 1, or 2 = [conditional | unconditional] sharing the same state */
  int32_t max_n;        /* largest crystal */
  int64_t n_edges;      /* E = sum n^2, ONE variant */
  const int32_t *node2graph;   /* [N] */
  const int32_t *node_base;    /* [N] first node of the node's crystal */
  const int32_t *node_n;       /* [N] atoms in the node's crystal */
  const int32_t *graph_off;    /* [B+1] node offsets */
  /* exact path: materialised edge rows in the reference's order (i outer, j inner) */
  const int32_t *edge_i;       /* [E] */
  const int32_t *edge_j;       /* [E] */
  const int64_t *node_eoff;    /* [N+1] edge-row offset of node i's segment */
  int32_t n_chunks;            /* exact path processes edges in chunks of whole segments */
  const int32_t *host_chunk_node_lo; /* HOST [n_chunks+1] node boundaries of the chunks */
  const int64_t *host_chunk_edge_lo; /* HOST [n_chunks+1] edge-row boundaries of the chunks */
  int64_t chunk_max_edges;     /* largest chunk, in edge rows */
  /* tensor-core path: tiles of 128 edge rows = whole segments of equal length */
  int32_t n_tiles;
  const int32_t *tile_row_i;   /* [n_tiles*128] node i of the row, -1 = padding */
  const int32_t *tile_row_j;   /* [n_tiles*128] node j of the row */
  const int32_t *tile_seg_n;   /* [n_tiles] segment length (atoms per crystal) of the tile */
} cb2_batch;

/* ---- one decoder forward: CSPNet.forward (cspnet.py:345-405) ---- */
typedef struct {
  const int64_t *atom_types;   /* [N] */
  const float *frac_coords;    /* [N,3] */
  const float *lattices;       /* [B,9] row-major 3x3 */
  const float *film_cond;      /* [V*B,1024] SiLU(mlp_cond(cat[t,text])) = [scale|shift]; NULL = no FiLM */
  float *head_out;             /* [V*N,128]: type logits 0..103, coords 104..106 */
  float *lattice_out;          /* [V*B,9] (may be NULL with coords_only) */
  float *node_features;        /* [V*N,512] after final LN (may be NULL) */
  int32_t coords_only;         /* corrector call: lattice head skipped */
  int32_t precision;           /* cb2_precision */
  int32_t *flags;              /* [B] optional: CB2_FLAG_TC_RANGE is OR-ed in for crystals outside the fp16 range */
} cb2_forward_io;

/* per-crystal guard flags (cb2_state.flags, cb2_forward_io.flags) */
#define CB2_FLAG_NONFINITE 1   /* the update produced NaN/Inf for this crystal */
#define CB2_FLAG_TC_RANGE 2    /* tensor-core mode only: max |W_ip vec(L L^T) + b1| > CB2_TC_RANGE_LIMIT.  The edge
                                * pre-activations and the aggregated edge features grow with this per-crystal term;
                                * with fp16 operands (10-bit mantissa, like TF32) the decoder outputs stay within the
                                * 1e-3 tolerance of the fp32 reference up to the limit (validated on reference
                                * fixtures with cells of 3..40 Angstrom) and drift past it
                                * beyond (1.2e-3 at 55 A cells; the reference itself discards cells > 60 A,
                                * evaluate.py:180).  A flagged crystal is outside the validated range: re-run it in
                                * exact mode. */
#define CB2_TC_RANGE_LIMIT 96.0f

/* ---- sampler state + one reverse-diffusion timestep (chemeleon.py:379-467) ---- */
typedef struct {
  int64_t *atom_types;         /* [N] a_t  -> a_{t-1} (in place) */
  float *frac_coords;          /* [N,3] x_t -> x_{t-1} (wrapped) */
  float *lattices;             /* [B,9] l_t -> l_{t-1} */
  int32_t *t_dev;              /* device scalar: current timestep; decremented by the step */
  int32_t *flags;              /* [B] guard bits CB2_FLAG_* (OR-ed in, never cleared by the library) */
} cb2_state;

typedef struct {
  const float *coef;           /* [T+1,16] coefficient table (schedules.py) */
  const float *text_part;      /* [V*B,1024] W_cond[:,128:] @ text + b  (cond rows then null rows); with text_row:
                                * [n_rows,1024], e.g. the n_prompts + 1 rows cb2_text_condition writes */
  const int32_t *text_row;     /* optional [V*B]: row of text_part conditioning (variant v, crystal g); crystals that
                                * share a prompt share a row, every unconditional row is the one null row */
  float cond_scale;
  int32_t timesteps;           /* T: the lattice is clipped to [-6,6] at t == T only (chemeleon.py:424-425) */
  int32_t precision;
  int32_t noise_mode;          /* 0 = injected tensors below, 1 = in-kernel Philox */
  /* injected noise, indexed by step s = t_start - t: (chemeleon.py:400-404,418,435,455) */
  int32_t t_start;             /* timestep of slice 0 of the injected tensors */
  const float *rand_a;         /* [S,N,104] uniform */
  const float *rand_l;         /* [S,B,9]   normal */
  const float *rand_x;         /* [S,N,3]   normal, predictor */
  const float *rand_x2;        /* [S,N,3]   normal, corrector */
  /* Philox: noise keyed by (seed, global sample id, atom, timestep, stream) */
  uint64_t seed;
  const uint64_t *seed_dev;    /* optional device scalar overriding `seed` (lets a captured graph be re-seeded) */
  const int64_t *graph_gid;    /* [B] global sample ids (sharding-invariant noise) */
} cb2_step_args;

/* Library / device checks. */
int cb2_abi_version(void);
const char *cb2_last_error(void);
int cb2_check_device(int device);            /* CB2_OK iff compute capability 10.x */

/* ---- operand images for the tensor-core path, packed on the HOST (pure CPU code) ----
 * What weights.pack_weights does in Python, for hosts without it: fp32 row-major [rows,K] (a torch
 * Linear weight) -> the fp16 image the kernels stream.  kind:
 *   CB2_PACK_KMAJOR      [K/8][rows][8]  ("K-major, no swizzle": core matrix = 8 rows x 16 B)   w_hij_t, wn1_t, wn2_t, film_wp_t
 *   CB2_PACK_FD          the same with the K columns of W_fd permuted from the reference order (sin block | cos block,
 *                        cspnet.py:49-51) to d*2F + 2k + {sin,cos}                              w_fd_t
 *   CB2_PACK_ROW_BLOCKS  one K-major image per block of 128 rows, of w / 2 (exact; the edge kernels'
 *                        GEMM2 accumulates x / 2 for SiLU(x) = x/2 (1 + tanh(x/2))): [rows/128][K/8][128][8]  w2_t
 *   CB2_PACK_HEAD_SPLIT  16-byte header {float 1/s} + image [3K/8][256][8] of s [w_hi | w_hi | w_lo]
 *                        (split precision, power-of-two scale s, rows padded to 256)             w_head_t */
#define CB2_PACK_KMAJOR 0
#define CB2_PACK_FD 1
#define CB2_PACK_ROW_BLOCKS 2
#define CB2_PACK_HEAD_SPLIT 3
size_t cb2_pack_bytes(int32_t kind, int32_t rows, int32_t K);
int cb2_pack_weights(int32_t kind, const float *w /*host [rows,K]*/, int32_t rows, int32_t K, void *out /*host*/,
                     size_t out_bytes);

/* Workspace one forward / step of model `m` needs for this batch (bytes). */
size_t cb2_workspace_bytes(const cb2_model *m, const cb2_batch *batch, int precision);

/* h = node_embedding(atom_types); replaces cspnet.py:357 (also used by tests). */
int cb2_embed_nodes(const cb2_model *m, const cb2_batch *b, const int64_t *atom_types,
                    float *h /*[V*N,512]*/, void *stream);

/* cond = SiLU(time_table[t] + text_part): FilmLayer.mlp_cond (cspnet.py:70-73,80-81)
 * with the time half tabulated.  t is read from *t_dev. */
int cb2_film_cond(const cb2_model *m, const cb2_batch *b, const float *text_part,
                  const int32_t *t_dev, float *film_cond /*[V*B,1024]*/, void *stream);

/* ---- text-conditioning tail: TextEncoder.get_text_embeds after the language model
 * (text_encoder/text_encoder.py:40-45,186-205) folded through the text half of FilmLayer.mlp_cond
 * (cspnet.py:70-73): once per distinct prompt, off the per-timestep path. ---- */
typedef struct {
  int32_t embed_dim;            /* TextEncoder.text_embed_dim (768) */
  int32_t text_dim;             /* TextEncoder.text_dim (512) */
  const float *w1, *b1;         /* text_emb.0: Linear(embed_dim, embed_dim) */
  const float *ln_g, *ln_b;     /* text_emb.1: LayerNorm(embed_dim); text_emb.2 = GELU */
  const float *w2, *b2;         /* text_emb.3: Linear(embed_dim, text_dim) */
  const float *null_embeds;     /* [1,embed_dim] TextEncoder.null_text_embeds (cond_drop_prob = 1 replaces every row by it) */
  const float *w_text;          /* [1024,text_dim] FilmLayer.mlp_cond.0.weight[:, time_dim:] */
  const float *b_cond;          /* [1024] FilmLayer.mlp_cond.0.bias */
} cb2_text_tail;

/* text_part[r] = W_text text_emb(r < n_prompts ? enc[r] : null_embeds) + b_cond for r = 0..n_prompts:
 * rows 0..n_prompts-1 condition on the language-model embeddings enc [n_prompts,embed_dim]
 * (get_text_embeds(cond_drop_prob=0)), row n_prompts is the unconditional row (cond_drop_prob=1). */
int cb2_text_condition(const cb2_text_tail *t, const float *enc, int32_t n_prompts,
                       float *text_part /*[n_prompts+1,1024]*/, void *workspace, size_t workspace_bytes, void *stream);
size_t cb2_text_condition_workspace_bytes(const cb2_text_tail *t, int32_t n_prompts);

/* C = act(A W^T + bias): the fp32 linear the exact path is built from (tests). */
int cb2_linear_f32(const float *A, int64_t lda, const float *W, const float *bias, float *C,
                   int64_t ldc, int64_t M, int32_t N, int32_t K, int32_t silu, void *stream);

/* C = act(A16 W16^T + bias) on the tensor cores (tcgen05, fp32 accumulate).  A16: fp16
 * row-major [M,lda]; Wt: fp16 operand image [K/8][Nw][8] of a torch Linear weight [Nw,K]
 * (weights.tile_k_major); K % 64 == 0, Nw % 256 == 0, lda % 8 == 0 (A is re-tiled into the row-panel layout of the pipeline first, into the caller's workspace).  The building block of the
 * node-level GEMMs (FilmLayer.proj, hoisted W1 blocks, node_mlp; cspnet.py:86,113,120). */
int cb2_linear_tc(const void *A16, int64_t lda, const void *Wt, int32_t Nw, const float *bias, float *C,
                  int64_t ldc, int64_t M, int32_t K, int32_t silu, void *workspace, size_t workspace_bytes,
                  void *stream);
/* ... its workspace (the row-panel copy of A): bytes for an [M,K] operand. */
size_t cb2_linear_tc_workspace_bytes(int64_t M, int32_t K);

/* Edge model + scatter_mean of ONE CSPLayer (CSPLayer.edge_model + the aggregation in
 * node_model, cspnet.py:129-160) from the hoisted node terms P [V*N,1024] = (P_i | P_j) = hn [W_hi;W_hj]^T
 * and the per-crystal lattice term cg [B,512] = W_ip vec(L L^T) + b1 (fp32 in both modes; NULL = 0):
 *   agg_i = mean_j SiLU(W2 SiLU(P_i[i] + cg[graph(i)] + P_j[j] + W_fd emb(x_j - x_i)) + b2).
 * precision FP32: P and agg are float; TC_F16: P and agg are fp16 (row-major, leading dimensions
 * 1024 and ld_agg). */
int cb2_edge_layer(const cb2_model *m, int32_t layer, const cb2_batch *b, const float *frac_coords,
                   const void *P, const float *cg, void *agg, int64_t ld_agg, int32_t precision, void *workspace,
                   size_t workspace_bytes, void *stream);

/* One CSPNet.forward (cspnet.py:345-405) for all V variants of the batch. */
int cb2_decoder_forward(const cb2_model *m, const cb2_batch *b, const cb2_forward_io *io,
                        void *workspace, size_t workspace_bytes, void *stream);

/* Predictor half of the update: CFG mix (chemeleon.py:288-290), D3PM.p_logits
 * (diff_utils.py:307-329), lattice DDPM step (chemeleon.py:413-425), coordinate
 * predictor (chemeleon.py:427-437).  x_t -> x_{t-1/2} (unwrapped) in place. */
int cb2_update_predictor(const cb2_batch *b, cb2_state *s, const cb2_step_args *a,
                         const float *head_out, const float *lattice_out, void *stream);

/* Corrector half: Langevin step + mod-1 wrap (chemeleon.py:452-463); decrements *t_dev. */
int cb2_update_corrector(const cb2_batch *b, cb2_state *s, const cb2_step_args *a,
                         const float *head_out, void *stream);

/* One full timestep of Chemeleon._sample_generator's loop body (chemeleon.py:379-467):
 * film_cond, predictor forward (cond+null), predictor update, corrector forward,
 * corrector update.  Capturable in a CUDA graph; replay T times. */
int cb2_sampler_step(const cb2_model *m, const cb2_batch *b, cb2_state *s, const cb2_step_args *a,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Streaming / trajectory wire format (replaces the per-step device->host copy + ase.Atoms construction of
 * TrajectoryContainer.get_atoms, schema.py:57-83, and the JSON of Atoms the reference server streams,
 * app/server.py:49-52): one compact frame per timestep, packed on the device so that a single async
 * copy into a pinned ring buffer moves it.  Layout (little endian):
 *   int32 t (timestep index the state belongs to), int32 n_nodes, int32 n_graphs, int32 0;
 *   uint8 types[n_nodes] (values > 103 -> 0 like schema.py:60-62), zero-padded to a multiple of 4 bytes;
 *   float frac_coords[n_nodes][3]; float lattice[n_graphs][9]. */
size_t cb2_frame_bytes(int32_t n_nodes, int32_t n_graphs);
int cb2_pack_frame(const cb2_batch *b, const cb2_state *s, void *frame /*device*/, size_t frame_bytes, void *stream);

/* Validity pre-filter of finished structures on the device: restates
 * chemeleon/scripts/evaluate.py:177-189 (max(lattice.abc) > max_length; smallest positive
 * periodic distance < min_distance) and sample_target_composition.py:57-62 (reduced
 * composition != target).  graph_off: [n_graphs+1] node offsets; target_reduced_counts: [104]
 * atoms per atomic number of the target's reduced formula, or NULL to skip that test.
 * flags[g] = OR of CB2_INVALID_*; min_dist[g] (inf if no positive distance), max_abc[g]. */
#define CB2_INVALID_LATTICE 1
#define CB2_INVALID_DISTANCE 2
#define CB2_INVALID_COMPOSITION 4
int cb2_validity_filter(const int64_t *atom_types, const float *frac_coords, const float *lattices,
                        const int32_t *graph_off, int32_t n_graphs, const int32_t *target_reduced_counts,
                        float max_length, float min_distance, int32_t *flags, float *min_dist, float *max_abc,
                        void *stream);

/* Number of kernels launched by this library in this process (bench: gpu_launches). */
uint64_t cb2_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CHEMELEON_B200_H */
