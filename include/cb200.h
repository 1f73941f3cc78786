/* cb200 -- C-ABI of the B200-native reverse-diffusion pose-sampling hot path.
 *
 * Drop-in boundary for LDeng0205/confidence-bootstrapping (reference paths relative to the
 * reference root).  The reference has no native code: every entry point below replaces a chain of
 * third-party PyTorch ops that the reference reaches from Python, cited per function.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless marked host.
 *   - all kernels are asynchronous on `stream` (a cudaStream_t passed as void*), allocate nothing,
 *     keep no global state except the last-error string, and never synchronise: the whole step is
 *     CUDA-graph capturable.  Edge counts live on the device (CSR row pointers).
 *   - return 0 on success; otherwise a non-zero code and cb_last_error() describes the failure
 *     (the Python host raises RuntimeError, which the reference's callers catch to halve the
 *     batch: finetune_train.py:187-195, inference.py:566-570).
 *   - indices are int32 inside the library (int64 at the Python surface, like the reference).
 */
#ifndef CB200_H
#define CB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* cb_last_error(void);
int cb_version(void);
/* sizeof of the argument structs, for binding self-checks: 0 edge_feat, 1 tp_segment, 2 tp_conv, 3 sde_step,
 * 4 tp_row, 5 tp_term, 6 tp_run */
int cb_sizeof(int which);

/* ------------------------------------------------------------------------------------------ K1
 * Neighbour search.  Replaces torch_cluster.radius / radius_graph at
 *   models/score_model.py:502 (ligand radius graph), :568-573 (cross graph, per-graph dynamic
 *   cutoff = positions divided by cutoff[batch], r = 1), :655 (bond-centre graph);
 *   models/all_atom_score_model.py:528,593-598,611-612,657.
 * Predicate (bit-exact): same graph AND fl(fl(dx*dx + dy*dy) + dz*dz) < fl(r*r), with
 * dx = fl(y/c) - fl(x/c) when `cutoff` is given.  Output is CSR over the queries y with the
 * candidate indices ascending; a query keeps at most `max_neighbors` lowest-index candidates.
 * `exclude_self` drops candidate == query AFTER truncation (radius_graph: max_neighbors+1, loop=False).
 *
 * cb_radius_count writes count[q] (post-truncation, post self-removal); the caller turns it into
 * rowptr with cb_exclusive_scan_i32; cb_radius_fill then writes row[e] (query) and col[e] (candidate).
 */
int cb_radius_count(const float* x, const int32_t* x_ptr, const float* y, const int32_t* y_batch,
                    const float* cutoff /* [B] or NULL */, float r, int32_t n_y, int32_t max_neighbors,
                    int32_t exclude_self, int32_t* count, void* stream);
int cb_radius_fill(const float* x, const int32_t* x_ptr, const float* y, const int32_t* y_batch,
                   const float* cutoff, float r, int32_t n_y, int32_t max_neighbors, int32_t exclude_self,
                   const int32_t* rowptr, int32_t* row, int32_t* col, void* stream);
/* out[0] = 0, out[i+1] = in[0] + ... + in[i]   (n <= 2^24); `scratch` holds >= 4096 int32. */
int cb_exclusive_scan_i32(const int32_t* in, int32_t* out, int32_t n, int32_t* scratch, void* stream);
/* Transposed edge list of a radius relation: for every candidate x (ascending) the queries y
 * (ascending) whose KEPT neighbour list contains x -- the same edge set as cb_radius_fill, grouped
 * by x.  Used for the flipped edge groups (score_model.py:356-357 `torch.flip(lr_edge_index)`).
 * `kept_rowptr/kept_col` are the outputs of the forward search (needed only when truncation
 * occurred; pass NULL when max_neighbors >= every segment size). */
int cb_radius_count_t(const float* x, const int32_t* x_batch, const float* y, const int32_t* y_ptr,
                      const float* cutoff, float r, int32_t n_x, int32_t exclude_self,
                      const int32_t* kept_rowptr, const int32_t* kept_col, int32_t* count, void* stream);
int cb_radius_fill_t(const float* x, const int32_t* x_batch, const float* y, const int32_t* y_ptr,
                     const float* cutoff, float r, int32_t n_x, int32_t exclude_self,
                     const int32_t* kept_rowptr, const int32_t* kept_col,
                     const int32_t* rowptr_t, int32_t* row_t, int32_t* col_t, void* stream);
/* ------------------------------------------------------------------------------------------ K2
 * Edge featurisation: Gaussian smearing + real spherical harmonics + the 2-layer edge-embedding
 * MLP in one pass.  Replaces GaussianSmearing (score_model.py:667-677), o3.spherical_harmonics
 * (:519,536,581-582,647,661) and {lig,rec,cross,center,final}_edge_embedding (:111-123,238-243,
 * 259-264) applied in build_*_conv_graph (:492-664).
 *   vec      = sh_sign * (pos_nbr[col[e]] - pos_agg[row[e]])
 *   feature  = the reference's concatenation of bond one-hot / sigma embedding / smearing; the
 *              sigma-embedding block is constant per graph, so the host folds it into b1_graph
 *   out_attr = W2 relu(W1 feature + b1) + b2           [E, ns]
 *   out_sh   = SH_{l<=lmax}(vec), 'component' normalised, unit vector   [E, (lmax+1)^2]
 * The number of edges is read from n_edges_dev[0] on the device.
 */
typedef struct {
    const int32_t* row;         /* [E] aggregation node of the edge                       */
    const int32_t* col;         /* [E] neighbour node                                     */
    const int32_t* n_edges_dev; /* device scalar: E                                       */
    int32_t e_cap;              /* capacity of the output buffers (upper bound for E)     */
    const float* pos_agg;       /* [N_agg,3]                                              */
    const float* pos_nbr;       /* [N_nbr,3]                                              */
    float sh_sign;              /* +1: nbr-agg, -1: agg-nbr                               */
    int32_t lmax;               /* 1 or 2                                                 */
    const int32_t* agg_graph;   /* [N_agg] graph id of the aggregation node               */
    const float* b1_graph;      /* [B or 1, ns]: b1 + W1[:, sigma cols] sigma_emb[g] (host-folded) */
    int32_t b1_graph_stride;    /* ns, or 0 when one row is shared by all graphs          */
    const float* extra;         /* [E, n_extra] raw per-edge features or NULL (zeros)     */
    int32_t n_extra;
    const float* smear_offset;  /* [n_gauss] Gaussian centres (GaussianSmearing.offset buffer) */
    float smear_coeff;          /* -0.5/(offset[1]-offset[0])^2 (score_model.py:672)       */
    int32_t n_gauss;
    const float* W1; int32_t ldw1; /* first Linear [ns, ldw1]                              */
    int32_t extra_off, smear_off;  /* column offsets of the two in-kernel feature blocks   */
    const float* W2; const float* b2; /* [ns, ns], [ns]                                   */
    int32_t ns;
    float* out_attr;            /* [e_cap, ns]                                            */
    float* out_sh;              /* [e_cap, (lmax+1)^2]                                    */
} cb_edge_feat_args;
int cb_edge_featurize(const cb_edge_feat_args* a, void* stream);

/* ------------------------------------------------------------------------------------------ K3
 * Tensor-product convolution layer.  Replaces TensorProductConvLayer.forward
 * (models/tensor_layers.py:195-217): radial MLP -> per-edge tensor-product weights ->
 * FasterTensorProduct (:66-117) / e3nn FullyConnectedTensorProduct (:185) -> scatter-mean over
 * edge_index[0] (:206) -> e3nn BatchNorm in eval mode (:211-212) -> zero-padded residual (:214-216).
 *
 * The bilinear structure is used exactly (fp re-association only): per aggregation node i and
 * edge segment s,   sum_e tp_e = T_s( sum_e f_e (x) [h_e ; 1] ),   f_e = CG products of x[col[e]]
 * and sh_e (the "program"), h_e = relu(b1 + P_agg[i] + P_nbr[col[e]] + W1e (e_attr[e] + e_post[g(i)])),
 * T_s = contraction with (W2, b2).  The [E, weight_numel] tensor is never materialised.
 * Two launches: (a) accumulate -- persistent CTAs walk the (node, slot) pairs; per chunk of 16 edges the hidden
 * layer and the rank-16 update A += F^T [h;1] run as tcgen05 3xTF32 MMAs with TMEM accumulators (accum_mode 2-4:
 * lock-step, transposed, warp-specialised; accum_mode 1 keeps A in fp32 FFMA register tiles), the finished tile is
 * written once to the workspace;
 * (b) transform + epilogue -- one CTA per 32 nodes streams the tile's accumulator rows and the W2a rows with
 * bulk copies through a shared-memory ring and finishes with mean / BatchNorm / residual.
 */
typedef struct {            /* one CG product term of an f-row: coef * x[x_idx] * sh[sh_idx] */
    int16_t x_idx; int16_t sh_idx; float coef;
} cb_tp_term;
typedef struct {            /* one f-row (intermediate channel feeding one output irrep copy)   */
    int32_t term_begin, term_end; /* slice of the term table                                  */
    int32_t w_base;         /* weight row for multiplicity m is w_base + m                       */
    int32_t out_base;       /* output channel for multiplicity m is out_base + m*out_step        */
    int32_t out_step;
    int32_t mul;            /* number of output multiplicities fed by this row                   */
    int32_t p_off;          /* prefix sum of mul (kept for evaluators / tests)                   */
    int32_t pad_;
} cb_tp_row;
typedef struct {            /* consecutive rows feeding the same outputs, with contiguous weight rows: */
    int32_t row_begin, row_end;   /* row r uses weight rows w_base0 + (r-row_begin)*mul + m, m < mul   */
    int32_t out_base, out_step, mul, w_base0;
    int32_t pad0_, pad1_;
} cb_tp_run;
typedef struct {
    const int32_t* rowptr;  /* CSR over aggregation nodes: edges of node i are rowptr[i-n0]..rowptr[i-n0+1] */
    const int32_t* col;     /* [E] neighbour index into x                                       */
    const float* e_attr;    /* [E, ne]                                                          */
    const float* e_post;    /* [B, ne] added to e_attr (per graph of the aggregation node) or NULL */
    const float* sh;        /* [E, S]                                                           */
    const float* P_agg;     /* [n_out, ldp_agg] projection of aggregation-side scalars (row = node id) or NULL */
    const float* P_nbr;     /* [N_in, ldp_nbr] projection of neighbour-side scalars or NULL      */
    int32_t ldp_agg, ldp_nbr;
    const float* W1e;       /* first Linear, pointing at the first edge-embedding column [H, ldw1] */
    int32_t ldw1;
    const float* b1;        /* [H]                                                              */
    const float* W2a;       /* [weight_numel, H+4] second Linear with its bias folded in: row w =
                               (W2[w][0..H), b2[w], 0, 0, 0) -- contiguous rows feed 1-D bulk copies */
    int32_t n0, n1;         /* aggregation-node range served by this segment                    */
    int32_t col_off;        /* added to col[e] to index x / P_nbr (node tables are concatenated)  */
    int32_t slot;           /* segments with the same slot share (n0,n1), the radial MLP and one accumulator;
                               slots are numbered 0.. in segment order, equal slots adjacent        */
    const int32_t* gate_rowptr; /* optional [n1-n0+1] CSR row pointer of ANOTHER edge list over the same node range: nodes
                               with no edge there are skipped (treated as degree 0) -- dead-output pruning, e.g. receptor
                               rows of the last full conv layer that no rec->lig edge reads (score_model.py:372-374) */
    const uint8_t* gate_mask;   /* optional [n1-n0] keep flags with the same effect (0 = skip the node in this segment)   */
    const float* W2t;       /* optional: W2a re-laid for the tensor-core transform (cb_tp_chain.w_off indexes it): per chain
                               the hi then the lo TF32 part of the chain's `npad` weight rows as K-major core-matrix tiles  */
} cb_tp_segment;
typedef struct {            /* one MMA chain of the tensor-core transform: accumulator rows row[c] (one per stacked
                               output component, -1 = unused) times the weight tile at W2t + w_off                         */
    int32_t row[3]; int32_t n_comp, w_off, npad, acc_col, first;
} cb_tp_chain;
typedef struct {            /* one output block (components sharing their weights): where its partial sums sit in TMEM and
                               which output channels they feed: out_base[c] + m * out_step                                  */
    int32_t n_comp, mul, npad, acc_col0, n_partials, out_step; int32_t out_base[3]; int32_t pad_[3];
} cb_tp_block;
#define CB_MAX_SEGS 12
typedef struct {
    const float* x;         /* [N_in, d_in] node features gathered at col                        */
    int32_t d_in, d_out, S, ne, H;
    int32_t n_out;          /* number of aggregation nodes (rows of out)                        */
    const int32_t* agg_graph; /* [n_out] graph id (for e_post) or NULL                           */
    const cb_tp_row* rows; int32_t n_rows;   /* device tables built once per layer               */
    const cb_tp_term* terms; int32_t n_terms;
    const cb_tp_run* runs; int32_t n_runs;
    int32_t n_segs;
    cb_tp_segment segs[CB_MAX_SEGS];
    /* epilogue: mean over all segments, BatchNorm(eval) affine, residual                         */
    const float* bn_scale;  /* [d_out] weight/sqrt(running_var+eps) expanded per channel, or NULL */
    const float* bn_shift;  /* [d_out] bias - running_mean*scale on 0e channels, 0 elsewhere      */
    const float* residual;  /* [n_out, ld_res] added to the first d_res channels, or NULL        */
    int32_t d_res, ld_res;
    float* out;             /* [n_out, d_out]                                                    */
    /* workspace for the per-(node, slot) accumulators A[n_rows][H+4], stored per (slot, tile of 32 aggregation
     * nodes counted from node_begin) row-major over (row, rank of the node among the tile's active nodes) so that
     * the transform kernel reads a tile's row group with one contiguous bulk copy; the call processes the
     * aggregation nodes [node_begin, node_end) and needs cb_tp_conv_items(a) * n_rows * (H+4) floats            */
    float* workspace; int64_t workspace_floats;
    int32_t node_begin, node_end;
    int32_t accum_mode;     /* accumulate kernel: 1 = fp32 FFMA register tiles, 2 = tcgen05 3xTF32 with TMEM accumulator,
                             * 3 = 2 with the accumulator transposed (D[hidden unit][f-row]; rows <= 240, H % 32 == 0),
                             * 4 = warp-specialised transposed kernel (additionally ne == 32); 3 and 4 fall back to the
                             * widest kernel the layer shape allows.  Results agree to fp32 rounding.              */
    int32_t flags;          /* CB_TP_RAW_SUM: write the un-normalised sums (no mean, BatchNorm, residual) */
    /* Sample-invariant contribution computed by an earlier CB_TP_RAW_SUM call: aggregation node i in
     * [pre_n0, pre_n1) additionally receives pre_sum[(i - pre_n0) % pre_period][:] before the mean and counts
     * pre_deg[(i - pre_n0) % pre_period] more incoming edges.  Use: the S copies of one receptor in a sampling batch
     * have identical rec->rec messages in the first conv layer (same receptor, same t: models/score_model.py:298-320,
     * 365-370), so they are aggregated and transformed once for one copy and shared by all S.  NULL = none.        */
    const float* pre_sum;   /* [pre_period, d_out]                                               */
    const int32_t* pre_deg; /* [pre_period]                                                      */
    int32_t pre_n0, pre_n1, pre_period, pad_;
    /* tensor-core transform plan (optional; built once per layer on the host, irreps.transform_plan): when present and every
     * segment carries W2t, launches with enough node tiles run the tcgen05 transform kernel instead of the FFMA one          */
    const cb_tp_chain* chains; int32_t n_chains;
    const cb_tp_block* blocks; int32_t n_blocks;
    int32_t kp, max_chain_bytes;   /* K padded to 8; largest (raw rows + W tile) staging footprint of a chain in bytes */
} cb_tp_conv_args;
#define CB_TP_RAW_SUM 1
/* number of (node, slot) accumulators the call will use (host arithmetic only) */
int64_t cb_tp_conv_items(const cb_tp_conv_args* a);
int cb_tp_conv_forward(const cb_tp_conv_args* a, void* stream);

/* ------------------------------------------------------------------------------------------ K4
 * Reverse-SDE pose update for B copies of one ligand topology.  Replaces, per step,
 * utils/sampling.py:119-141 (perturbation from score and noise) and modify_conformer_batch
 * (utils/diffusion_utils.py:60-78): axis_angle_to_matrix (utils/geometry.py:39-86), rigid move,
 * sequential bond rotations (utils/torsion.py:75-90), Kabsch re-alignment with reflection fix
 * (utils/geometry.py:246-276).
 *   tr  = c_tr_score  * tr_score  + c_tr_noise  * z_tr     (same for rot, tor)
 */
typedef struct {
    float* pos;                 /* [B, N, 3] in/out                                           */
    int32_t B, N, R;            /* graphs, atoms per ligand, rotatable bonds                  */
    const int32_t* bond_uv;     /* [R, 2] (u, v): v lies on the rotating side                 */
    const uint8_t* mask_rotate; /* [R, N]                                                     */
    const float* tr_score; const float* rot_score; const float* tor_score; /* [B,3],[B,3],[B*R] */
    const float* z_tr; const float* z_rot; const float* z_tor;            /* noise or NULL (0) */
    float c_tr_score, c_tr_noise, c_rot_score, c_rot_noise, c_tor_score, c_tor_noise;
    const float* coeffs_dev;    /* optional [6] device copy of the six coefficients above (same order); when non-NULL it
                                   overrides them, so a CUDA-graph replay can change the step's scalars          */
} cb_sde_step_args;
int cb_sde_step(const cb_sde_step_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif
