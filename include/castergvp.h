/*
 * castergvp.h -- C ABI of libcastergvp.so: B200 (sm_100a) kernels for the CASTER-DTA GVP hot path.
 *
 * The reference (stelleg/caster-dta) is pure Python/PyTorch and has NO FFI or plugin registry; its "operator API"
 * for this path is the nn.Module interface of models/gvp_layers.py.  Each entry point below therefore cites the
 * reference Python it replaces (paths relative to the reference root) and INTEGRATION.md shows the ctypes stub a
 * maintainer would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless named h_*; fp32 tensors are contiguous, row-major:
 *       scalars s:[rows, S], vectors V:[rows, C, 3] (xyz innermost)           (gvp_layers.py:66-77)
 *       edge_index:[2,E] int64, row 0 = source j, row 1 = target i            (gvp_layers.py:291-300, PyG flow)
 *   - The caller owns EVERY buffer including workspaces; the library never allocates or frees device memory,
 *     never synchronises the device and never changes the current device.  All work is enqueued on `stream`.
 *   - Return value: 0 = ok; < 0 = argument error detected before any launch (see cgvp_last_error());
 *     > 0 = cudaError_t of a failed launch.  No exceptions cross the ABI.  Thread-safe (thread-local error text).
 */
#ifndef CASTERGVP_H
#define CASTERGVP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cgvp_stream_t; /* cudaStream_t */

#define CGVP_MAX_CHAIN 4

enum { CGVP_ACT_NONE = 0, CGVP_ACT_RELU = 1, CGVP_ACT_SIGMOID = 2 };
enum { CGVP_AGGR_SUM = 0, CGVP_AGGR_MEAN = 1 };

/* One Geometric Vector Perceptron -- replaces GVP.__init__/forward, models/gvp_layers.py:123-175.
 * (si,vi) -> (so,vo) with h hidden vector channels (h = h_dim or max(vi,vo), :130).  vi == 0 selects the
 * scalar-only branch (:167-171), vo == 0 returns scalars only (:175). */
typedef struct CgvpGvpDesc {
    int32_t si, vi, so, vo, h;
    int32_t scalar_act;  /* CGVP_ACT_*  (activations[0]) */
    int32_t vector_act;  /* CGVP_ACT_*  (activations[1]) */
    int32_t vector_gate; /* 0/1 */
} CgvpGvpDesc;

/* nn.Linear parameters of one GVP in PyTorch layout ([out, in] row-major); NULL where the GVP has none.
 *   wh:[h,vi]  ws:[so,si+h] (+bs:[so])  wv:[vo,h]  wsv:[vo,so] (+bg:[vo])       (gvp_layers.py:129-137) */
typedef struct CgvpGvpWeights {
    const float *wh, *ws, *bs, *wv, *wsv, *bg;
} CgvpGvpWeights;

typedef struct CgvpGvpGrads {
    float *wh, *ws, *bs, *wv, *wsv, *bg;
} CgvpGvpGrads;

/* ---- library ---------------------------------------------------------------------------------------------- */
const char* cgvp_last_error(void);
int32_t cgvp_version(void);
int32_t cgvp_sm_count(void); /* SMs of the current device (grid sizing), 0 if no device */

/* ---- kernel selection (testing / measurement aid; no reference counterpart) -----------------------------------
 * Descriptors whose dims match a compiled-in specialisation (the CASTER-DTA checkpoint dims: message chain of
 * GVPConv on (16,4)/(32,1), gvp_node, gvp_edge, node update, readout) are served by register-resident kernels
 * (csrc/conv_reg.cu, csrc/rows_reg.cu); everything else by the generic shared-memory tile kernels.  Passing 0
 * forces the generic kernels (parity tests compare the two).  Default 1. */
int32_t cgvp_set_fast_paths(int32_t on);
/* Precision mode of the fused GVPConv forward.  0 (default): fp32 FFMA everywhere (<= 1e-4 of the reference).
 * 1: where a tensor-core kernel is compiled in (config-5 dims, nodes (100,16) / edges (32,1): csrc/conv_tc.cu) the
 * W_h / W_s / W_mu / gate projections of the message GVPs run as tcgen05.mma with bf16 operands and fp32
 * accumulation in TMEM (<= 1e-2 of the reference, the bound BASELINE.json states for tensor-core modes). */
int32_t cgvp_set_tensor_cores(int32_t on);
/* Dense GEMMs of the wide node update (csrc/rows_wide.cu).  1 (default): fp32-accurate split-precision ("3xTF32":
 * hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) tcgen05 GEMM (csrc/gemm_tc.cu); 0: register-tiled FFMA GEMM.
 * Both meet the fp32 parity bound (<= 1e-4 of the reference). */
int32_t cgvp_set_wide_gemm(int32_t tensor);

/* ---- kernel timing (measurement aid; no reference counterpart) ------------------------------------------------
 * When enabled, the library brackets the MAIN kernel of each operator with cudaEvents on the caller's stream.
 * cgvp_profile_collect synchronises on the recorded events and returns the summed device time and the launch
 * count of one kernel id, then forgets those records. */
enum { CGVP_K_CONV_FWD = 0, CGVP_K_CONV_BWD = 1, CGVP_K_ROWS_FWD = 2, CGVP_K_ROWS_BWD = 3, CGVP_K_SEGMENT_REDUCE = 4,
       CGVP_K_GATHER = 5, CGVP_K_FEATURIZE = 6, CGVP_K_COUNT = 7 };
int32_t cgvp_profile_enable(int32_t on);
int32_t cgvp_profile_collect(int32_t kernel_id, double* total_ms, int64_t* launches);

/* ---- weight packing ----------------------------------------------------------------------------------------
 * Kernels consume weights in a packed, zero-padded, K-major layout (and its transpose for the backward data path).
 * cgvp_gvp_packed_floats: floats of one packed GVP block.  cgvp_pack_weights: PyTorch layout -> packed blocks
 * (one launch for n GVPs; `packed[i]` points at a block of cgvp_gvp_packed_floats(&descs[i]) floats, 16-B aligned).
 * cgvp_unpack_grads: packed gradient blocks (same layout, forward half) -> PyTorch-layout gradient tensors. */
int64_t cgvp_gvp_packed_floats(const CgvpGvpDesc* desc);
int32_t cgvp_pack_weights(int32_t n, const CgvpGvpDesc* h_descs, const CgvpGvpWeights* h_weights,
                          float* const* h_packed, cgvp_stream_t stream);
int32_t cgvp_unpack_grads(int32_t n, const CgvpGvpDesc* h_descs, const float* const* h_packed_grads,
                          const CgvpGvpGrads* h_grads, cgvp_stream_t stream);

/* ---- graph plan ---------------------------------------------------------------------------------------------
 * Replaces the implicit index handling of PyG MessagePassing.propagate (used at gvp_layers.py:298-300): a stable
 * sort of the edges by target gives a CSR view (deterministic segmented aggregation instead of atomic
 * scatter_add_), and a second CSR view by source serves the backward pass.
 *   perm      [E] int32: dst-sorted position p -> original edge id
 *   src, dst  [E] int32: endpoints in dst-sorted order
 *   rowptr    [N+1] int32: edges with target n are positions rowptr[n] .. rowptr[n+1]-1
 *   sperm     [E] int32: source-sorted position q -> dst-sorted position p
 *   srowptr   [N+1] int32: edges with source n are sperm[srowptr[n] .. srowptr[n+1]-1]                          */
typedef struct CgvpPlan {
    int64_t num_edges, num_nodes;
    int32_t *perm, *src, *dst, *rowptr, *sperm, *srowptr;
} CgvpPlan;

int64_t cgvp_plan_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int32_t cgvp_plan_build(const int64_t* edge_index, const CgvpPlan* plan, void* ws, int64_t ws_bytes,
                        cgvp_stream_t stream);

/* ---- standalone gather / segmented reduce (the two HBM-bound primitives) ------------------------------------
 * cgvp_gather_message_input materialises what GVPConv.message builds with tuple_cat (gvp_layers.py:303-306):
 *   ms[e] = [s[src_e] ; es[e] ; s[dst_e]]  (2*ns+es)      mv[e] = [V[src_e] ; eV[e] ; V[dst_e]]  ((2*nv+ev) x 3)
 * in ORIGINAL edge order.  cgvp_segment_reduce is the aggregation of PyG propagate (sum / mean over the target,
 * dim_size = N) done as a deterministic CSR segmented reduction: out[n] = sum_{p in seg(n)} rows[index[p]]
 * (index == NULL means identity); mean divides by max(deg,1).  `beta` = 1 accumulates into out. */
int32_t cgvp_gather_message_input(const int64_t* edge_index, int64_t num_edges, int32_t ns, int32_t nv, int32_t es,
                                  int32_t ev, const float* s, const float* v, const float* e_s, const float* e_v,
                                  float* ms, float* mv, cgvp_stream_t stream);
int32_t cgvp_segment_reduce(const float* rows, int32_t width, const int32_t* rowptr, const int32_t* index,
                            int64_t num_nodes, int32_t aggr, int32_t beta, float* out, cgvp_stream_t stream);

/* ---- fused GVPConv -------------------------------------------------------------------------------------------
 * Replaces GVPConv.forward + message + PyG gather/aggregate (gvp_layers.py:291-308): per dst-sorted edge tile,
 * gather (s_j, e_s, s_i), (V_j, e_V, V_i) into shared memory, run the n_gvp stacked message GVPs, and reduce the
 * messages per target node.  Nothing of size E is written.
 *   gvp[0] must map (2*ns+es, 2*nv+ev) -> ...; the last GVP must produce (ns_out, nv_out).
 *   edge_sorted != 0: e_s/e_v (and d_e_s/d_e_v) are already in dst-sorted order (plan.perm is not applied). */
typedef struct CgvpConvDesc {
    int32_t ns, nv, es, ev;       /* node / edge feature dims */
    int32_t n_gvp;                /* 1..CGVP_MAX_CHAIN (message_func, gvp_layers.py:275-289) */
    CgvpGvpDesc gvp[CGVP_MAX_CHAIN];
    int32_t aggr;                 /* CGVP_AGGR_* */
    int32_t edge_sorted;
} CgvpConvDesc;

int64_t cgvp_conv_workspace_bytes(const CgvpConvDesc* desc, int64_t num_edges, int64_t num_nodes, int32_t backward);
int32_t cgvp_conv_fwd(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                      const float* e_s, const float* e_v, const float* const* h_packed, float* out_s, float* out_v,
                      void* ws, int64_t ws_bytes, cgvp_stream_t stream);
/* Backward of the above (autograd through gvp_layers.py:291-308).  d_x_* receive the FULL node gradient (target-
 * side and source-side contributions); d_e_* the edge-attribute gradient (accumulated into if accumulate_edge);
 * h_packed_grads[i] the packed weight-gradient block of GVP i (cgvp_unpack_grads turns it into PyTorch layout). */
int32_t cgvp_conv_bwd(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                      const float* e_s, const float* e_v, const float* const* h_packed, const float* d_out_s,
                      const float* d_out_v, float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v,
                      int32_t accumulate_edge, float* const* h_packed_grads, void* ws, int64_t ws_bytes,
                      cgvp_stream_t stream);
/* Optional training stash: the forward of the specialised kernels can leave the outputs of message GVPs 0 and 1 per
 * (sorted) edge in a caller-owned buffer of cgvp_conv_stash_bytes() bytes, and the backward of the SAME call pair reads them
 * instead of recomputing the first two GVPs (-21 % FMAs in the backward for 224 B/edge of extra traffic).  0 bytes = the
 * descriptor has no such path; stash = NULL reproduces cgvp_conv_fwd / cgvp_conv_bwd. */
int64_t cgvp_conv_stash_bytes(const CgvpConvDesc* desc, int64_t num_edges);
int32_t cgvp_conv_fwd_stash(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                            const float* e_s, const float* e_v, const float* const* h_packed, float* out_s, float* out_v,
                            void* ws, int64_t ws_bytes, void* stash, cgvp_stream_t stream);
int32_t cgvp_conv_bwd_stash(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                            const float* e_s, const float* e_v, const float* const* h_packed, const float* d_out_s,
                            const float* d_out_v, float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v,
                            int32_t accumulate_edge, float* const* h_packed_grads, void* ws, int64_t ws_bytes,
                            const void* stash, cgvp_stream_t stream);

/* ---- fused row program ---------------------------------------------------------------------------------------
 * One kernel for every per-row (per-node or per-edge) stage of the path:
 *     x  = [onehot(types) ; in_s], in_v                          (protein_gnn.py:139-152)
 *     x  = x + mask0 * h            if has_residual_in           (gvp_layers.py:407, Dropout :187-219)
 *     x  = LayerNorm0(x)            if pre_norm                  (gvp_layers.py:231-242)
 *     y  = GVP_{n-1}(...GVP_0(x))   n_gvp in 0..CGVP_MAX_CHAIN   (gvp_layers.py:142-175, ff_func :355-364)
 *     y  = x + mask1 * y            if post_residual             (gvp_layers.py:410)
 *     y  = LayerNorm1(y)            if post_norm
 * covering gvp_node / gvp_edge (+LayerNorm, protein_gnn.py:375-376), the GVPConvLayer node update
 * (gvp_layers.py:407-410), gvp_norm_before_scalar + gvp_to_scalar (protein_gnn.py:385-386), and the stand-alone
 * GVP / LayerNorm modules.  Dropout masks are supplied by the caller already scaled by 1/(1-p) (NULL = identity):
 * mask*_s:[rows,S], mask*_v:[rows,C].  `in_index` (int32[rows], NULL = identity) gathers input rows. */
typedef struct CgvpRowDesc {
    int32_t in_s, in_v;           /* dims of in_s/in_v (before the one-hot prefix) */
    int32_t onehot;               /* number of type classes prepended to the scalars (0 = none) */
    int32_t has_residual_in;
    int32_t pre_norm;
    int32_t n_gvp;
    CgvpGvpDesc gvp[CGVP_MAX_CHAIN];
    int32_t post_residual;
    int32_t post_norm;
} CgvpRowDesc;

typedef struct CgvpRowArgs {
    int64_t rows;
    const float *in_s, *in_v;          /* [rows_in, in_s], [rows_in, in_v, 3] */
    const int64_t* types;              /* [rows_in] (onehot > 0) */
    const int32_t* in_index;           /* optional gather of input rows */
    const float *h_s, *h_v;            /* residual addend (has_residual_in) */
    const float *mask0_s, *mask0_v, *mask1_s, *mask1_v;
    const float *ln0_w, *ln0_b, *ln1_w, *ln1_b;
    const float* const* h_packed;      /* host array of n_gvp device pointers */
    float *out_s, *out_v;
    float* stash;                      /* optional training stash [rows, cgvp_rows_stash_floats(desc)]: the forward leaves the
                                          pre-activation scalars s' of its GVPs, the backward (same pointer) then skips every
                                          W_s projection of its recompute; NULL = recompute */
} CgvpRowArgs;

typedef struct CgvpRowGradArgs {
    const float *d_out_s, *d_out_v;
    float *d_in_s, *d_in_v;            /* may be NULL (inputs are data); scattered through in_index */
    float *d_h_s, *d_h_v;              /* gradient of the residual addend (has_residual_in) */
    float *d_ln0_w, *d_ln0_b, *d_ln1_w, *d_ln1_b;
    float* const* h_packed_grads;
} CgvpRowGradArgs;

int64_t cgvp_rows_workspace_bytes(const CgvpRowDesc* desc, int64_t rows, int32_t backward);
/* floats per row of the training stash, 0 if the kernels serving `desc` keep none */
int64_t cgvp_rows_stash_floats(const CgvpRowDesc* desc);
int32_t cgvp_rows_fwd(const CgvpRowDesc* desc, const CgvpRowArgs* args, void* ws, int64_t ws_bytes,
                      cgvp_stream_t stream);
int32_t cgvp_rows_bwd(const CgvpRowDesc* desc, const CgvpRowArgs* args, const CgvpRowGradArgs* grads, void* ws,
                      int64_t ws_bytes, cgvp_stream_t stream);

/* ---- residue-graph featurizer --------------------------------------------------------------------------------
 * Replaces compute_residue_edge_features (utils/create_protein_features.py:201-357) + construct_graph
 * (utils/create_graphs.py:6-62) for a BATCH of proteins, sparse from the start (no n x n x 35 fp64 array):
 *   ca:[N,3] fp32 C-alpha coordinates of all proteins back to back, ptr:[B+1] int64 protein boundaries.
 *   thresh_type 0 = 'dist' (d <= thresh), 1 = 'num' (k nearest per source row), 2 = 'prop' (ceil(thresh*n)).
 * Phase 1 (count) writes row_offsets:[N+1] int64 (exclusive scan of edges per source residue; row_offsets[N] = E).
 * The caller reads E back, allocates, and calls phase 2 (fill), which writes edge_index:[2,E] int64 with batch-
 * global node ids sorted by (src,dst), edge_s:[E,32] = [RBF16 ; cos8 ; sin8], edge_v:[E,1,3] = unit(CA_src-CA_dst).
 * Distances are evaluated in fp64 with separately rounded operations so the edge set is bit-identical. */
int64_t cgvp_featurize_workspace_bytes(int64_t num_nodes, int64_t max_protein_len);
int32_t cgvp_featurize_count(const float* ca, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                             int64_t max_protein_len, double thresh, int32_t thresh_type, int32_t keep_self_loops,
                             int64_t* row_offsets, void* ws, int64_t ws_bytes, cgvp_stream_t stream);
int32_t cgvp_featurize_fill(const float* ca, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                            int64_t max_protein_len, double thresh, int32_t thresh_type, int32_t keep_self_loops,
                            const int64_t* row_offsets, int64_t* edge_index, int64_t num_edges, float* edge_s,
                            float* edge_v, void* ws, int64_t ws_bytes, cgvp_stream_t stream);

/* ---- residue node features ------------------------------------------------------------------------------------
 * Replaces compute_residue_node_features(vectorize_features=True, add_esm2_embeds=False)
 * (utils/create_protein_features.py:12-198) for a batch of proteins:
 *   res_coords:[N,4,3] fp32 backbone atoms (N, CA, C, O) of all proteins back to back, ptr:[B+1] int64.
 *   out_s:[N, 6 + num_props + 16*add_posenc] = [cos(phi,psi,omega) ; sin(phi,psi,omega) ; aa_table[idents] ; pos-enc],
 *   out_v:[N,3,3] = [forward ; backward ; virtual side chain] unit-vector features.
 * aa_table:[num_types,num_props] fp32 is the caller's amino-acid property table (utils/protein_definitions.py:71-277,
 * min-max normalised as the reference does); num_props = 0 skips the look-up (include_aa_props=False). */
int32_t cgvp_node_features(const float* res_coords, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                           const int64_t* idents, const float* aa_table, int32_t num_types, int32_t num_props,
                           int32_t add_posenc, float* out_s, float* out_v, cgvp_stream_t stream);

/* ---- cross-attention core on packed rows ---------------------------------------------------------------------
 * The scaled-dot-product stage of the two nn.MultiheadAttention modules of CrossAttentionModule
 * (models/joint_gnn.py:350-361; key_padding_mask = the other side's padding) without padding:
 *   q:[Nq, H*D]  k, v:[Nk, H*D] fp32 packed rows; graph g owns query rows [qptr[g], qptr[g+1]) and key rows
 *   [kptr[g], kptr[g+1]); qbatch:[Nq] / kbatch:[Nk] int64 graph id per row.
 *   out:[Nq, H*D]; stats:[Nq, H, 2] (row max, row sum) for the backward; weights:[B, lq_max, lk_max] (zero-filled
 *   by the caller) receives the head-averaged probabilities (need_weights=True) or is NULL; q_fill:[H*D] + w_fill:
 *   [B, lk_max] optionally give the map row of a PADDED query (the reference computes those from padding).
 * cgvp_attn_bwd recomputes the probabilities and writes dq, dk, dv (dsum:[Nq, H] is scratch).  Deterministic. */
int32_t cgvp_attn_supported(int32_t num_heads, int32_t head_dim);
int32_t cgvp_attn_fwd(const float* q, const float* k, const float* v, const int64_t* qptr, const int64_t* kptr,
                      const int64_t* qbatch, int64_t num_graphs, int64_t num_q, int64_t num_k, int32_t num_heads,
                      int32_t head_dim, float scale, const float* q_fill, int32_t lq_max, int32_t lk_max, float* out,
                      float* stats, float* weights, float* w_fill, cgvp_stream_t stream);
int32_t cgvp_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* stats,
                      const float* d_out, const int64_t* qptr, const int64_t* kptr, const int64_t* qbatch,
                      const int64_t* kbatch, int64_t num_graphs, int64_t num_q, int64_t num_k, int32_t num_heads,
                      int32_t head_dim, float scale, float* dsum, float* dq, float* dk, float* dv, cgvp_stream_t stream);

/* ---- dense-layer weight / bias gradient on the tensor cores ---------------------------------------------------
 * dW[N,K] = dY[M,N]^T X[M,K], db[N] = column sums of dY: the autograd of the nn.Linear layers around the encoder
 * (models/joint_gnn.py:172-288,321-408) for M = all residues of a batch.  fp32-accurate (3xTF32 split precision,
 * tcgen05 kind::tf32 with MN-major swizzled operands), deterministic.  db may be NULL.  Row-major, contiguous,
 * 16-byte aligned buffers; supported: M >= 1024, N % 128 == 0 (<= 1024), K % 32 == 0 (<= 256). */
int32_t cgvp_linear_wgrad_supported(int64_t M, int32_t N, int32_t K);
int64_t cgvp_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K);
int32_t cgvp_linear_wgrad(const float* dy, const float* x, int64_t M, int32_t N, int32_t K, float* dw, float* db,
                          void* ws, int64_t ws_bytes, cgvp_stream_t stream);
/* y[M,N] = x[M,K] w[N,K]^T + bias (nn.Linear forward) and dx[M,K] = dy[M,N] w[N,K] (its input gradient), same 3xTF32
 * arithmetic, persistent CTAs with the weight slice resident in shared memory.  cgvp_linear_gemm_supported(M, out, in)
 * answers for y = A B^T with `out` result columns and `in` reduction length (forward: (M, N, K); dgrad: (M, K, N)). */
int32_t cgvp_linear_gemm_supported(int64_t M, int32_t out_cols, int32_t red_len);
int32_t cgvp_linear_fwd(const float* x, const float* w, const float* bias, int64_t M, int32_t N, int32_t K, float* y,
                        cgvp_stream_t stream);
int32_t cgvp_linear_dgrad(const float* dy, const float* w, int64_t M, int32_t N, int32_t K, float* dx, cgvp_stream_t stream);

/* ---- row LayerNorm --------------------------------------------------------------------------------------------
 * nn.LayerNorm(D) (eps inside the square root, affine) on packed [rows, D] activations, D in {32, 64, 128, 256}: the
 * preattn_norm / ff_norm layers of CrossAttentionModule (models/joint_gnn.py:321-408).  stats:[rows, 2] = (mean, rstd).
 * The backward writes dx and the gamma / beta gradients (per-CTA partials summed in a fixed order: deterministic). */
int32_t cgvp_layernorm_supported(int32_t D);
int64_t cgvp_layernorm_workspace_bytes(int64_t rows, int32_t D);
int32_t cgvp_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int32_t D, float eps,
                           float* y, float* stats, cgvp_stream_t stream);
int32_t cgvp_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, int64_t rows, int32_t D,
                           float* dx, float* dgamma, float* dbeta, void* ws, int64_t ws_bytes, cgvp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CASTERGVP_H */
