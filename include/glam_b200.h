/*
 * glam_b200 — C ABI of the B200-native (sm_100a) GLAM message-passing hot path.
 *
 * This is the drop-in boundary: plain device pointers + sizes + a cudaStream_t, no torch types.
 * The reference (yvquanli/GLAM) is pure Python; its hot path sits behind the `layer.py` module
 * namespace.  `glam_b200/layer.py` mirrors that namespace and binds these entry
 * points with ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`;
 *   - matrices are row-major fp32 with an explicit leading dimension (`ld*`, in elements);
 *   - indices inside the library are int32; `edge_index` / `batch` come in as the reference's int64;
 *   - inputs are borrowed and never written; outputs / workspaces are caller-allocated;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - return value: 0 = ok, <0 = invalid argument (see glam_last_error()), >0 = cudaError_t;
 *   - no atomics on floating point anywhere: every reduction has a fixed order, results are
 *     bitwise reproducible run to run.
 *
 * Reference citations are `path:line` under the upstream tree (yvquanli/GLAM).
 */
#ifndef GLAM_B200_H
#define GLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLAM_B200_ABI_VERSION 14
#define GLAM_MAX_HEADS 4

int glam_abi_version(void);
/* Human-readable description of the last non-zero status returned on this host thread. */
const char* glam_last_error(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
int64_t glam_launch_count(void);
/* Arithmetic of the dense projections: 1 (default) = TF32 operands on the tcgen05 tensor cores with fp32
 * accumulation (what torch 1.10, the reference's pinned version, does by default for fp32 matmul on Ampere+);
 * 0 = exact fp32 FMAs on the CUDA cores.  Everything else (softmax, gates, reductions) is always fp32. */
int glam_set_math_mode(int mode);
int glam_get_math_mode(void);

/* ---------------------------------------------------------------------------------------------
 * (1) Destination-sorted CSR builder.
 * Replaces PyG's per-call `index_select(edge_index[0|1])` + atomic `scatter` (MessagePassing.propagate,
 * reached from src_1gp/layer.py:40,86) with one index build per batch, reused by every message step
 * and by backward.  Bit-exact against `torch.argsort(edge_index[1], stable=True)`:
 *   dst_perm[p]   original edge id of the p-th edge in (dst, original order) order
 *   dst_rowptr[i] .. dst_rowptr[i+1]  = in-edges of node i in that order
 *   dst_src[p]    source node of edge dst_perm[p]
 *   dst_dst[p]    destination node of that edge (optional, may be NULL; lets backward run edge-parallel)
 * and the same for sources (backward scatters by SOURCE):
 *   src_perm / src_rowptr  (stable by edge_index[0]);  src_pos[k] = position p (dst order) of edge
 *   src_perm[k];  src_dst[k] = destination node of that edge.
 * workspace: >= glam_csr_workspace_bytes(N, E) bytes.
 * --------------------------------------------------------------------------------------------- */
size_t glam_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int glam_build_csr(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                   int32_t* dst_rowptr, int32_t* dst_src, int32_t* dst_perm, int32_t* dst_dst,
                   int32_t* src_rowptr, int32_t* src_pos, int32_t* src_dst,
                   void* workspace, size_t workspace_bytes, void* stream);
/* graph_ptr[g] = first node of graph g (B+1 entries) from the sorted int64 `batch` vector
 * (PyG Batch.batch; replaces np.bincount(batch.cpu()) + cumsum, src_2gi_ddi/layer.py:271-272). */
int glam_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* graph_ptr, void* stream);
/* out[p, :] = in[perm[p], :]  (edge_attr into dst order, once per batch). */
int glam_gather_rows(const float* in, const int32_t* perm, int64_t rows, int64_t cols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (2) Small dense contractions (projections).  Y[M,N] = epilogue(X[M,K] * W + bias).
 * W(k,n) is addressed as W[k*w_sk + n*w_sn] so both `x @ W` (src_1gp/layer.py:37,59) and
 * `x @ W^T` (torch.nn.GRU / Linear weights) are served without a transpose copy.
 * epilogue: 0 none | 1 celu(alpha=1) (src_1gp/layer.py:261) | 2 Y = acc * celu'(aux) with aux the
 * saved celu OUTPUT | 3 Y += acc.
 * --------------------------------------------------------------------------------------------- */
int glam_gemm(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
              const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
              int epilogue, void* stream);
/* Same, with output columns [exact_col_begin, exact_col_end) always computed with exact fp32 FMAs even in TF32 math
 * mode: the attention-logit columns s_i | s_j of the extended node projection feed a softmax whose gradients are
 * zero-sum per destination, so operand rounding there is amplified in the attention-vector gradient. */
int glam_gemm_ex(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                 const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                 int epilogue, int exact_col_begin, int exact_col_end, void* stream);
/* Weight gradients: out[Ka,Kb] = sum_m A[m,Ka] * B[m,Kb]; fixed-order two-stage reduction.
 * workspace >= glam_gemm_tn_workspace_bytes(M, Ka, Kb). */
size_t glam_gemm_tn_workspace_bytes(int64_t M, int64_t Ka, int64_t Kb);
int glam_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb,
                 float* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream);
/* Same contraction with the bias gradient folded in: out = A^T B (written transposed when transpose_out != 0, i.e.
 * directly in torch.nn.GRU's [3C,C] weight layout) and, when colsum_b != NULL, colsum_b[Kb] = sum_m B[m,:].  In TF32
 * math mode this runs on the tensor cores (rows are the MMA K dimension, operands MN-major) and the column sums come
 * from a constant-1 feature appended to A; otherwise it is the exact fp32 path.  Deterministic either way. */
size_t glam_gemm_tn_ex_workspace_bytes(int64_t M, int64_t Ka, int64_t Kb, int want_colsum);
int glam_gemm_tn_ex(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb,
                    float* out, int64_t ldo, int transpose_out, float* colsum_b, void* workspace, size_t workspace_bytes,
                    void* stream);
/* Bias gradients: out[n] = sum_m G[m,n]; fixed order. workspace >= glam_colsum_workspace_bytes(M, N). */
size_t glam_colsum_workspace_bytes(int64_t M, int64_t N);
int glam_colsum(const float* G, int64_t ldg, int64_t M, int64_t N, float* out, void* workspace,
                size_t workspace_bytes, void* stream);

/* Tile descriptors of the windowed edge kernels (once per batch, next to the CSR).  Nodes are cut into tiles of
 * glam_edge_tile_rows(N) consecutive ids; a PyG batch is block-diagonal with graph-by-graph numbering
 * (src_1gp/dataset.py:75-87 + Batch.from_data_list), so the sources of a tile's in-edges (and the destinations of its
 * out-edges) span one short contiguous range of rows, which the kernels stage in shared memory with bulk async copies
 * instead of gathering row by row.  dst_tiles[t] = {lo, hi, e0, e1}: rows [lo,hi) cover the tile's own rows and every
 * source of its in-edges [e0,e1) (dst order); src_tiles[t] = {lo, hi, k0, k1}: rows [lo,hi) cover the destinations of
 * its out-edges [k0,k1) (src order).  Both arrays: 4 * glam_edge_tile_count(N) int32, 16-byte aligned; src_tiles may be
 * NULL (forward only).  The edge entry points below take them as an optional argument (NULL = per-edge gather path);
 * tiles whose window does not fit in shared memory are handled inside the same launch from global memory. */
int glam_edge_tile_rows(int64_t num_nodes);
int64_t glam_edge_tile_count(int64_t num_nodes);
int glam_build_edge_tiles(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr,
                          const int32_t* src_dst, int64_t num_nodes, int64_t num_edges, int32_t* dst_tiles,
                          int32_t* src_tiles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Triplet attention edge phase — TripletMessage.message + aggregate (src_1gp/layer.py:42-55, :17)
 * and TripletMessageLight.message (:88-97), restructured per SURVEY.md Appendix C.
 *
 * xpe [N, ldxp]: columns [0,HC) = x@weight_node, [HC,HC+H) = s_i (= <a_i, xp_h>), [HC+H,HC+2H) = s_j.
 * edge_attr is in DST order ([E,De]); w_edge [De,HC] (NULL for the Light layer: e_ij == 1 in the message);
 * att_edge [De,H] = per-head dot of w_edge rows with the edge slice of weight_triplet_att (Light: the raw
 * edge slice).  One warp per destination node: logits -> leaky_relu -> PyG softmax
 * exp(a-max)/(sum+1e-16) -> agg[i] = sum_e alpha * e_ij (.) x_j.  Writes alpha [E,H] (dst order).
 * --------------------------------------------------------------------------------------------- */
int glam_triplet_edge_fwd(const float* xpe, int64_t ldxp, const float* edge_attr, const float* w_edge,
                          const float* att_edge, const int32_t* dst_rowptr, const int32_t* dst_src,
                          const int32_t* dst_tiles, int64_t num_nodes, int64_t num_edges, int heads, int channels, int edge_dim,
                          float negative_slope, float* agg, float* alpha, void* stream);
/* Backward, destination pass: from g_agg [N,HC] computes g_logit [E,H] (dst order), g_s_i into
 * g_xpe[:, HC:HC+H], and per-warp partials of g_w_edge reduced in fixed order into g_w_edge [De,HC]
 * (skipped when w_edge == NULL).  workspace >= glam_triplet_bwd_workspace_bytes(...). */
size_t glam_triplet_bwd_workspace_bytes(int heads, int channels, int edge_dim);
int glam_triplet_edge_bwd_dst(const float* xpe, int64_t ldxp, const float* edge_attr, const float* w_edge,
                              const float* att_edge, const float* alpha, const float* g_agg,
                              const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* dst_dst,
                              const int32_t* dst_tiles, int64_t num_nodes, int64_t num_edges, int heads, int channels,
                              int edge_dim, float negative_slope, float* g_logit, float* g_xpe, float* g_w_edge,
                              void* workspace, size_t workspace_bytes, void* stream);
/* Backward, source pass: g_xpe[j, 0:HC] = sum over out-edges of alpha * g_agg[dst] (.) e_ij,
 * g_xpe[j, HC+H:HC+2H] = sum g_logit, pad columns zeroed. */
int glam_triplet_edge_bwd_src(const float* edge_attr, const float* w_edge, const float* alpha, const float* g_agg,
                              const float* g_logit, const int32_t* src_rowptr, const int32_t* src_pos,
                              const int32_t* src_dst, const int32_t* src_tiles, int64_t num_nodes, int64_t num_edges,
                              int heads, int channels, int edge_dim, float* g_xpe, int64_t ldxp, void* stream);

/* Derived weights of the triplet layers (parameter space, tiny): w_ext [C, ldxp] = weight_node | weight_node_h a_i,h |
 * weight_node_h a_j,h | 0 and att_edge [De,H] = weight_edge_h a_e,h, with a_i|a_e|a_j the three slices of
 * weight_triplet_att (concat order of src_1gp/layer.py:48).  light != 0: TripletMessageLight (:92), att is
 * [2C+De], weight_edge unused.  The backward chains (g_w_ext, g_att_edge, optional direct g_w_edge) to the
 * reference's parameters weight_node, weight_edge, weight_triplet_att. */
int glam_triplet_prep_fwd(const float* weight_node, const float* weight_edge, const float* att, int channels, int heads,
                          int edge_dim, int light, int ldxp, float* w_ext, float* att_edge, void* stream);
int glam_triplet_prep_bwd(const float* weight_node, const float* weight_edge, const float* att, const float* g_w_ext,
                          const float* g_att_edge, const float* g_w_edge_direct, int channels, int heads, int edge_dim,
                          int light, int ldxp, float* g_weight_node, float* g_weight_edge, float* g_att, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (4) GRU node update of MessageBlock (src_1gp/layer.py:260-266): gates from gi = m W_ih^T + b_ih,
 * gh = h W_hh^T + b_hh (both [N,3C], gate order r,z,n), h' = (1-z) n + z h, x_out = act(h' + identity).
 * In place: gi is overwritten with [r|z|n] (saved for backward); gh's n-chunk is kept.
 * act: 0 none | 1 relu | 2 leaky_relu(act_param) | 3 celu(alpha=1).  identity may be NULL (res=False).
 * --------------------------------------------------------------------------------------------- */
int glam_gru_gates_fwd(float* gi_rzn, const float* gh, const float* h, const float* identity, int64_t num_nodes,
                       int channels, int act, float act_param, float* h_new, float* x_out, void* stream);
/* Backward: g_x_out, g_h_carry (may be NULL) -> g_gi, g_gh [N,3C], g_h_prev (= g_h' * z), g_identity (may be NULL). */
int glam_gru_gates_bwd(const float* rzn, const float* gh, const float* h, const float* x_out, const float* g_x_out,
                       const float* g_h_carry, int64_t num_nodes, int channels, int act, float act_param,
                       float* g_gi, float* g_gh, float* g_h_prev, float* g_identity, void* stream);
/* Same with the n-gate hidden pre-activation passed on its own: row n at gh_n + n*ld_ghn (glam_gru_gates_bwd is this with
 * gh_n = gh + 2C, ld_ghn = 3C); glam_gru_fused_fwd writes gh_n as a compact [N,C] tensor. */
int glam_gru_gates_bwd_ex(const float* rzn, const float* gh_n, int64_t ld_ghn, const float* h, const float* x_out,
                          const float* g_x_out, const float* g_h_carry, int64_t num_nodes, int channels, int act,
                          float act_param, float* g_gi, float* g_gh, float* g_h_prev, float* g_identity, void* stream);
/* LSTM cell gates for Set2Set (torch.nn.LSTM inside PyG Set2Set; src_1gp/model.py:41): gates [B,4C] (i,f,g,o,
 * pre-activation, overwritten with the activated gates), c_prev -> c_new, h_new. */
int glam_lstm_gates_fwd(float* gates, const float* c_prev, int64_t rows, int channels, float* c_new, float* h_new,
                        void* stream);
int glam_lstm_gates_bwd(const float* gates_act, const float* c_prev, const float* c_new, const float* g_h,
                        const float* g_c_in, int64_t rows, int channels, float* g_gates, float* g_c_prev,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Per-graph attention pooling shared by GlobalLAPool/GlobalAttention (src_1gp/layer.py:206-220) and
 * one Set2Set step: e[n] = <x[n], q[g]> + q_bias, a = PyG softmax over the nodes of graph g,
 * r[g] = sum_n a[n] x[n], asum[g] = sum_n a[n].   q_stride = 0 shares one q between graphs.
 * --------------------------------------------------------------------------------------------- */
int glam_seg_attn_pool_fwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* q_bias,
                           const int32_t* graph_ptr, int64_t num_graphs, int channels,
                           float* a, float* r, int64_t ldr, float* asum, void* stream);
/* Backward: g_r [B,*], g_asum (may be NULL) -> g_x (+= if accumulate), g_q per graph [B,C], g_e [N]. */
int glam_seg_attn_pool_bwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* a,
                           const float* g_r, int64_t ldgr, const float* g_asum, const int32_t* graph_ptr,
                           int64_t num_graphs, int channels, int accumulate,
                           float* g_x, int64_t ldgx, float* g_q, float* g_e, void* stream);

/* Set2Set rounds (PyG Set2Set(in_channels=C, processing_steps=S) @1.7.2, src_1gp/model.py:2,41).  A round is
 *   gates = U [B,3C] x [W_ih | W_hh]^T + (b_ih + b_hh)         -- glam_gemm_ex (shared weights: tensor cores)
 *   glam_set2set_round_fwd: gate non-linearities (i,f,g,o), c' = f c + i g, h = o tanh c', a = softmax_n <x_n, h>
 *       (PyG softmax), r = sum_n a_n x_n, u_next = [h | r | h] (the next round's GEMM operand; q* = [h | r] is also
 *       written to q_star when non-NULL).  gates is activated in place; att [N] keeps this round's weights.
 * One warp per graph.  Backward per round (reverse order): g_u is the gradient of the round's output ([h|r], 2C columns,
 * for the last round; [h|r|h], 3C columns = G_next x [W_ih|W_hh], otherwise); g_c carries the cell-state gradient;
 * g_x is overwritten (accumulate == 0) or added to; G [B,4C] receives the gate pre-activation gradients, from which
 * [g_w_ih | g_w_hh] = sum_rounds G^T U and g_b_ih = g_b_hh = colsum(G) follow as one glam_gemm_tn_ex. */
int glam_set2set_round_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels,
                           float* gates, const float* c_prev, float* c_new, float* att, float* u_next, float* q_star,
                           void* stream);
int glam_set2set_round_bwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels,
                           const float* gates, const float* c_prev, const float* c_new, const float* att,
                           const float* g_u, int64_t ldgu, int gu_cols, float* g_c, float* g_x, int accumulate,
                           float* G, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (6) Cross-graph interaction pool — dot_and_global_pool2 (src_2gi_ddi/layer.py:270-283, identical in
 * src_2gi_dti_scr/layer.py): per pair g, S = Xa[g] Xb[g]^T, out[g] = [max S, mean S]; one CTA per
 * pair, no host sync.  Saves argmax (a,b) as node indices and the per-graph column sums for backward.
 * --------------------------------------------------------------------------------------------- */
int glam_pair_dot_pool_fwd(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                           int64_t num_pairs, int channels, float* out, int32_t* argmax, float* sum_a,
                           float* sum_b, void* stream);
/* Shared second operand (evaluation): pair g reads graph idx_b[g] of the b side, so a virtual-screening
 * batch holds every distinct protein graph ONCE (LIT-PCBA has one protein per target, src_2gi_dti_scr/dataset.py:297; the
 * reference collates one copy per pair, dataset.py:329-335).  ptr_b has one entry per DISTINCT graph (+1).  Forward only. */
int glam_pair_dot_pool_fwd_idx(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                               const int32_t* idx_b, int64_t num_pairs, int channels, float* out, int32_t* argmax,
                               float* sum_a, float* sum_b, void* stream);
/* Tensor-core variant of the forward for pairs whose second side is large (drug-target: 25 x ~500 x C is a real GEMM): S on
 * tcgen05 with TF32 operands ([128 x 256] accumulator tiles in TMEM, thread = ligand row scans for the max / first arg-max);
 * the mean and sum_a / sum_b stay exact fp32.  idx_b may be NULL.  Needs tf32 math mode, channels % 4 == 0 in [32, 64]. */
int glam_pair_dot_pool_tc_supported(int channels);
/* Small pairs (drug-drug, ~25 x 25 atoms; ABI v13): one warp per pair, eight pairs per CTA, the second graph's row in registers and
 * the first graph's tile transposed in shared memory — same values, arg-max rule and outputs as glam_pair_dot_pool_fwd (idx_b may
 * be NULL); channels in {32, 36, 48, 64}, 16-byte aligned rows.  Graphs over 32 rows are handled (32-row tiles) but belong on the
 * CTA-per-pair or the tensor-core entry point. */
int glam_pair_dot_pool_small_supported(int channels);
int glam_pair_dot_pool_fwd_small(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b, const int32_t* idx_b,
                                 int64_t num_pairs, int channels, float* out, int32_t* argmax, float* sum_a, float* sum_b,
                                 void* stream);
int glam_pair_dot_pool_fwd_tc(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                              const int32_t* idx_b, int64_t num_pairs, int channels, float* out, int32_t* argmax,
                              float* sum_a, float* sum_b, void* stream);
int glam_pair_dot_pool_bwd(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                           const float* g_out, const int32_t* argmax, const float* sum_a, const float* sum_b,
                           int64_t num_pairs, int channels, float* g_xa, float* g_xb, void* stream);

/* Fused GRU update (tf32 math mode, channels in {32,36,40,44}): both gate GEMMs (m W_ih^T, h W_hh^T) on the tensor cores
 * into one TMEM accumulator (columns permuted per channel to r_pre, z_pre, gi_n, gh_n; the h-side product accumulates
 * onto the m-side one), and the gate arithmetic of glam_gru_gates_fwd in the epilogue — the [N,3C] pre-activations
 * never reach HBM.  Writes rzn [N,3C] (r|z|n, saved for backward), gh_n [N,C] (for glam_gru_gates_bwd_ex), h_new and
 * x_out [N,C] (contiguous).  glam_gru_fused_supported(channels) tells whether the current math mode / shape can use it;
 * otherwise call glam_gemm_ex twice + glam_gru_gates_fwd. */
int glam_gru_fused_supported(int channels);
int glam_gru_fused_fwd(const float* m, int64_t ldm, const float* h, int64_t ldh, const float* identity,
                       const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int64_t N,
                       int channels, int act, float act_param, float* rzn, float* gh_n, float* h_new, float* x_out,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * (6b) "Next" rows of the scope table: the reference's default readout and its GCN tower.
 *
 * glam_pool5_fwd/bwd — GlobalPool5 (src_1gp/layer.py:197-203; default mol_readout, src_1gp/run.py:25):
 *   out [B,5C] = [mean | sum | PyG global_sort_pool(k=3): rows sorted by the last channel, descending, ties by node
 *   order, missing rows zero]; top_idx [B,3] = chosen node ids (-1 = missing), saved for backward.
 * glam_csr_aggregate — out[i,:] = (accumulate ? out[i,:] : 0) + row_scale[i] * (sum_{p in row i} edge_w[p] * Y[col[p],:]
 *   + self_w[i] * Y[i,:]) + bias; edge_w / self_w / row_scale / bias may be NULL (1 / 0 / 1 / 0).  Deterministic.
 * glam_gcn_norm — PyG GCNConv(in,out) normalisation @1.7.2 (`_GCNConv`, src_1gp/layer.py:143-149; default pro_block,
 *   src_2gi_dti_scr/run.py:19): self edges are replaced by one unit self loop per node, dinv[i] = (1 + non-self
 *   in-degree)^-1/2, w_dst[p] = dinv[dst] dinv[src] in destination order (0 on self edges), w_src likewise in source order
 *   (for the transposed aggregation of backward; may be NULL).  out = aggregate(x W, w_dst, self_w = dinv^2) + bias.
 * --------------------------------------------------------------------------------------------- */
int glam_pool5_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels, float* out,
                   int32_t* top_idx, void* stream);
int glam_pool5_bwd(const float* g_out, const int32_t* graph_ptr, const int32_t* top_idx, int64_t num_graphs, int channels,
                   float* g_x, int64_t ldgx, void* stream);
int glam_csr_aggregate(const float* Y, int64_t ldy, const int32_t* rowptr, const int32_t* col, const float* edge_w,
                       const float* self_w, const float* row_scale, const float* bias, int64_t num_rows, int features,
                       float* out, int64_t ldo, int accumulate, void* stream);
int glam_gcn_norm(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr, const int32_t* src_dst,
                  int64_t num_nodes, float* dinv, float* w_dst, float* w_src, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (7) Optimizer step of the data-parallel training step — torch.optim.Adam(model.parameters(), lr) as the reference
 * trainer builds it (src_1gp/trainer.py:49-50): one pass over flat fp32 buffers (parameters, the all-reduced gradient
 * bucket, both moments).  lr [1] and state [3] = {step, 1-beta1^step, 1-beta2^step} are device memory (CUDA-graph
 * replayable; state must start as zeros); the call first advances state, then updates.  grad_scale multiplies the
 * gradients on load (1/world_size after a sum all-reduce).
 * --------------------------------------------------------------------------------------------- */
int glam_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                   float* state, float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (8) The message stack as ONE kernel (csrc/mp_fused.cu): `steps` applications of the weight-tied MessageBlock —
 * TripletMessage (src_1gp/layer.py:36-61) -> CELU -> GRU (:260-263) -> +identity -> act (:265-266), the loop of
 * src_1gp/model.py:53-54 — on graph-aligned tiles.  A PyG batch is block-diagonal (src_1gp/dataset.py:75-87 +
 * Batch.from_data_list), so a tile of whole graphs (<= 128 nodes) is closed under message passing: projections run on
 * the tcgen05 tensor cores out of shared memory, the edge softmax / aggregation reads the projected rows there, and x, h
 * stay resident from step to step.  xp [N,HC+2H] and agg [N,HC] never reach HBM (unless saved for backward).
 *
 * glam_build_graph_tiles — greedy packing of consecutive whole graphs (graph_ptr, B+1 entries) into tiles of
 *   <= max_nodes rows and <= max_edges in-edges (glam_graph_tile_caps), 512 graphs per packing chunk (tiles do not span
 *   chunks).  tiles: 4*B int32 {n0, n1, e0, e1} per tile, 16-byte aligned; workspace >=
 *   glam_graph_tiles_workspace_bytes(B); meta: int32[4], ZEROED by the caller: meta[0] = tile count, meta[1] = OR of precondition violations
 *   (1 a graph has more rows than a tile, 2 more in-edges than a tile, 4 an edge crosses tiles i.e. graphs,
 *   8 an edge_attr row is not one-hot).  glam_edge_types — bond type per dst-ordered edge (index of the 1 in the
 *   one-hot row, src_1gp/dataset.py:82), ORs 8 into meta[1] otherwise.
 * glam_message_stack_fwd — w_ext / att_edge are glam_triplet_prep_fwd's derived weights.  h0 NULL: h = x on the first
 *   step (layer.py:253-254).  x_raw != NULL: the model's input LinearBlock (src_1gp/model.py:49) is applied while the rows are
 *   loaded — x0 = act(x_raw [N,raw_dim] w_pre^T + b_pre), w_pre [C,raw_dim] as torch.nn.Linear stores it, raw_dim <= 16,
 *   exact fp32 — and x0 may be NULL.  conv_only = 1: x_out = TripletMessage(x0) alone (steps must be 1, GRU arguments NULL).
 *   Eval (save_xpe == NULL): x_out [keep_all ? steps : 1][N][C] = the step outputs (last only unless keep_all), h_out [N][C]
 *   (may be NULL) = the final GRU state.  Training (save_xpe != NULL): what the backward kernels consume is written as
 *   stacked tensors — save_x, save_h [steps+1][N][C] (block inputs / states; entry 0 = x0 / h0), save_xpe [steps][N][ld],
 *   save_agg [steps][N][HC], save_alpha [steps][E][H], save_m [steps][N][C], save_rzn [steps][N][3C], save_gh [steps][N][C]
 *   (conv_only: save_xpe, save_agg, save_alpha and x_out only).  save_gt != NULL (then save_rzn / save_gh may be NULL): the
 *   gate-side tensors for glam_message_stack_bwd in its tile-blocked layout instead — [steps][N][7C]; inside step s the tile
 *   {n0, n1} owns floats [n0*7C, n1*7C) as 7*C/4 slots (r, z, n, gh_n, h, x', m; C/4 16-byte chunks each) of n1-n0 rows x
 *   16 bytes: a warp's 32 rows of one chunk are contiguous in the forward's stores and in the backward's loads; of save_h only
 *   entries 0 and `steps` (the initial and the final state) are written then — the states in between live in save_gt / save_mh.
 *   save_mh != NULL
 *   (then save_m may be NULL): [steps][N][2C+4] rows m | h_in | 1 0 0 0 — with glam_message_stack_bwd's g4 rows the ONE pair of
 *   operands of both GRU weight gradients ([m h 1]^T [g_r g_z g_n g_n r]; the constant column yields the bias gradients).  pn_batch !=
 *   NULL (evaluation only: no saves, h0 == NULL): PairNorm(scale=1, eps=pn_eps) per graph on every step's block input inside the
 *   kernel (_PairNorm, the reference's default graph_norm, src_1gp/layer.py:179-185,255; pn_batch = the int64 `batch` vector);
 *   the residual and the first step's h use the un-normalised rows, as the reference does (layer.py:253,264).  If meta[1] != 0 the outputs are filled with NaN: callers
 *   check meta[1] on the host when they can (outside CUDA-graph capture) and fall back to the per-op entry points.
 *   glam_message_stack_supported: tf32 math mode, heads == 3, channels in {32,36,40}, edge_dim <= 4.
 * --------------------------------------------------------------------------------------------- */
int glam_graph_tile_caps(int* max_nodes, int* max_edges);
/* HOST function (no GPU, no stream; ABI v12): the order of a batch's graphs in which CONSECUTIVE graphs fill the tiles — first-fit
 * decreasing over nodes[B] (and, optionally, in-edges edges[B]); perm[new position] = old graph index.  The kernels' time per
 * tile does not depend on how full the tile is, and a batch has no order of its own (the reference's DataLoader shuffles,
 * src_1gp/trainer.py:95-101): collating in this order makes 815 instead of 918 tiles of 4096 MoleculeNet-shaped graphs.  Empty
 * graphs come first, graphs over cap_nodes last.  O(B); glam_b200/synth.py::tile_order is the numpy statement of the same. */
int glam_tile_order(const int64_t* nodes, const int64_t* edges, int64_t num_graphs, int cap_nodes, int cap_edges, int64_t* perm);
size_t glam_graph_tiles_workspace_bytes(int64_t num_graphs);
int glam_build_graph_tiles(const int32_t* graph_ptr, int64_t num_graphs, const int32_t* dst_rowptr, const int32_t* dst_src,
                           int64_t num_nodes, int64_t num_edges, int32_t* tiles, int32_t* meta, void* workspace,
                           size_t workspace_bytes, void* stream);
/* perm (ABI v14; may be NULL): edge_attr is in the CALLER's edge order and the dst-ordered edge e is its row perm[e] (glam_build_csr's
 * dst_perm) — the evaluation path then never materialises the permuted edge_attr. */
int glam_edge_types(const float* edge_attr, const int32_t* perm, int64_t num_edges, int edge_dim, uint8_t* etype, int32_t* meta,
                    void* stream);
int glam_message_stack_supported(int channels, int heads, int edge_dim);
/* Profiling aid: when `cycles` (device memory, [148][32] unsigned 64-bit) is not NULL, every following glam_message_stack_fwd /
 * glam_message_stack_bwd launch has thread 0 of each CTA add the SM cycles spent between consecutive phase boundaries to
 * cycles[cta][phase] — forward: 0 tile load, 1 logits, 2 softmax, 3 projection wait, 4 TMEM -> xp tile, 5 aggregation, 6 agg
 * panels, 7 scale wait, 8 CELU epilogue, 9 GRU wait, 10 gates, 11 step outputs, 12 tile end, 13 MMA issue, 14 set-up;
 * backward: 16 index words, 17 gate backward, 18 G copy-out + GRU MMAs, 19 CELU' epilogue, 20 G_PRE copy-out + scale MMA,
 * 21 TMEM -> g_agg tile, 22 destination pass, 23 source pass, 24 G_XPE copy-out, 25 tile output; 17 is split further: 26 wait
 * for the previous step's g_x MMA + first loads, 27 the gate rounds (17 = the barrier behind them).  Results are unaffected.
 * NULL switches it off. */
int glam_message_stack_phase_clock(unsigned long long* cycles);
int glam_message_stack_fwd(const float* x0, const float* h0, const float* x_raw, int raw_dim, const float* w_pre,
                           const float* b_pre, int pre_act, float pre_act_param, const float* w_ext, int64_t ldw, const float* w_edge,
                           const float* att_edge, const float* w_scale, const float* bias, const float* w_ih,
                           const float* w_hh, const float* b_ih, const float* b_hh, const int32_t* tiles,
                           const int32_t* tile_meta, const int32_t* dst_rowptr, const int32_t* dst_src,
                           const uint8_t* etype, int64_t num_nodes, int64_t num_edges, int channels, int heads,
                           int edge_dim, int steps, float negative_slope, int act, float act_param, int res,
                           int conv_only, int keep_all, float* x_out, float* h_out, float* save_x, float* save_h,
                           float* save_xpe, float* save_agg, float* save_alpha, float* save_m, float* save_rzn,
                           float* save_gh, float* save_gt, float* save_mh, const int64_t* pn_batch, float pn_eps, void* stream);

/* glam_message_stack_bwd — backward of glam_message_stack_fwd's training mode (h0 == NULL, no conv_only) in ONE launch
 *   (csrc/mp_fused_bwd.cu), on the same tile table: gate backward, the four input-gradient projections, the edge backward
 *   by destination and by source, with the carried gradients g_x / g_h resident on the SM between steps (the reverse of the
 *   loop of src_1gp/model.py:53-54 over src_1gp/layer.py:252-267).  Reads what the forward saved (the gate-side tensors either
 *   row-major — save_h, save_x, save_m, save_rzn, save_gh — or, when save_gt != NULL, from the tile-blocked save); h_g_ext is a HOST array of
 *   `steps` device pointers (NULL entries allowed): the gradient arriving at every step's output x_{s+1} [N][C] from outside;
 *   g_h_final [N][C] (may be NULL) the gradient of the final GRU state.  Writes what the weight-gradient contractions read
 *   — g_gi, g_gh [steps][N][3C] (or, when g4 != NULL, g4 [steps][N][4C] = g_r | g_z | g_n | g_n r in their place), g_pre
 *   [steps][N][C], g_xpe [steps][N][ld] — plus g_x0 [N][C] (gradient of x0, both as the
 *   first block input and the first GRU state; when the forward was given its own h0 tensor pass g_h0 [N][C] != NULL and the two
 *   gradients come out apart) and the parameter gradients of the edge phase summed over steps: g_w_edge
 *   [edge_dim][HC], g_att_edge [edge_dim][H] (= d/d att_edge of glam_triplet_prep_fwd).  needs the source-side index of
 *   glam_build_csr.  workspace >= glam_message_stack_bwd_workspace_bytes(), 16-byte aligned.  If meta[1] != 0, g_x0 and the
 *   parameter gradients are filled with NaN.  Supported: tf32 math mode, heads == 3, channels in {32,36}, edge_dim <= 4,
 *   steps <= 8.
 *   pre_act != 0 (ABI v11): the forward applied the model's input LinearBlock in the kernel (x_raw / w_pre, activation pre_act
 *   with parameter pre_act_param; save_x[0] holds its output x0); g_x0 then leaves as the gradient of that block's
 *   PRE-activation rows, g_x0 * act'(x0), ready for the block's weight / bias contraction. */
int glam_message_stack_bwd_supported(int channels, int heads, int edge_dim, int steps);
size_t glam_message_stack_bwd_workspace_bytes(int channels, int heads, int edge_dim);
int glam_message_stack_bwd(const float* save_x, const float* save_h, const float* save_xpe, const float* save_alpha,
                           const float* save_m, const float* save_rzn, const float* save_gh, const float* save_gt,
                           const float* const* h_g_ext,
                           const float* g_h_final, const float* w_ext, int64_t ldw, const float* w_edge, const float* att_edge,
                           const float* w_scale, const float* w_ih, const float* w_hh, const int32_t* tiles,
                           const int32_t* tile_meta, const int32_t* dst_rowptr, const int32_t* dst_src, const uint8_t* etype,
                           const int32_t* src_rowptr, const int32_t* src_pos, const int32_t* src_dst, int64_t num_nodes,
                           int64_t num_edges, int channels, int heads, int edge_dim, int steps, float negative_slope, int act,
                           float act_param, int res, int pre_act, float pre_act_param, float* g_gi, float* g_gh, float* g4,
                           float* g_pre, float* g_xpe,
                           float* g_x0, float* g_h0, float* g_w_edge, float* g_att_edge, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (9) Packed graph store -> device index (csrc/packed.cu).  The reference ships every batch as fp32
 * features + int64 edge_index + fp32 one-hot edge_attr + int64 batch (src_1gp/dataset.py:60-97 + Batch.from_data_list,
 * ~2.6 KB per 25-atom molecule) and the index is rebuilt from it per batch.  The packed store (glam_b200/packed.py) keeps the
 * dst-sorted in-edge lists with graph-local uint8 source ids, uint8 bond types / degrees / atom features (~360 B per
 * molecule); unpacking is two scans + a warp per graph and yields graph_ptr [B+1], dst_rowptr [N+1], dst_src [E]
 * (bit-identical to glam_graph_ptr / glam_build_csr on the raw batch) and x [N,node_dim] fp32.  edge_ptr [B+1]: scratch
 * (first in-edge of every graph).  The bond types feed glam_message_stack_fwd as they are.
 * --------------------------------------------------------------------------------------------- */
int glam_unpack_graphs(const uint8_t* n_g, const uint16_t* e_g, const uint8_t* deg, const uint8_t* nbr, const uint8_t* xq,
                       int64_t num_graphs, int64_t num_nodes, int64_t num_edges, int node_dim, int32_t* graph_ptr,
                       int32_t* edge_ptr, int32_t* dst_rowptr, int32_t* dst_src, float* x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (10) PairNorm per graph (PyG PairNorm(scale=1, scale_individually=False, eps) @1.7.2 — the reference's default graph_norm,
 * src_1gp/run.py:28, src_1gp/layer.py:179-185,255): y = (x - mean_g(x)) / sqrt(eps + mean_g(sum_c (x - mean_g(x))^2)).
 * One warp per graph, fixed-order reductions (the torch formulation scatters with floating-point atomics).  Backward
 * recomputes the statistics from x; accumulate = 1 adds into g_x.
 * --------------------------------------------------------------------------------------------- */
int glam_pair_norm_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels, float eps,
                       float* y, int64_t ldy, void* stream);
int glam_pair_norm_bwd(const float* x, int64_t ldx, const float* g_y, int64_t ldg, const int32_t* graph_ptr, int64_t num_graphs,
                       int channels, float eps, float* g_x, int64_t ldgx, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLAM_B200_H */
