/* rc_b200.h — C ABI of the B200 (sm_100a) kernels behind raincast-gnn's training hot path.
 *
 * The reference (SohirMaskey/raincast-gnn) is pure Python on top of PyTorch + torch-geometric and
 * has no FFI of its own; each entry point below names the reference code (file:line under the
 * reference tree) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - Every function returns 0 on success, an RC_ERR_* code otherwise; rc_last_error() gives text.
 *   - The caller (PyTorch) owns every buffer: inputs, outputs, saved activations and workspaces.
 *     The library never allocates, frees or synchronises, so every call is CUDA-graph capturable.
 *   - Pointers are DEVICE pointers unless the function name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - Matrices are row-major float32 with an explicit leading dimension; index arrays are int32
 *     on the device (the reference's int64 edge_index is narrowed once, in rc_csr_build*).
 */
#ifndef RC_B200_H
#define RC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_OK 0
#define RC_ERR_ARG 1       /* bad size / null pointer / unsupported shape */
#define RC_ERR_CUDA 2      /* a CUDA runtime call failed (launch error etc.) */
#define RC_ERR_WORKSPACE 3 /* workspace too small */
#define RC_ERR_GRAPH 4     /* edge_index holds a node id outside [0, M) */

int rc_version(void);
const char* rc_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t rc_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Station graph: construction and layout
 * ---------------------------------------------------------------------------------------------- */

/* Radius graph from a dense distance matrix, utils/data.py:261-284 (build_edge_index_and_attr):
 * non-self edges with D[i,j] <= max_dist in row-major order (src=i ascending, then dst=j), attr =
 * (d / max d)^-1 in float32, then N self loops with attr 1.0.  Two calls: count, then fill. */
int rc_radius_graph_count_host(const float* dist, int n, float max_dist, int64_t* n_edges);
int rc_radius_graph_fill_host(const float* dist, int n, float max_dist, int64_t n_edges,
                              int64_t* edge_index /* [2, n_edges] */, float* edge_attr /* [n_edges] */);

/* Same edge order from 2-D coordinates with a cell grid (no N x N matrix), for graphs the dense
 * function cannot hold (BASELINE.json config 4: 100k nodes).  Distances are float64 Euclidean
 * rounded to float32, as the synthetic generator of SURVEY.md 8d defines them. */
int rc_radius_graph_coords_count_host(const double* xy, int n, double max_dist, int64_t* n_edges);
int rc_radius_graph_coords_fill_host(const double* xy, int n, double max_dist, int64_t n_edges,
                                     int64_t* edge_index, float* edge_attr);

/* dst-sorted CSR + src-sorted transpose + reverse-edge map of a (batched) edge list; replaces the
 * per-step PyG collate + scatter indices (train.py:155-156 -> Batch.from_data_list; SURVEY.md
 * Appendix B).  The sort is STABLE, so every row keeps the reference's edge order and
 *   stack(col, row_of_slot)[:, inverse(perm)] == edge_index   bit for bit.
 * Outputs (int32 / float32): rowptr[M+1], col[E], attr[E], perm[E]; t_rowptr[M+1], t_dst[E],
 * t_attr[E], t_perm[E], t_slot[E]; rev[E] (slot of the reverse edge, -1 if absent). */
typedef struct rc_csr {
  int32_t* rowptr; int32_t* col; float* attr; int32_t* perm;
  int32_t* t_rowptr; int32_t* t_dst; float* t_attr; int32_t* t_perm; int32_t* t_slot;
  int32_t* rev;
} rc_csr;

int rc_csr_build_host(const int64_t* edge_index, const float* edge_attr, int64_t n_edges, int num_nodes,
                      const rc_csr* out);
size_t rc_csr_build_workspace(int64_t n_edges, int num_nodes);
int rc_csr_build(const int64_t* edge_index, const float* edge_attr, int64_t n_edges, int num_nodes,
                 const rc_csr* out, void* workspace, size_t workspace_bytes, int32_t* err_flag /* device, 1 int */,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * GINE aggregation (PyG GINEConv message + 'add' aggregation + self term; call site
 * models/gnn.py:27-29,39-44)
 * ---------------------------------------------------------------------------------------------- */

/* h[i,:] = sum_{slot s in CSR row i} relu(x[col[s],:] + attr[s]*w_edge + b_edge) + (1+eps)*x[i,:]
 * Sum order inside a row is slot order (= reference edge order): deterministic, no atomics. */
int rc_gine_aggr_fwd(const float* x, const int32_t* rowptr, const int32_t* col, const float* attr,
                     const float* w_edge /* [H] = lin.weight[:,0] */, const float* b_edge /* [H] */,
                     const float* eps /* [1] */, float* h, int num_nodes, int hidden, void* stream);

/* Transpose gather (atomic-free):
 *   dx[j,:] = (1+eps)*g[j,:] + sum_{q in transpose row j} g[t_dst[q],:] * 1[x[j,:] + t_attr[q]*w + b > 0]
 *             (+ addend[j,:] when addend != NULL: the residual branch of models/gnn.py:44)
 * and per-block partial sums for d w_edge, d b_edge, d eps: partials[nblocks][3][H]. */
int rc_gine_aggr_bwd_nblocks(int num_nodes, int hidden);
int rc_gine_aggr_bwd(const float* g, const float* x, const int32_t* t_rowptr, const int32_t* t_dst,
                     const float* t_attr, const float* w_edge, const float* b_edge, const float* eps,
                     const float* addend, float* dx, float* partials, int num_nodes, int hidden, void* stream);
/* d_w[H], d_b[H], d_eps[1] from the partials (fixed summation order). */
int rc_gine_aggr_bwd_finalize(const float* partials, int nblocks, int hidden, float* d_w, float* d_b,
                              float* d_eps, void* stream);

/* Station tiles: the large-graph layout of the same aggregation (north_star item 1: the batching seam,
 * utils/dataset.py / train.py:155-156, emits the graph layout the kernels want).  Rows of a gather matrix
 * (rowptr/col/attr: the CSR above for forward, t_rowptr/t_dst/t_attr for backward) are clustered by a
 * breadth-first search over the graph so that the distinct rows a cluster gathers (its own rows + a halo)
 * number at most `max_src` and, with the cluster's records (at most `max_block_bytes`), fit in one CTA's
 * shared memory; a batched reference graph becomes one tile per 122-station graph with no halo.  Inside a
 * tile, rows are grouped three by three (a row with the two neighbours sharing most sources with it) and a
 * group's edges are stored per distinct source, so that a warp reads a gathered row from shared memory once
 * for all rows of the group that use it.
 *   tile_stage_ptr[T+1] -> stage_id[]: the rows a tile stages, own rows first (group by group), halo after
 *   tile_blk_ptr[T+1]   -> blocks[], in 16-byte units; per tile, 16-byte records:
 *       {rows owned, rows staged, groups, edges}
 *       per group {node id of row 0, 1, 2 (-1: none), byte offset of the group's first entry in the block}
 *                 {n1 | n2 << 16, n3 | n4 << 16, n5 | n6 << 16, n7}  entries per class (class = bit mask of
 *                                                                    the group's rows that use the source)
 *                 {degree of row 0, 1, 2 (float32), byte offset of row 0 among the staged rows}
 *       entries {staged-row index * row_bytes, attr for row 0, 1, 2 (float32 bits)}, class 1 first, 7 last;
 *       groups with most edges first.  A repeated (source, row) pair gets one entry per occurrence.
 * The builder takes worst-case output sizes (tile_*_ptr: M+1 ints, stage_id: M+E ints, blocks: 64*M + 16*E
 * bytes) and reports the counts (n_entries: row reads from shared memory; E / n_entries = reuse factor).
 * Fails with RC_ERR_ARG when a single row cannot fit a tile.  rc_gine_tiles_verify_host walks the tiles
 * the way the kernels do and compares them with the CSR (RC_ERR_GRAPH + message on any difference). */
typedef struct rc_gine_tiles {
  int32_t n_tiles; int32_t max_staged;       /* most rows any tile stages                      */
  int32_t max_block_bytes; int32_t row_bytes; /* largest tile block; 4 * hidden                 */
  const int32_t* tile_stage_ptr; const int32_t* tile_blk_ptr; const int32_t* stage_id; const int32_t* blocks;
  int32_t* sched;   /* device int32[8], zero when handed over: the forward kernel's tile counters (it leaves them
                       zero; one launch at a time per rc_gine_tiles)                                        */
} rc_gine_tiles;
/* max_src / max_block_bytes that fit one CTA at `hidden` columns (128 | hidden <= 512). */
int rc_gine_tiles_limits(int hidden, int* max_src, int* max_block_bytes);
int rc_gine_tiles_build_host(const int32_t* rowptr, const int32_t* col, const float* attr, int num_nodes,
                             int64_t n_edges, int max_src, int max_block_bytes, int row_bytes,
                             int32_t* tile_stage_ptr, int32_t* tile_blk_ptr, int32_t* stage_id, int32_t* blocks,
                             int32_t* n_tiles, int64_t* n_staged, int64_t* n_block_units, int32_t* max_staged,
                             int32_t* max_block_bytes_out, int64_t* n_entries);
int rc_gine_tiles_verify_host(const int32_t* rowptr, const int32_t* col, const float* attr, int num_nodes,
                              int64_t n_edges, int n_tiles, int max_staged, int max_block_bytes, int row_bytes,
                              const int32_t* tile_stage_ptr, const int32_t* tile_blk_ptr, const int32_t* stage_id,
                              const int32_t* blocks);
/* Same results as rc_gine_aggr_fwd / rc_gine_aggr_bwd to rounding (the sums run class by class instead of in
 * CSR slot order, and relu(x_j + a w + b) is evaluated as max(x_j + a w, -b) + b with the bias added once per
 * row); deterministic for given tiles.  Every gathered row travels HBM/L2 -> shared memory once per tile
 * (cp.async) and is read from shared memory once per group of rows that needs it.
 * 128 | hidden <= 512.  partials: [rc_gine_aggr_bwd_tiled_nblocks][3][H], finalised by
 * rc_gine_aggr_bwd_finalize. */
int rc_gine_aggr_fwd_tiled(const float* x, const rc_gine_tiles* tiles, const float* w_edge, const float* b_edge,
                           const float* eps, float* h, int num_nodes, int hidden, void* stream);
int rc_gine_aggr_bwd_tiled_nblocks(const rc_gine_tiles* tiles, int hidden);
int rc_gine_aggr_bwd_tiled(const float* g, const float* x, const rc_gine_tiles* t_tiles, const float* w_edge,
                           const float* b_edge, const float* eps, const float* addend, float* dx, float* partials,
                           int num_nodes, int hidden, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers: one fp32 SIMT GEMM family with fused prologues / epilogues (torch.nn.Linear,
 * BatchNorm1d and ReLU of models/gnn.py:21-26,51-62,113,123 and their backward)
 * ---------------------------------------------------------------------------------------------- */

/* operand layouts: how the stored matrix relates to D[i,j] = sum_r A(i,r) * B(r,j) */
#define RC_A_ROW 0 /* A stored [i][r]  (activations in forward / backward-data)            */
#define RC_A_RED 1 /* A stored [r][i]  (grad_out in the weight-gradient GEMM, r = sample)   */
#define RC_B_COL 0 /* B stored [j][r]  (nn.Linear weight [out,in] in forward)                */
#define RC_B_RED 1 /* B stored [r][j]  (weight in backward-data; activations in weight-grad) */

/* operand prologues, applied to stored element (row, col) while a tile is loaded */
#define RC_OP_NONE 0
#define RC_OP_BN_RELU 1 /* v = relu((v - p0[col]) * p1[col] * p2[col] + p3[col])   (mean, rstd, gamma, beta) */
#define RC_OP_BITMASK 2 /* v = bit(bits[row*ld_bits + col/32], col%32) ? v : 0                                */
#define RC_OP_AFFINE2 3 /* v = p0[col]*v + p1[col]*(aux[row*ld_aux + col] - p3[col]) + p2[col]  (BN backward) */
/* The GINE aggregation as the prologue of the layer's first Linear (north_star item 2: message, aggregation and the
 * (1+eps) x_i + aggr update fused into the node MLP; models/gnn.py:21-29 / PyG GINEConv): A operand only, A stored [i][r],
 *   A(i, r) = sum_{slot s of CSR row i} relu(x[idx1[s], r] + aux[s]*p0[r] + p1[r]) + (1 + p2[0]) * x[i, r]
 * with ptr = x (the layer input, ld % 4 == 0, 16-byte aligned), idx0 = rowptr, idx1 = col, aux = attr, p0 / p1 = the edge
 * Linear's weight / bias, p2 = eps.  Same arithmetic and slot order as rc_gine_aggr_fwd (bit-identical).  rc_gemm.a_out
 * (optional) receives the aggregated rows - the backward needs them for the weight gradient. */
#define RC_OP_GINE_AGGR 4

/* epilogues on D element (i, j), after `+ bias_scale*bias[j]` */
#define RC_EPI_NONE 0
#define RC_EPI_RELU 1       /* D = relu(v)                                                                    */
#define RC_EPI_RELU_RES 2   /* D = res[i,j] + relu(v); bits (optional) record v > 0        (models/gnn.py:44) */
#define RC_EPI_BN_STATS 3   /* D = v; per row-tile column mean / M2 -> stats[tile][2][N]  (BatchNorm forward) */
#define RC_EPI_MASK_POS 4   /* D = aux[i,j] > 0 ? v : 0                                     (ReLU backward)    */
#define RC_EPI_ADD_RES 6    /* D = v + res[i,j]   (dim_red: the half of the Linear that only needs the batch's x runs
                               early, off the critical path, and joins here; models/gnn.py:134-135)         */
#define RC_EPI_BN_RELU_BWD 5 /* z = p2*(aux-p0)*p1 + p3; D = z > 0 ? v : 0; column partial sums of D and
                               D*(aux-p0)*p1 -> stats[tile][2][N]               (ReLU + BatchNorm backward)   */

typedef struct rc_operand {
  const float* ptr; int ld;
  int op;                     /* RC_OP_* */
  const float* p0; const float* p1; const float* p2; const float* p3;   /* per stored column */
  const float* aux; int ld_aux;
  const uint32_t* bits; int ld_bits;
  const int32_t* idx0; const int32_t* idx1;   /* RC_OP_GINE_AGGR: CSR rowptr / col */
} rc_operand;

typedef struct rc_gemm {
  int m, n, k;                /* D is m x n, reduction length k (first segment) */
  int a_layout, b_layout;
  rc_operand a, b;
  /* optional second segment accumulated into the same D (dim_red on cat([x, emb]) without the cat,
   * models/gnn.py:134-135); same layouts, reduction length k2, no prologues */
  const float* a2; int lda2; const float* b2; int ldb2; int k2;
  float* d; int ldd;
  const float* bias; float bias_scale;
  int epi;                    /* RC_EPI_* */
  const float* res; int ld_res;           /* RC_EPI_RELU_RES, RC_EPI_ADD_RES */
  uint32_t* bits_out; int ld_bits_out;    /* RC_EPI_RELU / RC_EPI_RELU_RES (optional) */
  const float* e_aux; int ld_e_aux;       /* RC_EPI_MASK_POS, RC_EPI_BN_RELU_BWD */
  const float* e_p0; const float* e_p1; const float* e_p2; const float* e_p3;
  float* stats;                           /* RC_EPI_BN_STATS / RC_EPI_BN_RELU_BWD: [row_tiles][2][n] */
  /* split of the reduction over gridDim.z (weight-gradient GEMMs): slice z writes its partial tile to
   * d + z*split_stride, and (colsum_a != NULL, A stored [r][i]) the column sums of A over its slice to
   * colsum_a + z*m  (the bias gradient).  splits <= 1: no split. */
  int splits; long long split_stride;
  float* colsum_a;
  int rows_per_warp;          /* 0 = choose; else 1, 2, 4 or 8 (row tile = 8 * rows_per_warp) */
  /* tensor-core path (tcgen05, 3xTF32: fp32-accurate): taken for activation GEMMs (A stored [i][r]) with m >= 16384
   * when tc_ws holds rc_gemm_tc_workspace(g) bytes (the pre-split weight blocks live there for the duration of the
   * call), and for weight-gradient GEMMs (A stored [r][i], B stored [r][j]) with k >= 16384 samples (no workspace).
   * RC_GEMM_TC=0 in the environment keeps everything on the SIMT kernels. */
  void* tc_ws; size_t tc_ws_bytes;
  float* a_out; int ld_a_out;   /* the A operand after its prologue, written once (nullable): RC_OP_GINE_AGGR on the SIMT
                                   path, every prologue on the tensor-core activation path (the layer's weight-gradient
                                   GEMM then takes it as a plain operand).  Which path runs is rc_gemm_tc_workspace(g) > 0 -
                                   NOT the row tile: the SIMT kernels also pick 64-row tiles once those fill the SMs */
  int b_static;                 /* 1: the B operand is a parameter and res / e_aux are saved activations - the kernel
                                   launched just before this one on the stream writes none of them.  The kernel then fetches its first B tile BEFORE waiting for that
                                   kernel (programmatic dependent launch), overlapping the fetch with its tail */
} rc_gemm;

int rc_gemm_row_tile(const rc_gemm* g);  /* the row tile the launch would use (for stats sizing)   */
size_t rc_gemm_tc_workspace(const rc_gemm* g);   /* 0: the tensor-core path does not apply to this activation GEMM */
/* reduction splits the tensor-core weight-gradient kernel wants for D (m x n) over k samples; 0: not applicable */
int rc_gemm_tc_wgrad_splits(int m, int n, int k);
int rc_gemm_run(const rc_gemm* g, void* stream);

/* BatchNorm1d training-mode statistics from the RC_EPI_BN_STATS tiles (per-tile count / mean / M2 combined in
 * float64): mean[N], rstd[N] = 1/sqrt(var_biased + eps); running_mean/var updated with `momentum`
 * and the UNBIASED variance, num_batches_tracked += 1 (torch.nn.BatchNorm1d, models/gnn.py:23). */
int rc_bn_stats_finalize(const float* stats, int row_tiles, int row_tile, int m, int n, float eps,
                         float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                         int64_t* num_batches_tracked, void* stream);
/* Eval mode: mean = running_mean, rstd = 1/sqrt(running_var + eps). */
int rc_bn_eval_prepare(const float* running_mean, const float* running_var, int n, float eps, float* mean,
                       float* rstd, void* stream);
/* BatchNorm backward coefficients from the RC_EPI_BN_RELU_BWD tiles: d_gamma[N], d_beta[N] and the
 * per-column coefficients such that d t = c0*dz + c1*(t - mean) + c2 (the RC_OP_AFFINE2 prologue with
 * p3 = mean). */
int rc_bn_bwd_finalize(const float* stats, int row_tiles, int m, int n, int batch_stats, const float* gamma, const float* mean,
                       const float* rstd, float* d_gamma, float* d_beta, float* c0, float* c1, float* c2,
                       void* stream);

/* out[j] = scale * sum_{p < parts} src[p*stride + j]  (fixed order, float64 accumulator) for a
 * table of segments; one launch finishes every split weight/bias gradient of a backward pass. */
typedef struct rc_reduce_seg {
  const float* src; float* dst; long long stride; int parts; int n; float scale; int accumulate;
  int row_len; int dst_ld;   /* row_len > 0: element j lands at dst[(j / row_len) * dst_ld + j % row_len] */
} rc_reduce_seg;
/* `segs` is a HOST array; up to RC_REDUCE_MAX_SEGS descriptors travel as kernel arguments per launch
 * (no device-side table, so the call is graph-capturable without a staging copy). */
#define RC_REDUCE_MAX_SEGS 16
int rc_reduce_segments(const rc_reduce_seg* segs, int n_segs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DeepSets member MLP + pooling (models/gnn.py:48-68)
 * ---------------------------------------------------------------------------------------------- */

/* pooled[i,:] = sum_{e < members} relu(ens[i,e,:] @ w1^T + b1)      ens [M, members, F], w1 [H, F]
 * (the second phi Linear is applied AFTER the sum by rc_gemm_run with bias_scale = members: the sum
 *  over members commutes with it exactly in real arithmetic). */
int rc_deepsets_pool_fwd(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                         int members, int feats, int hidden, void* stream);
/* Same contraction with bf16 operands and fp32 accumulation on the tensor cores (tcgen05 + TMEM; BASELINE.json
 * config 5).  rc_deepsets_pool_fwd itself switches to the tensor cores with an fp32-accurate 3xTF32 split when
 * num_nodes*members >= 8192 and 128 | hidden (override: RC_DEEPSETS_TC=0/1). */
int rc_deepsets_pool_fwd_bf16(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                              int members, int feats, int hidden, void* stream);
/* d w1 / d b1 partials from d pooled (ReLU mask recomputed, nothing saved in forward):
 * partials[nblocks][H*F + H]. */
int rc_deepsets_pool_bwd_nblocks(int num_nodes, int members, int feats, int hidden);
/* With num_nodes*members >= 8192 and 11 or 51 members (the reference's ensembles) the backward also runs on the tensor
 * cores (tcgen05, 3xTF32; RC_DEEPSETS_TC=0/1 overrides).  mask_bits_out (nullable, tests): the ReLU mask the backward
 * used, bit (c % 32) of word [row * ceil(H/32) + c / 32] for member row `row`, channel c. */
int rc_deepsets_pool_bwd(const float* ens, const float* w1, const float* b1, const float* d_pooled,
                         float* partials, int num_nodes, int members, int feats, int hidden,
                         int bf16_operands /* 1 after rc_deepsets_pool_fwd_bf16: mask and inputs as the tensor cores saw them */,
                         uint32_t* mask_bits_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Output links + closed-form CRPS (models/model_utils.py:70-113, models/loss.py:6-68,71-272,335-369)
 * ---------------------------------------------------------------------------------------------- */
#define RC_LOSS_NORMAL 0        /* NormalCRPS,      2 columns: mu, sigma                   */
#define RC_LOSS_MIXED_NORMAL 1  /* MixedNormalCRPS, 3 columns: mu, sigma, p                */
#define RC_LOSS_MIXED 2         /* MixedLoss, fixed u, 4 columns: mu, sigma, p, sigma_u    */
#define RC_LOSS_MIXED_U 3       /* MixedLoss, learned u, 5 columns: ..., u                 */

/* PostProcess forward / backward (softplus+1e-6, sigmoid, 2.12*sigmoid). */
int rc_postprocess_fwd(const float* raw, float* post, int num_nodes, int kind, void* stream);
int rc_postprocess_bwd(const float* raw, const float* d_post, float* d_raw, int num_nodes, int kind, void* stream);

/* Per-node CRPS and its gradient in one pass.  `pred` holds post-processed parameters
 * (raw_input = 0, the public `crps(prediction, y)` signature) or raw head outputs (raw_input = 1: the
 * links are applied inside and d_pred is the gradient w.r.t. the raw outputs).  NaN targets are
 * skipped (models/loss.py:216,237-241).  The kernel writes per-block partial sums; finalize produces
 * loss_out[0] = mean over valid nodes (float64, like the reference's promoted result) and
 * n_valid[0]; d_pred is already divided by the number of valid nodes.
 * workspace: rc_crps_workspace(num_nodes) bytes.  xi must not be 1 or 2 (models/loss.py:121-124). */
size_t rc_crps_workspace(int num_nodes);
int rc_crps_fwd_bwd(const float* pred, const float* y, float* d_pred /* nullable */, double* loss_out,
                    int32_t* n_valid, int num_nodes, int kind, int raw_input, float u_fixed, float xi, float t,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Head Linear (models/gnn.py:123,139: `aggr`, H -> C) + links + CRPS + their backward in ONE launch, for batches of up to
 * 16 384 nodes and H = 128 or 256 (rc_head_crps_blocks returns 0 when it does not apply: use rc_gemm_run + rc_crps_fwd_bwd).
 *   h [M][H] last hidden state, w [C][H], b [C] (C = kind + 2), y [M] targets (NaN = missing)
 *   d_h [M][H]           gradient of the mean CRPS w.r.t. h
 *   partials [blocks][C*H + C]   per-CTA partial d w (row-major [C][H]) followed by d b: sum over blocks (rc_reduce_segments)
 *   loss_partials [blocks] float64 scratch; loss_out float64[1] = mean over valid nodes; n_valid int32[1]
 * y and w are read BEFORE the kernel waits for the launch in front of it on the stream (programmatic dependent launch):
 * that launch must not write them. */
int rc_head_crps_blocks(int num_nodes, int hidden);
int rc_head_crps_fwd_bwd(const float* h, const float* w, const float* b, const float* y, float* d_h, float* partials,
                         double* loss_partials, double* loss_out, int32_t* n_valid, int num_nodes, int hidden, int kind,
                         float u_fixed, float xi, float t, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser (train.py:66-69,185: torch.optim.AdamW defaults) on flat parameter / gradient buffers
 * ---------------------------------------------------------------------------------------------- */
/* p -= lr*wd*p; m,v update; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps); grad is first
 * multiplied by grad_scale (1/world_size after an all-reduce SUM).  `step` is a device int64 that is
 * incremented by the kernel, so the call is replayable inside a CUDA graph. */
int rc_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t* step,
                  long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, void* stream);

/* GPU-resident split (replaces the per-step PyG collate + 7 H2D copies of train.py:61-62,155-156 when
 * the whole split fits in HBM: the reference's train split is ~0.7 GB).  x_all [n_dates][x_len], ens_all
 * [n_dates][ens_len], y_all [n_dates][y_len] live on the device; `dates` (device int64 [n_batch]) selects the batch,
 * which lands in the step's inputs x [n_batch][x_len], ens, y in the order given (= PyG's concatenation order for a
 * static graph).  An index outside [0, n_dates) sets *bad = 1 and copies nothing for that slot. */
int rc_gather_dates(const float* x_all, const float* ens_all, const float* y_all, const int64_t* dates, int n_batch,
                    int n_dates, long long x_len, long long ens_len, long long y_len, float* x, float* ens, float* y,
                    int32_t* bad, void* stream);
/* The same gather for a CAPTURED step (train.py:55-74 as one CUDA graph per batch, nothing written by the host between
 * replays): `order` (device int64 [n_batches][n_batch]) holds the epoch's shuffled order, the batch taken is number
 * *step_count - epoch_base[0], where step_count is the optimiser's device-side step counter (rc_adamw_step /
 * rc_p2p_step advance it), epoch_base[0] its value when the epoch began and epoch_base[1] the number of batches of this
 * epoch (device int64 [2]).  A batch number outside [0, min(n_batches, epoch_base[1])) sets *bad = 2 and copies nothing. */
int rc_gather_dates_step(const float* x_all, const float* ens_all, const float* y_all, const int64_t* order, int n_batches,
                         const int64_t* step_count, const int64_t* epoch_base, int n_batch, int n_dates, long long x_len,
                         long long ens_len, long long y_len, float* x, float* ens, float* y, int32_t* bad, void* stream);

/* Data-parallel step over NVLink peer memory (the reference, train.py:55-74,185, is single device):
 * every rank keeps its flat gradient where the other ranks of the box can read it, and
 *     rc_p2p_barrier(slot 0)  ->  rc_p2p_adamw_step  ->  rc_p2p_barrier(slot 1)
 * replaces ncclAllReduce + rc_adamw_step.  rc_p2p_adamw_step sums the peers' gradients in rank order (bit-identical
 * on every rank), scales by 1/world and applies the same AdamW update as rc_adamw_step (device-side step counter).
 *   flags       device array [world] of pointers: rank r's flag block (int32 [2][16], zero before first use) as
 *               mapped in this process;  epochs: this rank's int32[2] (zero before first use);  timed_out:
 *               int32[1], set if a peer did not arrive within ~10 s (the kernel then gives up instead of hanging)
 *   peer_grads  device array [world] of pointers to the ranks' flat gradients (n floats, n % 4 == 0, 16-byte aligned)
 * No allocation, no host synchronisation: all three calls are CUDA-graph capturable.  world <= 16. */
int rc_p2p_barrier(int32_t* const* flags, int32_t* epochs, int rank, int world, int slot, int32_t* timed_out, void* stream);
int rc_p2p_adamw_step(float* param, const float* const* peer_grads, int world, float* exp_avg, float* exp_avg_sq,
                      int64_t* step, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                      void* stream);
/* The same exchange as ONE kernel after backward: rc_p2p_step waits (inside the kernel) until every rank has published
 * its gradients of exchange *epoch + 1, sums them in rank order, applies AdamW, advances *step and *epoch, publishes
 * "done reading" as soon as this rank holds the peers' gradients in registers, and ends only when every peer has done
 * the same - so the caller may overwrite its gradients as soon as the kernel has completed (no second call).
 * flags: as rc_p2p_barrier (slot 0 = published, slot 1 = done reading); epoch: this rank's int32[1], zero before first
 * use, never reset (it must advance in lock step on all ranks: every rank calls rc_p2p_step the same number of times).
 * The kernel fetches this rank's param / exp_avg / exp_avg_sq BEFORE it waits for the kernel launched before it on the
 * stream (programmatic dependent launch): that kernel must not write them (in a training step it writes gradients).
 * rc_p2p_wait_done is the stand-alone wait for slot 1 (callers that build the exchange from the barriers).
 * rc_p2p_flag_scope(1): device-scope fences + relaxed system-scope flag accesses instead of the library default
 * st.release.sys / fence.acq_rel.sys (7-8 us per step cheaper on an NVSwitch box; rc_p2p.cu states why it is sufficient
 * for flags that guard data written by earlier kernels). */
int rc_p2p_step(float* param, const float* const* peer_grads, int32_t* const* flags, int32_t* epoch, int rank, int world,
                float* exp_avg, float* exp_avg_sq, int64_t* step, long long n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int32_t* timed_out, void* stream);
int rc_p2p_wait_done(int32_t* const* flags, const int32_t* epoch, int rank, int world, int32_t* timed_out, void* stream);
int rc_p2p_flag_scope(int device_scope_fences);

/* ------------------------------------------------------------------------------------------------
 * Test instrumentation: the ReLU decisions of the backward kernels as bit masks (bit c % 32 of word c / 32 per row).
 * The parity tests force them into the float64 oracle so that gradients can be held to 1e-5 without an allowance for
 * units within rounding of their threshold (tests/test_gpu_masked_parity.py).  Not used by the product path.
 * ---------------------------------------------------------------------------------------------- */
/* message ReLU of the GINE aggregation backward, per TRANSPOSE slot q (t_rowptr / t_attr of rc_csr): bits_out[E][ceil(H/32)];
 * tiled = 1 evaluates the expression of rc_gine_aggr_bwd_tiled, 0 that of rc_gine_aggr_bwd */
int rc_debug_gine_msg_mask(const float* x, const int32_t* t_rowptr, const float* t_attr, const float* w_edge,
                           const float* b_edge, int num_nodes, int hidden, int tiled, uint32_t* bits_out, void* stream);
/* ReLU behind BatchNorm as the RC_EPI_BN_RELU_BWD epilogue evaluates it: bits_out[m][ceil(n/32)] */
int rc_debug_bn_relu_mask(const float* t, int ld, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          int m, int n, uint32_t* bits_out, void* stream);
/* fp32 FMA throughput probe (bench.py step_roofline: the measured fp32 peak); *flops = operations one launch executes */
int rc_debug_fma_peak(float* scratch, int iters, double* flops, void* stream);
/* tensor-core GEMM timeline of CTA 0 (tools/trace_gemm_tc.py) */
void rc_debug_tc_trace(void* device_buf);
/* rc_p2p_step: globaltimer (ns) of CTA 0 at entry, flags published, all ranks arrived, update done, exit; rc_p2p_wait_done:
 * entry, exit: int64[8] (set before a CUDA graph capture to trace replays) */
void rc_debug_p2p_trace(void* device_buf);
/* tensor-core DeepSets pool backward: clocks of CTA 0, int64 [8 events][32 stages] (tools/trace_pool_bwd.py) */
void rc_debug_ds_trace(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* RC_B200_H */
