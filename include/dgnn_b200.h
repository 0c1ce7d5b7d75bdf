/*
 * dgnn_b200 — C ABI of the B200-native DGNN cell-classification hot path.
 *
 * Drop-in boundary: these are the entry points a host binding (ctypes here; see
 * INTEGRATION.md) calls in place of the library kernels the reference reaches through
 * PyTorch / PyG / torch_scatter.  Every function
 *   - returns 0 on success, non-zero on failure (message: dgnn_last_error()),
 *   - takes raw DEVICE pointers unless a parameter says "host",
 *   - allocates nothing: the caller passes outputs and workspaces,
 *   - enqueues on the given cudaStream_t (passed as void*) and does not synchronise.
 * All feature matrices are row-major float32 with row length a multiple of 4 floats and a
 * 16-byte aligned base.  Indices are int32 on the device (int64 only where the reference's
 * own tensors are handed in unchanged).
 *
 * Reference interfaces replaced (file:line under /root/reference):
 *   learning/surfaceNetStaticEdgeFilters.py:66-96   SAGEConv.forward/message (+ PyG propagate,
 *                                                   torch_scatter mean)        -> dgnn_layer_fwd / _bwd_*
 *   learning/surfaceNetStaticEdgeFilters.py:116-123 BatchNorm / LayerNorm      -> dgnn_norm_finalize (+ affine-on-load)
 *   learning/surfaceNetStaticEdgeFilters.py:180-187 decoder                    -> dgnn_layer_fwd(no gather) + dgnn_rowdot_*
 *   learning/surfaceNetUpdatedEdgeFilters.py:147-176 SAGEConv (updated filters) -> dgnn_layer_fwd with chained edge state
 *   learning/runModel.py:163-211                    calcLossAndOA (kl)         -> dgnn_kl_loss_fwd / _bwd
 *   learning/runModel.py:109-160                    calcRegularization         -> dgnn_edge_reg_fwd / _bwd
 *   learning/runModel.py:290,282                    Adam step                  -> dgnn_adam_step
 *   processing/data.py:434-439                      adjacency -> edge_index    -> dgnn_ell_from_adjacency / dgnn_ell_build
 *   processing/generate_mesh.py:75,94-105           labels, interface facets   -> dgnn_argmax_labels / dgnn_interface_facets
 */
#ifndef DGNN_B200_H
#define DGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGNN_B200_VERSION 100

/* ---- library ------------------------------------------------------------------------- */
int dgnn_version(void);
const char* dgnn_last_error(void);
/* 0 if `device` is a compute-capability 10.x GPU; error otherwise (no fallback exists). */
int dgnn_device_check(int device);
/* number of SMs of the current device (grid sizing), <0 on error */
int dgnn_sm_count(void);
/* Leave n SMs free in every persistent-kernel grid launched by the calling thread from now on (0 = use all):
 * lets the halo pack + NCCL kernels of a communication stream run beside a layer kernel.  Returns the
 * previous value.  dgnn_sm_count / dgnn_tc_grid / dgnn_small_grid follow the reservation. */
int dgnn_reserve_sms(int n);

/* ---- graph layout (processing/data.py:434-439 re-laid to ELL-4) ------------------------ */
/* adjacencies int32[4n,2] (row 4i+k = (i, k-th facet neighbour)) -> nbr int32[n,4] and the
 * reverse-facet slot rslot uint8[n,4] (nbr[nbr[i,k], rslot[i,k]] == i).  *err_flag (device
 * int32, caller-zeroed) is set to 1 if a row's owner column is not i, 2 if not symmetric. */
int dgnn_ell_from_adjacency(const int32_t* adj, int64_t n, int32_t* nbr, uint8_t* rslot,
                            int32_t* err_flag, void* stream);
/* generic: edge list (src[e] -> tgt[e], int64 as in edge_index) -> per-target rows of <=4
 * in-edges in ascending edge id (by_source = 0) or ascending (source id, edge id) (by_source = 1:
 * the row order of PyG's SparseTensor(row=src, col=tgt).t(), which NeighborSampler walks).
 * nbr int32[n_rows,4] (source ids, -1 pad), eid int32[n_rows,4] (edge id e, -1 pad), cnt
 * int32[n_rows] (caller-zeroed).  *err_flag = 3 if a row has more than 4 in-edges, 4 if a
 * target is outside [0, n_rows) or a source outside [0, n_other). */
int dgnn_ell_build(const int64_t* src, const int64_t* tgt, int64_t n_edges, int64_t n_rows,
                   int64_t n_other, int by_source, int32_t* nbr, int32_t* eid, int32_t* cnt,
                   int32_t* err_flag, void* stream);
/* 3*bits-bit Morton code of pos float32[n,3] quantised over [lo,hi] (host float[3]) */
int dgnn_morton_codes(const float* pos, int64_t n, const float* lo_host, const float* hi_host,
                      int bits, uint64_t* codes, void* stream);
/* out[new,k] = inv[nbr[perm[new],k]] (negative ids kept) */
int dgnn_perm_apply_ell(const int32_t* nbr, const int32_t* perm, const int32_t* inv, int64_t n,
                        int32_t* out, void* stream);
/* dst[r,:] = src[idx[r],:]  (row_floats multiple of 4; idx<0 -> zeros) */
int dgnn_gather_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats,
                     float* dst, void* stream);
/* dst[idx[r],:] += src[r,:]; idx distinct within a call (gradient rows returned by one peer) */
int dgnn_add_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats,
                  float* dst, void* stream);
/* dst[idx[r],:] = src[r,:] */
int dgnn_scatter_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats,
                      float* dst, void* stream);
/* edge attributes ea float32[4n,fe] (reference row order 4i+k) -> incoming order
 * ea_in[new,k] = ea[4*nbr[i,k]+rslot[i,k]] and own-slot order ea_own[new,k] = ea[4i+k],
 * i = perm[new] (perm may be NULL = identity); nbr/rslot in ORIGINAL numbering. Either
 * output may be NULL. */
int dgnn_edge_relayout(const float* ea, const int32_t* nbr, const uint8_t* rslot,
                       const int32_t* perm, int64_t n, int fe, float* ea_in, float* ea_own,
                       void* stream);
/* Same with the batch's edge selection fused in: edge r of the batch is row e_id[r] of `ea`
 * (data.all.edge_attr[e_id], Static:217; e_id int64[4n], NULL = identity). */
int dgnn_edge_relayout_idx(const float* ea, const int64_t* e_id, const int32_t* nbr, const uint8_t* rslot,
                           const int32_t* perm, int64_t n, int fe, float* ea_in, float* ea_own,
                           void* stream);

/* ---- neighbour-closure sampler on the device (run.py:59-74: NeighborSampler(sizes=[-1]*hops)) -----------------
 * in_src / in_eid int32[N,4]: the in-edge table of dgnn_ell_build (sources and edge ids, ascending edge id, -1 pad).
 * One hop over the node list n_id int64[n_tgt] (targets = all of it), every step a grid-stride kernel:
 *   dgnn_sampler_degree  deg[t] = in-degree of n_id[t]                       (caller: offsets = exclusive prefix sum)
 *   dgnn_sampler_expand  edges in (target, ascending edge id) order: e_id, global source, local target
 *   dgnn_sampler_set_loc loc[ids[i]] = base + i  (base < 0: reset to -1); loc int32[N] is -1 for nodes not in n_id
 *   dgnn_sampler_mark    first[s] = min edge position whose source s is new (first int32[N], INT_MAX when idle);
 *                        flag[p] = 1 iff position p is that first appearance          (caller: rank = inclusive prefix sum)
 *   dgnn_sampler_assign  new_ids[rank-1] = source, loc[source] = n_tgt + rank - 1, first reset; src_local[p] = loc[src] */
int dgnn_sampler_degree(const int64_t* n_id, int64_t n_tgt, const int32_t* in_src, int64_t* deg, void* stream);
int dgnn_sampler_expand(const int64_t* n_id, int64_t n_tgt, const int32_t* in_src, const int32_t* in_eid,
                        const int64_t* offsets, int64_t* e_id, int64_t* src_global, int64_t* tgt_local, void* stream);
int dgnn_sampler_set_loc(const int64_t* ids, int64_t n, int base, int32_t* loc, void* stream);
int dgnn_sampler_mark(const int64_t* src_global, int64_t n_edges, const int32_t* loc, int32_t* first,
                      int64_t* flag, void* stream);
int dgnn_sampler_assign(const int64_t* src_global, int64_t n_edges, const int64_t* flag, const int64_t* rank,
                        int64_t n_tgt, int32_t* loc, int32_t* first, int64_t* new_ids, int64_t* src_local,
                        void* stream);

/* ---- loader: per-graph feature standardisation (processing/data.py:467-506, sklearn StandardScaler) -----------
 * Columns col0 .. col0+c-1 (c <= 64) of x float32[n, ld].  dgnn_column_moments writes per-block partial sums
 * double[n_blocks, 2, c] of (x - center_j) and (x - center_j)^2 (center NULL = 0): a first pass gives the means, a
 * second pass centred on them the variances (the numerically stable two-pass form sklearn uses).
 * dgnn_column_affine writes out[r, col0+j] = float((x[r, col0+j] - shift[j]) * inv_scale[j]); out may alias x. */
int dgnn_column_moments(const float* x, int64_t n, int ld, int col0, int c, const double* center,
                        double* partials, int n_blocks, void* stream);
int dgnn_column_affine(const float* x, int64_t n, int ld, int col0, int c, const double* shift,
                       const double* inv_scale, int ld_out, float* out, void* stream);
/* The reference keeps the features in float64 up to the final `torch.tensor(.., dtype=torch.float)`
 * (processing/data.py:444-519).  Same two-pass moments on a float64 matrix, and the transform as sklearn
 * writes it, `X -= mean_; X /= scale_`, evaluated in float64 and rounded to float32 once:
 * out[r, col0_out+j] = float((x[r, col0+j] - mean[j]) / scale[j]). */
int dgnn_column_moments_f64(const double* x, int64_t n, int ld, int col0, int c, const double* center,
                            double* partials, int n_blocks, void* stream);
int dgnn_column_standardize_f64(const double* x, int64_t n, int ld, int col0, int c, const double* mean,
                                const double* scale, int ld_out, int col0_out, float* out, void* stream);

/* ---- one message-passing layer, forward (Static:66-96 + norm/ReLU of the producer) ------
 * For target row t < n_tgt (targets are the first n_tgt source rows):
 *   h(s)   = in_scale ? relu?(x_in[s]*in_scale + in_shift) : relu?(x_in[s])   (relu iff relu_in)
 *   phi_k  = w_e ? (w_e . ea[t,k] + b_e) : 1
 *   agg[t] = (1/max(cnt,1)) * sum_{k: nbr[t,k]>=0} h(nbr[t,k]) (*) phi_k
 *   z[t]   = [agg[t] | h(t)] . wt_cat + bias          (wt_cat float32[k_total, f_out],
 *            k_total = 2*f_in, rows 0..f_in-1 = lin_j^T, rows f_in.. = lin_i^T)
 *   out[t] = out_scale ? relu?(z*out_scale + out_shift) : z            (relu iff relu_out)
 * With nbr == NULL the gather is skipped (dense layer: z = h(t) . wt_cat + bias, k_total = f_in).
 * agg_save (nullable) receives agg (training).  stats (nullable) = double[grid, 2, f_out]
 * per-CTA partial (sum z, sum z^2) over valid rows; grid = dgnn_layer_grid().
 * ea_stride = floats per edge row in `ea` (>= fe).   * relu_in: bit 0 = ReLU on load; bit 1 (dgnn_layer_fwd only, fe == 0) = self loops: every target is its own fifth
 * in-neighbour, agg = (sum_k h(nbr_k) + h(t)) / (cnt + 1)  (run.py:70-71,215-216 add_self_loops when edge_convs == 0). */
int dgnn_layer_grid(int f_in, int f_out);
int dgnn_layer_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                   const int32_t* nbr, const float* ea, int fe,
                   const float* w_e, const float* b_e,
                   const float* wt_cat, const float* bias,
                   const float* out_scale, const float* out_shift, int relu_out,
                   int64_t n_tgt, int f_in, int f_out,
                   float* out, float* agg_save, double* stats, void* stream);

/* ---- tensor-core (tcgen05 / TMEM, 3xTF32) variants ---------------------------------------------
 * Any width that is a multiple of 4 (configs/modelnet.yaml:56 ships [128,256,512,1024]): a launch covers
 * one slice of output columns (dgnn_tc_slice(0) = 128 forward, dgnn_tc_slice(1) = 256 backward), wider
 * layers are run slice by slice inside the call; the contraction length is unbounded.
 * Same semantics as dgnn_layer_fwd / dgnn_dense_bwd; the dense operand is pre-packed by
 * dgnn_pack_b_tf32 into 128B-swizzled K-atoms split into TF32 hi / lo parts:
 *   forward : w = [W_j | W_i] (float32[f_out, 2 f_in] or [f_out, f_in] for a dense layer),
 *             n_rows = f_out, ld = row stride, seg_len = f_in, n_segs = 2 (1 for dense)
 *   backward: w = [W_j | W_i]^T (float32[k_total, f_out]), n_rows = k_total, ld = f_out,
 *             seg_len = f_out, n_segs = 1
 * packed holds dgnn_tc_packed_floats(n_rows, seg_len, n_segs) floats, rows grouped in slices of
 * `slice` = dgnn_tc_slice(0 forward / 1 backward).  Partial-sum workspaces (stats, db_partials) have
 * dgnn_tc_grid() rows.
 * dgnn_dense_fwd_tc / dgnn_dense_bwd_tc (and dgnn_layer_fwd_tc without a table) move their activations with TMA tensor
 * maps (cp.async.bulk.tensor loads of [128 x 32] atoms, [32 x 32] tensor stores): agg / x_in / dy / z / out / d_agg /
 * d_self must be 16-byte aligned, row-contiguous float32 matrices.  With db_partials != NULL, or f_in not a multiple of
 * 32 next to a table, dgnn_dense_bwd_tc takes the older per-thread-load kernel (same results). */
int dgnn_tc_supported(int f_in, int f_out, int gather);
int dgnn_tc_grid(void);
int dgnn_tc_slice(int backward);
int dgnn_tc_packed_floats(int n_rows, int seg_len, int n_segs);
int dgnn_pack_b_tf32(const float* w, int n_rows, int ld, int seg_len, int n_segs, int slice,
                     float* packed, void* stream);
int dgnn_layer_fwd_tc(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                      const int32_t* nbr, const float* ea, int fe,
                      const float* w_e, const float* b_e,
                      const float* b_packed, const float* bias,
                      const float* out_scale, const float* out_shift, int relu_out,
                      int64_t n_tgt, int f_in, int f_out,
                      float* out, float* agg_save, double* stats, void* stream);
int dgnn_dense_bwd_tc(const float* dy, const float* z, const float* g, const float* a, const float* b,
                      const float* mean, const float* rstd,
                      const float* b_packed, const int32_t* nbr, int64_t n_tgt, int f_in, int f_out,
                      float* d_agg, float* d_self, double* db_partials, void* stream);

/* Aggregation with the edge filter on tensor cores (f_in <= 128, 4 <= fe <= 28):
 *   fwd: agg[t] = mean_k h(nbr[t,k]) (*) (w_e . ea[t,k] + b_e)
 *   bwd: dh[s] = d_self[s] (s < n_tgt) + sum_k (w_e . ea_own[s,k] + b_e) (*) d_agg[onbr[s,k]];
 *        dy_prev = relu'(z_prev*p_scale+p_shift) * dh, s_partials double[dgnn_tc_grid(), 2*f_in] = (S1, S2);
 *        dwe_partials float[dgnn_tc_grid(), f_in, 32] (nullable): columns 0..fe-1 = dW_e, column fe = db_e,
 *        accumulated in TMEM from dphi = h(s) (*) d_agg (dphi rounded to TF32, ea split hi/lo)
 * followed by dgnn_dense_fwd_tc: z = [agg | h(x_in)] . W^T (+ epilogue as dgnn_layer_fwd). */
int dgnn_gather_tc_supported(int f, int fe);
int dgnn_gather_tc_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                       const int32_t* nbr, const float* ea, int fe, const float* w_e, const float* b_e,
                       int64_t n_tgt, int f_in, float* agg, void* stream);
int dgnn_gather_tc_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const float* ea_own,
                       int fe, const float* w_e, const float* b_e, const float* z_prev,
                       const float* p_scale, const float* p_shift, const float* p_mean,
                       const float* p_rstd, int p_relu, int64_t n_src, int64_t n_tgt, int f_in,
                       float* dy_prev, double* s_partials, float* dwe_partials, void* stream);
int dgnn_dense_fwd_tc(const float* agg, const float* x_in, const float* in_scale, const float* in_shift,
                      int relu_in, const float* b_packed, const float* bias, const float* out_scale,
                      const float* out_shift, int relu_out, int64_t n_tgt, int f_in, int f_out, float* out,
                      double* stats, void* stream);

/* dW on tensor cores with a TMEM-resident accumulator (f_out <= 128, k_total <= 256);
 * partials float[dgnn_tc_grid(), f_out, k_total], summed by dgnn_reduce_partials_f32.
 * db_partials (nullable) double[dgnn_tc_grid(), f_out]: column sums of dz = the bias gradient (then
 * dgnn_dense_bwd_tc may be called with db_partials == NULL). */
int dgnn_dw_tc_supported(int f_out, int k_total);
int dgnn_dw_bwd_tc(const float* dy, const float* z, const float* g, const float* a, const float* b,
                   const float* mean, const float* rstd,
                   const float* agg, const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                   int64_t n_tgt, int f_in, int f_out, int k_total, float* partials, double* db_partials,
                   void* stream);

/* Updated-edge-filter variant (learning/surfaceNetUpdatedEdgeFilters.py:157-176): aggregation with a
 * materialised edge state phi float32[n_tgt,4,f]:  agg[t] = mean_k h(x_in[nbr[t,k]]) (*) phi[t,k,:] */
int dgnn_gather_phi_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                        const int32_t* nbr, const float* phi, int64_t n_tgt, int f, float* agg, void* stream);
/* Backward of that aggregation (autograd of Updated:157-176, :236-241 in the reference).
 *   dphi[t,k,:] = h(x_in[nbr[t,k]]) (*) d_agg[t]  +  (phi[t,k,:] > 0) * de_next[eid[t,k],:]
 *   h(s) = relu?(x_in[s]*in_scale + in_shift)  (in_scale NULL: no affine)
 * d_agg already divided by max(cnt,1) (dgnn_dense_bwd); de_next float32[E_all,f] indexed by the global
 * edge id eid int32[n_tgt,4] = gradient arriving through relu(edge state) from the next layer (NULL: none). */
int dgnn_upd_edge_bwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                      const int32_t* nbr, const float* d_agg,
                      const float* phi, const float* de_next, const int32_t* eid, int64_t n_tgt, int f,
                      float* dphi, void* stream);
/*   dx[s] = relu'(x_in[s]) * ( d_self[s] (s < n_tgt) + sum_j phi[orow[s,j],:] (*) d_agg[onbr[s,j]] )
 * onbr int32[n_src,4] = target of the j-th out-edge of s, orow = that edge's row 4t+k in phi (-1 = none). */
int dgnn_gather_phi_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const int32_t* orow,
                        const float* phi, const float* x_in, int relu_in, int64_t n_src, int64_t n_tgt,
                        int f, float* dx, void* stream);
/* dst = (z > 0) ? src : 0, elementwise over n_floats (multiple of 4); dst may alias src */
int dgnn_relu_mask(const float* src, const float* z, int64_t n_floats, float* dst, void* stream);

/* Reduce per-CTA (sum, sum^2) partials and produce the normalisation's per-channel affine.
 * mode 0 = BatchNorm1d training statistics (biased var for normalisation; running_mean /
 *          running_var (unbiased) updated with `momentum` when non-NULL),
 * mode 1 = PyG graph LayerNorm (scalar mean / population std over all rows x channels;
 *          eps added to the std).
 * Outputs (float[c] each): scale, shift (y = z*scale + shift), mean, rstd. */
int dgnn_norm_finalize(const double* stats, int n_partials, int64_t n_rows, int c,
                       const float* weight, const float* bias, float eps, float momentum,
                       int mode, float* running_mean, float* running_var,
                       float* scale, float* shift, float* mean, float* rstd, void* stream);
/* eval-mode BatchNorm affine from running statistics */
int dgnn_norm_eval_affine(const float* weight, const float* bias, const float* running_mean,
                          const float* running_var, float eps, int c, float* scale, float* shift,
                          void* stream);

/* out[r,o] = sum_f h(r,f) * w[o,f] + b[o], o < od <= 4, h as in dgnn_layer_fwd */
int dgnn_rowdot_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                    const float* w, const float* b, int64_t n, int f, int od, float* out,
                    void* stream);
/* out = h(x_in) materialised (decoder == 0) */
int dgnn_affine_relu(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                     int64_t n, int f, float* out, void* stream);

/* ---- loss (runModel.py:163-211, kl branch) ---------------------------------------------
 * weight_mode: 0 = w, 1 = sqrt(w), 2 = log(1+w), 3 = 1.  partials = double[grid,2]
 * (sum w*l, sum w), grid = dgnn_small_grid().  dgnn_kl_loss_finalize writes
 * out[0]=loss, out[1]=sum w*l, out[2]=sum w (float). */
int dgnn_small_grid(void);
int dgnn_kl_loss_fwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                     int weight_mode, int64_t n, double* partials, void* stream);
int dgnn_kl_loss_finalize(const double* partials, int n_partials, float* out, void* stream);
/* dlogits[r,k] = grad_out * w_r * (softmax_k * sum_k y - y_k) / sum_w ; sums = out of finalize */
int dgnn_kl_loss_bwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                     int weight_mode, int64_t n, const float* sums, const float* grad_out,
                     float* dlogits, void* stream);
/* runModel.py:181-188, one logit per cell: kind 0 = bce (BCEWithLogits against y, weighted like kl),
 * kind 1 = mse (mean of (sigmoid(z) - y)^2; the reference's weights cancel).  partials double[dgnn_small_grid(), 2]
 * = (numerator, normaliser), finalised by dgnn_kl_loss_finalize; y / w are strided columns. */
int dgnn_point_loss_fwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                        int weight_mode, int kind, int64_t n, double* partials, void* stream);
int dgnn_point_loss_bwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                        int weight_mode, int kind, int64_t n, const float* sums, const float* grad_out,
                        float* dlogits, void* stream);
/* runModel.py:109-160: partial sums of |p0[src]-p0[tgt]| over an edge list (int64) */
int dgnn_edge_reg_fwd(const float* logits, const int64_t* src, const int64_t* tgt, int64_t n_edges,
                      double* partials, void* stream);
/* its gradient (reg_loss is added to the training loss from regularization.edge_epoch on,
 * runModel.py:250-255): dlogits[c] = grad_out[0] * scale * p(1-p) * (sum of sign(p_src - p_tgt) over the
 * edges of c) * (+1, -1).  sign_count int32[n_rows], caller-zeroed workspace (integer atomics: exact and
 * order-independent); scale = edge_weight / n_edges; grad_out is a device scalar. */
int dgnn_edge_reg_bwd(const float* logits, const int64_t* src, const int64_t* tgt, int64_t n_edges,
                      int64_t n_rows, float scale, const float* grad_out, int32_t* sign_count,
                      float* dlogits, void* stream);

/* ---- backward ---------------------------------------------------------------------------
 * Normalisation backward is carried by per-channel coefficient vectors
 *   dz = g*dy - (a + xhat*b),  xhat = (z - mean)*rstd
 * produced by dgnn_norm_bwd_coeffs from the reduced (S1 = sum dy, S2 = sum dy*xhat). */
/* d(final linear): dy[r,f] = relu'(.) * sum_o dlogits[r,o]*w[o,f]; partials double[grid, od*f + od + 2*f]
 * = (dW[od,f], db[od], S1[f], S2[f]) */
int dgnn_rowdot_bwd(const float* dlogits, const float* z_in, const float* in_scale, const float* in_shift,
                    const float* mean, const float* rstd, int relu_in,
                    const float* w, int64_t n, int f, int od, float* dy, double* partials, void* stream);
/* elementwise: dy = relu'(z*scale+shift) * dh (relu_in) and the producer norm's (S1,S2);
 * partials double[dgnn_small_grid(), 2*f]; dy may alias dh. */
int dgnn_act_bwd(const float* dh, const float* z, const float* in_scale, const float* in_shift,
                 const float* mean, const float* rstd, int relu_in, int64_t n, int f, float* dy,
                 double* partials, void* stream);
int dgnn_reduce_partials(const double* partials, int n_partials, int len, float* out, void* stream);
int dgnn_norm_bwd_coeffs(const float* s1, const float* s2, int64_t n_rows, int c, const float* weight,
                         const float* rstd, int mode, float* g, float* a, float* b, void* stream);
/* dense part: dz = g*dy - (a + xhat*b) (g == NULL: dz = dy);  dA = dz . w_cat (w_cat
 * float32[f_out, k_total]);  d_agg[t] = dA[t,:f_in]/max(cnt,1) (nbr != NULL), d_self = rest.
 * db partial sums double[grid, f_out]. */
int dgnn_dense_bwd(const float* dy, const float* z, const float* g, const float* a, const float* b,
                   const float* mean, const float* rstd,
                   const float* w_cat, const int32_t* nbr, int64_t n_tgt, int f_in, int f_out, int k_total,
                   float* d_agg, float* d_self, double* db_partials, void* stream);
/* dW[f_out, k_total] = sum_t dz[t]^T [agg[t] | h(t)];  partials float[splits, f_out, k_total],
 * splits = dgnn_dw_splits(f_out, k_total). agg may be NULL (dense layer, k_total = f_in). */
int dgnn_dw_splits(int f_out, int k_total);
int dgnn_dw_bwd(const float* dy, const float* z, const float* g, const float* a, const float* b,
                const float* mean, const float* rstd,
                const float* agg, const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                int64_t n_tgt, int f_in, int f_out, int k_total, float* partials, void* stream);
int dgnn_reduce_partials_f32(const float* partials, int n_partials, int64_t len, float* out, void* stream);
/* gather part, atomic-free through the out-edge ELL table (onbr[s,k] = target of the k-th
 * out-edge of s, ea_own its attributes):
 *   dh[s]   = d_self[s] (s < n_tgt) + sum_k phi(ea_own[s,k]) (*) d_agg[onbr[s,k]]
 *   dphi_k  = h(s) (*) d_agg[onbr[s,k]] ;  dW_e += dphi_k (x) ea_own[s,k] ; db_e += dphi_k
 *   dy_prev[s] = relu'(.) * dh[s]   (written when dy_prev != NULL), S1/S2 of the producer norm.
 * partials double[grid, f_in*(fe+1) + 2*f_in] = (dW_e[f_in,fe], db_e[f_in], S1[f_in], S2[f_in]). */
int dgnn_gather_bwd_grid(int f_in);
int dgnn_gather_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const float* ea_own, int fe,
                    const float* w_e, const float* b_e,
                    const float* x_in, const float* in_scale, const float* in_shift,
                    const float* in_mean, const float* in_rstd, int relu_in,
                    int64_t n_src, int64_t n_tgt, int f_in,
                    float* dy_prev, double* partials, void* stream);

/* edge-filter gradients only (dW_e, db_e; the S1/S2 part of the partials is zero) */
int dgnn_edge_filter_bwd(const float* d_agg, const int32_t* onbr, const float* ea_own, int fe,
                         const float* w_e, const float* b_e, const float* x_in, const float* in_scale,
                         const float* in_shift, int relu_in, int64_t n_src, int64_t n_tgt, int f_in,
                         double* partials, void* stream);

/* ---- optimiser (torch.optim.Adam defaults, runModel.py:290) ------------------------------ */
int dgnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, int step, void* stream);

/* all tensors in one launch: table = device int64[n_tensors,5] rows (param, grad, exp_avg,
 * exp_avg_sq, numel); max_n = largest numel */
int dgnn_adam_multi(const int64_t* table, int n_tensors, int64_t max_n, float lr, float beta1,
                    float beta2, float eps, int step, void* stream);
/* The same update with lr / beta1 / beta2 / eps (hyper, device float[4]) and the step count (state, device int64[1], >= 1,
 * advanced by the caller on the stream before the launch) read from device memory: a CUDA graph captured around the
 * training step (runModel.GraphedStep) stays valid from replay to replay. */
int dgnn_adam_multi_dev(const int64_t* table, int n_tensors, int64_t max_n, const float* hyper,
                        const int64_t* state, void* stream);

/* ---- labels / facets (generate_mesh.py:75, 94-105) --------------------------------------- */
int dgnn_argmax_labels(const float* logits, int64_t n, int od, uint8_t* labels, void* stream);
/* dataLoader.exportScore (processing/data.py:521-535): sigmoid(logits) and softmax(logits, dim=-1), float32[n, od] each */
int dgnn_scores(const float* logits, int64_t n, int od, float* sigmoid, float* softmax, void* stream);
/* nfacets int32[f,2] in finite-cell numbering (-1 = infinite -> outside); flag[f] = 1 if the two
 * cells' labels differ */
int dgnn_interface_facets(const uint8_t* labels_finite, int64_t n_finite, const int32_t* nfacets,
                          int64_t n_facets, uint8_t* flag, void* stream);

/* ---- halo exchange staging (multi-GPU) ---------------------------------------- */
/* buf[r,:] = x[idx[r],:] / x[base + r,:] = buf[r,:] are dgnn_gather_rows / plain copies; the
 * NCCL send/recv itself stays with the host (torch.distributed) on the same stream. */

/* ---- graph-cut regularisation of the labels (processing/generate_mesh.py:15-58) --------------------------------
 * Two labels, E(l) = sum_c D(c,l_c) + w * #{finite-finite facets with different labels}, D(c,0) =
 * round(z[c,1]*unary_weight), D(c,1) = round(z[c,0]*unary_weight) (float32 product, round-half-even, as
 * `(prediction*unary_weight).round()` after the column swap of :25).  gco's alpha-expansion ends in a global minimum
 * of this submodular energy = one s-t minimum cut, computed by a lock-free push-relabel over the facet table
 * nbr int32[n,4] (-1 = none) / rslot uint8[n,4] (nbr[nbr[c,k], rslot[c,k]] == c):
 *   dgnn_gc_terminals    excess / sink_cap int64[n] from the logits (optionally the two cost columns d0 / d1)
 *   dgnn_gc_bfs_init / dgnn_gc_bfs_step   global relabelling: height = residual distance to the sink, hmax = none;
 *                        step `level` labels the cells one arc further, *changed (device int32) is set if any
 *   dgnn_gc_push_relabel `iters` push / relabel attempts per cell; cap int32[n,4] residual facet capacities (init w)
 *   dgnn_gc_active       adds the number of cells with excess that can still reach the sink to *count
 *   dgnn_gc_labels       labels uint8[n]: 1 (outside) if the cell reaches the sink, else 0 (inside)
 *   dgnn_gc_energy       per-block (data, 2 x smoothness) partial sums int64[dgnn_gc_energy_grid(), 2] */
int dgnn_gc_terminals(const float* logits, int64_t n, float unary_weight, int64_t* excess, int64_t* sink_cap,
                      int64_t* d0, int64_t* d1, void* stream);
int dgnn_gc_push_relabel(int64_t n, const int32_t* nbr, const uint8_t* rslot, int32_t* cap, int64_t* excess,
                         int64_t* sink_cap, int32_t* height, int hmax, int iters, void* stream);
int dgnn_gc_bfs_init(int64_t n, const int64_t* sink_cap, int32_t* height, int hmax, void* stream);
int dgnn_gc_bfs_step(int64_t n, const int32_t* nbr, const uint8_t* rslot, const int32_t* cap, int32_t* height,
                     int level, int hmax, int32_t* changed, void* stream);
int dgnn_gc_active(int64_t n, const int64_t* excess, const int32_t* height, int hmax, uint64_t* count,
                   void* stream);
int dgnn_gc_labels(int64_t n, const int32_t* height, int hmax, uint8_t* labels, void* stream);
int dgnn_gc_energy_grid(void);
int dgnn_gc_energy(int64_t n, const float* logits, float unary_weight, const int32_t* nbr, int binary_weight,
                   const uint8_t* labels, int64_t* partials, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGNN_B200_H */
