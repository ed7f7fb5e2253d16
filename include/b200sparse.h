/*
 * b200sparse.h — C ABI of libb200sparse.so, the sm_100a sparse-3D-convolution engine that sits
 * behind DODA's operator surface (spconv v1.2 / PG_OP / pointops2_cuda).
 *
 * Conventions
 *   - every entry point returns int: 0 = ok, <0 = error (see B200SP_E*); the message is available
 *     from b200sp_last_error() (thread-local).  Nothing here calls exit()/abort().
 *   - pointers marked "dev" are CUDA device pointers, "host" are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - the caller allocates every output (as the reference's Python wrappers do,
 *     lib/pointgroup_ops/functions/pointgroup_ops.py:59,72,138-139,273).
 *   - features are fp32 row-major [rows, C]; indices/rulebooks are int32; coords are
 *     int32 [M,4] = (batch, i0, i1, i2) in the axis order of spatial_shape.
 *
 * Each block cites the reference interface it replaces.
 */
#ifndef B200SPARSE_H_
#define B200SPARSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SP_OK 0
#define B200SP_EINVAL (-1)   /* bad argument */
#define B200SP_ECUDA (-2)    /* CUDA runtime / launch error */
#define B200SP_ENOMEM (-3)   /* workspace too small */
#define B200SP_EUNSUP (-4)   /* configuration not supported by this build */

const char* b200sp_last_error(void);
int b200sp_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
int64_t b200sp_launch_count(void);
/* name of the conv / weight-gradient kernel family the last b200sp_gather_gemm* / b200sp_wgrad* call of this thread
 * dispatched to ("k_conv_direct", "k_conv_tc", "k_gather_gemm", "k_wgrad_direct", "k_wgrad_os", "k_wgrad_tc",
 * "k_wgrad"); bench.py labels its per-kernel roofline rows with it. */
const char* b200sp_last_kernel(void);
/* fork / join of a side stream (cudaEventRecord + cudaStreamWaitEvent in one call): the host runs a layer's weight
 * gradient on a side stream next to its dgrad + BN backward.  Events are created once and reused. */
int b200sp_event_create(void** event_out);
int b200sp_stream_fork(void* main_stream, void* side_stream, void* event);
int b200sp_stream_join(void* main_stream, void* side_stream, void* event);

/* ------------------------------------------------------------------------------------------
 * Rulebook builder.  Replaces spconv v1.2 `ops.get_indice_pairs` (called from
 * SparseConvolution.forward; reference call sites model/unet.py:36, model/unet_block.py:26,29,70,78).
 *
 * Tables produced (all int32, -1 = none):
 *   nbr  [M, K]           SubM: nbr[q,k] = input row at site(q) + (k - centre)   (out -> in view)
 *   order [M]             SubM: the engine's processing order = rows stably sorted by their neighbour bitmask
 *                         (bit k set <=> nbr[q,k] >= 0);
 *   rowmask [M]           rowmask[i] = the K-bit mask of row order[i] (lets the conv kernel find a tile's present
 *                         offsets from 128 words instead of scanning 128 x K table entries);
 *   nbr_perm [M, K]       nbr_perm[i,k] = nbr[order[i],k].  The 128 rows of a tile then share (nearly) one set of
 *                         present offsets: absent offsets drop out per tile and most gathered rows are real
 *   fwd  [M_in, K]        strided conv: fwd[j,k] = output row reached by input j through offset k
 *   bwd  [M_out, K]       strided conv: bwd[o,k] = input row feeding output o through offset k
 *   pairs [2, K, M_in]    spconv layout, canonical order (ascending input row inside each offset)
 *   pairnum [K]
 * ------------------------------------------------------------------------------------------ */

/* bytes of device workspace needed by the two builders below for M_in rows and K offsets */
int64_t b200sp_rulebook_ws_bytes(int64_t M_in, int K, int cand_per_input);

/* SubM (stride 1, output sites == input sites). ksize/dil per axis; padding is k/2 as in spconv. */
int b200sp_rulebook_subm(const int32_t* coords_dev, int64_t M, int batch, const int32_t* shape_host /*[3]*/,
                         const int32_t* ksize_host /*[3]*/, const int32_t* dil_host /*[3]*/,
                         int32_t* nbr_dev /*[M,K] or NULL*/, int32_t* pairs_dev /*[2,K,M] or NULL*/,
                         int32_t* pairnum_dev /*[K] or NULL*/, int32_t* order_dev /*[M] or NULL*/,
                         int32_t* nbr_perm_dev /*[M,K] or NULL*/, int32_t* rowmask_dev /*[M] or NULL*/, void* ws_dev,
                         int64_t ws_bytes, void* stream);

/* Regular (strided) sparse conv.  Output sites are returned in ascending flattened index
 * (the spconv-CUDA convention).  out_coords/bwd must be sized for the upper bound
 * M_in * cand_per_input rows; *n_out_host receives the real count (this call synchronises the
 * stream once to read it). */
int b200sp_rulebook_conv(const int32_t* coords_dev, int64_t M_in, int batch, const int32_t* shape_host,
                         const int32_t* out_shape_host, const int32_t* ksize_host, const int32_t* stride_host,
                         const int32_t* pad_host, const int32_t* dil_host, int cand_per_input,
                         int32_t* out_coords_dev /*[ub,4]*/, int32_t* fwd_dev /*[M_in,K]*/,
                         int32_t* bwd_dev /*[ub,K]*/, int32_t* pairs_dev /*[2,K,M_in] or NULL*/,
                         int32_t* pairnum_dev /*[K] or NULL*/, int64_t* n_out_host, void* ws_dev,
                         int64_t ws_bytes, void* stream);

/* The same build in two halves, so that a host can start it early and never block on it: _begin hashes the
 * candidate output sites and starts an async copy of their count into *n_out_host (pinned host memory) without
 * waiting; after the caller has synchronised (stream or event) it passes the count to _finish together with the
 * SAME, untouched workspace.  b200sp_rulebook_conv == begin + cudaStreamSynchronize + finish. */
int b200sp_rulebook_conv_begin(const int32_t* coords_dev, int64_t M_in, int batch, const int32_t* shape_host,
                               const int32_t* out_shape_host, const int32_t* ksize_host,
                               const int32_t* stride_host, const int32_t* pad_host, const int32_t* dil_host,
                               int cand_per_input, int32_t* n_out_host /*pinned*/, void* ws_dev, int64_t ws_bytes,
                               void* stream);
int b200sp_rulebook_conv_finish(const int32_t* coords_dev, int64_t M_in, int batch, const int32_t* shape_host,
                                const int32_t* out_shape_host, const int32_t* ksize_host,
                                const int32_t* stride_host, const int32_t* pad_host, const int32_t* dil_host,
                                int cand_per_input, int64_t n_out, int32_t* out_coords_dev /*[ub,4]*/,
                                int32_t* fwd_dev /*[M_in,K]*/, int32_t* bwd_dev /*[ub,K]*/,
                                int32_t* pairs_dev /*[2,K,M_in] or NULL*/, int32_t* pairnum_dev /*[K] or NULL*/,
                                void* ws_dev, int64_t ws_bytes, void* stream);

/* pairs [2,K,M_in] (spconv layout) -> out-stationary table tab[n_out,K]; `inverse` swaps roles. */
int b200sp_pairs_to_table(const int32_t* pairs_dev, const int32_t* pairnum_dev, int K, int64_t M_in,
                          int inverse, int32_t* tab_dev /*[n_out,K]*/, int64_t n_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Convolution.  Replaces spconv v1.2 `ops.indice_conv` / `ops.indice_conv_backward`
 * (gather -> SGEMM -> scatter-add per offset) with one output-stationary gather-GEMM launch.
 *
 *   out[orow[r],:] (+)= sum_k  in[tab[r,k], :] @ W[k]    tab == NULL, K == 1  ->  plain GEMM (1x1 conv)
 *                                                         orow == NULL -> identity (table row r is output row r)
 *
 * W is the module's weight viewed as [K][Ci_w][Co_w] contiguous.  wflags = 0: forward (Cin = Ci_w, Cout = Co_w).
 * wflags bit0: use W[k]^T (dgrad: Cin = Co_w, Cout = Ci_w); bit1: mirrored offsets W[K-1-k] (SubM dgrad), with
 * the out->in table of the adjoint; bit2: W_dev is an image made by b200sp_prep_weights_batch with the same bits 0-1.
 * ------------------------------------------------------------------------------------------ */
int b200sp_gather_gemm(const float* in_dev, int64_t n_in, int Cin, const float* W_dev, int wflags,
                       const int32_t* tab_dev, const int32_t* orow_dev, const int32_t* rowmask_dev /*or NULL*/, int K,
                       float* out_dev, int64_t n_out, int Cout, int accumulate, void* ws_dev, int64_t ws_bytes,
                       void* stream);
/* the same with a residual: out = conv + res_dev ([n_out][Cout], must not alias out_dev) -- the `output.features +=
 * identity.features` of model/unet_block.py:37 folded into the conv's epilogue (same fp32 add, one launch less) */
int b200sp_gather_gemm_res(const float* in_dev, int64_t n_in, int Cin, const float* W_dev, int wflags,
                           const int32_t* tab_dev, const int32_t* orow_dev, const int32_t* rowmask_dev /*or NULL*/, int K,
                           float* out_dev, int64_t n_out, int Cout, const float* res_dev, void* ws_dev, int64_t ws_bytes,
                           void* stream);

/* pair-grouped variant (each output row written by exactly one pair; used for the non-overlapping
 * inverse conv forward and the strided conv dgrad):  out[po[k][i],:] = in[pi[k][i],:] @ W[k].
 * pairnum stays on the device (no host sync): n_upper >= max_k pairnum[k] sizes the grid and CTAs past
 * pairnum[k] exit immediately. */
int b200sp_gather_gemm_pairs(const float* in_dev, int Cin, const float* W_dev, int wflags,
                             const int32_t* pairs_in_dev /*[K,stride]*/, const int32_t* pairs_out_dev,
                             const int32_t* pairnum_dev, int64_t n_upper, int K, int64_t pair_stride, float* out_dev,
                             int Cout, int accumulate, void* ws_dev, int64_t ws_bytes, void* stream);

/* Weight images for the tensor path can be prepared ahead, for MANY layers in one launch (weights change once per
 * optimizer step, not per call).  desc_host: n rows of 6 x int64 {W_dev, image_dev, K, Ci_w, Co_w, wflags};
 * image_dev must hold b200sp_conv_prepared_bytes(K, Cin, Cout) bytes (0 = shape not covered: do not prepare);
 * desc_dev: device scratch of >= 64*n bytes.  A prepared image is then passed as W_dev with (wflags | 4). */
int64_t b200sp_conv_prepared_bytes(int K, int Cin, int Cout);
int b200sp_prep_weights_batch(const int64_t* desc_host, int n, void* desc_dev, int64_t desc_dev_bytes, void* stream);

/* device workspace both calls above need (pre-split tensor-core weight image / transposed weights) */
int64_t b200sp_conv_ws_bytes(int K, int Cin, int Cout);
/* 0 (default) = tcgen05 tensor-core path (3xTF32, fp32-accurate) wherever the shape is covered;
 * 1 = fp32 CUDA-core kernel only.  Also selectable with the environment variable B200SP_CONV_IMPL=fp32. */
int b200sp_set_conv_impl(int impl);
/* Inside the tensor path, narrow layers (Cin, Cout in {16, 32}, table mode) run the register-gather
 * kernel (conv_direct.cu: warp-level mma.m16n8k8 3xTF32 fed straight from gathered rows, reads the raw weights);
 * everything else runs the persistent tcgen05 kernel.  on = 0 switches the register-gather kernel off
 * (also B200SP_DIRECT=0); _covers tells which kernel a shape would get. */
int b200sp_set_conv_direct(int on);
/* on = 1: the tcgen05 conv kernel gathers its rows with TMA (cp.async.bulk.tensor tile::gather4, swizzled K-major tiles)
 * instead of per-thread cp.async; also B200SP_TC_TMA=1.  Same results bit for bit; off by default (measured slower for
 * the 64..128-byte rows of this path, csrc/conv_tc.cu). */
int b200sp_set_conv_tma(int on);
int b200sp_conv_direct_covers(int K, int Cin, int Cout);

/* weight gradient: dW[k][ci][co] += sum_i a[pa[k][i]][ci] * b[pb[k][i]][co].   pa/pb NULL -> identity
 * rows (1x1 conv: n_upper rows, pairnum ignored).  dW must be zero-initialised by the caller. */
int b200sp_wgrad(const float* a_dev, int Ca, const float* b_dev, int Cb, const int32_t* pa_dev, const int32_t* pb_dev,
                 const int32_t* pairnum_dev, int64_t n_upper, int K, int64_t pair_stride, float* dW_dev /*[K,Ca,Cb]*/,
                 void* stream);

/* Weight gradient in TABLE form (out-stationary, wgrad_direct.cu), for the shapes _covers reports (Ca, Cb in
 * {16, 32}, not both 32):  dW[k][ca][cb] += sum_r a[tab[r][k]][ca] * g[orow[r]][cb],  tab/orow/rowmask as in b200sp_gather_gemm
 * (SubM: nbr_perm/order/rowmask of b200sp_rulebook_subm; strided conv: bwd table, a = input, g = output gradient;
 * inverse conv: fwd table).  g is read once per row instead of once per pair.  dW must be zeroed by the caller. */
int b200sp_wgrad_table_covers(int K, int Ca, int Cb);
/* 1 when the table form is also the faster choice for a layer of n_rows table rows (the dispatch rule of the layer
 * executor; b200sp_wgrad_table itself accepts every covered shape) */
int b200sp_wgrad_table_prefers(int K, int Ca, int Cb, int64_t n_rows);
int b200sp_wgrad_table(const float* a_dev, int Ca, const float* g_dev, int Cb, const int32_t* tab_dev,
                       const int32_t* orow_dev, const int32_t* rowmask_dev, int64_t n_rows, int K, float* dW_dev,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer executor: ONE call per [BatchNorm -> ReLU ->] sparse conv layer and direction (csrc/layer.cu).  Replaces,
 * per layer, what spconv v1.2's SparseConvolution.forward + SubMConvFunction / SparseConvFunction /
 * SparseInverseConvFunction .forward/.backward and nn.BatchNorm1d + nn.ReLU run from Python
 * (model/unet_block.py:24-29,68-70,76-78): kernel choice (register-gather / tcgen05 table or pair mode / fp32),
 * side-stream weight gradient, BN statistics / apply / backward -- no allocation inside, every buffer from the caller.
 *
 * A rulebook is described by B200SP_RB_WORDS int64 words on the HOST (device pointers of its tables, 0 = absent):
 * ------------------------------------------------------------------------------------------ */
enum {
    B200SP_RB_NBR = 0,     /* SubM out->in table [M][K] */
    B200SP_RB_NBR_PERM,    /* the same in mask-sorted processing order */
    B200SP_RB_ORDER,       /* [M] output row of processing row r */
    B200SP_RB_ROWMASK,     /* [M] K-bit neighbour mask of processing row r */
    B200SP_RB_FWD,         /* strided conv: [n_fine][K] fine row -> coarse row per offset */
    B200SP_RB_BWD,         /* strided conv: [n_coarse][K] coarse row -> fine row per offset */
    B200SP_RB_PAIRS_IN,    /* indice_pairs[0] = [K][pstride] input rows (spconv layout, canonical order) */
    B200SP_RB_PAIRS_OUT,   /* indice_pairs[1] */
    B200SP_RB_PAIRNUM,     /* [K] */
    B200SP_RB_N_FINE,      /* rows of the rulebook's input sites */
    B200SP_RB_N_COARSE,    /* rows of its output sites (= N_FINE for SubM) */
    B200SP_RB_PSTRIDE,
    B200SP_RB_K,
    B200SP_RB_NONOVERLAP,  /* strided conv with one candidate per input (k == s): pair mode is atomics-free */
    B200SP_RB_WORDS = 16
};
/* args_host: int64 vector (pointers and sizes; eps / momentum as IEEE-754 double bit patterns), layout documented at
 * the definitions in csrc/layer.cu and built by doda_b200/ops.py (_layer_fwd / _layer_bwd). */
int b200sp_conv_layer_fwd(const int64_t* args_host, int n);
int b200sp_conv_layer_bwd(const int64_t* args_host, int n);

/* out[k'][co][ci] = W[k][ci][co], k' = mirror ? K-1-k : k   (weights for dgrad) */
int b200sp_weight_transpose(const float* W_dev, int K, int Cin, int Cout, int mirror, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm(+ReLU) on active sites.  Replaces nn.BatchNorm1d(eps=1e-4, momentum=0.1) + nn.ReLU on the
 * [N_active, C] feature matrix (model/unet.py:28,43; model/unet_block.py:24-28,46-47,68-69,76-77) and
 * DSNorm's F.batch_norm call (model/dsnorm.py:79-84).
 * ------------------------------------------------------------------------------------------ */
/* workspace for the two calls below: the caller zero-fills it ONCE after allocation (it holds a completion ticket
 * that every launch leaves at zero again); it may then be reused by any number of stream-ordered calls */
int64_t b200sp_bn_ws_bytes(int64_t M, int C);
/* training: batch statistics -> mean/invstd (saved for backward); y = [relu]((x-mean)*invstd*w + b).
 * running_mean/var (nullable) are updated in place with `momentum` and the unbiased variance exactly as
 * F.batch_norm does; num_batches_tracked (int64 on device, nullable) is incremented. */
int b200sp_bn_fwd_train(const float* x_dev, int64_t M, int C, const float* w_dev, const float* b_dev, float eps,
                        int relu, float* y_dev, float* mean_dev /*[C]*/, float* invstd_dev /*[C]*/,
                        float* running_mean_dev, float* running_var_dev, float momentum,
                        int64_t* num_batches_tracked_dev, void* ws_dev, int64_t ws_bytes, void* stream);
/* inference / fixed statistics: y = [relu](x*scale + shift) */
int b200sp_affine_relu(const float* x_dev, int64_t M, int C, const float* scale_dev, const float* shift_dev,
                       int relu, float* y_dev, void* stream);
int b200sp_bn_bwd(const float* x_dev, const float* dy_dev, int64_t M, int C, const float* w_dev, const float* b_dev,
                  const float* mean_dev, const float* invstd_dev, int relu, float* dx_dev, float* dw_dev,
                  float* db_dev, void* ws_dev, int64_t ws_bytes, void* stream);
/* the same, dx = BN backward + add_dev ([M][C] or NULL): the gradient of a second consumer of x (skip connection) */
int b200sp_bn_bwd_add(const float* x_dev, const float* dy_dev, int64_t M, int C, const float* w_dev, const float* b_dev,
                      const float* mean_dev, const float* invstd_dev, int relu, float* dx_dev, float* dw_dev,
                      float* db_dev, const float* add_dev, void* ws_dev, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Point <-> voxel.  Replaces PG_OP.voxelize_idx (CPU, lib/pointgroup_ops/src/voxelize/voxelize.cpp:11-155),
 * voxelize_fp / voxelize_bp / point_recover_fp / point_recover_bp (voxelize.cu:10-52, voxelize.cpp:185-205)
 * and the devoxelize gather `features[p2v]` (model/unet.py:62) with its scatter-add backward.
 * ------------------------------------------------------------------------------------------ */
/* CPU, re-entrant, never touches CUDA (runs in forked DataLoader workers). Two-call protocol:
 * pass out pointers NULL to obtain M and maxActive, then call again with buffers. */
int b200sp_voxelize_idx_cpu(const int64_t* coords_host, int64_t N, int ncol /*3 or 4*/, int batch_size, int mode,
                            int64_t* out_coords_host /*[M,ncol]*/, int32_t* input_map_host /*[N]*/,
                            int32_t* output_map_host /*[M,1+maxActive]*/, int64_t* M_out, int32_t* max_active_out);
/* out[v,:] (+)= mult * sum_i feats[map[v,1+i],:]   (mult = 1/count when average) ; no atomics */
int b200sp_voxelize_fp(const float* feats_dev, float* out_dev, const int32_t* map_dev, int average, int64_t M,
                       int max_active, int C, void* stream);
/* d_feats[map[v,1+i],:] += mult * d_out[v,:] */
int b200sp_voxelize_bp(const float* dout_dev, float* dfeats_dev, const int32_t* map_dev, int average, int64_t M,
                       int max_active, int C, void* stream);
/* dst[r, dst_col0 : dst_col0 + ncols] = src[r, src_col0 : src_col0 + ncols] for r < rows (row strides in floats): the
 * channel concat of the U-Net skip connection (model/unet_block.py:95) and the split of its gradient, one 2-D copy */
int b200sp_copy_cols(const float* src_dev, int64_t rows, int src_stride, int src_col0, int ncols, float* dst_dev,
                     int dst_stride, int dst_col0, void* stream);
/* out[i,:] = src[idx[i],:]   (idx int32 or int64) */
int b200sp_gather_rows(const float* src_dev, const void* idx_dev, int idx_is_i64, int64_t n, int C, float* out_dev,
                       void* stream);
/* dst[idx[i],:] += src[i,:] */
int b200sp_scatter_add_rows(const float* src_dev, const void* idx_dev, int idx_is_i64, int64_t n, int C,
                            float* dst_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Remaining PG_OP kernels (lib/pointgroup_ops/src/pointgroup_ops_api.cpp:7-26).
 * ------------------------------------------------------------------------------------------ */
/* sec_mean.cu:12-109 — CSR-segment reductions, offsets int32 [P+1] */
int b200sp_sec_mean(const float* inp_dev, const int32_t* offsets_dev, float* out_dev, int P, int C, void* stream);
int b200sp_sec_mean_bp(const float* dout_dev, const int32_t* offsets_dev, float* dinp_dev, int P, int C, void* stream);
int b200sp_sec_min(const float* inp_dev, const int32_t* offsets_dev, float* out_dev, int P, int C, void* stream);
int b200sp_sec_max(const float* inp_dev, const int32_t* offsets_dev, float* out_dev, int P, int C, void* stream);
/* bfs_cluster.cu:15-89 — radius neighbours inside the point's batch segment. Returns the total
 * neighbour count in *n_active_host (may exceed n*mean_active: caller retries, as the reference does). */
int64_t b200sp_ballquery_ws_bytes(int n);
int b200sp_ballquery_batch_p(const float* xyz_dev, const int32_t* batch_idxs_dev, const int32_t* batch_offsets_dev,
                             int32_t* idx_dev, int32_t* start_len_dev, int n, int mean_active, float radius,
                             void* ws_dev, int64_t ws_bytes, int32_t* n_active_host, void* stream);
/* Faster two-call form: give the size query input_map (out_coords = output_map = NULL) and it also returns the
 * point -> voxel map; _fill then builds out_coords / output_map from that map without hashing a second time. */
int b200sp_voxelize_idx_cpu_fill(const int64_t* coords_host, int64_t N, int ncol, int mode,
                                 const int32_t* input_map_host /*[N], from the first call*/, int64_t M,
                                 int32_t max_active, int64_t* out_coords_host, int32_t* output_map_host);
/* The same on the device (batch assembly without the serial CPU hash map), bit-identical results
 * (first-touch voxel order, ascending points per voxel).  coords: int64 [N, ncol] on the device, non-negative,
 * x, y, z < 2^20, batch index < 2^(64 - 3 * bits(largest coordinate)) - 1 (any batch up to a 2^17-wide grid; a
 * 2^20-wide grid leaves 4 bits).  Two calls: _begin writes input_map [N] and starts an async copy of
 * info[4] = {M, maxActive (modes 3/4), bad-coordinate flag, duplicate flag} to pinned host memory; after a stream
 * sync the caller sizes out_coords [M, ncol] / output_map [M, 1 + maxActive] (maxActive = 1 for modes 0-2) and calls
 * _finish with the SAME workspace (b200sp_voxelize_idx_gpu_ws_bytes(N) bytes). */
int64_t b200sp_voxelize_idx_gpu_ws_bytes(int64_t N);
int b200sp_voxelize_idx_gpu_begin(const int64_t* coords_dev, int64_t N, int ncol, int mode, int32_t* input_map_dev,
                                  int32_t* info_host /*[4] pinned*/, void* ws_dev, int64_t ws_bytes, void* stream);
int b200sp_voxelize_idx_gpu_finish(const int64_t* coords_dev, int64_t N, int ncol, int mode, int64_t M,
                                   int max_active, int64_t* out_coords_dev, int32_t* output_map_dev, void* ws_dev,
                                   int64_t ws_bytes, void* stream);

/* bfs_cluster.cpp:28-111 — CPU BFS connected components. Two-call protocol like voxelize_idx. */
int b200sp_bfs_cluster_cpu(const int32_t* sem_host, const int32_t* idx_host, const int32_t* start_len_host, int N,
                           int threshold, int32_t* cluster_idxs_host /*[sum,2]*/, int32_t* cluster_offsets_host,
                           int64_t* n_idx_out, int64_t* n_cluster_out);
/* roipool.cu:12-49 */
int b200sp_roipool_fp(const float* feats_dev, const int32_t* offsets_dev, float* out_dev, int32_t* maxidx_dev,
                      int P, int C, void* stream);
int b200sp_roipool_bp(const float* dout_dev, const int32_t* maxidx_dev, float* dfeats_dev, int P, int C, void* stream);
/* get_iou.cu:12-29 */
int b200sp_get_iou(const int32_t* proposals_idx_dev, const int32_t* proposals_offset_dev,
                   const int64_t* instance_labels_dev, const int32_t* instance_pointnum_dev, float* iou_dev, int P,
                   int I, void* stream);
/* knn.cu:7-50 */
int b200sp_knn_batch(const float* xyz_dev, const float* query_dev, const int32_t* batch_idxs_dev,
                     const int32_t* query_offsets_dev, int32_t* idx_dev, int n, int m, int k, void* stream);

/* ------------------------------------------------------------------------------------------
 * pointops2_cuda (lib/pointops2/src/pointops_api.cpp:13-23).  `offset` arguments are END offsets
 * without the leading 0 (functions/pointops2.py:47,66).
 * ------------------------------------------------------------------------------------------ */
int b200sp_knnquery(int m, int nsample, const float* xyz_dev, const float* new_xyz_dev, const int32_t* offset_dev,
                    const int32_t* new_offset_dev, int32_t* idx_dev, float* dist2_dev, void* stream);
int b200sp_furthestsampling(int b, int n_max, int dim, const float* xyz_dev, const int32_t* offset_dev,
                            const int32_t* new_offset_dev, float* tmp_dev, int32_t* idx_dev, void* stream);
int b200sp_grouping_fwd(int m, int nsample, int c, const float* in_dev, const int32_t* idx_dev, float* out_dev,
                        void* stream);
int b200sp_grouping_bwd(int m, int nsample, int c, const float* dout_dev, const int32_t* idx_dev, float* din_dev,
                        void* stream);
int b200sp_interpolation_fwd(int n, int c, int k, const float* in_dev, const int32_t* idx_dev,
                             const float* weight_dev, float* out_dev, void* stream);
int b200sp_interpolation_bwd(int n, int c, int k, const float* dout_dev, const int32_t* idx_dev,
                             const float* weight_dev, float* din_dev, void* stream);
int b200sp_subtraction_fwd(int n, int nsample, int c, const float* in1_dev, const float* in2_dev,
                           const int32_t* idx_dev, float* out_dev, void* stream);
int b200sp_subtraction_bwd(int n, int nsample, int c, const int32_t* idx_dev, const float* dout_dev,
                           float* din1_dev, float* din2_dev, void* stream);
int b200sp_aggregation_fwd(int n, int nsample, int c, int w_c, const float* in_dev, const float* pos_dev,
                           const float* w_dev, const int32_t* idx_dev, float* out_dev, void* stream);
int b200sp_aggregation_bwd(int n, int nsample, int c, int w_c, const float* in_dev, const float* pos_dev,
                           const float* w_dev, const int32_t* idx_dev, const float* dout_dev, float* din_dev,
                           float* dpos_dev, float* dw_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Loss epilogue: softmax cross-entropy over point logits, replaces the nn.CrossEntropyLoss call
 * of model/unet.py:168-170 (ignore_index, optional per-class weights, mean reduction).
 * out2[0] = loss, out2[1] = sum of the weights of the counted rows (kept for the backward).
 * ws: b200sp_cross_entropy_ws_bytes() bytes, zero-filled once by the caller.  C <= 64.
 * ------------------------------------------------------------------------------------------ */
int64_t b200sp_cross_entropy_ws_bytes(void);
int b200sp_cross_entropy_fwd(const float* logits_dev /*[N,C]*/, const int64_t* labels_dev /*[N]*/,
                             const float* weight_dev /*[C] or NULL*/, int64_t N, int C, int64_t ignore_index,
                             float* out2_dev /*[2]*/, void* ws_dev, int64_t ws_bytes, void* stream);
int b200sp_cross_entropy_bwd(const float* logits_dev, const int64_t* labels_dev, const float* weight_dev, int64_t N,
                             int C, int64_t ignore_index, const float* out2_dev, const float* dloss_dev /*[1]*/,
                             float* dlogits_dev /*[N,C]*/, void* stream);

/* Metric epilogue, replaces util/common_utils.py:233-247 (intersectionAndUnionGPU: three torch.histc calls on CPU
 * copies, i.e. three device->host syncs per iteration at tool/train.py:113-118).  pred / label: int64 [N];
 * out3k: float [3][K] = per-class intersection, union, target counts (float, as histc returns them), on the device,
 * no host sync.  ws: >= 12 K bytes of device scratch.  K <= 1024. */
int b200sp_intersection_union(const int64_t* pred_dev, const int64_t* label_dev, int64_t N, int K,
                              int64_t ignore_index, float* out3k_dev, void* ws_dev, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Augmentation hot spots, replacing the numpy / scipy bodies of
 * dataset/augmentor/augmentor_utils.py on the data-loading side.  Arithmetic as the reference: float32 noise grids
 * blurred with double accumulation per pass, double interpolation, double coordinates out.
 * ------------------------------------------------------------------------------------------ */
/* elastic, augmentor_utils.py:62-73: six 3-tap box passes (axes 0,1,2,0,1,2; zero outside the grid) over the three
 * noise grids noise_dev float[3][nx][ny][nz] in place; scratch_dev: same size. */
int b200sp_elastic_blur(float* noise_dev, float* scratch_dev, int nx, int ny, int nz, void* stream);
/* elastic, augmentor_utils.py:74-80: out = x + mag * trilinear(noise, x) on the axes linspace(-(n-1)*gran, (n-1)*gran, n);
 * points outside the grid get 0 (bounds_error=0, fill_value=0).  xyz_dev: float or double [N,3]; out_dev: double [N,3]. */
int b200sp_elastic_apply(const void* xyz_dev, int xyz_is_f64, int64_t N, const float* noise_dev, int nx, int ny, int nz,
                         double gran, double mag, double* out_dev, void* stream);
/* crop, augmentor_utils.py:459-470: xyz_offset = xyz + offset; valid &= (xyz_offset.min(1) >= 0) & all(xyz_offset <
 * full_scale); *count_dev = valid.sum().  valid_u8_dev: uint8 [N] in/out; xyz_offset_dev: double [N,3] or NULL. */
int b200sp_crop_mask(const double* xyz_dev, int64_t N, const double* offset3_host, const double* full_scale3_host,
                     void* valid_u8_dev, double* xyz_offset_dev, int32_t* count_dev, void* stream);
/* scene_aug, augmentor_utils.py:103: out = xyz @ m (m9_host row-major 3x3: jitter / flip / rotation) */
int b200sp_affine3(const void* xyz_dev, int xyz_is_f64, int64_t N, const double* m9_host, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPARSE_H_ */
