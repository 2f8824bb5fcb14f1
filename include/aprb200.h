/* aprb200.h — C-ABI of libaprb200.so: the B200-native (sm_100a) KPConv neighbourhood pipeline.
 *
 * Drop-in boundary for the hot path of Predator_APR's KFE encoder. Every entry point replaces one reference
 * interface (paths relative to /root/reference/Predator_APR):
 *
 *   aprb_grid_subsample_batch      <- batch_grid_subsampling()      cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.h:95-104
 *                                     (bound by cpp_subsampling/wrapper.cpp:62-333 as grid_subsampling.subsample_batch,
 *                                      and :338-566 as grid_subsampling.subsample with B = 1)
 *   aprb_radius_neighbors_batch    <- batch_nanoflann_neighbors()   cpp_wrappers/cpp_neighbors/neighbors/neighbors.h:24-29
 *                                     (bound by cpp_neighbors/wrapper.cpp:58-238 as radius_neighbors.batch_query;
 *                                      the [:, :max_neighbors] cut of datasets/dataloader.py:66-70 is folded in)
 *   aprb_kpconv_forward            <- KPConv.forward                models/blocks.py:229-374 (rigid, linear influence, sum)
 *   aprb_max_pool / aprb_closest_pool <- max_pool / closest_pool    models/blocks.py:86-102 / :71-83
 *   aprb_instnorm_lrelu            <- BatchNormBlock.forward (= InstanceNorm1d over all rows) + LeakyReLU(0.1)
 *                                                                   models/blocks.py:459-468, :496-510, :592, :669, :681
 *   aprb_linear_tf32               <- UnaryBlock.mlp (nn.Linear, no bias)  models/blocks.py:493, :500
 *
 * Conventions: every pointer named d_* is a DEVICE pointer; `stream` is a cudaStream_t passed as void*; calls are
 * stream-ordered and asynchronous unless stated. Every function returns 0 on success or a negative aprb_status;
 * aprb_last_error() returns a thread-local message for the last failure. No C++ exception crosses this boundary.
 * Points are packed [N,3] fp32 rows (12-byte stride, the layout of PointXYZ, cpp_utils/cloud/cloud.h:40-104).
 */
#ifndef APRB200_H
#define APRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    APRB_OK = 0,
    APRB_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unsupported shape)      */
    APRB_ERR_WORKSPACE = -2,   /* workspace too small for the request                                */
    APRB_ERR_CUDA = -3,        /* a CUDA runtime call or kernel launch failed                        */
    APRB_ERR_UNSUPPORTED = -4, /* valid in the reference but not implemented on this path            */
    APRB_ERR_EMPTY = -5        /* result is empty: the reference raises RuntimeError("Error") here   */
} aprb_status;

int aprb_version(void);
const char* aprb_last_error(void);
/* Kernel launches issued by this library since load (CUB calls count their internal kernels). */
long long aprb_launch_count(void);
/* Per-kernel timing with CUDA events on the launching stream: enable, run, then aprb_prof_report() synchronises the
 * device and writes "kernel_name launches total_ms" lines into buf and clears the records. */
int aprb_prof_enable(int on);
int aprb_prof_report(char* buf, size_t cap);
/* Tuning switches (process-wide, for A/B measurements; defaults are the measured-best settings):
 *   "kpconv_f16" 1, "act_f16" 1   fp16 operands / fp16 activation storage in aprb_kfe_forward (0: TF32 in fp32 storage)
 *   "fuse_stats" 1                InstanceNorm statistics from the GEMM epilogue's group partials
 *   "gemm_apply" 1                recompute path for the block-closing Linears (0: stored fp32 products; 2: every block)
 *   "kpconv_fused" 0              one-kernel fused KPConv (parity-green, slower)
 *   "blocking_sync" 1             host waits sleep (cudaEventBlockingSync) instead of spinning; read at aprb_kfe_create
 *   "gemm_persistent" 1, "gemm_bn" 0, "gemm_costages" 1, "gemm_cluster" 1, "gemm_stages" 0, "nrm_park" 0,
 *   "kpw_version" 4, "kpconv_chunk_mb" 0, "host_zero_copy" 0   kernel-selection experiments kept for reproducibility
 *   "dbg_skip_d2h" 0              diagnostic: skip the host-output copy
 * Unknown names return APRB_ERR_INVALID. */
int aprb_set_option(const char* name, int value);
/* Device properties the host side sizes grids with (SM count etc.). Returns status. */
int aprb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- K1: voxel-grid barycentre subsampling ------ */
/* Workspace bytes needed for a stacked batch of N points in B clouds (fdim feature columns, 0 if none). */
size_t aprb_grid_subsample_ws_bytes(int N, int B, int fdim);

/* d_pts [N,3] fp32 stacked clouds, d_lens [B] int32 (sum == N). Output rows are emitted per cloud in ascending
 * voxel-key order (canonical order; the reference's order is unordered_map iteration order, unspecified).
 * d_out_pts capacity N rows; d_out_lens [B]; d_out_M [1] int32 (device) receives the total row count.
 * Optional: d_feats [N,fdim] -> d_out_feats [<=N,fdim] (per-voxel mean), pass NULL/0 to skip.
 * max_p <= 0 means unlimited (grid_subsampling.cpp:133-134). Asynchronous: read d_out_M / d_out_lens after
 * synchronising `stream`. key_bits (32 or 64) is the width of the packed (cloud, iz, iy, ix) sort key: 32 is the fast
 * path and fits every LiDAR-sized grid. d_status [1] int32 (device, optional) receives 0 = ok, 2 = the grid needs
 * more than key_bits bits (outputs invalid: call again with key_bits = 64), 1 = it does not fit 64 bits either. */
int aprb_grid_subsample_batch(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p,
                              const float* d_feats, int fdim,
                              float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, float* d_out_feats,
                              int32_t* d_status, int key_bits, void* d_ws, size_t ws_bytes, void* stream);

/* The same with the per-voxel label vote of the reference (grid_subsampling.h:62-75 update_classes,
 * grid_subsampling.cpp:96-101): d_classes [N,ldim] int32 -> d_out_classes [<=N,ldim], the most frequent label of the
 * voxel's points per label dimension. Ties: the smallest label value (the reference: unordered_map iteration order,
 * unspecified). Pass d_classes = NULL / ldim = 0 to skip; same workspace as aprb_grid_subsample_batch. */
int aprb_grid_subsample_batch_labels(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p,
                                     const float* d_feats, int fdim, const int32_t* d_classes, int ldim,
                                     float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, float* d_out_feats,
                                     int32_t* d_out_classes, int32_t* d_status, int key_bits, void* d_ws, size_t ws_bytes,
                                     void* stream);

/* First-level voxelisation of RAW scans (the step before the path): d_raw [N, stride] fp32 rows whose first
 * three floats are x, y, z (KITTI .bin: stride 4 = x, y, z, reflectance, datasets/kitti.py:191-194), stacked clouds of
 * d_lens [B] points, with open3d's PointCloud.voxel_down_sample(voxel_size) semantics (kitti.py:468-471, :588-589;
 * open3d==0.10.0.0 is a third-party dependency absent from this tree: parity UNPINNED, the published algorithm is restated):
 * points widened to double, grid origin = min_bound - 0.5 * voxel_size, voxel = floor((p - origin) / voxel_size), per-voxel
 * double sum in point order, output = float(sum / count) — the narrowing datasets/dataloader.py:125/:163 applies anyway.
 * Rows come out in ascending (cloud, iz, iy, ix) order (open3d: unordered_map order). Outputs / status / key_bits as in
 * aprb_grid_subsample_batch. */
size_t aprb_voxel_downsample_ws_bytes(int N, int B);
int aprb_voxel_downsample_raw(const float* d_raw, int stride, const int32_t* d_lens, int B, int N, double voxel_size,
                              float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, int32_t* d_status, int key_bits,
                              void* d_ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- K2+K3: batched radius neighbour search ----- */
size_t aprb_radius_neighbors_ws_bytes(int Nq, int Ns, int B);

/* For every query, the supports of the same batch element with d2 < radius^2 (fp32, d2 = (dx*dx+dy*dy)+dz*dz,
 * no FMA), ascending by (d2, support index), truncated to `width` columns, padded with Ns.
 * d_out_idx is [Nq, ld] int32 row-major (ld >= width); columns [0,width) of every row are written.
 * d_counts [Nq] int32 (optional) receives the untruncated neighbour count of every query;
 * d_max_count [1] int32 (optional, device) receives max over queries (the reference's output width,
 * neighbors.cpp:296-304). width must be in [1, 16384]. Asynchronous. */
int aprb_radius_neighbors_batch(const float* d_q, const float* d_s, const int32_t* d_qlens, const int32_t* d_slens,
                                int B, int Nq, int Ns, float radius, int width,
                                int32_t* d_out_idx, int ld, int32_t* d_counts, int32_t* d_max_count,
                                void* d_ws, size_t ws_bytes, void* stream);

/* The same search split in two, so several query sets can share one cell list (in the KPConv pyramid the grid over
 * P_l with cell r_l serves conv(l) and pool(l); the grid over P_{l+1} with cell 2 r_l serves upsample(l), conv(l+1),
 * pool(l+1): 4 builds instead of 10). A grid built with `radius` serves any query radius <= that value.
 * d_grid is an opaque device buffer of aprb_cell_grid_bytes(Ns, B) bytes that must outlive its queries; queries on
 * one grid must be ordered on one stream (the grid keeps the query offsets). */
size_t aprb_cell_grid_bytes(int Ns, int B);
int aprb_cell_grid_build(const float* d_s, const int32_t* d_slens, int B, int Ns, float radius,
                         void* d_grid, size_t grid_bytes, void* stream);
int aprb_cell_grid_query(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                         int Nq, int Ns, float radius, int width, int32_t* d_out_idx, int ld,
                         int32_t* d_counts, int32_t* d_max_count, void* stream);

/* Nearest support within `radius` only: d_out_idx[n * ld] = column 0 of the matrix aprb_cell_grid_query would write (same
 * d2 arithmetic, ties by ascending index, pad Ns). All the reference reads of an upsample matrix (closest_pool: inds[:, 0],
 * models/blocks.py:71-83). */
int aprb_cell_grid_query_nearest(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                                 int Nq, int Ns, float radius, int32_t* d_out_idx, int ld, void* stream);
/* aprb_cell_grid_query that also records, per segment of clouds_per_segment consecutive clouds (one collated pair), the
 * width of the reference's matrix for that collate: d_seg_width[s] = min(max neighbour count in the segment, width). */
int aprb_cell_grid_query_seg(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                             int Nq, int Ns, float radius, int width, int32_t* d_out_idx, int ld,
                             int clouds_per_segment, int32_t* d_seg_width, void* stream);

/* ---------------------------------------------------------------- K5: KPConv forward ------------------------- */
/* Prepared weights: the [K,Cin,Cout] fp32 parameter re-laid as the K-major, TF32-rounded B operand
 * [Cout, K*Cin] used by the tcgen05 contraction. d_wprep has K*Cin*Cout floats. Run once per weight update. */
int aprb_kpconv_prepare_weights(const float* d_W, int K, int Cin, int Cout, float* d_wprep, void* stream);
/* fp16 form of the same operand (K*Cin*Cout halves), for mode 3 of aprb_kpconv_forward. */
int aprb_kpconv_prepare_weights_f16(const float* d_W, int K, int Cin, int Cout, void* d_wprep16, void* stream);

size_t aprb_kpconv_ws_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout);

/* out[n,:] = (sum_k (sum_h max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / extent) * x[idx[n,h],:]) @ W[k]) / max(1, nn[n])
 * with nn[n] = #{h : sum_c x[idx[n,h],c] > 0}; idx == Ns is the shadow neighbour (far point, zero feature row).
 * d_idx is [Nq, ld_idx] int32 (idx_is_i64 == 0) or int64 (idx_is_i64 != 0); only columns [0,H) are read.
 * d_W is the raw [K,Cin,Cout] parameter (used by the fp32 path), d_wprep the prepared operand (tensor path;
 * may be NULL to force the fp32 CUDA-core path). mode: 0 = auto (tensor path when supported), 1 = fp32 CUDA cores,
 * 2 = tcgen05 TF32 (error if unsupported shape), 3 = tcgen05 with fp16 operands: d_wprep is then the fp16 operand of
 * aprb_kpconv_prepare_weights_f16 and the weighted tile is produced in fp16 — the 10-bit mantissa of TF32 at half the
 * bytes (K*Cin % 64 == 0, Cin % 4 == 0, H <= 128). Every fp16 store of this path saturates: a weighted sum outside the
 * fp16 range clamps to +-65504 instead of becoming inf (features are expected to be O(1), e.g. InstanceNorm outputs).
 * 4 = as 3, and d_x itself is fp16 [Ns, Cin] (int32 indices, Cin a multiple of the producer's 32/64/128/256/512 slab);
 *     for Cin % 64 == 0 the weighting kernel multiplies with fma.rn.f32.f16: the influence weight is rounded to fp16 as well
 *     (fp32 accumulation; aprb_set_option("kpw_fh", 0) keeps fp32 weights).
 * 5 = fp16 features like 4, with the weighting stage itself on tcgen05 (kpconv_tc.cu) and d_wprep the ck-ordered operand
 *     of aprb_kpconv_prepare_weights_f16_ck; shapes per aprb_kpconv_tc_supported. */
int aprb_kpconv_forward(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                        const float* d_x, const float* d_kp, const float* d_W, const float* d_wprep,
                        float extent, int Nq, int Ns, int H, int K, int Cin, int Cout,
                        float* d_out, int mode, void* d_ws, size_t ws_bytes, void* stream);
/* Same operator; when d_gstat is given (aprb_group_stats_bytes(Nq, Cout) bytes) and the contraction finishes in one
 * tcgen05 GEMM, its epilogue also records the column statistics of every whole 32-row group of d_out —
 * d_gstat[g][0][c] = mean, d_gstat[g][1][c] = sum of squared deviations from that mean, rows [32g, 32g+32) — and sets
 * *stats_written = 1: the InstanceNorm that follows (models/blocks.py:459-468) then skips its own statistics pass
 * (aprb_instnorm_lrelu_seg_pre). *stats_written = 0 (fp32 path, split-K, row chunks) means d_gstat was not touched. */
size_t aprb_group_stats_bytes(int N, int C);
int aprb_kpconv_forward_stats(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                              const float* d_x, const float* d_kp, const float* d_W, const float* d_wprep,
                              float extent, int Nq, int Ns, int H, int K, int Cin, int Cout,
                              float* d_out, int mode, float* d_gstat, int* stats_written,
                              void* d_ws, size_t ws_bytes, void* stream);

/* Training path. Stage A+B alone: d_wf [Nq, K*Cin] = sum_h w[n,k,h] x[idx[n,h],:] and
 * d_inv_nn [Nq] = 1 / max(1, neighbor_num) (blocks.py:269-354, :369-371), so that autograd can keep wf for the weight
 * gradient dW = wf^T (dOut * inv_nn). Workspace: aprb_kpconv_weighted_ws_bytes(Ns). */
size_t aprb_kpconv_weighted_ws_bytes(int Ns);
int aprb_kpconv_weighted(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                         const float* d_x, const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin,
                         int round_tf32, float* d_wf, float* d_inv_nn, void* d_ws, size_t ws_bytes, void* stream);
/* Stage A+B in the native pipeline's storage format: d_x16 [Ns, Cin] fp16 -> fp16 weighted tile (+ d_inv_nn), int32
 * indices. layout_ck = 0: d_wf16 [Nq, K*Cin] (kernel-point-major) by the CUDA-core list kernel; layout_ck = 1: d_wf16
 * [Nq, Cin*16] (channel-major, kernel point minor, zero beyond K) by the tcgen05 weighting kernel (kpconv_tc.cu: per
 * query, D[c, k] = sum_h X[c, h] w[h, k] with the gathered feature rows as an MN-major operand; Cin in {64, 128, 256},
 * H <= 64 — aprb_kpconv_tc_supported). */
int aprb_kpconv_weighted_f16(const float* d_q, const float* d_s, const int32_t* d_idx, int ld_idx, const void* d_x16,
                             const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin, int layout_ck,
                             void* d_wf16, float* d_inv_nn, void* d_ws, size_t ws_bytes, void* stream);
int aprb_kpconv_tc_supported(int H, int K, int Cin, int Cout, int Ns);
/* Prepared operand for mode 5: [Cout, Cin*16] fp16, column c*16 + k = W[k, c, :] (zero for k >= K). */
int aprb_kpconv_prepare_weights_f16_ck(const float* d_W, int K, int Cin, int Cout, void* d_wprep16ck, void* stream);
/* Data gradient of stage A+B: d_dx [Ns,Cin] = scatter-add over the neighbour lists of w[n,k,h] * d_dwf[n,k,:]
 * (d_dx is zeroed by the call; fp32 red.add, so the summation order is not fixed). */
int aprb_kpconv_backward_data(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                              const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin,
                              const float* d_dwf, float* d_dx, void* stream);

/* ---------------------------------------------------------------- K4: pooling / upsampling gathers ----------- */
/* out[n,c] = max_h (x ++ 0)[idx[n,h], c] over h < min(H, *d_width if non-NULL). */
int aprb_max_pool(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H, int C,
                  const int32_t* d_width, float* d_out, void* stream);
/* Super-batched form: rows [d_seg_off[s], d_seg_off[s+1]) of the query set belong to segment s (one collated pair) and use
 * the width d_seg_width[s] — the width of the reference's pool matrix for that pair, min(max_count, limit)
 * (neighbors.cpp:296-304, dataloader.py:66-70): a fixed-width device matrix has extra all-pad columns when
 * max_count < limit, and those must not inject the zero shadow row into a row the reference sees full. S == 1 may pass
 * d_seg_off == NULL. d_x / d_out are fp32, or fp16 when x_is_f16 (C a multiple of 128, <= 1024; int32 indices). */
int aprb_max_pool_seg(const void* d_x, int x_is_f16, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H,
                      int C, const int32_t* d_seg_off, int S, const int32_t* d_seg_width, void* d_out, void* stream);
/* d_seg_width[s] = 1 + last valid column over the rows of segment s of an index matrix (0 for an all-pad segment): the
 * reference width recovered from a fixed-width matrix whose rows hold their valid entries first (radius search output). */
int aprb_pool_seg_widths(const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H, const int32_t* d_seg_off,
                         int S, int32_t* d_seg_width, void* stream);
/* Gradient of aprb_max_pool: d_dx [Ns,C] (zeroed by the call) += d_dy[n,c] at the first neighbour attaining the max. */
int aprb_max_pool_backward(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H, int C,
                           const float* d_dy, float* d_dx, void* stream);
/* out[n,:] = (x ++ 0)[idx[n,0], :] */
int aprb_closest_pool(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int C,
                      float* d_out, void* stream);

/* ---------------------------------------------------------------- K6: InstanceNorm (+residual) (+LeakyReLU) -- */
size_t aprb_instnorm_ws_bytes(int N, int C);
/* y = act( (x - mean_col) * rsqrt(var_col + eps) + (d_residual ? residual : 0) ), statistics per column over all
 * N rows (biased variance), act = LeakyReLU(slope) when slope != 1, identity when slope == 1.
 * norm_residual != 0 additionally standardises the residual with its own column statistics before adding
 * (ResnetBottleneckBlock: unary2 + unary_shortcut, models/blocks.py:672-681). round_tf32 != 0 stores y rounded to
 * TF32 (nearest), so the tcgen05 GEMMs that consume it read it exactly. In-place (d_y == d_x) allowed. */
int aprb_instnorm_lrelu(const float* d_x, int N, int C, float eps, float slope,
                        const float* d_residual, int norm_residual, int round_tf32, float* d_y,
                        void* d_ws, size_t ws_bytes, void* stream);

/* Training path: gradient of y = LeakyReLU_slope(InstanceNorm(x)) w.r.t. x from y itself (LeakyReLU is invertible, slope > 0;
 * slope == 1: no activation), the incoming gradient d_dy and the forward's per-column rstd [C]:
 * dx = rstd * (dz - mean(dz) - xhat * mean(dz * xhat)), dz = dy * act'. Two launches, fixed summation order. */
size_t aprb_instnorm_backward_ws_bytes(int C);
int aprb_instnorm_lrelu_backward(const float* d_y, const float* d_dy, const float* d_rstd, int N, int C, float slope,
                                 float* d_dx, void* d_ws, size_t ws_bytes, void* stream);

/* Segmented form for super-batched pairs: rows [d_seg_off[s], d_seg_off[s+1]) (device int32 [S+1]) are normalised with
 * their own column statistics, i.e. S independent BatchNormBlock calls in two launches. S == 1 with d_seg_off == NULL
 * is the plain form. Needs C % 4 == 0 and 16-byte aligned tensors. */
size_t aprb_instnorm_seg_ws_bytes(int N, int C, int S);
int aprb_instnorm_lrelu_seg(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps, float slope,
                            const float* d_residual, int norm_residual, int round_tf32, float* d_y,
                            void* d_ws, size_t ws_bytes, void* stream);
/* Same, with the statistics pass replaced by the producer's group partials where they exist: d_gstat_x for d_x and
 * d_gstat_res for d_residual (norm_residual != 0), each as written by aprb_linear_tf32_stats / aprb_kpconv_forward_stats
 * for exactly that tensor, or NULL to scan the tensor. Same result up to fp32 reassociation of the variance sums. */
int aprb_instnorm_lrelu_seg_pre(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps, float slope,
                                const float* d_residual, int norm_residual, int round_tf32, float* d_y,
                                const float* d_gstat_x, const float* d_gstat_res, void* d_ws, size_t ws_bytes, void* stream);
/* fp16 activation storage (aprb_kfe_forward keeps every normalised activation in fp16: the values are TF32-rounded, so
 * nothing is lost, and each such tensor costs half the HBM bytes to write, gather and feed to the tensor cores):
 * the same normalisation with an optional fp16 residual (a plain residual only) and an optional fp16 output. d_x is
 * always the fp32 output of a GEMM. */
int aprb_instnorm_lrelu_seg_f16(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps, float slope,
                                const void* d_residual, int residual_is_f16, int norm_residual, int round_tf32,
                                void* d_y, int out_is_f16, const float* d_gstat_x, const float* d_gstat_res,
                                void* d_ws, size_t ws_bytes, void* stream);
/* max_pool on fp16 features (C a multiple of 128, <= 1024; int32 indices). */
int aprb_max_pool_f16(const void* d_x16, const int32_t* d_idx, int ld_idx, int Nq, int Ns, int H, int C,
                      void* d_out16, void* stream);
/* Recompute path for the Linears that end a ResnetBottleneckBlock (unary2 and the shortcut Linear, each followed by
 * BatchNormBlock = InstanceNorm, models/blocks.py:496-510, :669-681): their fp32 output exists only to be standardised,
 * so it is never materialised. Three calls replace Linear -> norm -> (+ shortcut) -> LeakyReLU:
 *  1. aprb_linear_f16_stats_ragged: the product's group statistics (as aprb_linear_f16_stats); d_y receives only the
 *     rows the statistics need from the tensor itself (the ragged last 32-row group and the groups that straddle a
 *     segment boundary), everything else of d_y is left untouched;
 *  2. aprb_instnorm_seg_stats: d_stats[seg][t][mean | rstd][C] for one (d_x2 == NULL) or two tensors;
 *  3. aprb_linear_f16_norm_apply: repeats the contraction (bit-identical accumulators) and stores
 *       y = LeakyReLU((x16 @ W16^T - mean) * rstd + shortcut), rounded to the 10-bit mantissa, fp16 or fp32,
 *     where shortcut = d_res16 (plain fp16 rows), or (d_sc16 @ d_Wsc16^T - mean1) * rstd1 with tensor 1 of d_stats,
 *     or nothing. Cin, Csc % 64 == 0, Cout % 16 == 0. Same values as the three-kernel sequence. */
int aprb_linear_f16_stats_ragged(const void* d_x16, const void* d_W16, int N, int Cin, int Cout, float* d_y,
                                 float* d_gstat, const int32_t* d_seg_off, int S, void* stream);
int aprb_instnorm_seg_stats(const float* d_x, const float* d_x2, int N, int C, const int32_t* d_seg_off, int S, float eps,
                            const float* d_gstat_x, const float* d_gstat_x2, float* d_stats, void* stream);
int aprb_linear_f16_norm_apply(const void* d_x16, const void* d_W16, int N, int Cin, int Cout, const void* d_sc16,
                               const void* d_Wsc16, int Csc, const void* d_res16, const int32_t* d_seg_off, int S,
                               const float* d_stats, float slope, void* d_y, int out_is_f16, void* stream);
/* y (fp32) = x16 @ W16^T with fp16 operands on tcgen05 (Cin % 64 == 0, Cout % 16 == 0) + group statistics of y. */
int aprb_linear_f16_stats(const void* d_x16, const void* d_W16, int N, int Cin, int Cout, float* d_y,
                          float* d_gstat, int* stats_written, void* stream);
int aprb_f32_to_f16(const float* d_in, void* d_out16, size_t n, void* stream);
/* d_seg_off[s] = first stacked row of cloud s * clouds_per_segment; ceil(B / clouds_per_segment) + 1 entries. */
int aprb_segment_offsets(const int32_t* d_lens, int B, int clouds_per_segment, int32_t* d_seg_off, void* stream);

/* ---------------------------------------------------------------- Linear (UnaryBlock.mlp) on tcgen05 TF32 ---- */
/* y[N,Cout] = x[N,Cin] @ W[Cout,Cin]^T (nn.Linear layout, no bias), TF32 operands (round-to-nearest), fp32
 * accumulate in TMEM. Cin % 32 == 0 and Cout % 16 == 0 required, else APRB_ERR_UNSUPPORTED. The workspace (may be
 * NULL) holds split-K partial tiles when the output grid alone cannot fill the SMs; the reduction order is fixed. */
size_t aprb_linear_tf32_ws_bytes(int N, int Cin, int Cout);
int aprb_linear_tf32(const float* d_x, const float* d_W, int N, int Cin, int Cout, float* d_y,
                     void* d_ws, size_t ws_bytes, void* stream);
/* Linear + group statistics of y for the normalisation that follows (see aprb_kpconv_forward_stats). */
int aprb_linear_tf32_stats(const float* d_x, const float* d_W, int N, int Cin, int Cout, float* d_y,
                           float* d_gstat, int* stats_written, void* d_ws, size_t ws_bytes, void* stream);
/* out[i] = in[i] rounded to TF32 (round to nearest, ties away). Used once per weight update for mlp.weight. */
int aprb_round_tf32(const float* d_in, float* d_out, size_t n, void* stream);

/* ---------------------------------------------------------------- whole path: pyramid + KFE encoder ---------- */
/* One call = collate_fn_descriptor (datasets/dataloader.py:72-198: 3 subsamplings + 10 radius searches, cell lists
 * shared between searches) + the encoder loop of KPFCNN.forward (models/architectures.py:149-153) for one stacked
 * batch of clouds, driven from native code on one stream: ~3 us of host time per launch instead of a Python call, the
 * host reads the per-level point counts through events so it never waits for the encoder kernels, and several
 * handles can run concurrently from several host threads (one stream each). */
typedef struct {
    int type;                 /* 0 = SimpleBlock, 1 = ResnetBottleneckBlock                     (blocks.py:539, :596)   */
    int strided;              /* 'strided' in block_name: queries = next level, indices = pools (blocks.py:583-590)      */
    int layer;                /* layer_ind                                                                                */
    int in_dim, out_dim;      /* block dims as passed to block_decider (KPConv runs on out_dim/2 or out_dim/4 channels)   */
    float radius, extent;     /* KPConv radius / KP_extent of this block                                                  */
    const float* kp;          /* [K,3]                                                                                    */
    const float* kp_W;        /* [K,Cin,Cout] raw KPConv weights                                                          */
    const float* kp_Wprep;    /* prepared [Cout, K*Cin] (aprb_kpconv_prepare_weights) or NULL (fp32 CUDA-core path)       */
    const float* unary1_W;    /* [out/4, in]  TF32-rounded (aprb_round_tf32), NULL when the block has nn.Identity         */
    const float* unary2_W;    /* [out, out/4] TF32-rounded                                                                */
    const float* shortcut_W;  /* [out, in]    TF32-rounded, NULL when in_dim == out_dim                                   */
    const void* kp_Wprep16;   /* fp16 prepared [Cout, K*Cin] (aprb_kpconv_prepare_weights_f16) or NULL: when set, KPConv  */
                              /* runs with fp16 operands (mode 3) — its input here is always an InstanceNorm output       */
    const void* kp_Wprep16ck; /* fp16 prepared [Cout, Cin*16] (aprb_kpconv_prepare_weights_f16_ck) or NULL: when set and   */
                              /* the shape allows, the weighting stage runs on tcgen05 (mode 5)                           */
} aprb_kfe_block;

typedef struct {
    int num_layers;           /* pyramid levels (4)                                        configs/train/kitti.yaml:12    */
    int K;                    /* kernel points (15)                                                                       */
    float first_subsampling_dl, conv_radius;
    int limits[8];            /* neighbourhood limit per level (calibrate_neighbors, dataloader.py:200-232)               */
    int build_upsamples;      /* 1 = also run the 3 upsample searches, full [N_l, limit] matrices like the collate; 2 = only   */
                              /* their column 0 ([N_l, 1], the nearest support: all closest_pool reads); 0 = none */
    int in_feats_dim;         /* 1                                                                                         */
    int clouds_per_segment;   /* super-batching: P collated pairs stacked in one call = B = 2P clouds with 2 here, so that  */
                              /* BatchNormBlock keeps its per-pair statistics (blocks.py:459-468); 0 = all B clouds are ONE */
                              /* collate (the reference's own semantics for a single collate_fn_descriptor call)            */
} aprb_kfe_config;

typedef struct aprb_kfe aprb_kfe;

int aprb_kfe_create(const aprb_kfe_config* cfg, const aprb_kfe_block* blocks, int nblocks, aprb_kfe** out);
void aprb_kfe_destroy(aprb_kfe* h);
/* Device arena needed for a stacked batch of at most N points in B clouds. */
size_t aprb_kfe_arena_bytes(const aprb_kfe* h, int N, int B);
/* Same with level l bounded by N * ratio^l points (measured 0.40-0.42 per level; 1 = the unconditional bound above).
 * aprb_kfe_forward returns APRB_ERR_WORKSPACE, without writing past the arena, if a level turns out larger. */
size_t aprb_kfe_arena_bytes_est(const aprb_kfe* h, int N, int B, float ratio);
/* d_pts [N,3] level-0 points (already at first_subsampling_dl), d_lens [B]; d_feats [N,in_feats_dim] or NULL (= ones).
 * Runs on `stream`, returns after the last kernel is ENQUEUED (it waits only for the three point-count read-backs).
 * *out_feats receives a device pointer into the arena: encoder output [*out_rows, *out_cols] fp32. */
int aprb_kfe_forward(aprb_kfe* h, const float* d_pts, const int32_t* d_lens, const float* d_feats, int N, int B,
                     void* d_arena, size_t arena_bytes, const float** out_feats, int* out_rows, int* out_cols,
                     void* stream);
/* Same from HOST buffers (pinned memory recommended): H2D of points/lengths, forward, D2H of the encoder output into
 * h_out (capacity h_out_rows_cap rows of *out_cols floats); synchronises `stream` before returning. */
int aprb_kfe_forward_host(aprb_kfe* h, const float* h_pts, const int32_t* h_lens, int N, int B, void* d_arena,
                          size_t arena_bytes, float* h_out, int h_out_rows_cap, int* out_rows, int* out_cols,
                          void* stream);
/* Asynchronous form of the same call: returns once everything is queued. The final block writes into one of two device
 * buffers owned by the handle and a side stream copies it to h_out when the path is done, so the next call (other
 * buffer, other host buffer) can be queued and run while this output is still crossing PCIe. aprb_kfe_wait_host(h, t)
 * blocks until the copy of the call that returned ticket t has landed. At most two calls in flight per handle. */
int aprb_kfe_forward_host_async(aprb_kfe* h, const float* h_pts, const int32_t* h_lens, int N, int B, void* d_arena,
                                size_t arena_bytes, float* h_out, int h_out_rows_cap, int* out_rows, int* out_cols,
                                int* ticket, void* stream);
/* Host output of aprb_kfe_forward_host_async in fp16 (h_out then holds rows x cols halves): the final activation is
 * rounded to a 10-bit mantissa anyway, so fp16 converts back to the identical fp32 value for |v| >= 2^-14 (absolute
 * error <= 2^-25 below) at half the PCIe bytes. Only in the fp16 activation mode (the default). */
int aprb_kfe_set_host_output_f16(aprb_kfe* h, int on);
int aprb_kfe_wait_host(aprb_kfe* h, int ticket);
/* After a forward: pyramid tensors in the arena. what: 0 = points [n,3] f32, 1 = neighbors, 2 = pools, 3 = upsamples
 * (int32 [n, limit]), 4 = stack lengths [B] i32. Pointers stay valid until the next forward on this handle. */
int aprb_kfe_get(const aprb_kfe* h, int what, int level, const void** d_ptr, int* rows, int* cols);
/* Output of encoder block `block` of the last forward ([rows, cols], fp16 when *is_f16; a pointer into the arena, valid
 * until the next forward): the decoder half of KPFCNN concatenates the inputs of the strided blocks as skip features
 * (models/architectures.py:149-152, :195-197). */
int aprb_kfe_get_block_output(const aprb_kfe* h, int block, const void** d_ptr, int* rows, int* cols, int* is_f16);
/* Test hook: with a tap buffer set, every forward also copies intermediate tensors (device to device, stream ordered)
 * into d_buf: tag 4*b + 0 = output of encoder block b, 4*b + 1 = raw output of its KPConv (fp32), 4*b + 2 = the input of
 * its KPConv (fp16 activation mode only). aprb_kfe_get_tap(h, i) returns the i-th recorded tensor [rows, cols], fp16
 * when *is_f16. The per-block parity tests feed block b's oracle with block b-1's tap. d_buf == NULL: taps off. */
int aprb_kfe_set_tap(aprb_kfe* h, void* d_buf, size_t bytes);
int aprb_kfe_tap_count(const aprb_kfe* h);
int aprb_kfe_get_tap(const aprb_kfe* h, int i, const void** d_ptr, int* rows, int* cols, int* is_f16, int* tag);

#ifdef __cplusplus
}
#endif
#endif /* APRB200_H */
