/*
 * kdfusion_b200 -- C ABI of the B200 (sm_100a) kernels behind the camera+LiDAR
 * distillation training hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * reference (KELVIN-ASU/Lightweight-Multi-Modal-Scene-Understanding-via-
 * Knowledge-Distillation) is pure Python/PyTorch and has no FFI of its own, so
 * each entry point names the reference Python lines whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     nothing synchronises, nothing allocates (callers pass workspaces), so
 *     every call can be captured into a CUDA graph;
 *   - tensors are dense and row-major in the order written, e.g. [B,N,4];
 *   - return value 0 = success, otherwise an error code; kdf_last_error()
 *     gives the message of the last failure on the calling thread;
 *   - dtype codes: KDF_F32 / KDF_BF16 (features, logits, gradients);
 *     points, index math, statistics and loss scalars are always fp32.
 */
#ifndef KDFUSION_B200_H
#define KDFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDF_ABI_VERSION 1

enum { KDF_F32 = 0, KDF_BF16 = 1 };
enum { KDF_REDUCE_MAX = 0, KDF_REDUCE_MEAN = 1 };
enum { KDF_OK = 0, KDF_ERR_ARG = 1, KDF_ERR_CUDA = 2, KDF_ERR_UNSUPPORTED = 3 };

/* ---------------------------------------------------------------- library */
int         kdf_abi_version(void);
const char *kdf_last_error(void);
/* number of SMs / compute capability of the current device (host outputs) */
int         kdf_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------- (1) LiDAR -> BEV projection
 * Replaces SpatialLiDAREncoder.points_to_bev_coords + the index/scatter half
 * of forward_vectorized (reference src/models/lidar_encoder.py:42-55, 69-99).
 *
 *   xn = (x - x0) / xspan ; yn = (y - y0) / yspan         one IEEE fp32 rounding per op
 *   valid = 0 <= xn <= 1 && 0 <= yn <= 1                   closed range, NaN -> invalid
 *   col = trunc(xn * (W-1)), row = trunc(yn * (H-1))       clamped to the grid
 *   cell = row*W + col   (-1 when invalid)
 *
 * x0/xspan/y0/yspan are the fp32 values of x_range[0], x_range[1]-x_range[0], ...
 * formed on the host exactly as the reference's buffers do (int64 difference
 * first when the range is integral).
 */

/* Index / occupancy only (bit-exact outputs; also the rasterisation primitive).
 *   points    f32 [B,N,point_stride]  (x,y first; point_stride >= 2 floats, 4 for (x,y,z,i))
 *   cell      i32 [B,N]   out
 *   rank      i32 [B,N]   out, may be NULL: arrival rank of the point inside its cell
 *   count     i32 [B,H*W] out (zeroed by the call): points per cell
 */
int kdf_bev_index(const float *points, int B, int64_t N, int point_stride,
                  float x0, float xspan, float y0, float yspan, int H, int W,
                  int32_t *cell, int32_t *rank, int32_t *count, void *stream);

/* Range-view (spherical) projection: the same index / sort / per-cell reduce machinery over the cells of a range image
 * instead of the bird's-eye-view grid (BASELINE north star: "projection into the range/BEV grid").  The reference has no
 * range-view code, so this half is parity-unpinned; the convention is the one of range-image LiDAR networks:
 *   depth = |(x,y,z)| ; yaw = -atan2(y, x) ; pitch = asin(z / depth)
 *   col = floor(0.5 * (yaw/pi + 1) * W) ; row = floor((1 - (pitch - fov_down) / (fov_up - fov_down)) * H), clamped
 *   invalid (-1): depth == 0, non-finite coordinates, pitch outside [fov_down, fov_up]   (angles in radians)
 * kdf_range_index = kdf_bev_index, kdf_range_project_fwd = kdf_bev_project_fwd with that cell function (fp32 device
 * atan2f / asinf: cell ids agree with a float64 evaluation except for points within rounding of a cell boundary);
 * gradients go through kdf_bev_project_bwd unchanged (it only sees cells). */
int kdf_range_index(const float *points, int B, int64_t N, int point_stride, float fov_up, float fov_down, int H, int W,
                    int32_t *cell, int32_t *count, void *stream);
int kdf_range_project_fwd(const float *points, int point_stride, const void *feats, int dtype, int B, int64_t N, int C,
                          float fov_up, float fov_down, int H, int W, int reduce,
                          void *grid, int32_t *count, int32_t *cell, int32_t *ties, int32_t *order, int32_t *offsets,
                          void *workspace, size_t workspace_bytes, void *stream);

/* BEV label rasterisation (reference src/data_loading/pandaset_dataset.py:23-45, rasterize_bev):
 *   inside = x_min <= x <= x_max && y_min <= y <= y_max     on the RAW fp32 coordinates (:33)
 *   col = clip(trunc((x - x_min) / xspan * (W-1)), 0, W-1), row likewise (:39-40; xspan = x_max - x_min
 *         formed on the host like the reference's Python scalars)
 *   a cell keeps the FIRST non-zero label that lands in it, in point order (:42-44)
 *     = the label of the lowest-index non-zero-labelled point of the cell (an integer min-reduction).
 *   points   f32 [B,N,point_stride]     labels i64 [B,N]
 *   first_ws i32 [B,H*W] scratch        out i64 [B,H,W] (0 where no labelled point landed)
 */
int kdf_bev_rasterize(const float *points, int point_stride, const int64_t *labels, int B, int64_t N,
                      float x_min, float x_max, float y_min, float y_max, float xspan, float yspan,
                      int H, int W, int32_t *first_ws, int64_t *out, void *stream);

/* bytes of scratch kdf_bev_project_fwd needs */
size_t kdf_bev_workspace_bytes(int B, int64_t N, int H, int W);

/* Full projection: per-cell channel-wise max (scatter_reduce_ amax,
 * include_self=False into zeros, lidar_encoder.py:85-96) or mean.
 *   feats     dtype [B,N,C] point-major            (C % 4 == 0)
 *   grid      dtype [B,H,W,C] out  -- the NHWC memory of the [B,C,H,W] view the
 *                                     reference returns (lidar_encoder.py:99)
 *   count     i32 [B,H*W] out
 *   cell      i32 [B,N]   out
 *   ties      i32 [B,H*W,C] out, max only, may be NULL: sources equal to the max.  Pass NULL when
 *                          one row is 8/16/32 16-byte lanes (C = 64/128/256 bf16, 32/64/128 fp32): the
 *                          forward is then a pure packed maximum and kdf_bev_project_bwd counts the ties
 *                          itself from feats/grid (faster both ways); other C need it for the backward
 *   order     i32 [B,N]   out, may be NULL (then taken from workspace): point ids
 *                          grouped by cell (counting sort), offsets in `offsets`
 *   offsets   i32 [B,H*W+1] out, may be NULL (then taken from workspace)
 */
int kdf_bev_project_fwd(const float *points, int point_stride, const void *feats, int dtype,
                        int B, int64_t N, int C,
                        float x0, float xspan, float y0, float yspan, int H, int W, int reduce,
                        void *grid, int32_t *count, int32_t *cell, int32_t *ties,
                        int32_t *order, int32_t *offsets,
                        void *workspace, size_t workspace_bytes, void *stream);

/* The sort stage alone (index -> scan -> fill): cell ids, occupancy and the cell ordering, to be
 * shared by several reductions over the same sweep (teacher and student project the same points). */
int kdf_bev_build_order(const float *points, int point_stride, int B, int64_t N,
                        float x0, float xspan, float y0, float yspan, int H, int W,
                        int32_t *count, int32_t *cell, int32_t *order, int32_t *offsets,
                        void *workspace, size_t workspace_bytes, void *stream);

/* The sort stage with the POINTS THEMSELVES written in cell order (the point MLP then runs over cell-sorted
 * rows and the projection works on contiguous row segments): sorted_points f32 [B,N,4] -- frame b's rows
 * [offsets[b,c], offsets[b,c+1]) are the points of cell c, the points outside the grid follow from offsets[b,HW] on --,
 * cell_sorted i32 [B,N] = the GLOBAL cell id b*H*W + c of every sorted row (-1 outside), order (nullable) i32 [B,N] = the
 * permutation (sorted row -> point id).  count / cell / offsets as kdf_bev_build_order (cell stays in the caller's point
 * order: it is the reference's flat index, lidar_encoder.py:74-82).  Points must be dense (x, y, z, i), 16-byte aligned. */
int kdf_bev_build_sorted(const float *points, int B, int64_t N, float x0, float xspan, float y0, float yspan, int H, int W,
                         int32_t *count, int32_t *cell, int32_t *offsets, float *sorted_points, int32_t *cell_sorted,
                         int32_t *order, void *workspace, size_t workspace_bytes, void *stream);

/* The reduce stage alone, over a cell ordering produced earlier (by
 * kdf_bev_project_fwd with order/offsets supplied, e.g. to share one sort between
 * the teacher's and the student's projection of the same sweep, or to time the
 * dominant kernel in isolation).  Same outputs as kdf_bev_project_fwd's grid/ties. */
int kdf_bev_reduce(const void *feats, int dtype, const int32_t *order, const int32_t *offsets,
                   int B, int64_t N, int C, int H, int W, int reduce,
                   void *grid, int32_t *ties, void *stream);

/* Gradient of the projection w.r.t. feats.  max: the cell gradient is split
 * evenly among the sources equal to the max (ATen ScatterReduceBackward), with
 * ATen's quirk that a max of exactly 0.0 counts the zero-initialised output as
 * one more tie.  mean: grad / count.  Points outside the grid get 0.
 *   grad_grid dtype [B,H*W,C]; grid/ties from the forward (max only; ties may be NULL when the forward
 *   was run without them, see kdf_bev_project_fwd -- needs order/offsets);
 *   order/offsets: the forward's cell ordering (both or neither); with it the gradient is
 *   computed cell-major (per-cell rows read once), without it point-major;
 *   grad_feats dtype [B,N,C] out (every row written).
 */
int kdf_bev_project_bwd(const void *grad_grid, const void *feats, const void *grid,
                        const int32_t *ties, const int32_t *count, const int32_t *cell,
                        const int32_t *order, const int32_t *offsets,
                        int dtype, int B, int64_t N, int C, int H, int W, int reduce,
                        void *grad_feats, void *stream);

/* ---------------------------------------------------------------- BatchNorm (+activation) over rows
 * Every BatchNorm of the reference (BatchNorm1d in the point MLP, src/models/lidar_encoder.py:25-35;
 * BatchNorm2d in camera_encoder.py:19-41 and fusion_module.py:11-32) applied to channels-last
 * data seen as rows x [M,C] (M = B*N points or B*H*W pixels).  act: 0 none, 1 ReLU, 2 ReLU6.
 *
 * kdf_rowbn_stats    training statistics in one pass: mean / invstd (biased variance, eps inside the
 *                    sqrt), the folded scale = gamma*invstd and shift = beta - mean*scale, and the
 *                    nn.BatchNorm running-stat update (momentum, unbiased variance) when
 *                    running_mean/var are given.  gamma/beta may be NULL (1 / 0).  pre_bias (may be
 *                    NULL) is a per-channel bias the producing layer would have added to x: batch
 *                    normalisation cancels it exactly, so it only shifts the running mean.
 * kdf_rowbn_apply_fwd  y = act(x*scale + shift) [+ residual]      (residual may be NULL)
 * kdf_rowbn_bwd      d x, d gamma, d beta from grad_out: dy = g*act'(x*scale+shift);
 *                    batch_stats != 0 chains through the batch mean / variance (training mode),
 *                    batch_stats == 0 is eval mode (running statistics are constants).
 *                    Inputs of up to 250 MB (grad_out + x) run as ONE cooperative kernel (column sums, grid
 *                    barrier, coefficients, apply: the second read is an L2 hit); larger ones as two kernels.
 * Workspaces: kdf_rowbn_workspace_bytes(C) for stats / fwd_train, kdf_rowbn_bwd_workspace_bytes(C) for bwd
 * (a ticket + 8 fp64 accumulator sets [2][C], zeroed by each call; the column sums are spread over the sets because
 * same-address atomics serialise in L2).
 */
size_t kdf_rowbn_workspace_bytes(int C);
size_t kdf_rowbn_bwd_workspace_bytes(int C);
int kdf_rowbn_stats(const void *x, int dtype, int64_t M, int C, const float *gamma, const float *beta,
                    const float *pre_bias, float eps, float momentum, float *running_mean, float *running_var,
                    float *mean, float *invstd, float *scale, float *shift, void *workspace, void *stream);
int kdf_rowbn_apply_fwd(const void *x, const void *residual, int dtype, int64_t M, int C,
                        const float *scale, const float *shift, int act, void *y, void *stream);
int kdf_rowbn_bwd(const void *grad_out, const void *x, int dtype, int64_t M, int C,
                  const float *scale, const float *shift, const float *mean, const float *invstd,
                  int act, int batch_stats, void *grad_x, float *dgamma, float *dbeta,
                  void *workspace, void *stream);

/* Projection of a layer whose BatchNorm+ReLU has not been applied yet (fused point-MLP path): the feature the
 * reference scatters (lidar_encoder.py:32-34, 85-96) is a3 = bf16(relu(z*scale + shift)), formed on the fly from
 * the stored pre-BatchNorm rows z (bf16 [B,N,C]).  BatchNorm-apply, ReLU and rounding are monotonic, so
 * kdf_bev_reduce_affine reduces the raw rows to the per-cell EXTREME of z (max where scale >= 0, min where
 * scale < 0; grid_z bf16 [B,HW,C], optional) and stores its activation (grid bf16 [B,HW,C] = the per-cell max of
 * a3; empty cells 0).  kdf_bev_bwd_affine = gradient w.r.t. the BatchNorm OUTPUT y3 (ReLU folded in): the rows
 * whose z equals the cell's extreme share the cell's gradient evenly (k rows -> g/k each, bf16) when the
 * activation is positive, every other row of a point inside the grid gets zeros, and so do the rows of points
 * outside when `cell` is given (cell = NULL leaves those rows unwritten: the consumer masks them, see
 * kdf_mlp_layer_bwd's row_cell); sums f64 [2,C] = (sum dy, sum dy*z) for BatchNorm's backward (zeroed by the call).  kdf_point_moments: sums of (x,y,z,i) and their 10 pairwise products over all
 * points (f64 [14], zeroed by the call) -- the first MLP layer is linear in the point, so its BatchNorm
 * statistics follow from these. */
int kdf_bev_reduce_affine(const void *z_bf16, const float *scale, const float *shift,
                          const int32_t *order, const int32_t *offsets, int B, int64_t N, int C, int H, int W,
                          void *grid_bf16, void *grid_z_bf16, void *stream);
int kdf_bev_bwd_affine(const void *grad_grid_bf16, const void *z_bf16, const void *grid_bf16, const void *grid_z_bf16,
                       const int32_t *order, const int32_t *offsets, const int32_t *cell,
                       int B, int64_t N, int C, int H, int W, void *dy_bf16, double *sums, void *stream);
int kdf_point_moments(const float *points, int64_t M, double *out14, void *stream);

/* Cell-sorted rows (kdf_bev_build_sorted; kdf_bev_reduce_affine then takes order = NULL): the projection backward
 * without gradient rows.  share bf16 [B*H*W, C] = what every row at its cell's extreme receives (g/k, bf16; rows of empty
 * cells are not written), bits u8 [B*N, C/8] = per row and 8 channels one byte, bit q set when channel 8g+q of the row sits
 * at the extreme (rows of points outside are not written), sums as kdf_bev_bwd_affine.  kdf_mlp_layer_bwd_share is
 * kdf_mlp_layer_bwd (mode 1) with dy[row][c] = bit ? share[cell_sorted[row]][c] : 0 formed in its prologue: the
 * [B*N, C] gradient tensor of lidar_encoder.py:85-96's backward is never written or read. */
int kdf_bev_bwd_share(const void *grad_grid_bf16, const void *z_bf16, const void *grid_bf16, const void *grid_z_bf16,
                      const int32_t *offsets, int B, int64_t N, int C, int H, int W, void *share_bf16, void *bits_u8,
                      double *sums, void *stream);
int kdf_mlp_layer_bwd_share(const int32_t *row_cell, const void *share, const void *bits, const void *z,
                            const float *gs, const float *ga, const float *gb, const void *z_prev, int64_t M,
                            const float *pro_a, const float *pro_b, const void *W_bf16, void *dy_prev, double *sums, float *dW,
                            void *stream);

/* ---------------------------------------------------------------- fused point-MLP layers (tcgen05)
 * One tensor-core layer of the point MLP (src/models/lidar_encoder.py:25-35) as ONE kernel that keeps
 * only the pre-BatchNorm outputs in HBM:  z_out[M,128] (bf16) = prologue(input) . W^T, plus the fp64
 * column sums / sums of squares of the stored values (this layer's BatchNorm statistics).
 *   mode 0: input = raw points f32 [M,4]; the prologue recomputes the whole first layer
 *           a1[c] = relu(q[c,:].(x,y,z,i) + r[c]), q f32 [64,4] = scale1*W1, r f32 [64]; Kin = 64
 *   mode 1: input = previous z bf16 [M,Kin]; prologue relu(z*scale + shift); pro_a = scale, pro_b = shift; Kin = 128
 *   W_bf16 [Nout=128, Kin] row-major bf16;  stats f64 [2,128] (zeroed by the call).
 * kdf_bn_finalize turns such sums into mean / invstd / folded scale, shift and updates running stats.
 */
int kdf_mlp_layer_fwd(int mode, const void *input, int64_t M, const float *pro_a, const float *pro_b,
                      const void *W_bf16, int Kin, int Nout, void *z_out, double *stats, void *stream);
/* All three layers in ONE kernel when the BatchNorms use running statistics (eval mode -- the frozen teacher of the
 * distillation step): points f32 [M,4] -> z3 bf16 [M,128] (pre-BatchNorm-3 rows, what kdf_bev_reduce_affine takes).
 * q/r: folded first layer as in mode 0 above; scale2/shift2 f32 [128]: BatchNorm-2 with running statistics;
 * W2 bf16 [128,64], W3 bf16 [128,128].  z2 never exists in HBM (16 + 256 B per point instead of 16 + 768).
 * Bit-identical to kdf_mlp_layer_fwd(mode 0) followed by kdf_mlp_layer_fwd(mode 1). */
int kdf_mlp_eval3_fwd(const float *points, int64_t M, const float *q, const float *r, const void *W2_bf16,
                      const float *scale2, const float *shift2, const void *W3_bf16, void *z3_out, void *stream);
int kdf_bn_finalize(const double *stats, int64_t M, int C, const float *gamma, const float *beta, const float *pre_bias,
                    float eps, float momentum, float *running_mean, float *running_var,
                    float *mean, float *invstd, float *scale, float *shift, void *stream);

/* The same layer backwards, ONE kernel (dgrad + wgrad on the tensor cores, BatchNorm backward in the prologue):
 *   dz = gs*dy + ga + gb*z   (dy, z bf16 [M,128]; gs/ga/gb f32 [128]: BatchNorm backward of THIS layer, chained
 *                             through the batch statistics by the caller from the previous kernel's column sums)
 *   dW f32 [128,Kin] = dz^T . a_in            (a_in re-created in the prologue exactly as kdf_mlp_layer_fwd does)
 *   mode 1 (Kin=128, input = z_prev bf16 [M,128], pro_a/pro_b = scale/shift of the previous BatchNorm):
 *           dy_prev bf16 [M,128] = (dz . W) * (a_in > 0);  sums f64 [2,128] = (sum dy_prev, sum dy_prev*z_prev)
 *   mode 0 (Kin=64, input = raw points f32 [M,4], pro_a = q [64,4], pro_b = r [64]): nothing per point is stored;
 *           with dy1 = (dz . W) * (a1 > 0):  sums f64 [5,64] = (sum dy1, sum dy1*x, sum dy1*y, sum dy1*z, sum dy1*i),
 *           from which the caller forms the first layer's BatchNorm / weight gradients (it is linear in the point).
 * sums and dW are zeroed by the call.  row_cell (nullable, i32 [M]): rows with row_cell < 0 are points outside the
 * grid, whose dy is zero by construction -- their dy rows are ignored (kdf_bev_bwd_affine called with cell = NULL
 * leaves them unwritten).  Replaces autograd through Conv1d+BatchNorm1d+ReLU (lidar_encoder.py:25-35). */
int kdf_mlp_layer_bwd(int mode, const void *dy, const void *z, const float *gs, const float *ga, const float *gb,
                      const void *input, int64_t M, const float *pro_a, const float *pro_b, const void *W_bf16,
                      int Kin, void *dy_prev, double *sums, float *dW, const int32_t *row_cell, void *stream);

/* Per-channel algebra between the layer kernels, one tiny launch each (BatchNorm1d of lidar_encoder.py:27,30,33):
 *   kdf_mlp_l1_stats   batch statistics of layer 1 in closed form from the 14 point moments (it is linear in the point):
 *                      q f32 [64,4] = scale1*W1, r f32 [64] = shift1 (the folded first layer of the mode-0 prologues),
 *                      mean / invstd / scale f32 [64]; running statistics advanced when given (b1 joins the mean)
 *   kdf_bn_bwd_coeffs  BatchNorm backward through the batch statistics as coefficients dz = gs*dy + ga + gb*z, and
 *                      dgamma / dbeta, from sums f64 [2,C] = (sum dy, sum dy*z)
 *   kdf_mlp_l1_bwd     layer 1 backward in closed form from kdf_mlp_layer_bwd(mode 0)'s sums f64 [5,64] and the moments:
 *                      dW1 f32 [64,4], dgamma, dbeta f32 [64] */
int kdf_mlp_l1_stats(const double *moments14, int64_t M, const float *W1, const float *b1, const float *gamma,
                     const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                     float *q, float *r, float *mean, float *invstd, float *scale, void *stream);
int kdf_bn_bwd_coeffs(const double *sums, int C, int64_t M, const float *mean, const float *invstd, const float *scale,
                      float *gs, float *ga, float *gb, float *dgamma, float *dbeta, void *stream);
int kdf_mlp_l1_bwd(const double *sums5x64, const double *moments14, int64_t M, const float *W1, const float *mean,
                   const float *invstd, const float *scale, float *dW1, float *dgamma, float *dbeta, void *stream);

/* ---------------------------------------------------------------- pointwise (1x1) convolution layers (tcgen05)
 * Replaces the 1x1 nn.Conv2d + nn.BatchNorm2d + nn.ReLU/ReLU6 groups of the camera branch, the FPN, the fusion
 * projections and the head (reference src/models/camera_encoder.py:19-51, src/models/fusion_module.py:8-34, 51-64,
 * 111-120, 162-173) over channels-last maps, i.e. pixel-major rows:
 *
 *   a   = pro_act(x * pro_scale + pro_shift)          the PREVIOUS BatchNorm + activation, when pro_scale != NULL
 *   z   = a . W^T                                     bf16 operands, fp32 accumulation on the tensor cores
 *   out = z                                           and, when stats != NULL, stats[0][n] += sum_rows bf16(z),
 *                                                     stats[1][n] += sum_rows bf16(z)^2  (this layer's batch statistics)
 *   out = epi_act(z * epi_scale + epi_shift) [+ residual]     when epi_scale != NULL (running-statistics BatchNorm folded)
 *
 *   x        bf16 [M, K]   K a multiple of 64          W   bf16 [N, K]   N a multiple of 32
 *   pro_*    f32 [K]       epi_* f32 [N]               residual bf16 [M, N], nullable (only with epi_scale)
 *   out      bf16 [M, N]   stats f64 [2, N], nullable (zeroed by the call; not together with epi_scale)
 *   act codes: 0 none, 1 ReLU, 2 ReLU6
 */
int kdf_pw_conv_fwd(const void *x, int64_t M, int K, int N, const void *W,
                    const float *pro_scale, const float *pro_shift, int pro_act,
                    const float *epi_scale, const float *epi_shift, int epi_act, const void *residual,
                    void *out, double *stats, void *stream);

/* ---------------------------------------------------------------- (2) camera-LiDAR fusion
 * Inputs are the PRE-BatchNorm outputs of the two 1x1 projection convolutions
 * in pixel-major (NHWC) layout; BatchNorm is applied as y = x*scale + shift with
 * per-channel fp32 scale/shift (batch statistics in training, running statistics
 * in eval), followed by ReLU -- i.e. the tail of the reference's Conv1x1 block
 * (src/models/fusion_module.py:8-17) fused into the fusion itself.
 *
 * weighted  (WeightedFusion, fusion_module.py:107-136, executed inline at :248-253):
 *      cp = relu(cam*sc+sh), lp = relu(lid*sc+sh)
 *      a  = W2 . relu(W1 . [cp;lp] + b1) + b2 ;  w = softmax(a)   (2 logits per pixel)
 *      out = cp*w0 + lp*w1
 *   M pixels, C channels per branch (C % 32 == 0, C <= 256), hidden width = C.
 *   w1 f32 [C,2C], b1 f32 [C], w2 f32 [2,C], b2 f32 [2];  attn f32 [M,2] out.
 */
int kdf_fusion_weighted_fwd(const void *cam_pre, const void *lid_pre, int dtype, int64_t M, int C,
                            const float *cam_scale, const float *cam_shift,
                            const float *lid_scale, const float *lid_shift,
                            const float *w1, const float *b1, const float *w2, const float *b2,
                            void *out, float *attn, void *stream);

/* Backward of the above.  Outputs gradients w.r.t. the pre-BN inputs treating
 * scale/shift as independent inputs (their gradients are returned so autograd
 * can chain through the batch statistics):
 *   grad_cam_pre/grad_lid_pre dtype [M,C];
 *   grad_affine f32 [4,C]  = d/d(cam_scale, cam_shift, lid_scale, lid_shift)   (zeroed by the call)
 *   grad_w1 f32 [C,2C], grad_b1 f32 [C], grad_w2 f32 [2,C], grad_b2 f32 [2]    (zeroed by the call)
 */
int kdf_fusion_weighted_bwd(const void *grad_out, const void *cam_pre, const void *lid_pre,
                            int dtype, int64_t M, int C,
                            const float *cam_scale, const float *cam_shift,
                            const float *lid_scale, const float *lid_shift,
                            const float *w1, const float *b1, const float *w2, const float *b2,
                            const float *attn,
                            void *grad_cam_pre, void *grad_lid_pre, float *grad_affine,
                            float *grad_w1, float *grad_b1, float *grad_w2, float *grad_b2,
                            void *stream);

/* minimal (MinimalFusion, fusion_module.py:94-104; inline at :248-255):
 *      out = relu(cam*sc+sh) + relu(lid*sc+sh)
 * concat  (first half of ConcatenationFusion, fusion_module.py:89-91; inline :243-245):
 *      out[M,2C] = [relu(cam*sc+sh) ; relu(lid*sc+sh)]          (mode = 1)
 */
int kdf_fusion_affine_relu_pair_fwd(const void *cam_pre, const void *lid_pre, int dtype, int64_t M, int C,
                                    const float *cam_scale, const float *cam_shift,
                                    const float *lid_scale, const float *lid_shift,
                                    int mode /*0 = add, 1 = concat*/, void *out, void *stream);
int kdf_fusion_affine_relu_pair_bwd(const void *grad_out, const void *cam_pre, const void *lid_pre,
                                    int dtype, int64_t M, int C,
                                    const float *cam_scale, const float *cam_shift,
                                    const float *lid_scale, const float *lid_shift, int mode,
                                    void *grad_cam_pre, void *grad_lid_pre, float *grad_affine /*[4,C], zeroed*/,
                                    void *stream);

/* ---------------------------------------------------------------- (3) distillation loss
 * One pass over logits and mimic taps producing the loss terms AND their
 * gradients:
 *   ce  = sum_i w[y_i] * (-log softmax(z_s)_i[y_i]) / sum_i w[y_i]      over y_i != ignore
 *         (nn.CrossEntropyLoss(ignore_index=-1, weight=w), reference src/training/trainer.py:55,88)
 *   kl  = T^2 * sum_pixels KL(softmax(z_t/T) || softmax(z_s/T)) / (B*HW)   (not in the reference)
 *   mse = sum_taps mean((s - t)^2)                                          (not in the reference)
 *   loss = (1-alpha)*ce + alpha*kl + beta*mse
 *
 *   s_logits/t_logits dtype_logits [B,K,HW] (NCHW), K <= 8; t_logits NULL -> no KL term
 *   labels i64 [B,HW]; class_w f32 [K] or NULL
 *   taps: up to 2 (s_featX, t_featX dtype_feat, numelX elements; NULL/0 = unused)
 *   d_logits [B,K,HW] out; d_featX out (same dtype as feats)
 *   scalars f32 [8] out: loss, ce, kl, mse, wsum, mse_tap0, mse_tap1, n_valid
 *   grad_scale multiplies every gradient (upstream dL/dloss, 1/world_size, ...)
 *   workspace: kdf_kd_loss_workspace_bytes() bytes
 */
size_t kdf_kd_loss_workspace_bytes(void);
int kdf_kd_loss_fwd_bwd(const void *s_logits, const void *t_logits, const int64_t *labels,
                        const float *class_w, int B, int K, int64_t HW, int dtype_logits,
                        float T, float alpha, float beta, int64_t ignore_index,
                        const void *s_feat0, const void *t_feat0, void *d_feat0, int64_t numel0,
                        const void *s_feat1, const void *t_feat1, void *d_feat1, int64_t numel1,
                        int dtype_feat, float grad_scale,
                        void *d_logits, float *scalars, void *workspace, void *stream);
/* The label histogram (the CE normaliser sum_i w[y_i] needs it) depends on the labels only: it can be taken when the
 * batch arrives -- on a side stream, off the loss's critical path -- with kdf_kd_label_count into the SAME workspace,
 * followed later by kdf_kd_loss_fwd_bwd_counted (identical arguments and results, no histogram phase). */
int kdf_kd_label_count(const int64_t *labels, int B, int K, int64_t HW, int64_t ignore_index, void *workspace, void *stream);
int kdf_kd_loss_fwd_bwd_counted(const void *s_logits, const void *t_logits, const int64_t *labels,
                        const float *class_w, int B, int K, int64_t HW, int dtype_logits,
                        float T, float alpha, float beta, int64_t ignore_index,
                        const void *s_feat0, const void *t_feat0, void *d_feat0, int64_t numel0,
                        const void *s_feat1, const void *t_feat1, void *d_feat1, int64_t numel1,
                        int dtype_feat, float grad_scale,
                        void *d_logits, float *scalars, void *workspace, void *stream);

/* g[m,c] += bc[c]*x[m,c] + ac[c] over rows [M,C] (in place): the batch-statistics part of a BatchNorm backward
 * (nn.BatchNorm2d inside Conv1x1, fusion_module.py:11-15) whose per-row part and column sums came out of the fused
 * fusion backward; bc / ac f32 [C] are formed by the caller from those sums. */
int kdf_rows_axpb(void *g, const void *x, int dtype, int64_t M, int C, const float *bc, const float *ac, void *stream);

/* The classifier of the BEV-resolution head (fusion_module.py:162-173): nn.Conv2d(32, K, 1) with bias over bf16 pixel rows
 * x [M,32] (M = frames x HW), taps rounded to bf16 as the autocast convolution does, fp32 accumulation, written as planar
 * logits bf16 [B,K,HW] (what the loss reads).  K = 1..4.  kdf_cls_conv_bwd: dx bf16 [M,32] (nullable), grad_weight f32
 * [K,32] and grad_bias f32 [K] (nullable) -- both zeroed by the call -- from planar dlogits bf16 [B,K,HW]. */
int kdf_cls_conv_fwd(const void *x_bf16, const float *weight, const float *bias, int64_t M, int Cin, int K, int HW,
                     void *logits_bf16, void *stream);
int kdf_cls_conv_bwd(const void *x_bf16, const void *dlogits_bf16, const float *weight, int64_t M, int Cin, int K, int HW,
                     void *dx_bf16, float *grad_weight, float *grad_bias, void *stream);

/* The camera stem (camera_encoder.py:63-67): Conv2d(3, 32, 3, stride 2, padding 1, bias=False) read straight from the fp32
 * NCHW image [B,3,H,W] (inputs and taps rounded to bf16 as the autocast convolution does, fp32 accumulation) into bf16
 * pixel-major rows out [B,OH,OW,32].  stats (nullable, f64 [2,32], zeroed by the call): column sums of the stored values
 * for the train-mode BatchNorm that follows; post_scale / post_shift (nullable, f32 [32]) + post_act (0 none, 1 ReLU,
 * 2 ReLU6): the folded running-statistics BatchNorm + activation of inference instead.  kdf_stem_conv_bwd_weight: the weight
 * gradient f32 [32,3,3,3] (zeroed by the call) from the gradient rows bf16 [B,OH,OW,32]; the image needs no gradient. */
int kdf_stem_conv_fwd(const float *image, const float *weight, int B, int H, int W,
                      const float *post_scale, const float *post_shift, int post_act, void *out_bf16, double *stats, void *stream);
int kdf_stem_conv_bwd_weight(const float *image, const void *grad_out_bf16, int B, int H, int W, float *grad_weight, void *stream);

/* ---------------------------------------------------------------- depthwise 3x3 convolution
 * nn.Conv2d(C, C, 3, stride, padding=1, groups=C, bias=False) of the inverted-residual blocks
 * (camera_encoder.py:27-33), DWSeparableConv (fusion_module.py:24-27) and the concat fusion (:80-82), over
 * pixel-major maps: in [B,H,W,C], out [B,OH,OW,C] with OH = (H-1)/stride + 1; f32 or bf16 storage, fp32
 * arithmetic; weight f32 [C,9] (= the contiguous [C,1,3,3] parameter).  stride 1 or 2.
 *   kdf_dwconv3x3_fwd        out = conv(in, weight); flip=1 (stride 1 only) uses the taps reversed; stats (nullable,
 *                            f64 [2,C], zeroed by the call) receives the per-channel sum / sum of squares of the
 *                            stored outputs -- the batch statistics of the BatchNorm that follows (kdf_bn_finalize)
 *   kdf_dwconv3x3_bwd_data   grad_in [B,H,W,C] from grad_out [B,OH,OW,C]
 *   kdf_dwconv3x3_bwd_weight grad_weight f32 [C,9] (zeroed by the call) from in and grad_out */
int kdf_dwconv3x3_fwd(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                      int flip, void *out, double *stats, void *stream);
/* The same convolution followed, in the same kernel, by the BatchNorm (running statistics, given as per-channel
 * scale/shift f32 [C]) and activation (0 none, 1 ReLU, 2 ReLU6) that come after it in eval mode
 * (camera_encoder.py:26-33 with the module in eval()): out = act(round(conv(in)) * scale + shift), where round() is the
 * storage rounding the two-kernel sequence would have applied -- bit-identical to kdf_dwconv3x3_fwd + kdf_rowbn_apply_fwd. */
int kdf_dwconv3x3_affine_fwd(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                             const float *post_scale, const float *post_shift, int act, void *out, void *stream);
int kdf_dwconv3x3_bwd_data(const void *grad_out, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                           void *grad_in, void *stream);
int kdf_dwconv3x3_bwd_weight(const void *in, const void *grad_out, int dtype, int B, int H, int W, int C, int stride,
                             float *grad_weight, void *stream);

/* ---------------------------------------------------------------- FPN-lite merge
 * CameraFPNLite.forward (fusion_module.py:51-64): every lateral is resized to the largest map with
 * F.interpolate(mode="bilinear", align_corners=False) (:61-62) and summed (:63).  One pass here:
 *   out[B,H,W,C] = base[B,H,W,C] + bilinear(lo_a[B,h,w,C]) (+ bilinear(lo_b[B,h,w,C]) if non-null)
 * pixel-major rows, f32 or bf16 storage (fp32 arithmetic), ATen's source-index rule for any (h,w)->(H,W).
 * kdf_fpn_up2_bwd is the adjoint of the resize for the exact 2x case (H=2h, W=2w; what the model uses):
 *   grad_lo[B,h,w,C] = resize^T(grad_out[B,2h,2w,C]); the gradient w.r.t. base is grad_out itself. */
int kdf_fpn_merge_fwd(const void *base, const void *lo_a, const void *lo_b, int dtype,
                      int B, int H, int W, int h, int w, int C, void *out, void *stream);
int kdf_fpn_up2_bwd(const void *grad_out, int dtype, int B, int h, int w, int C, void *grad_lo, void *stream);

/* ---------------------------------------------------------------- training-step helpers
 * Confusion matrix of SegmentationMetrics.update (trainer.py:18-26):
 *   conf[t,p] += 1 over pixels with label t != ignore, 0 <= t < K, p = argmax_k logits.
 *   conf i64 [K,K] is ACCUMULATED into (caller zeroes it at reset()).
 */
int kdf_confusion_matrix(const void *logits, const int64_t *labels, int B, int K, int64_t HW,
                         int dtype_logits, int64_t ignore_index, int64_t *conf, void *stream);

/* AdamW over one flat parameter buffer (torch.optim.AdamW single-tensor maths,
 * trainer.py:56,90): decoupled weight decay, bias-corrected moments.
 *   hyper f32 [2] on device = {lr, step}  (step already incremented, >= 1)
 *   param_bf16  bf16 [n] out, nullable: the updated parameters rounded to bf16 (the operand copy the tensor-core
 *               layers read, so that no per-step cast of the weights is needed)
 */
int kdf_adamw_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                   const float *hyper, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, void *param_bf16, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KDFUSION_B200_H */
