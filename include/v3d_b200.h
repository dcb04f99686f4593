/*
 * v3d_b200.h -- C-ABI of the B200-native (sm_100a) vision3d per-frame LiDAR hot path.
 *
 * This is the drop-in boundary. The reference (jhultman/vision3d @ b9a50ee) has no C plugin
 * registry; its boundary is four Python import names. Each entry point
 * below is what a binding for one of those names calls; the reference interface it replaces
 * is cited as file:line relative to the reference tree. INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions (mirroring the in-tree reference ops, SURVEY.md 8b "Conventions"):
 *   - every pointer is a DEVICE pointer unless the name ends in _host; inputs are never
 *     mutated; outputs and workspaces are caller-allocated (no allocation inside the library);
 *   - `stream` is a cudaStream_t passed as void* (the reference launches on the current
 *     torch stream: box_iou_rotated_cuda.cu:71,102); calls are asynchronous and never
 *     synchronise the stream or the device, so the whole path can be CUDA-graph captured;
 *   - data-dependent sizes (kept boxes, voxels, active sites) are written to DEVICE counters
 *     and the outputs are sized by a static capacity, instead of the reference's blocking
 *     D2H copy (nms_rotated_cuda.cu:106);
 *   - return value: V3D_OK (0) or a negative v3d error; on error nothing is launched.
 *     The reference raises RuntimeError through AT_ASSERTM/AT_ERROR
 *     (box_iou_rotated_cuda.cu:69-70,85-87); the Python bindings do the same from the code.
 *   - fp32 data, int32 indices, row-major contiguous, exactly as the reference ops read them
 *     through raw data_ptr (box_iou_rotated_cuda.cu:81-82).
 *   - there is NO CPU path: every function needs a CUDA device of compute capability 10.x.
 */
#ifndef V3D_B200_H_
#define V3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define V3D_API
#else
#define V3D_API __attribute__((visibility("default")))
#endif

typedef void* v3d_stream_t; /* cudaStream_t */

enum {
  V3D_OK = 0,
  V3D_ERR_INVALID_ARGUMENT = -1,  /* bad size / null pointer / unsupported channel count */
  V3D_ERR_WORKSPACE_TOO_SMALL = -2,
  V3D_ERR_CUDA = -3,              /* a CUDA runtime call or launch failed (see v3d_last_cuda_error) */
  V3D_ERR_UNSUPPORTED_DEVICE = -4 /* not an sm_100 class device */
};

V3D_API int v3d_abi_version(void);
V3D_API const char* v3d_status_string(int status);
V3D_API const char* v3d_last_cuda_error(void);
/* cuda_version.cu / vision.cpp:21-32 get_cuda_version(): CUDART_VERSION the library was built with */
V3D_API int v3d_cudart_version(void);
/* 0 if the current device can run this library (compute capability 10.x), else an error */
V3D_API int v3d_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * a13  box_iou_rotated(boxes1[M,5], boxes2[N,5]) -> ious[M,N]
 * replaces vision3d._C.box_iou_rotated: csrc/vision.cpp:63, box_iou_rotated.h:20-32,
 * box_iou_rotated_cuda.cu:65-121. Boxes are (x_ctr, y_ctr, w, h, angle_degrees).
 * Arithmetic = single_box_iou_rotated as nvcc compiles it (box_iou_rotated_utils.h:313-340,
 * exchange-sort hull :197-214) evaluated WITHOUT fused multiply-add, i.e. bit-identical to
 * oracle variant 1 / the header compiled on the host.
 * ------------------------------------------------------------------------------------------- */
V3D_API int v3d_box_iou_rotated(const float* boxes1, int M, const float* boxes2, int N, float* ious,
                                v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a12  nms_rotated(dets[N,5], scores[N], thr) -> keep (indices into dets, descending score)
 * replaces vision3d._C.nms_rotated: csrc/vision.cpp:64, nms_rotated.h:22-36,
 * nms_rotated_cuda.cu:74-134 (bit set when IoU > thr, greedy scan over the sorted list).
 * Everything stays on the device: score ranking (stable: ties -> lower index first), upper-
 * triangular 64x64 mask tiles with an exact disjointness pre-test, and the greedy scan.
 * keep[0..*num_keep) is valid, keep must hold N int64; num_keep is one device int32.
 * ------------------------------------------------------------------------------------------- */
V3D_API size_t v3d_nms_rotated_workspace_bytes(int N);
V3D_API int v3d_nms_rotated(const float* dets, const float* scores, int N, float iou_threshold,
                            int64_t* keep, int* num_keep, void* workspace, size_t workspace_bytes,
                            v3d_stream_t stream);
/* Same result for inputs that are consecutive groups of `group_size` (<= 128) boxes which cannot overlap across
 * groups -- what batched_nms_rotated's per-group coordinate offsets (ops/iou_nms.py:121-132) build before it
 * calls nms_rotated: cross-group IoU is exactly 0, so the greedy NMS decomposes into one per group (same
 * comparator, same IoU arithmetic on the same offset coordinates). One CTA per group instead of an N x N mask.
 * The caller guarantees the disjointness; same workspace as v3d_nms_rotated. */
V3D_API int v3d_nms_rotated_grouped(const float* dets, const float* scores, int N, int group_size,
                                    float iou_threshold, int64_t* keep, int* num_keep, void* workspace,
                                    size_t workspace_bytes, v3d_stream_t stream);

/* Training-side consumer of a13: ProposalTargetAssigner.match_class_i (core/proposal_targets.py:53-60)
 * = box_iou_rotated(gt[M,5], anchors[N,5]) + Matcher.__call__ (ops/matcher.py:86-107) without materialising the M x N
 * matrix: matches[a] = first gt index of maximal IoU, labels[a] = label of the stratum [low, high) the maximum falls in
 * (strata evaluated in order, default label 1), matched_vals[a] (optional) = that maximum. M in 1..256 (M == 0 is the
 * reference's "no gt" shortcut, handled by the binding), strata are HOST arrays of n_strata <= 8 entries. */
V3D_API int v3d_match_anchors(const float* gt_boxes, int M, const float* anchors, int N, int n_strata,
                              const float* low_host, const float* high_host, const int* label_host,
                              int64_t* matches, signed char* labels, float* matched_vals, v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a1 (+a2)  point -> voxel for a whole batch in one call
 * replaces spconv.utils.VoxelGenerator(...).generate(points) per frame and the batch-index
 * prefix + concatenate of Preprocessor.generate_batch_voxels (core/preprocess.py:17-33), and
 * optionally VoxelFeatureExtractor.forward (detector/layers.py:10-17) as an epilogue.
 *   points        (total_points, C) f32, frames concatenated; first 3 columns are x,y,z
 *   frame_offsets (B+1) int32 device: frame b owns rows [off[b], off[b+1])
 *   max_frame_points  HOST upper bound on any frame's point count (<= total_points); sizes the grid
 *   points_capacity   the capacity the workspace was sized/initialised with (>= total_points)
 *   range_min[3], voxel_size[3] fp32 (xyz) and grid[3] cells per axis (xyz) are HOST values
 *   cap_policy    0 = stop the frame at the first point that would open voxel #max_voxels
 *                 (spconv v1.0/1.1), 1 = skip only such points (spconv >= 1.2)
 * outputs (capacity B*max_voxels rows; frames packed back to back like np.concatenate):
 *   voxels     (rows, max_pts, C) f32 zero padded, arrival order inside a voxel
 *   coords     (rows, 4) int32 (b, z, y, x); voxel order = first appearance inside the frame
 *   num_points (rows) int32
 *   voxel_offsets (B+1) int32 device: frame b owns voxel rows [vo[b], vo[b+1])
 *   mean       (rows, C) f32 = sum over slots / num_points, or NULL to skip (a2)
 * The workspace is persistent: initialise once with v3d_voxelize_workspace_init and pass the
 * same buffer to every call (it carries a device-side 32-bit call counter / epoch, so no per-call
 * clearing is needed and CUDA-graph replays stay correct); re-initialise after 2^32-1 calls.
 * ------------------------------------------------------------------------------------------- */
V3D_API size_t v3d_voxelize_workspace_bytes(int total_points_capacity, int B);
V3D_API int v3d_voxelize_workspace_init(void* workspace, size_t workspace_bytes,
                                        int total_points_capacity, int B, v3d_stream_t stream);
V3D_API int v3d_voxelize_batch(const float* points, int total_points, int max_frame_points, int C,
                               const int* frame_offsets, int B, const float* range_min_host,
                               const float* voxel_size_host,
                               const int* grid_host, int max_pts, int max_voxels, int cap_policy,
                               float* voxels, int* coords, int* num_points, int* voxel_offsets,
                               float* mean, void* workspace, size_t workspace_bytes,
                               int points_capacity, v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a4  rule book (spconv get_indice_pairs inside SubMConv3d / SparseConv3d.forward; layers
 * constructed at detector/sparse_cnn.py:15-30,151-175).
 * The rule table is output-stationary: nbr[kk * nbr_stride + o] = input row feeding output row
 * o through kernel offset kk = (kz*k1 + ky)*k2 + kx, or -1. Correlation convention of
 * torch.nn.functional.conv3d: in_pos = out_pos*stride - pad + k*dilation.
 * Active-site counts live on the device (n_in / n_out are device int32) so no host sync.
 *
 * A "site table" is a device hash of the active sites of one resolution level; it is built once
 * per level and shared by every SubM layer with the same indice_key and by the strided conv
 * that leaves the level.
 * ------------------------------------------------------------------------------------------- */
V3D_API size_t v3d_site_table_bytes(int capacity_rows);
V3D_API int v3d_site_table_init(void* table, size_t table_bytes, int capacity_rows, v3d_stream_t stream);
/* indices (cap,4) int32 b,z,y,x ; n_rows device int32 ; shape[3] (z,y,x) host */
V3D_API int v3d_site_table_build(void* table, const int* indices, const int* n_rows, int capacity_rows,
                                 const int* shape_host, v3d_stream_t stream);
/* SubM: outputs = inputs. nbr (KV, nbr_stride). */
V3D_API int v3d_rulebook_subm(const void* table, const int* indices, const int* n_rows,
                              int capacity_rows, const int* shape_host, const int* ksize_host,
                              const int* dilation_host, int* nbr, int nbr_stride, v3d_stream_t stream);
/* Strided conv: discovers the output sites (ascending flat (b,z,y,x) order), writes
 * out_indices (out_capacity,4), n_out (device) and nbr (KV, nbr_stride). in_table is the site
 * table of the INPUT level. workspace from v3d_rulebook_conv_workspace_bytes. */
V3D_API size_t v3d_rulebook_conv_workspace_bytes(int B, const int* out_shape_host, int capacity_rows,
                                                 int kernel_volume);
V3D_API void v3d_conv_out_shape(const int* shape, const int* ksize, const int* stride, const int* pad,
                                const int* dilation, int* out_shape);
V3D_API int v3d_rulebook_conv(const void* in_table, const int* indices, const int* n_rows,
                              int capacity_rows, int B, const int* shape_host, const int* ksize_host,
                              const int* stride_host, const int* pad_host, const int* dilation_host,
                              int* out_indices, int* n_out, int out_capacity, int* nbr, int nbr_stride,
                              void* workspace, size_t workspace_bytes, v3d_stream_t stream);

/* Rank-indexed variants: a level PRODUCED by v3d_rulebook_conv has its rows in ascending flat order, and
 * that call's workspace (bitmap + two-level popcount prefix) is a complete site index of the level
 * (row = number of active cells before the cell). These variants use it instead of a hash table:
 * `level_index` / `in_level_index` = the workspace pointer of the v3d_rulebook_conv call that produced the
 * level, `index_capacity` = the out_capacity passed to that call, shape = that level's shape. */
V3D_API int v3d_rulebook_subm_ranked(const void* level_index, int B, int index_capacity, const int* indices,
                                     const int* n_rows, int capacity_rows, const int* shape_host,
                                     const int* ksize_host, const int* dilation_host, int* nbr, int nbr_stride,
                                     v3d_stream_t stream);
V3D_API int v3d_rulebook_conv_ranked(const void* in_level_index, int in_index_capacity, const int* indices,
                                     const int* n_rows, int capacity_rows, int B, const int* shape_host,
                                     const int* ksize_host, const int* stride_host, const int* pad_host,
                                     const int* dilation_host, int* out_indices, int* n_out, int out_capacity,
                                     int* nbr, int nbr_stride, void* workspace, size_t workspace_bytes,
                                     v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a5 + a6  sparse convolution forward with the eval-mode BatchNorm1d + ReLU that
 * spconv.SparseSequential applies next folded into the epilogue
 * (SubMConvFunction / SparseConvFunction.forward; detector/sparse_cnn.py:15-30).
 *   feat (n_in rows, Cin) f32, weight (KV, Cin, Cout) f32 [spconv layout (k0,k1,k2,Cin,Cout)],
 *   out (out rows, Cout) f32 = relu?(scale * sum_kk feat[nbr[kk][o]] @ W[kk] + shift).
 *   scale/shift may be NULL (identity); n_out is a device int32.
 * ------------------------------------------------------------------------------------------- */
V3D_API int v3d_sparse_conv_fwd(const float* feat, const float* weight, const int* nbr, int nbr_stride,
                                const int* n_out, int out_capacity, int kernel_volume, int Cin,
                                int Cout, const float* scale, const float* shift, int relu, float* out,
                                v3d_stream_t stream);

/* Tensor-core path of the same op (tcgen05.mma kind::f16 on a 3-term bf16 split, fp32 accumulators in
 * TMEM; a value x is carried as h1 = bf16(x), h2 = bf16(x - h1) and a product as h1*g1 + h1*g2 + h2*g1:
 * relative error of a product <= ~3*2^-18, measured 4e-6 (Frobenius) per layer against fp64 -- inside the
 * 1e-4 of the contract; use v3d_sparse_conv_fwd when exact fp32 products are required).
 * Features travel between layers as "packed" rows [h1(0..C-1) | h2(0..C-1)] of bf16 (4*C bytes per row, the
 * fp32 footprint): v3d_feature_pack converts fp32 rows, and the conv writes fp32 rows (`out`), packed rows
 * (`out_packed`) or both (either may be NULL, not both). Weights are prepared once per layer into the
 * shared-memory image the kernel streams with 1-D TMA:
 *   v3d_sparse_conv_prepared_bytes returns 0 for shapes only the exact-fp32 path supports (Cin < 16 ...).
 * Supported: kernel_volume <= 27, Cin and Cout in {16, 32, 64}. */
V3D_API int v3d_sparse_conv_tc_variant(void); /* diagnostic: fetch scheme + 8 * cg + 16 * spin of the tcgen05 kernel
                                                 this process runs (V3D_TC_FETCH / V3D_TC_CG / V3D_TC_WAIT or the
                                                 built-in defaults, csrc/sparse_conv_tc.cu); no reference counterpart */
V3D_API size_t v3d_sparse_conv_prepared_bytes(int kernel_volume, int Cin, int Cout);
V3D_API int v3d_sparse_conv_prepare(const float* weight, int kernel_volume, int Cin, int Cout, void* prepared,
                                    size_t prepared_bytes, v3d_stream_t stream);
/* feat (rows, C_src) f32 -> packed (rows, 2*C) bf16, channels C_src..C-1 zero (C_src <= C, C_src % 4 == 0): a
 * narrow input (the 4-channel voxel means) is padded to the 16 channels the tensor-core path needs */
V3D_API int v3d_feature_pack(const float* feat, const int* n_rows, int capacity_rows, int C_src, int C, void* packed,
                             v3d_stream_t stream);
V3D_API int v3d_sparse_conv_fwd_tc(const void* feat_packed, const void* prepared, const int* nbr, int nbr_stride,
                                   const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                   const float* scale, const float* shift, int relu, float* out,
                                   void* out_packed, v3d_stream_t stream);

/* SURVEY 8f-1 -- backward of the sparse convolution (SparseConvFunction / SubMConvFunction.backward of spconv v1.x, the
 * autograd Functions behind the layers of detector/sparse_cnn.py:15-30; needed by train.py:57-72), exact fp32:
 *   dX : v3d_sparse_conv_fwd(grad_out, W^T per offset (KV, Cout, Cin), inv, ...) on the inverted rule table
 *        inv[k][i] = o  <=>  nbr[k][o] = i  (v3d_rulebook_invert; for SubM rule books inv[k] = nbr[KV-1-k], no build);
 *   dW : v3d_sparse_conv_bwd_weight: grad_weight[k] = sum_o feat[nbr[k][o]]^T grad_out[o] (zeroed, then accumulated
 *        with one atomicAdd per element and 1024-row chunk: fp32 summation order is not fixed). */
V3D_API int v3d_rulebook_invert(const int* nbr, int nbr_stride, const int* n_out, int out_capacity, int kernel_volume,
                                int* inv, int inv_stride, v3d_stream_t stream);
V3D_API int v3d_sparse_conv_bwd_weight(const float* feat, const float* grad_out, const int* nbr, int nbr_stride,
                                       const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                       float* grad_weight, v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a3  SparseConvTensor.dense(): (N,C) rows -> (B,C,D,H,W), zero filled (sparse_cnn.py:128-133).
 * workspace holds the (B,D,H,W) int32 cell->row map.
 * ------------------------------------------------------------------------------------------- */
V3D_API size_t v3d_sparse_to_dense_workspace_bytes(int B, const int* shape_host);
V3D_API int v3d_sparse_to_dense(const float* feat, const int* indices, const int* n_rows,
                                int capacity_rows, int C, int B, const int* shape_host, float* out,
                                void* workspace, size_t workspace_bytes, v3d_stream_t stream);

/* Same scatter, but written as the BEV map the RPN consumes, in channels-last memory:
 * out (B, H, W, C*D) with channel index c*D + d == `dense().view(B, C*D, H, W)` (sparse_cnn.py:131-132)
 * in torch's channels_last format. C must divide 256, D <= 8. */
V3D_API int v3d_sparse_to_dense_nhwc(const float* feat, const int* indices, const int* n_rows,
                                     int capacity_rows, int C, int B, const int* shape_host, float* out,
                                     void* workspace, size_t workspace_bytes, v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a7  furthest_point_sample(xyz[B,N,3], m) -> idx[B,m] int32   (detector/model.py:53)
 * a8  gather_operation(feat[B,C,N], idx[B,m]) -> [B,C,m]          (detector/model.py:54)
 * a9  ball_query(radius, nsample, xyz[B,N,3], new_xyz[B,M,3]) -> idx[B,M,nsample] int32
 * a10 grouping_operation(feat[B,C,N], idx[B,M,ns]) -> [B,C,M,ns]
 *     query_and_group: [xyz[idx]-new_xyz ; feat[idx]] -> [B,3+C,M,ns] (QueryAndGroup, use_xyz)
 * (pointnet2 ops used through PointnetSAModuleMSG: detector/model.py:39-43,64;
 *  detector/roi_grid_pool.py:28-32,68)
 * ------------------------------------------------------------------------------------------- */
V3D_API size_t v3d_fps_workspace_bytes(int B, int N);
V3D_API int v3d_fps(const float* xyz, int B, int N, int m, int* idx, void* workspace,
                    size_t workspace_bytes, v3d_stream_t stream);
V3D_API int v3d_gather(const float* feat, const int* idx, int B, int C, int N, int m, float* out,
                       v3d_stream_t stream);
V3D_API int v3d_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M, float radius,
                           int nsample, int* idx, v3d_stream_t stream);
V3D_API int v3d_group(const float* feat, const int* idx, int B, int C, int N, int M, int nsample,
                      float* out, v3d_stream_t stream);
V3D_API int v3d_query_and_group(const float* xyz, const float* new_xyz, const float* feat,
                                const int* idx, int B, int C, int N, int M, int nsample, float* out,
                                v3d_stream_t stream);

/* a7+a8 fused: PV_RCNN.sample_keypoints (detector/model.py:46-56). `points` rows are `point_stride` floats
 * (3 = xyz, 4 = x,y,z,intensity as core/preprocess.py emits them; stride 4 is read with 128-bit loads, base 16-byte
 * aligned). Writes idx[B,m] and, when keypoints != NULL, keypoints[B,m,3] = points[idx, :3] (gather_operation +
 * the two transposes of model.py:54-55). */
V3D_API int v3d_fps_keypoints(const float* points, int point_stride, int B, int N, int m, int* idx,
                              float* keypoints, v3d_stream_t stream);

/* a9 for a whole PointnetSAModuleMSG (detector/model.py:39-43,64; detector/roi_grid_pool.py:28-32,68): all
 * `n_radii` (<= 4) ball queries of the module in one pass over the sources, one warp per query, same results as
 * n_radii calls of v3d_ball_query. Sources: row_offsets == NULL -> dense (B, N, point_stride) batch; else frame b
 * owns rows [row_offsets[b], row_offsets[b+1]) of a packed (rows, point_stride) array (DEVICE int[B+1]) and the
 * indices written are relative to the frame's first row -- this replaces pad_batch (detector/sparse_cnn.py:118-126).
 * radii_host / nsamples_host / idx_host are HOST arrays of n_radii entries (idx_host[r] = DEVICE int[B,M,ns_r]). */
V3D_API int v3d_ball_query_msg(const float* xyz, int point_stride, const int* row_offsets, const float* new_xyz,
                               int B, int N, int M, int n_radii, const float* radii_host,
                               const int* nsamples_host, int* const* idx_host, v3d_stream_t stream);

/* a9 with culling, same results as v3d_ball_query_msg bit for bit: v3d_ball_query_bounds writes one axis-aligned box
 * per 32 consecutive source rows of every frame into `bounds` (v3d_ball_query_bounds_bytes(B, max_rows_per_frame)
 * bytes; max_rows_per_frame >= every frame's row count), v3d_ball_query_msg_culled skips the chunks whose box is
 * farther from the query than the largest radius. Pays off when consecutive rows are spatially close: the sparse
 * levels produced by strided convolutions are in ascending (b,z,y,x) order. */
V3D_API size_t v3d_ball_query_bounds_bytes(int B, int max_rows_per_frame);
V3D_API int v3d_ball_query_bounds(const float* xyz, int point_stride, const int* row_offsets, int B, int N,
                                  int max_rows_per_frame, void* bounds, v3d_stream_t stream);
V3D_API int v3d_ball_query_msg_culled(const float* xyz, int point_stride, const int* row_offsets, const void* bounds,
                                      int max_rows_per_frame, const float* new_xyz, int B, int N, int M, int n_radii,
                                      const float* radii_host, const int* nsamples_host, int* const* idx_host,
                                      v3d_stream_t stream);

/* a9 for sources whose rows are NOT spatially ordered (shuffled raw points, first-appearance voxels), same results bit
 * for bit: v3d_ball_query_sort_x buckets every frame's rows along x into `sorted` (float4 rows {x, y, z, original row
 * index}; same frame ranges as the source; workspace of v3d_ball_query_sort_workspace_bytes(B) bytes, zeroed once),
 * v3d_ball_query_bounds on `sorted` (stride 4) gives thin chunk boxes, and v3d_ball_query_msg_select culls with them
 * and keeps, per radius, the nsample (<= 32) smallest ORIGINAL indices among the hits = the first nsample hits of the
 * reference's index-order scan. x_min / x_max: the extent of the scene along x (values outside are clamped). */
V3D_API size_t v3d_ball_query_sort_workspace_bytes(int B);
V3D_API int v3d_ball_query_sort_x(const float* xyz, int point_stride, const int* row_offsets, int B, int N,
                                  int max_rows_per_frame, float x_min, float x_max, void* sorted, void* workspace,
                                  v3d_stream_t stream);
V3D_API int v3d_ball_query_msg_select(const void* sorted, const int* row_offsets, const void* bounds,
                                      int max_rows_per_frame, const float* new_xyz, int B, int N, int M, int n_radii,
                                      const float* radii_host, const int* nsamples_host, int* const* idx_host,
                                      v3d_stream_t stream);

/* BEVFeatureGatherer.forward (vision3d/detector/layers.py:30-50): bilinear grid_sample (align_corners, zero padding) of
 * the BEV map at the keypoints, with the reference's index arithmetic (pixel = (xy - offset) / pixel_size, clamp,
 * (size - 2) normaliser, x <-> W swap). map_nhwc (B, H, W, C) f32 = a channels_last (B, C, H, W) tensor, C % 4 == 0;
 * keypoints (B, M, 3); out (B, c_total, M) f32, channels [c_off, c_off + C) written. */
V3D_API int v3d_bev_gather(const float* map_nhwc, int B, int H, int W, int C, const float* keypoints, int M,
                           float x_offset, float y_offset, float pixel_x, float pixel_y, float* out, int c_total,
                           int c_off, v3d_stream_t stream);

/* a10 QueryAndGroup(use_xyz=True) reading ROW-major sources: xyz (rows, xyz_stride), feat rows of C floats
 * `feat_stride` floats apart (C may be 0; feat may alias xyz, e.g. the intensity column of (x,y,z,i) points),
 * dense (row = b*N + idx) or ragged (row = row_offsets[b] + idx) -> out[B, 3+C, M, nsample]. Spares the
 * `features.transpose(1, 2).contiguous()` of detector/model.py:62 for the row-major sparse levels. */
V3D_API int v3d_query_and_group_rows(const float* xyz, int xyz_stride, const float* feat, int feat_stride, int C, int N,
                                     const int* row_offsets, const float* new_xyz, const int* idx, int B, int M,
                                     int nsample, float* out, v3d_stream_t stream);

/* SURVEY 8f-2 -- fused set abstraction: grouping (a10) -> shared MLP (two 1x1 convolutions, BatchNorm folded, ReLU) ->
 * max over nsample, for one scale of a PointnetSAModuleMSG (detector/model.py:58-66, detector/roi_grid_pool.py:26-33,68),
 * on the tensor cores (bf16 3-term split, fp32 accumulate, <= 1e-4 of the output scale); the grouped tensor
 * (B, 3+C, M, nsample) and the MLP activations never reach HBM.
 *   feat_packed : source rows [h1(0..Cp-1) | h2(0..Cp-1)] bf16, Cp = feature channels rounded up to 8 (zero padded);
 *                 dense (row = b*N + idx) or ragged (row = row_offsets[b] + idx, DEVICE int[B+1]) like v3d_ball_query_msg
 *   xyz / new_xyz / idx : as v3d_query_and_group_rows; nsample must be 16 or 32
 *   w1_prepared : v3d_sa_mlp_prepare of the layer-1 weights as (ceil((Cp+3)/64), 64, N1) fp32 chunks in K order
 *                 [features 0..Cp-1 | dx dy dz | 0...] (the reference's input order is [xyz ; features]: permute rows)
 *   w2_prepared : v3d_sa_mlp_prepare of the layer-2 weights as (ceil(N1/64), 64, N2) chunks; b1[N1], b2[N2] folded biases
 *   out[B, c_total, M] receives channels [c_off, c_off + N2). (N1, N2) in {(16,16), (32,32), (64,64), (192,96)}. */
V3D_API size_t v3d_sa_mlp_prepared_bytes(int n_chunks, int N);
V3D_API int v3d_sa_mlp_prepare(const float* weight_chunks, int n_chunks, int N, void* prepared, size_t prepared_bytes,
                               v3d_stream_t stream);
V3D_API int v3d_sa_fused(const void* feat_packed, int Cp, const float* xyz, int xyz_stride, const int* row_offsets, int N,
                         const float* new_xyz, const int* idx, int B, int M, int nsample, const void* w1_prepared,
                         const float* b1, int N1, const void* w2_prepared, const float* b2, int N2, float* out,
                         int c_total, int c_off, v3d_stream_t stream);
/* fp32 feat[b, c, n] (element strides given) -> packed rows (B*N, 2*Cp) bf16 [h1 | h2], channels C..Cp-1 zero: the
 * gather source format of v3d_sa_fused for channel-major tensors (keypoint features (B, 512, 2048), point intensity) */
V3D_API int v3d_pack_channel_major(const float* feat, long long b_stride, long long c_stride, long long n_stride, int B,
                                   int C, int N, int Cp, void* packed, v3d_stream_t stream);

/* a15 on the device: offsets[b] = first row of frame b in a (b,z,y,x)-ordered index list, offsets[B] = n_rows
 * (= torchsearchsorted.searchsorted(batch_index, arange(B+1)), compute_pad_amounts, detector/sparse_cnn.py:107-116,
 * without its .cpu().numpy() round trip). */
V3D_API int v3d_batch_offsets(const int* indices, const int* n_rows, int capacity_rows, int B, int* offsets,
                              v3d_stream_t stream);

/* to_global (detector/sparse_cnn.py:91-105): xyz[i] = float(x,y,z of indices[i]) * voxel_size_host + offset_host
 * (voxel_size_host = base voxel size * level stride, fp32, xyz order). */
V3D_API int v3d_to_global(const int* indices, const int* n_rows, int capacity_rows, const float* voxel_size_host,
                          const float* offset_host, float* xyz, v3d_stream_t stream);

/* pad_batch / pad_for_batch (detector/sparse_cnn.py:118-126, core/preprocess.py:35-45): ragged rows (rows, C) ->
 * dense out[B, frame_capacity, C]; a frame with fewer rows is filled with uniformly drawn duplicates of its own
 * rows (counter-based RNG keyed by (seed, frame, slot): tensors padded with the same seed stay row-paired). */
V3D_API int v3d_pad_batch(const float* src, int C, const int* row_offsets, int B, int frame_capacity,
                          unsigned long long seed, float* out, v3d_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Glue between the RPN heads and a12 for the SECOND inference path (engine-internal, replaces ~55 tiny
 * torch launches): ProposalLayer._decode + the BEV slice + batched_nms_rotated's coordinate offsets
 * (detector/proposal.py:47-70, core/box_encode.py:13-23, ops/iou_nms.py:121-132), and the gather of the
 * kept boxes into packed result rows [7 box | score | frame | class | valid] (+ one row of counters).
 *   reg_map      conv_reg output (B, n_cls*7*n_yaw, ny, nx) f32 with element strides reg_strides_host[4]
 *   anchors      (n_cls, n_yaw, ny, nx, 7) f32 contiguous
 *   anchor_idx   (B, n_cls, topk) int64 indices into the flattened (n_yaw, ny, nx) grid (torch.topk)
 *   boxes (B*n_cls*topk, 7), nms_in (B*n_cls*topk, 5) outputs
 * ------------------------------------------------------------------------------------------- */
V3D_API int v3d_second_head_decode(const float* reg_map, const long long* reg_strides_host, const float* anchors,
                                   const int64_t* anchor_idx, int B, int n_cls, int n_yaw, int ny, int nx,
                                   int topk, float* boxes, float* nms_in, v3d_stream_t stream);
/* The same stage without materialising the head maps (detector/proposal.py:61-78 computes conv_cls and conv_reg
 * over all ny*nx*n_yaw anchors and then keeps topk of them):
 *   v3d_head_cls_logits  : 1x1 classification conv on the channels_last (B, ny, nx, 128) RPN map ->
 *                          logits (B, n_out, ny*nx) [= conv_cls output, NCHW], n_out <= 8, weight (n_out,128)
 *   v3d_topk_rows        : row-wise top-k, values descending, ties -> lower index (torch.topk semantics up to
 *                          the order of equal values), k <= 256; sigmoid is monotonic, so it runs on logits
 *   v3d_head_reg_gather  : conv_reg evaluated only at the topk anchors -> deltas (N,7); scores = sigmoid(logit)
 *   v3d_second_head_decode_compact : decode + BEV + NMS group offsets from the compact deltas */
V3D_API int v3d_head_cls_logits(const float* fmap_nhwc, int B, int hw, int C, const float* weight,
                                const float* bias, int n_out, float* logits, v3d_stream_t stream);
V3D_API size_t v3d_topk_rows_workspace_bytes(int rows, int k);
V3D_API int v3d_topk_rows(const float* values, int rows, int row_len, int k, float* out_values,
                          int64_t* out_index, void* workspace /* may be NULL: single-pass */,
                          size_t workspace_bytes, v3d_stream_t stream);
V3D_API int v3d_head_reg_gather(const float* fmap_nhwc, int C, const float* w_reg, const float* b_reg,
                                const float* top_logits, const int64_t* anchor_idx, int B, int n_cls, int n_yaw,
                                int ny, int nx, int topk, float* deltas, float* scores, v3d_stream_t stream);
V3D_API int v3d_second_head_decode_compact(const float* deltas, const float* anchors, const int64_t* anchor_idx,
                                           int B, int n_cls, int n_yaw, int ny, int nx, int topk, float* boxes,
                                           float* nms_in, v3d_stream_t stream);
V3D_API int v3d_pack_detections(const float* boxes, const float* scores, const int64_t* keep, const int* count,
                                const float* score_thresh, int N, int n_cls, int topk,
                                const int* const* counters_dev, int n_counters, float* result,
                                v3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* V3D_B200_H_ */
