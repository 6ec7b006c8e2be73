/*
 * fsf_b200.h — C-ABI of the B200-native FSF hot path (libfsf_b200.so).
 *
 * Plain pointers + sizes, no torch types.  Every pointer marked "dev" is a CUDA
 * device pointer owned by the caller; every pointer marked "host" is host memory
 * read synchronously during the call.  All kernels are enqueued on `stream`
 * (a cudaStream_t passed as void*); no call synchronises the device or the
 * stream unless it says so.  Return value: FSFB_OK (0) or a negative status;
 * fsfb_last_error() gives a thread-local message for the last failure.
 *
 * The reference (BraveGroup/FullySparseFusion) has no C FFI of its own: its hot
 * path calls Python extension modules (torch_scatter, spconv, TorchEx,
 * ingroup_indices, mmdet3d.ops.Voxelization) and ATen.  Each entry point below
 * cites the reference call site (file:line under /root/reference) whose
 * arithmetic it replaces; INTEGRATION.md shows the ctypes stub per symbol.
 */
#ifndef FSF_B200_H_
#define FSF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSFB_OK 0
#define FSFB_ERR_BADARG (-1)   /* null pointer, negative size, unsupported width */
#define FSFB_ERR_CAPACITY (-2) /* workspace / key space / output buffer too small */
#define FSFB_ERR_CUDA (-3)     /* a CUDA runtime call or launch failed */

#define FSFB_REDUCE_SUM 0
#define FSFB_REDUCE_MEAN 1
#define FSFB_REDUCE_MAX 2
#define FSFB_REDUCE_DENSE 0x100 /* OR-able hint: every segment in [0,m) has at least one row (index from a ranking) */

#define FSFB_ACT_NONE 0
#define FSFB_ACT_RELU 1
#define FSFB_ACT_GELU 2
#define FSFB_RESIDUAL_POST 0x100 /* OR into `act`: y = act(norm(x + bias)) + residual (default adds before act) */
/* OR into `act` (fsfb_gather_gemm_split with a row_order only): `nbr` is already permuted into the row order and padded —
 * nbr[k][i] = neighbour of output row row_order[i], row stride round_up(rows, 128), padding = -1 — so that a tile's table is
 * 27 contiguous 512-byte runs instead of 27 x 128 scattered words (the scattered form cost as much load/store-unit time as the
 * operand gather itself: profiles/r2_ncu_summary.md section 2). */
#define FSFB_NBR_ROW_ORDERED 0x200

#define FSFB_NORM_NONE 0
#define FSFB_NORM_LAYERNORM 1 /* per-row LN over channels (mmcv 'LN')            */
#define FSFB_NORM_AFFINE 2    /* per-channel scale/shift (eval-mode BN folded)  */

/* ABI version (bumped on any signature change). */
int fsfb_version(void);  /* currently 2 */
/* Thread-local text of the last error on this thread ("" if none). */
const char* fsfb_last_error(void);
/* Number of kernel launches issued through this library since process start
 * (bench.py's gpu_launches counter). */
int64_t fsfb_launch_count(void);

/* ---------------------------------------------------------------------------
 * a1  Dynamic voxelization.
 * Replaces mmdet3d.ops.Voxelization(max_num_points=-1) as called at
 * projects/mmdet3d_plugin/models/detectors/single_stage_fsd.py:217-219 and the
 * in-tree formula torch.div(p - min, vs, rounding_mode='floor') at :270, :444,
 * :591-593, :948.  Evaluated in IEEE fp32 with no contraction;
 *   floor_mode 0: c = floor((p - range_min) / voxel)        (the Voxelization kernel)
 *   floor_mode 1: c = torch.div(p - range_min, voxel, rounding_mode='floor')
 *                 (ATen div_floor_floating: exact fmod first — the in-tree formula)
 * a point with any coordinate outside [0, grid) (or NaN) gets (-1,-1,-1).
 *   pts      dev  [n, row_stride] f32, xyz in columns 0..2
 *   range_min, voxel  host [3] f32 (x,y,z);  grid host [3] i32 (x,y,z)
 *   coors_zyx dev [n,3] i32 (z,y,x) — written.  order_xyz = 1 writes (x,y,z) instead and
 *   check_range = 0 keeps raw (possibly negative / beyond-grid) coordinates: the ClusterAssigner's
 *   BEV voxels of voted centres (single_stage_fsd.py:946-950), which are not range-filtered.
 * ------------------------------------------------------------------------- */
int fsfb_voxelize(const float* pts, int64_t n, int64_t row_stride,
                  const float* range_min, const float* voxel, const int32_t* grid,
                  int floor_mode, int order_xyz, int check_range, int32_t* coors_zyx,
                  void* stream);

/* ---------------------------------------------------------------------------
 * a2  Row ranking = torch.unique(rows, dim=0, return_inverse, return_counts)
 * as called by scatter_v2 (projects/mmdet3d_plugin/ops/sst_ops.py:156,165),
 * SIR (models/backbones/sir.py:68), pre_voxelize (single_stage_fsd.py:595).
 * Sort-free: rows are linearised into a bounded key space (per-column lo/extent),
 * marked in an L2-resident bitmap, and ranked by a popcount prefix, which yields
 * exactly the lexicographic-ascending ranks torch.unique(dim=0) returns.
 * ------------------------------------------------------------------------- */

/* Per-column min and max of integer rows.  minmax dev [2*d] i64 = {min[d], max[d]}.
 * rows_i64: 1 → rows are int64, 0 → int32.  n == 0 leaves {INT64_MAX, INT64_MIN}. */
int fsfb_rows_minmax(const void* rows, int rows_i64, int64_t n, int d,
                     int64_t* minmax, void* stream);

/* Workspace bytes fsfb_rank_rows needs for n rows over `cells` key cells. */
int fsfb_rank_workspace_bytes(int64_t n, int64_t cells, size_t* bytes);

/* Rank rows.  lo/ext host [d]: column j must lie in [lo[j], lo[j]+ext[j]); rows
 * outside raise *status (dev i32, bit0) and get inv = -1.  prod(ext) <= 2^32-1.
 *   inv32     dev [n] i32 (nullable)   rank of each row
 *   inv64     dev [n] i64 (nullable)   same, torch dtype
 *   uniq      dev [cap_unique, d] same dtype as rows (nullable) — unique rows ascending
 *   counts    dev [cap_unique] i32 (nullable) — rows per unique row
 *   num_unique dev [1] i32 — M (read it after syncing the stream)
 *   status    dev [1] i32 — 0 ok; bit0 row out of bounds; bit1 cap_unique exceeded
 */
int fsfb_rank_rows(const void* rows, int rows_i64, int64_t n, int d,
                   const int64_t* lo, const int64_t* ext,
                   void* workspace, size_t workspace_bytes,
                   int32_t* inv32, int64_t* inv64,
                   void* uniq, int64_t cap_unique, int32_t* counts,
                   int32_t* num_unique, int32_t* status, void* stream);

/* ---------------------------------------------------------------------------
 * a2/a12/a13  Segment CSR ("scatter rulebook") and segmented reductions =
 * torch_scatter.scatter_max / scatter(reduce='mean'|'sum') over dim 0 as called
 * at projects/mmdet3d_plugin/ops/sst_ops.py:168,170.
 * ------------------------------------------------------------------------- */

int fsfb_csr_workspace_bytes(int64_t n, int64_t m, size_t* bytes);

/* Build the CSR of `index` (values in [0,m), or <0 = dropped row):
 *   offsets dev [m+1] i32, perm dev [n] i32 (source rows grouped by segment,
 *   ascending source row inside a segment — stable), seg dev [n] i32 (segment of
 *   each sorted position).  Rows with index<0 are placed after offsets[m].
 * index_i64: 1 → int64 index, 0 → int32. */
int fsfb_csr_build(const void* index, int index_i64, int64_t n, int64_t m,
                   int32_t* offsets, int32_t* perm, int32_t* seg,
                   void* workspace, size_t workspace_bytes, void* stream);

int fsfb_segment_reduce_workspace_bytes(int64_t n, int c, int with_arg, size_t* bytes);

/* out[s, :] = reduce over rows r with index[r]==s of feat[r, :].
 * mode FSFB_REDUCE_{SUM,MEAN,MAX} (| FSFB_REDUCE_DENSE to skip the empty-segment pass).  MAX: ties → lowest source row (torch_scatter's
 * sequential CPU rule); argmax (nullable) dev [m,c] i64.  Empty segment → 0 and
 * argmax = n.  perm == NULL means rows are already grouped (perm = identity).
 *   feat dev [n, feat_stride] f32 (c <= feat_stride), out dev [m, c] f32 */
int fsfb_segment_reduce(const float* feat, int64_t n, int c, int64_t feat_stride,
                        const int32_t* perm, const int32_t* seg, const int32_t* offsets,
                        int64_t m, int mode, float* out, int64_t* argmax,
                        void* workspace, size_t workspace_bytes, void* stream);

/* out[i, :] = idx[i] >= 0 ? src[idx[i], :] : fill.  Voxel→point gather of
 * models/necks/voxel2point_neck.py:42-50 and voxel_center[unq_inv]
 * (models/detectors/FSF.py:310-311).  idx_i64 selects the index dtype. */
int fsfb_gather_rows(const float* src, int64_t m, int c, int64_t src_stride, const void* idx, int idx_i64,
                     int64_t n, float fill, float* out, int64_t out_stride, void* stream);

/* In-group index: out[i] = number of j < i with group[j] == group[i]  (stable
 * variant of ingroup_indices.forward, projects/mmdet3d_plugin/ops/sst_ops.py:246-248;
 * oracle = get_inner_win_inds_slow, models/middle_encoders/sst_input_layer.py:200-208).
 * group dev [n] i64 values in [0, m); out dev [n] i64. */
int fsfb_ingroup_workspace_bytes(int64_t n, int64_t m, size_t* bytes);
int fsfb_ingroup_indices(const int64_t* group, int64_t n, int64_t m, int64_t* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * a7+a8  LiDAR→camera projection fused with nearest sampling of the instance-id
 * planes = FSF.prj_points_2d + FSF.points_in_mask
 * (projects/mmdet3d_plugin/models/detectors/FSF.py:169-226): homogeneous
 * projection by lidar2img, depth > 1e-3, clip to [1e-5,1e5], normalise to (-1,1),
 * invalid → -2, then F.grid_sample(mode='nearest', align_corners=False, zeros).
 * The id planes are sampled as stored (u8 or i32): no float copy is made.
 *   xyz       dev [n, xyz_stride] f32
 *   lidar2img dev [cams,4,4] f32 row-major
 *   mask      dev [cams, classes, H, W]  (mask_i32: 0 → uint8, 1 → int32)
 *   out_ids   dev [n, cams, classes] i64  — the reference's obj_id_tensor
 * ------------------------------------------------------------------------- */
int fsfb_project_sample(const float* xyz, int64_t n, int64_t xyz_stride,
                        const float* lidar2img, int cams,
                        const void* mask, int mask_i32, int classes, int H, int W,
                        int64_t* out_ids, void* stream);

/* Fused contract (a9: FSF.img_cross_attn, FSF.py:714-718): same sampling, but only
 * the camera with the largest id sum is kept (first maximal camera on ties):
 *   ids_sel dev [n, classes] i32 (i32 so AV2 ids fit), cam_sel dev [n] u8,
 *   fg dev [n] u8 = any id > 0 over all cams/classes (FSF.extract_fg_pts, :299-308);
 *   overlap dev [n] u8 (nullable) = number of (cam,class) slots with id > 0;
 *   scores dev [n, classes] f32 (nullable) = anno[id-1][anno_col] of the selected ids, 0 for id 0 —
 *   get_all_cls_preds_2d + encode_preds_2d's score column (FSF.py:506-535, 449-474) with
 *   anno dev [anno_rows, anno_cols] f32 = mask_anno of the sample.  Every output is nullable. */
int fsfb_project_sample_select(const float* xyz, int64_t n, int64_t xyz_stride,
                               const float* lidar2img, int cams,
                               const void* mask, int mask_i32, int classes, int H, int W,
                               int32_t* ids_sel, uint8_t* cam_sel, uint8_t* fg,
                               uint8_t* overlap, const float* anno, int anno_rows, int anno_cols,
                               int anno_col, float* scores, void* stream);

/* EXPERIMENTAL twin of fsfb_project_sample_select (same reference lines, same outputs) for u8 planes stored class-interleaved:
 * mask_hwc16 dev [cams, H, W, 16] u8, 16-byte aligned — byte k of a texel = id plane of class k (FSF.py:202-226 samples the
 * planes one by one; here a texel's ids are one 16-byte load).  fullysparsefusion_b200.loading.LoadMaskFromFiles(layout="hwc16")
 * decodes the sample straight into this layout.  Not validated on hardware yet. */
int fsfb_project_sample_select_hwc(const float* xyz, int64_t n, int64_t xyz_stride,
                                   const float* lidar2img, int cams, const void* mask_hwc16,
                                   int classes, int H, int W, int32_t* ids_sel, uint8_t* cam_sel,
                                   uint8_t* fg, uint8_t* overlap, const float* anno, int anno_rows,
                                   int anno_cols, int anno_col, float* scores, void* stream);

/* ---------------------------------------------------------------------------
 * a5 / a9 / a10 / a16 / a17  Gather-GEMM with fused epilogue — the one dense
 * contraction of the path.  It serves nn.Linear stacks built by build_mlp
 * (projects/mmdet3d_plugin/ops/sst_ops.py:808-833: Linear → norm → act), and the
 * sparse convolutions of SimpleSparseUNet (config
 * projects/configs/nuScenes/FSF_nuScenes_config.py:58-70; spconv SubMConv3d /
 * SparseConv3d / SparseInverseConv3d) in output-stationary form:
 *
 *   out[r, :] = epilogue( sum_k  a[nbr[k][r], :] @ w[k]^T )        r in [0, rows)
 *
 * nbr[k][r] < 0 means "no input row for offset k" (contributes zero); nbr == NULL
 * with koff == 1 is a plain Linear over rows of `a`.  w is [koff][cout][cin]
 * (nn.Linear's [out,in] per offset), pre-packed once by fsfb_gemm_prepack into the
 * tensor-core tile layout (tf32 hi/lo split, 128 B swizzle).  Arithmetic: tcgen05
 * kind::tf32 MMAs in 3xTF32 split form (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32
 * accumulation in TMEM): |error| ~ 1e-6 relative, inside the 1e-4 parity budget
 * that plain TF32 (5e-4) would break.
 *
 * row_order (nullable, dev [rows] i32): a permutation of the output rows; tile i of the grid then processes
 * rows row_order[128 i .. 128 i + 127].  fsfb_rulebook_row_order sorts rows by their set of active offsets
 * so that a tile only visits the offsets its rows share (~2x fewer MMA stages on LiDAR voxel sets); results are
 * identical for any order.
 *
 * epilogue(x) = act( norm(x + bias) + residual ):
 *   norm FSFB_NORM_LAYERNORM: nn.LayerNorm(cout, eps) with norm_w/norm_b   (cout <= 256)
 *   norm FSFB_NORM_AFFINE   : x * norm_w[c] + norm_b[c]  (eval-mode BatchNorm1d /
 *                             naiveSyncBN1d folded: w = gamma/sqrt(var+eps), b = beta - mean*w)
 *   act  FSFB_ACT_RELU | FSFB_ACT_GELU (exact erf form, nn.GELU()) | NONE
 * ------------------------------------------------------------------------- */
int fsfb_gemm_prepack_bytes(int koff, int cin, int cout, size_t* bytes);
/* w dev [koff, cout, cin] f32 row-major → packed dev (bytes from the call above). */
int fsfb_gemm_prepack(const float* w, int koff, int cin, int cout, void* packed, void* stream);

int fsfb_gather_gemm(const float* a, int64_t a_rows, int cin, int64_t a_stride,
                     const int32_t* nbr, const int32_t* row_order, int koff, int64_t rows,
                     const void* w_packed, int cout,
                     const float* bias, int norm, const float* norm_w, const float* norm_b,
                     float eps, const float* residual, int64_t residual_stride, int act,
                     float* out, int64_t out_stride, void* stream);

/* Same contract with the offsets cut into `splits` contiguous ranges that run as independent work units
 * (split-K over the 27 offsets): each unit writes raw partial sums to workspace[split][rows][round_up(cout,128)]
 * and a second kernel sums the slabs in a fixed order and applies the epilogue (deterministic).  For levels with
 * few rows and wide channels (the 1 k / 11 k-row U-Net levels, models cited above), where rows/128 tiles leave most
 * SMs idle.  splits == 1 is fsfb_gather_gemm.  Needs cout <= 128 or a multiple of 128, cout <= 1024. */
int fsfb_gather_gemm_splitk_bytes(int64_t rows, int cout, int splits, size_t* bytes);
int fsfb_gather_gemm_splitk(const float* a, int64_t a_rows, int cin, int64_t a_stride,
                            const int32_t* nbr, const int32_t* row_order, int koff, int64_t rows,
                            const void* w_packed, int cout,
                            const float* bias, int norm, const float* norm_w, const float* norm_b,
                            float eps, const float* residual, int64_t residual_stride, int act,
                            float* out, int64_t out_stride, int splits, void* workspace,
                            size_t workspace_bytes, void* stream);

/* fsfb_gather_gemm_splitk plus HOST copies (nullable, [cout] f32) of the per-channel vectors bias / norm_w / norm_b.
 * When every device vector that is present has its host copy, cout <= 128 and splits == 1, the vectors travel in the
 * kernel parameters and the epilogue reads them as constant-bank operands (no shared-memory or global loads in the
 * epilogue warps, whose memory instructions otherwise queue behind the gathers).  Results are identical. */
int fsfb_gather_gemm_hv(const float* a, int64_t a_rows, int cin, int64_t a_stride,
                        const int32_t* nbr, const int32_t* row_order, int koff, int64_t rows,
                        const void* w_packed, int cout,
                        const float* bias, int norm, const float* norm_w, const float* norm_b,
                        float eps, const float* residual, int64_t residual_stride, int act,
                        float* out, int64_t out_stride, int splits, void* workspace,
                        size_t workspace_bytes, const float* host_bias, const float* host_norm_w,
                        const float* host_norm_b, void* stream);

/* Pre-split operands for the 27-offset convolutions.  fsfb_split_rows converts fp32 rows [rows, c] (c % 32 == 0, 16-byte
 * aligned rows) into the fp16-split row format of the tensor-core operand: per 32 channels one 128-byte block
 * [32 x fp16(a) | 32 x fp16((a - hi) * 2048)], out dev [rows * c * 4] bytes (HBM-bound: 8 c bytes per row).
 * fsfb_gather_gemm_split is fsfb_gather_gemm_hv reading such rows (dense, 4 * cin bytes per row): the gather then is pure data
 * movement (16-byte cp.async copies into the operand ring, no conversion per gathered row), which matters when every row is
 * gathered by ~6 of the 27 offsets.  Results are bit-identical to fsfb_gather_gemm on the fp32 rows. */
int fsfb_split_rows(const float* a, int64_t rows, int c, int64_t a_stride, void* out, void* stream);
int fsfb_gather_gemm_split(const void* a_split, int64_t a_rows, int cin,
                           const int32_t* nbr, const int32_t* row_order, int koff, int64_t rows,
                           const void* w_packed, int cout,
                           const float* bias, int norm, const float* norm_w, const float* norm_b,
                           float eps, const float* residual, int64_t residual_stride, int act,
                           float* out, int64_t out_stride, int splits, void* workspace,
                           size_t workspace_bytes, const float* host_bias, const float* host_norm_w,
                           const float* host_norm_b, void* stream);

/* The default arithmetic of fsfb_gather_gemm* is the fp16-split form (csrc/gemm_ss.cu: a = hi + lo / 2048 with fp16 hi and
 * lo, three kind::f16 MMAs per product, the same 22 mantissa bits as 3xTF32; FSFB_GEMM_F16=0 in the environment restores
 * 3xTF32).  Inputs must lie inside fp16 range (|a| < 65504; the reference's features are normalised activations and metric
 * coordinates).  A launch that meets a larger magnitude produces Inf/NaN in the affected rows and bumps this counter:
 * *count = number of such launches on the current device since the library was loaded (synchronises the device). */
int fsfb_gemm_f16_overflows(unsigned int* count);

/* Diagnostics only: per-CTA role cycle counters [148][32] u32 of the last fsfb_gather_gemm launch made with
 * FSFB_GEMM_TIMERS=1 in the environment (layout in csrc/gemm_ts.cu). */
int fsfb_debug_gemm_timers(unsigned int* out);
/* Same for the shared-memory-operand kernel (layout in csrc/gemm_ss.cu). */
int fsfb_debug_gemm_ss_timers(unsigned int* out);

/* Same contract evaluated with plain fp32 FMAs on CUDA cores from the UNPACKED weights
 * (w dev [koff,cout,cin]).  Independent cross-check of the tensor-core path at sizes the
 * CPU oracle cannot reach; tests only — nothing in the product path calls it. */
int fsfb_gather_gemm_simt(const float* a, int64_t a_rows, int cin, int64_t a_stride,
                          const int32_t* nbr, int koff, int64_t rows,
                          const float* w, int cout,
                          const float* bias, int norm, const float* norm_w, const float* norm_b,
                          float eps, const float* residual, int64_t residual_stride, int act,
                          float* out, int64_t out_stride, void* stream);

/* Row-wise norm + activation in place-or-out: y = act(norm(x + bias) + residual) for
 * rows wider than the fused epilogue handles (the 1024-wide LN of the cluster heads,
 * dense_heads/sparse_cluster_head.py:75-77). */
int fsfb_rownorm_act(const float* x, int64_t rows, int c, int64_t x_stride, const float* bias,
                     int norm, const float* norm_w, const float* norm_b, float eps,
                     const float* residual, int64_t residual_stride, int act,
                     float* out, int64_t out_stride, void* stream);

/* ---------------------------------------------------------------------------
 * a5  Sparse-convolution rulebook (spconv's indice pairs) in output-stationary form.
 * Replaces the hash → rulebook step inside spconv.SubMConv3d / SparseConv3d /
 * SparseInverseConv3d as used by SimpleSparseUNet (config
 * projects/configs/nuScenes/FSF_nuScenes_config.py:58-70; import
 * projects/mmdet3d_plugin/ops/sst_ops.py:5).  The voxel index is the bitmap that
 * fsfb_rank_rows / fsfb_conv_out_index leave at the START of their workspace (keep that
 * buffer alive): a coordinate's row is found by bit test + popcount prefix, so there is no
 * hash table.  Offsets are numbered k = (kz*K_y + ky)*K_x + kx.
 *   nbr dev [koff, m_out] i32:  nbr[k][o] = input row read by offset k of output o, or -1
 *     forward    (transposed=0): in = o*stride - pad + k      (SubM: stride 1, in set == out set)
 *     transposed (transposed=1): in = (o + pad - k)/stride when divisible  (inverse conv:
 *                                the pairs of the forward strided conv with the same geometry)
 *   The reference's (in_row, out_row) pair lists per offset are {(nbr[k][o], o) : nbr[k][o] >= 0},
 *   already sorted by out_row.
 * coors are int32 [m,4] rows (b,z,y,x), 16-byte aligned.
 * ------------------------------------------------------------------------- */
int fsfb_conv_rulebook(const int32_t* out_coors, int64_t m_out, const void* in_index,
                       const int64_t* in_lo, const int64_t* in_ext, const int32_t* ksize,
                       const int32_t* stride, const int32_t* pad, int transposed, int32_t* nbr,
                       void* stream);

/* Row order for fsfb_gather_gemm: stable sort of output rows by their 27-bit mask of present offsets.
 * order dev [rows] i32. */
int fsfb_rulebook_order_workspace_bytes(int64_t rows, int koff, size_t* bytes);
int fsfb_rulebook_row_order(const int32_t* nbr, int koff, int64_t rows, int32_t* order, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Output site set of a strided SparseConv3d + its voxel index: every output cell reached by at
 * least one (input site, offset).  workspace (fsfb_rank_workspace_bytes(0, cells)) receives
 * the index of the OUTPUT set; out_coors dev [cap_out,4] i32 receives its rows in
 * lexicographic (b,z,y,x) order (== their row numbers); num_out/status as in fsfb_rank_rows. */
int fsfb_conv_out_index(const int32_t* in_coors, int64_t m_in, const int64_t* out_lo,
                        const int64_t* out_ext, const int32_t* ksize, const int32_t* stride,
                        const int32_t* pad, void* workspace, size_t workspace_bytes,
                        int32_t* out_coors, int64_t cap_out, int32_t* num_out, int32_t* status,
                        void* stream);

/* ---------------------------------------------------------------------------
 * a14  Connected-components clustering of voted centres.
 * Replaces find_connected_componets / find_connected_componets_single_batch
 * (projects/mmdet3d_plugin/models/detectors/single_stage_fsd.py:45-82: dense m x m
 * distance matrix on the GPU → host → scipy.sparse.csgraph.connected_components → device)
 * and torchex.connected_components(points, batch_idx, dist, 100, 2, False) (:37-43).
 * Adjacency: fp32 sqrt(dx^2 + dy^2) < dist over columns 0,1 of `points`.
 *   batch_idx dev [m] i32 or NULL.  NULL = the single-batch variant (:69-82, what the stock
 *   config runs at inference; batch ids ignored).  Non-NULL = per-sample clustering (:45-67):
 *   points of different samples are never adjacent and components are numbered sample by sample.
 *   labels dev [m] i32: component id, contiguous from 0, numbered in order of each component's
 *   lowest member index (scipy's first-visit order) — bit-exact with the reference.
 *   num_components dev [1] i32 (nullable).
 * ------------------------------------------------------------------------- */
int fsfb_ccl_workspace_bytes(int64_t m, size_t* bytes);
int fsfb_connected_components(const float* points, int64_t m, int64_t stride, const int32_t* batch_idx,
                              float dist, int32_t* labels, int32_t* num_components, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Per-point fused passes between the scatters and the GEMMs (one HBM pass each).
 * ------------------------------------------------------------------------- */

/* a3  DynamicScatterVFE point decoration (registry type named at
 * projects/configs/nuScenes/FSF_nuScenes_config.py:42-52, called at
 * models/detectors/single_stage_fsd.py:232):
 *   out[i] = [ feats[i, 0:cin] | xyz - voxel_mean[inv[i], 0:3] | xyz - (coor*vs + vs/2 + min) ]
 * coors dev [n,4] (b,z,y,x) i32/i64; inv dev [n] i32; voxel_mean dev [m,cin]; out dev [n, cin+3+3]. */
int fsfb_vfe_decorate(const float* feats, int64_t n, int cin, int64_t feat_stride, const void* coors,
                      int coors_i64, const int32_t* inv, const float* voxel_mean, const float* voxel_size,
                      const float* range_min, int with_cluster_center, int with_voxel_center, float* out,
                      void* stream);

/* a12  SIRLayer input (models/backbones/sir.py:41-62 builds it with xyz_normalizer and rel_mlp):
 *   out[i,c] = (c < 3 ? feats[i,c] / xyz_normalizer[c] : feats[i,c]) * (gate ? gate[i,c] : 1) */
int fsfb_sir_input(const float* feats, int64_t n, int c, int64_t feat_stride, const float* xyz_normalizer,
                   const float* gate, int64_t gate_stride, float* out, int64_t out_stride, void* stream);

/* out[i,c] = x[i,c] / divisors[c]   (f_cluster / rel_dist_scaler, sir.py:57; divisors on device) */
int fsfb_div_cols(const float* x, int64_t n, int c, int64_t x_stride, const float* divisors_dev, float* out,
                  int64_t out_stride, void* stream);

/* x[i,c] += y[i,c]   (SIRLayer shortcut; x_merge.features + x.features in the U-Net decoder) */
int fsfb_add_inplace(float* x, int64_t n, int c, int64_t x_stride, const float* y, int64_t y_stride, void* stream);

/* SparseUNet.reduce_channel: out[i,c] = sum_j x[i, c*(cin/cout) + j]  (features.view(n,cout,-1).sum(2)) */
int fsfb_reduce_channel(const float* x, int64_t n, int cin, int64_t x_stride, int cout, float* out, void* stream);

/* a6  Voxel2PointScatterNeck.forward (projects/mmdet3d_plugin/models/necks/voxel2point_neck.py:42-67):
 *   out[i] = [ voxel_feats[inv[i], :] | xyz - ((coor + 0.5)*vs + min) ],  mask[i] = !(gathered row == padding everywhere)
 * out dev [n, out_stride >= c+3]; mask dev [n] u8; dropped dev [1] i32 = number of masked-out points (the caller compacts
 * only when it is non-zero, as the reference's boolean indexing would). */
int fsfb_neck_points(const float* points, int64_t n, int64_t pts_stride, const void* coors, int coors_i64,
                     const float* voxel_feats, int64_t m, int c, const void* inv, int inv_i64,
                     const float* voxel_size, const float* range_min, float padding, float* out, int64_t out_stride,
                     uint8_t* mask, int32_t* dropped, void* stream);

/* a10  VoteSegHead.decode_vote_targets (models/decode_heads/segmentation_head.py:265-266): v * |v| */
int fsfb_vote_decode(const float* preds, int64_t total, float* out, void* stream);

/* Stable boolean-mask compaction: idx[k] = position of the k-th non-zero mask byte; *count = how many.
 * (points[mask] of FSF.extract_fg_pts, FSF.py:299-308; group_sample, single_stage_fsd.py:828-850) */
int fsfb_compact_workspace_bytes(int64_t n, size_t* bytes);
/* mask[i] = x[i*stride + col] > thr  (SingleStageFSD.get_fg_mask, single_stage_fsd.py:753-757) */
int fsfb_threshold_mask(const float* x, int64_t n, int64_t stride, int col, float thr, uint8_t* mask, void* stream);
/* mask[i] = counts[inv[i]] >= min_count  (filter_almost_empty, single_stage_fsd.py:31-35) */
int fsfb_count_mask(const int32_t* counts, const int32_t* inv, int64_t n, int min_count, uint8_t* mask, void* stream);
int fsfb_compact_indices(const uint8_t* mask, int64_t n, int32_t* idx, int32_t* count, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * a11 / a13 / a15  Query-generation helpers.
 * ------------------------------------------------------------------------- */

/* FSF.get_point_fg_weights (projects/mmdet3d_plugin/models/detectors/FSF.py:345-355) and
 * SingleStageFSD.group_sample (models/detectors/single_stage_fsd.py:802-865) in one pass:
 *   p = softmax(logits[i, 0:C+1]);  fg_weight[i] = 1 - p[C]
 *   group_score[i,g] = sum_{c in group g} p[c];   w = offset weights ('max' mode, :868-874)
 *   group_center[i,g,:] = xyz[i] + sum_{c in g} w_c * offsets[i, c, :]
 * group_lens host [n_groups], group_classes host [sum lens] (class indices).  Outputs nullable. */
int fsfb_group_sample(const float* logits, int64_t n, int num_classes_with_bg, const float* xyz, int64_t xyz_stride,
                      const float* offsets, const int32_t* group_lens, const int32_t* group_classes, int n_groups,
                      float* fg_weight, float* group_score, float* group_center, void* stream);

/* FSF.extract_fg_pts + double_overlap_pts + get_sir_coors (FSF.py:260-308, 357-365).
 * gather_overlap: ov32[j] = min(overlap[idx_fg[j]], 17) (input of fsfb_csr_build with m = 18 groups; group 17 = a point seen by
 * more than the 16 (camera, class) masks frustum_expand keeps per point — the caller must treat a non-empty group 17 as an error:
 * the reference appends overlap - 1 rows for any count).
 * frustum_expand: given that CSR (perm/seg/offsets over overlap counts), writes for every output row
 * the source point (rows_point, index into the full point array) and sir_coors (batch, 0, object id)
 * in the reference's row order; total rows = n_fg + sum_k cnt_k*(k-1).  status bit 2 = a point's
 * re-sampled id count disagrees with its overlap count (or exceeds 16). */
int fsfb_gather_overlap(const uint8_t* overlap, const int32_t* idx_fg, int64_t n_fg, int32_t* ov32, void* stream);
int fsfb_frustum_expand(const float* xyz, int64_t xyz_stride, const float* lidar2img, int cams, const void* mask,
                        int mask_i32, int classes, int H, int W, const int32_t* idx_fg, int64_t n_fg,
                        const int32_t* perm, const int32_t* seg, const int32_t* offsets, const int32_t* batch_idx,
                        int32_t* rows_point, int32_t* sir_coors, int32_t* status, void* stream);

/* get_cluster_delta_weighted (FSF.py:313-329): out4[r] = (xyz*w, w), w = clamp(weight[src], 1e-5). */
int fsfb_weighted_xyz(const float* xyz, int64_t xyz_stride, const float* weight, const int32_t* rows, int64_t n_rows,
                      float* out4, void* stream);
/* center[k] = mean[k,:3] / mean[k,3] (mean_cols == 4) or mean[k,:3] (== 3);
 * f_cluster[r] = xyz[src] - center[inv[r]]   (FSF.py:324-329; single_stage_fsd.py:460-462) */
int fsfb_cluster_delta(const float* xyz, int64_t xyz_stride, const int32_t* rows, int64_t n_rows, const float* mean,
                       int64_t k, int mean_cols, const int32_t* inv, float* center, float* f_cluster, void* stream);

/* get_single_cls_preds_2d + encode_preds_2d (FSF.py:449-504): preds_2d dev [k, anno_cols],
 * feat dev [k, 5 + num_classes + 1] = (bbox / (w,h,w,h), score, one-hot category). */
int fsfb_encode_preds_2d(const float* anno, int anno_rows, int anno_cols, const int32_t* obj_coors, int coor_stride,
                         int coor_col, int64_t k, float img_w, float img_h, int num_classes, float* preds_2d,
                         float* feat, void* stream);

/* a12  SIRLayer relative-position gate, fused: out = cat(xyz/xyz_normalizer, feats[:,3:]) * rel_mlp(f_cluster/scaler)
 * with rel_mlp = build_mlp(3, [h1, h2, c], LN(eps), act) (three Linear(bias=False) → LayerNorm → act blocks;
 * built by SIR at projects/mmdet3d_plugin/models/backbones/sir.py:41-62, rel_dist_scaler = 10 at :57).
 * w1 dev [h1,3], w2 dev [h2,h1], w3 dev [c,h2] (nn.Linear layouts), ln*_w/ln*_b the LayerNorm affines.
 * h1, h2 <= 32, 3 <= c <= 256.  feats_b (nullable): the input row is cat(feats[:, :c_a], feats_b[:, :c-c_a]) —
 * SIR.forward's torch.cat([points, out_feats], 1) (sir.py:76) without materialising it. */
int fsfb_sir_gate_input(const float* feats, int64_t n, int c, int64_t feat_stride, const float* feats_b,
                        int64_t feats_b_stride, int c_a, const float* f_cluster,
                        int64_t fc_stride, float rel_dist_scaler, const float* xyz_normalizer, int h1, int h2,
                        const float* w1, const float* ln1_w, const float* ln1_b, const float* w2,
                        const float* ln2_w, const float* ln2_b, const float* w3, const float* ln3_w,
                        const float* ln3_b, float eps, int act, float* out, int64_t out_stride, void* stream);

/* a13/a14 for ALL class groups in one pass (csrc/group_cluster.cu): the per-group selection of
 * SingleStageFSD.group_sample (projects/mmdet3d_plugin/models/detectors/single_stage_fsd.py:822-842) and
 * ClusterAssigner.forward_single_class (:936-982), which the reference runs as a Python loop over the class groups.
 * The candidate list is ordered (group, row); the group id is the batch column of the ranking and of the CCL.
 *   fsfb_group_flags     flags dev [G, n] u8 = score[:, g] > thresholds[g] (host [G]); a group with no candidate keeps
 *                        row 0 (:833-835); counts dev [G] i32 = candidates per group before that fallback
 *   fsfb_group_split     flat (ascending indices into flags, from fsfb_compact_indices) → grp, vox, cidx = vox*G + grp
 *   fsfb_group_voxelize  rows4 dev [t,4] i32 = (grp, floor_div(centre - range_min, voxel_sizes[grp]))  (:946-950);
 *                        voxel_sizes host [G,3]
 *   fsfb_group_keep      keep = counts[inv] >= min_points, and everything of a group where nothing survives (:953-955)
 *   fsfb_group_relabel   out[i] = labels[inv[i]] - (smallest label of that row's group): cluster ids from 0 per group;
 *                        batch = group id of each labelled row (stride in i32 elements), base dev [G] scratch
 *   fsfb_connected_components_groups   fsfb_connected_components with one distance per batch id (host [n_batches] <= 8)
 */
int fsfb_group_flags(const float* score, int64_t n, int64_t stride, int n_groups, const float* thresholds, uint8_t* flags,
                     int32_t* counts, void* stream);
int fsfb_group_split(const int32_t* flat, int64_t t, int64_t n, int n_groups, int32_t* grp, int32_t* vox, int32_t* cidx,
                     void* stream);
int fsfb_group_voxelize(const float* centers, int64_t t, const int32_t* grp, const float* range_min, const float* voxel_sizes,
                        int n_groups, int32_t* rows4, void* stream);
int fsfb_group_keep(const int32_t* counts, const int32_t* inv, const int32_t* grp, int64_t t, int n_groups, int min_points,
                    uint8_t* keep, int32_t* kept_per_group, void* stream);
int fsfb_group_relabel(const int32_t* labels, const int32_t* batch, int64_t batch_stride, int64_t m, int n_groups,
                       const int32_t* inv, int64_t t, int32_t* base, int32_t* out, void* stream);
int fsfb_connected_components_groups(const float* points, int64_t m, int64_t stride, const int32_t* batch_idx,
                                     const float* dist_per_batch, int n_batches, int32_t* labels, int32_t* num_components,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * f1  Dynamic point pooling of the query-refinement stage.
 * Replaces dynamic_point_pool_ext.forward(rois, pts, extra_wlh, max_inbox_point, out_pts_idx, out_roi_idx,
 * out_pts_feats) (projects/mmdet3d_plugin/ops/dynamic_point_pool_op.py:27-32; caller
 * models/roi_heads/roi_extractors/dynamic_point_roi_extractor.py:52-57).  Source un-vendored: semantics restated
 * from the published FSD kernel and the invariants asserted in-tree (dynamic_point_roi_extractor.py:84-92).
 *   rois dev [k,7] f32 (cx,cy,cz,w,l,h,rz), gravity centre; pts dev rows of >= 3 floats (x,y,z first).
 *   A point belongs to a roi when it lies inside the box enlarged by extra_wlh (host [3]: added to l, w, h).
 *   Outputs (caller-allocated, `capacity` rows, upstream 50000; rows past *num_out are left untouched so the caller's
 *   -1 / 0 prefill keeps its meaning): out_pts_idx / out_roi_idx dev i64, out_pts_feats dev [capacity,13] f32 =
 *   (x,y,z, local x,y,z, local + half dims (3), half dims - local (3), in-margin flag).
 *   Order: roi-major, point index ascending; at most max_inbox_point lowest-index points per roi; the first
 *   `capacity` entries of that order (upstream's atomics make both choices run-dependent).
 *   num_out dev [1] i32: rows written.
 * ------------------------------------------------------------------------- */
int fsfb_dynamic_point_pool_workspace_bytes(int64_t k, int max_inbox_point, size_t* bytes);
/* FSF.decode_stage_bboxes (projects/mmdet3d_plugin/models/detectors/FSF.py:1085-1095) = BasePointBBoxCoder.decode
 * (core/bbox/coders/base_point_bbox_coder.py:59-82) plus the batch column: reg dev [k, code_size] = (dxyz, log dims, sin, cos
 * [, vx, vy]), base_points dev [k, >=3]; rois dev [k, code_size] = (batch, xyz = dxyz + base, dims = exp(.) - 1e-6,
 * yaw = atan2(sin, cos)[, vx, vy]).  batch dev i32 (strided) or NULL = 0.  code_size 8 or 10. */
int fsfb_decode_boxes(const float* reg, int64_t k, int code_size, int64_t reg_stride, const float* base_points,
                      int64_t base_stride, const int32_t* batch, int64_t batch_stride, float* rois, void* stream);
int fsfb_dynamic_point_pool(const float* rois, int64_t k, const float* pts, int64_t n, int64_t pts_stride,
                            const float* extra_wlh, int max_inbox_point, int64_t capacity,
                            long long* out_pts_idx, long long* out_roi_idx, float* out_pts_feats,
                            int32_t* num_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * f3  Multi-class rotated BEV NMS: the tail of FrustumClusterHead._get_bboxes_single
 * (projects/mmdet3d_plugin/models/dense_heads/frustum_cluster_head.py:595-698) = sigmoid, box3d_multiclass_nms with
 * use_rotate_nms (mmdet3d fork / iou3d, un-vendored: rotated-rectangle IoU restated by polygon clipping).  Three steps with
 * the two list lengths read by the caller in between (as fsfb_compact_indices reports them):
 *   fsfb_nms_flags     scores dev [k, C] (sigmoid(logits) when apply_sigmoid), flags dev [C, k] u8 = score > score_thr,
 *                      counts dev [C] i32
 *   fsfb_nms_suppress  flat = ascending indices of the set flags (class-major candidates); per class: order by (score desc,
 *                      box index asc), suppress later boxes whose rotated BEV IoU with a kept one exceeds nms_thr;
 *                      keep dev [candidates] u8 in that order.  boxes dev rows (x, y, z, dx, dy, dz, yaw, ...).
 *                      max_class = largest counts[c] (<= 65536).
 *   fsfb_nms_emit      kept_idx = ascending indices of the set keep flags → boxes [n, box_dim], scores [n], labels [n] i64,
 *                      source row [n] i32 (nullable), n = min(kept, max_num): class-major in descending score, or the max_num
 *                      best scores in descending order when more survive.  Same workspace as fsfb_nms_suppress.
 * ------------------------------------------------------------------------- */
int fsfb_nms_flags(const float* logits, int64_t k, int num_classes, int64_t stride, int apply_sigmoid, float score_thr,
                   float* scores, uint8_t* flags, int32_t* counts, void* stream);
int fsfb_nms_workspace_bytes(int64_t candidates, int max_class, size_t* bytes);
int fsfb_nms_suppress(const float* boxes, int64_t k, int64_t box_stride, const float* scores, int num_classes, const int32_t* flat,
                      int64_t candidates, const int32_t* counts, int max_class, float nms_thr, uint8_t* keep, void* workspace,
                      size_t workspace_bytes, void* stream);
int fsfb_nms_emit(const float* boxes, int64_t box_stride, int box_dim, const int32_t* kept_idx, int64_t kept, int64_t candidates,
                  int max_class, int64_t max_num, float* out_boxes, float* out_scores, long long* out_labels, int32_t* out_box_idx,
                  void* workspace, size_t workspace_bytes, void* stream);

/* The neighbour table in the row order of fsfb_rulebook_row_order, padded with -1 to round_up(rows, 128) columns: the form
 * FSFB_NBR_ROW_ORDERED expects.  out dev [koff][round_up(rows, 128)] i32.  (No reference counterpart: spconv keeps pair lists.) */
int fsfb_permute_rulebook(const int32_t* nbr, int koff, int64_t rows, const int32_t* order, int32_t* out, void* stream);

/* -------------------------------------------------------------------------
 * Weight gradient of the gather-GEMM (training)
 * Replaces: the weight-gradient half of the backward of spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d (un-vendored;
 * the layers of SimpleSparseUNet, projects/configs/FSF_nuScenes_config.py:58-70) and of nn.Linear inside build_mlp
 * (projects/mmdet3d_plugin/ops/sst_ops.py:808-833), as run by tools/train.py:244-251.
 *   dw[k][co][ci] = sum over output rows r with 0 <= nbr[k][r] < a_rows of dy[r][co] * a[nbr[k][r]][ci]
 *   a dev [a_rows, cin] f32 (row stride a_stride), dy dev [rows, cout] f32 (row stride dy_stride), nbr dev [koff][rows] i32 or
 *   null with koff == 1 (Linear: every row pairs with itself), dw dev [koff][cout][cin] f32, overwritten.
 *   Deterministic: pairs are summed in ascending row order inside fixed row splits, splits in order (no atomics).
 * ------------------------------------------------------------------------- */
int fsfb_conv_wgrad_workspace_bytes(int64_t rows, int koff, int cin, int cout, size_t* bytes);
int fsfb_conv_wgrad(const float* a, int64_t a_rows, int cin, int64_t a_stride, const float* dy, int64_t rows, int cout,
                    int64_t dy_stride, const int32_t* nbr, int koff, float* dw, void* workspace, size_t workspace_bytes,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSF_B200_H_ */
