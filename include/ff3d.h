/*
 * ff3d.h -- C ABI of libff3d.so: the sm_100a kernels behind the FocalFormer3D per-scene forward path.
 *
 * Boundary rules
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated here.
 *   - every entry point returns 0 on success, a negative FF3D_E* code otherwise;
 *     ff3d_last_error() returns a thread-local message for the last failure.
 *   - all launches are asynchronous on the given stream; no host synchronisation happens inside,
 *     data-dependent sizes (voxel counts, active-site counts) stay in device memory (`*_dev` args).
 *
 * Each entry point names the reference interface it replaces.  Paths are relative to the reference
 * repository (NVlabs/FocalFormer3D @ d5bb555); [upstream] marks the pinned third-party op the
 * reference calls at that site (mmdet3d v0.17.1 / mmcv-full 1.3.18, not vendored in the reference tree).
 */
#ifndef FF3D_H_
#define FF3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FF3D_OK 0
#define FF3D_EINVAL (-1)   /* bad argument / unsupported shape */
#define FF3D_ECUDA (-2)    /* CUDA runtime error at launch */
#define FF3D_EWORKSPACE (-3)

#define FF3D_ACT_NONE 0
#define FF3D_ACT_RELU 1
#define FF3D_ACT_RELU6 2

#define FF3D_MAX_BATCH 64

typedef void* ff3d_stream_t; /* cudaStream_t */

const char* ff3d_last_error(void);
int ff3d_version(void);

/* ------------------------------------------------------------------------------------------------
 * Hard voxelisation + voxel feature mean.
 * Replaces: [upstream] mmdet3d.ops.voxel hard_voxelize (deterministic) as called per sample from
 *   projects/mmdet3d_plugin/models/detectors/focalformer3d.py:189-209 (FocalFormer3D.voxelize) and
 *   [upstream] HardSimpleVFE at focalformer3d.py:166.
 * points: [n_total, n_feat] fp32, the B samples concatenated; batch_offsets (HOST) [batch+1] row offsets.
 * Outputs are sized for cap = batch*max_voxels rows:
 *   voxels [cap, max_points, n_feat] (may be NULL), coors [cap,4] int32 (b,z,y,x), num_points [cap] int32,
 *   mean_feats [cap, mean_ld] (may be NULL; columns >= n_feat are written as 0),
 *   n_voxels_dev [1 + batch] int32: total, then per-sample counts.
 * Voxel order = sample-major, first-appearance order inside a sample (the reference's order).
 * Dynamic mode (max_points <= 0; [upstream] Voxelization(max_num_points=-1) + DynamicSimpleVFE as used at
 *   focalformer3d.py:159-163,213-238 by DeformFormer3D_L_dynamic): no per-voxel point cap -- pass the largest
 *   per-sample point count as max_voxels --, voxels must be NULL, num_points = all points of the voxel, mean_feats =
 *   their mean (fp64 atomic sums); same voxel order as above.
 */
size_t ff3d_voxelize_workspace_bytes(int n_total, int batch, int max_voxels, int max_points);
int ff3d_voxelize_hard(const float* points, int n_total, int n_feat, const int* batch_offsets_host, int batch,
                       const float* voxel_size3, const float* pc_range6, int max_points, int max_voxels,
                       float* voxels, int* coors, int* num_points, float* mean_feats, int mean_ld,
                       int* n_voxels_dev, void* workspace, size_t workspace_bytes, ff3d_stream_t stream);

/* HardVFE with one VFELayer and max pooling (Waymo configs): out[v,c] = max over all max_points slots of
 * relu(x[v,k,:] . w[:,c] + b[c]); padded slots are masked to zero first and therefore contribute relu(b[c]).
 * Replaces: [upstream] mmdet3d HardVFE / VFELayer at focalformer3d.py:166 (cfg FocalFormer3D_Waymo_L.py:141-152).
 * voxels [cap, max_points, n_feat], w [n_feat, C] and b [C] with BatchNorm1d(eval) folded in, out [cap, ldo]. */
int ff3d_vfe_hard(const float* voxels, const int* num_points, const int* n_voxels_dev, int cap, int max_points,
                  int n_feat, const float* w, const float* b, float* out, int ldo, int C, ff3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse-conv rulebooks (output-stationary neighbour maps).
 * Replaces: [upstream] mmdet3d.ops.spconv get_indice_pairs (indice_cuda.cu) as used by every
 *   SubMConv3d / SparseConv3d of SparseEncoder (focalformer3d.py:168; cfg FocalFormer3D_L.py:198-206).
 * A level is (coors [cap,4] int32 (b,z,y,x), n_dev int32[1], dense shape D,H,W, batch).  The hash table is
 *   (hkeys uint32[hsize], hvals int32[hsize]), hsize a power of two >= 2*cap; the caller memsets nothing:
 *   ff3d_sp_hash_build clears it.  nbr maps are [taps, cap] int32, -1 = no neighbour.
 */
int ff3d_sp_hash_build(const int* coors, const int* n_dev, int cap, int batch, int D, int H, int W,
                       uint32_t* hkeys, int* hvals, int hsize, ff3d_stream_t stream);
/* SubM k=3 p=1: nbr[t][o] = row of the site at coors[o] + (kz-1,ky-1,kx-1), t = (kz*3+ky)*3+kx */
int ff3d_sp_subm_map(const int* coors, const int* n_dev, int cap, int batch, int D, int H, int W,
                     const uint32_t* hkeys, const int* hvals, int hsize, int* nbr, ff3d_stream_t stream);
/* SparseConv3d (kernel k3, stride s3, padding p3): creates the output level (coors_out, n_out_dev, its hash)
 * and nbr_out [kvol, cap_out]: input row feeding output o through tap t (i = o*s - p + k).
 * overflow_dev int32[1] is set to 1 if more than cap_out output sites exist. */
int ff3d_sp_down_build(const int* coors_in, const int* n_in_dev, int cap_in, int batch, int D, int H, int W,
                       const uint32_t* hkeys_in, const int* hvals_in, int hsize_in,
                       const int* k3, const int* s3, const int* p3,
                       int* coors_out, int* n_out_dev, int cap_out, int Do, int Ho, int Wo,
                       uint32_t* hkeys_out, int* hvals_out, int hsize_out, int* nbr_out, int* overflow_dev,
                       ff3d_stream_t stream);
/* ---- mask-sorted rulebooks (round 2): the output rows of every sparse conv are ordered by their tap-presence mask so
 * that the 128-row tiles of the gather-GEMM are homogeneous and the kernel skips, per tile, every tap no row of the tile
 * uses (ff3d_gemm_desc.tile_mask).  Same [upstream] get_indice_pairs semantics as above (i = o*s - p + k); only the ROW
 * ORDER -- implementation-defined in spconv -- changes.
 *   ff3d_sp_down_sites   the site-creation half of ff3d_sp_down_build: insert the proposed output keys, then number the
 *                        occupied hash slots in slot order (block counts + scan: no contended row counter, deterministic)
 *   ff3d_sp_tap_keys     keys[o] = tap mask of output row o (probing the INPUT level's hash) with the bits permuted into
 *                        rarity order (corner taps most significant); kvol = k0*k1*k2 <= 27 key bits
 *   ff3d_sort_pairs      stable LSD radix sort of (key, value) with a device-side count; vals_in NULL -> 0..n-1
 *   ff3d_sp_level_permute  store a level in sorted order: coors_out[i] = coors_in[perm[i]], hash values := new rows
 *   ff3d_sp_nbr_build    nbr[t][j] for tile position j (source row o = perm ? perm[j] : j), tile_mask[j/128] = OR of
 *                        the natural-order tap masks of the tile's rows, and the row map y_off (y_mode 1: the ROW o, for desc->y_row;
 *                        y_mode 2: NHWC BEV element offset ((b*bev_h + y)*bev_w + x)*ldy + z*bev_c; 0: none)
 *   ff3d_sp_gather_rows  dst[i,:cols] = src[perm[i],:cols] (voxel features into the sorted level-1 order) */
int ff3d_sp_down_sites(const int* coors_in, const int* n_in_dev, int cap_in, int batch, int D, int H, int W,
                       const int* k3, const int* s3, const int* p3, int* coors_out, int* n_out_dev, int cap_out,
                       int Do, int Ho, int Wo, uint32_t* hkeys_out, int* hvals_out, int hsize_out, int* overflow_dev,
                       int* scratch /* ff3d_sp_down_sites_scratch_ints(hsize_out) ints */, ff3d_stream_t stream);
int ff3d_sp_down_sites_scratch_ints(int hsize);
int ff3d_sp_tap_keys(const int* coors_out, const int* n_out_dev, int cap_out, int D, int H, int W,
                     const uint32_t* hkeys_in, const int* hvals_in, int hsize_in, const int* k3, const int* s3,
                     const int* p3, uint32_t* keys, int* nbr_unsorted /* optional [kvol, cap_out]: the probed rows */,
                     ff3d_stream_t stream);
size_t ff3d_sort_workspace_bytes(int cap);
int ff3d_sort_pairs(const uint32_t* keys_in, const int* vals_in, const int* n_dev, int cap, int key_bits,
                    uint32_t* keys_out, int* vals_out, void* workspace, size_t workspace_bytes, ff3d_stream_t stream);
int ff3d_sp_level_permute(const int* coors_in, const int* perm, const int* n_dev, int cap, int D, int H, int W,
                          int* coors_out, const uint32_t* hkeys, int* hvals, int hsize, int* inv /* optional: old row -> new */,
                          ff3d_stream_t stream);
/* ff3d_sp_nbr_build without a second round of hash probes: re-orders the rows ff3d_sp_tap_keys kept (nbr_unsorted) into
 * tile order, remapping them through inv when the probed level itself was re-ordered (SubM) */
int ff3d_sp_nbr_permute(const int* nbr_unsorted, const int* perm, const int* inv, const int* coors_out, const int* n_out_dev,
                        int cap_out, int kvol, int* nbr, uint32_t* tile_mask, int* y_off, int y_mode, int ldy, int bev_h,
                        int bev_w, int bev_c, ff3d_stream_t stream);
int ff3d_sp_nbr_build(const int* coors_out, const int* perm, const int* n_out_dev, int cap_out, int D, int H, int W,
                      const uint32_t* hkeys_in, const int* hvals_in, int hsize_in, const int* k3, const int* s3,
                      const int* p3, int* nbr, uint32_t* tile_mask, int* y_off, int y_mode, int ldy, int bev_h, int bev_w,
                      int bev_c, ff3d_stream_t stream);
int ff3d_sp_gather_rows(const float* src, int ld_src, const int* perm, const int* n_dev, int cap, float* dst, int ld_dst,
                        int cols, ff3d_stream_t stream);
/* element offsets for scattering the last sparse level straight into the NHWC BEV grid:
 * off[o] = ((b*H + y)*W + x)*ld + z*C   (replaces SparseConvTensor.dense() + view, [upstream]) */
int ff3d_sp_bev_offsets(const int* coors, const int* n_dev, int cap, int H, int W, int ld, int C, int* off,
                        ff3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM: y[row(m), n] = act( sum_{t,c} x[src(m,t), c] * w[t, c, n] + bias[n] + res[m, n] )
 * One kernel family covers
 *   mode FF3D_GEMM_ROWS   : linear / 1x1 conv (src(m,0) = m)                      -- F.linear, nn.Conv1d(k=1)
 *   mode FF3D_GEMM_CONV2D : NHWC conv, kh x kw window, stride, zero padding        -- cuDNN conv2d in the reference
 *   mode FF3D_GEMM_SPARSE : rulebook gather (src = nbr[t][m])                      -- [upstream] spconv indice_conv
 * Replaces: [upstream] mmdet3d.ops.spconv indice_conv (gather + torch::mm + scatter-add per offset), the cuDNN
 *   convs of SECOND/SECONDFPN (focalformer3d.py:169-171), FocalEncoder (focal_encoder.py:204-219), the head's
 *   ConvModules (focal_decoder.py:588,637,819-823) and every nn.Linear / Conv1d of the decoder
 *   (focal_decoder.py:872-939).  BatchNorm(eval) is folded into w/bias by the host.
 */
#define FF3D_GEMM_ROWS 0
#define FF3D_GEMM_CONV2D 1
#define FF3D_GEMM_SPARSE 2

typedef struct ff3d_gemm_desc {
  int mode;
  int M;               /* rows (ROWS/CONV2D: exact; SPARSE: capacity) */
  const int* m_dev;    /* optional device row count (SPARSE); NULL -> M */
  int cin, cout, taps; /* cin % 4 == 0 */
  const float* x; int ldx;         /* input rows, ldx % 4 == 0 */
  const float* x2;                 /* optional second addend of the A operand (ROWS mode): A = x + x2 */
  const float* w; int ldw;         /* [taps, cin, ldw], ldw % 4 == 0, ldw >= cout */
  const float* bias;               /* [cout] or NULL */
  const float* res; int ldres;     /* residual rows (indexed like y rows without the offset map) or NULL */
  float* y; int ldy;
  int act;
  /* CONV2D geometry: input [B, H, W, cin] with batch stride x_bstride rows; output [B, Ho*uy, Wo*ux, *] */
  int B, H, W, Ho, Wo, kh, kw, stride, pad;
  long long x_bstride, y_bstride;  /* rows per batch element (0 -> H*W / Ho*uy*Wo*ux) */
  long long y_row0;                /* row offset added inside each batch element */
  int ux, uy, dx, dy;              /* output up-sampling lattice (transposed conv k=s): row=(oy*uy+dy, ox*ux+dx); 0 -> 1 */
  /* SPARSE */
  const int* nbr; int nbr_stride;  /* [taps, nbr_stride] */
  const int* y_off;                /* optional per-row ELEMENT offset into y (replaces m*ldy) */
  int res_after_act;               /* 0: act(acc+bias+res) (residual blocks); 1: act(acc+bias)+res (query_feat += roi_feat) */
  const uint32_t* tile_mask;       /* SPARSE, optional: [ceil(M/128)] OR of the tap masks (bit t = tap t) of each 128-row
                                      tile (ff3d_sp_nbr_build); taps whose bit is clear are skipped for the tile.  Every
                                      (row, tap) with nbr >= 0 must have its bit set.  NULL = all taps. */
  /* ---- split activations (round 2): a row of C channels stored as fp16 [hi(C) | lo(C)], value = hi + lo / 2048 ----- */
  const void* xs; int ldxs;        /* A rows in split form: halves [xs_rows, ldxs], hi plane at column 0, lo plane at xs_lo  */
  int xs_lo; int xs_rows;          /* (ff3d_tmagemm only; x may then be NULL).  xs_rows = rows addressable through xs        */
  const void* res_s; int ldres_s; int res_s_lo;   /* residual rows in split form (ff3d_tmagemm; instead of res)            */
  void* ys; int ldys; int ys_lo;   /* optional split copy of the output rows (both kernels); y may be NULL in ff3d_tmagemm  */
  const int* y_row;                /* SPARSE, optional: output ROW of tile position m (strided convs over mask-sorted rows)  */
  int zero_row;                    /* SPARSE + xs: index of an all-zero row of xs, read for absent neighbours                */
} ff3d_gemm_desc;

int ff3d_igemm(const ff3d_gemm_desc* desc, ff3d_stream_t stream);

/* Same contract on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, weights staged by
 * cp.async.bulk) with 3xTF32 split accumulation, i.e. fp32-grade accuracy: D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi.
 * Supported when ff3d_tcgemm_ntile(cin, cout) > 0 (cin in {8,16} or a multiple of 32; cout in {16,32,64} or a
 * multiple of 128).  `wimg` = [cout/ntile][ff3d_tcgemm_stages(cin,taps)][2][ntile*32] floats: per (N tile, K step)
 * the hi and lo TF32 parts of w as 128B-swizzled K-major shared-memory images (focalformer3d_b200/ops.py
 * tc_weight_images).  desc->w / ldw are ignored. */
int ff3d_tcgemm(const ff3d_gemm_desc* desc, const float* wimg, ff3d_stream_t stream);
/* same with an explicit N tile (16/32/64/128 dividing cout) the images were packed for; 0 = default */
int ff3d_tcgemm_bn(const ff3d_gemm_desc* desc, const float* wimg, int ntile, ff3d_stream_t stream);
int ff3d_tcgemm_ntile(int cin, int cout);
int ff3d_tcgemm_stages(int cin, int taps);
/* Same contract with an fp16 hi/lo operand split on kind::f16 (twice the MMA rate of kind::tf32, 64 K-values per 128-byte
 * shared-memory row): a = hi + 2^-11 lo, hi = rn_f16(a), lo = rn_f16((a - hi) 2^11); D_main += A_hi*B_hi,
 * D_cross += A_hi*B_lo + A_lo*B_hi, y = D_main + 2^-11 D_cross -- the same 22-bit products as the TF32 split.
 * Supported when ff3d_tcgemm_f16_ntile(cin, cout) > 0 (cin in {8,16,32} or a multiple of 64).  `wimg16` =
 * [cout/ntile][ff3d_tcgemm_f16_stages(cin,taps)][2][ntile*64] halves (focalformer3d_b200/ops.py tc_weight_images_f16).
 * Values beyond +-65504 saturate: *overflow_dev (int32, may be NULL) is OR-ed with 1 and the result is invalid -- the
 * caller must check the flag and fail (or rerun the layer through ff3d_tcgemm). */
int ff3d_tcgemm_f16(const ff3d_gemm_desc* desc, const void* wimg16, int ntile, int* overflow_dev, ff3d_stream_t stream);
int ff3d_tcgemm_f16_ntile(int cin, int cout);
int ff3d_tcgemm_f16_stages(int cin, int taps);
/* TMA-fed variant on PRE-SPLIT activations (desc->xs): the A operand is moved global -> swizzled shared memory by the TMA
 * unit (2-D tile loads for ROWS, 4-D NHWC box loads with out-of-bounds zero fill for stride-1 CONV2D, tile::gather4 of the
 * rulebook rows for SPARSE), no conversion and no producer warps; weights = the ff3d_tcgemm_f16 images.  Outputs: fp32 rows
 * (y) and / or split rows (ys) for the next TMA layer; residual from fp32 (res) or split (res_s) rows.
 * ff3d_tmagemm_supported: cin a multiple of 64, cout 64 or a multiple of 128, no x2, conv stride 1. */
int ff3d_tmagemm(const ff3d_gemm_desc* desc, const void* wimg16, int ntile, int* overflow_dev, ff3d_stream_t stream);
int ff3d_tmagemm_supported(const ff3d_gemm_desc* desc);
void ff3d_tmagemm_conv_patch(int Ho, int Wo, int* bw_out, int* bh_out);
/* fp32 rows [rows, C] -> split rows (for activations produced by non-GEMM kernels) and back (for non-GEMM consumers) */
int ff3d_split_rows(const float* x, int ldx, const int* n_dev, long long rows, int C, void* ys, int ldys, int ys_lo,
                    int* overflow_dev, ff3d_stream_t stream);
int ff3d_unsplit_rows(const void* xs, int ldxs, int xs_lo, const int* n_dev, long long rows, int C, float* y, int ldy,
                      ff3d_stream_t stream);

/* Depthwise 3x3 stride 1 pad 1, NHWC, folded BN + activation (torchvision InvertedResidual dw conv,
 * focal_encoder.py:36-38). x [B,H,W,C] (ldx), w [9, C], bias [C]. */
int ff3d_dwconv3x3(const float* x, int ldx, const float* w, const float* bias, float* y, int ldy, int B, int H, int W,
                   int C, int act, ff3d_stream_t stream);

/* same depthwise conv with the output in split form only: ys [B*H*W, 2C] fp16 [hi | lo] (A operand of the TMA-fed project conv) */
int ff3d_dwconv3x3_split(const float* x, int ldx, const float* w, const float* bias, void* ys, int B, int H, int W, int C, int act,
                         int* overflow_dev, ff3d_stream_t stream);

/* y = LayerNorm(x) * gamma + beta over the last dim C (nn.LayerNorm, eps) -- [upstream] mmcv BaseTransformerLayer norms */
int ff3d_layernorm(const float* x, const float* gamma, const float* beta, float* y, int rows, int C, float eps,
                   ff3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Hard-Instance-Probing stage: sigmoid * accumulated mask -> class-aware local-max NMS -> top-k ->
 * query gathers -> accumulated-mask update.  Replaces focal_decoder.py:631-782 (one HIP stage).
 * logits [B,H,W,ldl] NHWC (first C channels used); logits2 (optional, NULL for HIP stages): second heatmap whose
 * sigmoid is averaged with the first (single-stage DeformFormer3D head, focal_decoder.py:547-549); acc_mask [B,C,H,W] fp32 (in/out);
 * nms_heat [B,C,H,W] fp32 workspace/out (the NMS'ed masked heatmap of this stage);
 * feat [B,H,W,ldf] stage feature; cls_w [C, Cf] (class_encoding weight transposed), cls_b [Cf];
 * outputs at query slot q0..q0+k-1 of nq_total: top_idx [B,k] int32 (flat class*H*W + pos, canonical order:
 * descending value, ties -> lower index), query_feat [B, nq_total, Cf], query_pos [B, nq_total, 2],
 * query_score [B, nq_total, C], query_label [B, nq_total] int32 (token-major rows).
 * exempt_lo..exempt_hi: classes using a 1x1 window (nuScenes 8..9, Waymo 1..2), focal_decoder.py:678-683,777-780.
 */
size_t ff3d_hip_workspace_bytes(int B, int C, int H, int W);
int ff3d_hip_stage(const float* logits, int ldl, const float* logits2, int ldl2, float* acc_mask, float* nms_heat, const float* feat, int ldf, int Cf,
                   const float* cls_w, const float* cls_b, int B, int C, int H, int W, int k, int nms_kernel,
                   int exempt_lo, int exempt_hi, int q0, int nq_total, int* top_idx, float* query_feat,
                   float* query_pos, float* query_score, int* query_label, void* workspace, size_t workspace_bytes,
                   ff3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder pieces.
 */
/* gen_sineembed_for_position (models/utils/utils.py:40-66): pos [rows,2] (x,y) in BEV-cell units, divided by
 * (w, h) inside; dim_t [128] = 10000^(2*(j//2)/128) supplied by the host;
 * out [rows,256] = (sin/cos of y | sin/cos of x), scale 2*pi. */
int ff3d_sine_embed(const float* pos, float w, float h, const float* dim_t, float* out, int rows, ff3d_stream_t stream);

/* nn.MultiheadAttention core (softmax(QK^T/sqrt(d)) V) for the query self-attention:
 * q,k,v [B, Nq, heads*d] with row strides ldq/ldk/ldv; out [B,Nq,heads*d].  ([upstream] mmcv MultiheadAttention) */
int ff3d_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                  int B, int Nq, int heads, int d, ff3d_stream_t stream);

/* Multi-scale deformable attention sampling.
 * Replaces: [upstream] mmcv.ops ms_deform_attn_forward (ms_deform_attn_cuda.cu) called through
 *   MultiScaleDeformableAttention at focal_decoder.py:927-933.
 * value [B, n_tokens, ldv] (heads*d channels starting at column v_col0), levels (H_l, W_l, start_l) HOST arrays;
 * ref [B,Nq,2] (x,y), divided by (ref_w, ref_h) inside (focal_decoder.py:869); offs [B,Nq,heads*L*P*2]; attw logits [B,Nq,heads*L*P] (softmax over L*P done here);
 * out [B,Nq,heads*d]. */
int ff3d_msda(const float* value, int ldv, int v_col0, long long v_bstride, const int* lvl_h, const int* lvl_w,
              const int* lvl_start, int L, int P, const float* ref, float ref_w, float ref_h, const float* offs,
              int ldoffs, const float* attw, int ldattw, float* out, int B, int Nq, int heads, int d,
              ff3d_stream_t stream);

/* ROI feature sampling (focal_decoder.py:890-919): per query a g x g grid in the (expanded) box frame, rotated by
 * yaw, normalised by roi_range (x0,y0,x1,y1), clipped to [-2,2], bilinear (align_corners=False, zeros) on L levels.
 * query_box [B*Nq, box_ld] rows = (centre(2) in cell units, height, log-dims(3), sin, cos, ...) -- the previous
 * decoder stage's prediction row; value buffer [B, n_tokens, ldv] (first C columns) with level geometry as in
 * ff3d_msda; out [B*Nq, L*g*g*C] ordered (level, point, channel). */
int ff3d_roi_sample(const float* query_box, int box_ld, const float* value, int ldv, long long v_bstride,
                    const int* lvl_h, const int* lvl_w, const int* lvl_start, int L, int C, int g, float expand,
                    float cell_x, float cell_y, float origin_x, float origin_y, const float* roi_range4, float* out,
                    int B, int Nq, ff3d_stream_t stream);

/* same sampling with the output rows in split form: out_split [B*Nq, 2*K] fp16 = [hi(K) | lo(K)], K = L*g*g*C (the A
 * operand of the TMA-fed roi_mlp.0 GEMM: no 180 MB fp32 operand, no in-kernel conversion) */
int ff3d_roi_sample_split(const float* query_box, int box_ld, const float* value, int ldv, long long v_bstride,
                          const int* lvl_h, const int* lvl_w, const int* lvl_start, int L, int C, int g, float expand,
                          float cell_x, float cell_y, float origin_x, float origin_y, const float* roi_range4,
                          void* out_split, int B, int Nq, int* overflow_dev, ff3d_stream_t stream);

/* Box-state update after the prediction heads (focal_decoder.py:945-957). pred [rows, ldp] rows =
 * (center2, height1, dim3, rot2, [vel2], class logits...): center += query_pos; query_pos = center;
 * with prev != NULL (roi_based_reg): dim[:2] += prev.dim[:2], rot += prev.rot. */
int ff3d_head_update(float* pred, int ldp, float* query_pos, const float* prev, int ldprev, int rows,
                     ff3d_stream_t stream);

/* Class-aware regression select (focal_decoder.py:940-943, `classaware_reg=True`, FocalFormer3D_Waymo15_L): the
 * prediction heads emit num_classes regression sets per query (channel = class * k + d for each of the n_groups heads of
 * widths group_k[]), followed by `tail` pass-through columns (class logits); out keeps the set of label[row]. */
int ff3d_class_select(const float* full, int ld_full, const int* label, const int* group_k, int n_groups, int tail,
                      int num_classes, float* out, int ld_out, int rows, ff3d_stream_t stream);

/* One decoder stage in ONE launch: the stage's `n_layers` DeformableTransformer decoder layers and its prediction heads.
 * Replaces: focal_decoder.py:927-939 -- self.decoder[i](query_feat, ..., query_pos, value, reference_points) ([upstream]
 *   mmcv BaseTransformerLayer, operation_order (self_attn, norm, cross_attn, norm, ffn, norm): nn.MultiheadAttention over the
 *   scene's queries, MultiScaleDeformableAttention over the BEV pyramid, FFN 128 -> ffn -> 128, three LayerNorms) followed by
 *   self.prediction_heads[i] (Conv1d 128->64 + BN + ReLU, Conv1d 64->k per head, here one GEMM + one block-diagonal GEMM).
 * All weight matrices are nn.Linear-style [N, K] matrices packed in MMA B-fragment order with fp16 hi / lo parts
 * (focalformer3d_b200/model.py pack_frag): int32 [N/8][K/16][32 lanes][4] = {hi b0, hi b1, lo b0, lo b1}, N zero-padded to a
 * multiple of 16; biases fp32, zero-padded alike.  value = the batched value projection [B, n_tokens, ldv], layer j reading
 * columns [128 j, 128 j + 128); q_pos [B*nq, 2] in cells, divided by (ref_w, ref_h) inside (focal_decoder.py:869).
 * Outputs: x_out [B*nq, 128] (the stage's query features, may be NULL) and pred [B*nq, ld_pred] (first pred_cols columns).
 * Built for hidden = 128, heads = 8; workspace from ff3d_decoder_stage_workspace_bytes (K / V exchange + counters). */
typedef struct ff3d_decoder_layer_weights {
  const void* w_qkv; const float* b_qkv;        /* in_proj [384, 128]: rows (q | k | v)                                  */
  const void* w_o;   const float* b_o;          /* attention out_proj [128, 128]                                         */
  const void* w_oa;  const float* b_oa;         /* sampling_offsets (heads*L*P*2) | attention_weights (heads*L*P) rows   */
  const void* w_op;  const float* b_op;         /* MSDA output_proj [128, 128]                                           */
  const void* w_f1;  const float* b_f1;         /* FFN [ffn, 128]                                                        */
  const void* w_f2;  const float* b_f2;         /* FFN [128, ffn]                                                        */
  const float* ln_gamma[3]; const float* ln_beta[3];
} ff3d_decoder_layer_weights;
typedef struct ff3d_decoder_stage_desc {
  int B, nq, n_layers, hidden, heads, n_levels, n_points, ffn;
  int lvl_h[4], lvl_w[4], lvl_start[4];
  const float* x_in; const float* qpe; const float* q_pos;
  float ref_w, ref_h;
  const float* value; int ldv; long long v_bstride;
  ff3d_decoder_layer_weights layers[4];
  const void* w_h1; const float* b_h1; int n_h1;      /* heads, first layer: [n_h1, 128], ReLU                               */
  const void* w_h2; const float* b_h2; int n_pred;    /* heads, second layer: [n_pred, n_h1] (n_pred padded to 16)           */
  float* x_out; float* pred; int ld_pred, pred_cols;
  void* workspace; size_t workspace_bytes;
  int* overflow_dev;                                  /* raised when an activation leaves the fp16 range (like ff3d_tmagemm) */
} ff3d_decoder_stage_desc;
size_t ff3d_decoder_stage_workspace_bytes(int B, int nq, int n_layers);
int ff3d_decoder_stage(const ff3d_decoder_stage_desc* desc, ff3d_stream_t stream);

/* Final scoring + box decode (focal_decoder.py:1313-1321 + transfusion_bbox_coder.py:71-158, nms_type=None):
 * pred as above (class logits at column cls_col); boxes [rows, 9|7] = (x,y,z_bottom,w,l,h,yaw[,vx,vy]),
 * scores [rows], labels [rows] int32, keep [rows] uint8 (post_center_range test). */
int ff3d_box_decode(const float* pred, int ldp, int cls_col, int has_vel, const float* query_score,
                    const int* query_label, int rows, int C, float cell_x, float cell_y, float origin_x, float origin_y,
                    const float* post_range6, float* boxes, float* scores, int* labels, unsigned char* keep,
                    ff3d_stream_t stream);

/* ---- input side -------------------------------------------------------------------------------
 * Multi-sweep point assembly on the device.  Replaces: [upstream] mmdet3d v0.17.1 LoadPointsFromMultiSweeps (test mode) and
 * PointsRangeFilter of the data pipeline (projects/configs/focalformer3d/FocalFormer3D_L.py:100-111).
 * raw [n_total, n_feat_in] = the key frame followed by the sweeps as loaded from the files (x, y, z, intensity, ...);
 * sweep_offsets (HOST) [n_sweeps+1]; per sweep (HOST): rot [9] row-major sensor2lidar_rotation and trans [3] (float64, as in
 * the info files), dt (time lag written to column 4), remove_close (drop |x| < r and |y| < r), transform (apply rot / trans).
 * range6 (may be NULL): PointsRangeFilter bounds (strict).  out [n_total, 5]: dropped points are overwritten with pad_value
 * (outside every point_cloud_range, so the voxeliser discards them) instead of being compacted away: order and size fixed. */
int ff3d_assemble_sweeps(const float* raw, int n_feat_in, const int* sweep_offsets_host, int n_sweeps, const double* rot_host,
                         const double* trans_host, const float* dt_host, const unsigned char* remove_close_host,
                         const unsigned char* transform_host, float close_radius, const float* range6_host, float pad_value,
                         float* out, ff3d_stream_t stream);
/* Camera frames: uint8 [n, H, W, 3] (BGR, as cv2 decodes them) -> float32 planar [n, 3, pad_h, pad_w]: bilinear resize to
 * (out_h, out_w) with cv2.INTER_LINEAR's float arithmetic, BGR->RGB, (x - mean) * (1 / std), zero padding to a multiple of
 * size_divisor.  Replaces LoadMultiViewImageFromFiles(to_float32) + ScaleImageMultiViewImage + NormalizeMultiviewImage +
 * PadMultiViewImage + DefaultFormatBundle3D (projects/mmdet3d_plugin/datasets/pipelines/transform_3d.py:125-249). */
int ff3d_image_preprocess(const unsigned char* img, int n, int H, int W, int out_h, int out_w, const float* mean3,
                          const float* std3, int to_rgb, int size_divisor, float* out, int pad_h, int pad_w,
                          ff3d_stream_t stream);

/* ---- output side ------------------------------------------------------------------------------
 * Per-task NMS of get_bboxes when test_cfg.nms_type is set (focal_decoder.py:1333-1393): task t owns the classes of
 * class_masks[t] (bit c); radius[t] <= 0 keeps every box of the task (:1376).
 *   mode 0 'circle': [upstream] mmdet3d.core.circle_nms -- score-descending greedy, a kept box suppresses boxes whose SQUARED
 *     centre distance is <= radius[t]; at most post_max (83 = the upstream default) boxes per task.
 *   mode 1 'rotate': [upstream] mmdet3d.ops.iou3d nms_gpu on xywhr2xyxyr(boxes.bev) -- the pre_max best boxes, rotated-BEV
 *     IoU > radius[t] suppresses, at most post_max kept.
 * boxes [B, nq, box_ld] = (x, y, z, dx, dy, dz, yaw, ...), scores / labels / keep_in [B, nq]; keep_out [B, nq] uint8 =
 * keep_in AND survived.  nq <= 1024. */
int ff3d_nms_tasks(const float* boxes, int box_ld, const float* scores, const int* labels, const unsigned char* keep_in,
                   int B, int nq, int n_tasks, const unsigned int* class_masks_host, const float* radius_host, int mode,
                   int pre_max, int post_max, unsigned char* keep_out, ff3d_stream_t stream);
/* [upstream] mmdet3d.ops.iou3d boxes_iou_bev on (x, y, z, dx, dy, dz, yaw, ...) rows: iou [n, m] (merge_augs.py:148) */
int ff3d_boxes_iou_bev(const float* a, int lda, int n, const float* b, int ldb, int m, float* iou, ff3d_stream_t stream);
/* box voting of the TTA merge (merge_augs.py:149-157): out[i, :dim] = sum_j w_ij box_j / (sum_j w_ij + 1e-6) with
 * w = iou zeroed below vote_thresh; yaw = atan2 of the weighted sine / cosine sums. */
int ff3d_box_voting(const float* iou, int n_sel, int m, const float* boxes, int ld, int dim, float vote_thresh, float* out,
                    ff3d_stream_t stream);
/* [upstream] mmdet3d bbox3d_mapping_back for LiDAR boxes (merge_augs.py:91-94), in place */
int ff3d_boxes_map_back(float* boxes, int ld, int dim, int n, float scale_factor, int flip_horizontal, int flip_vertical,
                        ff3d_stream_t stream);

/* ---- camera branch (DeformFormer3D_C_R50) -------------------------------------------------
 * Image repack for the NHWC convolution path: x [n, C, H, W] planar (what extract_img_feat receives,
 * focalformer3d.py:133-141) -> y [n, H, W, ld], channels [C, ld) zero (ld % 4 == 0). */
int ff3d_nchw_to_nhwc(const float* x, float* y, int n, int C, int H, int W, int ld, ff3d_stream_t stream);
/* nn.MaxPool2d(3, stride 2, padding 1) of the ResNet stem ([upstream] mmdet ResNet.maxpool): x [n,H,W,C] ->
 * y [n, (H-1)/2+1, (W-1)/2+1, C]. */
int ff3d_maxpool3x3s2(const float* x, float* y, int n, int H, int W, int C, ff3d_stream_t stream);
/* FPN top-down step ([upstream] mmdet FPN.forward: laterals[i-1] += F.interpolate(laterals[i], size, 'nearest')):
 * dst [n,Hd,Wd,C] += src [n,Hs,Ws,C] at (min(floor(y*Hs/Hd),Hs-1), min(floor(x*Ws/Wd),Ws-1)). */
int ff3d_upsample_add(float* dst, const float* src, int n, int Hd, int Wd, int Hs, int Ws, int C, ff3d_stream_t stream);
/* Lift-Splat-Shoot lift + splat (lss.py:126-146 CamEncode soft-max / outer product, :228-271 get_geometry without
 * augmentation matrices, :324-362 voxel_pooling, :373-377 s2c) in one pass, no [B,N,D,H,W,C] volume in HBM.
 *   dn      [B*cams, fH, fW, ld] depthnet output, columns [0,64) context features, [64, 64+D) depth logits
 *   frustum [D, fH, fW, 3] (u, v, depth) as stored in the checkpoint;  rots [B*cams, 9], trans [B*cams, 3]
 *   bev     [B, ny, nx, nz*64] (zeroed here), channel = z*64 + c; cell = trunc((R (u d, v d, d) + t - lo) / dx)
 * Cell indices are bit-exact w.r.t. the oracle; per-cell sums are atomics (fp32 addition order is not fixed). */
int ff3d_lss_splat(const float* dn, int ld, const float* frustum, const float* rots, const float* trans, float* bev,
                   int B, int cams, int D, int fH, int fW, const float* lo3, const float* dx3, int nx, int ny, int nz,
                   ff3d_stream_t stream);

/* Local-context attention of the 'bevfusion' FocalEncoderLayer (LiDAR + camera configs): replaces the JIT extension
 * locatt_ops (similar_forward -> softmax(. / sqrt(C)) -> weighting_forward; models/utils/ops/locatt_ops/kernels.cuh:4-80,
 * encoder_utils.py:155-163) in one pass.  q/k/v/y NHWC [B,H,W,C] views (batch-dense), C in {128, 256}, K odd (9).
 * Neighbours outside the map score 0 inside the soft-max and add no value, exactly like the reference's kernels. */
int ff3d_local_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* y, int ldy,
                         int B, int H, int W, int C, int K, ff3d_stream_t stream);

/* Misc elementwise helpers used by the head glue (all fp32). */
int ff3d_add_rows(const float* a, const float* b, float* y, long long n, ff3d_stream_t stream);
/* y[b, r, :] = a[b, r, :] + p[r, :]   (value + cached BEV positional embedding, focal_decoder.py:886) */
int ff3d_add_bcast_rows(const float* a, const float* p, float* y, int B, long long rows, int C, ff3d_stream_t stream);
/* the same sum written in split form only: ys [B*rows, 2C] fp16 [hi | lo] (input of the TMA-fed value_proj GEMM) */
int ff3d_add_bcast_rows_split(const float* a, const float* p, void* ys, int B, long long rows, int C, int* overflow_dev,
                              ff3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FF3D_H_ */
