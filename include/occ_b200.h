/* occ_b200.h -- C ABI of libocc_b200.so: the B200 (sm_100a) hot path of OccFormer.
 *
 * The reference has no C FFI on this path; its boundary is Python (mmcv Registry names, nn.Module forward
 * signatures, one pybind11 torch extension).  This header is what the reference-side Python binds with ctypes
 * (see INTEGRATION.md).  Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout; P/ = projects/mmdet3d_plugin/, M/ = mmdetection3d/mmdet3d/).
 *
 * Conventions (all functions):
 *   - pointers are DEVICE pointers to contiguous fp32 buffers unless stated; no torch types cross the ABI;
 *   - return 0 on success, <0 for an argument error (-1 invalid, -2 unsupported shape, -3 driver entry point
 *     missing), >0 = cudaError_t of a failed runtime call / launch;
 *   - never allocate, never synchronise, never exit; all work is enqueued on `stream`;
 *   - the current device must already be set by the caller (torch does this).
 *
 * Numerics: every tensor-core contraction is an fp32 problem computed as three bf16 tensor-core passes on hi/lo split
 * operands (a ~= hi + lo, hi = bf16(a), lo = bf16(a - hi); a*b ~= hi*hi' + lo*hi' + hi*lo', fp32 accumulate): ~1e-5
 * relative error.  Operands of such contractions are exchanged in the "S32" split format: an fp32-container tensor
 * (rows, C), C % 32 == 0, whose every aligned 32-column chunk (128 bytes) holds [32 bf16 hi | 32 bf16 lo] of its 32
 * values (same bytes and row pitch as the fp32 tensor).  occ_split_rows / occ_unsplit_rows convert.  Arguments
 * documented "S32" are in this format; everything else is plain fp32.
 */
#ifndef OCC_B200_H_
#define OCC_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* occ_stream_t; /* = cudaStream_t */

int occ_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * LSS voxel pooling.
 * Replaces: mmdet3d.ops.bev_pool.bev_pool(feats, coords, B, D, H, W)      M/ops/bev_pool/bev_pool.py:83-97
 *           bev_pool_ext.bev_pool_forward                                  M/ops/bev_pool/src/bev_pool.cpp:22-47
 *           ViewTransformerLiftSplatShootVoxel.voxel_pooling / lift        P/occformer/image2bev/ViewTransformerLSSVoxel.py:77-100,110-115
 * Output layout is channel-last (B, X, Y, Z, C); the Python wrapper returns permuted views carrying the
 * reference's shapes ((B,C,Z,X,Y) for bev_pool, (B,C,X,Y,Z) for voxel_pooling).
 * Bookkeeping left in the workspace after a call (byte offsets from occ_voxel_pool_workspace_layout):
 *   vox_id[n_points] int32 (linear voxel id ((b*X+x)*Y+y)*Z+z, -1 = dropped), counts[B*X*Y*Z] int32 (points per
 *   voxel = the reference's interval lengths), head[B*X*Y*Z] / next[n_points] int32 (per-voxel point lists, ids are
 *   1-based, 0 terminates).
 */
size_t occ_voxel_pool_workspace_bytes(int n_points, int B, int X, int Y, int Z);
int occ_voxel_pool_workspace_layout(int n_points, int B, int X, int Y, int Z, size_t* off_counts,
                                    size_t* off_head, size_t* off_next, size_t* off_vox_id);
/* get_geometry (ViewTransformerLSSBEVDepth.py:117-150): frustum (P = D*fH*fW, 3) + camera matrices -> geom
 * (B, N, P, 3) ego-frame points.  intrins (B,N,intrin_rows,intrin_cols): 3x3 (nuScenes), 3x4 or 4x4 (KITTI P2,
 * semantic_kitti_lss_dataset.py:62-64 builds a 4x4); bda (B,d,d), d 3|4. */
int occ_lss_geometry(const float* frustum, int P, const float* rots, const float* trans, const float* intrins,
                     int intrin_rows, int intrin_cols, const float* post_rots, const float* post_trans,
                     const float* bda, int bda_dim, int B, int N, float* geom, occ_stream_t stream);
/* depth softmax over D + NCHW->NHWC of the context features (ViewTransformerLSSVoxel.py:108-110):
 * depth_logits (BN, D, HW), img_feat (BN, C, HW) -> depth_prob (BN, D, HW), feat_cl (BN, HW, C) */
int occ_lift_prologue(const float* depth_logits, const float* img_feat, float* depth_prob, float* feat_cl, int BN,
                      int D, int C, int HW, occ_stream_t stream);
/* fused lift-splat: out[b,x,y,z,:] = sum_{p in voxel} depth_prob[p] * feat_cl[pixel(p), :]; geom (B*N*D*HW, 3);
 * dx/bx/nx = the view transformer's float parameters (ViewTransformerLSSBEVDepth.py:21-25,81-83).
 * out_split (optional, C % 32 == 0): the same grid in S32 (operand of the encoder's first conv). */
int occ_lift_splat(const float* depth_prob, const float* feat_cl, const float* geom, float* out, float* out_split,
                   int B, int N, int D, int HW, int C, float dx0, float dx1, float dx2, float bx0, float bx1, float bx2,
                   float nx0, float nx1, float nx2, int X, int Y, int Z, void* workspace, size_t workspace_bytes,
                   occ_stream_t stream);
/* The whole ViewTransformerLiftSplatShootVoxel.forward after DepthNet (ViewTransformerLSSVoxel.py:107-119) in three
 * launches: depth softmax + get_geometry + voxel index + per-voxel lists + NCHW->NHWC of the context features in one
 * front kernel (the (B,N,D,fH,fW,3) geometry tensor is never written), then the pooling kernel.  depth_logits
 * (B*N, D, HW) and img_feat (B*N, C, HW) with logits_stride / feat_stride floats between cameras (channel slices of
 * DepthNet's (B*N, D + C, fH, fW) output are passed without a copy), frustum (D*HW, 3), camera matrices as
 * occ_lss_geometry; outputs: depth_prob (B*N, D, HW), feat_cl (B*N, HW, C), out (B,X,Y,Z,C) fp32, out_split (optional) S32. */
int occ_lift_splat_fused(const float* depth_logits, long long logits_stride, const float* img_feat,
                         long long feat_stride, const float* frustum, const float* rots,
                         const float* trans, const float* intrins, int intrin_rows, int intrin_cols,
                         const float* post_rots, const float* post_trans, const float* bda, int bda_dim,
                         float* depth_prob, float* feat_cl, float* out, float* out_split, int B, int N, int D, int HW,
                         int C, float dx0, float dx1, float dx2, float bx0, float bx1, float bx2, float nx0, float nx1,
                         float nx2, int X, int Y, int Z, void* workspace, size_t workspace_bytes, occ_stream_t stream);
/* voxel_pooling(geom, volume) with a materialised volume: feats (B*points_per_batch, C), geom (same rows, 3)
 * (ViewTransformerLSSVoxel.py:77-100) */
int occ_voxel_pool_geom(const float* feats, const float* geom, float* out, int B, int points_per_batch, int C,
                        float dx0, float dx1, float dx2, float bx0, float bx1, float bx2, float nx0, float nx1,
                        float nx2, int X, int Y, int Z, void* workspace, size_t workspace_bytes, occ_stream_t stream);
/* drop-in bev_pool: feats (n, C), coords (n, 4) int64 (x, y, z, b) -> out (B, X, Y, Z, C) */
int occ_bev_pool(const float* feats, const long long* coords, float* out, int n, int C, int B, int X, int Y, int Z,
                 void* workspace, size_t workspace_bytes, occ_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * fp32-faithful tensor-core GEMM / implicit-GEMM convolution (tcgen05 bf16 x 3 passes + TMA), fused epilogues.
 * Replaces the cuBLAS / cuDNN calls issued by nn.Linear / nn.Conv3d / nn.Conv2d inside
 *   DualpathTransformerBlock  P/occformer/backbones/dualpath_block.py:36-48,79
 *   SwinBlock / WindowMSA     P/occformer/backbones/modules/window_attention.py:65-67,336-344
 *   BottleNeckASPP / ASPP     P/occformer/backbones/modules/aspp.py:49-172
 *   MSDeformAttnPixelDecoder3D  P/occformer/necks/multiscale_deformattn_3d.py:60-141 (input/lateral/output convs, Linears)
 *   Mask2Former*OccHead       P/occformer/mask2former/mask2former_nusc_occ.py:446-457 (mask einsum), decoder K/V proj
 * out[M,N] = act(A[M,K] W[N,K]^T + bias) (+ residual); A, W in S32 (K % 32 == 0); act: 0 none, 1 ReLU, 2 GELU(erf);
 * split_out: write out in S32 (N % 32 == 0) instead of fp32; bias / residual fp32.
 * gn_stats (optional): fp64 (sum, sumsq) per (batch, group) of the raw accumulator, cpg = channels per group. */
int occ_gemm_bf16x3(const float* A, const float* W, float* out, int M, int N, int K, const float* bias,
                    const float* residual, int act, int split_out, double* gn_stats, int cpg, int rows_per_batch,
                    occ_stream_t stream);
/* x (B,X,Y,Z,Cin) channel-last S32, w2 (Cout, KX*KY*KZ*Cin) tap-major S32, out (B,Xo,Yo,Zo,Cout); K in {1,3}, stride in
 * {1,2}, "same" padding dil*(K-1)/2.  2-D convs: Z = KZ = 1.  workspace (optional): occ_conv_workspace_bytes() bytes,
 * zero before its first use, one per stream -- the split-K ordering counters of the deep stages (without it the conv
 * runs single-pass). */
size_t occ_conv_workspace_bytes(void);
int occ_conv_bf16x3(const float* x, const float* w2, float* out, int B, int X, int Y, int Z, int Cin, int Cout, int KX,
                    int KY, int KZ, int stride, int dil, const float* bias, const float* residual, int act,
                    int split_out, double* gn_stats, int cpg, void* workspace, size_t workspace_bytes,
                    occ_stream_t stream);
/* Several same-input stride-1 convolutions in ONE launch via an explicit tap table: taps = ntaps x (dx, dy, dz) input offsets
 * (HOST ints), w2 (Cout, ntaps*Cin) S32 tap-major in the table's order (zero where a branch does not use a tap).  The four
 * ASPP branches (1x1 + three dilated 3x3, P/occformer/backbones/modules/aspp.py:107-113) become one 25-tap launch. */
int occ_conv_taps_bf16x3(const float* x, const float* w2, float* out, int B, int X, int Y, int Z, int Cin, int Cout,
                         int ntaps, const int* taps, double* gn_stats, int cpg, occ_stream_t stream);
/* Tensor-core passes per 32-k operand block, library-wide launch-time setting (host side; every launcher reads it when it
 * is called).  3 (default) = fp32-faithful: hi*hi + lo*hi + hi*lo, the mode of every parity claim against the reference.
 * 1 = single-pass bf16 operands (hi halves only): BASELINE config 5's "bf16" -- a third of the tensor work, ~3e-3
 * relative error per contraction; the reference has no twin of it, it is validated against the fp32 oracle at
 * its own tolerance.  occ_set_mma_passes returns the previous value, or -1 for an unsupported count. */
int occ_set_mma_passes(int passes);
int occ_get_mma_passes(void);
/* fp32 rows (rows, C) <-> S32 (C % 32 == 0) */
int occ_split_rows(const float* in, float* out, long long rows, int C, occ_stream_t stream);
int occ_unsplit_rows(const float* in, float* out, long long rows, int C, occ_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Dual-path encoder block glue (P/occformer/backbones/dualpath_block.py:65-82).
 * Token rows: voxel tokens ((b*X+x)*Y+y)*Z+z, then BEV tokens B*X*Y*Z + (b*X+x)*Y+y. */
/* GroupNorm+ReLU of the raw conv output, Z-mean, LayerNorm1 (dualpath_block.py:43-48,69; window_attention.py:355):
 * tok fp32 (residual), tokn S32 (operand of the QKV GEMM).
 * win_shift < 0: tokn rows in token order;  win_shift = 0 / 1 (needs X, XY % X == 0): tokn in the WINDOW LAYOUT of the
 * un-shifted / shifted 7x7 partition -- token (img, x, y) at row window*64 + t with window = (wx*nWy + wy)*n_images + img,
 * occ_window_layout_rows() rows in total,
 * zero-initialised once by the caller (pad positions are never written) -- the operand of occ_swin_qkv_attention. */
int occ_gn_relu_zmean_ln(const float* y, const double* stats, const float* gn_w, const float* gn_b,
                         const float* ln_w, const float* ln_b, float* tok, float* tokn, int B, int XY, int Z, int C,
                         int groups, int X, int win_shift, occ_stream_t stream);
long long occ_window_layout_rows(int B, int X, int Y, int Z);
/* GroupNorm statistics from a conv / GEMM epilogue at a finer power-of-two grouping -> the module's groups:
 * out (B, groups_out, 2) = sums of `factor` consecutive entries of in (B, groups_out * factor, 2), fp64. */
int occ_stats_regroup(const double* in, double* out, int B, int groups_out, int factor, occ_stream_t stream);
int occ_layernorm(const float* in, const float* w, const float* b, float* out, long long rows, int C, int split_out,
                  occ_stream_t stream);
/* o = act(gn(in[row, c])) (+ residual[row, c]) -- ASPP / neck norms (aspp.py:42-46,117-120,166-172): out[row, c] = o
 * (fp32, optional) and / or out_split[row, out_off + c] = o in S32 with row pitch ldo (optional) */
int occ_gn_apply(const float* in, const double* stats, const float* w, const float* b, const float* residual,
                 float* out, float* out_split, long long rows, int rows_per_batch, int C, int groups, int ldo,
                 int out_off, int relu, occ_stream_t stream);
/* ASPP image-pooling branch: GAP -> 1x1 conv -> GN -> ReLU -> broadcast (aspp.py:89-95,113-114).
 * sums_ws: workspace of B*ch doubles followed by B*ch floats. */
int occ_aspp_gap_branch(const float* in, double* sums_ws, const float* wconv, const float* gw, const float* gb,
                        float* cat, int B, int rows_per_batch, int ch, int groups, int ldo, int out_off,
                        occ_stream_t stream);
/* coeff = sigmoid(<x,w>+b); out = x + coeff * bev + identity (identity optionally GroupNorm'd: strided skip path)
 * dualpath_block.py:36-41,79-82 */
int occ_dualpath_fuse(const float* x, const float* bev, const float* cw, float cbias, const float* identity,
                      int identity_split /*identity is S32*/, const double* id_stats, const float* id_w,
                      const float* id_b, int groups, float* out /*fp32, optional*/, float* out_split /*S32, optional*/,
                      int B, int XY, int Z, int C, occ_stream_t stream);
/* Swin block tail for C == 128 in one tensor-core kernel (csrc/swin_mlp_fused.cu): y1 = tok + att Wp^T + bp (WindowMSA.proj
 * + the ShiftWindowMSA residual, window_attention.py:105-107 / swin.py SwinBlock.forward), out = y1 + W2 GELU(W1 LN2(y1)
 * + b1) + b2 (norm2 + FFN + residual).  att and the weights ((out, in) row-major) in S32; tok / out fp32.  Returns -2 for
 * C != 128 (the caller then uses occ_gemm_bf16x3 / occ_layernorm). */
int occ_swin_proj_ffn(const float* att, const float* tok, const float* wp, const float* bp, const float* ln_w,
                      const float* ln_b, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                      long long M, int C, occ_stream_t stream);
/* (shifted) 7x7 window attention core over B*(Z+1) images: ShiftWindowMSA.forward + WindowMSA.forward
 * (window_attention.py:168-242, 69-107) minus the qkv / proj linears, on tcgen05 tensor cores.
 * bias_pad = relative_position_bias_table[relative_position_index] as (heads, 49*49 padded to 2404 floats).
 * qkv_head_major = 0: qkv / qkv_bias columns in the reference order [q|k|v][head][32] (WindowMSA.qkv);  1: [head][q|k|v][32]
 * (the caller permuted the rows of the qkv weight: one contiguous 384-byte run per (token, head)). */
int occ_window_attention(const float* qkv /*S32*/, const float* qkv_bias /*S32 row*/, const float* bias_pad,
                         float* out /*S32*/, int B, int X, int Y, int Z, int C, int heads, int shift,
                         int qkv_head_major, occ_stream_t stream);

/* The same attention with the QKV projection fused in (C == 128; csrc/swin_attn_fused.cu): tokn
 * (occ_window_layout_rows(B,X,Y,Z), 128) S32 = LayerNorm1'ed tokens in the window layout of THIS shift (see
 * occ_gn_relu_zmean_ln), out (B*X*Y*(Z+1), 128) S32 in token order; wqkv (384, 128) S32 / bqkv (384) fp32 = WindowMSA.qkv with HEAD-MAJOR rows [head][q|k|v][32]; the
 * q/k/v tensor never exists in HBM.  Returns -2 for C != 128 (use occ_gemm_bf16x3 + occ_window_attention). */
int occ_swin_qkv_attention(const float* tokn, const float* wqkv, const float* bqkv, const float* bias_pad, float* out,
                           int B, int X, int Y, int Z, int C, int heads, int shift, occ_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * MSDeformAttnPixelDecoder3D, the neck between encoder and head (P/occformer/necks/multiscale_deformattn_3d.py:143-248,
 * P/occformer/necks/multi_scale_deform_attn_3d.py:17-80,185-286).  Token rows are "level-major":
 * row(l, b, x, y, z) = B*start_l + b*n_l + (x*Y_l + y)*Z_l + z with the L levels ordered coarse -> fine (start_l = tokens
 * of the coarser levels of one sample), so every level is a contiguous channel-last (B, X_l, Y_l, Z_l, C) tensor.
 * `grids` = L x (X, Y, Z) ints in HOST memory.  The Linears / 1x1x1 / 3x3x3 convolutions go through occ_gemm_bf16x3 /
 * occ_conv_bf16x3. */
/* [LayerNorm (norms.k of the encoder layer) of each row ->] out_f32 = x (fp32), out_s32 = x (S32), out_pos = x + pos[token]
 * (S32; pos (Nq, C) = SinePositionalEncoding3D + level_encoding of the token's level: query_pos of the reference); any of
 * the three outputs may be NULL; ln_w / ln_b NULL = no normalisation. */
int occ_neck_token_prep(const float* in, const float* ln_w, const float* ln_b, const float* pos, float* out_f32,
                        float* out_s32, float* out_pos, int L, int B, const int* grids, int C, occ_stream_t stream);
/* multi_scale_deformable_attn_pytorch + the sampling-location arithmetic of MultiScaleDeformableAttention3D.forward:
 * value (rows, value_ld) = value_proj output, head h at columns [h*head_ld, h*head_ld + E/H) (head_ld = E/H, or E/H
 * rounded up to 32 floats so that a head slice is one 128-byte line); ow (rows, H*L*P*4) = [sampling_offsets (h,l,p,(z,y,x)) | attention logits (h,l,p)];
 * out (rows, E) S32 = sum_{l,p} softmax(logits) * trilinear(value level l)(ref + offset / (Z_l,Y_l,X_l)), zeros outside,
 * align_corners=False; strides = the L feature strides (HOST floats; reference-point arithmetic). */
int occ_ms_deform_attn(const float* value, int value_ld, int head_ld, const float* ow, float* out, int L, int B,
                       const int* grids, const float* strides, int E, int H, int P, occ_stream_t stream);
/* FPN step (:228-240): out_s (S32) = GroupNorm(cur raw lateral-conv output, stats) + trilinear upsample
 * (align_corners=False) of coarse (B, Xc, Yc, Zc, C) fp32 */
int occ_gn_upsample_add(const float* cur, const double* stats, const float* gw, const float* gb, int groups,
                        const float* coarse, float* out_s, int B, int X, int Y, int Z, int Xc, int Yc, int Zc, int C,
                        occ_stream_t stream);
/* GroupNorm statistics of a finished tensor x (B, rows_per_batch, C): stats (B, C/cpg, 2) fp64 += (sum, sumsq); for the
 * group sizes the conv epilogue does not accumulate itself (cpg not a power of two: 192 channels / 32 groups) */
int occ_gn_stats(const float* x, double* stats, int B, int rows_per_batch, int C, int cpg, occ_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Mask2Former-3D occupancy decoder head (P/occformer/mask2former/mask2former_nusc_occ.py, mask2former_occ.py).
 * Layouts: voxel memories channel-last (B, S, E); queries (B, Q, E); mask logits "query-last" (B, S, Q).
 * The voxel-side GEMMs (K/V projections :657-667 via nn.MultiheadAttention in_proj; mask einsum :457) go through
 * occ_gemm_bf16x3; the entry points below are the fused query-side / reduction kernels. */
/* SinePositionalEncoding3D.forward, normalize=True, all-False mask (positional_encoding.py:58-108): out (X*Y*Z, 3F) */
int occ_sine_pos3d(float* out, int X, int Y, int Z, int num_feats, float temperature, float scale, float eps,
                   float offset, occ_stream_t stream);
/* level prep (mask2former_nusc_occ.py:614-630): mem = in + level_embed, kpos = in + level_embed + pos, both S32;
 * in is channel-last (B,S,C) or the reference layout (B,C,S); level_embed / pos+kpos optional. */
int occ_head_prep(const float* in, int in_channel_last, const float* level_embed, const float* pos, float* mem,
                  float* kpos, int B, long long S, int C, occ_stream_t stream);
/* forward_head query side (:446-455): [optional: query = LN(norms.2)(query_in) -> query_state] -> post_norm LN ->
 * cls_embed -> cls_out (rows, NC); mask_embed MLP -> membed_out (rows, E) S32; [optional: the NEXT layer's
 * cross-attention query projection qh_out = ((query + query_pos) Wq^T + bq) * scale].  Weights K-major (in, out). */
int occ_query_head(const float* query_in, const float* n2w, const float* n2b, float* query_state, const float* pn_w,
                   const float* pn_b, const float* clsT, const float* cls_b, int NC, const float* m0T, const float* m0b,
                   const float* m1T, const float* m1b, const float* m2T, const float* m2b, float* cls_out,
                   float* membed_out, const float* query_pos, int Q, const float* wqT, const float* bq, float scale,
                   float* qh_out, const float* ffn_part /*(nparts, rows, E) FFN column-block partials, added to query_in
                   in the order 0..nparts-1; may be NULL with nparts = 0*/, int nparts, int rows, int E, occ_stream_t stream);
/* adaptive_max_pool3d of the mask logits (:463), general windows -> pooled (B, Xo*Yo*Zo, Q) as order-preserving ints
 * (see occ_mask_gemm_pool); row_flag[b*Q+q] = 1 iff some key of the row is un-blocked (pooled >= 0), else the row
 * attends everywhere (:652-653).  attn_mask == pooled < 0. */
int occ_mask_pool(const float* mask, int* pooled, int* row_flag, int B, int X, int Y, int Z, int Xo, int Yo, int Zo,
                  int Q, occ_stream_t stream);
/* mask einsum with fused attention-mask pooling (:457-466) when the pooling windows are powers of two >= 2 dividing the
 * grid: mask[b,v,q] = <mf[b,v,:], membed[b,q,:]> (written only if mask_out != NULL), pooled[b,cell,q] = window max as an
 * order-preserving int (i >= 0 ? i : i ^ 0x7FFFFFFF of the float bits; blocked <=> value < 0), flag[b,q] as occ_mask_pool */
int occ_mask_gemm_pool(const float* mf /*S32*/, const float* membed /*S32*/, float* mask_out, int* pooled, int* flag,
                       int B, int X, int Y, int Z, int E, int Q, int Xo, int Yo, int Zo, occ_stream_t stream);
/* masked cross attention (mmcv MultiheadAttention -> nn.MultiheadAttention, bool attn_mask) on the tcgen05 tensor cores
 * (128-key tiles, TMA operand loads, P in TMEM); qh fp32, Kp / Vp S32 (one chunk per (key, head)); writes
 * occ_cross_attn_tc_partials(S) partials per (sample, head): part (B, H, npart, Q, 34) = running max, running sum, 32
 * value accumulators */
int occ_cross_attn_tc_partials(int S);
/* pooled (B,S,Q) ordered-int mask logits -> bits (B, 4*ceil(S/128), Q): one "blocked" bit per (key, query) */
int occ_mask_bits(const int* pooled, unsigned* bits, int B, int S, int Q, occ_stream_t stream);
int occ_cross_attn_tc(const float* qh, const float* Kp, const float* Vp, int ld, int koff, int voff, const unsigned* bits,
                      const int* row_flag, float* part, int B, int S, int Q, int E, int H, occ_stream_t stream);
/* merge partials -> out_proj -> +identity -> LN(norms.0) -> query1; self-attention in_proj -> sa_qkv (rows, 3E) */
int occ_cross_merge(const float* part, int nchunk, int H, const float* query, const float* query_pos, int Q,
                    const float* woT, const float* bo, const float* n0w, const float* n0b, const float* sa_inT,
                    const float* sa_inb, float scale, float* query1, float* sa_qkv, int rows, int E,
                    occ_stream_t stream);
/* self attention over the Q queries -> out_proj -> +identity -> LN(norms.1) -> x1; FFN(ReLU) + identity accumulated
 * as F/E column blocks: ybuf = x1 + b2, ffn_part (F/E, rows, E) = the blocks' contributions; occ_query_head adds them in
 * a fixed order (bit-reproducible, no floating-point atomics) and applies the closing LN(norms.2). */
int occ_self_attn_ffn(const float* sa_qkv, const float* query1, int Q, const float* woT, const float* bo,
                      const float* n1w, const float* n1b, const float* f1T, const float* f1b, const float* f2T,
                      const float* f2b, int F, float* x1, float* ybuf, float* ffn_part, int rows, int E, int H,
                      occ_stream_t stream);
/* simple_test tail (:725-736, format_results :691-696): trilinear upsample (align_corners=True) -> sigmoid ->
 * einsum with softmax(cls)[..., :-1]; mask (B, X*Y*Z, Q), cls (B, Q, NC) -> out (B, NC-1, Xo, Yo, Zo); labels (optional,
 * (B, Xo, Yo, Zo) uint8) = argmax over the class axis (post_process_semantic, occupancyformer.py:238-243) */
int occ_classmix(const float* mask, const float* cls, float* out, unsigned char* labels, int B, int X, int Y, int Z,
                 int Xo, int Yo, int Zo, int Q, int NC, occ_stream_t stream);
/* (B, S, Q) -> (B, Q, S): the reference's mask_pred layout, for forward()'s return value */
int occ_transpose_sq(const float* in, float* out, int B, long long S, int Q, occ_stream_t stream);
/* forward_lidarseg eval branch (:505-542): grid_sample of one sample's class volume (K,X,Y,Z) at n points
 * (rows of pts_stride floats, xyz first) + softmax -> out (n, K) */
int occ_lidarseg_points(const float* vox, const float* pts, int pts_stride, int n, float x_min, float y_min,
                        float z_min, float x_max, float y_max, float z_max, int X, int Y, int Z, int K, int border,
                        float* out, occ_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Evaluation tail: the integer counts behind the path's single collective
 * (P/occformer/apis/test.py:195-212 sums them over ranks). */
/* SSCMetrics.get_score_completion + get_score_semantic_and_completion (P/utils/ssc_metric.py:104-168), masked by
 * target != ignore: pred / target n uint8 voxel labels -> out[0..2] = completion tp, fp, fn; out[3 + c], out[3 + K + c],
 * out[3 + 2K + c] = semantic tp, fp, fn of class c (3 + 3K int64).  conf_ws: K*K int64 scratch. */
int occ_ssc_counts(const unsigned char* pred, const unsigned char* target, long long n, int K, int ignore,
                   long long* conf_ws, long long* out, occ_stream_t stream);
/* OccupancyFormer.simple_evaluation_semantic (P/occformer/detectors/occupancyformer.py:219-224,246-254) + fast_hist_crop
 * (P/utils/metric_util.py:8-23): scores (n, K) point class scores, labels (n) int64 -> hist[(gt-1)*(K-1) + (pred-1)] += 1
 * for gt in 1..K-1, pred = 1 + argmax(scores[:, 1:]); hist ((K-1)^2 int64) is accumulated. */
int occ_lidarseg_hist(const float* scores, const long long* labels, int n, int K, long long* hist, occ_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OCC_B200_H_ */
