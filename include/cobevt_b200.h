/*
 * cobevt_b200.h -- C ABI of libcobevt_sm100.so (sm_100a only).
 *
 * The reference (DerrickXuNu/CoBEVT) has no FFI layer: its operator API is the
 * PyTorch nn.Module surface.  These entry points are what the Python mirror
 * modules in cobevt_b200/ bind through ctypes; each one names the reference
 * code it replaces (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers
 *     unless stated; activations are NHWC ("tokens x channels") 16-bit floats
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they
 *     never synchronise and never allocate device memory
 *   - return 0 on success, <0 on error: -1 bad shape/argument, -2 unsupported
 *     configuration, -3 CUDA error; cobevt_last_error() returns the message
 *   - dtype: 0 = fp16, 1 = bf16 (16-bit storage, fp32 accumulation)
 */
#ifndef COBEVT_B200_H_
#define COBEVT_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COBEVT_OK 0
#define COBEVT_ERR_SHAPE (-1)
#define COBEVT_ERR_UNSUPPORTED (-2)
#define COBEVT_ERR_CUDA (-3)

enum { COBEVT_F16 = 0, COBEVT_BF16 = 1 };
enum { COBEVT_ACT_NONE = 0, COBEVT_ACT_RELU = 1, COBEVT_ACT_GELU = 2, COBEVT_ACT_SWISH = 3 };

int cobevt_version(void);
const char* cobevt_last_error(void);
/* sizeof of a parameter struct in this build: 0 CobevtConvParams, 1 CobevtAttnParams, 2 CobevtMlpParams, 3 CobevtTokenMap
 * (-1 otherwise).  Bindings check their own declarations against it once, at load time (cobevt_b200/_lib.py does). */
int32_t cobevt_struct_bytes(int32_t which);

/* ---------------------------------------------------------------------------
 * Implicit-GEMM convolution / token GEMM on tcgen05 tensor cores.
 *   y[m, co] = act( sum_{ky,kx,ci} x[pix(m)+(ky,kx), ci] * w[co, ky, kx, ci]
 *                   + bias[co] + res[m, co] )
 * Replaces: torchvision ResNet convs + BN + ReLU + residual used by
 *   opv2v/opencood/models/backbones/resnet_ms.py:46-89, the Bottleneck /
 *   downsample convs of sub_modules/fax_modules.py:472-489, every nn.Linear on
 *   the token path (fax_modules.py:189-193,311-312; swap_fusion_modules.py:45-53;
 *   base_transformer.py:112-124), naive_decoder.py:62-91, bev_seg_head.py:35-61.
 * A 1x1 "conv" over a (rows,1,1,C) tensor is a plain GEMM.
 * -------------------------------------------------------------------------*/
typedef struct {
  const void* x;       /* (N, Hin, Win, Cin) 16-bit, Cin % 8 == 0                     */
  const void* w;       /* packed by cobevt_conv_pack_weight                           */
  const float* bias;   /* [Cout] fp32 or NULL                                         */
  const void* res;     /* (M, Cout) 16-bit residual added before act, or NULL         */
  const float* res32;  /* (M, Cout) fp32 residual, or NULL                            */
  void* y;             /* (M, Cout) 16-bit output or NULL                             */
  float* y32;          /* (M, Cout) fp32 output or NULL                               */
  void* y_ln;          /* (M, Cout) 16-bit LayerNorm of the output row, or NULL; requires
                          Cout == block_n in {32, 64, 128}                            */
  const float* ln_gamma; /* [Cout] affine of y_ln, or NULL for the plain normalised row
                          (consumers fold their own LayerNorm affine into their weights) */
  const float* ln_beta;
  int32_t N, Hin, Win, Cin; /* physical input extents                                  */
  int32_t OH, OW, Cout;     /* M = N*OH*OW                                             */
  int32_t KH, KW;
  int32_t stride_h, stride_w, pad_h, pad_w;
  int32_t res_rows;    /* >0: residual row index is (m % res_rows) (broadcast over batch)  */
  int32_t in_upsample; /* 1: logical input is the physical one nearest-upsampled x2    */
  int32_t act;         /* COBEVT_ACT_*                                                 */
  int32_t dtype;       /* COBEVT_F16 / COBEVT_BF16                                     */
  int32_t block_n;     /* N tile the weights were packed for: 16,32,64,128,256         */
  float ln_eps;
} CobevtConvParams;

/* bytes of the packed weight buffer for (Cout, KH*KW*Cin, block_n) */
int64_t cobevt_conv_packed_bytes(int32_t Cout, int32_t K, int32_t block_n);
/* HOST-side pack: w_host is [Cout][K] row-major 16-bit (K = KH*KW*Cin, ci fastest);
 * out_host receives the tile images the kernel bulk-copies (SWIZZLE_128B K-major). */
int cobevt_conv_pack_weight(const uint16_t* w_host, int32_t Cout, int32_t K, int32_t block_n,
                            uint16_t* out_host);
int cobevt_conv2d_nhwc_fwd(const CobevtConvParams* p, void* stream);

/* ---------------------------------------------------------------------------
 * Fused token MLP (PreNorm feed-forward) on tcgen05, the 2x-wide GELU activations never leave the SM:
 *   y = res + W2 gelu(W1 x_ln + b1) + b2 ;  y_ln = LayerNorm(y) [* gamma + beta]
 * Replaces mlp_1 / mlp_2 with prenorm_1/2 and postnorm of CrossViewSwapAttention (fax_modules.py:309-313,411,
 * 435-437) and PreNormResidual(FeedForward) of the FuseBEVT blocks (base_transformer.py:102-124,
 * swap_fusion_modules.py:144-163).  x_ln is the plain normalised row produced by the preceding kernel's y_ln output
 * (the LayerNorm affine is folded into w1 / b1).  Built for C = 128, H = 256.
 * -------------------------------------------------------------------------*/
typedef struct {
  const void* x_ln;      /* (M, C) 16-bit                                                         */
  const void* w1;        /* (H, C) packed by cobevt_conv_pack_weight with block_n = 128           */
  const float* b1;       /* [H] fp32                                                              */
  const void* w2;        /* (C, H) packed with block_n = 128                                      */
  const float* b2;       /* [C] fp32                                                              */
  const void* res;       /* (M, C) 16-bit residual stream, or NULL                                */
  void* y;               /* (M, C) 16-bit, or NULL                                                */
  void* y_ln;            /* (M, C) 16-bit LayerNorm of the output row, or NULL                    */
  const float* ln_gamma; /* [C] affine of y_ln or NULL (plain normalised row)                     */
  const float* ln_beta;
  int32_t M, C, H, dtype;
  float ln_eps;
} CobevtMlpParams;
int cobevt_token_mlp_fwd(const CobevtMlpParams* p, void* stream);

/* ---------------------------------------------------------------------------
 * Windowed / dilated-grid multi-head attention (head_dim 32) on tcgen05:
 *   S = scale * Q K^T (+ bias) (+ key mask) ; P = softmax(S) ; O = P V
 * Q, K, V are already projected token maps; window / grid partitions are pure
 * addressing (no rearranged copies).
 * Replaces: CrossWinAttention.forward fax_modules.py:198-248 (core :226-234) with
 *   the partitions of CrossViewSwapAttention.forward :399-433; Attention.forward
 *   fax_modules.py:132-176; swap_fusion_modules.Attention.forward :87-128.
 * -------------------------------------------------------------------------*/
typedef struct {
  int32_t views;      /* token planes per batch row (cameras / agents)                */
  int32_t H, W;       /* logical (padded) map extent the partition is defined on      */
  int32_t Hv, Wv;     /* valid extent stored in memory (<= H, W); outside = pad token */
  int32_t w1, w2;     /* window extent                                                */
  int32_t grid;       /* 0: window partition '(x w1) (y w2)'; 1: grid '(w1 x) (w2 y)' */
} CobevtTokenMap;

typedef struct {
  const void* q;      /* (B, q.views, q.Hv, q.Wv, :) 16-bit, head h at channels [32h, 32h+32) */
  const void* k;      /* (B, kv.views, kv.Hv, kv.Wv, C)                              */
  const void* v;
  void* out;          /* q_mean ? (B, q.Hv, q.Wv, C) : (B, q.views, q.Hv, q.Wv, C)   */
  const void* k_pad;  /* [C] 16-bit token used for padded key positions, or NULL     */
  const void* v_pad;
  const float* bias;  /* (heads, Lq, Lk) fp32 additive bias or NULL; Lk counts the keys of ALL kv_views_total
                         views of the window (a key-segment call adds its own slice of every row)     */
  const uint8_t* key_mask; /* (B, kv.Hv, kv.Wv, kv.views) 1 = keep, 0 = -inf; or NULL */
  CobevtTokenMap qmap, kvmap;
  int32_t B, heads;
  int32_t q_stride, kv_stride, out_stride; /* token strides in elements (>= heads*32);
                         lets q/k/v be column slices of one fused projection output    */
  int32_t q_mean;     /* 1: q views are separate passes averaged (CrossWinAttention
                         mean over n, fax_modules.py:243); 0: views are extra rows   */
  float scale;
  int32_t dtype;
  /* --- key-segment mode (windows whose keys do not fit one resident pass, e.g. the stage-2 global
   *     cross attention over 6 x 420 keys): the caller runs one call per group of K/V views and merges */
  float* part;            /* NULL, or fp32 partial results of THIS call: (B, windows, heads, Lq, 36) =
                             32 un-normalised output channels, running max (log2 domain), running sum, 2 pad;
                             `out` is not written, q views are plain rows (q_mean is applied by the merge) */
  int32_t kv_view0;       /* first view of the K/V tensors used by this call (kvmap.views = views used)   */
  int32_t kv_views_total; /* views of the K/V tensors, i.e. their batch stride; 0: kvmap.views            */
  int32_t kv_segments;    /* 0/1: one segment; n: ONE launch covers n consecutive segments of kvmap.views views
                             each (starting at kv_view0), segment s writing to part + s * part_segment_stride */
  int64_t part_segment_stride; /* in floats */
  int32_t bias_layout;    /* 0: bias is (heads, Lq, Lk) row-major.  1: key-group-major (heads, Lk/4, Lq, 4): the 32 query rows
                             of a warp read 512 contiguous bytes per key group instead of 32 scattered 16-byte pieces
                             (resident-key kernel only: Lk per call <= cobevt_attn_resident_keys(), Lk % 4 == 0)        */
} CobevtAttnParams;

int cobevt_window_attn_fwd(const CobevtAttnParams* p, void* stream);

/* Bytes of the caller-owned fp32 scratch `part` a key-segment call needs (kv_segments > 1), 0 otherwise, < 0 on a bad
 * query map: the library never allocates, callers size the workspace with this query. */
int64_t cobevt_attn_workspace_bytes(const CobevtAttnParams* p);

/* Largest number of keys one call keeps resident per window on the four-warpgroup kernel (longer multi-view
 * windows: key-segment mode above). */
int32_t cobevt_attn_resident_keys(void);

/* Merge `segments` partial results (consecutive at `segment_stride` floats) into the 16-bit output of the
 * attention they partition: log-sum-exp across segments per query token, then the mean over query views when
 * q_mean (fax_modules.py:243).  qmap / out layout as in cobevt_window_attn_fwd. */
int cobevt_attn_merge(const float* part, int32_t segments, int64_t segment_stride, void* out,
                      const CobevtTokenMap* qmap, int32_t B, int32_t heads, int32_t q_mean,
                      int32_t out_stride, int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------
 * HBM-bound element-wise / gather kernels (dtype: 0 f16, 1 bf16; permute also 2 = f32)
 * -------------------------------------------------------------------------*/
/* (npix,3) fp32 RGB -> (npix,8) 16-bit, channels 3..7 zero.  Replaces the
 * .permute(0,3,1,2).contiguous() copy of resnet_ms.py:62-65 (the stem reads NHWC8). */
int cobevt_pack_rgb8(const float* x, void* y, int64_t npix, int32_t dtype, void* stream);
/* (N,H,W,3) fp32 RGB -> (N,Hp,Wp,4) 16-bit, image placed at (pad_top, pad_left), zeros elsewhere:
 * input staging of the 7x7/s2 stem (resnet_ms.py:62-67) -- no bounds checks, 16-byte aligned taps */
int cobevt_pack_rgb4_padded(const float* x, void* y, int32_t N, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                            int32_t pad_top, int32_t pad_left, int32_t src_nchw, const float* mean3_host,
                            const float* std3_host, int32_t dtype, void* stream);
/* nn.Upsample(x2, bilinear, align_corners=True) on NHWC 16-bit; nuscenes .../decoder.py:13 */
int cobevt_upsample_bilinear2x(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t dtype,
                               void* stream);
/* y = relu(y + nearest_resize(up -> HxW)) in place: DecoderBlock.forward, nuscenes .../decoder.py:28-36 */
int cobevt_add_nearest_relu(void* y, const void* up, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w,
                            int32_t C, int32_t dtype, void* stream);
/* batched fp32 3x3 / 4x4 inverse (Gauss-Jordan, partial pivoting): intrinsics / extrinsics .inverse(),
 * nuscenes .../encoder_pyramid_axial.py:538-539 */
int cobevt_invert_matrices(const float* src, float* dst, int32_t n, int32_t dim, void* stream);
/* torchvision resnet maxpool (3x3, stride 2, pad 1) on NHWC, resnet_ms.py:71 */
int cobevt_maxpool3x3s2_nhwc(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C,
                             int32_t dtype, void* stream);
/* ya = relu(x*sa+ba), yb = relu(x*sb+bb): the eval BatchNorm+ReLU in front of feature_proj and
 * feature_linear, fax_modules.py:281-292,379-384 (yb may be NULL) */
int cobevt_bn_relu2(const void* x, const float* sa, const float* ba, void* ya, const float* sb,
                    const float* bb, void* yb, int64_t rows, int32_t C, int32_t dtype, void* stream);
/* normalised ray-direction embedding img_embed, fax_modules.py:346-358.
 * I (NV,3,3) intrinsics (inverted in-kernel when invert_intrinsic, :500-501) or their inverse,
 * E_inv (NV,4,4), w_img / w_cam (C,4) fp32 -> out (NV,h,w,C) */
int cobevt_ray_embed(const float* I, const float* E_inv, const float* w_img, const float* w_cam,
                     void* out, int32_t NV, int32_t h, int32_t w, float image_h, float image_w,
                     int32_t C, int32_t invert_intrinsic, int32_t dtype, void* stream);
/* out[b,n,p,:] = LayerNorm0( x[(b*P+p) % x_rows,:] + normalise(w_bev world_p + b_bev - w_cam c_{b,n}) )
 * fax_modules.py:370-375,387-390 + the LayerNorm of to_q (:189); world (2,P) fp32 */
int cobevt_bev_query(const void* x, int64_t x_rows, const float* world, const float* w_bev,
                     const float* b_bev, const float* w_cam, const float* E_inv, void* out, int32_t BL,
                     int32_t n, int32_t P, int32_t C, float eps, int32_t dtype, void* stream);
/* row LayerNorm (gamma/beta NULL -> plain normalisation) */
int cobevt_layernorm_rows(const void* x, const float* gamma, const float* beta, void* y, int64_t rows,
                          int32_t C, float eps, int32_t dtype, void* stream);
/* x (B,L,P,C) -> LayerNorm0(mean over L) (B,P,C): head of SwapFusionEncoder,
 * swap_fusion_modules.py:275-278 */
int cobevt_mean_views_ln(const void* x, void* y, int32_t B, int32_t L, int32_t P, int32_t C, float eps,
                         int32_t dtype, void* stream);
/* regroup (fuse_utils.py:8-61) + STTF bilinear warp (corpbevt.py:28-64) + ROI/cav mask
 * (torch_transformation_utils.py:11-49) in one pass, no host sync:
 * x (sum L, H, W, C), tm (B,L,4,4) fp32, record_len (B) int64 (device) ->
 * out (B,L,H,W,C), mask (B,H,W,L) uint8 */
int cobevt_sttf_regroup(const void* x, const float* tm, const int64_t* record_len, void* out,
                        uint8_t* mask, int32_t B, int32_t L, int32_t H, int32_t W, int32_t C,
                        float discrete_ratio, float downsample_rate, int32_t use_roi_mask,
                        int32_t dtype, void* stream);
/* uint8 camera frames (N,H,W,3) -> (N,Hp,Wp,4) 16-bit RGBx with a zero border, fusing the RgbPreProcessor
 * arithmetic of opv2v/opencood/data_utils/pre_processor/rgb_preprocessor.py:16-36 (optional BGR->RGB :38-46,
 * x/255 :33-34, (x-mean)/std :24-31) into the stem's input staging.  mean3/std3: HOST pointers
 * (3 floats) or NULL. */
int cobevt_stage_rgb_u8(const uint8_t* x, void* y, int32_t N, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                        int32_t pad_top, int32_t pad_left, int32_t bgr, const float* mean3_host,
                        const float* std3_host, int32_t dtype, void* stream);
/* prob = softmax(seg, dim=1), amap = argmax(prob, dim=1) for seg (N,C,HW) fp32 NCHW: softmax_argmax of
 * opv2v/opencood/data_utils/post_processor/camera_bev_postprocessor.py:55-59 (used by post_process_train :61-87) */
int cobevt_softmax_argmax(const float* seg, float* prob, int64_t* amap, int32_t N, int32_t C, int64_t HW, void* stream);
/* F.interpolate(x, scale_factor=s, recompute_scale_factor=False) in its default 'nearest' mode on NHWC 16-bit
 * (nuscenes .../encoder_pyramid_axial.py:492-493): src index = min(floor(dst * src_per_dst), in - 1) */
int cobevt_resize_nearest(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t OH, int32_t OW,
                          float src_per_dst_h, float src_per_dst_w, void* stream);
/* ---- MBConv pieces of the EfficientNet-B4 prefix (nuscenes .../backbones/efficientnet.py:56-69; arithmetic =
 * efficientnet_pytorch 0.7.1 MBConvBlock, package absent from the reference tree: parity unpinned) ------------- */
/* y = swish(depthwise KxK conv(x; w) + bias): NHWC 16-bit, w (K,K,C) fp32 with BN folded, asymmetric "static same"
 * padding (pad_top/pad_left before, the rest implied by OH/OW); sums (N,C) fp32 or NULL: += sum over pixels of y
 * (the squeeze of squeeze-excite; caller zeroes it) */
int cobevt_dwconv_swish(const void* x, const float* w, const float* bias, void* y, float* sums, int32_t N, int32_t H,
                        int32_t W, int32_t C, int32_t K, int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH,
                        int32_t OW, int32_t dtype, void* stream);
/* gate (N,C) = sigmoid(W2 swish(W1 (sums * inv_hw) + b1) + b2); w1 (Cse,C), w2 (C,Cse) fp32 */
int cobevt_se_gate(const float* sums, float inv_hw, const float* w1, const float* b1, const float* w2,
                   const float* b2, float* gate, int32_t N, int32_t C, int32_t Cse, void* stream);
/* x[n,p,c] *= gate[n,c] with the gate of cobevt_se_gate recomputed per block (one launch instead of two) */
int cobevt_se_scale(void* x, const float* sums, float inv_hw, const float* w1, const float* b1, const float* w2,
                    const float* b2, int32_t N, int64_t HW, int32_t C, int32_t Cse, int32_t dtype, void* stream);
/* y[n,p,c] = x[n,p,c] * gate[n,c] (in place allowed) */
int cobevt_scale_channels(const void* x, const float* gate, void* y, int32_t N, int64_t HW, int32_t C, int32_t dtype,
                          void* stream);
/* ---- analytic stage-0 query projection: q = W' LN0(x + normalise(bev_embed(world) - c_embed)) + b' without a per-token
 * reduction or GEMM (fax_modules.py:370-375,387-390 + to_q :189-190; derivation in csrc/eltwise2.cu).
 * chan_tab (C,20) fp32 = [a, b_, b_bev, w_cam(4), 0 | W'a, W'b_, W'1, b' | W'b_bev, W'w_cam(4), 0, 0, 0] (a, b_ = columns of
 * bev_embed.weight); n <= min(8, C/8) views;
 * stats (rows,12) fp32 = channel means of [x, x^2, x a, x b_, x b_bev, x w_cam(4)] from cobevt_row_stats9;
 * qx (x_rows,C) fp32 = W'x from cobevt_conv2d_nhwc_fwd (y32); x_rows = P (one prior shared by every row) or BL*P;
 * qa = a.a, qb = b_.b_, qd = 2 a.b_, ma = mean(a), mb = mean(b_); out (BL,n,P,C) 16-bit. */
int cobevt_row_stats9(const void* x, const float* chan_tab, float* stats, int64_t rows, int32_t C, int32_t dtype, void* stream);
int cobevt_bev_q_analytic(const float* qx, int64_t x_rows, const float* stats, const float* world, const float* chan_tab,
                          const float* E_inv, void* out, int32_t BL, int32_t n, int32_t P, int32_t C, float qa, float qb,
                          float qd, float ma, float mb, float eps, int32_t dtype, void* stream);
/* NCHW <-> NHWC(Cpad) with dtype conversion; module API boundaries only */
int cobevt_permute(const void* src, void* dst, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                   int32_t to_nhwc, int32_t src_dtype, int32_t dst_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COBEVT_B200_H_ */
