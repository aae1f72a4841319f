/* srb200 - C ABI of the B200-native super-resolution inference path.
 *
 * The reference (bgmanuel99/Super-Resolution-Images-for-3D-Printing-Defect-Detection) has no FFI
 * of its own: its hot path disappears into three third-party engine calls.  Each entry point below
 * names the reference call it replaces.  All pointers are DEVICE pointers unless the name ends in
 * `_host`; sizes are element counts; `stream` is a `cudaStream_t` (0 = legacy default stream).
 * Every function returns 0 on success or a negative `SRB_E_*` code and then `srb_last_error()`
 * returns a thread-local description.  No entry point owns caller memory; nothing falls back to
 * the CPU.
 *
 *   Keras Conv2D / Add / Lambda / depth_to_space / clip  -> srb_conv2d_nhwc
 *        SRModels/deep_learning_models/SRCNN_model.py:50-52, EDSR_model.py:55-125,
 *        ESRGAN_model.py:212-345, model.predict at SRCNN_model.py:210, EDSR_model.py:274
 *   cv2.resize(..., INTER_CUBIC)                          -> srb_bicubic_f32 / srb_bicubic_u8
 *        SRModels/classic_super_resolution_algorithms/classic_algorithms.py:11-13,
 *        SRModels/loading_methods.py:147, SRCNN_model.py:191
 *   tf.image.psnr / tf.image.ssim (max_val = 1)           -> srb_psnr_ssim_f32
 *        SRModels/metrics.py:3-7
 *   add_padding + patch loops + overlap-add               -> srb_pad_extract_f32 / srb_overlap_add_f32
 *        SRModels/loading_methods.py:6-26, EDSR_model.py:201-256, SRCNN_model.py:127-188
 */
#ifndef SRB200_H
#define SRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* srb_stream_t;          /* cudaStream_t */
typedef struct srb_conv_weights srb_conv_weights;   /* opaque, device-resident packed kernel */

enum { SRB_OK = 0, SRB_E_INVALID = -1, SRB_E_UNSUPPORTED = -2, SRB_E_CUDA = -3, SRB_E_NOMEM = -4 };
enum { SRB_F32 = 0, SRB_BF16 = 1, SRB_U8 = 2, SRB_F16 = 3, SRB_F8E5M2 = 4 };
enum { SRB_ACT_NONE = 0, SRB_ACT_RELU = 1, SRB_ACT_PRELU = 2, SRB_ACT_LEAKY = 3, SRB_ACT_TANH = 4 };
/* conv engine selection: AUTO picks tcgen05 when the shape is eligible, else the CUDA-core path */
enum { SRB_ENGINE_AUTO = 0, SRB_ENGINE_DIRECT = 1, SRB_ENGINE_TCGEN05 = 2 };

const char* srb_last_error(void);
int srb_version(void);
/* sm_count / cc of the current device; fails unless the device is sm_100 (B200). */
int srb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- PSNR + SSIM, fused, one pass over both images (metrics.py:3-7) --------------------------
 * a, b: [B, H, W, C] float32 NHWC.  psnr, ssim, mse: [B] float32 (each may be NULL); mse is the
 * per-image mean squared error (the Keras "mean_squared_error" loss of SRCNN_model.py:59).
 * sums: optional [4] float64 {sum psnr, sum ssim, count, sum mse}, ACCUMULATED (not overwritten) so
 * that sharded evaluation can all-reduce one 4-vector.  workspace: >= srb_psnr_ssim_workspace(B) bytes.
 * H and W must be >= 11 (tf.image.ssim raises otherwise) -> SRB_E_INVALID.
 * With ssim == NULL only the squared error is evaluated (tf.image.psnr alone, metrics.py:3-4): a streaming reduction at
 * the HBM roofline, any image size; the ssim entry of `sums` then receives nothing. */
size_t srb_psnr_ssim_workspace(int batch);
int srb_psnr_ssim_f32(const float* a, const float* b, int batch, int height, int width, int channels,
                      float max_val, float* psnr, float* ssim, float* mse, double* sums,
                      void* workspace, size_t workspace_bytes, srb_stream_t stream);
/* Same pass with the window chosen: SRB_SSIM_TF = srb_psnr_ssim_f32; SRB_SSIM_SKIMAGE = the metric definitions the
 * reference's classical benchmark uses (super_resolucion_clasica.ipynb cell 7: skimage.metrics.peak_signal_noise_ratio
 * and structural_similarity(..., data_range, channel_axis=2) with their defaults - 7 x 7 uniform window, sample
 * covariance N/(N-1), K1 = 0.01, K2 = 0.03, borders cropped by 3): psnr = 10 log10(max_val^2 / mse), ssim = mean over the
 * (H-6) x (W-6) valid windows and channels.  Needs H, W >= 7.
 * SRB_SSIM_TF on wide 1- / 3-channel images runs both filter passes on the tensor path with the Gaussian rounded to an
 * fp16 window that sums to exactly 1 (|dSSIM| <= ~3e-6 against the float32 Gaussian, data kept to 22 bits);
 * SRB_SSIM_TF_EXACT keeps the float32 Gaussian on the CUDA cores for every size (also: environment SRB_SSIM_EXACT=1). */
enum { SRB_SSIM_TF = 0, SRB_SSIM_SKIMAGE = 1, SRB_SSIM_TF_EXACT = 2 };
int srb_psnr_ssim_window_f32(const float* a, const float* b, int batch, int height, int width, int channels,
                             float max_val, int window, float* psnr /* [B] or NULL */, float* ssim /* [B] or NULL */,
                             float* mse /* [B] or NULL */, double* sums /* [4] or NULL, accumulated */,
                             void* workspace, size_t workspace_bytes, srb_stream_t stream);

/* ---- bicubic resampling == cv2.resize(src, (dst_w, dst_h), INTER_CUBIC) ------------------------
 * NHWC interleaved, any channel count, any ratio.  clip01: clamp to [0,1] (loading_methods.py:148).
 * u8: fixed_point = 0 -> saturate(rint(float path)) (OpenCV default dispatch);
 *     fixed_point = 1 -> OpenCV's 11-bit fixed-point path (cv2.setUseOptimized(False)), bit-exact. */
int srb_bicubic_f32(const float* src, int batch, int src_h, int src_w, int channels,
                    float* dst, int dst_h, int dst_w, int clip01, srb_stream_t stream);
/* cv2.resize(src, (dst_w, dst_h), interpolation) for the other codes of the reference's interpolation map
 * (classic_algorithms.py:7-21, loading_methods.py:133-147): SRB_INTER_LINEAR, SRB_INTER_AREA (up-scaling, where OpenCV
 * evaluates it as a bilinear filter with area coefficients) and SRB_INTER_CUBIC (== srb_bicubic_f32).  Same kernels:
 * the bilinear taps are (0, 1 - t, t, 0) in the four-tap table.  SRB_INTER_LANCZOS4 (classic_algorithms.py:19-21, 58-62)
 * is OpenCV's 8-tap interpolateLanczos4 filter on its own eight-tap table and streaming kernel. */
enum { SRB_INTER_LINEAR = 1, SRB_INTER_CUBIC = 2, SRB_INTER_AREA = 3, SRB_INTER_LANCZOS4 = 4 };   /* OpenCV's codes */
int srb_resize_f32(const float* src, int batch, int src_h, int src_w, int channels,
                   float* dst, int dst_h, int dst_w, int interpolation, int clip01, srb_stream_t stream);
int srb_bicubic_u8(const uint8_t* src, int batch, int src_h, int src_w, int channels,
                   uint8_t* dst, int dst_h, int dst_w, int fixed_point, srb_stream_t stream);
/* cv2.resize on uint8 images for the same four codes (super_resolucion_clasica.ipynb cell 7 calls interpolate_bilinear /
 * _area / _lanczos on uint8 arrays): OpenCV's 11-bit fixed-point linear, area (up-scaling) and Lanczos-4 paths, bit-exact
 * against cv2 4.13 golden outputs; SRB_INTER_CUBIC == srb_bicubic_u8(..., fixed_point = 0). */
int srb_resize_u8(const uint8_t* src, int batch, int src_h, int src_w, int channels,
                  uint8_t* dst, int dst_h, int dst_w, int interpolation, srb_stream_t stream);

/* ---- tiling (loading_methods.py:6-26; EDSR_model.py:201-256) -----------------------------------
 * pad_extract: reflect-pad bottom/right by the reference's rule and cut [ny*nx, P, P, C] patches at
 * stride S.  srb_tiling_geometry returns padded size and patch grid (host-only helper).
 * overlap_add: average the patches covering each output pixel (patch order = reference loop order),
 * crop to out_h x out_w, clip to [0,1]. */
int srb_tiling_geometry(int height, int width, int patch, int stride,
                        int* padded_h, int* padded_w, int* ny, int* nx);
int srb_pad_extract_f32(const float* image, int height, int width, int channels, int patch, int stride,
                        float* patches, srb_stream_t stream);
int srb_overlap_add_f32(const float* patches, int ny, int nx, int patch_out, int stride_out, int channels,
                        float* image, int out_h, int out_w, srb_stream_t stream);

/* ---- convolution (Keras Conv2D, padding="same", stride 1) with fused epilogue -------------------
 * y = clip( alpha * act(conv(x, W) + bias) + beta1 * res1 + beta2 * res2 ), optionally written
 * through tf.nn.depth_to_space(y, d2s) (DCR).  With d2s = r the output tensor is
 * [B, H*r, W*r, cout/(r*r)] and PReLU slopes are indexed by the post-shuffle channel. */
typedef struct srb_conv_args {
  const void* x;        int x_dtype;   int x_cstride;  int x_coffset;   /* input  NHWC, C-slice allowed */
  void*       y;        int y_dtype;   int y_cstride;  int y_coffset;   /* output NHWC, C-slice allowed; SRB_U8 (without
                                                                           y2) stores saturate(rint(255 v)): the [0, 1]
                                                                           image quantised for a 4x smaller read-back */
  void*       y2;       int y2_dtype;  int y2_cstride;  int y2_mode;    /* optional second output, same geometry:
                                                                           mode 0 = the same values in another dtype
                                                                           (fp32 trunk next to the 16-bit operand);
                                                                           mode 1 = v - round_to_y_dtype(v), the
                                                                           rounding error of y, so that y + y2 carries
                                                                           ~22 bits in two 16-bit tensors, or ~14 bits in
                                                                           3 bytes when y2 is SRB_F8E5M2 (allowed for y2,
                                                                           res1 and res2 only) */
  int batch, height, width;
  const srb_conv_weights* weights;
  int act;              float act_slope;               const float* prelu;   /* [cout / d2s^2] */
  float alpha;
  const void* res1;     int res1_dtype; int res1_cstride; float beta1;   /* same geometry as y (pre-d2s NOT supported) */
  const void* res2;     int res2_dtype; int res2_cstride; float beta2;
  int clip01;
  int d2s;              /* 1 = none, 2/3/4 = depth_to_space factor */
  int engine;           /* SRB_ENGINE_* */
} srb_conv_args;

/* hwio_host: Keras kernel [kh, kw, cin, cout] float32 on the HOST; bias_host: [cout] or NULL.
 * Packs (and rounds to bf16 and fp16 for the tcgen05 path) once; the result lives on the current device. */
int srb_conv_weights_create(const float* hwio_host, const float* bias_host, int kh, int kw, int cin, int cout,
                            srb_conv_weights** out);
void srb_conv_weights_destroy(srb_conv_weights* w);
int srb_conv2d_nhwc(const srb_conv_args* args, srb_stream_t stream);
/* which engine AUTO would pick for these args: SRB_ENGINE_DIRECT or SRB_ENGINE_TCGEN05 */
int srb_conv2d_engine(const srb_conv_args* args);
/* tcgen05 engine: launch every eligible Cin = 64 layer (>= 32-channel chunks) as CTA pairs - clusters of two CTAs issuing
 * cta_group::2 MMAs (M = 256) from the leader, each CTA keeping half of the weight rows.  Returns the previous
 * setting (on < 0 only queries).  Default off: pairs are then used only where they pay - 256-channel chunks (the
 * up-sampling convs) and inputs wider than 64 channels, where they double the MMA width per resident weight byte;
 * for Cin = 64 and N <= 128 they measured equal to single CTAs on B200. */
int srb_conv_tc_set_cta_pairs(int on);

/* ---- EDSR up-sampling tail as one composed 5 x 5 convolution (EDSR_model.py:76-95, 117-123) --------
 * Conv2D(64 -> 64 r'^2, 3x3) -> depth_to_space(r') [twice for x4] -> Conv2D(64 -> C, 3x3) has no activation in between, so it
 * is one linear map from the 64-channel low-resolution image to the r x r x C block of every low-resolution pixel with a
 * 5 x 5 footprint.  w_host: [3][3][5][5][cin][r*r*C] float32, the exact map for the nine border classes (vy, vx) = (first /
 * interior / last row) x (first / interior / last column) - the zero padding of the INTERMEDIATE images only reaches the
 * outermost low-resolution pixels (srb200/compose.py builds it in float64); output channel (i*r + j)*C + c is sub-pixel
 * (i, j), channel c.  bias_host: [3][3][r*r*C].  The 16-bit copies hold w * w_scale (a power of two); the kernel multiplies
 * the accumulators by 1 / w_scale.  cin must be 64, r*r*C <= 48.
 * srb_upsample_composed: x [B, H, W, 64] fp16 / bf16 NHWC (channel slice of a wider buffer allowed) ->
 * y [B, H*r, W*r, C] f32 / f16 / bf16 / u8 (u8 = saturate(rint(255 v)) of the [0, 1] image), optionally clipped to [0, 1].
 * H, W >= 2.  One launch of upsample5_fold_kernel (tcgen05). */
typedef struct srb_upsampler srb_upsampler;
int srb_upsampler_create(const float* w_host, const float* bias_host, int cin, int image_channels, int scale, float w_scale,
                         srb_upsampler** out);
void srb_upsampler_destroy(srb_upsampler* u);
int srb_upsample_composed(const srb_upsampler* u, const void* x, int x_dtype, int x_cstride, int x_coffset,
                          int batch, int height, int width, void* y, int y_dtype, int clip01, srb_stream_t stream);

/* ---- small layout / elementwise helpers used between layers -------------------------------------- */
int srb_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, float scale, float shift,
             srb_stream_t stream);                                   /* dst = src * scale + shift */
/* dst = max(src * scale + shift, 0): the ReLU that follows a convolution whose input channels were accumulated in
 * several passes of srb_conv2d_nhwc (64-channel slices of a wider input, fp32 partial sums through res1) */
int srb_cast_relu(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, float scale, float shift,
                  srb_stream_t stream);
int srb_maxpool2x2_nhwc(const void* x, int dtype, int batch, int height, int width, int channels, void* y,
                        srb_stream_t stream);                        /* VGG16_model.py:69 (MaxPooling2D) */
int srb_gap_dense_softmax(const void* x, int dtype, int batch, int hw, int channels,
                          const float* w1, const float* b1, int hidden, const float* w2, const float* b2,
                          int classes, float* probs, srb_stream_t stream);   /* VGG16_model.py:84-97 */
/* SelfAttention core (ESRGAN_model.py:48-70): o = softmax(g f^T) h per image; f,g: [B,HW,dk], h: [B,HW,dv] */
int srb_self_attention_f32(const float* f, const float* g, const float* h, int batch, int hw, int dk, int dv,
                           float* o, srb_stream_t stream);

/* The same contract on the tensor cores (the 16-bit precision modes): S = g f^T as a tcgen05 product of fp16 (hi, lo)
 * split operands (float32-grade scores), online softmax in float32, P in fp16 for the second tcgen05 product P h (h in fp16),
 * float32 accumulation; |o - exact| ~ 1e-3 * max|h|.  dk <= 16, dv in {16, 32, 64}.  workspace: >=
 * srb_self_attention_tc_workspace(...) bytes, 128-byte aligned (padded 16-bit operands and the transpose of h). */
size_t srb_self_attention_tc_workspace(int batch, int hw, int dk, int dv);
int srb_self_attention_tc(const float* f, const float* g, const float* h, int batch, int hw, int dk, int dv,
                          float* o, void* workspace, size_t workspace_bytes, srb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H */
