// srb_conv2d_nhwc: argument validation, weight packing and engine dispatch.
// Replaces Keras Conv2D(+Add/Lambda/depth_to_space/clip) as used by the reference's networks
// (SRCNN_model.py:50-52, EDSR_model.py:55-125, ESRGAN_model.py:212-345, VGG16_model.py:69-97).
#include "common.cuh"
#include "conv_common.cuh"
#include <stdlib.h>
#include <vector>

using namespace srb;

extern "C" int srb_conv_weights_create(const float* hwio, const float* bias, int kh, int kw, int cin, int cout,
                                       srb_conv_weights** out) {
  SRB_REQUIRE(hwio && out, "conv_weights_create: null pointer");
  SRB_REQUIRE(kh > 0 && kw > 0 && (kh & 1) && (kw & 1) && cin > 0 && cout > 0,
              "conv_weights_create: kernel must be odd and channel counts positive (got %dx%d, %d->%d)", kh, kw, cin, cout);
  srb_conv_weights* w = (srb_conv_weights*)calloc(1, sizeof(srb_conv_weights));
  if (!w) { set_error("conv_weights_create: out of host memory"); return SRB_E_NOMEM; }
  w->kh = kh; w->kw = kw; w->cin = cin; w->cout = cout;
  w->cout_pad4 = (cout + 3) & ~3;
  const int taps = kh * kw;
  std::vector<float> packed((size_t)taps * cin * w->cout_pad4, 0.f);
  for (int t = 0; t < taps; ++t)
    for (int c = 0; c < cin; ++c)
      for (int o = 0; o < cout; ++o)
        packed[((size_t)t * cin + c) * w->cout_pad4 + o] = hwio[((size_t)t * cin + c) * cout + o];
  std::vector<float> b(cout, 0.f);
  if (bias) for (int o = 0; o < cout; ++o) b[o] = bias[o];
  cudaError_t e;
  if ((e = cudaMalloc(&w->hwio, packed.size() * sizeof(float))) != cudaSuccess ||
      (e = cudaMemcpy(w->hwio, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&w->bias, (size_t)cout * sizeof(float))) != cudaSuccess ||
      (e = cudaMemcpy(w->bias, b.data(), (size_t)cout * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
    srb_conv_weights_destroy(w);
    return cuda_fail(e, "conv_weights_create");
  }
  // RGB head layers (cin == 3, cout == 64): im2col GEMM operand for the tcgen05 head kernel
  if (cin == 3 && cout == 64 && kh <= 9 && kw <= 9) {
    const int K = taps * 3, kpad = (K + 15) & ~15, n_kb = (kpad + 63) / 64;
    w->tc_head_kb = n_kb;
    std::vector<__nv_bfloat16> hb((size_t)n_kb * 64 * 64, __float2bfloat16_rn(0.f));
    std::vector<__half> hh((size_t)n_kb * 64 * 64, __float2half_rn(0.f));
    for (int k = 0; k < K; ++k)
      for (int o = 0; o < cout; ++o) {
        const float v = hwio[(size_t)k * cout + o];               // k = (dy*kw + dx)*3 + c is the HWIO flattening itself
        const size_t idx = ((size_t)(k / 64) * 64 + o) * 64 + (k % 64);
        hb[idx] = __float2bfloat16_rn(v);
        hh[idx] = __float2half_rn(v);
      }
    if ((e = cudaMalloc(&w->tc_head, hb.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc_head, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&w->tc_head_f16, hh.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc_head_f16, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
      srb_conv_weights_destroy(w);
      return cuda_fail(e, "conv_weights_create(tc head)");
    }
  }
  // small-filter RGB heads (cin <= 8, cout == 64, 3x3 / 5x5): B operand of the NHWC8 head kernel in the un-swizzled canonical
  // order - address(n, chunk) = dy block + (n / 8) * (KC * 128 B) + chunk * 128 B + (n % 8) * 16 B, chunk = dx, 8 channels each
  if (cin <= 8 && cout == 64 && kh == kw && (kh == 3 || kh == 5)) {
    const int kc = 2 * ((kw * 8 + 15) / 16);
    std::vector<__nv_bfloat16> hb((size_t)kh * 64 * kc * 8, __float2bfloat16_rn(0.f));
    std::vector<__half> hh((size_t)kh * 64 * kc * 8, __float2half_rn(0.f));
    for (int dy = 0; dy < kh; ++dy)
      for (int n = 0; n < 64; ++n)
        for (int dx = 0; dx < kw; ++dx)
          for (int c = 0; c < cin; ++c) {
            const float v = hwio[((size_t)(dy * kw + dx) * cin + c) * cout + n];
            const size_t idx = (((size_t)dy * 8 + n / 8) * kc + dx) * 64 + (size_t)(n % 8) * 8 + c;
            hb[idx] = __float2bfloat16_rn(v);
            hh[idx] = __float2half_rn(v);
          }
    if ((e = cudaMalloc(&w->tc_head8, hb.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc_head8, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&w->tc_head8_f16, hh.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc_head8_f16, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
      srb_conv_weights_destroy(w);
      return cuda_fail(e, "conv_weights_create(tc head8)");
    }
  }
  // tensor-core copies: [tap][cout_pad16][cin_pad64] K(cin)-major rows in bf16 and fp16.  cin = 64 is the native shape of
  // the tcgen05 engine; wider inputs (VGG16 blocks 2-5: 128 .. 512 channels, the ESRGAN dense blocks' growing concatenation:
  // 64 + j * growth) are walked in 64-channel K chunks with zero columns past cin.
  if (cin >= 32 && cin % 8 == 0 && cin <= 1024 && kh <= 9 && kw <= 9) {
    const int rows = (cout + 15) & ~15;
    const int kp = (cin + 63) & ~63;
    w->tc_cout_pad = rows;
    w->tc_cin_pad = kp;
    std::vector<__nv_bfloat16> tb((size_t)taps * rows * kp, __float2bfloat16_rn(0.f));
    std::vector<__half> th((size_t)taps * rows * kp, __float2half_rn(0.f));
    for (int t = 0; t < taps; ++t)
      for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c) {
          const float v = hwio[((size_t)t * cin + c) * cout + o];
          tb[((size_t)t * rows + o) * kp + c] = __float2bfloat16_rn(v);
          th[((size_t)t * rows + o) * kp + c] = __float2half_rn(v);
        }
    if ((e = cudaMalloc(&w->tc, tb.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc, tb.data(), tb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&w->tc_f16, th.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(w->tc_f16, th.data(), th.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
      srb_conv_weights_destroy(w);
      return cuda_fail(e, "conv_weights_create(tc)");
    }
  }
  if (cin == 64) {
    if (kh == 3 && kw == 3 && cout % 16 == 0 && cout >= 16 && cout <= 64) {
      // 64 -> 16 / 32 / 48 / 64 layers: horizontal taps folded into N = 3 * cout rows per vertical tap, row = dx * cout + co
      // (wide-tile kernel)
      const int n = 3 * cout;
      std::vector<__nv_bfloat16> fb((size_t)3 * n * cin, __float2bfloat16_rn(0.f));
      std::vector<__half> fh((size_t)3 * n * cin, __float2half_rn(0.f));
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
          for (int o = 0; o < cout; ++o)
            for (int c = 0; c < cin; ++c) {
              const float v = hwio[((size_t)(dy * 3 + dx) * cin + c) * cout + o];
              fb[((size_t)dy * n + dx * cout + o) * cin + c] = __float2bfloat16_rn(v);
              fh[((size_t)dy * n + dx * cout + o) * cin + c] = __float2half_rn(v);
            }
      if ((e = cudaMalloc(&w->tc_fold, fb.size() * 2)) != cudaSuccess ||
          (e = cudaMemcpy(w->tc_fold, fb.data(), fb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
          (e = cudaMalloc(&w->tc_fold_f16, fh.size() * 2)) != cudaSuccess ||
          (e = cudaMemcpy(w->tc_fold_f16, fh.data(), fh.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
        srb_conv_weights_destroy(w);
        return cuda_fail(e, "conv_weights_create(tc fold)");
      }
    }
    if (cout <= 4 && kh <= 9 && kw <= 9) {
      // few-channel layers (RGB tails, any odd filter up to 9 x 9): the horizontal taps are folded into N, row = dx * 4 + co,
      // N = kw * 4 rounded up to 16 rows per vertical tap (wide-tile kernel, mode 3)
      const int n = (kw * 4 + 15) & ~15;
      std::vector<__nv_bfloat16> fb((size_t)kh * n * cin, __float2bfloat16_rn(0.f));
      std::vector<__half> fh((size_t)kh * n * cin, __float2half_rn(0.f));
      for (int dy = 0; dy < kh; ++dy)
        for (int dx = 0; dx < kw; ++dx)
          for (int o = 0; o < cout; ++o)
            for (int c = 0; c < cin; ++c) {
              const float v = hwio[((size_t)(dy * kw + dx) * cin + c) * cout + o];
              fb[((size_t)dy * n + dx * 4 + o) * cin + c] = __float2bfloat16_rn(v);
              fh[((size_t)dy * n + dx * 4 + o) * cin + c] = __float2half_rn(v);
            }
      if ((e = cudaMalloc(&w->tc_fold, fb.size() * 2)) != cudaSuccess ||
          (e = cudaMemcpy(w->tc_fold, fb.data(), fb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
          (e = cudaMalloc(&w->tc_fold_f16, fh.size() * 2)) != cudaSuccess ||
          (e = cudaMemcpy(w->tc_fold_f16, fh.data(), fh.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
        srb_conv_weights_destroy(w);
        return cuda_fail(e, "conv_weights_create(tc fold)");
      }
    }
  }
  *out = w;
  return SRB_OK;
}

extern "C" void srb_conv_weights_destroy(srb_conv_weights* w) {
  if (!w) return;
  if (w->hwio) cudaFree(w->hwio);
  if (w->bias) cudaFree(w->bias);
  if (w->tc) cudaFree(w->tc);
  if (w->tc_f16) cudaFree(w->tc_f16);
  if (w->tc_fold) cudaFree(w->tc_fold);
  if (w->tc_fold_f16) cudaFree(w->tc_fold_f16);
  if (w->tc_head) cudaFree(w->tc_head);
  if (w->tc_head_f16) cudaFree(w->tc_head_f16);
  if (w->tc_head8) cudaFree(w->tc_head8);
  if (w->tc_head8_f16) cudaFree(w->tc_head8_f16);
  free(w);
}

static int fill_params(const srb_conv_args* a, ConvParams& p) {
  SRB_REQUIRE(a, "conv2d: null args");
  SRB_REQUIRE(a->x && a->y && a->weights, "conv2d: null pointer");
  const srb_conv_weights* w = a->weights;
  SRB_REQUIRE(a->batch >= 0 && a->height > 0 && a->width > 0, "conv2d: bad geometry %dx%dx%d", a->batch, a->height, a->width);
  auto float_dt = [](int d) { return d == SRB_F32 || d == SRB_BF16 || d == SRB_F16; };
  auto float_or_f8 = [&](int d) { return float_dt(d) || d == SRB_F8E5M2; };   // e5m2: second output / residuals only
  SRB_REQUIRE(float_dt(a->x_dtype), "conv2d: x dtype must be f32, bf16 or f16");
  // (u8 output: the [0, 1] image quantised as saturate(rint(255 v)) - the read-back format of the RGB tail layers)
  SRB_REQUIRE(float_dt(a->y_dtype) || (a->y_dtype == SRB_U8 && !a->y2), "conv2d: y dtype must be f32, bf16, f16 or (without y2) u8");
  SRB_REQUIRE(!a->y2 || float_or_f8(a->y2_dtype), "conv2d: y2 dtype must be f32, bf16, f16 or e5m2");
  SRB_REQUIRE(!a->res1 || float_or_f8(a->res1_dtype), "conv2d: res1 dtype must be f32, bf16, f16 or e5m2");
  SRB_REQUIRE(!a->res2 || float_or_f8(a->res2_dtype), "conv2d: res2 dtype must be f32, bf16, f16 or e5m2");
  const int d2s = a->d2s <= 0 ? 1 : a->d2s;
  SRB_REQUIRE(d2s >= 1 && d2s <= 4, "conv2d: depth_to_space factor must be 1..4 (got %d)", a->d2s);
  SRB_REQUIRE(w->cout % (d2s * d2s) == 0, "conv2d: cout %d is not divisible by d2s^2 = %d", w->cout, d2s * d2s);
  p.x = a->x; p.x_dtype = a->x_dtype;
  p.x_cstride = a->x_cstride > 0 ? a->x_cstride : w->cin; p.x_coffset = a->x_coffset;
  p.c_post = w->cout / (d2s * d2s);
  p.y = a->y; p.y_dtype = a->y_dtype;
  p.y_cstride = a->y_cstride > 0 ? a->y_cstride : p.c_post; p.y_coffset = a->y_coffset;
  p.y2 = a->y2; p.y2_dtype = a->y2_dtype; p.y2_cstride = a->y2_cstride > 0 ? a->y2_cstride : p.c_post;
  p.y2_mode = a->y2 ? a->y2_mode : 0;
  SRB_REQUIRE(p.y2_mode == 0 || p.y2_mode == 1, "conv2d: y2_mode must be 0 (copy) or 1 (rounding error of y)");
  SRB_REQUIRE(p.x_coffset >= 0 && p.x_coffset + w->cin <= p.x_cstride, "conv2d: input channel slice out of range");
  SRB_REQUIRE(p.y_coffset >= 0 && p.y_coffset + p.c_post <= p.y_cstride, "conv2d: output channel slice out of range");
  p.B = a->batch; p.H = a->height; p.W = a->width;
  p.kh = w->kh; p.kw = w->kw; p.cin = w->cin; p.cout = w->cout;
  p.w_hwio = w->hwio; p.w_cout_pad = w->cout_pad4;
  p.w_tc = a->x_dtype == SRB_F16 ? (const void*)w->tc_f16 : (const void*)w->tc; p.w_tc_rows = w->tc_cout_pad;
  p.w_tc_cin = w->tc_cin_pad;
  p.w_tc_fold = a->x_dtype == SRB_F16 ? (const void*)w->tc_fold_f16 : (const void*)w->tc_fold;
  p.w_tc_head = a->y_dtype == SRB_F16 ? (const void*)w->tc_head_f16 : (const void*)w->tc_head; p.w_tc_head_kb = w->tc_head_kb;
  p.w_tc_head8 = a->y_dtype == SRB_F16 ? (const void*)w->tc_head8_f16 : (const void*)w->tc_head8;
  p.bias = w->bias;
  p.act = a->act; p.act_slope = a->act_slope; p.prelu = a->prelu;
  SRB_REQUIRE(a->act >= SRB_ACT_NONE && a->act <= SRB_ACT_TANH, "conv2d: unknown activation %d", a->act);
  SRB_REQUIRE(a->act != SRB_ACT_PRELU || a->prelu, "conv2d: PReLU needs a slope vector");
  p.alpha = a->alpha;
  p.res1 = a->res1; p.res1_dtype = a->res1_dtype; p.res1_cstride = a->res1_cstride > 0 ? a->res1_cstride : p.c_post; p.beta1 = a->beta1;
  p.res2 = a->res2; p.res2_dtype = a->res2_dtype; p.res2_cstride = a->res2_cstride > 0 ? a->res2_cstride : p.c_post; p.beta2 = a->beta2;
  p.clip01 = a->clip01;
  p.d2s = d2s;
  return SRB_OK;
}

extern "C" int srb_conv2d_engine(const srb_conv_args* a) {
  ConvParams p;
  int rc = fill_params(a, p);
  if (rc) return rc;
  return (conv_tc_eligible(p) || conv_head8_eligible(p) || conv_headtc_eligible(p)) ? SRB_ENGINE_TCGEN05 : SRB_ENGINE_DIRECT;
}

extern "C" int srb_conv2d_nhwc(const srb_conv_args* a, srb_stream_t stream) {
  if (a && a->batch == 0 && a->weights) return SRB_OK;   // empty batch: nothing to do, pointers may be null
  ConvParams p;
  int rc = fill_params(a, p);
  if (rc) return rc;
  if (p.B == 0) return SRB_OK;
  const bool tc_ok = conv_tc_eligible(p);
  if (a->engine != SRB_ENGINE_DIRECT && !tc_ok && conv_head8_eligible(p)) {        // 3x3 / 5x5 RGB heads: NHWC8 rows as the A operand
    rc = conv_head8_launch(p, (cudaStream_t)stream);
    if (rc != SRB_E_UNSUPPORTED) return rc;
  }
  if (a->engine != SRB_ENGINE_DIRECT && !tc_ok && conv_headtc_eligible(p)) {       // RGB head layers: im2col GEMM on the tensor cores
    rc = conv_headtc_launch(p, (cudaStream_t)stream);
    if (rc != SRB_E_UNSUPPORTED || a->engine == SRB_ENGINE_TCGEN05) return rc;
  }
  if (a->engine == SRB_ENGINE_TCGEN05 && !tc_ok) {
    set_error("conv2d: shape not eligible for the tcgen05 engine (needs bf16/fp16 NHWC input with 16-byte aligned pixels, 32 <= cin <= 1024 "
              "in multiples of 8, odd filter up to 9x9)");
    return SRB_E_UNSUPPORTED;
  }
  if (a->engine == SRB_ENGINE_TCGEN05) return conv_tc_launch(p, (cudaStream_t)stream);
  if (a->engine == SRB_ENGINE_AUTO && tc_ok) {
    rc = conv_tc_launch(p, (cudaStream_t)stream);
    if (rc != SRB_E_UNSUPPORTED) return rc;          // (a shape whose staging does not fit shared memory falls through)
  }
  return conv_direct_launch(p, (cudaStream_t)stream);
}
