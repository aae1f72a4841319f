// Parameters and the fused epilogue shared by both convolution engines.
#pragma once
#include "common.cuh"

namespace srb {

struct ConvParams {
  const void* x; int x_dtype, x_cstride, x_coffset;
  void* y;       int y_dtype, y_cstride, y_coffset;
  void* y2;      int y2_dtype, y2_cstride, y2_mode;
  int B, H, W;
  int kh, kw, cin, cout;
  const float* w_hwio; int w_cout_pad;      // direct engine weights
  const void* w_tc; int w_tc_rows;          // tcgen05 engine weights [tap][rows][w_tc_cin] in the dtype of x
  int w_tc_cin;                             // cin rounded up to a multiple of 64 (the kernel walks it in 64-channel chunks)
  const void* w_tc_fold;                    // dx-folded weights ([dy][16][cin] for cout <= 4, [dy][192][cin] for cout == 64), or nullptr
  const void* w_tc_head; int w_tc_head_kb;  // cin == 3 im2col weights in the 16-bit dtype of y, or nullptr
  const void* w_tc_head8;                   // cin <= 8, 3x3 / 5x5: core-matrix-ordered weights of the NHWC8 head kernel, or nullptr
  const float* bias;                        // [cout], never null
  int act; float act_slope; const float* prelu;
  float alpha;
  const void* res1; int res1_dtype, res1_cstride; float beta1;
  const void* res2; int res2_dtype, res2_cstride; float beta2;
  int clip01;
  int d2s;                                  // 1, 2, 3, 4
  int c_post;                               // cout / (d2s*d2s)
};

__device__ __forceinline__ float load_elem(const void* base, int dtype, size_t idx) {
  if (dtype == SRB_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  if (dtype == SRB_F16) return __half2float(reinterpret_cast<const __half*>(base)[idx]);
  if (dtype == SRB_F8E5M2) {                                  // e5m2 = the upper byte of an IEEE half
    const __half_raw h = {(unsigned short)(reinterpret_cast<const uint8_t*>(base)[idx] << 8)};
    return __half2float(__half(h));
  }
  return __ldg(reinterpret_cast<const float*>(base) + idx);
}

// (e5m2 is rare on the scalar paths: kept out of line so that the common stores stay small)
static __device__ __noinline__ void store_e5m2(void* base, size_t idx, float v) {
  reinterpret_cast<uint8_t*>(base)[idx] = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2);
}
static __device__ __noinline__ float round_e5m2(float v) {
  const __half_raw h = {(unsigned short)(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2) << 8)};
  return __half2float(__half(h));
}

__device__ __forceinline__ void store_elem(void* base, int dtype, size_t idx, float v) {
  if (dtype == SRB_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else if (dtype == SRB_F16) reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
  else if (dtype == SRB_F8E5M2) store_e5m2(base, idx, v);
  else if (dtype == SRB_U8) reinterpret_cast<uint8_t*>(base)[idx] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
  else reinterpret_cast<float*>(base)[idx] = v;
}

// v as it reads back from a tensor of the given dtype
__device__ __forceinline__ float round_to(int dtype, float v) {
  if (dtype == SRB_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  if (dtype == SRB_F16) return __half2float(__float2half_rn(v));
  if (dtype == SRB_F8E5M2) return round_e5m2(v);
  return v;
}

// value-level epilogue: everything except the store address
__device__ __forceinline__ float epilogue_value(const ConvParams& p, float acc, int co, int c_out,
                                                size_t out_pix /* linear output pixel index */) {
  float v = acc + p.bias[co];
  const float slope = (p.act == SRB_ACT_PRELU) ? p.prelu[c_out] : p.act_slope;
  v = apply_act(v, p.act, slope);
  v *= p.alpha;
  if (p.res1) v = fmaf(p.beta1, load_elem(p.res1, p.res1_dtype, out_pix * p.res1_cstride + c_out), v);
  if (p.res2) v = fmaf(p.beta2, load_elem(p.res2, p.res2_dtype, out_pix * p.res2_cstride + c_out), v);
  if (p.clip01) v = fminf(fmaxf(v, 0.f), 1.f);
  return v;
}

// output pixel index and channel for conv-domain (b, y, x, co) under depth_to_space (DCR)
__device__ __forceinline__ void d2s_map(const ConvParams& p, int b, int y, int x, int co,
                                        size_t& out_pix, int& c_out) {
  if (p.d2s == 1) {
    out_pix = ((size_t)b * p.H + y) * p.W + x;
    c_out = co;
  } else {
    const int r = p.d2s;
    const int q = co / p.c_post;
    c_out = co - q * p.c_post;
    const int i = q / r, j = q - i * r;
    out_pix = ((size_t)b * p.H * r + (size_t)y * r + i) * ((size_t)p.W * r) + (size_t)x * r + j;
  }
}

__device__ __forceinline__ void epilogue_store(const ConvParams& p, int b, int y, int x, int co, float acc) {
  size_t out_pix; int c_out;
  d2s_map(p, b, y, x, co, out_pix, c_out);
  const float v = epilogue_value(p, acc, co, c_out, out_pix);
  store_elem(p.y, p.y_dtype, out_pix * p.y_cstride + p.y_coffset + c_out, v);
  if (p.y2) store_elem(p.y2, p.y2_dtype, out_pix * p.y2_cstride + c_out, p.y2_mode == 1 ? v - round_to(p.y_dtype, v) : v);
}

int conv_direct_launch(const ConvParams& p, cudaStream_t stream);
int conv_head_launch(const ConvParams& p, cudaStream_t stream);
bool conv_head_eligible(const ConvParams& p);
int conv_tc_launch(const ConvParams& p, cudaStream_t stream);
int conv_headtc_launch(const ConvParams& p, cudaStream_t stream);
bool conv_headtc_eligible(const ConvParams& p);
int conv_head8_launch(const ConvParams& p, cudaStream_t stream);
bool conv_head8_eligible(const ConvParams& p);
bool conv_tc_eligible(const ConvParams& p);

}  // namespace srb
