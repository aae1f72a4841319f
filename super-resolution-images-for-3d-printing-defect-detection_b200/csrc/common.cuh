// Shared host/device helpers for libsrb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/srb200.h"

namespace srb {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SRB_CUDA(call)                                            \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::srb::cuda_fail(_e, #call);    \
  } while (0)

#define SRB_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) { ::srb::set_error(__VA_ARGS__); return SRB_E_INVALID; } \
  } while (0)

inline int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return SRB_OK;
}

int sm_count();

template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <> __device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void store_from_float(T* p, float v);
template <> __device__ __forceinline__ void store_from_float<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void store_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ void store_from_float<uint8_t>(uint8_t* p, float v) {   // saturate(rint(v))
  *p = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f));
}
template <> __device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p) { return (float)*p; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case SRB_ACT_RELU:  return fmaxf(v, 0.f);
    case SRB_ACT_PRELU:
    case SRB_ACT_LEAKY: return v >= 0.f ? v : v * slope;
    case SRB_ACT_TANH:  return tanhf(v);
    default:            return v;
  }
}

inline size_t dtype_size(int dt) { return dt == SRB_F32 ? 4 : (dt == SRB_BF16 || dt == SRB_F16) ? 2 : 1; }   // (u8, e5m2: 1)

}  // namespace srb

// Packed conv weights (opaque to C callers).
struct srb_conv_weights {
  int kh, kw, cin, cout;
  float* hwio;            // [kh*kw][cin][cout_pad4] float32 (direct engine), cout padded to a multiple of 4
  int cout_pad4;
  float* bias;            // [cout] float32 (zeros when the layer has none)
  __nv_bfloat16* tc;      // [kh*kw][cout_pad][cin_pad] bf16, K-major rows (tcgen05 engine) or nullptr
  __half* tc_f16;         // same, IEEE half
  int tc_cout_pad;        // rows per tap in `tc` (multiple of 16)
  int tc_cin_pad;         // K extent of a row of `tc`: cin rounded up to a multiple of 64 (zero columns past cin)
  __nv_bfloat16* tc_fold; // cout <= 4 only: [dy][16 rows = dx*5+co][cin] (horizontal taps folded into N) or nullptr
  __half* tc_fold_f16;
  __nv_bfloat16* tc_head; // cin == 3, cout == 64 only: im2col GEMM operand [n_kb][64 cout rows][64 k], k = (dy*kw + dx)*3 + c, or nullptr
  __half* tc_head_f16;
  int tc_head_kb;         // 64-wide K blocks
  __nv_bfloat16* tc_head8;   // cin <= 8, cout == 64, 3x3 / 5x5 only: un-swizzled core-matrix order [dy][8-row group][K chunk = dx][row][8 ch]
  __half* tc_head8_f16;      //   for the NHWC8 head kernel (conv_head8.cu), or nullptr
};
