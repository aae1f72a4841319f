// cv2.resize(..., INTER_CUBIC) on the GPU (classic_algorithms.py:11-13, loading_methods.py:147,
// SRCNN_model.py:191).  Separable 4-tap Keys cubic (A = -0.75), half-pixel centres, tap index clamp.
//
// `bicubic_tables` evaluates the per-axis tap indices and coefficients once (the only place double precision is
// used).  `bicubic_swin_kernel` (default) stages the source footprint of a 256-element x 32-row output tile in shared
// memory with coalesced cp.async and lets every thread slide a 4-row register window of horizontal-pass results down
// its output column; `bicubic_stream_kernel` is the same loop reading global memory, for strong down-scaling.
//
// float path  : t from double, FMA-contracted coefficient polynomial, FMA accumulation in tap order
//               (OpenCV's default dispatch to <= 1e-6; uint8 = saturate(rint(.)) of the same path).
// fixed path  : OpenCV's 11-bit fixed-point uint8 path (== cv2.setUseOptimized(False)), bit-exact:
//               float32 f/t, plain float32 polynomial, int32 horizontal pass, float32 vertical pass.
#include "common.cuh"
#include <stdlib.h>

namespace srb {

struct AxisTap { int idx[4]; float coef[4]; };   // 32 bytes

__global__ void bicubic_tables(AxisTap* __restrict__ tab, int* __restrict__ base_out, int n_src, int n_dst, int fixed) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_dst) return;
  const double scale = 1.0 / ((double)n_dst / (double)n_src);
  const double fd = __dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const float A = -0.75f;
  int s;
  float c0, c1, c2, c3;
  if (fixed) {
    const float f = (float)fd;
    s = (int)floorf(f);
    const float t = __fsub_rn(f, (float)s);
    const float x1 = __fadd_rn(t, 1.f);
    c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.f * A), x1), 8.f * A), x1), 4.f * A);
    c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, t), A + 3.f), t), t), 1.f);
    const float u = __fsub_rn(1.f, t);
    c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, u), A + 3.f), u), u), 1.f);
    c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
    // INTER_RESIZE_COEF_SCALE = 2048, saturate_cast<short>(rint)
    c0 = fminf(fmaxf(rintf(__fmul_rn(c0, 2048.f)), -32768.f), 32767.f);
    c1 = fminf(fmaxf(rintf(__fmul_rn(c1, 2048.f)), -32768.f), 32767.f);
    c2 = fminf(fmaxf(rintf(__fmul_rn(c2, 2048.f)), -32768.f), 32767.f);
    c3 = fminf(fmaxf(rintf(__fmul_rn(c3, 2048.f)), -32768.f), 32767.f);
  } else {
    const double fl = floor(fd);
    s = (int)fl;
    const float t = (float)(fd - fl);
    const float x1 = __fadd_rn(t, 1.f);
    c0 = __fmaf_rn(__fmaf_rn(__fmaf_rn(A, x1, -5.f * A), x1, 8.f * A), x1, -4.f * A);
    c1 = __fmaf_rn(__fmul_rn(__fmaf_rn(A + 2.f, t, -(A + 3.f)), t), t, 1.f);
    const float u = __fsub_rn(1.f, t);
    c2 = __fmaf_rn(__fmul_rn(__fmaf_rn(A + 2.f, u, -(A + 3.f)), u), u, 1.f);
    c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
  }
  AxisTap a;
#pragma unroll
  for (int k = 0; k < 4; ++k) a.idx[k] = min(max(s - 1 + k, 0), n_src - 1);
  a.coef[0] = c0; a.coef[1] = c1; a.coef[2] = c2; a.coef[3] = c3;
  tab[d] = a;
  if (base_out) base_out[d] = s - 1;   // first tap row before clamping
}

constexpr int kTE = 256;   // interleaved output elements per block (= threads)

template <typename T> __device__ __forceinline__ float px_load(const T* p);
template <> __device__ __forceinline__ float px_load<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float px_load<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

// Streaming variant (fallback for strong down-scaling, where a tile's source footprint does not fit shared memory):
// a thread owns one interleaved output column (x*C + c) over a strip of output
// rows and keeps the horizontal-pass results of the four source rows under the vertical taps in registers.  When the
// tap window moves down by a source row, one new horizontal result is computed (4 L1-cached, block-coalesced loads +
// 4 FMAs) - so an up-scale by s costs (4 loads + 4 FMA)/s + 4 FMA + 1 coalesced store per output element, with no
// shared memory and no barriers.  Index clamping is applied when a row is fetched, which keeps the window a run of four
// consecutive (unclamped) rows at the image borders.  Same arithmetic order as OpenCV, so uint8 stays bit-exact.
template <typename T, bool FIXED>
__global__ void __launch_bounds__(kTE)
bicubic_stream_kernel(const T* __restrict__ src, T* __restrict__ dst, const AxisTap* __restrict__ xtab,
                      const AxisTap* __restrict__ ytab, const int* __restrict__ ybase, int src_h, int src_w, int C,
                      int dst_h, int dst_w, int rows_per_block, int clip01) {
  const int DE = dst_w * C, SE = src_w * C;
  const int e = blockIdx.x * kTE + threadIdx.x;
  if (e >= DE) return;
  const int x = e / C, c = e - x * C;
  const AxisTap xt = xtab[x];
  const int o0 = xt.idx[0] * C + c, o1 = xt.idx[1] * C + c, o2 = xt.idx[2] * C + c, o3 = xt.idx[3] * C + c;
  const int y0 = blockIdx.y * rows_per_block, y1 = min(y0 + rows_per_block, dst_h);
  const T* simg = src + (size_t)blockIdx.z * src_h * SE;
  T* out = dst + (size_t)blockIdx.z * dst_h * DE + (size_t)y0 * DE + e;

  auto hrow = [&](int r) -> float {                 // horizontal pass of (clamped) source row r for this column
    r = min(max(r, 0), src_h - 1);
    const T* row = simg + (size_t)r * SE;
    if (FIXED) {
      const int v = (int)row[o0] * (int)xt.coef[0] + (int)row[o1] * (int)xt.coef[1] +
                    (int)row[o2] * (int)xt.coef[2] + (int)row[o3] * (int)xt.coef[3];
      return __int_as_float(v);
    }
    float v = __fmul_rn(px_load(row + o0), xt.coef[0]);
    v = __fmaf_rn(px_load(row + o1), xt.coef[1], v);
    v = __fmaf_rn(px_load(row + o2), xt.coef[2], v);
    v = __fmaf_rn(px_load(row + o3), xt.coef[3], v);
    return v;
  };

  int u = __ldg(ybase + y0);
  float w0 = hrow(u), w1 = hrow(u + 1), w2 = hrow(u + 2), w3 = hrow(u + 3);
  for (int y = y0; y < y1; ++y, out += DE) {
    const int ub = __ldg(ybase + y);                // block-uniform
    while (u < ub) { w0 = w1; w1 = w2; w2 = w3; ++u; w3 = hrow(u + 3); }
    const float4 cy = __ldg(reinterpret_cast<const float4*>(ytab[y].coef));
    float v;
    if (FIXED) {
      const float sc = 1.f / (2048.f * 2048.f);
      v = __fmul_rn((float)__float_as_int(w0), __fmul_rn(cy.x, sc));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w1), __fmul_rn(cy.y, sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w2), __fmul_rn(cy.z, sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w3), __fmul_rn(cy.w, sc)));
    } else {
      v = __fmul_rn(w0, cy.x);
      v = __fmaf_rn(w1, cy.y, v);
      v = __fmaf_rn(w2, cy.z, v);
      v = __fmaf_rn(w3, cy.w, v);
    }
    if (sizeof(T) == 1) {
      const int qv = __float2int_rn(v);                      // round-half-even, then saturate
      *reinterpret_cast<uint8_t*>(out) = (uint8_t)min(max(qv, 0), 255);
    } else {
      if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
      *reinterpret_cast<float*>(out) = v;
    }
  }
}

// Shared-memory sliding-window variant (the default): stage 1 of the tiled kernel (coalesced cp.async of the source
// footprint) feeds the register window of the streaming kernel, so the horizontal pass reads shared memory instead of
// global memory and there is no horizontal-result buffer: per output element an up-scale by s costs 4/s shared gathers
// + 2 uniform shared loads (row coefficients, row base) + 1 coalesced store, and ~20 KB of shared memory per block.
template <typename T, bool FIXED>
__global__ void __launch_bounds__(kTE)
bicubic_swin_kernel(const T* __restrict__ src, T* __restrict__ dst, const AxisTap* __restrict__ xtab,
                    const AxisTap* __restrict__ ytab, const int* __restrict__ ybase, int src_h, int src_w, int C,
                    int dst_h, int dst_w, int tile_rows, int max_src_rows, int max_src_cols, int clip01) {
  extern __shared__ __align__(16) float wsm_[];
  float* S = wsm_;                                               // [max_src_rows][max_src_cols]
  float4* cy = reinterpret_cast<float4*>(S + (size_t)max_src_rows * max_src_cols);   // [tile_rows] vertical coefficients
  int* ub = reinterpret_cast<int*>(cy + tile_rows);              // [tile_rows] first (unclamped) tap row, tile-relative
  const int t = threadIdx.x;
  const int DE = dst_w * C, SE = src_w * C;
  const int e0 = blockIdx.x * kTE;
  const int e = e0 + t;
  const int y0 = blockIdx.y * tile_rows;
  const int y1 = min(y0 + tile_rows, dst_h);
  const size_t src_img = (size_t)blockIdx.z * src_h * SE;
  const int u0 = __ldg(ybase + y0);                              // unclamped first source row of the tile
  const int nr = min(__ldg(ybase + y1 - 1) + 4 - u0, max_src_rows);
  const int x_first = e0 / C, x_last = (min(e0 + kTE, DE) - 1) / C;
  const int c_lo = xtab[x_first].idx[0] * C;
  const int nc = min(xtab[x_last].idx[3] * C + C - c_lo, max_src_cols);

  // stage 1: rows u0 .. u0+nr-1 (clamped to the image) x columns c_lo .. c_lo+nc-1 -> shared memory
  for (int r = 0; r < nr; ++r) {
    const int gr = min(max(u0 + r, 0), src_h - 1);
    const T* row = src + src_img + (size_t)gr * SE + c_lo;
    for (int i = t; i < nc; i += kTE) {
      if (sizeof(T) == 4) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(S + r * max_src_cols + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(row + i) : "memory");
      } else {
        S[r * max_src_cols + i] = px_load(row + i);
      }
    }
  }
  for (int i = t; i < y1 - y0; i += kTE) {
    cy[i] = __ldg(reinterpret_cast<const float4*>(ytab[y0 + i].coef));
    ub[i] = __ldg(ybase + y0 + i) - u0;
  }
  if (sizeof(T) == 4) asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (e >= DE) return;

  const int x = e / C, c = e - x * C;
  const AxisTap xt = xtab[x];
  const int o0 = xt.idx[0] * C + c - c_lo, o1 = xt.idx[1] * C + c - c_lo, o2 = xt.idx[2] * C + c - c_lo,
            o3 = xt.idx[3] * C + c - c_lo;
  auto hrow = [&](int r) -> float {                    // horizontal pass of tile row r for this column
    const float* row = S + min(r, nr - 1) * max_src_cols;
    if (FIXED) {
      const int v = (int)row[o0] * (int)xt.coef[0] + (int)row[o1] * (int)xt.coef[1] +
                    (int)row[o2] * (int)xt.coef[2] + (int)row[o3] * (int)xt.coef[3];
      return __int_as_float(v);
    }
    float v = __fmul_rn(row[o0], xt.coef[0]);
    v = __fmaf_rn(row[o1], xt.coef[1], v);
    v = __fmaf_rn(row[o2], xt.coef[2], v);
    v = __fmaf_rn(row[o3], xt.coef[3], v);
    return v;
  };
  int u = 0;
  float w0 = hrow(0), w1 = hrow(1), w2 = hrow(2), w3 = hrow(3);
  T* out = dst + (size_t)blockIdx.z * dst_h * DE + (size_t)y0 * DE + e;
  for (int i = 0; i < y1 - y0; ++i, out += DE) {
    const int un = ub[i];                              // block-uniform
    while (u < un) { w0 = w1; w1 = w2; w2 = w3; ++u; w3 = hrow(u + 3); }
    const float4 k = cy[i];
    float v;
    if (FIXED) {
      const float sc = 1.f / (2048.f * 2048.f);
      v = __fmul_rn((float)__float_as_int(w0), __fmul_rn(k.x, sc));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w1), __fmul_rn(k.y, sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w2), __fmul_rn(k.z, sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(w3), __fmul_rn(k.w, sc)));
    } else {
      v = __fmul_rn(w0, k.x);
      v = __fmaf_rn(w1, k.y, v);
      v = __fmaf_rn(w2, k.z, v);
      v = __fmaf_rn(w3, k.w, v);
    }
    if (sizeof(T) == 1) {
      const int qv = __float2int_rn(v);                      // round-half-even, then saturate
      *reinterpret_cast<uint8_t*>(out) = (uint8_t)min(max(qv, 0), 255);
    } else {
      if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
      *reinterpret_cast<float*>(out) = v;
    }
  }
}

template <typename T, bool FIXED>
static int run_bicubic(const T* src, int batch, int sh, int sw, int C, T* dst, int dh, int dw, int clip01,
                       cudaStream_t stream) {
  SRB_REQUIRE(src && dst, "bicubic: null pointer");
  SRB_REQUIRE(batch >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && C > 0, "bicubic: bad geometry");
  if (batch == 0) return SRB_OK;
  AxisTap* tabs = nullptr;
  SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap) * ((size_t)dw + dh) + sizeof(int) * (size_t)dh, stream));
  int* ybase = reinterpret_cast<int*>(tabs + (size_t)dw + dh);
  AxisTap* xtab = tabs;
  AxisTap* ytab = tabs + dw;
  bicubic_tables<<<(dw + 127) / 128, 128, 0, stream>>>(xtab, nullptr, sw, dw, FIXED ? 1 : 0);
  bicubic_tables<<<(dh + 127) / 128, 128, 0, stream>>>(ytab, ybase, sh, dh, FIXED ? 1 : 0);
  int rc = launch_check("bicubic_tables");
  if (rc) return rc;
  {
    // default: shared-memory sliding window
    const double ratio = (double)sw / (double)dw;
    const int src_px = (int)((kTE / C + 2) * (ratio > 1.0 ? ratio : 1.0)) + 6;
    const int msc = ((src_px * C) + 3) & ~3;
    int tr = 32;
    auto rows_needed = [&](int n) { return (int)(((long)n * sh + dh - 1) / dh) + 6; };
    auto need = [&](int n) { return (size_t)rows_needed(n) * msc * sizeof(float) + (size_t)n * (sizeof(float4) + sizeof(int)); };
    while (tr > 4 && need(tr) > 40 * 1024) tr >>= 1;
    if (need(tr) <= 96 * 1024) {
      const size_t smem = need(tr);
      SRB_CUDA(cudaFuncSetAttribute(bicubic_swin_kernel<T, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dim3 grid((dw * C + kTE - 1) / kTE, (dh + tr - 1) / tr, batch);
      SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "bicubic: grid too large");
      bicubic_swin_kernel<T, FIXED><<<grid, kTE, smem, stream>>>(src, dst, xtab, ytab, ybase, sh, sw, C, dh, dw, tr,
                                                                 rows_needed(tr), msc, clip01);
      rc = launch_check("bicubic_swin_kernel");
    } else {
      // strong down-scaling: stream straight from global memory
      int rows = 64;
      const long target = 4L * sm_count();
      while (rows > 8 && (long)((dw * C + kTE - 1) / kTE) * ((dh + rows - 1) / rows) * batch < target) rows >>= 1;
      dim3 grid((dw * C + kTE - 1) / kTE, (dh + rows - 1) / rows, batch);
      SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "bicubic: grid too large");
      bicubic_stream_kernel<T, FIXED><<<grid, kTE, 0, stream>>>(src, dst, xtab, ytab, ybase, sh, sw, C, dh, dw, rows, clip01);
      rc = launch_check("bicubic_stream_kernel");
    }
  }
  SRB_CUDA(cudaFreeAsync(tabs, stream));
  return rc;
}

}  // namespace srb

using namespace srb;

extern "C" int srb_bicubic_f32(const float* src, int batch, int src_h, int src_w, int channels,
                               float* dst, int dst_h, int dst_w, int clip01, srb_stream_t stream) {
  return run_bicubic<float, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, clip01, (cudaStream_t)stream);
}

extern "C" int srb_bicubic_u8(const uint8_t* src, int batch, int src_h, int src_w, int channels,
                              uint8_t* dst, int dst_h, int dst_w, int fixed_point, srb_stream_t stream) {
  if (fixed_point)
    return run_bicubic<uint8_t, true>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, 0, (cudaStream_t)stream);
  return run_bicubic<uint8_t, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, 0, (cudaStream_t)stream);
}
