// cv2.resize(..., INTER_CUBIC) on the GPU (classic_algorithms.py:11-13, loading_methods.py:147,
// SRCNN_model.py:191).  Separable 4-tap Keys cubic (A = -0.75), half-pixel centres, tap index clamp.
//
// `bicubic_tables` evaluates the per-axis tap indices and coefficients once (the only place double
// precision is used).  `bicubic_tile_kernel` does the separable resampling for a tile of 256 interleaved
// output elements x TY output rows in three shared-memory stages: (1) the source rows/columns the tile
// needs are streamed in with coalesced cp.async, (2) the horizontal 4-tap pass runs smem -> smem, (3) the
// vertical 4-tap pass reads the thread's own column back and writes contiguous output rows.  The older
// `bicubic_kernel` (horizontal taps gathered straight from global memory) remains as the fallback for
// strong down-scaling, where the source footprint of a tile does not fit shared memory.
//
// float path  : t from double, FMA-contracted coefficient polynomial, FMA accumulation in tap order
//               (OpenCV's default dispatch to <= 1e-6; uint8 = saturate(rint(.)) of the same path).
// fixed path  : OpenCV's 11-bit fixed-point uint8 path (== cv2.setUseOptimized(False)), bit-exact:
//               float32 f/t, plain float32 polynomial, int32 horizontal pass, float32 vertical pass.
#include "common.cuh"
#include <stdlib.h>

namespace srb {

struct AxisTap { int idx[4]; float coef[4]; };   // 32 bytes

__global__ void bicubic_tables(AxisTap* __restrict__ tab, int n_src, int n_dst, int fixed) {
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
}

constexpr int kTE = 256;   // interleaved output elements per block (= threads)

template <typename T> __device__ __forceinline__ float px_load(const T* p);
template <> __device__ __forceinline__ float px_load<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float px_load<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

template <typename T, bool FIXED>
__global__ void __launch_bounds__(kTE)
bicubic_kernel(const T* __restrict__ src, T* __restrict__ dst, const AxisTap* __restrict__ xtab,
               const AxisTap* __restrict__ ytab, int src_h, int src_w, int C, int dst_h, int dst_w,
               int tile_rows, int max_src_rows, int clip01) {
  extern __shared__ float hbuf[];   // [max_src_rows][kTE]  (int32 bit patterns when FIXED)
  const int t = threadIdx.x;
  const int DE = dst_w * C;
  const int e = blockIdx.x * kTE + t;
  const int y0 = blockIdx.y * tile_rows;
  const int y1 = min(y0 + tile_rows, dst_h) - 1;
  const size_t src_img = (size_t)blockIdx.z * src_h * src_w * C;
  const size_t dst_img = (size_t)blockIdx.z * dst_h * DE;
  const int r_lo = ytab[y0].idx[0];
  const int r_hi = ytab[y1].idx[3];
  const int nr = min(r_hi - r_lo + 1, max_src_rows);
  const bool valid = e < DE;

  if (valid) {
    const int x = e / C, c = e - x * C;
    const AxisTap xt = xtab[x];
    const int o0 = xt.idx[0] * C + c, o1 = xt.idx[1] * C + c, o2 = xt.idx[2] * C + c, o3 = xt.idx[3] * C + c;
    const T* row = src + src_img + (size_t)r_lo * src_w * C;
    for (int r = 0; r < nr; ++r, row += (size_t)src_w * C) {
      if (FIXED) {
        const int v = (int)row[o0] * (int)xt.coef[0] + (int)row[o1] * (int)xt.coef[1] +
                      (int)row[o2] * (int)xt.coef[2] + (int)row[o3] * (int)xt.coef[3];
        hbuf[r * kTE + t] = __int_as_float(v);
      } else {
        float v = __fmul_rn(px_load(row + o0), xt.coef[0]);
        v = __fmaf_rn(px_load(row + o1), xt.coef[1], v);
        v = __fmaf_rn(px_load(row + o2), xt.coef[2], v);
        v = __fmaf_rn(px_load(row + o3), xt.coef[3], v);
        hbuf[r * kTE + t] = v;
      }
    }
  }
  // each thread only reads back its own column: no barrier needed
  if (!valid) return;
  for (int y = y0; y <= y1; ++y) {
    const AxisTap yt = ytab[y];
    float v;
    if (FIXED) {
      const float sc = 1.f / (2048.f * 2048.f);
      v = __fmul_rn((float)__float_as_int(hbuf[(yt.idx[0] - r_lo) * kTE + t]), __fmul_rn(yt.coef[0], sc));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(hbuf[(yt.idx[1] - r_lo) * kTE + t]), __fmul_rn(yt.coef[1], sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(hbuf[(yt.idx[2] - r_lo) * kTE + t]), __fmul_rn(yt.coef[2], sc)));
      v = __fadd_rn(v, __fmul_rn((float)__float_as_int(hbuf[(yt.idx[3] - r_lo) * kTE + t]), __fmul_rn(yt.coef[3], sc)));
    } else {
      v = __fmul_rn(hbuf[(yt.idx[0] - r_lo) * kTE + t], yt.coef[0]);
      v = __fmaf_rn(hbuf[(yt.idx[1] - r_lo) * kTE + t], yt.coef[1], v);
      v = __fmaf_rn(hbuf[(yt.idx[2] - r_lo) * kTE + t], yt.coef[2], v);
      v = __fmaf_rn(hbuf[(yt.idx[3] - r_lo) * kTE + t], yt.coef[3], v);
    }
    if (sizeof(T) == 1) {
      const int q = __float2int_rn(v);                       // round-half-even, then saturate
      reinterpret_cast<uint8_t*>(dst)[dst_img + (size_t)y * DE + e] = (uint8_t)min(max(q, 0), 255);
    } else {
      if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
      reinterpret_cast<float*>(dst)[dst_img + (size_t)y * DE + e] = v;
    }
  }
}


struct __align__(16) RowTap { int off[4]; float coef[4]; };   // vertical taps of one output row, offsets into the H buffer

// V consecutive output elements per thread (V = 4: 16-byte shared / global accesses, needs dst_w * C % 4 == 0),
// NT threads per block; a block covers NT * V interleaved output elements x tile_rows output rows.
template <typename T, bool FIXED, int V, int NT>
__global__ void __launch_bounds__(NT)
bicubic_tile_kernel(const T* __restrict__ src, T* __restrict__ dst, const AxisTap* __restrict__ xtab,
                    const AxisTap* __restrict__ ytab, int src_h, int src_w, int C, int dst_h, int dst_w,
                    int tile_rows, int max_src_rows, int max_src_cols, int clip01) {
  constexpr int kBE = NT * V;                                // output elements per block row
  extern __shared__ __align__(16) float tsm[];
  float* S = tsm;                                            // [max_src_rows][max_src_cols] source tile (as float)
  float* Hb = S + (size_t)max_src_rows * max_src_cols;       // [max_src_rows][kBE] horizontal-pass results
  RowTap* rt = reinterpret_cast<RowTap*>(Hb + (size_t)max_src_rows * kBE);   // [tile_rows]
  const int t = threadIdx.x;
  const int DE = dst_w * C, SE = src_w * C;
  const int e0 = blockIdx.x * kBE;
  const int e = e0 + t * V;                                  // first element of this thread
  const int y0 = blockIdx.y * tile_rows;
  const int y1 = min(y0 + tile_rows, dst_h) - 1;
  const size_t src_img = (size_t)blockIdx.z * src_h * SE;
  const size_t dst_img = (size_t)blockIdx.z * dst_h * DE;
  const int r_lo = ytab[y0].idx[0];
  const int nr = min(ytab[y1].idx[3] - r_lo + 1, max_src_rows);
  const int x_first = e0 / C, x_last = (min(e0 + kBE, DE) - 1) / C;
  const int c_lo = xtab[x_first].idx[0] * C;
  const int nc = min(xtab[x_last].idx[3] * C + C - c_lo, max_src_cols);
  const bool valid = e < DE;                                 // DE % V == 0, so all V elements are in range together

  // stage 1: source footprint -> shared memory (coalesced along the row)
  for (int r = 0; r < nr; ++r) {
    const T* row = src + src_img + (size_t)(r_lo + r) * SE + c_lo;
    for (int i = t; i < nc; i += NT) {
      if (sizeof(T) == 4) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(S + r * max_src_cols + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(row + i) : "memory");
      } else {
        S[r * max_src_cols + i] = px_load(row + i);
      }
    }
  }
  for (int i = t; i <= y1 - y0; i += NT) {
    const AxisTap a = ytab[y0 + i];
    RowTap q;
#pragma unroll
    for (int k = 0; k < 4; ++k) { q.off[k] = (a.idx[k] - r_lo) * kBE; q.coef[k] = a.coef[k]; }
    rt[i] = q;
  }
  if (sizeof(T) == 4) asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  if (!valid) return;
  // stage 2: horizontal pass, V output columns per thread, all source rows of the tile
  {
    int o[V][4];
    float cf[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int x = (e + v) / C, c = (e + v) - x * C;
      const AxisTap xt = xtab[x];
#pragma unroll
      for (int k = 0; k < 4; ++k) { o[v][k] = xt.idx[k] * C + c - c_lo; cf[v][k] = xt.coef[k]; }
    }
    const float* srow = S;
#pragma unroll 2
    for (int r = 0; r < nr; ++r, srow += max_src_cols) {
      float h[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if (FIXED) {
          const int iv = (int)srow[o[v][0]] * (int)cf[v][0] + (int)srow[o[v][1]] * (int)cf[v][1] +
                         (int)srow[o[v][2]] * (int)cf[v][2] + (int)srow[o[v][3]] * (int)cf[v][3];
          h[v] = __int_as_float(iv);
        } else {
          float a = __fmul_rn(srow[o[v][0]], cf[v][0]);
          a = __fmaf_rn(srow[o[v][1]], cf[v][1], a);
          a = __fmaf_rn(srow[o[v][2]], cf[v][2], a);
          a = __fmaf_rn(srow[o[v][3]], cf[v][3], a);
          h[v] = a;
        }
      }
      if (V == 4) *reinterpret_cast<float4*>(Hb + r * kBE + t * 4) = make_float4(h[0], h[1], h[2], h[3]);
      else Hb[r * kBE + t] = h[0];
    }
  }
  // stage 3: vertical pass; each thread only reads back its own columns of Hb (no barrier needed)
  const float* hcol = Hb + t * V;
  T* out = dst + dst_img + (size_t)y0 * DE + e;
#pragma unroll 2
  for (int i = 0; i <= y1 - y0; ++i, out += DE) {
    const RowTap q = rt[i];
    float hv[4][V];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (V == 4) {
        const float4 f = *reinterpret_cast<const float4*>(hcol + q.off[k]);
        hv[k][0] = f.x; hv[k][1 % V] = f.y; hv[k][2 % V] = f.z; hv[k][3 % V] = f.w;
      } else {
        hv[k][0] = hcol[q.off[k]];
      }
    }
    float res[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float a;
      if (FIXED) {
        const float sc = 1.f / (2048.f * 2048.f);
        a = __fmul_rn((float)__float_as_int(hv[0][v]), __fmul_rn(q.coef[0], sc));
        a = __fadd_rn(a, __fmul_rn((float)__float_as_int(hv[1][v]), __fmul_rn(q.coef[1], sc)));
        a = __fadd_rn(a, __fmul_rn((float)__float_as_int(hv[2][v]), __fmul_rn(q.coef[2], sc)));
        a = __fadd_rn(a, __fmul_rn((float)__float_as_int(hv[3][v]), __fmul_rn(q.coef[3], sc)));
      } else {
        a = __fmul_rn(hv[0][v], q.coef[0]);
        a = __fmaf_rn(hv[1][v], q.coef[1], a);
        a = __fmaf_rn(hv[2][v], q.coef[2], a);
        a = __fmaf_rn(hv[3][v], q.coef[3], a);
      }
      res[v] = a;
    }
    if (sizeof(T) == 1) {
      uint32_t pk = 0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int qv = min(max(__float2int_rn(res[v]), 0), 255);     // round-half-even, then saturate
        pk |= (uint32_t)qv << (8 * v);
      }
      if (V == 4) *reinterpret_cast<uint32_t*>(out) = pk;
      else *reinterpret_cast<uint8_t*>(out) = (uint8_t)pk;
    } else {
      if (clip01) {
#pragma unroll
        for (int v = 0; v < V; ++v) res[v] = fminf(fmaxf(res[v], 0.f), 1.f);
      }
      if (V == 4) *reinterpret_cast<float4*>(out) = make_float4(res[0], res[1 % V], res[2 % V], res[3 % V]);
      else *reinterpret_cast<float*>(out) = res[0];
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
  SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap) * ((size_t)dw + dh), stream));
  AxisTap* xtab = tabs;
  AxisTap* ytab = tabs + dw;
  bicubic_tables<<<(dw + 127) / 128, 128, 0, stream>>>(xtab, sw, dw, FIXED ? 1 : 0);
  bicubic_tables<<<(dh + 127) / 128, 128, 0, stream>>>(ytab, sh, dh, FIXED ? 1 : 0);
  int rc = launch_check("bicubic_tables");
  if (rc) return rc;
  int tile_rows = 32;
  auto src_rows = [&](int tr) { return (int)(((long)tr * sh + dh - 1) / dh) + 5; };
  // vector path: 4 consecutive elements per thread when rows of the output keep 16-byte (uint8: 4-byte) alignment
  // (opt-in: measured slower than the scalar mapping on B200 - 22 % vs 35 % of HBM at x2 - because 128-thread blocks
  //  with 40 KB of staging leave too few warps to cover the stage-1 load latency; kept for the double-buffered rewrite)
  static const bool vec4_enabled = getenv("SRB_BICUBIC_VEC4") != nullptr;
  const bool vec4 = vec4_enabled && ((dw * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) % (4 * sizeof(T))) == 0);
  const int NT = vec4 ? 128 : kTE, V = vec4 ? 4 : 1, BE = NT * V;
  // source columns one block can touch: its output pixels scaled back, plus the 4-tap support
  const double ratio = (double)sw / (double)dw;
  const int src_px = (int)((BE / C + 2) * (ratio > 1.0 ? ratio : 1.0)) + 6;
  const int msc = ((src_px * C) + 3) & ~3;
  auto tile_smem = [&](int tr) {
    return ((size_t)src_rows(tr) * (msc + BE)) * sizeof(float) + (size_t)tr * sizeof(RowTap);
  };
  if (vec4) tile_rows = 16;
  while (tile_rows > 4 && tile_smem(tile_rows) > 48 * 1024) tile_rows >>= 1;
  if (tile_smem(tile_rows) <= 96 * 1024) {
    const int msr = src_rows(tile_rows);
    const size_t smem = tile_smem(tile_rows);
    dim3 grid((dw * C + BE - 1) / BE, (dh + tile_rows - 1) / tile_rows, batch);
    SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "bicubic: grid too large");
    if (vec4) {
      SRB_CUDA(cudaFuncSetAttribute(bicubic_tile_kernel<T, FIXED, 4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bicubic_tile_kernel<T, FIXED, 4, 128><<<grid, 128, smem, stream>>>(src, dst, xtab, ytab, sh, sw, C, dh, dw, tile_rows, msr, msc, clip01);
    } else {
      SRB_CUDA(cudaFuncSetAttribute(bicubic_tile_kernel<T, FIXED, 1, kTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bicubic_tile_kernel<T, FIXED, 1, kTE><<<grid, kTE, smem, stream>>>(src, dst, xtab, ytab, sh, sw, C, dh, dw, tile_rows, msr, msc, clip01);
    }
  } else {
    tile_rows = 32;
    while (tile_rows > 1 && (size_t)src_rows(tile_rows) * kTE * sizeof(float) > 96 * 1024) tile_rows >>= 1;
    const int msr = src_rows(tile_rows);
    const size_t smem = (size_t)msr * kTE * sizeof(float);
    SRB_CUDA(cudaFuncSetAttribute(bicubic_kernel<T, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((dw * C + kTE - 1) / kTE, (dh + tile_rows - 1) / tile_rows, batch);
    bicubic_kernel<T, FIXED><<<grid, kTE, smem, stream>>>(src, dst, xtab, ytab, sh, sw, C, dh, dw, tile_rows, msr, clip01);
  }
  rc = launch_check("bicubic_kernel");
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
