// cv2.resize(..., INTER_CUBIC) on the GPU (classic_algorithms.py:11-13, loading_methods.py:147,
// SRCNN_model.py:191).  Separable 4-tap Keys cubic (A = -0.75), half-pixel centres, tap index clamp.
//
// `bicubic_tables` evaluates the per-axis tap indices and coefficients once (the only place double precision is
// used).  `bicubic_quad_kernel` (default) stages the source footprint of a 1,024-element x up-to-64-row output tile in
// shared memory and lets every thread slide a 4-row register window of horizontal-pass results down four adjacent
// output columns (packed fp32x2 math, 16-byte stores); `bicubic_stream_kernel` is the one-column loop reading global
// memory, for strong down-scaling where the footprint does not fit.
//
// float path  : t from double, FMA-contracted coefficient polynomial, FMA accumulation in tap order
//               (OpenCV's default dispatch to <= 1e-6; uint8 = saturate(rint(.)) of the same path).
// fixed path  : OpenCV's 11-bit fixed-point uint8 path (== cv2.setUseOptimized(False)), bit-exact:
//               float32 f/t, plain float32 polynomial, int32 horizontal pass, float32 vertical pass.
#include "common.cuh"
#include <stdlib.h>

namespace srb {

struct AxisTap { int idx[4]; float coef[4]; };   // 32 bytes

// mode: 0 cubic (float path), 1 cubic (OpenCV 11-bit fixed-point coefficients), 2 bilinear, 3 bilinear with INTER_AREA's
// up-scaling coefficients (cv2.resize treats INTER_AREA as that when the image grows).  The bilinear modes fill the same
// four-tap table with (0, 1 - t, t, 0), so every kernel below serves them unchanged.  Modes 4 / 5 are 2 / 3 without the
// fraction clamp at the image ends: OpenCV builds its y tables that way and clamps the row index in the row loop, which
// the uint8 fixed-point path can tell apart (a border row is blended with itself through two truncated products).
__global__ void bicubic_tables(AxisTap* __restrict__ tab, int* __restrict__ base_out, int n_src, int n_dst, int mode) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_dst) return;
  const double scale = 1.0 / ((double)n_dst / (double)n_src);
  const double fd = __dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const float A = -0.75f;
  int s;
  float c0, c1, c2, c3;
  if (mode >= 2) {
    // OpenCV resize.cpp, linear branch: fx in float, taps sx and sx + 1, both ends clamp to a pure copy
    float t;
    int sx;
    const bool clamp_t = mode < 4;
    if (mode == 2 || mode == 4) {
      const float f = (float)fd;
      sx = (int)floorf(f);
      t = __fsub_rn(f, (float)sx);
    } else {
      const double inv_scale = (double)n_dst / (double)n_src;
      sx = (int)floor((double)d * scale);
      t = (float)((double)(d + 1) - (double)(sx + 1) * inv_scale);
      t = t <= 0.f ? 0.f : __fsub_rn(t, floorf(t));
    }
    if (clamp_t) {
      if (sx < 0) { t = 0.f; sx = 0; }
      if (sx >= n_src - 1) { t = 0.f; sx = n_src - 1; }
    }
    s = sx;
    c0 = 0.f; c1 = __fsub_rn(1.f, t); c2 = t; c3 = 0.f;
  } else if (mode == 1) {
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

// Shared-memory quad kernel (the default).  A block owns kQT * kQE = 1,024 interleaved output elements (x*C + c) by
// `tile_rows` output rows.  Stage 1 copies the source footprint of the tile into shared memory as float with 16-byte
// (uint8: 4-byte) vector accesses; rows and columns outside the image are filled with the clamped (replicated) border
// pixels there, so the main loop needs no index clamps.  Every thread then produces kQE = 4 consecutive output
// elements per row - one 16-byte (uint8: 4-byte) coalesced store - and slides a 4-row register window of
// horizontal-pass results down its columns: an up-scale by s costs (16 shared gathers + 8 packed FMAs) / s + 8 packed
// FMAs + 1 store per four output elements.  The arithmetic is fp32x2 packed (FMUL2 / FFMA2), which is the same IEEE
// operation per lane as the scalar code, so the results are bit-identical to the streaming kernel:
//   float path : FMA chain in tap order (OpenCV default dispatch to <= 1e-6);
//   fixed path : the integer horizontal pass is evaluated in fp32 (|sum| < 2^24, so every product and sum is exact),
//                the vertical pass is OpenCV's float32 multiply-then-add with coefficients scaled by 2^-22.
constexpr int kQT = 256;   // threads per block
constexpr int kQE = 4;     // interleaved output elements per thread

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
};
template <> struct Vec4<uint8_t> {
  static __device__ __forceinline__ float4 load(const uint8_t* p) {
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
    return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  }
};

// MAP: which four of its warp's 128 elements a lane owns.  One shared gather of a warp reads 32 source floats spaced
// ~k / s apart (k = element stride between lanes, s = scale) plus a few floats of per-pixel jitter, and is free of bank
// conflicts only while that span stays below 32 floats:
//   1  elements l, 32 + l, 64 + l, 96 + l (k = 1): float32 always (four coalesced 4-byte stores per row); uint8 below x2.3
//      (four 1-byte stores, 32 contiguous bytes per warp and instruction);
//   2  pairs (2l, 2l + 1), (64 + 2l, 64 + 2l + 1) (k = 2): uint8 from x2.3 up (two 2-byte stores);
//   0  four consecutive elements (k = 4, one 4-byte store): uint8 with a channel count the other maps are not built for.
// With k = 4 the x2 gathers were two- to three-way conflicted and the kernel sat at 85 % of the shared-memory wavefront
// rate (profiles/r02_ncu_bicubic_u8.txt).
template <typename T, bool FIXED, int KC, int MAP>    // KC: compile-time channel count (tap offsets become immediates), 0 = run time
__global__ void __launch_bounds__(kQT)
bicubic_quad_kernel(const T* __restrict__ src, T* __restrict__ dst, const AxisTap* __restrict__ xtab,
                    const int* __restrict__ xbase, const AxisTap* __restrict__ ytab, const int* __restrict__ ybase,
                    int src_h, int src_w, int C_rt, int dst_h, int dst_w, int tile_rows, int max_src_rows, int pitch,
                    int vec_src, int vec_dst, int clip01) {
  const int C = KC ? KC : C_rt;
  extern __shared__ __align__(16) float qsm_[];
  float* S = qsm_;                                                          // [max_src_rows][pitch]
  float4* cyv = reinterpret_cast<float4*>(S + (size_t)max_src_rows * pitch);  // [tile_rows] vertical coefficients
  int* ub = reinterpret_cast<int*>(cyv + tile_rows);                        // [tile_rows] first tap row, tile-relative
  const int t = threadIdx.x;
  const int DE = dst_w * C, SE = src_w * C;
  const int e0 = blockIdx.x * (kQT * kQE);
  const int y0 = blockIdx.y * tile_rows;
  const int rows = min(tile_rows, dst_h - y0);
  const int u0 = __ldg(ybase + y0);                                         // unclamped first source row of the tile
  const int nr = min(__ldg(ybase + y0 + rows - 1) + 4 - u0, max_src_rows);
  const int x_first = e0 / C, x_last = (min(e0 + kQT * kQE, DE) - 1) / C;
  const int c_lo = (__ldg(xbase + x_first) * C) & ~3;                       // floor to a multiple of 4 (may be negative)
  const int nc4 = min(((__ldg(xbase + x_last) + 4) * C - c_lo + 3) >> 2, pitch >> 2);
  const T* simg = src + (size_t)blockIdx.z * src_h * SE;

  // stage 1: source rows u0 .. u0+nr-1 x elements c_lo .. c_lo+4*nc4-1, border-replicated, as float.  The (row, chunk)
  // space is flattened over the block (row = idx / nc4 by multiply-high); interior float chunks go through cp.async so
  // that every load of the tile is in flight at once.
  {
    const uint32_t total = (uint32_t)(nr * nc4);
    const uint32_t magic = 0xFFFFFFFFu / (uint32_t)nc4 + 1u;                // exact quotient for idx < 2^32 / nc4
    auto slow_chunk = [&](const T* row, int q, float* sdst) {               // border / unaligned chunk: clamped scalar gathers
      float f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int qq = q + k;
        int px = (qq + 8 * C) / C - 8;                                      // floor division (qq >= -2C - 3)
        const int ch = qq - px * C;
        px = min(max(px, 0), src_w - 1);
        f[k] = px_load(row + px * C + ch);
      }
      *reinterpret_cast<float4*>(sdst) = make_float4(f[0], f[1], f[2], f[3]);
    };
    if (sizeof(T) == 4) {
      for (uint32_t idx = t; idx < total; idx += kQT) {
        const int r = (int)__umulhi(idx, magic);
        const int i = (int)idx - r * nc4;
        const int gr = min(max(u0 + r, 0), src_h - 1);
        const T* row = simg + (size_t)gr * SE;
        float* sdst = S + r * pitch + 4 * i;
        const int q = c_lo + 4 * i;
        if (vec_src && q >= 0 && q + 4 <= SE) {
          const uint32_t d = (uint32_t)__cvta_generic_to_shared(sdst);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(row + q) : "memory");
        } else {
          slow_chunk(row, q, sdst);
        }
      }
    } else {
      // uint8: four chunks per thread and pass, all 4-byte loads issued before the first conversion (the loads of a
      // pass are independent; converting each right after its load serialised the whole stage on memory latency)
      constexpr int kU = 4;
      for (uint32_t base = t; base < total; base += kQT * kU) {
        uint32_t raw[kU];
        float* sd[kU];
        bool fast[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const uint32_t idx = base + (uint32_t)u * kQT;
          fast[u] = false;
          sd[u] = nullptr;
          raw[u] = 0u;
          if (idx < total) {
            const int r = (int)__umulhi(idx, magic);
            const int i = (int)idx - r * nc4;
            const int gr = min(max(u0 + r, 0), src_h - 1);
            const T* row = simg + (size_t)gr * SE;
            const int q = c_lo + 4 * i;
            sd[u] = S + r * pitch + 4 * i;
            fast[u] = vec_src && q >= 0 && q + 4 <= SE;
            if (fast[u]) raw[u] = __ldg(reinterpret_cast<const uint32_t*>(row + q));
            else slow_chunk(row, q, sd[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (fast[u])
            *reinterpret_cast<float4*>(sd[u]) = make_float4((float)(raw[u] & 0xFFu), (float)((raw[u] >> 8) & 0xFFu),
                                                            (float)((raw[u] >> 16) & 0xFFu), (float)(raw[u] >> 24));
      }
    }
    for (int i = t; i < rows; i += kQT) {
      float4 k = __ldg(reinterpret_cast<const float4*>(ytab[y0 + i].coef));
      if (FIXED) {
        const float sc = 1.f / (2048.f * 2048.f);
        k = make_float4(__fmul_rn(k.x, sc), __fmul_rn(k.y, sc), __fmul_rn(k.z, sc), __fmul_rn(k.w, sc));
      }
      cyv[i] = k;
      ub[i] = __ldg(ybase + y0 + i) - u0;
    }
    if (sizeof(T) == 4) asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  // Element mapping.  float: lane l of warp w owns elements e0 + 128 w + 32 j + l (j = 0..3), so one shared gather of a
  // warp covers 32 consecutive output elements (<= 32 consecutive source floats when up-scaling: no bank conflicts) and
  // each of the four 4-byte stores is a fully coalesced 128-byte line.  uint8: thread t owns the four consecutive
  // elements e0 + 4 t + j and packs them into one 4-byte store.
  static_assert(sizeof(T) == 1 || MAP == 1, "float32 uses the strided map");
  const int ef = MAP == 1 ? e0 + (t >> 5) * (32 * kQE) + (t & 31)                   // first of this thread's elements
               : MAP == 2 ? e0 + (t >> 5) * (32 * kQE) + 2 * (t & 31) : e0 + kQE * t;
  auto eoff = [](int j) { return MAP == 1 ? 32 * j : MAP == 2 ? (j >> 1) * 64 + (j & 1) : j; };   // offset of the j-th element
  if (ef >= DE) return;

  // per-element tap origin (unclamped: the staged tile replicates the borders) and packed horizontal coefficients
  int ob[kQE];
  float2 cx[4][kQE / 2];
#pragma unroll
  for (int j = 0; j < kQE; ++j) {
    const int e = min(ef + eoff(j), DE - 1);                                // (past the row end: repeat the last element)
    const int x = e / C, c = e - x * C;
    const float4 k = __ldg(reinterpret_cast<const float4*>(xtab[x].coef));
    ob[j] = __ldg(xbase + x) * C + c - c_lo;
    if (j & 1) { cx[0][j >> 1].y = k.x; cx[1][j >> 1].y = k.y; cx[2][j >> 1].y = k.z; cx[3][j >> 1].y = k.w; }
    else       { cx[0][j >> 1].x = k.x; cx[1][j >> 1].x = k.y; cx[2][j >> 1].x = k.z; cx[3][j >> 1].x = k.w; }
  }
  auto hrow = [&](int r, float2 (&w)[kQE / 2]) {                            // horizontal pass of tile row r
    const float* row = S + min(r, nr - 1) * pitch;
#pragma unroll
    for (int p = 0; p < kQE / 2; ++p) {
      const float* a = row + ob[2 * p];
      const float* b = row + ob[2 * p + 1];
      float2 v = __fmul2_rn(make_float2(a[0], b[0]), cx[0][p]);
      v = __ffma2_rn(make_float2(a[C], b[C]), cx[1][p], v);
      v = __ffma2_rn(make_float2(a[2 * C], b[2 * C]), cx[2][p], v);
      v = __ffma2_rn(make_float2(a[3 * C], b[3 * C]), cx[3][p], v);
      w[p] = v;
    }
  };
  float2 w0[kQE / 2], w1[kQE / 2], w2[kQE / 2], w3[kQE / 2];
  hrow(0, w0); hrow(1, w1); hrow(2, w2); hrow(3, w3);
  const bool all4 = ef + eoff(kQE - 1) < DE;
  T* out = dst + ((size_t)blockIdx.z * dst_h + y0) * DE + ef;

  // one output row from the window (a, b, c, d) = horizontal results of four consecutive source rows, oldest first
  auto emit = [&](const float2 (&a)[kQE / 2], const float2 (&b)[kQE / 2], const float2 (&c)[kQE / 2],
                  const float2 (&d)[kQE / 2], const float4 k) {
    float2 r[kQE / 2];
#pragma unroll
    for (int p = 0; p < kQE / 2; ++p) {
      float2 acc;
      if (FIXED) {
        // scalar on purpose: the product and the sum round separately (OpenCV's float32 vertical pass), and ptxas
        // contracts the packed mul.rn.f32x2 + add.rn.f32x2 pair into FFMA2 even when written as inline PTX
        acc.x = __fmul_rn(a[p].x, k.x);                      acc.y = __fmul_rn(a[p].y, k.x);
        acc.x = __fadd_rn(acc.x, __fmul_rn(b[p].x, k.y));    acc.y = __fadd_rn(acc.y, __fmul_rn(b[p].y, k.y));
        acc.x = __fadd_rn(acc.x, __fmul_rn(c[p].x, k.z));    acc.y = __fadd_rn(acc.y, __fmul_rn(c[p].y, k.z));
        acc.x = __fadd_rn(acc.x, __fmul_rn(d[p].x, k.w));    acc.y = __fadd_rn(acc.y, __fmul_rn(d[p].y, k.w));
      } else {
        acc = __fmul2_rn(a[p], make_float2(k.x, k.x));
        acc = __ffma2_rn(b[p], make_float2(k.y, k.y), acc);
        acc = __ffma2_rn(c[p], make_float2(k.z, k.z), acc);
        acc = __ffma2_rn(d[p], make_float2(k.w, k.w), acc);
      }
      r[p] = acc;
    }
    if (sizeof(T) == 1) {
      // saturate_cast<uchar>(rint(v)): adding 1.5 * 2^23 rounds half to even and leaves the integer in the low mantissa
      // bits (two's complement for negative v; |v| < 2^22 here), cvt.pack.sat then saturates two of them into bytes
      int q[kQE];
#pragma unroll
      for (int p = 0; p < kQE / 2; ++p) {
        const float2 m = __fadd2_rn(r[p], make_float2(12582912.f, 12582912.f));
        q[2 * p] = __float_as_int(m.x) - 0x4B400000;
        q[2 * p + 1] = __float_as_int(m.y) - 0x4B400000;
      }
      uint32_t hi, word;
      asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q[3]), "r"(q[2]), "r"(0));
      asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(q[1]), "r"(q[0]), "r"(hi));
      if (MAP == 1 && all4) {
#pragma unroll
        for (int j = 0; j < kQE; ++j) reinterpret_cast<uint8_t*>(out)[32 * j] = (uint8_t)(word >> (8 * j));
      } else if (MAP == 2 && vec_dst && all4) {
        *reinterpret_cast<uint16_t*>(out) = (uint16_t)word;
        *reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(out) + 64) = (uint16_t)(word >> 16);
      } else if (MAP == 0 && vec_dst) {
        *reinterpret_cast<uint32_t*>(out) = word;
      } else {
#pragma unroll
        for (int j = 0; j < kQE; ++j)
          if (ef + eoff(j) < DE) reinterpret_cast<uint8_t*>(out)[eoff(j)] = (uint8_t)(word >> (8 * j));
      }
    } else {
      float v[kQE];
#pragma unroll
      for (int p = 0; p < kQE / 2; ++p) { v[2 * p] = r[p].x; v[2 * p + 1] = r[p].y; }
      if (clip01) {
#pragma unroll
        for (int j = 0; j < kQE; ++j) v[j] = fminf(fmaxf(v[j], 0.f), 1.f);
      }
      if (all4) {
#pragma unroll
        for (int j = 0; j < kQE; ++j) reinterpret_cast<float*>(out)[eoff(j)] = v[j];
      } else {
#pragma unroll
        for (int j = 0; j < kQE; ++j)
          if (ef + eoff(j) < DE) reinterpret_cast<float*>(out)[eoff(j)] = v[j];
      }
    }
    out += DE;
  };
  // The window slides by renaming, not by moving registers: the four phases below are the four rotations of
  // (w0, w1, w2, w3).  A phase emits every output row whose first tap row is the window's, then replaces the oldest row.
  int i = 0, u = 0;                                                        // next output row / first source row of the window
#define SRB_QUAD_PHASE(A, B, C, D)                                     \
  while (i < rows && ub[i] <= u) { emit(A, B, C, D, cyv[i]); ++i; }    \
  if (i >= rows) break;                                                \
  ++u;                                                                 \
  hrow(u + 3, A);
  for (;;) {
    SRB_QUAD_PHASE(w0, w1, w2, w3)
    SRB_QUAD_PHASE(w1, w2, w3, w0)
    SRB_QUAD_PHASE(w2, w3, w0, w1)
    SRB_QUAD_PHASE(w3, w0, w1, w2)
  }
#undef SRB_QUAD_PHASE
}

// ---- cv2.resize(..., INTER_LANCZOS4) (classic_algorithms.py:19-21, 58-62; the interpolation map of
// loading_methods.py:133-145) ----------------------------------------------------------------------------------------
// Separable 8-tap Lanczos (a = 4) as OpenCV's interpolateLanczos4 evaluates it: position and fraction in float32, the
// eight sinc weights from one sin/cos pair in double through the pi/4 rotation table, 1e30 at a zero argument (which the
// normalisation turns into a one-hot row), float32 normalisation by the reciprocal of the float32 sum; taps
// floor(f) - 3 .. floor(f) + 4 with index clamping.
struct AxisTap8 { int idx[8]; float coef[8]; };   // 64 bytes

__global__ void lanczos4_tables(AxisTap8* __restrict__ tab, int* __restrict__ base_out, int n_src, int n_dst) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_dst) return;
  const double scale = 1.0 / ((double)n_dst / (double)n_src);
  const float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const int s = (int)floorf(f);
  const float t = __fsub_rn(f, (float)s);
  const double s45 = 0.70710678118654752440084436210485;
  const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double kPi = 3.1415926535897932384626433832795;
  const double y0 = -((double)t + 3.0) * kPi * 0.25;
  const double s0 = sin(y0), c0 = cos(y0);
  float c[8], sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float yi = __fsub_rn(__fadd_rn(t, 3.f), (float)i);
    if (fabsf(yi) >= 1e-6f) {
      const double y = -(double)yi * kPi * 0.25;
      c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else {
      c[i] = 1e30f;
    }
    sum = __fadd_rn(sum, c[i]);
  }
  sum = __fdiv_rn(1.f, sum);
  AxisTap8 a;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a.idx[i] = min(max(s - 3 + i, 0), n_src - 1);
    a.coef[i] = __fmul_rn(c[i], sum);
  }
  tab[d] = a;
  base_out[d] = s - 3;
}

// One interleaved output column per thread over a strip of output rows, with the horizontal-pass results of the eight
// source rows under the vertical taps kept in a register window (the 8-tap sibling of bicubic_stream_kernel).
__global__ void __launch_bounds__(kTE)
lanczos4_stream_kernel(const float* __restrict__ src, float* __restrict__ dst, const AxisTap8* __restrict__ xtab,
                       const AxisTap8* __restrict__ ytab, const int* __restrict__ ybase, int src_h, int src_w, int C,
                       int dst_h, int dst_w, int rows_per_block, int clip01) {
  const int DE = dst_w * C, SE = src_w * C;
  const int e = blockIdx.x * kTE + threadIdx.x;
  if (e >= DE) return;
  const int x = e / C, c = e - x * C;
  int o[8];
  float cx[8];
  {
    const AxisTap8 xt = xtab[x];
#pragma unroll
    for (int k = 0; k < 8; ++k) { o[k] = xt.idx[k] * C + c; cx[k] = xt.coef[k]; }
  }
  const int y0 = blockIdx.y * rows_per_block, y1 = min(y0 + rows_per_block, dst_h);
  const float* simg = src + (size_t)blockIdx.z * src_h * SE;
  float* out = dst + (size_t)blockIdx.z * dst_h * DE + (size_t)y0 * DE + e;
  auto hrow = [&](int r) -> float {
    r = min(max(r, 0), src_h - 1);
    const float* row = simg + (size_t)r * SE;
    float v = __fmul_rn(__ldg(row + o[0]), cx[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) v = __fmaf_rn(__ldg(row + o[k]), cx[k], v);
    return v;
  };
  int u = __ldg(ybase + y0);
  float w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = hrow(u + k);
  for (int y = y0; y < y1; ++y, out += DE) {
    const int ub = __ldg(ybase + y);                // block-uniform
    while (u < ub) {
#pragma unroll
      for (int k = 0; k < 7; ++k) w[k] = w[k + 1];
      ++u;
      w[7] = hrow(u + 7);
    }
    const float4 ca = __ldg(reinterpret_cast<const float4*>(ytab[y].coef));
    const float4 cb = __ldg(reinterpret_cast<const float4*>(ytab[y].coef + 4));
    float v = __fmul_rn(w[0], ca.x);
    v = __fmaf_rn(w[1], ca.y, v); v = __fmaf_rn(w[2], ca.z, v); v = __fmaf_rn(w[3], ca.w, v);
    v = __fmaf_rn(w[4], cb.x, v); v = __fmaf_rn(w[5], cb.y, v); v = __fmaf_rn(w[6], cb.z, v); v = __fmaf_rn(w[7], cb.w, v);
    if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
    *out = v;
  }
}

static int run_lanczos4(const float* src, int batch, int sh, int sw, int C, float* dst, int dh, int dw, int clip01,
                        cudaStream_t stream) {
  SRB_REQUIRE(src && dst, "lanczos4: null pointer");
  SRB_REQUIRE(batch >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && C > 0, "lanczos4: bad geometry");
  if (batch == 0) return SRB_OK;
  AxisTap8* tabs = nullptr;
  SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap8) * ((size_t)dw + dh) + sizeof(int) * ((size_t)dh + dw), stream));
  AxisTap8* xtab = tabs;
  AxisTap8* ytab = tabs + dw;
  int* ybase = reinterpret_cast<int*>(tabs + (size_t)dw + dh);
  int* xbase = ybase + dh;
  lanczos4_tables<<<(dw + 127) / 128, 128, 0, stream>>>(xtab, xbase, sw, dw);
  lanczos4_tables<<<(dh + 127) / 128, 128, 0, stream>>>(ytab, ybase, sh, dh);
  int rc = launch_check("lanczos4_tables");
  if (rc == SRB_OK) {
    int rows = 64;
    const long target = 4L * sm_count();
    while (rows > 8 && (long)((dw * C + kTE - 1) / kTE) * ((dh + rows - 1) / rows) * batch < target) rows >>= 1;
    dim3 grid((dw * C + kTE - 1) / kTE, (dh + rows - 1) / rows, batch);
    if (grid.y > 65535 || grid.z > 65535) {
      set_error("lanczos4: grid too large");
      rc = SRB_E_INVALID;
    } else {
      lanczos4_stream_kernel<<<grid, kTE, 0, stream>>>(src, dst, xtab, ytab, ybase, sh, sw, C, dh, dw, rows, clip01);
      rc = launch_check("lanczos4_stream_kernel");
    }
  }
  SRB_CUDA(cudaFreeAsync(tabs, stream));
  return rc;
}

// ---- uint8 INTER_LINEAR / INTER_AREA (up-scaling) / INTER_LANCZOS4: OpenCV's 11-bit fixed-point paths, bit-exact --------
// (classic_algorithms.py:7-9, 15-21 as super_resolucion_clasica.ipynb cell 7 calls them on uint8 images).  Coefficients are
// saturate_cast<short>(rint(c * 2048)); the horizontal pass is an int32 sum; the vertical pass is
//   linear / area : (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2      (VResizeLinear<uchar, int, short, ...>)
//   Lanczos-4     : (sum_k S_k * b_k + 2^21) >> 22                                       (FixedPtCast<int, uchar, 22>)
// One interleaved output column per thread; not a tuned path (the benchmark notebook resizes single images).
__device__ __forceinline__ int fix11(float c) { return (int)fminf(fmaxf(rintf(__fmul_rn(c, 2048.f)), -32768.f), 32767.f); }

__global__ void __launch_bounds__(kTE)
linear_u8_fixed_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const AxisTap* __restrict__ xtab,
                       const AxisTap* __restrict__ ytab, int src_h, int src_w, int C, int dst_h, int dst_w, int rows_per_block) {
  const int DE = dst_w * C, SE = src_w * C;
  const int e = blockIdx.x * kTE + threadIdx.x;
  if (e >= DE) return;
  const int x = e / C, c = e - x * C;
  const AxisTap xt = xtab[x];
  const int o1 = xt.idx[1] * C + c, o2 = xt.idx[2] * C + c;
  const int a1 = fix11(xt.coef[1]), a2 = fix11(xt.coef[2]);
  const int y0 = blockIdx.y * rows_per_block, y1 = min(y0 + rows_per_block, dst_h);
  const uint8_t* simg = src + (size_t)blockIdx.z * src_h * SE;
  uint8_t* out = dst + (size_t)blockIdx.z * dst_h * DE + (size_t)y0 * DE + e;
  for (int y = y0; y < y1; ++y, out += DE) {
    const AxisTap yt = ytab[y];
    const uint8_t* r0 = simg + (size_t)yt.idx[1] * SE;
    const uint8_t* r1 = simg + (size_t)yt.idx[2] * SE;
    const int s0 = (int)r0[o1] * a1 + (int)r0[o2] * a2, s1 = (int)r1[o1] * a1 + (int)r1[o2] * a2;
    const int b0 = fix11(yt.coef[1]), b1 = fix11(yt.coef[2]);
    const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
    *out = (uint8_t)min(max(v, 0), 255);
  }
}

__global__ void __launch_bounds__(kTE)
lanczos4_u8_fixed_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const AxisTap8* __restrict__ xtab,
                         const AxisTap8* __restrict__ ytab, int src_h, int src_w, int C, int dst_h, int dst_w, int rows_per_block) {
  const int DE = dst_w * C, SE = src_w * C;
  const int e = blockIdx.x * kTE + threadIdx.x;
  if (e >= DE) return;
  const int x = e / C, c = e - x * C;
  int o[8], a[8];
  {
    const AxisTap8 xt = xtab[x];
#pragma unroll
    for (int k = 0; k < 8; ++k) { o[k] = xt.idx[k] * C + c; a[k] = fix11(xt.coef[k]); }
  }
  const int y0 = blockIdx.y * rows_per_block, y1 = min(y0 + rows_per_block, dst_h);
  const uint8_t* simg = src + (size_t)blockIdx.z * src_h * SE;
  uint8_t* out = dst + (size_t)blockIdx.z * dst_h * DE + (size_t)y0 * DE + e;
  for (int y = y0; y < y1; ++y, out += DE) {
    const AxisTap8 yt = ytab[y];
    long long acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint8_t* row = simg + (size_t)yt.idx[k] * SE;
      int h = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) h += (int)row[o[j]] * a[j];
      acc += (long long)h * (long long)fix11(yt.coef[k]);
    }
    // (OpenCV accumulates in int32; |sum| < 2^31 for these coefficient magnitudes, so the wider accumulator agrees)
    const int v = (int)((acc + (1ll << 21)) >> 22);
    *out = (uint8_t)min(max(v, 0), 255);
  }
}

static int run_resize_u8_fixed(const uint8_t* src, int batch, int sh, int sw, int C, uint8_t* dst, int dh, int dw,
                               int interpolation, cudaStream_t stream) {
  SRB_REQUIRE(src && dst, "resize_u8: null pointer");
  SRB_REQUIRE(batch >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && C > 0, "resize_u8: bad geometry");
  if (batch == 0) return SRB_OK;
  int rows = 64;
  const long target = 4L * sm_count();
  while (rows > 8 && (long)((dw * C + kTE - 1) / kTE) * ((dh + rows - 1) / rows) * batch < target) rows >>= 1;
  dim3 grid((dw * C + kTE - 1) / kTE, (dh + rows - 1) / rows, batch);
  SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "resize_u8: grid too large");
  int rc;
  if (interpolation == SRB_INTER_LANCZOS4) {
    AxisTap8* tabs = nullptr;
    SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap8) * ((size_t)dw + dh) + sizeof(int) * ((size_t)dh + dw), stream));
    int* base = reinterpret_cast<int*>(tabs + (size_t)dw + dh);
    lanczos4_tables<<<(dw + 127) / 128, 128, 0, stream>>>(tabs, base + dh, sw, dw);
    lanczos4_tables<<<(dh + 127) / 128, 128, 0, stream>>>(tabs + dw, base, sh, dh);
    lanczos4_u8_fixed_kernel<<<grid, kTE, 0, stream>>>(src, dst, tabs, tabs + dw, sh, sw, C, dh, dw, rows);
    rc = launch_check("lanczos4_u8_fixed_kernel");
    SRB_CUDA(cudaFreeAsync(tabs, stream));
  } else {
    const int area = interpolation == SRB_INTER_AREA;
    AxisTap* tabs = nullptr;
    SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap) * ((size_t)dw + dh), stream));
    bicubic_tables<<<(dw + 127) / 128, 128, 0, stream>>>(tabs, nullptr, sw, dw, area ? 3 : 2);          // x: fraction clamped
    bicubic_tables<<<(dh + 127) / 128, 128, 0, stream>>>(tabs + dw, nullptr, sh, dh, area ? 5 : 4);     // y: row index clamped
    linear_u8_fixed_kernel<<<grid, kTE, 0, stream>>>(src, dst, tabs, tabs + dw, sh, sw, C, dh, dw, rows);
    rc = launch_check("linear_u8_fixed_kernel");
    SRB_CUDA(cudaFreeAsync(tabs, stream));
  }
  return rc;
}

template <typename T, bool FIXED>
static int run_bicubic(const T* src, int batch, int sh, int sw, int C, T* dst, int dh, int dw, int clip01,
                       cudaStream_t stream, int table_mode = FIXED ? 1 : 0) {
  SRB_REQUIRE(src && dst, "bicubic: null pointer");
  SRB_REQUIRE(batch >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && C > 0, "bicubic: bad geometry");
  if (batch == 0) return SRB_OK;
  {
    // the tap tables come from the stream-ordered allocator: keep its pool from handing memory back to the driver at every
    // synchronisation (the default release threshold is 0), which showed up as millisecond hiccups between launches
    static bool pool_set[64] = {};
    int dev = 0;
    SRB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !pool_set[dev]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_set[dev] = true;
    }
  }
  AxisTap* tabs = nullptr;
  SRB_CUDA(cudaMallocAsync(&tabs, sizeof(AxisTap) * ((size_t)dw + dh) + sizeof(int) * ((size_t)dh + dw), stream));
  int* ybase = reinterpret_cast<int*>(tabs + (size_t)dw + dh);
  int* xbase = ybase + dh;
  AxisTap* xtab = tabs;
  AxisTap* ytab = tabs + dw;
  bicubic_tables<<<(dw + 127) / 128, 128, 0, stream>>>(xtab, xbase, sw, dw, table_mode);
  bicubic_tables<<<(dh + 127) / 128, 128, 0, stream>>>(ytab, ybase, sh, dh, table_mode);
  int rc = launch_check("bicubic_tables");
  if (rc) return rc;
  {
    // default: shared-memory quad kernel; a block's source footprint is bounded from the scale ratios
    const int DE = dw * C, SE = sw * C;
    const double rx = (double)sw / (double)dw, ry = (double)sh / (double)dh;
    const int out_px = (kQT * kQE + C - 1) / C + 1;                       // output pixels a block can touch
    const int span_px = (int)(out_px * rx) + 7;                           // source pixels under their taps
    const int pitch = ((span_px * C + 4) + 3) & ~3;                       // (+4: c_lo is rounded down to a multiple of 4)
    auto rows_needed = [&](int n) { return (int)(n * ry) + 6; };
    auto need = [&](int n) { return (size_t)rows_needed(n) * pitch * sizeof(float) + (size_t)n * (sizeof(float4) + sizeof(int)); };
    int tr = 64;
    while (tr > 4 && (need(tr) > 48 * 1024 || tr >= 2 * dh)) tr >>= 1;
    if (need(tr) <= 96 * 1024) {
      const size_t smem = need(tr);
      static size_t configured = 0;                                       // (per template instantiation)
      if (smem > configured) {
        constexpr int kBase = sizeof(T) == 4 ? 1 : 0, kPair = sizeof(T) == 4 ? 1 : 2;   // (float32: every alias is map 1)
        SRB_CUDA(cudaFuncSetAttribute(bicubic_quad_kernel<T, FIXED, 0, kBase>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SRB_CUDA(cudaFuncSetAttribute(bicubic_quad_kernel<T, FIXED, 1, kBase>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SRB_CUDA(cudaFuncSetAttribute(bicubic_quad_kernel<T, FIXED, 4, kBase>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SRB_CUDA(cudaFuncSetAttribute(bicubic_quad_kernel<T, FIXED, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SRB_CUDA(cudaFuncSetAttribute(bicubic_quad_kernel<T, FIXED, 3, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      const uintptr_t salign = sizeof(T) == 4 ? 15u : 3u;
      const int vec_src = (SE % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & salign) == 0);
      const int vec_dst = (DE % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & salign) == 0);
      dim3 grid((DE + kQT * kQE - 1) / (kQT * kQE), (dh + tr - 1) / tr, batch);
      SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "bicubic: grid too large");
      constexpr int kBase = sizeof(T) == 4 ? 1 : 0, kPair = sizeof(T) == 4 ? 1 : 2;
      const bool pairs = rx <= 1.0 / 2.3;                 // (uint8 RGB only: element stride 2 keeps a gather's span under 32 floats)
      auto kern = C == 3 ? (pairs ? bicubic_quad_kernel<T, FIXED, 3, kPair> : bicubic_quad_kernel<T, FIXED, 3, 1>)
                : C == 1 ? bicubic_quad_kernel<T, FIXED, 1, kBase>
                : C == 4 ? bicubic_quad_kernel<T, FIXED, 4, kBase> : bicubic_quad_kernel<T, FIXED, 0, kBase>;
      kern<<<grid, kQT, smem, stream>>>(src, dst, xtab, xbase, ytab, ybase, sh, sw, C, dh, dw, tr, rows_needed(tr), pitch,
                                        vec_src, vec_dst, clip01);
      rc = launch_check("bicubic_quad_kernel");
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

extern "C" int srb_resize_f32(const float* src, int batch, int src_h, int src_w, int channels,
                              float* dst, int dst_h, int dst_w, int interpolation, int clip01, srb_stream_t stream) {
  if (interpolation == SRB_INTER_CUBIC)
    return run_bicubic<float, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, clip01, (cudaStream_t)stream);
  if (interpolation == SRB_INTER_LINEAR)
    return run_bicubic<float, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, clip01, (cudaStream_t)stream, 2);
  if (interpolation == SRB_INTER_AREA) {
    SRB_REQUIRE(dst_h >= src_h && dst_w >= src_w, "resize: INTER_AREA is built for up-scaling only (the reference's use); got %dx%d -> %dx%d",
                src_w, src_h, dst_w, dst_h);
    return run_bicubic<float, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, clip01, (cudaStream_t)stream, 3);
  }
  if (interpolation == SRB_INTER_LANCZOS4)
    return run_lanczos4(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, clip01, (cudaStream_t)stream);
  set_error("resize: unsupported interpolation code %d (linear = 1, cubic = 2, area = 3, lanczos4 = 4)", interpolation);
  return SRB_E_UNSUPPORTED;
}

extern "C" int srb_bicubic_u8(const uint8_t* src, int batch, int src_h, int src_w, int channels,
                              uint8_t* dst, int dst_h, int dst_w, int fixed_point, srb_stream_t stream) {
  if (fixed_point)
    return run_bicubic<uint8_t, true>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, 0, (cudaStream_t)stream);
  return run_bicubic<uint8_t, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, 0, (cudaStream_t)stream);
}

extern "C" int srb_resize_u8(const uint8_t* src, int batch, int src_h, int src_w, int channels,
                             uint8_t* dst, int dst_h, int dst_w, int interpolation, srb_stream_t stream) {
  if (interpolation == SRB_INTER_CUBIC)
    return run_bicubic<uint8_t, false>(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, 0, (cudaStream_t)stream);
  if (interpolation == SRB_INTER_AREA)
    SRB_REQUIRE(dst_h >= src_h && dst_w >= src_w, "resize_u8: INTER_AREA is built for up-scaling only (the reference's use); got %dx%d -> %dx%d",
                src_w, src_h, dst_w, dst_h);
  if (interpolation == SRB_INTER_LINEAR || interpolation == SRB_INTER_AREA || interpolation == SRB_INTER_LANCZOS4)
    return run_resize_u8_fixed(src, batch, src_h, src_w, channels, dst, dst_h, dst_w, interpolation, (cudaStream_t)stream);
  set_error("resize_u8: unsupported interpolation code %d (linear = 1, cubic = 2, area = 3, lanczos4 = 4)", interpolation);
  return SRB_E_UNSUPPORTED;
}
