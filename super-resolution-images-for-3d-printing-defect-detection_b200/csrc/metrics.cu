// Fused PSNR + SSIM (tf.image.psnr / tf.image.ssim semantics, max_val given; metrics.py:3-7).
//
// One pass over both images.  A block owns 128 consecutive interleaved output elements
// (x*C + c of the (H-10) x (W-10) SSIM map) and a strip of output rows.  Rows are streamed
// through a ring of shared-memory lines filled by cp.async four rows ahead of the arithmetic
// (no register staging, one barrier per row); each thread evaluates the 11-tap horizontal
// Gaussian of a, b, a^2+b^2 and a*b for its column and keeps the last 11 results in a register
// ring (the row loop is unrolled by 11 so ring slots are static), from which the vertical
// 11-tap pass and the SSIM point function follow.  The squared error for PSNR is accumulated
// from the same loaded lines over a non-overlapping ownership partition of the image.
// Per-image double-precision accumulators receive one atomicAdd per block and statistic.
#include "common.cuh"
#include "metrics_mma.cuh"
#include <stdlib.h>

namespace srb {

__constant__ float c_win[2][11];   // [0]: 11-tap Gaussian (sigma 1.5), [1]: 7-tap uniform window

constexpr int kCols = 128;  // output elements (threads) per block

template <int C, int K>   // K: window taps per axis (11: tf.image.ssim's Gaussian, 7: skimage's uniform window)
__global__ void __launch_bounds__(kCols)
psnr_ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int rows_per_strip,
                 float c1, float c2, float cov_norm, double* __restrict__ acc /* [B][2] = {sse, ssim_sum} */) {
  constexpr int kLine = kCols + (K - 1) * C;
  constexpr int kAhead = 4, kRing = 8;               // rows in flight / line buffers (power of two >= kAhead + 1)
  __shared__ float2 sab[kRing][kLine];               // (a, b) interleaved: one 8-byte shared load per tap
  __shared__ float red[2][kCols / 32];

  const int WE = W * C;            // interleaved floats per image row
  const int OE = (W - (K - 1)) * C;     // SSIM-map elements per row
  const int OH = H - (K - 1);
  const int t = threadIdx.x;
  const int e0 = blockIdx.x * kCols;
  const int y0 = blockIdx.y * rows_per_strip;
  const int rows_out = min(rows_per_strip, OH - y0);
  const int nin = rows_out + (K - 1);
  const bool last_x = (e0 + kCols >= OE);
  const bool last_y = (y0 + rows_per_strip >= OH);
  const size_t img_off = (size_t)blockIdx.z * H * WE;
  const float* pa = a + img_off;
  const float* pb = b + img_off;
  const bool col_valid = (e0 + t) < OE;

  // packed fp32x2 arithmetic (FFMA2 / FMUL2, sm_100): the maps are carried as (mu_a, mu_b), (E[a^2], E[b^2]) pairs
  // plus E[ab]; tf.image.ssim's E[a^2 + b^2] is the sum of the second pair
  float2 g2[K];
#pragma unroll
  for (int k = 0; k < K; ++k) g2[k] = make_float2(c_win[K == 11 ? 0 : 1][k], c_win[K == 11 ? 0 : 1][k]);

  float2 rab[K], rqq[K];
  float rp[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { rab[k] = rqq[k] = make_float2(0.f, 0.f); rp[k] = 0.f; }

  float sse = 0.f, ssim_sum = 0.f;

  // asynchronous fill of one line pair (row is block-uniform).  A thread always fetches the same (at most two)
  // line positions, so the range checks and addresses are hoisted; columns past the image stay zero
  static_assert(kLine <= 2 * kCols, "two fetch slots per thread");
  const bool ok0 = (e0 + t) < WE;
  const bool ok1 = (t + kCols) < kLine && (e0 + t + kCols) < WE;
  for (int i = t; i < kRing * kLine; i += kCols) (&sab[0][0])[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(&sab[0][t]);
  const float* fa = pa + (size_t)y0 * WE + e0 + t;
  const float* fb = pb + (size_t)y0 * WE + e0 + t;
  auto fetch = [&](int row) {
    if (row < nin) {
      const uint32_t d = s0 + (uint32_t)(row & (kRing - 1)) * (uint32_t)(kLine * sizeof(float2));
      if (ok0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(fa) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u), "l"(fb) : "memory");
      }
      if (ok1) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + (uint32_t)(kCols * sizeof(float2))), "l"(fa + kCols) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + (uint32_t)(kCols * sizeof(float2)) + 4u), "l"(fb + kCols) : "memory");
      }
      fa += WE;
      fb += WE;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per row, empty past the end
  };
#pragma unroll
  for (int r = 0; r < kAhead; ++r) fetch(r);

  for (int r = 0; r < nin; r += K) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int row = r + j;              // block-uniform
      if (row < nin) {
        const int buf = row & (kRing - 1);
        const bool row_owned = (row < rows_out) || last_y;
        asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 1) : "memory");   // this row's line has landed
        __syncthreads();                  // ... for every thread, and line (row-1) % kRing is no longer read
        fetch(row + kAhead);
        if (row_owned) {                  // squared error: every element of the row exactly once
          { const float2 v = sab[buf][t]; const float d = v.x - v.y; sse = fmaf(d, d, sse); }
          if (last_x && t < (K - 1) * C) { const float2 v = sab[buf][kCols + t]; const float d = v.x - v.y; sse = fmaf(d, d, sse); }
        }
        float2 hab = make_float2(0.f, 0.f), hqq = make_float2(0.f, 0.f);
        float hp = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float2 v = sab[buf][t + k * C];
          hab = __ffma2_rn(g2[k], v, hab);
          hqq = __ffma2_rn(g2[k], __fmul2_rn(v, v), hqq);
          hp = fmaf(g2[k].x, v.x * v.y, hp);
        }
        rab[j] = hab; rqq[j] = hqq; rp[j] = hp;
        if (row >= K - 1 && col_valid) {
          float2 mab = make_float2(0.f, 0.f), eqq = make_float2(0.f, 0.f);
          float ep = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int slot = (j + 1 + k) % K;    // oldest row first
            mab = __ffma2_rn(g2[k], rab[slot], mab);
            eqq = __ffma2_rn(g2[k], rqq[slot], eqq);
            ep = fmaf(g2[k].x, rp[slot], ep);
          }
          const float ma = mab.x, mb = mab.y, es = eqq.x + eqq.y;
          const float num0 = 2.f * ma * mb;
          const float den0 = fmaf(ma, ma, mb * mb);
          // cov_norm = 1 (tf.image.ssim, population moments) or N / (N - 1) (skimage's sample covariance)
          const float num = (num0 + c1) * fmaf(cov_norm, 2.f * ep - num0, c2);
          const float den = (den0 + c1) * fmaf(cov_norm, es - den0, c2);
          ssim_sum += num / den;
        }
      }
    }
  }

  sse = warp_sum(sse);
  ssim_sum = warp_sum(ssim_sum);
  if ((t & 31) == 0) { red[0][t >> 5] = sse; red[1][t >> 5] = ssim_sum; }
  __syncthreads();
  if (t == 0) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < kCols / 32; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
    atomicAdd(&acc[2 * blockIdx.z + 0], s0);
    atomicAdd(&acc[2 * blockIdx.z + 1], s1);
  }
}

// Wide-image variant: two horizontally adjacent map pixels per thread, one warp per (channel, 64-pixel column block).
// A warp de-interleaves its own channel plane when it fetches a line ((a, b) pairs per pixel, cp.async a few rows ahead),
// and every lane turns the pixels it fetched itself into a second plane of (a^2 + b^2, a*b) pairs one row ahead of the
// arithmetic, so the products are formed once per input element instead of once per tap.  The four maps tf.image.ssim
// filters (mu_a, mu_b, E[a^2 + b^2], E[ab]) are then exactly two packed fp32x2 FMAs per tap and output in both passes,
// and a lane's two outputs share 10 of their 11 tap loads (six contiguous 16-byte shared loads per plane and row).
// Warps never exchange data, so the row loop synchronises with __syncwarp only; the strided 4-byte fetches of the three
// channel warps of a block meet in L1.
constexpr int kPairPx = 64;                        // map pixels per warp row (2 per lane)
constexpr int kPairLW = kPairPx + 10;              // line width in pixels (even: 16-byte aligned planes)
constexpr int kPairAhead = 8;                      // rows in flight = line buffers (refilled after the row is retired)

// (Three resident blocks per SM, 157 registers.  Holding the allocation to five blocks (128 registers, no spills) or six
//  (96, spills) makes it SLOWER - 60 -> 49 -> 39 GP/s at 1K / 2K / 4K, profiles/r02_ssim_occupancy.txt: the kernel is bound
//  by shared-memory wavefronts (48 tap-load + 12 product + 6 fill wavefronts per 64 outputs against 44 FFMA2 issue cycles),
//  not by latency, so more warps only add instructions.)
template <int C>
__global__ void __launch_bounds__((C == 3 ? 3 : 4) * 32, 3)
psnr_ssim_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int rows_per_strip,
                      float c1, float c2, double* __restrict__ acc /* [B][2] = {sse, ssim_sum} */) {
  constexpr int NW = C == 3 ? 3 : 4;               // warps per block
  constexpr int CB = NW / C;                       // column blocks per block (4, 2, 1, 1 for C = 1..4)
  constexpr int LW = kPairLW, kAhead = kPairAhead;
  __shared__ __align__(16) float2 sab[NW][kAhead][LW];   // (a, b)
  __shared__ __align__(16) float2 ssp[NW][2][LW];        // (a^2 + b^2, a*b)
  __shared__ float red[2][NW];

  const int WE = W * C;
  const int OW = W - 10, OH = H - 10;
  const int t = threadIdx.x, wid = t >> 5, lane = t & 31;
  const int c = wid % C;
  const int X0 = (blockIdx.x * CB + wid / C) * kPairPx;
  const int y0 = blockIdx.y * rows_per_strip;
  const int rows_out = min(rows_per_strip, OH - y0);
  const int nin = X0 < OW ? rows_out + 10 : 0;     // (a warp past the right edge only joins the final reduction)
  const bool last_x = (X0 + kPairPx >= OW);
  const bool last_y = (y0 + rows_per_strip >= OH);
  const size_t img_off = (size_t)blockIdx.z * H * WE;

  float2 g2[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) g2[k] = make_float2(c_win[0][k], c_win[0][k]);

  // fetch slots: pixels lane, lane + 32 (the warp's own 64) and lane + 64 (the 10-pixel halo) of the warp's plane
  const bool in2 = lane < LW - 64;
  const bool ok0 = X0 + lane < W, ok1 = X0 + lane + 32 < W, ok2 = in2 && X0 + lane + 64 < W;
  for (int i = lane; i < kAhead * LW; i += 32) (&sab[wid][0][0])[i] = make_float2(0.f, 0.f);
  __syncwarp();
  const uint32_t sab0 = (uint32_t)__cvta_generic_to_shared(&sab[wid][0][lane]);
  constexpr uint32_t kLineBytes = LW * sizeof(float2);
  const float* fa = a + img_off + (size_t)y0 * WE + (size_t)(X0 + lane) * C + c;
  const float* fb = b + img_off + (size_t)y0 * WE + (size_t)(X0 + lane) * C + c;
  auto fetch = [&](int row) {
    if (row < nin) {
      const uint32_t d = sab0 + (uint32_t)(row & (kAhead - 1)) * kLineBytes;
      if (ok0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(fa) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u), "l"(fb) : "memory");
      }
      if (ok1) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 256u), "l"(fa + 32 * C) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 260u), "l"(fb + 32 * C) : "memory");
      }
      if (ok2) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 512u), "l"(fa + 64 * C) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 516u), "l"(fb + 64 * C) : "memory");
      }
      fa += WE;
      fb += WE;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float sse = 0.f;
  // products of this lane's own pixels of a landed row (its own cp.async groups are complete: no barrier needed)
  auto products = [&](int row) {
    if (row < nin) {
      const bool row_owned = (row < rows_out) || last_y;
      const float2* lab = &sab[wid][row & (kAhead - 1)][lane];
      float2* lsp = &ssp[wid][row & 1][lane];
#pragma unroll
      for (int s = 0; s < 3; ++s)
        if (s < 2 || in2) {
          const float2 v = lab[32 * s];
          const float2 sq = __fmul2_rn(v, v);
          lsp[32 * s] = make_float2(sq.x + sq.y, v.x * v.y);
          if (row_owned && (s < 2 || last_x)) { const float d = v.x - v.y; sse = fmaf(d, d, sse); }
        }
    }
  };

  float2 rab0[11], rab1[11], rsp0[11], rsp1[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) rab0[k] = rab1[k] = rsp0[k] = rsp1[k] = make_float2(0.f, 0.f);
  float2 ssim2 = make_float2(0.f, 0.f);
  const float2 valid2 = make_float2((X0 + 2 * lane) < OW ? 1.f : 0.f, (X0 + 2 * lane + 1) < OW ? 1.f : 0.f);
  const float2 c1_2 = make_float2(c1, c1), c2_2 = make_float2(c2, c2), two2 = make_float2(2.f, 2.f);

#pragma unroll
  for (int r = 0; r < kAhead; ++r) fetch(r);
  asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 1) : "memory");
  products(0);
  __syncwarp();

  for (int r = 0; r < nin; r += 11) {
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int row = r + j;              // warp-uniform
      if (row < nin) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 2) : "memory");   // this lane's part of row + 1 has landed
        products(row + 1);
        const float4* lab = reinterpret_cast<const float4*>(&sab[wid][row & (kAhead - 1)][2 * lane]);
        const float4* lsp = reinterpret_cast<const float4*>(&ssp[wid][row & 1][2 * lane]);
        float2 hab0 = make_float2(0.f, 0.f), hab1 = hab0, hsp0 = hab0, hsp1 = hab0;
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const float4 va = lab[m], vs = lsp[m];
          const float2 a0 = make_float2(va.x, va.y), a1 = make_float2(va.z, va.w);
          const float2 p0 = make_float2(vs.x, vs.y), p1 = make_float2(vs.z, vs.w);
          hab0 = __ffma2_rn(g2[2 * m], a0, hab0);
          hsp0 = __ffma2_rn(g2[2 * m], p0, hsp0);
          if (m >= 1) { hab1 = __ffma2_rn(g2[2 * m - 1], a0, hab1); hsp1 = __ffma2_rn(g2[2 * m - 1], p0, hsp1); }
          if (m <= 4) { hab0 = __ffma2_rn(g2[2 * m + 1], a1, hab0); hsp0 = __ffma2_rn(g2[2 * m + 1], p1, hsp0); }
          hab1 = __ffma2_rn(g2[2 * m], a1, hab1);
          hsp1 = __ffma2_rn(g2[2 * m], p1, hsp1);
        }
        rab0[j] = hab0; rab1[j] = hab1; rsp0[j] = hsp0; rsp1[j] = hsp1;
        if (row >= 10) {
          float2 mab0 = make_float2(0.f, 0.f), mab1 = mab0, esp0 = mab0, esp1 = mab0;
#pragma unroll
          for (int k = 0; k < 11; ++k) {
            const int slot = (j + 1 + k) % 11;   // oldest row first
            mab0 = __ffma2_rn(g2[k], rab0[slot], mab0);
            mab1 = __ffma2_rn(g2[k], rab1[slot], mab1);
            esp0 = __ffma2_rn(g2[k], rsp0[slot], esp0);
            esp1 = __ffma2_rn(g2[k], rsp1[slot], esp1);
          }
          // point function for both outputs at once
          const float2 ma = make_float2(mab0.x, mab1.x), mb = make_float2(mab0.y, mab1.y);
          const float2 es = make_float2(esp0.x, esp1.x), ep = make_float2(esp0.y, esp1.y);
          const float2 mm = __fmul2_rn(ma, mb);                       // mu_a mu_b
          const float2 den0 = __ffma2_rn(ma, ma, __fmul2_rn(mb, mb)); // mu_a^2 + mu_b^2
          const float2 ln = __ffma2_rn(two2, mm, c1_2);
          const float2 cn = __ffma2_rn(two2, make_float2(ep.x - mm.x, ep.y - mm.y), c2_2);
          const float2 ld = make_float2(den0.x + c1, den0.y + c1);
          const float2 cd = make_float2((es.x - den0.x) + c2, (es.y - den0.y) + c2);
          const float2 num = __fmul2_rn(ln, cn), den = __fmul2_rn(ld, cd);
          ssim2 = __ffma2_rn(make_float2(__fdividef(num.x, den.x), __fdividef(num.y, den.y)), valid2, ssim2);
        }
        __syncwarp();                     // row is retired by every lane; row + 1 (both planes) is visible
        fetch(row + kAhead);
      }
    }
  }

  sse = warp_sum(sse);
  const float ssim_sum = warp_sum(ssim2.x + ssim2.y);
  if (lane == 0) { red[0][wid] = sse; red[1][wid] = ssim_sum; }
  __syncthreads();
  if (t == 0) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
    atomicAdd(&acc[2 * blockIdx.z + 0], s0);
    atomicAdd(&acc[2 * blockIdx.z + 1], s1);
  }
}

// PSNR alone (ssim == NULL at the C ABI): the squared error is a pure streaming reduction - 24 B per RGB pixel, 16-byte
// loads, eight independent loads in flight per thread - and runs at the HBM roofline instead of the SSIM kernel's FP32 bound.
constexpr int kSseThreads = 256, kSseUnroll = 4;

__global__ void __launch_bounds__(kSseThreads)
sse_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n_per_image, int vec, double* __restrict__ acc) {
  __shared__ float red[kSseThreads / 32];
  const float* pa = a + (size_t)blockIdx.y * n_per_image;
  const float* pb = b + (size_t)blockIdx.y * n_per_image;
  float s0 = 0.f, s1 = 0.f;
  if (vec) {
    const size_t n4 = n_per_image >> 2;
    const float4* va = reinterpret_cast<const float4*>(pa);
    const float4* vb = reinterpret_cast<const float4*>(pb);
    const size_t step = (size_t)gridDim.x * kSseThreads;
    size_t i = (size_t)blockIdx.x * kSseThreads + threadIdx.x;
    for (; i + (kSseUnroll - 1) * step < n4; i += kSseUnroll * step) {
      float4 x[kSseUnroll], y[kSseUnroll];
#pragma unroll
      for (int u = 0; u < kSseUnroll; ++u) { x[u] = __ldcs(va + i + u * step); y[u] = __ldcs(vb + i + u * step); }
#pragma unroll
      for (int u = 0; u < kSseUnroll; ++u) {
        const float d0 = x[u].x - y[u].x, d1 = x[u].y - y[u].y, d2 = x[u].z - y[u].z, d3 = x[u].w - y[u].w;
        s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s0 = fmaf(d2, d2, s0); s1 = fmaf(d3, d3, s1);
      }
    }
    for (; i < n4; i += step) {
      const float4 x = __ldcs(va + i), y = __ldcs(vb + i);
      const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s0 = fmaf(d2, d2, s0); s1 = fmaf(d3, d3, s1);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * kSseThreads + threadIdx.x; i < n_per_image; i += (size_t)gridDim.x * kSseThreads) {
      const float d = __ldg(pa + i) - __ldg(pb + i);
      s0 = fmaf(d, d, s0);
    }
  }
  const float w = warp_sum(s0 + s1);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kSseThreads / 32; ++k) t += red[k];
    atomicAdd(&acc[2 * blockIdx.y], t);
  }
}

__global__ void psnr_ssim_finalize(const double* __restrict__ acc, int B, double n_pix, double n_map,
                                   float max_val, int skimage, float* psnr, float* ssim, float* mse_out, double* sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double mse = acc[2 * i] / n_pix;
  // tf.image.psnr: 20*log(max)/log(10) - 10/log(10)*log(mse), evaluated in float32;
  // skimage.metrics.peak_signal_noise_ratio: 10*log10(data_range^2 / mse) in float64
  const float p = skimage ? (float)(10.0 * log10((double)max_val * (double)max_val / mse))
                          : 20.f * log10f(max_val) - 10.f * log10f((float)mse);
  const float s = (float)(acc[2 * i + 1] / n_map);
  if (psnr) psnr[i] = p;
  if (ssim) ssim[i] = s;
  if (mse_out) mse_out[i] = (float)mse;
  if (sums) {
    atomicAdd(&sums[0], (double)p);
    atomicAdd(&sums[1], (double)s);
    atomicAdd(&sums[2], 1.0);
    atomicAdd(&sums[3], mse);
  }
}

static bool g_win_ready[64] = {};

static int upload_windows() {
  int dev = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && g_win_ready[dev]) return SRB_OK;
  double g[11], sum = 0.0;
  for (int i = 0; i < 11; ++i) { const double c = i - 5.0; g[i] = exp(-0.5 * c * c / (1.5 * 1.5)); sum += g[i]; }
  float wf[2][11] = {};
  for (int i = 0; i < 11; ++i) wf[0][i] = (float)(g[i] / sum);
  for (int i = 0; i < 7; ++i) wf[1][i] = (float)(1.0 / 7.0);
  SRB_CUDA(cudaMemcpyToSymbol(c_win, wf, sizeof(wf)));
  if (dev >= 0 && dev < 64) g_win_ready[dev] = true;
  return SRB_OK;
}

// window: SRB_SSIM_TF (11-tap Gaussian, population moments) or SRB_SSIM_SKIMAGE (7 x 7 uniform, sample covariance)
static int run_psnr_ssim(const float* a, const float* b, int batch, int height, int width, int channels, float max_val,
                         int window, float* psnr, float* ssim, float* mse, double* sums, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
  SRB_REQUIRE(a && b && workspace, "psnr_ssim: null pointer");
  SRB_REQUIRE(window == SRB_SSIM_TF || window == SRB_SSIM_SKIMAGE || window == SRB_SSIM_TF_EXACT,
              "psnr_ssim: unknown window kind %d", window);
  SRB_REQUIRE(batch >= 0 && channels >= 1 && channels <= 4, "psnr_ssim: channels must be 1..4 (got %d)", channels);
  const int K = window == SRB_SSIM_SKIMAGE ? 7 : 11;
  SRB_REQUIRE(workspace_bytes >= srb_psnr_ssim_workspace(batch), "psnr_ssim: workspace too small");
  double* acc = (double*)workspace;
  if (!ssim) {
    // PSNR / MSE only: streaming squared-error reduction (no window, so no minimum image size either)
    SRB_REQUIRE(height >= 1 && width >= 1, "psnr: bad geometry %dx%d", height, width);
    if (batch == 0) return SRB_OK;
    SRB_CUDA(cudaMemsetAsync(acc, 0, srb_psnr_ssim_workspace(batch), stream));
    const size_t n = (size_t)height * width * channels;
    const int vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    const size_t work = vec ? n / 4 : n;                                   // loads per image and array
    long per_image = (long)((work + (size_t)kSseThreads * kSseUnroll * 4 - 1) / ((size_t)kSseThreads * kSseUnroll * 4));
    const long cap = (32L * sm_count() + batch - 1) / batch;              // enough blocks to fill the chip, few atomics
    if (per_image > cap) per_image = cap;
    if (per_image < 1) per_image = 1;
    SRB_REQUIRE(batch <= 65535, "psnr: batch too large for one launch");
    dim3 grid((unsigned)per_image, batch);
    sse_kernel<<<grid, kSseThreads, 0, stream>>>(a, b, n, vec, acc);
    int rc = launch_check("sse_kernel");
    if (rc) return rc;
    psnr_ssim_finalize<<<(batch + 127) / 128, 128, 0, stream>>>(acc, batch, (double)n, 1.0, max_val,
                                                                 window == SRB_SSIM_SKIMAGE, psnr, nullptr, mse, sums);
    return launch_check("psnr_ssim_finalize");
  }
  SRB_REQUIRE(height >= K && width >= K, "psnr_ssim: image dimensions must be at least %dx%d (got %dx%d)", K, K, height, width);
  if (batch == 0) return SRB_OK;
  int rc = upload_windows();
  if (rc) return rc;
  SRB_CUDA(cudaMemsetAsync(acc, 0, srb_psnr_ssim_workspace(batch), stream));
  const int OH = height - (K - 1), OW = width - (K - 1), OE = OW * channels;
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  const float cov_norm = window == SRB_SSIM_SKIMAGE ? (float)(49.0 / 48.0) : 1.f;
  static const bool force_narrow = getenv("SRB_SSIM_NARROW") != nullptr;
  static const bool force_exact = getenv("SRB_SSIM_EXACT") != nullptr;
  if (window == SRB_SSIM_TF && !force_exact && !force_narrow && psnr_ssim_mma_eligible(a, b, height, width, channels, max_val)) {
    // wide RGB / grey images: both filter passes on the warp-level tensor path (fp16 window summing to 1, hi/lo-split data)
    rc = run_psnr_ssim_mma(a, b, batch, height, width, channels, c1, c2, acc, stream);
  } else if (window != SRB_SSIM_SKIMAGE && OW >= 96 && !force_narrow) {   // wide images: two map pixels per thread, one warp per plane
    const int px = kPairPx * (channels == 1 ? 4 : channels == 2 ? 2 : 1);   // map pixels per block row
    const int gxp = (OW + px - 1) / px;
    int rows = 256;                                // (10 halo rows per strip: 4 % extra horizontal work at 256)
    const long target = 8L * sm_count();
    while (rows > 16 && (long)gxp * ((OH + rows - 1) / rows) * batch < target) rows >>= 1;
    dim3 grid(gxp, (OH + rows - 1) / rows, batch);
    SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "psnr_ssim: grid too large");
    switch (channels) {
      case 1: psnr_ssim_pair_kernel<1><<<grid, 128, 0, stream>>>(a, b, height, width, rows, c1, c2, acc); break;
      case 2: psnr_ssim_pair_kernel<2><<<grid, 128, 0, stream>>>(a, b, height, width, rows, c1, c2, acc); break;
      case 3: psnr_ssim_pair_kernel<3><<<grid, 96, 0, stream>>>(a, b, height, width, rows, c1, c2, acc); break;
      default: psnr_ssim_pair_kernel<4><<<grid, 128, 0, stream>>>(a, b, height, width, rows, c1, c2, acc); break;
    }
    rc = launch_check("psnr_ssim_pair_kernel");
  } else {
    const int gx = (OE + kCols - 1) / kCols;
    int rows = 64;
    const long target = 2L * sm_count();
    while (rows > 16 && (long)gx * ((OH + rows - 1) / rows) * batch < target) rows >>= 1;
    dim3 grid(gx, (OH + rows - 1) / rows, batch);
    SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "psnr_ssim: grid too large");
#define SRB_SSIM_LAUNCH(CH, KK) psnr_ssim_kernel<CH, KK><<<grid, kCols, 0, stream>>>(a, b, height, width, rows, c1, c2, cov_norm, acc)
    if (K == 11) {
      switch (channels) {
        case 1: SRB_SSIM_LAUNCH(1, 11); break;
        case 2: SRB_SSIM_LAUNCH(2, 11); break;
        case 3: SRB_SSIM_LAUNCH(3, 11); break;
        default: SRB_SSIM_LAUNCH(4, 11); break;
      }
    } else {
      switch (channels) {
        case 1: SRB_SSIM_LAUNCH(1, 7); break;
        case 2: SRB_SSIM_LAUNCH(2, 7); break;
        case 3: SRB_SSIM_LAUNCH(3, 7); break;
        default: SRB_SSIM_LAUNCH(4, 7); break;
      }
    }
#undef SRB_SSIM_LAUNCH
    rc = launch_check("psnr_ssim_kernel");
  }
  if (rc) return rc;
  psnr_ssim_finalize<<<(batch + 127) / 128, 128, 0, stream>>>(
      acc, batch, (double)height * width * channels, (double)OH * OE, max_val, window == SRB_SSIM_SKIMAGE, psnr, ssim, mse, sums);
  return launch_check("psnr_ssim_finalize");
}

}  // namespace srb

using namespace srb;

extern "C" size_t srb_psnr_ssim_workspace(int batch) { return (size_t)(batch > 0 ? batch : 0) * 2 * sizeof(double); }

extern "C" int srb_psnr_ssim_f32(const float* a, const float* b, int batch, int height, int width, int channels,
                                 float max_val, float* psnr, float* ssim, float* mse, double* sums,
                                 void* workspace, size_t workspace_bytes, srb_stream_t stream) {
  return run_psnr_ssim(a, b, batch, height, width, channels, max_val, SRB_SSIM_TF, psnr, ssim, mse, sums, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

extern "C" int srb_psnr_ssim_window_f32(const float* a, const float* b, int batch, int height, int width, int channels,
                                        float max_val, int window, float* psnr, float* ssim, float* mse, double* sums,
                                        void* workspace, size_t workspace_bytes, srb_stream_t stream) {
  return run_psnr_ssim(a, b, batch, height, width, channels, max_val, window, psnr, ssim, mse, sums, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}
