// RGB head layers: 3x3, Cin = 3 convolutions on fp32 NHWC images (EDSR_model.py:102, ESRGAN_model.py:314,
// VGG16 block1_conv1).  K = 27 is far too small for the tensor cores and the layer is bandwidth-bound
// (12 B in, up to 384 B out per pixel), so this is a CUDA-core kernel built around coalesced stores:
//   * a block owns an 8 x 16 pixel tile; the 10 x 18 x 3 input halo and the [27][Cout] filter sit in smem;
//   * lanes are split as (pixel, channel quad): Cout/4 consecutive lanes hold the 4-channel groups of ONE
//     pixel, so every store instruction writes whole contiguous pixels (256 B fp32 / 128 B 16-bit for Cout=64);
//   * a thread accumulates 4 vertically adjacent pixels x 4 channels, re-using each filter row (9 taps x 4
//     channels, one 16-B shared load per tap) across the 4 pixels;
//   * bias, activation and alpha are fused; the result is written as fp32 and/or 16-bit (y, y2).
#include "common.cuh"
#include "conv_common.cuh"

namespace srb {

constexpr int kHT_H = 8, kHT_W = 16;       // tile
constexpr int kHaloW = kHT_W + 2, kHaloH = kHT_H + 2;

__device__ __forceinline__ void head_store4(void* base, int dtype, size_t idx, const float (&v)[4]) {
  if (dtype == SRB_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (dtype == SRB_F8E5M2) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[0], v[1]), __NV_SATFINITE, __NV_E5M2);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[2], v[3]), __NV_SATFINITE, __NV_E5M2);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(base) + idx) = lo | (hi << 16);
  } else if (dtype == SRB_BF16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  } else {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

__global__ void __launch_bounds__(256, 3)
conv_head_kernel(const ConvParams p, int tiles_x, int tiles_y) {
  extern __shared__ float hsm[];
  constexpr int kHaloN = kHaloH * kHaloW * 3;          // 540 floats (keeps everything after it 16-B aligned)
  float* halo2 = hsm;                                  // [2][kHaloH][kHaloW][3], double buffered
  float* wsm = hsm + 2 * kHaloN;                       // [27][cout]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lpp = p.cout >> 2;                         // lanes per pixel
  const int ppw = 32 / lpp;                            // pixels per warp instruction
  const int cq = lane % lpp, psub = lane / lpp;

  for (int i = tid; i < 27 * p.cout; i += 256) {
    const int k = i / p.cout, co = i - k * p.cout;
    wsm[i] = __ldg(p.w_hwio + (size_t)k * p.w_cout_pad + co);
  }
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias[j] = __ldg(p.bias + cq * 4 + j);

  // the (at most three) halo elements this thread fetches for every tile: position inside the halo is tile-invariant
  int h_off[3], h_dy[3], h_dx[3], h_c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = tid + k * 256;
    const int hp = i / 3;
    h_off[k] = i < kHaloN ? i : -1;
    h_c[k] = i - hp * 3;
    h_dx[k] = hp % kHaloW - 1;
    h_dy[k] = hp / kHaloW - 1;
  }
  const int tiles_per_img = tiles_x * tiles_y;
  const int total = p.B * tiles_per_img;
  const float* xin = reinterpret_cast<const float*>(p.x);
  auto fetch = [&](int tile, int buf) {                // asynchronous halo fill; out-of-image pixels are zero (same padding)
    if (tile < total) {
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
      const int y0 = (r / tiles_x) * kHT_H, x0 = (r % tiles_x) * kHT_W;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (h_off[k] >= 0) {
          float* d = halo2 + buf * kHaloN + h_off[k];
          const int gy = y0 + h_dy[k], gx = x0 + h_dx[k];
          if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
            const float* src = xin + (((size_t)b * p.H + gy) * p.W + gx) * p.x_cstride + p.x_coffset + h_c[k];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(d)), "l"(src) : "memory");
          } else {
            *d = 0.f;
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int y0 = (r / tiles_x) * kHT_H, x0 = (r % tiles_x) * kHT_W;
    const float* halo = halo2 + (it & 1) * kHaloN;
    __syncthreads();                                   // everyone is done with the buffer the next fetch overwrites
    fetch(tile + gridDim.x, (it + 1) & 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                   // this tile's halo (and, first time, the filter) is visible
    // columns of the tile are dealt to (warp, psub) pairs; each thread does 4 vertically adjacent pixels
    for (int col = warp * ppw + psub; col < kHT_W; col += 8 * ppw) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int row0 = half * 4;
        // packed fp32x2 FMAs (FFMA2): channel pairs (0,1) and (2,3) of each of the 4 pixels
        float2 acc01[4], acc23[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc01[q] = make_float2(bias[0], bias[1]); acc23[q] = make_float2(bias[2], bias[3]); }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float4 w[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const float4*>(wsm + (dy * 9 + t) * p.cout + cq * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float* hrow = halo + ((row0 + q + dy) * kHaloW + col) * 3;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float a = hrow[t];
              const float2 aa = make_float2(a, a);
              acc01[q] = __ffma2_rn(aa, make_float2(w[t].x, w[t].y), acc01[q]);
              acc23[q] = __ffma2_rn(aa, make_float2(w[t].z, w[t].w), acc23[q]);
            }
          }
        }
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[q][0] = acc01[q].x; acc[q][1] = acc01[q].y; acc[q][2] = acc23[q].x; acc[q][3] = acc23[q].y; }
        const int ox = x0 + col;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int oy = y0 + row0 + q;
          if (oy < p.H && ox < p.W) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float slope = (p.act == SRB_ACT_PRELU) ? __ldg(p.prelu + cq * 4 + j) : p.act_slope;
              v[j] = apply_act(acc[q][j], p.act, slope) * p.alpha;
              if (p.clip01) v[j] = fminf(fmaxf(v[j], 0.f), 1.f);
            }
            const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
            head_store4(p.y, p.y_dtype, pix * p.y_cstride + p.y_coffset + cq * 4, v);
            if (p.y2) {
              if (p.y2_mode == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] -= round_to(p.y_dtype, v[j]);
              }
              head_store4(p.y2, p.y2_dtype, pix * p.y2_cstride + cq * 4, v);
            }
          }
        }
      }
    }
  }
}

static bool head_aligned(const void* ptr, int dtype, int cstride, int coffset) {
  const size_t es = dtype == SRB_F32 ? 4 : dtype == SRB_F8E5M2 ? 1 : 2;
  return ((reinterpret_cast<uintptr_t>(ptr) + (size_t)coffset * es) % (4 * es) == 0) && (cstride % 4 == 0);
}

bool conv_head_eligible(const ConvParams& p) {
  if (p.kh != 3 || p.kw != 3 || p.cin != 3 || p.x_dtype != SRB_F32) return false;
  if (p.cout % 4 || p.cout > 128 || (32 % (p.cout / 4)) != 0) return false;
  if (p.d2s != 1 || p.res1 || p.res2 || p.y_dtype == SRB_F8E5M2 || p.y_dtype == SRB_U8) return false;
  if (!head_aligned(p.y, p.y_dtype, p.y_cstride, p.y_coffset)) return false;
  if (p.y2 && !head_aligned(p.y2, p.y2_dtype, p.y2_cstride, 0)) return false;
  return true;
}

int conv_head_launch(const ConvParams& p, cudaStream_t stream) {
  const int tiles_x = (p.W + kHT_W - 1) / kHT_W, tiles_y = (p.H + kHT_H - 1) / kHT_H;
  const long total = (long)p.B * tiles_x * tiles_y;
  const size_t smem = ((size_t)2 * kHaloH * kHaloW * 3 + 27 * (size_t)p.cout) * sizeof(float);
  const long cap = (long)sm_count() * 8;
  const int grid = (int)(total < cap ? total : cap);
  conv_head_kernel<<<grid, 256, smem, stream>>>(p, tiles_x, tiles_y);
  return launch_check("conv_head_kernel");
}

}  // namespace srb
