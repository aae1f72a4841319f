// CUDA-core convolution engine (Keras Conv2D padding="same", stride 1) - the general path.
//
// Covers every shape the tcgen05 engine does not (Cin = 3 head layers, odd channel counts, fp32
// mode with <= 1e-3 parity) with exact fp32 FMA accumulation (packed FFMA2).  Register-tiled implicit GEMM: a thread owns
// 8 pixels along x times 4 output channels; a block of 256 threads owns, depending on the layer's output count,
//   kCG = 8: 16 x 16 pixels x 32 output channels (the general case),
//   kCG = 2: 32 x 32 pixels x  8 output channels (growth convs of ESRGAN's dense blocks),
//   kCG = 1: 32 x 64 pixels x  4 output channels (RGB tails such as SRCNN's 5x5x32 -> 3: every thread a pixel group
//            instead of 7/8 of them multiplying zero filters).
// The input halo is staged per slab of kDCK channels (8; 3 for RGB inputs; 32 for 1x1 layers) as [c][y][x] in shared memory
// - with one 16-byte load per four channels when the input is float32 NHWC - and the filter rows (all of them at once when
// they fit 64 KB, else one row per pass) alongside it.  For 1 / 3 / 5 / 9-wide filters a thread reads the 8 + kw - 1 inputs of
// a halo row into registers once and slides over the taps.  The epilogue is the same fused bias / activation / scaled-
// residual / clip / depth_to_space as the tensor-core engine; float32 outputs leave as 16-byte stores (four channels of a
// pixel, or eight RGB pixels as six stores), everything else through the element-wise form.
#include "common.cuh"
#include "conv_common.cuh"

namespace srb {

// the element-wise epilogue, out of line: it is the fallback of the 16-byte stores, and inlining its 32 copies made the 9x9
// kernel 16.8k instructions long (10 % of its stall samples were instruction fetches)
static __device__ __noinline__ void epilogue_store_elem(const ConvParams& p, int b, int y, int x, int co, float acc) {
  epilogue_store(p, b, y, x, co, acc);
}

template <int kCG> struct DirectGeom {                 // kCG = output-channel groups of four per block (see the file header)
  static constexpr int kDN = 4 * kCG;                  // output channels per block
  static constexpr int kTCG = kCG == 8 ? 2 : kCG == 2 ? 4 : 8;   // eight-pixel groups per tile row
  static constexpr int kTR = 256 / kCG / kTCG;         // tile rows
  static constexpr int kTW = 8 * kTCG;                 // tile columns
};

// kDCK = input channels per slab: 8, or 3 for RGB inputs (a 9x9x3 head spent 5/8 of its FMAs on zero channels).
// kKW = filter width when it is 1 / 3 / 5 / 9 (0: any): a thread then reads the 8 + kKW - 1 inputs of a halo row once into
// registers and slides over the horizontal taps, instead of eight shared-memory loads per tap (the loop was LSU-bound).
template <int kDCK, int kCG, int kKW>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const ConvParams p, const int all_rows /* every filter row of a slab staged at once */) {
  constexpr int kDN = DirectGeom<kCG>::kDN, kTCG = DirectGeom<kCG>::kTCG, kTR = DirectGeom<kCG>::kTR, kTW = DirectGeom<kCG>::kTW;
  extern __shared__ float smem[];
  const int HH = kTR + p.kh - 1, HW = kTW + p.kw - 1;
  const int HWp = HW | 1;                               // odd row pitch
  float* halo = smem;                                   // [kDCK][HH][HWp]
  float* wsm_base = smem + ((kDCK * HH * HWp + 3) & ~3);   // [kh or 1][kw][kDCK][kDN], 16-byte aligned (the halo of a 3-channel slab need not be)

  const int tid = threadIdx.x;
  const int cg = tid % kCG, pg = tid / kCG;
  const int py = pg / kTCG, x0 = (pg % kTCG) * 8;
  const int n_chunks = (p.cout + kDN - 1) / kDN;
  const int b = blockIdx.z / n_chunks, cc = blockIdx.z % n_chunks;
  const int ty0 = blockIdx.y * kTR, tx0 = blockIdx.x * kTW;
  const int ph = p.kh / 2, pw = p.kw / 2;

  // packed fp32x2 accumulators (FFMA2, sm_100): output-channel pairs (0,1) and (2,3) of each of the 8 pixels
  float2 acc01[8], acc23[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc01[i] = acc23[i] = make_float2(0.f, 0.f);

  const size_t x_img = (size_t)b * p.H * p.W;
  const bool vec_in = p.x_dtype == SRB_F32 && ((p.x_cstride | p.x_coffset | p.cin) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
  for (int c0 = 0; c0 < p.cin; c0 += kDCK) {
    __syncthreads();
    if (kDCK % 4 == 0 && vec_in) {
      // float32 NHWC inputs with 16-byte channel groups: a thread keeps four channels and walks the halo positions - one
      // 16-byte load and four plane stores per position (a quarter of the loads, bounds checks and address arithmetic)
      constexpr int kG = (kDCK % 4 == 0 ? kDCK : 4) / 4, kStepV = 256 / kG;
      const int g4 = (tid % kG) * 4;
      int hp = tid / kG;
      int hy = hp / HW, hx = hp - hy * HW;
      const bool g_ok = c0 + g4 < p.cin;              // (cin % 4 == 0: a group is valid as a whole)
      const float* xf = reinterpret_cast<const float*>(p.x);
      for (; hp < HH * HW; hp += kStepV) {
        const int gy = ty0 + hy - ph, gx = tx0 + hx - pw;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g_ok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
          v = __ldg(reinterpret_cast<const float4*>(xf + (x_img + (size_t)gy * p.W + gx) * p.x_cstride + p.x_coffset + c0 + g4));
        float* hd = halo + (g4 * HH + hy) * HWp + hx;
        hd[0] = v.x; hd[HH * HWp] = v.y; hd[2 * HH * HWp] = v.z; hd[3 * HH * HWp] = v.w;
        hx += kStepV;
        while (hx >= HW) { hx -= HW; ++hy; }
      }
    } else if (256 % kDCK == 0) {
      // a thread keeps its channel and walks the halo positions in steps of 256 / kDCK: no division per element
      constexpr int kStep = 256 / (256 % kDCK == 0 ? kDCK : 1);
      const int c = tid % kDCK;
      int hp = tid / kDCK;
      int hy = hp / HW, hx = hp - hy * HW;
      const bool c_ok = c0 + c < p.cin;
      for (; hp < HH * HW; hp += kStep) {
        const int gy = ty0 + hy - ph, gx = tx0 + hx - pw;
        float v = 0.f;
        if (c_ok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
          v = load_elem(p.x, p.x_dtype, (x_img + (size_t)gy * p.W + gx) * p.x_cstride + p.x_coffset + c0 + c);
        halo[(c * HH + hy) * HWp + hx] = v;
        hx += kStep;
        while (hx >= HW) { hx -= HW; ++hy; }
      }
    } else {
      for (int idx = tid; idx < kDCK * HH * HW; idx += 256) {
        const int c = idx % kDCK;
        const int hp = idx / kDCK;
        const int hx = hp % HW, hy = hp / HW;
        const int gy = ty0 + hy - ph, gx = tx0 + hx - pw;
        float v = 0.f;
        if (c0 + c < p.cin && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
          v = load_elem(p.x, p.x_dtype, (x_img + (size_t)gy * p.W + gx) * p.x_cstride + p.x_coffset + c0 + c);
        halo[(c * HH + hy) * HWp + hx] = v;
      }
    }
    const int row_words = p.kw * kDCK * kDN;
    if (all_rows) {                       // small filters: one staging pass and one barrier per slab instead of one per row
      for (int idx = tid; idx < p.kh * row_words; idx += 256) {
        const int co = idx % kDN;
        const int c = (idx / kDN) % kDCK;
        const int tap = idx / (kDN * kDCK);          // dy * kw + dx
        float v = 0.f;
        const int gco = cc * kDN + co;
        if (c0 + c < p.cin && gco < p.w_cout_pad)
          v = __ldg(p.w_hwio + ((size_t)tap * p.cin + c0 + c) * p.w_cout_pad + gco);
        wsm_base[idx] = v;
      }
      __syncthreads();
    }
    for (int dy = 0; dy < p.kh; ++dy) {
      const float* wsm = all_rows ? wsm_base + dy * row_words : wsm_base;
      if (!all_rows) {
        __syncthreads();
        for (int idx = tid; idx < row_words; idx += 256) {
          const int co = idx % kDN;
          const int c = (idx / kDN) % kDCK;
          const int dx = idx / (kDN * kDCK);
          float v = 0.f;
          const int gco = cc * kDN + co;
          if (c0 + c < p.cin && gco < p.w_cout_pad)
            v = __ldg(p.w_hwio + ((size_t)(dy * p.kw + dx) * p.cin + c0 + c) * p.w_cout_pad + gco);
          wsm_base[idx] = v;
        }
        __syncthreads();
      }
      if (kKW > 0) {
#pragma unroll
        for (int c = 0; c < kDCK; ++c) {
          const float* hrow = halo + (c * HH + py + dy) * HWp + x0;
          float win[8 + (kKW > 0 ? kKW : 1) - 1];
#pragma unroll
          for (int i = 0; i < 8 + kKW - 1; ++i) win[i] = hrow[i];
#pragma unroll
          for (int dx = 0; dx < kKW; ++dx) {
            const float4 w4 = *reinterpret_cast<const float4*>(wsm + (dx * kDCK + c) * kDN + cg * 4);
            const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 aa = make_float2(win[i + dx], win[i + dx]);
              acc01[i] = __ffma2_rn(aa, w01, acc01[i]);
              acc23[i] = __ffma2_rn(aa, w23, acc23[i]);
            }
          }
        }
      } else {
        for (int dx = 0; dx < p.kw; ++dx) {
#pragma unroll
          for (int c = 0; c < kDCK; ++c) {
            const float4 w4 = *reinterpret_cast<const float4*>(wsm + (dx * kDCK + c) * kDN + cg * 4);
            const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
            const float* hrow = halo + (c * HH + py + dy) * HWp + x0 + dx;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = hrow[i];
              const float2 aa = make_float2(a, a);
              acc01[i] = __ffma2_rn(aa, w01, acc01[i]);
              acc23[i] = __ffma2_rn(aa, w23, acc23[i]);
            }
          }
        }
      }
    }
  }

  const int oy = ty0 + py;
  if (oy >= p.H) return;
  // float32 outputs without depth_to_space / second output: the thread's four channels of a pixel leave as one 16-byte store
  const int co0 = cc * kDN + cg * 4;
  // (with depth_to_space the four channels stay together when the post-shuffle channel count is a multiple of four)
  const bool vec4 = (p.d2s == 1 || (p.c_post & 3) == 0) && !p.y2 && p.y_dtype == SRB_F32 && ((p.y_cstride | p.y_coffset) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && co0 + 3 < p.cout;
  // RGB float32 outputs (cout = 3, packed NHWC): the thread's eight pixels are 24 consecutive floats = six 16-byte stores
  if (kCG == 1 && p.cout == 3 && p.d2s == 1 && !p.y2 && p.y_dtype == SRB_F32 && p.y_cstride == 3 && p.y_coffset == 0 &&
      (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && tx0 + x0 + 8 <= p.W) {
    const size_t pix0 = ((size_t)b * p.H + oy) * p.W + tx0 + x0;
    float v[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[3 * i + 0] = epilogue_value(p, acc01[i].x, 0, 0, pix0 + i);
      v[3 * i + 1] = epilogue_value(p, acc01[i].y, 1, 1, pix0 + i);
      v[3 * i + 2] = epilogue_value(p, acc23[i].x, 2, 2, pix0 + i);
    }
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + pix0 * 3);
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = tx0 + x0 + i;
    if (ox >= p.W) continue;
    if (vec4) {
      size_t out_pix; int c_out;
      d2s_map(p, b, oy, ox, co0, out_pix, c_out);
      float4 v;
      v.x = epilogue_value(p, acc01[i].x, co0, c_out, out_pix);
      v.y = epilogue_value(p, acc01[i].y, co0 + 1, c_out + 1, out_pix);
      v.z = epilogue_value(p, acc23[i].x, co0 + 2, c_out + 2, out_pix);
      v.w = epilogue_value(p, acc23[i].y, co0 + 3, c_out + 3, out_pix);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + out_pix * p.y_cstride + p.y_coffset + c_out) = v;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = cc * kDN + cg * 4 + j;
      if (co < p.cout) epilogue_store_elem(p, b, oy, ox, co, j == 0 ? acc01[i].x : j == 1 ? acc01[i].y : j == 2 ? acc23[i].x : acc23[i].y);
    }
  }
}

template <int kDCK, int kCG, int kKW>
static int conv_direct_launch_t(const ConvParams& p, cudaStream_t stream) {
  constexpr int kDN = DirectGeom<kCG>::kDN, kTR = DirectGeom<kCG>::kTR, kTW = DirectGeom<kCG>::kTW;
  const int HH = kTR + p.kh - 1, HW = kTW + p.kw - 1, HWp = HW | 1;
  const size_t halo_words = ((size_t)kDCK * HH * HWp + 3) & ~(size_t)3, row_words = (size_t)p.kw * kDCK * kDN;
  const int all_rows = (halo_words + p.kh * row_words) * sizeof(float) <= 64 * 1024;
  const size_t smem = (halo_words + (all_rows ? p.kh : 1) * row_words) * sizeof(float);
  SRB_REQUIRE(smem <= 200 * 1024, "conv(direct): kernel %dx%d too large for the shared-memory halo", p.kh, p.kw);
  static size_t configured = 0;
  if (smem > configured) {
    SRB_CUDA(cudaFuncSetAttribute(conv_direct_kernel<kDCK, kCG, kKW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int n_chunks = (p.cout + kDN - 1) / kDN;
  dim3 grid((p.W + kTW - 1) / kTW, (p.H + kTR - 1) / kTR, p.B * n_chunks);
  SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv(direct): grid too large");
  conv_direct_kernel<kDCK, kCG, kKW><<<grid, 256, smem, stream>>>(p, all_rows);
  return launch_check("conv_direct_kernel");
}

// Instantiated (slab, channel groups, filter width) combinations - the ones the reference's networks reach in the float32 mode;
// every other filter width runs the any-width form of its geometry.
int conv_direct_launch(const ConvParams& p, cudaStream_t stream) {
  if (conv_head_eligible(p)) return conv_head_launch(p, stream);   // RGB 3x3 head layers: coalesced-store kernel
  if (p.cout <= 4 && p.kh * p.kw <= 81 && p.W >= 32) {             // few outputs: every thread a pixel group (32 x 64 tiles)
    if (p.cin == 3) return conv_direct_launch_t<3, 1, 0>(p, stream);
    if (p.kw == 3) return conv_direct_launch_t<8, 1, 3>(p, stream);           // EDSR / ESRGAN RGB tails
    if (p.kw == 5) return conv_direct_launch_t<8, 1, 5>(p, stream);           // SRCNN 5x5x32 -> 3
    if (p.kw == 9) return conv_direct_launch_t<8, 1, 9>(p, stream);           // SRResNet 9x9x64 -> 3 tail
    return conv_direct_launch_t<8, 1, 0>(p, stream);
  }
  if (p.cout <= 8 && p.cin != 3 && p.kh * p.kw <= 25) {            // up to eight outputs: 32 x 32 tiles, two channel groups
    if (p.kw == 3) return conv_direct_launch_t<8, 2, 3>(p, stream);           // ESRGAN growth convs
    return conv_direct_launch_t<8, 2, 0>(p, stream);
  }
  if (p.cin == 3) {
    if (p.kw == 9) return conv_direct_launch_t<3, 8, 9>(p, stream);           // SRCNN / SRResNet 9x9 heads
    if (p.kw == 5) return conv_direct_launch_t<3, 8, 5>(p, stream);           // ESPCN 5x5 head
    return conv_direct_launch_t<3, 8, 0>(p, stream);
  }
  if (p.kh == 1 && p.kw == 1 && p.cin >= 32)                       // 1x1 layers: 32-channel slabs (a slab of 8 is three barriers
    return conv_direct_launch_t<32, 8, 1>(p, stream);              // and a staging pass per 128 packed FMAs of a thread)
  if (p.kw == 3) return conv_direct_launch_t<8, 8, 3>(p, stream);
  return conv_direct_launch_t<8, 8, 0>(p, stream);
}

}  // namespace srb
