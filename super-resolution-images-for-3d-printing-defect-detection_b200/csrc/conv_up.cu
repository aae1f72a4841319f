// EDSR's up-sampling tail as ONE 5 x 5 convolution on the tensor cores (EDSR_model.py:76-95, 117-123).
//
// Conv2D(64 -> 256) -> depth_to_space(2) [-> Conv2D(64 -> 256) -> depth_to_space(2)] -> Conv2D(64 -> C) has no activation
// in between, so it is a linear map from the 64-channel low-resolution feature image to the r x r x C block of every
// low-resolution pixel with a 5 x 5 footprint (srb200/compose.py builds the exact map in float64, nine border variants
// included).  At x4 that is 153,600 FLOP per low-resolution pixel instead of 1,529,856 and neither full-resolution
// 64-channel intermediate is written or read.
//
// Kernel = the wide-tile, dx-folded formulation of conv_fold.cu with five horizontal taps in N:
//     D[input pixel p, (dx, co)] = sum over dy, ci of X[p_y + dy - 2, p_x, ci] * W[dy, dx, ci, co]     (N = 5 * gw, 20 MMAs, M = 128)
//     out[y, x, co] = sum over dx of D[(y, x + dx - 2), (dx, co)]                                       (4 shuffles per channel)
// tile = 4 image rows x 32 input columns (28 outputs per row), halo = one TMA box 64 c x 32 w x 8 h, weights resident
// ([5][N][64] 16-bit, 150 KB at x4), double-buffered TMEM accumulators, 16 epilogue warps = (row, sub-row i of the r x r
// block); each lane stores the r * C contiguous values of its pixel's sub-row i straight into the image.
//
// Border variants: the persistent CTAs are split over nine SEGMENTS - interior, top / bottom row, left / right column, four
// corners - each with its own resident weights, output rectangle and tile list, all in one launch.  The column strips read
// the image through a TRANSPOSED tensor map (x and y strides swapped) with transposed weights, so that a column is a tile
// row like any other.
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct srb_upsampler {
  int scale, c_img, cin, cout;   // cout = scale^2 * c_img
  int gw, n, cw;                 // accumulator columns per dx group (cout rounded up to 16), N = 5 * gw, cw = scale * c_img
  float inv_wscale;
  void* w_bf16;                  // [9 segments][5 dy][N = dx * gw + co][64] K-major
  void* w_f16;
  float* bias;                   // [9][64]
};

namespace srb {

constexpr int kUW = 32, kUH = 4, kUK = 5, kUOut = kUW - (kUK - 1);
constexpr int kUpEpiWarps = 16;
constexpr int kUpThreads = 64 + 32 * kUpEpiWarps;
constexpr int kUpSegs = 9;
constexpr uint32_t kUpStage = (uint32_t)(kUH + kUK - 1) * kUW * 128u;     // 8 halo rows x 32 columns x 64 channels x 2 B

struct UpSeg {
  int cta0, ncta;                  // CTAs [cta0, cta0 + ncta) work on this segment
  int ry0, ry1, rx0, rx1;          // output rectangle in the segment's orientation (rows x columns)
  int tiles_x, tiles_per_img, total_tiles;
  uint32_t magic_tpi, magic_tx;
  int transposed;                  // rows = image x, columns = image y
};

struct UpParams {
  UpSeg seg[kUpSegs];
  int n, gw, cw, r, c_img;
  int stages;
  uint32_t tmem_cols, idesc;
  int B, H, W;
  void* y; int y_dtype; int clip01;
  float inv_wscale;
  const float* bias;               // [9][64]
  int reverse;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}

// kCW = 12: x4 RGB (each warp = one sub-row of 12 contiguous values, vector stores); kCW = 16: any other r / channel count
// (16 accumulator columns loaded per tap group, the first q.cw used, element-wise stores)
template <int kCW>
__global__ void __launch_bounds__(kUpThreads, 1)
upsample5_fold_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_xt,
                      const __grid_constant__ CUtensorMap tmap_w, const UpParams q) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t w_bytes = (uint32_t)kUK * (uint32_t)q.n * 128u;
  const uint32_t w_span = (w_bytes + 1023u) & ~1023u;
  const uint32_t w_smem = base, a_smem = base + w_span;
  uint8_t* tail = smem + w_span + (size_t)q.stages * kUpStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(kMaxStages + s); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 1 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 3 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);     // [64]

  // which segment this CTA serves
  int si = 0;
  UpSeg sg = q.seg[0];
#pragma unroll
  for (int i = 1; i < kUpSegs; ++i)
    if ((int)blockIdx.x >= q.seg[i].cta0 && (int)blockIdx.x < q.seg[i].cta0 + q.seg[i].ncta) { si = i; sg = q.seg[i]; }
  const int first_tile = (int)blockIdx.x - sg.cta0, tile_step = sg.ncta;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < q.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * q.r); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64; i += kUpThreads) bias_s[i] = q.bias[si * 64 + i];
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_xt); prefetch_tmap(&tmap_w); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), q.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto coords = [&](int tile, int& b, int& y0, int& x0) {
    const int tl = q.reverse ? sg.total_tiles - 1 - tile : tile;
    b = fast_div(tl, sg.tiles_per_img, sg.magic_tpi);
    const int rr_ = tl - b * sg.tiles_per_img;
    const int ty = fast_div(rr_, sg.tiles_x, sg.magic_tx);
    y0 = sg.ry0 + ty * kUH;
    x0 = sg.rx0 + (rr_ - ty * sg.tiles_x) * kUOut;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, w_bytes);
      for (int dy = 0; dy < kUK; ++dy)
        tma_load_2d(w_smem + (uint32_t)(dy * q.n) * 128u, &tmap_w, wfull_bar, 0, (si * kUK + dy) * q.n);
    }
    __syncwarp();
    const CUtensorMap* mx = sg.transposed ? &tmap_xt : &tmap_x;
    int s = 0; uint32_t ph = 0;
    for (int tile = first_tile; tile < sg.total_tiles; tile += tile_step) {
      int b, y0, x0;
      coords(tile, b, y0, x0);
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), kUpStage);
        tma_load_4d(a_smem + (uint32_t)s * kUpStage, mx, full_bar(s), 0, x0 - (kUK >> 1), y0 - (kUK >> 1), b);
      }
      __syncwarp();
      if (++s == q.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 5 vertical taps x 4 k-steps, N = 5 * gw =====================
    mbar_wait(wfull_bar, 0);
    int s = 0; uint32_t ph = 0; int it = 0;
    const uint64_t b_desc0 = make_desc(w_smem, 1024u);
    const uint32_t b_dy = ((uint32_t)q.n * 128u) >> 4;
    for (int tile = first_tile; tile < sg.total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint64_t a_desc0 = make_desc(a_smem + (uint32_t)s * kUpStage, 1024u);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * q.n);
      if (elect_one()) {
#pragma unroll
        for (int dy = 0; dy < kUK; ++dy) {
          const uint64_t ad = a_desc0 + (uint64_t)(dy * ((kUW * 128) >> 4)), bd = b_desc0 + (uint64_t)dy * b_dy;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2u * k, bd + 2u * k, q.idesc, (uint32_t)((dy | k) != 0));
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++s == q.stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: warp = (tile row = TMEM lane quadrant, sub-row i of the r x r block) =====================
    const int ew = warp - 2, quad = warp & 3, cq = ew >> 2;
    if (cq < q.r) {
      const int col0 = cq * q.cw;                        // first output channel of this warp: (i * r + 0) * C + 0
      const int r = q.r, c_img = q.c_img;
      const float inv = q.inv_wscale;
      float bb[kCW];
#pragma unroll
      for (int e = 0; e < kCW; ++e) bb[e] = (e < q.cw) ? bias_s[col0 + e] : 0.f;
      const size_t row_elems = (size_t)q.W * r * c_img;  // elements per output image row
      int it = 0;
      for (int tile = first_tile; tile < sg.total_tiles; tile += tile_step, ++it) {
        int b, y0, x0;
        coords(tile, b, y0, x0);
        const int acc = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * q.n + col0);
        uint32_t d[kUK][kCW];
        __syncwarp();
#pragma unroll
        for (int dx = 0; dx < kUK; ++dx) {
          const uint32_t t = t_row + (uint32_t)(dx * q.gw);
          if (kCW == 12) { tmem_ld8(t, &d[dx][0]); tmem_ld4(t + 8u, &d[dx][8]); }
          else tmem_ld16(t, *reinterpret_cast<uint32_t(*)[16]>(&d[dx][0]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        float v[kCW];
#pragma unroll
        for (int e = 0; e < kCW; ++e) {
          float a = __uint_as_float(d[0][e]);
#pragma unroll
          for (int dx = 1; dx < kUK; ++dx) a += __shfl_down_sync(0xffffffffu, __uint_as_float(d[dx][e]), dx);
          a = fmaf(a, inv, bb[e]);
          if (q.clip01) a = fminf(fmaxf(a, 0.f), 1.f);
          v[e] = a;
        }
        const int ty = y0 + quad, tx = x0 + lane;
        if (lane < kUOut && ty < sg.ry1 && tx < sg.rx1) {
          const int iy = sg.transposed ? tx : ty, ix = sg.transposed ? ty : tx;
          const size_t o = ((size_t)b * q.H * r + (size_t)iy * r + cq) * row_elems + (size_t)ix * r * c_img;
          if (kCW == 12) {
            if (q.y_dtype == SRB_F32) {
              float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(q.y) + o);
              dp[0] = make_float4(v[0], v[1], v[2], v[3]);
              dp[1] = make_float4(v[4], v[5], v[6], v[7]);
              dp[2] = make_float4(v[8], v[9], v[10], v[11]);
            } else if (q.y_dtype == SRB_U8) {
              uint32_t pk[3];
#pragma unroll
              for (int g = 0; g < 3; ++g) {
                uint32_t w4 = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  w4 |= (uint32_t)__float2int_rn(fminf(fmaxf(v[4 * g + u], 0.f), 1.f) * 255.f) << (8 * u);
                pk[g] = w4;
              }
              uint32_t* dp = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(q.y) + o);
              dp[0] = pk[0]; dp[1] = pk[1]; dp[2] = pk[2];
            } else {
              uint2* dp = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(q.y) + o);
#pragma unroll
              for (int g = 0; g < 3; ++g)
                dp[g] = make_uint2(pack2(v[4 * g], v[4 * g + 1], q.y_dtype), pack2(v[4 * g + 2], v[4 * g + 3], q.y_dtype));
            }
          } else {
#pragma unroll
            for (int e = 0; e < kCW; ++e)
              if (e < q.cw) store_elem(q.y, q.y_dtype, o + e, v[e]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, q.tmem_cols);
  }
}

}  // namespace srb

using namespace srb;

// segment order of the kernel: interior, top row, bottom row, left column, right column, then the four corners;
// {vy, vx} = border class of the rows / columns (0 first, 1 interior, 2 last); the column strips are stored transposed
static const int kSegVariant[kUpSegs][3] = {{1, 1, 0}, {0, 1, 0}, {2, 1, 0}, {1, 0, 1}, {1, 2, 1}, {0, 0, 0}, {0, 2, 0}, {2, 0, 0}, {2, 2, 0}};

extern "C" int srb_upsampler_create(const float* w, const float* bias, int cin, int c_img, int scale, float w_scale,
                                    srb_upsampler** out) {
  SRB_REQUIRE(w && bias && out, "upsampler_create: null pointer");
  SRB_REQUIRE(cin == 64, "upsampler_create: the tensor-core kernel needs 64 input channels (got %d)", cin);
  SRB_REQUIRE(scale >= 2 && scale <= 4 && c_img >= 1 && scale * scale * c_img <= 48 && scale * c_img <= 16,
              "upsampler_create: scale %d with %d image channels is not supported (scale^2 * channels <= 48)", scale, c_img);
  SRB_REQUIRE(w_scale > 0.f, "upsampler_create: w_scale must be positive");
  srb_upsampler* u = (srb_upsampler*)calloc(1, sizeof(srb_upsampler));
  if (!u) { set_error("upsampler_create: out of host memory"); return SRB_E_NOMEM; }
  u->scale = scale; u->c_img = c_img; u->cin = cin; u->cout = scale * scale * c_img;
  u->gw = (u->cout + 15) & ~15; u->n = kUK * u->gw; u->cw = scale * c_img;
  u->inv_wscale = 1.f / w_scale;
  const int cout = u->cout, n = u->n, gw = u->gw;
  std::vector<__nv_bfloat16> tb((size_t)kUpSegs * kUK * n * 64, __float2bfloat16_rn(0.f));
  std::vector<__half> th((size_t)kUpSegs * kUK * n * 64, __float2half_rn(0.f));
  std::vector<float> hb((size_t)kUpSegs * 64, 0.f);
  for (int s = 0; s < kUpSegs; ++s) {
    const int vy = kSegVariant[s][0], vx = kSegVariant[s][1], tr = kSegVariant[s][2];
    const float* wv = w + (size_t)(vy * 3 + vx) * kUK * kUK * cin * cout;
    for (int dy = 0; dy < kUK; ++dy)
      for (int dx = 0; dx < kUK; ++dx) {
        const int qy = tr ? dx : dy, qx = tr ? dy : dx;        // transposed segments: the kernel's rows run along image x
        for (int c = 0; c < cin; ++c)
          for (int o = 0; o < cout; ++o) {
            const float v = wv[((size_t)(qy * kUK + qx) * cin + c) * cout + o] * w_scale;
            const size_t idx = (((size_t)s * kUK + dy) * n + dx * gw + o) * 64 + c;
            tb[idx] = __float2bfloat16_rn(v);
            th[idx] = __float2half_rn(v);
          }
      }
    for (int o = 0; o < cout; ++o) hb[(size_t)s * 64 + o] = bias[(size_t)(vy * 3 + vx) * cout + o];
  }
  cudaError_t e;
  if ((e = cudaMalloc(&u->w_bf16, tb.size() * 2)) != cudaSuccess ||
      (e = cudaMemcpy(u->w_bf16, tb.data(), tb.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&u->w_f16, th.size() * 2)) != cudaSuccess ||
      (e = cudaMemcpy(u->w_f16, th.data(), th.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&u->bias, hb.size() * sizeof(float))) != cudaSuccess ||
      (e = cudaMemcpy(u->bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
    srb_upsampler_destroy(u);
    return cuda_fail(e, "upsampler_create");
  }
  *out = u;
  return SRB_OK;
}

extern "C" void srb_upsampler_destroy(srb_upsampler* u) {
  if (!u) return;
  if (u->w_bf16) cudaFree(u->w_bf16);
  if (u->w_f16) cudaFree(u->w_f16);
  if (u->bias) cudaFree(u->bias);
  free(u);
}

extern "C" int srb_upsample_composed(const srb_upsampler* u, const void* x, int x_dtype, int x_cstride, int x_coffset,
                                     int batch, int height, int width, void* y, int y_dtype, int clip01, srb_stream_t stream) {
  SRB_REQUIRE(u, "upsample_composed: null up-sampler");
  if (batch == 0) return SRB_OK;
  SRB_REQUIRE(x && y, "upsample_composed: null pointer");
  SRB_REQUIRE(x_dtype == SRB_F16 || x_dtype == SRB_BF16, "upsample_composed: x must be fp16 or bf16 (tensor-core operand)");
  SRB_REQUIRE(y_dtype == SRB_F32 || y_dtype == SRB_F16 || y_dtype == SRB_BF16 || y_dtype == SRB_U8,
              "upsample_composed: y dtype must be f32, f16, bf16 or u8");
  SRB_REQUIRE(batch > 0 && height >= 2 && width >= 2, "upsample_composed: images must be at least 2 x 2 (got %d x %d x %d)", batch, height, width);
  if (x_cstride <= 0) x_cstride = u->cin;
  SRB_REQUIRE(x_coffset >= 0 && x_coffset + u->cin <= x_cstride && x_cstride % 8 == 0 && x_coffset % 8 == 0 && aligned16(x),
              "upsample_composed: input channel slice must be 16-byte aligned and inside the pixel");
  SRB_REQUIRE(aligned16(y), "upsample_composed: y must be 16-byte aligned");
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("upsample_composed: cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }
  const int B = batch, H = height, W = width;

  UpParams q{};
  q.n = u->n; q.gw = u->gw; q.cw = u->cw; q.r = u->scale; q.c_img = u->c_img;
  q.B = B; q.H = H; q.W = W;
  q.y = y; q.y_dtype = y_dtype; q.clip01 = clip01; q.inv_wscale = u->inv_wscale; q.bias = u->bias;
  q.reverse = tc_next_reverse();
  const bool fast = u->cw == 12;
  const int kcw = fast ? 12 : 16;
  q.tmem_cols = 32;
  const int last_col = q.n + 4 * q.gw + (q.r - 1) * q.cw + kcw;                  // one past the last column any load touches
  while (q.tmem_cols < (uint32_t)(last_col > 2 * q.n ? last_col : 2 * q.n)) q.tmem_cols <<= 1;
  const uint32_t fmt = x_dtype == SRB_BF16 ? 1u : 0u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(q.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  // segments: {ry0, ry1, rx0, rx1} in the segment's own orientation
  const int rect[kUpSegs][4] = {
      {1, H - 1, 1, W - 1}, {0, 1, 1, W - 1}, {H - 1, H, 1, W - 1},
      {0, 1, 1, H - 1}, {W - 1, W, 1, H - 1},
      {0, 1, 0, 1}, {0, 1, W - 1, W}, {H - 1, H, 0, 1}, {H - 1, H, W - 1, W}};
  long total = 0;
  int nonempty = 0;
  for (int s = 0; s < kUpSegs; ++s) {
    UpSeg& g = q.seg[s];
    g.ry0 = rect[s][0]; g.ry1 = rect[s][1]; g.rx0 = rect[s][2]; g.rx1 = rect[s][3];
    g.transposed = kSegVariant[s][2];
    const int ny = g.ry1 - g.ry0, nx = g.rx1 - g.rx0;
    g.tiles_x = nx > 0 ? (nx + kUOut - 1) / kUOut : 0;
    const int tiles_y = ny > 0 ? (ny + kUH - 1) / kUH : 0;
    g.tiles_per_img = g.tiles_x * tiles_y;
    const long t = (long)B * g.tiles_per_img;
    SRB_REQUIRE(t < (1L << 30), "upsample_composed: too many tiles");
    g.total_tiles = (int)t;
    g.magic_tpi = g.tiles_per_img ? div_magic(t + 1, g.tiles_per_img) : 0;
    g.magic_tx = g.tiles_x ? div_magic((long)g.tiles_per_img, g.tiles_x) : 0;
    total += t;
    nonempty += t > 0;
  }
  // CTAs per segment: one each, then the rest one at a time to whichever segment has the most tiles per CTA
  int grid = sm_count();
  if ((long)grid > total) grid = (int)total;
  if (grid < nonempty) grid = nonempty;
  int used = 0;
  for (int s = 0; s < kUpSegs; ++s) { q.seg[s].ncta = q.seg[s].total_tiles > 0 ? 1 : 0; used += q.seg[s].ncta; }
  for (; used < grid; ++used) {
    int best = 0; double load = -1.0;
    for (int s = 0; s < kUpSegs; ++s)
      if (q.seg[s].ncta > 0) {
        const double l = (double)q.seg[s].total_tiles / q.seg[s].ncta;
        if (l > load) { load = l; best = s; }
      }
    ++q.seg[best].ncta;
  }
  int c0 = 0;
  for (int s = 0; s < kUpSegs; ++s) {
    q.seg[s].cta0 = c0; c0 += q.seg[s].ncta;
    if (q.seg[s].ncta == 0) q.seg[s].cta0 = -1 << 20;       // (never matches a block index)
  }

  int dev = 0, max_smem = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t w_bytes = ((size_t)kUK * q.n * 128 + 1023) & ~(size_t)1023;
  const size_t tail_bytes = (2 * kMaxStages + 6) * 8 + 64 * sizeof(float);
  auto smem_need = [&](int st) { return 1024 + w_bytes + (size_t)st * kUpStage + tail_bytes; };
  q.stages = 4;
  while (q.stages > 2 && smem_need(q.stages) > (size_t)max_smem) --q.stages;
  const size_t smem = smem_need(q.stages);
  if (smem > (size_t)max_smem) { set_error("upsample_composed: staging does not fit shared memory"); return SRB_E_UNSUPPORTED; }

  const CUtensorMapDataType tdt = x_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmx, tmxt, tmw;
  void* gptr = (void*)((const uint16_t*)x + x_coffset);
  const cuuint32_t es4[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {64, (cuuint32_t)kUW, (cuuint32_t)(kUH + kUK - 1), 1};
  {
    const cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)x_cstride * 2, (cuuint64_t)W * x_cstride * 2, (cuuint64_t)H * W * x_cstride * 2};
    CUresult r = encode(&tmx, tdt, 4, gptr, dims, strides, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("upsample_composed: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  {   // the same image with x and y swapped: dimension 1 walks image rows, dimension 2 image columns
    const cuuint64_t dims[4] = {64, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * x_cstride * 2, (cuuint64_t)x_cstride * 2, (cuuint64_t)H * W * x_cstride * 2};
    CUresult r = encode(&tmxt, tdt, 4, gptr, dims, strides, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("upsample_composed: cuTensorMapEncodeTiled(x transposed) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {64, (cuuint64_t)kUpSegs * kUK * q.n};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t wbox[2] = {64, (cuuint32_t)q.n};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmw, tdt, 2, x_dtype == SRB_BF16 ? u->w_bf16 : u->w_f16, dims, strides, wbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("upsample_composed: cuTensorMapEncodeTiled(w) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  typedef void (*UpFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const UpParams);
  const UpFn kernel = fast ? upsample5_fold_kernel<12> : upsample5_fold_kernel<16>;
  static size_t configured[2] = {0, 0};
  if (smem > configured[fast ? 0 : 1]) {
    SRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[fast ? 0 : 1] = smem;
  }
  SRB_CUDA(tc_launch(kernel, grid, kUpThreads, smem, (cudaStream_t)stream, false, tmx, tmxt, tmw, q));
  return launch_check("upsample5_fold_kernel");
}
