// Wide-tile, dx-folded tcgen05 kernel for 3x3 (and odd K x K few-channel) layers with Cin = 64 and few output channels per
// MMA: the res-block first convs and ESPCN's middle layer (Cout = 16 .. 64), the RGB tails (Cout <= 4), ESPCN's last layer
// (depth_to_space straight to the RGB image).  EDSR_model.py:61,121; see DESIGN.md section 4.1.
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace srb {

constexpr int kTileM = 128;               // GEMM rows of one tcgen05.mma (cta_group::1)

__device__ __noinline__ float act_generic_fold(float v, int act, float slope) { return apply_act(v, act, slope); }

// ===================================================================================================================
// Wide-tile, dx-folded kernel for 3x3 layers with few output channels per MMA (Cout = 64 and the RGB tails).
//
// A tcgen05.mma with M = 128 spends ~64 cycles fetching its 128 x 16 A block from shared memory whatever N is (measured:
// N = 16, 64 and 128 all issue at ~60-64 cycles per instruction), so the 36-MMA, N = 64 formulation of a 64 -> 64 layer
// cannot pass 50 % of the tensor peak.  Folding the three horizontal taps into N makes the instruction three times wider
// and the tile three times shorter:
//     D[input pixel p, (dx, co)] = sum over dy, ci of  X[p_y + dy - 1, p_x, ci] * W[dy, dx, ci, co]      (N = 3 * Cout, 12 MMAs)
//     out[y, x, co] = D[(y, x - 1), (0, co)] + D[(y, x), (1, co)] + D[(y, x + 1), (2, co)]
// The second line is a sum over neighbouring GEMM rows, i.e. neighbouring TMEM lanes: a tile is 4 image rows x 32
// consecutive input columns (one row per lane quadrant), lane l of a quadrant holds input column x0 - 1 + l, and the
// epilogue warp combines own / lane + 1 / lane + 2 with two shuffles per channel; lanes 0..29 hold the 30 output columns.
// The halo is one TMA box of 64 c x 32 w x 6 h; the dy shift is 4,096 bytes (swizzle-atom aligned).  Output rows move with
// TMA exactly as in the staged epilogue above (boxes of {32 channels, 30 pixels, 1 row}).
// Modes: 0 plain 16-bit y, 1 ReLU 16-bit y, 4 PReLU / leaky 16-bit y,
//        5 depth_to_space straight to a few-channel fp32 image (ESPCN: 48 -> 4 x 4 x RGB), float4 stores,
//        3 few channels (Cout <= 4, groups of 5 columns): bias / activation / alpha / clip, stored element-wise.
// ===================================================================================================================
constexpr int kFW = 32, kFH = 4, kFOut = 30;
constexpr int kFoldEpiWarps = 16;
constexpr int kFoldThreads = 64 + 32 * kFoldEpiWarps;

struct FoldParams {
  int n;                 // MMA N: 192 (3 x 64), or kw x 4 rounded up to 16 for the few-channel layers (16 / 32 / 48)
  int gw;                // accumulator columns per dx group: 64 or 4
  int kh, kw;            // filter size (3 x 3 for the 64-channel modes; odd, <= 9 for mode 3)
  int out_cols;          // output columns per tile row: 32 - (kw - 1)
  uint32_t stage_bytes;  // (kFH + kh - 1) halo rows x 32 columns x 128 B
  int tiles_x, tiles_y, total_tiles;
  int stages;
  uint32_t tmem_cols, idesc, epi_warp_bytes;
  int reverse;           // as TcParams::reverse
  uint32_t magic_tpi, magic_tx;   // as TcParams
};

template <int kMode, bool kGen>   // kGen: filter size from q.kh / q.kw (mode 3 only); otherwise 3 x 3 with compile-time geometry
__global__ void __launch_bounds__(kFoldThreads, 1)
conv3x3_fold_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ EpiMaps em, const FoldParams q, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int f_kh = kGen ? q.kh : 3, f_kw = kGen ? q.kw : 3;
  const int f_out = kGen ? q.out_cols : kFOut;
  const uint32_t f_stage = kGen ? q.stage_bytes : 6u * kFW * 128u;
  const uint32_t w_bytes = (uint32_t)f_kh * (uint32_t)q.n * 128u;
  const uint32_t w_span = (w_bytes + 1023u) & ~1023u;
  const uint32_t w_smem = base, a_smem = base + w_span;
  const uint32_t epi_smem = a_smem + (uint32_t)q.stages * f_stage;
  uint8_t* tail = smem + w_span + (size_t)q.stages * f_stage + (uint32_t)kFoldEpiWarps * q.epi_warp_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(kMaxStages + s); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 1 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 3 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);     // [64]
  float* slope_s = bias_s + 64;                                              // [64] PReLU / leaky slopes (mode 4)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < q.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kMode == 3 ? 4 : kFoldEpiWarps); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64; i += kFoldThreads) {
    bias_s[i] = i < p.cout ? p.bias[i] : 0.f;
    if (kMode == 4) slope_s[i] = (p.act == SRB_ACT_PRELU && i < p.cout) ? p.prelu[i] : p.act_slope;
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
  if (warp == 2 && lane == 0 && kMode != 3 && kMode != 5) prefetch_tmap(&em.y);
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), q.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = q.tiles_x * q.tiles_y;
  const int first_tile = (int)blockIdx.x, tile_step = (int)gridDim.x;
  auto coords = [&](int tile, int& b, int& y0, int& x0) {
    const int tl = q.reverse ? q.total_tiles - 1 - tile : tile;
    b = fast_div(tl, tiles_per_img, q.magic_tpi);
    const int rr_ = tl - b * tiles_per_img;
    const int ty = fast_div(rr_, q.tiles_x, q.magic_tx);
    y0 = ty * kFH;
    x0 = (rr_ - ty * q.tiles_x) * f_out;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, w_bytes);
      for (int dy = 0; dy < f_kh; ++dy) tma_load_2d(w_smem + (uint32_t)(dy * q.n) * 128u, &tmap_w, wfull_bar, 0, dy * q.n);
    }
    __syncwarp();
    int s = 0; uint32_t ph = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step) {
      int b, y0, x0;
      coords(tile, b, y0, x0);
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), f_stage);
        tma_load_4d(a_smem + (uint32_t)s * f_stage, &tmap_x, full_bar(s), 0, x0 - (f_kw >> 1), y0 - (f_kh >> 1), b);
      }
      __syncwarp();
      if (++s == q.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 3 vertical taps x 4 k-steps, N = 3 * group width =====================
    mbar_wait(wfull_bar, 0);
    int s = 0; uint32_t ph = 0; int it = 0;
    const uint64_t b_desc0 = make_desc(w_smem, 1024u);
    const uint32_t b_dy = ((uint32_t)q.n * 128u) >> 4;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint64_t a_desc0 = make_desc(a_smem + (uint32_t)s * f_stage, 1024u);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * q.n);
      if (!kGen) {                                        // 3 x 3: twelve MMAs, fully unrolled (descriptors in uniform registers)
        if (elect_one()) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t ad = a_desc0 + (uint64_t)(dy * ((kFW * 128) >> 4)), bd = b_desc0 + (uint64_t)dy * b_dy;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2u * k, bd + 2u * k, q.idesc, (uint32_t)((dy | k) != 0));
          }
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(acc));
        }
      } else if (elect_one()) {                          // taller filters (5 x 5, 9 x 9 tails): rolled over the vertical taps
        for (int dy = 0; dy < f_kh; ++dy) {
          const uint64_t ad = a_desc0 + (uint64_t)(dy * ((kFW * 128) >> 4)), bd = b_desc0 + (uint64_t)dy * b_dy;
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
    // ===================== epilogue =====================
    const int ew = warp - 2, quad = warp & 3, cq = ew >> 2;   // cq: warp set (mode 3) / 16-channel quarter (modes 0-2)
    auto release_tmem = [&](int acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    };
    if (kMode == 3) {
      // few channels: the two warp sets take alternate tiles; element-wise stores (lanes 0..29 = output columns x0 + lane)
      int it = 0;
      for (int tile = first_tile; tile < q.total_tiles && cq < 2; tile += tile_step, ++it) {
        const int acc = it & 1;
        if (acc != cq) continue;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        int b, y0, x0;
        coords(tile, b, y0, x0);
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        float v[4];
        const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * q.n);
        if (!kGen) {                                      // the hot case (RGB tail of EDSR): one TMEM load, two shuffles per channel
          uint32_t rr[16];                                // columns dx * 4 + co
          __syncwarp();
          tmem_ld16(t_acc, rr);
          tmem_ld_wait();
          release_tmem(acc);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float mid = __shfl_down_sync(0xffffffffu, __uint_as_float(rr[4 + e]), 1);
            const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(rr[8 + e]), 2);
            v[e] = (bias_s[e] + __uint_as_float(rr[e])) + (mid + right);
          }
        } else {
          uint32_t rr[48];
          __syncwarp();
          tmem_ld16(t_acc, *reinterpret_cast<uint32_t(*)[16]>(&rr[0]));
          if (q.n > 16) tmem_ld16(t_acc + 16u, *reinterpret_cast<uint32_t(*)[16]>(&rr[16]));
          if (q.n > 32) tmem_ld16(t_acc + 32u, *reinterpret_cast<uint32_t(*)[16]>(&rr[32]));
          tmem_ld_wait();
          release_tmem(acc);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = bias_s[e] + __uint_as_float(rr[e]);
#pragma unroll
          for (int dx = 1; dx < 9; ++dx) {
            if (dx < f_kw) {                              // (uniform) tap dx of output column l sits in lane l + dx
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] += __shfl_down_sync(0xffffffffu, __uint_as_float(rr[dx * 4 + e]), dx);
            }
          }
        }
        const int oy = y0 + quad, ox = x0 + lane;
        if (lane < f_out && oy < p.H && ox < p.W) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.act == SRB_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
            else if (p.act != SRB_ACT_NONE)
              v[e] = act_generic_fold(v[e], p.act, (p.act == SRB_ACT_PRELU && e < p.cout) ? __ldg(p.prelu + e) : p.act_slope);
            v[e] *= p.alpha;
            if (p.clip01) v[e] = fminf(fmaxf(v[e], 0.f), 1.f);
          }
          const size_t o = (((size_t)b * p.H + oy) * p.W + ox) * p.y_cstride + p.y_coffset;
          if (p.y_dtype == SRB_F32) {
            float* d = reinterpret_cast<float*>(p.y) + o;
#pragma unroll
            for (int e = 0; e < 4; ++e) if (e < p.cout) d[e] = v[e];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (e < p.cout) store_elem(p.y, p.y_dtype, o + e, v[e]);
          }
        }
      }
    } else {
      // 16 .. 64 output channels, 16-bit y: warp (quadrant = tile row, cq = 16 of the channels); rows move with TMA.
      // Sixteen epilogue warps (four per scheduler) hide the TMEM-load / shuffle latencies of this role.
      const int col0 = cq * 16;                           // first output channel of this warp
      constexpr uint32_t hb = 32u;                        // staged row bytes: 16 x 16-bit
      const uint32_t my_epi = epi_smem + (uint32_t)ew * q.epi_warp_bytes;
      const uint32_t row = (uint32_t)lane;
      const uint32_t h_row = row * hb, h_x = (row >> 2) & 1u;            // 32-byte swizzle
      const float alpha = p.alpha;
      const bool bf = p.y_dtype == SRB_BF16;
      float bb[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) bb[e] = bias_s[col0 + e];
      uint32_t d2s_off[4] = {};                           // mode 5: float offset of each 4-channel group inside the r x r block
      const bool d2s_plain = p.act == SRB_ACT_NONE && p.alpha == 1.f && !p.clip01;
      if (kMode == 5) {
        const int rg = p.d2s * p.c_post;                  // floats per output sub-row (a multiple of 4)
        const uint32_t sub_row = (uint32_t)(p.W * p.d2s) * (uint32_t)p.c_post;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cb = col0 + 4 * g, i = cb / rg;
          d2s_off[g] = (uint32_t)i * sub_row + (uint32_t)(cb - i * rg);
        }
      }
      __syncwarp();
      int it = 0;
      for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
        int b, y0, x0;
        coords(tile, b, y0, x0);
        const int acc = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        const uint32_t buf = my_epi;
        if (col0 >= p.cout) {                             // layers with fewer than 64 outputs: this channel quarter only keeps the protocol
          mbar_wait(tfull_bar(acc), acc_ph);
          release_tmem(acc);
          continue;
        }
        if (lane == 0 && kMode != 5) bulk_wait_read0();   // the previous store has finished reading the staging rows
        __syncwarp();
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * q.n + col0);
        uint32_t d0[16], d1[16], d2[16];
        __syncwarp();
        tmem_ld16(t_row, d0);
        tmem_ld16(t_row + (uint32_t)q.gw, d1);
        tmem_ld16(t_row + 2u * (uint32_t)q.gw, d2);
        tmem_ld_wait();
        release_tmem(acc);
        float a[16];                                      // conv + bias of output column x0 + lane (packed fp32x2 adds)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 mid = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(d1[2 * i]), 1),
                                         __shfl_down_sync(0xffffffffu, __uint_as_float(d1[2 * i + 1]), 1));
          const float2 right = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(d2[2 * i]), 2),
                                           __shfl_down_sync(0xffffffffu, __uint_as_float(d2[2 * i + 1]), 2));
          const float2 sum = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(d0[2 * i]), __uint_as_float(d0[2 * i + 1])), mid),
                                        __fadd2_rn(right, make_float2(bb[2 * i], bb[2 * i + 1])));
          a[2 * i] = sum.x; a[2 * i + 1] = sum.y;
        }
        if (kMode == 5) {
          // depth_to_space straight to the fp32 image: the 16 channels of this warp are four groups of four consecutive
          // floats inside the pixel's r x r block (offsets precomputed per warp); lanes 0..29 are output columns
          const int oy = y0 + quad, ox = x0 + lane;
          if (lane < kFOut && oy < p.H && ox < p.W) {
            const int r = p.d2s;
            float* drow = reinterpret_cast<float*>(p.y) +
                          (((size_t)b * p.H * r + (size_t)oy * r) * ((size_t)p.W * r) + (size_t)ox * r) * p.c_post;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float v[4] = {a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]};
              if (!d2s_plain) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float t = v[u];
                  if (p.act == SRB_ACT_RELU) t = fmaxf(t, 0.f);
                  else if (p.act != SRB_ACT_NONE) t = act_generic_fold(t, p.act, p.act_slope);
                  t *= alpha;
                  if (p.clip01) t = fminf(fmaxf(t, 0.f), 1.f);
                  v[u] = t;
                }
              }
              *reinterpret_cast<float4*>(drow + d2s_off[g]) = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
          continue;
        }
        const uint32_t a_h0 = buf + h_row + ((0u ^ h_x) << 4), a_h1 = buf + h_row + ((1u ^ h_x) << 4);
        uint32_t oh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v0 = a[2 * i], v1 = a[2 * i + 1];
          if (kMode == 4) {
            const float2 sl = *reinterpret_cast<const float2*>(slope_s + col0 + 2 * i);
            v0 = fmaf(sl.x, fminf(v0, 0.f), fmaxf(v0, 0.f)); v1 = fmaf(sl.y, fminf(v1, 0.f), fmaxf(v1, 0.f));
          }
          uint32_t pk = bf ? pack2(v0, v1, SRB_BF16) : pack2(v0, v1, SRB_F16);
          if (kMode == 1) {                              // ReLU on the packed pair (rounding is monotonic and 0 is exact)
            if (bf) { const __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&pk), __float2bfloat162_rn(0.f)); pk = *reinterpret_cast<const uint32_t*>(&r2); }
            else { const __half2 r2 = __hmax2(*reinterpret_cast<const __half2*>(&pk), __float2half2_rn(0.f)); pk = *reinterpret_cast<const uint32_t*>(&r2); }
          }
          oh[i] = pk;
        }
        sts128(a_h0, make_uint4(oh[0], oh[1], oh[2], oh[3]));
        sts128(a_h1, make_uint4(oh[4], oh[5], oh[6], oh[7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&em.y, buf, col0, x0, y0 + quad, b);               // rows 0..29 of the staging block
          bulk_commit();
        }
      }
      if (lane == 0) bulk_wait0();
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


// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int conv_fold_mode(const ConvParams& p) {          // -1: not eligible
  if (p.cin != 64 || !p.w_tc_fold || p.kh > 9 || p.kw > 9) return -1;
  auto dt16 = [](int d) { return d == SRB_BF16 || d == SRB_F16; };
  if (p.d2s != 1) {
    // mode 5: depth_to_space straight to a few-channel fp32 image (ESPCN's last layer: 48 -> 4 x 4 x RGB)
    const bool ok = p.kh == 3 && p.kw == 3 && p.cout % 16 == 0 && p.cout >= 16 && p.cout <= 64 && p.y_dtype == SRB_F32 &&
                    !p.res1 && !p.res2 && !p.y2 && p.act != SRB_ACT_PRELU && p.y_cstride == p.c_post && p.y_coffset == 0 &&
                    (p.d2s * p.c_post) % 4 == 0 && aligned16(p.y) &&
                    (long)p.B * p.H * p.W * p.cout < (1L << 31);          // (32-bit element offsets inside the image batch)
    return ok ? 5 : -1;
  }
  if (p.cout <= 4) return (!p.res1 && !p.res2 && !p.y2) ? 3 : -1;
  if (p.kh != 3 || p.kw != 3) return -1;
  if (p.cout % 16 || p.cout < 16 || p.cout > 64 || !dt16(p.y_dtype) || p.y_coffset % 8 || p.y_cstride % 8 || !aligned16(p.y)) return -1;
  if (!p.res1 && !p.res2 && !p.y2 && !p.clip01 && p.alpha == 1.f) {
    if (p.act == SRB_ACT_NONE) return 0;
    if (p.act == SRB_ACT_RELU) return 1;
    if (p.act == SRB_ACT_LEAKY || (p.act == SRB_ACT_PRELU && p.prelu)) return 4;
    return -1;
  }
  // (the pair8 residual layers stay on the 16 x 8-tile kernel: on wide tiles their epilogue - three TMEM loads, two shuffles
  //  per channel and the e5m2 conversions on top - measured 0.125 ms per 32 tiles of 192 x 192 against 0.111 ms there)
  return -1;
}

int conv_fold_launch(const ConvParams& p, int mode, cudaStream_t stream) {
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("conv(tcgen05): cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }
  FoldParams q{};
  q.gw = mode == 3 ? 4 : p.cout;                       // 16 / 32 / 48 / 64 output channels: N = 48 / 96 / 144 / 192
  q.kh = p.kh; q.kw = p.kw;
  q.n = mode == 3 ? ((p.kw * 4 + 15) & ~15) : 3 * p.cout;
  q.out_cols = kFW - (p.kw - 1);
  q.stage_bytes = (uint32_t)(kFH + p.kh - 1) * kFW * 128u;
  q.tiles_x = (p.W + q.out_cols - 1) / q.out_cols;
  q.tiles_y = (p.H + kFH - 1) / kFH;
  const long total = (long)p.B * q.tiles_x * q.tiles_y;
  SRB_REQUIRE(total < (1L << 30), "conv(tcgen05): too many tiles");
  q.total_tiles = (int)total;
  q.tmem_cols = 32;
  while (q.tmem_cols < (uint32_t)(2 * q.n)) q.tmem_cols <<= 1;
  const uint32_t fmt = p.x_dtype == SRB_BF16 ? 1u : 0u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(q.n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  q.epi_warp_bytes = (mode == 3 || mode == 5) ? 0u : 1024u;
  q.reverse = tc_next_reverse();
  q.magic_tpi = div_magic(total + 1, q.tiles_x * q.tiles_y);
  q.magic_tx = div_magic((long)q.tiles_x * q.tiles_y, q.tiles_x);
  int dev = 0, max_smem = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t w_bytes = ((size_t)p.kh * q.n * 128 + 1023) & ~(size_t)1023;
  const size_t tail_bytes = (2 * kMaxStages + 6) * 8 + 2 * 64 * sizeof(float);
  auto smem_need = [&](int st) { return 1024 + w_bytes + (size_t)st * q.stage_bytes + (size_t)kFoldEpiWarps * q.epi_warp_bytes + tail_bytes; };
  q.stages = 4;
  while (q.stages > 2 && smem_need(q.stages) > (size_t)max_smem) --q.stages;
  const size_t smem = smem_need(q.stages);
  if (smem > (size_t)max_smem) { set_error("conv(tcgen05, fold): staging does not fit shared memory"); return SRB_E_UNSUPPORTED; }

  const CUtensorMapDataType tdt = p.x_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmx, tmw;
  {
    const cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.x_cstride * 2, (cuuint64_t)p.W * p.x_cstride * 2, (cuuint64_t)p.H * p.W * p.x_cstride * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kFW, (cuuint32_t)(kFH + p.kh - 1), 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    void* gptr = (void*)((const uint16_t*)p.x + p.x_coffset);
    CUresult r = encode(&tmx, tdt, 4, gptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv(tcgen05, fold): cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {64, (cuuint64_t)p.kh * q.n};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)q.n};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmw, tdt, 2, (void*)p.w_tc_fold, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv(tcgen05, fold): cuTensorMapEncodeTiled(w) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (mode != 3 && mode != 5) {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.y_cstride * 2, (cuuint64_t)p.W * p.y_cstride * 2, (cuuint64_t)p.H * p.W * p.y_cstride * 2};
    const cuuint32_t box[4] = {16, (cuuint32_t)kFOut, 1, 1};
    const cuuint32_t es1[4] = {1, 1, 1, 1};
    void* g = (void*)((const uint16_t*)p.y + p.y_coffset);
    if (encode(&em.y, tdt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 && p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                      : (p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16),
               4, g, dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv(tcgen05, fold): cuTensorMapEncodeTiled(epilogue) failed");
      return SRB_E_CUDA;
    }
  }
  typedef void (*FoldFn)(const CUtensorMap, const CUtensorMap, const EpiMaps, const FoldParams, const ConvParams);
  static const FoldFn kernels[7] = {conv3x3_fold_kernel<0, false>, conv3x3_fold_kernel<1, false>, nullptr,
                                    conv3x3_fold_kernel<3, false>, conv3x3_fold_kernel<4, false>, conv3x3_fold_kernel<3, true>,
                                    conv3x3_fold_kernel<5, false>};
  static size_t configured[7] = {0, 0, 0, 0, 0, 0, 0};
  const int ki = (mode == 3 && (p.kh != 3 || p.kw != 3)) ? 5 : mode == 5 ? 6 : mode;
  if (smem > configured[ki]) {
    SRB_CUDA(cudaFuncSetAttribute(kernels[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ki] = smem;
  }
  int grid = sm_count();
  if ((long)grid > total) grid = (int)total;
  SRB_CUDA(tc_launch(kernels[ki], grid, kFoldThreads, smem, stream, false, tmx, tmw, em, q, p));
  return launch_check("conv3x3_fold_kernel");
}

}  // namespace srb
