// RGB head layers on the tensor cores: K x K, Cin = 3 -> 64 convolutions of fp32 NHWC images as an im2col GEMM
// (EDSR_model.py:102 3x3, ESRGAN_model.py:314 3x3; the 5x5 ESPCN and 9x9 SRResNet heads of BASELINE configs 2 and 4).
//
//   D[128 pixels x 64] = A[128 x Kpad] * W[Kpad x 64],   K = kh * kw * 3 padded to a multiple of 16 (27 -> 32, 75 -> 80, 243 -> 256)
//
// K is too irregular for TMA to build the A operand, so eight producer warps do it: the fp32 halo of a 16 x 8 pixel tile is
// fetched with cp.async (double buffered; zero outside the image), and every producer thread gathers eight consecutive k
// of one pixel through a k -> halo-offset table, converts them to the 16-bit operand type and writes one 16-byte chunk of
// the K-major, 128-byte-swizzled UMMA layout (64-wide K blocks of 128 rows x 128 B).  One warp issues the Kpad / 16
// tcgen05.mma per tile into double-buffered TMEM accumulators; eight epilogue warps apply bias / activation / alpha,
// produce the 16-bit output (and, for the pair8 trunk, the e5m2 rounding error of it) and store their 4-row x 8-pixel x
// 32-channel blocks with TMA, exactly like the 3x3 kernels of conv_tc.cu.  Padded k carry zero weights.
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace srb {

constexpr int kHTileH = 16, kHTileW = 8;
constexpr int kHProd = 8, kHEpi = 8;
constexpr int kHProdThreads = 32 * kHProd;
constexpr int kHThreads = 32 * (kHProd + 1 + kHEpi);     // 544
constexpr int kHStages = 2;
constexpr int kHRing = 2;                                // halo buffers: the next tile is fetched while this one is built
constexpr int kHMaxHalo = 5;                             // halo floats per producer thread (9x9: 24 x 16 x 3 = 1,152)

struct HeadTcParams {
  int K, kpad, n_kb;
  int halo_w, halo_h, halo_n;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_img, magic_x;                           // multiply-high reciprocals of tiles_per_img and tiles_x (0: divide)
  uint32_t idesc, tmem_cols, a_stage_bytes;
  int y2_f8;
};

struct HeadMaps { CUtensorMap y, y2; };

__global__ void __launch_bounds__(kHThreads, 1)
conv_headtc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ HeadMaps em, const HeadTcParams q,
                   const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t w_bytes = (uint32_t)q.n_kb * 8192u;
  const uint32_t w_smem = base, a_smem = base + w_bytes;
  const uint32_t epi_smem = a_smem + kHStages * q.a_stage_bytes;
  constexpr uint32_t kEpiWarpBytes = 3072u;               // 32 rows x 64 B (16-bit) + 32 rows x 32 B (e5m2)
  uint8_t* after = smem + w_bytes + (size_t)kHStages * q.a_stage_bytes + kHEpi * kEpiWarpBytes;
  float* halo = reinterpret_cast<float*>(after);                         // [kHRing][halo_n]
  int* koff = reinterpret_cast<int*>(halo + kHRing * q.halo_n);         // [kpad]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(koff + q.kpad) + 15) & ~(uintptr_t)15);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(2 + s); };
  const uint32_t wfull_bar = bar0 + 8u * 4u;
  auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(5 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(7 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* bias_s = reinterpret_cast<float*>(bars + 10);                   // [64]
  float* slope_s = bias_s + 64;                                          // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kHStages; ++s) { mbar_init(full_bar(s), kHProd); mbar_init(empty_bar(s), 1); }
    mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kHEpi); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64; i += kHThreads) {
    bias_s[i] = p.bias[i];
    slope_s[i] = p.act == SRB_ACT_PRELU ? p.prelu[i] : p.act_slope;
  }
  for (int k = threadIdx.x; k < q.kpad; k += kHThreads) {
    int off = 0;                                          // (padded k: any valid halo element - its weights are zero)
    if (k < q.K) { const int tap = k / 3, c = k - tap * 3, dy = tap / p.kw, dx = tap - dy * p.kw; off = (dy * q.halo_w + dx) * 3 + c; }
    koff[k] = off;
  }
  if (warp == kHProd) {
    if (lane == 0) { prefetch_tmap(&tmap_w); prefetch_tmap(&em.y); if (p.y2) prefetch_tmap(&em.y2); }
    tmem_alloc(smem_u32(tmem_slot), q.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = q.tiles_x * q.tiles_y;
  const int first_tile = (int)blockIdx.x, tile_step = (int)gridDim.x;
  auto coords = [&](int tile, int& b, int& y0, int& x0) {
    int rr_, ty;
    if (q.magic_img) {                                  // exact for every tile index of this launch (checked on the host)
      b = (int)__umulhi((uint32_t)tile, q.magic_img);
      rr_ = tile - b * tiles_per_img;
      ty = (int)__umulhi((uint32_t)rr_, q.magic_x);
    } else {
      b = tile / tiles_per_img;
      rr_ = tile - b * tiles_per_img;
      ty = rr_ / q.tiles_x;
    }
    y0 = ty * kHTileH;
    x0 = (rr_ - ty * q.tiles_x) * kHTileW;
  };

  if (warp < kHProd) {
    // ===================== producers: halo fetch + im2col into the UMMA operand layout =====================
    const int tid = threadIdx.x;
    if (tid == 0) {
      mbar_expect_tx(wfull_bar, w_bytes);
      for (int kb = 0; kb < q.n_kb; ++kb) tma_load_2d(w_smem + (uint32_t)kb * 8192u, &tmap_w, wfull_bar, 0, kb * 64);
    }
    // tile-invariant halo positions of this thread: offset of the element relative to the tile's first pixel
    int h_idx[kHMaxHalo], h_dy[kHMaxHalo], h_dx[kHMaxHalo], h_rel[kHMaxHalo];
#pragma unroll
    for (int j = 0; j < kHMaxHalo; ++j) {
      const int i = tid + j * kHProdThreads;
      const int hp = i / 3;
      h_idx[j] = i < q.halo_n ? i : -1;
      h_dx[j] = hp % q.halo_w - (p.kw >> 1);
      h_dy[j] = hp / q.halo_w - (p.kh >> 1);
      h_rel[j] = (h_dy[j] * p.W + h_dx[j]) * p.x_cstride + p.x_coffset + (i - hp * 3);
    }
    const float* xin = reinterpret_cast<const float*>(p.x);
    auto fetch = [&](int tile, int buf) {
      if (tile < q.total_tiles) {
        int b, y0, x0;
        coords(tile, b, y0, x0);
        const float* tile0 = xin + (((size_t)b * p.H + y0) * p.W + x0) * p.x_cstride;
        const uint32_t d0 = smem_u32(halo + buf * q.halo_n);
        // interior tiles (the whole halo inside the image) skip the per-element range checks
        const bool inside = y0 >= (p.kh >> 1) && y0 + kHTileH + (p.kh >> 1) <= p.H && x0 >= (p.kw >> 1) && x0 + kHTileW + (p.kw >> 1) <= p.W;
#pragma unroll
        for (int j = 0; j < kHMaxHalo; ++j) {
          if (h_idx[j] >= 0) {
            const uint32_t d = d0 + 4u * (uint32_t)h_idx[j];
            const int gy = y0 + h_dy[j], gx = x0 + h_dx[j];
            if (inside || (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)) {
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(tile0 + h_rel[j]) : "memory");
            } else {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(0.f) : "memory");
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int qd = warp & 3, kc0 = warp >> 2;
    const int m = qd * 32 + lane;                          // GEMM row = pixel (m >> 3, m & 7) of the tile
    const int origin = ((m >> 3) * q.halo_w + (m & 7)) * 3;
    const uint32_t row_off = (uint32_t)m * 128u, row_x = (uint32_t)(m & 7);
    const bool bf = p.y_dtype == SRB_BF16;
    const int n_kc = q.kpad >> 3;
#pragma unroll
    for (int j = 0; j < kHRing - 1; ++j) fetch(first_tile + j * tile_step, j);
    int it = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      asm volatile("bar.sync 1, 256;" ::: "memory");       // every producer is done reading the halo buffer refilled next
      fetch(tile + (kHRing - 1) * tile_step, (it + kHRing - 1) & (kHRing - 1));
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // (kHRing - 1 groups may stay in flight)
      asm volatile("bar.sync 1, 256;" ::: "memory");       // this tile's halo is complete and visible to all producers
      mbar_wait(empty_bar(s), ph ^ 1u);                    // the MMAs that read this operand stage have retired
      const float* hp = halo + (it & (kHRing - 1)) * q.halo_n + origin;
      const uint32_t a_stage = a_smem + (uint32_t)s * q.a_stage_bytes;
      for (int kc = kc0; kc < n_kc; kc += 2) {
        const int4 o0 = *reinterpret_cast<const int4*>(koff + kc * 8), o1 = *reinterpret_cast<const int4*>(koff + kc * 8 + 4);
        const float f0 = hp[o0.x], f1 = hp[o0.y], f2 = hp[o0.z], f3 = hp[o0.w], f4 = hp[o1.x], f5 = hp[o1.y], f6 = hp[o1.z], f7 = hp[o1.w];
        uint4 pk;
        if (bf) pk = make_uint4(pack2(f0, f1, SRB_BF16), pack2(f2, f3, SRB_BF16), pack2(f4, f5, SRB_BF16), pack2(f6, f7, SRB_BF16));
        else pk = make_uint4(pack2(f0, f1, SRB_F16), pack2(f2, f3, SRB_F16), pack2(f4, f5, SRB_F16), pack2(f6, f7, SRB_F16));
        sts128(a_stage + (uint32_t)(kc >> 3) * 16384u + row_off + ((((uint32_t)kc & 7u) ^ row_x) << 4), pk);
      }
      fence_proxy_async_smem();                            // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == kHProd) {
    // ===================== MMA issuer =====================
    mbar_wait(wfull_bar, 0);
    const int n_k = q.kpad >> 4;
    int it = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      const int s = it & 1, acc = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), ph ^ 1u);
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_stage = a_smem + (uint32_t)s * q.a_stage_bytes;
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
      if (elect_one()) {
        for (int k = 0; k < n_k; ++k) {
          const uint32_t kb = (uint32_t)k >> 2, ks = (uint32_t)k & 3u;
          umma_f16(d_tmem, make_desc(a_stage + kb * 16384u + ks * 32u, 1024u), make_desc(w_smem + kb * 8192u + ks * 32u, 1024u),
                   q.idesc, (uint32_t)(k != 0));
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM -> bias / activation -> 16-bit (+ e5m2 error) rows -> TMA stores =====================
    const int ew = warp - (kHProd + 1), quad = warp & 3, half = ew >> 2;
    const int col0 = half * 32;
    constexpr uint32_t hb = 64u, lb = 32u;
    const uint32_t buf0 = epi_smem + (uint32_t)ew * kEpiWarpBytes;
    const uint32_t row = (uint32_t)lane;
    const uint32_t h_row = row * hb, h_x = (row >> 1) & 3u;
    const uint32_t l_row = 32u * hb + row * lb, l_x = (row >> 2) & 1u;
    const bool bf = p.y_dtype == SRB_BF16;
    const int act = p.act;
    const float alpha = p.alpha;
    const bool clip = p.clip01 != 0, err_out = q.y2_f8 != 0;
    const bool scaled = alpha != 1.f || clip;
    float bias_r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[col0 + i];
    int it = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      int b, y0, x0;
      coords(tile, b, y0, x0);
      const int acc = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      const uint32_t buf = buf0;
      if (lane == 0) bulk_wait_read0();                   // the previous store has drained the staging rows
      __syncwarp();
      mbar_wait(tfull_bar(acc), ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64 + col0);
      uint32_t rr[2][16];
      __syncwarp();
      tmem_ld16(t_row, rr[0]);
      tmem_ld16(t_row + 16u, rr[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      // the activation is a warp-uniform switch around the fully unrolled 32-channel block (one code copy per kind)
      auto finish = [&](auto kind) {
        constexpr int kAct = decltype(kind)::value;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t oh[8], ol[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ci = 16 * c + 2 * i + u;
              float t = __uint_as_float(rr[c][2 * i + u]) + bias_r[ci];
              if (kAct == SRB_ACT_RELU) t = fmaxf(t, 0.f);
              else if (kAct == SRB_ACT_PRELU) t = fmaf(slope_s[col0 + ci], fminf(t, 0.f), fmaxf(t, 0.f));
              else if (kAct == SRB_ACT_TANH) t = tanhf(t);
              if (scaled) { t *= alpha; if (clip) t = fminf(fmaxf(t, 0.f), 1.f); }
              v[u] = t;
            }
            const uint32_t pk = bf ? pack2(v[0], v[1], SRB_BF16) : pack2(v[0], v[1], SRB_F16);
            oh[i] = pk;
            if (err_out) {
              float2 back;
              if (bf) back = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk));
              else back = __half22float2(*reinterpret_cast<const __half2*>(&pk));
              const uint32_t e2 = __nv_cvt_float2_to_fp8x2(make_float2(v[0] - back.x, v[1] - back.y), __NV_SATFINITE, __NV_E5M2);
              if (i & 1) ol[i >> 1] |= e2 << 16; else ol[i >> 1] = e2;
            }
          }
          sts128(buf + h_row + (((2u * (uint32_t)c) ^ h_x) << 4), make_uint4(oh[0], oh[1], oh[2], oh[3]));
          sts128(buf + h_row + (((2u * (uint32_t)c + 1u) ^ h_x) << 4), make_uint4(oh[4], oh[5], oh[6], oh[7]));
          if (err_out) sts128(buf + l_row + (((uint32_t)c ^ l_x) << 4), make_uint4(ol[0], ol[1], ol[2], ol[3]));
        }
      };
      if (act == SRB_ACT_RELU) finish(std::integral_constant<int, SRB_ACT_RELU>{});
      else if (act == SRB_ACT_PRELU || act == SRB_ACT_LEAKY) finish(std::integral_constant<int, SRB_ACT_PRELU>{});
      else if (act == SRB_ACT_TANH) finish(std::integral_constant<int, SRB_ACT_TANH>{});
      else finish(std::integral_constant<int, SRB_ACT_NONE>{});
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&em.y, buf, col0, x0, y0 + quad * 4, b);
        if (err_out) tma_store_4d(&em.y2, buf + 32u * hb, col0, x0, y0 + quad * 4, b);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kHProd) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, q.tmem_cols);
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool conv_headtc_eligible(const ConvParams& p) {
  static const bool enabled = getenv("SRB_NO_HEAD_TC") == nullptr;
  if (!enabled || p.cin != 3 || p.cout != 64 || !p.w_tc_head || p.x_dtype != SRB_F32) return false;
  if (p.kh > 9 || p.kw > 9 || p.d2s != 1 || p.res1 || p.res2) return false;
  if (p.y_dtype != SRB_F16 && p.y_dtype != SRB_BF16) return false;
  if (p.y_cstride % 8 || p.y_coffset % 8 || !al16(p.y)) return false;
  if (p.y2 && !(p.y2_dtype == SRB_F8E5M2 && p.y2_mode == 1 && p.y2_cstride % 16 == 0 && al16(p.y2))) return false;
  if ((kHTileH + p.kh - 1) * (kHTileW + p.kw - 1) * 3 > kHMaxHalo * kHProdThreads) return false;
  return true;
}

int conv_headtc_launch(const ConvParams& p, cudaStream_t stream) {
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("conv(head, tcgen05): cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }
  HeadTcParams q{};
  q.K = p.kh * p.kw * 3;
  q.kpad = (q.K + 15) & ~15;
  q.n_kb = (q.kpad + 63) / 64;
  if (q.n_kb != p.w_tc_head_kb) { set_error("conv(head, tcgen05): packed weights do not match the filter size"); return SRB_E_INVALID; }
  q.halo_w = kHTileW + p.kw - 1; q.halo_h = kHTileH + p.kh - 1; q.halo_n = q.halo_w * q.halo_h * 3;
  q.tiles_x = (p.W + kHTileW - 1) / kHTileW; q.tiles_y = (p.H + kHTileH - 1) / kHTileH;
  const long total = (long)p.B * q.tiles_x * q.tiles_y;
  SRB_REQUIRE(total < (1L << 30), "conv(head, tcgen05): too many tiles");
  q.total_tiles = (int)total;
  {
    // multiply-high division (idx * magic >> 32 with magic = floor((2^32 - 1) / d) + 1) is exact while idx * d < 2^32
    const uint64_t tpi = (uint64_t)q.tiles_x * q.tiles_y;
    const bool ok = (uint64_t)total * tpi < (1ull << 32) && tpi * (uint64_t)q.tiles_x < (1ull << 32) && tpi > 1 && q.tiles_x > 1;
    q.magic_img = ok ? (uint32_t)(0xFFFFFFFFull / tpi) + 1u : 0u;
    q.magic_x = ok ? (uint32_t)(0xFFFFFFFFull / (uint64_t)q.tiles_x) + 1u : 0u;
  }
  const uint32_t fmt = p.y_dtype == SRB_BF16 ? 1u : 0u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  q.tmem_cols = 128;
  q.a_stage_bytes = (uint32_t)q.n_kb * 16384u;
  q.y2_f8 = p.y2 ? 1 : 0;
  const size_t smem = 1024 + (size_t)q.n_kb * 8192 + (size_t)kHStages * q.a_stage_bytes + (size_t)kHEpi * 3072 +
                      (size_t)kHRing * q.halo_n * 4 + (size_t)q.kpad * 4 + 16 + 10 * 8 + 128 * 4;
  int dev = 0, max_smem = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem > (size_t)max_smem) { set_error("conv(head, tcgen05): staging does not fit shared memory"); return SRB_E_UNSUPPORTED; }

  const CUtensorMapDataType tdt = p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmw;
  {
    const cuuint64_t dims[2] = {64, (cuuint64_t)q.n_kb * 64};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 64};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmw, tdt, 2, (void*)p.w_tc_head, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv(head, tcgen05): cuTensorMapEncodeTiled(w) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  HeadMaps em;
  memset(&em, 0, sizeof(em));
  auto encode_out = [&](CUtensorMap* m, const void* ptr, int coffset, int cstride, bool f8) -> bool {
    const size_t es = f8 ? 1 : 2;
    const cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)cstride * es, (cuuint64_t)p.W * cstride * es, (cuuint64_t)p.H * p.W * cstride * es};
    const cuuint32_t box[4] = {32, (cuuint32_t)kHTileW, 4, 1};
    const cuuint32_t es1[4] = {1, 1, 1, 1};
    void* g = (void*)((const uint8_t*)ptr + (size_t)coffset * es);
    return encode(m, f8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : tdt, 4, g, dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  f8 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  bool ok = encode_out(&em.y, p.y, p.y_coffset, p.y_cstride, false);
  if (p.y2) ok = ok && encode_out(&em.y2, p.y2, 0, p.y2_cstride, true);
  if (!ok) { set_error("conv(head, tcgen05): cuTensorMapEncodeTiled(output) failed"); return SRB_E_CUDA; }

  static size_t configured = 0;
  if (smem > configured) {
    SRB_CUDA(cudaFuncSetAttribute(conv_headtc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = sm_count();
  if ((long)grid > total) grid = (int)total;
  SRB_CUDA(tc_launch(conv_headtc_kernel, grid, kHThreads, smem, stream, false, tmw, em, q, p));
  return launch_check("conv_headtc_kernel");
}

}  // namespace srb
