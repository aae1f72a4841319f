// RGB head layers with small filters (3x3, 5x5: ESPCN's conv1, the ESRGAN generator's initial conv) as an implicit GEMM on the
// tensor cores WITHOUT an im2col build: the float32 image is first padded to 8 channels of the 16-bit operand type (one
// 16-byte chunk per pixel), and a tile is 128 consecutive pixels of ONE image row.  For a vertical tap dy the A operand is
// then the image row itself: GEMM row r = pixel x0 + r, K = (dx, channel) = kw x 8 contiguous values starting at pixel
// x0 + r - kw/2.  In the un-swizzled K-major canonical layout a core matrix is 8 rows x 16 bytes, so with
//     leading byte offset (next 16-byte chunk of K) = 16 B      stride byte offset (next 8 rows) = 128 B
// consecutive rows simply OVERLAP in shared memory (row r + 1 starts one pixel after row r) and the TMA box of
// {8 channels, 128 + kw pixels, kh rows} is the whole operand for all kh x ceil(kw / 2) MMAs (M128, N64, K16) of the tile -
// `same` padding is the TMA out-of-bounds fill.  No producer warps, no gather: the kernel is warp 0 = TMA, warp 1 = MMA,
// eight epilogue warps (bias, activation, 16-bit rows leave by TMA).  The im2col kernel (conv_headtc.cu) keeps the 9x9 heads
// (K = 72 per row would be 45 MMAs per tile against its 16) and the layers that also emit the e5m2 trunk error.
// EDSR_model.py:102 / ESRGAN_model.py:314 (3x3), SURVEY row A14 (ESPCN 5x5).
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace srb {

constexpr int kH8Tile = 128;                 // pixels per tile (one image row segment)
constexpr int kH8Epi = 8;
constexpr int kH8Threads = 64 + 32 * kH8Epi;
constexpr int kH8Stages = 6;

struct Head8Params {
  int kh, kw, ksteps;        // ksteps = ceil(kw * 8 / 16)
  int box_w;                 // 128 + 2 * ksteps - 1 pixels per halo row
  uint32_t stage_bytes;      // kh * box_w * 16, rounded up to 128
  uint32_t w_bytes;          // kh * 64 * (2 * ksteps) * 16
  int tiles_x, total_tiles;
  uint32_t magic_row, magic_x;
  uint32_t idesc;
};

// un-swizzled K-major shared-memory matrix descriptor: [0,14) start >> 4, [16,30) leading byte offset >> 4 (K direction),
// [32,46) stride byte offset >> 4 (8-row groups), [46,48) version = 1, layout type 0
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// float32 [pixels][C] -> 16-bit [pixels][8], channels past C zero
__global__ void __launch_bounds__(256)
pad_to_nhwc8_kernel(const float* __restrict__ x, int cstride, int coffset, int C, size_t pixels, int bf16, uint4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= pixels) return;
  const float* p = x + i * cstride + coffset;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? __ldg(p + c) : 0.f;
  const int dt = bf16 ? SRB_BF16 : SRB_F16;
  out[i] = make_uint4(pack2(v[0], v[1], dt), pack2(v[2], v[3], dt), pack2(v[4], v[5], dt), pack2(v[6], v[7], dt));
}

__global__ void __launch_bounds__(kH8Threads, 1)
conv_head8_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                  const Head8Params q, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // [epilogue staging 2 x 16 KB][weights][A stages][barriers, bias, slopes]
  const uint32_t epi_smem = base, w_smem = base + 2u * 16384u;
  const uint32_t w_span = (q.w_bytes + 127u) & ~127u;
  const uint32_t a_smem = w_smem + w_span;
  uint8_t* tail = smem + 2u * 16384u + w_span + (size_t)kH8Stages * q.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(kH8Stages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kH8Stages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kH8Stages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kH8Stages + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kH8Stages + 5);        // [64]
  float* slope_s = bias_s + 64;                                               // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kH8Stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kH8Epi / 2); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64; i += kH8Threads) {
    bias_s[i] = p.bias[i];
    slope_s[i] = (p.act == SRB_ACT_PRELU && p.prelu) ? p.prelu[i] : p.act_slope;
  }
  {   // the packed weights (already in the canonical core-matrix order) -> shared memory, then visible to the tensor core
    const uint4* src = reinterpret_cast<const uint4*>(p.w_tc_head8);
    for (uint32_t i = threadIdx.x; i < q.w_bytes / 16u; i += kH8Threads) sts128(w_smem + 16u * i, __ldg(src + i));
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_y); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first_tile = (int)blockIdx.x, tile_step = (int)gridDim.x;
  auto coords = [&](int tile, int& b, int& y, int& x0) {       // tile -> (image, row, first pixel)
    const int row = fast_div(tile, q.tiles_x, q.magic_x);       // global row index b * H + y
    x0 = (tile - row * q.tiles_x) * kH8Tile;
    b = fast_div(row, p.H, q.magic_row);
    y = row - b * p.H;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0; uint32_t ph = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step) {
      int b, y, x0;
      coords(tile, b, y, x0);
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), (uint32_t)(q.kh * q.box_w * 16));
        tma_load_4d(a_smem + (uint32_t)s * q.stage_bytes, &tmap_x, full_bar(s), 0, x0 - (q.kw >> 1), y - (q.kh >> 1), b);
      }
      __syncwarp();
      if (++s == kH8Stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: kh vertical taps x ksteps (two pixels of the row per k16 step) =====================
    int s = 0; uint32_t ph = 0; int it = 0;
    const uint32_t kc_bytes = (uint32_t)(2 * q.ksteps) * 128u;             // one 8-row group of the weights: all K chunks
    const uint32_t w_dy = 8u * kc_bytes;                                   // 64 rows
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a0 = a_smem + (uint32_t)s * q.stage_bytes;
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
      if (elect_one()) {
        for (int dy = 0; dy < q.kh; ++dy) {
          const uint64_t ad = make_desc_ns(a0 + (uint32_t)(dy * q.box_w) * 16u, 16u, 128u);
          const uint64_t bd = make_desc_ns(w_smem + (uint32_t)dy * w_dy, 128u, kc_bytes);
          for (int k = 0; k < q.ksteps; ++k)
            umma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(16 * k), q.idesc, (uint32_t)((dy | k) != 0));
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++s == kH8Stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: two groups of four warps take alternate tiles (group g owns accumulator g and
    // staging buffer g); inside a group warp = TMEM lane quadrant = 32 pixels x all 64 channels =====================
    const int ew = warp - 2, quad = warp & 3, grp = ew >> 2;
    const bool bf = p.y_dtype == SRB_BF16;
    const uint32_t row = (uint32_t)(quad * 32 + lane);
    const bool leader = (ew & 3) == 0 && lane == 0;
    const uint32_t buf = epi_smem + (uint32_t)grp * 16384u;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    int it = 0;
    for (int tile = first_tile; tile < q.total_tiles; tile += tile_step, ++it) {
      if ((it & 1) != grp) continue;
      int b, y, x0;
      coords(tile, b, y, x0);
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      if (leader) bulk_wait_read0();                   // this group's previous store has finished reading its staging buffer
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      mbar_wait(tfull_bar(grp), acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * 64);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                 // 32 channels at a time (bounds the live registers)
        uint32_t ra[16], rb[16];
        __syncwarp();
        tmem_ld16(t_row + 32u * hf, ra);
        tmem_ld16(t_row + 32u * hf + 16u, rb);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(grp));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {                  // four 16-byte chunks = 8 channels each
          uint32_t w4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = 8 * g + 2 * u, c = 32 * hf + e;
            float v0 = __uint_as_float(e < 16 ? ra[e & 15] : rb[e & 15]) + bias_s[c];
            float v1 = __uint_as_float(e < 16 ? ra[(e + 1) & 15] : rb[(e + 1) & 15]) + bias_s[c + 1];
            if (p.act == SRB_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            else if (p.act == SRB_ACT_LEAKY || p.act == SRB_ACT_PRELU) {
              v0 = fmaf(slope_s[c], fminf(v0, 0.f), fmaxf(v0, 0.f));
              v1 = fmaf(slope_s[c + 1], fminf(v1, 0.f), fmaxf(v1, 0.f));
            } else if (p.act == SRB_ACT_TANH) { v0 = tanhf(v0); v1 = tanhf(v1); }
            w4[u] = pack2(v0, v1, bf ? SRB_BF16 : SRB_F16);
          }
          const uint32_t chunk = (uint32_t)(4 * hf + g);                    // 16-byte chunk of the pixel's 128-byte row
          sts128(buf + row * 128u + ((chunk ^ (row & 7u)) << 4), make_uint4(w4[0], w4[1], w4[2], w4[3]));
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        tma_store_4d(&tmap_y, buf, 0, x0, y, b);       // {64 channels, 128 pixels}: clipped at the right image edge
        bulk_commit();
      }
    }
    if (leader) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

bool conv_head8_eligible(const ConvParams& p) {
  static const bool enabled = getenv("SRB_NO_HEAD8") == nullptr;
  if (!enabled) return false;
  if (!p.w_tc_head8 || p.cin > 8 || p.cout != 64 || p.x_dtype != SRB_F32) return false;
  if (p.kh != p.kw || (p.kh != 3 && p.kh != 5) || p.d2s != 1 || p.res1 || p.res2 || p.y2 || p.clip01 || p.alpha != 1.f) return false;
  if (p.y_dtype != SRB_F16 && p.y_dtype != SRB_BF16) return false;
  if (p.y_cstride % 8 || p.y_coffset % 8 || !aligned16(p.y)) return false;
  if (p.act == SRB_ACT_PRELU && !p.prelu) return false;
  return (long)p.B * p.H < (1L << 24);
}

int conv_head8_launch(const ConvParams& p, cudaStream_t stream) {
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("conv(head8, tcgen05): cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }
  Head8Params q{};
  q.kh = p.kh; q.kw = p.kw;
  q.ksteps = (p.kw * 8 + 15) / 16;
  q.box_w = kH8Tile + 2 * q.ksteps - 1;
  q.stage_bytes = ((uint32_t)(p.kh * q.box_w * 16) + 127u) & ~127u;
  q.w_bytes = (uint32_t)(p.kh * 64 * 2 * q.ksteps * 16);
  q.tiles_x = (p.W + kH8Tile - 1) / kH8Tile;
  const long rows = (long)p.B * p.H, total = rows * q.tiles_x;
  SRB_REQUIRE(total < (1L << 30), "conv(head8, tcgen05): too many tiles");
  q.total_tiles = (int)total;
  q.magic_x = div_magic(total + 1, q.tiles_x);
  q.magic_row = div_magic(rows + 1, p.H);
  const uint32_t fmt = p.y_dtype == SRB_BF16 ? 1u : 0u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = 1024 + 2 * 16384 + ((q.w_bytes + 127u) & ~127u) + (size_t)kH8Stages * q.stage_bytes +
                      (2 * kH8Stages + 5) * 8 + 2 * 64 * sizeof(float) + 64;

  // the image in the operand layout: 8 channels of the 16-bit type per pixel (stream-ordered scratch)
  const size_t pixels = (size_t)p.B * p.H * p.W;
  {
    // keep the stream-ordered pool from handing the scratch back to the driver at every synchronisation (default threshold 0)
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
  uint4* x8 = nullptr;
  SRB_CUDA(cudaMallocAsync(&x8, pixels * 16, stream));
  pad_to_nhwc8_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float*>(p.x), p.x_cstride, p.x_coffset, p.cin,
                                                                           pixels, p.y_dtype == SRB_BF16, x8);
  int rc = launch_check("pad_to_nhwc8_kernel");
  const CUtensorMapDataType tdt = p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmx, tmy;
  if (rc == SRB_OK) {
    const cuuint64_t dims[4] = {8, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {16, (cuuint64_t)p.W * 16, (cuuint64_t)p.H * p.W * 16};
    const cuuint32_t box[4] = {8, (cuuint32_t)q.box_w, (cuuint32_t)p.kh, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&tmx, tdt, 4, x8, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv(head8, tcgen05): cuTensorMapEncodeTiled(x) failed");
      rc = SRB_E_CUDA;
    }
  }
  if (rc == SRB_OK) {
    const cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.y_cstride * 2, (cuuint64_t)p.W * p.y_cstride * 2, (cuuint64_t)p.H * p.W * p.y_cstride * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kH8Tile, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    void* g = (void*)((const uint16_t*)p.y + p.y_coffset);
    if (encode(&tmy, tdt, 4, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv(head8, tcgen05): cuTensorMapEncodeTiled(y) failed");
      rc = SRB_E_CUDA;
    }
  }
  if (rc == SRB_OK) {
    static size_t configured = 0;
    cudaError_t e = cudaSuccess;
    if (smem > configured) {
      e = cudaFuncSetAttribute(conv_head8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) configured = smem;
    }
    int grid = sm_count();
    if ((long)grid > total) grid = (int)total;
    if (e == cudaSuccess) e = tc_launch(conv_head8_kernel, grid, kH8Threads, smem, stream, false, tmx, tmy, q, p);
    rc = e == cudaSuccess ? launch_check("conv_head8_kernel") : cuda_fail(e, "conv_head8_kernel");
  }
  cudaFreeAsync(x8, stream);
  return rc;
}

}  // namespace srb
