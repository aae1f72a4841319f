// tcgen05 / TMEM implicit-GEMM engine for odd K x K convolutions on NHWC bf16 / fp16 activations with Cin = 64 (the EDSR /
// SRResNet / ESRGAN body, up-sampling and tail layers: EDSR_model.py:61,65,80-89,112,121) or wider (VGG16 blocks 2-5,
// VGG16_model.py:69-83; the ESRGAN dense blocks' growing concatenation, ESRGAN_model.py:230-246): 64-channel K chunks.
// The wide-tile, dx-folded sibling for layers with few output channels per MMA lives in conv_fold.cu.
//
// GEMM view per CTA tile:  D[128 pixels x N] += sum over 9 taps  A_tap[128 x 64] * W_tap[64 x N]
//   * a tile is 16 rows x 8 columns of output pixels of one image; pixel (ty, tx) is GEMM row ty*8+tx, so
//     every 8-row swizzle group of the A operand is one image row of the tile;
//   * the input halo (18 x 10 pixels x 64 channels = 128 B per pixel) is brought in by ONE 4-D TMA load
//     per tile with 128-byte swizzle; `same` zero padding is the TMA out-of-bounds fill;
//   * the nine taps are nine UMMA shared-memory descriptors into that one halo tile, shifted by
//     (dy * pitch + dx) pixels; K = 64 channels is four k16 steps inside the 128-byte swizzle row;
//   * the layer's weights [9][N][64] stay resident in shared memory for the whole persistent CTA;
//   * accumulators live in TMEM, double buffered, so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * the epilogue (4 warps, one pixel per thread) fuses bias, activation, alpha, two scaled residuals,
//     clip, the depth_to_space address permutation and an optional second-dtype copy of the output.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (two warps per TMEM lane quadrant, each taking half of the accumulator columns).
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace srb {

constexpr int kTileH = 16, kTileW = 8, kTileM = kTileH * kTileW;   // 128 GEMM rows
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct TcParams {
  int n_tile;            // output channels per CTA (multiple of 16, <= 128)
  int n_chunks;          // ceil(cout_pad / n_tile); CTA c owns chunk c % n_chunks
  int w_rows;            // rows per tap in the packed weight matrix
  int tiles_x, tiles_y, total_tiles;
  int stages;
  uint32_t stage_bytes;  // multiple of 1024
  int pitch;             // halo row pitch in pixels (= TMA box width)
  uint32_t load_bytes;   // bytes of the one halo tile a stage holds
  uint32_t tmem_cols;
  uint32_t idesc;
  int epi_mode;          // 0: generic scalar epilogue; 1: staged vector epilogue (smem transpose, coalesced 16-B accesses);
                         // 2: few-channel epilogue (cout <= 4, e.g. an RGB tail with a filter the fold kernel does not take);
                         // 3: 8 / 16 channels as a 16-bit channel slice (ESRGAN growth convs): 16-byte stores per pixel;
                         // 4: depth_to_space to few channels (ESPCN): r*c_post contiguous floats per output row
  int f_bufs;            // per-warp fp32 staging buffers (0, 1, or 2 when the residual is prefetched)
  int f_dst, h_dst;      // which output is fp32 / 16-bit: 0 none, 1 = y, 2 = y2
  int res_prefetch;      // 1: res1 is fp32 and is prefetched into the F buffers with cp.async; 2: 16-bit (res1, res2) pair
                         // prefetched the same way; 3: pair8 trunk (TMA epilogue)
  uint32_t epi_warp_bytes;
  int two_cta;           // launched as CTA pairs (cluster of 2, cta_group::2 MMAs)
  int n_kc;              // 64-channel K chunks of the input (cin_pad / 64): one halo load and one set of tap MMAs per chunk
  int kh, kw;            // filter size (odd, <= 9); 3x3 is the unrolled fast path
  int halo_rows;         // kTileH + kh - 1
  uint32_t magic_tpi, magic_tx;   // ceil(2^32 / tiles_per_img), ceil(2^32 / tiles_x) for multiply-high division, or 0
  int reverse;           // walk the tiles from the last to the first (alternate launches: the next layer starts on the
                         // rows the previous one wrote last, which are still in L2)
  int tma_epi;           // staged epilogue moves its rows with TMA (residual loads, output stores) instead of per-lane copies
};

// ---------------------------------------------------------------------------------------------------
// epilogue helpers: 8 consecutive channels at a time
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const void* base, int dtype, size_t idx, float (&v)[8]) {
  if (dtype == SRB_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == SRB_BF16) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      }
    }
  }
}

__device__ __forceinline__ void store8(void* base, int dtype, size_t idx, const float (&v)[8]) {
  if (dtype == SRB_F32) {
    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == SRB_BF16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
__device__ __noinline__ float act_generic(float v, int act, float slope) { return apply_act(v, act, slope); }

// generic per-element epilogue, kept out of line so that the staged path stays small in the instruction cache
__device__ __noinline__ void epilogue_store_generic(const ConvParams& p, int b, int y, int x, int co, float acc) {
  epilogue_store(p, b, y, x, co, acc);
}

// kSpec selects a compile-time specialisation of the staged epilogue (the runtime-flag version costs ~14
// instructions per channel, the specialised ones ~4):
//   0 generic (all flags read at run time)      1 act none, 16-bit y            2 ReLU, 16-bit y
//   3 fp32 residual, fp32 y + 16-bit y2         4 fp32 residual, 16-bit y
//   5 16-bit (hi, lo) residual pair -> 16-bit y + its rounding error y2 (compensated trunk, all in the F rows)
//   7 PReLU / leaky ReLU, 16-bit y (TMA epilogue only; slopes staged in shared memory)
//   6 (16-bit hi, e5m2 lo) residual pair -> 16-bit y + its rounding error y2 in e5m2 ("pair8" trunk: 3 bytes per
//     channel carry ~14 significant bits; staging = one 16-bit and one 8-bit row region per prefetch buffer)
// k2: cta_group::2 - a cluster of two CTAs issues M = 256 MMAs (128 pixels per CTA) from the leader; each CTA keeps
// only half of the weight rows (the tensor cores fetch the other half from the peer), which cuts the per-CTA
// B-operand shared-memory traffic and the resident weight footprint in half.
template <int kSpec, bool k2>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ EpiMaps em, const TcParams q, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is what the 128-byte swizzle pattern of TMA and UMMA is anchored to
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int n_taps = q.kh * q.kw;
  const uint32_t rank = k2 ? cluster_ctarank() : 0u;          // 0 = leader of the pair
  const uint32_t n_local = k2 ? (uint32_t)q.n_tile / 2u : (uint32_t)q.n_tile;   // weight rows held by this CTA
  const int n_kc = q.n_kc;
  const uint32_t w_kc_bytes = (uint32_t)n_taps * n_local * 128u;      // one K chunk of this CTA's weight rows, all taps
  const uint32_t w_bytes = w_kc_bytes * (uint32_t)n_kc;
  const uint32_t w_span = (w_bytes + 1023u) & ~1023u;
  const uint32_t w_smem = base;
  const uint32_t a_smem = base + w_span;
  const uint32_t epi_smem = a_smem + (uint32_t)q.stages * q.stage_bytes;
  uint8_t* tail = smem + w_span + (size_t)q.stages * q.stage_bytes + (uint32_t)kEpiWarps * q.epi_warp_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);               // full[S], empty[S], w_full, tfull[2], tempty[2]
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(kMaxStages + s); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 1 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * kMaxStages + 3 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);   // [n_tile], 16-byte aligned
  const uint32_t rbar0 = smem_u32(bias_s) + (uint32_t)q.n_tile * 4u;       // [kEpiWarps][2] residual-landed barriers (TMA epilogue)
  float* slope_s = bias_s + q.n_tile + 4 * kEpiWarps;                      // [n_tile] PReLU / leaky slopes (kSpec 7)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units: a CTA (or a CTA pair) owns channel chunk `unit % n_chunks` and every (units / n_chunks)-th tile (pair)
  const unsigned unit = k2 ? blockIdx.x >> 1 : blockIdx.x, units = k2 ? gridDim.x >> 1 : gridDim.x;
  const int chunk = (int)(unit % (unsigned)q.n_chunks);
  const int first_tile = (int)(unit / (unsigned)q.n_chunks) * (k2 ? 2 : 1) + (int)rank;
  const int tile_step = (int)(units / (unsigned)q.n_chunks) * (k2 ? 2 : 1);
  // both CTAs of a pair run the same number of iterations; the odd one may get a dummy tile past the end
  const int tile_end = q.total_tiles + (int)rank;
  const int co_base = chunk * q.n_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < q.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      // one arrival per participating epilogue warp (an elected lane, after the warp's TMEM reads have completed)
      mbar_init(tempty_bar(a), ((q.n_tile < 32 && q.epi_mode != 1) ? kEpiWarps / 2 : kEpiWarps) * (k2 ? 2 : 1));
    }
    if (q.tma_epi) for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(rbar0 + 8u * (uint32_t)i, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < q.n_tile; i += kThreads) bias_s[i] = (co_base + i < p.cout) ? p.bias[co_base + i] : 0.f;
  if (kSpec == 7) {
    for (int i = threadIdx.x; i < q.n_tile; i += kThreads) {
      const int co = co_base + i;
      slope_s[i] = (p.act == SRB_ACT_PRELU && co < p.cout) ? p.prelu[p.d2s > 1 ? co % p.c_post : co] : p.act_slope;
    }
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
  if (warp == 2 && lane == 0 && q.tma_epi) {
    prefetch_tmap(&em.y);
    if (kSpec == 6) { prefetch_tmap(&em.r1); prefetch_tmap(&em.r2); if (p.y2) prefetch_tmap(&em.y2); }
  }
  if (warp == 1) { if (k2) tmem_alloc_2sm(smem_u32(tmem_slot), q.tmem_cols); else tmem_alloc(smem_u32(tmem_slot), q.tmem_cols); }
  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();   // barrier inits must be visible to the peer before it signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = q.tiles_x * q.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // the whole warp runs the (warp-uniform) loop so that addresses live in uniform registers; one elected
    // lane issues the asynchronous copies
    if (elect_one()) {
      // k2: both CTAs' weight halves and A tiles report to the LEADER's barriers (it issues the MMAs for the pair)
      if (!k2) mbar_expect_tx(wfull_bar, w_bytes);
      else if (rank == 0) mbar_expect_tx(wfull_bar, 2u * w_bytes);
      for (int kc = 0; kc < n_kc; ++kc)
        for (int t = 0; t < n_taps; ++t) {
          const uint32_t dstw = w_smem + (uint32_t)kc * w_kc_bytes + (uint32_t)t * n_local * 128u;
          const int row = t * q.w_rows + co_base + (int)(rank * n_local);
          if (k2) tma_load_2d_2sm(dstw, &tmap_w, wfull_bar, kc * 64, row); else tma_load_2d(dstw, &tmap_w, wfull_bar, kc * 64, row);
        }
    }
    __syncwarp();
    int s = 0; uint32_t ph = 0;
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      const int tl = q.reverse ? q.total_tiles - 1 - tile : tile;
      const int b = fast_div(tl, tiles_per_img, q.magic_tpi), r = tl - b * tiles_per_img;   // (dummy tile: b == B, zero-filled by TMA)
      const int ty_ = fast_div(r, q.tiles_x, q.magic_tx);
      const int y0 = ty_ * kTileH, x0 = (r - ty_ * q.tiles_x) * kTileW;
      for (int kc = 0; kc < n_kc; ++kc) {              // one halo tile per 64-channel K chunk
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t dst = a_smem + (uint32_t)s * q.stage_bytes;
        if (elect_one()) {
          if (!k2) mbar_expect_tx(full_bar(s), q.load_bytes);
          else if (rank == 0) mbar_expect_tx(full_bar(s), 2u * q.load_bytes);
          if (k2) tma_load_4d_2sm(dst, &tmap_x, full_bar(s), kc * 64, x0 - (q.kw >> 1), y0 - (q.kh >> 1), b);
          else tma_load_4d(dst, &tmap_x, full_bar(s), kc * 64, x0 - (q.kw >> 1), y0 - (q.kh >> 1), b);
        }
        __syncwarp();
        if (++s == q.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop (descriptors in uniform registers); one elected lane issues the 36 MMAs of a tile.
    // k2: only the leader CTA issues (cta_group::2 instructions act on both CTAs' shared and tensor memory)
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc_flag) {
      if (k2) umma_f16_2sm(d, ad, bd, q.idesc, acc_flag); else umma_f16(d, ad, bd, q.idesc, acc_flag);
    };
    auto commit = [&](uint32_t bar) { if (k2) umma_commit_2sm(bar); else umma_commit(bar); };
    if (!k2 || rank == 0) {
    mbar_wait(wfull_bar, 0);
    int s = 0; uint32_t ph = 0; int it = 0;
    const uint32_t sbo = (uint32_t)q.pitch * 128u;
    const uint64_t b_desc0 = make_desc(w_smem, 1024u);
    const uint32_t b_tap_step = (n_local * 128u) >> 4;                  // descriptor address units (16 B)
    const uint32_t a_dy = (uint32_t)q.pitch * 8u;                       // one halo row, in 16-B units (one pixel = 8 units)
    const bool unrolled = q.kh == 3 && q.kw == 3;                       // the hot case: nine taps from uniform registers
    for (int tile = first_tile; tile < tile_end; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * q.n_tile);
      if (unrolled) {
        // 3 x 3: the 36 MMAs of a K chunk fully unrolled (the issuing thread spends ~10 instructions per MMA in the rolled
        // loop below, which is what bounds the narrow-N wide-input layers); wide inputs repeat the block per 64-channel chunk
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          // start-address field arithmetic: tap / k offsets never carry out of the 14-bit field (smem < 256 KB)
          const uint64_t a_desc0 = make_desc(a_smem + (uint32_t)s * q.stage_bytes, sbo);
          const uint64_t b_kc = b_desc0 + (uint64_t)((uint32_t)kc * (w_kc_bytes >> 4));
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t ad = a_desc0 + (uint64_t)((tap / 3) * a_dy + (tap % 3) * 8u);
              const uint64_t bd = b_kc + (uint64_t)tap * b_tap_step;
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_tmem, ad + 2u * k, bd + 2u * k, (tap | k) != 0 ? 1u : (uint32_t)(kc != 0));
            }
            commit(empty_bar(s));            // smem stage reusable once these MMAs have read it (both CTAs in k2)
            if (kc == n_kc - 1) commit(tfull_bar(acc));          // accumulator complete
          }
          __syncwarp();
          if (++s == q.stages) { s = 0; ph ^= 1u; }
        }
      } else {
        // general odd filter and / or wide input: the accumulator collects every (K chunk, tap) product; each 64-channel
        // chunk has its own halo stage and its own block of resident weight rows
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t a_desc0 = make_desc(a_smem + (uint32_t)s * q.stage_bytes, sbo);
          const uint64_t b_kc = b_desc0 + (uint64_t)((uint32_t)kc * (w_kc_bytes >> 4));
          if (elect_one()) {
            int tap = 0;
            for (int ty_ = 0; ty_ < q.kh; ++ty_)
              for (int tx_ = 0; tx_ < q.kw; ++tx_, ++tap) {
                const uint64_t ad = a_desc0 + (uint64_t)((uint32_t)ty_ * a_dy + (uint32_t)tx_ * 8u);
                const uint64_t bd = b_kc + (uint64_t)tap * b_tap_step;
#pragma unroll
                for (int k = 0; k < 4; ++k) mma(d_tmem, ad + 2u * k, bd + 2u * k, (uint32_t)((kc | tap | k) != 0));
              }
            commit(empty_bar(s));
            if (kc == n_kc - 1) commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++s == q.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    }
  } else if ((kSpec == 1 || kSpec == 2 || kSpec == 6 || kSpec == 7) && !(k2 && kSpec == 6) && q.tma_epi) {
    // ===================== epilogue, TMA flavour (16-bit y, d2s = 1, 64- or 128-channel chunks) =====================
    // Warp (quadrant, column half) owns 4 tile rows x 8 pixels x ncols channels.  Its staging rows are laid out exactly as
    // the {ncols, 8, 4, 1} TMA box with the 32/64/128-byte swizzle of the row size, so one elected lane moves the whole
    // block with one bulk-tensor copy: residual (hi, lo) loads one tile ahead into alternating buffers, signalled on a
    // per-warp mbarrier, and output stores tracked by the bulk async-group of that lane.  Out-of-image pixels are
    // clipped (stores) or zero-filled (loads) by the TMA unit - no per-lane address arithmetic or bounds predicates.
    constexpr bool P8 = kSpec == 6;
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    // a warp's columns are handled in sub-blocks of at most 64 channels (one 128-byte swizzle row): one for the 64- and
    // 128-channel chunks, two for the 256-channel chunks of CTA pairs
    const int wcols = q.n_tile >> 1, ncols = wcols > 64 ? 64 : wcols, n_sb = wcols / ncols;
    // depth_to_space(2): a sub-block's columns are one (i, j) sub-pixel; the output map is 5-D {c, j, x, i, (b, y)}
    const int d2s = p.d2s;
    const uint32_t hb = (uint32_t)ncols * 2u, lb = (uint32_t)ncols;          // row bytes: 16-bit rows, e5m2 rows
    const uint32_t h_sh = hb == 128u ? 0u : 1u, h_mask = (hb >> 4) - 1u;      // swizzle: chunk ^= (row >> sh) & mask
    const uint32_t l_sh = lb == 64u ? 1u : 2u, l_mask = (lb >> 4) - 1u;
    const uint32_t buf_bytes = 32u * hb + (P8 ? 32u * lb : 0u);
    const uint32_t my_epi = epi_smem + (uint32_t)ew * q.epi_warp_bytes;
    const uint32_t row = (uint32_t)lane;
    const uint32_t h_row = row * hb, h_x = (row >> h_sh) & h_mask;
    const uint32_t l_row = 32u * hb + row * lb, l_x = (row >> l_sh) & l_mask;
    const uint32_t my_rbar = rbar0 + 16u * (uint32_t)ew;
    const float alpha = p.alpha, beta1 = p.beta1, beta2 = p.beta2;
    const bool bf = p.y_dtype == SRB_BF16;
    auto release_tmem = [&](int acc) {                  // (k2: the MMA warp of the pair's leader waits for both CTAs)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (k2 && rank != 0) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
    };
    auto coords = [&](int tile, int& b, int& y0, int& x0) {
      const int tl = q.reverse ? q.total_tiles - 1 - tile : tile;
      b = fast_div(tl, tiles_per_img, q.magic_tpi);
      const int rr_ = tl - b * tiles_per_img;
      const int ty = fast_div(rr_, q.tiles_x, q.magic_tx);
      y0 = ty * kTileH + quad * 4;                      // first image row of this warp's quadrant
      x0 = (rr_ - ty * q.tiles_x) * kTileW;
    };
    const int c_res0 = co_base + half * wcols;          // (pair8: d2s = 1, one sub-block)
    auto load_res = [&](int tile, uint32_t nb) {        // lane 0: residual (hi, lo) of `tile` into buffer nb
      int b, y0, x0;
      coords(tile, b, y0, x0);
      const uint32_t buf = my_epi + nb * buf_bytes, bar = my_rbar + 8u * nb;
      mbar_expect_tx(bar, 32u * (hb + lb));
      tma_load_4d(buf, &em.r1, bar, c_res0, x0, y0, b);
      tma_load_4d(buf + 32u * hb, &em.r2, bar, c_res0, x0, y0, b);
    };
    int it = 0;
    if (P8 && lane == 0 && first_tile < q.total_tiles) load_res(first_tile, 0u);
    for (int tile = first_tile; tile < tile_end; tile += tile_step, ++it) {
      const bool live = tile < q.total_tiles;           // (k2: the pair's odd CTA may hold a dummy tile past the end)
      int b, y0, x0;
      coords(live ? tile : 0, b, y0, x0);
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const uint32_t buf = my_epi + (P8 ? (uint32_t)(it & 1) * buf_bytes : 0u);
      if (lane == 0) {
        bulk_wait_read0();                              // this lane's earlier stores have finished reading the staging rows
        if (P8 && tile + tile_step < q.total_tiles) load_res(tile + tile_step, (uint32_t)((it + 1) & 1));
      }
      __syncwarp();
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      if (P8) mbar_wait(my_rbar + 8u * (uint32_t)(it & 1), acc_ph);
      for (int sb = 0; sb < n_sb; ++sb) {
        const int col0 = half * wcols + sb * ncols;     // first accumulator column of this sub-block
        int c_out0 = co_base + col0, sub_i = 0, sub_j = 0;
        if (d2s > 1) { const int qq = c_out0 / p.c_post; c_out0 -= qq * p.c_post; sub_i = qq / d2s; sub_j = qq - sub_i * d2s; }
        if (sb > 0) {                                   // the staging rows are reused: the previous sub-block's store must have read them
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * q.n_tile + col0);
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 16) {
          uint32_t rr[16];
          float4 bv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bias_s + col0 + c0 + 4 * j);
          __syncwarp();
          tmem_ld16(t_row + (uint32_t)c0, rr);
          tmem_ld_wait();
          if (sb == n_sb - 1 && c0 + 16 >= ncols) release_tmem(acc);
          const float bb[16] = {bv[0].x, bv[0].y, bv[0].z, bv[0].w, bv[1].x, bv[1].y, bv[1].z, bv[1].w,
                                bv[2].x, bv[2].y, bv[2].z, bv[2].w, bv[3].x, bv[3].y, bv[3].z, bv[3].w};
          const uint32_t c16 = (uint32_t)c0 >> 4;
          const uint32_t a_h0 = buf + h_row + (((2u * c16) ^ h_x) << 4), a_h1 = buf + h_row + (((2u * c16 + 1u) ^ h_x) << 4);
          uint32_t oh[8];
          if (!P8) {
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v0 = __uint_as_float(rr[2 * i]) + bb[2 * i], v1 = __uint_as_float(rr[2 * i + 1]) + bb[2 * i + 1];
              if (kSpec == 2) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              if (kSpec == 7) {                             // PReLU / leaky: max(v, 0) + slope * min(v, 0)
                const float2 sl = *reinterpret_cast<const float2*>(slope_s + col0 + c0 + 2 * i);
                v0 = fmaf(sl.x, fminf(v0, 0.f), fmaxf(v0, 0.f)); v1 = fmaf(sl.y, fminf(v1, 0.f), fmaxf(v1, 0.f));
              }
              oh[i] = bf ? pack2(v0, v1, SRB_BF16) : pack2(v0, v1, SRB_F16);
            }
          } else {
            // pair8 trunk: v = alpha * (acc + bias) + beta1 * hi + beta2 * lo;  y = round16(v), y2 = e5m2(v - y)
            const uint32_t a_l = buf + l_row + ((c16 ^ l_x) << 4);
            const uint4 uh0 = lds128(a_h0), uh1 = lds128(a_h1), ul = lds128(a_l);
            const uint32_t wh[8] = {uh0.x, uh0.y, uh0.z, uh0.w, uh1.x, uh1.y, uh1.z, uh1.w};
            const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
            uint32_t ol[4];
            auto half_block = [&](auto is_bf) {
              constexpr bool kBf = decltype(is_bf)::value;
  #pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                float lo4[4], er[4];
                e5m2x4_to_float4(wl[i4], lo4);
  #pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int e0 = 4 * i4 + 2 * hh;
                  float2 fh;
                  if (kBf) fh = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wh[2 * i4 + hh]));
                  else fh = __half22float2(*reinterpret_cast<const __half2*>(&wh[2 * i4 + hh]));
                  const float a0 = (__uint_as_float(rr[e0]) + bb[e0]) * alpha, a1 = (__uint_as_float(rr[e0 + 1]) + bb[e0 + 1]) * alpha;
                  const float v0 = fmaf(beta2, lo4[2 * hh], fmaf(beta1, fh.x, a0));
                  const float v1 = fmaf(beta2, lo4[2 * hh + 1], fmaf(beta1, fh.y, a1));
                  const uint32_t pk = pack2(v0, v1, kBf ? SRB_BF16 : SRB_F16);
                  oh[2 * i4 + hh] = pk;
                  float2 back;
                  if (kBf) back = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk));
                  else back = __half22float2(*reinterpret_cast<const __half2*>(&pk));
                  er[2 * hh] = v0 - back.x; er[2 * hh + 1] = v1 - back.y;
                }
                ol[i4] = float4_to_e5m2x4(er[0], er[1], er[2], er[3]);
              }
            };
            if (bf) half_block(std::true_type{}); else half_block(std::false_type{});
            sts128(a_l, make_uint4(ol[0], ol[1], ol[2], ol[3]));
          }
          sts128(a_h0, make_uint4(oh[0], oh[1], oh[2], oh[3]));
          sts128(a_h1, make_uint4(oh[4], oh[5], oh[6], oh[7]));
        }
        fence_proxy_async_smem();                       // generic-proxy writes of the rows -> visible to the TMA unit
        __syncwarp();
        if (lane == 0 && live) {
          if (d2s > 1) tma_store_5d(&em.y, buf, c_out0, sub_j, x0, sub_i, b * p.H + y0);
          else tma_store_4d(&em.y, buf, c_out0, x0, y0, b);
          if (P8 && p.y2) tma_store_4d(&em.y2, buf + 32u * hb, c_out0, x0, y0, b);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait0();                        // all stores complete before the CTA's shared memory goes away
  } else {
    // ===================== epilogue: TMEM -> registers -> (smem transpose) -> global =====================
    // 8 warps: warp w may only touch TMEM lanes [32 * (w % 4), +32); the two warps of a lane quadrant split the
    // accumulator columns (output channels) in halves when the chunk has >= 32 channels.
    const int ew = warp - 2;
    const int quad = warp & 3;
    // narrow chunks (n_tile = 16) are not split: the two warp sets take alternate tiles instead (set h owns TMEM
    // accumulator h), which doubles the per-tile latency budget of the epilogue
    const int split = q.n_tile >= 32 ? 2 : 1;
    const int half = ew >> 2;
    const bool alt = q.n_tile < 32 && q.epi_mode != 1;  // (the staged epilogue keeps its per-warp tile sequence)
    const bool active = alt || half < split;
    const int ncols = q.n_tile / split;                 // channels this warp handles per pixel
    const int col0 = half * ncols * (split - 1);        // first accumulator column of this warp (0 when not split)
    const int m = quad * 32 + lane;                     // GEMM row = pixel within the tile
    // depth_to_space constants of this CTA's channel chunk (vector path: the chunk maps to one (i, j) sub-pixel)
    // (per warp: with 128-channel chunks the two column halves belong to different sub-pixels)
    const int r = p.d2s;
    const int c_first = co_base + col0;                 // first conv-domain channel of this warp
    int qi = 0, qj = 0, c_out0 = c_first;
    if (r > 1) { const int qq = c_first / p.c_post; c_out0 = c_first - qq * p.c_post; qi = qq / r; qj = qq - qi * r; }
    const size_t OW = (size_t)p.W * r, OH = (size_t)p.H * r;
    const uint32_t pitch_y = (uint32_t)(r * OW);        // output pixels between consecutive tile rows
    auto tile_pixel = [&](int b, int y0, int x0) -> size_t {
      return ((size_t)b * OH + (size_t)y0 * r + qi) * OW + (size_t)x0 * r + qj;
    };
    auto pix_off = [&](int mm) -> uint32_t { return (uint32_t)(mm >> 3) * pitch_y + (uint32_t)(mm & 7) * (uint32_t)r; };
    // per-warp staging: F = fp32 rows (residual prefetch / fp32 output), H = 16-bit output rows; XOR-swizzled 16-B chunks
    const uint32_t f_rb = (uint32_t)ncols * 4u, h_rb = (uint32_t)ncols * 2u;
    const uint32_t f_cpr = f_rb >> 4, h_cpr = h_rb >> 4;
    const uint32_t f_swz = f_cpr > 8 ? 7u : f_cpr - 1u, h_swz = h_cpr > 8 ? 7u : h_cpr - 1u;
    const uint32_t f_lg = 31u - (uint32_t)__clz((int)f_cpr), h_lg = 31u - (uint32_t)__clz((int)h_cpr);   // chunks per row are powers of 2
    const uint32_t my_epi = epi_smem + (uint32_t)ew * q.epi_warp_bytes;   // (n_tile = 16 staged layers: one buffer per warp, both sets)
    const uint32_t h_buf = my_epi + (uint32_t)q.f_bufs * 32u * f_rb;
    const int f_dst = q.f_dst, h_dst = q.h_dst;
    const int h_dtype = h_dst == 1 ? p.y_dtype : p.y2_dtype;
    const uint32_t my_row_sw = (uint32_t)lane;          // this lane's staging row

    constexpr bool G = kSpec == 0;
    const bool act_relu = G ? p.act == SRB_ACT_RELU : kSpec == 2;
    const bool act_other = G && p.act != SRB_ACT_NONE && p.act != SRB_ACT_RELU;
    const bool pair = G ? q.res_prefetch == 2 : kSpec == 5;     // F rows = [16-bit hi half | 16-bit lo half]
    const bool use_alpha = G ? p.alpha != 1.f : kSpec >= 3;
    const bool has_res1 = G ? p.res1 != nullptr : kSpec >= 3;
    const bool res_pref = G ? q.res_prefetch != 0 : kSpec >= 3;
    const bool has_res2 = G && p.res2 != nullptr && !pair;
    const bool do_clip = G && p.clip01;
    const bool f_on = G ? f_dst != 0 : kSpec == 3;
    const bool h_on = G ? h_dst != 0 : (kSpec != 5 && kSpec != 6);
    const int epi_mode = G ? q.epi_mode : 1;
    // cooperative (coalesced) move of 32 staged rows: lane -> (row within a group of 32/cpr rows, 16-B chunk)
    auto prefetch_res = [&](int tile, int fb) {
      const int tl = q.reverse ? q.total_tiles - 1 - tile : tile;
      const int b = fast_div(tl, tiles_per_img, q.magic_tpi), rr_ = tl - b * tiles_per_img;
      const int ty_ = fast_div(rr_, q.tiles_x, q.magic_tx);
      const int y0 = ty_ * kTileH, x0 = (rr_ - ty_ * q.tiles_x) * kTileW;
      const bool full = (y0 + kTileH <= p.H) && (x0 + kTileW <= p.W);
      const size_t tp = tile_pixel(b, y0, x0);
      const uint32_t buf = my_epi + (uint32_t)fb * 32u * f_rb;
      const uint32_t rows_per_it = 32u >> f_lg;
      const uint32_t ch = (uint32_t)lane & (f_cpr - 1u), rsub = (uint32_t)lane >> f_lg;
      // fp32 residual: one tensor fills the row; (hi, lo) pair: res1 fills the first half of the chunks, res2 the second
      const uint32_t hc = f_cpr >> 1;
      const bool second = pair && ch >= hc;
      const uint8_t* src0;
      uint32_t pix_bytes;
      if (!pair) {
        src0 = reinterpret_cast<const uint8_t*>(p.res1) + (tp * (size_t)p.res1_cstride + c_out0) * 4u + ch * 16u;
        pix_bytes = (uint32_t)p.res1_cstride * 4u;
      } else if (!second) {
        src0 = reinterpret_cast<const uint8_t*>(p.res1) + (tp * (size_t)p.res1_cstride + c_out0) * 2u + ch * 16u;
        pix_bytes = (uint32_t)p.res1_cstride * 2u;
      } else {
        src0 = reinterpret_cast<const uint8_t*>(p.res2) + (tp * (size_t)p.res2_cstride + c_out0) * 2u + (ch - hc) * 16u;
        pix_bytes = (uint32_t)p.res2_cstride * 2u;
      }
#pragma unroll 4
      for (uint32_t row = rsub; row < 32u; row += rows_per_it) {
        const int mm = quad * 32 + (int)row;
        if (full || (y0 + (mm >> 3) < p.H && x0 + (mm & 7) < p.W))
          cp_async16(buf + row * f_rb + ((ch ^ (row & f_swz)) << 4), src0 + (size_t)pix_off(mm) * pix_bytes);
      }
    };
    auto copy_out = [&](uint32_t buf, uint32_t rb, uint32_t cpr, uint32_t lg, uint32_t swz, void* dst, int cstride, int coffset,
                        uint32_t esize, size_t tpix, bool full, int y0, int x0) {
      uint8_t* dst0 = reinterpret_cast<uint8_t*>(dst) + (tpix * (size_t)cstride + (size_t)(coffset + c_out0)) * esize;
      const uint32_t pix_bytes = (uint32_t)cstride * esize;
      const uint32_t rows_per_it = 32u >> lg;
      const uint32_t ch = (uint32_t)lane & (cpr - 1u), rsub = (uint32_t)lane >> lg;
#pragma unroll 4
      for (uint32_t row = rsub; row < 32u; row += rows_per_it) {
        const int mm = quad * 32 + (int)row;
        if (full || (y0 + (mm >> 3) < p.H && x0 + (mm & 7) < p.W)) {
          const uint4 v = lds128(buf + row * rb + ((ch ^ (row & swz)) << 4));
          *reinterpret_cast<uint4*>(dst0 + (size_t)pix_off(mm) * pix_bytes + ch * 16u) = v;
        }
      }
    };

    // depth_to_space-to-image epilogue (epi_mode 4): float offset of every 4-column group inside its pixel's r x r block
    // (sub-row i = channel / (r * c_post), then the position inside the sub-row), and whether the point function is bias-only
    uint32_t epi4_off[8] = {};
    const bool epi4_plain = p.act == SRB_ACT_NONE && p.alpha == 1.f && !p.clip01;
    if (epi_mode == 4) {
      const int rg = r * p.c_post;                      // floats per output sub-row (a multiple of 4)
      const uint32_t sub_row = (uint32_t)OW * (uint32_t)p.c_post;   // floats between output rows
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const int cb = c_first + 4 * c4, i = cb / rg;
        epi4_off[c4] = (uint32_t)i * sub_row + (uint32_t)(cb - i * rg);
      }
    }
    int it = 0;
    const bool staged = epi_mode == 1 && active;
    const bool prefetch = staged && res_pref;
    if (prefetch) {
      if (first_tile < q.total_tiles) prefetch_res(first_tile, 0);
      cp_async_commit();
    }
    auto release_tmem = [&](int acc) {                  // accumulator drained -> the (leader's) MMA warp may overwrite it
      tc_fence_before();
      __syncwarp();                                     // every lane's tcgen05.wait::ld has retired
      if (lane == 0) { if (k2 && rank != 0) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
    };
    for (int tile = first_tile; tile < tile_end; tile += tile_step, ++it) {
      const bool live = tile < q.total_tiles;           // (k2: the pair's odd CTA may hold a dummy tile)
      const int tl = !live ? 0 : q.reverse ? q.total_tiles - 1 - tile : tile;   // (never negative: multiply-high division)
      const int b = fast_div(tl, tiles_per_img, q.magic_tpi), rr_ = tl - b * tiles_per_img;
      const int ty_ = fast_div(rr_, q.tiles_x, q.magic_tx);
      const int y0 = live ? ty_ * kTileH : p.H, x0 = (rr_ - ty_ * q.tiles_x) * kTileW;
      const bool full = (y0 + kTileH <= p.H) && (x0 + kTileW <= p.W);
      const bool valid = full || (y0 + (m >> 3) < p.H && x0 + (m & 7) < p.W);
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      if (alt && acc != half) continue;                 // the other warp set owns this tile
      uint32_t f_buf = my_epi;
      if (prefetch) {
        if (tile + tile_step < q.total_tiles) prefetch_res(tile + tile_step, (it + 1) & 1);
        cp_async_commit();
        cp_async_wait1();                               // this tile's residual rows have landed
        __syncwarp();
        f_buf = my_epi + (uint32_t)(it & 1) * 32u * f_rb;
      }
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      if (!active) {                                    // chunk too narrow to split: this warp only keeps the protocol
        release_tmem(acc);
        continue;
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * q.n_tile + col0);
      const size_t tpix = tile_pixel(b, y0, x0);
      const size_t my_pix = tpix + pix_off(m);
      if (epi_mode == 4) {
        // depth_to_space straight to a few-channel image (ESPCN: 48 -> 4x4 x RGB): the r*c_post channels of one
        // sub-row i are r*c_post contiguous floats of output row oy*r + i, and neighbouring lanes continue the run
        uint32_t ra[16], rb[16];
        __syncwarp();
        tmem_ld16(t_row, ra);
        tmem_ld16(t_row + 16u, rb);                     // (ncols <= 32; columns past ncols belong to the peer warp, unused)
        tmem_ld_wait();
        release_tmem(acc);
        if (valid) {
          const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
          float* drow = reinterpret_cast<float*>(p.y) + (((size_t)b * OH + (size_t)oy * r) * OW + (size_t)ox * r) * p.c_post;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {              // static register indices: groups of 4 columns
            if (4 * c4 < ncols) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + col0 + 4 * c4);
              float v[4];
              v[0] = __uint_as_float(c4 < 4 ? ra[(4 * c4) & 15] : rb[(4 * c4) & 15]) + bv.x;
              v[1] = __uint_as_float(c4 < 4 ? ra[(4 * c4 + 1) & 15] : rb[(4 * c4 + 1) & 15]) + bv.y;
              v[2] = __uint_as_float(c4 < 4 ? ra[(4 * c4 + 2) & 15] : rb[(4 * c4 + 2) & 15]) + bv.z;
              v[3] = __uint_as_float(c4 < 4 ? ra[(4 * c4 + 3) & 15] : rb[(4 * c4 + 3) & 15]) + bv.w;
              if (!epi4_plain) {                        // (ESPCN's last layer is linear: warp-uniform skip)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float t = v[u];
                  if (p.act == SRB_ACT_RELU) t = fmaxf(t, 0.f);
                  else if (p.act != SRB_ACT_NONE) t = act_generic(t, p.act, p.act_slope);
                  t *= p.alpha;
                  if (p.clip01) t = fminf(fmaxf(t, 0.f), 1.f);
                  v[u] = t;
                }
              }
              *reinterpret_cast<float4*>(drow + epi4_off[c4]) = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t rr[16];
        float4 bv[4];                                   // bias of these 16 channels, fetched ahead of the TMEM read
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bias_s + col0 + c0 + 4 * j);
        __syncwarp();                                   // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld16(t_row + (uint32_t)c0, rr);
        tmem_ld_wait();
        if (c0 + 16 >= ncols) release_tmem(acc);        // last read of this accumulator: hand it back to the MMA warp
        if (epi_mode == 1) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int cc = c0 + g * 8;                  // channel offset inside this warp's column range
            float v[8];
            const float4 b0 = bv[2 * g], b1 = bv[2 * g + 1];
            v[0] = __uint_as_float(rr[g * 8 + 0]) + b0.x; v[1] = __uint_as_float(rr[g * 8 + 1]) + b0.y;
            v[2] = __uint_as_float(rr[g * 8 + 2]) + b0.z; v[3] = __uint_as_float(rr[g * 8 + 3]) + b0.w;
            v[4] = __uint_as_float(rr[g * 8 + 4]) + b1.x; v[5] = __uint_as_float(rr[g * 8 + 5]) + b1.y;
            v[6] = __uint_as_float(rr[g * 8 + 6]) + b1.z; v[7] = __uint_as_float(rr[g * 8 + 7]) + b1.w;
            if (act_relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            } else if (act_other) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float slope = (p.act == SRB_ACT_PRELU) ? __ldg(p.prelu + c_out0 + cc + e) : p.act_slope;
                v[e] = act_generic(v[e], p.act, slope);
              }
            }
            if (use_alpha) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] *= p.alpha;
            }
            const uint32_t fc = (uint32_t)cc >> 2;      // first of the two fp32 chunks of these 8 channels
            const uint32_t f_a0 = f_buf + my_row_sw * f_rb + (((fc) ^ (my_row_sw & f_swz)) << 4);
            const uint32_t f_a1 = f_buf + my_row_sw * f_rb + (((fc + 1u) ^ (my_row_sw & f_swz)) << 4);
            if (pair) {
              // compensated 16-bit trunk: residual = hi + lo (two halves of the staged row); outputs y = round16(v)
              // and y2 = v - y go back into the same two chunks
              const uint32_t hcn = f_cpr >> 1, gc = (uint32_t)cc >> 3;
              const uint32_t a_hi = f_buf + my_row_sw * f_rb + (((gc) ^ (my_row_sw & f_swz)) << 4);
              const uint32_t a_lo = f_buf + my_row_sw * f_rb + (((gc + hcn) ^ (my_row_sw & f_swz)) << 4);
              const uint4 uh = lds128(a_hi), ul = lds128(a_lo);
              const uint32_t wh[4] = {uh.x, uh.y, uh.z, uh.w}, wl[4] = {ul.x, ul.y, ul.z, ul.w};
              uint32_t oh[4], ol[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float2 fh, fl;
                if (p.y_dtype == SRB_BF16) {
                  fh = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wh[i]));
                  fl = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wl[i]));
                } else {
                  fh = __half22float2(*reinterpret_cast<const __half2*>(&wh[i]));
                  fl = __half22float2(*reinterpret_cast<const __half2*>(&wl[i]));
                }
                const float v0 = v[2 * i] + fmaf(p.beta2, fl.x, p.beta1 * fh.x);
                const float v1 = v[2 * i + 1] + fmaf(p.beta2, fl.y, p.beta1 * fh.y);
                oh[i] = pack2(v0, v1, p.y_dtype);
                float2 back;
                if (p.y_dtype == SRB_BF16) back = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&oh[i]));
                else back = __half22float2(*reinterpret_cast<const __half2*>(&oh[i]));
                ol[i] = pack2(v0 - back.x, v1 - back.y, p.y_dtype);
              }
              sts128(a_hi, make_uint4(oh[0], oh[1], oh[2], oh[3]));
              if (p.y2) sts128(a_lo, make_uint4(ol[0], ol[1], ol[2], ol[3]));
            } else if (has_res1) {
              float rv[8];
              if (res_pref) {
                const uint4 u0 = lds128(f_a0), u1 = lds128(f_a1);
                rv[0] = __uint_as_float(u0.x); rv[1] = __uint_as_float(u0.y); rv[2] = __uint_as_float(u0.z); rv[3] = __uint_as_float(u0.w);
                rv[4] = __uint_as_float(u1.x); rv[5] = __uint_as_float(u1.y); rv[6] = __uint_as_float(u1.z); rv[7] = __uint_as_float(u1.w);
              } else if (valid) {
                load8(p.res1, p.res1_dtype, my_pix * p.res1_cstride + c_out0 + cc, rv);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) rv[e] = 0.f;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaf(p.beta1, rv[e], v[e]);
            }
            if (has_res2 && valid) {
              float rv[8];
              load8(p.res2, p.res2_dtype, my_pix * p.res2_cstride + c_out0 + cc, rv);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaf(p.beta2, rv[e], v[e]);
            }
            if (do_clip) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fminf(fmaxf(v[e], 0.f), 1.f);
            }
            if (f_on && !pair) {
              sts128(f_a0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
              sts128(f_a1, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
            }
            if (h_on && !pair) {
              const uint32_t hc = (uint32_t)cc >> 3;
              uint4 pk;
              if (h_dtype == SRB_BF16)
                pk = make_uint4(pack2(v[0], v[1], SRB_BF16), pack2(v[2], v[3], SRB_BF16), pack2(v[4], v[5], SRB_BF16),
                                pack2(v[6], v[7], SRB_BF16));
              else
                pk = make_uint4(pack2(v[0], v[1], SRB_F16), pack2(v[2], v[3], SRB_F16), pack2(v[4], v[5], SRB_F16),
                                pack2(v[6], v[7], SRB_F16));
              sts128(h_buf + my_row_sw * h_rb + ((hc ^ (my_row_sw & h_swz)) << 4), pk);
            }
          }
        } else if (epi_mode == 3) {
          // 8 or 16 output channels written as a 16-bit channel slice (the growth convs of the ESRGAN dense blocks,
          // ESRGAN_model.py:230-246): bias + none / ReLU / leaky, one or two 16-byte stores per pixel
          if (valid && c0 == 0) {
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float b_ = e < 4 ? (&bv[0].x)[e] : e < 8 ? (&bv[1].x)[e - 4] : e < 12 ? (&bv[2].x)[e - 8] : (&bv[3].x)[e - 12];
              float t = __uint_as_float(rr[e]) + b_;
              if (p.act == SRB_ACT_RELU) t = fmaxf(t, 0.f);
              else if (p.act == SRB_ACT_LEAKY) t = fmaf(p.act_slope, fminf(t, 0.f), fmaxf(t, 0.f));
              v[e] = t;
            }
            const int dt_ = p.y_dtype;
            if (dt_ == SRB_F32) {                       // (the f / g projections of SelfAttention: float32 for the softmax)
              float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + my_pix * p.y_cstride + p.y_coffset);
              d[0] = make_float4(v[0], v[1], v[2], v[3]);
              d[1] = make_float4(v[4], v[5], v[6], v[7]);
              if (p.cout > 8) { d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]); }
            } else {
              uint16_t* d = reinterpret_cast<uint16_t*>(p.y) + my_pix * p.y_cstride + p.y_coffset;
              *reinterpret_cast<uint4*>(d) = make_uint4(pack2(v[0], v[1], dt_), pack2(v[2], v[3], dt_), pack2(v[4], v[5], dt_), pack2(v[6], v[7], dt_));
              if (p.cout > 8)
                *reinterpret_cast<uint4*>(d + 8) = make_uint4(pack2(v[8], v[9], dt_), pack2(v[10], v[11], dt_), pack2(v[12], v[13], dt_),
                                                              pack2(v[14], v[15], dt_));
            }
          }
        } else if (epi_mode == 2) {
          // few output channels (the RGB tail layers): <= 4 channels per pixel, no residual, no shuffle
          if (valid && c0 == 0) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[e] = __uint_as_float(rr[e]) + bias_s[e];
              if (p.act == SRB_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
              else if (p.act != SRB_ACT_NONE)
                v[e] = act_generic(v[e], p.act, (p.act == SRB_ACT_PRELU && e < p.cout) ? __ldg(p.prelu + e) : p.act_slope);
              v[e] *= p.alpha;
              if (p.clip01) v[e] = fminf(fmaxf(v[e], 0.f), 1.f);
            }
            const size_t o = my_pix * p.y_cstride + p.y_coffset;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < p.cout) store_elem(p.y, p.y_dtype, o + e, v[e]);
          }
        } else if (valid) {
          const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int co = co_base + col0 + c0 + e;
            if (co < p.cout) epilogue_store_generic(p, b, oy, ox, co, __uint_as_float(rr[e]));   // adds p.bias[co] itself
          }
        }
      }
      if (epi_mode == 1 && pair) {
        __syncwarp();
        // chunks [0, cpr/2) of every staged row -> y, chunks [cpr/2, cpr) -> y2 (both 16-bit)
        const uint32_t hcn = f_cpr >> 1;
        const uint32_t rows_per_it = 32u >> f_lg;
        const uint32_t ch = (uint32_t)lane & (f_cpr - 1u), rsub = (uint32_t)lane >> f_lg;
        const bool second = ch >= hcn;
        uint8_t* dst0 = second
            ? reinterpret_cast<uint8_t*>(p.y2) + (tpix * (size_t)p.y2_cstride + (size_t)c_out0) * 2u + (ch - hcn) * 16u
            : reinterpret_cast<uint8_t*>(p.y) + (tpix * (size_t)p.y_cstride + (size_t)(p.y_coffset + c_out0)) * 2u + ch * 16u;
        const uint32_t pix_bytes = (uint32_t)(second ? p.y2_cstride : p.y_cstride) * 2u;
#pragma unroll 4
        for (uint32_t row = rsub; row < 32u; row += rows_per_it) {
          const int mm = quad * 32 + (int)row;
          if ((!second || p.y2) && (full || (y0 + (mm >> 3) < p.H && x0 + (mm & 7) < p.W)))
            *reinterpret_cast<uint4*>(dst0 + (size_t)pix_off(mm) * pix_bytes) = lds128(f_buf + row * f_rb + ((ch ^ (row & f_swz)) << 4));
        }
        __syncwarp();
      } else if (epi_mode == 1) {
        __syncwarp();                                   // rows written by their owner lanes -> read by all lanes
        if (f_on)
          copy_out(f_buf, f_rb, f_cpr, f_lg, f_swz, f_dst == 1 ? p.y : p.y2, f_dst == 1 ? p.y_cstride : p.y2_cstride,
                   f_dst == 1 ? p.y_coffset : 0, 4u, tpix, full, y0, x0);
        if (h_on)
          copy_out(h_buf, h_rb, h_cpr, h_lg, h_swz, h_dst == 1 ? p.y : p.y2, h_dst == 1 ? p.y_cstride : p.y2_cstride,
                   h_dst == 1 ? p.y_coffset : 0, 2u, tpix, full, y0, x0);
        __syncwarp();                                   // staging rows are free for the next tile
      }
    }
  }

  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();    // (k2: the peer may still be reading this CTA's smem / signalling its barriers)
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (k2) tmem_dealloc_2sm(tmem_base, q.tmem_cols); else tmem_dealloc(tmem_base, q.tmem_cols);
  }
}



// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static unsigned g_launch_parity = 0;
int tc_next_reverse() { return (int)(g_launch_parity++ & 1u); }
static int g_two_cta = 0;       // srb_conv_tc_set_cta_pairs: launch every eligible cin = 64 layer as CTA pairs (default: only
                                // where pairs pay - 256-channel chunks and inputs wider than 64 channels)

EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

bool conv_tc_eligible(const ConvParams& p) {
  if (!p.w_tc || p.w_tc_cin < 64 || p.kh > 9 || p.kw > 9) return false;
  const int n_kc = p.w_tc_cin / 64;
  // the smallest weight chunk (16 rows; 32 rows as 16 per CTA of a pair when the input is wider than 64 channels) for every tap
  // and K chunk + two halo stages must fit next to barriers and staging (9x9, cin = 64: 166 KB + 48 KB)
  if ((size_t)p.kh * p.kw * 16 * 128 * n_kc + (size_t)(n_kc > 1 ? 2 : 1) * (kTileH + p.kh - 1) * (kTileW + p.kw - 1) * 128 > 216 * 1024)
    return false;
  if (p.x_dtype != SRB_BF16 && p.x_dtype != SRB_F16) return false;
  if ((p.x_cstride % 8) || (p.x_coffset % 8) || !aligned16(p.x)) return false;
  if (p.W < 1 || p.H < 1) return false;
  return true;
}

// 8 consecutive channels per access: one 16-byte vector for 16-bit types, two for fp32
static bool vec_ok_for(const void* ptr, int dtype, int cstride, int coffset) {
  const int per16 = dtype == SRB_F32 ? 4 : dtype == SRB_F8E5M2 ? 16 : 8;
  return aligned16(ptr) && (cstride % per16 == 0) && (coffset % per16 == 0);
}


// ---- wide-tile fold kernel: eligibility and launch ----
int conv_fold_mode(const ConvParams& p);
int conv_fold_launch(const ConvParams& p, int mode, cudaStream_t stream);

int conv_tc_launch(const ConvParams& p, cudaStream_t stream) {
  { const int fm = conv_fold_mode(p); if (fm >= 0) return conv_fold_launch(p, fm, stream); }
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("conv(tcgen05): cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }

  const int rows = p.w_tc_rows;                       // cout padded to 16
  int dev = 0, max_smem = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const long total = (long)p.B * ((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
  SRB_REQUIRE(total < (1L << 30), "conv(tcgen05): too many tiles");
  auto dt16 = [](int d) { return d == SRB_BF16 || d == SRB_F16; };
  const int two_cta = g_two_cta;

  // channel-chunk width per CTA: the widest of {256 (pairs), 128, 64, 32 (wide inputs), rows, 16} that divides the padded cout
  // and whose weights (all K chunks), >= 2 A stages and epilogue staging fit shared memory
  TcParams q{};
  size_t smem = 0;
  // 256-channel chunks exist only as CTA pairs (each CTA holds 128 weight rows) with the TMA epilogue, for plain 16-bit
  // layers with >= 256 output channels (the up-sampling convs): halves the B-operand shared-memory traffic per CTA that
  // the epilogue competes with.
  const bool plain_layer = !p.res1 && !p.res2 && !p.y2 && !p.clip01 && p.alpha == 1.f && dt16(p.y_dtype) &&
                           (p.act == SRB_ACT_NONE || p.act == SRB_ACT_RELU || p.act == SRB_ACT_LEAKY || (p.act == SRB_ACT_PRELU && p.prelu));
  const int n_kc = p.w_tc_cin / 64;                   // 64-channel K chunks (weights of every chunk stay resident)
  // candidates, widest first.  Wide inputs (n_kc > 1) also try every width as a CTA pair: each CTA then keeps half of the
  // weight rows, so the pair's MMA is twice as wide for the same resident bytes (N is what pays for the A-operand fetch)
  const int cand[6] = {256, 128, 64, 32, rows < 64 ? rows : 16, 16};
  bool found = false;
  for (int ci = 0; ci < 12 && !found; ++ci) {
    const int nt = cand[ci >> 1];
    const bool try_pair = (ci & 1) == 0;               // even: as a CTA pair, odd: single CTAs
    if (nt > 256 || rows % nt) continue;
    if (nt == 32 && n_kc == 1) continue;               // (cin = 64 keeps its round-1 choices: 128 / 64 / rows / 16)
    if (nt == 256 && !try_pair) continue;
    if (nt == 256 && !(plain_layer && (n_kc > 1 || (p.kh == 3 && p.kw == 3)))) continue;
    if (try_pair && nt != 256 && !(two_cta || n_kc > 1)) continue;
    if (!try_pair && nt != 256 && two_cta && n_kc == 1 && nt >= 32) continue;   // (set_cta_pairs(1): pairs wherever possible)
    if (try_pair && nt < 32) continue;
    q = TcParams{};
    q.n_kc = n_kc;
    q.kh = p.kh; q.kw = p.kw;
    q.halo_rows = kTileH + p.kh - 1;
    q.n_tile = nt;
    q.n_chunks = rows / nt;
    q.w_rows = rows;
    q.tiles_x = (p.W + kTileW - 1) / kTileW;
    q.tiles_y = (p.H + kTileH - 1) / kTileH;
    q.total_tiles = (int)total;
    q.pitch = kTileW + p.kw - 1;
    q.load_bytes = (uint32_t)(q.halo_rows * q.pitch * 128);
    q.stage_bytes = (q.load_bytes + 1023u) & ~1023u;
    q.tmem_cols = 32;
    while (q.tmem_cols < (uint32_t)(2 * nt)) q.tmem_cols <<= 1;
    const uint32_t fmt = p.x_dtype == SRB_BF16 ? 1u : 0u;   // F16F32Format: 0 = F16, 1 = BF16
    const bool k2c = try_pair;
    q.two_cta = k2c ? 1 : 0;
    q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)((k2c ? 2 * kTileM : kTileM) >> 4) << 24);
    // staged vector epilogue: one fp32 and/or one 16-bit destination; a warp's columns map to one d2s sub-pixel
    const int warp_cols = nt >= 32 ? nt / 2 : nt;
    const int sb_cols = warp_cols > 64 ? 64 : warp_cols;      // TMA epilogue sub-block width (256-channel chunks: two per warp)
    bool vec = (nt == 16 || nt == 32 || nt == 64 || nt == 128 || nt == 256) && (p.cout % nt == 0) &&
               (p.d2s == 1 || p.c_post % (nt == 256 ? sb_cols : warp_cols) == 0) && vec_ok_for(p.y, p.y_dtype, p.y_cstride, p.y_coffset) &&
               (!p.y2 || vec_ok_for(p.y2, p.y2_dtype, p.y2_cstride, 0)) &&
               (!p.res1 || vec_ok_for(p.res1, p.res1_dtype, p.res1_cstride, 0)) &&
               (!p.res2 || vec_ok_for(p.res2, p.res2_dtype, p.res2_cstride, 0));
    // compensated 16-bit trunk: (hi, lo) residual pair in, y + rounding error out, all the same 16-bit dtype
    // (y2 may be absent: the layer that leaves the trunk only needs the rounded sum)
    const bool pair = dt16(p.y_dtype) && (!p.y2 || (p.y2_mode == 1 && p.y2_dtype == p.y_dtype)) && p.res1 && p.res2 &&
                      p.res1_dtype == p.y_dtype && p.res2_dtype == p.y_dtype && p.act == SRB_ACT_NONE && !p.clip01;
    // pair8 trunk: 16-bit hi + e5m2 lo residual pair in, 16-bit y + e5m2 rounding error out
    const bool pair8 = dt16(p.y_dtype) && p.res1 && p.res2 && p.res1_dtype == p.y_dtype && p.res2_dtype == SRB_F8E5M2 &&
                       (!p.y2 || (p.y2_mode == 1 && p.y2_dtype == SRB_F8E5M2)) && p.act == SRB_ACT_NONE && !p.clip01 && p.d2s == 1;
    const bool any_f8 = p.y_dtype == SRB_F8E5M2 || (p.y2 && p.y2_dtype == SRB_F8E5M2) || (p.res1 && p.res1_dtype == SRB_F8E5M2) ||
                        (p.res2 && p.res2_dtype == SRB_F8E5M2);
    if (any_f8 && !pair8) vec = false;                 // (the generic scalar epilogue handles e5m2 element by element)
    if (p.y_dtype == SRB_U8) vec = false;              // (quantised output: element-wise stores)
    if (p.y2 && !pair && !pair8 && (p.y2_mode != 0 || ((p.y_dtype == SRB_F32) == (p.y2_dtype == SRB_F32)))) vec = false;   // one of each kind
    q.epi_mode = vec ? 1 : 0;
    if (!vec && p.cout <= 4 && p.d2s == 1 && !p.res1 && !p.res2 && !p.y2) q.epi_mode = 2;
    // 8 / 16 output channels as a 16-bit or float32 channel slice (ESRGAN growth convs at growth 8 / 16, SelfAttention f / g)
    if (!vec && nt == 16 && (p.cout == 8 || p.cout == 16) && p.d2s == 1 && !p.res1 && !p.res2 && !p.y2 && !p.clip01 && p.alpha == 1.f &&
        (dt16(p.y_dtype) || p.y_dtype == SRB_F32) && (p.act == SRB_ACT_NONE || p.act == SRB_ACT_RELU || p.act == SRB_ACT_LEAKY) &&
        p.y_cstride % 8 == 0 && p.y_coffset % 8 == 0 && aligned16(p.y))
      q.epi_mode = 3;
    // depth_to_space onto a few-channel fp32 image: every warp's column range is whole sub-rows of r*c_post floats
    if (!vec && p.d2s > 1 && p.y_dtype == SRB_F32 && !p.res1 && !p.res2 && !p.y2 && p.act != SRB_ACT_PRELU &&
        p.y_cstride == p.c_post && p.y_coffset == 0 && (p.d2s * p.c_post) % 4 == 0 && aligned16(p.y) &&
        nt <= 64 && p.cout == nt && warp_cols % (p.d2s * p.c_post) == 0)
      q.epi_mode = 4;
    if (vec) {
      q.f_dst = p.y_dtype == SRB_F32 ? 1 : (p.y2 && p.y2_dtype == SRB_F32 ? 2 : 0);
      q.h_dst = dt16(p.y_dtype) ? 1 : (p.y2 && dt16(p.y2_dtype) ? 2 : 0);
      q.res_prefetch = (p.res1 && p.res1_dtype == SRB_F32) ? 1 : 0;
      if (pair) { q.res_prefetch = 2; q.f_dst = 0; q.h_dst = 0; }
      if (pair8) { q.res_prefetch = 3; q.f_dst = 0; q.h_dst = 0; }
      q.f_bufs = q.res_prefetch ? 2 : (q.f_dst ? 1 : 0);
      q.epi_warp_bytes = (uint32_t)(q.f_bufs * 32 * warp_cols * 4 + (q.h_dst ? 32 * warp_cols * 2 : 0));
      if (pair8) q.epi_warp_bytes = (uint32_t)(2 * 32 * warp_cols * 3);
      // TMA epilogue: 16-bit y, 64- / 128- / 256-channel chunks, the plain / ReLU / PReLU / pair8 layer types
      const bool slope_act = p.act == SRB_ACT_LEAKY || (p.act == SRB_ACT_PRELU && p.prelu);
      const bool plain16 = !p.res1 && !p.res2 && !p.y2 && !p.clip01 && p.alpha == 1.f && q.h_dst == 1 && q.f_dst == 0 &&
                           (p.act == SRB_ACT_NONE || p.act == SRB_ACT_RELU || slope_act);
      // (depth_to_space: plain layers whose warp columns are whole sub-pixels, images made of whole 16-row tiles because
      //  the 5-D output map merges the image and row dimensions)
      const bool d2s_ok = p.d2s == 1 || (plain16 && p.c_post % sb_cols == 0 && p.H % kTileH == 0 && p.y_coffset == 0 && p.y_cstride == p.c_post);
      if (d2s_ok && (nt == 64 || nt == 128 || nt == 256) && dt16(p.y_dtype) && ((plain16) || (pair8 && !k2c))) {
        q.tma_epi = 1;
        q.epi_warp_bytes = (uint32_t)(pair8 ? 2 * 32 * warp_cols * 3 : 32 * sb_cols * 2);   // multiples of 1,024 bytes
      }
      // (the pair8 trunk epilogue only exists in its TMA form; any other chunk width takes the generic scalar epilogue)
      if (pair8 && !q.tma_epi) { q.epi_mode = 0; q.res_prefetch = 0; q.f_bufs = 0; q.epi_warp_bytes = 0; }
    }
    if (nt == 256 && !q.tma_epi) continue;                    // (256-channel chunks only exist with the TMA epilogue)
    const size_t w_bytes = ((size_t)p.kh * p.kw * (q.two_cta ? nt / 2 : nt) * 128 * n_kc + 1023) & ~(size_t)1023;
    const size_t tail_bytes = (2 * kMaxStages + 6) * 8 + 2 * (size_t)nt * sizeof(float) + 2 * kEpiWarps * 8;
    auto smem_need = [&](int st) { return 1024 + w_bytes + (size_t)st * q.stage_bytes + (size_t)kEpiWarps * q.epi_warp_bytes + tail_bytes; };
    q.stages = kMaxStages;
    while (q.stages > 1 && smem_need(q.stages) > (size_t)max_smem) --q.stages;
    smem = smem_need(q.stages);
    found = smem <= (size_t)max_smem && (q.stages >= 2 || (n_kc == 1 && nt == 16));
  }
  if (!found) { set_error("conv(tcgen05): weights + one pipeline stage + epilogue staging do not fit shared memory"); return SRB_E_UNSUPPORTED; }

  // ---- tensor maps ----
  const CUtensorMapDataType tdt = p.x_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmx, tmw;
  {
    // channel extent: the K chunks the weights cover, but never past the end of the pixel (the buffer may be narrower than
    // the padded cin: TMA zero-fills what lies beyond it; channels between cin and the extent meet zero weight columns)
    const int c_avail = p.x_cstride - p.x_coffset;
    const cuuint64_t c_ext = (cuuint64_t)(p.w_tc_cin < c_avail ? p.w_tc_cin : c_avail);
    const cuuint64_t dims[4] = {c_ext, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.x_cstride * 2, (cuuint64_t)p.W * p.x_cstride * 2,
                                   (cuuint64_t)p.H * p.W * p.x_cstride * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)q.pitch, (cuuint32_t)q.halo_rows, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    void* gptr = (void*)((const uint16_t*)p.x + p.x_coffset);
    CUresult r = encode(&tmx, tdt, 4, gptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv(tcgen05): cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SRB_E_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.w_tc_cin, (cuuint64_t)(p.kh * p.kw) * rows};
    const cuuint64_t strides[1] = {(cuuint64_t)p.w_tc_cin * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)(q.two_cta ? q.n_tile / 2 : q.n_tile)};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmw, tdt, 2, (void*)p.w_tc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv(tcgen05): cuTensorMapEncodeTiled(w) failed with %d", (int)r); return SRB_E_CUDA; }
  }

  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (q.tma_epi) {
    const int wc = q.n_tile / 2 > 64 ? 64 : q.n_tile / 2;
    auto encode_epi = [&](CUtensorMap* m, const void* ptr, int coffset, int cstride, bool f8) -> bool {
      const size_t es = f8 ? 1 : 2;
      const cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
      const cuuint64_t strides[3] = {(cuuint64_t)cstride * es, (cuuint64_t)p.W * cstride * es, (cuuint64_t)p.H * p.W * cstride * es};
      const cuuint32_t box[4] = {(cuuint32_t)wc, (cuuint32_t)kTileW, 4, 1};
      const cuuint32_t es1[4] = {1, 1, 1, 1};
      const size_t row_bytes = (size_t)wc * es;
      const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                    : CU_TENSOR_MAP_SWIZZLE_32B;
      const CUtensorMapDataType dt = f8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                        : (p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
      void* g = (void*)((const uint8_t*)ptr + (size_t)coffset * es);
      return encode(m, dt, 4, g, dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok;
    if (p.d2s > 1) {
      // out[b, y*r + i, x*r + j, c]: dims {c, j, x, i, (b, y)}; a warp stores {wc, 1, 8, 1, 4}
      const cuuint64_t r = (cuuint64_t)p.d2s, pix = (cuuint64_t)p.c_post * 2, orow = (cuuint64_t)p.W * r * pix;
      const cuuint64_t dims[5] = {(cuuint64_t)p.c_post, r, (cuuint64_t)p.W, r, (cuuint64_t)p.H * p.B};
      const cuuint64_t strides[4] = {pix, r * pix, orow, r * orow};
      const cuuint32_t box[5] = {(cuuint32_t)wc, 1, (cuuint32_t)kTileW, 1, 4};
      const cuuint32_t es1[5] = {1, 1, 1, 1, 1};
      const CUtensorMapSwizzle sw = wc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      ok = encode(&em.y, p.y_dtype == SRB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, p.y, dims, strides,
                  box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    } else {
      ok = encode_epi(&em.y, p.y, p.y_coffset, p.y_cstride, false);
    }
    if (q.res_prefetch == 3) {
      ok = ok && encode_epi(&em.r1, p.res1, 0, p.res1_cstride, false) && encode_epi(&em.r2, p.res2, 0, p.res2_cstride, true);
      if (p.y2) ok = ok && encode_epi(&em.y2, p.y2, 0, p.y2_cstride, true);
    }
    if (!ok) { set_error("conv(tcgen05): cuTensorMapEncodeTiled(epilogue) failed"); return SRB_E_CUDA; }
  }

  int spec = 0;
  if (q.epi_mode == 1 && q.res_prefetch == 3) {
    spec = 6;
  } else if (q.epi_mode == 1 && q.res_prefetch == 2) {
    spec = 5;
  } else if (q.epi_mode == 1 && !p.res2 && !p.clip01) {
    if (!p.res1 && p.alpha == 1.f && q.f_dst == 0 && q.h_dst == 1 && !p.y2) {
      if (p.act == SRB_ACT_NONE) spec = 1;
      else if (p.act == SRB_ACT_RELU) spec = 2;
      else if (q.tma_epi && (p.act == SRB_ACT_LEAKY || p.act == SRB_ACT_PRELU)) spec = 7;   // (TMA epilogue only)
    } else if (q.res_prefetch && p.act == SRB_ACT_NONE) {
      if (q.f_dst == 1 && q.h_dst == 2) spec = 3;
      else if (q.f_dst == 0 && q.h_dst == 1 && !p.y2) spec = 4;
    }
  }
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const EpiMaps, const TcParams, const ConvParams);
  static const KernelFn kernels[2][8] = {
      {conv3x3_tc_kernel<0, false>, conv3x3_tc_kernel<1, false>, conv3x3_tc_kernel<2, false>, conv3x3_tc_kernel<3, false>,
       conv3x3_tc_kernel<4, false>, conv3x3_tc_kernel<5, false>, conv3x3_tc_kernel<6, false>, conv3x3_tc_kernel<7, false>},
      {conv3x3_tc_kernel<0, true>, conv3x3_tc_kernel<1, true>, conv3x3_tc_kernel<2, true>, conv3x3_tc_kernel<3, true>,
       conv3x3_tc_kernel<4, true>, conv3x3_tc_kernel<5, true>, conv3x3_tc_kernel<6, true>, conv3x3_tc_kernel<7, true>}};
  static size_t configured[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
  const int k2 = q.two_cta;
  q.reverse = k2 ? 0 : tc_next_reverse();
  q.magic_tpi = div_magic((long)q.total_tiles + 2, q.tiles_x * q.tiles_y);
  q.magic_tx = div_magic((long)q.tiles_x * q.tiles_y, q.tiles_x);
  if (smem > configured[k2][spec]) {
    SRB_CUDA(cudaFuncSetAttribute(kernels[k2][spec], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[k2][spec] = smem;
  }
  const int unit = q.n_chunks * (k2 ? 2 : 1);          // CTAs per (chunk-complete) scheduling unit
  int grid = sm_count();
  grid -= grid % unit;
  if (grid < unit) grid = unit;
  const long work = ((long)(k2 ? (q.total_tiles + 1) / 2 : q.total_tiles)) * unit;
  if ((long)grid > work) grid = (int)work;
  SRB_CUDA(tc_launch(kernels[k2][spec], grid, kThreads, smem, stream, k2 != 0, tmx, tmw, em, q, p));
  return launch_check("conv3x3_tc_kernel");
}

}  // namespace srb

extern "C" int srb_conv_tc_set_cta_pairs(int on) {
  const int prev = srb::g_two_cta ? 1 : 0;
  if (on >= 0) srb::g_two_cta = on ? 1 : 0;
  return prev;
}
