// Fused PSNR + SSIM with both passes of the separable 11-tap Gaussian on the warp-level tensor path (mma.sync m16n8k16,
// fp16 operands, fp32 accumulation) - tf.image.ssim semantics (metrics.py:3-7), wide images.
//
// The CUDA-core kernel (metrics.cu) needs 88 FP32 FMAs per map element and is bound by the FP32 pipe and the shared-memory
// port at ~60 GP/s.  Here a warp evaluates, per step, the four maps (mu_a, mu_b, E[a^2 + b^2], E[ab]) of a 16-row x 22-column
// block of one channel plane as two chained banded-Toeplitz products that never leave its registers:
//
//   pass 1 (vertical):    D1[16 x 32] = Tv[16 x 32] . X[32 rows x 32 cols]     A = Tv (constant), B = image data
//   pass 2 (horizontal):  Out[16 x 24] = D1[16 x 32] . Th[32 x 24]             A = D1 (the m16n8 accumulator layout IS the
//                                                                               m16k16 A layout, two n-blocks per k-block),
//                                                                               B = Th (constant)
//
// Tv[m, k] = h[k - m], Th[k, n] = h[k - n] (VALID correlation), so both Toeplitz operands are eight / eight registers per
// lane for the whole kernel.  fp32-grade maps from fp16 operands: every data operand is split as x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) (22 significant bits, both products accumulate into the same fp32 accumulator) - for the
// image data before pass 1 and for D1 before pass 2.  The window is NOT split: h is the Gaussian rounded to fp16 with six
// taps moved by 1-3 ulp so that the eleven taps sum to exactly 1 (tools/ssim_window16.py: |dSSIM| <= 2.5e-6 against the
// float64 Gaussian on the test images, no normalisation factor, constant images stay exact).  26 HMMAs per map and step,
// 104 per 352 map elements.  SRB_SSIM_TF_EXACT selects the CUDA-core kernels with the float32 Gaussian instead.
//
// A block is one warp per channel; the warps share the raw interleaved rows of their strip, which arrive by TMA (one
// {32 C + 4 floats, 16 rows} box per array and half, zero-filled past the image) into two slots guarded by full / empty
// mbarriers: step s reads half s + 1 while half s + 2 is in flight, and no block-wide barrier is left in the loop.  The B
// fragments a lane builds from the newest half are the k-block 0 operands of its next step, so they are parked in a
// lane-private shared-memory stash: every input row is squared, multiplied and split once.  The row stride (100 words for
// C = 3) makes the fragment loads (rows 2t + {0, 1, 8, 9}, columns g + 8 nb) conflict-free; the TMA box must start on a
// 16-byte boundary of global memory (an unaligned inner coordinate traps as "illegal instruction"), so the strip start is
// rounded down and the fragment offset takes the remainder.
// The squared error for PSNR is taken from the same fragment registers over a non-overlapping ownership partition.
#include "common.cuh"
#include "metrics_mma.cuh"
#include "tc_ptx.cuh"

namespace srb {

namespace {

constexpr int kMW = 22;       // output columns per strip (32 input columns)
constexpr int kMH = 16;       // output rows per step
constexpr int kSlots = 2;     // resident 16-row halves per array (the one being read, the one in flight)

__constant__ float c_h16[11];   // the fp16-exact window (as float)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816_first(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {   // d = a b
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

__device__ __forceinline__ uint32_t pack_h2(float lo_elem, float hi_elem) {
  const __half2 h = __floats2half2_rn(lo_elem, hi_elem);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// v = hi + lo, both fp16 pairs (element with the lower index in the low half); the subtraction is one packed FMA
__device__ __forceinline__ void split2(float2 v, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v.x, v.y);
  const float2 r = __ffma2_rn(__half22float2(h), make_float2(-1.f, -1.f), v);
  const __half2 l = __floats2half2_rn(r.x, r.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP alone (den >= c1 c2 > 0: no range fix-up needed)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float tap(int d) { return (d >= 0 && d <= 10) ? c_h16[d] : 0.f; }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// words per shared-memory row = floats per TMA box row: >= 32 C + 2, 16-byte multiple, and 2 RS = 8 (mod 32) so that the
// fragment loads (rows 2t + ..., columns (g + 8 nb) C) fall into 32 different banks
template <int C> struct Geom;
template <> struct Geom<1> { static constexpr int RS = 36; };
template <> struct Geom<3> { static constexpr int RS = 100; };

}  // namespace

template <int C>
__global__ void __launch_bounds__(32 * C, C == 3 ? 4 : 12)
psnr_ssim_mma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int H, int W,
                     int steps_per_chunk, float c1, float c2, double* __restrict__ acc /* [B][2] = {sse, ssim_sum} */) {
  constexpr int RS = Geom<C>::RS;
  constexpr int kHalf = 16 * RS;                    // words per half
  extern __shared__ __align__(128) float smem[];    // [a | b][kSlots][16][RS] raw rows, then the fragment stash
  float* sa = smem;
  float* sb = smem + kSlots * kHalf;
  __shared__ __align__(8) uint64_t bars[2 * kSlots];   // full[kSlots], empty[kSlots]
  __shared__ float red[2][C];

  const int OW = W - 10, OH = H - 10;
  const int tid = threadIdx.x, ch = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // stash[(m, nb)][thread]: the hi / lo B fragments this lane built from the rows of the newest half - they are the k-block 0
  // operands of the next step (same lane, same registers), so every input row is squared, multiplied and split ONCE
  uint4* stash = reinterpret_cast<uint4*>(smem + 2 * kSlots * kHalf) + tid;
  const int x0 = blockIdx.x * kMW;
  const int ych = blockIdx.y * steps_per_chunk * kMH;                 // first output row of this chunk
  const int nsteps = min(steps_per_chunk, (OH - ych + kMH - 1) / kMH);
  const bool last_strip = x0 + kMW >= OW;
  const bool last_chunk = blockIdx.y == gridDim.y - 1;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kSlots]);
  // TMA wants the box to start on a 16-byte boundary of global memory: the strip's first float (22 C per strip) is rounded
  // down to a multiple of four and the fragment offset takes the remainder (0 or 2; RS has room for it)
  const int xs = (x0 * C) & ~3, xsh = x0 * C - xs;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kSlots; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, C); }
    fence_barrier_init();
  }
  __syncthreads();

  // one TMA box per array and half: rows ych + 16 hf ... + 15, 32 C (+ pad) floats from column x0; rows / columns past
  // the image arrive as zeros.  Half hf lives in slot hf % kSlots; it is read once, by step hf - 1 (half 0: the prologue).
  auto fill = [&](int hf) {
    if (hf <= nsteps) {
      const int slot = hf & (kSlots - 1);
      if (hf >= kSlots) mbar_wait(empty0 + 8 * slot, ((hf / kSlots) - 1) & 1);
      mbar_expect_tx(full0 + 8 * slot, 2 * kHalf * 4);
      tma_load_3d(smem_u32(sa + slot * kHalf), &tm_a, full0 + 8 * slot, xs, ych + 16 * hf, blockIdx.z);
      tma_load_3d(smem_u32(sb + slot * kHalf), &tm_b, full0 + 8 * slot, xs, ych + 16 * hf, blockIdx.z);
    }
  };
  if (tid == 0) { fill(0); fill(1); }

  // constant Toeplitz fragments
  uint32_t tv[2][4];                                // A of pass 1: Tv[m, 16 kb + k] = h[16 kb + k - m]
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    const int k0 = 16 * kb + 2 * t;
    tv[kb][0] = pack_h2(tap(k0 - g), tap(k0 + 1 - g));
    tv[kb][1] = pack_h2(tap(k0 - g - 8), tap(k0 + 1 - g - 8));
    tv[kb][2] = pack_h2(tap(k0 + 8 - g), tap(k0 + 9 - g));
    tv[kb][3] = pack_h2(tap(k0 + 8 - g - 8), tap(k0 + 9 - g - 8));
  }
  uint32_t th[4][2];                                // B of pass 2 for column offsets o = -8, 0, 8, 16: Th[k, n] = h[k + o - n]
#pragma unroll
  for (int oi = 0; oi < 4; ++oi) {
    const int o = 8 * oi - 8;
    th[oi][0] = pack_h2(tap(2 * t + o - g), tap(2 * t + 1 + o - g));
    th[oi][1] = pack_h2(tap(2 * t + 8 + o - g), tap(2 * t + 9 + o - g));
  }

  // validity of this lane's pass-2 outputs (columns 8 nb2 + 2 t + {0, 1}) and ownership of its input columns (g + 8 nb)
  float2 colv[3];
#pragma unroll
  for (int nb2 = 0; nb2 < 3; ++nb2) {
    const int n = 8 * nb2 + 2 * t;
    colv[nb2] = make_float2((n < kMW && x0 + n < OW) ? 1.f : 0.f, (n + 1 < kMW && x0 + n + 1 < OW) ? 1.f : 0.f);
  }
  float colw[4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int n = 8 * nb + g;
    colw[nb] = ((n < kMW || last_strip) && x0 + n < W) ? 1.f : 0.f;
  }

  float2 sse_nb[4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) sse_nb[nb] = make_float2(0.f, 0.f);
  float2 ssim2 = make_float2(0.f, 0.f);
  const float2 c1_2 = make_float2(c1, c1), c2_2 = make_float2(c2, c2), two2 = make_float2(2.f, 2.f), neg1 = make_float2(-1.f, -1.f);
  const int frag_off = 2 * t * RS + xsh + g * C + ch;   // this lane's element of a half: row 2t, column g of its channel

  // rows 2 t + 8 i + {0, 1}, columns g + 8 nb of half hf -> registers; their squared error (every half is loaded exactly once
  // per chunk, the half two chunks share counts for the lower one; zero-filled rows / columns add nothing)
  auto load_half = [&](int hf, float2 (&va)[4][2], float2 (&vb)[4][2]) {
    const int slot = hf & (kSlots - 1);
    mbar_wait(full0 + 8 * slot, (hf / kSlots) & 1);
    const int so = slot * kHalf + frag_off;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        va[nb][i] = make_float2(sa[so + (8 * i) * RS + 8 * nb * C], sa[so + (8 * i + 1) * RS + 8 * nb * C]);
        vb[nb][i] = make_float2(sb[so + (8 * i) * RS + 8 * nb * C], sb[so + (8 * i + 1) * RS + 8 * nb * C]);
      }
    if (hf < nsteps || last_chunk) {
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float2 d = __ffma2_rn(vb[nb][i], neg1, va[nb][i]);
          sse_nb[nb] = __ffma2_rn(d, d, sse_nb[nb]);
        }
    }
  };
  // the four map inputs of one row pair, split into fp16 hi / lo B-fragment registers
  auto frag = [&](int m, const float2 (&va)[4][2], const float2 (&vb)[4][2], int nb, uint4& f) {
    uint32_t bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float2 v;
      if (m == 0) v = va[nb][i];
      else if (m == 1) v = vb[nb][i];
      else if (m == 2) v = __ffma2_rn(va[nb][i], va[nb][i], __fmul2_rn(vb[nb][i], vb[nb][i]));
      else v = __fmul2_rn(va[nb][i], vb[nb][i]);
      split2(v, bh[i], bl[i]);
    }
    f = make_uint4(bh[0], bh[1], bl[0], bl[1]);
  };
  auto release = [&](int hf) {                     // this warp has the half in registers
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * (hf & (kSlots - 1)));
  };

  {                                                // prologue: half 0 -> stash
    float2 va[4][2], vb[4][2];
    load_half(0, va, vb);
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        uint4 f;
        frag(m, va, vb, nb, f);
        stash[(m * 4 + nb) * (32 * C)] = f;
      }
    release(0);
  }

  for (int s = 0; s < nsteps; ++s) {
    if (tid == 0) fill(s + 2);
    const int y0 = ych + kMH * s;

    float d1[4][4][4];
    {
      float2 va[4][2], vb[4][2];
      load_half(s + 1, va, vb);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          uint4* st = stash + (m * 4 + nb) * (32 * C);
          const uint4 f0 = *st;                    // k-block 0: the rows the last step (or the prologue) split
          mma16816_first(d1[m][nb], tv[0], f0.x, f0.y);
          mma16816(d1[m][nb], tv[0], f0.z, f0.w);
          uint4 f1;
          frag(m, va, vb, nb, f1);                 // k-block 1: the new half
          mma16816(d1[m][nb], tv[1], f1.x, f1.y);
          mma16816(d1[m][nb], tv[1], f1.z, f1.w);
          *st = f1;
        }
      release(s + 1);
    }

    // pass 2: the accumulators of n-blocks 2 kb2, 2 kb2 + 1 are the A fragment of k-block kb2
    float out[4][3][4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int kb2 = 0; kb2 < 2; ++kb2) {
        split2(make_float2(d1[m][2 * kb2][0], d1[m][2 * kb2][1]), ah[kb2][0], al[kb2][0]);
        split2(make_float2(d1[m][2 * kb2][2], d1[m][2 * kb2][3]), ah[kb2][1], al[kb2][1]);
        split2(make_float2(d1[m][2 * kb2 + 1][0], d1[m][2 * kb2 + 1][1]), ah[kb2][2], al[kb2][2]);
        split2(make_float2(d1[m][2 * kb2 + 1][2], d1[m][2 * kb2 + 1][3]), ah[kb2][3], al[kb2][3]);
      }
      // (k-block, n-block) pairs inside the band: column offset o = 16 kb2 - 8 nb2 -> th[(o + 8) / 8]
      mma16816_first(out[m][0], ah[0], th[1][0], th[1][1]);  mma16816(out[m][0], al[0], th[1][0], th[1][1]);   // (0, 0): o = 0
      mma16816(out[m][0], ah[1], th[3][0], th[3][1]);        mma16816(out[m][0], al[1], th[3][0], th[3][1]);   // (1, 0): o = 16
      mma16816_first(out[m][1], ah[0], th[0][0], th[0][1]);  mma16816(out[m][1], al[0], th[0][0], th[0][1]);   // (0, 1): o = -8
      mma16816(out[m][1], ah[1], th[2][0], th[2][1]);        mma16816(out[m][1], al[1], th[2][0], th[2][1]);   // (1, 1): o = 8
      mma16816_first(out[m][2], ah[1], th[1][0], th[1][1]);  mma16816(out[m][2], al[1], th[1][0], th[1][1]);   // (1, 2): o = 0
    }

    // point function: out[.][nb2][0..1] = row g, columns 8 nb2 + 2t + {0, 1}; [2..3] = row g + 8
    const bool full_rows = y0 + kMH <= OH;          // warp-uniform
    const float rv0 = (y0 + g < OH) ? 1.f : 0.f, rv1 = (y0 + g + 8 < OH) ? 1.f : 0.f;
#pragma unroll
    for (int nb2 = 0; nb2 < 3; ++nb2)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const float2 ma = make_float2(out[0][nb2][2 * hrow], out[0][nb2][2 * hrow + 1]);
        const float2 mb = make_float2(out[1][nb2][2 * hrow], out[1][nb2][2 * hrow + 1]);
        const float2 es = make_float2(out[2][nb2][2 * hrow], out[2][nb2][2 * hrow + 1]);
        const float2 ep = make_float2(out[3][nb2][2 * hrow], out[3][nb2][2 * hrow + 1]);
        const float2 mm = __fmul2_rn(ma, mb);
        const float2 den0 = __ffma2_rn(ma, ma, __fmul2_rn(mb, mb));
        const float2 ln = __ffma2_rn(two2, mm, c1_2);
        const float2 cn = __ffma2_rn(two2, __ffma2_rn(mm, neg1, ep), c2_2);
        const float2 ld = make_float2(den0.x + c1, den0.y + c1);
        const float2 cdm = __ffma2_rn(den0, neg1, es);               // (es - den0) + c2 in this order: for a == b it is
        const float2 cd = make_float2(cdm.x + c2, cdm.y + c2);        // bit-identical to cn, so SSIM(a, a) = 1 exactly
        const float2 num = __fmul2_rn(ln, cn), den = __fmul2_rn(ld, cd);
        const float rv = hrow ? rv1 : rv0;             // (0 only in the last step of the image)
        const float2 w = full_rows ? colv[nb2] : make_float2(colv[nb2].x * rv, colv[nb2].y * rv);
        ssim2 = __ffma2_rn(make_float2(num.x * rcp_approx(den.x), num.y * rcp_approx(den.y)), w, ssim2);
      }
  }

  float sse = 0.f;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) sse = fmaf(sse_nb[nb].x + sse_nb[nb].y, colw[nb], sse);
  sse = warp_sum(sse);
  const float ssim_sum = warp_sum(ssim2.x + ssim2.y);
  if (lane == 0) { red[0][ch] = sse; red[1][ch] = ssim_sum; }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < C; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
    atomicAdd(&acc[2 * blockIdx.z + 0], s0);
    atomicAdd(&acc[2 * blockIdx.z + 1], s1);
  }
}

static bool g_h16_ready[64] = {};

// fp16 window: the Gaussian (sigma 1.5) rounded to fp16, taps 0..5 moved by (-2, +3, -1, +1, -2, +1) ulp: sums to exactly 1
static int upload_window16() {
  int dev = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && g_h16_ready[dev]) return SRB_OK;
  static const uint16_t bits[6] = {0x1434, 0x1fcb, 0x289b, 0x2f01, 0x32cf, 0x3443};
  float w[11];
  double sum = 0.0;
  for (int i = 0; i < 11; ++i) {
    const uint16_t hb = bits[i <= 5 ? i : 10 - i];
    __half_raw r; r.x = hb;
    w[i] = __half2float(__half(r));
    sum += w[i];
  }
  SRB_REQUIRE(sum == 1.0, "psnr_ssim: fp16 window does not sum to 1 (%.17g)", sum);
  SRB_CUDA(cudaMemcpyToSymbol(c_h16, w, sizeof(w)));
  if (dev >= 0 && dev < 64) g_h16_ready[dev] = true;
  return SRB_OK;
}

// (max_val: a^2 + b^2 and its filtered maps go through fp16 operands - images on a [0, 255] scale would overflow 65504;
//  the reference calls tf.image.ssim with max_val = 1.0 only, metrics.py:6-7)
bool psnr_ssim_mma_eligible(const float* a, const float* b, int height, int width, int channels, float max_val) {
  return (channels == 1 || channels == 3) && width - 10 >= 96 && height >= 11 && max_val <= 64.f &&
         (width * channels) % 4 == 0 && aligned16(a) && aligned16(b) && tc_encode_fn() != nullptr;   // TMA: 16-byte rows
}

template <int C>
static int launch_mma(const float* a, const float* b, int batch, int H, int W, float c1, float c2, double* acc, cudaStream_t stream) {
  const int OW = W - 10, OH = H - 10;
  const int strips = (OW + kMW - 1) / kMW;
  const int steps = (OH + kMH - 1) / kMH;
  int spc = 16;                                     // steps per chunk (10 halo rows per chunk: 4 % at 16)
  const long target = 8L * sm_count();
  while (spc > 2 && (long)strips * ((steps + spc - 1) / spc) * batch < target) spc >>= 1;
  dim3 grid(strips, (steps + spc - 1) / spc, batch);
  SRB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "psnr_ssim: grid too large");
  const size_t smem = (size_t)2 * kSlots * 16 * Geom<C>::RS * sizeof(float) + (size_t)16 * 32 * C * sizeof(uint4);
  EncodeTiledFn encode = tc_encode_fn();
  SRB_REQUIRE(encode != nullptr, "psnr_ssim: cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tma_, tmb_;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)W * C, (cuuint64_t)H, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[3] = {(cuuint32_t)Geom<C>::RS, 16, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const float* ptrs[2] = {a, b};
    CUtensorMap* maps[2] = {&tma_, &tmb_};
    for (int i = 0; i < 2; ++i)
      if (encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptrs[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_error("psnr_ssim: cuTensorMapEncodeTiled failed");
        return SRB_E_CUDA;
      }
  }
  static bool attr_done[64] = {};
  int dev = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  if (!(dev >= 0 && dev < 64 && attr_done[dev])) {
    SRB_CUDA(cudaFuncSetAttribute(psnr_ssim_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  psnr_ssim_mma_kernel<C><<<grid, 32 * C, smem, stream>>>(tma_, tmb_, H, W, spc, c1, c2, acc);
  return launch_check("psnr_ssim_mma_kernel");
}

int run_psnr_ssim_mma(const float* a, const float* b, int batch, int height, int width, int channels, float c1, float c2,
                      double* acc, cudaStream_t stream) {
  int rc = upload_window16();
  if (rc) return rc;
  if (channels == 1) return launch_mma<1>(a, b, batch, height, width, c1, c2, acc, stream);
  return launch_mma<3>(a, b, batch, height, width, c1, c2, acc, stream);
}

}  // namespace srb
