// Error channel, device probe and the small elementwise / layout helpers of libsrb200.
#include "common.cuh"
#include <string.h>

namespace srb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return SRB_E_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// dst = src * scale + shift with dtype conversion; 4 elements per thread when aligned
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n, float scale, float shift, int relu) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = fmaf(load_as_float<S>(src + i), scale, shift);
    if (relu) v = fmaxf(v, 0.f);
    store_from_float<D>(dst + i, v);
  }
}

template <typename T>
__global__ void maxpool2x2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2;
  const size_t total = (size_t)B * OH * OW * C;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const T* p = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    const float v = fmaxf(fmaxf(load_as_float<T>(p), load_as_float<T>(p + C)),
                          fmaxf(load_as_float<T>(p + (size_t)W * C), load_as_float<T>(p + (size_t)W * C + C)));
    store_from_float<T>(y + i, v);
  }
}

// 16-bit activations with C % 8 == 0 (every VGG16 block): one output row per block, eight channels (16 bytes) per thread -
// four coalesced 16-byte loads, packed max, one 16-byte store.  (The element-per-thread kernel above ran at a third of the
// HBM rate and was 18 % of the VGG16 classifier's time, profiles/r02_vgg_launches.md.)
template <typename T2>
__device__ __forceinline__ uint4 max4x8(uint4 a, uint4 b, uint4 c, uint4 d) {
  uint4 r;
  const T2* pa = reinterpret_cast<const T2*>(&a); const T2* pb = reinterpret_cast<const T2*>(&b);
  const T2* pc = reinterpret_cast<const T2*>(&c); const T2* pd = reinterpret_cast<const T2*>(&d);
  T2* pr = reinterpret_cast<T2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(__hmax2(pa[i], pb[i]), __hmax2(pc[i], pd[i]));
  return r;
}

template <typename T2>
__global__ void __launch_bounds__(256)
maxpool2x2_vec8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C8) {
  const int OH = H / 2, OW = W / 2;
  const int b = blockIdx.x / OH, oy = blockIdx.x % OH;
  const uint4* r0 = x + ((size_t)b * H + 2 * oy) * W * C8;
  const uint4* r1 = r0 + (size_t)W * C8;
  uint4* yo = y + (size_t)blockIdx.x * OW * C8;
  for (int j = threadIdx.x; j < OW * C8; j += blockDim.x) {
    const int ox = j / C8, c8 = j - ox * C8;
    const int i = 2 * ox * C8 + c8;
    yo[j] = max4x8<T2>(__ldg(r0 + i), __ldg(r0 + i + C8), __ldg(r1 + i), __ldg(r1 + i + C8));
  }
}

// GlobalAveragePooling2D -> Dense(hidden, relu) -> Dense(classes, softmax); one block per image.
template <typename T>
__global__ void __launch_bounds__(256)
gap_dense_softmax_kernel(const T* __restrict__ x, int HW, int C, const float* __restrict__ w1,
                         const float* __restrict__ b1, int hidden, const float* __restrict__ w2,
                         const float* __restrict__ b2, int classes, float* __restrict__ probs) {
  extern __shared__ float sm[];
  float* gap = sm;              // [C]
  float* hid = sm + C;          // [hidden]
  float* logit = hid + hidden;  // [classes]
  const T* xi = x + (size_t)blockIdx.x * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += load_as_float<T>(xi + (size_t)p * C + c);
    gap[c] = s / (float)HW;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < hidden; h += blockDim.x) {
    float s = b1[h];
    for (int c = 0; c < C; ++c) s = fmaf(gap[c], __ldg(w1 + (size_t)c * hidden + h), s);
    hid[h] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < classes; k += blockDim.x) {
    float s = b2[k];
    for (int h = 0; h < hidden; ++h) s = fmaf(hid[h], __ldg(w2 + (size_t)h * classes + k), s);
    logit[k] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int k = 0; k < classes; ++k) m = fmaxf(m, logit[k]);
    float z = 0.f;
    for (int k = 0; k < classes; ++k) z += expf(logit[k] - m);
    for (int k = 0; k < classes; ++k) probs[(size_t)blockIdx.x * classes + k] = expf(logit[k] - m) / z;
  }
}

// SelfAttention core (ESRGAN_model.py:58-66): o[q] = sum_k softmax_k(g[q] . f[k]) h[k], exact fp32 on the CUDA cores.
// One THREAD per query (a block = 128 consecutive queries of one image): the query vector, the running maximum / sum of the
// online softmax and the DV output accumulators live in registers, so the key loop needs no cross-lane reduction at all.
// Keys and values are staged through shared memory in tiles of 64 and read back as broadcasts (every lane of a warp reads
// the same key: one wavefront per 16-byte load).  Per (query, key) pair: DK + DV FMAs - the value update as packed fp32x2 -
// plus one exp; keys are scored in groups of eight so that the accumulators are rescaled at most once per group.
template <int DK, int DV>
__global__ void __launch_bounds__(128)
self_attention_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ h,
                      int HW, float* __restrict__ o) {
  constexpr int KT = 64, KG = 8;
  __shared__ __align__(16) float sf[KT * DK];
  __shared__ __align__(16) float sh[KT * DV];
  const int b = blockIdx.y;
  const int q = blockIdx.x * 128 + threadIdx.x;
  const float* fb = f + (size_t)b * HW * DK;
  const float* hb = h + (size_t)b * HW * DV;
  float gq[DK];
  {
    const float* gp = g + ((size_t)b * HW + (q < HW ? q : HW - 1)) * DK;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gp + d));
      gq[d] = v.x; gq[d + 1] = v.y; gq[d + 2] = v.z; gq[d + 3] = v.w;
    }
  }
  float m = -INFINITY, z = 0.f;
  float2 acc[DV / 2];
#pragma unroll
  for (int i = 0; i < DV / 2; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int k0 = 0; k0 < HW; k0 += KT) {
    const int nk = min(KT, HW - k0);
    __syncthreads();
    // the tile's keys and values are contiguous in global memory: 16-byte coalesced copies (rows past HW: zeros)
    for (int i = threadIdx.x; i < KT * DK / 4; i += 128)
      reinterpret_cast<float4*>(sf)[i] = (4 * i < nk * DK) ? __ldg(reinterpret_cast<const float4*>(fb + (size_t)k0 * DK) + i)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < KT * DV / 4; i += 128)
      reinterpret_cast<float4*>(sh)[i] = (4 * i < nk * DV) ? __ldg(reinterpret_cast<const float4*>(hb + (size_t)k0 * DV) + i)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
#pragma unroll 1
    for (int kg = 0; kg < nk; kg += KG) {
      float sc[KG];
      float gm = -INFINITY;
#pragma unroll
      for (int j = 0; j < KG; ++j) {
        const float4* kp = reinterpret_cast<const float4*>(sf + (kg + j) * DK);
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DK; d += 4) {
          const float4 v = kp[d / 4];
          a = fmaf(gq[d], v.x, a); a = fmaf(gq[d + 1], v.y, a); a = fmaf(gq[d + 2], v.z, a); a = fmaf(gq[d + 3], v.w, a);
        }
        sc[j] = (kg + j < nk) ? a : -INFINITY;
        gm = fmaxf(gm, sc[j]);
      }
      if (gm > m) {                                     // new running maximum: rescale what has been accumulated so far
        const float corr = __expf(m - gm);              // (m = -inf on the first group: corr = 0, and z, acc are 0 anyway)
        z *= corr;
        const float2 c2 = make_float2(corr, corr);
#pragma unroll
        for (int i = 0; i < DV / 2; ++i) acc[i] = __fmul2_rn(acc[i], c2);
        m = gm;
      }
#pragma unroll
      for (int j = 0; j < KG; ++j) {
        const float p = __expf(sc[j] - m);              // (exp(-inf) = 0 for the keys past HW)
        z += p;
        const float2 p2 = make_float2(p, p);
        const float4* vp = reinterpret_cast<const float4*>(sh + (kg + j) * DV);
#pragma unroll
        for (int i = 0; i < DV / 4; ++i) {
          const float4 v = vp[i];
          acc[2 * i] = __ffma2_rn(p2, make_float2(v.x, v.y), acc[2 * i]);
          acc[2 * i + 1] = __ffma2_rn(p2, make_float2(v.z, v.w), acc[2 * i + 1]);
        }
      }
    }
  }
  if (q < HW) {
    const float inv = 1.f / z;
    float4* op = reinterpret_cast<float4*>(o + ((size_t)b * HW + q) * DV);
#pragma unroll
    for (int i = 0; i < DV / 4; ++i)
      op[i] = make_float4(acc[2 * i].x * inv, acc[2 * i].y * inv, acc[2 * i + 1].x * inv, acc[2 * i + 1].y * inv);
  }
}

static int grid_for(size_t n, int threads) {
  const size_t want = (n + threads - 1) / threads;
  const size_t cap = (size_t)sm_count() * 16;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace srb

using namespace srb;

extern "C" const char* srb_last_error(void) { return g_err; }
extern "C" int srb_version(void) { return 100; }

extern "C" int srb_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  int n = 0, ma = 0, mi = 0;
  SRB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  SRB_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  SRB_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  if (ma != 10) { set_error("libsrb200 is built for sm_100a only; device is sm_%d%d", ma, mi); return SRB_E_UNSUPPORTED; }
  return SRB_OK;
}

static int cast_impl(const void* src, int sdt, void* dst, int ddt, size_t n, float scale, float shift, int relu, srb_stream_t stream_);

extern "C" int srb_cast(const void* src, int sdt, void* dst, int ddt, size_t n, float scale, float shift,
                        srb_stream_t stream_) {
  return cast_impl(src, sdt, dst, ddt, n, scale, shift, 0, stream_);
}

extern "C" int srb_cast_relu(const void* src, int sdt, void* dst, int ddt, size_t n, float scale, float shift,
                             srb_stream_t stream_) {
  return cast_impl(src, sdt, dst, ddt, n, scale, shift, 1, stream_);
}

static int cast_impl(const void* src, int sdt, void* dst, int ddt, size_t n, float scale, float shift, int relu, srb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRB_REQUIRE(src && dst, "cast: null pointer");
  if (n == 0) return SRB_OK;
  const int g = grid_for(n, 256);
#define SRB_CAST(S, D) cast_kernel<S, D><<<g, 256, 0, stream>>>((const S*)src, (D*)dst, n, scale, shift, relu)
#define SRB_CAST_FROM(S)                                                                       \
  if (ddt == SRB_F32) SRB_CAST(S, float);                                                      \
  else if (ddt == SRB_BF16) SRB_CAST(S, __nv_bfloat16);                                        \
  else if (ddt == SRB_F16) SRB_CAST(S, __half);                                                \
  else if (ddt == SRB_U8) SRB_CAST(S, uint8_t);                                                \
  else { set_error("cast: unsupported destination dtype %d", ddt); return SRB_E_UNSUPPORTED; }
  if (sdt == SRB_F32) { SRB_CAST_FROM(float) }
  else if (sdt == SRB_BF16) { SRB_CAST_FROM(__nv_bfloat16) }
  else if (sdt == SRB_F16) { SRB_CAST_FROM(__half) }
  else if (sdt == SRB_U8) { SRB_CAST_FROM(uint8_t) }
  else { set_error("cast: unsupported source dtype %d", sdt); return SRB_E_UNSUPPORTED; }
#undef SRB_CAST_FROM
#undef SRB_CAST
  return launch_check("cast_kernel");
}

extern "C" int srb_maxpool2x2_nhwc(const void* x, int dtype, int batch, int height, int width, int channels, void* y,
                                   srb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRB_REQUIRE(x && y, "maxpool: null pointer");
  SRB_REQUIRE(batch >= 0 && height >= 2 && width >= 2 && channels > 0, "maxpool: bad geometry");
  const size_t total = (size_t)batch * (height / 2) * (width / 2) * channels;
  if (total == 0) return SRB_OK;
  if ((dtype == SRB_F16 || dtype == SRB_BF16) && channels % 8 == 0 && (long)batch * (height / 2) < (1L << 31) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const int rows = batch * (height / 2), threads = min(256, max(32, ((width / 2) * (channels / 8) + 31) / 32 * 32));
    if (dtype == SRB_F16)
      maxpool2x2_vec8_kernel<__half2><<<rows, threads, 0, stream>>>((const uint4*)x, (uint4*)y, height, width, channels / 8);
    else
      maxpool2x2_vec8_kernel<__nv_bfloat162><<<rows, threads, 0, stream>>>((const uint4*)x, (uint4*)y, height, width, channels / 8);
    return launch_check("maxpool2x2_vec8_kernel");
  }
  const int g = grid_for(total, 256);
  if (dtype == SRB_F32) maxpool2x2_kernel<float><<<g, 256, 0, stream>>>((const float*)x, (float*)y, batch, height, width, channels);
  else if (dtype == SRB_BF16) maxpool2x2_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, batch, height, width, channels);
  else if (dtype == SRB_F16) maxpool2x2_kernel<__half><<<g, 256, 0, stream>>>((const __half*)x, (__half*)y, batch, height, width, channels);
  else { set_error("maxpool: unsupported dtype %d", dtype); return SRB_E_UNSUPPORTED; }
  return launch_check("maxpool2x2_kernel");
}

extern "C" int srb_gap_dense_softmax(const void* x, int dtype, int batch, int hw, int channels,
                                     const float* w1, const float* b1, int hidden, const float* w2, const float* b2,
                                     int classes, float* probs, srb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRB_REQUIRE(x && w1 && b1 && w2 && b2 && probs, "gap_dense_softmax: null pointer");
  SRB_REQUIRE(batch >= 0 && hw > 0 && channels > 0 && hidden > 0 && classes > 0, "gap_dense_softmax: bad geometry");
  if (batch == 0) return SRB_OK;
  const size_t smem = (size_t)(channels + hidden + classes) * sizeof(float);
  SRB_REQUIRE(smem <= 48 * 1024, "gap_dense_softmax: layer too wide");
  if (dtype == SRB_F32)
    gap_dense_softmax_kernel<float><<<batch, 256, smem, stream>>>((const float*)x, hw, channels, w1, b1, hidden, w2, b2, classes, probs);
  else if (dtype == SRB_BF16)
    gap_dense_softmax_kernel<__nv_bfloat16><<<batch, 256, smem, stream>>>((const __nv_bfloat16*)x, hw, channels, w1, b1, hidden, w2, b2, classes, probs);
  else if (dtype == SRB_F16)
    gap_dense_softmax_kernel<__half><<<batch, 256, smem, stream>>>((const __half*)x, hw, channels, w1, b1, hidden, w2, b2, classes, probs);
  else { set_error("gap_dense_softmax: unsupported dtype %d", dtype); return SRB_E_UNSUPPORTED; }
  return launch_check("gap_dense_softmax_kernel");
}

extern "C" int srb_self_attention_f32(const float* f, const float* g, const float* h, int batch, int hw, int dk, int dv,
                                      float* o, srb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRB_REQUIRE(f && g && h && o, "self_attention: null pointer");
  SRB_REQUIRE(batch >= 0 && hw > 0, "self_attention: bad geometry");
  if (batch == 0) return SRB_OK;
  SRB_REQUIRE(batch <= 65535, "self_attention: batch too large for one launch");
  SRB_REQUIRE(((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(h) |
                reinterpret_cast<uintptr_t>(o)) & 15) == 0, "self_attention: tensors must be 16-byte aligned");
  dim3 grid((hw + 127) / 128, batch);
  if (dk == 8 && dv == 32) self_attention_kernel<8, 32><<<grid, 128, 0, stream>>>(f, g, h, hw, o);
  else if (dk == 4 && dv == 16) self_attention_kernel<4, 16><<<grid, 128, 0, stream>>>(f, g, h, hw, o);
  else if (dk == 16 && dv == 64) self_attention_kernel<16, 64><<<grid, 128, 0, stream>>>(f, g, h, hw, o);
  else { set_error("self_attention: unsupported head sizes dk=%d dv=%d (channels must be 32, 64 or 128)", dk, dv); return SRB_E_UNSUPPORTED; }
  return launch_check("self_attention_kernel");
}
