// Device-side tiling for super_resolve_image (loading_methods.py:6-26; EDSR_model.py:201-256;
// SRCNN_model.py:127-188): reflect padding + sliding-window patch extraction in one gather kernel,
// and overlap-add reconstruction as a deterministic gather (each output pixel sums the patches that
// cover it in the reference's loop order, divides by the count, clips to [0,1]).
#include "common.cuh"

namespace srb {

__host__ __device__ inline int pad_amount(int n, int patch, int stride) {
  int p = (n % stride != 0) ? (patch - (n % stride)) % stride : 0;
  const int q = patch - stride;
  return p > q ? p : q;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  return i < n ? i : period - i;
}

__global__ void pad_extract_kernel(const float* __restrict__ img, int H, int W, int C, int P, int S, int nx,
                                   float* __restrict__ patches, size_t total) {
  const size_t stride_t = (size_t)gridDim.x * blockDim.x;
  const int row_elems = P * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
    const int ec = (int)(i % row_elems);
    size_t rest = i / row_elems;
    const int py = (int)(rest % P);
    const int patch = (int)(rest / P);
    const int px = ec / C, c = ec - px * C;
    const int y = reflect_idx((patch / nx) * S + py, H);
    const int x = reflect_idx((patch % nx) * S + px, W);
    patches[i] = __ldg(img + ((size_t)y * W + x) * C + c);
  }
}

__global__ void overlap_add_kernel(const float* __restrict__ patches, int ny, int nx, int PS, int SS, int C,
                                   float* __restrict__ out, int out_h, int out_w) {
  const size_t total = (size_t)out_h * out_w * C;
  const size_t stride_t = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int x = (int)(p % out_w), y = (int)(p / out_w);
    // patches (iy, ix) with iy*SS <= y < iy*SS + PS
    const int iy_hi = min(y / SS, ny - 1), ix_hi = min(x / SS, nx - 1);
    const int iy_lo = max(0, (y - PS + SS) / SS), ix_lo = max(0, (x - PS + SS) / SS);
    float acc = 0.f, cnt = 0.f;
    for (int iy = iy_lo; iy <= iy_hi; ++iy)
      for (int ix = ix_lo; ix <= ix_hi; ++ix) {
        const int ly = y - iy * SS, lx = x - ix * SS;
        if (ly < PS && lx < PS) {
          acc = __fadd_rn(acc, __ldg(patches + (((size_t)(iy * nx + ix) * PS + ly) * PS + lx) * C + c));
          cnt += 1.f;
        }
      }
    float v = cnt != 0.f ? __fdiv_rn(acc, cnt) : 0.f;
    out[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

}  // namespace srb

using namespace srb;

extern "C" int srb_tiling_geometry(int height, int width, int patch, int stride,
                                   int* padded_h, int* padded_w, int* ny, int* nx) {
  SRB_REQUIRE(height > 0 && width > 0 && patch > 0 && stride > 0, "tiling: bad geometry");
  const int ph = height + pad_amount(height, patch, stride), pw = width + pad_amount(width, patch, stride);
  if (padded_h) *padded_h = ph;
  if (padded_w) *padded_w = pw;
  if (ny) *ny = ph >= patch ? (ph - patch) / stride + 1 : 0;
  if (nx) *nx = pw >= patch ? (pw - patch) / stride + 1 : 0;
  return SRB_OK;
}

extern "C" int srb_pad_extract_f32(const float* image, int height, int width, int channels, int patch, int stride,
                                   float* patches, srb_stream_t stream) {
  SRB_REQUIRE(image && patches, "pad_extract: null pointer");
  int ph, pw, ny, nx;
  int rc = srb_tiling_geometry(height, width, patch, stride, &ph, &pw, &ny, &nx);
  if (rc) return rc;
  // np.pad(mode="reflect") with pad >= n needs multiple reflections; the kernel handles any distance
  const size_t total = (size_t)ny * nx * patch * patch * channels;
  if (total == 0) return SRB_OK;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < (size_t)sm_count() * 16 ? (total + threads - 1) / threads
                                                                                      : (size_t)sm_count() * 16);
  pad_extract_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(image, height, width, channels, patch, stride, nx,
                                                                  patches, total);
  return launch_check("pad_extract_kernel");
}

extern "C" int srb_overlap_add_f32(const float* patches, int ny, int nx, int patch_out, int stride_out, int channels,
                                   float* image, int out_h, int out_w, srb_stream_t stream) {
  SRB_REQUIRE(patches && image, "overlap_add: null pointer");
  SRB_REQUIRE(ny > 0 && nx > 0 && patch_out > 0 && stride_out > 0 && channels > 0 && out_h > 0 && out_w > 0,
              "overlap_add: bad geometry");
  const size_t total = (size_t)out_h * out_w * channels;
  const int threads = 256;
  const size_t want = (total + threads - 1) / threads;
  const int blocks = (int)(want < (size_t)sm_count() * 16 ? want : (size_t)sm_count() * 16);
  overlap_add_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(patches, ny, nx, patch_out, stride_out, channels,
                                                                  image, out_h, out_w);
  return launch_check("overlap_add_kernel");
}
