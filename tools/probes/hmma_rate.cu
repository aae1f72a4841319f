// Legacy warp-level tensor path on sm_100a: issue rate of mma.sync.m16n8k16 (fp16 in, fp32 accumulate), alone and
// interleaved with packed FP32 FMAs.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS, int FMAS>
__global__ void __launch_bounds__(256) k(float* out, int iters, uint32_t seed) {
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed * 3, seed * 5};
  float d[CHAINS][4] = {};
  float2 f[8];
  for (int i = 0; i < 8; ++i) f[i] = make_float2(threadIdx.x * 1e-3f, i);
  const float2 m = make_float2(1.0001f, 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      mma16816(d[c], a, b);
#pragma unroll
      for (int q = 0; q < FMAS; ++q) f[(c * FMAS + q) & 7] = __ffma2_rn(f[(c * FMAS + q) & 7], m, m);
    }
  }
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  for (int i = 0; i < 8; ++i) s += f[i].x + f[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS, int FMAS> void run(const char* name, int warps) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096, blocks = 148 * (warps > 8 ? 2 : 1), threads = warps > 8 ? warps * 16 : warps * 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHAINS, FMAS><<<blocks, threads>>>(out, iters, 1); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<CHAINS, FMAS><<<blocks, threads>>>(out, iters, 1);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)blocks * (threads / 32) * iters * CHAINS;
  printf("%-28s warps/SM %2d: %.3f ms  %.1f TFLOP/s (HMMA)  %.2f cycles@1.9GHz per HMMA per SMSP, FFMA2 %.1f TFLOP/s\n", name, warps, ms,
         mmas * 4096 / ms / 1e9, ms * 1e-3 * 1.9e9 / (mmas / (148.0 * 4)), mmas * FMAS * 128 / ms / 1e9);
  cudaFree(out);
}
int main() {
  run<8, 0>("hmma only, 8 chains", 4);
  run<8, 0>("hmma only, 8 chains", 8);
  run<8, 0>("hmma only, 8 chains", 16);
  run<8, 1>("hmma + 1 FFMA2 each", 8);
  run<8, 2>("hmma + 2 FFMA2 each", 8);
  run<8, 4>("hmma + 4 FFMA2 each", 8);
  run<8, 8>("hmma + 8 FFMA2 each", 8);
  run<8, 8>("hmma + 8 FFMA2 each", 16);
  return 0;
}
