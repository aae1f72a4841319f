// SelfAttention core of the ESRGAN generator (ESRGAN_model.py:58-66) on the tensor cores:
//     beta = softmax(g f^T)  (over the keys),   o = beta h        f, g: [B, HW, dk]   h: [B, HW, dv]
// used by the 16-bit precision modes (the float32 mode keeps the exact CUDA-core kernel in core.cu).
//
// One CTA owns 128 queries of one image and walks the keys in blocks of 128 (flash-attention style online softmax):
//   S  = Q K_j^T        tcgen05.mma M128 N128, accumulator in TMEM.  The scores keep float32-grade accuracy although the
//                        operands are fp16: Q rows are [hi | hi | lo], K rows [hi | lo | hi] with lo = fp16(x - hi), so one
//                        K = 48 product is hi.hi + hi.lo + lo.hi (the softmax of a trained generator is peaked: rounding
//                        g and f to 11 bits would move the result by far more than the fp16 rounding of P below);
//   P  = exp2((S - m) log2 e)  by four softmax warps, one query row per thread (TMEM lane = row): two passes over the
//                        accumulator (row maximum, then exponentials), P written to shared memory as the K-major,
//                        128-byte-swizzled A operand of the second product (fp16: P is in [0, 1]);
//   O_j = P V_j          tcgen05.mma M128 N = dv from V^T tiles (dv rows x 64 keys), accumulator in TMEM; the running output
//                        (dv floats per row) lives in the softmax threads' registers and is rescaled there.
// A preparation kernel builds the padded 16-bit operands (and the transpose of h) in a workspace; TMA moves every tile.
// Two CTAs per SM (96 KB of shared memory and 256 TMEM columns each) overlap one CTA's softmax with the other's MMAs.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <math.h>

namespace srb {

constexpr int kAB = 128;                 // queries per CTA = keys per block
constexpr int kAThreads = 192;           // warp 0: TMA, warp 1: MMA, warps 2-5: softmax
constexpr uint32_t kATile = 128u * 128u; // one 128-row x 128-byte operand tile

struct AttnParams {
  int n, n_blocks, dv;
  uint32_t idesc_qk, idesc_pv;
  float* o;                              // [B, n, dv]
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {     // MUFU.EX2 (flush-to-zero): exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// ---- operand preparation: fp32 f / g / h -> padded fp16 rows of Q, K (hi / lo split) and the transpose of h -------------
//   Q row (64 halves = one 128-byte swizzle row): [0, dk) hi(g), [16, 16 + dk) hi(g), [32, 32 + dk) lo(g), zeros elsewhere
//   K row:                                         [0, dk) hi(f), [16, 16 + dk) lo(f), [32, 32 + dk) hi(f)
//   Vt[b][d][key] = fp16(h[b][key][d]), keys padded with zeros to n_pad (a multiple of 64)
__global__ void __launch_bounds__(128)
attn_prepare_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ h, int n, int n_pad,
                    int dk, int dv, __half* __restrict__ Q, __half* __restrict__ K, __half* __restrict__ Vt) {
  const int b = blockIdx.y, r = blockIdx.x * 128 + threadIdx.x;
  if (r >= n_pad) return;
  __half* vt = Vt + (size_t)b * dv * n_pad + r;
  if (r >= n) {
    for (int d = 0; d < dv; ++d) vt[(size_t)d * n_pad] = __float2half_rn(0.f);
    return;
  }
  const size_t row = (size_t)b * n + r;
  __align__(16) __half q[64], k[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) q[i] = k[i] = __float2half_rn(0.f);
  for (int d = 0; d < dk; ++d) {
    const float gv = __ldg(g + row * dk + d), fv = __ldg(f + row * dk + d);
    const __half gh = __float2half_rn(gv), fh = __float2half_rn(fv);
    const __half gl = __float2half_rn(gv - __half2float(gh)), fl = __float2half_rn(fv - __half2float(fh));
    q[d] = gh; q[16 + d] = gh; q[32 + d] = gl;
    k[d] = fh; k[16 + d] = fl; k[32 + d] = fh;
  }
  uint4* qd = reinterpret_cast<uint4*>(Q + row * 64);
  uint4* kd = reinterpret_cast<uint4*>(K + row * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) { qd[i] = reinterpret_cast<const uint4*>(q)[i]; kd[i] = reinterpret_cast<const uint4*>(k)[i]; }
  for (int d = 0; d < dv; ++d) vt[(size_t)d * n_pad] = __float2half_rn(__ldg(h + row * dv + d));
}

template <int DV>
__global__ void __launch_bounds__(kAThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr uint32_t kVTile = (uint32_t)DV * 128u;                 // V^T tile: DV rows x 64 keys
  constexpr uint32_t kStage = kATile + 2u * kVTile;                // K block + two V^T tiles
  const uint32_t q_smem = base, kv_smem = base + kATile, p_smem = kv_smem + 2u * kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kATile + 2u * kStage + 2u * kATile);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, s_full = bar0 + 8u, p_ready = bar0 + 16u, o_full = bar0 + 24u;
  auto kv_full = [&](int s) { return bar0 + 32u + 8u * (uint32_t)s; };
  auto kv_empty = [&](int s) { return bar0 + 48u + 8u * (uint32_t)s; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = (int)blockIdx.x, b = (int)blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128u;
  const int nb = p.n_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, kATile);
      tma_load_3d(q_smem, &tm_q, q_full, 0, qb * kAB, b);
    }
    __syncwarp();
    for (int j = 0; j < nb; ++j) {
      const int s = j & 1;
      mbar_wait(kv_empty(s), (((uint32_t)j >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t dst = kv_smem + (uint32_t)s * kStage;
        mbar_expect_tx(kv_full(s), kStage);
        tma_load_3d(dst, &tm_k, kv_full(s), 0, j * kAB, b);
        tma_load_3d(dst + kATile, &tm_v, kv_full(s), j * kAB, 0, b);
        tma_load_3d(dst + kATile + kVTile, &tm_v, kv_full(s), j * kAB + 64, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(q_full, 0);
    const uint64_t q_desc = make_desc(q_smem, 1024u);
    const uint64_t p_desc0 = make_desc(p_smem, 1024u), p_desc1 = make_desc(p_smem + kATile, 1024u);
    for (int j = 0; j < nb; ++j) {
      const int s = j & 1;
      mbar_wait(kv_full(s), ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t st = kv_smem + (uint32_t)s * kStage;
      const uint64_t k_desc = make_desc(st, 1024u);
      if (elect_one()) {                                  // S = Q K^T: three k16 steps ([hi | hi | lo] . [hi | lo | hi])
#pragma unroll
        for (int k = 0; k < 3; ++k) umma_f16(tmem_s, q_desc + 2u * k, k_desc + 2u * k, p.idesc_qk, (uint32_t)(k != 0));
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, (uint32_t)j & 1u);               // P_j is in shared memory; S and the previous O block have been read
      tc_fence_after();
      const uint64_t v_desc0 = make_desc(st + kATile, 1024u), v_desc1 = make_desc(st + kATile + kVTile, 1024u);
      if (elect_one()) {                                  // O_j = P V_j: two 64-key tiles x four k16 steps
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_o, p_desc0 + 2u * k, v_desc0 + 2u * k, p.idesc_pv, (uint32_t)(k != 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_o, p_desc1 + 2u * k, v_desc1 + 2u * k, p.idesc_pv, 1u);
        umma_commit(o_full);
        umma_commit(kv_empty(s));
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float c = 1.4426950408889634f;                  // log2(e)
    float m = -INFINITY, l = 0.f;
    float O[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) O[i] = 0.f;
    const uint32_t p_row = p_smem + (uint32_t)r * 128u, p_x = (uint32_t)r & 7u;
    auto add_o_block = [&]() {                            // O += the finished P V product (scaled to the current maximum m)
      uint32_t ob[32];
#pragma unroll
      for (int c0 = 0; c0 < DV; c0 += 32) {
        __syncwarp();
        tmem_ld32(tmem_o + lane_base + (uint32_t)c0, ob);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < DV) O[c0 + i] += __uint_as_float(ob[i]);
      }
    };
    for (int j = 0; j < nb; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int key0 = j * kAB;
      const bool ragged = key0 + kAB > p.n;               // (last block: keys past n are masked)
      // pass A: row maximum of this block
      float mb = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tmem_s + lane_base + (uint32_t)(32 * cc), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float sv = (ragged && key0 + 32 * cc + e >= p.n) ? -INFINITY : __uint_as_float(v[e]);
          mb = fmaxf(mb, sv);
        }
      }
      const float m_new = fmaxf(m, mb);
      const float alpha = fast_exp2((m - m_new) * c);         // (first block: m = -inf -> 0; O and l are 0 anyway)
      if (j > 0) {
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);        // P V of the previous block is complete: add it, and P may be rewritten
        tc_fence_after();
        add_o_block();
      }
#pragma unroll
      for (int i = 0; i < DV; ++i) O[i] *= alpha;
      l *= alpha;
      m = m_new;
      // pass B: exponentials -> P (fp16, swizzled A operand), row sum
      const float mc = m * c;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tmem_s + lane_base + (uint32_t)(32 * cc), v);
        tmem_ld_wait();
        uint32_t pk[16];
        float ls = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float s0 = __uint_as_float(v[e]), s1 = __uint_as_float(v[e + 1]);
          if (ragged) {
            if (key0 + 32 * cc + e >= p.n) s0 = -INFINITY;
            if (key0 + 32 * cc + e + 1 >= p.n) s1 = -INFINITY;
          }
          const float p0 = fast_exp2(fmaf(s0, c, -mc)), p1 = fast_exp2(fmaf(s1, c, -mc));
          ls += p0 + p1;
          pk[e >> 1] = pack2(p0, p1, SRB_F16);
        }
        l += ls;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {                  // 16-byte chunk = 8 keys; chunk index 4 cc + q4 of the row's 16
          const uint32_t ch = (uint32_t)(4 * cc + q4), tile = ch >> 3, c8 = ch & 7u;
          sts128(p_row + tile * kATile + ((c8 ^ p_x) << 4), make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
        }
      }
      tc_fence_before();                                  // this thread's TMEM reads of S (and of the O block) are done
      fence_proxy_async_smem();                           // P rows (generic proxy) -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    mbar_wait(o_full, (uint32_t)(nb - 1) & 1u);
    tc_fence_after();
    add_o_block();
    tc_fence_before();
    const int row = qb * kAB + r;
    if (row < p.n) {
      const float inv = 1.f / l;
      float4* op = reinterpret_cast<float4*>(p.o + ((size_t)b * p.n + row) * DV);
#pragma unroll
      for (int i = 0; i < DV / 4; ++i) op[i] = make_float4(O[4 * i] * inv, O[4 * i + 1] * inv, O[4 * i + 2] * inv, O[4 * i + 3] * inv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static size_t attn_workspace(int batch, int hw, int dv) {
  const size_t n_pad = ((size_t)hw + 63) & ~(size_t)63;
  return (size_t)batch * hw * 128 * 2 + (size_t)batch * dv * n_pad * 2 + 256;
}

}  // namespace srb

using namespace srb;

extern "C" size_t srb_self_attention_tc_workspace(int batch, int hw, int dk, int dv) {
  (void)dk;
  if (batch <= 0 || hw <= 0 || dv <= 0) return 0;
  return attn_workspace(batch, hw, dv);
}

extern "C" int srb_self_attention_tc(const float* f, const float* g, const float* h, int batch, int hw, int dk, int dv,
                                     float* o, void* workspace, size_t workspace_bytes, srb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRB_REQUIRE(batch >= 0 && hw > 0, "self_attention_tc: bad geometry");
  if (batch == 0) return SRB_OK;
  SRB_REQUIRE(f && g && h && o && workspace, "self_attention_tc: null pointer");
  SRB_REQUIRE(dk >= 1 && dk <= 16, "self_attention_tc: dk must be 1..16 (got %d)", dk);
  SRB_REQUIRE(dv == 16 || dv == 32 || dv == 64, "self_attention_tc: dv must be 16, 32 or 64 (got %d)", dv);
  SRB_REQUIRE(workspace_bytes >= attn_workspace(batch, hw, dv), "self_attention_tc: workspace too small");
  SRB_REQUIRE(batch <= 65535, "self_attention_tc: batch too large for one launch");
  SRB_REQUIRE(aligned16(o) && (reinterpret_cast<uintptr_t>(workspace) & 127u) == 0, "self_attention_tc: o / workspace alignment");
  EncodeTiledFn encode = tc_encode_fn();
  if (!encode) { set_error("self_attention_tc: cuTensorMapEncodeTiled is not available from the driver"); return SRB_E_CUDA; }
  const int n = hw, n_pad = (hw + 63) & ~63;
  __half* Q = reinterpret_cast<__half*>(workspace);
  __half* K = Q + (size_t)batch * n * 64;
  __half* Vt = K + (size_t)batch * n * 64;
  dim3 pgrid((n_pad + 127) / 128, batch);
  attn_prepare_kernel<<<pgrid, 128, 0, stream>>>(f, g, h, n, n_pad, dk, dv, Q, K, Vt);
  int rc = launch_check("attn_prepare_kernel");
  if (rc) return rc;

  CUtensorMap tq, tk, tv;
  const cuuint32_t es[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)n, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {128, (cuuint64_t)n * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)kAB, 1};
    if (encode(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, Q, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&tk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, K, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("self_attention_tc: cuTensorMapEncodeTiled(q / k) failed");
      return SRB_E_CUDA;
    }
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)n_pad, (cuuint64_t)dv, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)n_pad * 2, (cuuint64_t)dv * n_pad * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)dv, 1};
    if (encode(&tv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, Vt, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("self_attention_tc: cuTensorMapEncodeTiled(v) failed");
      return SRB_E_CUDA;
    }
  }
  AttnParams p{};
  p.n = n; p.n_blocks = (n + kAB - 1) / kAB; p.dv = dv; p.o = o;
  p.idesc_qk = (1u << 4) | ((uint32_t)(kAB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc_pv = (1u << 4) | ((uint32_t)(dv >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = 1024 + kATile + 2 * (size_t)(kATile + 2 * dv * 128) + 2 * kATile + 128;
  typedef void (*AttnFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
  const AttnFn kernel = dv == 16 ? attention_tc_kernel<16> : dv == 32 ? attention_tc_kernel<32> : attention_tc_kernel<64>;
  static size_t configured[3] = {0, 0, 0};
  const int ki = dv == 16 ? 0 : dv == 32 ? 1 : 2;
  if (smem > configured[ki]) {
    SRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ki] = smem;
  }
  dim3 grid(p.n_blocks, batch);
  kernel<<<grid, kAThreads, smem, stream>>>(tq, tk, tv, p);
  return launch_check("attention_tc_kernel");
}
