// Denoiser global attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// One CTA = (128-query tile, head, object).  The object's packed token rows are a segment of the
// bf16 QKV activation matrix [M, 3C] (q | k | v, head h at columns h*64); fragments that are padded
// or merged away are simply not in the packed batch, which is what the reference's key mask does
// (diffusers AttnProcessor2_0 with gen_mask, attention.py:84).
//
//   TMA      Q tile [128 x 64], K and V blocks [128 keys x 64] -> 128B-swizzled smem (one tensor map)
//   MMA 1    S[128 x keys] = Q K^T       kind::f16, M=128, N=128 per key block, fp32 in TMEM (<=512 cols)
//   softmax  4 warps, one query row per thread (its TMEM lane): pass 1 row max, pass 2 exp2 ->
//            bf16 P block written to smem in the UMMA K-major SWIZZLE_128B layout (double-buffered)
//   MMA 2    O[128 x 64] += P_blk V_blk  (V consumed as an MN-major operand straight from the TMA tile);
//            O aliases the first 64 TMEM columns of S, which are dead once key block 0 is exponentiated
//   epilogue O / rowsum -> bf16 -> global
//
// Segment lengths up to 512 tokens (20 fragments x 25 latent points = 500 in the reference).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int AT_D = 64;
constexpr int AT_BQ = 128;
constexpr int AT_BK = 128;
constexpr int AT_MAXKB = 4;
constexpr uint32_t AT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// smem: Q | K[4] | V[4] | P[2][2 panels] | barriers
constexpr uint32_t AT_OFF_Q = 0;
constexpr uint32_t AT_OFF_K = AT_TILE_BYTES;
constexpr uint32_t AT_OFF_V = AT_OFF_K + AT_MAXKB * AT_TILE_BYTES;
constexpr uint32_t AT_OFF_P = AT_OFF_V + AT_MAXKB * AT_TILE_BYTES;
constexpr uint32_t AT_P_BYTES = 2 * AT_TILE_BYTES;  // one P block = 2 panels of [128 x 64] bf16
constexpr uint32_t AT_OFF_BAR = AT_OFF_P + 2 * AT_P_BYTES;
constexpr uint32_t AT_SMEM_BYTES = AT_OFF_BAR + 128 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// K-major SWIZZLE_128B operand tile ([rows x 64] bf16, 128 B per row, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: 64 contiguous MN elements per 128 B row, K rows 128 B apart,
// groups of 8 K rows 1024 B apart (SBO); a single 64-wide MN block, so LBO is not exercised.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// D=f32, A=B=bf16; S: both K-major, M=128, N=128.  O: A K-major, B MN-major (bit 16), M=128, N=64.
constexpr uint32_t AT_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t AT_IDESC_O =
    (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AT_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(128)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ seg_start,
                        const int* __restrict__ seg_len, int C, float scale_log2e, int block,
                        __nv_bfloat16* __restrict__ out, int ldo) {
  const int seg = blockIdx.z, h = blockIdx.y;
  const int len = seg_len[seg], st = seg_start[seg];
  const int q0 = blockIdx.x * AT_BQ;
  if (q0 >= len) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_qk = base + AT_OFF_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p0 = bar_qk + 24,
                 bar_p1 = bar_qk + 32, tmem_slot = bar_qk + 40;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (len + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p0, 1);
    mbar_init(bar_p1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(base_ptr + AT_OFF_BAR + 40);

  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_qk, (1 + nkb) * AT_TILE_BYTES);
    tma_load_2d(base + AT_OFF_Q, &map, bar_qk, h * AT_D, st + q0);
    for (int kb = 0; kb < nkb; ++kb) tma_load_2d(base + AT_OFF_K + kb * AT_TILE_BYTES, &map, bar_qk, C + h * AT_D, st + kb * AT_BK);
    mbar_expect_tx(bar_v, nkb * AT_TILE_BYTES);
    for (int kb = 0; kb < nkb; ++kb)
      tma_load_2d(base + AT_OFF_V + kb * AT_TILE_BYTES, &map, bar_v, 2 * C + h * AT_D, st + kb * AT_BK);
  }
  if (threadIdx.x == 32) {
    mbar_wait(bar_qk, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t dq = desc_kmajor(base + AT_OFF_Q);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint64_t dk = desc_kmajor(base + AT_OFF_K + kb * AT_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem + kb * AT_BK, dq + 2 * k, dk + 2 * k, AT_IDESC_S, k != 0);
    }
    umma_commit(bar_s);
  }
  __syncwarp();

  // ---- softmax: thread = query row (TMEM lane 32*warp + lane) ----
  mbar_wait(bar_s, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  // key k is visible to query q iff k < len and (no block structure or same block): block > 0 gives the
  // block-diagonal local attention (one 25-token block per fragment, denoiser_transformer.py:158-166)
  int klo = 0, khi = len;
  if (block > 0) {
    klo = ((q0 + row) / block) * block;
    khi = min(len, klo + block);
  }
  float mx = -INFINITY;
  for (int c = 0; c < nkb * (AT_BK / 32); ++c) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c * 32, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int key = c * 32 + j;
      if (key >= klo && key < khi) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
  }
  const float mxs = (mx == -INFINITY) ? 0.f : mx * scale_log2e;
  float sum = 0.f;
  for (int kb = 0; kb < nkb; ++kb) {
    const int pb = kb & 1;
    if (kb >= 2) mbar_wait(pb ? bar_p1 : bar_p0, ((kb >> 1) - 1) & 1);  // MMA of block kb-2 has drained this P buffer
    uint8_t* pbase = base_ptr + AT_OFF_P + pb * AT_P_BYTES;
#pragma unroll 1
    for (int c = 0; c < AT_BK / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + kb * AT_BK + c * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        int key = kb * AT_BK + c * 32 + j;
        float p0 = (key >= klo && key < khi) ? exp2f(__uint_as_float(v[j]) * scale_log2e - mxs) : 0.f;
        float p1 = (key + 1 >= klo && key + 1 < khi) ? exp2f(__uint_as_float(v[j + 1]) * scale_log2e - mxs) : 0.f;
        __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
        // the row sum uses the bf16-rounded probabilities that the PV product actually sees
        sum += __low2float(b2) + __high2float(b2);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
      }
      // 32 keys = 64 B = 4 x 16-byte chunks of panel (c>>1), chunk index (c&1)*4 + i, swizzled with row&7
      uint8_t* prow = pbase + (c >> 1) * AT_TILE_BYTES + row * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int chunk = ((c & 1) * 4 + i) ^ (row & 7);
        *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (kb == 0) {
        mbar_wait(bar_v, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t pa = base + AT_OFF_P + pb * AT_P_BYTES;
      const uint32_t va = base + AT_OFF_V + kb * AT_TILE_BYTES;
#pragma unroll
      for (int k = 0; k < AT_BK / 16; ++k) {
        const uint64_t dp = desc_kmajor(pa + (k >> 2) * AT_TILE_BYTES) + 2 * (k & 3);
        const uint64_t dv = desc_mnmajor(va + k * 2048);
        umma_bf16(tmem, dp, dv, AT_IDESC_O, (kb | k) != 0);
      }
      umma_commit(pb ? bar_p1 : bar_p0);
    }
    __syncwarp();
  }
  // last block's commit marks O complete
  {
    const int kb = nkb - 1;
    mbar_wait((kb & 1) ? bar_p1 : bar_p0, (kb >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const float inv = 1.0f / sum;
  const bool ok = q0 + row < len;
  __nv_bfloat16* orow = out + (size_t)(st + q0 + row) * ldo + h * AT_D;
#pragma unroll
  for (int c = 0; c < AT_D / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c * 32, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
        __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
        __nv_bfloat162 a3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&a0), pk.y = *reinterpret_cast<uint32_t*>(&a1);
        pk.z = *reinterpret_cast<uint32_t*>(&a2), pk.w = *reinterpret_cast<uint32_t*>(&a3);
        *reinterpret_cast<uint4*>(orow + c * 32 + j) = pk;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

}  // namespace

extern "C" int pfpp_attention_tc(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                                 int n_segments, int max_len, int heads, int block, void* out, int ldo,
                                 cudaStream_t stream) {
  PFPP_CHECK_ARG(qkv && seg_start && seg_len && out && heads > 0 && C == heads * AT_D);
  PFPP_CHECK_ARG(max_len <= AT_MAXKB * AT_BK && (ld % 8) == 0 && (ldo % 8) == 0 && ((uintptr_t)qkv & 15) == 0 &&
                 ((uintptr_t)out & 15) == 0);
  if (n_segments == 0 || max_len <= 0 || M == 0) return PFPP_OK;
  auto fn = encode_fn();
  if (!fn) return PFPP_EUNSUPPORTED;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * C), (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)AT_D, 128};
  cuuint32_t estr[2] = {1, 1};
  if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return PFPP_EINVAL;
  PFPP_ENSURE_SMEM(attention_tc_kernel, AT_SMEM_BYTES);
  dim3 grid(pfpp_cdiv(max_len, AT_BQ), heads, n_segments);
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)AT_D);
  attention_tc_kernel<<<grid, 128, AT_SMEM_BYTES, stream>>>(map, seg_start, seg_len, C, scale_log2e, block,
                                                           (__nv_bfloat16*)out, ldo);
  PFPP_RETURN_LAST();
}
