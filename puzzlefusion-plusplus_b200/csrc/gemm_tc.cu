// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
//   C[M, N'] = epilogue(A[M, K] * W[N, K]^T + bias[N]) (+ residual)
//
// A (activations) and W (nn.Linear / 1x1-conv weights) are both K-major bf16, so each operand tile
// is a [rows x 64] box loaded by one cp.async.bulk.tensor into a 128B-swizzled shared-memory tile
// that tcgen05.mma consumes directly through a shared-memory matrix descriptor.  The fp32
// accumulator tile (128 x BN) lives in tensor memory; after the K loop the four warps read their
// TMEM lane quarter with tcgen05.ld and apply the fused epilogue (bias, ReLU/GELU/SiLU/GEGLU,
// residual) in registers before storing fp32 or bf16.
//
// CTA = 128 threads = 4 warps: lane 0 of warp 0 is the TMA producer, lane 0 of warp 1 issues the
// MMAs (single-thread issue), all four warps are the epilogue.  A 3-stage mbarrier ring
// (full/empty) decouples TMA from MMA; ~97 KB of shared memory per CTA lets two CTAs share an SM so
// one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

#include "common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BN = 128;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 3;
constexpr int TC_UMMA_K = 16;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

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

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address       bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset bits [16,30) (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset  bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// instruction descriptor (kind::f16): D=f32, A=B=bf16, K-major both, M=128, N=TC_BN
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

template <int EPI>
__device__ __forceinline__ float tc_act(float v) {
  if (EPI == PFPP_EPI_RELU) return fmaxf(v, 0.f);
  if (EPI == PFPP_EPI_GELU) return gelu_erf(v);
  if (EPI == PFPP_EPI_SILU) return silu(v);
  return v;
}

template <int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(128)
    gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const float* __restrict__ bias, const float* residual, int ldr, void* Cout, int ldc, int M,
                        int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  // full[s] at bar_base + 8 s ; empty[s] at bar_base + 8 (STAGES + s) ; tmem_full at bar_base + 16 STAGES
  const uint32_t bar_tmem_full = bar_base + 16 * TC_STAGES;
  const uint32_t tmem_slot = bar_tmem_full + 8;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int num_kb = (K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (TC_STAGES + s), 1);
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ----
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t round = kb / TC_STAGES;
      mbar_wait(bar_base + 8 * (TC_STAGES + s), (round & 1) ^ 1);  // slot free (passes immediately in round 0)
      const uint32_t full = bar_base + 8 * s;
      mbar_expect_tx(full, TC_STAGE_BYTES);
      const uint32_t sa = smem_base + s * TC_STAGE_BYTES;
      tma_load_2d(sa, &map_a, full, kb * TC_BK, m0);
      tma_load_2d(sa + TC_A_BYTES, &map_b, full, kb * TC_BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t round = kb / TC_STAGES;
      mbar_wait(bar_base + 8 * s, round & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_base + s * TC_STAGE_BYTES;
      const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + TC_A_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
        // advance 32 B (16 bf16) inside the 128 B swizzle atom: +2 in the 16-byte-unit address field
        umma_bf16(tmem_base, da + 2 * k, db + 2 * k, TC_IDESC, (kb | k) != 0);
      }
      umma_commit(bar_base + 8 * (TC_STAGES + s));  // frees the smem slot when these MMAs retire
    }
    umma_commit(bar_tmem_full);  // accumulator complete
  }
  __syncwarp();

  // ---- epilogue: all four warps, warp w owns TMEM lanes [32w, 32w+32) = rows m0+32w .. ----
  mbar_wait(bar_tmem_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int c = 0; c < TC_BN / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int nb = n0 + c * 32;
    if (row < M && nb < N) {
      if (EPI == PFPP_EPI_GEGLU) {
        // interleaved (value, gate) column pairs -> 16 outputs at columns nb/2 .. nb/2+15
        float o[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int n = nb + j;
          float val = __uint_as_float(v[j]) + ((bias && n < N) ? bias[n] : 0.f);
          float gate = __uint_as_float(v[j + 1]) + ((bias && n + 1 < N) ? bias[n + 1] : 0.f);
          o[j >> 1] = val * gelu_erf(gate);
        }
        const size_t off = (size_t)row * ldc + (nb >> 1);
        if (nb + 32 <= N && OUT_BF16 && (ldc % 8) == 0) {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + off;
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 pk;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[j], o[j + 1]), p1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), p3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2), pk.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(cp + j) = pk;
          }
        } else {
          for (int j = 0; j < 16; ++j) {
            if (nb + 2 * j + 1 < N) {
              if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(Cout)[off + j] = __float2bfloat16_rn(o[j]);
              else reinterpret_cast<float*>(Cout)[off + j] = o[j];
            }
          }
        }
      } else {
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          int n = nb + j;
          float t = __uint_as_float(v[j]) + ((bias && n < N) ? bias[n] : 0.f);
          o[j] = tc_act<EPI>(t);
        }
        if (residual) {
          const float* rp = residual + (size_t)row * ldr + nb;
          if (nb + 32 <= N && (ldr % 4) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              o[j] += r4.x, o[j + 1] += r4.y, o[j + 2] += r4.z, o[j + 3] += r4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) o[j] += rp[j];
          }
        }
        if (OUT_BF16) {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)row * ldc + nb;
          if (nb + 32 <= N && (ldc % 8) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 p0 = __floats2bfloat162_rn(o[j], o[j + 1]), p1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), p3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2), pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(cp + j) = pk;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) cp[j] = __float2bfloat16_rn(o[j]);
          }
        } else {
          float* cp = reinterpret_cast<float*>(Cout) + (size_t)row * ldc + nb;
          if (nb + 32 <= N && (ldc % 4) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) cp[j] = o[j];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_BN));
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for the wide denoiser projections (N >= 256): 128 x 256 output tiles, a 4-stage
// TMA ring, the fp32 accumulator double-buffered in TMEM (2 x 256 columns) and dedicated epilogue
// warps, so the epilogue of tile i (bias / GEGLU / residual, global stores) overlaps the main loop of
// tile i+1.  One CTA per SM, static round-robin tile schedule with n fastest (the A row block is
// re-used from L2 by consecutive tiles).
//   warp 0 : TMA producer (lane 0)      warp 1 : MMA issuer (lane 0) + TMEM owner
//   warps 2-9 : epilogue, warp w reads TMEM lane quarter (w % 4); warps 2-5 take columns 0-127 of the
//               tile, warps 6-9 columns 128-255 (the epilogue, not the main loop, bounds these shapes)
// ------------------------------------------------------------------------------------------------
constexpr int T2_BN = 256;
constexpr int T2_STAGES = 4;
constexpr uint32_t T2_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t T2_B_BYTES = T2_BN * TC_BK * 2;
constexpr uint32_t T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;
constexpr uint32_t T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + 1024 + 256;
constexpr uint32_t T2_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T2_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int T2_THREADS = 64 + 256;

template <int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(T2_THREADS)
    gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const float* __restrict__ bias, const float* residual, int ldr, void* Cout, int ldc, int M,
                         int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + T2_STAGES * T2_STAGE_BYTES;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * T2_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * T2_STAGES;  // [2]
  const uint32_t bar_tempty = bar_tfull + 16;            // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + T2_BN - 1) / T2_BN, tiles_m = (M + TC_BM - 1) / TC_BM;
  const int num_tiles = tiles_n * tiles_m;
  const int num_kb = (K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * T2_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % T2_STAGES, round = it / T2_STAGES;
          mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
          mbar_expect_tx(bar_full + 8 * s, T2_STAGE_BYTES);
          const uint32_t sa = smem_base + s * T2_STAGE_BYTES;
          tma_load_2d(sa, &map_a, bar_full + 8 * s, kb * TC_BK, m0);
          tma_load_2d(sa + T2_A_BYTES, &map_b, bar_full + 8 * s, kb * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * T2_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % T2_STAGES, round = it / T2_STAGES;
          mbar_wait(bar_full + 8 * s, round & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * T2_STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + T2_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, T2_IDESC, (kb | k) != 0);
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;      // which 128-column half of the tile this warp drains
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * T2_BN;
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + buf * T2_BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = chalf * (T2_BN / 64); c < (chalf + 1) * (T2_BN / 64); ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int nb = n0 + c * 32;
        if (row < M && nb < N) {
          if (EPI == PFPP_EPI_GEGLU) {
            float o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int n = nb + j;
              float val = __uint_as_float(v[j]) + ((bias && n < N) ? bias[n] : 0.f);
              float gate = __uint_as_float(v[j + 1]) + ((bias && n + 1 < N) ? bias[n + 1] : 0.f);
              o[j >> 1] = val * gelu_erf(gate);
            }
            const size_t off = (size_t)row * ldc + (nb >> 1);
            if (nb + 32 <= N && OUT_BF16 && (ldc % 8) == 0) {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + off;
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                uint4 pk;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(o[j], o[j + 1]), p1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), p3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2), pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(cp + j) = pk;
              }
            } else {
              for (int j = 0; j < 16; ++j) {
                if (nb + 2 * j + 1 < N) {
                  if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(Cout)[off + j] = __float2bfloat16_rn(o[j]);
                  else reinterpret_cast<float*>(Cout)[off + j] = o[j];
                }
              }
            }
          } else {
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              o[j] = tc_act<EPI>(__uint_as_float(v[j]) + ((bias && n < N) ? bias[n] : 0.f));
            }
            if (residual) {
              const float* rp = residual + (size_t)row * ldr + nb;
              if (nb + 32 <= N && (ldr % 4) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                  o[j] += r4.x, o[j + 1] += r4.y, o[j + 2] += r4.z, o[j + 3] += r4.w;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < N) o[j] += rp[j];
              }
            }
            if (OUT_BF16) {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)row * ldc + nb;
              if (nb + 32 <= N && (ldc % 8) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk;
                  __nv_bfloat162 p0 = __floats2bfloat162_rn(o[j], o[j + 1]), p1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
                  __nv_bfloat162 p2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), p3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
                  pk.z = *reinterpret_cast<uint32_t*>(&p2), pk.w = *reinterpret_cast<uint32_t*>(&p3);
                  *reinterpret_cast<uint4*>(cp + j) = pk;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < N) cp[j] = __float2bfloat16_rn(o[j]);
              }
            } else {
              float* cp = reinterpret_cast<float*>(Cout) + (size_t)row * ldc + nb;
              if (nb + 32 <= N && (ldc % 4) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (nb + j < N) cp[j] = o[j];
              }
            }
          }
        }
      }
      // this warp is done with the accumulator buffer: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ----
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 64 cols], 128B swizzle
int make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) return PFPP_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PFPP_OK : PFPP_EINVAL;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int EPI>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, const float* residual, int ldr, void* C,
               int ldc, int c_bf16, int M, int N, int K, cudaStream_t stream) {
  const int tiles = pfpp_cdiv(N, T2_BN) * pfpp_cdiv(M, TC_BM);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (c_bf16) {
    PFPP_ENSURE_SMEM((gemm_bf16_tc2_kernel<EPI, true>), T2_SMEM_BYTES);
    gemm_bf16_tc2_kernel<EPI, true><<<grid, T2_THREADS, T2_SMEM_BYTES, stream>>>(ma, mb, bias, residual, ldr, C, ldc, M, N, K);
  } else {
    PFPP_ENSURE_SMEM((gemm_bf16_tc2_kernel<EPI, false>), T2_SMEM_BYTES);
    gemm_bf16_tc2_kernel<EPI, false><<<grid, T2_THREADS, T2_SMEM_BYTES, stream>>>(ma, mb, bias, residual, ldr, C, ldc, M, N, K);
  }
  PFPP_RETURN_LAST();
}

template <int EPI>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, const float* residual, int ldr, void* C,
              int ldc, int c_bf16, int M, int N, int K, cudaStream_t stream) {
  dim3 grid(pfpp_cdiv(N, TC_BN), pfpp_cdiv(M, TC_BM));
  if (c_bf16) {
    PFPP_ENSURE_SMEM((gemm_bf16_tc_kernel<EPI, true>), TC_SMEM_BYTES);
    gemm_bf16_tc_kernel<EPI, true><<<grid, 128, TC_SMEM_BYTES, stream>>>(ma, mb, bias, residual, ldr, C, ldc, M, N, K);
  } else {
    PFPP_ENSURE_SMEM((gemm_bf16_tc_kernel<EPI, false>), TC_SMEM_BYTES);
    gemm_bf16_tc_kernel<EPI, false><<<grid, 128, TC_SMEM_BYTES, stream>>>(ma, mb, bias, residual, ldr, C, ldc, M, N, K);
  }
  PFPP_RETURN_LAST();
}

}  // namespace

extern "C" int pfpp_has_tensor_core_path(void) { return 1; }

extern "C" int pfpp_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual,
                              int ldr, void* C, int ldc, int c_bf16, int M, int N, int K, int epilogue,
                              cudaStream_t stream) {
  PFPP_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  PFPP_CHECK_ARG((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0);
  PFPP_CHECK_ARG((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0);
  if (M == 0) return PFPP_OK;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, K, lda, TC_BM);
  if (rc) return rc;
  // persistent 128x256 kernel for the wide projections; the GEGLU epilogue (128 erf per row per tile) is
  // better spread over the 2 CTAs/SM of the 128x128 kernel (measured: 183 vs 254 us at M=16000, N=4096)
  static const int v2_mode = []() {
    const char* e = getenv("PFPP_GEMM_V2");  // 0 = never, 1 = default policy, 2 = always (tuning knob)
    return e ? atoi(e) : 1;
  }();
  const bool wide = N >= T2_BN && M >= 2 * TC_BM && v2_mode != 0 && (epilogue != PFPP_EPI_GEGLU || v2_mode == 2);
  rc = make_map(&mb, W, N, K, ldw, wide ? T2_BN : TC_BN);
  if (rc) return rc;
  if (wide) {
    switch (epilogue) {
      case PFPP_EPI_NONE:
        return launch_tc2<PFPP_EPI_NONE>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
      case PFPP_EPI_RELU:
        return launch_tc2<PFPP_EPI_RELU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
      case PFPP_EPI_GELU:
        return launch_tc2<PFPP_EPI_GELU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
      case PFPP_EPI_SILU:
        return launch_tc2<PFPP_EPI_SILU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
      case PFPP_EPI_GEGLU:
        return launch_tc2<PFPP_EPI_GEGLU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
      default:
        return PFPP_EINVAL;
    }
  }
  switch (epilogue) {
    case PFPP_EPI_NONE:
      return launch_tc<PFPP_EPI_NONE>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
    case PFPP_EPI_RELU:
      return launch_tc<PFPP_EPI_RELU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
    case PFPP_EPI_GELU:
      return launch_tc<PFPP_EPI_GELU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
    case PFPP_EPI_SILU:
      return launch_tc<PFPP_EPI_SILU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
    case PFPP_EPI_GEGLU:
      return launch_tc<PFPP_EPI_GEGLU>(ma, mb, bias, residual, ldr, C, ldc, c_bf16, M, N, K, stream);
    default:
      return PFPP_EINVAL;
  }
}
