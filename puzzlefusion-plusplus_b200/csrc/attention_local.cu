// Block-diagonal ("local") attention of the denoiser (attention.py:79 with self_mask, denoiser_transformer.py:168-171):
// every fragment's 25 latent tokens attend to each other only.  One WARP = one (block of <= 32 tokens, head):
//
//   Q, K, V  [blk x 64] bf16     cp.async (16-byte chunks) into three private 4 KB shared-memory tiles, XOR-swizzled
//   S = Q K^T, O = P V           warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate), operands via ldmatrix
//   softmax                      in the accumulator fragments (quad shuffles), P re-used as the A operand in registers
//
// The tcgen05 path (attention_ws.cu) needs 128-row tiles, so a 25 x 25 block costs a 128 x 64 score panel and a
// 128 x 64 x 64 PV product: >= 80 % of its tensor and MUFU work is masked away and the kernel ends up bound by its
// compulsory QKV read at a fifth of the HBM rate.  At this granularity the legacy warp-level MMA is the right tool:
// no padding beyond 32 x 32, no TMEM / mbarrier round trips, 16 independent warps per SM hide the load latency.
// Arithmetic is the same as the tcgen05 kernel's: bf16 operands, fp32 scores, exp2 with the row maximum, probabilities
// rounded to bf16 for the PV product, the row sum taken over the ROUNDED probabilities, fp32 output scaled by 1 / sum.
// A block's result does not depend on its neighbours (batch == the same objects one by one, bit for bit).
#include "common.cuh"
#include "mma_common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int AL_D = 64;            // head dimension
constexpr int AL_ROWS = 32;         // padded block
constexpr int AL_TILE = AL_ROWS * AL_D * 2;  // 4 KB
constexpr int AL_WARPS = 8;

__device__ __forceinline__ uint32_t al_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float al_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk c of row r inside a [32 x 128 B] tile
__device__ __forceinline__ uint32_t al_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(32 * AL_WARPS, 2)
    attention_local_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int C, long long n_tasks, int heads, int blk,
                           float scale_log2e, __nv_bfloat16* __restrict__ out, int ldo) {
  extern __shared__ __align__(1024) uint8_t al_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* tq = al_smem_raw + warp * 3 * AL_TILE;
  uint8_t *tk = tq + AL_TILE, *tv = tk + AL_TILE;
  const uint32_t sq = al_smem(tq), sk = al_smem(tk), sv = al_smem(tv);
  // rows blk..31 are never written by the copies: zero them once (masked keys must meet finite V rows: 0 * NaN = NaN)
  for (int i = lane; i < (AL_ROWS - blk) * 8 * 3; i += 32) {
    const int m = i / ((AL_ROWS - blk) * 8), rc = i - m * (AL_ROWS - blk) * 8;
    *reinterpret_cast<uint4*>(tq + m * AL_TILE + al_off(blk + rc / 8, rc & 7)) = make_uint4(0u, 0u, 0u, 0u);
  }
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, ri = lane & 7;  // ldmatrix: this lane addresses row ri of 8x8 matrix mi
  const int n_mt = blk > 16 ? 2 : 1;
  // Every shared-memory address below is (lane constant) + (compile-time constant): the XOR swizzle of chunk 2k + b on a
  // row with (row & 7) == ri is ((2k) ^ (b ^ ri)) << 4, so four chunk offsets per operand cover the k / dim steps.
  uint32_t aq[4], ak[4], av[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t xq = (uint32_t)(((2 * k) ^ ((mi >> 1) ^ ri)) << 4), xk = (uint32_t)(((2 * k) ^ ((mi & 1) ^ ri)) << 4);
    aq[k] = sq + ((mi & 1) * 8 + ri) * 128 + xq;   // Q: rows (mi & 1) * 8 + ri, chunk 2 k + (mi >> 1)
    ak[k] = sk + ((mi >> 1) * 8 + ri) * 128 + xk;  // K: rows (mi >> 1) * 8 + ri, chunk 2 k + (mi & 1)
    av[k] = sv + ((mi & 1) * 8 + ri) * 128 + xq;   // V: rows (mi & 1) * 8 + ri, chunk 2 k + (mi >> 1)
  }
  // copies: lane -> 16-byte chunk cc of rows r0 + 4 j; (row & 7) is r0 for even j and r0 ^ 4 for odd j
  const int r0 = lane >> 3, cc = lane & 7;
  const uint32_t cp_even = (uint32_t)(r0 * 128 + ((cc ^ r0) << 4)), cp_odd = (uint32_t)(r0 * 128 + ((cc ^ r0 ^ 4) << 4));
  const int n_tasks32 = (int)n_tasks;
  for (int task = blockIdx.x * AL_WARPS + warp; task < n_tasks32; task += gridDim.x * AL_WARPS) {
    const int b = task / heads;
    const int h = task - b * heads;
    const long long row0 = (long long)b * blk;
    __syncwarp();  // the previous task's reads of the tiles are complete
    const __nv_bfloat16* src0 = qkv + (size_t)(row0 + r0) * ld + h * AL_D + cc * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (r0 + 4 * j < blk) {
        const __nv_bfloat16* src = src0 + (size_t)(4 * j) * ld;
        const uint32_t dst = sq + j * 512 + ((j & 1) ? cp_odd : cp_even);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + AL_TILE), "l"(src + C) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 2 * AL_TILE), "l"(src + 2 * C) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    uint32_t ostage[2][8][2];  // bf16x2 output pairs: [m-tile][d-tile][row half]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (mt < n_mt) {
        // ---- S = Q K^T : 16 query rows x 32 keys
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4(aq[ks] + mt * 2048, a0, a1, a2, a3);
#pragma unroll
          for (int np = 0; np < 2; ++np) {  // key tiles 2 np, 2 np + 1
            uint32_t b0, b1, b2, b3;
            ldsm_x4(ak[ks] + np * 2048, b0, b1, b2, b3);
            mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
            mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
          }
        }
        // ---- softmax over the blk valid keys; rows g (regs 0, 1) and g + 8 (regs 2, 3)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt * 8 + 2 * t + e >= blk) s[nt][e] = s[nt][2 + e] = -INFINITY;
            mx0 = fmaxf(mx0, s[nt][e]);
            mx1 = fmaxf(mx1, s[nt][2 + e]);
          }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float nm0 = -mx0 * scale_log2e, nm1 = -mx1 * scale_log2e;
        uint32_t p[4][2];
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          __nv_bfloat162 q0 = __floats2bfloat162_rn(al_ex2(fmaf(s[nt][0], scale_log2e, nm0)), al_ex2(fmaf(s[nt][1], scale_log2e, nm0)));
          __nv_bfloat162 q1 = __floats2bfloat162_rn(al_ex2(fmaf(s[nt][2], scale_log2e, nm1)), al_ex2(fmaf(s[nt][3], scale_log2e, nm1)));
          l0 += __low2float(q0) + __high2float(q0);
          l1 += __low2float(q1) + __high2float(q1);
          p[nt][0] = *reinterpret_cast<uint32_t*>(&q0);
          p[nt][1] = *reinterpret_cast<uint32_t*>(&q1);
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        // ---- O = P V : 16 rows x 64 dims, keys in two k16 steps
        float o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {  // dim tiles 2 dp, 2 dp + 1
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(av[dp] + kk * 2048, b0, b1, b2, b3);
            mma_bf16(o[2 * dp], p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1], b0, b1);
            mma_bf16(o[2 * dp + 1], p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1], b2, b3);
          }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          __nv_bfloat162 r0 = __floats2bfloat162_rn(o[dt][0] * i0, o[dt][1] * i0);
          __nv_bfloat162 r1 = __floats2bfloat162_rn(o[dt][2] * i1, o[dt][3] * i1);
          ostage[mt][dt][0] = *reinterpret_cast<uint32_t*>(&r0);
          ostage[mt][dt][1] = *reinterpret_cast<uint32_t*>(&r1);
        }
      }
    }
    // ---- output: fragments -> the Q tile (every lane is done with Q) -> 16-byte coalesced stores of rows < blk
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (mt < n_mt) {
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          *reinterpret_cast<uint32_t*>(tq + al_off(mt * 16 + g, dt) + t * 4) = ostage[mt][dt][0];
          *reinterpret_cast<uint32_t*>(tq + al_off(mt * 16 + 8 + g, dt) + t * 4) = ostage[mt][dt][1];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (r0 + 4 * j < blk)
        *reinterpret_cast<uint4*>(out + (size_t)(row0 + r0 + 4 * j) * ldo + h * AL_D + cc * 8) =
            *reinterpret_cast<const uint4*>(tq + j * 512 + ((j & 1) ? cp_odd : cp_even));
    }
    // rows >= blk of the Q tile now hold finite leftovers of the staging: they only ever feed the discarded query rows
  }
}

}  // namespace

extern "C" int pfpp_attention_local(const void* qkv, long long M, int ld, int C, int heads, int block, void* out, int ldo,
                                    cudaStream_t stream) {
  PFPP_CHECK_ARG(qkv && out && heads > 0 && C == heads * AL_D && block > 0 && block <= AL_ROWS && M >= 0 && (M % block) == 0);
  PFPP_CHECK_ARG((ld % 8) == 0 && (ldo % 8) == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0);
  if (M == 0) return PFPP_OK;
  const long long n_tasks = M / block * heads;
  const int smem = AL_WARPS * 3 * AL_TILE;
  PFPP_ENSURE_SMEM(attention_local_kernel, smem);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  long long grid = (n_tasks + AL_WARPS - 1) / AL_WARPS;
  if (grid > 2LL * sms) grid = 2LL * sms;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)AL_D);
  attention_local_kernel<<<(unsigned)grid, 32 * AL_WARPS, smem, stream>>>((const __nv_bfloat16*)qkv, ld, C, n_tasks, heads, block,
                                                                          scale_log2e, (__nv_bfloat16*)out, ldo);
  PFPP_RETURN_LAST();
}
