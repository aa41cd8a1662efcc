// Split-operand (fp32-grade) GEMM for a few hundred rows: the denoiser's output heads (denoiser_transformer.py:121-147:
// mlp_out_trans / mlp_out_rot on the mean-pooled fragment tokens, M = fragments).
//
//   C[M, N] = act(A Wt + b),   A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo        (bf16 hi/lo pairs, fp32 accumulate, 2^-16)
//
// Same operand formats and semantics as the tcgen05 path of pfpp_gemm_bf16x3 (rows [hi | lo] of A and W, fp32 or split
// output), which is what pfpp_gemm_bf16x3 dispatches to below M = 1024: there a 256 x 256 CTA-pair tile with 24 k-blocks
// in flight costs ~28 us per launch for 0.1 - 0.4 GFLOP (fixed pipeline cost, one or two tiles busy), five launches per
// DDPM step.  Here a CTA owns a 32 x 64 tile, four warps run mma.sync m16n8k16 on ldmatrix operands, and the k-blocks
// (A_hi, A_lo, W_hi, W_lo: 24 KB) stream through a two-stage cp.async ring.
#include "common.cuh"
#include "mma_common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int GS_BM = 32, GS_BN = 64, GS_BK = 64;
constexpr int GS_A_BYTES = GS_BM * 128, GS_W_BYTES = GS_BN * 128;      // one [rows x 64] bf16 tile
constexpr int GS_STAGE = 2 * GS_A_BYTES + 2 * GS_W_BYTES;              // A_hi | A_lo | W_hi | W_lo = 24 KB
constexpr int GS_THREADS = 128;

template <int EPI>
__device__ __forceinline__ float gs_act(float v) {
  if (EPI == PFPP_EPI_RELU) return fmaxf(v, 0.f);
  if (EPI == PFPP_EPI_GELU) return gelu_erf(v);
  if (EPI == PFPP_EPI_SILU) return silu(v);
  return v;
}

template <int EPI, bool SPLIT_OUT>
__global__ void __launch_bounds__(GS_THREADS)
    gemm_small_x3_kernel(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ W, int ldw,
                         const float* __restrict__ bias, void* __restrict__ C, int ldc, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t gs_smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(gs_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
  const int nkb = (K + GS_BK - 1) / GS_BK;
  const __nv_bfloat16* A_lo = A + lda / 2;
  const __nv_bfloat16* W_lo = W + ldw / 2;

  // stage loader: 16-byte chunks; rows / k-chunks outside the problem are zero-filled (src-size 0)
  auto load_stage = [&](int kb, int st) {
    const uint32_t s0 = sbase + st * GS_STAGE;
    const int k0 = kb * GS_BK;
#pragma unroll
    for (int i = 0; i < (GS_BM * 8) / GS_THREADS; ++i) {  // A tiles: 32 rows x 8 chunks
      const int ch = tid + i * GS_THREADS, r = ch >> 3, c = ch & 7;
      const bool ok = m0 + r < M && k0 + c * 8 < K;
      const size_t off = ok ? (size_t)(m0 + r) * lda + k0 + c * 8 : 0;
      const int nbytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s0 + sw128(r, c)), "l"(A + off), "r"(nbytes) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s0 + GS_A_BYTES + sw128(r, c)), "l"(A_lo + off), "r"(nbytes)
                   : "memory");
    }
#pragma unroll
    for (int i = 0; i < (GS_BN * 8) / GS_THREADS; ++i) {  // W tiles: 64 rows x 8 chunks
      const int ch = tid + i * GS_THREADS, r = ch >> 3, c = ch & 7;
      const bool ok = n0 + r < N && k0 + c * 8 < K;
      const size_t off = ok ? (size_t)(n0 + r) * ldw + k0 + c * 8 : 0;
      const int nbytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s0 + 2 * GS_A_BYTES + sw128(r, c)), "l"(W + off), "r"(nbytes)
                   : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s0 + 2 * GS_A_BYTES + GS_W_BYTES + sw128(r, c)), "l"(W_lo + off),
                   "r"(nbytes)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // warp tile: rows wm * 16 .. + 15, columns wn * 32 .. + 31 (four n-tiles of 8)
  const int wm = warp >> 1, wn = warp & 1;
  const int mi = lane >> 3, ri = lane & 7;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // ldmatrix row / chunk of this lane: A (rows (mi & 1) * 8 + ri, chunk 2 k + (mi >> 1)); W (rows (mi >> 1) * 8 + ri,
  // chunk 2 k + (mi & 1)); (row & 7) == ri in both
  const int a_row = wm * 16 + (mi & 1) * 8 + ri, a_cx = (mi >> 1) ^ ri;
  const int w_row = wn * 32 + (mi >> 1) * 8 + ri, w_cx = (mi & 1) ^ ri;

  load_stage(0, 0);
  for (int kb = 0; kb < nkb; ++kb) {
    if (kb + 1 < nkb) {
      load_stage(kb + 1, (kb + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t s0 = sbase + (kb & 1) * GS_STAGE;
    const uint32_t sa_hi = s0 + a_row * 128, sa_lo = sa_hi + GS_A_BYTES;
    const uint32_t sw_hi = s0 + 2 * GS_A_BYTES + w_row * 128, sw_lo = sw_hi + GS_W_BYTES;
#pragma unroll
    for (int k = 0; k < GS_BK / 16; ++k) {
      const uint32_t ca = (uint32_t)(((2 * k) ^ a_cx) << 4), cw = (uint32_t)(((2 * k) ^ w_cx) << 4);
      uint32_t ah[4], al[4];
      ldsm_x4(sa_hi + ca, ah[0], ah[1], ah[2], ah[3]);
      ldsm_x4(sa_lo + ca, al[0], al[1], al[2], al[3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // n-tiles 2 np, 2 np + 1
        uint32_t bh[4], bl[4];
        ldsm_x4(sw_hi + np * 16 * 128 + cw, bh[0], bh[1], bh[2], bh[3]);
        ldsm_x4(sw_lo + np * 16 * 128 + cw, bl[0], bl[1], bl[2], bl[3]);
        mma_bf16(acc[2 * np], ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
        mma_bf16(acc[2 * np + 1], ah[0], ah[1], ah[2], ah[3], bh[2], bh[3]);
        mma_bf16(acc[2 * np], al[0], al[1], al[2], al[3], bh[0], bh[1]);
        mma_bf16(acc[2 * np + 1], al[0], al[1], al[2], al[3], bh[2], bh[3]);
        mma_bf16(acc[2 * np], ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
        mma_bf16(acc[2 * np + 1], ah[0], ah[1], ah[2], ah[3], bl[2], bl[3]);
      }
    }
    __syncthreads();  // every warp is done with this stage before it is refilled
  }

  // ---- epilogue: bias, activation, fp32 or bf16 hi/lo split store (accumulator rows g / g + 8, columns 2 t, 2 t + 1)
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = m0 + wm * 16 + g + 8 * h;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n0 + wn * 32 + nt * 8 + 2 * t + e;
        if (row < M && n < N) {
          const float v = gs_act<EPI>(acc[nt][2 * h + e] + (bias ? bias[n] : 0.f));
          if (SPLIT_OUT) {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(C) + (size_t)row * ldc + n;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            cp[0] = hi;
            cp[ldc / 2] = __float2bfloat16_rn(v - __bfloat162float(hi));
          } else {
            reinterpret_cast<float*>(C)[(size_t)row * ldc + n] = v;
          }
        }
      }
    }
  }
}

template <int EPI>
int gs_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const float* bias, void* C, int ldc, int c_split,
              int M, int N, int K, cudaStream_t stream) {
  const dim3 grid(pfpp_cdiv(N, GS_BN), pfpp_cdiv(M, GS_BM));
  const int smem = 2 * GS_STAGE;
  if (c_split) {
    PFPP_ENSURE_SMEM((gemm_small_x3_kernel<EPI, true>), smem);
    gemm_small_x3_kernel<EPI, true><<<grid, GS_THREADS, smem, stream>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  } else {
    PFPP_ENSURE_SMEM((gemm_small_x3_kernel<EPI, false>), smem);
    gemm_small_x3_kernel<EPI, false><<<grid, GS_THREADS, smem, stream>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  }
  PFPP_RETURN_LAST();
}

}  // namespace

// Internal entry of pfpp_gemm_bf16x3 for small row counts (same arguments; no residual, no GEGLU).
int pfpp_gemm_small_x3(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int c_split, int M,
                       int N, int K, int epilogue, cudaStream_t stream) {
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  switch (epilogue) {
    case PFPP_EPI_NONE: return gs_launch<PFPP_EPI_NONE>(a, lda, w, ldw, bias, C, ldc, c_split, M, N, K, stream);
    case PFPP_EPI_RELU: return gs_launch<PFPP_EPI_RELU>(a, lda, w, ldw, bias, C, ldc, c_split, M, N, K, stream);
    case PFPP_EPI_GELU: return gs_launch<PFPP_EPI_GELU>(a, lda, w, ldw, bias, C, ldc, c_split, M, N, K, stream);
    case PFPP_EPI_SILU: return gs_launch<PFPP_EPI_SILU>(a, lda, w, ldw, bias, C, ldc, c_split, M, N, K, stream);
    default: return PFPP_EINVAL;
  }
}
