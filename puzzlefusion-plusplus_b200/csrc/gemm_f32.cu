// fp32 SIMT GEMM with fused epilogues -- the "parity mode" contraction engine.
//
//   C[M, N'] = epilogue(A[M, K] * W[N, K]^T + bias[N]) (+ residual)
//
// A and W are row-major with K contiguous (nn.Linear / 1x1-conv weight layout).  fp32 FFMA with
// fp32 accumulation reproduces the reference's fp32 cuBLAS/cuDNN arithmetic up to summation order,
// which is what the discrete stages downstream (VQ argmin, verifier threshold) need for parity.
// The bf16 tcgen05 engine in gemm_tc.cu is the fast path with the same epilogues.
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile split as 2x2 blocks of 4x4 so
// that shared-memory reads are conflict-free LDS.128; operands are stored k-major in smem.
#include "common.cuh"
#include "../../include/pfpp.h"

#define GBM 128
#define GBN 128
#define GBK 16
#define GPAD 4

template <int EPI>
__device__ __forceinline__ float apply_act(float v) {
  if (EPI == PFPP_EPI_RELU) return fmaxf(v, 0.f);
  if (EPI == PFPP_EPI_GELU) return gelu_erf(v);
  if (EPI == PFPP_EPI_SILU) return silu(v);
  return v;
}

template <int EPI>
__global__ void __launch_bounds__(256)
    gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                    const float* __restrict__ bias, const float* residual, int ldr, float* C, int ldc, int M, int N,
                    int K) {
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Ws[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;

  // global -> smem mapping: each thread moves 2 float4 (along K) of A and of W per k-tile
  const int lrow = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12
  float4 ra[2], rw[2];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      ra[h] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x, As[buf][lk + 1][r] = ra[h].y, As[buf][lk + 2][r] = ra[h].z,
      As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x, Ws[buf][lk + 1][r] = rw[h].y, Ws[buf][lk + 2][r] = rw[h].z,
      Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + GBK - 1) / GBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * GBK);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      int gn = n0 + jb * 64 + tx * 4;
      if (EPI == PFPP_EPI_GEGLU) {
        // columns interleaved on the host: (2j, 2j+1) = (value j, gate j)  ->  out column j
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          int n = gn + j;
          if (n + 1 < N) {
            float v = acc[i][jb * 4 + j] + (bias ? bias[n] : 0.f);
            float g = acc[i][jb * 4 + j + 1] + (bias ? bias[n + 1] : 0.f);
            float o = v * gelu_erf(g);
            size_t off = (size_t)gm * ldc + (n >> 1);
            C[off] = o;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int n = gn + j;
          if (n < N) {
            float v = acc[i][jb * 4 + j] + (bias ? bias[n] : 0.f);
            v = apply_act<EPI>(v);
            if (residual) v += residual[(size_t)gm * ldr + n];
            C[(size_t)gm * ldc + n] = v;
          }
        }
      }
    }
  }
}

// Small-M variant: 32x32x16 CTA tile, 64 threads, 4x4 register micro-tile.  Used when the 128x128 grid would
// leave most of the 148 SMs idle (the denoiser's output heads: M = number of fragments).  With so few rows the
// GEMM is bound by the latency of its serial K loop, not by FLOPs; small tiles put 4-10 CTAs on every SM so that
// the loops of different CTAs hide each other's global-load latency.  Every output element is still one
// sequential fmaf chain over k = 0..K-1, so both variants produce bit-identical results.
#define SBM 32
#define SBN 32
#define SBK 64  // deep k-tiles: 16 independent 16-byte loads per thread in flight, 4x fewer serial iterations
template <int EPI>
__global__ void __launch_bounds__(64)
    gemm_f32_small_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                          const float* __restrict__ bias, const float* residual, int ldr, float* C, int ldc, int M,
                          int N, int K) {
  __shared__ __align__(16) float As[2][SBK][SBM + GPAD];
  __shared__ __align__(16) float Ws[2][SBK][SBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int lrow = tid >> 4;     // 0..3 (+4 h)
  const int lk = (tid & 15) * 4;  // 0,4,..,60: a quarter-warp reads 256 contiguous bytes of one row
  float4 ra[8], rw[8];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int gm = m0 + lrow + 4 * h, gn = n0 + lrow + 4 * h, gk = k0 + lk;
      ra[h] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int r = lrow + 4 * h;
      As[buf][lk + 0][r] = ra[h].x, As[buf][lk + 1][r] = ra[h].y, As[buf][lk + 2][r] = ra[h].z, As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x, Ws[buf][lk + 1][r] = rw[h].y, Ws[buf][lk + 2][r] = rw[h].z, Ws[buf][lk + 3][r] = rw[h].w;
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (K + SBK - 1) / SBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * SBK);
#pragma unroll 16
    for (int k = 0; k < SBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    const int gn = n0 + tx * 4;
    if (EPI == PFPP_EPI_GEGLU) {
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int n = gn + j;
        if (n + 1 < N) {
          const float v = acc[i][j] + (bias ? bias[n] : 0.f);
          const float g = acc[i][j + 1] + (bias ? bias[n + 1] : 0.f);
          C[(size_t)gm * ldc + (n >> 1)] = v * gelu_erf(g);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = gn + j;
        if (n < N) {
          float v = acc[i][j] + (bias ? bias[n] : 0.f);
          v = apply_act<EPI>(v);
          if (residual) v += residual[(size_t)gm * ldr + n];
          C[(size_t)gm * ldc + n] = v;
        }
      }
    }
  }
}

extern "C" int pfpp_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                             const float* residual, int ldr, float* C, int ldc, int M, int N, int K, int epilogue,
                             cudaStream_t stream) {
  PFPP_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  PFPP_CHECK_ARG((K % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0);
  PFPP_CHECK_ARG((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0);
  if (M == 0) return PFPP_OK;
  dim3 grid(pfpp_cdiv(N, GBN), pfpp_cdiv(M, GBM));
  dim3 sgrid(pfpp_cdiv(N, SBN), pfpp_cdiv(M, SBM));
  const bool small = grid.x * grid.y < 74;  // the big tiles would leave more than half of the SMs idle
#define PFPP_GEMM_CASE(E)                                                                                      \
  case E:                                                                                                      \
    if (small)                                                                                                 \
      gemm_f32_small_kernel<E><<<sgrid, 64, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K);  \
    else                                                                                                       \
      gemm_f32_kernel<E><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K);      \
    break;
  switch (epilogue) {
    PFPP_GEMM_CASE(PFPP_EPI_NONE)
    PFPP_GEMM_CASE(PFPP_EPI_RELU)
    PFPP_GEMM_CASE(PFPP_EPI_GELU)
    PFPP_GEMM_CASE(PFPP_EPI_SILU)
    PFPP_GEMM_CASE(PFPP_EPI_GEGLU)
    default:
      return PFPP_EINVAL;
  }
#undef PFPP_GEMM_CASE
  PFPP_RETURN_LAST();
}
