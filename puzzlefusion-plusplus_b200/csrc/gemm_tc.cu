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
#include "tc_common.cuh"
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

// OUT selects the output format of every kernel below: 0 = fp32, 1 = bf16, 2 = bf16 hi/lo SPLIT (the fp32 value v is
// stored as hi = bf16(v) at column c and lo = bf16(v - hi) at column c + lo_off of the same row: the operand format
// of the fp32-grade "bf16x3" contraction).  `parts` = 1: plain bf16 GEMM; 3: split operands, the K loop runs three
// times over (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) into the same fp32 accumulator (map_a2 / map_b2 address the lo
// halves): a*w = a_hi w_hi + a_lo w_hi + a_hi w_lo + O(2^-16 |a w|), every product exact in the fp32 accumulator.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h), lo = *reinterpret_cast<uint32_t*>(&l);
}

template <int EPI, int OUT>
__device__ __forceinline__ void epilogue_store(const uint32_t* v, int row, int nb, int M, int N, const float* __restrict__ bias,
                                               const float* residual, int ldr, void* Cout, int ldc, int lo_off);

template <int EPI, int OUT>
__global__ void __launch_bounds__(128)
    gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2, int parts,
                        const float* __restrict__ bias, const float* residual, int ldr, void* Cout, int ldc, int lo_off,
                        int M, int N, int K) {
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
  const int kb_part = (K + TC_BK - 1) / TC_BK;
  const int num_kb = kb_part * parts;

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
      const int part = kb / kb_part, kk = kb - part * kb_part;
      mbar_wait(bar_base + 8 * (TC_STAGES + s), (round & 1) ^ 1);  // slot free (passes immediately in round 0)
      const uint32_t full = bar_base + 8 * s;
      mbar_expect_tx(full, TC_STAGE_BYTES);
      const uint32_t sa = smem_base + s * TC_STAGE_BYTES;
      tma_load_2d(sa, part == 1 ? &map_a2 : &map_a, full, kk * TC_BK, m0);
      tma_load_2d(sa + TC_A_BYTES, part == 2 ? &map_b2 : &map_b, full, kk * TC_BK, n0);
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
    epilogue_store<EPI, OUT>(v, row, n0 + c * 32, M, N, bias, residual, ldr, Cout, ldc, lo_off);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_BN));
  }
}

// Coalesced epilogue for one 32-row x 32-column accumulator chunk owned by one warp.
// tcgen05.ld hands every thread ONE ROW (32 consecutive columns); storing that directly makes each warp
// store touch 32 different rows (16-byte pieces, 32 partial sectors) -- measured 2-3x slower than the main
// loop.  Instead the warp transposes the chunk through a private 4 KB shared-memory tile (16-byte chunks
// XOR-swizzled by row, conflict-free both ways) so that in the second phase 8 consecutive lanes cover one
// 128-byte row segment: every global load (residual) and store is a fully coalesced line segment.  Bias,
// activation / GEGLU and the fp32 residual add happen in the second phase on 4 consecutive columns per lane.
template <int EPI, int OUT>
__device__ __forceinline__ void epilogue_chunk_coalesced(const uint32_t* v, float* stage, int lane, int row0, int nb,
                                                         int M, int N, const float* __restrict__ bias,
                                                         const float* residual, int ldr, void* Cout, int ldc, int lo_off) {
  // phase 1: thread = row `lane` of the chunk; 8 x 16-byte chunks, chunk j stored at slot j ^ (lane & 7)
  uint4* srow = reinterpret_cast<uint4*>(stage) + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  // phase 2: lane -> (row r = lane/8 + 4 i, column group c = lane % 8 -> columns nb + 4c .. 4c+3)
  const int c = lane & 7;
  const int n = nb + 4 * c;
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) b4[j] = bias[n + j];
  }
  // residual rows are fetched up front (8 independent 16-byte loads in flight per lane): residual may alias
  // Cout, so the compiler cannot hoist these loads above the stores of earlier rows by itself
  float4 res[8];
  const bool vec_res = residual && (n + 4 <= N) && (ldr % 4) == 0;
  if (vec_res) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + (lane >> 3) + 4 * i;
      res[i] = row < M ? *reinterpret_cast<const float4*>(residual + (size_t)row * ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i;
    const int row = row0 + r;
    const uint4 raw = reinterpret_cast<const uint4*>(stage)[r * 8 + (c ^ (r & 7))];
    if (row >= M || n >= N) continue;
    float a[4] = {__uint_as_float(raw.x) + b4[0], __uint_as_float(raw.y) + b4[1], __uint_as_float(raw.z) + b4[2],
                  __uint_as_float(raw.w) + b4[3]};
    if (EPI == PFPP_EPI_GEGLU) {
      // interleaved (value, gate) pairs: columns (n, n+1) -> output column n/2, (n+2, n+3) -> n/2 + 1
      const float o0 = a[0] * (OUT == 1 ? gelu_tanh_fast(a[1]) : gelu_erf(a[1]));
      const float o1 = a[2] * (OUT == 1 ? gelu_tanh_fast(a[3]) : gelu_erf(a[3]));
      const size_t off = (size_t)row * ldc + (n >> 1);
      if (OUT == 2) {
        uint32_t hi, lo;
        split2(o0, o1, hi, lo);
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + off;
        if (n + 3 < N && (ldc % 2) == 0 && (lo_off % 2) == 0) {
          *reinterpret_cast<uint32_t*>(cp) = hi;
          *reinterpret_cast<uint32_t*>(cp + lo_off) = lo;
        } else {
          const __nv_bfloat162 h2 = *reinterpret_cast<__nv_bfloat162*>(&hi), l2 = *reinterpret_cast<__nv_bfloat162*>(&lo);
          if (n + 1 < N) cp[0] = h2.x, cp[lo_off] = l2.x;
          if (n + 3 < N) cp[1] = h2.y, cp[lo_off + 1] = l2.y;
        }
      } else if (OUT == 1) {
        if (n + 3 < N && (ldc % 2) == 0) {
          __nv_bfloat162 p = __floats2bfloat162_rn(o0, o1);
          *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(Cout) + off) = p;
        } else {
          if (n + 1 < N) reinterpret_cast<__nv_bfloat16*>(Cout)[off] = __float2bfloat16_rn(o0);
          if (n + 3 < N) reinterpret_cast<__nv_bfloat16*>(Cout)[off + 1] = __float2bfloat16_rn(o1);
        }
      } else {
        if (n + 1 < N) reinterpret_cast<float*>(Cout)[off] = o0;
        if (n + 3 < N) reinterpret_cast<float*>(Cout)[off + 1] = o1;
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = tc_act<EPI>(a[j]);
    const bool full = n + 4 <= N;
    if (vec_res) {
      a[0] += res[i].x, a[1] += res[i].y, a[2] += res[i].z, a[3] += res[i].w;
    } else if (residual) {
      const float* rp = residual + (size_t)row * ldr + n;
      for (int j = 0; j < 4; ++j)
        if (n + j < N) a[j] += rp[j];
    }
    if (OUT == 2) {
      __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)row * ldc + n;
      uint2 hi, lo;
      split2(a[0], a[1], hi.x, lo.x);
      split2(a[2], a[3], hi.y, lo.y);
      if (full && (ldc % 4) == 0 && (lo_off % 4) == 0) {
        *reinterpret_cast<uint2*>(cp) = hi;
        *reinterpret_cast<uint2*>(cp + lo_off) = lo;
      } else {
        const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&hi);
        const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&lo);
        for (int j = 0; j < 4; ++j)
          if (n + j < N) cp[j] = hp[j], cp[lo_off + j] = lp[j];
      }
    } else if (OUT == 1) {
      __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)row * ldc + n;
      if (full && (ldc % 4) == 0) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0], a[1]), p1 = __floats2bfloat162_rn(a[2], a[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(cp) = pk;
      } else {
        for (int j = 0; j < 4; ++j)
          if (n + j < N) cp[j] = __float2bfloat16_rn(a[j]);
      }
    } else {
      float* cp = reinterpret_cast<float*>(Cout) + (size_t)row * ldc + n;
      if (full && (ldc % 4) == 0) {
        *reinterpret_cast<float4*>(cp) = make_float4(a[0], a[1], a[2], a[3]);
      } else {
        for (int j = 0; j < 4; ++j)
          if (n + j < N) cp[j] = a[j];
      }
    }
  }
  __syncwarp();  // the staging tile is rewritten by the next chunk
}

// bf16 outputs: the warp converts its 32 rows x 64 output columns to bf16, lays them out as a 128B-swizzled
// [32 x 128 B] box in its private shared-memory tile and hands the box to the TMA engine
// (cp.async.bulk.tensor store): the SM issues no global stores at all and the write is full-line.
// NCH = accumulator chunks (32 columns each) feeding one 64-column output slab: 2, or 4 for GEGLU.
template <int EPI>
__device__ __forceinline__ void epilogue_slab_tma(uint32_t taddr, int acc_col0, uint8_t* stage, uint32_t stage_addr, int lane,
                                                  int row0, int out_col0, int n_acc0, int N, const float* __restrict__ bias,
                                                  const CUtensorMap* map_c) {
  constexpr int NCH = EPI == PFPP_EPI_GEGLU ? 4 : 2;
  uint32_t pk[32];  // 64 bf16 of this thread's row
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    uint32_t v[32];
    tmem_ld32(taddr + (uint32_t)(acc_col0 + ch * 32), v);
    const int nb = n_acc0 + ch * 32;
    float bl = 0.f;
    if (bias && nb + lane < N) bl = bias[nb + lane];  // one coalesced load; broadcast by shuffle below
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (EPI == PFPP_EPI_GEGLU) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float v0 = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j);
        const float g0 = __uint_as_float(v[j + 1]) + __shfl_sync(0xffffffffu, bl, j + 1);
        const float v1 = __uint_as_float(v[j + 2]) + __shfl_sync(0xffffffffu, bl, j + 2);
        const float g1 = __uint_as_float(v[j + 3]) + __shfl_sync(0xffffffffu, bl, j + 3);
        __nv_bfloat162 p = __floats2bfloat162_rn(v0 * gelu_tanh_fast(g0), v1 * gelu_tanh_fast(g1));
        pk[ch * 8 + (j >> 2)] = *reinterpret_cast<uint32_t*>(&p);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float a = tc_act<EPI>(__uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j));
        const float b = tc_act<EPI>(__uint_as_float(v[j + 1]) + __shfl_sync(0xffffffffu, bl, j + 1));
        __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
        pk[ch * 16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&p);
      }
    }
  }
  // the previous box of this warp must have been read out of shared memory before it is overwritten
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  uint4* srow = reinterpret_cast<uint4*>(stage) + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(map_c, stage_addr, out_col0, row0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// Shared epilogue for one 32-column chunk of one accumulator row (values v[32] straight from tcgen05.ld).
template <int EPI, int OUT>
__device__ __forceinline__ void epilogue_store(const uint32_t* v, int row, int nb, int M, int N, const float* __restrict__ bias,
                                               const float* residual, int ldr, void* Cout, int ldc, int lo_off) {
  if (!(row < M && nb < N)) return;
  float o[32];
  int n_out, col0;  // number of output values of this chunk and their first output column
  if (EPI == PFPP_EPI_GEGLU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const int n = nb + j;
      float val = __uint_as_float(v[j]) + ((bias && n < N) ? bias[n] : 0.f);
      float gate = __uint_as_float(v[j + 1]) + ((bias && n + 1 < N) ? bias[n + 1] : 0.f);
      o[j >> 1] = val * (OUT == 1 ? gelu_tanh_fast(gate) : gelu_erf(gate));
    }
    n_out = (N - nb) / 2 < 16 ? (N - nb) / 2 : 16;
    col0 = nb >> 1;
  } else {
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
    n_out = N - nb < 32 ? N - nb : 32;
    col0 = nb;
  }
  constexpr int NV = EPI == PFPP_EPI_GEGLU ? 16 : 32;
  const size_t off = (size_t)row * ldc + col0;
  if (OUT == 0) {
    float* cp = reinterpret_cast<float*>(Cout) + off;
    if (n_out == NV && (ldc % 4) == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else {
      for (int j = 0; j < NV; ++j)
        if (j < n_out) cp[j] = o[j];
    }
    return;
  }
  __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + off;
  const bool vec = n_out == NV && (ldc % 8) == 0 && (OUT == 1 || (lo_off % 8) == 0);
  if (vec) {
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
      uint4 hi, lo;
      if (OUT == 2) {
        split2(o[j], o[j + 1], hi.x, lo.x), split2(o[j + 2], o[j + 3], hi.y, lo.y);
        split2(o[j + 4], o[j + 5], hi.z, lo.z), split2(o[j + 6], o[j + 7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(cp + lo_off + j) = lo;
      } else {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[j], o[j + 1]), p1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), p3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
        hi.x = *reinterpret_cast<uint32_t*>(&p0), hi.y = *reinterpret_cast<uint32_t*>(&p1);
        hi.z = *reinterpret_cast<uint32_t*>(&p2), hi.w = *reinterpret_cast<uint32_t*>(&p3);
      }
      *reinterpret_cast<uint4*>(cp + j) = hi;
    }
  } else {
    for (int j = 0; j < NV; ++j) {
      if (j < n_out) {
        const __nv_bfloat16 h = __float2bfloat16_rn(o[j]);
        cp[j] = h;
        if (OUT == 2) cp[lo_off + j] = __float2bfloat16_rn(o[j] - __bfloat162float(h));
      }
    }
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
constexpr uint32_t T2_EPI_STAGE_BYTES = 8 * 4096;  // 8 epilogue warps x 4 KB transpose tile
constexpr uint32_t T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + T2_EPI_STAGE_BYTES + 1024 + 256;
constexpr uint32_t T2_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T2_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(map), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
constexpr int T2_THREADS = 64 + 256;

// MC = true: CTAs run as clusters of 2 that own vertically adjacent 128-row tiles of the same 256-column block.
// Each CTA loads its own A tile and one HALF of the shared W tile, multicast into both CTAs' shared memory
// (cp.async.bulk.tensor ... .multicast::cluster), so the W operand crosses the L2->SM fabric once per pair:
// operand bytes per tile drop from 393 KB to 262 KB.  Stage-release (tcgen05.commit) is multicast to both
// CTAs' empty barriers, since a stage is rewritten by both producers.
template <int EPI, bool OUT_BF16, bool MC>
__global__ void __launch_bounds__(T2_THREADS)
    gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const float* __restrict__ bias, const float* residual, int ldr, void* Cout, int ldc, int M,
                         int N, int K, int m_store) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_stage = smem_base + T2_STAGES * T2_STAGE_BYTES;
  const uint32_t bar_base = epi_stage + T2_EPI_STAGE_BYTES;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * T2_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * T2_STAGES;  // [2]
  const uint32_t bar_tempty = bar_tfull + 16;            // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + T2_BN - 1) / T2_BN, tiles_m = (M + TC_BM - 1) / TC_BM;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  // work items: single tiles, or (MC) vertical pairs of tiles handled by the 2 CTAs of a cluster
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  const int worker = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_items = (MC ? (tiles_m + 1) / 2 : tiles_m) * tiles_n;
  auto item_m0 = [&](int t) { return ((MC ? 2 * (t / tiles_n) + (int)crank : t / tiles_n)) * TC_BM; };
  auto item_n0 = [&](int t) { return (t % tiles_n) * T2_BN; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, MC ? 2 : 1);  // MC: released by the MMA issuers of both CTAs
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
  if (MC) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = worker; t < num_items; t += n_workers) {
        const int m0 = item_m0(t), n0 = item_n0(t);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % T2_STAGES, round = it / T2_STAGES;
          mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
          mbar_expect_tx(bar_full + 8 * s, T2_STAGE_BYTES);
          const uint32_t sa = smem_base + s * T2_STAGE_BYTES;
          tma_load_2d(sa, &map_a, bar_full + 8 * s, kb * TC_BK, m0);
          if (MC) {
            // my half of the W tile (128 of the 256 rows) -> both CTAs, signalling each CTA's own full barrier
            tma_load_2d_mc(sa + T2_A_BYTES + crank * (T2_B_BYTES / 2), &map_b, bar_full + 8 * s, kb * TC_BK,
                           n0 + (int)crank * (T2_BN / 2), (uint16_t)0x3);
          } else {
            tma_load_2d(sa + T2_A_BYTES, &map_b, bar_full + 8 * s, kb * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      for (int t = worker; t < num_items; t += n_workers, ++lt) {
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
          if (MC) umma_commit_mc(bar_empty + 8 * s, (uint16_t)0x3);
          else umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;      // which 128-column half of the tile this warp drains
    uint32_t lt = 0;
    for (int t = worker; t < num_items; t += n_workers, ++lt) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      const int m0 = item_m0(t), n0 = item_n0(t);
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + buf * T2_BN + ((uint32_t)(q * 32) << 16);
      float* stage = reinterpret_cast<float*>(smem_raw + (epi_stage - smem_u32(smem_raw)) + (warp - 2) * 4096);
#pragma unroll 1
      for (int c = chalf * (T2_BN / 64); c < (chalf + 1) * (T2_BN / 64); ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        epilogue_chunk_coalesced<EPI, OUT_BF16 ? 1 : 0>(v, stage, lane, m0 + q * 32, n0 + c * 32, m_store, N, bias, residual,
                                                        ldr, Cout, ldc, 0);
      }
      // this warp is done with the accumulator buffer: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MC) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it / signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile with
// ONE MMA stream issued by the leader CTA.  Each CTA stages only its 128 rows of A and its 128-row half of
// the W tile (32 KB per k-block instead of 48 KB), the tensor cores exchange the halves, and each CTA ends
// up with its own 128 x 256 accumulator in its own TMEM.  This halves the shared-memory traffic per FLOP --
// at 128 x 256 per single CTA the main loop is bound by shared-memory bandwidth (TMA fill + operand reads
// ~190 B/clk/SM against 128 B/clk), not by L2 or the tensor pipe.  6-stage ring (192 KB), accumulators
// double-buffered in TMEM, 8 epilogue warps per CTA.
//   full[s]   : leader CTA only; both CTAs' TMA loads complete_tx on it (cta_group::2 loads)
//   empty[s]  : per CTA; released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[b]  : per CTA; leader's commit multicast;   tempty[b] : leader only, 16 arrivals (8 warps x 2 CTAs)
// ------------------------------------------------------------------------------------------------
constexpr int T3_STAGES = 6;
constexpr uint32_t T3_HALF_BYTES = 128 * TC_BK * 2;            // one [128 x 64] operand half = 16 KB
constexpr uint32_t T3_STAGE_BYTES = 2 * T3_HALF_BYTES;          // A half + W half per CTA
constexpr uint32_t T3_SMEM_BYTES = T3_STAGES * T3_STAGE_BYTES + T2_EPI_STAGE_BYTES + 1024 + 256;
constexpr uint32_t T3_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
template <int EPI, int OUT>
__global__ void __launch_bounds__(T2_THREADS)
    gemm_bf16_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2, int parts,
                         const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias, const float* residual,
                         int ldr, void* Cout, int ldc, int lo_off, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_stage = smem_base + T3_STAGES * T3_STAGE_BYTES;
  const uint32_t bar_base = epi_stage + T2_EPI_STAGE_BYTES;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * T3_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * T3_STAGES;  // [2]
  const uint32_t bar_tempty = bar_tfull + 16;            // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int tiles_n = (N + 255) / 256, tiles_mp = (M + 255) / 256;
  const int num_items = tiles_mp * tiles_n;
  const int kb_part = (K + TC_BK - 1) / TC_BK;
  const int num_kb = kb_part * parts;  // parts = 3: (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) into one accumulator
  const int worker = (int)(blockIdx.x >> 1), n_workers = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < T3_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 16);  // 8 epilogue warps of each of the two CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = worker; t < num_items; t += n_workers) {
        const int m0 = (t / tiles_n) * 256 + (int)crank * 128, n0 = (t % tiles_n) * 256 + (int)crank * 128;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % T3_STAGES, round = it / T3_STAGES;
          mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
          const uint32_t lead_full = (bar_full + 8 * s) & PEER_BIT_MASK;
          if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * T3_STAGE_BYTES);  // both CTAs' halves
          const uint32_t sa = smem_base + s * T3_STAGE_BYTES;
          const int part = kb / kb_part, kk = kb - part * kb_part;
          tma_load_2d_2sm(sa, part == 1 ? &map_a2 : &map_a, lead_full, kk * TC_BK, m0);
          tma_load_2d_2sm(sa + T3_HALF_BYTES, part == 2 ? &map_b2 : &map_b, lead_full, kk * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      uint32_t it = 0, lt = 0;
      for (int t = worker; t < num_items; t += n_workers, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);  // both CTAs' epilogues have drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % T3_STAGES, round = it / T3_STAGES;
          mbar_wait(bar_full + 8 * s, round & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * T3_STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + T3_HALF_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) umma_bf16_2sm(acc, da + 2 * k, db + 2 * k, T3_IDESC, (kb | k) != 0);
          umma_commit_2sm_mc(bar_empty + 8 * s, (uint16_t)0x3);
        }
        umma_commit_2sm_mc(bar_tfull + 8 * buf, (uint16_t)0x3);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;      // which 128-column half of the tile this warp drains
    uint32_t lt = 0;
    for (int t = worker; t < num_items; t += n_workers, ++lt) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      const int m0 = (t / tiles_n) * 256 + (int)crank * 128, n0 = (t % tiles_n) * 256;
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
      float* stage = reinterpret_cast<float*>(smem_raw + (epi_stage - smem_u32(smem_raw)) + (warp - 2) * 4096);
      if (OUT == 1 && residual == nullptr) {
        // bf16 outputs (QKV, GEGLU hidden): shared-memory box + TMA store, 64 output columns per box
        const uint32_t stage_addr = epi_stage + (warp - 2) * 4096;
        if (EPI == PFPP_EPI_GEGLU) {
          epilogue_slab_tma<EPI>(taddr, chalf * 128, reinterpret_cast<uint8_t*>(stage), stage_addr, lane, m0 + q * 32,
                                 (n0 >> 1) + chalf * 64, n0 + chalf * 128, N, bias, &map_c);
        } else {
#pragma unroll 1
          for (int sl = 0; sl < 2; ++sl)
            epilogue_slab_tma<EPI>(taddr, chalf * 128 + sl * 64, reinterpret_cast<uint8_t*>(stage), stage_addr, lane,
                                   m0 + q * 32, n0 + chalf * 128 + sl * 64, n0 + chalf * 128 + sl * 64, N, bias, &map_c);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(bar_tempty + 8 * buf);
          else mbar_arrive_remote(bar_tempty + 8 * buf, 0);
        }
        continue;
      }
#pragma unroll 1
      for (int c = chalf * 4; c < (chalf + 1) * 4; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        epilogue_chunk_coalesced<EPI, OUT>(v, stage, lane, m0 + q * 32, n0 + c * 32, M, N, bias, residual, ldr, Cout, ldc,
                                           lo_off);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_tempty + 8 * buf);
        else mbar_arrive_remote(bar_tempty + 8 * buf, 0);
      }
    }
  }
  if (OUT == 1 && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // TMA stores done
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
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

template <int EPI, bool OUT_BF16, bool MC>
int launch_tc2_impl(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, const float* residual, int ldr,
                    void* C, int ldc, int M, int N, int K, cudaStream_t stream) {
  auto kern = gemm_bf16_tc2_kernel<EPI, OUT_BF16, MC>;
  PFPP_ENSURE_SMEM(kern, T2_SMEM_BYTES);
  const int tiles_n = pfpp_cdiv(N, T2_BN), tiles_m = pfpp_cdiv(M, TC_BM);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = T2_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (MC) {
    const int items = ((tiles_m + 1) / 2) * tiles_n;
    int pairs = items < sm_count() / 2 ? items : sm_count() / 2;
    pairs = pfpp_cdiv(items, pfpp_cdiv(items, pairs));  // no more pairs than the round count needs
    cfg.gridDim = dim3(2 * pairs);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int tiles = tiles_n * tiles_m;
    int ctas = tiles < sm_count() ? tiles : sm_count();
    ctas = pfpp_cdiv(tiles, pfpp_cdiv(tiles, ctas));
    cfg.gridDim = dim3(ctas);
  }
  static const int nostore = getenv("PFPP_GEMM_DEBUG_NOSTORE") ? 1 : 0;  // ablation knob (results are not written)
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, bias, residual, ldr, C, ldc, M, N, K, nostore ? 0 : M);
  if (e != cudaSuccess) return (int)e;
  PFPP_RETURN_LAST();
}

template <int EPI>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias, const float* residual, int ldr, void* C,
               int ldc, int c_bf16, int M, int N, int K, bool mc, cudaStream_t stream) {
  if (mc) {
    return c_bf16 ? launch_tc2_impl<EPI, true, true>(ma, mb, bias, residual, ldr, C, ldc, M, N, K, stream)
                  : launch_tc2_impl<EPI, false, true>(ma, mb, bias, residual, ldr, C, ldc, M, N, K, stream);
  }
  return c_bf16 ? launch_tc2_impl<EPI, true, false>(ma, mb, bias, residual, ldr, C, ldc, M, N, K, stream)
                : launch_tc2_impl<EPI, false, false>(ma, mb, bias, residual, ldr, C, ldc, M, N, K, stream);
}

struct GemmMaps {
  CUtensorMap a, b, a2, b2;
  int parts;
};

template <int EPI, int OUT>
int launch_tc3_impl(const GemmMaps& m, const float* bias, const float* residual, int ldr, void* C, int ldc, int lo_off,
                    int M, int N, int K, cudaStream_t stream) {
  auto kern = gemm_bf16_tc3_kernel<EPI, OUT>;
  CUtensorMap mc = m.a;  // placeholder when the TMA-store path is not taken
  if (OUT == 1 && residual == nullptr) {
    // output boxes of [32 rows x 64 bf16], 128B-swizzled like the staging tile
    auto fn = get_encode_fn();
    if (!fn) return PFPP_EUNSUPPORTED;
    const int n_out = EPI == PFPP_EPI_GEGLU ? N / 2 : N;
    cuuint64_t dims[2] = {(cuuint64_t)n_out, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldc * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&mc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PFPP_EINVAL;
  }
  PFPP_ENSURE_SMEM(kern, T3_SMEM_BYTES);
  const int items = pfpp_cdiv(M, 256) * pfpp_cdiv(N, 256);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = T3_SMEM_BYTES;
  cfg.stream = stream;
  // persistent CTA pairs: every pair runs ceil(items / pairs) tiles, so only as many pairs as that round count needs are
  // launched (228 tiles on 74 pairs = 4 rounds = 57 pairs): same latency, and the SMs that would idle through the last
  // round stay free for the kernels of the other streams
  const int max_pairs = sm_count() / 2;
  int pairs = items < max_pairs ? items : max_pairs;
  pairs = pfpp_cdiv(items, pfpp_cdiv(items, pairs));
  cfg.gridDim = dim3(2 * pairs);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m.a, m.b, m.a2, m.b2, m.parts, mc, bias, residual, ldr, C, ldc, lo_off, M,
                                     N, K);
  if (e != cudaSuccess) return (int)e;
  PFPP_RETURN_LAST();
}

template <int EPI>
int launch_tc3(const GemmMaps& m, const float* bias, const float* residual, int ldr, void* C, int ldc, int lo_off, int out,
               int M, int N, int K, cudaStream_t stream) {
  if (out == 2) return launch_tc3_impl<EPI, 2>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream);
  return out ? launch_tc3_impl<EPI, 1>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream)
             : launch_tc3_impl<EPI, 0>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream);
}

template <int EPI, int OUT>
int launch_tc_impl(const GemmMaps& m, const float* bias, const float* residual, int ldr, void* C, int ldc, int lo_off, int M,
                   int N, int K, cudaStream_t stream) {
  dim3 grid(pfpp_cdiv(N, TC_BN), pfpp_cdiv(M, TC_BM));
  PFPP_ENSURE_SMEM((gemm_bf16_tc_kernel<EPI, OUT>), TC_SMEM_BYTES);
  gemm_bf16_tc_kernel<EPI, OUT><<<grid, 128, TC_SMEM_BYTES, stream>>>(m.a, m.b, m.a2, m.b2, m.parts, bias, residual, ldr, C,
                                                                      ldc, lo_off, M, N, K);
  PFPP_RETURN_LAST();
}

template <int EPI>
int launch_tc(const GemmMaps& m, const float* bias, const float* residual, int ldr, void* C, int ldc, int lo_off, int out,
              int M, int N, int K, cudaStream_t stream) {
  if (out == 2) return launch_tc_impl<EPI, 2>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream);
  return out ? launch_tc_impl<EPI, 1>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream)
             : launch_tc_impl<EPI, 0>(m, bias, residual, ldr, C, ldc, lo_off, M, N, K, stream);
}

// out: 0 fp32, 1 bf16, 2 bf16 hi/lo split.  A_lo / W_lo != nullptr: split operands (3 passes over K).
int gemm_dispatch(const void* A, const void* A_lo, int lda, const void* W, const void* W_lo, int ldw, const float* bias,
                  const float* residual, int ldr, void* C, int ldc, int lo_off, int out, int M, int N, int K,
                  int epilogue, cudaStream_t stream) {
  GemmMaps m;
  m.parts = (A_lo && W_lo) ? 3 : 1;
  int rc = make_map(&m.a, A, M, K, lda, TC_BM);
  if (rc) return rc;
  static const int v2_mode = []() {
    // 0 = 128x128 kernel only, 1 = persistent 128x256, 2 = + cluster multicast of W, 3 = CTA-pair (cta_group::2)
    const char* e = getenv("PFPP_GEMM_V2");
    return e ? atoi(e) : 3;
  }();
  const bool wide = N >= T2_BN && M >= 2 * TC_BM && v2_mode != 0;  // persistent 128x256 kernel for the wide projections
  const bool mc = wide && v2_mode == 2 && (N % T2_BN) == 0 && m.parts == 1 && out != 2;
  const bool pair = wide && (v2_mode >= 3 || m.parts == 3 || out == 2) &&
                    (out != 1 || (ldc % 8) == 0);  // cta_group::2 kernel (bf16 C through TMA stores)
  const int box_b = wide ? ((mc || pair) ? T2_BN / 2 : T2_BN) : TC_BN;
  rc = make_map(&m.b, W, N, K, ldw, box_b);
  if (rc) return rc;
  m.a2 = m.a, m.b2 = m.b;
  if (m.parts == 3) {
    rc = make_map(&m.a2, A_lo, M, K, lda, TC_BM);
    if (rc) return rc;
    rc = make_map(&m.b2, W_lo, N, K, ldw, box_b);
    if (rc) return rc;
  }
#define PFPP_EPI_SWITCH(FN, ...)                                    \
  switch (epilogue) {                                               \
    case PFPP_EPI_NONE: return FN<PFPP_EPI_NONE>(__VA_ARGS__);      \
    case PFPP_EPI_RELU: return FN<PFPP_EPI_RELU>(__VA_ARGS__);      \
    case PFPP_EPI_GELU: return FN<PFPP_EPI_GELU>(__VA_ARGS__);      \
    case PFPP_EPI_SILU: return FN<PFPP_EPI_SILU>(__VA_ARGS__);      \
    case PFPP_EPI_GEGLU: return FN<PFPP_EPI_GEGLU>(__VA_ARGS__);    \
    default: return PFPP_EINVAL;                                    \
  }
  if (pair) { PFPP_EPI_SWITCH(launch_tc3, m, bias, residual, ldr, C, ldc, lo_off, out, M, N, K, stream) }
  if (wide && m.parts == 1 && out != 2) {
    PFPP_EPI_SWITCH(launch_tc2, m.a, m.b, bias, residual, ldr, C, ldc, out, M, N, K, mc, stream)
  }
  PFPP_EPI_SWITCH(launch_tc, m, bias, residual, ldr, C, ldc, lo_off, out, M, N, K, stream)
#undef PFPP_EPI_SWITCH
}

}  // namespace

int pfpp_gemm_small_x3(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int c_split, int M,
                       int N, int K, int epilogue, cudaStream_t stream);  // gemm_small.cu

extern "C" int pfpp_has_tensor_core_path(void) { return 1; }

extern "C" int pfpp_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual,
                              int ldr, void* C, int ldc, int c_bf16, int M, int N, int K, int epilogue,
                              cudaStream_t stream) {
  PFPP_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  PFPP_CHECK_ARG((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0);
  PFPP_CHECK_ARG((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0);
  if (M == 0) return PFPP_OK;
  return gemm_dispatch(A, nullptr, lda, W, nullptr, ldw, bias, residual, ldr, C, ldc, 0, c_bf16 ? 1 : 0, M, N, K, epilogue,
                       stream);
}

extern "C" int pfpp_gemm_bf16x3(const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual,
                                int ldr, void* C, int ldc, int c_split, int M, int N, int K, int epilogue,
                                cudaStream_t stream) {
  PFPP_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  PFPP_CHECK_ARG((K % 8) == 0 && (lda % 16) == 0 && (ldw % 16) == 0 && lda >= 2 * K && ldw >= 2 * K);
  PFPP_CHECK_ARG((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0);
  PFPP_CHECK_ARG(!c_split || (ldc % 2) == 0);
  if (M == 0) return PFPP_OK;
  // a few hundred rows (the output heads: M = fragments): 32 x 64 tiles on warp-level MMAs instead of one or two
  // 256 x 256 CTA-pair tiles whose fixed pipeline cost dominates (gemm_small.cu)
  if (M <= 1024 && residual == nullptr && epilogue != PFPP_EPI_GEGLU)
    return pfpp_gemm_small_x3(A, lda, W, ldw, bias, C, ldc, c_split, M, N, K, epilogue, stream);
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  return gemm_dispatch(a, a + lda / 2, lda, w, w + ldw / 2, ldw, bias, residual, ldr, C, ldc, ldc / 2, c_split ? 2 : 0, M, N,
                       K, epilogue, stream);
}
