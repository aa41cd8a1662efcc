// Residual projection fused with the following (Ada)LayerNorm (attention.py:75-92 through diffusers Attention.to_out /
// FeedForward.net[2], followed by MyAdaLayerNorm :5-25 or nn.LayerNorm):
//
//     h   <- h + A W^T + bias                      (fp32 residual stream, in place)
//     ln  <- LN(h) * (1 + scale) + shift  |  LN(h) * gamma + beta        (bf16: the A operand of the next projection)
//
// for the N = C = 512 projections (attention out-proj, GEGLU FF2).  One CTA owns 128 FULL rows: its fp32 accumulator
// tile [128 x 512] fills the SM's tensor memory (512 columns), so the row statistics never leave the CTA, the
// residual stream is read once and written once per sub-layer and the separate LayerNorm pass (one more read of h and
// 18 launches per DDPM step) disappears.
//
// CTAs run as pairs (tcgen05 cta_group::2, cluster of 2): the pair computes 256 rows; each CTA stages its 128 rows of A
// and HALF of the W k-block (2 x [128 x 64]: its 128-row share of the column halves [0,256) and [256,512)), the
// tensor cores exchange the halves -- 48 KB instead of 80 KB of operand fill per CTA and k-block, which is what
// bounds this shape (every CTA streams the whole weight matrix).
//
//   warp 0      TMA producer (both CTAs; all loads signal the leader's `full` barrier), 3-stage ring
//   warp 1      MMA issuer (leader CTA only): per k16 step two M 256 x N 256 instructions (the two column halves)
//   warps 2-9   epilogue: warp -> TMEM lane quarter (warp % 4) and column half ((warp - 2) / 4); thread = row.
//               pass 1, per 32-column chunk: the residual tile [32 rows x 32 cols] arrives by TMA (128B-swizzled, a
//               ring of 5 tiles per warp: two filled under the main loop, three in the drained operand ring);
//               accumulator + bias + residual -> shifted row sums -> value parked back in TMEM and written over the
//               residual in its tile -> TMA store of h; (mean, M2) of the two column halves combined (Chan) through
//               shared memory;
//               pass 2, per 64-column slab: normalise + modulate (next TMEM load in flight) -> bf16 -> swizzled box in one
//               of the warp's two dedicated tiles -> TMA store.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int GL_C = 512;                               // row width (d_model)
constexpr int GL_BM = 128, GL_BK = 64, GL_STAGES = 3;
constexpr uint32_t GL_BOX_BYTES = 128 * GL_BK * 2;      // one [128 x 64] bf16 box = 16 KB
constexpr uint32_t GL_STAGE_BYTES = 3 * GL_BOX_BYTES;   // A rows + two W shares
// residual tiles [32 rows x 32 fp32] (4 KB, SWIZZLE_128B) of the epilogue warps: per warp a ring of GL_RT tiles, the first
// two in a dedicated area (filled while the main loop runs), the others in the operand ring once the MMAs are done
constexpr int GL_RT = 5;
constexpr uint32_t GL_OFF_EPI = GL_STAGES * GL_STAGE_BYTES;          // 8 warps x 2 tiles
constexpr uint32_t GL_OFF_STAT = GL_OFF_EPI + 8 * 2 * 4096;          // [2 halves][128 rows] (mean, M2)
constexpr uint32_t GL_OFF_TAB = GL_OFF_STAT + 2 * 128 * 8;           // bias | multiplier | offset, 512 floats each
constexpr uint32_t GL_OFF_BAR = GL_OFF_TAB + 3 * GL_C * 4;           // full[3] empty[3] acc | tmem slot | res[8][GL_RT]
constexpr uint32_t GL_SMEM_BYTES = GL_OFF_BAR + 512 + 1024;
static_assert(8 * (GL_RT - 2) * 4096 <= GL_STAGES * GL_STAGE_BYTES, "late residual tiles live in the operand ring");
constexpr int GL_THREADS = 64 + 256;
// D = f32, A = B = bf16, both K-major, M = 256 (128 per CTA), N = 256
constexpr uint32_t GL_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct LnParams {
  const float* bias;       // [512] or nullptr
  float* h;                // [M, 512] fp32 residual stream, updated in place
  const float* mod;        // AdaLN table [G, 1024] (scale | shift) or nullptr
  const int* row_group;    // row r uses mod[row_group[r / rows_per_group]]
  int rows_per_group;
  const float* gamma;      // affine LayerNorm (mod == nullptr)
  const float* beta;
  int M, K;
};

__global__ void __launch_bounds__(GL_THREADS, 1)
    gemm_res_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_h,
                       const LnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + GL_OFF_BAR, bar_empty = bar_full + 8 * GL_STAGES, bar_acc = bar_empty + 8 * GL_STAGES;
  const uint32_t tmem_slot = bar_acc + 8;
  const uint32_t bar_res = tmem_slot + 8;  // [8 warps][GL_RT]: a residual tile has landed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int m0 = (int)(blockIdx.x >> 1) * 2 * GL_BM + (int)crank * GL_BM;
  const int num_kb = (p.K + GL_BK - 1) / GL_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GL_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    for (int i = 0; i < 8 * GL_RT; ++i) mbar_init(bar_res + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(bp + GL_OFF_BAR + 16 * GL_STAGES + 8);

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % GL_STAGES;
        const uint32_t round = kb / GL_STAGES;
        mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
        const uint32_t lead_full = (bar_full + 8 * s) & PEER_BIT_MASK;
        if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * GL_STAGE_BYTES);  // both CTAs' boxes
        const uint32_t sa = base + s * GL_STAGE_BYTES;
        tma_load_2d_2sm(sa, &map_a, lead_full, kb * GL_BK, m0);
        tma_load_2d_2sm(sa + GL_BOX_BYTES, &map_w, lead_full, kb * GL_BK, (int)crank * 128);
        tma_load_2d_2sm(sa + 2 * GL_BOX_BYTES, &map_w, lead_full, kb * GL_BK, 256 + (int)crank * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % GL_STAGES;
        const uint32_t round = kb / GL_STAGES;
        mbar_wait(bar_full + 8 * s, round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * GL_STAGE_BYTES;
        const uint64_t da = desc_kmajor(sa);
        const uint64_t db0 = desc_kmajor(sa + GL_BOX_BYTES), db1 = desc_kmajor(sa + 2 * GL_BOX_BYTES);
#pragma unroll
        for (int k = 0; k < GL_BK / 16; ++k) {
          umma_bf16_2sm(tmem, da + 2 * k, db0 + 2 * k, GL_IDESC, (kb | k) != 0);
          umma_bf16_2sm(tmem + 256, da + 2 * k, db1 + 2 * k, GL_IDESC, (kb | k) != 0);
        }
        umma_commit_2sm_mc(bar_empty + 8 * s, (uint16_t)0x3);
      }
      umma_commit_2sm_mc(bar_acc, (uint16_t)0x3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3, ch = (warp - 2) >> 2;
    const int row = m0 + q * 32 + lane;                       // this thread's row
    const int col0 = ch * 256;                                // this warp's column half
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + col0;
    const int ew = warp - 2;                                  // epilogue warp 0..7
    // residual tile t of this warp: t < 2 dedicated, t >= 2 in the (by then idle) operand ring
    auto tile_off = [&](int t) -> uint32_t {
      return t < 2 ? GL_OFF_EPI + (uint32_t)(ew * 2 + t) * 4096u : (uint32_t)(ew * (GL_RT - 2) + (t - 2)) * 4096u;
    };
    const uint32_t my_res = bar_res + 8 * (ew * GL_RT);
    float* tab = reinterpret_cast<float*>(bp + GL_OFF_TAB);   // [0,512) bias, [512,1024) multiplier, [1024,1536) offset
    // the residual h[32 rows x 32 cols] of chunk c arrives by TMA, 128B-swizzled: the thread that owns a row reads it
    // with conflict-free 16-byte loads, overwrites it in place with the new h and the TMA engine stores the tile back --
    // no transposes through the LSU, no global loads / stores issued by the SM
    auto fetch = [&](int c) {  // lane 0
      const int t = c % GL_RT;
      mbar_expect_tx(my_res + 8 * t, 4096);
      tma_load_2d(base + tile_off(t), &map_h, my_res + 8 * t, col0 + c * 32, m0 + q * 32);
    };
    if (lane == 0) {
      fetch(0);
      fetch(1);
    }
    // tables of the CTA's first row's modulation group (the DDPM step: every row of the batch shares one timestep)
    const float* mrow0 = nullptr;
    if (p.mod) mrow0 = p.mod + (size_t)p.row_group[(m0 < p.M ? m0 : p.M - 1) / p.rows_per_group] * 2 * GL_C;
    for (int i = threadIdx.x - 64; i < GL_C; i += 256) {
      tab[i] = p.bias ? p.bias[i] : 0.f;
      tab[GL_C + i] = p.mod ? 1.0f + mrow0[i] : p.gamma[i];
      tab[2 * GL_C + i] = p.mod ? mrow0[GL_C + i] : p.beta[i];
    }
    asm volatile("bar.sync 5, 256;" ::: "memory");            // tables written
    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (lane == 0) {  // every MMA has completed: the operand ring is free
#pragma unroll
      for (int c = 2; c < GL_RT; ++c) fetch(c);
    }
    // ---- pass 1: h_new = acc + bias + h ; shifted row sums (shift = the row's first value of this half)
    float x0 = 0.f, sd = 0.f, sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      const int n0 = col0 + c * 32;
      const int t = c % GL_RT;
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      uint4* tile = reinterpret_cast<uint4*>(bp + tile_off(t)) + lane * 8;  // this thread's row
      mbar_wait(my_res + 8 * t, (c / GL_RT) & 1);
      uint4 r4[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r4[j] = tile[j ^ (lane & 7)];
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(tab + n0 + 4 * j);
        v[4 * j] = __float_as_uint((__uint_as_float(v[4 * j]) + b4.x) + __uint_as_float(r4[j].x));
        v[4 * j + 1] = __float_as_uint((__uint_as_float(v[4 * j + 1]) + b4.y) + __uint_as_float(r4[j].y));
        v[4 * j + 2] = __float_as_uint((__uint_as_float(v[4 * j + 2]) + b4.z) + __uint_as_float(r4[j].z));
        v[4 * j + 3] = __float_as_uint((__uint_as_float(v[4 * j + 3]) + b4.w) + __uint_as_float(r4[j].w));
      }
      if (c == 0) x0 = __uint_as_float(v[0]);
      {  // four independent chains per sum (a single one is 32 dependent adds per chunk)
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(v[j]) - x0;
          s4[j & 3] += d;
          q4[j & 3] = fmaf(d, d, q4[j & 3]);
        }
        sd += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        sq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
      }
      tmem_st32(taddr + c * 32, v);  // parked for pass 2
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (m0 + q * 32 < p.M) tma_store_2d(&map_h, base + tile_off(t), n0, m0 + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (c >= 2 && c + GL_RT - 2 < 8) {
          // the store of chunk c - 2 has finished reading its tile: refill it with the residual of chunk c + GL_RT - 2
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          fetch(c + GL_RT - 2);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // ---- row statistics: this half's (mean, M2) -> shared memory -> combined with the other half (Chan et al.)
    const float mean_h = x0 + sd * (1.0f / 256.0f);
    const float m2_h = sq - sd * sd * (1.0f / 256.0f);
    float2* stat = reinterpret_cast<float2*>(bp + GL_OFF_STAT);
    stat[ch * 128 + q * 32 + lane] = make_float2(mean_h, m2_h);
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // the two warps that share these 32 rows
    const float2 o = stat[(ch ^ 1) * 128 + q * 32 + lane];
    const float delta = o.x - mean_h;
    const float mean = mean_h + 0.5f * delta;
    const float var = fmaxf(m2_h + o.y + delta * delta * 128.0f, 0.f) * (1.0f / GL_C);
    const float rstd = rsqrtf(var + 1e-5f);
    // ---- pass 2: normalise + modulate -> bf16 -> TMA store, 64 columns at a time
    const float* mrow = nullptr;
    if (p.mod) mrow = row < p.M ? p.mod + (size_t)p.row_group[row / p.rows_per_group] * 2 * GL_C : mrow0;
    // the shared tables hold the first row's group: valid for this warp when all of its rows are in that group
    const bool uniform = __all_sync(0xffffffffu, mrow == mrow0);
    // 32 columns of this thread's row -> 16 packed bf16 pairs
    auto norm_chunk = [&](const uint32_t* v, int n, uint32_t* pk) {
      if (uniform) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 a4 = *reinterpret_cast<const float4*>(tab + GL_C + n + 4 * j);
          const float4 b4 = *reinterpret_cast<const float4*>(tab + 2 * GL_C + n + 4 * j);
          const float y0 = (__uint_as_float(v[4 * j]) - mean) * rstd * a4.x + b4.x;
          const float y1 = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * a4.y + b4.y;
          const float y2 = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * a4.z + b4.z;
          const float y3 = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * a4.w + b4.w;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
          pk[2 * j] = *reinterpret_cast<uint32_t*>(&p0);
          pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&p1);
        }
      } else {  // rows of different timesteps inside one warp (pfpp_denoiser_forward with mixed steps)
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float y0 = (__uint_as_float(v[j]) - mean) * rstd * (1.0f + mrow[n + j]) + mrow[GL_C + n + j];
          const float y1 = (__uint_as_float(v[j + 1]) - mean) * rstd * (1.0f + mrow[n + j + 1]) + mrow[GL_C + n + j + 1];
          __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
    };
    // the TMEM load of the next 32 columns is in flight while the current ones are normalised; the output boxes
    // alternate between the warp's two dedicated tiles so that a slab only waits for the store two slabs back
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
#pragma unroll 1
    for (int sl = 0; sl < 4; ++sl) {
      const int n0 = col0 + sl * 64;
      uint32_t pk[32];
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(taddr + sl * 64 + 32, vb);
      norm_chunk(va, n0, pk);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (sl + 1 < 4) tmem_ld32(taddr + (sl + 1) * 64, va);
      norm_chunk(vb, n0 + 32, pk + 16);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the box of slab sl - 2 has been read
      __syncwarp();
      uint4* srow = reinterpret_cast<uint4*>(bp + tile_off(sl & 1)) + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (m0 + q * 32 < p.M) tma_store_2d(&map_o, base + tile_off(sl & 1), n0, m0 + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // no CTA exits while its peer's tensor cores can still read its operand tiles
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

PFN_cuTensorMapEncodeTiled_v12000 gl_encode_fn() {
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

// bf16 row-major [rows, cols], leading dimension ld (elements); box [box_rows x 64], 128B swizzle
int gl_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  auto fn = gl_encode_fn();
  if (!fn) return PFPP_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PFPP_OK : PFPP_EINVAL;
}

}  // namespace

extern "C" int pfpp_gemm_res_ln(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int M, int K,
                                const float* mod, const int* row_group, int rows_per_group, const float* gamma,
                                const float* beta, void* ln_out, cudaStream_t stream) {
  PFPP_CHECK_ARG(A && W && h && ln_out && M >= 0 && K > 0 && (K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0);
  PFPP_CHECK_ARG((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0 && (((uintptr_t)ln_out) & 15) == 0 &&
                 (((uintptr_t)h) & 15) == 0);
  PFPP_CHECK_ARG(mod ? (row_group && rows_per_group > 0) : (gamma && beta));
  if (M == 0) return PFPP_OK;
  CUtensorMap ma, mw, mo, mh;
  int rc = gl_map(&ma, A, M, K, lda, GL_BM);
  if (rc) return rc;
  rc = gl_map(&mw, W, GL_C, K, ldw, 128);
  if (rc) return rc;
  rc = gl_map(&mo, ln_out, M, GL_C, GL_C, 32);
  if (rc) return rc;
  {  // the residual stream as [32 x 32] fp32 boxes (128-byte rows, 128B swizzle)
    auto fn = gl_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)GL_C, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)GL_C * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PFPP_EINVAL;
  }
  LnParams p{bias, h, mod, row_group, rows_per_group, gamma, beta, M, K};
  PFPP_ENSURE_SMEM(gemm_res_ln_kernel, GL_SMEM_BYTES);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(GL_THREADS);
  cfg.dynamicSmemBytes = GL_SMEM_BYTES;
  cfg.stream = stream;
  cfg.gridDim = dim3(2 * pfpp_cdiv(M, 2 * GL_BM));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_res_ln_kernel, ma, mw, mo, mh, p);
  if (e != cudaSuccess) return (int)e;
  PFPP_RETURN_LAST();
}
