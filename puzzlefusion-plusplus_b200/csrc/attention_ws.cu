// Denoiser attention on the 5th-gen tensor cores: warp-specialised, two query tiles in flight per CTA.
//
// One CTA = (segment, head).  A segment is a run of packed token rows of the bf16 QKV activation matrix
// [M, 3C] (q | k | v, head h at columns h*64):
//   * global attention (attention.py:84, key mask of denoiser_transformer.py:158-166): segment = one object's
//     valid fragments (<= 512 tokens), every query sees every key of the segment; padded / merged-away
//     fragments are simply not packed, which is what the reference's key mask does;
//   * local attention (attention.py:79 with self_mask): block-diagonal `block`-token blocks (25 = one fragment);
//     a tile holds floor(128/block) whole blocks (125 tokens) and query tile i only needs key tile i.
//
// Roles (352 threads):
//   warps 0-3   softmax group 0: query tiles 0, 2      one query row per thread (= its TMEM lane)
//   warps 4-7   softmax group 1: query tiles 1, 3
//   warp  8/9   MMA issuer of group 0 / 1 (one elected lane each): S = Q K^T (M128 N64 K64) two steps ahead of
//               the softmax, O += P V (M128 N64 K64) as soon as a P panel is ready
//   warp  10    TMA producer: K/V tiles once per CTA, Q tiles into one slot per group
// A *step* is 64 keys.  Per group S and P are double-buffered, so the tensor pipe computes S(n+1) and
// P(n-1) V while the group exponentiates step n, and the two groups overlap each other's gaps (the MUFU
// pipe bounds this kernel at head_dim 64: 128 x 64 exps per 2 x 128 tensor cycles).
//
// TMEM (512 columns): S_w[buf] at (2w + buf) * 64 (fp32) | O_w at 256 + 64 w.
// Softmax is online over the steps with a lazy rescale of O (only when the running maximum grows by more
// than 2^8: the fp32 accumulator and the bf16 probabilities keep their relative precision under a common
// scale); P goes to shared memory as a K-major SWIZZLE_128B UMMA operand panel, V is consumed MN-major
// straight from its TMA tile.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr int AW_D = 64;
constexpr int AW_BQ = 128;
constexpr int AW_TK = 128;                        // keys per K/V tile
constexpr int AW_STEP = 64;                       // keys per softmax step
constexpr int AW_MAXT = 4;                        // tiles (query tiles == key tiles) per segment
constexpr uint32_t AW_TILE = 128 * 64 * 2;        // 16 KB: one [128 x 64] bf16 operand tile
constexpr uint32_t AW_OFF_Q = 0;                  // one slot per group
constexpr uint32_t AW_OFF_K = AW_OFF_Q + 2 * AW_TILE;
constexpr uint32_t AW_OFF_V = AW_OFF_K + AW_MAXT * AW_TILE;
constexpr uint32_t AW_OFF_P = AW_OFF_V + AW_MAXT * AW_TILE;  // per group: 2 panels of [128 rows x 64 keys]
constexpr uint32_t AW_OFF_BAR = AW_OFF_P + 2 * 2 * AW_TILE;
constexpr uint32_t AW_SMEM_BYTES = AW_OFF_BAR + 512 + 1024;
constexpr int AW_THREADS = 352;
constexpr uint32_t AW_COL_S = 0, AW_COL_O = 256;
constexpr float AW_LAZY = 8.0f;  // log2 units

// barrier slots (8 bytes each); per-group barriers are indexed [w], per-group-per-buffer ones [2 w + buf]
enum { B_Q = 0, B_QFREE = 2, B_K = 4, B_V = 8, B_S = 12, B_SFREE = 16, B_P = 20, B_PV = 24, B_TOK = 28, B_KFREE = 30, B_VFREE = 34, B_COUNT = 38 };

// non-blocking probe (try_wait may suspend the thread for a system-dependent time, which is wrong for the MMA
// issuer that polls several barriers round-robin)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major SWIZZLE_128B operand tile ([rows x 64] bf16, 128 B per row, 8-row groups 1024 B apart)
// MN-major SWIZZLE_128B operand: 64 contiguous MN elements per 128 B row, K rows 128 B apart,
// groups of 8 K rows 1024 B apart (SBO); a single 64-wide MN block, so LBO is not exercised.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// D=f32, A=B=bf16, M=128, N=64.  S: both operands K-major.  O: A (P) K-major, B (V) MN-major (bit 16).
constexpr uint32_t AW_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AW_STEP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t AW_IDESC_O =
    (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AW_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// Row maximum over the visible keys [klo, khi) of one 64-key step; invisible scores are overwritten with -inf
// so that the exponential pass needs no predicate (ex2(-inf) = 0).  In the masked variant only the 8-key
// chunks [clo, chi) (warp-uniform: the union of the windows of the warp's 32 rows) are touched at all.
template <bool FULL>
__device__ __forceinline__ float row_max_mask(uint32_t* s, int klo, int khi, int clo, int chi) {
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains
  if (FULL) {
#pragma unroll
    for (int j = 0; j < AW_STEP; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(s[j]));
  } else {
#pragma unroll
    for (int c = 0; c < AW_STEP / 8; ++c) {
      if (c >= clo && c < chi) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = c * 8 + i;
          const float v = (j >= klo && j < khi) ? __uint_as_float(s[j]) : -INFINITY;
          s[j] = __float_as_uint(v);
          mx[i & 3] = fmaxf(mx[i & 3], v);
        }
      }
    }
  }
  return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// exponentials -> bf16 P panel (K-major SWIZZLE_128B: 16-byte chunk index ^ (row & 7)); returns the row sum
template <bool FULL>
__device__ __forceinline__ float exp_store(const uint32_t* s, float scale_log2e, float nm, uint8_t* prow, int row, int clo,
                                           int chi) {
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int c = 0; c < AW_STEP / 8; ++c) {
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (FULL || (c >= clo && c < chi)) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int key = c * 8 + j;
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[key]), scale_log2e, nm));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[key + 1]), scale_log2e, nm));
        __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
        if (FULL) {
          l0 += p0;
          l1 += p1;
        } else {
          // block-diagonal mode: sum the bf16-rounded probabilities the PV product actually sees.  A sum of a
          // few 8-bit-mantissa values is exact in fp32, so the result does not depend on where inside the
          // 128-key tile the block lies (a batch of objects == the same objects run one by one, bit for bit)
          l0 += __low2float(b2);
          l1 += __high2float(b2);
        }
      }
    }
    *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  return l0 + l1;
}

#define AW_TRACE(slot)                                                       \
  do {                                                                       \
    if (trace) trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 48 + (slot)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(AW_THREADS, 1)
    attention_ws_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ seg_start,
                        const int* __restrict__ seg_len, int C, float scale_log2e, int block, int tile_stride,
                        __nv_bfloat16* __restrict__ out, int ldo, long long* trace) {
  const int seg = blockIdx.y, h = blockIdx.x;
  if (threadIdx.x == 0 && trace) {
    AW_TRACE(0);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 48 + 12] = (long long)gt;
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 48 + 14] = smid;
  }
  const int len = seg_len[seg], st = seg_start[seg];
  if (len <= 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + AW_OFF_BAR;
  const uint32_t tmem_slot = bars + B_COUNT * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Short segments (<= 4 tiles, grid.z == 1): K/V of the whole segment stay resident, group w runs query tiles
  // w and w + 2.  Long segments (global attention only, grid.z > 1): CTA z runs query tiles 2z and 2z + 1 (one per
  // group) and streams the segment's key tiles through the 4 K and 4 V slots as a ring (slot = tile & 3).
  const bool long_mode = gridDim.z > 1;
  const int nkt = (len + tile_stride - 1) / tile_stride;       // tiles of the segment
  const int nt = long_mode ? nkt : min(AW_MAXT, nkt);          // query tiles == key tiles that are processed
  const int jq_step = long_mode ? (1 << 20) : 2;               // long mode: one query tile per group
  const int jq_base = long_mode ? 2 * (int)blockIdx.z : 0;     // group w starts at query tile jq_base + w
  if (jq_base >= nt) return;
  // steps (64 keys) of key tile kb that hold at least one key of the segment
  auto tile_steps = [&](int kb) { return (min(tile_stride, len - kb * tile_stride) + AW_STEP - 1) / AW_STEP; };

  if (threadIdx.x == 0) {
    const uint32_t n_issuers = jq_base + 1 < nt ? 2u : 1u;  // consumers of a K/V ring slot (long mode)
    for (int i = 0; i < B_COUNT; ++i)
      mbar_init(bars + 8 * i, i >= B_KFREE ? n_issuers : (((i >= B_SFREE && i < B_PV) || i >= B_TOK) ? 128u : 1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(base_ptr + AW_OFF_BAR + B_COUNT * 8);
  if (threadIdx.x == 0) AW_TRACE(1);

  if (warp == 10) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int qc = h * AW_D, kc = C + h * AW_D, vc = 2 * C + h * AW_D;
      if (long_mode) {
        for (int w = 0; w < 2 && jq_base + w < nt; ++w) {
          mbar_expect_tx(bars + 8 * (B_Q + w), AW_TILE);
          tma_load_2d(base + AW_OFF_Q + w * AW_TILE, &map, bars + 8 * (B_Q + w), qc, st + (jq_base + w) * tile_stride);
        }
        for (int kb = 0; kb < nt; ++kb) {
          const int slot = kb & 3;
          // a slot is refilled once every S (K) / PV (V) product of both groups that read tile kb - 4 has completed
          if (kb >= 4) mbar_wait(bars + 8 * (B_KFREE + slot), ((kb >> 2) - 1) & 1);
          mbar_expect_tx(bars + 8 * (B_K + slot), AW_TILE);
          tma_load_2d(base + AW_OFF_K + slot * AW_TILE, &map, bars + 8 * (B_K + slot), kc, st + kb * tile_stride);
          if (kb >= 4) mbar_wait(bars + 8 * (B_VFREE + slot), ((kb >> 2) - 1) & 1);
          mbar_expect_tx(bars + 8 * (B_V + slot), AW_TILE);
          tma_load_2d(base + AW_OFF_V + slot * AW_TILE, &map, bars + 8 * (B_V + slot), vc, st + kb * tile_stride);
        }
        AW_TRACE(2);
      } else {
      mbar_expect_tx(bars + 8 * (B_Q + 0), AW_TILE);
      tma_load_2d(base + AW_OFF_Q, &map, bars + 8 * (B_Q + 0), qc, st);
      mbar_expect_tx(bars + 8 * (B_K + 0), AW_TILE);
      tma_load_2d(base + AW_OFF_K, &map, bars + 8 * (B_K + 0), kc, st);
      if (nt > 1) {
        mbar_expect_tx(bars + 8 * (B_Q + 1), AW_TILE);
        tma_load_2d(base + AW_OFF_Q + AW_TILE, &map, bars + 8 * (B_Q + 1), qc, st + tile_stride);
      }
      // local mode: group 1 starts on key tile 1, so fetch the tiles in the order K0 K1 V0 V1 ...
      if (block && nt > 1) {
        mbar_expect_tx(bars + 8 * (B_K + 1), AW_TILE);
        tma_load_2d(base + AW_OFF_K + AW_TILE, &map, bars + 8 * (B_K + 1), kc, st + tile_stride);
      }
      mbar_expect_tx(bars + 8 * (B_V + 0), AW_TILE);
      tma_load_2d(base + AW_OFF_V, &map, bars + 8 * (B_V + 0), vc, st);
      for (int kb = 1; kb < nt; ++kb) {
        if (!(block && kb == 1)) {
          mbar_expect_tx(bars + 8 * (B_K + kb), AW_TILE);
          tma_load_2d(base + AW_OFF_K + kb * AW_TILE, &map, bars + 8 * (B_K + kb), kc, st + kb * tile_stride);
        }
        mbar_expect_tx(bars + 8 * (B_V + kb), AW_TILE);
        tma_load_2d(base + AW_OFF_V + kb * AW_TILE, &map, bars + 8 * (B_V + kb), vc, st + kb * tile_stride);
      }
      AW_TRACE(2);
      for (int w = 0; w < 2; ++w) {
        if (w + 2 < nt) {
          mbar_wait(bars + 8 * (B_QFREE + w), 0);  // every S product that reads Q slot w has completed
          mbar_expect_tx(bars + 8 * (B_Q + w), AW_TILE);
          tma_load_2d(base + AW_OFF_Q + w * AW_TILE, &map, bars + 8 * (B_Q + w), qc, st + (w + 2) * tile_stride);
        }
      }
      }  // short mode
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ MMA issuer of group w
    const int w = warp - 8;
    if (lane == 0 && jq_base + w < nt) {
      // S cursor (tile jq, key tile kb, half hf) runs two steps ahead of the PV cursor
      int s_jq = jq_base + w, s_kb = block ? s_jq : 0, s_hf = 0, s_n = 0, s_job = 0;
      bool s_more = true;
      auto issue_s = [&]() {
        const int kb_last = block ? s_jq : nt - 1;
        if (s_n >= 2) mbar_wait(bars + 8 * (B_SFREE + 2 * w + (s_n & 1)), ((s_n >> 1) - 1) & 1);
        if (s_hf == 0) {
          if (s_kb == (block ? s_jq : 0)) mbar_wait(bars + 8 * (B_Q + w), s_job & 1);
          mbar_wait(bars + 8 * (B_K + (s_kb & 3)), (s_kb >> 2) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dq = desc_kmajor(base + AW_OFF_Q + w * AW_TILE);
        const uint64_t dk = desc_kmajor(base + AW_OFF_K + (s_kb & 3) * AW_TILE + s_hf * (AW_TILE / 2));
        const uint32_t d = tmem + AW_COL_S + (2 * w + (s_n & 1)) * AW_STEP;
#pragma unroll
        for (int k = 0; k < AW_D / 16; ++k) umma_bf16(d, dq + 2 * k, dk + 2 * k, AW_IDESC_S, k != 0);
        umma_commit(bars + 8 * (B_S + 2 * w + (s_n & 1)));
        if (w == 0 && s_n == 0) AW_TRACE(3);
        ++s_n;
        if (s_hf + 1 < tile_steps(s_kb)) {
          ++s_hf;
        } else {
          if (long_mode) umma_commit(bars + 8 * (B_KFREE + (s_kb & 3)));  // this group is done with K tile s_kb
          if (s_kb < kb_last) {
            ++s_kb;
            s_hf = 0;
          } else if (s_jq + jq_step < nt) {
            umma_commit(bars + 8 * (B_QFREE + w));  // last S of this query tile: its Q slot may be refilled
            s_jq += jq_step;
            ++s_job;
            s_kb = block ? s_jq : 0;
            s_hf = 0;
          } else {
            s_more = false;
          }
        }
      };
      issue_s();
      if (s_more) issue_s();
      int p_n = 0;
      for (int jq = jq_base + w; jq < nt; jq += jq_step) {
        const int kb_first = block ? jq : 0, kb_last = block ? jq : nt - 1;
        for (int kb = kb_first; kb <= kb_last; ++kb) {
          const int nh = tile_steps(kb);
          for (int hf = 0; hf < nh; ++hf, ++p_n) {
            const int b = p_n & 1;
            mbar_wait(bars + 8 * (B_P + 2 * w + b), (p_n >> 1) & 1);
            if (hf == 0) mbar_wait(bars + 8 * (B_V + (kb & 3)), (kb >> 2) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t pa = base + AW_OFF_P + (2 * w + b) * AW_TILE;
            const uint32_t va = base + AW_OFF_V + (kb & 3) * AW_TILE + hf * (AW_TILE / 2);
#pragma unroll
            for (int k = 0; k < AW_STEP / 16; ++k)
              umma_bf16(tmem + AW_COL_O + w * AW_D, desc_kmajor(pa) + 2 * k, desc_mnmajor(va + k * 2048), AW_IDESC_O,
                        !(kb == kb_first && hf == 0) || k != 0);
            umma_commit(bars + 8 * (B_PV + 2 * w + b));
            if (long_mode && hf + 1 == nh) umma_commit(bars + 8 * (B_VFREE + (kb & 3)));  // done with V tile kb
            if (s_more) issue_s();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int w = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t o_addr = lane_addr + AW_COL_O + w * AW_D;
    int n = 0;  // steps consumed by this group so far
    // Global mode: the two groups take turns on the MUFU pipe (B_TOK + w is completed by the OTHER group when it has
    // finished an exponential phase): without the hand-over they drift into lockstep, exponentiate at the same
    // time at half rate each and then idle in their TMEM-load / max / barrier phases at the same time as well.
    // Group 0 runs tiles 0,2 and group 1 tiles 1,3 with the same number of steps per tile, so group 1 never has
    // more steps than group 0.
    const bool take_turns = (block == 0) && jq_base + 1 < nt;
    int steps_other = 0;  // total steps of the other group
    if (take_turns) {
      int per_tile = 0;
      for (int kb = 0; kb < nt; ++kb) per_tile += tile_steps(kb);
      steps_other = long_mode ? per_tile : per_tile * ((nt - (1 - w) + 1) / 2);
    }
    const int blk_lo = block ? (row / block) * block : 0;  // first key (tile-relative) of this row's block
    for (int jq = jq_base + w; jq < nt; jq += jq_step) {
      const int tile_len = min(tile_stride, len - jq * tile_stride);
      const int kb_first = block ? jq : 0, kb_last = block ? jq : nt - 1;
      float m_used = -INFINITY, l = 0.f;
      for (int kb = kb_first; kb <= kb_last; ++kb) {
        const int nh = tile_steps(kb);
        for (int hf = 0; hf < nh; ++hf, ++n) {
          const int b = n & 1;
          // keys of this step visible to this query row: [klo, khi) relative to the step
          int klo = 0, khi = min(AW_STEP, len - kb * tile_stride - hf * AW_STEP);
          if (block) {
            const int lo = blk_lo;
            klo = max(lo - hf * AW_STEP, 0);
            khi = min(min(lo + block, tile_len) - hf * AW_STEP, AW_STEP);
          }
          mbar_wait(bars + 8 * (B_S + 2 * w + b), (n >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t s[AW_STEP];
          const uint32_t s_addr = lane_addr + AW_COL_S + (2 * w + b) * AW_STEP;
          tmem_ld32(s_addr, s);
          tmem_ld32(s_addr + 32, s + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(bars + 8 * (B_SFREE + 2 * w + b));  // this S buffer may be overwritten (by step n + 2)
          const bool full = (klo == 0 && khi == AW_STEP);  // warp-uniform in the global mode
          int clo = 0, chi = AW_STEP / 8;
          if (!full) {
            clo = __reduce_min_sync(0xffffffffu, klo) >> 3;
            chi = (__reduce_max_sync(0xffffffffu, max(khi, klo)) + 7) >> 3;
          }
          float mx = full ? row_max_mask<true>(s, klo, khi, 0, 8) : row_max_mask<false>(s, klo, khi, clo, chi);
          if (block && hf == 0 && nh == 2) {
            // a block may straddle the two 64-key steps of its tile: take the row maximum over the whole block
            // up front (the second step's scores already sit in the other S buffer), so that the probabilities
            // do not depend on the position of the block inside the tile
            mbar_wait(bars + 8 * (B_S + 2 * w + (b ^ 1)), ((n + 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int klo2 = max(blk_lo - AW_STEP, 0), khi2 = min(min(blk_lo + block, tile_len) - AW_STEP, AW_STEP);
            const int hlo = __reduce_min_sync(0xffffffffu, klo2) >> 5;
            const int hhi = (__reduce_max_sync(0xffffffffu, max(khi2, klo2)) + 31) >> 5;
#pragma unroll 1
            for (int hh = hlo; hh < min(hhi, 2); ++hh) {
              uint32_t s2[32];
              tmem_ld32(lane_addr + AW_COL_S + (2 * w + (b ^ 1)) * AW_STEP + hh * 32, s2);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (hh * 32 + j >= klo2 && hh * 32 + j < khi2) mx = fmaxf(mx, __uint_as_float(s2[j]));
            }
          }
          const float mxs = mx * scale_log2e;
          float alpha = 1.f;
          bool rescale = false;
          if (m_used == -INFINITY) {
            m_used = mxs;  // nothing accumulated for this row yet (O row and l are exactly 0)
          } else if (mxs > m_used + AW_LAZY) {
            alpha = ex2_approx(m_used - mxs);
            m_used = mxs;
            rescale = true;
          }
          const bool first = (kb == kb_first && hf == 0);
          if (!first && __any_sync(0xffffffffu, rescale)) {
            // every PV product of this tile issued so far (steps < n) must have landed in O_w
            mbar_wait(bars + 8 * (B_PV + 2 * w + (b ^ 1)), ((n - 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              tmem_ld32(o_addr + hh * 32, o);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
              tmem_st32(o_addr + hh * 32, o);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          l *= alpha;
          if (n >= 2) mbar_wait(bars + 8 * (B_PV + 2 * w + b), ((n >> 1) - 1) & 1);  // P panel b drained by step n - 2
          uint8_t* prow = base_ptr + AW_OFF_P + (2 * w + b) * AW_TILE + row * 128;
          const float nm = (m_used == -INFINITY) ? 0.f : -m_used;
          if (take_turns) {
            // group 0 waits for group 1's phase n-1, group 1 for group 0's phase n
            const int need = w == 0 ? n - 1 : n;
            if (need >= 0 && need < steps_other) mbar_wait(bars + 8 * (B_TOK + w), need & 1);
          }
          l += full ? exp_store<true>(s, scale_log2e, nm, prow, row, 0, 8)
                    : exp_store<false>(s, scale_log2e, nm, prow, row, clo, chi);
          if (take_turns) mbar_arrive(bars + 8 * (B_TOK + (w ^ 1)));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor-core reads
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(bars + 8 * (B_P + 2 * w + b));
        }
      }
      // ---- epilogue of this query tile: O / l -> bf16 -> global
      mbar_wait(bars + 8 * (B_PV + 2 * w + ((n - 1) & 1)), ((n - 1) >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t o[64];
      tmem_ld32(o_addr, o);
      tmem_ld32(o_addr + 32, o + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < tile_len) {
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = out + (size_t)(st + jq * tile_stride + row) * ldo + h * AW_D;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          uint4 pk;
          __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
          __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
          __nv_bfloat162 a2 = __floats2bfloat162_rn(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
          __nv_bfloat162 a3 = __floats2bfloat162_rn(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t*>(&a0), pk.y = *reinterpret_cast<uint32_t*>(&a1);
          pk.z = *reinterpret_cast<uint32_t*>(&a2), pk.w = *reinterpret_cast<uint32_t*>(&a3);
          *reinterpret_cast<uint4*>(orow + j) = pk;
        }
      }
    }
  }
  if (threadIdx.x == 0) AW_TRACE(10);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && trace) {
    AW_TRACE(11);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 48 + 13] = (long long)gt;
  }
  if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
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

static int attention_launch(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                            int n_segments, int max_len, int heads, int block, void* out, int ldo, long long* trace,
                            cudaStream_t stream) {
  PFPP_CHECK_ARG(qkv && seg_start && seg_len && out && heads > 0 && C == heads * AW_D);
  PFPP_CHECK_ARG(block >= 0 && block <= AW_TK);
  const int tile_stride = block ? (AW_TK / block) * block : AW_TK;
  // segments of more than 4 tiles: global attention only (K/V streamed through the 4-slot ring, 2 query tiles per CTA)
  const bool long_segments = max_len > AW_MAXT * tile_stride;
  PFPP_CHECK_ARG((!long_segments || block == 0) && (ld % 8) == 0 && (ldo % 8) == 0 && ((uintptr_t)qkv & 15) == 0 &&
                 ((uintptr_t)out & 15) == 0);
  if (n_segments == 0 || max_len <= 0 || M == 0) return PFPP_OK;
  auto fn = encode_fn();
  if (!fn) return PFPP_EUNSUPPORTED;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * C), (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)AW_D, 128};
  cuuint32_t estr[2] = {1, 1};
  if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return PFPP_EINVAL;
  PFPP_ENSURE_SMEM(attention_ws_kernel, AW_SMEM_BYTES);
  dim3 grid(heads, n_segments, long_segments ? (pfpp_cdiv(max_len, AW_TK) + 1) / 2 : 1);
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)AW_D);
  attention_ws_kernel<<<grid, AW_THREADS, AW_SMEM_BYTES, stream>>>(map, seg_start, seg_len, C, scale_log2e, block,
                                                                  tile_stride, (__nv_bfloat16*)out, ldo, trace);
  PFPP_RETURN_LAST();
}

extern "C" int pfpp_attention_tc(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                                 int n_segments, int max_len, int heads, int block, void* out, int ldo,
                                 cudaStream_t stream) {
  return attention_launch(qkv, M, ld, C, seg_start, seg_len, n_segments, max_len, heads, block, out, ldo, nullptr, stream);
}

// Debug variant: per-CTA clock64 / globaltimer stamps of the pipeline phases into trace[n_ctas][16]
// (tools/bench_attention.py --trace); not part of the product path.
extern "C" int pfpp_attention_tc_trace(const void* qkv, long long M, int ld, int C, const int* seg_start,
                                       const int* seg_len, int n_segments, int max_len, int heads, int block, void* out,
                                       int ldo, long long* trace, cudaStream_t stream) {
  return attention_launch(qkv, M, ld, C, seg_start, seg_len, n_segments, max_len, heads, block, out, ldo, trace, stream);
}
