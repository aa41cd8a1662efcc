// Fused PointNet++ set-abstraction MLP on the 5th-gen tensor cores:
//   grouping gather -> 3 x (1x1 conv + folded BN + ReLU) -> max over nsample      in ONE kernel
// (utils/pn2_utils.py:139-148 and :209-214).  Activations never leave the SM: the gathered rows and
// the two hidden layers live in shared memory as bf16 UMMA operand tiles, accumulators in TMEM.
//
// One CTA owns SUB x 128 grouped rows (SUB = 2 for levels 2/3: both sub-tiles consume every weight stage, which
// halves the L2 weight traffic that bounds the kernel; 128 rows = 4 groups of 32 samples or 2 groups of 64).
//   layers 0,1  D[rows x C] = X[rows x K] . W^T   rows on the M axis; the epilogue thread owns one row,
//               adds the bias, applies ReLU and writes the bf16 row into the next layer's K-major
//               SWIZZLE_128B operand tile (16-byte stores).
//   layer 2     D[C x rows] = W[C x K] . X^T      channels on the M axis (128 per MMA block); the
//               epilogue thread owns one channel, so the max over a group's nsample rows is a plain
//               register reduction over TMEM columns; bias + ReLU commute with the max.
// CTA = 4*SUB compute warps (gather, MMA issue by lane 0 of warp 1, epilogues; warp w works on sub-tile w/4 and
// TMEM lane quarter w%4) + one TMA weight-producer warp.  Weight tiles [128 x 64] stream through a ring (full/empty mbarriers) that runs ahead across
// layer boundaries; X tiles are written by the CTA's own threads (generic proxy) and published to the
// tensor core with fence.proxy.async.
//
// Layer 0 is split: the D feature columns go through the tensor cores as 64-wide K panels (aligned 16-byte
// gathers); the three centroid-offset columns (dx,dy,dz) go through ONE extra K=16 MMA step on a small
// un-swizzled operand that holds each offset and each weight as a bf16 hi/lo pair:
//     w.d ~= w_hi d_hi + w_hi d_lo + w_lo d_hi        (9 of the 16 K slots; the products are exact in the fp32
// accumulator, the dropped w_lo d_lo term is 2^-16 relative) -- fp32-grade geometry terms without the 256
// broadcast LDS.128 per row that an epilogue FMA formulation costs (that epilogue was bound by shared-memory
// return bandwidth).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfpp.h"

namespace {

constexpr uint32_t SF_TILE = 128 * 64 * 2;  // one [128 x 64] bf16 operand panel / weight stage = 16 KB

// K-major SWIZZLE_128B tile: [rows x 64] bf16, 128 B per row, 8-row groups 1024 B apart
// K-major, no swizzle: core matrix = 8 rows x 16 B stored contiguously; [rows x 16] bf16 operand laid out as
// (row / 8) * 256 + (k / 8) * 128 + (row % 8) * 16 + (k % 8) * 2   ->  LBO (next core matrix in K) = 128 B,
// SBO (next 8-row group) = 256 B
__device__ __forceinline__ uint64_t desc_kmajor_nosw(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t xyzoff(int row, int kchunk) {
  return (uint32_t)(row >> 3) * 256u + (uint32_t)kchunk * 128u + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// D=f32, A=B=bf16, both K-major, M=128, N=128
constexpr uint32_t SF_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// byte offset of element (row, k) inside a K-major SW128 operand made of 64-wide panels
__device__ __forceinline__ uint32_t xoff(int row, int k) {
  return (uint32_t)(k >> 6) * SF_TILE + (uint32_t)row * 128u + (uint32_t)((((k & 63) >> 3) ^ (row & 7)) << 4) +
         (uint32_t)((k & 7) << 1);
}

struct SaParams {
  const float* xyz;        // [K, N, 3] source points of this level
  const float* new_xyz;    // [K, S, 3] centroids
  const __nv_bfloat16* feats;  // [K, N, D] previous level features (nullptr when D == 0)
  const int* gidx;         // [K, S, NS]
  const float* wxyz;       // [C1, 4] fp32: layer-0 weights of the (dx,dy,dz) columns
  const float* b0;
  const float* b1;
  const float* b2;
  __nv_bfloat16* out;      // [K*S, C3]
  int N, S;
  long long groups;        // K * S
  long long* trace;        // debug: per-CTA clock64 stamps [grid][16] (nullptr in the product path)
};

template <int NS, int D, int C1, int C2, int C3, int STAGES, int SUB>
struct SaCfg {
  static constexpr int ROWS = 128 * SUB;                    // grouped rows per CTA (SUB sub-tiles of 128)
  static constexpr int THREADS = ROWS + 32;                 // compute warps + one TMA producer warp
  static constexpr int K0 = D;                              // layer-0 tensor-core K (feature columns only)
  static constexpr int P0 = D / 64;                         // panels of X0 (0 for the first level)
  static constexpr int P1 = (C1 + 63) / 64, P2 = (C2 + 63) / 64;
  static constexpr int NB1 = (C1 + 127) / 128, NB2 = (C2 + 127) / 128, MB3 = C3 / 128;
  static constexpr int X_PANELS = (P0 > P1 ? (P0 > P2 ? P0 : P2) : (P1 > P2 ? P1 : P2));
  static constexpr uint32_t X_BYTES = X_PANELS * SF_TILE;   // one sub-tile's operand buffer (X0 -> X1 -> X2 in place)
  static constexpr uint32_t OFF_X = 0;
  static constexpr uint32_t OFF_W = SUB * X_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_W + STAGES * SF_TILE;
  static constexpr uint32_t OFF_ROWS = OFF_BAR + 128;        // [ROWS] int32 gather offsets
  static constexpr uint32_t OFF_XYZ = OFF_ROWS + 4 * ROWS;   // [SUB][128 rows x 16] bf16 hi/lo offsets (un-swizzled operand)
  static constexpr int WXYZ_ROWS = ((C1 + 127) / 128) * 128;
  static constexpr uint32_t OFF_WXYZ = OFF_XYZ + SUB * 4096; // [WXYZ_ROWS x 16] bf16 hi/lo weights of (dx,dy,dz)
  static constexpr uint32_t OFF_CONST = OFF_WXYZ + WXYZ_ROWS * 32;  // b0 [C1] | b1 [C2] | b2 [C3]  (fp32)
  static constexpr uint32_t CONST_BYTES = (C1 + C2 + C3) * 4;
  static constexpr uint32_t SMEM = OFF_CONST + CONST_BYTES + 1024 /*align slack*/;
  static constexpr int CW = (NB1 > NB2 ? NB1 : NB2) * 128;  // TMEM columns per sub-tile in layers 0/1
  static constexpr int TMEM_NEED = SUB * (CW > 128 ? CW : 128);
  static constexpr int TMEM_COLS = TMEM_NEED > 256 ? 512 : (TMEM_NEED > 128 ? 256 : 128);
  // CTAs that can share an SM (shared memory and TMEM bound): the register budget is set to allow them
  static constexpr int BY_SMEM = (227 * 1024) / (int)(SMEM + 1024);
  static constexpr int BY_TMEM = 512 / TMEM_COLS;
  static constexpr int MINB = BY_SMEM < BY_TMEM ? (BY_SMEM < 1 ? 1 : BY_SMEM) : BY_TMEM;
  static constexpr int GS = 128 / NS;                       // groups per sub-tile
  static constexpr int G = GS * SUB;                        // groups per CTA
};

template <int NS, int D, int C1, int C2, int C3, int STAGES, int SUB>
__global__ void __launch_bounds__(128 * SUB + 32, SaCfg<NS, D, C1, C2, C3, STAGES, SUB>::MINB)
    sa_fused_kernel(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                    const __grid_constant__ CUtensorMap map_w2, const SaParams p) {
  using Cfg = SaCfg<NS, D, C1, C2, C3, STAGES, SUB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + Cfg::OFF_BAR;             // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * STAGES;          // [STAGES]
  const uint32_t bar_acc = bar_empty + 8 * STAGES;           // accumulators of the current layer complete
  const uint32_t tmem_slot = bar_acc + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp >> 2, wq = warp & 3;                  // sub-tile and TMEM lane quarter of a compute warp
  const long long n_tiles = (p.groups + Cfg::G - 1) / Cfg::G;  // persistent CTA: tiles blockIdx.x, + gridDim.x, ...

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(bp + Cfg::OFF_BAR + 16 * STAGES + 8);

  // ---------------- weight producer: one elected thread, runs ahead over all three layers ----------
  // stage order == consumption order of the MMA issuer below; every stage feeds all SUB sub-tiles
  if (warp == 4 * SUB) {
    if (lane != 0) return;
    int it = 0;
    auto push = [&](const CUtensorMap* m, int kcol, int row) {
      const int s = it % STAGES;
      const uint32_t round = it / STAGES;
      mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
      mbar_expect_tx(bar_full + 8 * s, SF_TILE);
      tma_load_2d(base + Cfg::OFF_W + s * SF_TILE, m, bar_full + 8 * s, kcol, row);
      ++it;
    };
    // RESIDENT: a tile's weights fill the ring exactly once (level 1: 2 tiles), so they are loaded for the first
    // tile only and stay in shared memory for every later tile of this CTA
    constexpr int TOT = (D > 0 ? Cfg::NB1 * Cfg::P0 : 0) + Cfg::NB2 * Cfg::P1 + Cfg::MB3 * Cfg::P2;
    constexpr bool RESIDENT = TOT == STAGES;
    for (long long tile = blockIdx.x; tile < (RESIDENT ? (long long)blockIdx.x + 1 : n_tiles); tile += gridDim.x) {
      if (D > 0)
        for (int nb = 0; nb < Cfg::NB1; ++nb)
          for (int kp = 0; kp < Cfg::P0; ++kp) push(&map_w0, kp * 64, nb * 128);
      for (int nb = 0; nb < Cfg::NB2; ++nb)
        for (int kp = 0; kp < Cfg::P1; ++kp) push(&map_w1, kp * 64, nb * 128);
      for (int mb = 0; mb < Cfg::MB3; ++mb)
        for (int kp = 0; kp < Cfg::P2; ++kp) push(&map_w2, kp * 64, mb * 128);
    }
    return;  // the ring drains on its own; shared memory stays live until the compute warps exit
  }

#define SA_BAR() asm volatile("bar.sync 1, %0;" ::"n"(Cfg::ROWS) : "memory")
#define SA_TRACE(slot)                                                                  \
  do {                                                                                 \
    if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + (slot)] = clock64(); \
  } while (0)
  SA_TRACE(0);

  {  // per-channel constants -> shared memory once per CTA
    float* cst = reinterpret_cast<float*>(bp + Cfg::OFF_CONST);
    for (int i = threadIdx.x; i < C1; i += Cfg::ROWS) cst[i] = p.b0[i];
    for (int i = threadIdx.x; i < C2; i += Cfg::ROWS) cst[C1 + i] = p.b1[i];
    for (int i = threadIdx.x; i < C3; i += Cfg::ROWS) cst[C1 + C2 + i] = p.b2[i];
    // weight side of the (dx,dy,dz) K=16 step: k = [x_hi y_hi z_hi | x_hi y_hi z_hi | x_lo y_lo z_lo | 0 ...]
    for (int n = threadIdx.x; n < Cfg::WXYZ_ROWS; n += Cfg::ROWS) {
      float w[3] = {0.f, 0.f, 0.f}, hi[3], lo[3];
      if (n < C1) w[0] = p.wxyz[n * 4], w[1] = p.wxyz[n * 4 + 1], w[2] = p.wxyz[n * 4 + 2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        hi[i] = __bfloat162float(__float2bfloat16_rn(w[i]));
        lo[i] = w[i] - hi[i];
      }
      uint8_t* wb = bp + Cfg::OFF_WXYZ;
      *reinterpret_cast<uint4*>(wb + xyzoff(n, 0)) =
          make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[0]), pack_bf16(hi[1], hi[2]), pack_bf16(lo[0], lo[1]));
      *reinterpret_cast<uint4*>(wb + xyzoff(n, 1)) = make_uint4(pack_bf16(lo[2], 0.f), 0u, 0u, 0u);
    }
  }
  // Weight-ring bookkeeping: a tile consumes RING_TOT stages = RING_ROUNDS whole trips around the ring, so the
  // stage index of every load is a compile-time constant inside a tile (w_it restarts at 0) and only the phase
  // parity carries over from tile to tile (ring_par).  Keeping w_it itself as a loop-carried runtime counter
  // costs ~60 registers (every unrolled MMA descriptor becomes an address computation), i.e. a resident CTA.
  constexpr int RING_TOT = (D > 0 ? Cfg::NB1 * Cfg::P0 : 0) + Cfg::NB2 * Cfg::P1 + Cfg::MB3 * Cfg::P2;
  static_assert(RING_TOT % STAGES == 0, "a tile must consume whole trips around the weight ring");
  constexpr int RING_ROUNDS = RING_TOT / STAGES;
  constexpr bool RESIDENT = RING_TOT == STAGES;  // weights stay in the ring after the first tile (see the producer)
  uint32_t ring_par = 0;
  constexpr int ACC_USES = 2 + Cfg::MB3;  // bar_acc completions per tile
  uint32_t acc_par = 0;  // phase parity of bar_acc at the start of the tile (flips per tile when ACC_USES is odd)
  const float* cst = reinterpret_cast<const float*>(bp + Cfg::OFF_CONST);

  // Row geometry of the NEXT tile is fetched while the current tile computes: the neighbour index at the top of a tile,
  // the two dependent loads (the neighbour's xyz, the group's centroid) before the layer-1 accumulator wait -- the
  // gather phase of a tile then starts with its source rows in registers and their coordinates in L1 (two L2 round
  // trips less).
  int nx_idx = -1, nx_off = -1;
  auto fetch_idx = [&](long long t) {  // stage A: neighbour index of this thread's row in tile t
    const long long g = t * Cfg::G + threadIdx.x / NS;
    nx_idx = (t < n_tiles && g < p.groups) ? p.gidx[g * NS + (threadIdx.x % NS)] : -1;
  };
  auto fetch_xyz = [&](long long t) {  // stage B: source row; its coordinates and the centroid are pulled into L1
    nx_off = -1;
    if (nx_idx >= 0) {
      const long long g = t * Cfg::G + threadIdx.x / NS;
      nx_off = (int)(g / p.S) * p.N + nx_idx;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.xyz + (long long)nx_off * 3));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.new_xyz + g * 3));
    }
  };
  fetch_idx(blockIdx.x);
  fetch_xyz(blockIdx.x);
  // Feature gather of a tile: every thread publishes its source row, then all threads stream the rows as 16-byte
  // cp.async chunks (consecutive threads = consecutive chunks of a row) straight into the swizzled operand tile, every
  // chunk in flight at once.  With EARLY_GATHER it is ISSUED one tile ahead -- for the first tile here, for every later tile as soon as
  // the last layer-2 MMA block of the previous tile has completed (the operand buffer is free from then on) -- so its
  // latency runs under that tile's last max-epilogue instead of in front of the layer-0 MMAs.
  int* src_row = reinterpret_cast<int*>(bp + Cfg::OFF_ROWS);  // source point index in [0, K*N), -1 = padding
  auto issue_gather = [&]() {
    src_row[threadIdx.x] = nx_off;
    SA_BAR();
    constexpr int CPR = (D > 0 ? D : 8) / 8;   // 16-byte chunks per row
    constexpr int TOTAL = Cfg::ROWS * CPR;
#pragma unroll 1
    for (int ch = threadIdx.x; ch < TOTAL; ch += Cfg::ROWS) {
      const int r = ch / CPR, c8 = ch % CPR;
      const int off = src_row[r];
      const uint32_t dst = base + Cfg::OFF_X + (r >> 7) * Cfg::X_BYTES + xoff(r & 127, c8 * 8);
      const void* src = off >= 0 ? (const void*)(p.feats + (long long)off * D + c8 * 8) : (const void*)p.feats;
      const int nbytes = off >= 0 ? 16 : 0;  // padding rows: zero fill
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // (level 3 runs one CTA per SM: nothing shares the SM during the early issue, and measured it is slower there --
  // 374 vs 353 us -- so the early issue is used only when CTAs share an SM: level 2, 436 -> 415 us)
  constexpr bool EARLY_GATHER = D > 0 && Cfg::MINB >= 2;
  if (EARLY_GATHER && (long long)blockIdx.x < n_tiles) issue_gather();

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const long long g0 = tile * Cfg::G;
  const bool tr0 = tile == blockIdx.x;  // trace stamps: first tile only
  int w_it = 0;  // consumer position in the weight ring within this tile (thread 32 only)
  uint32_t acc_phase = 0;
  // re-derived per tile behind an opaque barrier: otherwise the compiler hoists every swizzled operand address
  // of the epilogues out of the tile loop and keeps ~60 of them live in registers (which costs a resident CTA)
  int row = wq * 32 + lane;                                            // row inside this thread's sub-tile
  asm volatile("" : "+r"(row));
  uint8_t* xsub = bp + Cfg::OFF_X + sub * Cfg::X_BYTES;                // this sub-tile's operand buffer
  // ---------------- operands of layer 0: the (dx,dy,dz) hi/lo rows now, the feature rows issued one tile ahead ----------
  {
    const int r = threadIdx.x;  // 0 .. ROWS-1
    const int off = nx_off;     // -1 = padding row
    float d[3] = {0.f, 0.f, 0.f}, hi[3], lo[3];  // this row: xyz[idx] - centroid (fp32), then its bf16 hi/lo split
    if (off >= 0) {
      const float* px = p.xyz + (long long)off * 3;
      const float* pc = p.new_xyz + (g0 + r / NS) * 3;
      d[0] = fsub(px[0], pc[0]), d[1] = fsub(px[1], pc[1]), d[2] = fsub(px[2], pc[2]);
    }
    fetch_idx(tile + gridDim.x);  // stage A for the next tile (consumed by fetch_xyz below)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      hi[i] = __bfloat162float(__float2bfloat16_rn(d[i]));
      lo[i] = d[i] - hi[i];
    }
    // k = [x_hi y_hi z_hi | x_lo y_lo z_lo | x_hi y_hi z_hi | 0 ...] against the weight row built above
    uint8_t* xb = bp + Cfg::OFF_XYZ + (r >> 7) * 4096;
    *reinterpret_cast<uint4*>(xb + xyzoff(r & 127, 0)) =
        make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], lo[0]), pack_bf16(lo[1], lo[2]), pack_bf16(hi[0], hi[1]));
    *reinterpret_cast<uint4*>(xb + xyzoff(r & 127, 1)) = make_uint4(pack_bf16(hi[2], 0.f), 0u, 0u, 0u);
    if (D > 0 && !EARLY_GATHER) issue_gather();
    if (D > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");  // this tile's feature rows have landed
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  SA_BAR();
  if (tr0) SA_TRACE(1);

  // ---------------- layers 0 and 1: rows on M ----------------
#pragma unroll
  for (int layer = 0; layer < 2; ++layer) {
    const int PANELS = layer == 0 ? Cfg::P0 : Cfg::P1;
    const int NB = layer == 0 ? Cfg::NB1 : Cfg::NB2;
    const int COUT = layer == 0 ? C1 : C2;
    const float* bias = cst + (layer == 0 ? 0 : C1);
    if (threadIdx.x == 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int nb = 0; nb < NB; ++nb) {
        for (int kp = 0; kp < PANELS; ++kp) {
          const int s = w_it % STAGES;
          if (!RESIDENT || tr0) mbar_wait(bar_full + 8 * s, ((w_it / STAGES) & 1) ^ ring_par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t db = desc_kmajor(base + Cfg::OFF_W + s * SF_TILE);
#pragma unroll
          for (int t = 0; t < SUB; ++t) {
            const uint64_t da = desc_kmajor(base + Cfg::OFF_X + t * Cfg::X_BYTES + kp * SF_TILE);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + t * Cfg::CW + nb * 128, da + 2 * k, db + 2 * k, SF_IDESC, (kp | k) != 0);
          }
          if (!RESIDENT) umma_commit(bar_empty + 8 * s);
          ++w_it;
        }
        if (layer == 0) {  // the (dx,dy,dz) columns: one K=16 step on the un-swizzled hi/lo operands
          const uint64_t db = desc_kmajor_nosw(base + Cfg::OFF_WXYZ + nb * 4096);
#pragma unroll
          for (int t = 0; t < SUB; ++t)
            umma_bf16(tmem + t * Cfg::CW + nb * 128, desc_kmajor_nosw(base + Cfg::OFF_XYZ + t * 4096), db, SF_IDESC, PANELS != 0);
        }
      }
      umma_commit(bar_acc);
    }
    if (layer == 1) fetch_xyz(tile + gridDim.x);  // stage B for the next tile, under the layer-1 MMAs
    __syncwarp();
    mbar_wait(bar_acc, acc_phase ^ acc_par);
    acc_phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tr0) SA_TRACE(2 + 2 * layer);
    // epilogue: thread = row; (+ xyz terms) + bias + ReLU -> bf16 -> next operand tile (in place).  The TMEM load of
    // chunk c + 1 is in flight while chunk c is converted (two register buffers, ping-pong).
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16) + sub * Cfg::CW;
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    auto epi_chunk = [&](const uint32_t* v, int c) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + j);
        // ReLU on the packed pair after rounding (== rounding after ReLU: monotone, 0 is exact): one HMNMX2 for two values
        __nv_bfloat162 p0 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y), zero2);
        __nv_bfloat162 p1 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w), zero2);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&p0);
        pk[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(xsub + xoff(row, c * 32 + i * 8)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    };
    {
      uint32_t va[32], vb[32];
      tmem_ld32(lane_addr, va);
#pragma unroll 1
      for (int c = 0; c < COUT / 32; c += 2) {  // COUT / 32 is even (channel counts are multiples of 64)
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tmem_ld32(lane_addr + (c + 1) * 32, vb);
        epi_chunk(va, c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 2 < COUT / 32) tmem_ld32(lane_addr + (c + 2) * 32, va);
        epi_chunk(vb, c + 1);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SA_BAR();
    if (tr0) SA_TRACE(3 + 2 * layer);
  }

  // ---------------- layer 2: channels on M (one 128-channel block at a time), rows on N ----------------
  {
    const float* b2s = cst + C1 + C2;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16) + sub * 128;
#pragma unroll 1
    for (int mb = 0; mb < Cfg::MB3; ++mb) {
      if (threadIdx.x == 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kp = 0; kp < Cfg::P2; ++kp) {
          const int s = w_it % STAGES;
          if (!RESIDENT || tr0) mbar_wait(bar_full + 8 * s, ((w_it / STAGES) & 1) ^ ring_par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = desc_kmajor(base + Cfg::OFF_W + s * SF_TILE);
#pragma unroll
          for (int t = 0; t < SUB; ++t) {
            const uint64_t db = desc_kmajor(base + Cfg::OFF_X + t * Cfg::X_BYTES + kp * SF_TILE);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + t * 128, da + 2 * k, db + 2 * k, SF_IDESC, (kp | k) != 0);
          }
          if (!RESIDENT) umma_commit(bar_empty + 8 * s);
          ++w_it;
        }
        umma_commit(bar_acc);
      }
      __syncwarp();
      mbar_wait(bar_acc, acc_phase ^ acc_par);
      acc_phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tr0 && mb < 4) SA_TRACE(6 + 2 * mb);
      // every MMA of this tile has completed: the operand buffer is free -> start the next tile's feature gather
      if (EARLY_GATHER && mb == Cfg::MB3 - 1 && tile + gridDim.x < n_tiles) issue_gather();
      const int ch = mb * 128 + wq * 32 + lane;
      const float bias = b2s[ch];
      {  // max over each group's NS rows = TMEM columns; the load of chunk q + 1 is in flight while chunk q is reduced
        constexpr int CPG = NS / 32;  // 32-column chunks per group; a sub-tile has 128 columns = 4 chunks
        uint32_t va[32], vb[32];
        float m = -INFINITY;
        auto reduce = [&](const uint32_t* v, int q) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
          m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          if ((q + 1) % CPG == 0) {  // last chunk of group q / CPG
            const long long grp = g0 + sub * Cfg::GS + q / CPG;
            if (grp < p.groups) p.out[grp * C3 + ch] = __float2bfloat16_rn(fmaxf(m + bias, 0.f));
            m = -INFINITY;
          }
        };
        tmem_ld32(lane_addr, va);
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld32(lane_addr + (q + 1) * 32, vb);
          reduce(va, q);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (q + 2 < 4) tmem_ld32(lane_addr + (q + 2) * 32, va);
          reduce(vb, q + 1);
        }
      }
      // the next channel block reuses the same TMEM columns: every compute warp must have drained them
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      SA_BAR();
      if (tr0 && mb < 4) SA_TRACE(7 + 2 * mb);
    }
  }
  if (!RESIDENT) ring_par ^= (RING_ROUNDS & 1);
  acc_par ^= (ACC_USES & 1);
  }  // tile loop
  SA_TRACE(14);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS));
#undef SA_BAR
#undef SA_TRACE
}

// ------------------------------------------------------------------------------------------------------------
// Resident-weight variant (levels whose three weight matrices fit in shared memory next to the operand tiles:
// level 1 = 32 KB, level 2 = 128 KB).  One persistent CTA per SM holds ALL weight tiles for the whole launch
// (one TMA burst at CTA start, no weight ring, no producer warp, no per-tile L2 traffic for weights) and runs
// NWG independent warpgroups.  Each warpgroup walks its own sequence of 128-row tiles with its own operand
// buffer, TMEM columns, accumulator barrier and named barrier, so the groups drift out of phase and the tensor
// pipe works for one group while the others gather or run epilogues -- the overlap that two co-resident CTAs
// of the streaming kernel give, without streaming the weights twice.
// ------------------------------------------------------------------------------------------------------------
template <int NS, int D, int C1, int C2, int C3, int NWG>
struct SaResCfg {
  static constexpr int THREADS = 128 * NWG;
  static constexpr int P0 = D / 64, P1 = (C1 + 63) / 64, P2 = (C2 + 63) / 64;
  static constexpr int NB1 = (C1 + 127) / 128, NB2 = (C2 + 127) / 128, MB3 = C3 / 128;
  static constexpr int X_PANELS = (P0 > P1 ? (P0 > P2 ? P0 : P2) : (P1 > P2 ? P1 : P2));
  static constexpr int W0_TILES = NB1 * P0, W1_TILES = NB2 * P1, W2_TILES = MB3 * P2;
  static constexpr uint32_t OFF_W0 = 0;
  static constexpr uint32_t OFF_W1 = OFF_W0 + W0_TILES * SF_TILE;
  static constexpr uint32_t OFF_W2 = OFF_W1 + W1_TILES * SF_TILE;
  static constexpr uint32_t OFF_X = OFF_W2 + W2_TILES * SF_TILE;        // [NWG][X_PANELS] operand panels
  static constexpr uint32_t X_BYTES = X_PANELS * SF_TILE;
  static constexpr uint32_t OFF_XYZ = OFF_X + NWG * X_BYTES;            // [NWG][128 x 16] bf16 hi/lo offsets
  static constexpr int WXYZ_ROWS = NB1 * 128;
  static constexpr uint32_t OFF_WXYZ = OFF_XYZ + NWG * 4096;
  static constexpr uint32_t OFF_BAR = OFF_WXYZ + WXYZ_ROWS * 32;        // bar_w | bar_acc[NWG] | tmem slot
  static constexpr uint32_t OFF_ROWS = OFF_BAR + 128;                   // [NWG][128] int32 gather offsets
  static constexpr uint32_t OFF_CONST = OFF_ROWS + NWG * 512;           // b0 | b1 | b2 (fp32)
  static constexpr uint32_t SMEM = OFF_CONST + (C1 + C2 + C3) * 4 + 1024;
  static constexpr int TMEM_COLS = NWG * 128 > 256 ? 512 : (NWG * 128 > 128 ? 256 : 128);
  static constexpr int GS = 128 / NS;                                   // groups per tile
  static_assert(NB1 == 1 && NB2 == 1, "layers 0/1 must fit one 128-column accumulator block");
  static_assert(SMEM <= 227 * 1024, "resident weights + operand tiles exceed shared memory");
};

template <int NS, int D, int C1, int C2, int C3, int NWG>
__global__ void __launch_bounds__(128 * NWG, 1)
    sa_resident_kernel(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                       const __grid_constant__ CUtensorMap map_w2, const SaParams p) {
  using Cfg = SaResCfg<NS, D, C1, C2, C3, NWG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_w = base + Cfg::OFF_BAR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2, wq = warp & 3;       // warpgroup and TMEM lane quarter
  const int t128 = threadIdx.x & 127;             // thread inside its warpgroup
  const uint32_t bar_acc = bar_w + 8 + 8 * wg;
  const uint32_t tmem_slot = bar_w + 8 + 8 * NWG;
  const long long n_tiles = (p.groups + Cfg::GS - 1) / Cfg::GS;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int g = 0; g < NWG; ++g) mbar_init(bar_w + 8 + 8 * g, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // every weight tile of the level, once
    mbar_expect_tx(bar_w, (Cfg::W0_TILES + Cfg::W1_TILES + Cfg::W2_TILES) * SF_TILE);
    if (D > 0)
      for (int kp = 0; kp < Cfg::P0; ++kp) tma_load_2d(base + Cfg::OFF_W0 + kp * SF_TILE, &map_w0, bar_w, kp * 64, 0);
    for (int kp = 0; kp < Cfg::P1; ++kp) tma_load_2d(base + Cfg::OFF_W1 + kp * SF_TILE, &map_w1, bar_w, kp * 64, 0);
    for (int mb = 0; mb < Cfg::MB3; ++mb)
      for (int kp = 0; kp < Cfg::P2; ++kp)
        tma_load_2d(base + Cfg::OFF_W2 + (mb * Cfg::P2 + kp) * SF_TILE, &map_w2, bar_w, kp * 64, mb * 128);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {  // per-channel constants and the weight side of the (dx,dy,dz) step (see sa_fused_kernel)
    float* cst = reinterpret_cast<float*>(bp + Cfg::OFF_CONST);
    for (int i = threadIdx.x; i < C1; i += Cfg::THREADS) cst[i] = p.b0[i];
    for (int i = threadIdx.x; i < C2; i += Cfg::THREADS) cst[C1 + i] = p.b1[i];
    for (int i = threadIdx.x; i < C3; i += Cfg::THREADS) cst[C1 + C2 + i] = p.b2[i];
    for (int n = threadIdx.x; n < Cfg::WXYZ_ROWS; n += Cfg::THREADS) {
      float w[3] = {0.f, 0.f, 0.f}, hi[3], lo[3];
      if (n < C1) w[0] = p.wxyz[n * 4], w[1] = p.wxyz[n * 4 + 1], w[2] = p.wxyz[n * 4 + 2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        hi[i] = __bfloat162float(__float2bfloat16_rn(w[i]));
        lo[i] = w[i] - hi[i];
      }
      uint8_t* wb = bp + Cfg::OFF_WXYZ;
      *reinterpret_cast<uint4*>(wb + xyzoff(n, 0)) =
          make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[0]), pack_bf16(hi[1], hi[2]), pack_bf16(lo[0], lo[1]));
      *reinterpret_cast<uint4*>(wb + xyzoff(n, 1)) = make_uint4(pack_bf16(lo[2], 0.f), 0u, 0u, 0u);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(bp + Cfg::OFF_BAR + 8 + 8 * NWG) + wg * 128;  // this group's columns

#define SAR_BAR() asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory")
  const float* cst = reinterpret_cast<const float*>(bp + Cfg::OFF_CONST);
  uint8_t* xbuf = bp + Cfg::OFF_X + wg * Cfg::X_BYTES;
  const uint32_t xaddr = base + Cfg::OFF_X + wg * Cfg::X_BYTES;
  const uint32_t xyz_addr = base + Cfg::OFF_XYZ + wg * 4096;
  int* src_row = reinterpret_cast<int*>(bp + Cfg::OFF_ROWS + wg * 512);
  const int row = wq * 32 + lane;
  const bool issuer = t128 == 32;  // lane 0 of the group's second warp issues the group's MMAs
  uint32_t acc_phase = 0;
  if (issuer) {
    mbar_wait(bar_w, 0);  // the resident weights have landed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // row geometry of the group's NEXT tile is fetched under the current tile (see sa_fused_kernel)
  int nx_idx = -1, nx_off = -1;
  const long long tile_step = (long long)gridDim.x * NWG;
  auto fetch_idx = [&](long long t) {
    const long long g = t * Cfg::GS + t128 / NS;
    nx_idx = (t < n_tiles && g < p.groups) ? p.gidx[g * NS + (t128 % NS)] : -1;
  };
  auto fetch_xyz = [&](long long t) {
    nx_off = -1;
    if (nx_idx >= 0) {
      const long long g = t * Cfg::GS + t128 / NS;
      nx_off = (int)(g / p.S) * p.N + nx_idx;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.xyz + (long long)nx_off * 3));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.new_xyz + g * 3));
    }
  };
  fetch_idx((long long)blockIdx.x * NWG + wg);
  fetch_xyz((long long)blockIdx.x * NWG + wg);

  for (long long tile = (long long)blockIdx.x * NWG + wg; tile < n_tiles; tile += tile_step) {
    const long long g0 = tile * Cfg::GS;
    // ---------------- gather (thread = row): neighbour index, hi/lo offsets, then the feature rows by cp.async
    {
      const int r = t128;
      const int off = nx_off;  // -1 = padding row
      float d[3] = {0.f, 0.f, 0.f}, hi[3], lo[3];
      if (off >= 0) {
        const float* px = p.xyz + (long long)off * 3;
        const float* pc = p.new_xyz + (g0 + r / NS) * 3;
        d[0] = fsub(px[0], pc[0]), d[1] = fsub(px[1], pc[1]), d[2] = fsub(px[2], pc[2]);
      }
      fetch_idx(tile + tile_step);  // stage A for the next tile
      src_row[r] = off;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        hi[i] = __bfloat162float(__float2bfloat16_rn(d[i]));
        lo[i] = d[i] - hi[i];
      }
      uint8_t* xb = bp + Cfg::OFF_XYZ + wg * 4096;
      *reinterpret_cast<uint4*>(xb + xyzoff(r, 0)) =
          make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], lo[0]), pack_bf16(lo[1], lo[2]), pack_bf16(hi[0], hi[1]));
      *reinterpret_cast<uint4*>(xb + xyzoff(r, 1)) = make_uint4(pack_bf16(hi[2], 0.f), 0u, 0u, 0u);
      if (D > 0) {
        SAR_BAR();
        constexpr int CPR = D / 8;  // 16-byte chunks per row
#pragma unroll 2
        for (int ch = t128; ch < 128 * CPR; ch += 128) {
          const int rr = ch / CPR, c8 = ch % CPR;
          const int o2 = src_row[rr];
          const void* src = o2 >= 0 ? (const void*)(p.feats + (long long)o2 * D + c8 * 8) : (const void*)p.feats;
          const int nbytes = o2 >= 0 ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xaddr + xoff(rr, c8 * 8)), "l"(src), "r"(nbytes)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SAR_BAR();

    // ---------------- layers 0 and 1: rows on M, one 128-column accumulator block ----------------
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
      const int PANELS = layer == 0 ? Cfg::P0 : Cfg::P1;
      const uint32_t woff = layer == 0 ? Cfg::OFF_W0 : Cfg::OFF_W1;
      const int COUT = layer == 0 ? C1 : C2;
      const float* bias = cst + (layer == 0 ? 0 : C1);
      if (issuer) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kp = 0; kp < PANELS; ++kp) {
          const uint64_t da = desc_kmajor(xaddr + kp * SF_TILE), db = desc_kmajor(base + woff + kp * SF_TILE);
          for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, SF_IDESC, (kp | k) != 0);
        }
        if (layer == 0) umma_bf16(tmem, desc_kmajor_nosw(xyz_addr), desc_kmajor_nosw(base + Cfg::OFF_WXYZ), SF_IDESC, PANELS != 0);
        umma_commit(bar_acc);
      }
      if (layer == 1) fetch_xyz(tile + tile_step);  // stage B for the next tile, under the layer-1 MMAs
      __syncwarp();
      mbar_wait(bar_acc, acc_phase);
      acc_phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      auto epi_chunk = [&](const uint32_t* v, int c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + j);
          // ReLU on the packed pair after rounding (== rounding after ReLU: monotone, 0 is exact): one HMNMX2 for two values
          __nv_bfloat162 p0 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y), zero2);
          __nv_bfloat162 p1 = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w), zero2);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&p0);
          pk[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(xbuf + xoff(row, c * 32 + i * 8)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      };
      {  // the TMEM load of chunk c + 1 is in flight while chunk c is converted
        uint32_t va[32], vb[32];
        tmem_ld32(lane_addr, va);
#pragma unroll 1
        for (int c = 0; c < COUT / 32; c += 2) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld32(lane_addr + (c + 1) * 32, vb);
          epi_chunk(va, c);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 2 < COUT / 32) tmem_ld32(lane_addr + (c + 2) * 32, va);
          epi_chunk(vb, c + 1);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      SAR_BAR();
    }

    // ---------------- layer 2: channels on M (128 per block), the tile's 128 rows on N ----------------
    {
      const float* b2s = cst + C1 + C2;
      const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
      for (int mb = 0; mb < Cfg::MB3; ++mb) {
        if (issuer) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kp = 0; kp < Cfg::P2; ++kp) {
            const uint64_t da = desc_kmajor(base + Cfg::OFF_W2 + (mb * Cfg::P2 + kp) * SF_TILE);
            const uint64_t db = desc_kmajor(xaddr + kp * SF_TILE);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, SF_IDESC, (kp | k) != 0);
          }
          umma_commit(bar_acc);
        }
        __syncwarp();
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ch = mb * 128 + wq * 32 + lane;
        const float bias = b2s[ch];
        {  // max over each group's NS rows = TMEM columns; the load of chunk q + 1 is in flight while chunk q is reduced
          constexpr int CPG = NS / 32;  // 32-column chunks per group; the tile has 128 columns = 4 chunks
          uint32_t va[32], vb[32];
          float m = -INFINITY;
          auto reduce = [&](const uint32_t* v, int q) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
            m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
            if ((q + 1) % CPG == 0) {  // last chunk of group q / CPG
              const long long grp = g0 + q / CPG;
              if (grp < p.groups) p.out[grp * C3 + ch] = __float2bfloat16_rn(fmaxf(m + bias, 0.f));
              m = -INFINITY;
            }
          };
          tmem_ld32(lane_addr, va);
#pragma unroll
          for (int q = 0; q < 4; q += 2) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_ld32(lane_addr + (q + 1) * 32, vb);
            reduce(va, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (q + 2 < 4) tmem_ld32(lane_addr + (q + 2) * 32, va);
            reduce(vb, q + 1);
          }
        }
        // the next channel block / tile reuses the same TMEM columns: the whole group must have drained them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        SAR_BAR();
      }
    }
  }
#undef SAR_BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem - wg * 128), "r"(Cfg::TMEM_COLS));
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

int weight_map(CUtensorMap* map, const void* w, int rows, int cols, int ld) {
  auto fn = encode_fn();
  if (!fn) return PFPP_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PFPP_OK : PFPP_EINVAL;
}

template <int NS, int D, int C1, int C2, int C3, int STAGES, int SUB>
int launch_sa(const SaParams& p, const void* w0, const void* w1, const void* w2, cudaStream_t stream) {
  using Cfg = SaCfg<NS, D, C1, C2, C3, STAGES, SUB>;
  static_assert(D % 64 == 0 && C1 % 64 == 0 && C2 % 64 == 0 && C3 % 128 == 0, "panel-aligned channel counts");
  CUtensorMap m0, m1, m2;
  int rc = D > 0 ? weight_map(&m0, w0, C1, D, D) : weight_map(&m0, w1, C2, C1, C1);  // m0 unused when D == 0
  if (rc) return rc;
  rc = weight_map(&m1, w1, C2, C1, C1);
  if (rc) return rc;
  rc = weight_map(&m2, w2, C3, C2, C2);
  if (rc) return rc;
  auto kern = sa_fused_kernel<NS, D, C1, C2, C3, STAGES, SUB>;
  PFPP_ENSURE_SMEM(kern, Cfg::SMEM);
  const long long tiles = (p.groups + Cfg::G - 1) / Cfg::G;
  // persistent CTAs: as many as fit on the GPU at once (shared memory and TMEM bound), each walks tiles
  // blockIdx.x, blockIdx.x + gridDim.x, ... so barrier / TMEM / constant set-up is paid once per CTA
  static int per_sm = 0, n_sm = 0;
  if (!per_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int by_regs = 65536 / (fa.numRegs * ((Cfg::THREADS + 31) / 32 * 32));
    per_sm = Cfg::MINB < by_regs ? Cfg::MINB : by_regs;
    per_sm = per_sm < 1 ? 1 : per_sm;
  }
  // two CTAs per resident slot: the second half starts as the first retires, which keeps co-resident CTAs out of
  // phase with each other (measured: better than exactly one persistent CTA per slot)
  long long grid = (long long)2 * per_sm * n_sm;
  if (grid > tiles) grid = tiles;
  static const int grid_mode = []() {
    const char* e = getenv("PFPP_SA_GRID");  // tuning aid: 0 = one CTA per tile, n > 0 = n CTAs per SM
    return e ? atoi(e) : -1;
  }();
  if (grid_mode == 0) grid = tiles;
  if (grid_mode > 0) grid = tiles < (long long)grid_mode * n_sm ? tiles : (long long)grid_mode * n_sm;
  kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, stream>>>(m0, m1, m2, p);
  PFPP_RETURN_LAST();
}

template <int NS, int D, int C1, int C2, int C3, int NWG>
int launch_sa_resident(const SaParams& p, const void* w0, const void* w1, const void* w2, cudaStream_t stream) {
  using Cfg = SaResCfg<NS, D, C1, C2, C3, NWG>;
  static_assert(D % 64 == 0 && C1 % 64 == 0 && C2 % 64 == 0 && C3 % 128 == 0, "panel-aligned channel counts");
  CUtensorMap m0, m1, m2;
  int rc = D > 0 ? weight_map(&m0, w0, C1, D, D) : weight_map(&m0, w1, C2, C1, C1);  // m0 unused when D == 0
  if (rc) return rc;
  rc = weight_map(&m1, w1, C2, C1, C1);
  if (rc) return rc;
  rc = weight_map(&m2, w2, C3, C2, C2);
  if (rc) return rc;
  auto kern = sa_resident_kernel<NS, D, C1, C2, C3, NWG>;
  PFPP_ENSURE_SMEM(kern, Cfg::SMEM);
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long tiles = (p.groups + Cfg::GS - 1) / Cfg::GS;
  long long grid = (tiles + NWG - 1) / NWG;
  if (grid > n_sm) grid = n_sm;
  kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, stream>>>(m0, m1, m2, p);
  PFPP_RETURN_LAST();
}

}  // namespace

static int sa_dispatch(int level, const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K,
                       int N, int S, const void* w0_feat, const float* w0_xyz, const float* b0, const void* w1,
                       const float* b1, const void* w2, const float* b2, void* out, long long* trace,
                       cudaStream_t stream) {
  PFPP_CHECK_ARG(xyz && new_xyz && gidx && w0_xyz && w1 && w2 && b0 && b1 && b2 && out && K >= 0);
  if (K == 0) return PFPP_OK;
  SaParams p{xyz, new_xyz, (const __nv_bfloat16*)feats, gidx, w0_xyz, b0, b1, b2, (__nv_bfloat16*)out, N, S,
             (long long)K * S, trace};
  // Level 1 (32 KB of weights) runs the resident-weight kernel with 4 independent warpgroups per SM.  Level 2's
  // weights (128 KB) also fit, but then only two 128-row tiles are in flight per SM and the kernel loses more
  // latency hiding than it gains (measured 975 us vs 785 us), so levels 2 and 3 stream their weights.
  // PFPP_SA_RESIDENT (tuning aid): bit 0 = level 1, bit 1 = level 2.
  static const int resident = []() {
    const char* e = getenv("PFPP_SA_RESIDENT");
    return e ? atoi(e) : 1;
  }();
  switch (level) {
    case 1:
      if (resident & 1) return launch_sa_resident<32, 0, 64, 64, 128, 4>(p, w0_feat, w1, w2, stream);
      return launch_sa<32, 0, 64, 64, 128, 2, 1>(p, w0_feat, w1, w2, stream);
    case 2:
      PFPP_CHECK_ARG(feats && w0_feat);
      if (resident & 2) return launch_sa_resident<64, 128, 128, 128, 256, 2>(p, w0_feat, w1, w2, stream);
      return launch_sa<64, 128, 128, 128, 256, 2, 2>(p, w0_feat, w1, w2, stream);
    case 3:
      PFPP_CHECK_ARG(feats && w0_feat);
      return launch_sa<64, 256, 256, 256, 512, 2, 2>(p, w0_feat, w1, w2, stream);
    default:
      return PFPP_EINVAL;
  }
}

extern "C" int pfpp_sa_fused(int level, const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K,
                             int N, int S, const void* w0_feat, const float* w0_xyz, const float* b0, const void* w1,
                             const float* b1, const void* w2, const float* b2, void* out, cudaStream_t stream) {
  return sa_dispatch(level, xyz, new_xyz, feats, gidx, K, N, S, w0_feat, w0_xyz, b0, w1, b1, w2, b2, out, nullptr, stream);
}

// Debug variant: per-CTA clock64 stamps of the kernel phases into trace[grid][16] (tools/bench_sa.py --trace).
extern "C" int pfpp_sa_fused_trace(int level, const float* xyz, const float* new_xyz, const void* feats, const int* gidx,
                                   int K, int N, int S, const void* w0_feat, const float* w0_xyz, const float* b0,
                                   const void* w1, const float* b1, const void* w2, const float* b2, void* out,
                                   long long* trace, cudaStream_t stream) {
  return sa_dispatch(level, xyz, new_xyz, feats, gidx, K, N, S, w0_feat, w0_xyz, b0, w1, b1, w2, b2, out, trace, stream);
}
