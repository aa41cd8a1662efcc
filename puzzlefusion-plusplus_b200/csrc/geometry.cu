// Geometric stages of the fragment encoder: quaternion rotate, farthest point sampling,
// ball query, grouping gather, group max-pool, vector quantisation.
//
// These are scan/latency kernels over <=24 KB point tiles (HBM/L2-bound, no tensor cores):
// xyz tiles are staged once in shared memory (SoA), distances live in registers, reductions are
// warp shuffles.  All floating-point decisions replicate the oracle's op order with individually
// rounded fp32 ops so that indices are bit-identical (see oracle/third_party.py, oracle/encoder.py).
#include <climits>

#include "common.cuh"
#include "../../include/pfpp.h"

// ---------------------------------------------------------------------------------------------
// quaternion rotate (auto_aggl.py:70-78 + pytorch3d quaternion_apply)
// ---------------------------------------------------------------------------------------------
struct Quat {
  float w, x, y, z;
};

__device__ __forceinline__ Quat load_quat_normalised(const float* q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  // torch CPU norm over 4 elements: sequential fp32 sum of squares, then sqrt
  float n = __fsqrt_rn(fadd(fadd(fadd(fmul(w, w), fmul(x, x)), fmul(y, y)), fmul(z, z)));
  return Quat{fdiv(w, n), fdiv(x, n), fdiv(y, n), fdiv(z, n)};
}

// out = (q (x) (0,p) (x) q*)[1:], term order of quaternion_raw_multiply, no FMA.
__device__ __forceinline__ void quat_apply(const Quat& a, float px, float py, float pz, float& ox, float& oy,
                                           float& oz) {
  const float bw = 0.0f;
  float tw = fsub(fsub(fsub(fmul(a.w, bw), fmul(a.x, px)), fmul(a.y, py)), fmul(a.z, pz));
  float tx = fsub(fadd(fadd(fmul(a.w, px), fmul(a.x, bw)), fmul(a.y, pz)), fmul(a.z, py));
  float ty = fadd(fadd(fsub(fmul(a.w, py), fmul(a.x, pz)), fmul(a.y, bw)), fmul(a.z, px));
  float tz = fadd(fsub(fadd(fmul(a.w, pz), fmul(a.x, py)), fmul(a.y, px)), fmul(a.z, bw));
  float cw = a.w, cx = -a.x, cy = -a.y, cz = -a.z;
  ox = fsub(fadd(fadd(fmul(tw, cx), fmul(tx, cw)), fmul(ty, cz)), fmul(tz, cy));
  oy = fadd(fadd(fsub(fmul(tw, cy), fmul(tx, cz)), fmul(ty, cw)), fmul(tz, cx));
  oz = fadd(fsub(fadd(fmul(tw, cz), fmul(tx, cy)), fmul(ty, cx)), fmul(tz, cw));
}

// pose apply used by the verify/merge stage: out = quat_apply(q or q/|q|, p * scale) + t
__global__ void pose_apply_kernel(const float* __restrict__ pts, const int* __restrict__ seg_start,
                                  const int* __restrict__ seg_len, const int* __restrict__ seg_pose,
                                  const float* __restrict__ pose, const float* __restrict__ seg_scale,
                                  int normalise, float* __restrict__ out) {
  int s = blockIdx.x;
  const float* pp = pose + 7 * seg_pose[s];
  Quat q;
  if (normalise) {
    q = load_quat_normalised(pp + 3);
  } else {
    q = Quat{pp[3], pp[4], pp[5], pp[6]};
  }
  float sc = seg_scale ? seg_scale[s] : 1.0f;
  int st = seg_start[s], n = seg_len[s];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* p = pts + 3 * (size_t)(st + i);
    float x = p[0], y = p[1], z = p[2];
    if (seg_scale) {
      x = fmul(x, sc);
      y = fmul(y, sc);
      z = fmul(z, sc);
    }
    float ox, oy, oz;
    quat_apply(q, x, y, z, ox, oy, oz);
    float* o = out + 3 * (size_t)(st + i);
    o[0] = fadd(ox, pp[0]);
    o[1] = fadd(oy, pp[1]);
    o[2] = fadd(oz, pp[2]);
  }
}

extern "C" int pfpp_pose_apply(const float* pts, const int* seg_start, const int* seg_len, const int* seg_pose,
                               const float* pose, const float* seg_scale, int normalise, int n_segments,
                               float* out, cudaStream_t stream) {
  PFPP_CHECK_ARG(pts && seg_start && seg_len && seg_pose && pose && out && n_segments >= 0);
  if (n_segments == 0) return PFPP_OK;
  pose_apply_kernel<<<n_segments, 256, 0, stream>>>(pts, seg_start, seg_len, seg_pose, pose, seg_scale, normalise,
                                                     out);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// farthest point sampling (torch_cluster fps_kernel semantics)
//   one CTA (256 threads) per cloud; thread t owns points t, t+256, ... (the same ownership as
//   torch_cluster, which fixes the tie-break: larger dist, then lower thread, then lower index);
//   coordinates + running min-distance live in registers (PPT points per thread).
// ---------------------------------------------------------------------------------------------
#define FPS_THREADS 256

// Block-wide arg-max of (distance, then lower owner thread, then lower point index) with two warp-level
// redux.sync passes instead of shuffle trees.  Distances are >= 0, so their bit patterns order like signed
// integers; points that do not exist carry the distance -1.0f (a negative integer) and can never win.  Within
// a thread the caller keeps the first (lowest-index) maximum; between lanes / warps the lowest one wins, which
// is the (lower thread, lower index) tie-break of torch_cluster's strided scan + tree reduce.
// The partials are double-buffered by round parity -> one barrier per round.
// `red` points at this CTA's [2][2][FPS_THREADS/32] int partials (parity, {distance, index}, warp) as a 32-bit
// shared-window address, so the loads/stores are plain LDS/STS with immediate offsets.
__device__ __forceinline__ int fps_block_argmax(int ud, int n, int par, uint32_t red, int lane, int warp) {
  constexpr int W = FPS_THREADS / 32;
  const int m = __reduce_max_sync(0xffffffffu, ud);
  const unsigned win = __ballot_sync(0xffffffffu, ud == m);
  const uint32_t slot = red + (uint32_t)par * (2 * W * 4);
  if (lane == __ffs(win) - 1) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + warp * 4), "r"(m) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + (W + warp) * 4), "r"(n) : "memory");
  }
  __syncthreads();
  int pd = INT_MIN, pk = 0;
  if (lane < W) {
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pd) : "r"(slot + lane * 4) : "memory");
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pk) : "r"(slot + (W + lane) * 4) : "memory");
  }
  const int m2 = __reduce_max_sync(0xffffffffu, pd);
  const unsigned win2 = __ballot_sync(0xffffffffu, pd == m2);
  return __shfl_sync(0xffffffffu, pk, __ffs(win2) - 1);
}

template <int PPT, bool ROTATE>
__global__ void __launch_bounds__(FPS_THREADS)
    fps_kernel(const float* __restrict__ src, const int* __restrict__ src_slot, int N, int S,
               const float* __restrict__ quat, int quat_stride, const int* __restrict__ start,
               float* __restrict__ rot_out, int* __restrict__ out_idx, float* __restrict__ out_xyz) {
  extern __shared__ float sm[];
  float* sx = sm;
  float* sy = sm + N;
  float* sz = sm + 2 * N;
  int* hist = reinterpret_cast<int*>(sm + 3 * N);  // [S] selected indices, written out after the loop
  __shared__ int red_buf[2 * 2 * (FPS_THREADS / 32)];
  const uint32_t red = (uint32_t)__cvta_generic_to_shared(red_buf);

  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = src_slot ? src_slot[k] : k;
  const float* p = src + (size_t)slot * N * 3;

  Quat q;
  if (ROTATE) q = load_quat_normalised(quat + (size_t)slot * quat_stride);
  // coalesced AoS read -> SoA smem (+ rotated copy to global for the later stages)
  for (int i = tid; i < N; i += FPS_THREADS) {
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    if (ROTATE) {
      float ox, oy, oz;
      quat_apply(q, x, y, z, ox, oy, oz);
      x = ox, y = oy, z = oz;
      float* r = rot_out + ((size_t)k * N + i) * 3;
      r[0] = x, r[1] = y, r[2] = z;
    }
    sx[i] = x, sy[i] = y, sz[i] = z;
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    int n = tid + i * FPS_THREADS;
    bool ok = n < N;
    px[i] = ok ? sx[n] : 0.f;
    py[i] = ok ? sy[n] : 0.f;
    pz[i] = ok ? sz[n] : 0.f;
    dist[i] = ok ? 5e4f : -1.0f;  // a slot without a point keeps -1 (min(-1, d) = -1) and never wins the arg-max
  }
  int cur = start ? start[k] : 0;

  // The loop body is issue-bound (every resident warp runs it S times), so it is branch-free: no per-point
  // validity test, selects instead of branches for the per-thread best, and the outputs are written after the loop.
  for (int m = 0; m + 1 < S; ++m) {
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
    if (tid == 0) hist[m] = cur;
    int bd = 0, bi = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float dx = fsub(cx, px[i]), dy = fsub(cy, py[i]), dz = fsub(cz, pz[i]);
      float dd = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
      dd = fminf(dist[i], dd);
      dist[i] = dd;
      const int u = __float_as_int(dd);
      const bool better = (i == 0) || (u > bd);  // strict: the first (lowest-index) maximum of the thread wins
      bd = better ? u : bd;
      bi = better ? i : bi;
    }
    cur = fps_block_argmax(bd, tid + bi * FPS_THREADS, m & 1, red, lane, warp);
  }
  if (tid == 0) hist[S - 1] = cur;
  __syncthreads();
  int* oi = out_idx + (size_t)k * S;
  float* ox = out_xyz ? out_xyz + (size_t)k * S * 3 : nullptr;
  for (int m = tid; m < S; m += FPS_THREADS) {
    const int c = hist[m];
    oi[m] = c;
    if (ox) ox[3 * m] = sx[c], ox[3 * m + 1] = sy[c], ox[3 * m + 2] = sz[c];
  }
}

// large clouds (merge stage, up to 2^20 points): distances in a global scratch array
__global__ void __launch_bounds__(FPS_THREADS)
    fps_large_kernel(const float* __restrict__ src, const int* __restrict__ cloud_start,
                     const int* __restrict__ cloud_len, const int* __restrict__ n_samples,
                     const int* __restrict__ start, float* __restrict__ dist, const int* __restrict__ out_start,
                     int* __restrict__ out_idx) {
  __shared__ int red_buf[2 * 2 * (FPS_THREADS / 32)];
  const uint32_t red = (uint32_t)__cvta_generic_to_shared(red_buf);
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = cloud_len[k], S = n_samples[k];
  const float* p = src + 3 * (size_t)cloud_start[k];
  float* d = dist + cloud_start[k];
  int* oi = out_idx + out_start[k];
  for (int n = tid; n < N; n += FPS_THREADS) d[n] = 5e4f;
  int cur = start ? start[k] : 0;
  for (int m = 0; m < S; ++m) {
    if (tid == 0) oi[m] = cur;
    if (m + 1 == S) break;
    float cx = p[3 * cur], cy = p[3 * cur + 1], cz = p[3 * cur + 2];
    int bd = INT_MIN;  // a thread without points never wins
    int bn = 0;
    for (int n = tid; n < N; n += FPS_THREADS) {
      float dx = fsub(cx, p[3 * n]), dy = fsub(cy, p[3 * n + 1]), dz = fsub(cz, p[3 * n + 2]);
      float dd = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
      dd = fminf(d[n], dd);
      d[n] = dd;
      const int u = __float_as_int(dd);
      if (u > bd) bd = u, bn = n;
    }
    cur = fps_block_argmax(bd, bn, m & 1, red, lane, warp);
  }
}

template <bool ROTATE>
static int launch_fps(const float* src, const int* src_slot, int K, int N, int S, const float* quat, int quat_stride,
                      const int* start, float* rot_out, int* out_idx, float* out_xyz, cudaStream_t stream) {
  size_t smem = (size_t)3 * N * sizeof(float) + (size_t)S * sizeof(int);
#define PFPP_FPS_CASE(PPT)                                                                                   \
  fps_kernel<PPT, ROTATE><<<K, FPS_THREADS, smem, stream>>>(src, src_slot, N, S, quat, quat_stride, start, \
                                                             rot_out, out_idx, out_xyz)
  if (N <= 256) {
    PFPP_FPS_CASE(1);
  } else if (N <= 512) {
    PFPP_FPS_CASE(2);
  } else if (N <= 1024) {
    PFPP_FPS_CASE(4);
  } else if (N <= 2048) {
    PFPP_FPS_CASE(8);
  } else if (N <= 4096) {
    PFPP_ENSURE_SMEM((fps_kernel<16, ROTATE>), smem);
    PFPP_FPS_CASE(16);
  } else {
    return PFPP_EUNSUPPORTED;
  }
#undef PFPP_FPS_CASE
  PFPP_RETURN_LAST();
}

extern "C" int pfpp_fps(const float* xyz, int K, int N, int S, const int* start, int* out_idx, float* out_xyz,
                        cudaStream_t stream) {
  PFPP_CHECK_ARG(xyz && out_idx && K >= 0 && N > 0 && S > 0 && S <= N);
  if (K == 0) return PFPP_OK;
  return launch_fps<false>(xyz, nullptr, K, N, S, nullptr, 0, start, nullptr, out_idx, out_xyz, stream);
}

extern "C" int pfpp_rotate_fps(const float* part_pcs, const int* frag_slot, int K, int N, int S, const float* quat,
                               int quat_stride, float* rot_out, int* out_idx, float* out_xyz,
                               cudaStream_t stream) {
  PFPP_CHECK_ARG(part_pcs && quat && rot_out && out_idx && K >= 0 && N > 0 && S > 0 && S <= N);
  if (K == 0) return PFPP_OK;
  return launch_fps<true>(part_pcs, frag_slot, K, N, S, quat, quat_stride, nullptr, rot_out, out_idx, out_xyz,
                          stream);
}

extern "C" int pfpp_fps_ragged(const float* xyz, const int* cloud_start, const int* cloud_len, const int* n_samples,
                               const int* start, int n_clouds, float* dist_scratch, const int* out_start,
                               int* out_idx, cudaStream_t stream) {
  PFPP_CHECK_ARG(xyz && cloud_start && cloud_len && n_samples && dist_scratch && out_start && out_idx);
  if (n_clouds <= 0) return PFPP_OK;
  fps_large_kernel<<<n_clouds, FPS_THREADS, 0, stream>>>(xyz, cloud_start, cloud_len, n_samples, start,
                                                          dist_scratch, out_start, out_idx);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// ball query (pn2_utils.py:92-112): lowest-index `nsample` points with NOT(d > r^2), padded with
// the first hit.  d = -2*((sx*px + sy*py) + sz*pz) + |s|^2 + |p|^2, ops rounded individually.
//   CTA = (cloud, block of centroids); the cloud tile (xyz + |p|^2) is staged in smem; one warp
//   per centroid scans 32 points per iteration and compacts hits with ballot/popc.
// ---------------------------------------------------------------------------------------------
#define BQ_WARPS 8
#define BQ_CENTROIDS_PER_CTA 32

__global__ void __launch_bounds__(BQ_WARPS * 32)
    ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, int nsample,
                      float r2, int* __restrict__ out_idx) {
  extern __shared__ float sm[];
  float* sx = sm;
  float* sy = sm + N;
  float* sz = sm + 2 * N;
  float* sn = sm + 3 * N;
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + (size_t)k * N * 3;
  for (int i = tid; i < N; i += BQ_WARPS * 32) {
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    sx[i] = x, sy[i] = y, sz[i] = z;
    sn[i] = fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z));
  }
  __syncthreads();
  const int c0 = blockIdx.y * BQ_CENTROIDS_PER_CTA;
  for (int c = c0 + warp; c < min(S, c0 + BQ_CENTROIDS_PER_CTA); c += BQ_WARPS) {
    const float* q = new_xyz + ((size_t)k * S + c) * 3;
    float qx = q[0], qy = q[1], qz = q[2];
    float qn = fadd(fadd(fmul(qx, qx), fmul(qy, qy)), fmul(qz, qz));
    int* o = out_idx + ((size_t)k * S + c) * nsample;
    int cnt = 0, first = 0;
    for (int base = 0; base < N && cnt < nsample; base += 32) {
      int n = base + lane;
      bool in = false;
      if (n < N) {
        float dot = fadd(fadd(fmul(qx, sx[n]), fmul(qy, sy[n])), fmul(qz, sz[n]));
        float d = fmul(-2.0f, dot);
        d = fadd(d, qn);
        d = fadd(d, sn[n]);
        in = !(d > r2);
      }
      unsigned bal = __ballot_sync(0xffffffffu, in);
      if (bal) {
        if (cnt == 0) first = base + __ffs(bal) - 1;
        int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (in && pos < nsample) o[pos] = n;
        cnt += __popc(bal);
      }
    }
    if (cnt > nsample) cnt = nsample;
    for (int j = cnt + lane; j < nsample; j += 32) o[j] = first;
  }
}

extern "C" int pfpp_ball_query(const float* xyz, const float* new_xyz, int K, int N, int S, float radius_sq,
                               int nsample, int* out_idx, cudaStream_t stream) {
  PFPP_CHECK_ARG(xyz && new_xyz && out_idx && K >= 0 && N > 0 && S > 0 && nsample > 0);
  if (K == 0) return PFPP_OK;
  size_t smem = (size_t)4 * N * sizeof(float);
  if (smem > 200 * 1024) return PFPP_EUNSUPPORTED;
  PFPP_ENSURE_SMEM(ball_query_kernel, smem);
  dim3 grid(K, pfpp_cdiv(S, BQ_CENTROIDS_PER_CTA));
  ball_query_kernel<<<grid, BQ_WARPS * 32, smem, stream>>>(xyz, new_xyz, N, S, nsample, radius_sq, out_idx);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// grouping gather (pn2_utils.py:139-148): row (k,s,j) = [xyz[idx]-new_xyz[s], feats[idx]] padded to ld.
// One warp per output row; lanes run over channels (coalesced feature-row reads and writes).
// OutT = float (parity path) or __nv_bfloat16 (tensor-core path).
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OutT, typename FeatT>
__global__ void group_gather_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                    const FeatT* __restrict__ feats, const int* __restrict__ gidx, int N, int S,
                                    int ns, int D, int ld, long long rows, OutT* __restrict__ out) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  long long ks = row / ns;  // k*S + s
  int k = (int)(ks / S);
  int idx = gidx[row];
  OutT* o = out + row * ld;
  if (lane < 3) {
    float v = fsub(xyz[((size_t)k * N + idx) * 3 + lane], new_xyz[ks * 3 + lane]);
    o[lane] = to_out<OutT>(v);
  }
  const FeatT* f = feats ? feats + ((size_t)k * N + idx) * D : nullptr;
  for (int c = lane; c < ld - 3; c += 32) {
    float v = (f && c < D) ? (float)f[c] : 0.0f;
    o[3 + c] = to_out<OutT>(v);
  }
}

// split variant (fp32-grade tensor-core path): features are bf16 hi/lo rows [K, N, 2D] (lo at D + c) and are copied
// as they are; the centroid offsets are split here.  out rows hold 2*ld elements: hi at c, lo at ld + c.
__global__ void group_gather_split_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                          const __nv_bfloat16* __restrict__ feats, const int* __restrict__ gidx, int N,
                                          int S, int ns, int D, int ld, long long rows, __nv_bfloat16* __restrict__ out) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  long long ks = row / ns;  // k*S + s
  int k = (int)(ks / S);
  int idx = gidx[row];
  __nv_bfloat16* o = out + row * (2 * ld);
  if (lane < 3) {
    float v = fsub(xyz[((size_t)k * N + idx) * 3 + lane], new_xyz[ks * 3 + lane]);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    o[lane] = h;
    o[ld + lane] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  const __nv_bfloat16* f = feats ? feats + ((size_t)k * N + idx) * (2 * D) : nullptr;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int c = lane; c < ld - 3; c += 32) {
    o[3 + c] = (f && c < D) ? f[c] : zero;
    o[ld + 3 + c] = (f && c < D) ? f[D + c] : zero;
  }
}

extern "C" int pfpp_group_gather(const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K,
                                 int N, int S, int ns, int D, int ld, int out_bf16, void* out,
                                 cudaStream_t stream) {
  PFPP_CHECK_ARG(xyz && new_xyz && gidx && out && ld >= 3 + D && (feats || D == 0));
  long long rows = (long long)K * S * ns;
  if (rows == 0) return PFPP_OK;
  int grid = pfpp_cdiv(rows, 8);
  if (out_bf16 == 2)
    group_gather_split_kernel<<<grid, 256, 0, stream>>>(xyz, new_xyz, (const __nv_bfloat16*)feats, gidx, N, S, ns, D, ld,
                                                        rows, (__nv_bfloat16*)out);
  else if (out_bf16)
    group_gather_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        xyz, new_xyz, (const __nv_bfloat16*)feats, gidx, N, S, ns, D, ld, rows, (__nv_bfloat16*)out);
  else
    group_gather_kernel<float, float><<<grid, 256, 0, stream>>>(xyz, new_xyz, (const float*)feats, gidx, N, S, ns, D,
                                                                ld, rows, (float*)out);
  PFPP_RETURN_LAST();
}

// max over the nsample rows of each group (pn2_utils.py:214): in [G*ns, C] -> out [G, C]
template <typename T>
__global__ void group_max_kernel(const T* __restrict__ in, int ns, int C, int ld_in, long long G, T* __restrict__ out,
                                 int ld_out) {
  long long g = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const T* p = in + (g * ns) * (long long)ld_in + c;
    float m = (float)p[0];
    for (int j = 1; j < ns; ++j) m = fmaxf(m, (float)p[(long long)j * ld_in]);
    out[g * ld_out + c] = (T)m;
  }
}

// fp32 rows in, bf16 hi/lo split rows out (ld_out elements: hi at c, lo at ld_out/2 + c)
__global__ void group_max_split_kernel(const float* __restrict__ in, int ns, int C, int ld_in, long long G,
                                       __nv_bfloat16* __restrict__ out, int ld_out) {
  long long g = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = in + (g * ns) * (long long)ld_in + c;
    float m = p[0];
    for (int j = 1; j < ns; ++j) m = fmaxf(m, p[(long long)j * ld_in]);
    const __nv_bfloat16 h = __float2bfloat16_rn(m);
    out[g * ld_out + c] = h;
    out[g * ld_out + ld_out / 2 + c] = __float2bfloat16_rn(m - __bfloat162float(h));
  }
}

extern "C" int pfpp_group_max(const void* in, long long G, int ns, int C, int ld_in, int is_bf16, void* out,
                              int ld_out, cudaStream_t stream) {
  PFPP_CHECK_ARG(in && out && ns > 0 && C > 0);
  if (G == 0) return PFPP_OK;
  int threads = C >= 256 ? 256 : 128;
  if (is_bf16 == 2)  // fp32 in -> split out
    group_max_split_kernel<<<(unsigned)G, threads, 0, stream>>>((const float*)in, ns, C, ld_in, G, (__nv_bfloat16*)out,
                                                               ld_out);
  else if (is_bf16)
    group_max_kernel<__nv_bfloat16><<<(unsigned)G, threads, 0, stream>>>((const __nv_bfloat16*)in, ns, C, ld_in, G,
                                                                       (__nv_bfloat16*)out, ld_out);
  else
    group_max_kernel<float><<<(unsigned)G, threads, 0, stream>>>((const float*)in, ns, C, ld_in, G, (float*)out,
                                                                ld_out);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// vector quantisation (quantizer.py:42-63): per 16-d chunk, argmin_c (|z|^2 + |e_c|^2) - 2 z.e_c
// (first minimum), output z + (e - z).  Codebook (64 KB) + |e|^2 staged in shared memory; all threads of a
// warp read the same code row (smem broadcast).  A broadcast LDS.128 returns 512 B per warp, so the scan is
// bound by shared-memory return bandwidth unless every code row is reused: each thread scans for VQ_R chunks
// at once (the arithmetic per (chunk, code) pair is unchanged).  The VQ_WARPS warps of a CTA work on the SAME
// 32 x VQ_R chunks and split the code range (warp w scans codes [w, w+1) * n_codes / VQ_WARPS in increasing
// order); the partial (distance, code) winners are merged in warp order with a strict <, which keeps the
// first-minimum rule.  This gives 4x the warps of a one-warp-per-chunk-set layout at the same shared-memory traffic.
// ---------------------------------------------------------------------------------------------
#define VQ_DIM 16
#define VQ_WARPS 4
#define VQ_THREADS (32 * VQ_WARPS)
#define VQ_R 4
#define VQ_CHUNKS_PER_CTA (32 * VQ_R)

template <typename InT>
__global__ void __launch_bounds__(VQ_THREADS)
    vq_kernel(const InT* __restrict__ z, long long n_chunks, const float* __restrict__ codebook, int n_codes,
              float* __restrict__ out, int* __restrict__ codes) {
  extern __shared__ float sm[];
  float* cb = sm;                                   // [n_codes][16]
  float* cn = sm + (size_t)n_codes * VQ_DIM;        // [n_codes]
  float* pd = cn + n_codes;                         // [VQ_WARPS][VQ_CHUNKS_PER_CTA] partial best distance
  int* pi = reinterpret_cast<int*>(pd + VQ_WARPS * VQ_CHUNKS_PER_CTA);  // partial best code
  for (int i = threadIdx.x; i < n_codes * VQ_DIM / 4; i += VQ_THREADS)
    reinterpret_cast<float4*>(cb)[i] = reinterpret_cast<const float4*>(codebook)[i];
  __syncthreads();
  for (int c = threadIdx.x; c < n_codes; c += VQ_THREADS) {
    float s = 0.f;
    for (int d = 0; d < VQ_DIM; ++d) s += cb[c * VQ_DIM + d] * cb[c * VQ_DIM + d];
    cn[c] = s;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c_lo = (int)((long long)n_codes * warp / VQ_WARPS), c_hi = (int)((long long)n_codes * (warp + 1) / VQ_WARPS);
  for (long long base = (long long)blockIdx.x * VQ_CHUNKS_PER_CTA; base < n_chunks;
       base += (long long)gridDim.x * VQ_CHUNKS_PER_CTA) {
    float zv[VQ_R][VQ_DIM], zz[VQ_R], best[VQ_R];
    int bi[VQ_R];
#pragma unroll
    for (int r = 0; r < VQ_R; ++r) {
      const long long i = base + r * 32 + lane;
      zz[r] = 0.f, best[r] = INFINITY, bi[r] = c_lo;
#pragma unroll
      for (int d = 0; d < VQ_DIM; ++d) {
        zv[r][d] = i < n_chunks ? (float)z[i * VQ_DIM + d] : 0.f;
        zz[r] += zv[r][d] * zv[r][d];
      }
    }
    for (int c = c_lo; c < c_hi; ++c) {
      const float4* e = reinterpret_cast<const float4*>(cb + c * VQ_DIM);
      const float4 e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
      const float en = cn[c];
#pragma unroll
      for (int r = 0; r < VQ_R; ++r) {
        float dot = 0.f;
        dot += zv[r][0] * e0.x + zv[r][1] * e0.y + zv[r][2] * e0.z + zv[r][3] * e0.w;
        dot += zv[r][4] * e1.x + zv[r][5] * e1.y + zv[r][6] * e1.z + zv[r][7] * e1.w;
        dot += zv[r][8] * e2.x + zv[r][9] * e2.y + zv[r][10] * e2.z + zv[r][11] * e2.w;
        dot += zv[r][12] * e3.x + zv[r][13] * e3.y + zv[r][14] * e3.z + zv[r][15] * e3.w;
        const float dist = (zz[r] + en) - 2.0f * dot;
        if (dist < best[r]) best[r] = dist, bi[r] = c;
      }
    }
    __syncthreads();  // the partial tables of the previous round have been consumed
#pragma unroll
    for (int r = 0; r < VQ_R; ++r) {
      pd[warp * VQ_CHUNKS_PER_CTA + r * 32 + lane] = best[r];
      pi[warp * VQ_CHUNKS_PER_CTA + r * 32 + lane] = bi[r];
    }
    __syncthreads();
    // warp w finalises the chunks r == w: merge the partial winners in code order (strict <: first minimum)
    {
      const int r = warp;
      const int slot = r * 32 + lane;
      const long long i = base + slot;
      float bd = pd[slot];
      int bc = pi[slot];
#pragma unroll
      for (int w2 = 1; w2 < VQ_WARPS; ++w2) {
        const float d2 = pd[w2 * VQ_CHUNKS_PER_CTA + slot];
        if (d2 < bd) bd = d2, bc = pi[w2 * VQ_CHUNKS_PER_CTA + slot];
      }
      if (i < n_chunks) {
        if (codes) codes[i] = bc;
        // z is re-read here (indexing zv[] with the runtime warp id would push the whole array to local memory)
#pragma unroll
        for (int d = 0; d < VQ_DIM; ++d) {
          const float zd = (float)z[i * VQ_DIM + d];
          out[i * VQ_DIM + d] = fadd(zd, fsub(cb[bc * VQ_DIM + d], zd));
        }
      }
    }
  }
}

extern "C" int pfpp_vq(const void* z, int z_is_bf16, long long n_chunks, const float* codebook, int n_codes,
                       float* out, int* codes, cudaStream_t stream) {
  PFPP_CHECK_ARG(z && codebook && out && n_codes > 0);
  if (n_chunks == 0) return PFPP_OK;
  static_assert(VQ_R == VQ_WARPS, "warp w finalises chunk set w");
  size_t smem = (size_t)n_codes * (VQ_DIM + 1) * sizeof(float) + (size_t)VQ_WARPS * VQ_CHUNKS_PER_CTA * 8;
  if (smem > 200 * 1024) return PFPP_EUNSUPPORTED;
  int grid = (int)((n_chunks + VQ_CHUNKS_PER_CTA - 1) / VQ_CHUNKS_PER_CTA);
  if (grid > 148 * 3) grid = 148 * 3;
  if (z_is_bf16) {
    PFPP_ENSURE_SMEM(vq_kernel<__nv_bfloat16>, smem);
    vq_kernel<__nv_bfloat16><<<grid, VQ_THREADS, smem, stream>>>((const __nv_bfloat16*)z, n_chunks, codebook, n_codes,
                                                                 out, codes);
  } else {
    PFPP_ENSURE_SMEM(vq_kernel<float>, smem);
    vq_kernel<float><<<grid, VQ_THREADS, smem, stream>>>((const float*)z, n_chunks, codebook, n_codes, out, codes);
  }
  PFPP_RETURN_LAST();
}
