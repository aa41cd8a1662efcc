// Merge-stage geometry (utils/node_merge_utils.py:159-222, remove_intersect_points_and_fps_ds):
//   1. per-point normals by kNN-20 PCA (pytorch3d estimate_pointcloud_normals),
//   2. for every ordered pair of member clouds (i, j): drop point k of cloud i when
//      NN^2(i_k -> j) + NN^2(j_k -> i) < threshold (index-aligned sum, App. C.6) and the normals
//      n_i[k], n_j[k] oppose each other.
// pfpp_merge_filter returns the keep-mask of ONE component (compaction / FPS / renormalisation on the caller's
// stream); pfpp_merge runs the whole merge stage (auto_aggl.py:234-286) for ALL components of ALL objects of a
// batch in one asynchronous launch sequence with no host round trip.
#include "common.cuh"
#include "../../include/pfpp.h"

#define MERGE_MAX_KNN 32

// smallest-eigenvalue eigenvector of a symmetric 3x3 matrix, cyclic Jacobi in fp32
__device__ void smallest_eigvec3(float a00, float a01, float a02, float a11, float a12, float a22, float* v) {
  float A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 8; ++sweep) {
    float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[1][2]);
    if (off < 1e-30f) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        float apq = A[p][q];
        if (fabsf(apq) < 1e-38f) continue;
        float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
        float t = (theta >= 0 ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
        float c = rsqrtf(t * t + 1.0f), s = t * c;
        for (int k = 0; k < 3; ++k) {
          float akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          float apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          float vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (A[1][1] < A[m][m]) m = 1;
  if (A[2][2] < A[m][m]) m = 2;
  float n = rsqrtf(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
  v[0] = V[0][m] * n, v[1] = V[1][m] * n, v[2] = V[2][m] * n;
}

__global__ void __launch_bounds__(128)
    normals_kernel(const float* __restrict__ pcs, int N, int K, float* __restrict__ normals) {
  extern __shared__ float sm[];
  float* sx = sm;
  float* sy = sm + N;
  float* sz = sm + 2 * N;
  __shared__ float mean[3];
  const int cloud = blockIdx.y;
  const float* p = pcs + (size_t)cloud * N * 3;
  // centre the cloud (estimate_pointcloud_normals does; translation-invariant up to rounding)
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) s += p[3 * i + threadIdx.x];
    mean[threadIdx.x] = s / (float)N;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sx[i] = p[3 * i] - mean[0], sy[i] = p[3 * i + 1] - mean[1], sz[i] = p[3 * i + 2] - mean[2];
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float px = sx[i], py = sy[i], pz = sz[i];
  float bd[MERGE_MAX_KNN];
  int bi[MERGE_MAX_KNN];
  for (int k = 0; k < K; ++k) bd[k] = INFINITY, bi[k] = 0;
  for (int j = 0; j < N; ++j) {
    float dx = px - sx[j], dy = py - sy[j], dz = pz - sz[j];
    float d = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    if (d < bd[K - 1]) {  // insertion into the sorted list (stable: earlier index wins ties)
      int k = K - 1;
      while (k > 0 && bd[k - 1] > d) {
        bd[k] = bd[k - 1], bi[k] = bi[k - 1];
        --k;
      }
      bd[k] = d, bi[k] = j;
    }
  }
  float mx = 0, my = 0, mz = 0;
  for (int k = 0; k < K; ++k) mx += sx[bi[k]], my += sy[bi[k]], mz += sz[bi[k]];
  mx /= K, my /= K, mz /= K;
  float c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
  for (int k = 0; k < K; ++k) {
    float dx = sx[bi[k]] - mx, dy = sy[bi[k]] - my, dz = sz[bi[k]] - mz;
    c00 += dx * dx, c01 += dx * dy, c02 += dx * dz, c11 += dy * dy, c12 += dy * dz, c22 += dz * dz;
  }
  float inv = 1.0f / K;
  float n[3];
  smallest_eigvec3(c00 * inv, c01 * inv, c02 * inv, c11 * inv, c12 * inv, c22 * inv, n);
  int pos = 0;
  for (int k = 0; k < K; ++k) {
    float pr = (sx[bi[k]] - px) * n[0] + (sy[bi[k]] - py) * n[1] + (sz[bi[k]] - pz) * n[2];
    pos += pr > 0.f;
  }
  float sgn = ((float)pos < 0.5f * K) ? -1.0f : 1.0f;
  float* o = normals + ((size_t)cloud * N + i) * 3;
  o[0] = sgn * n[0], o[1] = sgn * n[1], o[2] = sgn * n[2];
}

// pair_i/pair_j != nullptr: blockIdx.y indexes an explicit list of ordered cloud pairs (batched merge); otherwise
// blockIdx.y = i*Pc + j over the Pc clouds of one component.
__global__ void __launch_bounds__(128)
    intersect_kernel(const float* __restrict__ pcs, const float* __restrict__ normals, int N, int Pc, float thr,
                     const int* __restrict__ pair_i, const int* __restrict__ pair_j,
                     unsigned char* __restrict__ keep) {
  extern __shared__ float sm[];
  float* ax = sm;
  float* ay = ax + N;
  float* az = ay + N;
  float* bx = az + N;
  float* by = bx + N;
  float* bz = by + N;
  const int i = pair_i ? pair_i[blockIdx.y] : blockIdx.y / Pc, j = pair_j ? pair_j[blockIdx.y] : blockIdx.y % Pc;
  if (i == j) return;
  const float* a = pcs + (size_t)i * N * 3;
  const float* b = pcs + (size_t)j * N * 3;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    ax[k] = a[3 * k], ay[k] = a[3 * k + 1], az[k] = a[3 * k + 2];
    bx[k] = b[3 * k], by[k] = b[3 * k + 1], bz[k] = b[3 * k + 2];
  }
  __syncthreads();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  float d1 = INFINITY, d2 = INFINITY;
  float pax = ax[k], pay = ay[k], paz = az[k], pbx = bx[k], pby = by[k], pbz = bz[k];
  for (int t = 0; t < N; ++t) {
    float dx = fsub(pax, bx[t]), dy = fsub(pay, by[t]), dz = fsub(paz, bz[t]);
    d1 = fminf(d1, fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
    dx = fsub(pbx, ax[t]), dy = fsub(pby, ay[t]), dz = fsub(pbz, az[t]);
    d2 = fminf(d2, fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
  }
  if (fadd(d1, d2) < thr) {
    const float* ni = normals + ((size_t)i * N + k) * 3;
    const float* nj = normals + ((size_t)j * N + k) * 3;
    float dot = fadd(fadd(fmul(ni[0], nj[0]), fmul(ni[1], nj[1])), fmul(ni[2], nj[2]));
    if (dot < 0.f) keep[(size_t)i * N + k] = 0;
  }
}

extern "C" int pfpp_merge_filter(const float* pcs, int n_clouds, int n_points, int knn, float threshold,
                                 unsigned char* keep, float* normals, cudaStream_t stream) {
  PFPP_CHECK_ARG(pcs && keep && normals && n_clouds > 0 && n_points > 0 && knn > 0 && knn <= MERGE_MAX_KNN &&
                 knn <= n_points);
  cudaError_t e = cudaMemsetAsync(keep, 1, (size_t)n_clouds * n_points, stream);
  if (e != cudaSuccess) return (int)e;
  size_t smem_n = sizeof(float) * 3 * (size_t)n_points, smem_i = 2 * smem_n;
  if (smem_i > 200 * 1024) return PFPP_EUNSUPPORTED;
  PFPP_ENSURE_SMEM(normals_kernel, smem_n);
  PFPP_ENSURE_SMEM(intersect_kernel, smem_i);
  dim3 g1(pfpp_cdiv(n_points, 128), n_clouds);
  normals_kernel<<<g1, 128, smem_n, stream>>>(pcs, n_points, knn, normals);
  if (n_clouds > 1) {
    dim3 g2(pfpp_cdiv(n_points, 128), n_clouds * n_clouds);
    intersect_kernel<<<g2, 128, smem_i, stream>>>(pcs, normals, n_points, n_clouds, threshold, nullptr, nullptr, keep);
  }
  PFPP_RETURN_LAST();
}


// ---------------------------------------------------------------------------------------------
// Batched merge stage (auto_aggl.py:234-286; utils/node_merge_utils.py:125-135,159-222) for every connected
// component of every object at once.  Tables are built by the host from the agglomeration graph:
//   comp_start [n_comp+1] -> member_slot [n_clouds]: the valid member fragments of component c in concatenation
//   order (their posed clouds are rows of `posed` [slots, N, 3]); comp_pivot_slot [n_comp]; pair_i / pair_j
//   [n_pairs]: every ordered pair of member clouds of the same component (indices into the member list);
//   uniform [n_comp]: the torch.rand(1) draw of node_merge_utils.py:219 (random_start FPS).
// Kernel sequence (all on `stream`, nothing returns to the host):
//   centroid (fp64 accumulation) -> recentre + concatenate -> kNN-20 PCA normals -> pairwise intersect filter ->
//   ordered compaction per component -> FPS metadata (start = int(u*M), n = ceil(M * float(N/M))) ->
//   ragged FPS -> gather N samples, max-abs scale, write the pivot slot of part_pcs / scale.
// Results the host needs later (centroid for init_pose, max-abs scale, kept-point count) land in `result`
// [n_comp, 8] = {cx, cy, cz, mscale, kept, n_out, start, 0} and are read back with the next pose download.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
    merge_centroid_kernel(const float* __restrict__ posed, int N, const int* __restrict__ comp_start,
                          const int* __restrict__ member_slot, float* __restrict__ result) {
  const int c = blockIdx.x;
  const int m0 = comp_start[c], m1 = comp_start[c + 1];
  double sx = 0, sy = 0, sz = 0;
  for (int m = m0; m < m1; ++m) {
    const float* p = posed + (size_t)member_slot[m] * N * 3;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sx += p[3 * i], sy += p[3 * i + 1], sz += p[3 * i + 2];
  }
  __shared__ double red[3][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) red[0][warp] = sx, red[1][warp] = sy, red[2][warp] = sz;
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    result[c * 8 + threadIdx.x] = (float)(s / (double)((long long)(m1 - m0) * N));
  }
}

// merged[m] = posed[member_slot[m]] - centroid[comp of m]  (grid: n_clouds)
__global__ void merge_recentre_kernel(const float* __restrict__ posed, int N, const int* __restrict__ member_slot,
                                      const int* __restrict__ member_comp, const float* __restrict__ result,
                                      float* __restrict__ merged) {
  const int m = blockIdx.x;
  const float* ctr = result + 8 * member_comp[m];
  const float cx = ctr[0], cy = ctr[1], cz = ctr[2];
  const float* p = posed + (size_t)member_slot[m] * N * 3;
  float* o = merged + (size_t)m * N * 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    o[3 * i] = fsub(p[3 * i], cx), o[3 * i + 1] = fsub(p[3 * i + 1], cy), o[3 * i + 2] = fsub(p[3 * i + 2], cz);
  }
}

// ordered compaction of the kept points of component c (concatenation order) + FPS metadata
__global__ void __launch_bounds__(1024)
    merge_compact_kernel(const float* __restrict__ merged, const unsigned char* __restrict__ keep, int N,
                         const int* __restrict__ comp_start, const float* __restrict__ uniform,
                         float* __restrict__ compact, float* __restrict__ result, int* __restrict__ fps_meta,
                         int n_comp) {
  const int c = blockIdx.x;
  const long long p0 = (long long)comp_start[c] * N, p1 = (long long)comp_start[c + 1] * N;
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (long long s = p0; s < p1; s += blockDim.x) {
    const long long i = s + threadIdx.x;
    const bool k = i < p1 && keep[i];
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (k) {
      const long long o = p0 + off + __popc(bal & ((1u << lane) - 1u));
      compact[3 * o] = merged[3 * i], compact[3 * o + 1] = merged[3 * i + 1], compact[3 * o + 2] = merged[3 * i + 2];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < nw; ++w) t += wsum[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int M = base_s;
    // node_merge_utils.py:216-220: ratio = float32(N / M) (Python double division), n = ceil(float32(M) * ratio),
    // start = int64(rand(1) * float(M))
    const float ratio = (float)((double)N / (double)(M > 0 ? M : 1));
    int n_out = (int)ceilf(fmul((float)M, ratio));
    int start = (int)fmul(uniform[c], (float)M);
    if (start >= M) start = M > 0 ? M - 1 : 0;
    int n_s = n_out < N ? n_out : N;  // only the first N samples are used (nmu:221)
    if (n_s > M) n_s = M;
    // fps_meta columns (each [n_comp]): cloud_start, cloud_len, n_samples, start, out_start
    fps_meta[c] = (int)p0;
    fps_meta[n_comp + c] = M;
    fps_meta[2 * n_comp + c] = n_s;
    fps_meta[3 * n_comp + c] = start;
    fps_meta[4 * n_comp + c] = c * N;
    result[c * 8 + 4] = (float)M;
    result[c * 8 + 5] = (float)n_out;
    result[c * 8 + 6] = (float)start;
    result[c * 8 + 7] = 0.f;
  }
}

// ds = compact[fps idx][:N]; mscale = max |ds|; part_pcs[pivot] = ds / mscale; scale[pivot] = mscale
__global__ void __launch_bounds__(256)
    merge_finalize_kernel(const float* __restrict__ compact, const int* __restrict__ fps_meta,
                          const int* __restrict__ fps_idx, int N, int n_comp, const int* __restrict__ comp_pivot_slot,
                          float* __restrict__ part_pcs, float* __restrict__ scale, float* __restrict__ result) {
  const int c = blockIdx.x;
  const float* src = compact + 3 * (size_t)fps_meta[c];
  const int n_s = fps_meta[2 * n_comp + c];
  const int* idx = fps_idx + (size_t)c * N;
  float mx = 0.f;
  for (int i = threadIdx.x; i < n_s; i += blockDim.x) {
    const float* p = src + 3 * (size_t)idx[i];
    mx = fmaxf(mx, fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2]))));
  }
  __shared__ float red[8];
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  float* o = part_pcs + (size_t)comp_pivot_slot[c] * N * 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    // fewer than N kept points (degenerate; the reference would fail on the shape): pad with the first sample
    const float* p = src + 3 * (size_t)idx[i < n_s ? i : 0];
    o[3 * i] = fdiv(p[0], mx), o[3 * i + 1] = fdiv(p[1], mx), o[3 * i + 2] = fdiv(p[2], mx);
  }
  if (threadIdx.x == 0) {
    scale[comp_pivot_slot[c]] = mx;
    result[c * 8 + 3] = mx;
  }
}

// by-area clouds of merged nodes (auto_aggl.py:259-262): dst[seg] = src[seg] - centroid[seg_comp]
__global__ void segment_shift_kernel(const float* __restrict__ src, const int* __restrict__ seg_start,
                                     const int* __restrict__ seg_len, const int* __restrict__ seg_comp,
                                     const float* __restrict__ result, float* __restrict__ dst) {
  const int s = blockIdx.x;
  const float* ctr = result + 8 * seg_comp[s];
  const float cx = ctr[0], cy = ctr[1], cz = ctr[2];
  const size_t st = seg_start[s];
  for (int i = threadIdx.x; i < seg_len[s]; i += blockDim.x) {
    const float* p = src + 3 * (st + i);
    float* o = dst + 3 * (st + i);
    o[0] = fsub(p[0], cx), o[1] = fsub(p[1], cy), o[2] = fsub(p[2], cz);
  }
}

extern "C" size_t pfpp_merge_workspace_bytes(int n_clouds, int n_comp, int n_points) {
  const size_t pts = (size_t)n_clouds * n_points;
  // merged, normals, compact [pts,3] f32; dist [pts] f32; keep [pts] u8 (padded); fps_meta [5,n_comp]; fps_idx [n_comp,N]
  return pts * 3 * 4 * 3 + pts * 4 + ((pts + 255) / 256) * 256 + (size_t)n_comp * 5 * 4 + (size_t)n_comp * n_points * 4 + 1024;
}

extern "C" int pfpp_merge(const float* posed, int n_points, int n_comp, int n_clouds, const int* comp_start,
                          const int* member_slot, const int* member_comp, const int* comp_pivot_slot, int n_pairs,
                          const int* pair_i, const int* pair_j, const float* uniform, float threshold, int knn,
                          int n_area_segs, const int* area_seg_start, const int* area_seg_len,
                          const int* area_seg_comp, const float* by_area_posed, float* by_area, float* part_pcs,
                          float* scale, float* result, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  PFPP_CHECK_ARG(n_comp >= 0 && n_clouds >= 0 && n_points > 0 && knn > 0 && knn <= MERGE_MAX_KNN && knn <= n_points);
  if (n_comp == 0) return PFPP_OK;
  PFPP_CHECK_ARG(posed && comp_start && member_slot && member_comp && comp_pivot_slot && uniform && part_pcs && scale &&
                 result && workspace && (n_pairs == 0 || (pair_i && pair_j)));
  if (ws_bytes < pfpp_merge_workspace_bytes(n_clouds, n_comp, n_points)) return PFPP_EWORKSPACE;
  const int N = n_points;
  const size_t pts = (size_t)n_clouds * N;
  char* w = (char*)workspace;
  float* merged = (float*)w;
  w += pts * 12;
  float* normals = (float*)w;
  w += pts * 12;
  float* compact = (float*)w;
  w += pts * 12;
  float* dist = (float*)w;
  w += pts * 4;
  int* fps_meta = (int*)w;
  w += (size_t)n_comp * 5 * 4;
  int* fps_idx = (int*)w;
  w += (size_t)n_comp * N * 4;
  unsigned char* keep = (unsigned char*)w;
  size_t smem_n = sizeof(float) * 3 * (size_t)N, smem_i = 2 * smem_n;
  if (smem_i > 200 * 1024) return PFPP_EUNSUPPORTED;
  PFPP_ENSURE_SMEM(normals_kernel, smem_n);
  PFPP_ENSURE_SMEM(intersect_kernel, smem_i);
  cudaError_t e = cudaMemsetAsync(keep, 1, pts, stream);
  if (e != cudaSuccess) return (int)e;
  merge_centroid_kernel<<<n_comp, 512, 0, stream>>>(posed, N, comp_start, member_slot, result);
  merge_recentre_kernel<<<n_clouds, 256, 0, stream>>>(posed, N, member_slot, member_comp, result, merged);
  if (n_area_segs > 0)
    segment_shift_kernel<<<n_area_segs, 256, 0, stream>>>(by_area_posed, area_seg_start, area_seg_len, area_seg_comp,
                                                          result, by_area);
  normals_kernel<<<dim3(pfpp_cdiv(N, 128), n_clouds), 128, smem_n, stream>>>(merged, N, knn, normals);
  for (int p0 = 0; p0 < n_pairs; p0 += 32768) {  // gridDim.y <= 65535
    const int np = n_pairs - p0 < 32768 ? n_pairs - p0 : 32768;
    intersect_kernel<<<dim3(pfpp_cdiv(N, 128), np), 128, smem_i, stream>>>(merged, normals, N, 0, threshold, pair_i + p0,
                                                                           pair_j + p0, keep);
  }
  merge_compact_kernel<<<n_comp, 1024, 0, stream>>>(merged, keep, N, comp_start, uniform, compact, result, fps_meta,
                                                    n_comp);
  int rc = pfpp_fps_ragged(compact, fps_meta, fps_meta + n_comp, fps_meta + 2 * n_comp, fps_meta + 3 * n_comp, n_comp,
                           dist, fps_meta + 4 * n_comp, fps_idx, stream);
  if (rc != 0) return rc;
  merge_finalize_kernel<<<n_comp, 256, 0, stream>>>(compact, fps_meta, fps_idx, N, n_comp, comp_pivot_slot, part_pcs,
                                                    scale, result);
  PFPP_RETURN_LAST();
}
