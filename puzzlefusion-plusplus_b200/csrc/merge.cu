// Merge-stage geometry (utils/node_merge_utils.py:159-222, remove_intersect_points_and_fps_ds):
//   1. per-point normals by kNN-20 PCA (pytorch3d estimate_pointcloud_normals),
//   2. for every ordered pair of member clouds (i, j): drop point k of cloud i when
//      NN^2(i_k -> j) + NN^2(j_k -> i) < threshold (index-aligned sum, App. C.6) and the normals
//      n_i[k], n_j[k] oppose each other.
// Output is the keep-mask; compaction / FPS / renormalisation follow on the caller's stream.
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

__global__ void __launch_bounds__(128)
    intersect_kernel(const float* __restrict__ pcs, const float* __restrict__ normals, int N, int Pc, float thr,
                     unsigned char* __restrict__ keep) {
  extern __shared__ float sm[];
  float* ax = sm;
  float* ay = ax + N;
  float* az = ay + N;
  float* bx = az + N;
  float* by = bx + N;
  float* bz = by + N;
  const int i = blockIdx.y / Pc, j = blockIdx.y % Pc;
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
    intersect_kernel<<<g2, 128, smem_i, stream>>>(pcs, normals, n_points, n_clouds, threshold, keep);
  }
  PFPP_RETURN_LAST();
}
