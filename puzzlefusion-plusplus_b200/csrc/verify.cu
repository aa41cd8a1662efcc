// Verify-stage geometry: per-edge matched-point Chamfer histogram -> verifier edge features
// (auto_aggl.py:181-201,385-389; utils/node_merge_utils.py:62-89; chamferdist semantics App. B.4).
//
// One CTA per pre-computed edge.  The matched fracture points of both fragments (K_e pairs, already
// posed by pfpp_pose_apply) are gathered into shared memory; thread k computes
//   cd_k = min_j |a_k - b_j|^2 + min_j |b_k - a_j|^2      (index-aligned sum, as the reference does)
// with individually rounded fp32 ops, buckets it into [0,1e-3,5e-3,1e-2,5e-2,1e-1,100) and the CTA
// emits the 7 features [count_i / max(total,1) (6), total].  Replaces <=190 Python iterations.
#include "common.cuh"
#include "../../include/pfpp.h"

#define EDGE_THREADS 128

__global__ void __launch_bounds__(EDGE_THREADS)
    edge_features_kernel(const float* __restrict__ pts, const int* __restrict__ pair_src,
                         const int* __restrict__ pair_tgt, const int* __restrict__ edge_start,
                         const int* __restrict__ edge_len, const int* __restrict__ edge_row, int max_pairs,
                         float* __restrict__ feat) {
  extern __shared__ float sm[];
  float* ax = sm;
  float* ay = ax + max_pairs;
  float* az = ay + max_pairs;
  float* bx = az + max_pairs;
  float* by = bx + max_pairs;
  float* bz = by + max_pairs;
  __shared__ int hist[6];
  const int e = blockIdx.x, st = edge_start[e], n = edge_len[e];
  if (threadIdx.x < 6) hist[threadIdx.x] = 0;
  for (int k = threadIdx.x; k < n; k += EDGE_THREADS) {
    const float* a = pts + 3 * (size_t)pair_src[st + k];
    const float* b = pts + 3 * (size_t)pair_tgt[st + k];
    ax[k] = a[0], ay[k] = a[1], az[k] = a[2];
    bx[k] = b[0], by[k] = b[1], bz[k] = b[2];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += EDGE_THREADS) {
    float d1 = INFINITY, d2 = INFINITY;
    float pax = ax[k], pay = ay[k], paz = az[k], pbx = bx[k], pby = by[k], pbz = bz[k];
    for (int j = 0; j < n; ++j) {
      float dx = fsub(pax, bx[j]), dy = fsub(pay, by[j]), dz = fsub(paz, bz[j]);
      d1 = fminf(d1, fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
      dx = fsub(pbx, ax[j]), dy = fsub(pby, ay[j]), dz = fsub(pbz, az[j]);
      d2 = fminf(d2, fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
    }
    float cd = fadd(d1, d2);
    // torch.bucketize(right=True) then counts[1:7]
    const float b[7] = {0.0f, 1e-3f, 5e-3f, 1e-2f, 5e-2f, 1e-1f, 100.0f};
    int bin = -1;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (cd >= b[i] && cd < b[i + 1]) bin = i;
    if (bin >= 0) atomicAdd(&hist[bin], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < 6; ++i) tot += hist[i];
    float* o = feat + 7 * (size_t)edge_row[e];
    float den = (float)(tot == 0 ? 1 : tot);
    for (int i = 0; i < 6; ++i) o[i] = fdiv((float)hist[i], den);
    o[6] = (float)tot;
  }
}

extern "C" int pfpp_edge_features(const float* pts, const int* pair_src, const int* pair_tgt, const int* edge_start,
                                  const int* edge_len, const int* edge_row, int n_edges, int max_pairs,
                                  long long n_rows, float* feat, cudaStream_t stream) {
  PFPP_CHECK_ARG(feat && n_rows >= 0 && n_edges >= 0);
  cudaError_t e = cudaMemsetAsync(feat, 0, sizeof(float) * 7 * (size_t)n_rows, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_edges == 0) return PFPP_OK;
  PFPP_CHECK_ARG(pts && pair_src && pair_tgt && edge_start && edge_len && edge_row && max_pairs > 0);
  size_t smem = sizeof(float) * 6 * (size_t)max_pairs;
  if (smem > 200 * 1024) return PFPP_EUNSUPPORTED;
  PFPP_ENSURE_SMEM(edge_features_kernel, smem);
  edge_features_kernel<<<n_edges, EDGE_THREADS, smem, stream>>>(pts, pair_src, pair_tgt, edge_start, edge_len, edge_row,
                                                                 max_pairs, feat);
  PFPP_RETURN_LAST();
}

// Verifier token assembly (verifier_transformer.py:49-56): tok = Linear(7->C)(feat) + [pe[i], pe[j]]
// for the packed valid edges; done here as a tiny SIMT kernel (K = 7 is not GEMM-shaped).
__global__ void verifier_embed_kernel(const float* __restrict__ feat, const int* __restrict__ tok_row,
                                      const int* __restrict__ tok_i, const int* __restrict__ tok_j,
                                      const float* __restrict__ W, const float* __restrict__ bias,
                                      const float* __restrict__ pe, int C, float* __restrict__ out) {
  int t = blockIdx.x;
  const float* f = feat + 7 * (size_t)tok_row[t];
  float fv[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) fv[i] = f[i];
  int half = C / 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) a = fmaf(fv[i], W[c * 7 + i], a);
    a += bias[c];
    float p = c < half ? pe[(size_t)tok_i[t] * half + c] : pe[(size_t)tok_j[t] * half + (c - half)];
    out[(size_t)t * C + c] = p + a;
  }
}

extern "C" int pfpp_verifier_embed(const float* feat, const int* tok_row, const int* tok_i, const int* tok_j,
                                   int n_tokens, const float* W, const float* bias, const float* pe, int C,
                                   float* out, cudaStream_t stream) {
  PFPP_CHECK_ARG(feat && tok_row && tok_i && tok_j && W && bias && pe && out);
  if (n_tokens == 0) return PFPP_OK;
  verifier_embed_kernel<<<n_tokens, 128, 0, stream>>>(feat, tok_row, tok_i, tok_j, W, bias, pe, C, out);
  PFPP_RETURN_LAST();
}

// logits[row] = tok . w + b for the packed tokens, scattered back to the dense [B, E] layout
// (verifier_transformer.py:63); one warp per token.
__global__ void verifier_head_kernel(const float* __restrict__ h, const int* __restrict__ tok_row, int n_tokens,
                                     const float* __restrict__ w, const float* __restrict__ b, int C,
                                     float* __restrict__ logits) {
  int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tokens) return;
  int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(h[(size_t)t * C + c], w[c], s);
  s = warp_sum(s);
  if (lane == 0) logits[tok_row[t]] = s + b[0];
}

extern "C" int pfpp_verifier_head(const float* h, const int* tok_row, int n_tokens, const float* w, const float* b,
                                  int C, float* logits, cudaStream_t stream) {
  PFPP_CHECK_ARG(h && tok_row && w && b && logits);
  if (n_tokens == 0) return PFPP_OK;
  verifier_head_kernel<<<pfpp_cdiv(n_tokens, 4), 128, 0, stream>>>(h, tok_row, n_tokens, w, b, C, logits);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// Brute-force nearest-neighbour squared distance between batched clouds (chamferdist / pytorch3d
// knn_points K=1 semantics): out[b,i] = min_j |a[b,i]-b[b,j]|^2.
// Used by the evaluation metrics (evaluator.py:108,137).  The target cloud streams through shared
// memory in 1024-point tiles; each thread owns one query point.  ((dx^2+dy^2)+dz^2, ops rounded
// individually, so the result is bit-identical to the oracle).
// ---------------------------------------------------------------------------------------------
#define NN_TILE 1024
#define NN_Q 4  // query points per thread: every target point read from shared memory serves 4 distance evaluations

// The kernel is issue-bound (9 individually rounded flops + 1 min per pair); the target tile is stored as float4
// so that one broadcast LDS.128 feeds NN_Q x 10 arithmetic instructions.
__global__ void __launch_bounds__(256)
    nn_sqdist_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int M, float* __restrict__ out) {
  __shared__ float4 bt[NN_TILE];
  const int batch = blockIdx.y;
  const int i0 = blockIdx.x * (256 * NN_Q) + threadIdx.x;  // this thread's queries: i0 + q * 256
  float ax[NN_Q], ay[NN_Q], az[NN_Q], best[NN_Q];
#pragma unroll
  for (int q = 0; q < NN_Q; ++q) {
    const int i = i0 + q * 256;
    const float* pa = a + ((size_t)batch * N + (i < N ? i : 0)) * 3;
    ax[q] = pa[0], ay[q] = pa[1], az[q] = pa[2];
    best[q] = INFINITY;
  }
  const float* pb = b + (size_t)batch * M * 3;
  for (int t0 = 0; t0 < M; t0 += NN_TILE) {
    const int nt = min(NN_TILE, M - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < nt; j += blockDim.x)
      bt[j] = make_float4(pb[(size_t)(t0 + j) * 3], pb[(size_t)(t0 + j) * 3 + 1], pb[(size_t)(t0 + j) * 3 + 2], 0.f);
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < nt; ++j) {
      const float4 t = bt[j];
#pragma unroll
      for (int q = 0; q < NN_Q; ++q) {
        const float dx = fsub(ax[q], t.x), dy = fsub(ay[q], t.y), dz = fsub(az[q], t.z);
        best[q] = fminf(best[q], fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NN_Q; ++q) {
    const int i = i0 + q * 256;
    if (i < N) out[(size_t)batch * N + i] = best[q];
  }
}

extern "C" int pfpp_nn_sqdist(const float* a, const float* b, int batches, int N, int M, float* out,
                              cudaStream_t stream) {
  PFPP_CHECK_ARG(a && b && out && N > 0 && M > 0 && batches >= 0);
  if (batches == 0) return PFPP_OK;
  dim3 grid(pfpp_cdiv(N, 256 * NN_Q), batches);
  nn_sqdist_kernel<<<grid, 256, 0, stream>>>(a, b, N, M, out);
  PFPP_RETURN_LAST();
}
