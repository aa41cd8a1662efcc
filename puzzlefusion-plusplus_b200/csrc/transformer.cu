// Token-level kernels of the denoiser / verifier transformers (fp32 SIMT):
// NeRF embeddings + token assembly, (Ada)LayerNorm, variable-length multi-head attention,
// mean-pool, DDPM posterior step.  The contractions around them go through pfpp_gemm_*.
#include "common.cuh"
#include "../../include/pfpp.h"

// ---------------------------------------------------------------------------------------------
// NeRF embedding features (utils/model_utils.py:40-69; denoiser_transformer.py:117-135)
//   token features  [latent(64) | nerf(xyz) 63 | nerf(scale) 21 | 0-pad]  -> feat_tok [F*L, ld_tok]
//   param features  [nerf(x) 147 | 0-pad]                                 -> feat_par [F,   ld_par]
// nerf(v) = [v, sin(v*2^0), cos(v*2^0), ..., sin(v*2^9), cos(v*2^9)] with v a d-vector, i.e. the
// output is ordered [v(d) | sin(f0 v)(d) | cos(f0 v)(d) | ...].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float nerf_feature(const float* v, int d, int c) {
  // c in [0, d*21)
  if (c < d) return v[c];
  int r = c - d;
  int blk = r / d, comp = r - blk * d;
  float f = (float)(1 << (blk >> 1));
  float a = fmul(v[comp], f);
  return (blk & 1) ? cosf(a) : sinf(a);
}

template <typename OutT>
__device__ __forceinline__ OutT cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// bf16 hi/lo split of an fp32 value (operand format of the fp32-grade bf16x3 contraction, gemm_tc.cu):
// hi = bf16(v), lo = bf16(v - hi); hi + lo carries 16 mantissa bits of v.
__device__ __forceinline__ void split_store(__nv_bfloat16* o, int lo_off, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  o[0] = h;
  o[lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// x [rows, C] fp32 (leading dimension ldx) -> out [rows, ldo] bf16: hi at column c, lo at ldo/2 + c, zero padding
__global__ void split_bf16_kernel(const float* __restrict__ x, long long rows, int C, int ldx, __nv_bfloat16* __restrict__ out,
                                  int ldo) {
  const int half = ldo >> 1;
  const long long total = rows * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int c = (int)(i - r * half);
    const float v = c < C ? x[r * ldx + c] : 0.f;
    split_store(out + r * ldo + c, half, v);
  }
}

extern "C" int pfpp_split_bf16(const float* x, long long rows, int C, int ldx, void* out, int ldo, cudaStream_t stream) {
  PFPP_CHECK_ARG(x && out && rows >= 0 && C > 0 && ldx >= C && (ldo % 2) == 0 && ldo / 2 >= C);
  if (rows == 0) return PFPP_OK;
  long long total = rows * (ldo / 2);
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  split_bf16_kernel<<<grid, 256, 0, stream>>>(x, rows, C, ldx, (__nv_bfloat16*)out, ldo);
  PFPP_RETURN_LAST();
}

// SPLIT: OutT is bf16 and every value is written as a hi/lo pair (lo at column offset ld/2 of the same row)
template <typename OutT, bool SPLIT>
__device__ __forceinline__ void put_out(OutT* o, int lo_off, float v) {
  if constexpr (SPLIT) split_store(o, lo_off, v);
  else *o = cvt_out<OutT>(v);
}

template <typename OutT, bool SPLIT>
__global__ void embed_features_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      const int* __restrict__ frag_slot, const float* __restrict__ latent,
                                      const float* __restrict__ xyz, int F, int L, int latent_dim,
                                      OutT* __restrict__ feat_tok, int ld_tok, OutT* __restrict__ feat_par,
                                      int ld_par) {
  int row = blockIdx.x;  // 0..F*L-1 : token rows ; F*L..F*L+F-1 : param rows
  if (row < F * L) {
    int f = row / L;
    int slot = frag_slot[f];
    float p[3] = {xyz[(size_t)row * 3], xyz[(size_t)row * 3 + 1], xyz[(size_t)row * 3 + 2]};
    float s = scale[slot];
    OutT* o = feat_tok + (size_t)row * ld_tok;
    const int w_tok = SPLIT ? ld_tok / 2 : ld_tok;
    for (int c = threadIdx.x; c < w_tok; c += blockDim.x) {
      float v = 0.f;
      if (c < latent_dim) v = latent[(size_t)row * latent_dim + c];
      else if (c < latent_dim + 63) v = nerf_feature(p, 3, c - latent_dim);
      else if (c < latent_dim + 84) v = nerf_feature(&s, 1, c - latent_dim - 63);
      put_out<OutT, SPLIT>(o + c, w_tok, v);
    }
  } else {
    int f = row - F * L;
    int slot = frag_slot[f];
    float v7[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) v7[i] = x[(size_t)slot * 7 + i];
    OutT* o = feat_par + (size_t)f * ld_par;
    const int w_par = SPLIT ? ld_par / 2 : ld_par;
    for (int c = threadIdx.x; c < w_par; c += blockDim.x)
      put_out<OutT, SPLIT>(o + c, w_par, c < 147 ? nerf_feature(v7, 7, c) : 0.f);
  }
}

extern "C" int pfpp_embed_features(const float* x, const float* scale, const int* frag_slot, const float* latent,
                                   const float* xyz, int F, int L, int latent_dim, int out_bf16, void* feat_tok,
                                   int ld_tok, void* feat_par, int ld_par, cudaStream_t stream) {
  PFPP_CHECK_ARG(x && scale && frag_slot && latent && xyz && feat_tok && feat_par);
  const int div = out_bf16 == 2 ? 2 : 1;  // out_bf16 == 2: bf16 hi/lo split rows of ld_* elements (two halves)
  PFPP_CHECK_ARG(ld_tok / div >= latent_dim + 84 && ld_par / div >= 147 && out_bf16 >= 0 && out_bf16 <= 2);
  if (F == 0) return PFPP_OK;
  if (out_bf16 == 2)
    embed_features_kernel<__nv_bfloat16, true><<<F * L + F, 64, 0, stream>>>(x, scale, frag_slot, latent, xyz, F, L,
                                                                            latent_dim, (__nv_bfloat16*)feat_tok, ld_tok,
                                                                            (__nv_bfloat16*)feat_par, ld_par);
  else if (out_bf16)
    embed_features_kernel<__nv_bfloat16, false><<<F * L + F, 64, 0, stream>>>(x, scale, frag_slot, latent, xyz, F, L,
                                                                             latent_dim, (__nv_bfloat16*)feat_tok, ld_tok,
                                                                             (__nv_bfloat16*)feat_par, ld_par);
  else
    embed_features_kernel<float, false><<<F * L + F, 64, 0, stream>>>(x, scale, frag_slot, latent, xyz, F, L, latent_dim,
                                                                     (float*)feat_tok, ld_tok, (float*)feat_par, ld_par);
  PFPP_RETURN_LAST();
}

// h[f*L+l] = shape_emb[f*L+l] + (x_emb[f] + ref_emb[ref[slot]]) + pe[slot % P]
// (denoiser_transformer.py:150-156,173-185; same association order as the reference)
__global__ void combine_embed_kernel(const float* __restrict__ shape_emb, const float* __restrict__ x_emb,
                                     const float* __restrict__ ref_emb, const float* __restrict__ pe,
                                     const int* __restrict__ frag_slot, const unsigned char* __restrict__ ref, int P,
                                     int L, int C, long long rows, float* __restrict__ h) {
  // one warp per token row, float4 per lane, grid-stride
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int f = (int)(row / L);
    const int slot = frag_slot[f];
    const float* re = ref_emb + (size_t)(ref[slot] ? 1 : 0) * C;
    const float* pp = pe + (size_t)(slot % P) * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 xe = *reinterpret_cast<const float4*>(x_emb + (size_t)f * C + c);
      const float4 r4 = *reinterpret_cast<const float4*>(re + c);
      const float4 se = *reinterpret_cast<const float4*>(shape_emb + (size_t)row * C + c);
      const float4 p4 = *reinterpret_cast<const float4*>(pp + c);
      float4 o;
      o.x = fadd(fadd(fadd(xe.x, r4.x), se.x), p4.x);
      o.y = fadd(fadd(fadd(xe.y, r4.y), se.y), p4.y);
      o.z = fadd(fadd(fadd(xe.z, r4.z), se.z), p4.z);
      o.w = fadd(fadd(fadd(xe.w, r4.w), se.w), p4.w);
      *reinterpret_cast<float4*>(h + (size_t)row * C + c) = o;
    }
  }
}

extern "C" int pfpp_combine_embed(const float* shape_emb, const float* x_emb, const float* ref_emb, const float* pe,
                                  const int* frag_slot, const unsigned char* ref, int F, int P, int L, int C, float* h,
                                  cudaStream_t stream) {
  PFPP_CHECK_ARG(shape_emb && x_emb && ref_emb && pe && frag_slot && ref && h);
  PFPP_CHECK_ARG((C % 4) == 0);
  if (F == 0) return PFPP_OK;
  const long long rows = (long long)F * L;
  int grid = pfpp_cdiv(rows, 8);
  if (grid > 148 * 8) grid = 148 * 8;
  combine_embed_kernel<<<grid, 256, 0, stream>>>(shape_emb, x_emb, ref_emb, pe, frag_slot, ref, P, L, C, rows, h);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// LayerNorm / AdaLayerNorm (attention.py:5-25): one warp per row, eps 1e-5, biased variance.
//   y = LN(x) * gamma + beta                         (affine)            or
//   y = LN(x) * (1 + mod[g, 0:C]) + mod[g, C:2C]     (AdaLN, g = row_group[row / rows_per_group])
// Optionally post-LN residual form y = LN(x + r).
// ---------------------------------------------------------------------------------------------
template <int C, typename OutT, bool SPLIT>
__global__ void __launch_bounds__(256)
    layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ mod, const int* __restrict__ row_group,
                     int rows_per_group, long long rows, OutT* __restrict__ y, float* __restrict__ sum_out) {
  const int lane = threadIdx.x & 31;
  constexpr int PER = C / 32;
  // grid-stride over rows: the grid is capped at one resident wave (a tail wave of a few blocks would cost a whole
  // extra wave of latency on a kernel that lasts ~10 us)
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    float4 t = *reinterpret_cast<const float4*>(x + row * C + (i / 4) * 128 + lane * 4);
    if (res) {
      float4 r = *reinterpret_cast<const float4*>(res + row * C + (i / 4) * 128 + lane * 4);
      t.x += r.x, t.y += r.y, t.z += r.z, t.w += r.w;
    }
    v[i] = t.x, v[i + 1] = t.y, v[i + 2] = t.z, v[i + 3] = t.w;
    s += t.x + t.y + t.z + t.w;
  }
  if (sum_out) {
#pragma unroll
    for (int i = 0; i < PER; i += 4)
      *reinterpret_cast<float4*>(sum_out + row * C + (i / 4) * 128 + lane * 4) =
          make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    float d = v[i] - mean;
    q += d * d;
  }
  float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
  const float* md = mod ? mod + (size_t)row_group[row / rows_per_group] * 2 * C : nullptr;
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    const int c = (i / 4) * 128 + lane * 4;
    float o[4] = {(v[i] - mean) * rstd, (v[i + 1] - mean) * rstd, (v[i + 2] - mean) * rstd, (v[i + 3] - mean) * rstd};
    if (md) {
      const float4 sc = *reinterpret_cast<const float4*>(md + c), sh = *reinterpret_cast<const float4*>(md + C + c);
      o[0] = o[0] * (1.0f + sc.x) + sh.x, o[1] = o[1] * (1.0f + sc.y) + sh.y;
      o[2] = o[2] * (1.0f + sc.z) + sh.z, o[3] = o[3] * (1.0f + sc.w) + sh.w;
    } else if (gamma) {
      const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
      o[0] = o[0] * ga.x + be.x, o[1] = o[1] * ga.y + be.y, o[2] = o[2] * ga.z + be.z, o[3] = o[3] * ga.w + be.w;
    }
    if constexpr (SPLIT) {  // bf16 hi/lo rows of 2C elements: hi at c, lo at C + c
      __nv_bfloat16* ys = reinterpret_cast<__nv_bfloat16*>(y) + row * (2 * C) + c;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      __nv_bfloat162 l0 = __floats2bfloat162_rn(o[0] - f0.x, o[1] - f0.y), l1 = __floats2bfloat162_rn(o[2] - f1.x, o[3] - f1.y);
      *reinterpret_cast<uint2*>(ys) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      *reinterpret_cast<uint2*>(ys + C) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
      continue;
    }
    OutT* yo = y + row * C + c;
    if (sizeof(OutT) == 2) {  // bf16: one 8-byte store per lane
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
      *reinterpret_cast<uint2*>(yo) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
    } else {
      *reinterpret_cast<float4*>(yo) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  }  // row loop
}

extern "C" int pfpp_layernorm(const float* x, const float* residual, const float* gamma, const float* beta,
                              const float* mod, const int* row_group, int rows_per_group, long long rows, int C,
                              int out_bf16, void* y, float* sum_out, cudaStream_t stream) {
  PFPP_CHECK_ARG(x && y && (C == 512 || C == 256));
  PFPP_CHECK_ARG(!mod || (row_group && rows_per_group > 0));
  if (rows == 0) return PFPP_OK;
  int grid = pfpp_cdiv(rows, 8);
  if (grid > 148 * 8) grid = 148 * 8;  // one resident wave (256 threads, <= 48 registers: 8 blocks per SM)
#define PFPP_LN_CASE(CC, T, SP)                                                                                   \
  layernorm_kernel<CC, T, SP><<<grid, 256, 0, stream>>>(x, residual, gamma, beta, mod, row_group, rows_per_group, \
                                                        rows, (T*)y, sum_out)
  if (C == 512) {
    if (out_bf16 == 2) PFPP_LN_CASE(512, __nv_bfloat16, true);
    else if (out_bf16) PFPP_LN_CASE(512, __nv_bfloat16, false);
    else PFPP_LN_CASE(512, float, false);
  } else {
    if (out_bf16 == 2) PFPP_LN_CASE(256, __nv_bfloat16, true);
    else if (out_bf16) PFPP_LN_CASE(256, __nv_bfloat16, false);
    else PFPP_LN_CASE(256, float, false);
  }
#undef PFPP_LN_CASE
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// variable-length multi-head attention (fp32 SIMT, flash-style online softmax).
//   tokens are packed [M, ld] with q/k/v at column offsets; segment s owns rows
//   [seg_start[s], seg_start[s]+seg_len[s]) and attends fully within itself.  This one kernel is
//   the denoiser's block-diagonal local attention (segments = fragments, 25 tokens), its global
//   attention over the valid fragments of an object (<=500 tokens, padded fragments are simply
//   not in the packed batch, which is what the reference's key mask achieves), and the verifier's
//   key-padded attention over valid edges.
//   Four lanes per query row (q and the output accumulator in registers, a quarter each), K/V tiles of 32 keys
//   staged in shared memory.
// ---------------------------------------------------------------------------------------------
#define ATT_KT 32
#define ATT_QPB 32  // queries per block (4 lanes each)
// Four lanes share one query row: lane `sub` owns the 16-byte chunks c = 4 i + sub of the head dimension (so the four
// lanes of a query read 64 contiguous bytes of a K / V row: one conflict-free wavefront), i.e. D/4 of the q values
// and of the output accumulator.  A score is the sum of the four partial dot products (two xor-shuffles).  Compared
// with one thread per query this quarters the registers per thread (q + acc = D/2 instead of 2 D), which is what
// bounds the occupancy -- and with it the latency hiding -- of this kernel.
template <int D, typename InT, typename OutT, bool SPLIT>
__global__ void __launch_bounds__(4 * ATT_QPB)
    attn_varlen_kernel(const InT* __restrict__ qkv, int ld, int q_off, int k_off, int v_off,
                       const int* __restrict__ seg_start, const int* __restrict__ seg_len, float scale,
                       OutT* __restrict__ out, int ldo) {
  constexpr int NC = D / 16;  // 16-byte chunks per lane
  __shared__ __align__(16) float Ks[ATT_KT][D];
  __shared__ __align__(16) float Vs[ATT_KT][D];
  const int s = blockIdx.z, h = blockIdx.y;
  const int len = seg_len[s], st = seg_start[s];
  const int q0 = blockIdx.x * ATT_QPB;
  if (q0 >= len) return;
  const int sub = threadIdx.x & 3;
  const int qi = q0 + (threadIdx.x >> 2);
  const bool active = qi < len;
  float q[NC * 4], acc[NC * 4];
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = (4 * i + sub) * 4 + e;
      q[4 * i + e] = active ? (float)qkv[(size_t)(st + qi) * ld + q_off + h * D + d] * scale : 0.f;
      acc[4 * i + e] = 0.f;
    }
  float m = -INFINITY, l = 0.f;
  for (int kt = 0; kt < len; kt += ATT_KT) {
    const int nk = min(ATT_KT, len - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) {
      const int r = i / D, d = i - r * D;
      const size_t base = (size_t)(st + kt + r) * ld + h * D + d;
      Ks[r][d] = (float)qkv[base + k_off];
      Vs[r][d] = (float)qkv[base + v_off];
    }
    __syncthreads();
    float sc[ATT_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      float a = -INFINITY;
      if (j < nk) {  // block-uniform
        a = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const float4 kv = *reinterpret_cast<const float4*>(&Ks[j][(4 * i + sub) * 4]);
          a += q[4 * i] * kv.x + q[4 * i + 1] * kv.y + q[4 * i + 2] * kv.z + q[4 * i + 3] * kv.w;
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        tmax = fmaxf(tmax, a);
      }
      sc[j] = a;
    }
    const float mn = fmaxf(m, tmax);
    const float corr = expf(m - mn);  // exp(-inf) = 0 on the first tile
    l *= corr;
#pragma unroll
    for (int d = 0; d < NC * 4; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      if (j < nk) {
        const float p = expf(sc[j] - mn);
        l += p;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][(4 * i + sub) * 4]);
          acc[4 * i] += p * vv.x, acc[4 * i + 1] += p * vv.y, acc[4 * i + 2] += p * vv.z, acc[4 * i + 3] += p * vv.w;
        }
      }
    }
    m = mn;
  }
  if (active) {
    const float inv = 1.0f / l;
    OutT* o = out + (size_t)(st + qi) * ldo + h * D;
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) put_out<OutT, SPLIT>(o + (4 * i + sub) * 4 + e, ldo / 2, acc[4 * i + e] * inv);
  }
}

extern "C" int pfpp_attention_varlen(const void* qkv, int ld, int q_off, int k_off, int v_off, const int* seg_start,
                                     const int* seg_len, int n_segments, int max_len, int heads, int head_dim,
                                     int io_bf16, void* out, int ldo, cudaStream_t stream) {
  PFPP_CHECK_ARG(qkv && seg_start && seg_len && out && heads > 0 && (head_dim == 64 || head_dim == 32));
  if (n_segments == 0 || max_len == 0) return PFPP_OK;
  const int threads = 4 * ATT_QPB;
  dim3 grid(pfpp_cdiv(max_len, ATT_QPB), heads, n_segments);
  float scale = 1.0f / sqrtf((float)head_dim);
#define PFPP_ATT_CASE(DD, TI, TO, SP)                                                                        \
  attn_varlen_kernel<DD, TI, TO, SP><<<grid, threads, 0, stream>>>((const TI*)qkv, ld, q_off, k_off, v_off, \
                                                                   seg_start, seg_len, scale, (TO*)out, ldo)
  // io_bf16: 0 = fp32 in / fp32 out, 1 = bf16 / bf16, 2 = fp32 in / bf16 hi-lo split out (rows of ldo = 2C elements)
  if (head_dim == 64) {
    if (io_bf16 == 2) PFPP_ATT_CASE(64, float, __nv_bfloat16, true);
    else if (io_bf16) PFPP_ATT_CASE(64, __nv_bfloat16, __nv_bfloat16, false);
    else PFPP_ATT_CASE(64, float, float, false);
  } else {
    if (io_bf16 == 2) PFPP_ATT_CASE(32, float, __nv_bfloat16, true);
    else if (io_bf16) PFPP_ATT_CASE(32, __nv_bfloat16, __nv_bfloat16, false);
    else PFPP_ATT_CASE(32, float, float, false);
  }
#undef PFPP_ATT_CASE
  PFPP_RETURN_LAST();
}

// mean over the L tokens of each fragment (denoiser_transformer.py:141-142): [F*L, C] -> [F, C]
template <typename OutT, bool SPLIT>
__global__ void mean_pool_kernel(const float* __restrict__ h, int L, int C, OutT* __restrict__ out) {
  const int f = blockIdx.x;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < L; ++l) {  // same summation order per channel as a scalar loop over l
      const float4 v = *reinterpret_cast<const float4*>(h + ((size_t)f * L + l) * C + c);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    OutT* o = out + (size_t)f * (SPLIT ? 2 * C : C) + c;
    put_out<OutT, SPLIT>(o, C, s.x / (float)L);
    put_out<OutT, SPLIT>(o + 1, C, s.y / (float)L);
    put_out<OutT, SPLIT>(o + 2, C, s.z / (float)L);
    put_out<OutT, SPLIT>(o + 3, C, s.w / (float)L);
  }
}
extern "C" int pfpp_mean_pool(const float* h, int F, int L, int C, int out_bf16, void* out, cudaStream_t stream) {
  PFPP_CHECK_ARG(h && out);
  if (F == 0) return PFPP_OK;
  if (out_bf16 == 2) mean_pool_kernel<__nv_bfloat16, true><<<F, 128, 0, stream>>>(h, L, C, (__nv_bfloat16*)out);  // [F, 2C] split
  else if (out_bf16) mean_pool_kernel<__nv_bfloat16, false><<<F, 128, 0, stream>>>(h, L, C, (__nv_bfloat16*)out);
  else mean_pool_kernel<float, false><<<F, 128, 0, stream>>>(h, L, C, (float*)out);
  PFPP_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// DDPM posterior step + reference-part clamp (diffusers DDPMScheduler.step, SURVEY App. B.1;
// auto_aggl.py:149-150).  eps is [F,7] in packed fragment order, x/noise/ref_pose are [slots,7].
// coef = {sqrt(1-abar_t), sqrt(abar_t), c_x0, c_x, sigma}; add_noise = (t > 0).
// ---------------------------------------------------------------------------------------------
__global__ void ddpm_step_kernel(const float* __restrict__ eps, int ld_eps, const int* __restrict__ frag_slot,
                                 const float* __restrict__ coef, const int* __restrict__ frag_coef, int add_noise,
                                 const float* __restrict__ noise, long long noise_step_stride,
                                 const unsigned char* __restrict__ ref, const float* __restrict__ ref_pose, int F,
                                 float* __restrict__ x, float* __restrict__ x_hist, long long hist_step_stride) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * 7) return;
  int f = i / 7, c = i - f * 7;
  int slot = frag_slot[f];
  const int step = frag_coef ? frag_coef[f] : 0;
  const float* cf = coef + 5 * step;
  float xv = x[(size_t)slot * 7 + c];
  float e = eps[(size_t)f * ld_eps + c];
  float x0 = fdiv(fsub(xv, fmul(cf[0], e)), cf[1]);
  float prev = fadd(fmul(cf[2], x0), fmul(cf[3], xv));
  if (add_noise) prev = fadd(prev, fmul(cf[4], noise[(size_t)step * noise_step_stride + (size_t)slot * 7 + c]));
  if (ref[slot]) prev = ref_pose[(size_t)slot * 7 + c];
  x[(size_t)slot * 7 + c] = prev;
  if (x_hist) x_hist[(size_t)step * hist_step_stride + (size_t)slot * 7 + c] = prev;
}

extern "C" int pfpp_ddpm_step(const float* eps, int ld_eps, const int* frag_slot, const float* coef,
                              const int* frag_coef, int add_noise, const float* noise, long long noise_step_stride,
                              const unsigned char* ref, const float* ref_pose, int F, float* x, float* x_hist,
                              long long hist_step_stride, cudaStream_t stream) {
  PFPP_CHECK_ARG(eps && frag_slot && coef && ref && ref_pose && x && (!add_noise || noise));
  if (F == 0) return PFPP_OK;
  ddpm_step_kernel<<<pfpp_cdiv(F * 7, 128), 128, 0, stream>>>(eps, ld_eps, frag_slot, coef, frag_coef, add_noise, noise,
                                                             noise_step_stride, ref, ref_pose, F, x, x_hist,
                                                             hist_step_stride);
  PFPP_RETURN_LAST();
}

// Device-side DDPM step counter: lets one captured CUDA graph serve all T steps of an outer iteration.
// pfpp_step_broadcast: out[i] = *step (per-fragment AdaLN row / coefficient row / noise row index);
// pfpp_step_advance:   *step += 1.
__global__ void step_broadcast_kernel(const int* __restrict__ step, int* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = *step;
}
__global__ void step_advance_kernel(int* step) { *step += 1; }

extern "C" int pfpp_step_broadcast(const int* step, int* out, int n, cudaStream_t stream) {
  PFPP_CHECK_ARG(step && out && n >= 0);
  if (n == 0) return PFPP_OK;
  step_broadcast_kernel<<<pfpp_cdiv(n, 256), 256, 0, stream>>>(step, out, n);
  PFPP_RETURN_LAST();
}
extern "C" int pfpp_step_advance(int* step, cudaStream_t stream) {
  PFPP_CHECK_ARG(step);
  step_advance_kernel<<<1, 1, 0, stream>>>(step);
  PFPP_RETURN_LAST();
}
