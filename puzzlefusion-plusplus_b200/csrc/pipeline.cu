// Coarse entry points of the drop-in boundary: one C call per stage of
// AutoAgglomerative.test_step -- the fragment encoder, the denoiser forward, one whole DDPM step, the verifier --
// sequencing the kernels of this library over a caller-provided workspace.  Weights arrive as flat structs of
// device pointers (include/pfpp.h) filled by the host packer from the reference checkpoints' state_dicts.
// No allocation, no host synchronisation, no global mutable state: every call is a fixed launch sequence on
// `stream` (capturable into a CUDA graph).
#include "common.cuh"
#include "../../include/pfpp.h"

namespace {

#define PF(call)            \
  do {                      \
    int rc__ = (call);      \
    if (rc__ != 0) return rc__; \
  } while (0)

// bump carving of the workspace; with base == nullptr it only measures
struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base((char*)b) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
  void* take_bytes(size_t n) { return take<char>(n); }
};

inline int pad_to(int k, int m) { return (k + m - 1) / m * m; }
inline size_t act_size(int mode) { return mode == 0 ? 4 : 2; }

// out = epi(a W^T + b) (+ residual) on the contraction engine of `mode` (engine.Engine.gemm).
// out_act: the output is an activation in the mode's operand format (fp32 / bf16 / split); otherwise fp32.
int gemm(int mode, const void* a, int lda, const PfppLinear& lin, void* out, int ldc, int M, int epi, const float* residual,
         int ldr, bool out_act, cudaStream_t s) {
  if (mode == 2)
    return pfpp_gemm_bf16x3(a, 2 * lda, lin.w, 2 * lin.k, lin.bias, residual, ldr, out, out_act ? 2 * ldc : ldc, out_act ? 1 : 0,
                            M, lin.n, lin.k, epi, s);
  if (mode == 1)
    return pfpp_gemm_bf16(a, lda, lin.w, lin.k, lin.bias, residual, ldr, out, ldc, out_act ? 1 : 0, M, lin.n, lin.k, epi, s);
  return pfpp_gemm_f32((const float*)a, lda, (const float*)lin.w, lin.k, lin.bias, residual, ldr, (float*)out, ldc, M, lin.n,
                       lin.k, epi, s);
}
int gemm_f32(const float* a, int lda, const PfppLinear& lin, float* out, int ldc, int M, int epi, cudaStream_t s) {
  return pfpp_gemm_f32(a, lda, (const float*)lin.w, lin.k, lin.bias, nullptr, 0, out, ldc, M, lin.n, lin.k, epi, s);
}

// ------------------------------------------------------------------------------------------------ encoder
struct EncBufs {
  float *z_e, *zq_tmp, *xyz_tmp;
  float* rot;
  int* gidx;
  void *X, *B1, *B2, *B3;
  int* idx[3];
  float* cxyz[3];
  void* feats[3];
  int Fc;
  bool fused;
};

bool enc_fused(const PfppEncoderWeights* w) {
  if (w->mode != 1 || !w->fused_sa) return false;
  static const int lv[3][5] = {{32, 0, 64, 64, 128}, {64, 128, 128, 128, 256}, {64, 256, 256, 256, 512}};
  int d = 0;
  for (int i = 0; i < 3; ++i) {
    if (w->nsample[i] != lv[i][0] || d != lv[i][1] || w->sa[i][0].n != lv[i][2] || w->sa[i][1].n != lv[i][3] ||
        w->sa[i][2].n != lv[i][4])
      return false;
    d = w->sa[i][2].n;
  }
  return true;
}

EncBufs plan_encoder(Bump& b, const PfppEncoderWeights* w, int F, int N) {
  EncBufs e{};
  const int L = w->latent_points, wm = w->mode == 2 ? 2 : 1;
  const size_t es = act_size(w->mode);
  e.fused = enc_fused(w);
  e.Fc = e.fused ? F : (w->chunk_frags < F ? w->chunk_frags : F);
  if (e.Fc < 1) e.Fc = 1;
  e.z_e = b.take<float>((size_t)F * L * w->latent_dim);
  e.zq_tmp = b.take<float>((size_t)F * L * w->latent_dim);
  e.xyz_tmp = b.take<float>((size_t)F * L * 3);
  e.rot = b.take<float>((size_t)e.Fc * N * 3);
  size_t max_rows = 0, mx = 0, m1 = 0, m2 = 0, m3 = 0;
  const int kmult = w->mode == 0 ? 4 : 8;
  int d = 0;
  for (int i = 0; i < 3; ++i) {
    const size_t rows = (size_t)e.Fc * w->npoint[i] * w->nsample[i];
    max_rows = rows > max_rows ? rows : max_rows;
    const size_t kin = pad_to(3 + d, kmult);
    mx = rows * kin > mx ? rows * kin : mx;
    m1 = rows * w->sa[i][0].n > m1 ? rows * w->sa[i][0].n : m1;
    m2 = rows * w->sa[i][1].n > m2 ? rows * w->sa[i][1].n : m2;
    m3 = rows * w->sa[i][2].n > m3 ? rows * w->sa[i][2].n : m3;
    d = w->sa[i][2].n;
  }
  e.gidx = b.take<int>(max_rows);
  if (!e.fused) {
    e.X = b.take_bytes(wm * mx * es);
    e.B1 = b.take_bytes(wm * m1 * es);
    e.B2 = b.take_bytes(wm * m2 * es);
    e.B3 = b.take_bytes(m3 * (w->mode == 2 ? 4 : es));
  }
  for (int i = 0; i < 3; ++i) {
    e.idx[i] = b.take<int>((size_t)e.Fc * w->npoint[i]);
    e.cxyz[i] = b.take<float>((size_t)e.Fc * w->npoint[i] * 3);
    e.feats[i] = b.take_bytes((size_t)e.Fc * w->npoint[i] * wm * w->sa[i][2].n * es);
  }
  return e;
}

// dst row pos[r] = src row r  (rows of `elems` floats): results of a compacted encoder pass into packed positions
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ pos, int elems,
                                    float* __restrict__ dst) {
  const float* s = src + (size_t)blockIdx.x * elems;
  float* d = dst + (size_t)pos[blockIdx.x] * elems;
  for (int i = threadIdx.x; i < elems; i += blockDim.x) d[i] = s[i];
}

// out_pos != nullptr: fragment f's results go to rows out_pos[f] of z_q / xyz_out (encoded into scratch, then
// scattered); nullptr: rows f.
int run_encoder(const PfppEncoderWeights* w, const float* part_pcs, const int* frag_slot, const float* x, int F, int N,
                float* z_q, float* xyz_out, int* codes, const int* out_pos, Bump& b, cudaStream_t s) {
  EncBufs e = plan_encoder(b, w, F, N);
  float *z_dst = z_q, *xyz_dst = xyz_out;
  if (out_pos) {
    z_q = e.zq_tmp;
    xyz_out = e.xyz_tmp;
  }
  const int L = w->latent_points, mode = w->mode, wm = mode == 2 ? 2 : 1, kmult = mode == 0 ? 4 : 8;
  for (int c0 = 0; c0 < F; c0 += e.Fc) {
    const int K = e.Fc < F - c0 ? e.Fc : F - c0;
    const float* src_xyz = e.rot;
    int src_n = N, src_d = 0;
    const void* src_feat = nullptr;
    for (int li = 0; li < 3; ++li) {
      const int S = w->npoint[li], ns = w->nsample[li];
      // the last level's centroids are the encoder's xyz output: written in place
      float* cx = li == 2 ? xyz_out + (size_t)c0 * L * 3 : e.cxyz[li];
      if (li == 0)
        PF(pfpp_rotate_fps(part_pcs, frag_slot + c0, K, N, S, x + 3, 7, e.rot, e.idx[0], cx, s));
      else
        PF(pfpp_fps(src_xyz, K, src_n, S, nullptr, e.idx[li], cx, s));
      PF(pfpp_ball_query(src_xyz, cx, K, src_n, S, w->radius_sq[li], ns, e.gidx, s));
      const PfppLinear &l0 = w->sa[li][0], &l1 = w->sa[li][1], &l2 = w->sa[li][2];
      if (e.fused) {
        PF(pfpp_sa_fused(li + 1, src_xyz, cx, src_feat, e.gidx, K, src_n, S, w->sa_w0_feat[li], w->sa_w0_xyz[li], l0.bias, l1.w,
                         l1.bias, l2.w, l2.bias, e.feats[li], s));
      } else {
        const int rows = K * S * ns, ld = pad_to(3 + src_d, kmult);
        PF(pfpp_group_gather(src_xyz, cx, src_feat, e.gidx, K, src_n, S, ns, src_d, ld, mode, e.X, s));
        PF(gemm(mode, e.X, ld, l0, e.B1, l0.n, rows, PFPP_EPI_RELU, nullptr, 0, true, s));
        PF(gemm(mode, e.B1, l0.n, l1, e.B2, l1.n, rows, PFPP_EPI_RELU, nullptr, 0, true, s));
        PF(gemm(mode, e.B2, l1.n, l2, e.B3, l2.n, rows, PFPP_EPI_RELU, nullptr, 0, mode != 2, s));
        PF(pfpp_group_max(e.B3, (long long)K * S, ns, l2.n, l2.n, mode, e.feats[li], wm * l2.n, s));
      }
      src_xyz = cx, src_n = S, src_feat = e.feats[li], src_d = l2.n;
    }
    // conv6 (pn2.py:65): fp32 output for the code search
    PF(gemm(mode, e.feats[2], w->conv6.k, w->conv6, e.z_e + (size_t)c0 * L * w->latent_dim, w->latent_dim, K * L, PFPP_EPI_NONE,
            nullptr, 0, false, s));
  }
  PF(pfpp_vq(e.z_e, 0, (long long)F * L * (w->latent_dim / 16), w->codebook, w->n_codes, z_q, codes, s));
  if (out_pos) {
    scatter_rows_kernel<<<F, 128, 0, s>>>(z_q, out_pos, L * w->latent_dim, z_dst);
    scatter_rows_kernel<<<F, 96, 0, s>>>(xyz_out, out_pos, L * 3, xyz_dst);
  }
  PFPP_RETURN_LAST();
}

// ------------------------------------------------------------------------------------------------ denoiser
struct DenBufs {
  void *feat_tok, *feat_par, *ln, *qkv, *ao, *ff;
  float *shape_emb, *x_emb, *h, *pooled, *h0, *ht, *hr;
  int *loc_start, *loc_len;
  int ld_tok, ld_par, n_loc;
};

__global__ void local_segments_kernel(int n, int per_tokens, int total_tokens, int* start, int* len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int st = i * per_tokens;
  start[i] = st;
  len[i] = total_tokens - st < per_tokens ? total_tokens - st : per_tokens;
}

DenBufs plan_denoiser(Bump& b, const PfppDenoiserWeights* w, int F) {
  DenBufs d{};
  const int L = w->L, C = w->C, mode = w->mode, wm = mode == 2 ? 2 : 1, kmult = mode == 0 ? 4 : 8;
  const size_t es = act_size(mode), M = (size_t)F * L;
  d.ld_tok = pad_to(w->latent_dim + 84, kmult);
  d.ld_par = pad_to(147, kmult);
  d.feat_tok = b.take_bytes(M * wm * d.ld_tok * es);
  d.feat_par = b.take_bytes((size_t)F * wm * d.ld_par * es);
  d.shape_emb = b.take<float>(M * C);
  d.x_emb = b.take<float>((size_t)F * C);
  d.h = b.take<float>(M * C);
  d.ln = b.take_bytes(M * wm * C * es);
  d.qkv = b.take_bytes(M * 3 * C * (mode == 2 ? 4 : es));
  d.ao = b.take_bytes(M * wm * C * es);
  d.ff = b.take_bytes(M * wm * 4 * C * es);
  d.pooled = b.take<float>((size_t)F * C);
  d.h0 = b.take<float>((size_t)F * 2 * C);
  d.ht = b.take<float>((size_t)F * (C / 2));
  d.hr = b.take<float>((size_t)F * (C / 2));
  const int per = 5 * (w->local_tiles > 0 ? w->local_tiles : 1);
  d.n_loc = (F + per - 1) / per;
  d.loc_start = b.take<int>(d.n_loc);
  d.loc_len = b.take<int>(d.n_loc);
  return d;
}

int run_denoiser(const PfppDenoiserWeights* w, const float* x, const float* scale, const unsigned char* ref,
                 const int* frag_slot, const int* frag_step, const float* latent, const float* xyz, const int* frag_seg_start,
                 const int* frag_seg_len, const int* obj_seg_start, const int* obj_seg_len, int F, int n_obj, int max_global,
                 float* eps, Bump& b, cudaStream_t s) {
  DenBufs d = plan_denoiser(b, w, F);
  const int L = w->L, C = w->C, H = w->heads, mode = w->mode, wm = mode == 2 ? 2 : 1, D = C / H;
  const int M = F * L;
  PF(pfpp_embed_features(x, scale, frag_slot, latent, xyz, F, L, w->latent_dim, mode, d.feat_tok, wm * d.ld_tok, d.feat_par,
                         wm * d.ld_par, s));
  PF(gemm(mode, d.feat_tok, d.ld_tok, w->shape_embedding, d.shape_emb, C, M, PFPP_EPI_NONE, nullptr, 0, false, s));
  PF(gemm(mode, d.feat_par, d.ld_par, w->param_fc, d.x_emb, C, F, PFPP_EPI_NONE, nullptr, 0, false, s));
  PF(pfpp_combine_embed(d.shape_emb, d.x_emb, w->ref_emb, w->pe, frag_slot, ref, F, w->P, L, C, d.h, s));
  const bool mma_local = mode == 1 && D == 64 && w->tc_attention && w->local_tiles == 0 && L <= 32;
  const bool tc_local = !mma_local && mode == 1 && D == 64 && w->tc_attention && 5 * L <= 128;
  const bool tc_global = mode == 1 && D == 64 && w->tc_attention;
  if (tc_local) {
    local_segments_kernel<<<pfpp_cdiv(d.n_loc, 128), 128, 0, s>>>(d.n_loc, 5 * w->local_tiles * L, M, d.loc_start, d.loc_len);
  }
  // bf16 mode: the residual projections (out-proj, FF2) also emit the NEXT (Ada)LayerNorm's output (pfpp_gemm_res_ln):
  // 17 of the 18 LayerNorm passes of a step and one read of the residual stream per sub-layer disappear
  const bool fuse = mode == 1 && w->fused_ln && C == 512;
  for (int li = 0; li < w->n_layers; ++li) {
    const PfppDenoiserLayer& lw = w->layers[li];
    for (int which = 0; which < 2; ++which) {
      const float* mod = w->mod + (size_t)(li * 2 + which) * w->T * 2 * C;
      if (!fuse || (li == 0 && which == 0))
        PF(pfpp_layernorm(d.h, nullptr, nullptr, nullptr, mod, frag_step, L, M, C, mode, d.ln, nullptr, s));
      PF(gemm(mode, d.ln, C, lw.qkv[which], d.qkv, 3 * C, M, PFPP_EPI_NONE, nullptr, 0, mode != 2, s));
      if (which == 1 && tc_global) {
        PF(pfpp_attention_tc(d.qkv, M, 3 * C, C, obj_seg_start, obj_seg_len, n_obj, max_global, H, 0, d.ao, C, s));
      } else if (which == 0 && mma_local) {
        // block-diagonal local attention: one warp per (fragment, head)
        PF(pfpp_attention_local(d.qkv, M, 3 * C, C, H, L, d.ao, C, s));
      } else if (which == 0 && tc_local) {
        // block-diagonal local attention: 5 fragments (125 tokens) per 128-row tensor-core tile
        PF(pfpp_attention_tc(d.qkv, M, 3 * C, C, d.loc_start, d.loc_len, d.n_loc, 5 * L * w->local_tiles, H, L, d.ao, C, s));
      } else if (which == 0) {
        PF(pfpp_attention_varlen(d.qkv, 3 * C, 0, C, 2 * C, frag_seg_start, frag_seg_len, F, L, H, D, mode, d.ao, wm * C, s));
      } else {
        PF(pfpp_attention_varlen(d.qkv, 3 * C, 0, C, 2 * C, obj_seg_start, obj_seg_len, n_obj, max_global, H, D, mode, d.ao,
                                 wm * C, s));
      }
      if (fuse && which == 0) {  // -> AdaLN of the global attention
        PF(pfpp_gemm_res_ln(d.ao, C, lw.out[0].w, lw.out[0].k, lw.out[0].bias, d.h, M, lw.out[0].k, mod + (size_t)w->T * 2 * C,
                            frag_step, L, nullptr, nullptr, d.ln, s));
      } else if (fuse) {         // -> norm3
        PF(pfpp_gemm_res_ln(d.ao, C, lw.out[1].w, lw.out[1].k, lw.out[1].bias, d.h, M, lw.out[1].k, nullptr, nullptr, 0,
                            lw.norm3_w, lw.norm3_b, d.ln, s));
      } else {
        PF(gemm(mode, d.ao, C, lw.out[which], d.h, C, M, PFPP_EPI_NONE, d.h, C, false, s));
      }
    }
    if (!fuse) PF(pfpp_layernorm(d.h, nullptr, lw.norm3_w, lw.norm3_b, nullptr, nullptr, 0, M, C, mode, d.ln, nullptr, s));
    PF(gemm(mode, d.ln, C, lw.ff1, d.ff, 4 * C, M, PFPP_EPI_GEGLU, nullptr, 0, true, s));
    if (fuse && li + 1 < w->n_layers) {  // -> AdaLN of the next layer's local attention
      PF(pfpp_gemm_res_ln(d.ff, 4 * C, lw.ff2.w, lw.ff2.k, lw.ff2.bias, d.h, M, lw.ff2.k,
                          w->mod + (size_t)((li + 1) * 2) * w->T * 2 * C, frag_step, L, nullptr, nullptr, d.ln, s));
    } else {
      PF(gemm(mode, d.ff, 4 * C, lw.ff2, d.h, C, M, PFPP_EPI_NONE, d.h, C, false, s));
    }
  }
  if (mode == 0) {
    PF(pfpp_mean_pool(d.h, F, L, C, 0, d.pooled, s));
    PF(gemm_f32(d.pooled, C, w->head0, d.h0, 2 * C, F, PFPP_EPI_SILU, s));
    PF(gemm_f32(d.h0, 2 * C, w->head_t2, d.ht, C / 2, F, PFPP_EPI_SILU, s));
    PF(gemm_f32(d.h0 + C, 2 * C, w->head_r2, d.hr, C / 2, F, PFPP_EPI_SILU, s));
    PF(gemm_f32(d.ht, C / 2, w->head_t4, eps, 8, F, PFPP_EPI_NONE, s));
    PF(gemm_f32(d.hr, C / 2, w->head_r4, eps + 3, 8, F, PFPP_EPI_NONE, s));
    return PFPP_OK;
  }
  // tensor-core modes: the output heads (M = fragments: latency-bound as SIMT GEMMs) run as split-operand bf16x3
  // GEMMs, fp32-grade; the same fp32 buffers hold the split rows (2 bf16 per fp32 slot)
  __nv_bfloat16 *pooled = (__nv_bfloat16*)d.pooled, *h0 = (__nv_bfloat16*)d.h0, *ht = (__nv_bfloat16*)d.ht,
                *hr = (__nv_bfloat16*)d.hr;
  PF(pfpp_mean_pool(d.h, F, L, C, 2, pooled, s));
  PF(gemm(2, pooled, C, w->head0, h0, 2 * C, F, PFPP_EPI_SILU, nullptr, 0, true, s));
  // the trans / rot branches read the two halves of head0's output: hi at column offset 0 / C, lo 2C further
  PF(pfpp_gemm_bf16x3(h0, 4 * C, w->head_t2.w, 2 * w->head_t2.k, w->head_t2.bias, nullptr, 0, ht, C, 1, F, w->head_t2.n,
                      w->head_t2.k, PFPP_EPI_SILU, s));
  PF(pfpp_gemm_bf16x3(h0 + C, 4 * C, w->head_r2.w, 2 * w->head_r2.k, w->head_r2.bias, nullptr, 0, hr, C, 1, F, w->head_r2.n,
                      w->head_r2.k, PFPP_EPI_SILU, s));
  PF(gemm(2, ht, C / 2, w->head_t4, eps, 8, F, PFPP_EPI_NONE, nullptr, 0, false, s));
  PF(gemm(2, hr, C / 2, w->head_r4, eps + 3, 8, F, PFPP_EPI_NONE, nullptr, 0, false, s));
  return PFPP_OK;
}

// ------------------------------------------------------------------------------------------------ verifier
struct VerBufs {
  float *h, *h2, *qkv, *t1;
  void *hs, *ao, *ffb;
};

VerBufs plan_verifier(Bump& b, const PfppVerifierWeights* w, int n) {
  VerBufs v{};
  const size_t C = w->C, ffn = w->ffn;
  v.h = b.take<float>(n * C);
  v.h2 = b.take<float>(n * C);
  v.qkv = b.take<float>(n * 3 * C);
  v.t1 = b.take<float>(n * C);
  if (w->tc) {
    v.hs = b.take_bytes(n * 2 * C * 2);
    v.ao = b.take_bytes(n * 2 * C * 2);
    v.ffb = b.take_bytes(n * 2 * ffn * 2);
  } else {
    v.ao = b.take_bytes(n * C * 4);
    v.ffb = b.take_bytes(n * ffn * 4);
  }
  return v;
}

}  // namespace

// sizeof of the weight structs as this library was compiled (which: 0 PfppLinear, 1 encoder, 2 denoiser layer,
// 3 denoiser, 4 verifier layer, 5 verifier): lets a foreign-language binding verify its struct layout.
extern "C" size_t pfpp_struct_bytes(int which) {
  switch (which) {
    case 0: return sizeof(PfppLinear);
    case 1: return sizeof(PfppEncoderWeights);
    case 2: return sizeof(PfppDenoiserLayer);
    case 3: return sizeof(PfppDenoiserWeights);
    case 4: return sizeof(PfppVerifierLayer);
    case 5: return sizeof(PfppVerifierWeights);
    default: return 0;
  }
}

extern "C" size_t pfpp_encoder_workspace_bytes(const PfppEncoderWeights* w, int F, int N) {
  if (!w || F < 0 || N <= 0) return 0;
  Bump b(nullptr);
  plan_encoder(b, w, F, N);
  return b.off + 256;
}

extern "C" int pfpp_encoder_forward(const PfppEncoderWeights* w, const float* part_pcs, const int* frag_slot, const float* x,
                                    int F, int N, float* z_q, float* xyz, int* codes, const int* out_pos, void* workspace,
                                    size_t ws_bytes, cudaStream_t stream) {
  PFPP_CHECK_ARG(w && part_pcs && frag_slot && x && z_q && xyz && workspace && F >= 0 && N > 0);
  PFPP_CHECK_ARG(w->mode >= 0 && w->mode <= 2 && w->latent_dim % 16 == 0);
  if (F == 0) return PFPP_OK;
  if (ws_bytes < pfpp_encoder_workspace_bytes(w, F, N)) return PFPP_EWORKSPACE;
  Bump b(workspace);
  return run_encoder(w, part_pcs, frag_slot, x, F, N, z_q, xyz, codes, out_pos, b, stream);
}

extern "C" size_t pfpp_denoiser_workspace_bytes(const PfppDenoiserWeights* w, int F) {
  if (!w || F < 0) return 0;
  Bump b(nullptr);
  plan_denoiser(b, w, F);
  return b.off + 256;
}

extern "C" int pfpp_denoiser_forward(const PfppDenoiserWeights* w, const float* x, const float* scale,
                                     const unsigned char* ref, const int* frag_slot, const int* frag_step,
                                     const float* latent, const float* xyz, const int* frag_seg_start,
                                     const int* frag_seg_len, const int* obj_seg_start, const int* obj_seg_len, int F,
                                     int n_obj, int max_global, float* eps, void* workspace, size_t ws_bytes,
                                     cudaStream_t stream) {
  PFPP_CHECK_ARG(w && x && scale && ref && frag_slot && frag_step && latent && xyz && frag_seg_start && frag_seg_len &&
                 obj_seg_start && obj_seg_len && eps && workspace && F >= 0 && n_obj >= 0);
  PFPP_CHECK_ARG(w->mode >= 0 && w->mode <= 2 && w->n_layers >= 0 && w->n_layers <= PFPP_MAX_LAYERS);
  if (F == 0) return PFPP_OK;
  if (ws_bytes < pfpp_denoiser_workspace_bytes(w, F)) return PFPP_EWORKSPACE;
  Bump b(workspace);
  return run_denoiser(w, x, scale, ref, frag_slot, frag_step, latent, xyz, frag_seg_start, frag_seg_len, obj_seg_start,
                      obj_seg_len, F, n_obj, max_global, eps, b, stream);
}

extern "C" size_t pfpp_step_workspace_bytes(const PfppEncoderWeights* we, const PfppDenoiserWeights* wd, int F, int N) {
  if (!we || !wd || F < 0 || N <= 0) return 0;
  Bump b(nullptr);
  b.take<float>((size_t)F * 8);  // eps
  plan_encoder(b, we, F, N);
  plan_denoiser(b, wd, F);
  return b.off + 256;
}

extern "C" int pfpp_denoiser_step(const PfppEncoderWeights* we, const PfppDenoiserWeights* wd, const float* part_pcs, float* x,
                                  const float* scale, const unsigned char* ref, const float* ref_pose, const int* frag_slot,
                                  int* frag_step, int* step_counter, const float* noise, long long noise_step_stride,
                                  float* x_hist, long long hist_step_stride, const int* frag_seg_start,
                                  const int* frag_seg_len, const int* obj_seg_start, const int* obj_seg_len, int F, int n_obj,
                                  int max_global, int N, const int* enc_slot, const int* enc_pos, int F_enc, float* latent,
                                  float* xyz, float* eps_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  PFPP_CHECK_ARG(we && wd && part_pcs && x && scale && ref && ref_pose && frag_slot && frag_step && step_counter && noise &&
                 frag_seg_start && frag_seg_len && obj_seg_start && obj_seg_len && latent && xyz && workspace && F >= 0 &&
                 N > 0 && F_enc >= 0 && F_enc <= F && (F_enc == 0 || !enc_slot == !enc_pos));
  if (F == 0) return PFPP_OK;
  if (ws_bytes < pfpp_step_workspace_bytes(we, wd, F, N)) return PFPP_EWORKSPACE;
  Bump b(workspace);
  float* eps = b.take<float>((size_t)F * 8);
  if (eps_out) eps = eps_out;
  // frag_step[f] = *step_counter: selects the AdaLN row, the scheduler coefficients, the noise and history rows
  PF(pfpp_step_broadcast(step_counter, frag_step, F, stream));
  if (enc_slot) {
    // only the fragments whose pose changes from step to step are re-encoded; the rows of the others (reference
    // parts: their pose is clamped every step, auto_aggl.py:150) were written once when the iteration began
    if (F_enc > 0) PF(run_encoder(we, part_pcs, enc_slot, x, F_enc, N, latent, xyz, nullptr, enc_pos, b, stream));
  } else {
    PF(run_encoder(we, part_pcs, frag_slot, x, F, N, latent, xyz, nullptr, nullptr, b, stream));
  }
  Bump b2(workspace);
  b2.off = (size_t)F * 8 * sizeof(float);
  {  // the denoiser's buffers follow the (F-sized) encoder plan, whatever F_enc is
    Bump skip(nullptr);
    skip.off = b2.off;
    plan_encoder(skip, we, F, N);
    b2.off = skip.off;
  }
  PF(run_denoiser(wd, x, scale, ref, frag_slot, frag_step, latent, xyz, frag_seg_start, frag_seg_len, obj_seg_start,
                  obj_seg_len, F, n_obj, max_global, eps, b2, stream));
  PF(pfpp_ddpm_step(eps, 8, frag_slot, wd->coef, frag_step, 1, noise, noise_step_stride, ref, ref_pose, F, x, x_hist,
                    hist_step_stride, stream));
  return pfpp_step_advance(step_counter, stream);
}

extern "C" size_t pfpp_verifier_workspace_bytes(const PfppVerifierWeights* w, int n_tokens) {
  if (!w || n_tokens < 0) return 0;
  Bump b(nullptr);
  plan_verifier(b, w, n_tokens);
  return b.off + 256;
}

extern "C" int pfpp_verifier_forward(const PfppVerifierWeights* w, const float* feat, const int* tok_row, const int* tok_i,
                                     const int* tok_j, int n_tokens, const int* seg_start, const int* seg_len, int n_segments,
                                     int max_len, long long n_rows, float* logits, void* workspace, size_t ws_bytes,
                                     cudaStream_t stream) {
  PFPP_CHECK_ARG(w && feat && tok_row && tok_i && tok_j && seg_start && seg_len && logits && workspace && n_tokens >= 0 &&
                 n_rows >= 0 && w->n_layers >= 0 && w->n_layers <= PFPP_MAX_LAYERS);
  cudaError_t e = cudaMemsetAsync(logits, 0, (size_t)n_rows * 4, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_tokens == 0) return PFPP_OK;
  if (ws_bytes < pfpp_verifier_workspace_bytes(w, n_tokens)) return PFPP_EWORKSPACE;
  Bump b(workspace);
  VerBufs v = plan_verifier(b, w, n_tokens);
  const int C = w->C, H = w->heads, n = n_tokens, ffn = w->ffn, tc = w->tc, mode = tc ? 2 : 0;
  PF(pfpp_verifier_embed(feat, tok_row, tok_i, tok_j, n, w->emb_w, w->emb_b, w->pe, C, v.h, stream));
  for (int li = 0; li < w->n_layers; ++li) {
    const PfppVerifierLayer& lw = w->layers[li];
    const void* a = v.h;
    if (tc) {
      PF(pfpp_split_bf16(v.h, n, C, C, v.hs, 2 * C, stream));
      a = v.hs;
    }
    PF(gemm(mode, a, C, lw.qkv, v.qkv, 3 * C, n, PFPP_EPI_NONE, nullptr, 0, false, stream));
    PF(pfpp_attention_varlen(v.qkv, 3 * C, 0, C, 2 * C, seg_start, seg_len, n_segments, max_len, H, C / H, tc ? 2 : 0, v.ao,
                             tc ? 2 * C : C, stream));
    PF(gemm(mode, v.ao, C, lw.out, v.t1, C, n, PFPP_EPI_NONE, nullptr, 0, false, stream));
    // x = LN1(x + attn)      (post-LN TransformerEncoderLayer)
    PF(pfpp_layernorm(v.h, v.t1, lw.n1w, lw.n1b, nullptr, nullptr, 0, n, C, 0, v.h2, nullptr, stream));
    a = v.h2;
    if (tc) {
      PF(pfpp_split_bf16(v.h2, n, C, C, v.hs, 2 * C, stream));
      a = v.hs;
    }
    PF(gemm(mode, a, C, lw.l1, v.ffb, ffn, n, PFPP_EPI_GELU, nullptr, 0, true, stream));
    PF(gemm(mode, v.ffb, ffn, lw.l2, v.t1, C, n, PFPP_EPI_NONE, nullptr, 0, false, stream));
    // x = LN2(x + ff)
    PF(pfpp_layernorm(v.h2, v.t1, lw.n2w, lw.n2b, nullptr, nullptr, 0, n, C, 0, v.h, nullptr, stream));
  }
  return pfpp_verifier_head(v.h, tok_row, n, w->out_w, w->out_b, C, logits, stream);
}
