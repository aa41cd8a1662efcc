/* pfpp.h -- C ABI of libpfpp_sm100.so: the B200-native kernels of PuzzleFusion++'s
 * denoise-and-verify hot path.
 *
 * This is the drop-in boundary.  The reference is a Python program whose
 * heavy lifting happens in un-vendored native extensions (torch_cluster, pytorch3d, chamferdist,
 * cuBLAS/cuDNN through torch); this library replaces those call sites.  Every entry point below
 * cites the reference call site(s) it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers and sizes, no torch types, no allocation inside, caller owns
 *     every buffer; `void*` operands are float32 or bfloat16 as selected by the *_bf16 flag;
 *   - every call is asynchronous on the given cudaStream_t;
 *   - return 0 = ok, negative = argument error (PFPP_E*), positive = cudaError_t;
 *   - thread-compatible (no global mutable state); calls that share a workspace must be ordered
 *     by the caller (stream order is enough);
 *   - indices are int32, row-major dense layouts, leading dimensions in ELEMENTS.
 *
 * Packed-batch vocabulary: a batch of B fractured objects, each with P fragment SLOTS (padded).
 * The F valid fragments of the whole batch are packed contiguously (object-major, slot-minor);
 * frag_slot[f] = b*P + p maps a packed fragment to its slot.  Token row of fragment f, latent
 * point l is f*L + l (L = 25).
 */
#ifndef PFPP_H_
#define PFPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define PFPP_VERSION 100

/* GEMM epilogues */
#define PFPP_EPI_NONE 0
#define PFPP_EPI_RELU 1  /* 1x1 conv + folded BatchNorm + ReLU (utils/pn2_utils.py:209-212) */
#define PFPP_EPI_GELU 2  /* verifier linear1 (verifier_transformer.py:17-24, torch gelu) */
#define PFPP_EPI_SILU 3  /* denoiser output heads (denoiser_transformer.py:87-103) */
#define PFPP_EPI_GEGLU 4 /* diffusers GEGLU: out[:, j] = (v_j + b) * gelu(g_j + b), weight rows interleaved (v0,g0,v1,g1,..) */

int pfpp_version(void);
/* 1 when the library was built with the tcgen05/TMA (bf16) GEMM engine. */
int pfpp_has_tensor_core_path(void);

/* ---- geometry (encoder) ---------------------------------------------------------------- */

/* AutoAgglomerative._apply_rots (puzzlefusion_plusplus/auto_aggl.py:70-78, pytorch3d
 * quaternion_apply) fused with the first farthest-point-sampling level of sample_and_group
 * (utils/pn2_utils.py:131-139, torch_cluster.fps random_start=False).
 * part_pcs [slots,N,3]; quat points at x[0][3] with row stride quat_stride (7);
 * out: rot_out [K,N,3] rotated clouds, out_idx [K,S], out_xyz [K,S,3] (may be NULL). */
int pfpp_rotate_fps(const float* part_pcs, const int* frag_slot, int K, int N, int S, const float* quat,
                    int quat_stride, float* rot_out, int* out_idx, float* out_xyz, cudaStream_t stream);

/* torch_cluster.fps on K equal-size clouds (utils/pn2_utils.py:134; levels 2 and 3).
 * start [K] = first index per cloud or NULL for 0 (random_start=False). N <= 4096. */
int pfpp_fps(const float* xyz, int K, int N, int S, const int* start, int* out_idx, float* out_xyz,
             cudaStream_t stream);

/* torch_cluster.fps on ragged, large clouds (utils/node_merge_utils.py:218-220, merge stage).
 * dist_scratch has one float per input point. */
int pfpp_fps_ragged(const float* xyz, const int* cloud_start, const int* cloud_len, const int* n_samples,
                    const int* start, int n_clouds, float* dist_scratch, const int* out_start, int* out_idx,
                    cudaStream_t stream);

/* query_ball_point (utils/pn2_utils.py:92-112): out_idx [K,S,nsample]. */
int pfpp_ball_query(const float* xyz, const float* new_xyz, int K, int N, int S, float radius_sq, int nsample,
                    int* out_idx, cudaStream_t stream);

/* index_points + centroid subtraction + concat (utils/pn2_utils.py:139-148):
 * out [K*S*ns, ld] = [xyz[idx]-new_xyz | feats[idx] | 0-pad]. feats [K,N,D] (NULL when D == 0). */
int pfpp_group_gather(const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K, int N, int S,
                      int ns, int D, int ld, int out_bf16, void* out, cudaStream_t stream);

/* torch.max over nsample (utils/pn2_utils.py:214): in [G*ns, ld_in] -> out [G, ld_out]. */
int pfpp_group_max(const void* in, long long G, int ns, int C, int ld_in, int is_bf16, void* out, int ld_out,
                   cudaStream_t stream);

/* VectorQuantizer.forward, eval (vqvae/model/modules/quantizer.py:42-63): z [n_chunks,16] ->
 * out = z + (e[argmin] - z), codes (may be NULL). */
int pfpp_vq(const void* z, int z_is_bf16, long long n_chunks, const float* codebook, int n_codes, float* out,
            int* codes, cudaStream_t stream);

/* sample_and_group's gather + PointNetSetAbstraction's 3 x relu(bn(conv1x1)) + max over nsample
 * (utils/pn2_utils.py:139-148,209-214) fused in one tcgen05 kernel; activations stay in shared
 * memory / TMEM.  level in {1,2,3} selects (nsample, D, C1, C2, C3) = (32,0,64,64,128),
 * (64,128,128,128,256), (64,256,256,256,512).  feats [K,N,D] bf16.  Layer 0 is split: w0_feat [C1, D]
 * bf16 (feature columns; NULL for level 1) and w0_xyz [C1, 4] fp32 (dx,dy,dz columns: the kernel splits weights
 * and centroid offsets into bf16 hi/lo pairs and applies them as ONE extra K=16 tcgen05 step,
 * w d = w_hi d_hi + w_hi d_lo + w_lo d_hi, 2^-16 relative); w1 [C2,C1], w2 [C3,C2] bf16; all with BatchNorm folded;
 * out [K*S, C3] bf16. */
int pfpp_sa_fused(int level, const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K, int N,
                  int S, const void* w0_feat, const float* w0_xyz, const float* b0, const void* w1, const float* b1,
                  const void* w2, const float* b2, void* out, cudaStream_t stream);
/* debug variant of pfpp_sa_fused: per-CTA clock stamps of the kernel phases into trace[grid][16] */
int pfpp_sa_fused_trace(int level, const float* xyz, const float* new_xyz, const void* feats, const int* gidx, int K, int N,
                  int S, const void* w0_feat, const float* w0_xyz, const float* b0, const void* w1, const float* b1,
                  const void* w2, const float* b2, void* out, long long* trace, cudaStream_t stream);

/* ---- contractions ---------------------------------------------------------------------- */

/* C[M,N'] = epi(A[M,K] W[N,K]^T + bias) (+ residual); fp32 SIMT (parity mode).  Replaces
 * nn.Linear / 1x1 Conv2d / Conv1d calls (utils/pn2_utils.py:210-212, vqvae/model/modules/pn2.py:65,
 * denoiser/model/modules/attention.py:79-89 via diffusers, verifier_transformer.py:62).
 * K, lda, ldw multiples of 4; residual may alias C. */
int pfpp_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias, const float* residual, int ldr,
                  float* C, int ldc, int M, int N, int K, int epilogue, cudaStream_t stream);

/* Same contract on the 5th-gen tensor cores: bf16 operands (A [M,K], W [N,K], K-major), fp32
 * accumulation in TMEM, operands staged by TMA.  K multiple of 64, lda/ldw multiples of 8,
 * 16-byte aligned bases.  C is fp32 or bf16 (c_bf16); residual is fp32 and may alias an fp32 C.
 * tensor maps are encoded per call on the host (driver entry point resolved at first use). */
int pfpp_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual, int ldr,
                   void* C, int ldc, int c_bf16, int M, int N, int K, int epilogue, cudaStream_t stream);

/* fp32-grade contraction on the tensor cores ("bf16x3"): operands in the bf16 hi/lo SPLIT format -- a row of a
 * split tensor holds ld elements, the hi half bf16(v) at columns [0, ld/2) and the lo half bf16(v - hi) at
 * [ld/2, ld) -- and three tcgen05 passes A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T into one fp32 TMEM accumulator
 * (relative error ~2^-16 per product; the reference's fp32 nn.Linear / conv1x1, same call sites as pfpp_gemm_f32).
 * A [M, lda >= 2K], W [N, ldw >= 2K] split; K multiple of 8, lda / ldw multiples of 16.  C is fp32 [M, ldc]
 * (c_split = 0; residual fp32, may alias C) or split [M, ldc] (c_split = 1: the operand of the next layer).
 * GELU / GEGLU epilogues use the exact erf form.  M <= 1024 without a residual (the output heads, M = fragments) runs the
 * same three products as 32 x 64 tiles on warp-level MMAs (csrc/gemm_small.cu): a 256 x 256 tcgen05 tile pair is mostly
 * fixed pipeline cost at that size. */
int pfpp_gemm_bf16x3(const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual, int ldr,
                     void* C, int ldc, int c_split, int M, int N, int K, int epilogue, cudaStream_t stream);

/* fp32 [rows, C] (leading dimension ldx) -> split bf16 [rows, ldo] (hi at c, lo at ldo/2 + c, zero padding). */
int pfpp_split_bf16(const float* x, long long rows, int C, int ldx, void* out, int ldo, cudaStream_t stream);

/* Output-format flag of the kernels below and of pfpp_group_gather / pfpp_group_max (`out_bf16`, `is_bf16`,
 * `io_bf16`): 0 = fp32, 1 = bf16, 2 = bf16 hi/lo split rows (ld counts BOTH halves).  pfpp_group_gather(2): feats are
 * split [K,N,2D], out rows hold 2*ld elements; pfpp_group_max(2): fp32 in, split out; pfpp_attention_varlen(2): fp32
 * q/k/v in, split out. */

/* ---- denoiser tokens -------------------------------------------------------------------- */

/* EmbedderNerf features (utils/model_utils.py:40-69) for DenoiserTransformer._gen_cond
 * (denoiser_transformer.py:117-135): feat_tok [F*L, ld_tok] = [latent | nerf(xyz) | nerf(scale)],
 * feat_par [F, ld_par] = nerf(x[slot]). */
int pfpp_embed_features(const float* x, const float* scale, const int* frag_slot, const float* latent,
                        const float* xyz, int F, int L, int latent_dim, int out_bf16, void* feat_tok, int ld_tok,
                        void* feat_par, int ld_par, cudaStream_t stream);

/* _add_ref_part_emb + broadcast + PositionalEncoding (denoiser_transformer.py:150-156,173-185). */
int pfpp_combine_embed(const float* shape_emb, const float* x_emb, const float* ref_emb, const float* pe,
                       const int* frag_slot, const unsigned char* ref, int F, int P, int L, int C, float* h,
                       cudaStream_t stream);

/* nn.LayerNorm / MyAdaLayerNorm (attention.py:5-25) / post-LN residual (torch TransformerEncoderLayer).
 * mod [G, 2C] = pre-tabulated Linear(SiLU(Embedding[t])) rows; row r uses mod[row_group[r / rows_per_group]].
 * residual != NULL computes LN(x + residual) and optionally stores the sum in sum_out. C in {256,512}. */
int pfpp_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, const float* mod,
                   const int* row_group, int rows_per_group, long long rows, int C, int out_bf16, void* y,
                   float* sum_out, cudaStream_t stream);

/* Residual projection fused with the (Ada)LayerNorm that follows it (attention.py:75-92 -- diffusers Attention.to_out /
 * FeedForward.net[2] + residual -- then MyAdaLayerNorm attention.py:5-25 or norm3):
 *   h[M,512] += A[M,K] W[512,K]^T + bias  (fp32, in place);  ln_out[M,512] (bf16) = LN(h) modulated like pfpp_layernorm
 * (mod / row_group / rows_per_group, or gamma / beta when mod == NULL).  bf16 operands, tcgen05 CTA pairs, one CTA per
 * 128 full rows (the 128 x 512 fp32 accumulator fills tensor memory).  K % 8 == 0. */
int pfpp_gemm_res_ln(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int M, int K,
                     const float* mod, const int* row_group, int rows_per_group, const float* gamma, const float* beta,
                     void* ln_out, cudaStream_t stream);

/* F.scaled_dot_product_attention with the reference's masks (diffusers AttnProcessor2_0 at
 * attention.py:79,84; torch MultiheadAttention at verifier_transformer.py:62) expressed as
 * full attention inside packed segments. head_dim in {32,64}. */
int pfpp_attention_varlen(const void* qkv, int ld, int q_off, int k_off, int v_off, const int* seg_start,
                          const int* seg_len, int n_segments, int max_len, int heads, int head_dim, int io_bf16,
                          void* out, int ldo, cudaStream_t stream);

/* The denoiser's attention (attention.py:75-92 through diffusers Attention) on the tensor cores: bf16 qkv
 * [M, 3C] (q | k | v, head h at columns h*64), head_dim 64, one CTA per (segment, head); S = QK^T and
 * O = PV run as tcgen05.mma with fp32 accumulators in TMEM, operands by TMA, online softmax.
 * block == 0: global attention (attention.py:84 with the key mask): every token of a segment (one object's
 *   valid fragments) attends to the whole segment.  Segments of <= 512 tokens keep K/V resident in shared memory
 *   (one CTA per segment and head); longer segments run two query tiles per CTA and stream K/V through a ring.
 * block  > 0: the block-diagonal local attention (attention.py:79 with self_mask,
 *   denoiser_transformer.py:158-166): tokens attend inside aligned blocks of `block` tokens; a segment holds
 *   up to 4 tiles of floor(128/block) blocks (block 25: 125-token tiles, segments <= 500 tokens) and
 *   seg_start/seg_len must be multiples of the tile length (the last segment may be short).
 * out is bf16 [M, C]. */
int pfpp_attention_tc(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                      int n_segments, int max_len, int heads, int block, void* out, int ldo, cudaStream_t stream);
/* The same block-diagonal local attention (attention.py:79 with self_mask) at its natural granularity: one warp per
 * (block of `block` <= 32 consecutive tokens, head), warp-level mma.sync m16n8k16 + ldmatrix, no padding to 128-row tiles.
 * Rows [i*block, (i+1)*block) of the packed [M, 3C] bf16 QKV matrix attend to each other; M % block == 0, head_dim 64.
 * Same arithmetic as pfpp_attention_tc (bf16 operands, fp32 scores, bf16 probabilities, row sum of the rounded values). */
int pfpp_attention_local(const void* qkv, long long M, int ld, int C, int heads, int block, void* out, int ldo,
                         cudaStream_t stream);
/* debug variant of pfpp_attention_tc: per-CTA clock stamps of the pipeline phases into trace[n_ctas][48] */
int pfpp_attention_tc_trace(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                            int n_segments, int max_len, int heads, int block, void* out, int ldo, long long* trace,
                            cudaStream_t stream);

/* mean over L (denoiser_transformer.py:141-142); out_bf16 == 2: split rows [F, 2C]. */
int pfpp_mean_pool(const float* h, int F, int L, int C, int out_bf16, void* out, cudaStream_t stream);

/* DDPMScheduler.step (diffusers 0.21.4; auto_aggl.py:149) + reference-part clamp (auto_aggl.py:150) +
 * trajectory record (auto_aggl.py:151, kept on the device).  coef rows = {sqrt(1-abar_t), sqrt(abar_t),
 * c_x0, c_x, sigma}; frag_coef[f] = DDPM step index of fragment f (NULL = 0): selects the coef row, the noise
 * row noise[step * noise_step_stride ...] and the history row x_hist[step * hist_step_stride ...] (x_hist may
 * be NULL).  x, noise rows, ref_pose are [slots,7]; only valid slots are touched. */
int pfpp_ddpm_step(const float* eps, int ld_eps, const int* frag_slot, const float* coef, const int* frag_coef,
                   int add_noise, const float* noise, long long noise_step_stride, const unsigned char* ref,
                   const float* ref_pose, int F, float* x, float* x_hist, long long hist_step_stride,
                   cudaStream_t stream);

/* Device-side step counter so that one captured CUDA graph serves every DDPM step of an outer iteration
 * (replaces the host loop variable of auto_aggl.py:137): out[i] = *step ; *step += 1. */
int pfpp_step_broadcast(const int* step, int* out, int n, cudaStream_t stream);
int pfpp_step_advance(int* step, cudaStream_t stream);

/* ---- verify / merge geometry ------------------------------------------------------------ */

/* get_final_pose_pts / get_final_pose_pts_dynamic (utils/node_merge_utils.py:16-53): for ragged
 * segments, out = quat_apply(q, p * seg_scale) + t with pose row seg_pose[s]; normalise=1 divides
 * q by its norm first (get_final_pose_pts), 0 applies it raw (the _dynamic variant, App. C.2). */
int pfpp_pose_apply(const float* pts, const int* seg_start, const int* seg_len, const int* seg_pose,
                    const float* pose, const float* seg_scale, int normalise, int n_segments, float* out,
                    cudaStream_t stream);

/* get_distance_for_matching_pts + _make_cd_to_bins + feature normalisation
 * (utils/node_merge_utils.py:62-89; auto_aggl.py:181-201,385-389). feat [n_rows,7] is zeroed first. */
int pfpp_edge_features(const float* pts, const int* pair_src, const int* pair_tgt, const int* edge_start,
                       const int* edge_len, const int* edge_row, int n_edges, int max_pairs, long long n_rows,
                       float* feat, cudaStream_t stream);

/* VerifierTransformer input embedding and output head (verifier_transformer.py:49-56,63). */
int pfpp_verifier_embed(const float* feat, const int* tok_row, const int* tok_i, const int* tok_j, int n_tokens,
                        const float* W, const float* bias, const float* pe, int C, float* out, cudaStream_t stream);
int pfpp_verifier_head(const float* h, const int* tok_row, int n_tokens, const float* w, const float* b, int C,
                       float* logits, cudaStream_t stream);

/* Per-point NN^2 both ways between equal-size clouds and estimate_pointcloud_normals (kNN-20 PCA)
 * for remove_intersect_points_and_fps_ds (utils/node_merge_utils.py:159-222). */
int pfpp_merge_filter(const float* pcs, int n_clouds, int n_points, int knn, float threshold, unsigned char* keep,
                      float* normals, cudaStream_t stream);

/* The whole merge stage of one outer iteration (auto_aggl.py:234-286: merge_node, by-area shift,
 * remove_intersect_points_and_fps_ds, renormalisation; utils/node_merge_utils.py:125-135,159-222) for ALL
 * connected components of ALL objects of a batch, asynchronously on `stream`, no host round trip.
 *   posed [slots,N,3]: fragment clouds posed by pfpp_pose_apply(normalise=1, seg_scale);
 *   comp_start [n_comp+1] -> member_slot [n_clouds] (valid members of component c, concatenation order),
 *   member_comp [n_clouds], comp_pivot_slot [n_comp]; pair_i/pair_j [n_pairs]: ordered pairs (i != j) of member-list
 *   indices inside one component; uniform [n_comp]: the torch.rand(1) draws of the random-start FPS (nmu:219);
 *   area segments: by_area[seg] = by_area_posed[seg] - centroid[area_seg_comp] (auto_aggl.py:259-262).
 * Writes part_pcs[pivot slot] = ds / max|ds|, scale[pivot slot] = max|ds| and result [n_comp,8] =
 * {centroid xyz, max|ds|, kept points M, n_out = ceil(M * float32(N/M)), fps start, 0} (floats).
 * workspace >= pfpp_merge_workspace_bytes(n_clouds, n_comp, n_points). */
size_t pfpp_merge_workspace_bytes(int n_clouds, int n_comp, int n_points);
int pfpp_merge(const float* posed, int n_points, int n_comp, int n_clouds, const int* comp_start,
               const int* member_slot, const int* member_comp, const int* comp_pivot_slot, int n_pairs,
               const int* pair_i, const int* pair_j, const float* uniform, float threshold, int knn,
               int n_area_segs, const int* area_seg_start, const int* area_seg_len, const int* area_seg_comp,
               const float* by_area_posed, float* by_area, float* part_pcs, float* scale, float* result,
               void* workspace, size_t ws_bytes, cudaStream_t stream);

/* chamferdist / pytorch3d knn_points(K=1) squared distances for the evaluation metrics
 * (denoiser/evaluation/evaluator.py:108,137): out[b,i] = min_j |a[b,i] - b[b,j]|^2. */
int pfpp_nn_sqdist(const float* a, const float* b, int batches, int N, int M, float* out, cudaStream_t stream);

/* ---- coarse entry points: one call per stage ----------------------------------
 *
 * What a C caller binds to run the reference's test_step (test.py:19-43 -> AutoAgglomerative.test_step) without the
 * Python engine: weights are flat structs of DEVICE pointers (filled once per checkpoint by the host packer,
 * weights.py -> engine.py), every call sequences this library's kernels over a caller-owned workspace of
 * pfpp_*_workspace_bytes() bytes, asynchronously on `stream` (capturable in a CUDA graph), with no allocation and
 * no host synchronisation.  `mode` selects the contraction engine: 0 = fp32 SIMT, 1 = bf16 tcgen05 (fused set
 * abstraction + tensor-core attention), 2 = bf16x3 split operands on tcgen05 (fp32-grade). */
#define PFPP_MAX_LAYERS 8

typedef struct PfppLinear {
  const void* w;     /* [n, k] fp32 (mode 0) | [n, k] bf16 (mode 1) | [n, 2k] bf16 hi/lo split (mode 2); BatchNorm folded */
  const float* bias; /* [n] or NULL */
  int n, k;          /* k: the padded K the GEMM runs over (multiple of 4 fp32 / 8 bf16) */
} PfppLinear;

/* VQVAE.encode (vqvae/model/modules/vq_vae.py:52-68 -> pn2.py:57-68): 3 set-abstraction levels + conv6 + codebook */
typedef struct PfppEncoderWeights {
  int mode;
  int npoint[3], nsample[3];
  float radius_sq[3];
  PfppLinear sa[3][3];
  const void* sa_w0_feat[3]; /* mode 1, fused kernel: layer-0 feature columns [C1, D] bf16 (NULL at level 1) */
  const float* sa_w0_xyz[3]; /* mode 1, fused kernel: layer-0 centroid-offset columns [C1, 4] fp32 */
  PfppLinear conv6;
  const float* codebook; /* [n_codes, 16] */
  int n_codes, latent_points, latent_dim;
  int chunk_frags; /* fragments per pass of the unfused path (bounds its [rows, C] intermediates) */
  int fused_sa;    /* mode 1: use pfpp_sa_fused */
} PfppEncoderWeights;

typedef struct PfppDenoiserLayer {
  PfppLinear qkv[2], out[2]; /* [0] = self_attn (block-diagonal), [1] = global_attn (attention.py:75-92) */
  PfppLinear ff1, ff2;       /* GEGLU rows interleaved (value_j, gate_j) */
  const float* norm3_w;
  const float* norm3_b;
} PfppDenoiserLayer;

/* DenoiserTransformer (denoiser_transformer.py:13-202) + the scheduler's coefficient table */
typedef struct PfppDenoiserWeights {
  int mode, C, heads, n_layers, P, L, latent_dim, T;
  int tc_attention, local_tiles; /* mode 1: tcgen05 attention; local attention: 0 = one warp per fragment and head
                                  * (pfpp_attention_local), n > 0 = tcgen05 with n 125-token tiles per CTA */
  int fused_ln; /* mode 1: out-proj / FF2 + residual + the following (Ada)LayerNorm as ONE kernel (pfpp_gemm_res_ln) */
  PfppLinear shape_embedding, param_fc;
  const float* ref_emb; /* [2, C] */
  const float* pe;      /* [P, C] */
  const float* mod;     /* [2 * n_layers, T, 2C]: AdaLN rows Linear(SiLU(Embedding[t])) for the T inference timesteps */
  const float* coef;    /* [T, 5] = {sqrt(1-abar_t), sqrt(abar_t), c_x0, c_x, sigma} (pfpp_ddpm_step) */
  PfppDenoiserLayer layers[PFPP_MAX_LAYERS];
  PfppLinear head0, head_t2, head_r2, head_t4, head_r4; /* fp32 (mode 0) | split, as mode 2 (modes 1 and 2) */
} PfppDenoiserWeights;

typedef struct PfppVerifierLayer {
  PfppLinear qkv, out, l1, l2;
  const float *n1w, *n1b, *n2w, *n2b;
} PfppVerifierLayer;

/* VerifierTransformer (verifier_transformer.py:9-65) */
typedef struct PfppVerifierWeights {
  int C, heads, n_layers, ffn;
  int tc; /* 1: projections / FFN as split-operand tcgen05 GEMMs (weights in the mode-2 format), 0: fp32 SIMT */
  const float *emb_w, *emb_b, *pe, *out_w, *out_b;
  PfppVerifierLayer layers[PFPP_MAX_LAYERS];
} PfppVerifierWeights;

/* AutoAgglomerative._apply_rots + _extract_features (auto_aggl.py:70-92) for the F packed fragments frag_slot[f]:
 * part_pcs [slots,N,3], x [slots,7] pose rows (the quaternion is normalised here) -> z_q [F*L, latent_dim],
 * xyz [F,L,3], codes [F*L*latent_dim/16] (may be NULL).  out_pos != NULL: fragment f's rows are written at row
 * block out_pos[f] of z_q / xyz instead of f (a compacted pass filling its rows of a larger packed batch). */
size_t pfpp_encoder_workspace_bytes(const PfppEncoderWeights* w, int F, int N);
int pfpp_encoder_forward(const PfppEncoderWeights* w, const float* part_pcs, const int* frag_slot, const float* x, int F,
                         int N, float* z_q, float* xyz, int* codes, const int* out_pos, void* workspace, size_t ws_bytes,
                         cudaStream_t stream);

/* DenoiserTransformer.forward (denoiser_transformer.py:169-202) on the packed batch -> eps [F, 8] (columns 0..6).
 * frag_step[f] = index of fragment f's timestep in the inference schedule; attention segments over the packed
 * tokens: per fragment (start f*L, length L) and per object (its valid fragments, <= max_global tokens). */
size_t pfpp_denoiser_workspace_bytes(const PfppDenoiserWeights* w, int F);
int pfpp_denoiser_forward(const PfppDenoiserWeights* w, const float* x, const float* scale, const unsigned char* ref,
                          const int* frag_slot, const int* frag_step, const float* latent, const float* xyz,
                          const int* frag_seg_start, const int* frag_seg_len, const int* obj_seg_start,
                          const int* obj_seg_len, int F, int n_obj, int max_global, float* eps, void* workspace,
                          size_t ws_bytes, cudaStream_t stream);

/* One whole DDPM step of auto_aggl.py:137-151 for the packed batch: step index = *step_counter (device) ->
 * encoder -> denoiser -> DDPMScheduler.step + reference clamp + history row -> *step_counter += 1.
 * noise [T, slots, 7] / x_hist [T, slots, 7] (row stride given; x_hist may be NULL); eps_out [F,8] may be NULL.
 * latent [F*L, latent_dim] / xyz [F,L,3] are caller-owned and persist from step to step: with enc_slot / enc_pos
 * (F_enc entries: slot and packed position of the fragments to re-encode) only those fragments are encoded this
 * step -- the reference parts, whose pose is clamped every step (auto_aggl.py:150), keep the rows the caller wrote
 * once with pfpp_encoder_forward(out_pos) when the outer iteration began (same arithmetic, same results).
 * enc_slot == NULL: every fragment of frag_slot is encoded. */
size_t pfpp_step_workspace_bytes(const PfppEncoderWeights* we, const PfppDenoiserWeights* wd, int F, int N);
int pfpp_denoiser_step(const PfppEncoderWeights* we, const PfppDenoiserWeights* wd, const float* part_pcs, float* x,
                       const float* scale, const unsigned char* ref, const float* ref_pose, const int* frag_slot,
                       int* frag_step, int* step_counter, const float* noise, long long noise_step_stride, float* x_hist,
                       long long hist_step_stride, const int* frag_seg_start, const int* frag_seg_len,
                       const int* obj_seg_start, const int* obj_seg_len, int F, int n_obj, int max_global, int N,
                       const int* enc_slot, const int* enc_pos, int F_enc, float* latent, float* xyz, float* eps_out,
                       void* workspace, size_t ws_bytes, cudaStream_t stream);

/* VerifierTransformer.forward (verifier_transformer.py:42-65) on packed valid-edge tokens (tok_row = dense edge row
 * b*E + e, tok_i / tok_j = fragment indices, one segment per object) -> logits [n_rows] (zero on invalid rows). */
size_t pfpp_verifier_workspace_bytes(const PfppVerifierWeights* w, int n_tokens);
int pfpp_verifier_forward(const PfppVerifierWeights* w, const float* feat, const int* tok_row, const int* tok_i,
                          const int* tok_j, int n_tokens, const int* seg_start, const int* seg_len, int n_segments,
                          int max_len, long long n_rows, float* logits, void* workspace, size_t ws_bytes,
                          cudaStream_t stream);

/* sizeof of the structs above as compiled into the library (which: 0 PfppLinear, 1 PfppEncoderWeights,
 * 2 PfppDenoiserLayer, 3 PfppDenoiserWeights, 4 PfppVerifierLayer, 5 PfppVerifierWeights; 0 for anything else). */
size_t pfpp_struct_bytes(int which);

/* Hash of the sources this library was built from (build.py: FNV-1a 64 of csrc + this header), so a caller can
 * tie the loaded binary to a source tree. */
unsigned long long pfpp_build_id(void);

/* ---- Chamfer distance with gradients ----------------------------------- */

/* chamfer_cuda.chamfer_forward (Jigsaw_matching/utils/chamfer/cuda/chamfer_kernel.cu:31-173): nearest neighbour of
 * every point of xyz1 [B,n1,3] in xyz2 [B,n2,3] and vice versa: squared distances and the FIRST minimising index. */
int pfpp_chamfer_forward(const float* xyz1, const float* xyz2, int batches, int n1, int n2, float* dist1, int* idx1,
                         float* dist2, int* idx2, cudaStream_t stream);
/* chamfer_cuda.chamfer_backward (chamfer_kernel.cu:175-260): grad_xyz1 [B,n1,3], grad_xyz2 [B,n2,3] (zeroed here). */
int pfpp_chamfer_backward(const float* grad_dist1, const float* grad_dist2, const float* xyz1, const float* xyz2,
                          const int* idx1, const int* idx2, int batches, int n1, int n2, float* grad_xyz1,
                          float* grad_xyz2, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PFPP_H_ */
