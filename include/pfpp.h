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
 * GELU / GEGLU epilogues use the exact erf form. */
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
/* debug variant of pfpp_attention_tc: per-CTA clock stamps of the pipeline phases into trace[n_ctas][48] */
int pfpp_attention_tc_trace(const void* qkv, long long M, int ld, int C, const int* seg_start, const int* seg_len,
                            int n_segments, int max_len, int heads, int block, void* out, int ldo, long long* trace,
                            cudaStream_t stream);

/* mean over L (denoiser_transformer.py:141-142). */
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
