/* esmdiff_b200 -- C ABI of the B200-native ESMDiff ddpm sampling path.
 *
 * The reference (lujiarui/esmdiff) is pure Python: it has no FFI.  The boundary this library
 * stands behind is three Python call sites; each entry point below names the
 * reference interface it replaces.  Plain pointers and sizes only -- no torch types.
 *
 *   - All `*_dev` pointers are device pointers on the context's device; the library borrows
 *     them for the duration of the (stream-ordered, asynchronous) call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Every function returns 0 on success, non-zero on failure; esmdiff_last_error() gives text.
 *   - One context per device; not thread-safe; calls are ordered on the given stream.
 *   - There is no CPU fallback: on a machine without a CUDA device esmdiff_create() fails.
 */
#ifndef ESMDIFF_B200_H
#define ESMDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESMDIFF_ABI_VERSION 2

#define ESMDIFF_STRUCTURE_MASK_TOKEN 4096 /* esm.utils.constants.esm3.STRUCTURE_MASK_TOKEN; model.py:381 */

typedef struct esmdiff_ctx esmdiff_ctx;

/* Hyper-parameters: CustomizedESM3.__init__ (slm/models/net.py:323-334), configs/experiment/mdlm.yaml:26-58 */
typedef struct esmdiff_cfg {
    int32_t d_model;            /* 1536 */
    int32_t n_heads;            /* 24   (d_head must be 64) */
    int32_t n_layers;           /* 48   (residual scale = sqrt(n_layers / 36)) */
    int32_t ffn_hidden;         /* 4096 = ceil(8/3 d / 256) * 256 */
    int32_t n_structure_heads;  /* 4101 */
    int32_t seq_vocab;          /* 64 */
    int32_t struct_vocab;       /* 4101 */
    int32_t time_freq_dim;      /* 256, TimestepEmbedder.frequency_embedding_size (net.py:487) */
    int32_t time_conditioning;  /* mdlm.yaml:40; 0 -> sigma is zeroed (model.py:538-539) */
    int32_t model_kind;         /* 0: the sampling network above.  1: the VQ-VAE structure token decoder
                                 * (esm StructureTokenDecoder, ESM3_structure_decoder_v0: d_model 1280, 20 heads,
                                 * 30 blocks, ffn_hidden 3584, scale_residue=False); then n_structure_heads is
                                 * the width of Dim6RotStructureHead.proj (23) and struct_vocab the rows of
                                 * `embed` (4101); seq_vocab, time_* are unused */
    int32_t n_aux_out;          /* model_kind 1: bins of the pLDDT head (50), 0 = head absent */
    int32_t v_heads;            /* 256: heads of block 0's geometric attention (net.py:341); 0 = its weights are dropped
                                 * and esmdiff_set_structure_coords is unavailable (the ddpm path never needs them) */
    int32_t reserved[4];
} esmdiff_cfg;

enum esmdiff_dtype { ESMDIFF_F32 = 0, ESMDIFF_BF16 = 1 };

int esmdiff_abi_version(void);
const char* esmdiff_last_error(const esmdiff_ctx* ctx); /* ctx may be NULL: last create() error */

/* Replaces hydra.utils.instantiate(cfg.model) + .to(device) (slm/utils/checkpoint_utils.py:59,72). */
int esmdiff_create(const esmdiff_cfg* cfg, int device, esmdiff_ctx** out);
int esmdiff_destroy(esmdiff_ctx* ctx);

/* MaskedDiffusionLanguageModeling(time_conditioning=...) (model.py:333, 538-539) can differ from the
 * value the network was created with: 0 -> sigma is zeroed before the time embedding. */
int esmdiff_set_time_conditioning(esmdiff_ctx* ctx, int on);

/* Replaces model.load_state_dict(all_params) (checkpoint_utils.py:63-64).  `key` is the state-dict
 * key of the DeepSpeed ['module'] dict ("net.transformer.blocks.0.attn.out_proj.weight",
 * "sigma_embedder.mlp.0.bias", ...).  Data is copied (and converted: GEMM weights are kept in
 * bf16, FFN W1 rows are interleaved gate/up per 128).  Keys the ddpm path never reads
 * (function/residue embeddings; geom_attn.* when cfg.v_heads is 0) are accepted and dropped; block 0's
 * geom_attn.* are kept for esmdiff_set_structure_coords otherwise.  Unknown keys fail. */
int esmdiff_set_weight(esmdiff_ctx* ctx, const char* key, const void* data, int on_device,
                       int dtype, const int64_t* shape, int ndim);
/* Strict check that every key of the path has been set; builds derived constants. */
int esmdiff_finalize_weights(esmdiff_ctx* ctx);

/* Replaces sigma_embedder(sigma) (model.py:466-471, net.py:519-522) for one sigma shared by the
 * batch: cond_out_dev[d_model].  */
int esmdiff_time_embed(esmdiff_ctx* ctx, float sigma, float* cond_out_dev, void* stream);

/* Replaces self.net(structure_tokens=xt, sequence_tokens=seq, auxiliary_embeddings=aux,
 * labels=None).structure_logits (model.py:475-481; net.py:371-483).
 *   seq_dev, xt_dev : int64 [B*T]
 *   aux_dev         : fp32, row m at aux_dev + m*aux_row_stride (stride 0 = one vector for all
 *                     rows); NULL = no auxiliary embedding
 *   logits_out_dev  : fp32 [B*T, n_structure_heads], row-major contiguous
 *   embeddings_out_dev : fp32 [B*T, d_model] pre-final-norm residual stream, or NULL */
int esmdiff_forward(esmdiff_ctx* ctx, const int64_t* seq_dev, const int64_t* xt_dev, int B, int T,
                    const float* aux_dev, int64_t aux_row_stride, float* logits_out_dev,
                    float* embeddings_out_dev, void* stream);
/* Replaces the structure_coords argument of CustomizedESM3.forward (net.py:385, 433-441): backbone
 * coordinates of the batch the NEXT forwards / sampling loops run on -> build_affine3d_from_coordinates, and
 * block 0's geometric attention (esm GeometricReasoningOriginalImpl, v_heads 256, mask_and_zero_frameless=True;
 * restated in oracle/geom_ref.py, parity unpinned) becomes live: x += out_proj(geom(s_norm(x))) / scale between the
 * attention and the FFN of block 0.
 *   coords_dev : fp32 [B, T, 3, 3]  N, CA, C per token position (BOS / EOS and unknown residues: NaN or inf), or
 *                NULL = back to the default of the ddpm path (NaN everywhere: the branch is exactly 0 and skipped)
 * Needs the geom_attn.* weights of block 0 and cfg.v_heads. */
int esmdiff_set_structure_coords(esmdiff_ctx* ctx, const float* coords_dev, int B, int T, void* stream);
/* Same with the time embedding computed inside from sigma (_model_wrapper, model.py:464-481). */
int esmdiff_forward_sigma(esmdiff_ctx* ctx, const int64_t* seq_dev, const int64_t* xt_dev, int B,
                          int T, float sigma, float* logits_out_dev, void* stream);

/* Replaces logits_parameterization (model.py:527-533): logp_out_dev[B*T, V] (may alias logits). */
int esmdiff_logits_parameterization(esmdiff_ctx* ctx, const float* logits_dev,
                                    const int64_t* xt_dev, int B, int T, float* logp_out_dev,
                                    void* stream);
/* Replaces logits_parameterization + the tail of _ddpm_update + _sample_categorical
 * (model.py:527-533, 602-607, 24-28).  x_inout_dev int64 [B*T] is updated in place.
 * u_dev: the uniforms torch.rand_like(q_xs) would draw, fp32 [B*T, V]; NULL = library Philox
 * stream keyed by (seed, step).  mc_t/mc_s: move chances 1-exp(-sigma) (model.py:592-593). */
int esmdiff_sample_step(esmdiff_ctx* ctx, int64_t* x_inout_dev, const float* logits_dev,
                        const float* u_dev, float mc_t, float mc_s, int B, int T, uint64_t seed,
                        uint32_t step, void* stream);
/* Replaces the noise-removal argmax (model.py:575-579). */
int esmdiff_denoise_argmax(esmdiff_ctx* ctx, int64_t* x_inout_dev, const float* logits_dev, int B,
                           int T, void* stream);

/* LogLinearNoise schedule of ddpm_sample/_ddpm_update (model.py:564-567, 584-593;
 * noise_utils.py:205-206) in C float arithmetic: sigma[steps+1], mc_t[steps], mc_s[steps]. */
int esmdiff_schedule(int steps, float eps, float noise_eps, float* sigma, float* mc_t, float* mc_s);

/* Replaces MaskedDiffusionLanguageModeling.ddpm_sample (model.py:543-581), device resident.
 *   prior_dev : int64 [B*T] or NULL (all MASK);  out_dev : int64 [B*T]
 *   sigma[steps+1], mc_t[steps], mc_s[steps] : host arrays (esmdiff_schedule or the caller's)
 *   uniforms come from the library Philox stream (seed); noise_removal as model.py:575. */
int esmdiff_ddpm_sample(esmdiff_ctx* ctx, const int64_t* seq_dev, const int64_t* prior_dev, int B,
                        int T, int steps, const float* sigma, const float* mc_t, const float* mc_s,
                        uint64_t seed, int noise_removal, int64_t* out_dev, void* stream);
/* Same through HOST buffers (pageable or pinned): H2D of seq/prior, the loop, D2H of the ids,
 * stream synchronised on return.  This is the end-to-end call bench.py times as `e2e`. */
int esmdiff_ddpm_sample_host(esmdiff_ctx* ctx, const int64_t* seq_host, const int64_t* prior_host,
                             int B, int T, int steps, float eps, uint64_t seed, int noise_removal,
                             int64_t* out_host);

/* --mode gibbs (reference slm/sample_esmdiff.py:66-130 -> esm.utils.generation.iterative_sampling_raw with
 * GenerationConfig(track="structure", num_steps, temperature, top_p); esm==3.0.4, restated in oracle/gibbs_ref.py).
 * One decoding step on device-resident tokens, after a forward WITHOUT time conditioning:
 *   every still-masked row: top-p filter of the raw logits, special ids (>= 4096) excluded, candidate id =
 *   argmax softmax(l / temperature) / Exp(1) (= torch.multinomial(p, 1)), entropy of the filtered distribution;
 *   then the k masked positions of lowest entropy of every sample take their candidate
 *   (esm.utils.generation._get_iterative_sampling_mask_for_prompt_and_step, strategy "entropy").
 *   noise_dev : fp32 [B*T, 4101] Exp(1) draws of the caller (torch's exponential_()), or NULL = library Philox stream */
int esmdiff_gibbs_step(esmdiff_ctx* ctx, int64_t* x_inout_dev, const float* logits_dev, const float* noise_dev,
                       int B, int T, float temperature, float top_p, int k, uint64_t seed, uint32_t step,
                       void* stream);
/* The whole loop, device resident: out = prior (BOS, MASK..., EOS or a partly masked prompt); steps x {forward, step}.
 *   k_per_step[steps] : host array, positions to unmask at each step (the cosine schedule, esmdiff_b200/gibbs.py) */
int esmdiff_gibbs_sample(esmdiff_ctx* ctx, const int64_t* seq_dev, const int64_t* prior_dev, int B, int T, int steps,
                         const int* k_per_step, float temperature, float top_p, uint64_t seed, int64_t* out_dev,
                         void* stream);

/* Replaces the structure half of esm3_model.decode(prot) (slm/sample_esmdiff.py:56-61 -> esm
 * ESM3.decode -> StructureTokenDecoder.decode + ProteinChain.infer_oxygen), BATCHED over all samples
 * instead of the reference's serial B=1 loop (sample_esmdiff.py:225-230).  Context of model_kind 1.
 *   tokens_dev   : int64 [B*T] structure tokens INCLUDING the BOS (4098) / EOS (4097) positions
 *   bb_out_dev   : fp32 [B*T, 3, 3]  N, CA, C of every position (rows of BOS/EOS are meaningless)
 *   o_out_dev    : fp32 [B*T, 3]     inferred carbonyl O (NaN at BOS/EOS and the last residue), or NULL
 *   plddt_out_dev: fp32 [B*T]        CategoricalMixture mean of the pLDDT head in [0,1], or NULL
 *   affine_out_dev: fp32 [B*T, n_structure_heads] raw Dim6RotStructureHead.proj output, or NULL
 * Weight keys: StructureTokenDecoder.state_dict() names ("embed.weight",
 * "decoder_stack.blocks.0.attn.out_proj.weight", "affine_output_projection.proj.bias",
 * "plddt_head.3.weight", ...); "pairwise_classification_head.*" (pTM / PAE, not written to the PDB)
 * is accepted and dropped. */
int esmdiff_decode_structure(esmdiff_ctx* ctx, const int64_t* tokens_dev, int B, int T,
                             float* bb_out_dev, float* o_out_dev, float* plddt_out_dev,
                             float* affine_out_dev, void* stream);

/* ---- VQ-VAE structure encoder: the inpainting front end ------------------------------------------------
 * Replaces the structure half of esm3_model.encode(ESMProtein(sequence, coordinates)) in protseq_to_data
 * (slm/models/utils.py:136-137; consumers slm/sample_esmdiff.py:166-175, 197-209) -> esm tokenize_structure ->
 * StructureTokenEncoder.encode (esm==3.0.4, restated in oracle/vqvae_enc_ref.py, parity unpinned).
 * ESM3_structure_encoder_v0: d_model 1024, v_heads 128, 2 blocks (geometric attention + SwiGLU FFN, hidden 2816),
 * d_out 128, 4096 codes, 16 nearest neighbours, 32 relative-position bins.  Computes in fp32 (the result is an
 * index).  Weight keys: StructureTokenEncoder.state_dict() names ("transformer.blocks.0.geom_attn.proj.weight",
 * "transformer.blocks.0.ffn.1.weight", "transformer.norm.weight", "pre_vq_proj.bias", "codebook.embeddings",
 * "relative_positional_embedding.embedding.weight", ...); the codebook's EMA buffers are accepted and dropped. */
typedef struct esmdiff_encoder esmdiff_encoder;
typedef struct esmdiff_encoder_cfg {
    int32_t d_model, v_heads, n_layers, ffn_hidden, d_out, n_codes, knn, rel_bins;
    int32_t reserved[8];
} esmdiff_encoder_cfg;
int esmdiff_encoder_create(const esmdiff_encoder_cfg* cfg, int device, esmdiff_encoder** out);
int esmdiff_encoder_destroy(esmdiff_encoder* enc);
const char* esmdiff_encoder_last_error(const esmdiff_encoder* enc);   /* enc may be NULL: last create() error */
int esmdiff_encoder_set_weight(esmdiff_encoder* enc, const char* key, const void* data, int on_device, int dtype,
                               const int64_t* shape, int ndim);      /* dtype: ESMDIFF_F32 */
int esmdiff_encoder_finalize(esmdiff_encoder* enc);
/*   coords_dev        : fp32 [B, L, 3, 3]  N, CA, C per residue; NaN / inf = unknown (mask_ids: utils.py:121)
 *   residue_index_dev : int64 [B, L] or NULL (= positions; tokenize_structure passes 1..L, same differences)
 *   codes_out_dev     : int64 [B, L]  code per residue (frameless residues: the code nearest to pre_vq_proj.bias)
 *   z_out_dev         : fp32 [B, L, d_out] pre-quantisation vectors, or NULL
 *   edges_out_dev     : int32 [B, L, min(knn, L)] neighbour lists, or NULL */
int esmdiff_encode_structure(esmdiff_encoder* enc, const float* coords_dev, const int64_t* residue_index_dev, int B,
                             int L, int64_t* codes_out_dev, float* z_out_dev, int32_t* edges_out_dev, void* stream);
/* Single kernels of that path (unit parity tests; the same kernels serve block 0 of the sampling network when
 * coordinates are given, esmdiff_set_structure_coords):
 *   build_affine3d_from_coordinates (net.py:441): rot [B*L, 9] row-major, trans [B*L, 3], mask uint8 [B*L];
 *   GeometricReasoningOriginalImpl between proj and out_proj: proj_dev fp32 [G*S, 15 H] -> out fp32 [G*S, 3 H]
 *   for G groups of S rows; frame_idx int32 [G*S] row -> frame (NULL: the row itself); work_dev [G*S, 15 H]. */
int esmdiff_op_backbone_frames(const float* coords_dev, int B, int L, float* rot_out_dev, float* trans_out_dev,
                               uint8_t* mask_out_dev, void* stream);
int esmdiff_op_geometric_attention(const float* proj_dev, const float* rot_dev, const float* trans_dev,
                                   const uint8_t* mask_dev, const int32_t* frame_idx_dev, const float* rot_scale_dev,
                                   const float* dist_scale_dev, int G, int S, int H, int zero_frameless,
                                   float* work_dev, float* out_dev, void* stream);

/* Waits for the stream and reports asynchronous failures: CUDA errors, out-of-range token ids
 * (the reference raises IndexError in nn.Embedding), pipeline watchdog trips. */
int esmdiff_synchronize(esmdiff_ctx* ctx, void* stream);
/* Number of kernels this library has launched since create (bench.py's gpu_launches claim). */
int64_t esmdiff_launch_count(const esmdiff_ctx* ctx);

/* Per-kernel device timing for roofline reports (bench.py): while enabled, every launch of the
 * kinds below is bracketed by CUDA events on its own stream.  esmdiff_profile_read sums, for one
 * kind, the event-measured milliseconds, the algorithmic work (FLOPs for the tensor-core kernels,
 * bytes for the HBM-bound ones) and the launch count since the last enable(1). */
enum esmdiff_prof_kind {
    ESMDIFF_PROF_GEMM_STORE_BF16 = 0, /* QKV projection                         FLOPs */
    ESMDIFF_PROF_GEMM_RESID_F32 = 1,  /* out_proj / FFN W2 + residual           FLOPs */
    ESMDIFF_PROF_GEMM_SWIGLU = 2,     /* FFN W1 + SwiGLU                        FLOPs */
    ESMDIFF_PROF_GEMM_BIAS_GELU = 3,  /* head Linear + GELU                     FLOPs */
    ESMDIFF_PROF_GEMM_BIAS = 4,       /* head output Linear                     FLOPs */
    ESMDIFF_PROF_ATTENTION = 5,       /*                                        FLOPs */
    ESMDIFF_PROF_SAMPLER = 6,         /* logits (+ uniforms) read               bytes */
    ESMDIFF_PROF_LAYERNORM = 7,       /* fp32 in + bf16 out                     bytes */
    ESMDIFF_PROF_QK_NORM_ROPE = 8,    /* q,k bf16 in + out                      bytes */
    ESMDIFF_PROF_EMBED = 9            /* two gathers + one store, fp32          bytes */
};
int esmdiff_profile_enable(esmdiff_ctx* ctx, int on);
int esmdiff_profile_read(esmdiff_ctx* ctx, int kind, double* ms, double* work, int64_t* launches);

/* ---- single-kernel entry points (unit parity tests, roofline timing) ---------------------- */
/* C = A[M,K] (bf16) * W[N,K]^T (bf16), epilogue: 0 store bf16, 1 resid fp32 (out += acc/scale),
 * 2 SwiGLU bf16 (W rows pre-interleaved, out [M, N/2]), 3 bias+GELU fp32, 4 bias fp32. */
int esmdiff_op_gemm(esmdiff_ctx* ctx, int epilogue, const void* a_bf16_dev, const void* w_bf16_dev,
                    int M, int N, int K, void* out_dev, int64_t ldo, const float* bias_dev,
                    float scale, void* stream);
/* The LayerNorm-folded epilogues (the block pre-LNs of esm UnifiedTransformerBlock, applied as
 * LN(x) W^T = rstd (x (gamma.W)^T - mean colsum(gamma.W)) + beta W^T):
 *   5 store bf16 with row statistics, 6 resid fp32 + bf16 copy of the new rows (xb_out [M,N]) +
 *   partial statistics (stats_out), 7 SwiGLU bf16 with row statistics.
 *   stats_out: float2, CAPACITY [M, N/96]: (mean, M2) per span of the updated rows, densely packed with
 *   N/span entries per row, span = 128 or -- when the library runs the residual GEMM with 192-wide tiles
 *   (finer wave quantisation, chosen per launch) -- 96; esmdiff_op_stats_span() tells which after the
 *   call, and a consumer call given the same pointer as stats_in uses it automatically.
 *   stats_in: such a buffer (any other pointer is taken as 128-column spans, float2 [M, K/128]);
 *   colsum/bias: [N]. */
int esmdiff_op_gemm_ln(esmdiff_ctx* ctx, int epilogue, const void* a_bf16_dev, const void* w_bf16_dev,
                       int M, int N, int K, void* out_dev, int64_t ldo, const float* bias_dev,
                       float scale, const void* stats_in_dev, const float* colsum_dev,
                       void* stats_out_dev, void* xb_out_dev, void* stream);
/* dst bf16 [rows, cols] = W * gamma (columns), colsum[rows] = row sums of dst, bias[rows] = W beta
 * (beta may be NULL); optional SwiGLU row interleave as esmdiff_op_convert_bf16. */
int esmdiff_op_fold_layernorm(esmdiff_ctx* ctx, const float* w_dev, const float* gamma_dev,
                              const float* beta_dev, void* dst_bf16_dev, float* colsum_dev,
                              float* bias_dev, int64_t rows, int64_t cols, int swiglu_hidden,
                              void* stream);
/* y bf16 [M,D] = LayerNorm(x fp32 [M,D]) * w + b (b may be NULL), eps 1e-5. */
int esmdiff_op_layernorm(esmdiff_ctx* ctx, const float* x_dev, const float* w_dev,
                         const float* b_dev, void* y_bf16_dev, int M, int D, void* stream);
/* in place on qkv bf16 [B*T, 3D]: q_ln/k_ln over D then rotary per 64-wide head. */
int esmdiff_op_qk_norm_rope(esmdiff_ctx* ctx, void* qkv_bf16_dev, const float* q_w_dev,
                            const float* k_w_dev, int B, int T, int D, void* stream);
/* ctx bf16 [B*T, D] = softmax(q k^T / 8) v per head, from qkv bf16 [B*T, 3D]. */
int esmdiff_op_attention(esmdiff_ctx* ctx, const void* qkv_bf16_dev, void* ctx_bf16_dev, int B,
                         int T, int H, void* stream);
/* The same with q_ln / k_ln and the rotary embedding folded in (esm MultiHeadAttention: LayerNorm
 * over the full width of q and of k, weight only, then rotate-half RoPE per 64-wide head):
 *   - esmdiff_op_fold_layernorm_centered: as esmdiff_op_fold_layernorm, with the column means of
 *     each block of `center_block` rows removed from the first `center_rows` rows first (the q and
 *     k thirds of the QKV weight), so that the GEMM yields q - mean(q), k - mean(k);
 *   - esmdiff_op_gemm_qkv_rope: epilogue 8.  out bf16 [M, N]: columns < n_rope hold
 *     rope(gamma * y) with y the centred, pre-LN-folded projection, the rest (v) y itself;
 *     qk_sumsq_out fp32 [M, n_rope/128]: sum of y^2 over each 128-column span;
 *   - esmdiff_op_attention_ln: attention on that layout; 1/sqrt(mean y^2 + eps) of the q and k rows
 *     is applied inside (qk_sumsq [B*T, 2*H*64/128]: q spans then k spans). */
int esmdiff_op_stats_span(const esmdiff_ctx* ctx);   /* span of the statistics the last epilogue-6 call left */
int esmdiff_op_fold_layernorm_centered(esmdiff_ctx* ctx, const float* w_dev, const float* gamma_dev,
                                       const float* beta_dev, void* dst_bf16_dev, float* colsum_dev,
                                       float* bias_dev, int64_t rows, int64_t cols, int64_t center_rows,
                                       int64_t center_block, void* stream);
int esmdiff_op_gemm_qkv_rope(esmdiff_ctx* ctx, const void* a_bf16_dev, const void* w_bf16_dev, int M,
                             int N, int K, void* out_bf16_dev, int64_t ldo, const float* bias_dev,
                             const void* stats_in_dev, const float* colsum_dev,
                             const float* qk_gamma_dev, float* qk_sumsq_out_dev, int T, int n_rope,
                             void* stream);
int esmdiff_op_attention_ln(esmdiff_ctx* ctx, const void* qkv_bf16_dev, const float* qk_sumsq_dev,
                            void* ctx_bf16_dev, int B, int T, int H, void* stream);
/* fp32 [rows, cols] -> bf16, optional SwiGLU row interleave (swiglu_hidden > 0). */
int esmdiff_op_convert_bf16(esmdiff_ctx* ctx, const float* src_dev, void* dst_bf16_dev,
                            int64_t rows, int64_t cols, int swiglu_hidden, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESMDIFF_B200_H */
