/*
 * apx.h — C ABI of libapx.so, the sm_100a implementation of AP-adapter's hot path.
 *
 * Every entry point returns 0 on success and a non-zero APX_E* code on failure; the
 * message is available from apx_last_error() (thread-local).  Nothing throws across
 * the ABI, nothing falls back to a CPU or library path: an unsupported shape / dtype
 * is an error.  All pointers are DEVICE pointers unless the name ends in `_host`.
 * All launches go to the `stream` argument (a cudaStream_t passed as void*), perform
 * no allocation and no host synchronisation, and are CUDA-graph capturable.
 *
 * Reference interfaces replaced (paths relative to the AP-adapter repository):
 *   apx_ipattn_fwd(_add)  APadapter/ap_adapter/attention_processor.py:347-470
 *                         (IPAttnProcessor2_0.__call__: to_q, two SDPAs, scale-add, to_out)
 *   apx_ipattn_bwd        autograd of the same call in train_apadapter_v2.py:939-958 (accelerator.backward through the
 *                         processor: dX, d to_k_ip.weight, d to_v_ip.weight)
 *   apx_ipattn_pack       one-time re-layout of attn.to_q.weight / attn.to_out[0].weight (:387,:457) for the site kernel
 *   apx_ctx_kv            attention_processor.py:400-415,435-439 (to_k/to_v on the text rows,
 *                         to_k_ip/to_v_ip on the audio rows; timestep-invariant, cached)
 *   apx_linear            the nn.Linear dispatches on the path (attention_processor.py:387,457;
 *                         timm Block qkv/proj/fc1/fc2 used by audio_encoder/models_mae.py:220-231)
 *   apx_attention         F.scaled_dot_product_attention call sites
 *   apx_flash_attention   the same for long key sequences (the UNet's self-attention sites)
 *                         (attention_processor.py:274,429,443; timm Attention in models_mae.py:566)
 *   apx_audiomae_fwd      audio_encoder/AudioMAE.py:190-212 (forward) -> :37-55 ->
 *                         audio_encoder/models_mae.py:548-570 (encoder) -> AudioMAE.py:148-182 (pool)
 *   apx_groupnorm_nhwc    nn.GroupNorm + F.silu (+ the conv-bias / time-embedding adds in front of it) inside diffusers' ResnetBlock2D /
 *   apx_add_bias_nhwc     Transformer2DModel as wired by pipeline/modeling_audioldm2.py:1047-1058,1223-1233,1399-1410 (third-party
 *                         leaves)
 *   apx_fbank(_batch)     audio_encoder/AudioMAE.py:356-394 (extract_kaldi_fbank_feature; torchaudio's Kaldi fbank is a
 *                         third-party dependency of the reference, restated from torchaudio 2.x compliance/kaldi.py)
 *   apx_resample          audio_encoder/AudioMAE.py:360-366 (torchaudio.functional.resample to 16 kHz)
 *   apx_patch_embed       audio_encoder/models_mae.py:19-44 (PatchEmbed_org) + :552,561-563
 *   apx_layernorm         nn.LayerNorm(eps=1e-6) inside timm Block / models_mae.py:568
 *   apx_pool_tokens       audio_encoder/AudioMAE.py:148-182 (drop CLS, (avg+max)/2 pooling)
 */
#ifndef APX_H_
#define APX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------------------- */
#define APX_OK            0
#define APX_EINVAL        1   /* bad argument (null pointer, negative size, misalignment)   */
#define APX_EUNSUPPORTED  2   /* shape / dtype outside what the sm_100a kernels implement    */
#define APX_ECUDA         3   /* a CUDA runtime / driver call failed                         */
#define APX_EWORKSPACE    4   /* caller-provided workspace too small                         */

/* ---- element types ----------------------------------------------------------------- */
#define APX_F16   0
#define APX_BF16  1
#define APX_F32   2

/* ---- epilogue flags for apx_linear -------------------------------------------------- */
#define APX_EPI_NONE      0
#define APX_EPI_GELU      1   /* exact erf GELU, applied after the bias                      */
#define APX_EPI_GEGLU     2   /* adjacent column pairs (a, g) -> a * gelu_erf(g); `out` has N/2 columns (diffusers GEGLU with the
                                 projection rows interleaved by the host: row 2i = hidden_i, row 2i+1 = gate_i)                 */
#define APX_EPI_RES16     4   /* flag, OR-ed in: `residual` holds `out_dtype` (16-bit) elements instead of fp32; the GEMM value is
                                 rounded to `out_dtype` before the add, as a separate 16-bit `y + h` would                    */

#define APX_EPI_LNFOLD    8   /* internal to apx_linear_lnfold: LayerNorm of the A rows folded into the GEMM                  */

/* ---- kernel families ---------------------------------------------------------------- */
#define APX_IMPL_TC           0   /* tcgen05 + TMA + TMEM (product path; single site kernel when packed weights are given) */
#define APX_IMPL_SIMT         1   /* plain CUDA-core kernels: device-side cross-check for tests                     */
#define APX_IMPL_TC_UNFUSED   2   /* tcgen05 path as three kernels (to_q GEMM, attention core, to_out GEMM)         */

const char* apx_last_error(void);
int         apx_version(void);            /* ABI version, bumps on any signature change   */
int         apx_device_ok(void);          /* 1 iff the current device is compute 10.x      */
int         apx_set_impl(int impl);       /* APX_IMPL_TC (default), _SIMT or _TC_UNFUSED   */
int         apx_get_impl(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches)     */
uint64_t    apx_launch_count(void);

/*
 * Y[M,N] = act(A[M,K] · W[N,K]^T + bias[N]) + residual[M,N]
 *   A, W      : `in_dtype` (APX_F16 | APX_BF16), row-major, K contiguous, lda/ldw in elements
 *   bias      : fp32[N] or NULL
 *   residual  : fp32[M,N] (ld = ldr) or NULL; may alias `out` when out_dtype == APX_F32; with APX_EPI_RES16 it is 16-bit
 *               (`out_dtype`) and may alias a 16-bit `out`; with APX_EPI_GEGLU it is [M, N/2] like `out`
 *   out       : `out_dtype` (APX_F16 | APX_BF16 | APX_F32), ld = ldo
 * Requirements (tcgen05 path): K % 64 == 0, N % 64 == 0, 16-byte aligned rows.
 */
int apx_linear(const void* A, int64_t lda, const void* W, int64_t ldw,
               const float* bias, const float* residual, int64_t ldr,
               void* out, int64_t ldo,
               int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
               void* stream);

/*
 * Y = act( LayerNorm(A) · W^T + b ) [+ residual]  with the LayerNorm FOLDED into the GEMM (the `norm1` / `norm3` in front of the
 * QKV / GEGLU projections of diffusers' BasicTransformerBlock, and timm Block's norm1 / norm2: one kernel instead of two, and the
 * normalised activations never exist in memory):
 *   Wg  = W * gamma[None, :]  (the caller folds the LayerNorm weight into the projection once per weight version), [N, K] `in_dtype`
 *   c1  = fp32[N], c1[n] = sum_k Wg[n, k]                (of the ROUNDED Wg, so that a constant row normalises to exactly beta W^T + b)
 *   c2  = fp32[N], c2[n] = sum_k beta[k] W[n, k] + b[n]
 *   out = rstd_m (A · Wg^T - mean_m c1) + c2,  mean_m / rstd_m over the K elements of row m of A (biased variance, eps), computed
 *         inside the kernel from the A tiles as they stream through shared memory.
 * Same epilogue flags, dtypes and alignment rules as apx_linear; no split-K (a CTA must see whole rows).
 */
int apx_linear_lnfold(const void* A, int64_t lda, const void* Wg, int64_t ldw, const float* c1, const float* c2, float eps,
                      const float* residual, int64_t ldr, void* out, int64_t ldo,
                      int M, int N, int K, int in_dtype, int out_dtype, int epilogue, void* stream);

/*
 * Segmented multi-head attention on token-major tensors.
 *   O[b,s,h,:] = w0 · softmax_{l<L0}(q·k_l / sqrt(d)) v_l  +  w1 · softmax_{L0<=l<L}(q·k_l / sqrt(d)) v_l
 * With L0 == 0 or L0 == L this is plain softmax attention weighted by w1 / w0.
 *   Q : [B, S, H*d] view, element strides (q_bs, q_rs), head h at column h*d
 *   K, V : [B, L, H*d] views with their own strides; O : [B, S, H*d] view
 * dtype APX_F16 | APX_BF16 for all four.  d % 8 == 0, d <= 96, L <= 544.
 */
int apx_attention(const void* Q, int64_t q_bs, int64_t q_rs,
                  const void* K, int64_t k_bs, int64_t k_rs,
                  const void* V, int64_t v_bs, int64_t v_rs,
                  void* O, int64_t o_bs, int64_t o_rs,
                  int B, int S, int H, int d, int L, int L0, float w0, float w1,
                  int dtype, void* stream);

/*
 * Plain softmax attention over ANY number of keys (flash-style: keys streamed in blocks of 128, running max / sum in registers) —
 * the F.scaled_dot_product_attention call of AttnProcessor2_0 (APadapter/ap_adapter/attention_processor.py:267-274) at the
 * self-attention sites of the UNet (S = L = 1000 / 252 / 64 tokens), no mask:
 *   O[b,s,h,:] = softmax_l(q·k_l / sqrt(d)) v_l
 *   Q, O : [B, S, C] views, K, V : [B, L, C] views, C = H*d, element strides (*_bs per batch element, *_rs per row; multiples of 8),
 *   columns contiguous — the q / k / v column slices of one fused QKV projection are such views.
 * dtype APX_F16 | APX_BF16.  d % 16 == 0, 16 <= d <= 80 (APX_EUNSUPPORTED otherwise; no fallback).
 */
int apx_flash_attention(const void* Q, int64_t q_bs, int64_t q_rs,
                        const void* K, int64_t k_bs, int64_t k_rs,
                        const void* V, int64_t v_bs, int64_t v_rs,
                        void* O, int64_t o_bs, int64_t o_rs,
                        int B, int S, int L, int C, int H, int dtype, void* stream);

/*
 * Context K/V cache of one adapter site (timestep-invariant; build once per generation).
 * ctx [Bc, Lt+La, Dctx] (`dtype`, contiguous); Wk/Wv [C, Dctx] are the UNet's attn2.to_k/to_v
 * (applied to rows < Lt), Wk_ip/Wv_ip [C, Dctx] the adapter's to_k_ip/to_v_ip (rows >= Lt).
 * `cache` receives apx_ctx_kv_bytes(...) bytes in the layout apx_ipattn_fwd consumes.
 */
size_t apx_ctx_kv_bytes(int Bc, int Lt, int La, int C, int H);
int apx_ctx_kv(const void* ctx, const void* Wk, const void* Wv, const void* Wk_ip, const void* Wv_ip,
               void* cache, int Bc, int Lt, int La, int Dctx, int C, int H, int dtype,
               void* workspace, size_t workspace_bytes, void* stream);
size_t apx_ctx_kv_workspace_bytes(int Bc, int Lt, int La, int C);

/*
 * Pre-packed to_q / to_out weights of one adapter site for the single-kernel path (build once per weight version).
 * Wq, Wo [C, C] `dtype` contiguous (attn.to_q.weight, attn.to_out[0].weight).  apx_ipattn_pack_bytes returns 0
 * when the site kernel does not cover (C, H) (then pass wpack = NULL below).
 */
size_t apx_ipattn_pack_bytes(int C, int H);
int apx_ipattn_pack(const void* Wq, const void* Wo, void* wpack, int C, int H, int dtype, void* stream);
/* 1 iff apx_ipattn_fwd runs this (C, H, context) as ONE fused kernel when given a pack */
int apx_ipattn_fused_ok(int C, int H, int Lt, int La);
/* Which kernel apx_ipattn_fwd picks for a call with a pack: 0 = three kernels, 1 = all-heads site kernel (one CTA per 128-query
 * tile: C <= 256), 2 = one-head-per-CTA site kernel (a cluster of H CTAs per tile: any C = H*d, H <= 8), 3 = the all-heads kernel with
 * four softmax warpgroups (one tile per SM: batch >= ~5 clips at the S = 1000 sites, 16 + 128 padded context columns).  The choice
 * follows the shape and the grid size; apx_set_site_variant pins it (measurements / tests): APX_SITE_AUTO, _ALLHEADS, _ONEHEAD. */
#define APX_SITE_AUTO      0
#define APX_SITE_ALLHEADS  1
#define APX_SITE_ONEHEAD   2
#define APX_SITE_ALLHEADS4 3   /* all-heads kernel with four softmax warpgroups (score rows split over two threads): full grids */
int apx_ipattn_site_kernel(int Bc, int S, int C, int H, int Lt, int La);
int apx_set_site_variant(int variant);

/*
 * One adapter site, one denoise step (attention_processor.py:387-459):
 *   Y = ( SDPA(X Wq^T, K_t, V_t) + ap_scale · SDPA(X Wq^T, K_a, V_a) ) Wo^T + bo
 * X, Y [Bc, S, C] contiguous `dtype`; Wq, Wo [C, C] `dtype`; bo fp32[C] or NULL;
 * `cache` from apx_ctx_kv with identical (Bc, Lt, La, C, H); `wpack` from apx_ipattn_pack or NULL.
 * With a pack and a covered shape this is ONE tcgen05 kernel (to_q + both softmax(QK^T)V + ap_scale add + to_out; contexts of up to
 * 256 padded tokens, H <= 8 or H*d_pad <= 256); otherwise three (GEMM, attention core, GEMM) using `workspace` for Q and O.  Wq / Wo may be NULL when the fused
 * kernel is guaranteed (apx_ipattn_fused_ok).
 * workspace: apx_ipattn_workspace_bytes(Bc, S, C) bytes.
 */
size_t apx_ipattn_workspace_bytes(int Bc, int S, int C);
int apx_ipattn_fwd(const void* X, const void* Wq, const void* cache, const void* Wo, const void* wpack, const float* bo,
                   void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream);
/*
 * Same site with the transformer block's residual add folded in (the caller's `attn2(norm2(h)) + h`, diffusers
 * BasicTransformerBlock as used by pipeline_audioldm2.py's UNet):  Y <- Y + site(X), Y holding the residual stream h on entry.
 * The site's value is rounded to `dtype` and added to Y with one more `dtype` rounding, exactly as the caller's elementwise add
 * would (fused kernel: TMA reduce-add of the output boxes; three-kernel form: 16-bit residual operand of the to_out GEMM).
 */
int apx_ipattn_fwd_add(const void* X, const void* Wq, const void* cache, const void* Wo, const void* wpack, const float* bo,
                       void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same site with the transformer block's `norm2` LayerNorm folded into the kernel (diffusers BasicTransformerBlock:
 * `attn2(norm2(h), ctx) + h`, the call sites pipeline/modeling_audioldm2.py:1140-1156, 1317-1333, 1492-1508):
 *   X = the UN-normalised stream h; `wpack` = apx_ipattn_pack of (Wq * gamma[None, :], Wo); ln_c1[n] = sum_k (Wq*gamma)[n, k] (of the rounded
 *   product), ln_c2[n] = sum_k beta[k] Wq[n, k], both fp32[C]; q = rstd (x Wq'^T - mean c1) + c2 with the row statistics taken inside the
 *   kernel from the X tiles it has staged for the projection.  accumulate != 0: Y <- Y + site (Y may alias X: the block's residual stream).
 * Only the all-heads site kernel folds the norm: call it when apx_ipattn_site_kernel(...) == APX_SITE_ALLHEADS, else APX_EUNSUPPORTED.
 */
int apx_ipattn_fwd_ln(const void* X, const void* cache, const void* wpack, const float* bo, const float* ln_c1, const float* ln_c2, float ln_eps,
                      void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype, int accumulate, void* stream);

/*
 * Backward of one adapter site for adapter-only training (train_apadapter_v2.py:892-979: only to_k_ip / to_v_ip train, the
 * UNet is frozen, and the loss must still reach the adapter sites further up, hence dX):
 *   dY, X, dX [Bc, S, C] `dtype`; Wq, Wo [C, C] `dtype` (as in apx_ipattn_fwd); `cache` = the context cache the forward used
 *   (apx_ctx_kv with the CURRENT adapter weights); ctx [Bc, Lt+La, Dctx] `dtype`; dWk_ip, dWv_ip fp32 [C, Dctx], overwritten.
 * Requires La > 0, Lt + La <= 544 (the un-pooled 8 + 512 context included), d % 8 == 0, C % 64 == 0.
 */
size_t apx_ipattn_bwd_workspace_bytes(int Bc, int S, int C, int La);
int apx_ipattn_bwd(const void* dY, const void* X, const void* Wq, const void* Wo, const void* cache, const void* ctx, void* dX,
                   float* dWk_ip, float* dWv_ip, int Bc, int S, int C, int H, int Lt, int La, int Dctx, float ap_scale, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- UNet block neighbours of the adapter sites (pipeline/modeling_audioldm2.py:1140-1156, 1317-1333, 1492-1508:
 *      ResnetBlock2D / Transformer2DModel around every attn2 call site), on channels-last 16-bit activations ------------------- */
/*
 * y = [SiLU]( GroupNorm_groups( x + chan_bias[c] + bc_add[b, c] ) * gamma[c] + beta[c] )
 *   x, y [B, HW, C] `dtype` (NHWC); gamma, beta fp32[C]; chan_bias fp32[C] or NULL (the bias of the convolution that produced x);
 *   bc_add [B, C] `dtype`, rows bc_stride elements apart, or NULL (ResnetBlock2D's time-embedding projection).  Statistics in fp32 over the SUMMED value, two
 *   exact passes; one rounding at the end.  Replaces conv-bias add + temb add + nn.GroupNorm (3 ATen kernels) + F.silu.
 *   Requires (C / groups) % 4 == 0 and HW * (C / groups) * 4 <= 1600 KB (a group is staged in the shared memory of a cluster of <= 8 CTAs).
 */
int apx_groupnorm_nhwc(const void* x, const float* gamma, const float* beta, const float* chan_bias, const void* bc_add, int64_t bc_stride,
                       void* y, int B, int HW, int C, int groups, float eps, int silu, int dtype, void* stream);
/* y = (a + b + bias[c]) * scale over [rows, C] `dtype` (C % 8 == 0); bias fp32[C] or NULL.  ResnetBlock2D's output:
 * (shortcut(x) + conv2(h) + conv biases) / output_scale_factor. */
int apx_add_bias_nhwc(const void* a, const void* b, const float* bias, void* y, int64_t rows, int C, float scale, int dtype, void* stream);
/* y_sum = a + b (rounded to `dtype`, the new residual stream) and y_ln = LayerNorm(y_sum) * gamma + beta, in one kernel; b == NULL (then
 * y_sum == NULL): y_ln = LayerNorm(a).  a, b, y_sum, y_ln [rows, C] `dtype` (C % 8 == 0, C <= 2048); gamma, beta [C] in `w_dtype` (= dtype or
 * APX_F32).  Statistics in fp32 over the rounded sum, two exact passes.  Replaces "hidden_states = attn_output + hidden_states;
 * norm_hidden_states = self.norm2(hidden_states)" of diffusers' BasicTransformerBlock as wired around every attn2 call site by
 * pipeline/modeling_audioldm2.py:1140-1156, 1317-1333, 1492-1508 (third-party leaf). */
int apx_add_layernorm(const void* a, const void* b, const void* gamma, const void* beta, int w_dtype, void* y_sum, void* y_ln, int64_t rows,
                      int C, float eps, int dtype, void* stream);

/* ---- AudioMAE encoder ---------------------------------------------------------------- */
typedef struct apx_vit_block {
    const float* ln1_w; const float* ln1_b;           /* fp32[768]                           */
    const void*  qkv_w; const float* qkv_b;           /* [2304,768] wdtype, fp32[2304]       */
    const void*  proj_w; const float* proj_b;         /* [768,768],  fp32[768]               */
    const float* ln2_w; const float* ln2_b;
    const void*  fc1_w; const float* fc1_b;           /* [3072,768], fp32[3072]              */
    const void*  fc2_w; const float* fc2_b;           /* [768,3072], fp32[768]               */
    /* optional (ABI v3): norm1 / norm2 folded into the qkv / fc1 GEMMs (see apx_linear_lnfold): Wg [N,768] wdtype, c1, c2 fp32[N].
     * All six NULL = separate LayerNorm kernels. */
    const void*  qkv_wg; const float* qkv_c1; const float* qkv_c2;
    const void*  fc1_wg; const float* fc1_c1; const float* fc1_c2;
} apx_vit_block;

typedef struct apx_vit_weights {
    int depth, embed, heads, mlp;                      /* 12, 768, 12, 3072                   */
    int wdtype;                                        /* APX_F16 | APX_BF16 (GEMM operands)  */
    const float* patch_wT;                             /* fp32[256, 768]: Conv2d(1,768,16,16) weight
                                                          viewed [768,256] and TRANSPOSED by the host  */
    const float* patch_b;                              /* fp32[768]                            */
    const float* pos_embed;                            /* fp32[513, 768]                       */
    const float* cls_token;                            /* fp32[768]                            */
    const float* norm_w; const float* norm_b;          /* final LayerNorm                      */
    const apx_vit_block* blocks;                       /* host array of `depth` entries        */
} apx_vit_weights;

size_t apx_audiomae_workspace_bytes(int B);
/* mel fp32 [B,1024,128] -> emb fp32 [B, 512/(tp*fp), 768]; LayerNorm eps = 1e-6. */
int apx_audiomae_fwd(const float* mel, const apx_vit_weights* w_host, float* emb,
                     int B, int time_pool, int freq_pool,
                     void* workspace, size_t workspace_bytes, void* stream);

/* same, stopping before the pooling: tokens fp32 [B, 513, 768] = final-LayerNorm output incl. the CLS row
 * (audio_encoder/models_mae.py:548-570, forward_encoder_no_random_mask_no_average)                      */
int apx_audiomae_encode(const float* mel, const apx_vit_weights* w_host, float* tokens, int B,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Kaldi-compatible log-mel front-end of the audio prompt (audio_encoder/AudioMAE.py:356-394 after the resampling:
 * torchaudio.compliance.kaldi.fbank(htk_compat, hanning, 128 bins, 25 ms / 10 ms, no dither) -> cut / zero-pad to
 * target_len frames -> (x + 4.2677393) / (2 * 4.5689974)).
 *   wave16k fp32[n_samples] at 16 kHz (the global DC offset is removed inside); mel_fbT fp32[257, 128] = the mel filterbank
 *   transposed (host-built, ap_adapter_b200/audiomae.py::kaldi_mel_banks); out fp32[target_len, 128]; scratch >= 4 bytes.
 */
int apx_fbank(const float* wave16k, int n_samples, const float* mel_fbT, float* out, int target_len, float* scratch, void* stream);
/* the same for a batch of equally long clips (the training collate, train_apadapter_v2.py:426-468): wave16k fp32[B, n_samples],
 * out fp32[B, target_len, 128], scratch >= 4*B bytes */
int apx_fbank_batch(const float* wave16k, int B, int n_samples, const float* mel_fbT, float* out, int target_len, float* scratch, void* stream);
/*
 * Resampling in front of the filterbank (audio_encoder/AudioMAE.py:360-366: torchaudio.functional.resample(waveform, sr, 16000);
 * torchaudio's windowed-sinc polyphase resampler is a third-party dependency of the reference).  With orig / new the two rates
 * divided by their gcd, `taps` fp32[new, 2*width + orig] the host-built windowed-sinc table (ap_adapter_b200/audiomae.py::sinc_resample_taps):
 *   out[b, j*new + p] = sum_t taps[p, t] * wave[b, j*orig + t - width]   (wave = 0 outside [0, n_samples)),  out fp32[B, ceil(new*n_samples/orig)]
 */
int apx_resample(const float* wave, int B, int n_samples, const float* taps, int orig, int new_, int width, float* out, void* stream);

/* individual stages (also used by the parity tests) */
int apx_patch_embed(const float* mel, const float* patch_wT, const float* patch_b,
                    const float* pos_embed, const float* cls_token, float* x, int B, void* stream);
int apx_layernorm(const float* x, const float* w, const float* b, void* y, int rows, int cols,
                  float eps, int out_dtype, void* stream);
/* xn: LayerNorm-ed tokens fp32 [B,513,768] -> emb fp32 [B, (64/tp)*(8/fp), 768] */
int apx_pool_tokens(const float* xn, float* emb, int B, int time_pool, int freq_pool, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APX_H_ */
