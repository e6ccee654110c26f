/*
 * istvt_b200 — C ABI of the B200-native ISTVT forward hot path.
 *
 * One entry point per kernel family.  Every function takes raw DEVICE pointers, plain sizes and the
 * CUDA stream to launch on (a `cudaStream_t` passed as `void*`), and returns 0 on success, a positive
 * `cudaError_t` value, or a negative ISTVT_ERR_* code.  No C++ exceptions, no Python / torch objects
 * cross this boundary.  All functions are re-entrant across host threads bound to different devices
 * (the `nn.DataParallel` use at train_CNN.py:185-186): they launch on the caller's current device and
 * the given stream and keep only immutable per-device caches.
 *
 * The reference (Vill-Lab/2023-TIFS-ISTVT) has no FFI of its own — every op on its hot path is an ATen
 * call made from Python.  Each entry below therefore cites the reference Python call site(s) whose
 * arithmetic it replaces (paths relative to the reference root).
 *
 * Layout conventions (differ from the reference's NCHW on purpose):
 *   - images / feature maps: NHWC, channel innermost;
 *   - token sequences: row-major [rows, dim], row = ((b * F + f) * P + p), F = T+1 frames, P = 362;
 *   - `dtype` arguments: ISTVT_BF16 (0) or ISTVT_F32 (1) select the activation element type.
 */
#ifndef ISTVT_B200_H
#define ISTVT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISTVT_OK 0
#define ISTVT_ERR_INVALID_ARG (-1)
#define ISTVT_ERR_NO_DRIVER (-2) /* cuTensorMapEncodeTiled entry point not available */
#define ISTVT_ERR_TMAP (-3)      /* tensor-map encoding rejected the shape/strides */
#define ISTVT_ERR_UNSUPPORTED (-4)

#define ISTVT_BF16 0
#define ISTVT_F32 1

#define ISTVT_ACT_NONE 0
#define ISTVT_ACT_RELU 1
#define ISTVT_ACT_GELU 2 /* exact erf GELU, nn.GELU() default (network/vivit/module.py:28) */

typedef void* istvt_stream_t; /* cudaStream_t */

/* Library ABI version (bumped when a signature changes). */
int istvt_abi_version(void);
/* Human-readable text for a return code of any function below (static storage). */
const char* istvt_error_string(int code);
/* Number of kernel launches issued through this library by the calling process (all threads). */
int64_t istvt_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm(dim) over the last axis, eps inside the sqrt, affine gamma/beta (fp32).
 * Replaces: PreNorm.norm at network/vivit/module.py:18,21 (LN before spatial attention and before
 * the MLP), STTransformer.norm at network/vivit/vivit.py:89,101.
 * x: [rows, dim] (x_dtype), y: [rows, dim] (y_dtype).
 * ------------------------------------------------------------------------------------------- */
int istvt_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                        int y_dtype, int64_t rows, int dim, float eps, istvt_stream_t stream);
/* The same with row pitches ldx / ldy (elements, >= dim, % 4 == 0).  The engine keeps its [rows, 728] bf16 GEMM
 * operands at a pitch of 768 elements: every 128-byte TMA box row of an operand is then ONE 128-byte line instead of
 * straddling two (7 rows of 8 at pitch 728), which is worth 13-25 % on the K = 728 GEMMs (profiles/README.md r6n). */
int istvt_layernorm_fwd_ld(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* y,
                           int y_dtype, int64_t ldy, int64_t rows, int dim, float eps, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm fused with the temporal self-subtract of TemporalResidualAttention.
 * Replaces: PreNorm.norm (module.py:21) followed by module.py:191-194:
 *   xn   = LN(x)                                             -> `xn`   (operand of to_v,  module.py:196)
 *   diff = cat(xn[:, :2], xn[:, 2:] - xn[:, 1:-1], dim=1)    -> `diff` (operand of to_qk, module.py:195)
 * x: fp32 [batch, frames, tokens, dim]; xn/diff: [batch, frames, tokens, dim] (out_dtype).
 * The difference is formed in fp32 before rounding to out_dtype.
 * ------------------------------------------------------------------------------------------- */
int istvt_layernorm_diff_fwd(const float* x, const float* gamma, const float* beta, void* xn, void* diff,
                             int out_dtype, int batch, int frames, int tokens, int dim, float eps,
                             istvt_stream_t stream);
/* The same with row pitches: ldx of x, ld_out of xn AND diff (elements, >= dim, % 4 == 0). */
int istvt_layernorm_diff_fwd_ld(const float* x, int64_t ldx, const float* gamma, const float* beta, void* xn, void* diff,
                                int out_dtype, int64_t ld_out, int batch, int frames, int tokens, int dim, float eps,
                                istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * C[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] ) + residual[m, n]
 * tcgen05 / TMEM tensor-core GEMM, bf16 operands, fp32 accumulation, TMA-fed.
 * Replaces every nn.Linear on the path — to_qk / to_v / to_out (module.py:182-188,195-196,206),
 * to_qkv / to_out (module.py:74-79,83,92), FeedForward.net (module.py:27-31,34) with the residual adds of
 * vivit.py:99-100 — and every 1x1 convolution + BatchNorm(+ReLU) of the Xception entry flow
 * (SeparableConv2d.pointwise xception.py:44,48; Block.skip/skipbn xception.py:57-58,94-96) once BN is
 * folded into W and bias.
 * A: bf16 [m, k] row pitch lda; W: bf16 [n, k] row pitch ldw (nn.Linear weight layout);
 * C: c_dtype [m, n] row pitch ldc; bias: fp32 [n] or NULL; residual: fp32 [m, n] row pitch ldr or NULL
 * (may alias C when c_dtype is F32).  Pitches are in elements; lda, ldw, ldc must be multiples of 8.
 * ------------------------------------------------------------------------------------------- */
int istvt_gemm_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc, int c_dtype,
                   int64_t m, int n, int k, const float* bias, const float* residual, int64_t ldr, int act,
                   istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MLP fusions of the training step (FeedForward, module.py:27-34; train_CNN.py:532 `loss.backward()`), CTA-pair kernel,
 * bf16, n >= 256:
 *   istvt_gemm_act_dual_fwd: c_pre = A W^T + bias (what the backward of the activation needs) AND c_act = act(c_pre)
 *     (the next Linear's operand) from one accumulator — no stand-alone GELU pass in the training forward;
 *   istvt_gemm_dgelu_fwd:    c = (A W^T) o gelu'(pre) — the data gradient of the second Linear multiplied by the
 *     activation's derivative in the epilogue (pre: bf16 [m, n], pitch ld_pre) — no stand-alone GELU backward pass.
 * ------------------------------------------------------------------------------------------- */
int istvt_gemm_act_dual_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c_act, int64_t ldc, void* c_pre,
                            int64_t ldc_pre, int64_t m, int n, int k, const float* bias, int act, istvt_stream_t stream);
int istvt_gemm_dgelu_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc, const void* pre,
                         int64_t ld_pre, int64_t m, int n, int k, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm folded around a pair of GEMMs:  Z = LN(Y) W2^T  with  Y = A W1^T + b1  computed WITHOUT the LN pass.
 * Replaces: PreNorm of the spatial attention, module.py:15-21 on vivit.py:93,99 — `norm(y1)` between the temporal
 * attention's to_out (module.py:185-188,206) and the spatial to_qkv (module.py:74,83).  LN's input is the bf16 GEMM
 * output, so folding is precision-neutral:
 *   LN(y) W^T = rstd (y (gamma o W)^T - mu rowsum(gamma o W)) + W beta
 * istvt_gemm_rowstats_fwd: C = A W^T + bias (bf16), and for every row m and 64-column group g of the output
 *   row_stats[m, g] = {sum, sum of squared deviations from the group mean} (fp32 [m, ceil(n / 64), 2]; every slot is
 *   written exactly once: no atomics, bitwise reproducible; taken from the fp32 values before the bf16 rounding).
 * istvt_ln_stats_finalize: mu_rstd[m] = {mean, 1 / sqrt(var + eps)} over `dim` features from the ceil(dim / 64)
 *   partials of row m, combined with Chan's parallel-variance formula (no E[x^2] - mu^2 cancellation).
 * istvt_gemm_lnfold_fwd: C[m,n] = rstd_m (acc[m,n] - mu_m w_rowsum[n]) + shift[n];
 *   w = bf16(gamma o W) [n, k], w_rowsum = its fp32 row sums, shift = W beta (fp32 [n]), mu_rstd fp32 [m, 2].
 * GEMMs: n >= 256, bf16 output, the usual alignment rules of istvt_gemm_fwd.
 * ------------------------------------------------------------------------------------------- */
int istvt_gemm_rowstats_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc, int64_t m,
                            int n, int k, const float* bias, float* row_stats, istvt_stream_t stream);
int istvt_ln_stats_finalize(const float* partials, float* mu_rstd, int64_t m, int dim, float eps, istvt_stream_t stream);
int istvt_gemm_lnfold_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc, int64_t m,
                          int n, int k, const float* mu_rstd, const float* w_rowsum, const float* shift,
                          istvt_stream_t stream);

/* Same contract in fp32 (SIMT FFMA kernel; the 1e-4 validation mode, not the performance path).
 * A, W, C, residual all fp32. */
int istvt_gemm_f32_fwd(const float* a, int64_t lda, const float* w, int64_t ldw, float* c, int64_t ldc, int64_t m,
                       int n, int k, const float* bias, const float* residual, int64_t ldr, int act,
                       istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense 3x3 stride-1 pad-0 convolution as an implicit GEMM on the tensor cores, + folded BN + act.
 * Replaces: conv2 + bn2 + relu, xception.py:122-123,198-200.
 * x: NHWC [n, h, w, cin] (dtype); wt: [cout, 3, 3, cin] (dtype, BN scale folded in); bias fp32 [cout];
 * y: NHWC [n, h-2, w-2, cout] (dtype).  bf16 runs on tcgen05, fp32 on the SIMT kernel.
 * ------------------------------------------------------------------------------------------- */
int istvt_conv3x3_fwd(const void* x, const void* wt, const float* bias, void* y, int dtype, int n, int h, int w,
                      int cin, int cout, int act, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The same convolution for conv2's shape (bf16, 32 -> 64 channels) on PIXEL PAIRS: two adjacent NHWC pixels are one
 * 128-byte row of a [n h w / 2, 64] matrix, so the implicit GEMM fetches whole 128-byte lines (6 or 7 k-blocks of 64
 * for two output pixels instead of 9 x 32 for one) against a rearranged, zero-padded weight matrix.
 * Replaces: conv2 + bn2 + relu, xception.py:122-123,198-200 (same results as istvt_conv3x3_fwd up to fp32 summation
 * order).  n * h * w must be even.
 * istvt_conv3x3_pair_pack: wt bf16 [64, 3, 3, 32] (BN scale folded) -> wpair bf16 [128, taps * 64], taps = 7 when the
 *   input width w_in is odd and 6 when it is even (the layout depends on w_in's parity only).
 * istvt_conv3x3_pair_fwd:  x NHWC bf16 [n, h, w, 32]; bias fp32 [64]; y NHWC bf16 [n, h-2, w-2, 64]; act none / ReLU.
 * ------------------------------------------------------------------------------------------- */
int istvt_conv3x3_pair_pack(const void* wt, void* wpair, int w_in, istvt_stream_t stream);
int istvt_conv3x3_pair_fwd(const void* x, const void* wpair, const float* bias, void* y, int n, int h, int w, int act,
                           istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stem: 3x3 stride-2 pad-0 convolution 3 -> cout on an NCHW fp32 clip + folded BN + ReLU, NHWC out.
 * Replaces: conv1 + bn1 + relu, xception.py:118-120,194-196 (input rearrange vivit.py:204 is a view).
 * x: fp32 NCHW [n, 3, h, w]; wt: fp32 [cout, 3, 3, 3] (BN scale folded); bias fp32 [cout];
 * y: NHWC [n, (h-3)/2+1, (w-3)/2+1, cout] (dtype).  cout must be 32.
 * ------------------------------------------------------------------------------------------- */
int istvt_conv_stem_fwd(const float* x, const float* wt, const float* bias, void* y, int dtype, int n, int h,
                        int w, int cout, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stem on decoded frames: the same convolution reading uint8 NHWC [n, h, w, 3] (the layout video decoders and
 * the reference's absent `dataset.transform` input, train_CNN.py:18-21,172-173, start from).  The per-channel
 * input normalisation (x / 255 - mean) / std — xception.py:12-13 documents mean = std = 0.5 — is affine and conv1
 * has no padding, so the caller folds it into `wt` / `bias` together with the BN scale; the kernel converts bytes.
 * cuts the host->device copy and the stem's HBM read 4x.
 * x: uint8 [n, h, w, 3]; wt: fp32 [cout, 3, 3, 3]; bias fp32 [cout]; y: NHWC [n, (h-3)/2+1, (w-3)/2+1, cout].
 * ------------------------------------------------------------------------------------------- */
int istvt_conv_stem_u8_fwd(const uint8_t* x, const float* wt, const float* bias, void* y, int dtype, int n, int h,
                           int w, int cout, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Depthwise 3x3 stride-1 pad-1 convolution, optional ReLU applied to the input on load.
 * Replaces: SeparableConv2d.conv1 (xception.py:43,47) and the ReLU that precedes it inside Block.rep
 * (xception.py:66-67,72-73,82-85).
 * x, y: NHWC [n, h, w, c] (dtype); wt: fp32 [3, 3, c] (tap-major, channel innermost). c % 8 == 0.
 * ------------------------------------------------------------------------------------------- */
int istvt_dwconv3x3_fwd(const void* x, const float* wt, void* y, int dtype, int n, int h, int w, int c,
                        int relu_in, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SeparableConv2d + folded BatchNorm (+ ReLU) in ONE kernel: the depthwise result is produced straight into the
 * pointwise tensor-core GEMM's shared-memory A tiles and never goes to HBM.
 * Replaces: SeparableConv2d.forward (xception.py:46-49: conv1 = depthwise 3x3 pad 1 groups=C, pointwise = 1x1) together
 * with the BatchNorm2d after it (xception.py:69-75, eval mode: scale folded into pw, shift = bias) and the ReLU that
 * precedes the next separable conv; relu_in = the ReLU in front of THIS one (xception.py:82-85).
 *   y = act( pw . depthwise3x3( relu_in ? relu(x) : x ) + bias )
 * x: bf16 NHWC [n, h, w, c]; dw: fp32 [3, 3, c]; pw: bf16 [n_out, c], row pitch ld_pw; bias: fp32 [n_out];
 * y: bf16 NHWC [n, h, w, n_out]; act: ISTVT_ACT_NONE or ISTVT_ACT_RELU.
 * Supported: c in {64, 128, 192, 256}, n_out in {64, 128, 192, 256}, weights + rings within 227 KB of shared memory
 * (every separable conv of entry-flow blocks 1 and 2 except 256 -> 256); otherwise ISTVT_ERR_UNSUPPORTED and the
 * caller runs istvt_dwconv3x3_fwd + istvt_gemm_fwd.
 * ------------------------------------------------------------------------------------------- */
int istvt_sepconv_fused_fwd(const void* x, const float* dw, const void* pw, int64_t ld_pw, const float* bias, void* y,
                            int n, int h, int w, int c, int n_out, int relu_in, int act, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Pixel subsampling by 2 in both axes (rows/cols 0, 2, 4, ...): the gather of a stride-2 1x1 convolution.
 * Replaces the stride of Block.skip (xception.py:57, 94); the 1x1 itself is istvt_gemm_fwd.
 * x: NHWC [n, h, w, c]; y: NHWC [n, (h-1)/2+1, (w-1)/2+1, c].
 * ------------------------------------------------------------------------------------------- */
int istvt_subsample2_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MaxPool2d(3, stride 2, pad 1) of the block body + residual add of the skip branch.
 * Replaces: nn.MaxPool2d at xception.py:87-88 and `x += skip` at xception.py:100.
 * x: NHWC [n, h, w, c]; skip: NHWC [n, ho, wo, c], ho = (h-1)/2+1; y: NHWC [n, ho, wo, c].
 * ------------------------------------------------------------------------------------------- */
int istvt_pool_add_fwd(const void* x, const void* skip, void* y, int dtype, int n, int h, int w, int c,
                       istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Block-3 variant of pool+add that writes straight into the transformer token buffer and adds the
 * positional embedding, so the NCHW->tokens permute copy of the reference disappears.
 * Replaces: xception.py:87-88,100 + Rearrange 'b t c h w -> b t (h w) c' (vivit.py:115,133) + the patch
 * part of `x += pos_embedding` (vivit.py:138).
 * x: NHWC [batch*t, h, w, c]; skip: NHWC [batch*t, ho, wo, c]; pos_emb: fp32 [t, ho*wo+1, c];
 * tokens: fp32 [batch, t+1, ho*wo+1, c]; element (b, f+1, 1+p, :) = pool(x)+skip+pos_emb[f, 1+p, :].
 * ------------------------------------------------------------------------------------------- */
int istvt_pool_add_tokens_fwd(const void* x, const void* skip, const float* pos_emb, float* tokens, int dtype,
                              int batch, int t, int h, int w, int c, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Class-token fill of the token buffer.
 * Replaces: vivit.py:136-140 — tokens[b, 0, p, :] = temporal_token (no positional embedding, all p);
 * tokens[b, f+1, 0, :] = space_token + pos_emb[f, 0, :].
 * ------------------------------------------------------------------------------------------- */
int istvt_token_fill_fwd(float* tokens, const float* space_token, const float* temporal_token,
                         const float* pos_emb, int batch, int t, int tokens_per_frame, int dim,
                         istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Temporal self-attention across frames at each token position (F x F per (clip, head, position)).
 * Replaces: module.py:197-205 (the rearranges, einsum, softmax, einsum, rearrange).
 * qk: [batch*frames*tokens, 2*heads*64] (q columns then k columns, head-major); v: [rows, heads*64];
 * out: [rows, heads*64]; probs (optional, may be NULL): fp32 [batch, heads, tokens, frames, frames].
 * q/k/v are read in place: consecutive frames of one position are `tokens` rows apart.
 * bf16: one warp per (clip, position, head) on mma.sync, registers only — frames <= 8 (T = 6) in one m16 tile,
 * frames <= 48 (the long-clip configuration, T = 32) with the frame axis tiled; fp32 and frames > 48: SIMT kernel.
 * ------------------------------------------------------------------------------------------- */
int istvt_attn_temporal_fwd(const void* qk, const void* v, void* out, float* probs, int dtype, int batch,
                            int frames, int tokens, int heads, float scale, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Spatial self-attention over the tokens of each frame (tokens x tokens per (clip, head, frame)).
 * Replaces: module.py:84-91.  bf16: tcgen05 QK^T and PV with the softmax fused in between, q/k/v tiles
 * fetched by TMA straight from the packed [rows, 3*heads*64] projection output (no permute copies).
 * qkv: [batch_frames*tokens, 3*heads*64]; out: [rows, heads*64];
 * probs (optional, may be NULL): fp32 [batch_frames, heads, tokens, tokens], laid out so that a
 * [batch, frames] split of batch_frames reproduces the reference's [b, h, t, hw, hw] after a transpose.
 * dim_head is fixed at 64; tokens <= 384.  fp32 dtype runs the SIMT validation kernel.
 * ------------------------------------------------------------------------------------------- */
int istvt_attn_spatial_fwd(const void* qkv, void* out, float* probs, int dtype, int batch_frames, int tokens,
                           int heads, float scale, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Joint self-attention over all tokens of a sequence, ANY sequence length (the
 * ablation transformers).  Replaces: module.py:53-63 (`Attention.forward`: chunk, rearranges, einsum, softmax,
 * einsum, rearrange), as used by `Transformer` (vivit.py:10-25) inside `ViViT` (vivit.py:29-81) and `VanillaTr`
 * (vivit.py:150-191; 6*361+1 = 2167 tokens per clip).
 * bf16: tcgen05 QK^T / PV with the keys streamed in blocks of 128 and an online softmax (the score matrix of a long
 * sequence does not fit TMEM); q/k/v tiles fetched by TMA from the packed projection output, no permute copies.
 * qkv: [batch*tokens, 3*heads*64] (q | k | v columns, head-major); out: [batch*tokens, heads*64].
 * dim_head is fixed at 64; scale > 0.  fp32 dtype runs the SIMT validation kernel.
 * ------------------------------------------------------------------------------------------- */
int istvt_attn_joint_fwd(const void* qkv, void* out, int dtype, int batch, int tokens, int heads, float scale,
                         istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Token-sequence assembly of the ablation transformers: class token in slot 0, the n patch rows behind it, positional
 * embedding added, in one pass.  Replaces the repeat / torch.cat / `x += pos_embedding` triples at vivit.py:64-66
 * (ViViT, per frame: pos_period = T), vivit.py:73-74 (ViViT, per clip: pos = NULL) and vivit.py:183-185 (VanillaTr,
 * per clip: pos_period = 1).
 * tokens[s, 0, :] = cls + pos[s % pos_period, 0, :];  tokens[s, 1+i, :] = src[s*n + i, :] + pos[s % pos_period, 1+i, :].
 * src: [sequences*n, dim] (src_dtype bf16 or fp32); cls: fp32 [dim]; pos: fp32 [pos_period, n+1, dim] or NULL;
 * tokens: fp32 [sequences, n+1, dim].  dim % 4 == 0.
 * ------------------------------------------------------------------------------------------- */
int istvt_token_build_fwd(const void* src, int src_dtype, const float* cls, const float* pos, float* tokens,
                          int sequences, int n, int dim, int pos_period, istvt_stream_t stream);

/* Mean over the n token rows of every sequence: out[s, :] = mean_r x[s, r, :] (fp32 in, fp32 out; dim % 4 == 0).
 * Replaces: vivit.py:79 (`x.mean(dim = 1)`, ViViT with pool = 'mean'). */
int istvt_mean_rows_fwd(const float* x, float* out, int sequences, int n, int dim, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-frame Xception baseline (model_selection('xception'), train_CNN.py:924-929).
 * The middle flow (blocks 4-11), block 12 and conv3 / conv4 reuse istvt_dwconv3x3_fwd, istvt_gemm_fwd (1x1 + folded
 * BN + ReLU), istvt_subsample2_fwd and istvt_pool_add_fwd; these two entries are the only additions.
 *
 * istvt_add_fwd: y = a + b elementwise (count % 8 == 0) — the identity-skip residual `x += inp` of a Block without
 * skip convolution, xception.py:92-101.
 * istvt_pool_linear_fwd: logits, xception.py:208-221 — ReLU, adaptive_avg_pool2d(1, 1), last_linear (Dropout is
 * the identity in eval mode).  x: NHWC [n, hw, c] (dtype); w: fp32 [ncls, c]; bias fp32 [ncls] or NULL;
 * out: fp32 [n, ncls].
 * ------------------------------------------------------------------------------------------- */
int istvt_add_fwd(const void* a, const void* b, void* y, int dtype, int64_t count, istvt_stream_t stream);
int istvt_pool_linear_fwd(const void* x, int dtype, const float* w, const float* bias, float* out, int n, int hw,
                          int c, int ncls, int relu, istvt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Classification head on token (0, 0) of every clip: final transformer LayerNorm (row-wise, so only
 * that row is needed), mlp_head LayerNorm, Linear(dim -> 1).
 * Replaces: vivit.py:101 (restricted to the rows read afterwards), vivit.py:144-148.
 * tokens: fp32 [batch, rows_per_clip, dim]; logits: fp32 [batch].
 * ------------------------------------------------------------------------------------------- */
int istvt_head_fwd(const float* tokens, int64_t rows_per_clip, const float* norm_g, const float* norm_b,
                   const float* head_g, const float* head_b, const float* head_w, const float* head_bias,
                   float* logits, int batch, int dim, float eps, istvt_stream_t stream);


/* =============================================================================================
 * Training step (train_CNN.py:513-533: zero_grad, forward, BCE-with-logits, loss.backward(), AdamW.step()).
 * The reference gets its backward from torch autograd over the modules cited above; each entry below is
 * the hand-written gradient of the named forward.  bf16 activations / gradients, fp32 parameters, fp32
 * parameter gradients (accumulated: callers zero the gradient buffers once per step).
 * ============================================================================================= */

/* Spatial attention forward that also saves the per-row log-sum-exp (log2 domain) for the backward.
 * lse: fp32 [batch_frames, heads, tokens].  Same kernel as istvt_attn_spatial_fwd (bf16). */
int istvt_attn_spatial_fwd_lse(const void* qkv, void* out, float* lse, int batch_frames, int tokens, int heads,
                               float scale, istvt_stream_t stream);

/* Backward of module.py:84-91.  qkv / dqkv: bf16 [rows, 3*heads*64]; o (forward output) / dout: bf16
 * [rows, heads*64]; dq_scratch: fp32 [rows, heads*64] workspace. */
int istvt_attn_spatial_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                           float* dq_scratch, int batch_frames, int tokens, int heads, float scale,
                           istvt_stream_t stream);

/* Backward of module.py:197-205 (frames <= 48; ISTVT_ERR_UNSUPPORTED beyond).  dqk: bf16 [rows, 2*heads*64], dv: bf16
 * [rows, heads*64].  frames <= 8: registers only; 9..48: operands staged once in shared memory, ldmatrix fragments. */
int istvt_attn_temporal_bwd(const void* qk, const void* v, const void* dout, void* dqk, void* dv, int batch,
                            int frames, int tokens, int heads, float scale, istvt_stream_t stream);

/* LayerNorm backward (module.py:18,21; vivit.py:89,101).  dy: bf16 [rows, dim]; x: the forward input (x_dtype).
 * dy2 (optional): gradient w.r.t. the self-subtract output `diff` of istvt_layernorm_diff_fwd; the backward of
 * module.py:192 is then fused in: dy_total[f] = dy[f] + dy2[f] - dy2[f+1] (the last term for frames 1..F-2).
 * Exactly one of g_accum / dx_out is non-NULL:
 *   g_accum: fp32 [rows, dim] residual-stream gradient, += dx; g_bf16 (optional) receives a bf16 copy of it;
 *   dx_out : bf16 [rows, dim] = dx.
 * dgamma / dbeta: fp32 [dim], accumulated.  dim % 4 == 0, dim <= 768. */
int istvt_layernorm_bwd(const void* dy, const void* dy2, int frames, int tokens_per_frame, const void* x, int x_dtype,
                        const float* gamma, float* g_accum, void* g_bf16, void* dx_out, float* dgamma, float* dbeta,
                        int64_t rows, int dim, float eps, istvt_stream_t stream);
/* The same with row pitches (elements, >= dim, % 4 == 0): ld_dy of dy AND dy2, ld_x of x, ld_g of g_accum, ld_gb of
 * g_bf16, ld_dx of dx_out (the pitches of absent operands are ignored). */
int istvt_layernorm_bwd_ld(const void* dy, const void* dy2, int64_t ld_dy, int frames, int tokens_per_frame,
                           const void* x, int x_dtype, int64_t ld_x, const float* gamma, float* g_accum, int64_t ld_g,
                           void* g_bf16, int64_t ld_gb, void* dx_out, int64_t ld_dx, float* dgamma, float* dbeta,
                           float* out_colsum, int64_t rows, int dim, float eps, istvt_stream_t stream);
/* out_colsum (optional, fp32 [dim], +=): column sums of the rows this call produces (the updated g_accum, or dx_out) —
 * the bias gradient of the nn.Linear whose output gradient they are (net.3 / spatial to_out / temporal to_out). */

/* exact-erf GELU (module.py:28) on bf16, elementwise: forward (training keeps the pre-activation) and backward. */
int istvt_gelu_fwd(const void* x, void* y, int64_t n, istvt_stream_t stream);
int istvt_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, istvt_stream_t stream);
/* The same on a [rows, cols] matrix, also accumulating colsum[c] (fp32, +=) = sum_r dx[r, c]: the bias gradient of
 * FeedForward.net[0] (module.py:27) without a second pass over dx.  cols % 8 == 0, cols <= 8192. */
int istvt_gelu_bwd_colsum(const void* dy, const void* x, void* dx, float* colsum, int64_t rows, int cols,
                          istvt_stream_t stream);

/* fp32 -> bf16 cast (gradient of the fp32 residual stream as a GEMM operand). */
int istvt_cast_f32_bf16(const float* x, void* y, int64_t n, istvt_stream_t stream);
/* The same for a [rows, cols] matrix with row pitches (elements; cols, ldx, ldy % 4 == 0): bf16 GEMM operands of the
 * training step at the 128-byte aligned pitch (see istvt_layernorm_fwd_ld). */
int istvt_cast_f32_bf16_rows(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols,
                             istvt_stream_t stream);

/* out[c, m] = in[m, c] (bf16; out row pitch ldo >= m, multiple of 8, pad zero-filled); colsum (optional, fp32 [c])
 * += column sums of `in` — the bias gradient of the nn.Linear whose output gradient is being transposed.
 * Produces the K-major operands of the weight-gradient GEMMs. */
int istvt_transpose_colsum(const void* in, void* out, float* colsum, int64_t m, int c, int64_t ldo,
                           istvt_stream_t stream);

/* C[m, n] += sum_k A[m, k] * W[n, k], bf16 operands, fp32 C, split-K over CTA pairs with red.global accumulation:
 * dW = dY^T X for every nn.Linear / 1x1 convolution (A = dY^T [out_features, rows], W = X^T [in_features, rows]). */
int istvt_gemm_splitk_accum(const void* a, int64_t lda, const void* w, int64_t ldw, float* c, int64_t ldc, int64_t m,
                            int n, int64_t k, istvt_stream_t stream);

/* Head backward (vivit.py:101 on token (0,0), vivit.py:144-148): writes g[b, 0, :] (the caller zero-fills g) and
 * accumulates the gradients of transformer.norm, mlp_head.0 (LayerNorm) and mlp_head.1 (Linear). */
int istvt_head_bwd(const float* tokens, int64_t rows_per_clip, const float* dlogits, const float* norm_g,
                   const float* norm_b, const float* head_g, const float* head_b, const float* head_w, float* g,
                   float* d_norm_g, float* d_norm_b, float* d_head_g, float* d_head_b, float* d_head_w,
                   float* d_head_bias, int batch, int dim, float eps, istvt_stream_t stream);

/* Token-build backward (vivit.py:136-140): g fp32 [batch, t+1, tokens_per_frame, dim] ->
 * d_pos [t, tokens_per_frame, dim], d_space [dim], d_temporal [dim] (all accumulated). */
int istvt_token_bwd(const float* g, float* d_pos, float* d_space, float* d_temporal, int batch, int t,
                    int tokens_per_frame, int dim, istvt_stream_t stream);

/* torch.optim.AdamW step (train_CNN.py:199) over flat fp32 buffers; grad_scale multiplies the gradient first
 * (1 / world_size after an all-reduce SUM).  n % 4 == 0. */
int istvt_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                     istvt_stream_t stream);

/* ---- Xception entry flow in training mode (BatchNorm batch statistics; xception.py:52-101,193-206) ---- */

/* conv1 without the ReLU (xception.py:194): BatchNorm with batch statistics follows as separate kernels. */
int istvt_conv_stem_raw_fwd(const float* x, const float* wt, const float* bias, void* y, int dtype, int n, int h,
                            int w, int cout, istvt_stream_t stream);

/* nn.BatchNorm2d in training mode (xception.py:58,69,75,119,123) on x: bf16 [m, c] (NHWC flattened):
 *   stats    : sum[c] += sum_m x, sumsq[c] += sum_m x^2 (fp32, caller zeroes them)
 *   finalize : mean, rstd, scale = gamma*rstd, shift = beta - mean*scale; running_mean/var update (momentum,
 *              unbiased variance) when the running pointers are non-NULL
 *   apply    : y = x*scale + shift (+ReLU), bf16 */
int istvt_bn_stats_fwd(const void* x, float* sum, float* sumsq, int64_t m, int c, istvt_stream_t stream);
int istvt_bn_finalize_fwd(const float* sum, const float* sumsq, const float* gamma, const float* beta, float* scale,
                          float* shift, float* mean, float* rstd, float* running_mean, float* running_var, int64_t m,
                          int c, float eps, float momentum, istvt_stream_t stream);
int istvt_bn_apply_fwd(const void* x, const float* scale, const float* shift, void* y, int64_t m, int c, int relu,
                       istvt_stream_t stream);
/* Backward of BatchNorm(+ReLU): dgamma/dbeta (fp32 [c], must be zero on entry) and dx (bf16). */
int istvt_bn_bwd(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                 const float* rstd, float* dgamma, float* dbeta, void* dx, int64_t m, int c, int relu,
                 istvt_stream_t stream);

/* MaxPool2d(3,2,1) + skip add (xception.py:87-88,100) that also records each window's arg-max tap (uint8
 * [n, ho, wo, c]); exactly one of y (bf16 NHWC) / tokens (fp32 token buffer, + pos_emb, as
 * istvt_pool_add_tokens_fwd) is non-NULL. */
int istvt_pool_add_idx_fwd(const void* x, const void* skip, void* y, const float* pos_emb, float* tokens,
                           void* argmax, int n, int t_frames, int h, int w, int c, istvt_stream_t stream);
/* Max-pool backward (gather over the <= 4 windows containing each input pixel). dy: bf16 [n, ho, wo, c]. */
int istvt_pool_bwd(const void* dy, const void* argmax, void* dx, int n, int h, int w, int c, istvt_stream_t stream);
/* bf16 NHWC [batch*t, 19, 19, c] gradient of the block-3 output from the fp32 token gradient (vivit.py:133-138). */
int istvt_token_grad_gather(const float* g, void* d_out, int batch, int t, int tokens_per_frame, int c,
                            istvt_stream_t stream);
/* Depthwise 3x3 weight gradient, dw: fp32 [3, 3, c] accumulated (xception.py:43,47). */
int istvt_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int c, int relu_in,
                          istvt_stream_t stream);
/* Gradient of a Block's input: main branch (masked by the leading ReLU, xception.py:82-85) + scatter of the
 * stride-2 skip branch (xception.py:94). */
int istvt_block_input_grad(const void* d_main, const void* x_in, const void* d_skip, void* dx, int n, int h, int w,
                           int c, int relu_in, istvt_stream_t stream);
/* K-major im2col operands of the conv2 / conv1 weight-gradient GEMMs (xception.py:118,122). */
int istvt_im2col_t(const void* x, void* out, int n, int h, int w, int cin, int64_t ldo, istvt_stream_t stream);
int istvt_im2col_t_stem(const float* x, void* out, int n, int h, int w, int64_t ldo, istvt_stream_t stream);

/* ---- Relevance propagation (visualize_rel.py:206,257-262; the reference's implementation, `tfe...LRP`, is NOT in its
 * tree — this is the published gradient-weighted attention rollout, see DESIGN.md "relevance pass") ---- */

/* Attention backward kernels that also accumulate cam += relu(dA o A) / heads (A = attention probabilities,
 * dA = gradient of the logit w.r.t. them).  cam is fp32 and zero-filled by the caller:
 * spatial [batch_frames, tokens, tokens], temporal [batch, tokens, frames, frames]. */
int istvt_attn_spatial_bwd_cam(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                               float* dq_scratch, float* cam, int batch_frames, int tokens, int heads, float scale,
                               istvt_stream_t stream);
int istvt_attn_temporal_bwd_cam(const void* qk, const void* v, const void* dout, void* dqk, void* dv, float* cam,
                                int batch, int frames, int tokens, int heads, float scale, istvt_stream_t stream);
/* One rollout step on the class-token row: v[n, :] <- v[n, :] (I + cmat[n]), cmat fp32 [n, len, len]. */
int istvt_rollout_row(float* v, const float* cmat, int64_t n, int len, istvt_stream_t stream);

/* Strided row gather (data movement only): dst [n_outer, rows, row_bytes] contiguous <- src + o*outer_stride +
 * r*row_stride.  Used by the pruned last transformer layer: after layer 12 only token (0,0) of every clip is read
 * (vivit.py:144-148), so its spatial attention / MLP run on the frame-0 rows / the class-token row only. */
int istvt_gather_rows(const void* src, void* dst, int64_t n_outer, int64_t outer_stride_bytes, int64_t rows,
                      int64_t row_stride_bytes, int64_t row_bytes, istvt_stream_t stream);

/* Weight gradient WITHOUT transposed operand copies: dW[n_out, k_in] (fp32, +=) = dY[rows, n_out]^T X[rows, k_in].
 * Both bf16 operands are read in place as MN-major tcgen05 operand tiles (the token rows are the contraction
 * dimension); split-K over CTA pairs with red.global accumulation.  rows > 64.  n_out, k_in, pitches % 8 == 0. */
int istvt_gemm_wgrad_accum(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, float* dw, int64_t ld_dw,
                           int64_t rows, int n_out, int k_in, istvt_stream_t stream);
/* colsum[c] (fp32, +=) = sum_m x[m, c] (bf16): the bias gradient from an output gradient. */
int istvt_colsum(const void* x, float* colsum, int64_t m, int c, istvt_stream_t stream);
/* The same with a row pitch ld (elements, >= c, % 8 == 0). */
int istvt_colsum_ld(const void* x, int64_t ld, float* colsum, int64_t m, int c, istvt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ISTVT_B200_H */
