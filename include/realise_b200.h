/*
 * realise_b200.h — C ABI of librealise_b200.so: the sm_100a kernels behind
 * SpellBertPho2ResArch3.forward (ReaLiSe multimodal hot path).
 *
 * The reference (DaDaMrX/ReaLiSe) has no native layer; its "FFI" for this path is the set of
 * torch.nn calls made by src/models.py:806-870, src/char_cnn.py:9-55 and
 * transformers/modeling_bert.py:155-745.  Each entry point below names the reference call
 * site(s) it replaces.  Conventions:
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer owned by the caller,
 *     the library never allocates, frees or retains device memory;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no host sync;
 *   - return 0 on success, negative RL_E* for host-side argument errors, positive = cudaError_t;
 *     rl_last_error() returns a thread-local message for the last non-zero return;
 *   - bf16 = __nv_bfloat16 storage, f32 = float.  "K-major" = row-major with K contiguous.
 *   - the library holds NO mutable process-wide state: the 16-bit storage format of a tensor (bf16 or IEEE fp16) and
 *     the optional device-resident dropout step counter are per-call arguments (`*_dtype`, `drop_counter`); the only
 *     statics are one-time cudaFuncSetAttribute flags, cached occupancy queries and the thread-local error string.
 *
 * 16-bit formats: every "bf16" tensor below may instead be IEEE fp16 when the call says so (RL_DT_F16): same
 * tensor-core rate, three more mantissa bits.  The two operands of one MMA must share ONE format: a tcgen05.mma
 * kind::f16 whose descriptor mixes bf16 and fp16 raises "illegal instruction" on B200 (probed: tools/probe_mixed_mma.py).
 *
 * Dropout: masks are pure functions of (seed, site, element).  With a non-NULL `drop_counter` (device pointer) the
 * kernel uses seed + *drop_counter, read at run time: a captured CUDA graph of the whole train step draws fresh masks
 * per replay by bumping the counter on the stream.
 */
#ifndef REALISE_B200_H
#define REALISE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define RL_API extern "C" __attribute__((visibility("default")))
#else
#define RL_API
#endif

#define RL_OK 0
#define RL_EINVAL (-1)   /* bad shape / null pointer / unsupported size */
#define RL_EALIGN (-2)   /* pointer or stride alignment */
#define RL_EDRIVER (-3)  /* driver entry point / tensor-map encode failure */

/* ---- library ------------------------------------------------------------------------------ */
RL_API int rl_version(void);
RL_API const char* rl_last_error(void);

/* ---- tcgen05 GEMM with fused epilogue -------------------------------------------------------
 * out[m, n] = act( (sum_k A[m,k] * B[n,k]) * scale[n] + bias[n] + res[m,n] )
 * Replaces: every nn.Linear on the path — BertSelfAttention q/k/v (modeling_bert.py:221-232,
 * fused to one [2304,768] weight), BertSelfOutput.dense (:274), BertIntermediate.dense + gelu
 * (:327-328, :125-131), BertOutput.dense (:340), nn.GRU hidden projection (src/models.py:825),
 * classifier (src/models.py:859) — and, in conv mode, every nn.Conv2d + eval-mode BatchNorm2d
 * (+ReLU, + residual add) of CharResNet blocks (src/char_cnn.py:15-32) as an im2col-free
 * implicit GEMM whose A tiles are fetched tap by tap with 5-D TMA boxes.
 */
enum {
  RL_ACT_NONE = 0, RL_ACT_GELU = 1, RL_ACT_RELU = 2, RL_ACT_TANH = 3,
  RL_ACT_GELU_GRAD = 4, /* out = (acc*scale+bias) * gelu'(res): data gradient through GELU, res = saved pre-activation */
  RL_ACT_GELU_SAVE = 5  /* out = gelu(pre), out2 = pre (bf16): training forward keeps the pre-activation */
};
enum { RL_DT_BF16 = 0, RL_DT_F32 = 1, RL_DT_F16 = 2 };

typedef struct rl_gemm_desc {
  const void* a; /* bf16.  a_mode 0: [M, K] row-major, row stride lda (elements).
                    a_mode 1: activation tensor [NIMG][P][H][W][C] (C contiguous) */
  const void* b; /* bf16 [N, K] K-major, row stride ldb.  conv: K = ntaps * C, tap-major */
  int64_t M, N, K;
  int64_t lda, ldb;
  /* conv addressing; the GEMM row m = (img, oh, ow) with oh < H, ow < W */
  int32_t a_mode;
  int32_t conv_C, conv_W, conv_H, conv_P, conv_NIMG;
  int32_t ntaps;
  int8_t tap_dw[12], tap_dh[12], tap_plane[12];
  /* epilogue */
  void* out;      /* [M, N] (or remapped rows), dtype out_dtype, row stride ldo */
  int64_t ldo;
  int32_t out_dtype;
  void* out2;     /* optional bf16 copy of the result, row stride ldo2 (may be NULL) */
  int64_t ldo2;
  const float* scale; /* [N] or NULL */
  const float* bias;  /* [N] or NULL */
  const void* res;    /* [M, N] residual or NULL, dtype res_dtype, row stride ldr */
  int64_t ldr;
  int32_t res_dtype;
  int32_t act;
  int32_t out_remap; /* 0: row m -> m.  1: parity-split rows for a following stride-2 conv:
                        [img][oh&1][ow&1][oh/2][ow/2] */
  int32_t remap_plane; /* out_remap 2: GEMM row (img, h, w) -> row ((img*4 + remap_plane)*H + h)*W + w: one parity plane of
                          the data gradient of a stride-2 conv */
  int32_t conv_Cuse;   /* conv mode: channels [0, conv_Cuse) of every tap are the GEMM's K (0 = all conv_C) */
  int32_t split_k;     /* 0: off.  > 0: that many K ranges, -1: auto.  out (f32, no epilogue operands) is ACCUMULATED
                          INTO: out += A*B, partial sums of the K ranges meeting through f32 atomics, so the caller
                          zero-initialises it once per step (weight gradients: few output tiles, K = tokens or
                          pixels; also sums a gradient over time steps / shared weights for free) */
  int32_t a_major;   /* 0: A stored [M, K] (K-major).  1: A stored [K, M] with row stride lda (MN-major), e.g.
                        A = dY^T for a weight gradient straight from dY [tokens, out] */
  float drop_p;        /* training: dropout on (acc*scale+bias) BEFORE the residual add (BertSelfOutput/BertOutput, */
  uint32_t drop_site;  /* modeling_bert.py:275,341); the mask is the pure function keep(seed, site, m*N+n) also used */
  uint64_t drop_seed;  /* by the backward kernels.  drop_p = 0 disables it. */
  int32_t b_major;   /* 0: B stored [N, K].  1: B stored [K, N] with row stride ldb, e.g. B = W^T for a data
                        gradient straight from W [out, in], or B = X^T for a weight gradient */
  int32_t a_dtype, b_dtype; /* RL_DT_BF16 or RL_DT_F16, both the same (out_dtype / res_dtype take RL_DT_F16 too,
                               independently; out2 shares out's 16-bit format) */
  const uint64_t* drop_counter; /* optional device step counter added to drop_seed at run time (NULL = plain seed) */
  int32_t tune_tile_n;  /* 0: the cost model picks the N tile.  64 / 128 / 256: force it (tuning and tests; results do
                           not depend on it) */
  int32_t tune_no_pair; /* 0: cost model.  1: never use the CTA-pair (cta_group::2) kernels.  2: pairs but no 4-CTA
                           clusters (two pairs sharing the B tile through TMA multicast).  3: 4-CTA clusters whenever
                           eligible (M >= 512, 256-wide tiles, plain B).  4: no long-K (5-stage) variant.  5: no compile-time
                           specialised epilogues.  Results do not depend on it. */
  float* colsum;     /* optional f32 [N], ACCUMULATED into: sum over the M rows of the final result (after act): the bias
                        gradient of the layer below comes out of the data-gradient GEMM, BatchNorm's batch sum out of the
                        conv GEMM.  Needs split_k = 0. */
  float* colsumsq;   /* optional f32 [N], accumulated: sum over rows of result^2 (BatchNorm batch statistics) */
  int32_t sm_reserve; /* SMs this (persistent, one-CTA-per-SM) launch leaves free: the grid, the wave model and the split-K
                         chooser work with num_SMs - sm_reserve.  Data parallelism sets it while an NCCL all-reduce of a
                         gradient bucket runs concurrently with the rest of the backward (src/run.py:165-167, :200). */
  int32_t b_mode;    /* 0: B is a 2-D matrix.  1 (needs b_major = 1, a_mode = 0): B is the im2col matrix of the conv
                        activation `b` ([NIMG, P, H, W, C], geometry / taps in the conv_* fields), never materialised:
                        B[k = output pixel (img, oh, ow), n = tap * Cuse + c] = x[img, plane_t, oh+dh_t, ow+dw_t, c],
                        one 5-D TMA box per (64 pixels, 64 channels), zero outside the map.  With A = dY^T (a_major 1)
                        and split_k this is a conv weight gradient dW[co, tap, ci] in ONE GEMM (K = NIMG*H*W). */
} rl_gemm_desc;

RL_API int rl_gemm_bf16(const rl_gemm_desc* d, void* stream);
/* ---- fused attention core -------------------------------------------------------------------
 * ctx[b*L+q, h*64:(h+1)*64] = softmax(Q K^T / 8 + (1 - mask[b,:]) * -10000) V  per head.
 * qkv: bf16 [B*L, 3*heads*64] = [Q | K | V] as written by the fused QKV projection GEMM;
 * mask: int64 [B, L] (batch['masks']); ctx: bf16 [B*L, heads*64].  L <= 256, head_dim == 64.
 * Replaces BertSelfAttention.forward score/softmax/context path (modeling_bert.py:234-260) and the
 * extended-mask construction of BertModel.forward (:687, :696-697).  Dropout is identity (eval). */
RL_API int rl_attention_fwd(const void* qkv, const int64_t* mask, void* ctx,
                            float* row_lse /* optional [B, heads, L] f32: log2-domain logsumexp of every query row, saved
                                              for rl_attention_bwd */,
                            int64_t B, int64_t L, int64_t heads, int64_t head_dim,
                            int32_t act_dtype /* format of qkv / ctx: RL_DT_BF16 or RL_DT_F16 */, float drop_p,
                            uint64_t drop_seed, uint32_t drop_site, const uint64_t* drop_counter, void* stream);

/* ---- LayerNorm (biased variance, eps inside sqrt) over rows of f32 [rows, H] ------------------
 * Replaces BertLayerNorm in BertSelfOutput/BertOutput (modeling_bert.py:276, :342) and
 * resnet_layernorm (src/models.py:838).  Writes f32 and/or bf16 (either may be NULL). */
RL_API int rl_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out_f32,
                            void* out_bf16, int64_t rows, int64_t H, float eps, float drop_p, uint64_t drop_seed,
                            uint32_t drop_site, const uint64_t* drop_counter, int32_t drop_f32 /* mask the f32 output too */,
                            int32_t out16_dtype /* format of out_bf16: RL_DT_BF16 or RL_DT_F16 */, void* stream);

/* ---- BertEmbeddings.forward (modeling_bert.py:169-193) ---------------------------------------
 * out = LN(src + pos[position] + type[0]); src = word[ids[row]] when inputs_embeds is NULL else
 * inputs_embeds[row].  pos_mode 0: position = row % L (default arange); 1: position 0 for every
 * token (output_block, src/models.py:852-854). */
RL_API int rl_embed_ln_fwd(const int64_t* ids, const float* word, const float* inputs_embeds,
                           const float* pos, const float* type0, const float* gamma, const float* beta,
                           float* out_f32, void* out_bf16, float* pre_ln_out /* optional: the summed embedding */,
                           int64_t rows, int64_t L, int64_t H, int32_t pos_mode, float eps, float drop_p,
                           uint64_t drop_seed, uint32_t drop_site, const uint64_t* drop_counter, int32_t out16_dtype,
                           void* stream);

/* ---- gated fusion (src/models.py:840-850; src/models_abla.py:243-279) -------------------------
 * m0 = bert_hiddens, m1/m2 = the other present modalities in the reference's concat order, all
 * f32 [B*L, H].  gate_w: f32 [G, (G+1)*H], gate_b: [G], G = num_modal.  sum_mode != 0 is
 * fusion='sum'.  mean_dot_ws: f32 [B*3] scratch.  gates_out (optional): f32 [B*L, 3]. */
RL_API int rl_gate_fuse_fwd(const float* m0, const float* m1, const float* m2, int32_t num_modal,
                            int32_t sum_mode, const int64_t* mask, const float* gate_w,
                            const float* gate_b, float* mean_dot_ws, float* out, float* gates_out,
                            int64_t B, int64_t L, int64_t H, void* stream);

/* ---- masked CrossEntropyLoss (src/models.py:862-868) -----------------------------------------
 * loss = mean over rows with loss_mask == 1 of (logsumexp(logits[row]) - logits[row, tgt[row]]).
 * logits f32 (or fp16: the train step's choice — half the bytes written by the classifier GEMM and re-read here and by
 * the backward) [rows, V] with row stride ld (elements); row_loss_ws f32 [rows] scratch; loss f32 [1]. */
RL_API int rl_masked_ce_fwd(const void* logits, int32_t logits_dtype /* RL_DT_F32 or RL_DT_F16 */, const int64_t* tgt,
                            const int64_t* loss_mask, float* row_loss_ws, float* loss, float* row_lse_out /* optional [rows] */,
                            float* count_out /* optional [1] */, int64_t rows, int64_t V, int64_t ld,
                            void* stream);

/* ---- row-wise argmax over the vocabulary (first maximum wins) --------------------------------
 * Replaces the host-side `logits.detach().cpu().numpy()` + argmax of src/test.py:140-145 and
 * src/run.py:262-263 so that only [B, L] int64 ids cross PCIe.  out: int64 [rows]. */
RL_API int rl_argmax_rows(const float* logits, int64_t* out, int64_t rows, int64_t V, int64_t ld,
                          void* stream);

/* ---- pinyin GRU (src/models.py:818-826: nn.Embedding -> pack_padded_sequence -> nn.GRU) -------
 * rl_gru_input_table: table[v, :] = W_ih emb[v] + b_ih  (f32 [V=33, 3H]; gate order r,z,n).
 * rl_gru_step_fwd: one time step t for every token row; gh = h_{t-1} W_hh^T + b_hh (f32 [rows,3H],
 * produced by rl_gemm_bf16) or NULL with h_prev NULL at t = 0 (h_0 = 0).  Rows with lens[row] <= t
 * keep their state, which reproduces the packed-sequence final hidden.  pho_idx int64 [rows, T]. */
RL_API int rl_gru_input_table(const float* emb, const float* w_ih, const float* b_ih, float* table,
                              int64_t V, int64_t H, void* stream);
RL_API int rl_gru_step_fwd(const float* gh, const float* b_hh, const float* table,
                           const int64_t* pho_idx, const int32_t* lens, const float* h_prev,
                           float* h_out, void* h_out_bf16, int64_t rows, int64_t H, int64_t T,
                           int64_t t, int32_t out16_dtype, void* stream);

/* ---- glyph stem (src/models.py:829-834 gather; src/char_cnn.py:15-29 for res_block1) ----------
 * For every token: image = glyphs[ids[i]] (f32 [C,32,32]);  y1 = relu(bn1(conv3x3 s2 p1)),
 * ysc = bn_sc(conv1x1 s2), both bf16 NHWC [n_img,16,16,64].  BatchNorm (eval) is passed folded:
 * scale = gamma / sqrt(running_var + 1e-5), shift = beta - running_mean * scale.
 * w1: f32 [64, C, 3, 3], wsc: f32 [64, C] (the reference weight layouts). */
RL_API int rl_glyph_stem_fwd(const float* glyphs, const int64_t* ids, const float* w1, const float* wsc,
                             const float* scale1, const float* shift1, const float* scale_sc,
                             const float* shift_sc, void* y1, void* ysc, int64_t n_img, int32_t C,
                             void* stream);

/* ---- fused CharResNet block 1, eval mode (src/models.py:829-834 + src/char_cnn.py:15-32) -------
 * glyph gather -> conv3x3/s2+BN+ReLU -> conv3x3+BN (+ conv1x1/s2 shortcut+BN) -> ReLU in ONE persistent
 * tcgen05 kernel; HBM traffic per glyph = C*4 KB in + 32 KB out.  Weights are passed packed, bf16,
 * with the BatchNorm scale folded in (scale = gamma / sqrt(running_var + 1e-5)):
 *   w1_packed  [64, 32]   : [co, c*9+kh*3+kw] = conv1.weight[co,c,kh,kw] * scale1[co], zero padded
 *   wsc_packed [64, 32]   : [co, c*9+4]       = shortcut.weight[co,c]    * scale_sc[co]
 *   w2_packed  [64, 576]  : [co, (kh*3+kw)*64+ci] = conv2.weight[co,ci,kh,kw] * scale2[co]
 *   t1 [64] = BN1 shift;  t2s [64] = BN2 shift + shortcut-BN shift.
 * out: bf16 [n_img*256, 64], rows parity-split ([img][oh&1][ow&1][oh/2][ow/2]) for block 2. */
RL_API int rl_glyph_block1_fwd(const float* glyphs, const int64_t* ids, const void* w1_packed,
                               const void* wsc_packed, const void* w2_packed, const float* t1,
                               const float* t2s, void* out, int64_t n_img, int32_t C, void* stream);

/* ======================= training path (src/run.py:191-212) =========================================== */

/* ---- attention backward: dqkv = [dQ | dK | dV] from dctx, recomputing the probabilities ----
 * Differentiates BertSelfAttention.forward (modeling_bert.py:234-260) including its dropout on the probabilities.
 * qkv / ctx are the forward tensors (ctx is needed for delta = rowsum(dO o O)); every 16-bit tensor of the call
 * (qkv, ctx, dctx, dqkv) has the format act_dtype; row_lse
 * (optional) is the logsumexp saved by rl_attention_fwd: P = exp2(s - lse) is then recomputed in one pass over S.
 * dbias (optional, seq_len <= 128): f32 [3*heads*64], ACCUMULATED into — the column sums of dqkv, i.e. the bias gradient of
 * the fused query/key/value projection (BertSelfAttention.query/key/value.bias), summed from the staged 16-bit tiles while
 * their TMA stores are in flight; saves a separate pass over dqkv.
 * seq_len <= 256. */
RL_API int rl_attention_bwd(const void* qkv, const int64_t* mask, const void* ctx, const void* dctx, void* dqkv,
                            float* dbias, const float* row_lse, int64_t B, int64_t L, int64_t heads, int64_t head_dim, int32_t act_dtype,
                            float drop_p, uint64_t drop_seed, uint32_t drop_site, const uint64_t* drop_counter,
                            void* stream);

/* ---- LayerNorm backward: x = LN input (f32), dy = grad of the LN output; dx (+= add_in) in f32 and/or bf16;
 * dgamma/dbeta/dxsum (column sums of dy*xhat, dy, dx) are ACCUMULATED into (caller zeroes them per step). */
RL_API int rl_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* add_in, float* dx,
                            void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int64_t rows, int64_t H,
                            float eps, float drop_p, uint64_t drop_seed, uint32_t site_in /* 0 = none: dy is masked */,
                            uint32_t site_out /* 0 = none: dx_bf16 and dxsum are masked */, const uint64_t* drop_counter,
                            void* stream);

/* ---- keep mask of a dropout site as bytes (tests feed the kernel's masks to the CPU oracle) ---- */
RL_API int rl_dropout_mask(uint8_t* out, int64_t n, float drop_p, uint64_t drop_seed, uint32_t drop_site,
                           const uint64_t* drop_counter, void* stream);

/* ---- out[c] += sum_r x[r, c] for a bf16 matrix (bias gradients of nn.Linear) ---- */
RL_API int rl_colsum_bf16(const void* x, float* out, int64_t rows, int64_t cols, int64_t ld, void* stream);

/* ---- masked CE backward: dlogits (bf16 [rows, ldd], columns >= V zeroed) from the logits, the per-row
 * logsumexp and the active-row count saved by rl_masked_ce_fwd; gscale = d(loss) (device scalar, may be NULL). */
RL_API int rl_masked_ce_bwd(const void* logits, int32_t logits_dtype, const int64_t* tgt, const int64_t* loss_mask,
                            const float* row_lse, const float* count, const float* gscale, void* dlogits, int64_t rows,
                            int64_t V, int64_t ld, int64_t ldd, void* stream);

/* ---- BertEmbeddings backward: dword[ids[row]] += de[row], dpos[position(row)] += de[row] (vector atomics) ---- */
RL_API int rl_embed_bwd(const float* de, const int64_t* ids, float* dword, float* dpos, int64_t rows, int64_t L,
                        int64_t H, int32_t pos_mode, void* stream);

/* ---- gated fusion backward (src/models.py:840-850).  gates = f32 [B*L, 3] saved by rl_gate_fuse_fwd.
 * dm0..2 (f32 [B*L, H]) are written; dgate_w / dgate_b are accumulated.  ws: f32 scratch, B*L*3 + 2*B*H. */
RL_API int rl_gate_fuse_bwd(const float* dhid, const float* m0, const float* m1, const float* m2, int32_t num_modal,
                            const int64_t* mask, const float* gates, const float* gate_w, float* dm0, float* dm1,
                            float* dm2, float* dgate_w, float* dgate_b, float* ws, int64_t B, int64_t L, int64_t H,
                            void* stream);

/* ---- pinyin GRU backward through time (nn.GRU over packed sequences, src/models.py:818-826) ----
 * One step t: gates are recomputed from gh (saved forward recurrent pre-activations; NULL with h_prev NULL at
 * t = 0), dh_prev = dh * z (the dgh W_hh term is added by a following rl_gemm_bf16 with res = dh_prev), dgi / dgh
 * bf16 [rows, 3H] and the bf16 one-hot [rows, 64] of the step's symbol are GEMM operands for
 * dW_hh += dgh^T h_prev, dtable += onehot^T dgi.  rl_gru_table_bwd turns dtable [V, 3H] into db_ih, dW_ih, demb
 * (accumulating). */
RL_API int rl_gru_step_bwd(const float* dh, const float* gh, const float* b_hh, const float* table,
                           const int64_t* pho_idx, const int32_t* lens, const float* h_prev, float* dh_prev, void* dgi,
                           void* dgh, void* onehot, int64_t rows, int64_t H, int64_t T, int64_t t, void* stream);
RL_API int rl_gru_table_bwd(const float* dtable, const float* emb, const float* w_ih, float* dw_ih, float* db_ih,
                            float* demb, int64_t V, int64_t H, void* stream);

/* ---- CharResNet training pieces (src/char_cnn.py:9-55 with nn.BatchNorm2d in batch-statistics mode) -------------
 * rl_bn_stats: sums[0:C] += sum_m x, sums[C:2C] += sum_m x^2 over a raw conv output [M, C] (f32 in training; caller zeroes sums).
 * rl_bn_finalize: mean, biased var -> scale = gamma*rstd, shift = beta - mean*scale, mean/rstd saved for the backward;
 *   running_mean/var updated with `momentum` (unbiased var) and num_batches_tracked += 1, as nn.BatchNorm2d does.
 * rl_bn_apply: out = [relu](x1*scale1 + shift1 [+ x2*scale2 + shift2]); remap = 1 writes parity-split rows.
 * rl_bn_bwd: BatchNorm backward for dy (f32/bf16; rows parity-split when remap) optionally masked by the ReLU that
 *   followed (act_out > 0); dbeta/dgamma accumulated, dx written as bf16 with row stride ldx.
 *   fwd_scale / fwd_shift (optional, the rl_bn_finalize outputs the forward applied): the ReLU mask is re-derived from the
 *   raw conv outputs as (x*scale + shift [+ x2*scale2 + shift2]) > 0 — rl_bn_apply's own arithmetic, so the same bits —
 *   and act_out is not read at all (one tensor-sized HBM stream less in each of the two passes).
 * rl_im2col_bf16: col[m, t*C + ci] = x[img, plane_t, oh+dh_t, ow+dw_t, ci] so that a conv weight gradient
 *   dW[co, t, ci] is ONE split-K rl_gemm_bf16 (A = dY MN-major, B = col MN-major).
 * rl_glyph_im2col: the same for res_block1 straight from the glyph table: col1 [n*256, 32] (27 used), colsc [n*256, 8]
 *   (optional, may be NULL: the shortcut's pixel is the centre tap of col1). */
RL_API int rl_bn_stats(const void* x, int32_t x_dtype, float* sums, int64_t M, int64_t C, int64_t ld, void* stream);
RL_API int rl_bn_finalize(const float* sums, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int64_t* num_batches_tracked, float* scale, float* shift, float* mean_out,
                          float* rstd_out, int64_t M, int64_t C, float momentum, float eps, void* stream);
RL_API int rl_bn_apply(const void* x1, const float* scale1, const float* shift1, const void* x2, const float* scale2,
                       const float* shift2, int32_t x_dtype, void* out, int32_t out_dtype, int32_t relu, int64_t M,
                       int64_t C, int32_t remap, int32_t map_h, int32_t map_w, void* stream);
RL_API int rl_bn_bwd(const void* dy, int32_t dy_dtype, const void* act_out, int32_t act_dtype, const void* x,
                     int32_t x_dtype, const float* mean, const float* rstd, const float* gamma, float* dbeta, float* dgamma, void* dx,
                     int64_t ldx, int64_t M, int64_t C, int32_t remap, int32_t map_h, int32_t map_w, const float* fwd_scale,
                     const float* fwd_shift, void* stream);
/* Two BatchNorms that fed the same ReLU (out = relu(bn2(c2) + bn_s(cs)), src/char_cnn.py:31-32) differentiated in one
 * reduce + one apply pass: dy / act_out are read once for both branches.  x2 == NULL: single branch (== rl_bn_bwd).
 * x1 / x2 are f32 or bf16 (x_dtype; res_block1-2 keep their raw conv outputs in bf16 to halve the BatchNorm traffic).  C must be 8*2^k (< 256) or a multiple of 256; all pointers 16-byte aligned. */
RL_API int rl_bn_bwd2(const void* dy, int32_t dy_dtype, const void* act_out, int32_t act_dtype, int32_t x_dtype, const void* x1,
                      const float* mean1, const float* rstd1, const float* gamma1, float* dbeta1, float* dgamma1, void* dx1,
                      int64_t ldx1, const void* x2, const float* mean2, const float* rstd2, const float* gamma2,
                      float* dbeta2, float* dgamma2, void* dx2, int64_t ldx2, int64_t M, int64_t C, int32_t remap,
                      int32_t map_h, int32_t map_w, const float* fwd_scale1, const float* fwd_shift1,
                      const float* fwd_scale2, const float* fwd_shift2, void* stream);
RL_API int rl_im2col_bf16(const void* x, void* col, int64_t n_img, int32_t C, int32_t W, int32_t H, int32_t P,
                          int32_t ntaps, const int8_t* tap_dw, const int8_t* tap_dh, const int8_t* tap_plane,
                          void* stream);
RL_API int rl_glyph_im2col(const float* glyphs, const int64_t* ids, void* col1, void* colsc, int64_t n_img, int32_t C,
                           void* stream);

/* ---- multi-tensor grad-norm and fused clip + AdamW (src/run.py:207 clip_grad_norm_,
 * transformers/optimization.py:113-169).  table: device array of {float* p; const float* g; float* m;
 * float* v; bf16* shadow; float* shadow32; int64 n; float wd; int shadow_f16 (the shadow is fp16, not bf16)}; chunks: device array of int2 {tensor, chunk} covering
 * every 4096-element block.  rl_mt_sumsq writes sum(g^2) to out[0]: one partial per CTA into partials_ws (ws_floats >= 1;
 * rl_workspace_bytes("mt_sumsq") gives the size that keeps every SM busy), then a fixed-order second stage — the result
 * is bit-reproducible, so data-parallel replicas that hold identical gradients apply identical clip coefficients;
 * rl_mt_adamw applies g *= min(1, max_norm / (sqrt(sumsq)/grad_div + 1e-6)) / grad_div, then the AdamW update
 * with bias corrections bias_corr1 = 1-beta1^t, bias_corr2 = 1-beta2^t, and refreshes the bf16 shadow copy. */
RL_API int rl_mt_sumsq(const void* table, const void* chunks, int64_t num_chunks, float* out, float* partials_ws,
                       int64_t ws_floats, void* stream);
RL_API int rl_mt_adamw(const void* table, const void* chunks, int64_t num_chunks, const float* sumsq, float max_norm,
                       float lr, float beta1, float beta2, float eps, float bias_corr1, float bias_corr2,
                       float grad_div, void* stream);
/* Multi-tensor gather + cast in ONE launch: for every table entry {void* dst; const int32* map; int64 n; int32 dst_dtype;
 * int32 pad} and i < n: dst[i] = cast(srcs[map[i] >> 24][map[i] & 0xFFFFFF]), or 0 where map[i] < 0.  srcs: device array of
 * f32 pointers (<= 128 sources of < 2^24 elements each); chunks: device int2 {entry, 4096-element chunk}.  Used to
 * re-derive the conv operand layouts (tap-major, transposed, parity-plane matrices) from the fp32 master weights after
 * each optimizer step, and to scatter the tap-major weight-gradient GEMM outputs into the [cout, cin, kh, kw] gradients
 * (src/char_cnn.py:15-29 weights). */
RL_API int rl_mt_gather(const void* table, const void* chunks, int64_t num_chunks, const void* srcs, void* stream);

/* out[r, :] = table[ids[r], :] for f32 rows of H (inference glyph cache: in eval mode CharResNet + its LayerNorm input
 * is a pure function of the token id, so a [vocab, 768] table replaces the CNN; src/models.py:829-838). */
RL_API int rl_gather_rows_f32(const float* table, const int64_t* ids, float* out, int64_t rows, int64_t H, void* stream);

/* Split-precision operand for the tied classifier (src/models.py:859): out[r] = [hi | lo | hi] with hi = bf16(x[r]),
 * lo = bf16(x[r] - hi), width 3*cols.  Against B = [W_hi | W_hi | W_lo] one rl_gemm_bf16 with K = 3*cols computes
 * x W^T with ~16-bit mantissa operands: the 21128-way logits no longer carry the bf16 rounding of seq and E. */
RL_API int rl_split3_bf16(const float* x, void* out, int64_t rows, int64_t cols, int32_t out_dtype, void* stream);

/* GELU (erf form, transformers/modeling_bert.py:125-131) as element-wise passes next to the K = 768 GEMMs of
 * BertIntermediate: h = u * Phi(u) over n bf16 elements; and its backward fused with the bias gradient:
 * t[r, c] <- t[r, c] * gelu'(u[r, c]) in place (t = dy2 W2), dbias[c] += sum_r of the fp32 products. */
RL_API int rl_gelu_fwd(const void* u, void* h, int64_t n, int32_t dtype /* of u and h */, void* stream);
RL_API int rl_gelu_bwd_colsum(void* t /* bf16 gradient */, const void* u, float* dbias, int64_t rows, int64_t cols, int64_t ld,
                              int32_t u_dtype, void* stream);

/* Same update with the step's schedule read from device memory: hyper = {lr, 1 - beta1^t, 1 - beta2^t}.  Lets a
 * CUDA graph that contains the optimizer be replayed while LambdaLR (src/run.py:153-160) and the bias corrections move. */
RL_API int rl_mt_adamw_dev(const void* table, const void* chunks, int64_t num_chunks, const float* sumsq, float max_norm,
                           const float* hyper, float beta1, float beta2, float eps, float grad_div, void* stream);

/* Scratch sizes (bytes) of the entry points that take a caller-owned workspace: op is one of "gate_fuse_fwd"
 * (mean_dot_ws), "gate_fuse_bwd" (ws), "masked_ce_fwd" (row_loss_ws), "mt_sumsq" (partials_ws; B, L, H ignored).  Returns -1
 * for an unknown op. */
RL_API int64_t rl_workspace_bytes(const char* op, int64_t B, int64_t L, int64_t H);

#endif /* REALISE_B200_H */
