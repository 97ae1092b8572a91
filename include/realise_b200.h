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
enum { RL_ACT_NONE = 0, RL_ACT_GELU = 1, RL_ACT_RELU = 2, RL_ACT_TANH = 3 };
enum { RL_DT_BF16 = 0, RL_DT_F32 = 1 };

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
} rl_gemm_desc;

RL_API int rl_gemm_bf16(const rl_gemm_desc* d, void* stream);

#endif /* REALISE_B200_H */
