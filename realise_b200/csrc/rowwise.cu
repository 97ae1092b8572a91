// Row-wise (HBM-bound) kernels of the ReaLiSe path: embedding-sum + LayerNorm, LayerNorm,
// gated fusion, masked cross-entropy.  One warp per 768-wide row, 16-byte vector accesses.
#include "common.cuh"

namespace {

constexpr int MAX_V4 = 8;  // per-lane float4 count: H <= 1024

__device__ __forceinline__ void ln_row_store(const float4* x, int nv, float mean, float rstd,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             float* out_f32, __nv_bfloat16* out_bf16, int lane,
                                             const rl::DropSpec& drop, bool drop_f32, long long elem0, int f16 = 0) {
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      float4 y;
      y.x = (x[i].x - mean) * rstd * g.x + b.x;
      y.y = (x[i].y - mean) * rstd * g.y + b.y;
      y.z = (x[i].z - mean) * rstd * g.z + b.z;
      y.w = (x[i].w - mean) * rstd * g.w + b.w;
      float4 yd = y;
      if (drop.thresh) {  // dropout after the LayerNorm (BertEmbeddings.forward :192, src/models.py:858)
        yd = y;
        rl::drop_apply4(drop, elem0 + col, yd);
      }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + col) = drop_f32 ? yd : y;
      if (out_bf16)
        *reinterpret_cast<uint2*>(out_bf16 + col) = make_uint2(rl::pack_h(yd.x, yd.y, f16), rl::pack_h(yd.z, yd.w, f16));
    }
  }
}

__device__ __forceinline__ void ln_stats(const float4* x, int nv, int H, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) s += x[i].x + x[i].y + x[i].z + x[i].w;
  mean = rl::warp_sum(s) / (float)H;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
      v += a * a + b * b + c * c + d * d;
    }
  v = rl::warp_sum(v) / (float)H;
  rstd = rsqrtf(v + eps);
}

// LayerNorm over the last dim (biased variance, eps inside the sqrt) — nn.LayerNorm /
// BertLayerNorm.  x: f32 [rows, H]
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* out_f32,
                                                        __nv_bfloat16* out_bf16, long long rows, int H, float eps,
                                                        rl::DropSpec drop, int drop_f32, int f16) {
  rl::drop_resolve(drop);
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H / 128;
  float4 v[MAX_V4];
  const float4* src = reinterpret_cast<const float4*>(x + row * H);
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) v[i] = src[i * 32 + lane];
  float mean, rstd;
  ln_stats(v, nv, H, eps, mean, rstd);
  ln_row_store(v, nv, mean, rstd, gamma, beta, out_f32 ? out_f32 + row * H : nullptr,
               out_bf16 ? out_bf16 + row * H : nullptr, lane, drop, drop_f32 != 0, row * H, f16);
}

// BertEmbeddings.forward: LN(word[ids] (or inputs_embeds) + pos[position] + type[0])
// pos_mode 0: position = token index within the sentence (default arange), 1: position 0 for all.
__global__ void __launch_bounds__(256)
embed_ln_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ inputs_embeds,
                const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* out_f32, __nv_bfloat16* out_bf16, float* pre_out, long long rows,
                int L, int H, int pos_mode, float eps, rl::DropSpec drop, int f16) {
  rl::drop_resolve(drop);
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H / 128;
  const float4* src = inputs_embeds ? reinterpret_cast<const float4*>(inputs_embeds + row * H)
                                    : reinterpret_cast<const float4*>(word + ids[row] * (long long)H);
  const int position = pos_mode == 0 ? (int)(row % L) : 0;
  const float4* ps = reinterpret_cast<const float4*>(pos + (long long)position * H);
  const float4* ts = reinterpret_cast<const float4*>(type0);
  float4 v[MAX_V4];
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) {
      const float4 a = src[i * 32 + lane], b = __ldg(ps + i * 32 + lane), c = __ldg(ts + i * 32 + lane);
      v[i] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
      if (pre_out) reinterpret_cast<float4*>(pre_out + row * H)[i * 32 + lane] = v[i];
    }
  float mean, rstd;
  ln_stats(v, nv, H, eps, mean, rstd);
  ln_row_store(v, nv, mean, rstd, gamma, beta, out_f32 ? out_f32 + row * H : nullptr,
               out_bf16 ? out_bf16 + row * H : nullptr, lane, drop, true, row * H, f16);
}

// ---- gated fusion (src/models.py:840-850) ----
// pass 1: per sentence, masked mean of bert_hiddens and its contribution to each gate logit
__global__ void __launch_bounds__(256)
gate_mean_kernel(const float* __restrict__ bert_h, const long long* __restrict__ mask, const float* __restrict__ gate_w,
                 float* __restrict__ mean_dot, int L, int H, int G) {
  extern __shared__ float s_mean[];  // [H]
  __shared__ float s_red[8][4];
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int l = 0; l < L; ++l) cnt += (float)mask[(long long)b * L + l];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += bert_h[((long long)b * L + l) * H + c] * (float)mask[(long long)b * L + l];
    s_mean[c] = s / cnt;
  }
  __syncthreads();
  float d[3] = {0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < H; c += blockDim.x)
    for (int g = 0; g < G; ++g) d[g] += s_mean[c] * gate_w[(long long)g * (G + 1) * H + (long long)G * H + c];
  for (int g = 0; g < G; ++g) {
    d[g] = rl::warp_sum(d[g]);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][g] = d[g];
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += s_red[w][threadIdx.x];
    mean_dot[b * 3 + threadIdx.x] = s;
  }
}

// pass 2: per token, G dot products of length G*H, sigmoid, weighted sum of the modalities
__global__ void __launch_bounds__(256)
gate_fuse_kernel(const float* __restrict__ m0, const float* __restrict__ m1, const float* __restrict__ m2,
                 const float* __restrict__ gate_w, const float* __restrict__ gate_b, const float* __restrict__ mean_dot,
                 float* __restrict__ out, float* __restrict__ gates_out, long long rows, int L, int H, int G, int sum_mode) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H / 128;
  const float* mods[3] = {m0, m1, m2};
  float4 x[3][MAX_V4];
  float dots[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    if (m < G) {
      const float4* src = reinterpret_cast<const float4*>(mods[m] + row * H);
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i)
        if (i < nv) x[m][i] = src[i * 32 + lane];
    }
  }
  float g[3] = {1.f, 1.f, 1.f};
  if (!sum_mode) {
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
      if (gi < G) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          if (m < G) {
            const float4* w = reinterpret_cast<const float4*>(gate_w + (long long)gi * (G + 1) * H + (long long)m * H);
#pragma unroll
            for (int i = 0; i < MAX_V4; ++i)
              if (i < nv) {
                const float4 ww = __ldg(w + i * 32 + lane);
                dots[gi] += x[m][i].x * ww.x + x[m][i].y * ww.y + x[m][i].z * ww.z + x[m][i].w * ww.w;
              }
          }
        }
        const float z = rl::warp_sum(dots[gi]) + mean_dot[(row / L) * 3 + gi] + gate_b[gi];
        g[gi] = 1.0f / (1.0f + expf(-z));
      }
    }
    if (gates_out && lane < G) gates_out[row * 3 + lane] = g[lane];
  }
  float4* dst = reinterpret_cast<float4*>(out + row * H);
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int m = 0; m < 3; ++m)
        if (m < G) {
          y.x += g[m] * x[m][i].x; y.y += g[m] * x[m][i].y; y.z += g[m] * x[m][i].z; y.w += g[m] * x[m][i].w;
        }
      dst[i * 32 + lane] = y;
    }
}

// ---- masked cross entropy (src/models.py:862-868): mean over loss_mask==1 of lse - logit[tgt] ----
// logits may be f32 or IEEE fp16 (F16: the train step keeps its [tokens, vocab] logits in fp16 — 0.69 GB instead of
// 1.38 GB written by the classifier GEMM and read again here and by the backward; nobody reads train-mode logits at
// f32 precision, and 11 mantissa bits put 2^-12 relative error on exp(x - lse), far below the bf16 rounding of dlogits)
template <bool F16>
__device__ __forceinline__ float ce_ld(const void* x, long long i) {
  return F16 ? __half2float(reinterpret_cast<const __half*>(x)[i]) : reinterpret_cast<const float*>(x)[i];
}
// 8 consecutive logits starting at element 8*g of a 16-byte aligned row
template <bool F16>
__device__ __forceinline__ void ce_ld8(const void* x, int g, float (&v)[8]) {
  if (F16) {
    const uint4 a = reinterpret_cast<const uint4*>(x)[g];
    v[0] = rl::half_lo(a.x, 1); v[1] = rl::half_hi(a.x, 1); v[2] = rl::half_lo(a.y, 1); v[3] = rl::half_hi(a.y, 1);
    v[4] = rl::half_lo(a.z, 1); v[5] = rl::half_hi(a.z, 1); v[6] = rl::half_lo(a.w, 1); v[7] = rl::half_hi(a.w, 1);
  } else {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * g], b = reinterpret_cast<const float4*>(x)[2 * g + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <bool F16>
__global__ void __launch_bounds__(256)
ce_row_kernel(const void* __restrict__ logits, const long long* __restrict__ tgt, const long long* __restrict__ loss_mask,
              float* __restrict__ row_loss, float* __restrict__ row_lse, long long rows, int V, long long ld) {
  const long long row = blockIdx.x;
  __shared__ float s_red[8];
  if (loss_mask[row] != 1) {
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }
  const void* x = reinterpret_cast<const char*>(logits) + row * ld * (F16 ? 2 : 4);
  // one pass, online softmax: running (max, sum of exp) per thread, 8 logits per step when the row is 16-byte aligned
  float mx = -INFINITY, s = 0.f;
  const bool vec = ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int V8 = vec ? (V >> 3) : 0;
  for (int i = threadIdx.x; i < V8; i += blockDim.x) {
    float v[8];
    ce_ld8<F16>(x, i, v);
    const float m8 = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
    if (m8 > mx) {
      s *= __expf(mx - m8);
      mx = m8;
    }
    s += ((__expf(v[0] - mx) + __expf(v[1] - mx)) + (__expf(v[2] - mx) + __expf(v[3] - mx))) +
         ((__expf(v[4] - mx) + __expf(v[5] - mx)) + (__expf(v[6] - mx) + __expf(v[7] - mx)));
  }
  for (int i = V8 * 8 + threadIdx.x; i < V; i += blockDim.x) {
    const float v = ce_ld<F16>(x, i);
    if (v > mx) {
      s *= __expf(mx - v);
      mx = v;
    }
    s += __expf(v - mx);
  }
  // combine (max, sum) pairs: warp shuffle, then the 8 warp results
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float m = fmaxf(mx, m2);
    s = (mx == -INFINITY ? 0.f : s * __expf(mx - m)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - m));
    mx = m;
  }
  __shared__ float s_sum[8];
  if ((threadIdx.x & 31) == 0) {
    s_red[threadIdx.x >> 5] = mx;
    s_sum[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w] == -INFINITY ? 0.f : s_sum[w] * __expf(s_red[w] - m);
    const float lse = logf(t) + m;
    if (row_lse) row_lse[row] = lse;
    row_loss[row] = lse - ce_ld<F16>(x, tgt[row]);
  }
}

__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ loss_mask, float* __restrict__ loss,
                 float* __restrict__ count_out, long long rows) {
  __shared__ float s_sum[32], s_cnt[32];
  float s = 0.f, c = 0.f;
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
    s += row_loss[i];
    c += loss_mask[i] == 1 ? 1.f : 0.f;
  }
  s = rl::warp_sum(s);
  c = rl::warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = s;
    s_cnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int w = 0; w < 32; ++w) {
      ts += s_sum[w];
      tc += s_cnt[w];
    }
    loss[0] = ts / tc;
    if (count_out) count_out[0] = tc;
  }
}

// row-wise argmax (first maximum wins, like numpy/torch): the post-processing step of
// src/test.py:140-145 / src/run.py:262-263 moved on-device so only [B, L] ids cross PCIe
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ logits, long long* __restrict__ out, int V, long long ld) {
  const long long row = blockIdx.x;
  const float* x = logits + row * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = x[i];
    if (v > best) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) {
        best = s_v[w];
        bi = s_i[w];
      }
    out[row] = bi;
  }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(unsigned char* out, long long n, rl::DropSpec d) {
  rl::drop_resolve(d);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (d.thresh == 0u || rl::drop_keep(d.seed, d.site, (unsigned long long)i, d.thresh)) ? 1 : 0;
}

bool h_ok(int64_t H) { return H > 0 && H % 128 == 0 && H <= 128 * MAX_V4; }

}  // namespace

extern "C" int rl_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out_f32,
                                void* out_bf16, int64_t rows, int64_t H, float eps, float drop_p, uint64_t drop_seed,
                                uint32_t drop_site, const uint64_t* drop_counter, int32_t drop_f32, int32_t out16_dtype,
                                void* stream) {
  RL_REQUIRE(x && gamma && beta && (out_f32 || out_bf16), RL_EINVAL, "rl_layernorm_fwd: null pointer");
  RL_REQUIRE(h_ok(H), RL_EINVAL, "rl_layernorm_fwd: H=%lld must be a multiple of 128, <= 1024", (long long)H);
  if (rows <= 0) return 0;
  const int wpb = 8;
  layernorm_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      x, gamma, beta, out_f32, (__nv_bfloat16*)out_bf16, rows, (int)H, eps, rl::make_drop(drop_p, drop_seed, drop_site, drop_counter),
      drop_f32, out16_dtype == RL_DT_F16);
  return rl_check_launch("rl_layernorm_fwd");
}

extern "C" int rl_embed_ln_fwd(const int64_t* ids, const float* word, const float* inputs_embeds, const float* pos,
                               const float* type0, const float* gamma, const float* beta, float* out_f32,
                               void* out_bf16, float* pre_ln_out, int64_t rows, int64_t L, int64_t H, int32_t pos_mode,
                               float eps, float drop_p, uint64_t drop_seed, uint32_t drop_site,
                               const uint64_t* drop_counter, int32_t out16_dtype, void* stream) {
  RL_REQUIRE((inputs_embeds || (ids && word)) && pos && type0 && gamma && beta && (out_f32 || out_bf16), RL_EINVAL,
             "rl_embed_ln_fwd: null pointer");
  RL_REQUIRE(h_ok(H) && L > 0, RL_EINVAL, "rl_embed_ln_fwd: bad H/L");
  if (rows <= 0) return 0;
  const int wpb = 8;
  embed_ln_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, word, inputs_embeds, pos, type0, gamma, beta, out_f32, (__nv_bfloat16*)out_bf16, pre_ln_out,
      rows, (int)L, (int)H, pos_mode, eps, rl::make_drop(drop_p, drop_seed, drop_site, drop_counter),
      out16_dtype == RL_DT_F16);
  return rl_check_launch("rl_embed_ln_fwd");
}

extern "C" int rl_gate_fuse_fwd(const float* m0, const float* m1, const float* m2, int32_t num_modal,
                                int32_t sum_mode, const int64_t* mask, const float* gate_w, const float* gate_b,
                                float* mean_dot_ws, float* out, float* gates_out, int64_t B, int64_t L, int64_t H,
                                void* stream) {
  RL_REQUIRE(m0 && out && num_modal >= 1 && num_modal <= 3, RL_EINVAL, "rl_gate_fuse_fwd: bad arguments");
  RL_REQUIRE((num_modal < 2 || m1) && (num_modal < 3 || m2), RL_EINVAL, "rl_gate_fuse_fwd: missing modality");
  RL_REQUIRE(h_ok(H), RL_EINVAL, "rl_gate_fuse_fwd: bad H");
  if (!sum_mode) RL_REQUIRE(mask && gate_w && gate_b && mean_dot_ws, RL_EINVAL, "rl_gate_fuse_fwd: gate needs mask/weights/ws");
  if (B <= 0 || L <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!sum_mode) {
    gate_mean_kernel<<<(unsigned)B, 256, (size_t)H * 4, st>>>(m0, (const long long*)mask, gate_w, mean_dot_ws, (int)L,
                                                              (int)H, num_modal);
    int rc = rl_check_launch("rl_gate_fuse_fwd(mean)");
    if (rc) return rc;
  }
  const long long rows = B * L;
  const int wpb = 8;
  gate_fuse_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(m0, m1, m2, gate_w, gate_b, mean_dot_ws, out,
                                                                           gates_out, rows, (int)L, (int)H, num_modal,
                                                                           sum_mode);
  return rl_check_launch("rl_gate_fuse_fwd");
}

extern "C" int rl_masked_ce_fwd(const void* logits, int32_t logits_dtype, const int64_t* tgt, const int64_t* loss_mask,
                                float* row_loss_ws, float* loss, float* row_lse_out, float* count_out, int64_t rows, int64_t V,
                                int64_t ld, void* stream) {
  RL_REQUIRE(logits && tgt && loss_mask && row_loss_ws && loss, RL_EINVAL, "rl_masked_ce_fwd: null pointer");
  RL_REQUIRE(rows > 0 && V > 0 && ld >= V, RL_EINVAL, "rl_masked_ce_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  RL_REQUIRE(logits_dtype == RL_DT_F32 || logits_dtype == RL_DT_F16, RL_EINVAL, "rl_masked_ce_fwd: logits must be f32 or fp16");
  if (logits_dtype == RL_DT_F16)
    ce_row_kernel<true><<<(unsigned)rows, 256, 0, st>>>(logits, (const long long*)tgt, (const long long*)loss_mask, row_loss_ws,
                                                        row_lse_out, rows, (int)V, ld);
  else
    ce_row_kernel<false><<<(unsigned)rows, 256, 0, st>>>(logits, (const long long*)tgt, (const long long*)loss_mask, row_loss_ws,
                                                         row_lse_out, rows, (int)V, ld);
  int rc = rl_check_launch("rl_masked_ce_fwd(rows)");
  if (rc) return rc;
  ce_reduce_kernel<<<1, 1024, 0, st>>>(row_loss_ws, (const long long*)loss_mask, loss, count_out, rows);
  return rl_check_launch("rl_masked_ce_fwd");
}

// x (f32) -> [hi | lo | hi] bf16 rows of width 3*cols: hi = bf16(x), lo = bf16(x - hi).  With B = [W_hi | W_hi | W_lo]
// one K = 3*cols GEMM evaluates x W^T with 16-bit mantissa operands (hi*hi + lo*hi + hi*lo) on the bf16 tensor path.
__global__ void __launch_bounds__(256)
split3_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int cols, int f16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over rows * cols/4
  const int c4 = cols >> 2;
  if (i >= rows * c4) return;
  const long long r = i / c4;
  const int c = (int)(i - r * c4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * cols + c);
  const uint32_t h0 = rl::pack_h(v.x, v.y, f16), h1 = rl::pack_h(v.z, v.w, f16);
  const uint32_t l0 = rl::pack_h(v.x - rl::half_lo(h0, f16), v.y - rl::half_hi(h0, f16), f16);
  const uint32_t l1 = rl::pack_h(v.z - rl::half_lo(h1, f16), v.w - rl::half_hi(h1, f16), f16);
  __nv_bfloat16* o = out + r * 3 * cols + c;
  *reinterpret_cast<uint2*>(o) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(o + cols) = make_uint2(l0, l1);
  *reinterpret_cast<uint2*>(o + 2 * cols) = make_uint2(h0, h1);
}

// out[r] = table[ids[r]] (f32 rows of H): the inference-time glyph cache lookup (CharResNet in eval mode is a pure
// function of the token id, src/models.py:829-838)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ ids, float* __restrict__ out, long long rows,
                   int H) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(table + ids[row] * (long long)H);
  float4* dst = reinterpret_cast<float4*>(out + row * H);
  for (int i = lane; i < H / 4; i += 32) dst[i] = __ldg(src + i);
}

extern "C" int rl_gather_rows_f32(const float* table, const int64_t* ids, float* out, int64_t rows, int64_t H, void* stream) {
  RL_REQUIRE(table && ids && out && H > 0 && H % 4 == 0 && ((((uintptr_t)table | (uintptr_t)out) & 15) == 0), RL_EALIGN,
             "rl_gather_rows_f32: H must be a multiple of 4 and the pointers 16-byte aligned");
  if (rows <= 0) return 0;
  gather_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(table, (const long long*)ids, out, rows, (int)H);
  return rl_check_launch("rl_gather_rows_f32");
}

extern "C" int rl_split3_bf16(const float* x, void* out, int64_t rows, int64_t cols, int32_t out_dtype, void* stream) {
  RL_REQUIRE(x && out && cols > 0 && cols % 4 == 0 && (((uintptr_t)x & 15) == 0) && (((uintptr_t)out & 7) == 0), RL_EALIGN,
             "rl_split3_bf16: cols must be a multiple of 4 and the pointers aligned");
  if (rows <= 0) return 0;
  const long long n = rows * (cols / 4);
  split3_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, rows, (int)cols, out_dtype == RL_DT_F16);
  return rl_check_launch("rl_split3_bf16");
}

extern "C" int rl_argmax_rows(const float* logits, int64_t* out, int64_t rows, int64_t V, int64_t ld, void* stream) {
  RL_REQUIRE(logits && out && V > 0 && ld >= V, RL_EINVAL, "rl_argmax_rows: bad arguments");
  if (rows <= 0) return 0;
  argmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logits, (long long*)out, (int)V, ld);
  return rl_check_launch("rl_argmax_rows");
}

extern "C" int rl_dropout_mask(uint8_t* out, int64_t n, float drop_p, uint64_t drop_seed, uint32_t drop_site,
                               const uint64_t* drop_counter, void* stream) {
  RL_REQUIRE(out && n >= 0, RL_EINVAL, "rl_dropout_mask: bad arguments");
  if (n == 0) return 0;
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, rl::make_drop(drop_p, drop_seed, drop_site, drop_counter));
  return rl_check_launch("rl_dropout_mask");
}
