// Pinyin GRU pieces (input-projection table, per-step gate update) and the glyph stem
// (glyph gather + res_block1 conv3x3/s2 + 1x1/s2 shortcut with folded BatchNorm).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// GRU (src/models.py:661-669, :818-826).  The input projection W_ih x_t + b_ih only ever sees the
// 33 rows of pho_embeddings, so it is a [33, 3H] lookup table.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gru_table_kernel(const float* __restrict__ emb, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                 float* __restrict__ table, int V, int H) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // output column in [0, 3H)
  if (j >= 3 * H) return;
  float w[32];
  const int per = H / 32;  // <= 32
  for (int i = 0; i < per; ++i) w[i] = w_ih[(long long)j * H + i * 32 + lane];
  for (int v = 0; v < V; ++v) {
    float s = 0.f;
    for (int i = 0; i < per; ++i) s += w[i] * __ldg(emb + (long long)v * H + i * 32 + lane);
    s = rl::warp_sum(s);
    if (lane == 0) table[(long long)v * 3 * H + j] = s + b_ih[j];
  }
}

// one GRU time step for all rows; gh = h_{t-1} W_hh^T + b_hh (f32 [N,3H]) or NULL at t = 0 (h = 0)
__global__ void __launch_bounds__(256)
gru_step_kernel(const float* __restrict__ gh, const float* __restrict__ b_hh, const float* __restrict__ table,
                const long long* __restrict__ pho_idx, const int* __restrict__ lens, const float* __restrict__ h_prev,
                float* __restrict__ h_out, __nv_bfloat16* __restrict__ h_out_bf16, long long rows, int H, int T, int t, int f16) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bool active = lens[row] > t;
  const int nv = H / 128;
  float4* ho = reinterpret_cast<float4*>(h_out + row * H);
  uint2* hb = reinterpret_cast<uint2*>(h_out_bf16 + row * H);
  if (!active) {
    // finished (or padding) sequences keep their hidden state
    for (int i = 0; i < nv; ++i) {
      const float4 h = h_prev ? reinterpret_cast<const float4*>(h_prev + row * H)[i * 32 + lane]
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      ho[i * 32 + lane] = h;
      hb[i * 32 + lane] = make_uint2(rl::pack_h(h.x, h.y, f16), rl::pack_h(h.z, h.w, f16));
    }
    return;
  }
  const long long sym = pho_idx[row * T + t];
  const float* gi = table + sym * 3 * H;
  for (int i = 0; i < nv; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 ir = __ldg(reinterpret_cast<const float4*>(gi + c));
    const float4 iz = __ldg(reinterpret_cast<const float4*>(gi + H + c));
    const float4 in = __ldg(reinterpret_cast<const float4*>(gi + 2 * H + c));
    float4 hr, hz, hn, h;
    if (gh) {
      hr = *reinterpret_cast<const float4*>(gh + row * 3 * H + c);
      hz = *reinterpret_cast<const float4*>(gh + row * 3 * H + H + c);
      hn = *reinterpret_cast<const float4*>(gh + row * 3 * H + 2 * H + c);
      h = *reinterpret_cast<const float4*>(h_prev + row * H + c);
    } else {
      hr = __ldg(reinterpret_cast<const float4*>(b_hh + c));
      hz = __ldg(reinterpret_cast<const float4*>(b_hh + H + c));
      hn = __ldg(reinterpret_cast<const float4*>(b_hh + 2 * H + c));
      h = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 o;
#define RL_GRU_ELT(f)                                              \
  {                                                                \
    const float r = 1.0f / (1.0f + expf(-(ir.f + hr.f)));          \
    const float z = 1.0f / (1.0f + expf(-(iz.f + hz.f)));          \
    const float n = tanhf(in.f + r * hn.f);                        \
    o.f = (1.0f - z) * n + z * h.f;                                \
  }
    RL_GRU_ELT(x) RL_GRU_ELT(y) RL_GRU_ELT(z) RL_GRU_ELT(w)
#undef RL_GRU_ELT
    ho[i * 32 + lane] = o;
    hb[i * 32 + lane] = make_uint2(rl::pack_h(o.x, o.y, f16), rl::pack_h(o.z, o.w, f16));
  }
}

// ---------------------------------------------------------------------------------------------
// Glyph stem: images = table[src_idx] (src/models.py:829-834) -> res_block1.residual_function.0
// (conv3x3 s2 p1, C->64) + BN + ReLU and res_block1.shortcut (conv1x1 s2, C->64) + BN
// (src/char_cnn.py:15-29), BN folded into scale/shift.  One CTA per glyph, one thread per output
// pixel, outputs NHWC bf16 [N,16,16,64] ready to be TMA-loaded by the block-1 conv2 implicit GEMM.
// ---------------------------------------------------------------------------------------------
constexpr int STEM_CO = 64;

template <int C>
__global__ void __launch_bounds__(256)
glyph_stem_kernel(const float* __restrict__ glyphs, const long long* __restrict__ ids, const float* __restrict__ w1,
                  const float* __restrict__ wsc, const float* __restrict__ scale1, const float* __restrict__ shift1,
                  const float* __restrict__ scale_sc, const float* __restrict__ shift_sc, __nv_bfloat16* __restrict__ y1,
                  __nv_bfloat16* __restrict__ ysc) {
  __shared__ float s_img[C * 1024];
  __shared__ __align__(16) float s_w1[C * 9 * STEM_CO];  // [c][kh][kw][co]
  __shared__ __align__(16) float s_wsc[C * STEM_CO];     // [c][co]
  __shared__ __align__(16) float s_aff[4 * STEM_CO];
  const int tid = threadIdx.x;
  const long long img = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(glyphs + ids[img] * (long long)(C * 1024));
  for (int i = tid; i < C * 256; i += 256) reinterpret_cast<float4*>(s_img)[i] = __ldg(src + i);
  for (int i = tid; i < C * 9 * STEM_CO; i += 256) {
    const int co = i % STEM_CO, tap = i / STEM_CO;  // tap = c*9 + kh*3 + kw ; w1 is [co][c][kh][kw]
    s_w1[i] = w1[co * (C * 9) + tap];
  }
  for (int i = tid; i < C * STEM_CO; i += 256) s_wsc[i] = wsc[(i % STEM_CO) * C + i / STEM_CO];
  if (tid < STEM_CO) {
    s_aff[tid] = scale1[tid];
    s_aff[STEM_CO + tid] = shift1[tid];
    s_aff[2 * STEM_CO + tid] = scale_sc[tid];
    s_aff[3 * STEM_CO + tid] = shift_sc[tid];
  }
  __syncthreads();
  const int oh = tid >> 4, ow = tid & 15;
  float acc[STEM_CO];
  // ---- conv3x3 stride 2 pad 1 ----
#pragma unroll
  for (int j = 0; j < STEM_CO; ++j) acc[j] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = 2 * ow + kw - 1;
        const float x = (ih >= 0 && iw >= 0) ? s_img[c * 1024 + ih * 32 + iw] : 0.f;
        const float4* w = reinterpret_cast<const float4*>(s_w1 + ((c * 3 + kh) * 3 + kw) * STEM_CO);
#pragma unroll
        for (int j = 0; j < STEM_CO / 4; ++j) {
          const float4 ww = w[j];
          acc[4 * j] += x * ww.x; acc[4 * j + 1] += x * ww.y; acc[4 * j + 2] += x * ww.z; acc[4 * j + 3] += x * ww.w;
        }
      }
    }
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(y1 + (img * 256 + tid) * STEM_CO);
#pragma unroll
    for (int g = 0; g < STEM_CO / 8; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[8 * g + j] * s_aff[8 * g + j] + s_aff[STEM_CO + 8 * g + j], 0.f);
      dst[g] = make_uint4(rl::pack_bf16(v[0], v[1]), rl::pack_bf16(v[2], v[3]), rl::pack_bf16(v[4], v[5]),
                          rl::pack_bf16(v[6], v[7]));
    }
  }
  // ---- shortcut conv1x1 stride 2 ----
#pragma unroll
  for (int j = 0; j < STEM_CO; ++j) acc[j] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float x = s_img[c * 1024 + (2 * oh) * 32 + 2 * ow];
    const float4* w = reinterpret_cast<const float4*>(s_wsc + c * STEM_CO);
#pragma unroll
    for (int j = 0; j < STEM_CO / 4; ++j) {
      const float4 ww = w[j];
      acc[4 * j] += x * ww.x; acc[4 * j + 1] += x * ww.y; acc[4 * j + 2] += x * ww.z; acc[4 * j + 3] += x * ww.w;
    }
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(ysc + (img * 256 + tid) * STEM_CO);
#pragma unroll
    for (int g = 0; g < STEM_CO / 8; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = acc[8 * g + j] * s_aff[2 * STEM_CO + 8 * g + j] + s_aff[3 * STEM_CO + 8 * g + j];
      dst[g] = make_uint4(rl::pack_bf16(v[0], v[1]), rl::pack_bf16(v[2], v[3]), rl::pack_bf16(v[4], v[5]),
                          rl::pack_bf16(v[6], v[7]));
    }
  }
}


// ---------------------------------------------------------------------------------------------
// GRU backward through time (one step).  Recomputes the gates from the saved gh_t (or b_hh at t = 0) and the
// input table, then for active rows:
//   dn = dh (1-z), dz = dh (h_prev - n), dh_prev = dh z, dn_pre = dn (1-n^2), dr_pre = dn_pre gh_n r (1-r),
//   dz_pre = dz z (1-z);   dgi = [dr_pre, dz_pre, dn_pre],  dgh = [dr_pre, dz_pre, dn_pre r]
// Rows whose sequence already ended pass dh through unchanged.  dgi / dgh are written as bf16 GEMM operands
// (data gradient dgh W_hh, weight gradient dgh^T h_prev, table gradient onehot^T dgi), together with the bf16
// one-hot row of the step's pinyin symbol.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gru_step_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gh, const float* __restrict__ b_hh,
                    const float* __restrict__ table, const long long* __restrict__ pho_idx, const int* __restrict__ lens,
                    const float* __restrict__ h_prev, float* __restrict__ dh_prev, __nv_bfloat16* __restrict__ dgi,
                    __nv_bfloat16* __restrict__ dgh, __nv_bfloat16* __restrict__ onehot, long long rows, int H, int T, int t) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bool active = lens[row] > t;
  const int nv = H / 128;
  const long long sym = active ? pho_idx[row * T + t] : -1;
  // one-hot [64] bf16: two lanes-worth of uint32 pairs
  reinterpret_cast<uint32_t*>(onehot + row * 64)[lane] =
      (sym == 2 * lane ? 0x3F80u : 0u) | (sym == 2 * lane + 1 ? 0x3F800000u : 0u);
  const float* gi = table + (active ? sym : 0) * 3 * H;
  for (int i = 0; i < nv; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 d = *reinterpret_cast<const float4*>(dh + row * H + c);
    uint2 z2 = make_uint2(0u, 0u);
    if (!active) {
      *reinterpret_cast<float4*>(dh_prev + row * H + c) = d;
      *reinterpret_cast<uint2*>(dgi + row * 3 * H + c) = z2;
      *reinterpret_cast<uint2*>(dgi + row * 3 * H + H + c) = z2;
      *reinterpret_cast<uint2*>(dgi + row * 3 * H + 2 * H + c) = z2;
      *reinterpret_cast<uint2*>(dgh + row * 3 * H + c) = z2;
      *reinterpret_cast<uint2*>(dgh + row * 3 * H + H + c) = z2;
      *reinterpret_cast<uint2*>(dgh + row * 3 * H + 2 * H + c) = z2;
      continue;
    }
    const float4 ir = __ldg(reinterpret_cast<const float4*>(gi + c));
    const float4 iz = __ldg(reinterpret_cast<const float4*>(gi + H + c));
    const float4 in = __ldg(reinterpret_cast<const float4*>(gi + 2 * H + c));
    float4 hr, hz, hn, hp;
    if (gh) {
      hr = *reinterpret_cast<const float4*>(gh + row * 3 * H + c);
      hz = *reinterpret_cast<const float4*>(gh + row * 3 * H + H + c);
      hn = *reinterpret_cast<const float4*>(gh + row * 3 * H + 2 * H + c);
      hp = *reinterpret_cast<const float4*>(h_prev + row * H + c);
    } else {
      hr = __ldg(reinterpret_cast<const float4*>(b_hh + c));
      hz = __ldg(reinterpret_cast<const float4*>(b_hh + H + c));
      hn = __ldg(reinterpret_cast<const float4*>(b_hh + 2 * H + c));
      hp = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 o_prev, g_r, g_z, g_n, g_nr;
#define RL_GRU_BWD(f)                                              \
  {                                                                \
    const float r = 1.0f / (1.0f + expf(-(ir.f + hr.f)));          \
    const float z = 1.0f / (1.0f + expf(-(iz.f + hz.f)));          \
    const float n = tanhf(in.f + r * hn.f);                        \
    const float dn_pre = d.f * (1.0f - z) * (1.0f - n * n);        \
    const float dz_pre = d.f * (hp.f - n) * z * (1.0f - z);        \
    const float dr_pre = dn_pre * hn.f * r * (1.0f - r);           \
    o_prev.f = d.f * z;                                            \
    g_r.f = dr_pre; g_z.f = dz_pre; g_n.f = dn_pre; g_nr.f = dn_pre * r; \
  }
    RL_GRU_BWD(x) RL_GRU_BWD(y) RL_GRU_BWD(z) RL_GRU_BWD(w)
#undef RL_GRU_BWD
    *reinterpret_cast<float4*>(dh_prev + row * H + c) = o_prev;
    const uint2 pr = make_uint2(rl::pack_bf16(g_r.x, g_r.y), rl::pack_bf16(g_r.z, g_r.w));
    const uint2 pz = make_uint2(rl::pack_bf16(g_z.x, g_z.y), rl::pack_bf16(g_z.z, g_z.w));
    *reinterpret_cast<uint2*>(dgi + row * 3 * H + c) = pr;
    *reinterpret_cast<uint2*>(dgi + row * 3 * H + H + c) = pz;
    *reinterpret_cast<uint2*>(dgi + row * 3 * H + 2 * H + c) = make_uint2(rl::pack_bf16(g_n.x, g_n.y), rl::pack_bf16(g_n.z, g_n.w));
    *reinterpret_cast<uint2*>(dgh + row * 3 * H + c) = pr;
    *reinterpret_cast<uint2*>(dgh + row * 3 * H + H + c) = pz;
    *reinterpret_cast<uint2*>(dgh + row * 3 * H + 2 * H + c) = make_uint2(rl::pack_bf16(g_nr.x, g_nr.y), rl::pack_bf16(g_nr.z, g_nr.w));
  }
}

// table[v] = W_ih emb[v] + b_ih  =>  db_ih = sum_v dT[v],  dW_ih[j, :] = sum_v dT[v, j] emb[v, :],  demb = dT W_ih
__global__ void __launch_bounds__(256)
gru_table_bwd_w_kernel(const float* __restrict__ dT, const float* __restrict__ emb, float* __restrict__ dw_ih,
                       float* __restrict__ db_ih, int V, int H) {
  const int j = blockIdx.x;  // row of W_ih, in [0, 3H)
  __shared__ float s_d[64];
  if (threadIdx.x < V) s_d[threadIdx.x] = dT[(long long)threadIdx.x * 3 * H + j];
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f;
    for (int v = 0; v < V; ++v) b += s_d[v];
    db_ih[j] += b;
  }
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc += s_d[v] * __ldg(emb + (long long)v * H + c);
    dw_ih[(long long)j * H + c] += acc;
  }
}
// demb[v, c] += sum_j dT[v, j] W_ih[j, c]: one CTA per (symbol v, 128-row slab of W_ih), partial sums by atomics
// (V = 33 symbols alone would leave most SMs idle)
__global__ void __launch_bounds__(256)
gru_table_bwd_e_kernel(const float* __restrict__ dT, const float* __restrict__ w_ih, float* __restrict__ demb, int V, int H) {
  const int v = blockIdx.x;
  const int j0 = blockIdx.y * 128;
  const int j1 = min(j0 + 128, 3 * H);
  __shared__ float s_d[128];
  if (threadIdx.x < 128) s_d[threadIdx.x] = j0 + threadIdx.x < j1 ? dT[(long long)v * 3 * H + j0 + threadIdx.x] : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = j0; j < j1; ++j) acc = fmaf(s_d[j - j0], __ldg(w_ih + (long long)j * H + c), acc);
    atomicAdd(demb + (long long)v * H + c, acc);
  }
}

}  // namespace

extern "C" int rl_gru_step_bwd(const float* dh, const float* gh, const float* b_hh, const float* table,
                               const int64_t* pho_idx, const int32_t* lens, const float* h_prev, float* dh_prev, void* dgi,
                               void* dgh, void* onehot, int64_t rows, int64_t H, int64_t T, int64_t t, void* stream) {
  RL_REQUIRE(dh && b_hh && table && pho_idx && lens && dh_prev && dgi && dgh && onehot, RL_EINVAL, "rl_gru_step_bwd: null pointer");
  RL_REQUIRE((gh == nullptr) == (h_prev == nullptr), RL_EINVAL, "rl_gru_step_bwd: gh and h_prev go together");
  RL_REQUIRE(H % 128 == 0 && t >= 0 && t < T, RL_EINVAL, "rl_gru_step_bwd: bad shape");
  if (rows <= 0) return 0;
  gru_step_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      dh, gh, b_hh, table, (const long long*)pho_idx, lens, h_prev, dh_prev, (__nv_bfloat16*)dgi, (__nv_bfloat16*)dgh,
      (__nv_bfloat16*)onehot, rows, (int)H, (int)T, (int)t);
  return rl_check_launch("rl_gru_step_bwd");
}

extern "C" int rl_gru_table_bwd(const float* dtable, const float* emb, const float* w_ih, float* dw_ih, float* db_ih,
                                float* demb, int64_t V, int64_t H, void* stream) {
  RL_REQUIRE(dtable && emb && w_ih && dw_ih && db_ih && demb && V > 0 && V <= 64, RL_EINVAL, "rl_gru_table_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  gru_table_bwd_w_kernel<<<(unsigned)(3 * H), 256, 0, st>>>(dtable, emb, dw_ih, db_ih, (int)V, (int)H);
  gru_table_bwd_e_kernel<<<dim3((unsigned)V, (unsigned)((3 * H + 127) / 128)), 256, 0, st>>>(dtable, w_ih, demb, (int)V, (int)H);
  return rl_check_launch("rl_gru_table_bwd");
}

namespace {

}  // namespace

extern "C" int rl_gru_input_table(const float* emb, const float* w_ih, const float* b_ih, float* table, int64_t V,
                                  int64_t H, void* stream) {
  RL_REQUIRE(emb && w_ih && b_ih && table, RL_EINVAL, "rl_gru_input_table: null pointer");
  RL_REQUIRE(H % 32 == 0 && H <= 1024 && V > 0, RL_EINVAL, "rl_gru_input_table: bad shape");
  const int wpb = 8;
  gru_table_kernel<<<(unsigned)((3 * H + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(emb, w_ih, b_ih, table,
                                                                                             (int)V, (int)H);
  return rl_check_launch("rl_gru_input_table");
}

extern "C" int rl_gru_step_fwd(const float* gh, const float* b_hh, const float* table, const int64_t* pho_idx,
                               const int32_t* lens, const float* h_prev, float* h_out, void* h_out_bf16, int64_t rows,
                               int64_t H, int64_t T, int64_t t, int32_t out16_dtype, void* stream) {
  RL_REQUIRE(table && pho_idx && lens && h_out && h_out_bf16 && b_hh, RL_EINVAL, "rl_gru_step_fwd: null pointer");
  RL_REQUIRE((gh == nullptr) == (h_prev == nullptr), RL_EINVAL, "rl_gru_step_fwd: gh and h_prev go together");
  RL_REQUIRE(H % 128 == 0 && t >= 0 && t < T, RL_EINVAL, "rl_gru_step_fwd: bad shape");
  if (rows <= 0) return 0;
  const int wpb = 8;
  gru_step_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      gh, b_hh, table, (const long long*)pho_idx, lens, h_prev, h_out, (__nv_bfloat16*)h_out_bf16, rows, (int)H, (int)T,
      (int)t, out16_dtype == RL_DT_F16);
  return rl_check_launch("rl_gru_step_fwd");
}

extern "C" int rl_glyph_stem_fwd(const float* glyphs, const int64_t* ids, const float* w1, const float* wsc,
                                 const float* scale1, const float* shift1, const float* scale_sc,
                                 const float* shift_sc, void* y1, void* ysc, int64_t n_img, int32_t C, void* stream) {
  RL_REQUIRE(glyphs && ids && w1 && wsc && scale1 && shift1 && scale_sc && shift_sc && y1 && ysc, RL_EINVAL,
             "rl_glyph_stem_fwd: null pointer");
  RL_REQUIRE(C >= 1 && C <= 3, RL_EINVAL, "rl_glyph_stem_fwd: num_fonts must be 1, 2 or 3 (got %d)", C);
  RL_REQUIRE(((uintptr_t)glyphs & 15) == 0 && ((uintptr_t)y1 & 15) == 0 && ((uintptr_t)ysc & 15) == 0, RL_EALIGN,
             "rl_glyph_stem_fwd: alignment");
  if (n_img <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3)
    glyph_stem_kernel<3><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, w1, wsc, scale1, shift1,
                                                         scale_sc, shift_sc, (__nv_bfloat16*)y1, (__nv_bfloat16*)ysc);
  else if (C == 2)
    glyph_stem_kernel<2><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, w1, wsc, scale1, shift1,
                                                         scale_sc, shift_sc, (__nv_bfloat16*)y1, (__nv_bfloat16*)ysc);
  else
    glyph_stem_kernel<1><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, w1, wsc, scale1, shift1,
                                                         scale_sc, shift_sc, (__nv_bfloat16*)y1, (__nv_bfloat16*)ysc);
  return rl_check_launch("rl_glyph_stem_fwd");
}
