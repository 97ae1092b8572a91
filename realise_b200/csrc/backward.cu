// Backward / training-step kernels that are not GEMMs: LayerNorm backward, bias (column) sums, masked-CE
// backward, embedding scatter, gated-fusion backward, global grad-norm + fused clip/AdamW (multi-tensor).
// Reference semantics: transformers/modeling_bert.py (BertLayerNorm, BertEmbeddings), src/models.py:840-868,
// src/run.py:207 (clip_grad_norm_), transformers/optimization.py:113-169 (AdamW.step).
#include "common.cuh"

namespace {

constexpr int MAX_V4 = 8;

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward.  x = LN input (f32), dy = grad of LN output.  dx = rstd * (g - mean(g) - xhat*mean(g*xhat)),
// g = dy*gamma.  Optionally dx += add_in (gradient arriving through the residual connection), a bf16 copy of dx
// (operand of the following weight/data-gradient GEMMs), and the column sums dgamma += sum dy*xhat,
// dbeta += sum dy, dxsum += sum dx (bias gradient of the linear layer that produced x / type-embedding grad).
// ---------------------------------------------------------------------------------------------------------
// Persistent layout: 3 CTAs per SM, each CTA owns a contiguous slab of rows (warp w takes rows w, w+8, ... of the
// slab).  The three column partials (dgamma, dbeta, dxsum) of a warp live in that warp's PRIVATE shared-memory slice
// ([3][NV][32 lanes] float4 = 9 KB at H = 768: plain LDS.128 / FADD / STS.128, no atomics — shared-memory float
// atomics are CAS loops), which keeps the kernel at ~80 registers so that 24 warps per SM hide the HBM latency of
// the row loads.  At the end the 8 slices are summed and flushed with one global atomic per column per CTA.
template <int NV>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ add_in, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum, long long rows, int H,
              float eps, rl::DropSpec drop_in, rl::DropSpec drop_out, int rows_per_cta) {
  extern __shared__ float4 s_acc4[];   // [8 warps][3][NV][32]
  rl::drop_resolve(drop_in);
  rl::drop_resolve(drop_out);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = NV < MAX_V4 ? NV : H / 128;
  float4* mine = s_acc4 + (size_t)warp * 3 * NV * 32 + lane;
#pragma unroll
  for (int i = 0; i < 3 * NV; ++i) mine[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool want_g = dgamma != nullptr, want_b = dbeta != nullptr, want_x = dxsum != nullptr;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  for (int rr = warp; rr < rows_per_cta; rr += 8) {
    const long long row = row_begin + rr;
    if (row >= rows) break;
    float4 xv[NV], gv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        xv[i] = reinterpret_cast<const float4*>(x + row * H)[i * 32 + lane];
        gv[i] = reinterpret_cast<const float4*>(dy + row * H)[i * 32 + lane];
      }
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        if (drop_in.thresh) {  // the LN output went through dropout in the forward: dy_eff = dy * keep / (1-p)
          const long long e0 = row * H + (i * 32 + lane) * 4;
          rl::drop_apply4(drop_in, e0, gv[i]);
        }
        s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
      }
    const float mean = rl::warp_sum(s) / (float)H;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        var += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      }
    const float rstd = rsqrtf(rl::warp_sum(var) / (float)H + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
        if (want_g) {
          float4 a = mine[i * 32];
          a.x = fmaf(gv[i].x, xv[i].x, a.x); a.y = fmaf(gv[i].y, xv[i].y, a.y);
          a.z = fmaf(gv[i].z, xv[i].z, a.z); a.w = fmaf(gv[i].w, xv[i].w, a.w);
          mine[i * 32] = a;
        }
        if (want_b) {
          float4 a = mine[(NV + i) * 32];
          a.x += gv[i].x; a.y += gv[i].y; a.z += gv[i].z; a.w += gv[i].w;
          mine[(NV + i) * 32] = a;
        }
        gv[i].x *= gm.x; gv[i].y *= gm.y; gv[i].z *= gm.z; gv[i].w *= gm.w;  // g = dy * gamma
        m1 += gv[i].x + gv[i].y + gv[i].z + gv[i].w;
        m2 += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
      }
    m1 = rl::warp_sum(m1) / (float)H;
    m2 = rl::warp_sum(m2) / (float)H;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        float4 d;
        d.x = rstd * (gv[i].x - m1 - xv[i].x * m2);
        d.y = rstd * (gv[i].y - m1 - xv[i].y * m2);
        d.z = rstd * (gv[i].z - m1 - xv[i].z * m2);
        d.w = rstd * (gv[i].w - m1 - xv[i].w * m2);
        float4 dm = d;  // gradient of the producing linear layer's output: masked when that output was dropped out
        if (drop_out.thresh) {
          const long long e0 = row * H + (i * 32 + lane) * 4;
          rl::drop_apply4(drop_out, e0, dm);
        }
        if (want_x) {
          float4 a = mine[(2 * NV + i) * 32];
          a.x += dm.x; a.y += dm.y; a.z += dm.z; a.w += dm.w;
          mine[(2 * NV + i) * 32] = a;
        }
        if (dx_bf16)
          reinterpret_cast<uint2*>(dx_bf16 + row * H)[i * 32 + lane] = make_uint2(rl::pack_bf16(dm.x, dm.y), rl::pack_bf16(dm.z, dm.w));
        if (add_in) {
          const float4 a = reinterpret_cast<const float4*>(add_in + row * H)[i * 32 + lane];
          d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
        }
        if (dx) reinterpret_cast<float4*>(dx + row * H)[i * 32 + lane] = d;
      }
  }
  __syncthreads();
  // flush: float slot t of a slice = [which][i][lane][comp] -> column (i*32 + lane)*4 + comp
  const float* s_f = reinterpret_cast<const float*>(s_acc4);
  for (int t = threadIdx.x; t < 3 * NV * 128; t += 256) {
    const int which = t / (NV * 128), u = t % (NV * 128);
    if (u < H) {   // u == column
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += s_f[w * 3 * NV * 128 + t];
      float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
      if (dst) atomicAdd(dst + u, tot);
    }
  }
}

// column sums of a bf16 [rows, cols] matrix (bias gradient): out[c] += sum_r x[r, c].
// Vector path (cols, ld multiples of 8, 16-byte aligned base): a warp reads 512 contiguous bytes of one row (8 bf16
// per lane), the 8 warps of the CTA take rows r, r+8, ... of the CTA's row slab, 4 rows in flight per warp.
__global__ void __launch_bounds__(256)
colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int cols, long long ld,
                       int rows_per_cta) {
  __shared__ float s[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols) {
    const __nv_bfloat16* base = x + c0;
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(base + (r + 8 * u) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += rl::bf16_lo(v[u].x); acc[1] += rl::bf16_hi(v[u].x); acc[2] += rl::bf16_lo(v[u].y); acc[3] += rl::bf16_hi(v[u].y);
        acc[4] += rl::bf16_lo(v[u].z); acc[5] += rl::bf16_hi(v[u].z); acc[6] += rl::bf16_lo(v[u].w); acc[7] += rl::bf16_hi(v[u].w);
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + r * ld);
      acc[0] += rl::bf16_lo(v.x); acc[1] += rl::bf16_hi(v.x); acc[2] += rl::bf16_lo(v.y); acc[3] += rl::bf16_hi(v.y);
      acc[4] += rl::bf16_lo(v.z); acc[5] += rl::bf16_hi(v.z); acc[6] += rl::bf16_lo(v.w); acc[7] += rl::bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// GELU as stand-alone passes.  Inside a GEMM epilogue the erf polynomial is issued by 8 warps per SM and costs more
// than the K = 768 main loop it should hide behind; as an element-wise pass every warp of the SM shares it and the
// kernel runs at HBM speed.  gelu_fwd: h = u * Phi(u) (transformers/modeling_bert.py:125-131).
// gelu_bwd_colsum: du = t * gelu'(u) in place over t (t = dy2 W2), and dbias[c] += sum_r du[r, c] (the bias gradient
// of BertIntermediate.dense) in the same pass — same row-slab layout as colsum_bf16_vec_kernel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ h, long long n8, int f16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(u)[i];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const rl::f2 r = rl::gelu_erf2(rl::f2{rl::half_lo(w[j], f16), rl::half_hi(w[j], f16)});
    o[j] = rl::pack_h(r.x, r.y, f16);
  }
  reinterpret_cast<uint4*>(h)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(__nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ u, float* __restrict__ dbias,
                       long long rows, int cols, long long ld, int rows_per_cta, int u_f16) {
  __shared__ float s[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols) {
    for (long long r = r0 + warp; r < r1; r += 16) {
      const bool two = r + 8 < r1;
      const uint4 ta = *reinterpret_cast<const uint4*>(t + r * ld + c0);
      const uint4 ua = *reinterpret_cast<const uint4*>(u + r * ld + c0);
      uint4 tb = make_uint4(0, 0, 0, 0), ub = tb;
      if (two) {
        tb = *reinterpret_cast<const uint4*>(t + (r + 8) * ld + c0);
        ub = *reinterpret_cast<const uint4*>(u + (r + 8) * ld + c0);
      }
      const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
      const uint32_t uw[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
      uint32_t ow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const rl::f2 d = rl::mul2(rl::f2{rl::bf16_lo(tw[k]), rl::bf16_hi(tw[k])},
                                  rl::gelu_grad2(rl::f2{rl::half_lo(uw[k], u_f16), rl::half_hi(uw[k], u_f16)}));
        ow[k] = rl::pack_bf16(d.x, d.y);
        acc[(k & 3) * 2] += d.x;      // bias gradient from the unrounded products (fp32, like the reference's autograd)
        acc[(k & 3) * 2 + 1] += d.y;
      }
      *reinterpret_cast<uint4*>(t + r * ld + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      if (two) *reinterpret_cast<uint4*>(t + (r + 8) * ld + c0) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols && dbias) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s[w][threadIdx.x];
    atomicAdd(dbias + c, tot);
  }
}

__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int cols, long long ld) {
  // block = 32 columns x 8 row-groups; grid = (cols/32 rounded up, row chunks)
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * 1024;
  float acc = 0.f;
  if (col < cols)
    for (long long r = r0 + ry; r < r0 + 1024 && r < rows; r += 8) acc += __bfloat162float(x[r * ld + col]);
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][cx];
    atomicAdd(out + col, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// masked CE backward: dlogits[r, j] = gscale * mask_r / count * (exp(logit_rj - lse_r) - [j == tgt_r]), bf16,
// columns [V, ldd) zero (K padding of the following GEMMs).
// ---------------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const void* __restrict__ logits, const long long* __restrict__ tgt, const long long* __restrict__ loss_mask,
              const float* __restrict__ row_lse, const float* __restrict__ count, const float* __restrict__ gscale,
              __nv_bfloat16* __restrict__ dlogits, int V, long long ld, long long ldd) {
  const long long row = blockIdx.x;
  __nv_bfloat16* d = dlogits + row * ldd;
  if (loss_mask[row] != 1) {
    for (long long j = threadIdx.x * 8; j < ldd; j += blockDim.x * 8) *reinterpret_cast<uint4*>(d + j) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float coef = (gscale ? gscale[0] : 1.0f) / count[0];
  const float lse = row_lse[row];
  const int t = (int)tgt[row];
  const char* x = reinterpret_cast<const char*>(logits) + row * ld * (F16 ? 2 : 4);
  int j0 = 0;
  if ((ld & 7) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {   // 8 logits -> one 16-byte bf16 store
    const int V8 = V >> 3;
    for (int g = threadIdx.x; g < V8; g += blockDim.x) {
      float v[8];
      if (F16) {
        const uint4 a = reinterpret_cast<const uint4*>(x)[g];
        v[0] = rl::half_lo(a.x, 1); v[1] = rl::half_hi(a.x, 1); v[2] = rl::half_lo(a.y, 1); v[3] = rl::half_hi(a.y, 1);
        v[4] = rl::half_lo(a.z, 1); v[5] = rl::half_hi(a.z, 1); v[6] = rl::half_lo(a.w, 1); v[7] = rl::half_hi(a.w, 1);
      } else {
        const float4 a = reinterpret_cast<const float4*>(x)[2 * g], b = reinterpret_cast<const float4*>(x)[2 * g + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = coef * __expf(v[k] - lse);
      if ((t >> 3) == g) v[t & 7] -= coef;
      *reinterpret_cast<uint4*>(d + 8 * g) =
          make_uint4(rl::pack_bf16(v[0], v[1]), rl::pack_bf16(v[2], v[3]), rl::pack_bf16(v[4], v[5]), rl::pack_bf16(v[6], v[7]));
    }
    j0 = V8 * 8;
  }
  for (int j = j0 + threadIdx.x; j < (int)ldd; j += blockDim.x) {
    float v = 0.f;
    if (j < V) {
      const float xj = F16 ? __half2float(reinterpret_cast<const __half*>(x)[j]) : reinterpret_cast<const float*>(x)[j];
      v = coef * (__expf(xj - lse) - (j == t ? 1.0f : 0.0f));
    }
    d[j] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// embedding backward: de [rows, H] f32 -> dword[ids[row]] += de, dpos[position] += de   (vector atomics)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ de, const long long* __restrict__ ids, float* __restrict__ dword,
                 float* __restrict__ dpos, long long rows, int L, int H, int pos_mode) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H / 128;
  const int position = pos_mode == 0 ? (int)(row % L) : 0;
  for (int i = 0; i < nv; ++i) {
    const float4 g = reinterpret_cast<const float4*>(de + row * H)[i * 32 + lane];
    if (dword) atomicAdd(reinterpret_cast<float4*>(dword + ids[row] * (long long)H) + i * 32 + lane, g);
    if (dpos) atomicAdd(reinterpret_cast<float4*>(dpos + (long long)position * H) + i * 32 + lane, g);
  }
}

// ---------------------------------------------------------------------------------------------------------
// gated fusion backward (src/models.py:840-850).  Forward: z_i = W_i . [m_0..m_{G-1}, mean_b] + b_i, g = sigmoid(z),
// hid = sum_i g_i m_i.  Given dhid: dz_i = (dhid . m_i) g_i (1-g_i);
//   dm_j = g_j dhid + sum_i dz_i W[i, jH:(j+1)H];   dmean_b += sum_{l,i} dz_i W[i, GH:];   db_i = sum dz_i
//   dbert_l += mask_l / cnt_b * dmean_b   (second kernel);  dW = dz^T . cat (third kernel).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_bwd_token_kernel(const float* __restrict__ dhid, const float* __restrict__ m0, const float* __restrict__ m1,
                      const float* __restrict__ m2, const float* __restrict__ gates, const float* __restrict__ gate_w,
                      float* __restrict__ dm0, float* __restrict__ dm1, float* __restrict__ dm2, float* __restrict__ dz_out,
                      float* __restrict__ dmean, float* __restrict__ dgate_b, long long rows, int L, int H, int G) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H / 128;
  const float* mods[3] = {m0, m1, m2};
  float* dms[3] = {dm0, dm1, dm2};
  float4 dh[MAX_V4];
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) dh[i] = reinterpret_cast<const float4*>(dhid + row * H)[i * 32 + lane];
  float g[3] = {0.f, 0.f, 0.f}, dz[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (j < G) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i)
        if (i < nv) {
          const float4 mv = reinterpret_cast<const float4*>(mods[j] + row * H)[i * 32 + lane];
          dot += dh[i].x * mv.x + dh[i].y * mv.y + dh[i].z * mv.z + dh[i].w * mv.w;
        }
      dot = rl::warp_sum(dot);
      g[j] = gates[row * 3 + j];
      dz[j] = dot * g[j] * (1.0f - g[j]);
    }
  if (lane < G) {
    dz_out[row * 3 + lane] = dz[lane];
    atomicAdd(dgate_b + lane, dz[lane]);
  }
  const long long b = row / L;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j < G) {
          float4 o = make_float4(g[j] * dh[i].x, g[j] * dh[i].y, g[j] * dh[i].z, g[j] * dh[i].w);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < G) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(gate_w + (long long)k * (G + 1) * H + (long long)j * H + c));
              o.x += dz[k] * w.x; o.y += dz[k] * w.y; o.z += dz[k] * w.z; o.w += dz[k] * w.w;
            }
          reinterpret_cast<float4*>(dms[j] + row * H)[i * 32 + lane] = o;
        }
      float4 dmn = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < G) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(gate_w + (long long)k * (G + 1) * H + (long long)G * H + c));
          dmn.x += dz[k] * w.x; dmn.y += dz[k] * w.y; dmn.z += dz[k] * w.z; dmn.w += dz[k] * w.w;
        }
      atomicAdd(reinterpret_cast<float4*>(dmean + b * H + c), dmn);
    }
}

// dbert[row] += mask[row] / cnt_b * dmean_b
__global__ void __launch_bounds__(256)
gate_bwd_mean_kernel(float* __restrict__ dm0, const float* __restrict__ dmean, const long long* __restrict__ mask,
                     long long rows, int L, int H) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long b = row / L;
  float cnt = 0.f;
  for (int l = lane; l < L; l += 32) cnt += (float)mask[b * L + l];
  cnt = rl::warp_sum(cnt);
  const float w = (float)mask[row] / cnt;
  if (w == 0.f) return;
  for (int i = 0; i < H / 128; ++i) {
    float4* p = reinterpret_cast<float4*>(dm0 + row * H) + i * 32 + lane;
    const float4 d = reinterpret_cast<const float4*>(dmean + b * H)[i * 32 + lane];
    float4 o = *p;
    o.x += w * d.x; o.y += w * d.y; o.z += w * d.z; o.w += w * d.w;
    *p = o;
  }
}

// dW[k, j*H + c] += sum_rows dz[row, k] * cat_j[row, c]  (cat_j = modality j, or the broadcast masked mean for j == G)
__global__ void __launch_bounds__(256)
gate_bwd_weight_kernel(const float* __restrict__ dz, const float* __restrict__ src, const float* __restrict__ mean_src,
                       float* __restrict__ dgate_w, long long rows, int L, int H, int G, int j) {
  // grid: (H/32, row chunks of 512); block: 32 columns x 8 row lanes
  __shared__ float s[3][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * 512;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
  for (long long r = r0 + ry; r < r0 + 512 && r < rows; r += 8) {
    const float v = mean_src ? mean_src[(r / L) * H + col] : src[r * H + col];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (k < G) acc[k] += dz[r * 3 + k] * v;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) s[k][ry][cx] = acc[k];
  __syncthreads();
  if (ry < G) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[ry][i][cx];
    atomicAdd(dgate_w + (long long)ry * (G + 1) * H + (long long)j * H + col, t);
  }
}

// masked mean of bert_hiddens per sentence (needed again for dW of the mean block)
__global__ void __launch_bounds__(256)
masked_mean_kernel(const float* __restrict__ x, const long long* __restrict__ mask, float* __restrict__ mean, int L, int H) {
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int l = 0; l < L; ++l) cnt += (float)mask[(long long)b * L + l];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += x[((long long)b * L + l) * H + c] * (float)mask[(long long)b * L + l];
    mean[(long long)b * H + c] = s / cnt;
  }
}

// ---------------------------------------------------------------------------------------------------------
// multi-tensor grad-norm + clip + AdamW.  A device table describes every parameter; the grid runs over 4096-
// element chunks listed in a chunk table.  The step is two launches: sum of squares, then the update which
// reads the clip coefficient min(1, max_norm / (norm + 1e-6)) computed on device (no host sync).
// ---------------------------------------------------------------------------------------------------------
struct TensorEntry {   // mirrored by realise_b200/optim.py (ctypes) — keep in sync
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* shadow;  // optional 16-bit operand copy refreshed in the same pass (bf16, or fp16 when shadow_f16)
  float* shadow32;        // optional f32 copy (e.g. the slice of a fused QKV bias vector)
  long long n;
  float wd;
  int shadow_f16;
};
constexpr int OPT_CHUNK = 4096;

// grid-stride over the 4096-element chunks: ~8 CTAs per SM, ONE atomic per CTA (one atomic per chunk is 50 k atomics
// on a single address, which serialise in L2 and cost more than reading the 816 MB of gradients)
__global__ void __launch_bounds__(256)
mt_sumsq_kernel(const TensorEntry* __restrict__ tab, const int2* __restrict__ chunks, float* __restrict__ out,
                long long num_chunks) {
  float s = 0.f;
  for (long long c = blockIdx.x; c < num_chunks; c += gridDim.x) {
    const int2 ck = chunks[c];
    const TensorEntry e = tab[ck.x];
    const long long base = (long long)ck.y * OPT_CHUNK;
    if (base + OPT_CHUNK <= e.n && ((reinterpret_cast<uintptr_t>(e.g) & 15) == 0)) {
      const float4* g4 = reinterpret_cast<const float4*>(e.g + base);
#pragma unroll
      for (int i = 0; i < OPT_CHUNK / 4 / 256; ++i) {
        const float4 g = g4[i * 256 + threadIdx.x];
        s += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
      }
    } else {
      for (int i = threadIdx.x; i < OPT_CHUNK; i += 256) {
        const long long idx = base + i;
        if (idx < e.n) {
          const float g = e.g[idx];
          s += g * g;
        }
      }
    }
  }
  s = rl::warp_sum(s);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[blockIdx.x] = t;     // one partial per CTA, summed in a FIXED order below: float atomics would make the clip
  }                          // coefficient — and after it every parameter — differ in the last bits between ranks
}

// deterministic second stage: total of the per-CTA partials (double accumulation, fixed strided order + fixed tree)
__global__ void __launch_bounds__(256)
mt_sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sh[0];
}

__global__ void __launch_bounds__(256)
mt_adamw_kernel(const TensorEntry* __restrict__ tab, const int2* __restrict__ chunks, const float* __restrict__ sumsq,
                float max_norm, float lr, float beta1, float beta2, float eps, float bc1, float bc2, float grad_div,
                const float* __restrict__ hyper) {
  if (hyper) {  // device-resident schedule (CUDA-graph replay): {lr, 1 - beta1^t, 1 - beta2^t}
    lr = hyper[0];
    bc1 = hyper[1];
    bc2 = hyper[2];
  }
  const int2 ck = chunks[blockIdx.x];
  const TensorEntry e = tab[ck.x];
  const long long base = (long long)ck.y * OPT_CHUNK;
  float coef = 1.0f / grad_div;
  if (max_norm > 0.f) {
    const float norm = sqrtf(sumsq[0]) / grad_div;
    const float c = max_norm / (norm + 1e-6f);
    if (c < 1.0f) coef *= c;
  }
  const float step_size = lr * sqrtf(bc2) / bc1;
  // full, 16-byte aligned chunks: float4 / 8-byte 16-bit stores (28 B per parameter of pure streaming)
  const bool vec = base + OPT_CHUNK <= e.n &&
                   (((reinterpret_cast<uintptr_t>(e.p + base) | reinterpret_cast<uintptr_t>(e.g + base) |
                      reinterpret_cast<uintptr_t>(e.m + base) | reinterpret_cast<uintptr_t>(e.v + base) |
                      (e.shadow32 ? reinterpret_cast<uintptr_t>(e.shadow32 + base) : 0)) & 15) == 0) &&
                   (!e.shadow || (reinterpret_cast<uintptr_t>(e.shadow + base) & 7) == 0);
  if (vec) {
#pragma unroll
    for (int it = 0; it < OPT_CHUNK / 4 / 256; ++it) {
      const long long i4 = base / 4 + it * 256 + threadIdx.x;
      const float4 g4 = reinterpret_cast<const float4*>(e.g)[i4];
      float4 m4 = reinterpret_cast<float4*>(e.m)[i4], v4 = reinterpret_cast<float4*>(e.v)[i4], p4 = reinterpret_cast<float4*>(e.p)[i4];
      float gg[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
      float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mm[k] = beta1 * mm[k] + (1.0f - beta1) * gg[k];
        vv[k] = beta2 * vv[k] + (1.0f - beta2) * gg[k] * gg[k];
        pp[k] -= step_size * mm[k] / (sqrtf(vv[k]) + eps);
        if (e.wd > 0.f) pp[k] -= lr * e.wd * pp[k];
      }
      reinterpret_cast<float4*>(e.m)[i4] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(e.v)[i4] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      reinterpret_cast<float4*>(e.p)[i4] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      if (e.shadow)
        reinterpret_cast<uint2*>(e.shadow)[i4] = make_uint2(rl::pack_h(pp[0], pp[1], e.shadow_f16), rl::pack_h(pp[2], pp[3], e.shadow_f16));
      if (e.shadow32) reinterpret_cast<float4*>(e.shadow32)[i4] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    }
    return;
  }
  for (int i = threadIdx.x; i < OPT_CHUNK; i += 256) {
    const long long idx = base + i;
    if (idx < e.n) {
      const float g = e.g[idx] * coef;
      const float m = beta1 * e.m[idx] + (1.0f - beta1) * g;
      const float v = beta2 * e.v[idx] + (1.0f - beta2) * g * g;
      float pv = e.p[idx];
      pv -= step_size * m / (sqrtf(v) + eps);
      if (e.wd > 0.f) pv -= lr * e.wd * pv;
      e.m[idx] = m;
      e.v[idx] = v;
      e.p[idx] = pv;
      if (e.shadow) {
        if (e.shadow_f16) reinterpret_cast<__half*>(e.shadow)[idx] = __float2half_rn(pv);
        else e.shadow[idx] = __float2bfloat16(pv);
      }
      if (e.shadow32) e.shadow32[idx] = pv;
    }
  }
}

// ---- multi-tensor gather + cast: dst[i] = cast(srcs[map[i] >> 24][map[i] & 0xFFFFFF]) (0 where map[i] < 0) ----------
// One launch re-derives every operand LAYOUT of the conv weights from the fp32 masters (tap-major / transposed /
// parity-plane matrices, zero padded), and one more scatters the tap-major weight-gradient GEMM outputs back into the
// parameters' [cout, cin, kh, kw] gradient layout — instead of ~100 cat / transpose / cast / index_put launches per step.
struct GatherEntry {   // mirrored by realise_b200/train.py (ctypes) — keep in sync
  void* dst;
  const int* map;
  long long n;
  int dst_dtype;       // RL_DT_BF16 / RL_DT_F32 / RL_DT_F16
  int pad;
};

__global__ void __launch_bounds__(256)
mt_gather_kernel(const GatherEntry* __restrict__ tab, const int2* __restrict__ chunks, const float* const* __restrict__ srcs) {
  const int2 ck = chunks[blockIdx.x];
  const GatherEntry e = tab[ck.x];
  const long long base = (long long)ck.y * OPT_CHUNK;
  for (int i = threadIdx.x; i < OPT_CHUNK; i += 256) {
    const long long idx = base + i;
    if (idx >= e.n) break;
    const int m = __ldg(e.map + idx);
    const float v = m < 0 ? 0.f : __ldg(srcs[m >> 24] + (m & 0xFFFFFF));
    if (e.dst_dtype == RL_DT_F32) reinterpret_cast<float*>(e.dst)[idx] = v;
    else if (e.dst_dtype == RL_DT_F16) reinterpret_cast<__half*>(e.dst)[idx] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(e.dst)[idx] = __float2bfloat16(v);
  }
}

bool h_ok(int64_t H) { return H > 0 && H % 128 == 0 && H <= 128 * MAX_V4; }

}  // namespace

extern "C" int rl_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* add_in, float* dx,
                                void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int64_t rows, int64_t H,
                                float eps, float drop_p, uint64_t drop_seed, uint32_t site_in, uint32_t site_out,
                                const uint64_t* drop_counter, void* stream) {
  RL_REQUIRE(dy && x && gamma && (dx || dx_bf16), RL_EINVAL, "rl_layernorm_bwd: null pointer");
  RL_REQUIRE(h_ok(H), RL_EINVAL, "rl_layernorm_bwd: bad H");
  if (rows <= 0) return 0;
  // 8 rows per CTA at least (one per warp); at most 2 CTAs per SM (128 registers: no spills)
  const long long max_ctas = 2LL * rl_num_sms();
  long long ctas = (rows + 7) / 8;
  if (ctas > max_ctas) ctas = max_ctas;
  const int rows_per_cta = (int)((rows + ctas - 1) / ctas);
  ctas = (rows + rows_per_cta - 1) / rows_per_cta;
  const rl::DropSpec din = rl::make_drop(site_in ? drop_p : 0.f, drop_seed, site_in, drop_counter);
  const rl::DropSpec dout = rl::make_drop(site_out ? drop_p : 0.f, drop_seed, site_out, drop_counter);
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaFuncSetAttribute(ln_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 6 * 512);
    cudaFuncSetAttribute(ln_bwd_kernel<MAX_V4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * MAX_V4 * 512);
    configured = true;
  }
  if (H == 768)
    ln_bwd_kernel<6><<<(unsigned)ctas, 256, 8 * 3 * 6 * 512, (cudaStream_t)stream>>>(
        dy, x, gamma, add_in, dx, (__nv_bfloat16*)dx_bf16, dgamma, dbeta, dxsum, rows, (int)H, eps, din, dout, rows_per_cta);
  else
    ln_bwd_kernel<MAX_V4><<<(unsigned)ctas, 256, 8 * 3 * MAX_V4 * 512, (cudaStream_t)stream>>>(
        dy, x, gamma, add_in, dx, (__nv_bfloat16*)dx_bf16, dgamma, dbeta, dxsum, rows, (int)H, eps, din, dout, rows_per_cta);
  return rl_check_launch("rl_layernorm_bwd");
}

extern "C" int rl_colsum_bf16(const void* x, float* out, int64_t rows, int64_t cols, int64_t ld, void* stream) {
  RL_REQUIRE(x && out && cols > 0 && ld >= cols, RL_EINVAL, "rl_colsum_bf16: bad arguments");
  if (rows <= 0) return 0;
  if (cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    const long long col_blocks = (cols + 255) / 256;
    const long long wave = (long long)rl_ctas_per_sm((const void*)colsum_bf16_vec_kernel, 256, 0) * rl_num_sms();
    long long row_chunks = wave / col_blocks;   // one full wave of resident CTAs, never a partial second one
    if (row_chunks > (rows + 63) / 64) row_chunks = (rows + 63) / 64;
    if (row_chunks < 1) row_chunks = 1;
    const int rows_per_cta = (int)((rows + row_chunks - 1) / row_chunks);
    row_chunks = (rows + rows_per_cta - 1) / rows_per_cta;
    dim3 grid((unsigned)col_blocks, (unsigned)row_chunks);
    colsum_bf16_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, rows, (int)cols, ld, rows_per_cta);
    return rl_check_launch("rl_colsum_bf16");
  }
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 1023) / 1024));
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, rows, (int)cols, ld);
  return rl_check_launch("rl_colsum_bf16");
}

extern "C" int rl_gelu_fwd(const void* u, void* h, int64_t n, int32_t dtype, void* stream) {
  RL_REQUIRE(u && h && n >= 0 && n % 8 == 0 && (((uintptr_t)u | (uintptr_t)h) & 15) == 0, RL_EALIGN,
             "rl_gelu_fwd: n must be a multiple of 8 and the pointers 16-byte aligned");
  if (n == 0) return 0;
  const long long n8 = n / 8;
  gelu_fwd_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)u, (__nv_bfloat16*)h, n8,
                                                                                dtype == RL_DT_F16);
  return rl_check_launch("rl_gelu_fwd");
}

extern "C" int rl_gelu_bwd_colsum(void* t, const void* u, float* dbias, int64_t rows, int64_t cols, int64_t ld,
                                  int32_t u_dtype, void* stream) {
  RL_REQUIRE(t && u && cols > 0 && ld >= cols && cols % 8 == 0 && ld % 8 == 0 && (((uintptr_t)t | (uintptr_t)u) & 15) == 0,
             RL_EALIGN, "rl_gelu_bwd_colsum: cols / ld must be multiples of 8 and the pointers 16-byte aligned");
  if (rows <= 0) return 0;
  const long long col_blocks = (cols + 255) / 256;
  const long long wave = (long long)rl_ctas_per_sm((const void*)gelu_bwd_colsum_kernel, 256, 0) * rl_num_sms();
  long long row_chunks = wave / col_blocks;   // one full wave of resident CTAs, never a partial second one
  if (row_chunks > (rows + 63) / 64) row_chunks = (rows + 63) / 64;
  if (row_chunks < 1) row_chunks = 1;
  const int rows_per_cta = (int)((rows + row_chunks - 1) / row_chunks);
  row_chunks = (rows + rows_per_cta - 1) / rows_per_cta;
  dim3 grid((unsigned)col_blocks, (unsigned)row_chunks);
  gelu_bwd_colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)t, (const __nv_bfloat16*)u, dbias, rows, (int)cols,
                                                                ld, rows_per_cta, u_dtype == RL_DT_F16);
  return rl_check_launch("rl_gelu_bwd_colsum");
}

extern "C" int rl_masked_ce_bwd(const void* logits, int32_t logits_dtype, const int64_t* tgt, const int64_t* loss_mask,
                                const float* row_lse, const float* count, const float* gscale, void* dlogits, int64_t rows,
                                int64_t V, int64_t ld, int64_t ldd, void* stream) {
  RL_REQUIRE(logits && tgt && loss_mask && row_lse && count && dlogits, RL_EINVAL, "rl_masked_ce_bwd: null pointer");
  RL_REQUIRE(ldd >= V && ldd % 8 == 0 && ((uintptr_t)dlogits & 15) == 0, RL_EALIGN, "rl_masked_ce_bwd: ldd must be >= V, %%8");
  if (rows <= 0) return 0;
  RL_REQUIRE(logits_dtype == RL_DT_F32 || logits_dtype == RL_DT_F16, RL_EINVAL, "rl_masked_ce_bwd: logits must be f32 or fp16");
  if (logits_dtype == RL_DT_F16)
    ce_bwd_kernel<true><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)tgt, (const long long*)loss_mask,
                                                                         row_lse, count, gscale, (__nv_bfloat16*)dlogits, (int)V, ld, ldd);
  else
    ce_bwd_kernel<false><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)tgt, (const long long*)loss_mask,
                                                                          row_lse, count, gscale, (__nv_bfloat16*)dlogits, (int)V, ld, ldd);
  return rl_check_launch("rl_masked_ce_bwd");
}

extern "C" int rl_embed_bwd(const float* de, const int64_t* ids, float* dword, float* dpos, int64_t rows, int64_t L,
                            int64_t H, int32_t pos_mode, void* stream) {
  RL_REQUIRE(de && (dword == nullptr || ids) && h_ok(H) && L > 0, RL_EINVAL, "rl_embed_bwd: bad arguments");
  if (rows <= 0) return 0;
  embed_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(de, (const long long*)ids, dword, dpos, rows,
                                                                               (int)L, (int)H, pos_mode);
  return rl_check_launch("rl_embed_bwd");
}

extern "C" int rl_gate_fuse_bwd(const float* dhid, const float* m0, const float* m1, const float* m2, int32_t num_modal,
                                const int64_t* mask, const float* gates, const float* gate_w, float* dm0, float* dm1,
                                float* dm2, float* dgate_w, float* dgate_b, float* ws, int64_t B, int64_t L, int64_t H,
                                void* stream) {
  // ws: f32 scratch of size B*L*3 (dz) + 2*B*H (dmean, mean)
  RL_REQUIRE(dhid && m0 && dm0 && gates && gate_w && dgate_w && dgate_b && ws && mask, RL_EINVAL, "rl_gate_fuse_bwd: null pointer");
  RL_REQUIRE(num_modal >= 1 && num_modal <= 3 && h_ok(H), RL_EINVAL, "rl_gate_fuse_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = B * L;
  float* dz = ws;
  float* dmean = ws + rows * 3;
  float* mean = dmean + B * H;
  cudaMemsetAsync(dmean, 0, (size_t)B * H * sizeof(float), st);
  gate_bwd_token_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(dhid, m0, m1, m2, gates, gate_w, dm0, dm1, dm2, dz, dmean,
                                                                  dgate_b, rows, (int)L, (int)H, num_modal);
  int rc = rl_check_launch("rl_gate_fuse_bwd(token)");
  if (rc) return rc;
  gate_bwd_mean_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(dm0, dmean, (const long long*)mask, rows, (int)L, (int)H);
  masked_mean_kernel<<<(unsigned)B, 256, 0, st>>>(m0, (const long long*)mask, mean, (int)L, (int)H);
  const float* srcs[3] = {m0, m1, m2};
  dim3 grid((unsigned)(H / 32), (unsigned)((rows + 511) / 512));
  for (int j = 0; j <= num_modal; ++j)
    gate_bwd_weight_kernel<<<grid, 256, 0, st>>>(dz, j < num_modal ? srcs[j] : nullptr, j < num_modal ? nullptr : mean, dgate_w,
                                                rows, (int)L, (int)H, num_modal, j);
  return rl_check_launch("rl_gate_fuse_bwd");
}

extern "C" int rl_mt_sumsq(const void* table, const void* chunks, int64_t num_chunks, float* out, float* partials_ws,
                           int64_t ws_floats, void* stream) {
  RL_REQUIRE(table && chunks && out && partials_ws, RL_EINVAL, "rl_mt_sumsq: null pointer");
  long long grid = 8LL * rl_num_sms();
  if (grid > num_chunks) grid = num_chunks;
  if (grid > ws_floats) grid = ws_floats;          // fewer CTAs, same result up to fp32 partial rounding
  RL_REQUIRE(num_chunks <= 0 || grid >= 1, RL_EINVAL, "rl_mt_sumsq: partials workspace is empty");
  if (num_chunks <= 0) {
    cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)stream);
    return rl_check_launch("rl_mt_sumsq");
  }
  mt_sumsq_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const TensorEntry*)table, (const int2*)chunks, partials_ws,
                                                                  num_chunks);
  mt_sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials_ws, (int)grid, out);
  return rl_check_launch("rl_mt_sumsq");
}

extern "C" int rl_mt_adamw(const void* table, const void* chunks, int64_t num_chunks, const float* sumsq, float max_norm,
                           float lr, float beta1, float beta2, float eps, float bias_corr1, float bias_corr2,
                           float grad_div, void* stream) {
  RL_REQUIRE(table && chunks && sumsq, RL_EINVAL, "rl_mt_adamw: null pointer");
  if (num_chunks <= 0) return 0;
  mt_adamw_kernel<<<(unsigned)num_chunks, 256, 0, (cudaStream_t)stream>>>((const TensorEntry*)table, (const int2*)chunks, sumsq,
                                                                        max_norm, lr, beta1, beta2, eps, bias_corr1,
                                                                        bias_corr2, grad_div, nullptr);
  return rl_check_launch("rl_mt_adamw");
}

extern "C" int rl_mt_gather(const void* table, const void* chunks, int64_t num_chunks, const void* srcs, void* stream) {
  RL_REQUIRE(table && chunks && srcs, RL_EINVAL, "rl_mt_gather: null pointer");
  if (num_chunks <= 0) return 0;
  mt_gather_kernel<<<(unsigned)num_chunks, 256, 0, (cudaStream_t)stream>>>((const GatherEntry*)table, (const int2*)chunks,
                                                                         (const float* const*)srcs);
  return rl_check_launch("rl_mt_gather");
}

extern "C" int rl_mt_adamw_dev(const void* table, const void* chunks, int64_t num_chunks, const float* sumsq, float max_norm,
                               const float* hyper, float beta1, float beta2, float eps, float grad_div, void* stream) {
  RL_REQUIRE(table && chunks && sumsq && hyper, RL_EINVAL, "rl_mt_adamw_dev: null pointer");
  if (num_chunks <= 0) return 0;
  mt_adamw_kernel<<<(unsigned)num_chunks, 256, 0, (cudaStream_t)stream>>>((const TensorEntry*)table, (const int2*)chunks, sumsq,
                                                                        max_norm, 0.f, beta1, beta2, eps, 1.f, 1.f, grad_div, hyper);
  return rl_check_launch("rl_mt_adamw_dev");
}
