// Training-mode pieces of CharResNet (src/char_cnn.py:9-55): BatchNorm2d with batch statistics (forward and
// backward), the im2col gathers that turn conv weight gradients into plain tcgen05 GEMMs, and the raw
// (pre-BatchNorm) glyph stem.  Convolutions themselves (forward, data gradients) run through rl_gemm_bf16.
//
// Layouts: activations are NHWC bf16.  "plain" rows are (img, h, w); "parity-split" rows are
// [img][h&1][w&1][h/2][w/2] (the input layout of a stride-2 conv).  Raw conv outputs are always plain.
#include "common.cuh"

namespace {

__device__ __forceinline__ long long split_row(long long row, int hw_shift, int w_shift) {
  const int hw = 1 << hw_shift, w = 1 << w_shift;
  const long long img = row >> hw_shift;
  const int pix = (int)(row & (hw - 1));
  const int oh = pix >> w_shift, ow = pix & (w - 1);
  const int h2 = (hw >> w_shift) >> 1, w2 = w >> 1;
  return ((img * 4 + (oh & 1) * 2 + (ow & 1)) * h2 + (oh >> 1)) * w2 + (ow >> 1);
}

__device__ __forceinline__ float load_raw(const void* x, int x_f32, long long i) {
  return x_f32 ? reinterpret_cast<const float*>(x)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
}
__device__ __forceinline__ void load_raw8(const void* x, int x_f32, long long i, float (&v)[8]) {
  if (x_f32) {
    const float4 a = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i)[0];
    const float4 b = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + i);
    v[0] = rl::bf16_lo(a.x); v[1] = rl::bf16_hi(a.x); v[2] = rl::bf16_lo(a.y); v[3] = rl::bf16_hi(a.y);
    v[4] = rl::bf16_lo(a.z); v[5] = rl::bf16_hi(a.z); v[6] = rl::bf16_lo(a.w); v[7] = rl::bf16_hi(a.w);
  }
}

// ---- BatchNorm forward: per-channel sum / sum of squares of a raw conv output [M, C] (f32 or bf16) ----------
// (raw conv outputs are kept in f32 in training: batch statistics over few samples cancel catastrophically in bf16)
__global__ void __launch_bounds__(256)
bn_stats_kernel(const void* __restrict__ x, int x_f32, float* __restrict__ sums, long long M, int C, long long ld) {
  __shared__ float s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * 2048;
  float a = 0.f, b = 0.f;
  if (col < C)
    for (long long r = r0 + ry; r < r0 + 2048 && r < M; r += 8) {
      const float v = load_raw(x, x_f32, r * ld + col);
      a += v;
      b += v * v;
    }
  s1[ry][cx] = a;
  s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && col < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t1 += s1[i][cx];
      t2 += s2[i][cx];
    }
    atomicAdd(sums + col, t1);
    atomicAdd(sums + C + col, t2);
  }
}

// mean / biased var -> scale = gamma*rstd, shift = beta - mean*scale; running stats with momentum (unbiased var)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, long long M, int C, float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
  if (c >= C) return;
  const float mean = sums[c] / (float)M;
  float var = sums[C + c] / (float)M - mean * mean;
  var = fmaxf(var, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * ((float)M / (float)(M > 1 ? M - 1 : 1));
  }
}

// out = act(x1*scale1 + shift1 [+ x2*scale2 + shift2]); 8 channels per thread; optional parity-split row remap
__global__ void __launch_bounds__(256)
bn_apply_kernel(const void* __restrict__ x1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                const void* __restrict__ x2, const float* __restrict__ sc2, const float* __restrict__ sh2, int x_f32,
                void* __restrict__ out, int out_f32, int relu, long long M, int C, int remap, int hw_shift, int w_shift) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over M * C/8
  const int c8 = C >> 3;
  if (idx >= M * c8) return;
  const long long row = idx / c8;
  const int c = (int)(idx - row * c8) * 8;
  float v[8];
  load_raw8(x1, x_f32, row * C + c, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], __ldg(sc1 + c + j), __ldg(sh1 + c + j));
  if (x2) {
    float w[8];
    load_raw8(x2, x_f32, row * C + c, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += fmaf(w[j], __ldg(sc2 + c + j), __ldg(sh2 + c + j));
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  const long long orow = remap ? split_row(row, hw_shift, w_shift) : row;
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * C + c);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * C + c) =
        make_uint4(rl::pack_bf16(v[0], v[1]), rl::pack_bf16(v[2], v[3]), rl::pack_bf16(v[4], v[5]), rl::pack_bf16(v[6], v[7]));
  }
}

// ---- BatchNorm backward -----------------------------------------------------------------------------------
// dy (f32 or bf16, rows possibly parity-split) masked by the ReLU that followed (act_out > 0, same rows as dy);
// x = raw conv output (plain rows).  reduce: dbeta += sum dy, dgamma += sum dy*xhat.
// apply: dx = gamma*rstd*(dy - dbeta/M - xhat*dgamma/M) as bf16, written at column offset into a wider matrix.
__device__ __forceinline__ float load_dy(const void* dy, int dy_f32, long long i) {
  return dy_f32 ? reinterpret_cast<const float*>(dy)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy)[i]);
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ act_out, int act_f32,
                     const void* __restrict__ x, int x_f32, const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dbeta, float* __restrict__ dgamma, long long M, int C, int remap, int hw_shift,
                     int w_shift) {
  __shared__ float s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * 2048;
  float a = 0.f, b = 0.f;
  if (col < C) {
    const float mu = mean[col], rs = rstd[col];
    for (long long r = r0 + ry; r < r0 + 2048 && r < M; r += 8) {
      const long long dr = remap ? split_row(r, hw_shift, w_shift) : r;
      float g = load_dy(dy, dy_f32, dr * C + col);
      if (act_out && !(load_dy(act_out, act_f32, dr * C + col) > 0.f)) g = 0.f;
      const float xh = (load_raw(x, x_f32, r * C + col) - mu) * rs;
      a += g;
      b += g * xh;
    }
  }
  s1[ry][cx] = a;
  s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && col < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t1 += s1[i][cx];
      t2 += s2[i][cx];
    }
    atomicAdd(dbeta + col, t1);
    atomicAdd(dgamma + col, t2);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ act_out, int act_f32,
                    const void* __restrict__ x, int x_f32, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                    __nv_bfloat16* __restrict__ dx, long long ldx, long long M, int C, int remap, int hw_shift, int w_shift) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over M * C/8
  const int c8 = C >> 3;
  if (idx >= M * c8) return;
  const long long row = idx / c8;
  const int c = (int)(idx - row * c8) * 8;
  const long long dr = remap ? split_row(row, hw_shift, w_shift) : row;
  float xv[8];
  load_raw8(x, x_f32, row * C + c, xv);
  const float invM = 1.0f / (float)M;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float g = load_dy(dy, dy_f32, dr * C + c + j);
    if (act_out && !(load_dy(act_out, act_f32, dr * C + c + j) > 0.f)) g = 0.f;
    const float rs = __ldg(rstd + c + j);
    const float xh = (xv[j] - __ldg(mean + c + j)) * rs;
    o[j] = __ldg(gamma + c + j) * rs * (g - __ldg(dbeta + c + j) * invM - xh * __ldg(dgamma + c + j) * invM);
  }
  *reinterpret_cast<uint4*>(dx + row * ldx + c) =
      make_uint4(rl::pack_bf16(o[0], o[1]), rl::pack_bf16(o[2], o[3]), rl::pack_bf16(o[4], o[5]), rl::pack_bf16(o[6], o[7]));
}


// ---- vectorised variants (C % 8 == 0, dense rows): 8 channels per lane, a warp spans min(C, 256) channels --------
// lanes_per_row = min(32, C/8) (a power of two); a warp covers 32/lanes_per_row consecutive rows per iteration and the
// CTA's 8 warps stride through the CTA's row slab.  Column partials stay in registers for the whole slab; one smem
// reduction over the 8 warps and one atomic per (lane-column) at the end.
__device__ __forceinline__ void load8(const void* p, int is_f32, long long i, float (&v)[8]) { load_raw8(p, is_f32, i, v); }

__global__ void __launch_bounds__(256)
bn_stats_vec_kernel(const void* __restrict__ x, int x_f32, float* __restrict__ sums, long long M, int C, int lpr_shift,
                    int rows_per_cta) {
  __shared__ float s1[8][256], s2[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpr = 1 << lpr_shift, rpw = 32 >> lpr_shift;
  const int col = blockIdx.x * 256 + (lane & (lpr - 1)) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  const int step = 8 * rpw;
  long long r = r0 + warp * rpw + (lane >> lpr_shift);
  for (; r + 3 * step < r1; r += 4 * step) {   // four row groups in flight (64 B per lane even with bf16 rows)
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_raw8(x, x_f32, (r + u * step) * C + col, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += v[u][j];
        b[j] = fmaf(v[u][j], v[u][j], b[j]);
      }
  }
  for (; r < r1; r += step) {
    float v[8];
    load_raw8(x, x_f32, r * C + col, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] += v[j];
      b[j] = fmaf(v[j], v[j], b[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[warp][lane * 8 + j] = a[j];
    s2[warp][lane * 8 + j] = b[j];
  }
  __syncthreads();
  const int t = threadIdx.x;
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    t1 += s1[w][t];
    t2 += s2[w][t];
  }
  const int c = blockIdx.x * 256 + ((t >> 3) & (lpr - 1)) * 8 + (t & 7);
  atomicAdd(sums + c, t1);
  atomicAdd(sums + C + c, t2);
}

struct BnBranch {          // one BatchNorm whose output fed the (shared) ReLU
  const void* x;           // raw conv output [M, C], plain rows; f32, or bf16 when x_f32 == 0
  int x_f32;
  const float* mean;
  const float* rstd;
  const float* gamma;
  float* dbeta;
  float* dgamma;
  __nv_bfloat16* dx;       // [M, ldx] bf16, plain rows
  long long ldx;
  const float* sc;         // optional: the forward's scale / shift (rl_bn_finalize).  When branch 0 carries them the ReLU mask
  const float* sh;         // is re-derived as (x*sc + sh [+ x2*sc2 + sh2]) > 0 — rl_bn_apply's arithmetic — and act_out is not read
};

// ReLU mask of 8 channels from the raw conv outputs (same fmaf / add order as bn_apply_vec_kernel)
template <int NB>
__device__ __forceinline__ void relu_mask8(const float (&x0)[8], const float (&x1)[8], const float* s_m, int lc, float (&g)[8]) {
  float a[8], b[8], y[8];
  *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(s_m + lc);
  *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(s_m + lc + 4);
  *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(s_m + 256 + lc);
  *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(s_m + 256 + lc + 4);
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = fmaf(x0[j], a[j], b[j]);
  if (NB == 2) {
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(s_m + 512 + lc);
    *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(s_m + 512 + lc + 4);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(s_m + 768 + lc);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(s_m + 768 + lc + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] += fmaf(x1[j], a[j], b[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = y[j] > 0.f ? g[j] : 0.f;
}

// the forward scale / shift of the CTA's 256-column block -> s_m[NB][2][256]
template <int NB>
__device__ __forceinline__ void fill_mask_coef(float* s_m, const BnBranch& b0, const BnBranch& b1, int C) {
  const int ct = blockIdx.x * 256 + threadIdx.x;
  s_m[threadIdx.x] = ct < C ? b0.sc[ct] : 0.f;
  s_m[256 + threadIdx.x] = ct < C ? b0.sh[ct] : 0.f;
  if (NB == 2) {
    s_m[512 + threadIdx.x] = ct < C ? b1.sc[ct] : 0.f;
    s_m[768 + threadIdx.x] = ct < C ? b1.sh[ct] : 0.f;
  }
}

// reduce: dbeta = sum g, dgamma = rstd * (sum g*x - mean * sum g) — mean/rstd are applied once at the end, so the loop
// carries only 8 * (1 + NB) accumulators (4 CTAs per SM); two row groups in flight per warp.
template <int NB>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_vec_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ act_out, int act_f32, BnBranch b0,
                         BnBranch b1, long long M, int C, int lpr_shift, int rows_per_cta, int remap, int hw_shift, int w_shift) {
  __shared__ float sm[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpr = 1 << lpr_shift, rpw = 32 >> lpr_shift;
  const int col = blockIdx.x * 256 + (lane & (lpr - 1)) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float sg[8], sx[NB][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sg[j] = 0.f;
    sx[0][j] = 0.f;
    if (NB == 2) sx[NB - 1][j] = 0.f;
  }
  const int step = 8 * rpw;
  long long r = r0 + warp * rpw + (lane >> lpr_shift);
  __shared__ __align__(16) float s_m[NB * 2 * 256];
  const bool recompute = b0.sc != nullptr;      // kernel-uniform
  const int lc = (lane & (lpr - 1)) * 8;
  if (recompute) {
    fill_mask_coef<NB>(s_m, b0, b1, C);
    __syncthreads();
    // mask from the raw conv outputs: three (two) streams instead of four (three), two row groups in flight
    constexpr int U2 = 2;
    for (; r + (U2 - 1) * step < r1; r += U2 * step) {
      float g[U2][8], x0[U2][8], x1[U2][8];
#pragma unroll
      for (int u = 0; u < U2; ++u) {
        const long long rr = r + u * step;
        const long long dr = remap ? split_row(rr, hw_shift, w_shift) : rr;
        load8(dy, dy_f32, dr * C + col, g[u]);
        load_raw8(b0.x, b0.x_f32, rr * C + col, x0[u]);
        if (NB == 2) load_raw8(b1.x, b1.x_f32, rr * C + col, x1[u]);
      }
#pragma unroll
      for (int u = 0; u < U2; ++u) {
        relu_mask8<NB>(x0[u], x1[u], s_m, lc, g[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sg[j] += g[u][j];
          sx[0][j] = fmaf(g[u][j], x0[u][j], sx[0][j]);
          if (NB == 2) sx[NB - 1][j] = fmaf(g[u][j], x1[u][j], sx[NB - 1][j]);
        }
      }
    }
    for (; r < r1; r += step) {
      const long long dr = remap ? split_row(r, hw_shift, w_shift) : r;
      float g[8], x0[8], x1[8];
      load8(dy, dy_f32, dr * C + col, g);
      load_raw8(b0.x, b0.x_f32, r * C + col, x0);
      if (NB == 2) load_raw8(b1.x, b1.x_f32, r * C + col, x1);
      relu_mask8<NB>(x0, x1, s_m, lc, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[j] += g[j];
        sx[0][j] = fmaf(g[j], x0[j], sx[0][j]);
        if (NB == 2) sx[NB - 1][j] = fmaf(g[j], x1[j], sx[NB - 1][j]);
      }
    }
  }
  constexpr int U = 4;   // row groups in flight per warp (bf16 rows are only 16 B per lane and tensor)
  for (; r + (U - 1) * step < r1; r += U * step) {
    float g[U][8], xv[U][8];
    long long dr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * step;
      dr[u] = remap ? split_row(rr, hw_shift, w_shift) : rr;
      load8(dy, dy_f32, dr[u] * C + col, g[u]);
      load_raw8(b0.x, b0.x_f32, rr * C + col, xv[u]);
    }
    if (act_out) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a[8];
        load8(act_out, act_f32, dr[u] * C + col, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[u][j] = a[j] > 0.f ? g[u][j] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[j] += g[u][j];
        sx[0][j] = fmaf(g[u][j], xv[u][j], sx[0][j]);
      }
    if (NB == 2) {
#pragma unroll
      for (int u = 0; u < U; ++u) load_raw8(b1.x, b1.x_f32, (r + u * step) * C + col, xv[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) sx[NB - 1][j] = fmaf(g[u][j], xv[u][j], sx[NB - 1][j]);
    }
  }
  for (; r < r1; r += step) {
    const long long dr = remap ? split_row(r, hw_shift, w_shift) : r;
    float g[8], xv[8];
    load8(dy, dy_f32, dr * C + col, g);
    if (act_out) {
      float a[8];
      load8(act_out, act_f32, dr * C + col, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
    }
    load_raw8(b0.x, b0.x_f32, r * C + col, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sg[j] += g[j];
      sx[0][j] = fmaf(g[j], xv[j], sx[0][j]);
    }
    if (NB == 2) {
      load_raw8(b1.x, b1.x_f32, r * C + col, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) sx[NB - 1][j] = fmaf(g[j], xv[j], sx[NB - 1][j]);
    }
  }
  const int t = threadIdx.x;
  const int c = blockIdx.x * 256 + ((t >> 3) & (lpr - 1)) * 8 + (t & 7);
  float tot_g = 0.f;
#pragma unroll
  for (int q = 0; q < 1 + NB; ++q) {
    if (q) __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[warp][lane * 8 + j] = q == 0 ? sg[j] : sx[q - 1][j];
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sm[w][t];
    if (q == 0) {
      tot_g = tot;
      atomicAdd(b0.dbeta + c, tot);
      if (NB == 2) atomicAdd(b1.dbeta + c, tot);
    } else {
      const BnBranch& b = q == 1 ? b0 : b1;
      atomicAdd(b.dgamma + c, b.rstd[c] * (tot - b.mean[c] * tot_g));
    }
  }
}

// apply: dx = A*g + B*x + D per channel with A = gamma*rstd, B = -gamma*rstd^2*dgamma/M,
// D = gamma*rstd*(mean*rstd*dgamma/M - dbeta/M).  The CTA derives the coefficients of its 256-column block once into
// shared memory (so they cost no registers) and every thread walks the rows of the slab two row groups at a time: with
// bf16 raw conv outputs a single group is only 64 bytes in flight per lane, too little to cover the HBM latency.
template <int NB>
__global__ void __launch_bounds__(256, NB == 2 ? 2 : 3)
bn_bwd_apply_vec_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ act_out, int act_f32, BnBranch b0,
                        BnBranch b1, long long M, int C, int lpr_shift, int rows_per_cta, int remap, int hw_shift, int w_shift) {
  __shared__ __align__(16) float s_coef[NB][3][256];
  __shared__ __align__(16) float s_m[NB * 2 * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpr = 1 << lpr_shift, rpw = 32 >> lpr_shift;
  const bool recompute = b0.sc != nullptr;      // kernel-uniform
  if (recompute) fill_mask_coef<NB>(s_m, b0, b1, C);
  {
    const int ct = blockIdx.x * 256 + threadIdx.x;
    const float invM = 1.0f / (float)M;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const BnBranch& b = q == 0 ? b0 : b1;
      float A = 0.f, Bc = 0.f, D = 0.f;
      if (ct < C) {
        const float rs = b.rstd[ct], mu = b.mean[ct], gm = b.gamma[ct];
        const float dg = b.dgamma[ct] * invM, db = b.dbeta[ct] * invM;
        A = gm * rs;
        Bc = -gm * rs * rs * dg;
        D = gm * rs * (mu * rs * dg - db);
      }
      s_coef[q][0][threadIdx.x] = A;
      s_coef[q][1][threadIdx.x] = Bc;
      s_coef[q][2][threadIdx.x] = D;
    }
  }
  __syncthreads();
  const int lc = (lane & (lpr - 1)) * 8;   // column inside the CTA's 256-column block
  const int col = blockIdx.x * 256 + lc;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  const int step = 8 * rpw;
  for (long long r = r0 + warp * rpw + (lane >> lpr_shift); r < r1; r += 2 * step) {
    const long long rb = r + step;
    const bool two = rb < r1;
    const long long da = remap ? split_row(r, hw_shift, w_shift) : r;
    const long long db = two ? (remap ? split_row(rb, hw_shift, w_shift) : rb) : da;
    const long long xb_row = two ? rb : r;
    float g[2][8], xv[2][NB][8];
    load8(dy, dy_f32, da * C + col, g[0]);
    load8(dy, dy_f32, db * C + col, g[1]);
    load_raw8(b0.x, b0.x_f32, r * C + col, xv[0][0]);
    load_raw8(b0.x, b0.x_f32, xb_row * C + col, xv[1][0]);
    if (NB == 2) {
      load_raw8(b1.x, b1.x_f32, r * C + col, xv[0][NB - 1]);
      load_raw8(b1.x, b1.x_f32, xb_row * C + col, xv[1][NB - 1]);
    }
    if (recompute) {
      relu_mask8<NB>(xv[0][0], xv[0][NB - 1], s_m, lc, g[0]);
      relu_mask8<NB>(xv[1][0], xv[1][NB - 1], s_m, lc, g[1]);
    } else if (act_out) {
      float a0[8], a1[8];
      load8(act_out, act_f32, da * C + col, a0);
      load8(act_out, act_f32, db * C + col, a1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[0][j] = a0[j] > 0.f ? g[0][j] : 0.f;
        g[1][j] = a1[j] > 0.f ? g[1][j] : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const BnBranch& b = q == 0 ? b0 : b1;
      float A[8], Bc[8], D[8];
      *reinterpret_cast<float4*>(A) = *reinterpret_cast<const float4*>(&s_coef[q][0][lc]);
      *reinterpret_cast<float4*>(A + 4) = *reinterpret_cast<const float4*>(&s_coef[q][0][lc + 4]);
      *reinterpret_cast<float4*>(Bc) = *reinterpret_cast<const float4*>(&s_coef[q][1][lc]);
      *reinterpret_cast<float4*>(Bc + 4) = *reinterpret_cast<const float4*>(&s_coef[q][1][lc + 4]);
      *reinterpret_cast<float4*>(D) = *reinterpret_cast<const float4*>(&s_coef[q][2][lc]);
      *reinterpret_cast<float4*>(D + 4) = *reinterpret_cast<const float4*>(&s_coef[q][2][lc + 4]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], g[h][j], fmaf(Bc[j], xv[h][q][j], D[j]));
        *reinterpret_cast<uint4*>(b.dx + (h ? rb : r) * b.ldx + col) =
            make_uint4(rl::pack_bf16(o[0], o[1]), rl::pack_bf16(o[2], o[3]), rl::pack_bf16(o[4], o[5]), rl::pack_bf16(o[6], o[7]));
      }
    }
  }
}

// vector-path geometry for a [M, C] matrix: lanes per row (log2) or -1 when C does not fit the scheme
int vec_lpr_shift(long long C) {
  if (C % 8) return -1;
  if (C >= 256) return (C % 256 == 0) ? 5 : -1;
  const int l = (int)(C / 8);
  if (l & (l - 1)) return -1;
  int s = 0;
  while ((1 << s) < l) ++s;
  return s;
}

void vec_grid(long long M, long long C, int lpr_shift, dim3* grid, int* rows_per_cta, const void* func) {
  const long long col_blocks = (C + 255) / 256;
  const int rows_per_iter = 8 * (32 >> lpr_shift);
  // one full wave of resident CTAs (occupancy x SMs): a handful of CTAs beyond it would cost a whole extra pass
  long long chunks = (long long)rl_ctas_per_sm(func, 256, 0) * rl_num_sms() / col_blocks;
  const long long max_chunks = (M + 4 * rows_per_iter - 1) / (4 * rows_per_iter);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  long long rpc = (M + chunks - 1) / chunks;
  rpc = (rpc + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
  chunks = (M + rpc - 1) / rpc;
  *rows_per_cta = (int)rpc;
  *grid = dim3((unsigned)col_blocks, (unsigned)chunks);
}

// forward apply, vector path: out = [relu](x1*sc1 + sh1 [+ x2*sc2 + sh2]); the CTA's 256-column scale/shift block sits in
// shared memory, every thread owns 8 channels and walks the rows of its slab two row groups at a time
template <int NX>
__global__ void __launch_bounds__(256, NX == 2 ? 3 : 4)
bn_apply_vec_kernel(const void* __restrict__ x1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                    const void* __restrict__ x2, const float* __restrict__ sc2, const float* __restrict__ sh2, int x_f32,
                    void* __restrict__ out, int out_f32, int relu, long long M, int C, int lpr_shift, int rows_per_cta, int remap,
                    int hw_shift, int w_shift) {
  __shared__ __align__(16) float s_c[NX][2][256];
  {
    const int ct = blockIdx.x * 256 + threadIdx.x;
    s_c[0][0][threadIdx.x] = ct < C ? sc1[ct] : 0.f;
    s_c[0][1][threadIdx.x] = ct < C ? sh1[ct] : 0.f;
    if (NX == 2) {
      s_c[NX - 1][0][threadIdx.x] = ct < C ? sc2[ct] : 0.f;
      s_c[NX - 1][1][threadIdx.x] = ct < C ? sh2[ct] : 0.f;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpr = 1 << lpr_shift, rpw = 32 >> lpr_shift;
  const int lc = (lane & (lpr - 1)) * 8;
  const int col = blockIdx.x * 256 + lc;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  const int step = 8 * rpw;
  for (long long r = r0 + warp * rpw + (lane >> lpr_shift); r < r1; r += 2 * step) {
    const long long rb = r + step;
    const bool two = rb < r1;
    const long long rb_ld = two ? rb : r;
    float v[2][8], w[2][8];
    load_raw8(x1, x_f32, r * C + col, v[0]);
    load_raw8(x1, x_f32, rb_ld * C + col, v[1]);
    if (NX == 2) {
      load_raw8(x2, x_f32, r * C + col, w[0]);
      load_raw8(x2, x_f32, rb_ld * C + col, w[1]);
    }
    float a[8], b[8];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&s_c[0][0][lc]);
    *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&s_c[0][0][lc + 4]);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&s_c[0][1][lc]);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&s_c[0][1][lc + 4]);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[h][j] = fmaf(v[h][j], a[j], b[j]);
    if (NX == 2) {
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&s_c[NX - 1][0][lc]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&s_c[NX - 1][0][lc + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&s_c[NX - 1][1][lc]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&s_c[NX - 1][1][lc + 4]);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[h][j] += fmaf(w[h][j], a[j], b[j]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[h][j] = fmaxf(v[h][j], 0.f);
      }
      const long long row = h ? rb : r;
      const long long orow = remap ? split_row(row, hw_shift, w_shift) : row;
      if (out_f32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * C + col);
        o[0] = make_float4(v[h][0], v[h][1], v[h][2], v[h][3]);
        o[1] = make_float4(v[h][4], v[h][5], v[h][6], v[h][7]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * C + col) =
            make_uint4(rl::pack_bf16(v[h][0], v[h][1]), rl::pack_bf16(v[h][2], v[h][3]), rl::pack_bf16(v[h][4], v[h][5]),
                       rl::pack_bf16(v[h][6], v[h][7]));
      }
    }
  }
}

// ---- im2col for weight gradients: col[m, t*C + ci] = x[img, plane_t, oh+dh_t, ow+dw_t, ci] (0 outside) ------
// so that dW[co, t, ci] = sum_m dY[m, co] * col[m, t*C + ci] is one plain (split-K) GEMM; the small result is
// permuted to the reference's [co, ci, kh, kw] layout afterwards.
struct TapTable {
  int n;
  signed char dw[12], dh[12], plane[12];
};

__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, long long M, int C, int W, int H, int P,
              int hw_shift, int w_shift, TapTable taps) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over M * C/8
  const int c8 = C >> 3;
  if (idx >= M * c8) return;
  const long long row = idx / c8;
  const int c = (int)(idx - row * c8) * 8;
  const long long img = row >> hw_shift;
  const int pix = (int)(row & ((1 << hw_shift) - 1));
  const int oh = pix >> w_shift, ow = pix & ((1 << w_shift) - 1);
  const int T = taps.n;
  __nv_bfloat16* dst = col + row * (long long)(C * T) + c;   // tap-major columns: col[m, t*C + ci], 16-byte copies
  for (int t = 0; t < T; ++t) {
    const int ih = oh + taps.dh[t], iw = ow + taps.dw[t];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
      v = *reinterpret_cast<const uint4*>(x + ((((img * P + taps.plane[t]) * H + ih) * W + iw) * (long long)C + c));
    *reinterpret_cast<uint4*>(dst + (long long)t * C) = v;
  }
}

// glyph patches for the block-1 weight gradients: col1[m, c*9 + kh*3 + kw] (32 columns, 27 used) and the
// centre pixel colsc[m, c] (8 columns, C used); m = (img, oh, ow) over the 16x16 output map
template <int C>
__global__ void __launch_bounds__(256)
glyph_im2col_kernel(const float* __restrict__ glyphs, const long long* __restrict__ ids, __nv_bfloat16* __restrict__ col1,
                    __nv_bfloat16* __restrict__ colsc) {
  __shared__ float s_img[C * 1024];
  const int tid = threadIdx.x;
  const long long img = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(glyphs + ids[img] * (long long)(C * 1024));
  for (int i = tid; i < C * 256; i += 256) reinterpret_cast<float4*>(s_img)[i] = __ldg(src + i);
  __syncthreads();
  const int oh = tid >> 4, ow = tid & 15;
  float a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ih = 2 * oh + kh - 1, iw = 2 * ow + kw - 1;
        if (ih >= 0 && iw >= 0) a[c * 9 + kh * 3 + kw] = s_img[c * 1024 + ih * 32 + iw];
      }
  uint4* d1 = reinterpret_cast<uint4*>(col1 + (img * 256 + tid) * 32);
#pragma unroll
  for (int g = 0; g < 4; ++g)
    d1[g] = make_uint4(rl::pack_bf16(a[8 * g], a[8 * g + 1]), rl::pack_bf16(a[8 * g + 2], a[8 * g + 3]),
                       rl::pack_bf16(a[8 * g + 4], a[8 * g + 5]), rl::pack_bf16(a[8 * g + 6], a[8 * g + 7]));
  if (colsc) {
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = k < C ? s_img[k * 1024 + (2 * oh) * 32 + 2 * ow] : 0.f;
    *reinterpret_cast<uint4*>(colsc + (img * 256 + tid) * 8) =
        make_uint4(rl::pack_bf16(s[0], s[1]), rl::pack_bf16(s[2], s[3]), rl::pack_bf16(s[4], s[5]), rl::pack_bf16(s[6], s[7]));
  }
}

int ilog2x(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return ((1 << s) == v) ? s : -1;
}

}  // namespace

extern "C" int rl_bn_stats(const void* x, int32_t x_dtype, float* sums, int64_t M, int64_t C, int64_t ld, void* stream) {
  RL_REQUIRE(x && sums && M > 0 && C > 0 && ld >= C, RL_EINVAL, "rl_bn_stats: bad arguments");
  const int ls = vec_lpr_shift(C);
  if (ld == C && ls >= 0 && ((uintptr_t)x & 15) == 0) {
    dim3 vg;
    int rpc;
    vec_grid(M, C, ls, &vg, &rpc, (const void*)bn_stats_vec_kernel);
    bn_stats_vec_kernel<<<vg, 256, 0, (cudaStream_t)stream>>>(x, x_dtype == RL_DT_F32, sums, M, (int)C, ls, rpc);
    return rl_check_launch("rl_bn_stats");
  }
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((M + 2047) / 2048));
  bn_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_dtype == RL_DT_F32, sums, M, (int)C, ld);
  return rl_check_launch("rl_bn_stats");
}

extern "C" int rl_bn_finalize(const float* sums, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, int64_t* num_batches_tracked, float* scale, float* shift, float* mean_out,
                              float* rstd_out, int64_t M, int64_t C, float momentum, float eps, void* stream) {
  RL_REQUIRE(sums && gamma && beta && scale && shift && mean_out && rstd_out && M > 0 && C > 0, RL_EINVAL,
             "rl_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      sums, gamma, beta, running_mean, running_var, (long long*)num_batches_tracked, scale, shift, mean_out, rstd_out, M, (int)C,
      momentum, eps);
  return rl_check_launch("rl_bn_finalize");
}

extern "C" int rl_bn_apply(const void* x1, const float* scale1, const float* shift1, const void* x2, const float* scale2,
                           const float* shift2, int32_t x_dtype, void* out, int32_t out_dtype, int32_t relu, int64_t M,
                           int64_t C, int32_t remap, int32_t map_h, int32_t map_w, void* stream) {
  RL_REQUIRE(x1 && scale1 && shift1 && out && M > 0 && C > 0 && C % 8 == 0, RL_EINVAL, "rl_bn_apply: bad arguments");
  RL_REQUIRE(!x2 || (scale2 && shift2), RL_EINVAL, "rl_bn_apply: second operand needs scale/shift");
  int hs = 0, ws = 0;
  if (remap) {
    hs = ilog2x(map_h);
    ws = ilog2x(map_w);
    RL_REQUIRE(hs >= 1 && ws >= 1, RL_EINVAL, "rl_bn_apply: parity split needs a power-of-two map >= 2x2");
  }
  const int ls = vec_lpr_shift(C);
  if (ls >= 0 && (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) & 15) == 0) {
    dim3 vg;
    int rpc;
    vec_grid(M, C, ls, &vg, &rpc, x2 ? (const void*)bn_apply_vec_kernel<2> : (const void*)bn_apply_vec_kernel<1>);
    if (x2)
      bn_apply_vec_kernel<2><<<vg, 256, 0, (cudaStream_t)stream>>>(x1, scale1, shift1, x2, scale2, shift2, x_dtype == RL_DT_F32, out,
                                                                  out_dtype == RL_DT_F32, relu, M, (int)C, ls, rpc, remap, hs + ws, ws);
    else
      bn_apply_vec_kernel<1><<<vg, 256, 0, (cudaStream_t)stream>>>(x1, scale1, shift1, nullptr, nullptr, nullptr, x_dtype == RL_DT_F32,
                                                                  out, out_dtype == RL_DT_F32, relu, M, (int)C, ls, rpc, remap,
                                                                  hs + ws, ws);
    return rl_check_launch("rl_bn_apply");
  }
  const long long n = M * (C / 8);
  bn_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x1, scale1, shift1, x2, scale2, shift2, x_dtype == RL_DT_F32, out, out_dtype == RL_DT_F32, relu, M, (int)C, remap, hs + ws,
      ws);
  return rl_check_launch("rl_bn_apply");
}

extern "C" int rl_bn_bwd2(const void* dy, int32_t dy_dtype, const void* act_out, int32_t act_dtype, int32_t x_dtype, const void* x1,
                          const float* mean1, const float* rstd1, const float* gamma1, float* dbeta1, float* dgamma1, void* dx1,
                          int64_t ldx1, const void* x2, const float* mean2, const float* rstd2, const float* gamma2,
                          float* dbeta2, float* dgamma2, void* dx2, int64_t ldx2, int64_t M, int64_t C, int32_t remap,
                          int32_t map_h, int32_t map_w, const float* fwd_scale1, const float* fwd_shift1,
                          const float* fwd_scale2, const float* fwd_shift2, void* stream) {
  RL_REQUIRE(!fwd_scale1 || (fwd_shift1 && (!x2 || (fwd_scale2 && fwd_shift2))), RL_EINVAL,
             "rl_bn_bwd2: the mask is re-derived from the forward scale AND shift of every branch");
  RL_REQUIRE(dy && x1 && mean1 && rstd1 && gamma1 && dbeta1 && dgamma1 && dx1 && M > 0 && C > 0 && ldx1 >= C && ldx1 % 8 == 0,
             RL_EINVAL, "rl_bn_bwd2: bad arguments");
  RL_REQUIRE(!x2 || (mean2 && rstd2 && gamma2 && dbeta2 && dgamma2 && dx2 && ldx2 >= C && ldx2 % 8 == 0), RL_EINVAL,
             "rl_bn_bwd2: incomplete second branch");
  const int ls = vec_lpr_shift(C);
  RL_REQUIRE(ls >= 0, RL_EINVAL, "rl_bn_bwd2: C must be 8*2^k (< 256) or a multiple of 256");
  RL_REQUIRE((((uintptr_t)dy | (uintptr_t)x1 | (uintptr_t)dx1 | (uintptr_t)x2 | (uintptr_t)dx2 | (uintptr_t)act_out) & 15) == 0,
             RL_EALIGN, "rl_bn_bwd2: pointers must be 16-byte aligned");
  int hs = 0, ws = 0;
  if (remap) {
    hs = ilog2x(map_h);
    ws = ilog2x(map_w);
    RL_REQUIRE(hs >= 1 && ws >= 1, RL_EINVAL, "rl_bn_bwd2: parity split needs a power-of-two map >= 2x2");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int xf = x_dtype == RL_DT_F32;
  BnBranch b0{x1, xf, mean1, rstd1, gamma1, dbeta1, dgamma1, (__nv_bfloat16*)dx1, ldx1, fwd_scale1, fwd_shift1};
  BnBranch b1{x2, xf, mean2, rstd2, gamma2, dbeta2, dgamma2, (__nv_bfloat16*)dx2, ldx2, fwd_scale2, fwd_shift2};
  dim3 vg, va;
  int rpc, rpa;
  vec_grid(M, C, ls, &vg, &rpc, x2 ? (const void*)bn_bwd_reduce_vec_kernel<2> : (const void*)bn_bwd_reduce_vec_kernel<1>);
  vec_grid(M, C, ls, &va, &rpa, x2 ? (const void*)bn_bwd_apply_vec_kernel<2> : (const void*)bn_bwd_apply_vec_kernel<1>);
  const int df = dy_dtype == RL_DT_F32, af = act_dtype == RL_DT_F32;
  if (x2) {
    bn_bwd_reduce_vec_kernel<2><<<vg, 256, 0, st>>>(dy, df, act_out, af, b0, b1, M, (int)C, ls, rpc, remap, hs + ws, ws);
    int rc = rl_check_launch("rl_bn_bwd2(reduce)");
    if (rc) return rc;
    bn_bwd_apply_vec_kernel<2><<<va, 256, 0, st>>>(dy, df, act_out, af, b0, b1, M, (int)C, ls, rpa, remap, hs + ws, ws);
  } else {
    bn_bwd_reduce_vec_kernel<1><<<vg, 256, 0, st>>>(dy, df, act_out, af, b0, b1, M, (int)C, ls, rpc, remap, hs + ws, ws);
    int rc = rl_check_launch("rl_bn_bwd2(reduce)");
    if (rc) return rc;
    bn_bwd_apply_vec_kernel<1><<<va, 256, 0, st>>>(dy, df, act_out, af, b0, b1, M, (int)C, ls, rpa, remap, hs + ws, ws);
  }
  return rl_check_launch("rl_bn_bwd2");
}

extern "C" int rl_bn_bwd(const void* dy, int32_t dy_dtype, const void* act_out, int32_t act_dtype, const void* x,
                         int32_t x_dtype, const float* mean, const float* rstd, const float* gamma, float* dbeta, float* dgamma, void* dx,
                         int64_t ldx, int64_t M, int64_t C, int32_t remap, int32_t map_h, int32_t map_w, const float* fwd_scale,
                         const float* fwd_shift, void* stream) {
  RL_REQUIRE(dy && x && mean && rstd && gamma && dbeta && dgamma && dx && M > 0 && C > 0 && C % 8 == 0 && ldx >= C, RL_EINVAL,
             "rl_bn_bwd: bad arguments");
  int hs = 0, ws = 0;
  if (remap) {
    hs = ilog2x(map_h);
    ws = ilog2x(map_w);
    RL_REQUIRE(hs >= 1 && ws >= 1, RL_EINVAL, "rl_bn_bwd: parity split needs a power-of-two map >= 2x2");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_lpr_shift(C) >= 0 && ldx % 8 == 0 && (((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx) & 15) == 0 &&
      (!act_out || ((uintptr_t)act_out & 15) == 0))
    return rl_bn_bwd2(dy, dy_dtype, act_out, act_dtype, x_dtype, x, mean, rstd, gamma, dbeta, dgamma, dx, ldx, nullptr, nullptr,
                      nullptr, nullptr, nullptr, nullptr, nullptr, 0, M, C, remap, map_h, map_w, fwd_scale,
                      fwd_scale ? fwd_shift : nullptr, nullptr, nullptr, stream);
  RL_REQUIRE(act_out || !fwd_scale, RL_EINVAL, "rl_bn_bwd: the scalar fallback path masks with act_out (pass it too)");
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((M + 2047) / 2048));
  bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dy, dy_dtype == RL_DT_F32, act_out, act_dtype == RL_DT_F32, x, x_dtype == RL_DT_F32,
                                             mean, rstd, dbeta, dgamma, M, (int)C, remap, hs + ws, ws);
  int rc = rl_check_launch("rl_bn_bwd(reduce)");
  if (rc) return rc;
  const long long n = M * (C / 8);
  bn_bwd_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dy, dy_dtype == RL_DT_F32, act_out, act_dtype == RL_DT_F32,
                                                                  x, x_dtype == RL_DT_F32, mean, rstd, gamma, dbeta, dgamma,
                                                                  (__nv_bfloat16*)dx, ldx, M, (int)C, remap, hs + ws, ws);
  return rl_check_launch("rl_bn_bwd");
}

extern "C" int rl_im2col_bf16(const void* x, void* col, int64_t n_img, int32_t C, int32_t W, int32_t H, int32_t P,
                              int32_t ntaps, const int8_t* tap_dw, const int8_t* tap_dh, const int8_t* tap_plane,
                              void* stream) {
  RL_REQUIRE(x && col && tap_dw && tap_dh && tap_plane && ntaps >= 1 && ntaps <= 12 && C % 8 == 0, RL_EINVAL,
             "rl_im2col_bf16: bad arguments");
  const int ws = ilog2x(W), hs = ilog2x(H);
  RL_REQUIRE(ws >= 0 && hs >= 0, RL_EINVAL, "rl_im2col_bf16: map must be power-of-two");
  TapTable t;
  t.n = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    t.dw[i] = tap_dw[i];
    t.dh[i] = tap_dh[i];
    t.plane[i] = tap_plane[i];
  }
  const long long M = n_img * W * H;
  const long long n = M * (C / 8);
  if (n == 0) return 0;
  im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, M, C,
                                                                              W, H, P, ws + hs, ws, t);
  return rl_check_launch("rl_im2col_bf16");
}

extern "C" int rl_glyph_im2col(const float* glyphs, const int64_t* ids, void* col1, void* colsc, int64_t n_img, int32_t C,
                               void* stream) {
  RL_REQUIRE(glyphs && ids && col1 && C >= 1 && C <= 3, RL_EINVAL, "rl_glyph_im2col: bad arguments (num_fonts 1..3)");  // colsc optional
  if (n_img <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3)
    glyph_im2col_kernel<3><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, (__nv_bfloat16*)col1, (__nv_bfloat16*)colsc);
  else if (C == 2)
    glyph_im2col_kernel<2><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, (__nv_bfloat16*)col1, (__nv_bfloat16*)colsc);
  else
    glyph_im2col_kernel<1><<<(unsigned)n_img, 256, 0, st>>>(glyphs, (const long long*)ids, (__nv_bfloat16*)col1, (__nv_bfloat16*)colsc);
  return rl_check_launch("rl_glyph_im2col");
}
