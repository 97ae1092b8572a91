// Fused CharResNet block 1 (eval mode): glyph gather -> conv3x3/s2 (C->64)+BN+ReLU -> conv3x3 (64->64)+BN,
// + conv1x1/s2 shortcut+BN, ReLU.  One persistent CTA per SM streams glyphs: 12 KB in (fp32 bitmap),
// 32 KB out (bf16 NHWC, parity-split for the stride-2 conv of block 2); nothing else touches HBM.
//
// Everything runs on tcgen05.  Per glyph:
//   A1   = im2col of the bitmap, [256 px x 32 (27 taps, zero padded)] bf16, built by the 256 pixel threads
//   D1   = A1 . W1'^T            (conv1, BN1 scale folded into W1')            TMEM cols   0..127 (2 halves)
//   D2   = A1 . Wsc'^T           (shortcut as the centre-tap columns of A1)     TMEM cols 128..255
//   Y    = relu(D1 + t1) as bf16, written three times into shared memory, shifted by dw = -1, 0, +1 and
//          padded by one zero row above/below, so that every 3x3 tap of conv2 is a contiguous
//          SWIZZLE_128B K-major operand tile (im2col-free: the taps are just descriptor offsets)
//   D2  += sum over 9 taps  Y[dw](rows shifted by dh) . W2'[tap]^T               (conv2, BN2 scale folded)
//   out  = relu(D2 + t2 + t_sc) as bf16
// BatchNorm is the eval-mode affine map; scales are folded into the bf16 weights on the host, shifts are added
// in the epilogues.  Reference: src/models.py:829-834 (gather), src/char_cnn.py:15-32 (BasicBlock).
#include "common.cuh"

namespace {

constexpr int B1_THREADS = 288;  // 8 pixel warps + 1 control warp
constexpr int Y_COPY_BYTES = 18 * 16 * 128;  // 18 rows (1 zero row top/bottom) x 16 px x 64 ch bf16
constexpr int W2_BYTES = 9 * 64 * 128;
constexpr int A1_BYTES = 256 * 64;   // 64-byte rows (K = 32), SWIZZLE_64B
constexpr int B1_BYTES = 64 * 64;

struct B1Params {
  const float* glyphs;
  const long long* ids;
  const float* t1;   // [64]  BN1 shift
  const float* t2s;  // [64]  BN2 shift + shortcut-BN shift
  __nv_bfloat16* out;  // [n_img*256, 64], parity-split rows
  int n_img;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                           // LBO (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

template <int C>
__global__ void __launch_bounds__(B1_THREADS, 1)
glyph_block1_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmWsc,
                    const __grid_constant__ CUtensorMap tmW2, const B1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (rl::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sY = smem;                          // 3 copies, 1024-aligned (36864 = 36 * 1024)
  uint8_t* sW2 = sY + 3 * Y_COPY_BYTES;        // 9 tap tiles [64 x 64] bf16 SW128
  uint8_t* sA1 = sW2 + W2_BYTES;               // [256 x 32] bf16 SW64
  uint8_t* sB1 = sA1 + A1_BYTES;               // [64 x 32] bf16 SW64
  uint8_t* sBsc = sB1 + B1_BYTES;
  float* s_img = reinterpret_cast<float*>(sBsc + B1_BYTES);  // [C][32][32] fp32
  float* s_t1 = s_img + C * 1024;
  float* s_t2 = s_t1 + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_t2 + 64);
  uint64_t* bar_w = &bars[0];     // weights landed (TMA)
  uint64_t* bar_a1 = &bars[1];    // im2col tile written (256 arrivals)
  uint64_t* bar_m1 = &bars[2];    // conv1 + shortcut MMAs done
  uint64_t* bar_y = &bars[3];     // Y copies written (256 arrivals)
  uint64_t* bar_m2 = &bars[4];    // conv2 MMAs done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 256) {
    rl::tma_prefetch_desc(&tmW1);
    rl::tma_prefetch_desc(&tmWsc);
    rl::tma_prefetch_desc(&tmW2);
    rl::mbar_init(bar_w, 1);
    rl::mbar_init(bar_a1, 256);
    rl::mbar_init(bar_m1, 1);
    rl::mbar_init(bar_y, 256);
    rl::mbar_init(bar_m2, 1);
    rl::fence_barrier_init();
  }
  if (warp == 8) rl::tmem_alloc(tmem_ptr, 256);
  // zero the Y copies once: the halo rows and the shifted-out columns are never written again
  for (int i = tid; i < 3 * Y_COPY_BYTES / 16; i += B1_THREADS) reinterpret_cast<uint4*>(sY)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 64) {
    s_t1[tid] = p.t1[tid];
    s_t2[tid] = p.t2s[tid];
  }
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    if (lane == 0) {
      // ---------------- control thread: weights once, then the MMAs of every glyph ----------------
      rl::mbar_expect_tx(bar_w, W2_BYTES + 2 * B1_BYTES);
      for (int t = 0; t < 9; ++t) rl::tma_load_2d(sW2 + t * 8192, &tmW2, bar_w, t * 64, 0);
      rl::tma_load_2d(sB1, &tmW1, bar_w, 0, 0);
      rl::tma_load_2d(sBsc, &tmWsc, bar_w, 0, 0);
      rl::mbar_wait(bar_w, 0);
      constexpr uint32_t idesc = rl::make_idesc_bf16(128, 64);
      const uint32_t a1 = rl::smem_u32(sA1), b1 = rl::smem_u32(sB1), bsc = rl::smem_u32(sBsc);
      const uint32_t yb = rl::smem_u32(sY), w2 = rl::smem_u32(sW2);
      uint32_t ph = 0;
      for (int img = blockIdx.x; img < p.n_img; img += gridDim.x, ph ^= 1) {
        rl::mbar_wait(bar_a1, ph);
        rl::tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            rl::tc_mma_f16(tmem_base + h * 64, make_desc(a1 + h * 8192 + k * 32, 512, 4), make_desc(b1 + k * 32, 512, 4),
                           idesc, k);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            rl::tc_mma_f16(tmem_base + 128 + h * 64, make_desc(a1 + h * 8192 + k * 32, 512, 4),
                           make_desc(bsc + k * 32, 512, 4), idesc, k);
        }
        rl::tc_commit(bar_m1);
        rl::mbar_wait(bar_y, ph);
        rl::tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          for (int t = 0; t < 9; ++t) {
            const int dh = t / 3 - 1, dwi = t % 3;  // tap (kh, kw): dh = kh - 1, copy index = kw
            const uint32_t ya = yb + dwi * Y_COPY_BYTES + (h * 8 + 1 + dh) * 2048;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              rl::tc_mma_f16(tmem_base + 128 + h * 64, make_desc(ya + k * 32, 1024, 2),
                             make_desc(w2 + t * 8192 + k * 32, 1024, 2), idesc, 1);
          }
        }
        rl::tc_commit(bar_m2);
      }
    }
  } else {
    // ---------------- pixel threads: thread t <-> output pixel t of the 16x16 map ----------------
    const int px = tid;              // 0..255
    const int oh = px >> 4, ow = px & 15;
    const int half = px >> 7;        // accumulator half (TMEM columns), lanes = px & 127
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph = 0;
    // software prefetch: the next glyph's bitmap is pulled into registers while this one is processed
    constexpr int NV = (C * 256 + 255) / 256;  // float4 per thread
    float4 pre[NV];
    {
      const float4* src = reinterpret_cast<const float4*>(p.glyphs + p.ids[blockIdx.x] * (long long)(C * 1024));
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (tid + i * 256 < C * 256) pre[i] = __ldg(src + tid + i * 256);
    }
    for (int img = blockIdx.x; img < p.n_img; img += gridDim.x, ph ^= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (tid + i * 256 < C * 256) reinterpret_cast<float4*>(s_img)[tid + i * 256] = pre[i];
      if (img + (int)gridDim.x < p.n_img) {
        const float4* src =
            reinterpret_cast<const float4*>(p.glyphs + p.ids[img + gridDim.x] * (long long)(C * 1024));
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (tid + i * 256 < C * 256) pre[i] = __ldg(src + tid + i * 256);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // im2col row of this pixel: k = c*9 + kh*3 + kw -> bitmap[c][2*oh+kh-1][2*ow+kw-1]
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
      {
        uint8_t* rowp = sA1 + px * 64;
        const int sw = (px >> 1) & 3;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(rowp + ((g ^ sw) << 4)) =
              make_uint4(rl::pack_bf16(a[8 * g], a[8 * g + 1]), rl::pack_bf16(a[8 * g + 2], a[8 * g + 3]),
                         rl::pack_bf16(a[8 * g + 4], a[8 * g + 5]), rl::pack_bf16(a[8 * g + 6], a[8 * g + 7]));
      }
      rl::fence_proxy_async();
      rl::tc_fence_before();      // orders this thread's earlier tcgen05.ld of D1/D2 before the next MMAs
      rl::mbar_arrive(bar_a1);

      // ---- epilogue 1: Y = relu(conv1 + t1) -> three shifted bf16 copies ----
      rl::mbar_wait(bar_m1, ph);
      rl::tc_fence_after();
      {
        const int R = (oh + 1) * 16 + ow;  // row of this pixel in a padded copy
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          rl::tmem_ld_32x32(t_lane + half * 64 + cc * 32, v);
          rl::tmem_ld_wait();
          uint4 pk[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(__uint_as_float(v[8 * g + j]) + s_t1[cc * 32 + 8 * g + j], 0.f);
            pk[g] = make_uint4(rl::pack_bf16(y[0], y[1]), rl::pack_bf16(y[2], y[3]), rl::pack_bf16(y[4], y[5]),
                               rl::pack_bf16(y[6], y[7]));
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            // copy d holds y shifted by dw = d - 1:  copy[h][w'] = y[h][w' + dw]  ->  y[h][w] lands at w' = w - dw
            const int wq = ow - (d - 1);
            if (wq >= 0 && wq <= 15) {
              const int r = R - (d - 1);
              uint8_t* rowp = sY + d * Y_COPY_BYTES + r * 128;
#pragma unroll
              for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(rowp + (((cc * 4 + g) ^ (r & 7)) << 4)) = pk[g];
            }
          }
        }
      }
      rl::fence_proxy_async();
      rl::tc_fence_before();
      rl::mbar_arrive(bar_y);

      // ---- epilogue 2: out = relu(conv2 + shortcut + t2 + t_sc), parity-split NHWC row ----
      rl::mbar_wait(bar_m2, ph);
      rl::tc_fence_after();
      {
        const long long orow = (((long long)img * 4 + (oh & 1) * 2 + (ow & 1)) * 8 + (oh >> 1)) * 8 + (ow >> 1);
        uint4* dst = reinterpret_cast<uint4*>(p.out + orow * 64);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          rl::tmem_ld_32x32(t_lane + 128 + half * 64 + cc * 32, v);
          rl::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(__uint_as_float(v[8 * g + j]) + s_t2[cc * 32 + 8 * g + j], 0.f);
            dst[cc * 4 + g] = make_uint4(rl::pack_bf16(y[0], y[1]), rl::pack_bf16(y[2], y[3]),
                                         rl::pack_bf16(y[4], y[5]), rl::pack_bf16(y[6], y[7]));
          }
        }
      }
    }
  }
  rl::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, 256);
  }
}

template <int C>
constexpr int b1_smem_bytes() {
  return 3 * Y_COPY_BYTES + W2_BYTES + A1_BYTES + 2 * B1_BYTES + C * 4096 + 2 * 64 * 4 + 5 * 8 + 16 + 1024;
}

template <int C>
int launch_b1(const CUtensorMap& t1, const CUtensorMap& tsc, const CUtensorMap& t2, const B1Params& p, cudaStream_t st) {
  constexpr int smem = b1_smem_bytes<C>();
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(glyph_block1_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      rl_set_error("rl_glyph_block1_fwd: cudaFuncSetAttribute(%d B) failed: %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.n_img < rl_num_sms() ? p.n_img : rl_num_sms();
  glyph_block1_kernel<C><<<grid, B1_THREADS, smem, st>>>(t1, tsc, t2, p);
  return rl_check_launch("rl_glyph_block1_fwd");
}

}  // namespace

extern "C" int rl_glyph_block1_fwd(const float* glyphs, const int64_t* ids, const void* w1_packed,
                                   const void* wsc_packed, const void* w2_packed, const float* t1,
                                   const float* t2s, void* out, int64_t n_img, int32_t C, void* stream) {
  RL_REQUIRE(glyphs && ids && w1_packed && wsc_packed && w2_packed && t1 && t2s && out, RL_EINVAL,
             "rl_glyph_block1_fwd: null pointer");
  RL_REQUIRE(C >= 1 && C <= 3, RL_EINVAL, "rl_glyph_block1_fwd: num_fonts must be 1, 2 or 3 (got %d: 9*C taps must fit K = 32)", C);
  RL_REQUIRE(((uintptr_t)glyphs & 15) == 0 && ((uintptr_t)out & 15) == 0, RL_EALIGN, "rl_glyph_block1_fwd: alignment");
  if (n_img <= 0) return 0;
  CUtensorMap m1, msc, m2;
  {
    uint64_t dims[2] = {32, 64};
    uint64_t strides[1] = {64};
    uint32_t box[2] = {32, 64};
    int rc = rl_make_tmap(&m1, w1_packed, RL_TMAP_BF16, 64, 2, dims, strides, box);
    if (rc) return rc;
    rc = rl_make_tmap(&msc, wsc_packed, RL_TMAP_BF16, 64, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {576, 64};
    uint64_t strides[1] = {576 * 2};
    uint32_t box[2] = {64, 64};
    int rc = rl_make_tmap(&m2, w2_packed, RL_TMAP_BF16, 128, 2, dims, strides, box);
    if (rc) return rc;
  }
  B1Params p;
  p.glyphs = glyphs;
  p.ids = reinterpret_cast<const long long*>(ids);
  p.t1 = t1;
  p.t2s = t2s;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.n_img = (int)n_img;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return C == 3 ? launch_b1<3>(m1, msc, m2, p, st) : C == 2 ? launch_b1<2>(m1, msc, m2, p, st) : launch_b1<1>(m1, msc, m2, p, st);
}
