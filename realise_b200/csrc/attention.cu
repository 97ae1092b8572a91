// Fused BertSelfAttention core on tcgen05:  ctx = softmax(Q K^T / sqrt(d) + (1-mask)*-10000) V
// One CTA per (q-tile of 128 rows, head, sentence).  S = Q K^T accumulates in TMEM (128 lanes x Lkv
// fp32 columns), each of the 128 threads owns one query row (tcgen05.ld 32x32b) and does the masked
// softmax in registers, P is written as bf16 into SWIZZLE_128B K-major shared tiles, and O = P V
// is a second tcgen05 MMA whose B operand is V as TMA delivered it ([kv, d] = MN-major).
// Q/K/V are read straight out of the fused QKV projection [tokens, 3*H] with 2-D TMA boxes.
#include "common.cuh"

namespace {

constexpr int ATT_THREADS = 128;
constexpr int HEAD_DIM = 64;

struct AttParams {
  const long long* mask;  // [B, L] 1 = token, 0 = padding (int64 as in the reference batch)
  __nv_bfloat16* ctx;     // [B*L, H]
  int L, H, lkv16;
  float scale_log2;       // (1/sqrt(d)) * log2(e)
  int heads;
  rl::DropSpec drop;      // dropout on the attention probabilities (modeling_bert.py:250)
  int f16;                // Q/K/V, P and ctx are fp16 instead of bf16 (act_dtype == RL_DT_F16)
  float* lse;             // optional [B, heads, L]: log2-domain logsumexp of every query row (saved for the backward)
};

template <int LKV_MAX>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const AttParams p) {
  // O (64 fp32 columns) reuses the first S columns once every thread has consumed its S row, and for
  // Lkv <= 128 the bf16 P tiles reuse the Q and K staging, so a CTA needs 128 TMEM columns and
  // ~49 KB of shared memory: four CTAs are resident per SM and hide each other's TMA/MMA latency.
  constexpr uint32_t TMEM_COLS = LKV_MAX <= 128 ? 128 : 256;
  constexpr uint32_t O_COL = 0;
  constexpr int Q_BYTES = 128 * 128;
  constexpr int KV_BYTES = LKV_MAX * 128;
  constexpr int P_CHUNKS = LKV_MAX / 64;
  constexpr bool P_ALIAS = LKV_MAX <= 128;

  extern __shared__ uint8_t smem_raw[];
  // pad to 1024 B by pointer arithmetic (keeps the shared address space visible to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (rl::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sP = P_ALIAS ? sQ : sV + KV_BYTES;  // P_CHUNKS tiles of [128 x 64] bf16, SW128 K-major
  float* s_mask = reinterpret_cast<float*>(sV + KV_BYTES + (P_ALIAS ? 0 : P_CHUNKS * Q_BYTES));
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + LKV_MAX);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  uint64_t* bar_qk = &bars[0];
  uint64_t* bar_v = &bars[1];
  uint64_t* bar_s = &bars[2];
  uint64_t* bar_o = &bars[3];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int L = p.L, lkv16 = p.lkv16;
  const int row0 = b * L;

  if (tid == 0) {
    rl::tma_prefetch_desc(&tmQ);
    rl::tma_prefetch_desc(&tmKV);
    rl::mbar_init(bar_qk, 1);
    rl::mbar_init(bar_v, 1);
    rl::mbar_init(bar_s, 1);
    rl::mbar_init(bar_o, 1);
    rl::fence_barrier_init();
  }
  if (warp == 0) rl::tmem_alloc(tmem_ptr, TMEM_COLS);
  // additive mask, exactly the reference's (1 - mask) * -10000 (modeling_bert.py:696-697), in log2 units
  for (int j = tid; j < LKV_MAX; j += ATT_THREADS) {
    float m = -INFINITY;
    if (j < L) m = p.mask[(long long)b * L + j] != 0 ? 0.0f : -10000.0f * 1.4426950408889634f;   // attention masks are 0 / 1 (no I2F.S64)
    s_mask[j] = m;
  }
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    rl::mbar_expect_tx(bar_qk, Q_BYTES + lkv16 * 128);
    rl::tma_load_2d(sQ, &tmQ, bar_qk, head * HEAD_DIM, row0 + q0);
    rl::tma_load_2d(sK, &tmKV, bar_qk, p.H + head * HEAD_DIM, row0);
    rl::mbar_expect_tx(bar_v, lkv16 * 128);
    rl::tma_load_2d(sV, &tmKV, bar_v, 2 * p.H + head * HEAD_DIM, row0);
    // S = Q K^T : M=128, N=lkv16, K=64
    rl::mbar_wait(bar_qk, 0);
    rl::tc_fence_after();
    const uint32_t idesc_s = rl::make_idesc_bf16(128, lkv16, 0, 0, p.f16);
    const uint32_t qa = rl::smem_u32(sQ), ka = rl::smem_u32(sK);
#pragma unroll
    for (int k = 0; k < HEAD_DIM / 16; ++k) {
      rl::tc_mma_f16(tmem_base, rl::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                     rl::make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k != 0);
    }
    rl::tc_commit(bar_s);
  }

  // ---- softmax: thread r owns query row r of the tile ----
  rl::mbar_wait(bar_s, 0);
  rl::tc_fence_after();
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int nchunk = (lkv16 + 31) / 32;
  const int nfull = L / 32;  // chunks whose 32 columns are all real keys
  const float4* m4 = reinterpret_cast<const float4*>(s_mask);
  float mx = -INFINITY;
  for (int c = 0; c < nchunk; ++c) {
    uint32_t v[32];
    rl::tmem_ld_32x32(t_row + c * 32, v);
    rl::tmem_ld_wait();
    if (c < nfull) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 m = m4[c * 8 + (j >> 2)];
        mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), p.scale_log2, m.x));
        mx = fmaxf(mx, fmaf(__uint_as_float(v[j + 1]), p.scale_log2, m.y));
        mx = fmaxf(mx, fmaf(__uint_as_float(v[j + 2]), p.scale_log2, m.z));
        mx = fmaxf(mx, fmaf(__uint_as_float(v[j + 3]), p.scale_log2, m.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;  // columns >= L (tile padding / stale TMEM) never contribute
        const float sc = col < L ? fmaf(__uint_as_float(v[j]), p.scale_log2, s_mask[col]) : -INFINITY;
        mx = fmaxf(mx, sc);
      }
    }
  }
  float sum = 0.0f;
  const int r = tid;  // row within the 128-row tile
  for (int c = 0; c < nchunk; ++c) {
    uint32_t v[32];
    rl::tmem_ld_32x32(t_row + c * 32, v);
    rl::tmem_ld_wait();
    float e[32];
    if (c < nfull) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 m = m4[c * 8 + (j >> 2)];
        e[j] = rl::ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, m.x) - mx);
        e[j + 1] = rl::ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, m.y) - mx);
        e[j + 2] = rl::ex2(fmaf(__uint_as_float(v[j + 2]), p.scale_log2, m.z) - mx);
        e[j + 3] = rl::ex2(fmaf(__uint_as_float(v[j + 3]), p.scale_log2, m.w) - mx);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        const float sc = col < L ? fmaf(__uint_as_float(v[j]), p.scale_log2, s_mask[col]) : -INFINITY;
        e[j] = rl::ex2(sc - mx);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) sum += (e[j] + e[j + 1]) + (e[j + 2] + e[j + 3]);
    if (p.drop.thresh) {  // the row sum stays that of the undropped probabilities (softmax, then dropout)
      const unsigned long long e0 = (((unsigned long long)b * p.heads + head) * L + (q0 + r)) * L + c * 32;
      rl::DropSpec dsp = p.drop;
      rl::drop_resolve(dsp);
      rl::drop_apply32(dsp, e0, e);
    }
    // columns c*32 .. c*32+31 of P -> chunk tile (c/2), 16-byte pieces (c&1)*4 .. +3, swizzled by row
    uint8_t* tile = sP + (c >> 1) * Q_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int piece = ((c & 1) * 4 + g) ^ (r & 7);
      *reinterpret_cast<uint4*>(tile + piece * 16) =
          make_uint4(rl::pack_h(e[8 * g], e[8 * g + 1], p.f16), rl::pack_h(e[8 * g + 2], e[8 * g + 3], p.f16),
                     rl::pack_h(e[8 * g + 4], e[8 * g + 5], p.f16), rl::pack_h(e[8 * g + 6], e[8 * g + 7], p.f16));
    }
  }
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    // O = P V : M=128, N=64, K=lkv16;  A = P (K-major tiles of 64 columns), B = V (MN-major)
    rl::tc_fence_after();
    rl::mbar_wait(bar_v, 0);
    rl::tc_fence_after();
    const uint32_t idesc_o = rl::make_idesc_bf16(128, HEAD_DIM, 0, 1, p.f16);
    const uint32_t pa = rl::smem_u32(sP), va = rl::smem_u32(sV);
    const int nk = lkv16 / 16;
    for (int k = 0; k < nk; ++k) {
      const uint32_t a_addr = pa + (k >> 2) * Q_BYTES + (k & 3) * 32;
      const uint32_t b_addr = va + k * 2048;  // 16 kv rows of 128 B
      rl::tc_mma_f16(tmem_base + O_COL, rl::make_smem_desc_sw128(a_addr, 16, 1024),
                     rl::make_smem_desc_sw128(b_addr, 1024, 1024), idesc_o, k != 0);
    }
    rl::tc_commit(bar_o);
  }
  rl::mbar_wait(bar_o, 0);
  rl::tc_fence_after();
  {
    const float inv = 1.0f / sum;
    const int q = q0 + r;
    if (p.lse && q < L) p.lse[((long long)b * p.heads + head) * L + q] = mx + log2f(sum);
    __nv_bfloat16* dst = p.ctx + (long long)(row0 + q) * p.H + head * HEAD_DIM;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      rl::tmem_ld_32x32(t_row + O_COL + c * 32, v);
      rl::tmem_ld_wait();
      if (q < L) {
        uint4* o = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          o[g] = make_uint4(rl::pack_h(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv, p.f16));
      }
    }
  }
  rl::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int LKV_MAX>
constexpr int att_smem_bytes() {
  return 128 * 128 + 2 * LKV_MAX * 128 + (LKV_MAX <= 128 ? 0 : (LKV_MAX / 64) * 128 * 128) + LKV_MAX * 4 + 4 * 8 + 16 +
         1024;
}

template <int LKV_MAX>
int launch_att(const CUtensorMap& tq, const CUtensorMap& tkv, const AttParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = att_smem_bytes<LKV_MAX>();
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<LKV_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      rl_set_error("rl_attention_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  attention_kernel<LKV_MAX><<<grid, ATT_THREADS, smem, st>>>(tq, tkv, p);
  return rl_check_launch("rl_attention_fwd");
}

}  // namespace

extern "C" int rl_attention_fwd(const void* qkv, const int64_t* mask, void* ctx, float* row_lse, int64_t B, int64_t L,
                                int64_t heads, int64_t head_dim, int32_t act_dtype, float drop_p, uint64_t drop_seed,
                                uint32_t drop_site, const uint64_t* drop_counter, void* stream) {
  RL_REQUIRE(qkv && mask && ctx, RL_EINVAL, "rl_attention_fwd: null pointer");
  RL_REQUIRE(head_dim == HEAD_DIM, RL_EINVAL, "rl_attention_fwd: head_dim must be 64, got %lld", (long long)head_dim);
  RL_REQUIRE(B > 0 && heads > 0 && L > 0, RL_EINVAL, "rl_attention_fwd: empty problem");
  RL_REQUIRE(L <= 256, RL_EINVAL, "rl_attention_fwd: seq_len %lld > 256 not supported", (long long)L);
  RL_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)ctx & 15) == 0, RL_EALIGN, "rl_attention_fwd: alignment");
  const int H = (int)(heads * head_dim);
  const int lkv16 = (int)((L + 15) / 16 * 16);
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {(uint64_t)(3 * H), (uint64_t)(B * L)};
  uint64_t strides[1] = {(uint64_t)(3 * H) * 2};
  uint32_t boxq[2] = {64, 128};
  uint32_t boxkv[2] = {64, (uint32_t)lkv16};
  int rc = rl_make_tmap_bf16(&tq, qkv, 2, dims, strides, boxq);
  if (rc) return rc;
  rc = rl_make_tmap_bf16(&tkv, qkv, 2, dims, strides, boxkv);
  if (rc) return rc;
  AttParams p;
  p.mask = reinterpret_cast<const long long*>(mask);
  p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.L = (int)L;
  p.H = H;
  p.lkv16 = lkv16;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.heads = (int)heads;
  p.drop = rl::make_drop(drop_p, drop_seed, drop_site, drop_counter);
  p.lse = row_lse;
  p.f16 = act_dtype == RL_DT_F16;
  dim3 grid((unsigned)((L + 127) / 128), (unsigned)heads, (unsigned)B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (lkv16 <= 128) return launch_att<128>(tq, tkv, p, grid, st);
  return launch_att<256>(tq, tkv, p, grid, st);
}
