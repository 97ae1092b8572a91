// BertSelfAttention backward on tcgen05 (training path; seq_len <= 128, one CTA per (sentence, head)).
//   given dO = d(ctx):   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),  delta_q = sum_d dO[q,d] O[q,d],
//                        dQ = dS K / 8,  dK = dS^T Q / 8            (P = softmax(QK^T/8 + mask) is recomputed)
// Seven small GEMMs per head, all on the tensor core: S and dP accumulate in TMEM (thread = query row does the
// softmax algebra), P and dS are written as bf16 K-major smem tiles; the transposed products (P^T dO, dS^T Q)
// read the same tiles through MN-major descriptors, so nothing is transposed in memory.
// Reference math: transformers/modeling_bert.py:234-260 differentiated (dropout = identity, p = 0).
#include "common.cuh"

namespace {

constexpr int ATT_THREADS = 128;
constexpr int HEAD_DIM = 64;
constexpr int T16K = 128 * 128;  // one [128 x 64] bf16 tile

struct AttBwdParams {
  const long long* mask;
  const __nv_bfloat16* ctx;   // forward output O, [B*L, H]
  const __nv_bfloat16* dctx;  // dO
  __nv_bfloat16* dqkv;        // [B*L, 3H]
  int L, H, lkv16;
  float scale_log2;
  int heads;
  rl::DropSpec drop;
  const float* lse;           // optional [B, heads, L] log2-domain logsumexp saved by the forward: skips two passes over S
  float* dbias;               // optional [3H] f32, accumulated: column sums of dqkv (the fused q/k/v bias gradient)
  int f16;                    // every 16-bit tensor of the call (Q/K/V, ctx, dO, dqkv, the P / dS tiles) is fp16 instead of
                              // bf16: tcgen05 kind::f16 needs ONE format for both operands of an MMA (probed on B200)
};

__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmDQKV, const AttBwdParams p) {
  // Two CTAs per SM: 256 TMEM columns and 7 x 16 KB of shared memory each.  S / dP (phase 1) are dead once every
  // thread has turned them into the bf16 P / dS tiles, so the phase-2 accumulators dV / dK / dQ reuse their columns;
  // V is dead once dP = dO V^T has completed (bar_s), so kv-tile 0 of P overwrites it.
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 0, COL_DK = 64, COL_DQ = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((rl::smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B tiles need 1 KB alignment (no slack is budgeted)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T16K;
  uint8_t* sV = sK + T16K;
  uint8_t* sDO = sV + T16K;
  uint8_t* sP1 = sDO + T16K;       // P, kv tile 1 ([128 q x 64 kv]); kv tile 0 aliases sV
  uint8_t* sDS = sP1 + T16K;       // dS, two tiles
  float* s_mask = reinterpret_cast<float*>(sDS + 2 * T16K);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + 128);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);
  uint64_t* bar_ld = &bars[0];
  uint64_t* bar_s = &bars[1];
  uint64_t* bar_o = &bars[2];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.x, b = blockIdx.y;
  const int L = p.L, lkv16 = p.lkv16;
  const int row0 = b * L;

  if (tid == 0) {
    rl::tma_prefetch_desc(&tmQ);
    rl::tma_prefetch_desc(&tmKV);
    rl::tma_prefetch_desc(&tmDO);
    rl::tma_prefetch_desc(&tmO);
    rl::tma_prefetch_desc(&tmDQKV);
    rl::mbar_init(bar_ld, 1);
    rl::mbar_init(bar_s, 1);
    rl::mbar_init(bar_o, 1);
    rl::fence_barrier_init();
  }
  if (warp == 0) rl::tmem_alloc(tmem_ptr, 256);
  for (int j = tid; j < 128; j += ATT_THREADS) {
    float m = -INFINITY;
    if (j < L) m = p.mask[(long long)b * L + j] != 0 ? 0.0f : -10000.0f * 1.4426950408889634f;   // attention masks are 0 / 1
    s_mask[j] = m;
  }
  // zero P tile 1 and dS tile 1 once: chunk tiles or rows that are never written must not feed NaNs into the MMAs
  // (dS tile 0 first receives the O tile; every thread clears its own row there once it has read it)
  for (int i = tid; i < T16K / 16; i += ATT_THREADS) {
    reinterpret_cast<uint4*>(sP1)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sDS + T16K)[i] = make_uint4(0, 0, 0, 0);
  }
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    rl::mbar_expect_tx(bar_ld, 3 * T16K + 2 * lkv16 * 128);
    rl::tma_load_2d(sDS, &tmO, bar_ld, head * HEAD_DIM, row0);   // O, for delta: lives in dS tile 0 until the rows are read
    rl::tma_load_2d(sQ, &tmQ, bar_ld, head * HEAD_DIM, row0);
    rl::tma_load_2d(sK, &tmKV, bar_ld, p.H + head * HEAD_DIM, row0);
    rl::tma_load_2d(sV, &tmKV, bar_ld, 2 * p.H + head * HEAD_DIM, row0);
    rl::tma_load_2d(sDO, &tmDO, bar_ld, head * HEAD_DIM, row0);
    rl::mbar_wait(bar_ld, 0);
    rl::tc_fence_after();
    const uint32_t idesc = rl::make_idesc_h(128, lkv16, 0, 0, p.f16, p.f16);      // S = Q K^T
    const uint32_t idesc_dp = rl::make_idesc_h(128, lkv16, 0, 0, p.f16, p.f16);   // dP = dO V^T
    const uint32_t qa = rl::smem_u32(sQ), ka = rl::smem_u32(sK), va = rl::smem_u32(sV), da = rl::smem_u32(sDO);
#pragma unroll
    for (int k = 0; k < 4; ++k)  // S = Q K^T
      rl::tc_mma_f16(tmem_base + COL_S, rl::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                     rl::make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)  // dP = dO V^T
      rl::tc_mma_f16(tmem_base + COL_DP, rl::make_smem_desc_sw128(da + k * 32, 16, 1024),
                     rl::make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_dp, k != 0);
    rl::tc_commit(bar_s);
  }

  // delta = rowsum(dO o O) for this thread's query row, from the two TMA-loaded tiles (row-strided global loads cost 32
  // cache lines per instruction and were 19 % of the kernel's stall samples).  Row r of a SWIZZLE_128B tile: 16-byte piece g
  // sits at ((g ^ (r & 7)) << 4) of the 128-byte row.
  const int r = tid;
  float delta = 0.f;
  if (tid != 0) rl::mbar_wait(bar_ld, 0);   // (thread 0 has waited above)
  {
    const uint8_t* orow = sDS + (r >> 3) * 1024 + (r & 7) * 128;
    const uint8_t* grow = sDO + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 a = *reinterpret_cast<const uint4*>(orow + ((i ^ (r & 7)) << 4));
      const uint4 c = *reinterpret_cast<const uint4*>(grow + ((i ^ (r & 7)) << 4));
      delta += rl::half_lo(a.x, p.f16) * rl::half_lo(c.x, p.f16) + rl::half_hi(a.x, p.f16) * rl::half_hi(c.x, p.f16) +
               rl::half_lo(a.y, p.f16) * rl::half_lo(c.y, p.f16) + rl::half_hi(a.y, p.f16) * rl::half_hi(c.y, p.f16) +
               rl::half_lo(a.z, p.f16) * rl::half_lo(c.z, p.f16) + rl::half_hi(a.z, p.f16) * rl::half_hi(c.z, p.f16) +
               rl::half_lo(a.w, p.f16) * rl::half_lo(c.w, p.f16) + rl::half_hi(a.w, p.f16) * rl::half_hi(c.w, p.f16);
    }
    // my row of dS tile 0 held O: clear it (kv chunks this sentence does not reach stay zero)
    uint8_t* zrow = sDS + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(zrow + (i << 4)) = make_uint4(0, 0, 0, 0);
  }
  if (r >= L) delta = 0.f;

  rl::mbar_wait(bar_s, 0);
  rl::tc_fence_after();
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int nchunk = (lkv16 + 31) / 32;
  float mx = -INFINITY;
  float inv = 0.f;
  if (p.lse) {
    // P = exp2(s - lse) directly; rows beyond the sentence get lse = +inf -> P = 0 (nothing for dK / dV)
    mx = r < L ? p.lse[((long long)b * p.heads + head) * L + r] : INFINITY;
    inv = 1.0f;
  } else {
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[32];
      rl::tmem_ld_32x32(t_row + COL_S + c * 32, v);
      rl::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        const float sc = col < L ? fmaf(__uint_as_float(v[j]), p.scale_log2, s_mask[col]) : -INFINITY;
        mx = fmaxf(mx, sc);
      }
    }
    float sum = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[32];
      rl::tmem_ld_32x32(t_row + COL_S + c * 32, v);
      rl::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        const float sc = col < L ? fmaf(__uint_as_float(v[j]), p.scale_log2, s_mask[col]) : -INFINITY;
        sum += rl::ex2(sc - mx);
      }
    }
    inv = r < L ? 1.0f / sum : 0.0f;  // query rows beyond the sentence contribute nothing to dK / dV
  }
  rl::DropSpec dsp = p.drop;
  rl::drop_resolve(dsp);
  for (int c = 0; c < nchunk; ++c) {
    uint32_t v[32], w[32];
    rl::tmem_ld_32x32(t_row + COL_S + c * 32, v);
    rl::tmem_ld_32x32(t_row + COL_DP + c * 32, w);
    rl::tmem_ld_wait();
    float pr[32], ds[32];
    unsigned int keep_bits = 0xFFFFFFFFu;
    if (dsp.thresh) keep_bits = rl::drop_bits32(dsp, (((unsigned long long)b * p.heads + head) * L + r) * L + c * 32);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float sc = col < L ? fmaf(__uint_as_float(v[j]), p.scale_log2, s_mask[col]) : -INFINITY;
      pr[j] = rl::ex2(sc - mx) * inv;
      float dp = __uint_as_float(w[j]);  // gradient wrt the (dropped-out) probabilities
      if (p.drop.thresh) {
        const bool keep = (keep_bits >> j) & 1u;
        dp = keep ? dp * p.drop.scale : 0.f;
        ds[j] = col < L ? pr[j] * (dp - delta) : 0.f;
        pr[j] = keep ? pr[j] * p.drop.scale : 0.f;  // the tile used for dV = P_drop^T dO
      } else {
        ds[j] = col < L ? pr[j] * (dp - delta) : 0.f;
      }
    }
    uint8_t* tp = ((c >> 1) ? sP1 : sV) + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* td = sDS + (c >> 1) * T16K + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int piece = (((c & 1) * 4 + g) ^ (r & 7)) << 4;
      *reinterpret_cast<uint4*>(tp + piece) =
          make_uint4(rl::pack_h(pr[8 * g], pr[8 * g + 1], p.f16), rl::pack_h(pr[8 * g + 2], pr[8 * g + 3], p.f16),
                     rl::pack_h(pr[8 * g + 4], pr[8 * g + 5], p.f16), rl::pack_h(pr[8 * g + 6], pr[8 * g + 7], p.f16));
      *reinterpret_cast<uint4*>(td + piece) =
          make_uint4(rl::pack_h(ds[8 * g], ds[8 * g + 1], p.f16), rl::pack_h(ds[8 * g + 2], ds[8 * g + 3], p.f16),
                     rl::pack_h(ds[8 * g + 4], ds[8 * g + 5], p.f16), rl::pack_h(ds[8 * g + 6], ds[8 * g + 7], p.f16));
    }
  }
  if (nchunk < 2) {  // kv columns 32..63 of P tile 0 still hold V: clear them (their dV rows are never stored, but
                     // stale bit patterns must not reach the tensor core as NaNs)
    uint8_t* tp = sV + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(tp + (((4 + g) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    rl::tc_fence_after();
    const uint32_t pa = rl::smem_u32(sV), dsa = rl::smem_u32(sDS), qa = rl::smem_u32(sQ), ka = rl::smem_u32(sK),
                   da = rl::smem_u32(sDO);
    // dV[kv, d] = sum_q P[q, kv] dO[q, d] : A = P^T (MN-major: kv contiguous, 64-kv blocks one tile apart), B = dO^T
    const uint32_t idesc_t = rl::make_idesc_h(128, HEAD_DIM, 1, 1, p.f16, p.f16);   // P^T dO
    const uint32_t idesc_k = rl::make_idesc_h(128, HEAD_DIM, 1, 1, p.f16, p.f16);   // dS^T Q
#pragma unroll
    for (int k = 0; k < 8; ++k)
      rl::tc_mma_f16(tmem_base + COL_DV, rl::make_smem_desc_sw128(pa + k * 2048, 2 * T16K, 1024),  // P tiles: sV, sP1
                     rl::make_smem_desc_sw128(da + k * 2048, 1024, 1024), idesc_t, k != 0);
    // dK[kv, d] = sum_q dS[q, kv] Q[q, d]
#pragma unroll
    for (int k = 0; k < 8; ++k)
      rl::tc_mma_f16(tmem_base + COL_DK, rl::make_smem_desc_sw128(dsa + k * 2048, T16K, 1024),
                     rl::make_smem_desc_sw128(qa + k * 2048, 1024, 1024), idesc_k, k != 0);
    // dQ[q, d] = sum_kv dS[q, kv] K[kv, d] : A = dS K-major over kv, B = K^T (MN-major)
    const uint32_t idesc_q = rl::make_idesc_h(128, HEAD_DIM, 0, 1, p.f16, p.f16);   // dS K
    const int nk = lkv16 / 16;
    for (int k = 0; k < nk; ++k)
      rl::tc_mma_f16(tmem_base + COL_DQ, rl::make_smem_desc_sw128(dsa + (k >> 2) * T16K + (k & 3) * 32, 16, 1024),
                     rl::make_smem_desc_sw128(ka + k * 2048, 1024, 1024), idesc_q, k != 0);
    rl::tc_commit(bar_o);
  }
  rl::mbar_wait(bar_o, 0);
  rl::tc_fence_after();
  {
    // dQ / dK / dV leave through the dead Q / K / dO tiles (SWIZZLE_128B rows) and one TMA store each; the 3-D map
    // [B][L][3H] clips the rows beyond the sentence.  (Row-strided 16-byte global stores were 18 % of the stall samples.)
    uint8_t* stage[3] = {sQ, sK, sDO};
    const uint32_t cols[3] = {COL_DQ, COL_DK, COL_DV};
    const float scl[3] = {0.125f, 0.125f, 1.0f};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      uint8_t* srow = stage[t] + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        rl::tmem_ld_32x32(t_row + cols[t] + c * 32, v);
        rl::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(srow + (((c * 4 + g) ^ (r & 7)) << 4)) =
              make_uint4(rl::pack_h(__uint_as_float(v[8 * g]) * scl[t], __uint_as_float(v[8 * g + 1]) * scl[t], p.f16),
                         rl::pack_h(__uint_as_float(v[8 * g + 2]) * scl[t], __uint_as_float(v[8 * g + 3]) * scl[t], p.f16),
                         rl::pack_h(__uint_as_float(v[8 * g + 4]) * scl[t], __uint_as_float(v[8 * g + 5]) * scl[t], p.f16),
                         rl::pack_h(__uint_as_float(v[8 * g + 6]) * scl[t], __uint_as_float(v[8 * g + 7]) * scl[t], p.f16));
      }
    }
  }
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();
  if (p.dbias && tid >= 32) {
    // bias gradient of the fused QKV projection = column sums of the three staged tiles (rows beyond the sentence are exact
    // zeros).  Warps 1-3 take one tile each, a lane two adjacent columns (one conflict-free 4-byte word per row), while
    // thread 0 issues the TMA stores of the same tiles; 192 atomics per CTA.  (A 31-shuffle butterfly per 32-column chunk
    // before the staging cost the kernel 12 %; this costs ~2 %.)
    const int t = (tid >> 5) - 1, cp = tid & 31;
    const uint8_t* src = (t == 0 ? sQ : t == 1 ? sK : sDO) + (cp & 3) * 4;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int row = 0; row < 128; ++row) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(src + (row >> 3) * 1024 + (row & 7) * 128 + (((cp >> 2) ^ (row & 7)) << 4));
      s0 += rl::half_lo(w, p.f16);
      s1 += rl::half_hi(w, p.f16);
    }
    float* o = p.dbias + t * p.H + head * HEAD_DIM + 2 * cp;
    atomicAdd(o, s0);
    atomicAdd(o + 1, s1);
  }
  if (tid == 0) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint8_t* src = t == 0 ? sQ : t == 1 ? sK : sDO;
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&tmDQKV)),
                   "r"(rl::smem_u32(src)), "r"(t * p.H + head * HEAD_DIM), "r"(0), "r"(b)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tiles are read before the CTA's smem goes away
  }
  if (warp == 0) {
    __syncwarp();
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, 256);
  }
}

constexpr int ATT_BWD_SMEM = 7 * T16K + 128 * 4 + 3 * 8 + 16;   // 115,240 B: two CTAs per SM

// ------------------------------------------------------------------------------------------------------------------
// 128 < seq_len <= 256: the same algebra blocked 128 x 128.  One CTA per (sentence, head) keeps both q tiles and both kv
// tiles of Q / K / V / dO resident (8 x 16 KB) and walks the four (kv tile j, q tile i) blocks:
//     S_ij = Q_i K_j^T, dP_ij = dO_i V_j^T  ->  P_ij, dS_ij (smem)  ->  dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j
// dQ_0 / dQ_1 accumulate over j, dK_j / dV_j over i — all in TMEM (512 columns, nothing aliased); the saved logsumexp of
// the forward is required (P = exp2(s - lse) in one pass).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd256_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO, const AttBwdParams p) {
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;   // dQ_i at COL_DQ + 64 i
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((rl::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sQ = smem;                 // [2][128 x 64]
  uint8_t* sK = sQ + 2 * T16K;
  uint8_t* sV = sK + 2 * T16K;
  uint8_t* sDO = sV + 2 * T16K;
  uint8_t* sP = sDO + 2 * T16K;       // P block: two [128 q x 64 kv] tiles
  uint8_t* sDS = sP + 2 * T16K;       // dS block
  float* s_mask = reinterpret_cast<float*>(sDS + 2 * T16K);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + 256);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);
  uint64_t* bar_ld = &bars[0];
  uint64_t* bar_s = &bars[1];
  uint64_t* bar_o = &bars[2];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.x, b = blockIdx.y;
  const int L = p.L;
  const int row0 = b * L;

  if (tid == 0) {
    rl::tma_prefetch_desc(&tmQ);
    rl::tma_prefetch_desc(&tmDO);
    rl::mbar_init(bar_ld, 1);
    rl::mbar_init(bar_s, 1);
    rl::mbar_init(bar_o, 1);
    rl::fence_barrier_init();
  }
  if (warp == 0) rl::tmem_alloc(tmem_ptr, 512);
  for (int j = tid; j < 256; j += ATT_THREADS) {
    float m = -INFINITY;
    if (j < L) m = p.mask[(long long)b * L + j] != 0 ? 0.0f : -10000.0f * 1.4426950408889634f;   // attention masks are 0 / 1 (no I2F.S64)
    s_mask[j] = m;
  }
  for (int i = tid; i < 4 * T16K / 16; i += ATT_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  rl::fence_proxy_async();
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    rl::mbar_expect_tx(bar_ld, 8 * T16K);
#pragma unroll
    for (int t = 0; t < 2; ++t) {   // rows beyond the tensor read zeros; rows of the next sentence are masked out below
      rl::tma_load_2d(sQ + t * T16K, &tmQ, bar_ld, head * HEAD_DIM, row0 + t * 128);
      rl::tma_load_2d(sK + t * T16K, &tmQ, bar_ld, p.H + head * HEAD_DIM, row0 + t * 128);
      rl::tma_load_2d(sV + t * T16K, &tmQ, bar_ld, 2 * p.H + head * HEAD_DIM, row0 + t * 128);
      rl::tma_load_2d(sDO + t * T16K, &tmDO, bar_ld, head * HEAD_DIM, row0 + t * 128);
    }
  }
  // per-row constants of both q tiles: delta = rowsum(dO o O), lse
  float delta[2], lse[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = i * 128 + tid;
    delta[i] = 0.f;
    lse[i] = INFINITY;               // rows beyond the sentence: P = 0
    if (q < L) {
      lse[i] = p.lse[((long long)b * p.heads + head) * L + q];
      const uint4* o = reinterpret_cast<const uint4*>(p.ctx + (long long)(row0 + q) * p.H + head * HEAD_DIM);
      const uint4* g = reinterpret_cast<const uint4*>(p.dctx + (long long)(row0 + q) * p.H + head * HEAD_DIM);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 a = o[k], c = g[k];
        delta[i] += rl::half_lo(a.x, p.f16) * rl::half_lo(c.x, p.f16) + rl::half_hi(a.x, p.f16) * rl::half_hi(c.x, p.f16) +
                    rl::half_lo(a.y, p.f16) * rl::half_lo(c.y, p.f16) + rl::half_hi(a.y, p.f16) * rl::half_hi(c.y, p.f16) +
                    rl::half_lo(a.z, p.f16) * rl::half_lo(c.z, p.f16) + rl::half_hi(a.z, p.f16) * rl::half_hi(c.z, p.f16) +
                    rl::half_lo(a.w, p.f16) * rl::half_lo(c.w, p.f16) + rl::half_hi(a.w, p.f16) * rl::half_hi(c.w, p.f16);
      }
    }
  }
  rl::DropSpec dsp = p.drop;
  rl::drop_resolve(dsp);
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t qa = rl::smem_u32(sQ), ka = rl::smem_u32(sK), va = rl::smem_u32(sV), da = rl::smem_u32(sDO),
                 pa = rl::smem_u32(sP), dsa = rl::smem_u32(sDS);
  const uint32_t idesc_s = rl::make_idesc_h(128, 128, 0, 0, p.f16, p.f16);
  const uint32_t idesc_t = rl::make_idesc_h(128, HEAD_DIM, 1, 1, p.f16, p.f16);
  const uint32_t idesc_q = rl::make_idesc_h(128, HEAD_DIM, 0, 1, p.f16, p.f16);
  uint32_t ph_s = 0, ph_o = 0;
  const int r = tid;
  int blk = 0;
  for (int j = 0; j < 2; ++j) {
    for (int i = 0; i < 2; ++i, ++blk) {
      if (tid == 0) {
        if (blk == 0) {
          rl::mbar_wait(bar_ld, 0);
        }
        rl::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S = Q_i K_j^T
          rl::tc_mma_f16(tmem_base + COL_S, rl::make_smem_desc_sw128(qa + i * T16K + k * 32, 16, 1024),
                         rl::make_smem_desc_sw128(ka + j * T16K + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP = dO_i V_j^T
          rl::tc_mma_f16(tmem_base + COL_DP, rl::make_smem_desc_sw128(da + i * T16K + k * 32, 16, 1024),
                         rl::make_smem_desc_sw128(va + j * T16K + k * 32, 16, 1024), idesc_s, k != 0);
        rl::tc_commit(bar_s);
      }
      rl::mbar_wait(bar_s, ph_s);
      ph_s ^= 1;
      rl::tc_fence_after();
      const int q = i * 128 + r;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], w[32];
        rl::tmem_ld_32x32(t_row + COL_S + c * 32, v);
        rl::tmem_ld_32x32(t_row + COL_DP + c * 32, w);
        rl::tmem_ld_wait();
        float pr[32], ds[32];
        unsigned int keep_bits = 0xFFFFFFFFu;
        const int col0 = j * 128 + c * 32;
        if (dsp.thresh && q < L)
          keep_bits = rl::drop_bits32(dsp, (((unsigned long long)b * p.heads + head) * L + q) * L + col0);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int col = col0 + jj;
          const float sc = col < L ? fmaf(__uint_as_float(v[jj]), p.scale_log2, s_mask[col]) : -INFINITY;
          pr[jj] = rl::ex2(sc - lse[i]);
          float dp = __uint_as_float(w[jj]);
          if (p.drop.thresh) {
            const bool keep = (keep_bits >> jj) & 1u;
            dp = keep ? dp * p.drop.scale : 0.f;
            ds[jj] = col < L ? pr[jj] * (dp - delta[i]) : 0.f;
            pr[jj] = keep ? pr[jj] * p.drop.scale : 0.f;
          } else {
            ds[jj] = col < L ? pr[jj] * (dp - delta[i]) : 0.f;
          }
        }
        uint8_t* tp = sP + (c >> 1) * T16K + (r >> 3) * 1024 + (r & 7) * 128;
        uint8_t* td = sDS + (c >> 1) * T16K + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int piece = (((c & 1) * 4 + g) ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(tp + piece) =
              make_uint4(rl::pack_h(pr[8 * g], pr[8 * g + 1], p.f16), rl::pack_h(pr[8 * g + 2], pr[8 * g + 3], p.f16),
                         rl::pack_h(pr[8 * g + 4], pr[8 * g + 5], p.f16), rl::pack_h(pr[8 * g + 6], pr[8 * g + 7], p.f16));
          *reinterpret_cast<uint4*>(td + piece) =
              make_uint4(rl::pack_h(ds[8 * g], ds[8 * g + 1], p.f16), rl::pack_h(ds[8 * g + 2], ds[8 * g + 3], p.f16),
                         rl::pack_h(ds[8 * g + 4], ds[8 * g + 5], p.f16), rl::pack_h(ds[8 * g + 6], ds[8 * g + 7], p.f16));
        }
      }
      rl::fence_proxy_async();
      rl::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        rl::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV_j[kv, d] (+)= sum_q P[q, kv] dO_i[q, d]
          rl::tc_mma_f16(tmem_base + COL_DV, rl::make_smem_desc_sw128(pa + k * 2048, T16K, 1024),
                         rl::make_smem_desc_sw128(da + i * T16K + k * 2048, 1024, 1024), idesc_t, (i != 0 || k != 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK_j[kv, d] (+)= sum_q dS[q, kv] Q_i[q, d]
          rl::tc_mma_f16(tmem_base + COL_DK, rl::make_smem_desc_sw128(dsa + k * 2048, T16K, 1024),
                         rl::make_smem_desc_sw128(qa + i * T16K + k * 2048, 1024, 1024), idesc_t, (i != 0 || k != 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ_i[q, d] (+)= sum_kv dS[q, kv] K_j[kv, d]
          rl::tc_mma_f16(tmem_base + COL_DQ + i * 64, rl::make_smem_desc_sw128(dsa + (k >> 2) * T16K + (k & 3) * 32, 16, 1024),
                         rl::make_smem_desc_sw128(ka + j * T16K + k * 2048, 1024, 1024), idesc_q, (j != 0 || k != 0));
        rl::tc_commit(bar_o);
      }
      // the next block overwrites S / dP (TMEM) and the P / dS tiles (smem): its MMAs and stores wait for this commit
      rl::mbar_wait(bar_o, ph_o);
      ph_o ^= 1;
      rl::tc_fence_after();
    }
    // dK_j, dV_j complete: rows kv = j*128 + r
    {
      const int kv = j * 128 + r;
      unsigned short* base = reinterpret_cast<unsigned short*>(p.dqkv) + (long long)(row0 + kv) * 3 * p.H + head * HEAD_DIM;
      const uint32_t cols[2] = {COL_DK, COL_DV};
      const float scl[2] = {0.125f, 1.0f};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          rl::tmem_ld_32x32(t_row + cols[t] + c * 32, v);
          rl::tmem_ld_wait();
          if (kv < L) {
            uint4* o = reinterpret_cast<uint4*>(base + (t + 1) * p.H + c * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              o[g] = make_uint4(rl::pack_h(__uint_as_float(v[8 * g]) * scl[t], __uint_as_float(v[8 * g + 1]) * scl[t], p.f16),
                                rl::pack_h(__uint_as_float(v[8 * g + 2]) * scl[t], __uint_as_float(v[8 * g + 3]) * scl[t], p.f16),
                                rl::pack_h(__uint_as_float(v[8 * g + 4]) * scl[t], __uint_as_float(v[8 * g + 5]) * scl[t], p.f16),
                                rl::pack_h(__uint_as_float(v[8 * g + 6]) * scl[t], __uint_as_float(v[8 * g + 7]) * scl[t], p.f16));
          }
        }
      }
      rl::tc_fence_before();
      __syncthreads();   // every thread has read dK_j / dV_j before the next j's MMAs overwrite them
      rl::tc_fence_after();
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = i * 128 + r;
    unsigned short* base = reinterpret_cast<unsigned short*>(p.dqkv) + (long long)(row0 + q) * 3 * p.H + head * HEAD_DIM;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      rl::tmem_ld_32x32(t_row + COL_DQ + i * 64 + c * 32, v);
      rl::tmem_ld_wait();
      if (q < L) {
        uint4* o = reinterpret_cast<uint4*>(base + c * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          o[g] = make_uint4(rl::pack_h(__uint_as_float(v[8 * g]) * 0.125f, __uint_as_float(v[8 * g + 1]) * 0.125f, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 2]) * 0.125f, __uint_as_float(v[8 * g + 3]) * 0.125f, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 4]) * 0.125f, __uint_as_float(v[8 * g + 5]) * 0.125f, p.f16),
                            rl::pack_h(__uint_as_float(v[8 * g + 6]) * 0.125f, __uint_as_float(v[8 * g + 7]) * 0.125f, p.f16));
      }
    }
  }
  rl::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, 512);
  }
}

constexpr int ATT_BWD256_SMEM = 12 * T16K + 256 * 4 + 3 * 8 + 16;   // 197,672 B: one CTA per SM

}  // namespace

extern "C" int rl_attention_bwd(const void* qkv, const int64_t* mask, const void* ctx, const void* dctx, void* dqkv,
                                float* dbias, const float* row_lse, int64_t B, int64_t L, int64_t heads, int64_t head_dim, int32_t act_dtype,
                                float drop_p, uint64_t drop_seed, uint32_t drop_site, const uint64_t* drop_counter,
                                void* stream) {
  RL_REQUIRE(qkv && mask && ctx && dctx && dqkv, RL_EINVAL, "rl_attention_bwd: null pointer");
  RL_REQUIRE(head_dim == HEAD_DIM, RL_EINVAL, "rl_attention_bwd: head_dim must be 64");
  RL_REQUIRE(B > 0 && heads > 0 && L > 0 && L <= 256, RL_EINVAL, "rl_attention_bwd: seq_len %lld not in 1..256", (long long)L);
  RL_REQUIRE(L <= 128 || row_lse, RL_EINVAL, "rl_attention_bwd: seq_len > 128 needs the row_lse saved by rl_attention_fwd");
  RL_REQUIRE(L <= 128 || !dbias, RL_EINVAL, "rl_attention_bwd: the fused bias gradient (dbias) is built for seq_len <= 128");
  const int H = (int)(heads * head_dim);
  const int lkv16 = (int)((L + 15) / 16 * 16);
  CUtensorMap tq, tkv, tdo;
  uint64_t dims[2] = {(uint64_t)(3 * H), (uint64_t)(B * L)};
  uint64_t strides[1] = {(uint64_t)(3 * H) * 2};
  uint32_t boxq[2] = {64, 128};
  uint32_t boxkv[2] = {64, (uint32_t)(lkv16 <= 128 ? lkv16 : 128)};
  int rc = rl_make_tmap_bf16(&tq, qkv, 2, dims, strides, boxq);
  if (rc) return rc;
  rc = rl_make_tmap_bf16(&tkv, qkv, 2, dims, strides, boxkv);
  if (rc) return rc;
  uint64_t dimso[2] = {(uint64_t)H, (uint64_t)(B * L)};
  uint64_t strideso[1] = {(uint64_t)H * 2};
  rc = rl_make_tmap_bf16(&tdo, dctx, 2, dimso, strideso, boxq);
  if (rc) return rc;
  if (L > 128) {
    static std::atomic<bool> configured256{false};
    if (!configured256) {
      cudaError_t e = cudaFuncSetAttribute(attention_bwd256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD256_SMEM);
      if (e != cudaSuccess) {
        rl_set_error("rl_attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return (int)e;
      }
      configured256 = true;
    }
    AttBwdParams p;
    p.mask = reinterpret_cast<const long long*>(mask);
    p.ctx = reinterpret_cast<const __nv_bfloat16*>(ctx);
    p.dctx = reinterpret_cast<const __nv_bfloat16*>(dctx);
    p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
    p.L = (int)L;
    p.H = H;
    p.lkv16 = lkv16;
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.heads = (int)heads;
    p.drop = rl::make_drop(drop_p, drop_seed, drop_site, drop_counter);
    p.f16 = act_dtype == RL_DT_F16;
    p.lse = row_lse;
    p.dbias = nullptr;
    attention_bwd256_kernel<<<dim3((unsigned)heads, (unsigned)B), ATT_THREADS, ATT_BWD256_SMEM, (cudaStream_t)stream>>>(tq, tdo, p);
    return rl_check_launch("rl_attention_bwd(256)");
  }
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM);
    if (e != cudaSuccess) {
      rl_set_error("rl_attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  CUtensorMap to, tdq;
  rc = rl_make_tmap_bf16(&to, ctx, 2, dimso, strideso, boxq);
  if (rc) return rc;
  {
    uint64_t dims3[3] = {(uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    uint64_t strides3[2] = {(uint64_t)(3 * H) * 2, (uint64_t)L * (3 * H) * 2};
    uint32_t box3[3] = {64, 128, 1};
    rc = rl_make_tmap_bf16(&tdq, dqkv, 3, dims3, strides3, box3);
    if (rc) return rc;
  }
  AttBwdParams p;
  p.mask = reinterpret_cast<const long long*>(mask);
  p.ctx = reinterpret_cast<const __nv_bfloat16*>(ctx);
  p.dctx = reinterpret_cast<const __nv_bfloat16*>(dctx);
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.L = (int)L;
  p.H = H;
  p.lkv16 = lkv16;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.heads = (int)heads;
  p.drop = rl::make_drop(drop_p, drop_seed, drop_site, drop_counter);
  p.f16 = act_dtype == RL_DT_F16;
  p.lse = row_lse;
  p.dbias = dbias;
  attention_bwd_kernel<<<dim3((unsigned)heads, (unsigned)B), ATT_THREADS, ATT_BWD_SMEM, (cudaStream_t)stream>>>(tq, tkv, tdo, to, tdq, p);
  return rl_check_launch("rl_attention_bwd");
}
