// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//   out = act((A . B^T) * scale + bias + res),  A:[M,K] bf16 (or 5-D conv activation), B:[N,K] bf16
// Roles per CTA (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 =
// TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> global).  The fp32 accumulator is
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Operands are staged by TMA with SWIZZLE_128B into a STAGES-deep ring of {A 128x64, B BNx64} tiles.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int STAGE_BYTES = 8 * (8192 + 1024);  // per epilogue warp: two 32x32 fp32 staging tiles + scale/bias table

struct KParams {
  int M, N, num_kb;
  int k_splits, kb_per_split;  // split-K: tile index = (m, n, split); splits accumulate with f32 atomics
  int atomic_out;              // split_k requested: out += A*B through atomics even when one split suffices
  int tiles_m, tiles_n;
  int a_mode;
  int hw_shift, w_shift;
  int cin_blocks;
  int8_t tap_dw[12], tap_dh[12], tap_plane[12];
  void* out;
  long long ldo;
  int out_f32;
  __nv_bfloat16* out2;
  long long ldo2;
  const float* scale;
  const float* bias;
  const void* res;
  long long ldr;
  int res_f32;
  int act;
  int out_remap;
  int remap_plane;  // out_remap == 2: rows (img, h, w) of the GEMM go to plane `remap_plane` of a parity-split tensor
  int tma_store;  // epilogue writes through tmC (plain row-major outputs)
  int tma_res;    // residual tiles arrive through tmR (TMA load into the staging tile the result leaves from): per-thread
                  // row loads touch 32 cache lines per instruction and made the f32-residual epilogues L1-bound
  int deep;       // 16-bit result leaving by TMA: every chunk of a tile has its OWN 2 KB staging tile, the stores are issued
                  // back to back and only drained when the warp starts its NEXT tile (a whole main loop later).  With two
                  // alternating tiles each chunk waited for the TMA store of two chunks earlier to finish READING its tile —
                  // behind the main loop's loads that takes thousands of cycles and made every K = 768 GEMM epilogue-bound.
  int res_all;    // deep + tma_res (16-bit residual AND result): every chunk of a tile has its own 2 KB staging tile (see issue_residual)
  int tma_out2;   // GELU_SAVE: the pre-activation tile leaves through tmC2 from the upper half of the staging tile
  int sms;          // SMs this launch may occupy (num_SMs - sm_reserve)
  float* colsum;    // optional [N]: += sum over rows of the result (bias gradients; BatchNorm batch statistics)
  float* colsumsq;  // optional [N]: += sum over rows of result^2
  int vec_store;  // direct path may use 16-byte stores
  rl::DropSpec drop;  // dropout on the linear output before the residual add (BertSelfOutput / BertOutput)
  int a_f16, b_f16;   // operand formats of the MMA: IEEE fp16 instead of bf16 (tcgen05 kind::f16 takes either, per operand)
  int o_f16, r_f16;   // 16-bit outputs (out, out2) / 16-bit residual stored as fp16 instead of bf16
  int b_mode;     // 1: B tiles are gathered from a conv activation (implicit im2col, weight gradients)
  int ks_major;   // tile order: split index outermost, so that the N tiles sharing a K range run side by side (L2 reuse)
  int nimg;       // conv: number of images (an image index >= nimg makes a TMA box read zeros)
  int ntaps;
  int a_mn, b_mn; // operand stored MN-major: A as [K, M] (M contiguous), B as [K, N] (N contiguous)
  uint32_t eflags; // the epilogue's mode switches packed into one word (EF_* below; host: pack_eflags)
  int8_t halo_t[12]; // conv64_halo_kernel: weight tap block of (dw + 1) * 3 + (dh + 1)
};

// Epilogue mode bits.  The per-chunk code of an epilogue warp used to test ~25 KParams fields, each an LDC -> ISETP -> BRA
// chain of 30-90 cycles that two warps per scheduler cannot hide: warp-state sampling of the K = 768 GEMMs (QKV, 16384 x
// 2304 x 768) showed 3.2 k cycles per 32-column chunk for ~260 instructions, the epilogue outlasting the 9 k-cycle main
// loop and the MMA warp blocked on tmem_empty 19 % of the kernel.  The switches now live in ONE register (EpiState::flags,
// loaded once per warp), and the hottest configurations of the 256 x 256 pair kernel are compiled with the word as a
// template constant (SPEC != 0), which folds every test and frees the registers of the paths not taken.
enum : uint32_t {
  EF_TMA_RES = 1u << 0, EF_RES_F32 = 1u << 1, EF_R_F16 = 1u << 2, EF_DROP = 1u << 3, EF_TMA_STORE = 1u << 4,
  EF_DEEP = 1u << 5, EF_RES_ALL = 1u << 6, EF_TMA_OUT2 = 1u << 7, EF_OUT_F32 = 1u << 8, EF_O_F16 = 1u << 9,
  EF_ATOMIC = 1u << 10, EF_SCALE = 1u << 11, EF_RES = 1u << 12, EF_OUT2 = 1u << 13, EF_VEC = 1u << 14,
  EF_COLSUM = 1u << 15, EF_COLSUMSQ = 1u << 16, EF_ACT_SHIFT = 24, EF_VALID = 1u << 31
};
// compile-time specialisations of the pair kernel (training formats: bf16 operands and results)
constexpr uint32_t SPEC_OUT16 = EF_VALID | EF_TMA_STORE | EF_DEEP | EF_VEC;                  // (bias) -> bf16: QKV, dgrads
constexpr uint32_t SPEC_GELU_SAVE = EF_VALID | EF_TMA_STORE | EF_TMA_OUT2 | EF_OUT2 | EF_VEC | ((uint32_t)RL_ACT_GELU_SAVE << EF_ACT_SHIFT);
constexpr uint32_t SPEC_GELU_GRAD = EF_VALID | EF_TMA_STORE | EF_DEEP | EF_TMA_RES | EF_RES_ALL | EF_RES | EF_VEC |
                                    ((uint32_t)RL_ACT_GELU_GRAD << EF_ACT_SHIFT);          // du = (dy W2) o gelu'(u) -> bf16
constexpr uint32_t SPEC_GELU_GRAD_CS = SPEC_GELU_GRAD | EF_COLSUM;                          // ... and db1 = column sums of du
// forward-only (eval) formats: fp16 operands and 16-bit results
constexpr uint32_t SPEC_OUT16_F16 = SPEC_OUT16 | EF_O_F16;                                                      // QKV
constexpr uint32_t SPEC_GELU16_F16 = SPEC_OUT16 | EF_O_F16 | ((uint32_t)RL_ACT_GELU << EF_ACT_SHIFT);            // FFN1 + GELU
constexpr uint32_t SPEC_RES32_ND = EF_VALID | EF_TMA_STORE | EF_TMA_RES | EF_RES | EF_RES_F32 | EF_OUT_F32 | EF_VEC;   // f32 residual -> f32, no dropout (dgrads)
constexpr uint32_t SPEC_SPLITK = EF_VALID | EF_TMA_STORE | EF_OUT_F32 | EF_ATOMIC | EF_VEC;     // split-K weight gradients (TMA reduce-add)
constexpr uint32_t SPEC_RES32 = EF_VALID | EF_TMA_STORE | EF_TMA_RES | EF_RES | EF_RES_F32 | EF_OUT_F32 | EF_DROP | EF_VEC;  // bias + dropout + f32 residual -> f32

using rl::fast_erf;
using rl::gelu_grad;

__device__ __forceinline__ long long remap_row(const KParams& p, int row) {
  if (p.out_remap == 2) {  // data gradient of a stride-2 conv: this GEMM produces one parity plane of dX
    const int hw = 1 << p.hw_shift;
    const long long img = row >> p.hw_shift;
    return (img * 4 + p.remap_plane) * hw + (row & (hw - 1));
  }
  if (p.out_remap != 1) return row;
  const int hw = 1 << p.hw_shift, w = 1 << p.w_shift;
  const int img = row >> p.hw_shift, pix = row & (hw - 1);
  const int oh = pix >> p.w_shift, ow = pix & (w - 1);
  const int h2 = (hw >> p.w_shift) >> 1, w2 = w >> 1;
  return (((long long)img * 4 + (oh & 1) * 2 + (ow & 1)) * h2 + (oh >> 1)) * w2 + (ow >> 1);
}

// ---- epilogue -------------------------------------------------------------------------------------
// One epilogue warp owns the TMEM lanes of its quarter (thread = output row) and one half of the tile's BN
// columns, processed in 32-column chunks.  Everything that does not depend on the accumulator is fetched
// early: scale/bias of the warp's columns go to a small smem table and the first chunk's residual goes to
// registers BEFORE the warp waits for the MMAs of the tile; the residual of chunk c+1 is loaded while chunk c
// is being computed.  Results leave through a swizzled 32x32 smem tile + one TMA store per chunk (coalesced,
// clips the M/N tails) or, for remapped / oddly aligned outputs, through direct row stores.
__device__ __forceinline__ void load_residual(const KParams& p, uint32_t F, int row, bool row_ok, int nb, float (&x)[32]) {
  const bool r_f16 = F & EF_R_F16;
  if ((F & EF_RES) && row_ok && nb < p.N) {
    if (nb + 32 <= p.N) {
      if (F & EF_RES_F32) {
        const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + (long long)row * p.ldr + nb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = r[j];
          x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
        }
      } else {
        const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + (long long)row * p.ldr + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 t = r[j];
          x[8 * j] = rl::half_lo(t.x, r_f16); x[8 * j + 1] = rl::half_hi(t.x, r_f16);
          x[8 * j + 2] = rl::half_lo(t.y, r_f16); x[8 * j + 3] = rl::half_hi(t.y, r_f16);
          x[8 * j + 4] = rl::half_lo(t.z, r_f16); x[8 * j + 5] = rl::half_hi(t.z, r_f16);
          x[8 * j + 6] = rl::half_lo(t.w, r_f16); x[8 * j + 7] = rl::half_hi(t.w, r_f16);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = 0.f;
        if (nb + j < p.N)
          x[j] = (F & EF_RES_F32) ? reinterpret_cast<const float*>(p.res)[(long long)row * p.ldr + nb + j]
                                  : rl::half_lo((uint32_t)reinterpret_cast<const unsigned short*>(p.res)[(long long)row * p.ldr + nb + j], r_f16);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = 0.f;
  }
}

struct EpiState {
  int stg_sel;       // staging tile of the next chunk (alternates per chunk ACROSS tiles)
  uint32_t rphase;   // bit b: phase of the residual-arrival barrier of staging tile b
  // column reductions (p.colsum / p.colsumsq): lane l accumulates column l of each of this warp's chunks across the
  // tiles of the persistent loop and flushes with one atomic per column when the column block changes (a conv GEMM
  // with one N tile flushes ONCE per CTA — per-tile atomics on 64 addresses would serialise in L2)
  float cs[4], cq[4];
  int cs_n0;
  uint32_t flags;    // KParams::eflags in a register (see EF_*)
  int sb_n0;         // column block whose scale/bias the warp's table holds (-1: none): a GEMM with ONE N tile (the convs)
                     // fills it once per CTA instead of paying a global-load latency on every 128-row tile
  int deep_set;      // EF_DEEP with fewer than 4 chunks per warp and tile (BN = 64 / 128): which SET of 2 KB staging tiles this
                     // tile uses.  Rotating over 4 / CH sets lets a tile start while the TMA stores of the previous ones are
                     // still reading shared memory; waiting for them (wait_group.read 0) serialised every 128-row tile of
                     // the short-K conv GEMMs behind a store round trip (~2.5 k cycles per tile, 3.5 TB/s on an HBM-bound layer)
};

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// before a tile reuses its staging set: the stores of every earlier tile that used this set have been read
template <int CH>
__device__ __forceinline__ void deep_tile_begin(bool rotate, int lane) {
  constexpr int SETS = 4 / CH;
  if (lane == 0) {
    if (rotate) bulk_wait_read<(SETS - 1) * CH>();
    else bulk_wait_read<0>();
  }
  __syncwarp();
}
template <int CH>
__device__ __forceinline__ void deep_tile_end(bool rotate, EpiState& st) {
  constexpr int SETS = 4 / CH;
  if (rotate) st.deep_set = (st.deep_set + 1) & (SETS - 1);
}

// scale/bias of this warp's BN/2 columns -> sb[0..BN/2) and sb[128..128+BN/2); first residual chunk -> xr.
// All global loads are issued before the first shared store: the rolled loop this replaces paid one global-load latency
// per 32 columns (4 x ~500 cycles per 256-wide tile, 14 % of the epilogue warps' time in the K = 768 GEMMs).
template <int BN, int NSTG = 2>
__device__ __forceinline__ void epilogue_prefetch(const KParams& p, uint32_t F, float* sb, int row0, int n0, int half, int lane,
                                                  float (&xr)[32], EpiState& st) {
  constexpr int HC = BN / 2;
  const int c0 = n0 + half * HC;
  __syncwarp();
  if (NSTG != 1 && st.sb_n0 != n0) {
    st.sb_n0 = n0;
    float sv[HC / 32], bv[HC / 32];
    const float* bias = p.bias;
#pragma unroll
    for (int k = 0; k < HC / 32; ++k) {
      const int n = c0 + lane + 32 * k;
      sv[k] = ((F & EF_SCALE) && n < p.N) ? __ldg(p.scale + n) : 1.0f;
      bv[k] = (bias && n < p.N) ? __ldg(bias + n) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < HC / 32; ++k) {
      if (F & EF_SCALE) sb[lane + 32 * k] = sv[k];
      sb[128 + lane + 32 * k] = bv[k];
    }
  }
  if (!(F & EF_TMA_RES)) load_residual(p, F, row0 + lane, row0 + lane < p.M, c0, xr);
  __syncwarp();
}


// KParams::eflags as an opaque register value: the compiler may not re-derive it from constant memory at each use
__device__ __forceinline__ uint32_t load_eflags(const KParams& p) {
  uint32_t f = p.eflags;
  asm volatile("mov.b32 %0, %0;" : "+r"(f));
  return f;
}

using rl::warp_colsum32;

template <int CH>
__device__ __forceinline__ void colsum_flush(const KParams& p, EpiState& st, int half, int lane) {
  if (st.cs_n0 < 0) return;
#pragma unroll
  for (int cc = 0; cc < CH; ++cc) {
    const int col = st.cs_n0 + (half * CH + cc) * 32 + lane;
    if (col < p.N) {
      if (p.colsum) atomicAdd(p.colsum + col, st.cs[cc]);
      if (p.colsumsq) atomicAdd(p.colsumsq + col, st.cq[cc]);
    }
    st.cs[cc] = 0.f;
    st.cq[cc] = 0.f;
  }
  st.cs_n0 = -1;
}

// lane 0: fetch the residual tile of the chunk at column nb into staging tile `buf` (its previous TMA store must be done)
// Two schemes: (a) 4 KB tiles (f32 residual / f32 result): two tiles, the next chunk's residual is fetched while this chunk
// is processed; (b) all-16-bit (EF_RES_ALL): the tile's CH <= 4 residual chunks fit the 8 KB staging region as 2 KB tiles and
// are ALL fetched when the warp starts on the tile, a whole main loop before they are needed (one chunk of lead does not
// cover the ~1 us TMA latency).
__device__ __forceinline__ void issue_residual(uint32_t F, const CUtensorMap* tmR_ptr, uint8_t* stg_base,
                                               uint64_t* rbar, int buf, int nb, int row0) {
  rl::mbar_expect_tx(&rbar[buf], (F & EF_RES_F32) ? 4096u : 2048u);
  rl::tma_load_2d(stg_base + buf * ((F & EF_RES_ALL) ? 2048 : 4096), tmR_ptr, &rbar[buf], nb, row0);
}

// x = acc * scale + bias for the 32 columns of a chunk.  Scale/bias come from the warp's shared-memory table, or (NSTG == 1:
// the long-K variant has no table — its shared memory went into pipeline stages) straight from global memory
// (warp-uniform addresses, L1 hits).  Without a scale vector (everything but the BatchNorm-folded convs) it is one add.
template <int NSTG>
__device__ __forceinline__ void affine32(const KParams& p, uint32_t F, const float* sb, int cc, int nb,
                                         const uint32_t (&v)[32], float (&x)[32]) {
  if (NSTG == 1) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = nb + j;
      if (n + 4 <= p.N) {
        if (F & EF_SCALE) sc = __ldg(reinterpret_cast<const float4*>(p.scale + n));
        if (p.bias) bi = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      } else {
        float s4[4] = {1.f, 1.f, 1.f, 1.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (n + k < p.N) {
            if (F & EF_SCALE) s4[k] = __ldg(p.scale + n + k);
            if (p.bias) b4[k] = __ldg(p.bias + n + k);
          }
        sc = make_float4(s4[0], s4[1], s4[2], s4[3]);
        bi = make_float4(b4[0], b4[1], b4[2], b4[3]);
      }
      x[j] = fmaf(__uint_as_float(v[j]), sc.x, bi.x);
      x[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, bi.y);
      x[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, bi.z);
      x[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, bi.w);
    }
  } else if (F & EF_SCALE) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 sc = *reinterpret_cast<const float4*>(sb + cc * 32 + j);
      const float4 bi = *reinterpret_cast<const float4*>(sb + 128 + cc * 32 + j);
      x[j] = fmaf(__uint_as_float(v[j]), sc.x, bi.x);
      x[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, bi.y);
      x[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, bi.z);
      x[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, bi.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bi = *reinterpret_cast<const float4*>(sb + 128 + cc * 32 + j);
      x[j] = __uint_as_float(v[j]) + bi.x;
      x[j + 1] = __uint_as_float(v[j + 1]) + bi.y;
      x[j + 2] = __uint_as_float(v[j + 2]) + bi.z;
      x[j + 3] = __uint_as_float(v[j + 3]) + bi.w;
    }
  }
}

// One output tile of an epilogue warp.  SPEC != 0: the mode word is a compile-time constant (see EF_*).  `release()` hands
// the accumulator buffer back to the MMA warp; it is called as soon as the LAST chunk has left TMEM, so the math and the
// stores of that chunk overlap the MMAs of the tile after next.
template <int BN, bool COLS, int NSTG = 2, uint32_t SPEC = 0, class Release>
__device__ __forceinline__ void epilogue_tile(const KParams& p, const CUtensorMap* tmC_ptr, const CUtensorMap* tmC2_ptr,
                                              const CUtensorMap* tmR_ptr, uint8_t* stg_base, uint64_t* rbar,
                                              const float* sb, uint32_t taddr, int row0, int n0, int half, int lane,
                                              float (&xr)[32], EpiState& st, Release release) {
  constexpr int CH = BN / 64;  // 32-column chunks per half
  const uint32_t F = SPEC ? SPEC : st.flags;
  const int act = (int)((F >> EF_ACT_SHIFT) & 7u);
  const bool tma_res = F & EF_TMA_RES, res_all = F & EF_RES_ALL, deep = F & EF_DEEP, tma_store = F & EF_TMA_STORE;
  const bool tma_out2 = F & EF_TMA_OUT2, out_f32 = F & EF_OUT_F32, o_f16 = F & EF_O_F16, r_f16 = F & EF_R_F16;
  const bool atomic_out = F & EF_ATOMIC, vec_store = F & EF_VEC, has_out2 = F & EF_OUT2, has_res = F & EF_RES;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const long long orow = remap_row(p, row);
#pragma unroll 1
  for (int cc = 0; cc < CH; ++cc) {
    const int c = half * CH + cc;
    const int nb = n0 + c * 32;
    const bool live = nb < p.N && row0 < p.M;  // warp-uniform
    uint32_t v[32];
    rl::tmem_ld_32x32(taddr + c * 32, v);
    float xn[32];
    if (tma_res) {
      // residual of THIS chunk: TMA delivered it into the staging tile the result will leave from (issued one chunk
      // ahead); read my row, then prefetch the next chunk's tile into the other staging tile
      const int buf = NSTG == 1 ? 0 : res_all ? cc : (st.stg_sel & 1);
      if (live) {
        if (NSTG == 1) {
          // single staging tile: fetch THIS chunk's residual once the previous chunk's store has drained (the latency is
          // hidden behind the >= 24 k-block main loop of the next tile)
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_residual(F, tmR_ptr, stg_base, rbar, 0, nb, row0);
          }
        } else if (!res_all && cc + 1 < CH && nb + 32 < p.N && lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store that last read tile buf^1
          issue_residual(F, tmR_ptr, stg_base, rbar, buf ^ 1, nb + 32, row0);
        }
        rl::mbar_wait(&rbar[buf], (st.rphase >> buf) & 1u);
        st.rphase ^= 1u << buf;
        const uint8_t* rt = stg_base + buf * (res_all ? 2048 : 4096);
        if (F & EF_RES_F32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(rt + lane * 128 + ((g ^ (lane & 7)) << 4));
            xr[4 * g] = t.x; xr[4 * g + 1] = t.y; xr[4 * g + 2] = t.z; xr[4 * g + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 t = *reinterpret_cast<const uint4*>(rt + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4));
            xr[8 * g] = rl::half_lo(t.x, r_f16); xr[8 * g + 1] = rl::half_hi(t.x, r_f16);
            xr[8 * g + 2] = rl::half_lo(t.y, r_f16); xr[8 * g + 3] = rl::half_hi(t.y, r_f16);
            xr[8 * g + 4] = rl::half_lo(t.z, r_f16); xr[8 * g + 5] = rl::half_hi(t.z, r_f16);
            xr[8 * g + 6] = rl::half_lo(t.w, r_f16); xr[8 * g + 7] = rl::half_hi(t.w, r_f16);
          }
        }
        __syncwarp();   // every lane has its residual row in registers before anyone overwrites the tile with results
      }
    } else if (has_res && cc + 1 < CH) {
      load_residual(p, F, row, row_ok, nb + 32, xn);  // overlaps the TMEM load and this chunk's math
    }
    rl::tmem_ld_wait();
    if (cc == CH - 1) {   // the tile has left TMEM: the MMA warp may start the tile after next on this buffer
      rl::tc_fence_before();
      __syncwarp();
      release();
    }
    if (live) {
      float x[32];
      affine32<NSTG>(p, F, sb, cc, nb, v, x);
      if (act == RL_ACT_GELU_GRAD) {
        // data gradient through GELU: the `res` operand carries the saved pre-activation u
#pragma unroll
        for (int j = 0; j < 32; j += 2) {   // packed f32x2 polynomial (FFMA2 / FMUL2): half the issue slots of the scalar form
          const rl::f2 g = rl::gelu_grad2(rl::f2{xr[j], xr[j + 1]});
          x[j] *= g.x;
          x[j + 1] *= g.y;
        }
      } else {
        if (F & EF_DROP) {
          const unsigned long long e0 = (unsigned long long)row * p.N + nb;
          rl::DropSpec dsp = p.drop;
          rl::drop_resolve(dsp);
          rl::drop_apply32(dsp, e0, x);
        }
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += xr[j];
        }
      }
      uint8_t* stg = nullptr;
      if (tma_store) {
        // two swizzled staging tiles per warp, alternated per CHUNK ACROSS TILES (st.stg_sel lives in the tile loop: a
        // BN = 64 tile has one chunk per warp, so alternating on the chunk index alone reused the tile the previous
        // store was still reading): the TMA store issued two chunks ago must have finished reading
        stg = NSTG == 1 ? stg_base : deep ? stg_base + (st.deep_set * CH + cc) * 2048 : stg_base + (st.stg_sel & 1) * 4096;
        st.stg_sel ^= 1;
        if (NSTG == 1) {
          if (!tma_res) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
        } else if (!tma_res && !deep) {   // (tma_res / deep: the tile's previous store was drained earlier)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
      }
      if (act == RL_ACT_GELU_SAVE && tma_out2) {
        // training forward: the pre-activation tile (16-bit) goes to the upper half of the staging tile
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(stg + 2048 + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(rl::pack_h(x[8 * g], x[8 * g + 1], o_f16), rl::pack_h(x[8 * g + 2], x[8 * g + 3], o_f16),
                         rl::pack_h(x[8 * g + 4], x[8 * g + 5], o_f16), rl::pack_h(x[8 * g + 6], x[8 * g + 7], o_f16));
      } else if (act == RL_ACT_GELU_SAVE && row_ok && has_out2) {
        // training forward: keep the pre-activation (bf16) for the backward pass, then activate
        if (nb + 32 <= p.N && vec_store) {
          uint4* o = reinterpret_cast<uint4*>(p.out2 + orow * p.ldo2 + nb);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_uint4(rl::pack_h(x[8 * j], x[8 * j + 1], o_f16), rl::pack_h(x[8 * j + 2], x[8 * j + 3], o_f16),
                              rl::pack_h(x[8 * j + 4], x[8 * j + 5], o_f16), rl::pack_h(x[8 * j + 6], x[8 * j + 7], o_f16));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) reinterpret_cast<unsigned short*>(p.out2)[orow * p.ldo2 + nb + j] = (unsigned short)(rl::pack_h(x[j], 0.f, o_f16) & 0xFFFFu);
        }
      }
      if (act == RL_ACT_GELU || act == RL_ACT_GELU_SAVE) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const rl::f2 g = rl::gelu_erf2(rl::f2{x[j], x[j + 1]});
          x[j] = g.x;
          x[j + 1] = g.y;
        }
      } else if (act == RL_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
      } else if (act == RL_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = tanhf(x[j]);
      }
      if (COLS) {
        if (st.cs_n0 != n0) {
          colsum_flush<CH>(p, st, half, lane);
          st.cs_n0 = n0;
        }
        float t[32];
        if (F & EF_COLSUM) {
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = row_ok ? x[j] : 0.f;
          const float tot = warp_colsum32(t, lane);
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (k == cc) st.cs[k] += tot;
        }
        if (F & EF_COLSUMSQ) {
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = row_ok ? x[j] * x[j] : 0.f;
          const float tot = warp_colsum32(t, lane);
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (k == cc) st.cq[k] += tot;
        }
      }
      if (tma_store) {
        if (out_f32) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(rl::pack_h(x[8 * g], x[8 * g + 1], o_f16), rl::pack_h(x[8 * g + 2], x[8 * g + 3], o_f16),
                           rl::pack_h(x[8 * g + 4], x[8 * g + 5], o_f16), rl::pack_h(x[8 * g + 6], x[8 * g + 7], o_f16));
        }
        rl::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (atomic_out)  // split-K: the tile is ADDED to the (zero-initialised) f32 output by the TMA unit itself
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(tmC_ptr)),
                         "r"(rl::smem_u32(stg)), "r"(nb), "r"(row0)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(tmC_ptr)),
                         "r"(rl::smem_u32(stg)), "r"(nb), "r"(row0)
                         : "memory");
          if (tma_out2)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(tmC2_ptr)),
                         "r"(rl::smem_u32(stg + 2048)), "r"(nb), "r"(row0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (atomic_out) {
        // split-K: partial sums of the different K ranges meet in the (zero-initialised) f32 output
        if (row_ok) {
          float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + nb;
          if (nb + 32 <= p.N && vec_store) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              atomicAdd(reinterpret_cast<float4*>(o) + j, make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) atomicAdd(o + j, x[j]);
          }
        }
      } else if (row_ok) {
        if (nb + 32 <= p.N && vec_store) {
          if (out_f32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + nb);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = make_uint4(rl::pack_h(x[8 * j], x[8 * j + 1], o_f16), rl::pack_h(x[8 * j + 2], x[8 * j + 3], o_f16),
                                rl::pack_h(x[8 * j + 4], x[8 * j + 5], o_f16), rl::pack_h(x[8 * j + 6], x[8 * j + 7], o_f16));
          }
          if (has_out2 && act != RL_ACT_GELU_SAVE) {
            uint4* o = reinterpret_cast<uint4*>(p.out2 + orow * p.ldo2 + nb);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = make_uint4(rl::pack_h(x[8 * j], x[8 * j + 1], o_f16), rl::pack_h(x[8 * j + 2], x[8 * j + 3], o_f16),
                                rl::pack_h(x[8 * j + 4], x[8 * j + 5], o_f16), rl::pack_h(x[8 * j + 6], x[8 * j + 7], o_f16));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nb + j < p.N) {
              if (out_f32)
                reinterpret_cast<float*>(p.out)[orow * p.ldo + nb + j] = x[j];
              else
                reinterpret_cast<unsigned short*>(p.out)[orow * p.ldo + nb + j] = (unsigned short)(rl::pack_h(x[j], 0.f, o_f16) & 0xFFFFu);
              if (has_out2 && act != RL_ACT_GELU_SAVE) reinterpret_cast<unsigned short*>(p.out2)[orow * p.ldo2 + nb + j] = (unsigned short)(rl::pack_h(x[j], 0.f, o_f16) & 0xFFFFu);
            }
          }
        }
      }
    }
    if (has_res && cc + 1 < CH && !tma_res) {
#pragma unroll
      for (int j = 0; j < 32; ++j) xr[j] = xn[j];
    }
  }
}

template <int BN, int STAGES, bool COLS, uint32_t SPEC = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                 const __grid_constant__ CUtensorMap tmR, const KParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  // pad to 1024 B by pointer arithmetic (keeps the shared address space visible to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (rl::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_stage = smem_b + STAGES * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 2);   // [8 epilogue warps][up to 4 staging tiles]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    rl::tma_prefetch_desc(&tmA);
    rl::tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 32; ++s) rl::mbar_init(&res_bar[s], 1);
    for (int s = 0; s < STAGES; ++s) {
      rl::mbar_init(&full_bar[s], 1);
      rl::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      rl::mbar_init(&tmem_full[s], 1);
      rl::mbar_init(&tmem_empty[s], 8);
    }
    rl::fence_barrier_init();
  }
  if (warp == 2) rl::tmem_alloc(tmem_ptr, TMEM_COLS);
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tiles_mn = p.tiles_m * p.tiles_n;
        const int mn = p.ks_major ? tile % tiles_mn : tile / p.k_splits;
        const int ks = p.ks_major ? tile / tiles_mn : tile - mn * p.k_splits;
        const int m_blk = mn / p.tiles_n;
        const int n_blk = mn - m_blk * p.tiles_n;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int m0 = m_blk * BM;
        const int n0 = n_blk * BN;
        int img0 = 0, h0 = 0;
        if (p.a_mode == 1) {
          img0 = m0 >> p.hw_shift;
          h0 = (m0 & ((1 << p.hw_shift) - 1)) >> p.w_shift;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          rl::mbar_wait(&empty_bar[stage], phase ^ 1);
          {
          rl::mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          if (p.a_mn) {
            rl::tma_load_2d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], m0, kb * BK);
            rl::tma_load_2d(smem_a + stage * A_BYTES + 8192, &tmA, &full_bar[stage], m0 + 64, kb * BK);
          } else if (p.a_mode == 0) {
            rl::tma_load_2d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
            const int t = kb / p.cin_blocks;
            const int cb = kb - t * p.cin_blocks;
            rl::tma_load_5d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], cb * BK,
                            (int)p.tap_dw[t], h0 + (int)p.tap_dh[t], (int)p.tap_plane[t], img0);
          }
          if (p.b_mode == 1) {
            // implicit im2col: k-block = 64 consecutive output pixels, each 64-wide N block = 64 channels of one tap
            const int pix0 = kb * BK;
            const int imgk = pix0 >> p.hw_shift;
            const int hk = (pix0 & ((1 << p.hw_shift) - 1)) >> p.w_shift;
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) {
              const int nb = (n0 >> 6) + b;
              const int t = nb / p.cin_blocks;
              const int cb = nb - t * p.cin_blocks;
              if (t < p.ntaps)
                rl::tma_load_5d(smem_b + stage * B_BYTES + b * 8192, &tmB, &full_bar[stage], cb * 64, (int)p.tap_dw[t],
                                hk + (int)p.tap_dh[t], (int)p.tap_plane[t], imgk);
              else  // N tail of the last tile: a box beyond the last image reads zeros (and still counts its bytes)
                rl::tma_load_5d(smem_b + stage * B_BYTES + b * 8192, &tmB, &full_bar[stage], 0, 0, 0, 0, p.nimg);
            }
          } else if (p.b_mn) {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              rl::tma_load_2d(smem_b + stage * B_BYTES + b * 8192, &tmB, &full_bar[stage], n0 + b * 64, kb * BK);
          } else {
            rl::tma_load_2d(smem_b + stage * B_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
          }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer =====================
      // The whole warp runs the loop converged (addresses stay in uniform registers); one elected lane issues.
      const uint32_t idesc = rl::make_idesc_h(BM, BN, p.a_mn, p.b_mn, p.a_f16, p.b_f16);
      const uint32_t a_base = rl::smem_u32(smem_a), b_base = rl::smem_u32(smem_b);
      // K-major: 128-byte rows, 8-row atoms 1024 B apart, a k-step of 16 is +32 B.  MN-major: 64-wide MN blocks
      // 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), a k-step of 16 rows is +2048 B.
      const uint32_t a_lbo = p.a_mn ? 8192 : 16, b_lbo = p.b_mn ? 8192 : 16;
      const uint32_t a_kstep = p.a_mn ? 128 : 2, b_kstep = p.b_mn ? 128 : 2;  // in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ks = p.ks_major ? tile / (p.tiles_m * p.tiles_n) : tile % p.k_splits;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        rl::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        rl::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          rl::mbar_wait(&full_bar[stage], phase);
          rl::tc_fence_after();
          if (rl::elect_one()) {
            const uint64_t adesc = rl::make_smem_desc_sw128(a_base + stage * A_BYTES, a_lbo, 1024);
            const uint64_t bdesc = rl::make_smem_desc_sw128(b_base + stage * B_BYTES, b_lbo, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              rl::tc_mma_f16(d_tmem, adesc + a_kstep * k, bdesc + b_kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            rl::tc_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (rl::elect_one()) rl::tc_commit(&tmem_full[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // Warp w may touch TMEM lanes 32*(w%4)..+31 (thread = output row); the two warps of a lane
    // quarter split the BN columns in halves.  Per 32-column chunk: TMEM -> registers, fused
    // scale/bias/residual/activation with 16-byte accesses along the thread's own row, then either
    //  (a) a swizzled 32x32 smem tile + one TMA store per warp (coalesced, clips the M/N tails), or
    //  (b) direct 16-byte row stores (remapped conv outputs, odd strides, the optional bf16 copy).
    const int ew = warp - 4;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* stg = smem_stage + ew * (8192 + 1024);
    float* sb = reinterpret_cast<float*>(stg + 8192);
    int acc = 0;
    EpiState est{0, 0u, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, -1, SPEC ? SPEC : load_eflags(p), -1, 0};
    const uint32_t F = SPEC ? SPEC : est.flags;
    uint64_t* rbar = res_bar + ew * 4;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = p.ks_major ? tile % (p.tiles_m * p.tiles_n) : tile / p.k_splits;
      const int m_blk = mn / p.tiles_n;
      const int n_blk = mn - m_blk * p.tiles_n;
      const int row0 = m_blk * BM + q * 32;
      const int n0 = n_blk * BN;
      float xr[32];
      epilogue_prefetch<BN>(p, F, sb, row0, n0, half, lane, xr, est);
      // staging-set rotation (see EpiState::deep_set): every chunk of every tile must issue a store, i.e. no N tail
      const bool rotate = BN < 256 && (F & EF_DEEP) && !(F & EF_TMA_RES) && p.N % BN == 0 && p.M % BM == 0;
      if ((F & EF_DEEP) && !(F & EF_TMA_RES)) deep_tile_begin<BN / 64>(rotate, lane);   // the staging tiles about to be reused have been read
      if ((F & EF_TMA_RES) && lane == 0 && row0 < p.M && n0 + half * (BN / 2) < p.N) {
        if (F & EF_RES_ALL) {
          // every residual chunk of this output tile, now: the stores of the previous tile have long drained
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < BN / 64; ++cc)
            if (n0 + half * (BN / 2) + cc * 32 < p.N) issue_residual(F, &tmR, stg, rbar, cc, n0 + half * (BN / 2) + cc * 32, row0);
        } else {
          // first residual tile of this output tile: its staging tile was last read by the store two chunks ago
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          issue_residual(F, &tmR, stg, rbar, est.stg_sel & 1, n0 + half * (BN / 2), row0);
        }
      }
      rl::mbar_wait(&tmem_full[acc], acc_phase);
      rl::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      uint64_t* done = &tmem_empty[acc];
      epilogue_tile<BN, COLS, 2, SPEC>(p, &tmC, &tmC2, &tmR, stg, rbar, sb, taddr, row0, n0, half, lane, xr, est,
                                       [done, lane] { if (lane == 0) rl::mbar_arrive(done); });
      deep_tile_end<BN / 64>(rotate, est);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (COLS) colsum_flush<BN / 64>(p, est, half, lane);
    if ((F & EF_TMA_STORE) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  rl::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return STAGES * (A_BYTES + BN * BK * 2) + STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + 256 + 1024;
}

template <int BN, int STAGES, bool COLS = false, uint32_t SPEC = 0>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                const CUtensorMap& tmR, const KParams& p, cudaStream_t st) {
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, COLS, SPEC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      rl_set_error("rl_gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int tiles = p.tiles_m * p.tiles_n * p.k_splits;
  const int grid = tiles < p.sms ? tiles : p.sms;
  gemm_bf16_kernel<BN, STAGES, COLS, SPEC><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, tmC2, tmR, p);
  return rl_check_launch("rl_gemm_bf16");
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolution, 64 -> 64 channels on 16 x 16 maps, stride 1 (res_block1.conv2 and its data gradient: 4.19 M output
// pixels per train step, the two longest launches of the step).  As nine-tap implicit GEMM through the generic kernel
// every 128-pixel tile pulled 9 x 16 KB of (overlapping) activation windows and 9 x 8 KB of weights through L2 — 216 KB
// per tile at the ~5 KB/clk the L2 delivers chip-wide, 3x the HBM time of the layer.  Here
//   * the weights (72 KB) are loaded ONCE per CTA and stay in shared memory;
//   * per dw in {-1, 0, +1} ONE TMA box of 10 image rows (h0-1 .. h0+8) x 16 pixels x 64 channels (20 KB) lands in a stage,
//     and the three dh taps are the SAME buffer read through UMMA descriptors offset by dh * 16 pixels * 128 B = 2 KB
//     (a multiple of the 1 KB swizzle period, so the plain SWIZZLE_128B K-major descriptor applies);
// 60 KB per tile instead of 216 KB.  Image borders are TMA out-of-bounds zeros, as in the generic conv path.
// Roles and epilogue are those of gemm_bf16_kernel (BN = 64; results leave as 16-bit tiles through TMA: SPEC_OUT16).
constexpr int CH_STAGES = 5;
constexpr int CH_A_BYTES = 160 * 128;
constexpr int CH_B_BYTES = 9 * 64 * 128;
constexpr int CH_WARP_STG = 5120;   // per epilogue warp: two 2 KB result tiles (alternating per tile) + 1 KB scale/bias table
constexpr int CONV_HALO_SMEM = CH_B_BYTES + CH_STAGES * CH_A_BYTES + 8 * CH_WARP_STG + (2 * CH_STAGES + 5) * 8 + 16 + 32 * 8;

__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv64_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const KParams p) {
  constexpr int BN = 64;
  constexpr uint32_t SPEC = SPEC_OUT16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((rl::smem_u32(smem_raw) & 1023u) != 0u) __trap();   // SWIZZLE_128B tiles need 1 KB alignment (no slack is budgeted)
  uint8_t* smem_b = smem_raw;
  uint8_t* smem_a = smem_b + CH_B_BYTES;
  uint8_t* smem_stage = smem_a + CH_STAGES * CH_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + 8 * CH_WARP_STG);
  uint64_t* empty_bar = full_bar + CH_STAGES;
  uint64_t* tmem_full = empty_bar + CH_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_bar + 1);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 2);   // (unused by this epilogue mode; the interface wants it)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    rl::tma_prefetch_desc(&tmA);
    rl::tma_prefetch_desc(&tmB);
    rl::tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      rl::mbar_init(&full_bar[s], 1);
      rl::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      rl::mbar_init(&tmem_full[s], 1);
      rl::mbar_init(&tmem_empty[s], 8);
    }
    rl::mbar_init(b_bar, 1);
    rl::fence_barrier_init();
  }
  if (warp == 2) rl::tmem_alloc(tmem_ptr, 128);
  rl::tc_fence_before();
  __syncthreads();
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int num_tiles = p.tiles_m;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      rl::mbar_expect_tx(b_bar, CH_B_BYTES);
      for (int t = 0; t < 9; ++t) rl::tma_load_2d(smem_b + t * 8192, &tmB, b_bar, t * BK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * BM;
        const int img0 = m0 >> 8, h0 = (m0 & 255) >> 4;
        for (int dwi = 0; dwi < 3; ++dwi) {
          rl::mbar_wait(&empty_bar[stage], phase ^ 1);
          rl::mbar_expect_tx(&full_bar[stage], CH_A_BYTES);
          rl::tma_load_5d(smem_a + stage * CH_A_BYTES, &tmA, &full_bar[stage], 0, dwi - 1, h0 - 1, 0, img0);
          if (++stage == CH_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = rl::make_idesc_h(BM, BN, 0, 0, p.a_f16, p.b_f16);
    const uint32_t a_base = rl::smem_u32(smem_a), b_base = rl::smem_u32(smem_b);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    rl::mbar_wait(b_bar, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      rl::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      rl::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int dwi = 0; dwi < 3; ++dwi) {
        rl::mbar_wait(&full_bar[stage], phase);
        rl::tc_fence_after();
        if (rl::elect_one()) {
#pragma unroll
          for (int dhi = 0; dhi < 3; ++dhi) {
            const int t = p.halo_t[dwi * 3 + dhi];
            const uint64_t adesc = rl::make_smem_desc_sw128(a_base + stage * CH_A_BYTES + dhi * 2048, 16, 1024);
            const uint64_t bdesc = rl::make_smem_desc_sw128(b_base + t * 8192, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              rl::tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dwi > 0 || dhi > 0 || k > 0) ? 1u : 0u);
          }
          rl::tc_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == CH_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (rl::elect_one()) rl::tc_commit(&tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 4;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* stg = smem_stage + ew * CH_WARP_STG;
    float* sb = reinterpret_cast<float*>(stg + 4096);
    int acc = 0;
    EpiState est{0, 0u, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, -1, SPEC, -1, 0};
    uint64_t* rbar = res_bar + ew * 4;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row0 = tile * BM + q * 32;
      float xr[32];
      epilogue_prefetch<BN>(p, SPEC, sb, row0, 0, half, lane, xr, est);
      if (lane == 0) bulk_wait_read<1>();   // the store of the tile before the previous one has read this staging tile
      __syncwarp();
      rl::mbar_wait(&tmem_full[acc], acc_phase);
      rl::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      uint64_t* done = &tmem_empty[acc];
      epilogue_tile<BN, false, 2, SPEC>(p, &tmC, &tmC, &tmC, stg, rbar, sb, taddr, row0, 0, half, lane, xr, est,
                                        [done, lane] { if (lane == 0) rl::mbar_arrive(done); });
      est.deep_set ^= 1;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  rl::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    rl::tc_fence_after();
    rl::tmem_dealloc(tmem_base, 128);
  }
}

int launch_conv64_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const KParams& p, cudaStream_t st) {
  static_assert(CONV_HALO_SMEM <= 232448, "dynamic shared memory of conv64_halo_kernel exceeds 227 KB");
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv64_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_HALO_SMEM);
    if (e != cudaSuccess) {
      rl_set_error("rl_gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.tiles_m < p.sms ? p.tiles_m : p.sms;
  conv64_halo_kernel<<<grid, GEMM_THREADS, CONV_HALO_SMEM, st>>>(tmA, tmB, tmC, p);
  return rl_check_launch("rl_gemm_bf16(conv 3x3 c64 halo)");
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA stages its own
// 128 rows of A and HALF of the B tile, the leader's single thread issues M=256 tcgen05.mma that read both
// CTAs' shared memory and write 128 accumulator lanes in each CTA's TMEM.  Per SM this cuts the operand bytes
// pulled through L2 per flop by 1.5x (the 1-CTA kernel is bound by L2->SM bandwidth, not by the tensor pipe).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in CTA 0 of the pair

__device__ __forceinline__ void tma2_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(rl::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(rl::smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(rl::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(rl::smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
// 2-D tile load delivered to the same CTA-relative smem offset of every CTA in `mask` (multicast through the cluster);
// the bytes are credited to the full barrier of each destination's pair leader
__device__ __forceinline__ void tma2_load_2d_mc(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(rl::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(rl::smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar, uint16_t mask = 3) {  // arrive on the same barrier in every CTA of mask
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          rl::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // from either CTA of the pair
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rl::smem_u32(bar) & kPeerBitMask) : "memory");
}

// CL = CTAs per cluster: 2 = one pair; 4 = two pairs stacked along M that SHARE the B tile: each CTA fetches only a
// quarter of it and multicasts the quarter to its counterpart in the other pair, so a CTA pulls 24 KB instead of 32 KB
// through L2 per 64-deep k-block (the pair kernel is bound by L2 -> SM bytes, not by the tensor pipe).
// NSTG = 32x32 fp32 staging tiles per epilogue warp: 2 (default), or 1 for the LONG-K variant, which spends the shared
// memory on a fifth pipeline stage instead: ncu shows the main loop bound by bytes in flight (4 x 32 KB per SM at ~2 us of
// TMA latency = 900 clk per k-block, 57 % tensor-pipe), and behind >= 24 k-blocks a serialised epilogue is invisible.
// SPEC != 0: the epilogue's mode word as a compile-time constant (EF_*, SPEC_*).
template <int BN, int STAGES, int CL, bool COLS, int NSTG = 2, uint32_t SPEC = 0>
__device__ __forceinline__ void gemm2_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                           const CUtensorMap& tmC2, const CUtensorMap& tmR, const KParams& p) {
  constexpr int BH_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the B tile
  constexpr int PAIRS = CL / 2;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // NSTG == 2: pad to 1024 B by pointer arithmetic.  NSTG == 1 (long-K variant) has no slack to pad with: the dynamic
  // shared-memory window of a kernel without static shared memory starts 1024-byte aligned (trap if it ever does not)
  uint8_t* smem = smem_raw + (NSTG == 1 ? 0u : ((1024u - (rl::smem_u32(smem_raw) & 1023u)) & 1023u));
  if (NSTG == 1 && (rl::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_stage = smem_b + STAGES * BH_BYTES;
  constexpr int WARP_STG = NSTG == 1 ? 4096 : NSTG * 4096 + 1024;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + 8 * WARP_STG);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 2);   // [8 epilogue warps][up to 4 staging tiles]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // 0 .. CL-1
  const uint32_t rank = crank & 1;            // rank inside the CTA pair
  const uint32_t pair = crank >> 1;           // pair inside the cluster (0 when CL == 2)
  const bool leader = rank == 0;
  const uint16_t all_mask = (uint16_t)((1u << CL) - 1u);
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const uint16_t bcast_mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));   // CL == 4: my counterpart in the other pair

  if (warp == 0 && lane == 0) {
    rl::tma_prefetch_desc(&tmA);
    rl::tma_prefetch_desc(&tmB);
    rl::tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 32; ++s) rl::mbar_init(&res_bar[s], 1);
    for (int s = 0; s < STAGES; ++s) {
      rl::mbar_init(&full_bar[s], 1);
      rl::mbar_init(&empty_bar[s], PAIRS);   // a stage is free once EVERY pair that reads it (own A, shared B) has consumed it
    }
    for (int s = 0; s < 2; ++s) {
      rl::mbar_init(&tmem_full[s], 1);
      rl::mbar_init(&tmem_empty[s], 16);  // 8 epilogue warps in each CTA of the pair
    }
    rl::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rl::smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  rl::tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  rl::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;  // tiles_m counts (CL * 128)-row cluster tiles here
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int tiles_mn = p.tiles_m * p.tiles_n;
        const int mn = p.ks_major ? tile % tiles_mn : tile / p.k_splits;
        const int ks = p.ks_major ? tile / tiles_mn : tile - mn * p.k_splits;
        const int m_blk = mn / p.tiles_n;
        const int n_blk = mn - m_blk * p.tiles_n;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int m0 = m_blk * CL * BM + (int)crank * BM;
        const int n0 = n_blk * BN + (int)rank * (BN / 2);
        int img0 = 0, h0 = 0;
        if (p.a_mode == 1) {
          img0 = m0 >> p.hw_shift;
          h0 = (m0 & ((1 << p.hw_shift) - 1)) >> p.w_shift;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          rl::mbar_wait(&empty_bar[stage], phase ^ 1);
          {
          if (leader) rl::mbar_expect_tx(&full_bar[stage], 2 * (A_BYTES + BH_BYTES));
          if (p.a_mn) {
            tma2_load_2d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], m0, kb * BK);
            tma2_load_2d(smem_a + stage * A_BYTES + 8192, &tmA, &full_bar[stage], m0 + 64, kb * BK);
          } else if (p.a_mode == 0) {
            tma2_load_2d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
            const int t = kb / p.cin_blocks;
            const int cb = kb - t * p.cin_blocks;
            tma2_load_5d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], cb * BK, (int)p.tap_dw[t],
                         h0 + (int)p.tap_dh[t], (int)p.tap_plane[t], img0);
          }
          if (p.b_mode == 1) {
            // implicit im2col (see gemm_bf16_kernel): this CTA gathers its half of the pair's N tile
            const int pix0 = kb * BK;
            const int imgk = pix0 >> p.hw_shift;
            const int hk = (pix0 & ((1 << p.hw_shift) - 1)) >> p.w_shift;
#pragma unroll
            for (int b = 0; b < BN / 128; ++b) {
              const int nb = (n0 >> 6) + b;
              const int t = nb / p.cin_blocks;
              const int cb = nb - t * p.cin_blocks;
              if (t < p.ntaps)
                tma2_load_5d(smem_b + stage * BH_BYTES + b * 8192, &tmB, &full_bar[stage], cb * 64, (int)p.tap_dw[t],
                             hk + (int)p.tap_dh[t], (int)p.tap_plane[t], imgk);
              else
                tma2_load_5d(smem_b + stage * BH_BYTES + b * 8192, &tmB, &full_bar[stage], 0, 0, 0, 0, p.nimg);
            }
          } else if (CL == 4) {
            // quarter `pair` of this half of B (64 n-rows / n-columns = 8 KB), delivered to both pairs
            if (p.b_mn)
              tma2_load_2d_mc(smem_b + stage * BH_BYTES + pair * 8192, &tmB, &full_bar[stage], n0 + (int)pair * 64, kb * BK,
                              bcast_mask);
            else
              tma2_load_2d_mc(smem_b + stage * BH_BYTES + pair * 8192, &tmB, &full_bar[stage], kb * BK, n0 + (int)pair * 64,
                              bcast_mask);
          } else if (p.b_mn) {
#pragma unroll
            for (int b = 0; b < BN / 128; ++b)
              tma2_load_2d(smem_b + stage * BH_BYTES + b * 8192, &tmB, &full_bar[stage], n0 + b * 64, kb * BK);
          } else {
            tma2_load_2d(smem_b + stage * BH_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
          }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      // The whole warp runs the loop converged (addresses stay in uniform registers); one elected lane issues.
      const uint32_t idesc = rl::make_idesc_h(2 * BM, BN, p.a_mn, p.b_mn, p.a_f16, p.b_f16);
      const uint32_t a_base = rl::smem_u32(smem_a), b_base = rl::smem_u32(smem_b);
      const uint32_t a_lbo = p.a_mn ? 8192 : 16, b_lbo = p.b_mn ? 8192 : 16;
      const uint32_t a_kstep = p.a_mn ? 128 : 2, b_kstep = p.b_mn ? 128 : 2;  // in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int ks = p.ks_major ? tile / (p.tiles_m * p.tiles_n) : tile % p.k_splits;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        rl::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        rl::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          rl::mbar_wait(&full_bar[stage], phase);
          rl::tc_fence_after();
          if (rl::elect_one()) {
            const uint64_t adesc = rl::make_smem_desc_sw128(a_base + stage * A_BYTES, a_lbo, 1024);
            const uint64_t bdesc = rl::make_smem_desc_sw128(b_base + stage * BH_BYTES, b_lbo, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc2_mma_f16(d_tmem, adesc + a_kstep * k, bdesc + b_kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            tc2_commit_mc(&empty_bar[stage], all_mask);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (rl::elect_one()) tc2_commit_mc(&tmem_full[acc], pair_mask);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps in each CTA; rows of this CTA's half of the pair tile) =====
    const int ew = warp - 4;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* stg = smem_stage + ew * WARP_STG;
    float* sb = reinterpret_cast<float*>(stg + NSTG * 4096);
    int acc = 0;
    EpiState est{0, 0u, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, -1, SPEC ? SPEC : load_eflags(p), -1, 0};
    const uint32_t F = SPEC ? SPEC : est.flags;
    uint64_t* rbar = res_bar + ew * 4;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int mn = p.ks_major ? tile % (p.tiles_m * p.tiles_n) : tile / p.k_splits;
      const int m_blk = mn / p.tiles_n;
      const int n_blk = mn - m_blk * p.tiles_n;
      const int row0 = m_blk * CL * BM + (int)crank * BM + q * 32;
      const int n0 = n_blk * BN;
      float xr[32];
      epilogue_prefetch<BN, NSTG>(p, F, sb, row0, n0, half, lane, xr, est);
      const bool rotate = NSTG == 2 && BN < 256 && (F & EF_DEEP) && !(F & EF_TMA_RES) && p.N % BN == 0 && p.M % (CL * BM) == 0;
      if (NSTG == 2 && (F & EF_DEEP) && !(F & EF_TMA_RES)) deep_tile_begin<BN / 64>(rotate, lane);
      if (NSTG == 2 && (F & EF_TMA_RES) && lane == 0 && row0 < p.M && n0 + half * (BN / 2) < p.N) {
        if (F & EF_RES_ALL) {
          // every residual chunk of this output tile, now: the stores of the previous tile have long drained
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < BN / 64; ++cc)
            if (n0 + half * (BN / 2) + cc * 32 < p.N) issue_residual(F, &tmR, stg, rbar, cc, n0 + half * (BN / 2) + cc * 32, row0);
        } else {
          // first residual tile of this output tile: its staging tile was last read by the store two chunks ago
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          issue_residual(F, &tmR, stg, rbar, est.stg_sel & 1, n0 + half * (BN / 2), row0);
        }
      }
      rl::mbar_wait(&tmem_full[acc], acc_phase);
      rl::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      uint64_t* done = &tmem_empty[acc];
      epilogue_tile<BN, COLS, NSTG, SPEC>(p, &tmC, &tmC2, &tmR, stg, rbar, sb, taddr, row0, n0, half, lane, xr, est,
                                          [done, lane] { if (lane == 0) mbar_arrive_leader(done); });
      deep_tile_end<BN / 64>(rotate, est);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (COLS) colsum_flush<BN / 64>(p, est, half, lane);
    if ((F & EF_TMA_STORE) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  rl::tc_fence_before();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / arriving on its barriers
  if (warp == 2) {
    rl::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int STAGES, bool COLS, int NSTG = 2, uint32_t SPEC = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                  const __grid_constant__ CUtensorMap tmR, const KParams p) {
  gemm2_body<BN, STAGES, 2, COLS, NSTG, SPEC>(tmA, tmB, tmC, tmC2, tmR, p);
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm4_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                  const __grid_constant__ CUtensorMap tmR, const KParams p) {
  gemm2_body<BN, STAGES, 4, false>(tmA, tmB, tmC, tmC2, tmR, p);
}

template <int BN, int STAGES, int NSTG = 2>
constexpr int gemm2_smem_bytes() {
  return STAGES * (A_BYTES + (BN / 2) * BK * 2) + (NSTG == 1 ? 8 * 4096 : 8 * (NSTG * 4096 + 1024) + 1024) + (2 * STAGES + 4) * 8 + 16 + 256;
}

template <int BN, int STAGES, bool COLS = false, int NSTG = 2, uint32_t SPEC = 0>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                 const CUtensorMap& tmR, const KParams& p, cudaStream_t st) {
  constexpr int smem = gemm2_smem_bytes<BN, STAGES, NSTG>();
  static_assert(smem <= 232448, "dynamic shared memory of the pair kernel exceeds 227 KB");
  static std::atomic<bool> configured{false};  // idempotent attribute set: a second thread racing here only repeats it
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<BN, STAGES, COLS, NSTG, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      rl_set_error("rl_gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int tiles = p.tiles_m * p.tiles_n * p.k_splits;
  const int max_clusters = p.sms / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  gemm2_bf16_kernel<BN, STAGES, COLS, NSTG, SPEC><<<2 * clusters, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, tmC2, tmR, p);
  return rl_check_launch("rl_gemm_bf16(cta_group::2)");
}

// co-resident 4-CTA clusters of the quad kernel (a cluster must sit inside one GPC: fewer than SMs / 4 may fit)
template <int BN, int STAGES>
int gemm4_max_clusters() {
  static std::atomic<int> cached{-1};
  int v = cached.load();
  if (v >= 0) return v;
  constexpr int smem = gemm2_smem_bytes<BN, STAGES>();
  cudaFuncSetAttribute(gemm4_bf16_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(rl_num_sms() / 4 * 4), 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 4;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm4_bf16_kernel<BN, STAGES>, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 0;   // no device / query failed: the quad kernel is never selected
  }
  cached.store(n);
  return n;
}

template <int BN, int STAGES>
int launch_gemm4(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                 const CUtensorMap& tmR, const KParams& p, cudaStream_t st) {
  constexpr int smem = gemm2_smem_bytes<BN, STAGES>();
  const int tiles = p.tiles_m * p.tiles_n * p.k_splits;
  const int max_clusters = gemm4_max_clusters<BN, STAGES>();
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  gemm4_bf16_kernel<BN, STAGES><<<4 * clusters, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, tmC2, tmR, p);
  return rl_check_launch("rl_gemm_bf16(4-CTA cluster)");
}

int ilog2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return ((1 << s) == v) ? s : -1;
}

}  // namespace

extern "C" int rl_gemm_bf16(const rl_gemm_desc* d, void* stream) {
  RL_REQUIRE(d != nullptr, RL_EINVAL, "rl_gemm_bf16: null descriptor");
  RL_REQUIRE(d->a && d->b && d->out, RL_EINVAL, "rl_gemm_bf16: null operand pointer");
  RL_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, RL_EINVAL, "rl_gemm_bf16: empty problem %lld x %lld x %lld",
             (long long)d->M, (long long)d->N, (long long)d->K);
  RL_REQUIRE(d->K % 8 == 0, RL_EINVAL, "rl_gemm_bf16: K=%lld must be a multiple of 8 (16-byte rows; the K tail of a"
             " 64-deep block is zero-filled by TMA)", (long long)d->K);
  RL_REQUIRE((d->a_major == 0 || d->a_major == 1) && (d->b_major == 0 || d->b_major == 1), RL_EINVAL,
             "rl_gemm_bf16: a_major / b_major must be 0 (K-major) or 1 (MN-major)");
  RL_REQUIRE(!(d->a_major == 1 && d->a_mode != 0), RL_EINVAL, "rl_gemm_bf16: MN-major A needs a_mode 0");
  RL_REQUIRE(d->M < (1ll << 31) && d->N < (1ll << 31), RL_EINVAL, "rl_gemm_bf16: M/N too large");
  RL_REQUIRE(((uintptr_t)d->a & 15) == 0 && ((uintptr_t)d->b & 15) == 0, RL_EALIGN,
             "rl_gemm_bf16: a/b must be 16-byte aligned");
  RL_REQUIRE(d->ldb % 8 == 0, RL_EALIGN, "rl_gemm_bf16: ldb must be a multiple of 8 elements");
  RL_REQUIRE(((uintptr_t)d->out & 3) == 0, RL_EALIGN, "rl_gemm_bf16: out must be 4-byte aligned");
  if (d->scale) RL_REQUIRE(((uintptr_t)d->scale & 15) == 0, RL_EALIGN, "rl_gemm_bf16: scale alignment");
  if (d->bias) RL_REQUIRE(((uintptr_t)d->bias & 15) == 0, RL_EALIGN, "rl_gemm_bf16: bias alignment");

  KParams p{};
  p.M = (int)d->M;
  p.N = (int)d->N;
  p.num_kb = (int)((d->K + BK - 1) / BK);
  p.a_mn = d->a_major;
  p.b_mn = d->b_major;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.a_mode = d->a_mode;
  RL_REQUIRE((d->a_dtype == RL_DT_BF16 || d->a_dtype == RL_DT_F16) && (d->b_dtype == RL_DT_BF16 || d->b_dtype == RL_DT_F16),
             RL_EINVAL, "rl_gemm_bf16: a_dtype / b_dtype must be RL_DT_BF16 or RL_DT_F16");
  // Probed on B200 (tools/probe_mixed_mma.py): an instruction descriptor whose a_format and b_format differ (bf16 x fp16)
  // raises "illegal instruction" — kind::f16 multiplies two operands of ONE 16-bit format.
  RL_REQUIRE(d->a_dtype == d->b_dtype, RL_EINVAL, "rl_gemm_bf16: A and B must share one 16-bit format (bf16 x fp16 is not "
             "executable by tcgen05.mma kind::f16)");
  p.a_f16 = d->a_dtype == RL_DT_F16;
  p.b_f16 = d->b_dtype == RL_DT_F16;
  p.o_f16 = d->out_dtype == RL_DT_F16;
  p.r_f16 = d->res_dtype == RL_DT_F16;
  p.b_mode = d->b_mode;
  p.ks_major = 0;
  p.nimg = d->conv_NIMG;
  p.ntaps = d->ntaps;
  RL_REQUIRE(d->b_mode == 0 || (d->b_mode == 1 && d->b_major == 1 && d->a_mode == 0), RL_EINVAL,
             "rl_gemm_bf16: b_mode 1 (implicit im2col B) needs b_major = 1 and a_mode = 0");
  p.out = d->out;
  p.ldo = d->ldo;
  p.out_f32 = d->out_dtype == RL_DT_F32;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(d->out2);
  p.ldo2 = d->ldo2;
  p.scale = d->scale;
  p.bias = d->bias;
  p.res = d->res;
  p.ldr = d->ldr;
  p.res_f32 = d->res_dtype == RL_DT_F32;
  p.act = d->act;
  p.out_remap = d->out_remap;
  p.remap_plane = d->remap_plane;
  p.drop = rl::make_drop(d->drop_p, d->drop_seed, d->drop_site, d->drop_counter);
  RL_REQUIRE(d->sm_reserve >= 0 && d->sm_reserve < rl_num_sms() - 1, RL_EINVAL, "rl_gemm_bf16: sm_reserve %d out of range", d->sm_reserve);
  p.sms = (rl_num_sms() - d->sm_reserve) & ~1;   // an even number: CTA pairs
  p.colsum = d->colsum;
  p.colsumsq = d->colsumsq;
  RL_REQUIRE(!(d->colsum || d->colsumsq) || d->split_k == 0, RL_EINVAL, "rl_gemm_bf16: colsum / colsumsq need split_k = 0");

  // Tile / kernel selection.  Cost model per 64-deep k-block of one CTA tile (cycles): the tensor pipe needs
  // 2*bn, the operand bytes need bytes / 42.6 (measured L2->SM ingress per SM, ~6.3 KB/clk chip-wide);
  // a CTA pair (cta_group::2) stages only half of B per CTA.  Total = waves * max(mma, load).
  int bn = 256;
  int pair = 0;
  const int g_force_bn = d->tune_tile_n;          // 0 = cost model; 64 / 128 / 256 force the N tile (tuning, tests)
  const int g_pair_mode = d->tune_no_pair == 1 ? 0 : 1;  // tune_no_pair 1: never use the cta_group::2 kernels
  {
    const int sms = p.sms;
    double best = 1e30;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 1 && (!g_pair_mode || d->M <= BM)) continue;
      for (int cand = 128; cand <= 256; cand += 128) {
        if (g_force_bn == 128 || g_force_bn == 256) {
          if (cand != g_force_bn) continue;
        } else if (cand == 256 && p.N <= 128) {
          continue;
        }
        const long long tn = (p.N + cand - 1) / cand;
        const long long tm = mode ? 2 * ((d->M + 2 * BM - 1) / (2 * BM)) : (d->M + BM - 1) / BM;  // CTA tiles along M
        const long long waves = (tm * tn + sms - 1) / sms;
        const double bytes = A_BYTES + (mode ? cand / 2 : cand) * BK * 2.0;
        const double per_kb = bytes / 42.6 > 2.0 * cand ? bytes / 42.6 : 2.0 * cand;
        double cost = waves * (per_kb * p.num_kb + 1500.0);  // + per-tile epilogue / pipeline bubble
        // split-K fills the machine whatever the tile count (the split chooser below sizes the waves): what matters
        // is the total operand traffic, tiles x bytes per k-block — the larger tile moves fewer bytes per FLOP
        if (d->split_k != 0) cost = (double)(tm * tn) * per_kb;
        if (cost < best) {
          best = cost;
          bn = cand;
          pair = mode;
        }
      }
    }
    if (p.N <= 64 && g_force_bn == 0) {
      bn = 64;
      pair = 0;
    }
    if (d->b_mode == 1 && bn == 256 && p.N % 256 != 0 && p.N % 128 == 0) bn = 128;   // no padded N tile of gathers
    if (g_force_bn == 64) {
      bn = 64;
      pair = 0;
    }
  }
  // 4-CTA clusters (two pairs sharing the B tile through TMA multicast): same cost model, 24 KB instead of 32 KB per
  // CTA and k-block, but only as many clusters as fit the GPCs and 512-row tiles
  int quad = 0;
  if (pair && bn == 256 && d->b_mode == 0 && d->tune_no_pair != 2 && p.M >= 4 * BM) {
    const int maxc4 = gemm4_max_clusters<256, 4>();
    if (maxc4 > 0) {
      const long long tn = (p.N + 255) / 256;
      const long long t2 = ((p.M + 2 * BM - 1) / (2 * BM)) * tn, t4 = ((p.M + 4 * BM - 1) / (4 * BM)) * tn;
      const long long u2 = p.sms / 2, u4 = maxc4;
      const double kb2 = (A_BYTES + 128 * BK * 2.0) / 42.6, kb4_l2 = (A_BYTES + 64 * BK * 2.0) / 42.6;
      const double kb4 = kb4_l2 > 512.0 ? kb4_l2 : 512.0;
      double c2, c4;
      if (d->split_k != 0) {   // split-K fills whole waves: compare machine-wide throughput
        c2 = (double)t2 * kb2 / (double)u2;
        c4 = (double)t4 * kb4 / (double)u4;
      } else {
        c2 = (double)((t2 + u2 - 1) / u2) * (kb2 * p.num_kb + 1500.0);
        c4 = (double)((t4 + u4 - 1) / u4) * (kb4 * p.num_kb + 1500.0);
      }
      (void)c2; (void)c4;
      // Measured (tools/gemm_bench.py, profiles/): the 4-CTA kernel is correct but NOT faster — multicast cuts L2 reads, the
      // bytes entering each SM stay the same — so the cost model never picks it; tune_no_pair = 3 forces it (tests, bench).
      quad = d->tune_no_pair == 3 ? 1 : 0;
    }
  }
  if (pair) p.tiles_m = quad ? (p.M + 4 * BM - 1) / (4 * BM) : (p.M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (p.N + bn - 1) / bn;
  // split-K (weight gradients: few output tiles, K = tokens / pixels): ~2 CTAs per SM worth of tiles, >= 8 k-blocks each
  p.k_splits = 1;
  p.kb_per_split = p.num_kb;
  p.atomic_out = d->split_k != 0;
  if (d->split_k != 0) {
    RL_REQUIRE(d->out_dtype == RL_DT_F32 && !d->res && !d->bias && !d->scale && d->act == RL_ACT_NONE && !d->out2 &&
                   d->out_remap == 0 && d->drop_p == 0.f,
               RL_EINVAL, "rl_gemm_bf16: split_k needs a plain f32 accumulate-into output (no epilogue operands)");
    // Work items = output tiles x splits, executed in waves over the SMs (CTA pairs in pair mode).  Pick the split
    // count that minimises waves x (k-blocks per split + the tile's fixed cost: pipeline fill and the atomic
    // epilogue, ~10 k-blocks' worth), e.g. 27 tiles on 74 pairs: 8 splits (3 full waves of 32 k-blocks) beat 6
    // (3 ragged waves of 43).
    const long long tiles = (long long)p.tiles_m * p.tiles_n;
    const long long units = quad ? gemm4_max_clusters<256, 4>() : pair ? p.sms / 2 : p.sms;
    int maxs = p.num_kb / 8;
    if (maxs < 1) maxs = 1;
    int want = 1;
    if (d->split_k > 0) {
      want = d->split_k;
    } else {
      double best_cost = 1e30;
      for (int sp = 1; sp <= maxs && sp <= 512; ++sp) {
        const long long per = (p.num_kb + sp - 1) / sp;
        const long long real = (p.num_kb + per - 1) / per;          // splits that actually exist
        const long long waves = (tiles * real + units - 1) / units;
        const double cost = (double)waves * ((double)per + 10.0);
        if (cost < best_cost - 1e-9) {
          best_cost = cost;
          want = sp;
        }
      }
    }
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    p.kb_per_split = (p.num_kb + want - 1) / want;
    p.k_splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    // split index outermost: the output tiles of one K range run side by side and share its operand panels in L2
    // (tile-major order makes every output tile re-read its full A / B panels from HBM)
    if (p.k_splits > 1) p.ks_major = 1;
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (d->a_mode == 0 && d->a_major == 1) {
    RL_REQUIRE(d->lda % 8 == 0, RL_EALIGN, "rl_gemm_bf16: lda must be a multiple of 8 elements");
    uint64_t dims[2] = {(uint64_t)d->M, (uint64_t)d->K};  // stored [K, M], M contiguous
    uint64_t strides[1] = {(uint64_t)d->lda * 2};
    uint32_t box[2] = {64, BK};
    rc = rl_make_tmap_bf16(&tmA, d->a, 2, dims, strides, box);
    if (rc) return rc;
  } else if (d->a_mode == 0) {
    RL_REQUIRE(d->lda % 8 == 0, RL_EALIGN, "rl_gemm_bf16: lda must be a multiple of 8 elements");
    uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
    uint64_t strides[1] = {(uint64_t)d->lda * 2};
    uint32_t box[2] = {BK, BM};
    rc = rl_make_tmap_bf16(&tmA, d->a, 2, dims, strides, box);
    if (rc) return rc;
  } else if (d->a_mode == 1) {
    const int C = d->conv_C, W = d->conv_W, H = d->conv_H, P = d->conv_P, NI = d->conv_NIMG;
    const int Cuse = d->conv_Cuse > 0 ? d->conv_Cuse : C;  // channels [0, Cuse) of each tap feed the GEMM
    RL_REQUIRE(C > 0 && C % 8 == 0 && Cuse % BK == 0 && Cuse <= C, RL_EINVAL,
               "rl_gemm_bf16(conv): C=%d / used channels %d (must be a multiple of 64)", C, Cuse);
    RL_REQUIRE(d->ntaps >= 1 && d->ntaps <= 12, RL_EINVAL, "rl_gemm_bf16(conv): ntaps=%d", d->ntaps);
    RL_REQUIRE(d->K == (int64_t)d->ntaps * Cuse, RL_EINVAL, "rl_gemm_bf16(conv): K != ntaps*C");
    const int ws = ilog2_exact(W), hs = ilog2_exact(H);
    RL_REQUIRE(ws >= 0 && hs >= 0 && W * H <= 256, RL_EINVAL,
               "rl_gemm_bf16(conv): map %dx%d must be power-of-two with <= 256 pixels", W, H);
    RL_REQUIRE(d->M == (int64_t)NI * W * H, RL_EINVAL, "rl_gemm_bf16(conv): M != NIMG*H*W");
    RL_REQUIRE(P == 1 || P == 4, RL_EINVAL, "rl_gemm_bf16(conv): P must be 1 or 4");
    p.hw_shift = ws + hs;
    p.w_shift = ws;
    p.cin_blocks = Cuse / BK;
    for (int t = 0; t < d->ntaps; ++t) {
      p.tap_dw[t] = d->tap_dw[t];
      p.tap_dh[t] = d->tap_dh[t];
      p.tap_plane[t] = d->tap_plane[t];
      RL_REQUIRE(d->tap_plane[t] >= 0 && d->tap_plane[t] < P, RL_EINVAL, "rl_gemm_bf16(conv): tap plane");
    }
    const int hw = W * H;
    uint32_t box[5];
    box[0] = BK;
    box[1] = (uint32_t)W;
    box[2] = (uint32_t)(hw >= BM ? BM / W : H);
    box[3] = 1;
    box[4] = (uint32_t)(hw >= BM ? 1 : BM / hw);
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)P, (uint64_t)NI};
    uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)C * W * 2, (uint64_t)C * W * H * 2,
                           (uint64_t)C * W * H * P * 2};
    rc = rl_make_tmap_bf16(&tmA, d->a, 5, dims, strides, box);
    if (rc) return rc;
  } else {
    RL_REQUIRE(false, RL_EINVAL, "rl_gemm_bf16: unknown a_mode %d", d->a_mode);
  }
  if (d->out_remap == 1) {
    RL_REQUIRE(d->a_mode == 1, RL_EINVAL, "rl_gemm_bf16: out_remap=1 needs conv geometry");
    RL_REQUIRE(d->conv_W >= 2 && d->conv_H >= 2, RL_EINVAL, "rl_gemm_bf16: parity split needs >=2x2 map");
  }
  if (d->out_remap == 2)
    RL_REQUIRE(d->a_mode == 1 && d->remap_plane >= 0 && d->remap_plane < 4, RL_EINVAL,
               "rl_gemm_bf16: out_remap=2 needs conv geometry and a plane in 0..3");
  if (d->b_mode == 1) {
    const int C = d->conv_C, W = d->conv_W, H = d->conv_H, P = d->conv_P, NI = d->conv_NIMG;
    const int Cuse = d->conv_Cuse > 0 ? d->conv_Cuse : C;
    RL_REQUIRE(C > 0 && C % 8 == 0 && Cuse % 64 == 0 && Cuse <= C, RL_EINVAL,
               "rl_gemm_bf16(im2col B): C=%d / used channels %d (must be a multiple of 64)", C, Cuse);
    RL_REQUIRE(d->ntaps >= 1 && d->ntaps <= 12, RL_EINVAL, "rl_gemm_bf16(im2col B): ntaps=%d", d->ntaps);
    RL_REQUIRE(d->N == (int64_t)d->ntaps * Cuse, RL_EINVAL, "rl_gemm_bf16(im2col B): N != ntaps*C");
    const int ws = ilog2_exact(W), hs = ilog2_exact(H);
    RL_REQUIRE(ws >= 0 && hs >= 0 && W <= 64 && W * H <= 256, RL_EINVAL,
               "rl_gemm_bf16(im2col B): map %dx%d must be power-of-two with <= 256 pixels", W, H);
    RL_REQUIRE(d->K == (int64_t)NI * W * H, RL_EINVAL, "rl_gemm_bf16(im2col B): K != NIMG*H*W");
    RL_REQUIRE(P == 1 || P == 4, RL_EINVAL, "rl_gemm_bf16(im2col B): P must be 1 or 4");
    p.hw_shift = ws + hs;
    p.w_shift = ws;
    p.cin_blocks = Cuse / 64;
    p.ks_major = 1;
    for (int t = 0; t < d->ntaps; ++t) {
      p.tap_dw[t] = d->tap_dw[t];
      p.tap_dh[t] = d->tap_dh[t];
      p.tap_plane[t] = d->tap_plane[t];
      RL_REQUIRE(d->tap_plane[t] >= 0 && d->tap_plane[t] < P, RL_EINVAL, "rl_gemm_bf16(im2col B): tap plane");
    }
    const int hw = W * H;
    uint32_t box[5];
    box[0] = 64;
    box[1] = (uint32_t)W;
    box[2] = (uint32_t)(hw >= BK ? BK / W : H);
    box[3] = 1;
    box[4] = (uint32_t)(hw >= BK ? 1 : BK / hw);
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)P, (uint64_t)NI};
    uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)C * W * 2, (uint64_t)C * W * H * 2, (uint64_t)C * W * H * P * 2};
    rc = rl_make_tmap_bf16(&tmB, d->b, 5, dims, strides, box);
    if (rc) return rc;
  } else if (d->b_major == 1) {
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->K};  // stored [K, N], N contiguous
    uint64_t strides[1] = {(uint64_t)d->ldb * 2};
    uint32_t box[2] = {64, BK};
    rc = rl_make_tmap_bf16(&tmB, d->b, 2, dims, strides, box);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)d->ldb * 2};
    uint32_t box[2] = {BK, (uint32_t)(quad ? bn / 4 : pair ? bn / 2 : bn)};
    rc = rl_make_tmap_bf16(&tmB, d->b, 2, dims, strides, box);
    if (rc) return rc;
  }
  // output path: TMA store for plain row-major outputs with 16-byte aligned rows
  const int oelt = p.out_f32 ? 4 : 2;
  const bool aligned16 = ((uintptr_t)d->out & 15) == 0 && (d->ldo * oelt) % 16 == 0;
  const bool out2_tma = d->out2 != nullptr && d->act == RL_ACT_GELU_SAVE && !p.out_f32 && ((uintptr_t)d->out2 & 15) == 0 &&
                        (d->ldo2 * 2) % 16 == 0;
  p.tma_store = (d->out_remap == 0 && (d->out2 == nullptr || out2_tma) && aligned16) ? 1 : 0;  // else direct stores / scalar atomics
  p.tma_out2 = (p.tma_store && out2_tma) ? 1 : 0;
  // (split-K with a TMA-able output: the per-chunk store becomes a cp.reduce.async.bulk .add — no per-element atomics)
  p.vec_store = (aligned16 && (d->out2 == nullptr || (((uintptr_t)d->out2 & 15) == 0 && d->ldo2 % 8 == 0))) ? 1 : 0;
  if (d->res) {
    const int relt = p.res_f32 ? 4 : 2;
    RL_REQUIRE(((uintptr_t)d->res & 15) == 0 && (d->ldr * relt) % 16 == 0, RL_EALIGN,
               "rl_gemm_bf16: res / ldr must be 16-byte aligned");
  }
  CUtensorMap tmC = tmB;  // placeholder when the direct-store path is used
  if (p.tma_store) {
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
    uint64_t strides[1] = {(uint64_t)d->ldo * oelt};
    uint32_t box[2] = {32, 32};
    rc = rl_make_tmap(&tmC, d->out, p.out_f32 ? RL_TMAP_F32 : RL_TMAP_BF16, p.out_f32 ? 128 : 64, 2, dims, strides, box);
    if (rc) return rc;
  }
  CUtensorMap tmC2 = tmB, tmR = tmB;
  if (p.tma_out2) {
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
    uint64_t strides[1] = {(uint64_t)d->ldo2 * 2};
    uint32_t box[2] = {32, 32};
    rc = rl_make_tmap(&tmC2, d->out2, RL_TMAP_BF16, 64, 2, dims, strides, box);
    if (rc) return rc;
  }
  // residual tiles by TMA: plain row-major result leaving by TMA, residual rows 16-byte aligned (checked above), and the
  // residual tile must fit the staging tile the result leaves from (f32 residual -> f32 result)
  p.tma_res = (d->res && p.tma_store && !p.atomic_out && !p.tma_out2 && (p.out_f32 || !p.res_f32)) ? 1 : 0;
  const bool long_k = pair && bn == 256 && !(d->colsum || d->colsumsq) && !p.tma_out2 && p.kb_per_split >= 24 && d->tune_no_pair != 4;
  p.deep = (p.tma_store && !p.out_f32 && !p.tma_out2 && !p.atomic_out && !long_k) ? 1 : 0;
  p.res_all = (p.deep && p.tma_res) ? 1 : 0;
  if (p.tma_res) {
    uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
    uint64_t strides[1] = {(uint64_t)d->ldr * (p.res_f32 ? 4 : 2)};
    uint32_t box[2] = {32, 32};
    rc = rl_make_tmap(&tmR, d->res, p.res_f32 ? RL_TMAP_F32 : RL_TMAP_BF16, p.res_f32 ? 128 : 64, 2, dims, strides, box);
    if (rc) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool cols = d->colsum || d->colsumsq;   // separate instantiations: the reductions cost registers in the epilogue
  p.eflags = EF_VALID | (p.tma_res ? EF_TMA_RES : 0u) | (p.res && p.res_f32 ? EF_RES_F32 : 0u) | (p.res && p.r_f16 ? EF_R_F16 : 0u) |
             (p.drop.thresh ? EF_DROP : 0u) | (p.tma_store ? EF_TMA_STORE : 0u) | (p.deep ? EF_DEEP : 0u) |
             (p.res_all ? EF_RES_ALL : 0u) | (p.tma_out2 ? EF_TMA_OUT2 : 0u) | (p.out_f32 ? EF_OUT_F32 : 0u) |
             (p.o_f16 ? EF_O_F16 : 0u) | (p.atomic_out ? EF_ATOMIC : 0u) | (p.scale ? EF_SCALE : 0u) | (p.res ? EF_RES : 0u) |
             (p.out2 ? EF_OUT2 : 0u) | (p.vec_store ? EF_VEC : 0u) | (p.colsum ? EF_COLSUM : 0u) |
             (p.colsumsq ? EF_COLSUMSQ : 0u) | ((uint32_t)p.act << EF_ACT_SHIFT);
  // res_block1's 3x3 / 64-channel / 16x16 convolution (forward conv2 and its data gradient): resident weights + row-halo A
  if (d->a_mode == 1 && d->ntaps == 9 && d->conv_C == 64 && (d->conv_Cuse == 0 || d->conv_Cuse == 64) && d->conv_W == 16 &&
      d->conv_H == 16 && d->conv_P == 1 && p.N == 64 && d->K == 576 && bn == 64 && !pair && !cols && d->split_k == 0 &&
      d->b_major == 0 && d->b_mode == 0 && p.M % BM == 0 && p.eflags == SPEC_OUT16 && d->tune_no_pair != 5) {
    bool ok = true;
    for (int i = 0; i < 9; ++i) p.halo_t[i] = -1;
    for (int t = 0; t < 9 && ok; ++t) {
      const int dw = d->tap_dw[t], dh = d->tap_dh[t];
      ok = dw >= -1 && dw <= 1 && dh >= -1 && dh <= 1 && d->tap_plane[t] == 0 && p.halo_t[(dw + 1) * 3 + dh + 1] < 0;
      if (ok) p.halo_t[(dw + 1) * 3 + dh + 1] = (int8_t)t;
    }
    if (ok) {
      CUtensorMap tmH;
      uint32_t box[5] = {BK, 16, 10, 1, 1};
      uint64_t dims[5] = {64, 16, 16, 1, (uint64_t)d->conv_NIMG};
      uint64_t strides[4] = {64 * 2, 64 * 16 * 2, 64 * 16 * 16 * 2, 64 * 16 * 16 * 2};
      rc = rl_make_tmap_bf16(&tmH, d->a, 5, dims, strides, box);
      if (rc) return rc;
      return launch_conv64_halo(tmH, tmB, tmC, p, st);
    }
  }
  if (quad && !cols) return launch_gemm4<256, 4>(tmA, tmB, tmC, tmC2, tmR, p, st);
  if (pair && bn == 256 && !cols && !long_k && d->tune_no_pair != 5) {
    // the hot short-K configurations of a train step, with the epilogue's mode word folded at compile time
    if (p.eflags == SPEC_OUT16) return launch_gemm2<256, 4, false, 2, SPEC_OUT16>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_GELU_SAVE) return launch_gemm2<256, 4, false, 2, SPEC_GELU_SAVE>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_RES32) return launch_gemm2<256, 4, false, 2, SPEC_RES32>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_GELU_GRAD) return launch_gemm2<256, 4, false, 2, SPEC_GELU_GRAD>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_RES32_ND) return launch_gemm2<256, 4, false, 2, SPEC_RES32_ND>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_OUT16_F16) return launch_gemm2<256, 4, false, 2, SPEC_OUT16_F16>(tmA, tmB, tmC, tmC2, tmR, p, st);
    if (p.eflags == SPEC_GELU16_F16) return launch_gemm2<256, 4, false, 2, SPEC_GELU16_F16>(tmA, tmB, tmC, tmC2, tmR, p, st);
  }
  if (pair && bn == 256 && cols && d->tune_no_pair != 5 && p.eflags == SPEC_GELU_GRAD_CS)
    return launch_gemm2<256, 4, true, 2, SPEC_GELU_GRAD_CS>(tmA, tmB, tmC, tmC2, tmR, p, st);
  if (pair) {
    // long-K variant (5 stages; one staging tile per epilogue warp, no scale/bias table): >= 24 k-blocks per work item.
    // Measured (tools/gemm_bench.py): 4 -> 5 stages = -5..10 % on K >= 2304 GEMMs and split-K weight gradients, 5 -> 6 nothing
    // more: ~750 clk per k-block is what 32 KB per k-block costs at the ~81 GB/s an SM can ingest (68 % tensor-pipe ceiling
    // of 256x256 pair tiles)
    if (bn == 256 && !cols && !p.tma_out2 && p.kb_per_split >= 24 && d->tune_no_pair != 4) {
      if (d->tune_no_pair != 5) {
        if (p.eflags == SPEC_RES32) return launch_gemm2<256, 5, false, 1, SPEC_RES32>(tmA, tmB, tmC, tmC2, tmR, p, st);
        if (p.eflags == SPEC_RES32_ND) return launch_gemm2<256, 5, false, 1, SPEC_RES32_ND>(tmA, tmB, tmC, tmC2, tmR, p, st);
        if (p.eflags == SPEC_SPLITK) return launch_gemm2<256, 5, false, 1, SPEC_SPLITK>(tmA, tmB, tmC, tmC2, tmR, p, st);
      }
      return launch_gemm2<256, 5, false, 1>(tmA, tmB, tmC, tmC2, tmR, p, st);
    }
    if (bn == 256) return cols ? launch_gemm2<256, 4, true>(tmA, tmB, tmC, tmC2, tmR, p, st)
                               : launch_gemm2<256, 4>(tmA, tmB, tmC, tmC2, tmR, p, st);
    return cols ? launch_gemm2<128, 6, true>(tmA, tmB, tmC, tmC2, tmR, p, st) : launch_gemm2<128, 6>(tmA, tmB, tmC, tmC2, tmR, p, st);
  }
  if (bn == 256) return cols ? launch_gemm<256, 3, true>(tmA, tmB, tmC, tmC2, tmR, p, st) : launch_gemm<256, 3>(tmA, tmB, tmC, tmC2, tmR, p, st);
  if (bn == 64 && !cols && p.eflags == SPEC_OUT16 && d->tune_no_pair != 5)   // the glyph stem convs: one chunk per warp and tile, epilogue-bound
    return launch_gemm<64, 6, false, SPEC_OUT16>(tmA, tmB, tmC, tmC2, tmR, p, st);
  if (bn == 64) return cols ? launch_gemm<64, 6, true>(tmA, tmB, tmC, tmC2, tmR, p, st) : launch_gemm<64, 6>(tmA, tmB, tmC, tmC2, tmR, p, st);
  return cols ? launch_gemm<128, 4, true>(tmA, tmB, tmC, tmC2, tmR, p, st) : launch_gemm<128, 4>(tmA, tmB, tmC, tmC2, tmR, p, st);
}
