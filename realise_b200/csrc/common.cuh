// Shared device-side PTX wrappers (mbarrier / TMA / tcgen05 / TMEM) and host-side
// error plumbing for librealise_b200.so.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/realise_b200.h"

// ----------------------------------------------------------------------------------
// host-side error handling (thread-local message, C ABI returns an int code)
// ----------------------------------------------------------------------------------
void rl_set_error(const char* fmt, ...);
int rl_check_launch(const char* what);  // returns 0 or positive cudaError_t
int rl_num_sms();
// resident CTAs per SM of a kernel (cudaOccupancyMaxActiveBlocksPerMultiprocessor, cached per function; 2 when the query
// fails, e.g. without a device).  Slab kernels size their grid to ONE full wave: a few CTAs beyond it cost a whole extra pass.
int rl_ctas_per_sm(const void* func, int threads, int dyn_smem);

#define RL_REQUIRE(cond, code, ...)    \
  do {                                 \
    if (!(cond)) {                     \
      rl_set_error(__VA_ARGS__);       \
      return (code);                   \
    }                                  \
  } while (0)

// driver entry point for cuTensorMapEncodeTiled, resolved lazily through the runtime so
// the library has no link-time dependency on libcuda.so (it must dlopen on a CPU-only box)
typedef CUresult (*rl_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
rl_tmap_encode_fn rl_get_tmap_encode();

// Build a tiled tensor map, dims fastest-first.  swizzle_bytes in {0, 32, 64, 128}.  Returns 0 / error code.
enum { RL_TMAP_BF16 = 0, RL_TMAP_F32 = 1 };
int rl_make_tmap(CUtensorMap* out, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box);
inline int rl_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
  return rl_make_tmap(out, base, RL_TMAP_BF16, 128, rank, dims, strides_bytes, box);
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------
namespace rl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA (cp.async.bulk.tensor, tile mode, completes on an mbarrier) ----
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// commit all prior tcgen05.mma of this thread to an mbarrier (arrive::one)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16/fp16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (e.g. softmax probabilities), B from smem
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// instruction descriptor: bf16 x bf16 -> f32, M x N tile, majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0,
                                                       int b_mn_major = 0, int f16 = 0) {
  return (1u << 4)                        // c_format = F32
         | ((f16 ? 0u : 1u) << 7)         // a_format: 0 = F16, 1 = BF16
         | ((f16 ? 0u : 1u) << 10)        // b_format
         | ((uint32_t)a_mn_major << 15)   // a_major
         | ((uint32_t)b_mn_major << 16)   // b_major
         | ((uint32_t)(N >> 3) << 17)     // n_dim
         | ((uint32_t)(M >> 4) << 24);    // m_dim
}

// the same with the 16-bit format chosen per operand (tcgen05 kind::f16 multiplies bf16 by fp16 operands as they are:
// training keeps forward tensors in fp16 and gradients in bf16)
__host__ __device__ constexpr uint32_t make_idesc_h(int M, int N, int a_mn_major, int b_mn_major, int a_f16, int b_f16) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// shared-memory matrix descriptor, SWIZZLE_128B.  K-major: rows of 128 B (64 bf16), 8-row atoms,
// SBO = byte stride between consecutive 8-row groups.  MN-major: 128-B rows along MN, 8 k-rows per
// atom; LBO = byte stride between 64-element MN blocks, SBO = byte stride between 8-k groups.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// TMEM -> registers, 32 lanes x 32 columns of 32-bit (one row per thread, 32 consecutive columns)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// registers -> TMEM, 32 lanes x 32 columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
      "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
// 16-bit storage format of a tensor: bf16 or IEEE fp16 (RL_DT_F16: three more mantissa bits for the same tensor-core
// rate; forward activations / weights of this model stay far inside the fp16 range, gradients stay bf16)
__device__ __forceinline__ uint32_t pack_h(float a, float b, int f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16(a, b);
}
__device__ __forceinline__ float half_lo(uint32_t v, int f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))) : bf16_lo(v);
}
__device__ __forceinline__ float half_hi(uint32_t v, int f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(v >> 16))) : bf16_hi(v);
}

// 2^x, one MUFU op (ex2.approx: <= 2 ulp; -inf -> 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Counter-based dropout mask: keep(seed, site, idx) is a pure function, so the backward kernels regenerate
// exactly the mask the forward used.  One splitmix64 finalizer over (seed, site, idx / 4) yields four 16-bit
// lanes, one per element of an aligned group of 4; an element is dropped when its lane falls below
// thresh = round(p * 2^16) (p = 0.1 -> 0.100006).  Kernels that walk aligned groups hash once per 4 elements.
__host__ __device__ __forceinline__ unsigned long long drop_hash4(unsigned long long seed, unsigned int site,
                                                                  unsigned long long idx4) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)site + 1ull) + idx4 * 0xD1342543DE82EF95ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned int site, unsigned long long idx,
                                                   unsigned int thresh) {
  const unsigned long long h = drop_hash4(seed, site, idx >> 2);
  return ((unsigned int)(h >> ((unsigned int)(idx & 3ull) * 16u)) & 0xFFFFu) >= thresh;
}
struct DropSpec {   // passed by value to kernels; thresh == 0 means "no dropout"
  unsigned long long seed;
  const unsigned long long* seed_ptr;   // optional device-resident step counter added to `seed` at run time, so that a
                                        // captured CUDA graph draws fresh masks on every replay (the drop_counter argument)
  unsigned int site;
  unsigned int thresh;
  float scale;      // 1 / (1 - p)
};
__host__ __device__ __forceinline__ DropSpec make_drop(float p, unsigned long long seed, unsigned int site,
                                                       const uint64_t* counter = nullptr) {
  DropSpec d;
  d.seed = seed;
  d.seed_ptr = reinterpret_cast<const unsigned long long*>(counter);
  d.site = site;
  d.thresh = p > 0.f ? (unsigned int)(p * 65536.0 + 0.5) : 0u;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}
// Effective seed of a launch: kernels call this ONCE (per thread) and keep the result in `seed`.
__device__ __forceinline__ void drop_resolve(DropSpec& d) {
  if (d.thresh != 0u && d.seed_ptr != nullptr) d.seed += __ldg(d.seed_ptr);
  d.seed_ptr = nullptr;
}
__device__ __forceinline__ float drop_apply(const DropSpec& d, unsigned long long idx, float x) {
  if (d.thresh == 0u) return x;
  return drop_keep(d.seed, d.site, idx, d.thresh) ? x * d.scale : 0.0f;
}

// 4 consecutive elements starting at e0 (e0 % 4 == 0): one hash
__device__ __forceinline__ void drop_apply4(const DropSpec& d, unsigned long long e0, float4& v) {
  if (d.thresh == 0u) return;
  const unsigned long long h = drop_hash4(d.seed, d.site, e0 >> 2);
  const unsigned int lo = (unsigned int)h, hi = (unsigned int)(h >> 32);
  v.x = (lo & 0xFFFFu) >= d.thresh ? v.x * d.scale : 0.f;
  v.y = (lo >> 16) >= d.thresh ? v.y * d.scale : 0.f;
  v.z = (hi & 0xFFFFu) >= d.thresh ? v.z * d.scale : 0.f;
  v.w = (hi >> 16) >= d.thresh ? v.w * d.scale : 0.f;
}
// 32 consecutive elements starting at e0: 8 hashes when e0 % 4 == 0, per-element otherwise
__device__ __forceinline__ void drop_apply32(const DropSpec& d, unsigned long long e0, float (&x)[32]) {
  if (d.thresh == 0u) return;
  if ((e0 & 3ull) == 0ull) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const unsigned long long h = drop_hash4(d.seed, d.site, (e0 >> 2) + g);
      const unsigned int lo = (unsigned int)h, hi = (unsigned int)(h >> 32);
      x[4 * g] = (lo & 0xFFFFu) >= d.thresh ? x[4 * g] * d.scale : 0.f;
      x[4 * g + 1] = (lo >> 16) >= d.thresh ? x[4 * g + 1] * d.scale : 0.f;
      x[4 * g + 2] = (hi & 0xFFFFu) >= d.thresh ? x[4 * g + 2] * d.scale : 0.f;
      x[4 * g + 3] = (hi >> 16) >= d.thresh ? x[4 * g + 3] * d.scale : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = drop_keep(d.seed, d.site, e0 + j, d.thresh) ? x[j] * d.scale : 0.f;
  }
}
// keep bits (bit j = element e0 + j kept) of 32 consecutive elements
__device__ __forceinline__ unsigned int drop_bits32(const DropSpec& d, unsigned long long e0) {
  unsigned int bits = 0u;
  if ((e0 & 3ull) == 0ull) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const unsigned long long h = drop_hash4(d.seed, d.site, (e0 >> 2) + g);
      const unsigned int lo = (unsigned int)h, hi = (unsigned int)(h >> 32);
      bits |= ((lo & 0xFFFFu) >= d.thresh ? 1u : 0u) << (4 * g);
      bits |= ((lo >> 16) >= d.thresh ? 1u : 0u) << (4 * g + 1);
      bits |= ((hi & 0xFFFFu) >= d.thresh ? 1u : 0u) << (4 * g + 2);
      bits |= ((hi >> 16) >= d.thresh ? 1u : 0u) << (4 * g + 3);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) bits |= (drop_keep(d.seed, d.site, e0 + j, d.thresh) ? 1u : 0u) << j;
  }
  return bits;
}

// per-column sum over the 32 rows held by the warp (lane = row, v[c] = column c): butterfly transpose-reduce, 31 shuffles;
// on return lane l holds the total of column l in v[0]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w];
      const float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the result)
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float y = 1.0f - poly * t * __expf(-ax * ax);
  return copysignf(y, x);
}

// d/dx [x * Phi(x)] = Phi(x) + x * phi(x).  erf(u/sqrt2) and phi(u) share one exponential: exp(-(u/sqrt2)^2) = exp(-u^2/2).
__device__ __forceinline__ float gelu_grad(float u) {
  const float x = u * 0.70710678118654752440f;
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = __expf(-ax * ax);
  const float erfv = copysignf(1.0f - poly * t * e, x);
  return fmaf(u, 0.3989422804014327f * e, fmaf(0.5f, erfv, 0.5f));
}

__device__ __forceinline__ float gelu_erf(float x) { return x * 0.5f * (1.0f + fast_erf(x * 0.70710678118654752440f)); }

// ---- packed f32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of fp32 work) ----------
// The element-wise GELU passes are bound by instruction issue, not by the fp32 pipe or HBM: evaluating the erf
// polynomial on two elements per instruction halves their issue count (22 -> 12.5 SASS instructions per element).
struct f2 {
  float x, y;
};
__device__ __forceinline__ unsigned long long f2_pack(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long r) {
  f2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 bc2(float c) { return f2{c, c}; }
__device__ __forceinline__ float rcp_approx(float x) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// shared part of gelu / gelu': y = 1 - erf(|u|/sqrt2) (Abramowitz-Stegun 7.1.26) and e = exp(-u^2/2), two elements at once
__device__ __forceinline__ void erfc_exp2(f2 u, f2& ax, f2& y, f2& e) {
  const f2 x = mul2(u, bc2(0.70710678118654752440f));
  ax = f2{fabsf(x.x), fabsf(x.y)};
  const f2 den = fma2(bc2(0.3275911f), ax, bc2(1.0f));
  const f2 t = f2{rcp_approx(den.x), rcp_approx(den.y)};
  f2 poly = fma2(bc2(1.061405429f), t, bc2(-1.453152027f));
  poly = fma2(poly, t, bc2(1.421413741f));
  poly = fma2(poly, t, bc2(-0.284496736f));
  poly = fma2(poly, t, bc2(0.254829592f));
  const f2 a2 = mul2(mul2(ax, bc2(-1.4426950408889634f)), ax);   // -x^2 * log2(e)
  e = f2{ex2(a2.x), ex2(a2.y)};
  y = mul2(mul2(poly, t), e);
}
// u * Phi(u) = u/2 + |u|/2 * erf(|u|/sqrt2)
__device__ __forceinline__ f2 gelu_erf2(f2 u) {
  f2 ax, y, e;
  erfc_exp2(u, ax, y, e);
  const f2 ahx = mul2(ax, bc2(0.70710678118654752440f));   // |u| / 2
  return fma2(f2{-ahx.x, -ahx.y}, y, add2(mul2(u, bc2(0.5f)), ahx));
}
// Phi(u) + u * phi(u), Phi(u) = 1 - y/2 (u >= 0) or y/2 (u < 0)
__device__ __forceinline__ f2 gelu_grad2(f2 u) {
  f2 ax, y, e;
  erfc_exp2(u, ax, y, e);
  const f2 hy = mul2(y, bc2(0.5f));
  const f2 up = fma2(bc2(-1.0f), hy, bc2(1.0f));
  const f2 cdf = f2{u.x >= 0.f ? up.x : hy.x, u.y >= 0.f ? up.y : hy.y};
  return fma2(u, mul2(e, bc2(0.3989422804014327f)), cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace rl
#endif  // __CUDACC__
