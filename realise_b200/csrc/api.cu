// Library-level plumbing of librealise_b200.so: error strings, device attributes, TMA descriptors.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void rl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int rl_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    rl_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int rl_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      sms = 148;
  }
  return sms;
}

int rl_ctas_per_sm(const void* func, int threads, int dyn_smem) {
  static std::mutex mu;
  static const void* keys[64];
  static int vals[64];
  static int n = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n; ++i)
    if (keys[i] == func) return vals[i];
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, (size_t)dyn_smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 2;
  }
  if (n < 64) {
    keys[n] = func;
    vals[n] = occ;
    ++n;
  }
  return occ;
}

rl_tmap_encode_fn rl_get_tmap_encode() {
  static rl_tmap_encode_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                         &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<rl_tmap_encode_fn>(p);
  });
  return fn;
}

int rl_make_tmap(CUtensorMap* out, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box) {
  rl_tmap_encode_fn enc = rl_get_tmap_encode();
  RL_REQUIRE(enc != nullptr, RL_EDRIVER, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dtype == RL_TMAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RL_REQUIRE(r == CUDA_SUCCESS, RL_EDRIVER,
             "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0,
             rank > 2 ? box[2] : 0);
  return 0;
}

extern "C" int64_t rl_workspace_bytes(const char* op, int64_t B, int64_t L, int64_t H) {
  if (!op) return -1;
  const std::string s(op);
  if (s == "gate_fuse_fwd") return B * 3 * (int64_t)sizeof(float);
  if (s == "gate_fuse_bwd") return (B * L * 3 + 2 * B * H) * (int64_t)sizeof(float);
  if (s == "masked_ce_fwd") return B * L * (int64_t)sizeof(float);
  if (s == "mt_sumsq") return 8LL * rl_num_sms() * (int64_t)sizeof(float);
  return -1;
}

extern "C" int rl_version(void) { return 200; }
extern "C" const char* rl_last_error(void) { return g_err; }
