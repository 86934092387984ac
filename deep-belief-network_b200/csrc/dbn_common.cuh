// dbn_common.cuh -- shared host/device helpers of the dbn_b200 C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <string.h>
#include <stdlib.h>
#include <utility>

#include "dbn_b200.h"

typedef __nv_bfloat16 bf16;

namespace dbn {

// ---- error plumbing (never throws across the C boundary) --------------------------------------
inline std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  char buf[512];
  snprintf(buf, sizeof(buf), fmt, a, b);
  last_error_ref() = buf;
  return code;
}
inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> n{0};
  return n;
}

#define DBN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return ::dbn::fail(DBN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define DBN_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    ::dbn::launch_counter().fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) return ::dbn::fail(DBN_ERR_CUDA, "kernel launch %s: %s", __func__, cudaGetErrorString(_e)); \
  } while (0)

#define DBN_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) return ::dbn::fail(DBN_ERR_INVALID_ARG, "%s: %s", __func__, msg);          \
  } while (0)

#define DBN_TRY(expr)                                                                       \
  do {                                                                                      \
    int _rc = (expr);                                                                       \
    if (_rc != DBN_OK) return _rc;                                                          \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- per-device one-time state --------------------------------------------------------------------
// Function attributes (opt-in shared memory), SM counts and occupancy answers belong to a DEVICE, not to the
// process: a caller may fit on cuda:0 and then on cuda:1.  Caches are therefore indexed by the current device.
constexpr int DBN_MAX_DEVICES = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); d = 0; }
  return (d < 0 || d >= DBN_MAX_DEVICES) ? 0 : d;
}
// true exactly once per (mask, current device)
inline bool first_use_on_device(std::atomic<uint64_t>& mask) {
  const uint64_t bit = uint64_t(1) << current_device();
  return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every contraction kernel is launched with programmatic stream serialization: it may start (and run
// its prologue: barrier init, TMEM allocation, tensor-map prefetch) while its predecessor in the
// stream is still executing, and touches global memory only after pdl_wait(), which returns once the
// predecessor grid has completed and flushed.  At batch 256 the step is a chain of ~10 us kernels, so
// hiding launch latency and prologue is worth ~25 % of the step.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_PDL");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- typed matrix access ------------------------------------------------------------------------
// two floats <-> one 32-bit word of a 16-bit matrix (bf16 or fp16)
__device__ __forceinline__ uint32_t pack16x2(int dtype, float a, float b) {
  if (dtype == DBN_F16) {   // saturating: an out-of-range input loses precision instead of turning into inf
    const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.0f), 65504.0f), fminf(fmaxf(b, -65504.0f), 65504.0f));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack16x2(int dtype, uint32_t w, float& a, float& b) {
  if (dtype == DBN_F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    a = f.x; b = f.y;
  } else {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w);
    a = __low2float(h); b = __high2float(h);
  }
}
__device__ __forceinline__ float mat_load(const dbn_mat& m, int r, int c) {
  size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (m.dtype == DBN_F32) return reinterpret_cast<const float*>(m.ptr)[i];
  if (m.dtype == DBN_F16) return __half2float(reinterpret_cast<const __half*>(m.ptr)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(m.ptr)[i]);
}
__device__ __forceinline__ void mat_store(const dbn_mat& m, int r, int c, float v) {
  size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (m.dtype == DBN_F32) reinterpret_cast<float*>(m.ptr)[i] = v;
  else if (m.dtype == DBN_F16) reinterpret_cast<__half*>(m.ptr)[i] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
  else reinterpret_cast<bf16*>(m.ptr)[i] = __float2bfloat16_rn(v);
}
// 4 consecutive columns starting at a multiple of 4; vectorised when the row pitch allows it.
__device__ __forceinline__ void mat_store4(const dbn_mat& m, int r, int c, const float v[4], int nvalid) {
  size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (nvalid == 4 && (m.ld & 3) == 0) {
    if (m.dtype == DBN_F32) {
      float* p = reinterpret_cast<float*>(m.ptr) + i;
      if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        return;
      }
    } else {
      bf16* p = reinterpret_cast<bf16*>(m.ptr) + i;
      if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        uint2 pk;
        pk.x = pack16x2(m.dtype, v[0], v[1]);
        pk.y = pack16x2(m.dtype, v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = pk;
        return;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < nvalid) mat_store(m, r, c + j, v[j]);
}
__device__ __forceinline__ void mat_load4(const dbn_mat& m, int r, int c, float v[4], int nvalid) {
  size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (nvalid == 4 && (m.ld & 3) == 0) {
    if (m.dtype == DBN_F32) {
      const float* p = reinterpret_cast<const float*>(m.ptr) + i;
      if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        return;
      }
    } else {
      const bf16* p = reinterpret_cast<const bf16*>(m.ptr) + i;
      if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        uint2 pk = *reinterpret_cast<const uint2*>(p);
        unpack16x2(m.dtype, pk.x, v[0], v[1]);
        unpack16x2(m.dtype, pk.y, v[2], v[3]);
        return;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (j < nvalid) ? mat_load(m, r, c + j) : 0.f;
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------------
// Counter-based generator (Salmon et al., SC'11).  counter = (row, col/4, epoch, stream), key = seed.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 24-bit uniform in [0,1): exactly representable in fp32, so `u < p` is the same comparison in
// the fp32 kernels and in the float64 oracle.
__host__ __device__ __forceinline__ float u01_from_bits(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// 4 uniforms for columns n4..n4+3 (n4 % 4 == 0) of global row `row`.
__device__ __forceinline__ void rng_uniform4(const dbn_rng& g, uint32_t epoch, int m, int n4, float u[4]) {
  uint32_t o[4];
  philox4x32_10(g.row0 + (uint32_t)m, (g.col0 + (uint32_t)n4) >> 2, epoch, g.stream, (uint32_t)g.seed,
                (uint32_t)(g.seed >> 32), o);
#pragma unroll
  for (int j = 0; j < 4; ++j) u[j] = u01_from_bits(o[j]);
}

}  // namespace dbn
