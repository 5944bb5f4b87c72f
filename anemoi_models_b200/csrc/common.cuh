// Shared device/host helpers for libanemoi_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/anemoi_b200.h"

namespace ab2 {

// ---- error plumbing ---------------------------------------------------------------------------------
char* last_error_buf();  // thread-local, defined in abi.cu
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define AB2_CUDA_OK(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return ab2::fail(AB2_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
long long* launch_counter();  // process-wide count of kernel launches issued by this library (defined in abi.cu)
#define AB2_LAUNCH_OK(name)                                                                            \
  do {                                                                                                 \
    __atomic_add_fetch(ab2::launch_counter(), 1, __ATOMIC_RELAXED);                                    \
    cudaError_t _e = cudaGetLastError();                                                               \
    if (_e != cudaSuccess) return ab2::fail(AB2_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel instantiation, device), under a
// mutex (calls arrive from the main thread and from autograd's worker threads).  Expands inside the launcher of ONE kernel
// instantiation, so the function-local statics are per kernel.  Evaluates to false when the attribute cannot be set.
#define AB2_ENSURE_DYN_SMEM(kern, bytes)                                                                          \
  ([&]() -> bool {                                                                                                \
    static std::mutex mu_;                                                                                        \
    static bool done_[64] = {false};                                                                              \
    int dev_ = 0;                                                                                                 \
    if (cudaGetDevice(&dev_) != cudaSuccess) return false;                                                        \
    std::lock_guard<std::mutex> lock_(mu_);                                                                       \
    if (dev_ >= 0 && dev_ < 64 && done_[dev_]) return true;                                                       \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess) return false; \
    if (dev_ >= 0 && dev_ < 64) done_[dev_] = true;                                                               \
    return true;                                                                                                  \
  }())

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- 16-byte vectors of T ---------------------------------------------------------------------------
template <typename T>
struct Vec;  // VEC elements of T in one 16-byte load
template <>
struct Vec<float> {
  static constexpr int N = 4;
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
};

// read-only, streaming 16-byte load (rows are touched once or twice; keep L1 for the index arrays)
__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// default-cached 16-byte load (for rows that are re-read by neighbouring CTAs: q, g, k, v)
__device__ __forceinline__ uint4 ldg16_keep(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
__device__ __forceinline__ void unpack(const uint4& r, float (&f)[Vec<T>::N]);
template <>
__device__ __forceinline__ void unpack<float>(const uint4& r, float (&f)[4]) {
  f[0] = __uint_as_float(r.x);
  f[1] = __uint_as_float(r.y);
  f[2] = __uint_as_float(r.z);
  f[3] = __uint_as_float(r.w);
}
template <>
__device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& r, float (&f)[8]) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(r.x << 16);
  f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16);
  f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16);
  f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16);
  f[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T>
__device__ __forceinline__ uint4 pack(const float (&f)[Vec<T>::N]);
template <>
__device__ __forceinline__ uint4 pack<float>(const float (&f)[4]) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <>
__device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <typename T>
__device__ __forceinline__ float to_f(T x);
template <>
__device__ __forceinline__ float to_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sum over the LPH consecutive lanes of a head group (LPH is a power of two <= 32); `mask` names exactly those lanes
template <int LPH>
__device__ __forceinline__ float group_sum(float x, unsigned mask) {
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
  return x;
}

constexpr float kLog2e = 1.4426950408889634f;

}  // namespace ab2
