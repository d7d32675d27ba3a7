// common.cuh - shared host/device helpers for libb200ldm (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <cstdarg>
#include <cstring>

#include "../../include/b200ldm.h"

namespace b200ldm {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define B2_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::b200ldm::fail(B200LDM_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr,         \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define B2_REQUIRE(cond, code, ...)                                 \
  do {                                                              \
    if (!(cond)) return ::b200ldm::fail((code), __VA_ARGS__);       \
  } while (0)

// count + check a kernel launch (cudaPeekAtLastError is capture-safe)
#define B2_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    ::b200ldm::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess)                                                                 \
      return ::b200ldm::fail(B200LDM_CUDA_ERROR, "kernel launch failed: %s (%s:%d)",       \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

// ---------------------------------------------------------------- per-device state (host)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), the SM count and measured kernel choices belong to ONE device; a
// process may drive several (a serving replica looping over GPUs, torch.cuda.set_device(1) after a run on device 0),
// so every cached launch setup is keyed by the current device.
constexpr int kMaxDevices = 64;
int current_device();   // cudaGetDevice (thread-local, no driver round trip), clamped to [0, kMaxDevices)
int sm_count();         // multiprocessors of the current device (cached per device; 148 on B200)

// opt a kernel in to > 48 KB of dynamic shared memory once per device (thread-safe: the attribute call is idempotent,
// the bit is only an optimisation).  Variadic so that template-ids with commas can be passed.
#define B2_SET_SMEM_ONCE(bytes, ...)                                                                        \
  do {                                                                                                      \
    static std::atomic<uint64_t> _b2_done{0};                                                               \
    const uint64_t _b2_bit = 1ull << ::b200ldm::current_device();                                           \
    if (!(_b2_done.load(std::memory_order_acquire) & _b2_bit)) {                                            \
      B2_CHECK_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes))); \
      _b2_done.fetch_or(_b2_bit, std::memory_order_release);                                                \
    }                                                                                                       \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- device scalar helpers
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// fast path: ex2.approx + rcp.approx (2 ulp) instead of the ~10-instruction IEEE division
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// bf16 outputs: x*sigmoid(x) = h + h*tanh(h), h = x/2: ONE MUFU op (tanh.approx, rel. err 2^-11 < bf16 ulp) instead of
// ex2 + rcp - the GroupNorm+SiLU pass was XU-pipe limited (ncu: 66 % XU at 53 % DRAM, profiles/r01_ncu_full_final.md)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
template <typename T>
__device__ __forceinline__ float silu_for(float x) { return silu_f(x); }
template <>
__device__ __forceinline__ float silu_for<__nv_bfloat16>(float x) { return silu_tanh(x); }

// 8 consecutive elements <-> 8 floats, 16-byte (bf16) / 2x16-byte (fp32) vector access
template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200ldm
