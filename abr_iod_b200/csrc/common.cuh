// common.cuh -- shared helpers of libabr_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/abr_b200.h"

namespace abr {

// ---- host-side error plumbing (thread-local text, integer codes; see abr_b200.h) ----
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define ABR_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::abr::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define ABR_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::abr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return ABR_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define ABR_CHECK_LAUNCH(name)                                                         \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::abr::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));      \
      return ABR_ERR_CUDA;                                                             \
    }                                                                                  \
    ::abr::count_launch();                                                             \
  } while (0)

// Process-wide tuning switches (api.cu: abr_set_option; initial values from the environment).
struct Options {
  int roi_v2;       // -1: v2 kernels where the staged ones do not apply (output larger than 8x7); 0: never; 1: whenever supported
  int fwd_tma, bwd_tma, ard_cluster;  // 0 disables the TMA-staged ROIAlign kernels / the cluster ARD kernels
};
Options& options();

// Stage timing (api.cu; abr_stage_timing_begin / _end): stage_mark(st, k) records event k of the current call on `st`
// when timing is on (k = 0 opens a call, k = n_stages closes it); a no-op otherwise.
void stage_mark(cudaStream_t st, int k);

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- device-side element access: V consecutive elements <-> float[V] ----
template <typename T, int V>
struct VecIO;

template <>
struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
  static __device__ __forceinline__ void store_stream(float* p, const float (&v)[1]) { __stcs(p, v[0]); }
  static __device__ __forceinline__ void red_add(float* p, const float (&v)[1]) { atomicAdd(p, v[0]); }
};
template <>
struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  // write-once tensors (pooled outputs, gradients of pooled tensors): st.global.cs, evict-first in L2, so that the
  // stream of output lines does not push the re-read feature maps out of the cache
  static __device__ __forceinline__ void store_stream(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  // one 16-byte reduction per thread: a warp issues a single contiguous 512 B request
  static __device__ __forceinline__ void red_add(float* p, const float (&v)[4]) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
  }
};
template <>
struct VecIO<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float (&v)[1]) {
    __stcs(reinterpret_cast<unsigned short*>(p), __bfloat16_as_ushort(__float2bfloat16_rn(v[0])));
  }
  static __device__ __forceinline__ void red_add(__nv_bfloat16* p, const float (&v)[1]) {
    atomicAdd(p, __float2bfloat16_rn(v[0]));
  }
};
template <>
struct VecIO<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    __stcs(reinterpret_cast<uint4*>(p), t);
  }
  static __device__ __forceinline__ void red_add(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w)
                 : "memory");
  }
};

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

}  // namespace abr
