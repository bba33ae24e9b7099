// emu_shim.h -- lets abr_iod_b200/csrc/roi_v2.cuh compile as plain host C++ (g++ -ffp-contract=off) so that the very
// device logic of the v2 ROIAlign kernels can be checked against the CPU oracle without a GPU.  TEST INFRASTRUCTURE.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ABR_EMU 1
#define ABR_MAX_LEVELS 8
#define ABR_DEV static inline
#define ABR_DEV_COLD static inline
#define ABR_DEVM inline
#define ABR_HD static inline
#define ABR_HOSTDEV static inline
#define __restrict__
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#define ABR_LDG4I(p) (*reinterpret_cast<const int4*>(p))
#define ABR_LDGI(p) (*(p))
#define ABR_LDG2F(p) (*(p))

namespace abr {
template <typename T, int V>
struct VecIO {
  static void load(const T* p, float (&v)[V]) { for (int i = 0; i < V; i++) v[i] = p[i]; }
  static void store(T* p, const float (&v)[V]) { for (int i = 0; i < V; i++) p[i] = v[i]; }
  static void red_add(T* p, const float (&v)[V]) { for (int i = 0; i < V; i++) p[i] += v[i]; }
};
}  // namespace abr
