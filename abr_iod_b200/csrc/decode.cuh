// decode.cuh -- BoxCoder.decode + clip_to_image for one box (shared by rpn.cu and box_post.cu).
#pragma once
#include "common.cuh"

namespace abr {

struct BoxCoderParams {
  float wx, wy, ww, wh, clip;
};

// modeling/box_coder.py:52-95 of the reference for one (deltas, reference box) pair, one rounding per tensor op (no FMA
// contraction), followed by structures/bounding_box.py:214-219 clip_to_image(remove_empty=False) to [0, w-1] x [0, h-1].
__device__ __forceinline__ float4 decode_and_clip(const float4 an, float r0, float r1, float r2, float r3, const BoxCoderParams& c,
                                                  float xmax, float ymax) {
  const float widths = __fadd_rn(__fsub_rn(an.z, an.x), 1.f);
  const float heights = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
  const float ctr_x = __fadd_rn(an.x, __fmul_rn(0.5f, widths));
  const float ctr_y = __fadd_rn(an.y, __fmul_rn(0.5f, heights));
  const float dx = __fdiv_rn(r0, c.wx), dy = __fdiv_rn(r1, c.wy);
  float dw = __fdiv_rn(r2, c.ww), dh = __fdiv_rn(r3, c.wh);
  dw = dw > c.clip ? c.clip : dw;  // torch.clamp(max=): NaN stays NaN
  dh = dh > c.clip ? c.clip : dh;
  const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x);
  const float pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
  const float pw = __fmul_rn(expf(dw), widths);
  const float ph = __fmul_rn(expf(dh), heights);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  float y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f);
  float y2 = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f);
  x1 = fminf(fmaxf(x1, 0.f), xmax); y1 = fminf(fmaxf(y1, 0.f), ymax);
  x2 = fminf(fmaxf(x2, 0.f), xmax); y2 = fminf(fmaxf(y2, 0.f), ymax);
  return make_float4(x1, y1, x2, y2);
}

// order-preserving 32-bit key of a float (larger float = larger key; NaN largest; -0.0 == +0.0) and its inverse
__device__ __forceinline__ unsigned ordered_bits(float s) {
  unsigned u = __float_as_uint(s);
  if (s != s) return 0xFFFFFFFFu;
  if (u == 0x80000000u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

}  // namespace abr
