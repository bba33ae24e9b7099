// logit_losses.cu -- the two logit-level losses of the incremental step, forward + backward in one kernel each (sm_100a).
//
//  * abr_roi_distillation_id : calculate_roi_distillation_losses(dist='id') (distillation/distillation.py:164-241 of the
//    reference): unbiased cross-entropy between teacher and student class logits (the student's background = its own
//    background + every class the teacher never saw) + L2 between the old classes' box deltas.
//  * abr_fastrcnn_loss       : FastRCNNLossComputation.__call__ (modeling/roi_heads/box_head/loss.py:122-184): inclusive
//    classification loss (dist_type 'id': background = background + the n_old old classes, old-class columns score 0) or
//    plain cross-entropy, + smooth-L1 on the box deltas of the positive rows' own class.
// The reference builds each from ~25 tiny PyTorch kernels and autograd replays ~40 more; the tensors are [R, 21]-sized, so
// that is pure launch latency.  Here one warp owns one row (lanes over classes), the gradient is the closed form, and the
// per-row loss terms are reduced in fixed order by the last CTA to finish (deterministic, no float atomics).
#include "common.cuh"

namespace abr {

constexpr int kRowsPerCta = 8;  // warps per CTA, one row each

struct LossTail {
  float* partials;        // [R][2]
  unsigned int* counter;  // zeroed before launch
  float* loss_out;
};

// Last CTA: out[0..2] from the per-row partials (double accumulation, fixed order).
// mode 0 (distillation): cls = -sum0 / R, box = sum1 / R, out = {cls + box, cls, box}
// mode 1 (box head):     cls = -sum0 / n_valid, box = sum1 / R, out = {cls, box}
__device__ __forceinline__ void finish_losses(const LossTail& t, int R, int mode, int n_valid) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(t.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    const volatile float* part = t.partials;
    for (int i = threadIdx.x; i < R; i += 32) { a += (double)part[2 * i]; b += (double)part[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      if (mode == 0) {
        const double cls = -a / (double)R, box = b / (double)R;
        t.loss_out[0] = (float)(cls + box);
        t.loss_out[1] = (float)cls;
        t.loss_out[2] = (float)box;
      } else {
        t.loss_out[0] = (float)(-a / (double)n_valid);  // 0/0 = NaN when every label is ignored, like F.nll_loss
        t.loss_out[1] = (float)(b / (double)R);
      }
    }
  }
}

__global__ void __launch_bounds__(32 * kRowsPerCta) roi_distill_id_kernel(
    const float* __restrict__ s_scores, const float* __restrict__ s_boxes, const float* __restrict__ t_scores,
    const float* __restrict__ t_boxes, int R, int Co, int Ct, float grad_scale, float* __restrict__ g_scores,
    float* __restrict__ g_boxes, LossTail tail) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * kRowsPerCta + warp;
  if (r < R) {
    const float* ts = t_scores + (size_t)r * Ct;
    const float* ss = s_scores + (size_t)r * Co;
    // student: max, sum over all classes, sum over the background set {0} u [Co, Ct)
    float mt = -INFINITY;
    for (int k = lane; k < Ct; k += 32) mt = fmaxf(mt, ts[k]);
    mt = warp_max(mt);
    float S = 0.f, SB = 0.f;
    for (int k = lane; k < Ct; k += 32) {
      const float e = expf(ts[k] - mt);
      S += e;
      if (k == 0 || k >= Co) SB += e;
    }
    S = warp_sum(S); SB = warp_sum(SB);
    const float logS = logf(S), den = mt + logS;
    // teacher softmax
    float ms = -INFINITY;
    for (int c = lane; c < Co; c += 32) ms = fmaxf(ms, ss[c]);
    ms = warp_max(ms);
    float Z = 0.f;
    for (int c = lane; c < Co; c += 32) Z += expf(ss[c] - ms);
    Z = warp_sum(Z);
    const float p0 = expf(ss[0] - ms) / Z;
    float acc = 0.f;
    for (int c = 1 + lane; c < Co; c += 32) acc = fmaf(expf(ss[c] - ms) / Z, ts[c] - den, acc);
    acc = warp_sum(acc);
    const float cls_row = (p0 * (logf(SB) - logS) + acc) / (float)Co;
    if (g_scores) {
      const float k0 = -grad_scale / ((float)R * (float)Co);
      float* g = g_scores + (size_t)r * Ct;
      for (int k = lane; k < Ct; k += 32) {
        const float e = expf(ts[k] - mt);
        const float a = (k == 0 || k >= Co) ? p0 * (e / SB) : expf(ss[k] - ms) / Z;
        g[k] = k0 * (a - e / S);
      }
    }
    // boxes of the old foreground classes 1..Co-1: sum of squares, mean over the classes
    const float4* tb = reinterpret_cast<const float4*>(t_boxes) + (size_t)r * Ct;
    const float4* sb = reinterpret_cast<const float4*>(s_boxes) + (size_t)r * Co;
    float4* gb = g_boxes ? reinterpret_cast<float4*>(g_boxes) + (size_t)r * Ct : nullptr;
    const float kb = grad_scale * 2.f / ((float)R * (float)(Co - 1));
    float sq = 0.f;
    for (int c = lane; c < Ct; c += 32) {
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c >= 1 && c < Co) {
        const float4 a = tb[c], b = sb[c];
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        sq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        gv = make_float4(kb * d0, kb * d1, kb * d2, kb * d3);
      }
      if (gb) gb[c] = gv;
    }
    sq = warp_sum(sq);
    if (lane == 0) {
      tail.partials[2 * r] = cls_row;
      tail.partials[2 * r + 1] = sq / (float)(Co - 1);  // 0/0 = NaN for a teacher with background only, like torch.mean of nothing
    }
  }
  finish_losses(tail, R, 0, R);
}

__global__ void __launch_bounds__(32 * kRowsPerCta) fastrcnn_loss_kernel(
    const float* __restrict__ logits, const float* __restrict__ regression, int reg_stride, const long long* __restrict__ labels,
    const float* __restrict__ targets, int R, int C, int n_old, int cls_agnostic, float beta, float scale_cls, float scale_box,
    float* __restrict__ g_logits, float* __restrict__ g_reg, LossTail tail) {
  __shared__ int valid_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // rows that count in the mean of F.nll_loss / F.cross_entropy (ignore_index = -100); every CTA counts them itself
  if (threadIdx.x == 0) valid_sh = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) mine += labels[i] != -100;
  mine = (int)warp_sum((float)mine);  // R < 2^24 rows
  if (lane == 0 && mine) atomicAdd(&valid_sh, mine);
  __syncthreads();
  const int n_valid = valid_sh;
  const int r = blockIdx.x * kRowsPerCta + warp;
  if (r < R) {
    const float* l = logits + (size_t)r * C;
    const long long y = labels[r];
    const bool valid = y != -100;
    float m = -INFINITY;
    for (int k = lane; k < C; k += 32) m = fmaxf(m, l[k]);
    m = warp_max(m);
    float S = 0.f, So = 0.f;
    for (int k = lane; k < C; k += 32) {
      const float e = expf(l[k] - m);
      S += e;
      if (k <= n_old) So += e;
    }
    S = warp_sum(S); So = warp_sum(So);
    const float logS = logf(S);
    // which log-probability the row contributes: 0 = merged background, 1 = its own class, 2 = nothing (an old class under 'id')
    int kind = 1;
    if (n_old >= 0) kind = (y == 0) ? 0 : (y > n_old ? 1 : 2);
    float out = 0.f;
    if (valid && y >= 0 && y < C) {
      if (kind == 0) out = logf(So) - logS;
      else if (kind == 1) out = l[y] - m - logS;
    }
    if (g_logits) {
      float* g = g_logits + (size_t)r * C;
      const float k0 = -scale_cls / (float)n_valid;
      for (int k = lane; k < C; k += 32) {
        float v = 0.f;
        if (valid && kind != 2 && y >= 0 && y < C) {
          const float e = expf(l[k] - m);
          const float a = kind == 0 ? (k <= n_old ? e / So : 0.f) : (k == (int)y ? 1.f : 0.f);
          v = k0 * (a - e / S);
        }
        g[k] = v;
      }
    }
    // smooth-L1 on the positive rows' own class columns (loss.py:166-180), divided by ALL rows
    float box = 0.f;
    const bool pos = y > 0 && y < C;
    const int col0 = cls_agnostic ? 4 : 4 * (int)(pos ? y : 0);
    if (g_reg) {
      float* g = g_reg + (size_t)r * reg_stride;
      for (int k = lane; k < reg_stride; k += 32) {
        float v = 0.f;
        if (pos && k >= col0 && k < col0 + 4) {
          const float d = regression[(size_t)r * reg_stride + k] - targets[(size_t)r * 4 + (k - col0)];
          const float n = fabsf(d);
          v = (n < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f))) * (scale_box / (float)R);
        }
        g[k] = v;
      }
    }
    if (pos && lane < 4) {
      const float d = regression[(size_t)r * reg_stride + col0 + lane] - targets[(size_t)r * 4 + lane];
      const float n = fabsf(d);
      box = n < beta ? 0.5f * n * n / beta : n - 0.5f * beta;
    }
    box = warp_sum(box);
    if (lane == 0) {
      tail.partials[2 * r] = out;
      tail.partials[2 * r + 1] = box;
    }
  }
  finish_losses(tail, R, 1, n_valid);
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_logit_loss_workspace_bytes(int R) {
  if (R < 0) return 0;
  return 256 + (size_t)R * 2 * sizeof(float);
}

static int loss_tail(LossTail& t, void* workspace, size_t workspace_bytes, int R, float* out, cudaStream_t st, const char* who) {
  ABR_REQUIRE(workspace && workspace_bytes >= abr_logit_loss_workspace_bytes(R), ABR_ERR_WORKSPACE, "%s: workspace %zu B < %zu B", who,
              workspace_bytes, abr_logit_loss_workspace_bytes(R));
  t.counter = static_cast<unsigned int*>(workspace);
  t.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  t.loss_out = out;
  ABR_CUDA_OK(cudaMemsetAsync(t.counter, 0, sizeof(unsigned int), st));
  return ABR_OK;
}

int abr_roi_distillation_id(const float* soften_scores, const float* soften_bboxes, const float* target_scores,
                            const float* target_bboxes, int R, int C_old, int C_total, float grad_scale, float* grad_scores,
                            float* grad_bboxes, float* loss3, void* workspace, size_t workspace_bytes, abr_stream_t stream) {
  ABR_REQUIRE(R > 0 && C_old >= 1 && C_total > C_old, ABR_ERR_BAD_ARG,
              "roi_distillation: R=%d, teacher classes %d, student classes %d (the student must know more classes)", R, C_old, C_total);
  ABR_REQUIRE(soften_scores && soften_bboxes && target_scores && target_bboxes && loss3, ABR_ERR_BAD_ARG, "roi_distillation: null pointer");
  ABR_REQUIRE(((reinterpret_cast<uintptr_t>(soften_bboxes) | reinterpret_cast<uintptr_t>(target_bboxes) | reinterpret_cast<uintptr_t>(grad_bboxes)) & 15) == 0,
              ABR_ERR_BAD_ARG, "roi_distillation: box tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossTail t;
  int rc = loss_tail(t, workspace, workspace_bytes, R, loss3, st, "roi_distillation");
  if (rc) return rc;
  roi_distill_id_kernel<<<ceil_div(R, kRowsPerCta), 32 * kRowsPerCta, 0, st>>>(soften_scores, soften_bboxes, target_scores, target_bboxes, R,
                                                                             C_old, C_total, grad_scale, grad_scores, grad_bboxes, t);
  ABR_CHECK_LAUNCH("roi_distillation_id");
  return ABR_OK;
}

int abr_fastrcnn_loss(const float* class_logits, const float* box_regression, int reg_row_stride, const int64_t* labels,
                      const float* regression_targets, int R, int num_classes, int n_old, int cls_agnostic, float beta,
                      float grad_scale_cls, float grad_scale_box, float* grad_logits, float* grad_regression, float* loss2,
                      void* workspace, size_t workspace_bytes, abr_stream_t stream) {
  ABR_REQUIRE(R > 0 && num_classes >= 1, ABR_ERR_BAD_ARG, "fastrcnn_loss: R=%d C=%d", R, num_classes);
  ABR_REQUIRE(class_logits && box_regression && labels && regression_targets && loss2, ABR_ERR_BAD_ARG, "fastrcnn_loss: null pointer");
  ABR_REQUIRE(reg_row_stride >= (cls_agnostic ? 8 : 4 * num_classes), ABR_ERR_BAD_ARG, "fastrcnn_loss: box_regression rows of %d floats, need %d",
              reg_row_stride, cls_agnostic ? 8 : 4 * num_classes);
  ABR_REQUIRE(n_old < num_classes, ABR_ERR_BAD_ARG, "fastrcnn_loss: n_old=%d with %d classes", n_old, num_classes);
  ABR_REQUIRE(beta > 0.f, ABR_ERR_BAD_ARG, "fastrcnn_loss: beta must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossTail t;
  int rc = loss_tail(t, workspace, workspace_bytes, R, loss2, st, "fastrcnn_loss");
  if (rc) return rc;
  fastrcnn_loss_kernel<<<ceil_div(R, kRowsPerCta), 32 * kRowsPerCta, 0, st>>>(class_logits, box_regression, reg_row_stride,
                                                                            reinterpret_cast<const long long*>(labels), regression_targets,
                                                                            R, num_classes, n_old < 0 ? -1 : n_old, cls_agnostic, beta,
                                                                            grad_scale_cls, grad_scale_box, grad_logits, grad_regression, t);
  ABR_CHECK_LAUNCH("fastrcnn_loss");
  return ABR_OK;
}

}  // extern "C"
