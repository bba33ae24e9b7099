// prototype.cu -- Prototype Box Selection on the device (sm_100a).
//
//  * abr_channel_mean        : the per-RoI descriptor of tools/prototype_box_selection.py:96-101 of the reference
//                              (torch.mean(roi_align_features.cpu(), dim=1): [R,C,7,7] -> [R,7,7]) without the 411 MB
//                              device-to-host copy it is computed after -- one streaming pass, HBM-bound.
//  * abr_prototype_distances : the scoring of Mem.mean_feature_sampling (tools/extract_memory.py:111-147): class mean of
//                              the descriptors, normalised; descriptors divided by the Frobenius norm of ALL of them;
//                              Euclidean distance of each to the mean -- in float64 like the numpy code.  The caller sorts
//                              the n distances (ascending; the closest num_bbox_per_cls boxes become the prototypes).
#include "common.cuh"

namespace abr {

// NHWC: one warp per (RoI, position) row of C contiguous channels.
template <typename T, int V>
__global__ void __launch_bounds__(256) channel_mean_nhwc_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* p = x + (size_t)row * C;
  float acc = 0.f;
  for (int c = lane * V; c < C; c += 32 * V) {
    float v[V];
    VecIO<T, V>::load(p + c, v);
#pragma unroll
    for (int i = 0; i < V; i++) acc += v[i];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc / (float)C;
}

// NCHW: one CTA per RoI; thread (g, p) owns position p and walks channels g, g+G, ... (consecutive threads touch
// consecutive elements), partial sums combined through shared memory.
template <typename T>
__global__ void __launch_bounds__(1024) channel_mean_nchw_kernel(const T* __restrict__ x, int C, int HW, int G, float* __restrict__ out) {
  extern __shared__ float part[];  // [G][HW]
  const int r = blockIdx.x, tid = threadIdx.x;
  const int g = tid / HW, p = tid - g * HW;
  const T* base = x + (size_t)r * C * HW;
  if (g < G) {
    float acc = 0.f;
    for (int c = g; c < C; c += G) {
      float v[1];
      VecIO<T, 1>::load(base + (size_t)c * HW + p, v);
      acc += v[0];
    }
    part[g * HW + p] = acc;
  }
  __syncthreads();
  if (tid < HW) {
    float acc = 0.f;
    for (int k = 0; k < G; k++) acc += part[k * HW + tid];
    out[(size_t)r * HW + tid] = acc / (float)C;
  }
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nwarp; i++) t += scratch[i];  // same order in every thread
  return t;
}

// One CTA: mu = mean_i f_i, mu /= |mu|;  phi_i = f_i / |F|_Frobenius;  dist_i = |mu - phi_i|  (extract_memory.py:125-141)
__global__ void __launch_bounds__(256) prototype_distance_kernel(const float* __restrict__ f, int n, int F, double* __restrict__ mu,
                                                                double* __restrict__ dist) {
  __shared__ double scratch[8];
  const int tid = threadIdx.x;
  double sq_all = 0.0;
  for (long long i = tid; i < (long long)n * F; i += blockDim.x) { const double v = (double)f[i]; sq_all += v * v; }
  const double fro = sqrt(block_sum(sq_all, scratch));
  double sq_mu = 0.0;
  for (int k = tid; k < F; k += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < n; i++) acc += (double)f[(size_t)i * F + k];
    acc /= (double)n;
    mu[k] = acc;
    sq_mu += acc * acc;
  }
  const double nmu = sqrt(block_sum(sq_mu, scratch));
  for (int k = tid; k < F; k += blockDim.x) mu[k] /= nmu;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < F; k++) { const double d = mu[k] - (double)f[(size_t)i * F + k] / fro; acc += d * d; }
    dist[i] = sqrt(acc);
  }
}

}  // namespace abr

using namespace abr;

extern "C" {

int abr_channel_mean(const void* pooled, int R, int C, int HW, int dtype, int layout, float* out, abr_stream_t stream) {
  ABR_REQUIRE(R >= 0 && C > 0 && HW > 0, ABR_ERR_BAD_ARG, "channel_mean: R=%d C=%d HW=%d", R, C, HW);
  if (R == 0) return ABR_OK;
  ABR_REQUIRE(pooled && out, ABR_ERR_BAD_ARG, "channel_mean: null pointer");
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "channel_mean: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "channel_mean: layout %d not supported", layout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == ABR_NHWC) {
    const long long rows = (long long)R * HW;
    const unsigned blocks = (unsigned)ceil_div<long long>(rows, 8);
    const bool aligned = (reinterpret_cast<uintptr_t>(pooled) & 15) == 0;
    if (dtype == ABR_F32) {
      if (aligned && C % 4 == 0) channel_mean_nhwc_kernel<float, 4><<<blocks, 256, 0, st>>>(static_cast<const float*>(pooled), rows, C, out);
      else channel_mean_nhwc_kernel<float, 1><<<blocks, 256, 0, st>>>(static_cast<const float*>(pooled), rows, C, out);
    } else {
      if (aligned && C % 8 == 0) channel_mean_nhwc_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(pooled), rows, C, out);
      else channel_mean_nhwc_kernel<__nv_bfloat16, 1><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(pooled), rows, C, out);
    }
  } else {
    ABR_REQUIRE(HW <= 1024, ABR_ERR_UNSUPPORTED, "channel_mean: %d positions per RoI (max 1024)", HW);
    int G = 1024 / HW;
    if (G > C) G = C;
    const int threads = ceil_div(G * HW, 32) * 32;
    const size_t smem = (size_t)G * HW * sizeof(float);
    if (dtype == ABR_F32) channel_mean_nchw_kernel<float><<<R, threads, smem, st>>>(static_cast<const float*>(pooled), C, HW, G, out);
    else channel_mean_nchw_kernel<__nv_bfloat16><<<R, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(pooled), C, HW, G, out);
  }
  ABR_CHECK_LAUNCH("channel_mean");
  return ABR_OK;
}

int abr_prototype_distances(const float* features, int n, int F, double* mean_out, double* dist, abr_stream_t stream) {
  ABR_REQUIRE(n > 0 && F > 0, ABR_ERR_BAD_ARG, "prototype_distances: n=%d F=%d", n, F);
  ABR_REQUIRE(features && mean_out && dist, ABR_ERR_BAD_ARG, "prototype_distances: null pointer");
  prototype_distance_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(features, n, F, mean_out, dist);
  ABR_CHECK_LAUNCH("prototype_distances");
  return ABR_OK;
}

}  // extern "C"
