// Speed-of-light probes for the ROIAlign kernels on B200: how fast can SMs (a) gather 512 B-per-warp pixel segments from
// an L2-resident NHWC map, (b) push 512 B-per-warp vector reductions (red.global.add.v4.f32) into it.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_limits l2_limits.cu && ./l2_limits
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each warp reads `iters` pixel segments (32 lanes x 16 B) at pseudo-random pixels; ILP independent loads in flight
template <int ILP>
__global__ void gather_kernel(const float4* __restrict__ map, int npix, int slices, int iters, float4* sink) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float4 acc = make_float4(0, 0, 0, 0);
  unsigned st = hash(warp * 2654435761u + 1);
  for (int i = 0; i < iters; i += ILP) {
    float4 v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      st = hash(st + k);
      const unsigned pix = st % npix, sl = (st >> 20) % slices;
      v[k] = __ldg(map + ((size_t)pix * slices + sl) * 32 + lane);
    }
#pragma unroll
    for (int k = 0; k < ILP; k++) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

template <int ILP>
__global__ void red_kernel(float4* map, int npix, int slices, int iters) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  unsigned st = hash(warp * 2654435761u + 7);
  for (int i = 0; i < iters; i++) {
    st = hash(st + i);
    const unsigned pix = st % npix, sl = (st >> 20) % slices;
    float* p = reinterpret_cast<float*>(map + ((size_t)pix * slices + sl) * 32 + lane);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
  }
}

__global__ void red_scalar_kernel(float* map, int npix, int slices, int iters) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  unsigned st = hash(warp * 2654435761u + 7);
  for (int i = 0; i < iters; i++) {
    st = hash(st + i);
    const unsigned pix = st % npix, sl = (st >> 20) % slices;
    atomicAdd(map + (((size_t)pix * slices + sl) * 32 + lane) * 4, 1.f);
  }
}

int main() {
  const int npix = 4 * 50 * 76, slices = 8;  // [4,50,76,1024] fp32 = 62 MB
  const size_t n4 = (size_t)npix * slices * 32;
  float4* map; float4* sink;
  CK(cudaMalloc(&map, n4 * 16)); CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(map, 0, n4 * 16));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int wps = 8; wps <= 64; wps *= 2) {            // resident warps per SM
    const int blocks = 148 * wps / 8, iters = 2048;
    const double bytes = (double)blocks * 8 * iters * 512;
    float ms;
    gather_kernel<4><<<blocks, 256>>>(map, npix, slices, iters, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a); gather_kernel<4><<<blocks, 256>>>(map, npix, slices, iters, sink); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b); printf("gather  ILP4 %2d warps/SM: %7.1f GB/s\n", wps, bytes / ms / 1e6);
    cudaEventRecord(a); gather_kernel<1><<<blocks, 256>>>(map, npix, slices, iters, sink); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b); printf("gather  ILP1 %2d warps/SM: %7.1f GB/s\n", wps, bytes / ms / 1e6);
    cudaEventRecord(a); red_kernel<1><<<blocks, 256>>>(map, npix, slices, iters); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b); printf("red.v4       %2d warps/SM: %7.1f GB/s payload\n", wps, bytes / ms / 1e6);
    cudaEventRecord(a); red_scalar_kernel<<<blocks, 256>>>(reinterpret_cast<float*>(map), npix, slices, iters); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b); printf("red.f32 x1   %2d warps/SM: %7.1f GB/s payload (128 B per warp request)\n", wps, bytes / 4 / ms / 1e6);
  }
  return 0;
}
