// Speed-of-light probe: does a TMA bulk reduction (cp.reduce.async.bulk ... .add.f32, shared -> global) push more
// fp32 adds per second into an L2-resident map than red.global.add.v4.f32 issued from registers?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_reduce tma_reduce.cu && ./tma_reduce
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// one elected thread per warp issues bulk reductions of `bytes` from the CTA's shared buffer to pseudo-random,
// `bytes`-aligned places of the map; `depth` bulk groups in flight per issuing thread
__global__ void bulk_reduce_kernel(float* map, size_t map_bytes, int bytes, int iters, int depth) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) buf[i] = 1.f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;
  const unsigned nchunks = (unsigned)(map_bytes / bytes);
  unsigned st = hash(warp * 2654435761u + 7);
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(buf);
  for (int i = 0; i < iters; i++) {
    st = hash(st + i);
    char* dst = reinterpret_cast<char*>(map) + (size_t)(st % nchunks) * bytes;
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (i >= depth) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

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

int main() {
  const int npix = 4 * 50 * 76, slices = 8;  // [4,50,76,1024] fp32 = 62 MB
  const size_t bytes_total = (size_t)npix * slices * 512;
  float* map;
  CK(cudaMalloc(&map, bytes_total));
  CK(cudaMemset(map, 0, bytes_total));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  {
    const int iters = 2000, blocks = 148 * 8, threads = 256;
    red_kernel<<<blocks, threads>>>(reinterpret_cast<float4*>(map), npix, slices, 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    red_kernel<<<blocks, threads>>>(reinterpret_cast<float4*>(map), npix, slices, iters);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("red.v4.f32 512B/warp: %.2f TB/s payload\n", (double)blocks * threads / 32 * iters * 512 / ms * 1e-9);
  }
  for (int bytes = 512; bytes <= 16384; bytes *= 2) {
    for (int warps = 1; warps <= 8; warps *= 2) {
      const int blocks = 148 * 4, threads = warps * 32, iters = (int)(64.0 * 1024 * 1024 / bytes / warps) + 8;
      CK(cudaFuncSetAttribute(bulk_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
      bulk_reduce_kernel<<<blocks, threads, bytes>>>(map, bytes_total, bytes, 8, 8);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      bulk_reduce_kernel<<<blocks, threads, bytes>>>(map, bytes_total, bytes, iters, 8);
      cudaEventRecord(b); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("bulk reduce %5d B x %d issuing warps/CTA x %d CTAs: %.2f TB/s payload (%.1f M ops/s)\n", bytes, warps, blocks,
             (double)blocks * warps * iters * bytes / ms * 1e-9, (double)blocks * warps * iters / ms * 1e-3);
    }
  }
  return 0;
}
