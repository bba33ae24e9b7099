// Stand-in for <THC/THC.h> (removed from PyTorch >= 1.11) -- TEST INFRASTRUCTURE, used only to compile the REFERENCE's
// own csrc/cuda/*.cu in place into oracle/_ref/libabr_ref_cuda.so (oracle/ref_cuda_shim.cu).  Provides exactly the five
// names those files use: THCState, THCudaMalloc, THCudaFree, THCudaCheck and (THCDeviceUtils.cuh) THCCeilDiv.
#pragma once
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAException.h>
#include <cuda_runtime.h>

struct THCState {};
inline void* THCudaMalloc(THCState*, size_t bytes) { return c10::cuda::CUDACachingAllocator::raw_alloc(bytes); }
inline void THCudaFree(THCState*, void* p) { c10::cuda::CUDACachingAllocator::raw_delete(p); }
#define THCudaCheck(expr) C10_CUDA_CHECK(expr)
