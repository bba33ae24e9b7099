// Stand-in for <THC/THCDeviceUtils.cuh>: THCCeilDiv only.  TEST INFRASTRUCTURE (see THC.h next to this file).
#pragma once
template <typename T>
__host__ __device__ __forceinline__ T THCCeilDiv(T a, T b) { return (a + b - 1) / b; }
