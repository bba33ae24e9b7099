// Stand-in for <THC/THCAtomics.cuh>: the reference kernels only need atomicAdd(float*/double*, .), which CUDA provides.
// TEST INFRASTRUCTURE (see THC.h next to this file).
#pragma once
#include <cuda_runtime.h>
