/* TEST INFRASTRUCTURE ONLY.  Lets g++ compile the reference's bare CUDA kernel
 * (/root/reference/vector_adjust.cu) as plain C++ so the reference's own source can be
 * executed on host cores: __global__ vanishes and the built-in index variables become
 * thread-local globals that oracle/vector_adjust_driver.cpp sweeps over the launch grid. */
#pragma once
struct hlv_shim_dim3 { unsigned x, y, z; };
extern thread_local hlv_shim_dim3 blockIdx, blockDim, threadIdx, gridDim;
#define __global__
