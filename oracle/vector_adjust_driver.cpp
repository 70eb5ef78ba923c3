// TEST INFRASTRUCTURE ONLY.  Host-side "launcher" for the reference kernel compiled through
// oracle/cuda_cpu_shim.h: replays the launch geometry of gpt_hessian_cuda.py:38-52
// (block=(256,1,1), grid=(ceil(n/256),1)) one thread at a time.
#include "cuda_cpu_shim.h"

thread_local hlv_shim_dim3 blockIdx, blockDim, threadIdx, gridDim;

extern "C" void vector_adjust(const float* grad_vector, const float* V, const float* eigvals,
                              float* adjusted_grad_vector, int num_eigenvalues, int vec_len, float delta);

extern "C" void ref_vector_adjust_cpu(const float* grad_vector, const float* V, const float* eigvals,
                                      float* adjusted_grad_vector, int num_eigenvalues, int vec_len,
                                      float delta, int block_size) {
    const unsigned grid = (unsigned)((vec_len + block_size - 1) / block_size);
    blockDim = {(unsigned)block_size, 1, 1};
    gridDim = {grid, 1, 1};
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < (unsigned)block_size; ++t) {
            blockIdx = {b, 0, 0};
            threadIdx = {t, 0, 0};
            vector_adjust(grad_vector, V, eigvals, adjusted_grad_vector, num_eigenvalues, vec_len, delta);
        }
}
