#pragma once
#include "common.cuh"

namespace nb200 {
int se_block(cudaStream_t st, __half* x, int n, int H, int W, int C, const float* w1, const float* b1, const float* w2,
             const float* b2, float* partial, float* scale);
size_t se_partial_floats(int n, int H, int W, int C);
int tail_conv(cudaStream_t st, int mode, int epi, const __half* x, const float* wt, const float* bias, __half* out,
              const __half* z1, int n, int Hi, int Wi, int z1H, int z1W, int clip);
}  // namespace nb200
