#pragma once
#include "common.cuh"

namespace nb200 {
int stem_conv3x3(cudaStream_t st, const __half* x, const float* wt, const float* bias, __half* out, int n, int Hi, int Wi,
                 int cout_pad, int ldo);
int window_attention(cudaStream_t st, const __half* qkv, const float* bias_table, __half* out, int B, int H, int W, int C,
                     int shift);
int to_image(cudaStream_t st, const __half* y, __half* z, int n, int Hs, int Ws, int cs, int r, int down);
}  // namespace nb200
