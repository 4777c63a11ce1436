#pragma once
#include "common.cuh"

namespace nb200 {
int stem_conv3x3(cudaStream_t st, const __half* x, const float* wt, const float* bias, __half* out, int n, int Hi, int Wi,
                 int cout_pad, int ldo);
constexpr int BIAS_FRAG_FLOATS = 6 * 3 * 6 * 32 * 4;  // per Swin block, see swin_attention_mma.cu
int build_bias_frag(cudaStream_t st, const float* table_121x6, float* frag);
// qkv: three dense planes q | k | v, each [B][H][W][C], `plane` elements apart
int window_attention(cudaStream_t st, const __half* qkv, const float* bias_frag, __half* out, int B, int H, int W, int C,
                     int shift, size_t plane);
// z: fp16 [n][3][S][S] for down == 1, fp32 for down in {2, 4}
int to_image(cudaStream_t st, const __half* y, void* z, int n, int Hs, int Ws, int cs, int r, int down);
}  // namespace nb200
