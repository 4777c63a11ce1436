// Fused Swin-block kernels (swin_fused_mlp.cu, swin_fused_attn.cu): two launches per SwinTransformerBlock.
#pragma once
#include "common.cuh"

namespace nb200 {

// x <- x1 + fc2(gelu(fc1(x1))),  x1 = x + att.Wp^T + bp  (att == nullptr: x1 = x).  x is [T][C] fp16, updated in place.
struct FusedMlp {
    __half* x = nullptr;
    const __half* att = nullptr;
    long long T = 0;
    int C = 0;
    const __half* wp = nullptr;   // [C][C]
    const float* bp = nullptr;
    const __half* w1 = nullptr;   // [2C][C]
    const float* b1 = nullptr;
    const __half* w2 = nullptr;   // [C][2C]
    const float* b2 = nullptr;
    // chunk-major copies [K/BK][rows][BK] (BK = 64 for C = 192, 32 for C = 96) for the half-SM kernel (swin_fused_mlp2.cu)
    const __half* w1_cm = nullptr;
    const __half* wp_cm = nullptr;
};
int swin_mlp_fused(cudaStream_t st, const FusedMlp& f);
// two CTAs per SM; C = 96: proj fused (att, wp_cm, bp), C = 192: MLP only (att must be null; the proj Linear runs as a GEMM)
int swin_mlp_fused2(cudaStream_t st, const FusedMlp& f);
int pack_chunk_major(cudaStream_t st, const __half* w, __half* out, int rows, int K, int bk);

// att <- window_attention(x.Wqkv^T + bqkv)  (everything of shifted_window_attention but the proj Linear).
// x, att: [B][H][W][C] fp16; wqkv_packed: rows regrouped per head pair (pack_qkv_for_fused); bias_tab: [6][36][40] fp32
// (log2e * relative position bias, columns 36..39 = -1e30).
struct FusedAttn {
    const __half* x = nullptr;
    __half* att = nullptr;
    int B = 0, H = 0, W = 0, C = 0, shift = 0;
    const __half* wqkv = nullptr;   // [3C][C], row order (pair, {q,k,v}, head-in-pair, d)
    const float* bqkv = nullptr;    // [3C], same order
    const float* bias_tab = nullptr;
    // swin_attn_tc.cu: rows ordered (unit, {q,k,v}, head-in-unit, d) with 96 rows per unit (swin_attn_tc_src_row); the bias
    // is the raw relative_position_bias_table [121][6] fp32
    const __half* wqkv_tc = nullptr;
    const float* bqkv_tc = nullptr;
    const float* bias_tab_tc = nullptr;
};
int swin_attn_fused(cudaStream_t st, const FusedAttn& f);
// the same operator with QK^T and PV on tcgen05 as well (swin_attn_tc.cu); uses the *_tc operands
int swin_attn_tc(cudaStream_t st, const FusedAttn& f);
// packed row pr of the tc layout -> row of the reference qkv weight ([q | k | v], head-major inside each)
__host__ __device__ inline int swin_attn_tc_src_row(int pr, int C) {
    const int D = C / 6, NHU = 32 / D;
    const int u = pr / 96, rem = pr % 96;
    const int m = rem / 32, hh = (rem % 32) / D, d = rem % D;
    return m * C + (NHU * u + hh) * D + d;
}

}  // namespace nb200
