// Alpha-aware pre-processing and test-time augmentation around the tiled render
// (the remaining device work of Waifu2x.convert, waifu2x/utils.py:255-297):
//   * AlphaBorderPadding (nunif/utils/alpha.py:32-57): `offset` rounds of "fill still-transparent pixels with the
//     3x3 box average of the already-filled ones, grow the mask by one pixel".  One fused pass per round
//     (mask box-sum + rgb box-sum + divide + select + next mask) instead of 2 depthwise convs + 5 elementwise ops.
//   * tta_split / tta_merge (nunif/transforms/tta.py:20-48): the 8 dihedral views of an image and the average of
//     the 8 inverse-transformed results; merge is a single gather pass (sum in the reference's order, * 1/8, clamp).
// HBM-bound elementwise work; algorithmic bytes per round of the padding = 8 planes (4 in, 4 out) * 4 B.
#include "common.cuh"
#include "../../include/nunif_b200.h"

namespace nb200 {

__global__ void __launch_bounds__(256) alpha_init_kernel(const float* __restrict__ rgb, const float* __restrict__ alpha,
                                                          float* __restrict__ rgb_out, float* __restrict__ mask, size_t plane) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    const float m = alpha[i] > 0.f ? 1.f : 0.f;          // alpha.py:43-44
    mask[i] = m;
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb_out[c * plane + i] = m < 1.f ? 0.f : rgb[c * plane + i];   // :45-46
}

// one round (alpha.py:47-54).  Zero padding: out-of-image taps contribute 0 to both sums.
__global__ void __launch_bounds__(256) alpha_round_kernel(const float* __restrict__ rgb, const float* __restrict__ mask,
                                                           float* __restrict__ rgb_out, float* __restrict__ mask_out,
                                                           int H, int W, int clamp_out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t plane = (size_t)H * W, i = (size_t)y * W + x;
    float mw = 0.f, s[3] = {0.f, 0.f, 0.f};
    const float m0 = mask[i];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            const size_t j = (size_t)yy * W + xx;
            mw += __ldg(mask + j);
            if (m0 < 1.f) {
#pragma unroll
                for (int c = 0; c < 3; ++c) s[c] += __ldg(rgb + c * plane + j);
            }
        }
    }
    const float den = __fadd_rn(mw, 1e-7f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = m0 < 1.f ? __fdiv_rn(s[c], den) : rgb[c * plane + i];
        if (clamp_out) v = clamp01(v);
        rgb_out[c * plane + i] = v;
    }
    mask_out[i] = mw > 0.f ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) clamp_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = clamp01(src[i]);
}

// source coordinate (sy, sx) in an H x W image for element (i, j) of dihedral view k (tta.py:24-33).
// views 0-3 keep the H x W shape, views 4-7 are W x H.
__device__ __forceinline__ void tta_src(int k, int i, int j, int H, int W, int& sy, int& sx) {
    switch (k) {
        case 0: sy = i; sx = j; break;
        case 1: sy = i; sx = W - 1 - j; break;                 // hflip
        case 2: sy = H - 1 - i; sx = j; break;                 // vflip
        case 3: sy = H - 1 - i; sx = W - 1 - j; break;         // vflip, hflip
        case 4: sy = j; sx = W - 1 - i; break;                 // rot90(1, (1, 2))
        case 5: sy = H - 1 - j; sx = W - 1 - i; break;         // rot90, hflip
        case 6: sy = j; sx = i; break;                         // rot90, vflip
        default: sy = H - 1 - j; sx = i; break;                // rot90, vflip, hflip
    }
}

__global__ void __launch_bounds__(256) tta_transform_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H,
                                                             int W, int k) {
    const int oh = k < 4 ? H : W, ow = k < 4 ? W : H;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ow) return;
    int sy, sx;
    tta_src(k, i, j, H, W, sy, sx);
    for (int c = 0; c < C; ++c) out[((size_t)c * oh + i) * ow + j] = __ldg(x + ((size_t)c * H + sy) * W + sx);
}

struct TtaMergeParams {
    const float* z[8];
    float* out;
    int C, H, W;
};

// out[c][y][x] = clamp(1/8 * sum_k z_k[c][pos_k(y, x)]), summed in the reference's order (tta.py:36-48).
// Element (y, x) of the inverse-transformed view k is element (i, j) of z_k with tta_src(k, i, j) == (y, x):
__global__ void __launch_bounds__(256) tta_merge_kernel(TtaMergeParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.W) return;
    const int H = p.H, W = p.W;
    for (int c = 0; c < p.C; ++c) {
        const size_t cp = (size_t)c * H * W;
        float a = __ldg(p.z[0] + cp + (size_t)y * W + x);
        a += __ldg(p.z[1] + cp + (size_t)y * W + (W - 1 - x));
        a += __ldg(p.z[2] + cp + (size_t)(H - 1 - y) * W + x);
        a += __ldg(p.z[3] + cp + (size_t)(H - 1 - y) * W + (W - 1 - x));
        // views 4-7 are stored W x H (row length H)
        a += __ldg(p.z[4] + cp + (size_t)(W - 1 - x) * H + y);
        a += __ldg(p.z[5] + cp + (size_t)(W - 1 - x) * H + (H - 1 - y));
        a += __ldg(p.z[6] + cp + (size_t)x * H + y);
        a += __ldg(p.z[7] + cp + (size_t)x * H + (H - 1 - y));
        p.out[cp + (size_t)y * W + x] = clamp01(a * 0.125f);
    }
}

}  // namespace nb200

using namespace nb200;

extern "C" size_t nb200_alpha_border_padding_workspace(int H, int W) {
    return (size_t)H * W * sizeof(float) * (3 + 2);   // one rgb ping buffer + two mask planes
}

extern "C" int nb200_alpha_border_padding(const float* rgb, const float* alpha, int H, int W, int offset, float* out,
                                          void* workspace, void* stream) {
    NB_CHECK(rgb && alpha && out && workspace, "null pointer");
    NB_CHECK(H > 0 && W > 0 && offset >= 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)H * W;
    float* tmp = reinterpret_cast<float*>(workspace);
    float* mask_a = tmp + 3 * plane;
    float* mask_b = mask_a + plane;
    // ping-pong so that the last round writes `out`: round r reads buf[r & 1 ^ start]
    float* bufs[2] = {out, tmp};
    int cur = offset % 2 == 0 ? 0 : 1;   // the init pass writes bufs[cur]; after `offset` swaps the result is in bufs[0]
    ProfScope ps(st, PC_OTHER, (double)plane * 4 * 8 * (offset + 1));
    alpha_init_kernel<<<(unsigned)cdiv64((int64_t)plane, 256), 256, 0, st>>>(rgb, alpha, bufs[cur], mask_a, plane);
    NB_LAUNCHED();
    float *mc = mask_a, *mn = mask_b;
    for (int r = 0; r < offset; ++r) {
        alpha_round_kernel<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(bufs[cur], mc, bufs[cur ^ 1], mn, H, W, r == offset - 1);
        NB_LAUNCHED();
        cur ^= 1;
        float* t = mc; mc = mn; mn = t;
    }
    if (offset == 0) {   // only the clamp of alpha.py:56 remains (in place)
        clamp_copy_kernel<<<(unsigned)cdiv64((int64_t)plane * 3, 256), 256, 0, st>>>(out, out, plane * 3);
        NB_LAUNCHED();
    }
    return 0;
}

extern "C" int nb200_tta_transform(const float* x, int C, int H, int W, int k, float* out, void* stream) {
    NB_CHECK(x && out, "null pointer");
    NB_CHECK(C > 0 && H > 0 && W > 0 && k >= 0 && k < 8, "bad argument");
    const int oh = k < 4 ? H : W, ow = k < 4 ? W : H;
    tta_transform_kernel<<<dim3(cdiv(ow, 256), oh), 256, 0, (cudaStream_t)stream>>>(x, out, C, H, W, k);
    NB_LAUNCHED();
    return 0;
}

extern "C" int nb200_tta_merge(const float* const* views, int C, int H, int W, float* out, void* stream) {
    NB_CHECK(views && out, "null pointer");
    NB_CHECK(C > 0 && H > 0 && W > 0, "bad argument");
    TtaMergeParams p;
    for (int k = 0; k < 8; ++k) {
        NB_CHECK(views[k], "null view");
        p.z[k] = views[k];
    }
    p.out = out; p.C = C; p.H = H; p.W = W;
    ProfScope ps((cudaStream_t)stream, PC_OTHER, (double)C * H * W * 4 * 9);
    tta_merge_kernel<<<dim3(cdiv(W, 256), H), 256, 0, (cudaStream_t)stream>>>(p);
    NB_LAUNCHED();
    return 0;
}
