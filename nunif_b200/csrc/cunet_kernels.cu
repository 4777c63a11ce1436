// Non-GEMM kernels of the CUNet path (NHWC fp16 activations):
//   SE block (nunif/modules/attention.py:38-44): deterministic global average pool,
//   the two 1x1 convs + ReLU + sigmoid on the pooled vector, channel scaling in place;
//   the 3-channel tail convolutions (cunet.py:43 conv_bottom 3x3, :41 deconv 4x4 s2 p3),
//   fused with the cascade's clamp / crop-add (cunet.py:149-163).
#include "common.cuh"
#include "cunet_kernels.h"

namespace nb200 {

// ---- SE: pool ------------------------------------------------------------------------------
// partial[b][chunk][c] = sum over the chunk's pixels; chunks are fixed-size => deterministic.
constexpr int SE_CHUNK = 2048;  // pixels per partial

template <int C>
__global__ void __launch_bounds__(256) se_pool_partial_kernel(const __half* __restrict__ x, float* __restrict__ partial, int HW,
                                                              int nchunks) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = chunk * SE_CHUNK, p1 = min(p0 + SE_CHUNK, HW);
    constexpr int LANES = C / 8;             // threads covering one pixel (8 channels each)
    constexpr int ROWS = 256 / LANES;        // pixels processed per step
    const int lane = threadIdx.x % LANES, row = threadIdx.x / LANES;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __half* xb = x + (size_t)b * HW * C;
    for (int p = p0 + row; p < p1; p += ROWS) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * C + lane * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            acc[2 * j] += f.x;
            acc[2 * j + 1] += f.y;
        }
    }
    __shared__ float sm[ROWS][C];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[row][lane * 8 + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int r = 0; r < ROWS; ++r) s += sm[r][c];
        partial[((size_t)b * nchunks + chunk) * C + c] = s;
    }
}

// ---- SE: finalize pool + fc1 + relu + fc2 + sigmoid -> scale[b][c] ---------------------------
template <int C>
__global__ void se_fc_kernel(const float* __restrict__ partial, int nchunks, int HW, const float* __restrict__ w1,
                             const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                             float* __restrict__ scale) {
    constexpr int R = C / 8;
    __shared__ float mean[C];
    __shared__ float hid[R];
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nchunks; ++k) s += partial[((size_t)b * nchunks + k) * C + c];
        mean[c] = s / (float)HW;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float s = b1[r];
        for (int c = 0; c < C; ++c) s += mean[c] * w1[r * C + c];
        hid[r] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = b2[c];
        for (int r = 0; r < R; ++r) s += hid[r] * w2[c * R + r];
        scale[(size_t)b * C + c] = 1.f / (1.f + __expf(-s));
    }
}

// ---- SE: x *= scale[b][c] ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_scale_kernel(__half* __restrict__ x, const float* __restrict__ scale, int C, size_t HW,
                                                       size_t total_vec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_vec) return;
    const int vpp = C / 8;
    const size_t pix = i / vpp;
    const int c0 = (int)(i % vpp) * 8;
    const size_t b = pix / HW;
    uint4 v = reinterpret_cast<uint4*>(x)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
    const float* s = scale + b * C + c0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        h[j] = __floats2half2_rn(f.x * s[2 * j], f.y * s[2 * j + 1]);
    }
    reinterpret_cast<uint4*>(x)[i] = v;
}

int se_block(cudaStream_t st, __half* x, int n, int H, int W, int C, const float* w1, const float* b1, const float* w2,
             const float* b2, float* partial, float* scale) {
    const int HW = H * W, nchunks = cdiv(HW, SE_CHUNK);
    ProfScope ps(st, PC_SE, (double)n * HW * C * 2 * 3);
    if (C == 64) {
        se_pool_partial_kernel<64><<<dim3(nchunks, n), 256, 0, st>>>(x, partial, HW, nchunks);
        NB_LAUNCHED();
        se_fc_kernel<64><<<n, 64, 0, st>>>(partial, nchunks, HW, w1, b1, w2, b2, scale);
    } else if (C == 128) {
        se_pool_partial_kernel<128><<<dim3(nchunks, n), 256, 0, st>>>(x, partial, HW, nchunks);
        NB_LAUNCHED();
        se_fc_kernel<128><<<n, 128, 0, st>>>(partial, nchunks, HW, w1, b1, w2, b2, scale);
    } else {
        return fail("se_block: unsupported channel count");
    }
    NB_LAUNCHED();
    const size_t total_vec = (size_t)n * HW * (C / 8);
    se_scale_kernel<<<(unsigned)cdiv64(total_vec, 256), 256, 0, st>>>(x, scale, C, (size_t)HW, total_vec);
    NB_LAUNCHED();
    return 0;
}

size_t se_partial_floats(int n, int H, int W, int C) { return (size_t)n * cdiv(H * W, SE_CHUNK) * C; }

// ---- 3-channel tails --------------------------------------------------------------------------
// MODE 0: conv 3x3 valid 64->3 ; MODE 1: ConvTranspose 4x4 s2 p3 64->3
// EPI 0: out = NHWC8 fp16 (3 ch + 5 zeros), optional clamp(0,1)   (z1 feeding unet2, cunet.py:150-153)
// EPI 1: out = planar fp16 z [n][3][Ho][Wo] = clamp(acc + z1[crop 20])  (cunet.py:154-163)
template <int MODE, int EPI>
__global__ void __launch_bounds__(128) tail_conv_kernel(const __half* __restrict__ x, const float* __restrict__ wt,
                                                        const float* __restrict__ bias, __half* __restrict__ out,
                                                        const __half* __restrict__ z1, int n, int Hi, int Wi, int Ho, int Wo,
                                                        int z1H, int z1W, int clip) {
    constexpr int TAPS = MODE == 0 ? 9 : 16;
    __shared__ __align__(16) float sw[TAPS * 64 * 3];  // [tap][ci][co]
    for (int i = threadIdx.x; i < TAPS * 64 * 3; i += blockDim.x) sw[i] = wt[i];
    __syncthreads();
    const size_t total = (size_t)n * Ho * Wo;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), b = (int)(i / ((size_t)Wo * Ho));
    float acc[3] = {bias[0], bias[1], bias[2]};
    const __half* xb = x + (size_t)b * Hi * Wi * 64;
    if (MODE == 0) {
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            const __half* px = xb + ((size_t)(oy + t / 3) * Wi + (ox + t % 3)) * 64;
            const float* w = sw + t * 192;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(px) + c8);
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    const float* w0 = w + (c8 * 8 + 2 * j) * 3;
                    acc[0] += f.x * w0[0] + f.y * w0[3];
                    acc[1] += f.x * w0[1] + f.y * w0[4];
                    acc[2] += f.x * w0[2] + f.y * w0[5];
                }
            }
        }
    } else {
        // oy = 2*iy - 3 + ky  =>  ky parity = (oy+3)&1, iy = (oy + 3 - ky) / 2
        const int py = (oy + 3) & 1, pxp = (ox + 3) & 1;
#pragma unroll 1
        for (int a = 0; a < 2; ++a) {
            const int ky = py + 2 * a, iy = (oy + 3 - ky) >> 1;
            if (iy < 0 || iy >= Hi) continue;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int kx = pxp + 2 * c, ix = (ox + 3 - kx) >> 1;
                if (ix < 0 || ix >= Wi) continue;
                const __half* px = xb + ((size_t)iy * Wi + ix) * 64;
                const float* w = sw + (ky * 4 + kx) * 192;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(px) + c8);
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(h[j]);
                        const float* w0 = w + (c8 * 8 + 2 * j) * 3;
                        acc[0] += f.x * w0[0] + f.y * w0[3];
                        acc[1] += f.x * w0[1] + f.y * w0[4];
                        acc[2] += f.x * w0[2] + f.y * w0[5];
                    }
                }
            }
        }
    }
    if (EPI == 0) {
        __align__(16) __half o[8];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = acc[k];
            if (clip) v = clamp01(v);
            o[k] = __float2half_rn(v);
        }
#pragma unroll
        for (int k = 3; k < 8; ++k) o[k] = __float2half_rn(0.f);
        reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(o);
    } else {
        const __half* zp = z1 + (((size_t)b * z1H + oy + 20) * z1W + ox + 20) * 8;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = acc[k] + __half2float(zp[k]);
            out[(((size_t)b * 3 + k) * Ho + oy) * Wo + ox] = __float2half_rn(clamp01(v));
        }
    }
}

// Tensor-core version of the tails (production path): the 3 output channels are padded to the n=8 of
// mma.sync.m16n8k16; a warp owns 16 output pixels of one image row (for the ConvTranspose: 16 pixels of one column
// parity, so that all rows of the MMA share the same 2x2 subset of the 4x4 taps), A fragments are gathered straight
// from the NHWC fp16 activation (implicit im2col, 4-byte loads that hit L1), B fragments come from a shared-memory copy
// of the weights in fp16 (the reference runs these convs in fp16 under autocast as well).
// 16 instructions per output pixel instead of ~3500 for the SIMT kernel above (kept as the fallback / cross-check).
template <int MODE, int EPI>
__global__ void __launch_bounds__(128) tail_conv_mma_kernel(const __half* __restrict__ x, const float* __restrict__ wt,
                                                            const float* __restrict__ bias, __half* __restrict__ out,
                                                            const __half* __restrict__ z1, int Hi, int Wi, int Ho, int Wo,
                                                            int z1H, int z1W, int clip) {
    constexpr int TAPS = MODE == 0 ? 9 : 16;
    // sB[tap][kc (4 chunks of 16 channels)][n (8)][16 halves]; rows n >= 3 are zero
    __shared__ __align__(16) __half sB[TAPS * 4 * 8 * 16];
    for (int i = threadIdx.x; i < TAPS * 4 * 8 * 16; i += blockDim.x) {
        const int k = i & 15, nn = (i >> 4) & 7, kc = (i >> 7) & 3, t = i >> 9;
        sB[i] = __float2half_rn(nn < 3 ? wt[(t * 64 + kc * 16 + k) * 3 + nn] : 0.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int oy = blockIdx.y % Ho, b = blockIdx.y / Ho;
    const int par = MODE == 1 ? blockIdx.z : 0;                  // column parity class of the ConvTranspose
    const int j0 = (blockIdx.x * 4 + warp) * 16;                 // first pixel (index within the parity class) of this warp
    const int npx = MODE == 1 ? Wo / 2 : Wo;
    if (j0 >= npx) return;
    const int jr[2] = {j0 + g, j0 + g + 8};                      // the two accumulator rows of this thread
    int ox[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) ox[r] = MODE == 1 ? 2 * min(jr[r], npx - 1) + par : min(jr[r], npx - 1);
    const __half* xb = x + (size_t)b * Hi * Wi * 64;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int NT = MODE == 0 ? 9 : 4;
#pragma unroll 1
    for (int tt = 0; tt < NT; ++tt) {
        int tap, iy, ix[2];
        bool ok[2];
        if (MODE == 0) {
            tap = tt;
            iy = oy + tt / 3;
            ix[0] = ox[0] + tt % 3; ix[1] = ox[1] + tt % 3;
            ok[0] = ok[1] = true;
        } else {
            // oy = 2*iy - 3 + ky  =>  ky has the parity of oy+3; same for x (cunet.py:41 ConvTranspose2d(64, 3, 4, 2, 3))
            const int ky = ((oy + 3) & 1) + 2 * (tt >> 1), kx = ((par + 3) & 1) + 2 * (tt & 1);
            tap = ky * 4 + kx;
            iy = (oy + 3 - ky) >> 1;
            const bool yok = iy >= 0 && iy < Hi;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                ix[r] = (ox[r] + 3 - kx) >> 1;
                ok[r] = yok && ix[r] >= 0 && ix[r] < Wi;
            }
            iy = min(max(iy, 0), Hi - 1);
        }
        const __half* p0 = xb + ((size_t)iy * Wi + min(max(ix[0], 0), Wi - 1)) * 64 + 2 * t4;
        const __half* p1 = xb + ((size_t)iy * Wi + min(max(ix[1], 0), Wi - 1)) * 64 + 2 * t4;
        const __half* wb = sB + (size_t)tap * 512 + g * 16 + 2 * t4;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t a[4];
            a[0] = ok[0] ? __ldg(reinterpret_cast<const uint32_t*>(p0 + kc * 16)) : 0u;
            a[1] = ok[1] ? __ldg(reinterpret_cast<const uint32_t*>(p1 + kc * 16)) : 0u;
            a[2] = ok[0] ? __ldg(reinterpret_cast<const uint32_t*>(p0 + kc * 16 + 8)) : 0u;
            a[3] = ok[1] ? __ldg(reinterpret_cast<const uint32_t*>(p1 + kc * 16 + 8)) : 0u;
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wb + kc * 128);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wb + kc * 128 + 8);
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }
    }
    // accumulator layout: acc[0..1] = row g, channels 2*t4, 2*t4+1; acc[2..3] = row g+8.  Channels 0,1 live in t4 == 0,
    // channel 2 in t4 == 1; gather the three into the t4 == 0 lane of each row.
    const float c2a = __shfl_down_sync(0xffffffffu, acc[0], 1), c2b = __shfl_down_sync(0xffffffffu, acc[2], 1);
    if (t4 != 0) return;
    const float bv[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (jr[r] >= npx) continue;
        float v[3] = {(r ? acc[2] : acc[0]) + bv[0], (r ? acc[3] : acc[1]) + bv[1], (r ? c2b : c2a) + bv[2]};
        const size_t pix = ((size_t)b * Ho + oy) * Wo + ox[r];
        if (EPI == 0) {
            __align__(16) __half o[8];
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = __float2half_rn(clip ? clamp01(v[k]) : v[k]);
#pragma unroll
            for (int k = 3; k < 8; ++k) o[k] = __float2half_rn(0.f);
            reinterpret_cast<uint4*>(out)[pix] = *reinterpret_cast<const uint4*>(o);
        } else {
            const __half* zp = z1 + (((size_t)b * z1H + oy + 20) * z1W + ox[r] + 20) * 8;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                out[(((size_t)b * 3 + k) * Ho + oy) * Wo + ox[r]] = __float2half_rn(clamp01(v[k] + __half2float(zp[k])));
        }
    }
}

extern int g_tune[8];  // gemm.cu; [7] != 0 selects the SIMT tail kernel (tests)

int tail_conv(cudaStream_t st, int mode, int epi, const __half* x, const float* wt, const float* bias, __half* out,
              const __half* z1, int n, int Hi, int Wi, int z1H, int z1W, int clip) {
    const int Ho = mode == 0 ? Hi - 2 : 2 * Hi - 4, Wo = mode == 0 ? Wi - 2 : 2 * Wi - 4;
    const size_t total = (size_t)n * Ho * Wo;
    const unsigned blocks = (unsigned)cdiv64(total, 128);
    ProfScope ps(st, PC_TAIL, (double)n * Hi * Wi * 128 + (double)total * 16);
    if (g_tune[7] == 0 && (mode == 0 || Wo % 2 == 0)) {
        const int npx = mode == 1 ? Wo / 2 : Wo;
        const dim3 grid(cdiv(npx, 64), (unsigned)(n * Ho), mode == 1 ? 2 : 1);
        if (mode == 0 && epi == 0) tail_conv_mma_kernel<0, 0><<<grid, 128, 0, st>>>(x, wt, bias, out, z1, Hi, Wi, Ho, Wo, z1H, z1W, clip);
        else if (mode == 0 && epi == 1) tail_conv_mma_kernel<0, 1><<<grid, 128, 0, st>>>(x, wt, bias, out, z1, Hi, Wi, Ho, Wo, z1H, z1W, clip);
        else if (mode == 1 && epi == 0) tail_conv_mma_kernel<1, 0><<<grid, 128, 0, st>>>(x, wt, bias, out, z1, Hi, Wi, Ho, Wo, z1H, z1W, clip);
        else return fail("tail_conv: unsupported mode");
        NB_LAUNCHED();
        return 0;
    }
    if (mode == 0 && epi == 0) tail_conv_kernel<0, 0><<<blocks, 128, 0, st>>>(x, wt, bias, out, z1, n, Hi, Wi, Ho, Wo, z1H, z1W, clip);
    else if (mode == 0 && epi == 1) tail_conv_kernel<0, 1><<<blocks, 128, 0, st>>>(x, wt, bias, out, z1, n, Hi, Wi, Ho, Wo, z1H, z1W, clip);
    else if (mode == 1 && epi == 0) tail_conv_kernel<1, 0><<<blocks, 128, 0, st>>>(x, wt, bias, out, z1, n, Hi, Wi, Ho, Wo, z1H, z1W, clip);
    else return fail("tail_conv: unsupported mode");
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb200
