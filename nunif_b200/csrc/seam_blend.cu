// Tiling engine: integer planner, tile unfold (replicate pad fused), closed-form
// overlap blend as a gather.
//
// Replaces nunif/utils/seam_blending.py:
//   create_config :109-143  -> nb200_tile_config_create (host, integers, bit-exact)
//   F.pad replicate :82 + per-tile slice copies :83-92 -> tile_unfold_kernel (one launch per batch)
//   update :156-174 (5 elementwise passes per tile over two fp32 accumulators the size of
//   the output) + get_output :39-40 -> tile_gather_blend_kernel: every output pixel gathers
//   the <=4 tiles that cover it, out = clamp(sum w*z / sum w).  No accumulators exist at all;
//   traffic = read each z once (fp16) + write the output once.
#include "common.cuh"
#include "../../include/nunif_b200.h"

namespace nb200 {

// dst[n][T][T][cpad] fp16; 8 channels per 16-byte store when cpad == 8.
__global__ void __launch_bounds__(256) tile_unfold_kernel(const float* __restrict__ x, int C, int H, int W, int pad_l, int pad_t,
                                                          int w_blocks, int step, int T, int tile0, int n, int cpad,
                                                          __half* __restrict__ dst) {
    const size_t total = (size_t)n * T * T;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int px = (int)(i % T);
    const int py = (int)((i / T) % T);
    const int k = (int)(i / ((size_t)T * T));
    const int t = tile0 + k;
    const int hi = t / w_blocks, wi = t % w_blocks;
    // padded coordinate -> clamp-to-edge source coordinate (F.pad mode='replicate')
    const int sy = min(max(hi * step + py - pad_t, 0), H - 1);
    const int sx = min(max(wi * step + px - pad_l, 0), W - 1);
    __half* d = dst + i * cpad;
    if (cpad == 8) {
        __align__(16) __half v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __float2half_rn(c < C ? __ldg(x + ((size_t)c * H + sy) * W + sx) : 0.f);
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(v);
    } else {
        for (int c = 0; c < cpad; ++c) d[c] = __float2half_rn(c < C ? __ldg(x + ((size_t)c * H + sy) * W + sx) : 0.f);
    }
}

struct BlendParams {
    const void* z;        // tile outputs: fp16 (models that return fp16 under autocast) or fp32 (the 4x-derived 2x / 1x models)
    float* out;
    int C, S, y_h, y_w, h_blocks, w_blocks, step_out, blend, y0;
    float ring[65];  // ring[d] = weight at distance d from the tile edge, d < blend
};

__device__ __forceinline__ float blend_weight(const BlendParams& p, int u, int v) {
    // create_blend_filter :146-153: ones inside, rings of 1-(i+1)/(blend+1) growing outward
    int d = min(min(u, v), min(p.S - 1 - u, p.S - 1 - v));
    return d >= p.blend ? 1.f : p.ring[d];
}

template <typename ZT> __device__ __forceinline__ float z_at(const void* z, size_t i);
template <> __device__ __forceinline__ float z_at<__half>(const void* z, size_t i) { return __half2float(reinterpret_cast<const __half*>(z)[i]); }
template <> __device__ __forceinline__ float z_at<float>(const void* z, size_t i) { return __ldg(reinterpret_cast<const float*>(z) + i); }
template <typename ZT> __device__ __forceinline__ void z_at4(const void* z, size_t i, float (&v)[4]);
template <> __device__ __forceinline__ void z_at4<__half>(const void* z, size_t i, float (&v)[4]) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(z) + i));
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}
template <> __device__ __forceinline__ void z_at4<float>(const void* z, size_t i, float (&v)[4]) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(z) + i));
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}

template <typename ZT>
__global__ void __launch_bounds__(256) tile_gather_blend_kernel(BlendParams p) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y + p.y0;
    if (X >= p.y_w) return;
    // tiles covering Y: hi with hi*step <= Y < hi*step + S
    const int hi1 = min(Y / p.step_out, p.h_blocks - 1);
    const int wi1 = min(X / p.step_out, p.w_blocks - 1);
    int his[2], wis[2], nh = 0, nw = 0;
    if (hi1 > 0 && Y - (hi1 - 1) * p.step_out < p.S) his[nh++] = hi1 - 1;
    if (Y - hi1 * p.step_out < p.S) his[nh++] = hi1;
    if (wi1 > 0 && X - (wi1 - 1) * p.step_out < p.S) wis[nw++] = wi1 - 1;
    if (X - wi1 * p.step_out < p.S) wis[nw++] = wi1;
    const size_t zplane = (size_t)p.S * p.S;
    for (int c = 0; c < p.C; ++c) {
        float num = 0.f, den = 0.f;
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                const int u = Y - his[a] * p.step_out, v = X - wis[b] * p.step_out;
                const size_t t = (size_t)his[a] * p.w_blocks + wis[b];
                const float zv = z_at<ZT>(p.z, (t * p.C + c) * zplane + (size_t)u * p.S + v);
                if (p.blend > 0) {
                    const float w = blend_weight(p, u, v);
                    num += w * zv;
                    den += w;
                } else {  // plain store, last tile in raster order wins (:173)
                    num = zv;
                    den = 1.f;
                }
            }
        p.out[((size_t)c * p.y_h + Y) * p.y_w + X] = clamp01(num / den);
    }
}

// 4 horizontally adjacent output pixels per thread (valid when y_w, the tile step and S are multiples of 4, so a
// group never straddles a tile edge): 8-byte fp16 loads, one float4 store per colour plane.
template <typename ZT>
__global__ void __launch_bounds__(256) tile_gather_blend4_kernel(BlendParams p) {
    const int X = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int Y = blockIdx.y + p.y0;
    if (X >= p.y_w) return;
    const int hi1 = min(Y / p.step_out, p.h_blocks - 1);
    const int wi1 = min(X / p.step_out, p.w_blocks - 1);
    int his[2], wis[2], nh = 0, nw = 0;
    if (hi1 > 0 && Y - (hi1 - 1) * p.step_out < p.S) his[nh++] = hi1 - 1;
    if (Y - hi1 * p.step_out < p.S) his[nh++] = hi1;
    if (wi1 > 0 && X - (wi1 - 1) * p.step_out < p.S) wis[nw++] = wi1 - 1;
    if (X - wi1 * p.step_out < p.S) wis[nw++] = wi1;
    const size_t zplane = (size_t)p.S * p.S;
    for (int c = 0; c < p.C; ++c) {
        float num[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                const int u = Y - his[a] * p.step_out, v = X - wis[b] * p.step_out;
                const size_t t = (size_t)his[a] * p.w_blocks + wis[b];
                float zv[4];
                z_at4<ZT>(p.z, (t * p.C + c) * zplane + (size_t)u * p.S + v, zv);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (p.blend > 0) {
                        const float w = blend_weight(p, u, v + k);
                        num[k] += w * zv[k];
                        den[k] += w;
                    } else {
                        num[k] = zv[k];
                        den[k] = 1.f;
                    }
                }
            }
        *reinterpret_cast<float4*>(p.out + ((size_t)c * p.y_h + Y) * p.y_w + X) =
            make_float4(clamp01(num[0] / den[0]), clamp01(num[1] / den[1]), clamp01(num[2] / den[2]), clamp01(num[3] / den[3]));
    }
}

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_tile_config_create(int x_h, int x_w, int scale, int offset, int tile_size, int blend_size,
                                        nb200_tile_config* out) {
    NB_CHECK(out, "null pointer");
    NB_CHECK(x_h > 0 && x_w > 0 && scale > 0 && offset >= 0 && tile_size > 0 && blend_size >= 0, "bad argument");
    const int input_offset = (offset + scale - 1) / scale;       // math.ceil(offset / scale)
    const int input_blend = (blend_size + scale - 1) / scale;    // math.ceil(blend_size / scale)
    const int step = tile_size - (input_offset * 2 + input_blend);
    NB_CHECK(step > 0, "tile_size too small for this offset/blend");
    int h_blocks = 0, w_blocks = 0, input_h = 0, input_w = 0;
    while (input_h < x_h + input_offset * 2) { input_h = h_blocks * step + tile_size; ++h_blocks; }
    while (input_w < x_w + input_offset * 2) { input_w = w_blocks * step + tile_size; ++w_blocks; }
    out->y_h = x_h * scale;
    out->y_w = x_w * scale;
    out->h_blocks = h_blocks;
    out->w_blocks = w_blocks;
    out->pad_l = input_offset;
    out->pad_r = input_w - (x_w + input_offset);
    out->pad_t = input_offset;
    out->pad_b = input_h - (x_h + input_offset);
    out->y_buffer_h = input_h * scale;
    out->y_buffer_w = input_w * scale;
    out->input_tile_step = step;
    out->output_tile_step = step * scale;
    return 0;
}

extern "C" int nb200_tile_unfold(const float* x, int C, int H, int W, const nb200_tile_config* cfg, int tile_size,
                                 int tile0, int n, void* dst, int cpad, void* stream) {
    NB_CHECK(x && cfg && dst, "null pointer");
    NB_CHECK(C <= cpad, "cpad must be >= C");
    NB_CHECK(tile0 >= 0 && n > 0 && tile0 + n <= cfg->h_blocks * cfg->w_blocks, "tile range out of bounds");
    const size_t total = (size_t)n * tile_size * tile_size;
    ProfScope ps((cudaStream_t)stream, PC_UNFOLD, (double)total * (C * 4 + cpad * 2));
    tile_unfold_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, C, H, W, cfg->pad_l, cfg->pad_t, cfg->w_blocks, cfg->input_tile_step, tile_size, tile0, n, cpad, (__half*)dst);
    NB_LAUNCHED();
    return 0;
}

extern "C" int nb200_tile_gather_blend(const void* z_all, int C, const nb200_tile_config* cfg, int scale, int offset,
                                       int tile_size, int blend_size, float* out, void* stream) {
    NB_CHECK(cfg, "null pointer");
    return nb200::tile_gather_blend_rows(z_all, 0, C, cfg, scale, offset, tile_size, blend_size, out, 0, cfg->y_h, stream);
}

// rows [y0, y1) of the blended output; every tile covering those rows must already be in z_all
int nb200::tile_gather_blend_rows(const void* z_all, int z_f32, int C, const nb200_tile_config* cfg, int scale, int offset, int tile_size,
                                  int blend_size, float* out, int y0, int y1, void* stream) {
    NB_CHECK(z_all && cfg && out, "null pointer");
    NB_CHECK(0 <= y0 && y0 < y1 && y1 <= cfg->y_h, "bad row range");
    NB_CHECK(blend_size >= 0 && blend_size <= 64, "blend_size out of range");
    BlendParams p;
    p.z = z_all;
    p.out = out;
    p.C = C;
    p.S = tile_size * scale - 2 * offset;
    p.y_h = cfg->y_h; p.y_w = cfg->y_w; p.h_blocks = cfg->h_blocks; p.w_blocks = cfg->w_blocks;
    p.step_out = cfg->output_tile_step;
    p.blend = blend_size;
    p.y0 = y0;
    NB_CHECK(p.S > 0 && p.S - p.step_out <= p.step_out, "tile overlap larger than the tile step is not supported");
    for (int d = 0; d < blend_size; ++d) {
        // ring index i = blend-1-d (outermost ring is added last); value = 1 - (1/(blend+1))*(i+1)
        const int i = blend_size - 1 - d;
        p.ring[d] = (float)(1.0 - (1.0 / (blend_size + 1)) * (i + 1));
    }
    const double frac = (double)(y1 - y0) / p.y_h;
    ProfScope ps((cudaStream_t)stream, PC_BLEND,
                 frac * ((double)C * p.y_h * p.y_w * 4 + (double)p.h_blocks * p.w_blocks * C * p.S * p.S * 2));
    const bool vec4 = p.y_w % 4 == 0 && p.step_out % 4 == 0 && p.S % 4 == 0 && ((uintptr_t)out & 15) == 0;
    const dim3 g4(cdiv(p.y_w / 4, 256), y1 - y0), g1(cdiv(p.y_w, 256), y1 - y0);
    cudaStream_t st = (cudaStream_t)stream;
    if (z_f32) {
        if (vec4) tile_gather_blend4_kernel<float><<<g4, 256, 0, st>>>(p);
        else tile_gather_blend_kernel<float><<<g1, 256, 0, st>>>(p);
    } else {
        if (vec4) tile_gather_blend4_kernel<__half><<<g4, 256, 0, st>>>(p);
        else tile_gather_blend_kernel<__half><<<g1, 256, 0, st>>>(p);
    }
    NB_LAUNCHED();
    return 0;
}
