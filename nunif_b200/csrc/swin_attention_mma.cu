// Shifted-window attention core on tensor cores (mma.sync m16n8k16, fp16 in / fp32 accumulate).
//
// One CTA per 6x6 window, one warp per head (6 warps), 4 CTAs per SM (6 for d = 16).  q/k/v rows are staged with
// cp.async (16-byte LDGSTS, no register round trip).  The 36 tokens are padded to 48 MMA rows by clamping
// the row index (padded keys are masked by select, padded probabilities are exactly 0):
//   S = (Q*scale) K^T : M=48 (3 m16 tiles), N=48 (6 n8 tiles), K=d (d/16 steps)
//   softmax on the accumulator fragments (quad shuffles), relative-position bias and the
//   shift mask added per element, padded keys masked out
//   O = P V           : the S fragments are re-packed in registers as the A operand (no smem trip),
//                       V fragments come from ldmatrix.trans
// torchvision swin_transformer.py:166-221 (roll, partition, bias :190, mask :193-209, softmax :211,
// attn@v :214, un-roll) - everything but the qkv / proj Linears, which run on the tcgen05 GEMM.
//
// The 36x36 problem per head is far too small for a tcgen05 tile (M=128), and the op moves
// 8*C bytes/token for ~144*C FLOP/token: it is HBM/L2-bound, so the warp-level HMMA path is the
// right instrument here.
#include "common.cuh"
#include "swin_kernels.h"
#include <map>
#include <mutex>

namespace nb200 {

extern int g_tune[16];  // gemm.cu (nb200_tune_set)

namespace {
constexpr int WS = 6, WTOK = 36, WPAD = 48, HEADS = 6;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
}  // namespace

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

constexpr int NKT = 5;   // key tiles of 8 columns covering the 36 keys (columns 36..39 are masked by the bias table)

__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x1_trans(uint32_t& r0, const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}

template <int D>
struct AttnCtx {
    __half* sq;
    const int* sreg;
    const float4* bf;
    const __half* kbase[NKT];
    const __half* vbase[3];
    int hc, g, t4;
    float scale;
    bool boundary;
};

// One 16-row query tile of one head: S = Q K^T, bias/mask, base-2 softmax numerators, O = P V, normalise, stage.
// LAST: query rows 32..35 only (accumulator rows g+8 are padding and are not evaluated).
template <int D, bool LAST>
__device__ __forceinline__ void attn_mtile(const AttnCtx<D>& cx, int mt) {
    constexpr int C = D * HEADS, LD = C + 8;
    constexpr uint32_t ONES = 0x3C003C00u;   // half2(1, 1)
    constexpr int HL = LAST ? 1 : 2;         // accumulator row halves in use
    const int g = cx.g, t4 = cx.t4, hc = cx.hc;
    const int row0 = min(mt * 16 + g, WTOK - 1), row1 = min(mt * 16 + g + 8, WTOK - 1);
    // ---- S = Q K^T
    float s[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) s[nt][r] = 0.f;
#pragma unroll
    for (int kt = 0; kt < D / 16; ++kt) {
        uint32_t a[4];
        const __half* p0 = cx.sq + row0 * LD + hc + kt * 16 + 2 * t4;
        const __half* p1 = cx.sq + row1 * LD + hc + kt * 16 + 2 * t4;
        a[0] = *reinterpret_cast<const uint32_t*>(p0);
        a[1] = *reinterpret_cast<const uint32_t*>(p1);
        a[2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
            const __half* pk = cx.kbase[nt] + kt * 16;
            mma16816(s[nt], a, *reinterpret_cast<const uint32_t*>(pk), *reinterpret_cast<const uint32_t*>(pk + 8));
        }
    }
    // ---- scale + bias (+ mask).  element r of tile nt: row = mt*16 + g + 8*(r>>1), col = nt*8 + 2*t4 + (r&1).
    // bias_frag holds log2(e)*relative_position_bias in exactly this fragment order (padded keys = -1e30), so
    // scale + bias + key padding is one FMA per element and the softmax runs in base 2.
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
        const float4 bv = __ldg(cx.bf + (mt * 6 + nt) * 32);
        s[nt][0] = fmaf(s[nt][0], cx.scale, bv.x);
        s[nt][1] = fmaf(s[nt][1], cx.scale, bv.y);
        if (!LAST) {
            s[nt][2] = fmaf(s[nt][2], cx.scale, bv.z);
            s[nt][3] = fmaf(s[nt][3], cx.scale, bv.w);
        }
    }
    if (cx.boundary) {  // -100 across regions (:193-209)
        const int q0 = cx.sreg[row0], q1 = cx.sreg[row1];
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cr = cx.sreg[min(nt * 8 + 2 * t4 + e, WTOK - 1)];
                if (cr != q0) s[nt][e] += -100.0f * 1.4426950408889634f;
                if (!LAST && cr != q1) s[nt][2 + e] += -100.0f * 1.4426950408889634f;
            }
    }
    // ---- softmax numerators (base 2).  The row sums come out of the P V product itself (a ones column appended to V),
    // i.e. they are the sums of exactly the fp16 probabilities that multiply V; 1/sum is applied to the output rows.
#pragma unroll
    for (int hlf = 0; hlf < HL; ++hlf) {
        float mx = -1e30f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * hlf], s[nt][2 * hlf + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float pe;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(s[nt][2 * hlf + e] - mx));
                s[nt][2 * hlf + e] = pe;
            }
    }
    if (LAST) {
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) s[nt][2] = s[nt][3] = 0.f;
    }
    // ---- O = P V  (P in fp16, un-normalised: values in [0, 1]); osum = P 1
    float o[D / 8][4], osum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[nt][r] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
        uint32_t a[4];
        a[0] = pack_half2(s[2 * kt][0], s[2 * kt][1]);
        a[1] = pack_half2(s[2 * kt][2], s[2 * kt][3]);
        a[2] = pack_half2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        a[3] = pack_half2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
        mma16816(osum, a, ONES, ONES);
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            uint32_t b0, b1;
            ldmatrix_x2_trans(b0, b1, cx.vbase[kt] + nt * 8);
            mma16816(o[nt], a, b0, b1);
        }
    }
    {
        const uint32_t a0 = pack_half2(s[4][0], s[4][1]), a1 = pack_half2(s[4][2], s[4][3]);
        mma1688(osum, a0, a1, ONES);
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            uint32_t b0;
            ldmatrix_x1_trans(b0, cx.vbase[2] + nt * 8);
            mma1688(o[nt], a0, a1, b0);
        }
    }
    const float inv0 = __fdividef(1.f, osum[0]);
    const float inv1 = LAST ? 0.f : __fdividef(1.f, osum[2]);
    // ---- stage this head's output columns into sq.  Each warp only ever reads and writes its own columns of
    // rows [16*mt, 16*mt+16) here, and those q rows are dead once this m-tile's S is done.
    __syncwarp();
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        if (!LAST || r0 < WTOK) *reinterpret_cast<uint32_t*>(cx.sq + r0 * LD + hc + nt * 8 + 2 * t4) = pack_half2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (!LAST) *reinterpret_cast<uint32_t*>(cx.sq + r1 * LD + hc + nt * 8 + 2 * t4) = pack_half2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

template <int D>
__global__ void __launch_bounds__(192, D == 16 ? 6 : 4) window_attention_mma_kernel(const __half* __restrict__ qkv, const float4* __restrict__ bias_frag,
                                                                   __half* __restrict__ out, int H, int W, int shift,
                                                                   size_t plane) {
    if (threadIdx.x == 0) NB_PDL_TRIGGER();
    constexpr int C = D * HEADS;
    constexpr int LD = C + 8;  // padded row: (C+8)*2 bytes = 4 words mod 32 banks -> conflict-free fragment loads
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sq = reinterpret_cast<__half*>(smem_raw);  // [WTOK][LD]  q, later the output tile
    __half* sk = sq + WTOK * LD;
    __half* sv = sk + WTOK * LD;
    int* stok = reinterpret_cast<int*>(sv + WTOK * LD);      // [WTOK]
    int* sreg = stok + WTOK;                                 // [WPAD]
    const int nww = W / WS;
    const int wx = blockIdx.x % nww, wy = blockIdx.x / nww, b = blockIdx.y;
    const int tid = threadIdx.x;
    if (tid < WPAD) {
        int reg = -1;
        if (tid < WTOK) {
            const int ry = wy * WS + tid / WS, rx = wx * WS + tid % WS;  // rolled coordinates
            const int y = (ry + shift) % H, x = (rx + shift) % W;        // torch.roll(-shift) :166-167
            stok[tid] = (b * H + y) * W + x;
            int hr = 0, wr = 0;
            if (shift > 0) {
                hr = ry < H - WS ? 0 : (ry < H - shift ? 1 : 2);
                wr = rx < W - WS ? 0 : (rx < W - shift ? 1 : 2);
            }
            reg = hr * 3 + wr;
        }
        sreg[tid] = reg;
    }
    __syncthreads();
    // ---- stage q, k, v: three dense [T][C] planes (the qkv GEMM writes them split, OUT_SPLIT), 16-byte cp.async each.
    // thread -> (16-byte column vv, row group rg); it walks tokens rg, rg+RG, ... of all three planes
    constexpr int VPT = C / 8;         // 16-byte columns per token row: 24 (C=192) or 12 (C=96)
    constexpr int RG = 192 / VPT;      // 8 or 16 row groups
    const int vv = tid % VPT, rg = tid / VPT;
    for (int t = rg; t < WTOK; t += RG) {
        const __half* src = qkv + (size_t)stok[t] * C + vv * 8;
        __half* dst = sq + t * LD + vv * 8;
        cp_async16(dst, src);
        cp_async16(dst + WTOK * LD, src + plane);
        cp_async16(dst + 2 * WTOK * LD, src + 2 * plane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int head = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int hc = head * D;
    const float scale = ((D == 16) ? 0.25f : 0.17677669529663687f) * 1.4426950408889634f;  // (C//heads)**-0.5 (:187) * log2(e)
    const bool boundary = shift > 0 && (wy == gridDim.x / nww - 1 || wx == nww - 1);  // only these windows mix mask regions
    const float4* bf = bias_frag + (size_t)head * (3 * 6 * 32) + lane;
    AttnCtx<D> cx;
    cx.sq = sq; cx.sreg = sreg; cx.bf = bf; cx.hc = hc; cx.g = g; cx.t4 = t4; cx.scale = scale; cx.boundary = boundary;
    // per-thread fragment base pointers (row clamps and column offsets resolved once).  36 keys = 4.5 n8 tiles: S uses
    // 5 key tiles (40 columns), P V uses two k16 steps and one k8 step.
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) cx.kbase[nt] = sk + min(nt * 8 + g, WTOK - 1) * LD + hc + 2 * t4;
    cx.vbase[0] = sv + (lane & 15) * LD + hc;
    cx.vbase[1] = sv + (16 + (lane & 15)) * LD + hc;
    cx.vbase[2] = sv + min(32 + (lane & 7), WTOK - 1) * LD + hc;
    // One 16-row m-tile at a time (rows 0-15, 16-31, 32-47): keeps the live state at 20 + 4*D/8 (+4) accumulators so that
    // 4 CTAs fit per SM; K / V fragments are re-read from shared memory per m-tile (cheap, conflict-free).
    // The last tile holds query rows 32..35 only: its upper half (rows 40..47) is skipped.
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) attn_mtile<D, false>(cx, mt);
    attn_mtile<D, true>(cx, 2);
    __syncthreads();
    for (int t = rg; t < WTOK; t += RG)
        *reinterpret_cast<uint4*>(out + (size_t)stok[t] * C + vv * 8) = *reinterpret_cast<const uint4*>(sq + t * LD + vv * 8);
}

// bias_frag[head][mt][nt][lane] (float4 = accumulator fragment order) = log2(e) * table[rel_index(row, col)][head];
// padded key columns (>= 36) hold -1e30 so they vanish in the softmax; padded query rows hold 0.
__global__ void build_bias_frag_kernel(const float* __restrict__ table, float4* __restrict__ frag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // ((head*3 + mt)*6 + nt)*32 + lane
    if (i >= HEADS * 3 * 6 * 32) return;
    const int lane = i & 31, nt = (i >> 5) % 6, mt = (i / (32 * 6)) % 3, head = i / (32 * 6 * 3);
    const int g = lane >> 2, t4 = lane & 3;
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = mt * 16 + g + 8 * (r >> 1), col = nt * 8 + 2 * t4 + (r & 1);
        if (col >= WTOK) v[r] = -1e30f;
        else if (row >= WTOK) v[r] = 0.f;
        else {
            const int qy = row / WS, qx = row % WS, ky = col / WS, kx = col % WS;
            // relative_position_index (swin_transformer.py:267-279)
            v[r] = 1.4426950408889634f * table[((qy - ky + WS - 1) * (2 * WS - 1) + (qx - kx + WS - 1)) * HEADS + head];
        }
    }
    frag[i] = make_float4(v[0], v[1], v[2], v[3]);
}

int build_bias_frag(cudaStream_t st, const float* table, float* frag) {
    build_bias_frag_kernel<<<(HEADS * 3 * 6 * 32 + 255) / 256, 256, 0, st>>>(table, reinterpret_cast<float4*>(frag));
    NB_LAUNCHED();
    return 0;
}

template <int D>
static size_t attn_smem_bytes() {
    constexpr int C = D * HEADS, LD = C + 8;
    return (size_t)3 * WTOK * LD * 2 + (WTOK + WPAD) * 4 + 16;
}

// opt-in shared memory + carveout, per (device, kernel): both attributes are per-device state
static int set_attn_attrs(const void* func, size_t smem, int carveout) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> done;
    int dev = 0;
    NB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    int& have = done[{dev, func}];
    if (have != carveout) {
        NB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        have = carveout;
    }
    return 0;
}

int window_attention(cudaStream_t st, const __half* qkv, const float* bias_frag_f, __half* out, int B, int H, int W, int C,
                     int shift, size_t plane) {
    const float4* bias_table = reinterpret_cast<const float4*>(bias_frag_f);
    NB_CHECK(H % WS == 0 && W % WS == 0, "feature map must be a multiple of the 6x6 window");
    NB_CHECK(C == 96 || C == 192, "window attention supports C=96 (d=16) and C=192 (d=32)");
    if (WS >= H) shift = 0;  // torchvision :151-155
    dim3 grid((H / WS) * (W / WS), B);
    ProfScope ps(st, PC_ATTN, (double)B * H * W * C * 4 * 2, (double)B * H * W * C * 3 * 2, (double)B * H * W * C * 2);  // q,k,v in; out
    // shared-memory carveout: just enough for the 4 resident CTAs, the rest stays L1 (the per-head bias fragments,
    // 55 KB per layer, are re-read by every window and should hit there).  g_tune[6] overrides the percentage.
    if (C == 96) {
        const int want = g_tune[6] > 0 ? g_tune[6] : 72;   // 6 CTAs x 23.6 KB
        if (set_attn_attrs((const void*)window_attention_mma_kernel<16>, attn_smem_bytes<16>(), want)) return 1;
        window_attention_mma_kernel<16><<<grid, 192, attn_smem_bytes<16>(), st>>>(qkv, bias_table, out, H, W, shift, plane);
    } else {
        const int want = g_tune[6] > 0 ? g_tune[6] : 86;
        if (set_attn_attrs((const void*)window_attention_mma_kernel<32>, attn_smem_bytes<32>(), want)) return 1;
        window_attention_mma_kernel<32><<<grid, 192, attn_smem_bytes<32>(), st>>>(qkv, bias_table, out, H, W, shift, plane);
    }
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb200
