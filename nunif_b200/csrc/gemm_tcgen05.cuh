// tcgen05 implicit-GEMM for NHWC fp16 activations (sm_100a).
//
//   D[m, n] = act( sum_taps sum_c A[pixel(m) + tap, c] * Wt[n, tap, c] + bias[n] ) (+ residual)
//
// One kernel covers every dense contraction on path A:
//   * Linear layers of the Swin blocks (1 tap, "pixel" = token)            swin_transformer.py:177,228,444
//   * 3x3 valid convolutions of CUNet / the Swin patch stem (9 taps)       cunet.py:14-17,38,41 ; swin_unet.py:133-136
//   * 2x2 stride-2 convolutions (2 taps over a (2C, W/2, 2, H/2, B) view)  cunet.py:36,78,80 ; swin_unet.py:49
//   * ConvTranspose 2x2 s2 / Linear+pixel_shuffle(2) (N = 4*Cout, scatter) cunet.py:38,82,84 ; swin_unet.py:69-82
//
// Structure (Blackwell native):
//   warp 0   : TMA producer - one 5-D box load per (tap, channel chunk) for A, one 2-D box for B,
//              128B/64B hardware swizzle, mbarrier complete_tx
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BLOCK_N, K=16 per instr),
//              tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-9: epilogue - tcgen05.ld accumulator rows, bias + activation (+ residual tile fetched by TMA),
//              fp16 results staged in 128B-swizzled shared memory and written with TMA tensor stores
//              (cp.async.bulk.tensor ... bulk_group), so every global access of the kernel is a full-line bulk copy
// Non-persistent, one 128 x BLOCK_N output tile per CTA; shared memory is sized so that two CTAs
// are co-resident per SM, which overlaps one tile's epilogue with the next tile's loads.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace nb200 {

enum : int { ACT_NONE = 0, ACT_LRELU01 = 1, ACT_GELU = 2, ACT_RELU = 3 };
enum : int { OUT_NHWC = 0, OUT_PIXSHUF2 = 1, OUT_SPLIT = 2 };

struct GemmParams {
    // output tiling: the M dimension is (b, y, x) over Ho x Wo pixels, tiled TH x TW (TH*TW == 128)
    int B, Ho, Wo, TH, TW, tiles_x, tiles_y;
    int N;               // output channels (GEMM N)
    int taps, cpt;       // taps and BK-chunks per tap (K = taps*cpt*BK)
    int8_t tap_dx[16], tap_dy[16], tap_dyi[16];
    int n_tiles;
    // epilogue (output / residual tensors are described by the tensor maps in GemmMaps)
    const float* bias;   // [N] or null
    int act;
    int out_mode, cout;  // OUT_PIXSHUF2: N = 4*cout ordered (dy,dx,co); maps o[g]/r[g] are the stride-2 views
                         // OUT_SPLIT: N = nsplit*cout, block g goes to its own dense [M][cout] plane (maps o[g])
    int has_res, res_cy, res_cx;
    int res_before_act;  // 0: out = act(acc+bias) + res ; 1: out = act(acc+bias+res)
};

// All tensor maps of one launch (a single __grid_constant__ parameter).
struct GemmMaps {
    CUtensorMap a, b;
    CUtensorMap o[4];    // output: NHWC view (c, x, y, b); pixel-shuffle mode: one stride-2 view per (dy,dx)
    CUtensorMap r[4];    // residual, same tiling as the output
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cold path of mbar_wait, kept out of line: the call sites of a warp-specialised kernel are many and instruction-cache space is
// what those kernels run out of first (profiles/r2/attn_tc_ncu_v7_source.txt: most stalls of the working warps are no-instruction)
static __device__ __noinline__ void mbar_timeout() {
    printf("nb200: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    // try_wait WITH a suspend-time hint: the thread sleeps in hardware until the phase completes (or 20 us pass) instead of
    // re-issuing the instruction.  Without the hint the default time limit is short and every waiting warp of a warp-specialised
    // CTA spins in the issue slots of its scheduler: with ~5 waiting warps per scheduler the working warps of swin_attn_tc.cu ran
    // at ~10 cycles per instruction (profiles/r2/attn_tc_timeline_v5.txt).  Bounded: a protocol bug traps instead of hanging the GPU.
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 18); ++it) {     // (the compiler unrolled this loop 32x at every call site: 2048 try_waits per kernel)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (done) return;
    }
    mbar_timeout();
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
    constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2 : (SWIZZLE_BYTES == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * SWIZZLE_BYTES) >> 4;  // 8 rows of one swizzle span
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Exact-erf GELU with ONE special-function op.  erfc(|x|/sqrt2) = 2^-q(|x|) with q a degree-5 polynomial
// (least-squares/minimax fit on [0, 6], max |gelu error| 5.3e-7 - below fp32 rounding of the surrounding math,
// and far below the fp16 rounding of the stored result).  gelu(x) = x * Phi(x), Phi = 1 - E/2 (x>0) or E/2 (x<0).
// nn.GELU (erf form) in torchvision's MLP, swin_transformer.py:444.  erff() costs ~40 instructions and two MUFU ops,
// which made the fc1 epilogue ALU/MUFU-bound; this is ~12 instructions and one MUFU.EX2.
__device__ __forceinline__ float gelu_erf(float x) {
    const float a = fabsf(x);
    float q = fmaf(a, 4.88118734e-04f, -7.19881030e-03f);
    q = fmaf(q, a, 5.21468017e-02f);
    q = fmaf(q, a, 4.59595724e-01f);
    q = fmaf(q, a, 1.15100057e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(-q, a, -1.0f)));   // Phi(-|x|) = erfc(|x|/sqrt2)/2: one MUFU.EX2
    // x*Phi(x) = max(x, 0) - |x|*Phi(-|x|) on both sides of zero: no select, 9 instructions per value
    return fmaxf(x, 0.f) - a * e;
}

// The same GELU on two values with packed fp32 instructions (FFMA2 / fma.rn.f32x2, sm_100+): identical IEEE results per
// lane, but the polynomial, the exponent FMA and the final FMA issue once per PAIR.  On B200 a scalar FFMA already issues
// at 1/clk/SMSP and FFMA2 at 1 per 2 clk (profiles/r1/ffma2_tput.json), so this adds no FMA throughput, it only frees issue
// slots (~13 -> ~7.5 instructions per value).  MEASURED SLOWER in the GEMM epilogue (fc1 247 -> 267 us: the register-pair
// moves and the longer dependent chain cost more than the issue slots save), so apply_act16 uses the scalar gelu_erf;
// kept for the record of the experiment.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const unsigned long long a = pk2(a0, a1);
    unsigned long long q = ffma2(a, pk2(4.88118734e-04f, 4.88118734e-04f), pk2(-7.19881030e-03f, -7.19881030e-03f));
    q = ffma2(q, a, pk2(5.21468017e-02f, 5.21468017e-02f));
    q = ffma2(q, a, pk2(4.59595724e-01f, 4.59595724e-01f));
    q = ffma2(q, a, pk2(1.15100057e+00f, 1.15100057e+00f));
    const unsigned long long na = pk2(-a0, -a1);
    const unsigned long long t = ffma2(q, na, pk2(-1.0f, -1.0f));            // -(q*a) - 1
    float t0, t1, e0, e1;
    upk2(t, t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    const unsigned long long r = ffma2(na, pk2(e0, e1), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));   // max(x,0) - |x|*e
    upk2(r, x0, x1);
}

// activation over a 16-value fragment; `act` is warp-uniform, so the switch is hoisted out of the element loop
__device__ __forceinline__ void apply_act16(float (&v)[16], int act) {
    switch (act) {
        case ACT_LRELU01:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.1f * v[j];
            break;
        case ACT_GELU:
            // scalar form: the packed-FMA variant below measured 8 % slower on the fc1 shapes (267 vs 247 us)
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
            break;
        case ACT_RELU:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            break;
        default: break;
    }
}

constexpr int GEMM_EPI_WARPS = 8;                       // non-persistent kernel: two warps per TMEM lane group
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

template <int BLOCK_N, int BK>
struct GemmCfg {
    static constexpr int SWIZZLE = BK * 2;  // bytes per K-row of a stage: 128 (BK=64) or 64 (BK=32)
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = BLOCK_N * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
    static constexpr int STAGES = (98304 / STAGE_BYTES) < 2 ? 2 : ((98304 / STAGE_BYTES) > 6 ? 6 : (98304 / STAGE_BYTES));
    // epilogue staging: BLOCK_N/CW chunks of [128 rows][CW cols] fp16, swizzle span = CW*2 bytes
    static constexpr int CW = (BLOCK_N % 64 == 0) ? 64 : ((BLOCK_N % 32 == 0) ? 32 : 16);
    static constexpr int NCH = BLOCK_N / CW;
    static constexpr int CH_BYTES = 128 * CW * 2;
    static_assert(NCH * CH_BYTES <= STAGES * STAGE_BYTES, "epilogue staging must fit in the (drained) pipeline stages");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
};

// byte offset of 16-byte chunk j of row r inside a [128][CW] staging tile with the TMA swizzle of span CW*2
template <int CW>
__device__ __forceinline__ uint32_t stage_off(int r, int j) {
    if (CW == 64) return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
    if (CW == 32) return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
    return (uint32_t)(r * 32 + ((j ^ ((r >> 2) & 1)) << 4));
}

template <int BLOCK_N, int BK>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_conv_kernel(const __grid_constant__ GemmMaps maps,
                                                                 const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, BK>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CW = Cfg::CW;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* res_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // n tile is the fast grid index so CTAs sharing an A tile run together (A re-reads hit L2)
    const int n_tile = blockIdx.x % p.n_tiles;
    const int tile = blockIdx.x / p.n_tiles;
    const int tx_i = tile % p.tiles_x;
    const int ty_i = (tile / p.tiles_x) % p.tiles_y;
    const int b = tile / (p.tiles_x * p.tiles_y);
    const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
    const int n0 = n_tile * BLOCK_N;
    const int k_iters = p.taps * p.cpt;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.o[0]);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(res_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            for (int it = 0; it < k_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int tap = it / p.cpt, ch = it - tap * p.cpt;
                uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                mbar_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_5d(&maps.a, &full_bar[s], sa, ch * BK, x0 + p.tap_dx[tap], p.tap_dyi[tap], y0 + p.tap_dy[tap], b);
                tma_load_2d(&maps.b, &full_bar[s], sb, it * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BLOCK_N);
        for (int it = 0; it < k_iters; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t ad = make_kmajor_desc<Cfg::SWIZZLE>(sa + k * 32);
                    const uint64_t bd = make_kmajor_desc<Cfg::SWIZZLE>(sb + k * 32);
                    umma_f16(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);                         // frees the smem stage when the MMAs retire
                if (it == k_iters - 1) umma_commit(tmem_full_bar);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int lane_grp = warp & 3;       // TMEM lanes [32*lane_grp, +32) are the ones this warp may read
        const int half = (warp - 2) >> 2;    // which 16-column blocks of the row this warp converts
        const int r = lane_grp * 32 + lane;  // accumulator row == pixel within the tile == staging row
        const bool leader = (warp == 2 && lane == 0);
        mbar_wait(tmem_full_bar, 0);         // all MMAs retired => every pipeline stage is drained and reusable
        tc_fence_after();
        uint8_t* stg = smem;
        if (p.has_res) {
            if (leader) {
                mbar_expect_tx(res_bar, Cfg::NCH * Cfg::CH_BYTES);
#pragma unroll 1
                for (int c = 0; c < Cfg::NCH; ++c) {
                    const int n = n0 + c * CW;
                    const int g = p.out_mode != OUT_NHWC ? n / p.cout : 0;
                    const int co = p.out_mode != OUT_NHWC ? n - g * p.cout : n;
                    tma_load_4d(&maps.r[g], res_bar, stg + c * Cfg::CH_BYTES, co, x0 + p.res_cx, y0 + p.res_cy, b);
                }
            }
            mbar_wait(res_bar, 0);
        }
        const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        const int act = p.act;
        const bool has_res = p.has_res != 0, res_first = p.res_before_act != 0;
#pragma unroll 1
        for (int sb = half; sb < BLOCK_N / 16; sb += 2) {
            uint32_t acc[16];
            tmem_ld16(trow + sb * 16, acc);
            tmem_ld_wait();
            const int c = (sb * 16) / CW, sub = sb - c * (CW / 16);
            uint8_t* buf = stg + c * Cfg::CH_BYTES;
            uint4* s0 = reinterpret_cast<uint4*>(buf + stage_off<CW>(r, 2 * sub));
            uint4* s1 = reinterpret_cast<uint4*>(buf + stage_off<CW>(r, 2 * sub + 1));
            float v[16];
            if (p.bias) {
                const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + sb * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 bq = __ldg(bp + q);
                    v[4 * q] = __uint_as_float(acc[4 * q]) + bq.x;
                    v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + bq.y;
                    v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + bq.z;
                    v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + bq.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
            }
            if (has_res) {
                float rv[16];
                const uint4 r0 = *s0, r1 = *s1;
                const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a = __half22float2(h0[j]), d = __half22float2(h1[j]);
                    rv[2 * j] = a.x; rv[2 * j + 1] = a.y; rv[8 + 2 * j] = d.x; rv[8 + 2 * j + 1] = d.y;
                }
                if (res_first) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += rv[j];
                    apply_act16(v, act);
                } else {
                    apply_act16(v, act);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += rv[j];
                }
            } else {
                apply_act16(v, act);
            }
            __align__(16) __half2 o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            *s0 = reinterpret_cast<const uint4*>(o)[0];
            *s1 = reinterpret_cast<const uint4*>(o)[1];
        }
        fence_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        epi_bar_sync();
        if (leader) {
#pragma unroll 1
            for (int c = 0; c < Cfg::NCH; ++c) {
                const int n = n0 + c * CW;
                const int g = p.out_mode != OUT_NHWC ? n / p.cout : 0;
                const int co = p.out_mode != OUT_NHWC ? n - g * p.cout : n;
                // out-of-range rows/cols of edge tiles are clipped by TMA
                tma_store_4d(&maps.o[g], stg + c * Cfg::CH_BYTES, co, x0, y0, b);
            }
            tma_store_commit();
            tma_store_wait_read();  // smem must stay valid until the bulk stores have read it
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace nb200
