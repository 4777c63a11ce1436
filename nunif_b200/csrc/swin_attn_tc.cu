// Fused Swin block head, ALL contractions on tcgen05 (sm_100a):   att = shifted_window_attention_core( x . Wqkv^T + bqkv )
//
// torchvision shifted_window_attention (swin_transformer.py:166-221) up to, not including, the `proj` Linear: roll, 6x6 window
// partition, qkv Linear (:177), scale, QK^T, relative position bias (:190), shift mask (:193-209), softmax (:211), attn@v
// (:214), window reverse, un-roll.  swin_fused_attn.cu (round 2, first half) ran the qkv GEMM on tcgen05 but the 36x36xd
// attention of every (window, head) on mma.sync in 16 latency-bound warps: its clock64 timeline
// (profiles/r2/fused_timeline_attn_192_v2.txt) shows 19 000 cycles per 3-window tile, ~15 600 of them in those warps, against
// ~5 000 cycles of tensor work.  Here S = QK^T and O = PV are UMMAs as well:
//
//   a tile = 3 windows = 108 tokens = the 128 rows of one UMMA; a UNIT = 96 qkv columns = one head (C = 192, d = 32) or a
//   head pair (C = 96, d = 16), packed [q | k | v] per unit at load time.
//     G(u)  D[u&1] (128 x 96)  = X (128 x C) . Wu^T                         K = C        accumulator ring of 2 x 96 TMEM columns
//     E(u)  D -> + bias -> fp16 -> Q[u&1], K[u&1] (128 x 32, 64B swizzle), V^T[u&1] (d+16 x 128 tokens, row d = 1)
//     S(h)  S[h&1] (128 x 112) = Qh . Kh^T                                   K = d        key j of window w = column 36 w + j
//     softmax(h): row r reads ITS window's 36 columns, scale + bias (+ shift mask), base-2 softmax, P[h&1] (128 x 128 fp16,
//           block diagonal: row r writes columns 36 w .. 36 w + 35, everything else stays zero) -> shared memory
//     PV(h) O[h&1] (128 x d+16) = P . Vh                                      K = 112      column d of O = row sum (ones row of V^T)
//     O epilogue: O / rowsum -> fp16 -> att at the un-rolled token position (run one head late, under the next PV)
//   q, k, v, S and P never reach HBM; the kernel reads x and writes att.
//
// 24 warps: 0-1 weight producers (ring slot s belongs to producer s & 1: bulk-tensor ops of one warp execute one after the
// other, ~0.34 us each, profiles/r2/tma_inflight.json), 2 / 3 / 4 the tcgen05.mma issuers of the G, S and PV queues (each a
// blocking loop on its own barriers, so no queue ever blocks another), 5-7 activation producers (one window
// each; the window gather is done by the TMA unit, a window that wraps around the rolled image is 2 or 4 boxes and E() restores
// the token order), 8-15 two E warpgroups (one per accumulator buffer), 16-23 two softmax warpgroups (heads alternate; each
// also runs the O epilogue of its head).
#include "gemm_tcgen05.cuh"
#include "swin_fused.h"
#include "tmap.h"

namespace nb200 {

extern unsigned long long* g_timeline;   // gemm.cu (nb200_debug_timeline)

namespace {

constexpr int WS = 6, WTOK = 36, HEADS = 6, WPT = 3;
constexpr int SN = 112;                  // N of S: 108 keys (window w at columns 36 w ..) rounded up to 16
constexpr int RT = 121;                  // (2 * 6 - 1)^2 relative positions
constexpr int BT_FLOATS = HEADS * RT;
constexpr int UN = 96;                   // qkv columns per unit
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// Everything below addresses shared memory by its 32-bit shared-space address: generic pointers cost a cvta and 64-bit
// arithmetic at every use, and instruction-cache footprint is what this kernel runs out of first (see the header).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// try_wait in a PTX-level loop (a suspend-time hint measured no faster and triples the code of every call site).  UNBOUNDED: the protocol below is
// validated by tests/test_gpu_fused.py::test_swin_attn_tc; build with -DNB200_TC_BOUNDED_WAITS (traps after ~5 s) when changing it.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef NB200_TC_BOUNDED_WAITS
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 18) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (!done) mbar_timeout();
#else
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// explicit shared-space accesses
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, __half h) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(__half_as_ushort(h)) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int C>
struct TcCfg {
    static constexpr int D = C / HEADS;                     // 32 | 16
    static constexpr int NHU = 32 / D;                      // heads per unit: 1 | 2
    static constexpr int UPT = HEADS / NHU;                 // units per tile: 6 | 3
    static constexpr int KCH = (C + 63) / 64;               // K-chunks of the qkv GEMM (the last one may be half out of bounds)
    static constexpr int KLAST = (C - (KCH - 1) * 64) / 16;
    static constexpr int NV = D + 16;                       // rows of V^T: d values, the ones row, 15 zero rows
    static constexpr int XCH = 128 * 128;
    static constexpr int XB = KCH * XCH;
    static constexpr int WST = UN * 128;                    // weight ring stage [96][64]
    static constexpr int STAGES = 4;
    static constexpr int QB = 128 * 64, KB = SN * 64;       // 64B-swizzled [rows][32 halfs]
    static constexpr int VCH = NV * 128;                    // one 64-token chunk of V^T
    static constexpr int VTB = 2 * VCH;
    static constexpr int PCH = 128 * 128, PB = 2 * PCH;
    static constexpr int TM_D = 0, TM_S = 192, TM_SSTRIDE = 160, TM_O = SN;   // slot = S (112 columns) | O (48)
    static constexpr int THREADS = 768;
    static constexpr int OFF_W = XB, OFF_Q = OFF_W + STAGES * WST, OFF_K = OFF_Q + 2 * QB, OFF_V = OFF_K + 2 * KB,
                         OFF_P = OFF_V + 2 * NHU * VTB, OFF_BT = OFF_P + 2 * PB, OFF_BIAS = OFF_BT + 3072,
                         OFF_BAR = OFF_BIAS + 3 * C * 4, DATA = OFF_BAR + 512;
    static constexpr size_t SMEM = 1024 + (size_t)DATA;
    static_assert(SMEM <= 232448, "shared memory of one SM");
    static_assert(OFF_Q % 1024 == 0 && OFF_K % 512 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0 && VCH % 1024 == 0, "swizzle atoms");
};

struct TcMaps {
    CUtensorMap x66, x36, x63, x33;   // x as (c, x, y, b); boxes (64, 6, 6), (64, 3 wide, 6), (64, 6, 3 tall), (64, 3, 3)
    CUtensorMap w;                    // packed Wqkv [3C][C], box (64, 96)
};
struct TcParams {
    unsigned long long* tl;
    int B, H, W, shift;
    int nww, nwh, tpi, tiles;   // tpi = tiles per image
    const float* bqkv;       // packed order
    const float* bias_tab;   // relative_position_bias_table [121][6]
    __half* att;
};

struct WinInfo {
    int b, wy, wx;
    bool valid, xs, ys;      // xs / ys: the window wraps around the rolled image in x / y
};
// out of line: three roles decode windows once per tile; two integer divisions by run-time values are ~100 instructions inlined
__device__ __noinline__ WinInfo win_info(const TcParams& p, int tile, int w) {
    // tiles never straddle images (p.tpi tiles per image, the last one of every image may be partial): a window's slot in
    // its tile - and with it the accumulation order of its PV product - depends only on its index inside its own image, so
    // the result is bit-identical for every batch size (tests/test_gpu_models.py::test_swin_tiled_render_golden)
    WinInfo wi;
    const int wpi = p.nww * p.nwh;
    wi.b = tile / p.tpi;
    const int win = (tile - wi.b * p.tpi) * WPT + w;
    wi.valid = win < wpi;
    const int rem = wi.valid ? win : 0;
    wi.wy = rem / p.nww;
    wi.wx = rem - wi.wy * p.nww;
    wi.xs = p.shift > 0 && wi.wx == p.nww - 1;
    wi.ys = p.shift > 0 && wi.wy == p.nwh - 1;
    return wi;
}
// landed row l of a window (order of the TMA boxes) -> token index i = y * 6 + x of the window
__device__ __forceinline__ int landed_to_token(int l, bool xs, bool ys) {
    if (!xs) return l;                                   // one box, or two boxes of 3 whole rows each
    if (!ys) { const int hx = l >= 18 ? 1 : 0, rem = l - 18 * hx; const int y = rem / 3; return y * 6 + hx * 3 + (rem - 3 * y); }
    const int qd = l / 9, rem = l - 9 * qd, yy = rem / 3, xx = rem - 3 * yy;
    return ((qd >> 1) * 3 + yy) * 6 + (qd & 1) * 3 + xx;
}

template <int C>
__global__ void __launch_bounds__(TcCfg<C>::THREADS, 1) swin_attn_tc_kernel(const __grid_constant__ TcMaps maps,
                                                                            const __grid_constant__ TcParams p) {
    using Cfg = TcCfg<C>;
    constexpr int D = Cfg::D, NHU = Cfg::NHU, UPT = Cfg::UPT, KCH = Cfg::KCH, NV = Cfg::NV, XCH = Cfg::XCH, WST = Cfg::WST, S = Cfg::STAGES;
    constexpr int THREADS = Cfg::THREADS;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);   // generic: prologue fills only
    const uint32_t sb = smem_u32(smem);
    const uint32_t sX = sb;                        // [KCH][128][64] 128B-swizzled; rows 108..127 stay zero
    const uint32_t sW = sb + Cfg::OFF_W;           // weight ring
    const uint32_t sQ = sb + Cfg::OFF_Q;           // 2 x [128][32]
    const uint32_t sK = sb + Cfg::OFF_K;           // 2 x [112][32]
    const uint32_t sV = sb + Cfg::OFF_V;           // 2 x NHU x V^T
    const uint32_t sP = sb + Cfg::OFF_P;           // 2 x [2 chunks][128][64]
    const uint32_t sBT = sb + Cfg::OFF_BT;         // [6][121] fp32, log2(e) folded in
    const uint32_t sBias = sb + Cfg::OFF_BIAS;     // [3C] fp32, packed order
    // mbarriers (8 bytes each)
    const uint32_t w_full = sb + Cfg::OFF_BAR;     // [S]
    const uint32_t w_empty = w_full + 8 * S;       // [S]
    const uint32_t x_full = w_empty + 8 * S;
    const uint32_t x_empty = x_full + 8;
    const uint32_t d_full = x_empty + 8;           // [2]
    const uint32_t d_empty = d_full + 16;          // [2]
    const uint32_t qk_full = d_empty + 16;         // [2]
    const uint32_t qk_empty = qk_full + 16;        // [2]
    const uint32_t v_full = qk_empty + 16;         // [2]
    const uint32_t v_empty = v_full + 16;          // [2]
    const uint32_t s_full = v_empty + 16;          // [2]
    const uint32_t s_free = s_full + 16;           // [2]
    const uint32_t o_full = s_free + 16;           // [2]
    const uint32_t p_full = o_full + 16;           // [2]
    const uint32_t p_empty = p_full + 16;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 8 * (2 * S + 2 + 22));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // debug timeline (CTA 0): tracks 0 MMA issuer, 1 E warp 8, 2 softmax warp 16, 3 softmax warp 20, 4 X producer 5, 5 W producer 0
    unsigned long long* tlb = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    int tli = 0;
#ifdef NB200_TC_TIMELINE   // ~30 probes of ~12 instructions: compiled out of the shipped kernel (instruction-cache footprint)
#define TTL(track, tag, aux) do { if (tlb && tli < 2048) { tlb[(track) * 2048 + tli] = ((unsigned long long)(tag) << 56) | ((unsigned long long)((aux) & 0xffff) << 40) | (clock64() & 0xffffffffffull); ++tli; } } while (0)
#else
#define TTL(track, tag, aux) do { } while (0)
    (void)tlb; (void)tli;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.x66); tma_prefetch_desc(&maps.x36); tma_prefetch_desc(&maps.x63); tma_prefetch_desc(&maps.x33);
        tma_prefetch_desc(&maps.w);
        for (int s = 0; s < S; ++s) { mbar_init((w_full + 8 * (s)), 1); mbar_init((w_empty + 8 * (s)), 1); }
        mbar_init(x_full, WPT); mbar_init(x_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init((d_full + 8 * (s)), 1); mbar_init((d_empty + 8 * (s)), 4);
            mbar_init((qk_full + 8 * (s)), 4); mbar_init((qk_empty + 8 * (s)), 1); mbar_init((v_full + 8 * (s)), 4); mbar_init((v_empty + 8 * (s)), 1);
            mbar_init((s_full + 8 * (s)), 1); mbar_init((s_free + 8 * (s)), 4); mbar_init((o_full + 8 * (s)), 1);
            mbar_init((p_full + 8 * (s)), 4); mbar_init((p_empty + 8 * (s)), 1);
        }
        fence_barrier_init();
    }
    if (threadIdx.x == 0) NB_PDL_TRIGGER();
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    // launch constants and zero fills: bias table, qkv bias; x tile (its 20 pad rows are never written again), Q / K (rows that
    // no token owns), P (everything outside the diagonal blocks), V^T (rows d+1.. and the ones row d)
    for (int i = threadIdx.x; i < BT_FLOATS; i += THREADS) reinterpret_cast<float*>(smem + Cfg::OFF_BT)[i] = LOG2E * __ldg(p.bias_tab + (i % RT) * HEADS + i / RT);
    for (int i = threadIdx.x; i < 3 * C; i += THREADS) reinterpret_cast<float*>(smem + Cfg::OFF_BIAS)[i] = __ldg(p.bqkv + i);
    for (int i = threadIdx.x; i < Cfg::XB / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (Cfg::OFF_BT - Cfg::OFF_Q) / 16; i += THREADS) reinterpret_cast<uint4*>(smem + Cfg::OFF_Q)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * NHU * 2 * 8; i += THREADS)     // ones row: 8 x 16 B per chunk (the swizzle permutes 16 B units inside a row)
        reinterpret_cast<uint4*>(smem + Cfg::OFF_V + (i >> 3) * Cfg::VCH + D * 128)[i & 7] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_async_smem();     // the fills are read by the tensor core through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.x, stride = gridDim.x;
    const int ntl = (p.tiles - first + stride - 1) / stride;   // tiles of this CTA (grid <= tiles)
    const int total_units = ntl * UPT, total_heads = ntl * HEADS;

    if (warp < 2) {
        // ===================== weight producers: global stage gs -> ring slot gs % 4, owned by warp gs & 1 =====================
        if (elect_one()) {
            const int nst = total_units * KCH;
            for (int gs = warp; gs < nst; gs += 2) {
                const int gu = gs / KCH, kc = gs - gu * KCH, u = gu % UPT, ws = gs % S;
                if (warp == 0) TTL(5, 50, u);
                mbar_wait((w_empty + 8 * (ws)), ((gs / S) & 1) ^ 1);
                if (warp == 0) TTL(5, 51, u);
                mbar_expect_tx((w_full + 8 * (ws)), WST);
                tma_load_2d(&maps.w, (w_full + 8 * (ws)), sW + ws * WST, kc * 64, u * UN);
            }
        }
    } else if (warp == 2) {
        // ===================== G issuer: D[u & 1] = X . Wu^T, one K-chunk per weight stage =====================
        // Three issuer warps (G, S, PV), each a plain blocking loop: an idle issuer sleeps in mbarrier.try_wait and costs no issue
        // slots.  (History, profiles/r2/attn_tc_timeline_v*.txt: ONE warp polling all queues spent ~60 instructions per pass, and
        // at the ~19 cycles per instruction a warp of this kernel gets (ncu: 17 % no-instruction, 40 % scoreboard) that is one issued
        // item per ~1000 cycles - every hand-off of the pipeline waited on the issuer.)
        const uint32_t idesc_g = make_idesc_f16(UN);
        const uint32_t aX = sX, aW = sW;
        int gstage = 0;
        for (int gu = 0; gu < total_units; ++gu) {
            const int b = gu & 1, u = gu % UPT;
            mbar_wait((d_empty + 8 * (b)), ((gu >> 1) & 1) ^ 1);
            if (u == 0) mbar_wait(x_full, (gu / UPT) & 1);
#pragma unroll 1
            for (int kc = 0; kc < KCH; ++kc, ++gstage) {
                const int ws = gstage % S;
                mbar_wait((w_full + 8 * (ws)), (gstage / S) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t td = tmem_base + Cfg::TM_D + b * UN;
                    const int ksteps = kc == KCH - 1 ? Cfg::KLAST : 4;
                    TTL(0, 1, gu * 4 + kc);
                    for (int k = 0; k < ksteps; ++k)
                        umma_f16(td, make_kmajor_desc<128>(aX + kc * XCH + k * 32), make_kmajor_desc<128>(aW + ws * WST + k * 32), idesc_g,
                                 (kc > 0 || k > 0) ? 1u : 0u);
                    umma_commit((w_empty + 8 * (ws)));
                    if (kc == KCH - 1) {
                        umma_commit((d_full + 8 * (b)));
                        if (u == UPT - 1) umma_commit(x_empty);   // the activation tile may be overwritten
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ===================== S issuer: S[h & 1] = Qh . Kh^T =====================
        const uint32_t idesc_s = make_idesc_f16(SN);
        const uint32_t aQ = sQ, aK = sK;
        for (int h = 0; h < total_heads; ++h) {
            const int gu = h / NHU, hh = h - gu * NHU, b = gu & 1, sl = h & 1;
            if (hh == 0) mbar_wait((qk_full + 8 * (b)), (gu >> 1) & 1);
            mbar_wait((s_free + 8 * (sl)), ((h >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t td = tmem_base + Cfg::TM_S + sl * Cfg::TM_SSTRIDE;
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_f16(td, make_kmajor_desc<64>(aQ + b * Cfg::QB + hh * 2 * D + k * 32), make_kmajor_desc<64>(aK + b * Cfg::KB + hh * 2 * D + k * 32),
                             idesc_s, k > 0 ? 1u : 0u);
                umma_commit((s_full + 8 * (sl)));
                if (hh == NHU - 1) umma_commit((qk_empty + 8 * (b)));
            }
            __syncwarp();
        }
    } else if (warp == 4) {
        // ===================== PV issuer: O[h & 1] = P[h & 1] . Vh (K = 112: tokens 0..111) =====================
        const uint32_t idesc_o = make_idesc_f16(NV);
        const uint32_t aV = sV, aP = sP;
        for (int h = 0; h < total_heads; ++h) {
            const int gu = h / NHU, hh = h - gu * NHU, b = gu & 1, sl = h & 1;
            if (hh == 0) mbar_wait((v_full + 8 * (b)), (gu >> 1) & 1);
            mbar_wait((p_full + 8 * (sl)), (h >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t td = tmem_base + Cfg::TM_S + sl * Cfg::TM_SSTRIDE + Cfg::TM_O;
                const uint32_t av = aV + (b * NHU + hh) * Cfg::VTB, ap = aP + sl * Cfg::PB;
                TTL(0, 3, h);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int kc2 = k >> 2, kk = k & 3;
                    umma_f16(td, make_kmajor_desc<128>(ap + kc2 * Cfg::PCH + kk * 32), make_kmajor_desc<128>(av + kc2 * Cfg::VCH + kk * 32), idesc_o,
                             k > 0 ? 1u : 0u);
                }
                umma_commit((o_full + 8 * (sl)));
                umma_commit((p_empty + 8 * (sl)));
                if (hh == NHU - 1) umma_commit((v_empty + 8 * (b)));
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ===================== activation producers: window j of every tile =====================
        const int j = warp - 5;
        if (elect_one()) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int ti = 0; ti < ntl; ++ti) {
                const int tile = first + ti * stride;
                if (j == 0) TTL(4, 40, ti);
                mbar_wait(x_empty, (ti & 1) ^ 1);
                if (j == 0) TTL(4, 41, ti);
                const WinInfo wi = win_info(p, tile, j);
                if (!wi.valid) { mbar_arrive(x_full); continue; }
                mbar_expect_tx(x_full, (uint32_t)(WTOK * 128 * KCH));
                const int y0 = wi.wy * WS + p.shift, x0 = wi.wx * WS + p.shift;   // torch.roll(-shift): window row r <- row (r + shift) % H
                const uint32_t dst = sX + j * WTOK * 128;
#pragma unroll 1
                for (int kc = 0; kc < KCH; ++kc) {
                    const uint32_t d = dst + kc * XCH;
                    if (!wi.xs && !wi.ys) {
                        tma_load_4d(&maps.x66, x_full, d, kc * 64, x0, y0, wi.b);
                    } else if (!wi.xs) {          // two boxes of 3 whole window rows
                        tma_load_4d(&maps.x63, x_full, d, kc * 64, x0, y0, wi.b);
                        tma_load_4d(&maps.x63, x_full, d + 18 * 128, kc * 64, x0, 0, wi.b);
                    } else if (!wi.ys) {          // left / right halves: landed order (half, y, x % 3)
                        tma_load_4d(&maps.x36, x_full, d, kc * 64, x0, y0, wi.b);
                        tma_load_4d(&maps.x36, x_full, d + 18 * 128, kc * 64, 0, y0, wi.b);
                    } else {                      // four 3x3 quadrants: landed order (hy, hx, y % 3, x % 3)
                        tma_load_4d(&maps.x33, x_full, d, kc * 64, x0, y0, wi.b);
                        tma_load_4d(&maps.x33, x_full, d + 9 * 128, kc * 64, 0, y0, wi.b);
                        tma_load_4d(&maps.x33, x_full, d + 18 * 128, kc * 64, x0, 0, wi.b);
                        tma_load_4d(&maps.x33, x_full, d + 27 * 128, kc * 64, 0, 0, wi.b);
                    }
                }
            }
        }
    } else if (warp < 16) {
        // ===================== E warpgroups (warps 8-11: accumulator 0, 12-15: accumulator 1) =====================
        const int b = (warp - 8) >> 2, q = warp & 3;
        const int r = q * 32 + lane;                         // landed row of the tile = TMEM lane
        const int w = r < 36 ? 0 : (r < 72 ? 1 : 2);
        const int l = r - 36 * w;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::TM_D + b * UN;
        const uint32_t qb = sQ + b * Cfg::QB, kb = sK + b * Cfg::KB, vb = sV + b * NHU * Cfg::VTB;
        int cur_ti = -1;
        bool rv = false;
        uint32_t qoff[4] = {0, 0, 0, 0};     // byte offset of the four 16-byte pieces of row R in a 64B-swizzled [rows][32] operand
        uint32_t vsw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, vcol = 0;   // V^T: column bytes of token R, swizzled 16-byte unit per (d & 7)
        for (int gu = b; gu < total_units; gu += 2) {
            const int ti = gu / UPT, u = gu - ti * UPT;
            if (ti != cur_ti) {
                cur_ti = ti;
                const WinInfo wi = win_info(p, first + ti * stride, w);
                rv = r < WPT * WTOK && wi.valid;
                const int R = 36 * w + landed_to_token(l < WTOK ? l : 0, wi.xs, wi.ys);
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) qoff[pc] = stage_off<32>(R, pc);
                const int tt = R & 63;
                vcol = (uint32_t)((R >> 6) * Cfg::VCH + (tt & 7) * 2);
#pragma unroll
                for (int j = 0; j < 8; ++j) vsw[j] = (uint32_t)(((tt >> 3) ^ j) << 4);
            }
            const bool tle = warp == 8 && lane == 0;
            if (tle) TTL(1, 10, u);
            mbar_wait((d_full + 8 * (b)), (gu >> 1) & 1);
            tc_fence_after();
            if (tle) TTL(1, 11, u);
            const uint32_t bia = sBias + 4 * (u * UN);
            const uint32_t par_e = ((gu >> 1) & 1) ^ 1;
            // ---- q and k: 32 columns each -> one 64-byte row of the 64B-swizzled operand
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                uint32_t acc[2][16];
                tmem_ld16(tlane + m * 32, acc[0]);
                tmem_ld16(tlane + m * 32 + 16, acc[1]);
                tmem_ld_wait();
                if (tle) TTL(1, 14, m);
                if (m == 0) mbar_wait((qk_empty + 8 * (b)), par_e);
                if (rv) {
                    const uint32_t base = m == 0 ? qb : kb;
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        const uint32_t* a = &acc[pc >> 1][(pc & 1) * 8];
                        const float4 b0 = lds128(bia + (m * 32 + pc * 8) * 4), b1 = lds128(bia + (m * 32 + pc * 8 + 4) * 4);
                        uint4 o;
                        o.x = pack_h2(__uint_as_float(a[0]) + b0.x, __uint_as_float(a[1]) + b0.y);
                        o.y = pack_h2(__uint_as_float(a[2]) + b0.z, __uint_as_float(a[3]) + b0.w);
                        o.z = pack_h2(__uint_as_float(a[4]) + b1.x, __uint_as_float(a[5]) + b1.y);
                        o.w = pack_h2(__uint_as_float(a[6]) + b1.z, __uint_as_float(a[7]) + b1.w);
                        sts128(base + qoff[pc], o);
                    }
                }
            }
            if (tle) TTL(1, 15, u);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive((qk_full + 8 * (b)));
            if (tle) TTL(1, 12, u);
            // ---- v: transposed, V^T[head][dd][token R]
            {
                uint32_t acc[2][16];
                tmem_ld16(tlane + 64, acc[0]);
                tmem_ld16(tlane + 80, acc[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive((d_empty + 8 * (b)));       // every column of D has been pulled by this warp
                if (tle) TTL(1, 16, u);
                mbar_wait((v_empty + 8 * (b)), par_e);
                if (tle) TTL(1, 17, u);
                if (rv) {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 bv = lds128(bia + (64 + c4 * 4) * 4);
                        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = c4 * 4 + e, hh = c / D, dd = c % D;
                            const float v = __uint_as_float(acc[c >> 4][c & 15]) + bb[e];
                            sts16(vb + vcol + (hh * Cfg::VTB + dd * 128) + vsw[dd & 7], __float2half_rn(v));
                        }
                    }
                }
            }
            if (tle) TTL(1, 18, u);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive((v_full + 8 * (b)));
            if (tle) TTL(1, 13, u);
        }
    } else {
        // ===================== softmax warpgroups (warps 16-19: even heads, 20-23: odd heads) + O epilogue =====================
        const int kk = (warp - 16) >> 2, q = warp & 3;
        const int r = q * 32 + lane;
        const int w = r < 36 ? 0 : (r < 72 ? 1 : 2);
        const int i = r < WPT * WTOK ? r - 36 * w : 0;         // token of the window; query (yq, xq)
        const int yq = i / WS, xq = i - yq * WS;
        const int wA = (32 * q) / 36, wB = (32 * q + 31) / 36 > 2 ? 2 : (32 * q + 31) / 36;   // windows of the first / last row of this warp
        const bool mixed = wA != wB, useB = w != wA;
        const uint32_t tslot = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::TM_S + kk * Cfg::TM_SSTRIDE;
        const float scale = ((D == 16) ? 0.25f : 0.17677669529663687f) * LOG2E;   // d^-0.5 (:187) * log2(e)
        uint32_t paddr[9];                                      // the nine 8-byte pieces of this row's 36 probabilities
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            const int cb = 72 * w + 8 * m, within = cb & 127;   // byte column of P
            paddr[m] = sP + kk * Cfg::PB + r * 128 + (cb >> 7) * Cfg::PCH + ((((within >> 4) ^ (r & 7))) << 4) + (within & 15);
        }
        const int rbase = (yq + WS - 1) * (2 * WS - 1) + xq + WS - 1;   // relative position index of key (0, 0)
        int cur_ti = -1;
        bool rv = false, anyb = false;
        float madd[4] = {0.f, 0.f, 0.f, 0.f};
        size_t tok = 0;
        // pending O epilogue (one head late: its PV runs under the softmax of the next head of this warpgroup)
        bool o_pending = false, o_rv = false;
        __half* o_dst = nullptr;
        uint32_t o_par = 0;
        const bool tls = q == 0 && lane == 0;
        const int ttr = 2 + kk;
        auto o_epilogue = [&]() {
            mbar_wait((o_full + 8 * (kk)), o_par);
            tc_fence_after();
            if (tls) TTL(ttr, 28, 0);
            uint32_t o[D], os[4];
#pragma unroll
            for (int c = 0; c < D / 16; ++c) tmem_ld16(tslot + Cfg::TM_O + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&o[c * 16]));
            tmem_ld4(tslot + Cfg::TM_O + D, os);
            tmem_ld_wait();
            tc_fence_before();
            if (tls) TTL(ttr, 29, 0);
            if (o_rv) {
                const float inv = __fdividef(1.f, __uint_as_float(os[0]));
#pragma unroll
                for (int c = 0; c < D / 8; ++c) {
                    uint4 v;
                    v.x = pack_h2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
                    v.y = pack_h2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
                    v.z = pack_h2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
                    v.w = pack_h2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
                    reinterpret_cast<uint4*>(o_dst)[c] = v;
                }
            }
        };
        for (int gh = kk;; gh += 2) {
            const bool live = gh < total_heads;     // the pass after the last head only drains the pending O epilogue
            const int ti = gh / HEADS, hd = gh - ti * HEADS;
            uint32_t pk[WTOK / 2];
            if (live) {
            if (ti != cur_ti) {
                cur_ti = ti;
                const WinInfo wi = win_info(p, first + ti * stride, w);
                rv = r < WPT * WTOK && wi.valid;
                // shift mask (:193-209): only the last window row / column mixes regions; key class = (ky < 3) * 2 + (kx < 3)
#pragma unroll
                for (int cls = 0; cls < 4; ++cls) {
                    const bool ka = (cls & 2) != 0, kb2 = (cls & 1) != 0;
                    const bool masked = (wi.ys && ((yq < 3) != ka)) || (wi.xs && ((xq < 3) != kb2));
                    madd[cls] = masked ? -100.0f * LOG2E : 0.f;
                }
                anyb = __any_sync(0xffffffffu, wi.xs || wi.ys);
                int y = wi.wy * WS + yq + p.shift, x = wi.wx * WS + xq + p.shift;
                if (y >= p.H) y -= p.H;
                if (x >= p.W) x -= p.W;
                tok = ((size_t)wi.b * p.H + y) * p.W + x;
            }
            if (tls) TTL(ttr, 20, hd);
            mbar_wait((s_full + 8 * (kk)), (gh >> 1) & 1);
            tc_fence_after();
            if (tls) TTL(ttr, 21, hd);
            // ---- the 36 columns of this row's window (36 w ..): warps that straddle two windows read both and select
            uint32_t sa[WTOK];
#pragma unroll
            for (int m = 0; m < 9; ++m) tmem_ld4(tslot + 36 * wA + 4 * m, *reinterpret_cast<uint32_t(*)[4]>(&sa[4 * m]));
            tmem_ld_wait();
            if (mixed) {
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    uint32_t sb[12];
#pragma unroll
                    for (int m = 0; m < 3; ++m) tmem_ld4(tslot + 36 * wB + 12 * part + 4 * m, *reinterpret_cast<uint32_t(*)[4]>(&sb[4 * m]));
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 12; ++c) sa[12 * part + c] = useB ? sb[c] : sa[12 * part + c];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive((s_free + 8 * (kk)));     // S(h + 2) may overwrite the S columns (O has its own)
            if (tls) TTL(ttr, 22, hd);
            float s[WTOK];
            const uint32_t bt = sBT + 4 * (hd * RT + rbase);
#pragma unroll
            for (int c = 0; c < WTOK; ++c) s[c] = fmaf(__uint_as_float(sa[c]), scale, lds32(bt - 4 * ((c / WS) * (2 * WS - 1) + (c % WS))));
            if (anyb) {
#pragma unroll
                for (int c = 0; c < WTOK; ++c) s[c] += madd[((c / WS < 3) ? 2 : 0) | ((c % WS < 3) ? 1 : 0)];
            }
            float mx = s[0];
#pragma unroll
            for (int c = 1; c < WTOK; ++c) mx = fmaxf(mx, s[c]);
#pragma unroll
            for (int c = 0; c < WTOK / 2; ++c) pk[c] = pack_h2(ex2f(s[2 * c] - mx), ex2f(s[2 * c + 1] - mx));
            if (tls) TTL(ttr, 23, hd);
            }
            if (o_pending) o_epilogue();                 // head gh - 2: its PV has had a whole softmax to finish
            if (!live) break;
            if (tls) TTL(ttr, 24, hd);
            mbar_wait((p_empty + 8 * (kk)), ((gh >> 1) & 1) ^ 1);
            if (tls) TTL(ttr, 26, hd);
            if (rv) {
#pragma unroll
                for (int m = 0; m < 9; ++m) sts64(paddr[m], pk[2 * m], pk[2 * m + 1]);
            }
            if (tls) TTL(ttr, 27, hd);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive((p_full + 8 * (kk)));
            if (tls) TTL(ttr, 25, hd);
            o_pending = true; o_rv = rv; o_dst = p.att + tok * C + hd * D; o_par = (gh >> 1) & 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
#undef TTL
}

template <int C>
static int launch_tc(cudaStream_t st, const FusedAttn& f) {
    using Cfg = TcCfg<C>;
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)f.W * C * 2, (cuuint64_t)f.H * f.W * C * 2};
        cuuint32_t b66[4] = {64, 6, 6, 1}, b36[4] = {64, 3, 6, 1}, b63[4] = {64, 6, 3, 1}, b33[4] = {64, 3, 3, 1};
        if (encode(&maps.x66, f.x, 4, dims, strides, b66, 128)) return 1;
        if (encode(&maps.x36, f.x, 4, dims, strides, b36, 128)) return 1;
        if (encode(&maps.x63, f.x, 4, dims, strides, b63, 128)) return 1;
        if (encode(&maps.x33, f.x, 4, dims, strides, b33, 128)) return 1;
        cuuint64_t wd[2] = {(cuuint64_t)C, (cuuint64_t)3 * C};
        cuuint64_t wst[1] = {(cuuint64_t)C * 2};
        cuuint32_t wb[2] = {64, (cuuint32_t)UN};
        if (encode(&maps.w, f.wqkv_tc, 2, wd, wst, wb, 128)) return 1;
    }
    TcParams p;
    p.B = f.B; p.H = f.H; p.W = f.W;
    p.shift = (WS >= f.H || WS >= f.W) ? 0 : f.shift;   // torchvision :151-155: no shift when the window covers the map
    p.nww = f.W / WS; p.nwh = f.H / WS;
    p.tpi = (p.nww * p.nwh + WPT - 1) / WPT;
    p.tiles = f.B * p.tpi;
    p.bqkv = f.bqkv_tc; p.bias_tab = f.bias_tab_tc; p.att = f.att;
    p.tl = g_timeline;
    if (ensure_dyn_smem((const void*)swin_attn_tc_kernel<C>, Cfg::SMEM)) return 1;
    int grid = device_sm_count();
    if (grid > p.tiles) grid = p.tiles;
    const double T = (double)f.B * f.H * f.W;
    ProfScope ps(st, PC_FUSED_ATTN, 2.0 * T * C * 3.0 * C + 4.0 * T * WTOK * C, T * C * 2.0 + 3.0 * C * C * 2.0, T * C * 2.0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NB_CUDA(cudaLaunchKernelEx(&cfg, swin_attn_tc_kernel<C>, maps, p));
    NB_LAUNCHED();
    return 0;
}

// reference row order (q | k | v, head-major inside each) -> packed (unit, {q,k,v}, head-in-unit, d)
__global__ void pack_qkv_tc_kernel(const __half* __restrict__ w, const float* __restrict__ b, __half* __restrict__ wp, float* __restrict__ bp, int C) {
    const int pr = blockIdx.x;
    const int src = swin_attn_tc_src_row(pr, C);
    for (int k = threadIdx.x; k < C; k += blockDim.x) wp[(size_t)pr * C + k] = w[(size_t)src * C + k];
    if (threadIdx.x == 0) bp[pr] = b[src];
}

}  // namespace

int swin_attn_tc(cudaStream_t st, const FusedAttn& f) {
    NB_CHECK(f.x && f.att && f.wqkv_tc && f.bqkv_tc && f.bias_tab_tc, "null pointer");
    NB_CHECK(f.B > 0 && f.H > 0 && f.W > 0, "empty input");
    NB_CHECK(f.H % WS == 0 && f.W % WS == 0, "feature map must be a multiple of the 6x6 window");
    NB_CHECK(f.C == 96 || f.C == 192, "fused window attention supports C = 96 (d = 16) and C = 192 (d = 32)");
    NB_CHECK(f.shift == 0 || f.shift == 3, "shift must be 0 or window/2");
    return f.C == 192 ? launch_tc<192>(st, f) : launch_tc<96>(st, f);
}

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_swin_attn_tc_f16(const void* x, const void* wqkv, const float* bqkv, const float* bias_table, void* att,
                                      int B, int H, int W, int C, int shift, void* stream) {
    NB_CHECK(x && wqkv && bqkv && bias_table && att, "null pointer");
    NB_CHECK(C == 96 || C == 192, "C must be 96 or 192");
    cudaStream_t st = (cudaStream_t)stream;
    __half* wp = nullptr;
    float* bp = nullptr;
    NB_CUDA(cudaMallocAsync((void**)&wp, (size_t)3 * C * C * 2, st));
    NB_CUDA(cudaMallocAsync((void**)&bp, (size_t)3 * C * 4, st));
    pack_qkv_tc_kernel<<<3 * C, 96, 0, st>>>((const __half*)wqkv, bqkv, wp, bp, C);
    FusedAttn f;
    f.x = (const __half*)x; f.att = (__half*)att; f.B = B; f.H = H; f.W = W; f.C = C; f.shift = shift;
    f.wqkv_tc = wp; f.bqkv_tc = bp; f.bias_tab_tc = bias_table;
    const int rc = swin_attn_tc(st, f);
    cudaFreeAsync(wp, st); cudaFreeAsync(bp, st);
    return rc;
}
