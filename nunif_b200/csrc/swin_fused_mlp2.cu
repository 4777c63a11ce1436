// Fused Swin MLP, DUAL-PIPELINE variant: two independent half-SM pipelines in one CTA (sm_100a).
//
//   x <- x1 + fc2(gelu(fc1(x1))),   x1 = x + att.Wp^T + bp  (PROJ, C = 96)   |   x1 = x  (C = 192; its proj runs as a GEMM)
//
// Why: a clock64 timeline of the one-CTA-per-SM kernel (swin_fused_mlp.cu; profiles/r2/fused_timeline_192.txt) shows
// 23.5k cycles per 128-token tile of which the epilogue warps are busy for ~11.5k: with ONE tile in flight every stage waits
// for the previous one (x load -> G0 -> E0 -> G1 -> E1 -> G2 -> E2 -> store).  Shared memory does not allow a second tile
// per pipeline, but the pipeline can be shrunk to half an SM - 8 epilogue warps, 256 TMEM columns, <= 112 KB shared memory.
// Two such pipelines run side by side and fill each other's bubbles.  They live in ONE CTA (24 warps; warp / 12 selects the
// pipeline, each with its own barriers, shared-memory half and TMEM half): as two CTAs per SM the C = 192 configuration does
// not fit (each CTA pays its own 1 KB reservation and alignment slack; measured occupancy 1).
//
// Per pipeline (12 warps):
//   warps 0,10,11  TMA producers (bulk ops of one warp serialise at ~0.34 us each; fixed ownership: ring slot s -> producer s,
//                  activation K-chunk kc -> producer kc).  Weights are PRE-PACKED chunk-major ([K/BK][rows][BK]) so that all
//                  K-chunks of one GEMM chunk arrive with a single 3-D box op into one ring stage.
//   warp 1         tcgen05.mma issuer:  G0 D0 = att.Wp^T (PROJ) ; G1 D1 = x1.W1[j]^T (ONE accumulator, HCH columns) ;
//                  G2 D2 += H[j].W2[:, j]^T.  Order G0, G1(0), { G1(j+1) after E1(j) drained D1, G2(j) }.
//   warps 2-9      8 epilogue warps (4 TMEM lane groups x 2 column halves): E0 / E1 (GELU) / E2 as in swin_fused_mlp.cu;
//                  E1 hands D1 back right after its TMEM loads, before the GELU math.
#include "gemm_tcgen05.cuh"
#include "swin_fused.h"
#include "tmap.h"
#include <cstdlib>

namespace nb200 {

namespace {

__device__ __forceinline__ void tmem_ld8b(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tma_store_2db(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3db(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int C>
struct F2Cfg {
    static constexpr bool PROJ = (C == 96);                 // C = 192: shared memory has no room for the att tile
    static constexpr int BK = (C % 64 == 0) ? 64 : 32;
    static constexpr int SW = BK * 2;
    static constexpr int KCH = C / BK;                      // 3
    static constexpr int HID = 2 * C;
    static constexpr int HCH = (C == 192) ? 64 : 96;        // hidden units per D1 accumulator
    static constexpr int NCH = HID / HCH;                   // 6 | 2
    static constexpr int HSUB = HCH / BK;                   // 1 | 3
    static constexpr int XCH = 128 * BK * 2;
    static constexpr int XB = KCH * XCH;
    static constexpr int G1B = KCH * HCH * BK * 2;          // one G1 chunk, all K-chunks: 24576 | 18432
    static constexpr int G2B = C * BK * 2;                  // one G2 sub-chunk: 24576 | 6144
    static constexpr int G0B = KCH * C * BK * 2;            // proj, all K-chunks (C = 96): 18432
    static constexpr int WST = G1B > G2B ? G1B : G2B;
    static constexpr int STAGES = 2;
    static constexpr int NHB = (C == 192) ? 1 : 2;          // H sub-chunk ring
    static constexpr int D1COL = 0, D2COL = HCH;            // TMEM: D1 | D2 ; D0 aliases D1 (PROJ: C <= HCH)
    static constexpr int TMEM_COLS = 256;
    static_assert(HCH + C <= TMEM_COLS, "TMEM budget of half an SM");
    static_assert(!PROJ || C <= HCH, "D0 must fit in the D1 columns");
    static_assert(KCH == 3, "one activation K-chunk per producer warp");
    static constexpr int EPW = 8, PRODUCERS = 3;
    static constexpr int WPH = 2 + EPW + PRODUCERS - 1;     // warps per pipeline (12)
    static constexpr int THREADS = 2 * 32 * WPH;            // two pipelines per CTA
    static constexpr int DATA = XB * (PROJ ? 2 : 1) + NHB * XCH + STAGES * WST;
    static constexpr int SMEM = 2 * DATA + 1024 + 512;      // alignment slack + 2 x 256 B of barriers
    static_assert(SMEM <= 232448, "shared memory of one SM");
};

struct Fused2Maps {
    CUtensorMap x, att, wp, w1, w2;
};
struct Fused2Params {
    int tiles;
    const float *bp, *b1, *b2;
};

// accumulator columns [col0, col0 + NCOLS) of this warp's 32 rows: + bias + residual (swizzled smem tile) -> fp16 in place
template <int C, int NCOLS>
__device__ __forceinline__ void epi2_residual(uint32_t tcol, int col0, uint8_t* sX, const float* __restrict__ bias, int r) {
    using Cfg = F2Cfg<C>;
    constexpr int BK = Cfg::BK, GP = 4;
    static_assert(NCOLS % (8 * GP) == 0 || NCOLS == 48, "piece groups");
    constexpr int NP = NCOLS / 8;
#pragma unroll 1
    for (int g0 = 0; g0 < NP; g0 += GP) {
        uint32_t acc[GP][8];
#pragma unroll
        for (int i = 0; i < GP; ++i)
            if (g0 + i < NP) tmem_ld8b(tcol + (uint32_t)(col0 + (g0 + i) * 8), acc[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < GP; ++i) {
            if (g0 + i >= NP) break;
            const int c0 = col0 + (g0 + i) * 8;
            const int ch = c0 / BK, jj = (c0 % BK) >> 3;
            uint4* ptr = reinterpret_cast<uint4*>(sX + ch * Cfg::XCH + stage_off<BK>(r, jj));
            const uint4 rv = *ptr;
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            __align__(16) __half2 o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 rr = __half22float2(rh[k]);
                o[k] = __floats2half2_rn(__uint_as_float(acc[i][2 * k]) + bb[2 * k] + rr.x,
                                         __uint_as_float(acc[i][2 * k + 1]) + bb[2 * k + 1] + rr.y);
            }
            *ptr = *reinterpret_cast<const uint4*>(o);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(F2Cfg<C>::THREADS, 1) swin_mlp_fused2_kernel(const __grid_constant__ Fused2Maps maps,
                                                                               const __grid_constant__ Fused2Params p) {
    using Cfg = F2Cfg<C>;
    constexpr bool PROJ = Cfg::PROJ;
    constexpr int BK = Cfg::BK, SW = Cfg::SW, KCH = Cfg::KCH, HCH = Cfg::HCH, NCH = Cfg::NCH, HSUB = Cfg::HSUB;
    constexpr int XCH = Cfg::XCH, XB = Cfg::XB, WST = Cfg::WST, NHB = Cfg::NHB, S = Cfg::STAGES, EPW = Cfg::EPW;

    const int half = (threadIdx.x >> 5) / Cfg::WPH;          // pipeline 0 / 1
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem = smem0 + half * Cfg::DATA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem0 + 2 * Cfg::DATA + half * 256);
    uint8_t* sX = smem;
    uint8_t* sATT = sX + XB;
    uint8_t* sH = sATT + (PROJ ? XB : 0);
    uint8_t* sW = sH + NHB * XCH;
    uint64_t* w_full = bars;            // [S]
    uint64_t* w_empty = w_full + S;     // [S]
    uint64_t* h_full = w_empty + S;     // [2]
    uint64_t* h_empty = h_full + 2;     // [2]
    uint64_t* d1_full = h_empty + 2;
    uint64_t* d1_empty = d1_full + 1;
    uint64_t* att_full = d1_empty + 1;
    uint64_t* att_empty = att_full + 1;
    uint64_t* x_full = att_empty + 1;
    uint64_t* x_empty = x_full + 1;
    uint64_t* d0_full = x_empty + 1;
    uint64_t* x1_ready = d0_full + 1;
    uint64_t* d2_full = x1_ready + 1;

    const int warp = (threadIdx.x >> 5) - half * Cfg::WPH, lane = threadIdx.x & 31;   // warp index within the pipeline
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.x);
        tma_prefetch_desc(&maps.w1);
        tma_prefetch_desc(&maps.w2);
        if (PROJ) { tma_prefetch_desc(&maps.att); tma_prefetch_desc(&maps.wp); }
        for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&h_full[s], EPW); mbar_init(&h_empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(d1_empty, EPW);
        mbar_init(att_full, KCH); mbar_init(att_empty, 1);
        mbar_init(x_full, KCH); mbar_init(x_empty, KCH);
        mbar_init(d0_full, 1); mbar_init(x1_ready, EPW); mbar_init(d2_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 0) NB_PDL_TRIGGER();
    uint32_t* tmem_slot0 = reinterpret_cast<uint32_t*>(smem0 + 2 * Cfg::DATA + 248);   // pipeline 0's slot: one allocation per CTA
    if (warp == 1 && half == 0) tmem_alloc<2 * Cfg::TMEM_COLS>(tmem_slot0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot0 + (uint32_t)(half * Cfg::TMEM_COLS);
    const int first = 2 * blockIdx.x + half, stride = 2 * gridDim.x;   // each pipeline walks its own tile sequence

    if (warp == 0 || warp >= 2 + EPW) {
        // ===================== TMA producers =====================
        const int pid = warp == 0 ? 0 : warp - (2 + EPW) + 1;
        if (elect_one() && first < p.tiles) {
            int ws = 0;
            uint32_t wph = 0;
            // one op per GEMM chunk; ring slot ws is owned by producer ws (S == 2: producers 0 and 1)
            auto stage = [&](auto&& issue, uint32_t bytes) {
                if (ws == pid) {
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    mbar_expect_tx(&w_full[ws], bytes);
                    issue(&w_full[ws], sW + ws * WST);
                }
                if (++ws == S) { ws = 0; wph ^= 1; }
            };
            auto load_act = [&](const CUtensorMap* m, uint64_t* full, uint64_t* empty, uint8_t* dst, int row0, uint32_t par) {
                mbar_wait(empty, par ^ 1);
                mbar_expect_tx(full, XCH);
                tma_load_2d(m, full, dst + pid * XCH, pid * BK, row0);     // K-chunk pid
            };
            auto g0w = [&]() { stage([&](uint64_t* bar, uint8_t* dst) { tma_load_3db(&maps.wp, bar, dst, 0, 0, 0); }, Cfg::G0B); };
            auto g1w = [&](int j) { stage([&](uint64_t* bar, uint8_t* dst) { tma_load_3db(&maps.w1, bar, dst, 0, j * HCH, 0); }, Cfg::G1B); };
            auto g2w = [&](int j) {
                for (int s2 = 0; s2 < HSUB; ++s2)
                    stage([&](uint64_t* bar, uint8_t* dst) { tma_load_2d(&maps.w2, bar, dst, (j * HSUB + s2) * BK, 0); }, Cfg::G2B);
            };
            bool waited = false;
            uint32_t par = 0;
            for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
                const int row0 = tile * 128;
                if (PROJ && !waited) g0w();                           // weights are launch constants
                if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
                if (PROJ) {
                    load_act(&maps.att, att_full, att_empty, sATT, row0, par);
                    if (waited) g0w();
                }
                waited = true;
                load_act(&maps.x, x_full, x_empty, sX, row0, par);
                g1w(0);
                for (int j = 0; j < NCH; ++j) {
                    if (j + 1 < NCH) g1w(j + 1);
                    g2w(j);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_c = make_idesc_f16(C), idesc_h = make_idesc_f16(HCH);
        int ws = 0, hb = 0;
        uint32_t wph = 0, hph = 0, par = 0, d1n = 0;
        const uint32_t aX = smem_u32(sX), aATT = smem_u32(sATT), aH = smem_u32(sH), aW = smem_u32(sW);
        auto next_stage = [&]() { if (++ws == S) { ws = 0; wph ^= 1; } };
        for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
            if (PROJ) {
                mbar_wait(att_full, par);
                mbar_wait(&w_full[ws], wph);
                tc_fence_after();
                if (elect_one()) {
                    for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(tmem_base + Cfg::D1COL, make_kmajor_desc<SW>(aATT + kc * XCH + k * 32),
                                     make_kmajor_desc<SW>(aW + ws * WST + kc * (C * BK * 2) + k * 32), idesc_c, (kc > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&w_empty[ws]);
                    umma_commit(d0_full);
                    umma_commit(att_empty);
                }
                __syncwarp();
                next_stage();
                mbar_wait(x1_ready, par);
            } else {
                mbar_wait(x_full, par);
            }
            tc_fence_after();
            auto g1 = [&](int j) {
                // the single D1 accumulator: E1 of the previous chunk has pulled its values out of TMEM
                mbar_wait(d1_empty, (d1n & 1) ^ 1);
                ++d1n;
                mbar_wait(&w_full[ws], wph);
                tc_fence_after();
                if (elect_one()) {
                    for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(tmem_base + Cfg::D1COL, make_kmajor_desc<SW>(aX + kc * XCH + k * 32),
                                     make_kmajor_desc<SW>(aW + ws * WST + kc * (HCH * BK * 2) + k * 32), idesc_h, (kc > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&w_empty[ws]);
                    umma_commit(d1_full);
                }
                __syncwarp();
                next_stage();
            };
            auto g2 = [&](int j) {
                for (int s2 = 0; s2 < HSUB; ++s2) {
                    mbar_wait(&h_full[hb], hph);
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(tmem_base + Cfg::D2COL, make_kmajor_desc<SW>(aH + hb * XCH + k * 32),
                                     make_kmajor_desc<SW>(aW + ws * WST + k * 32), idesc_c, (j > 0 || s2 > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&w_empty[ws]);
                        umma_commit(&h_empty[hb]);
                        if (j == NCH - 1 && s2 == HSUB - 1) umma_commit(d2_full);
                    }
                    __syncwarp();
                    next_stage();
                    if (++hb == NHB) { hb = 0; hph ^= 1; }
                }
            };
            g1(0);
            for (int j = 0; j < NCH; ++j) {
                if (j + 1 < NCH) g1(j + 1);
                g2(j);
            }
        }
    } else {
        // ===================== epilogue warps 2..9 =====================
        const int q = (warp - 2) >> 2;        // column half
        const int g = warp & 3;               // TMEM lane group
        const int r = g * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(g * 32) << 16);
        int hb = 0;
        uint32_t hph = 0, par = 0, d1n = 0;
        for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
            mbar_wait(x_full, par);
            if (PROJ) {
                mbar_wait(d0_full, par);
                tc_fence_after();
                epi2_residual<C, C / 2>(tlane + Cfg::D1COL, q * (C / 2), sX, p.bp, r);
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(x1_ready);   // (D0 aliased D1: G1(0) is ordered behind this through x1_ready)
            }
#pragma unroll 1
            for (int j = 0; j < NCH; ++j) {
                mbar_wait(d1_full, d1n & 1);
                ++d1n;
                tc_fence_after();
                constexpr int CPW = BK / 2;         // columns per warp in a [128][BK] sub-chunk: 32 (BK=64) or 16 (BK=32)
#pragma unroll 1
                for (int s2 = 0; s2 < HSUB; ++s2) {
                    const int cl = s2 * BK + q * CPW;
                    const uint32_t tcol = tlane + Cfg::D1COL + (uint32_t)cl;
                    const float* bia = p.b1 + j * HCH + cl;
                    float v[CPW];
                    {
                        uint32_t acc[CPW / 8][8];
#pragma unroll
                        for (int i = 0; i < CPW / 8; ++i) tmem_ld8b(tcol + i * 8, acc[i]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < CPW / 8; ++i) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bia + i * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bia + i * 8 + 4));
                            v[i * 8 + 0] = __uint_as_float(acc[i][0]) + b0.x; v[i * 8 + 1] = __uint_as_float(acc[i][1]) + b0.y;
                            v[i * 8 + 2] = __uint_as_float(acc[i][2]) + b0.z; v[i * 8 + 3] = __uint_as_float(acc[i][3]) + b0.w;
                            v[i * 8 + 4] = __uint_as_float(acc[i][4]) + b1.x; v[i * 8 + 5] = __uint_as_float(acc[i][5]) + b1.y;
                            v[i * 8 + 6] = __uint_as_float(acc[i][6]) + b1.z; v[i * 8 + 7] = __uint_as_float(acc[i][7]) + b1.w;
                        }
                    }
                    if (s2 == HSUB - 1) {   // every column of D1 has been pulled by this warp: the accumulator may be overwritten
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d1_empty);
                    }
#pragma unroll
                    for (int k = 0; k < CPW; ++k) v[k] = gelu_erf(v[k]);
                    mbar_wait(&h_empty[hb], hph ^ 1);
                    uint8_t* hbuf = sH + hb * XCH;
#pragma unroll
                    for (int pc = 0; pc < CPW / 8; ++pc) {
                        __align__(16) __half2 o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = __floats2half2_rn(v[pc * 8 + 2 * k], v[pc * 8 + 2 * k + 1]);
                        *reinterpret_cast<uint4*>(hbuf + stage_off<BK>(r, (q * CPW) / 8 + pc)) = *reinterpret_cast<const uint4*>(o);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_full[hb]);
                    if (++hb == NHB) { hb = 0; hph ^= 1; }
                }
            }
            // ---- E2
            mbar_wait(d2_full, par);
            tc_fence_after();
            epi2_residual<C, C / 2>(tlane + Cfg::D2COL, q * (C / 2), sX, p.b2, r);
            tc_fence_before();
            fence_async_smem();
            asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
            if (warp < 2 + KCH && lane == 0) {
                const int kc = warp - 2;
                tma_store_2db(&maps.x, sX + kc * XCH, kc * BK, tile * 128);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(x_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1 && half == 0) {
        tc_fence_after();
        tmem_dealloc<2 * Cfg::TMEM_COLS>(tmem_base);
    }
}

int encode_rows2(CUtensorMap* m, const void* base, long long rows, int cols, int ld, int box_cols, int box_rows, int sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    return encode(m, base, 2, dims, strides, box, sw);
}
// chunk-major weights [K/bk][rows][bk]: 3-D view (bk, rows, K/bk), box = (bk, box_rows, K/bk) -> [K/bk][box_rows][bk] in smem
int encode_cm(CUtensorMap* m, const void* base, int rows, int K, int bk, int box_rows, int sw) {
    cuuint64_t dims[3] = {(cuuint64_t)bk, (cuuint64_t)rows, (cuuint64_t)(K / bk)};
    cuuint64_t strides[2] = {(cuuint64_t)bk * 2, (cuuint64_t)rows * bk * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, (cuuint32_t)(K / bk)};
    return encode(m, base, 3, dims, strides, box, sw);
}

template <int C>
int launch_mlp2(cudaStream_t st, const FusedMlp& f) {
    using Cfg = F2Cfg<C>;
    NB_CHECK(f.w1_cm, "chunk-major fc1 weights missing");
    NB_CHECK(!Cfg::PROJ || (f.att && f.wp_cm && f.bp), "proj operands missing");
    NB_CHECK(Cfg::PROJ || !f.att, "C = 192: the half-SM kernel has no proj stage");
    Fused2Maps maps;
    memset(&maps, 0, sizeof(maps));
    if (encode_rows2(&maps.x, f.x, f.T, C, C, Cfg::BK, 128, Cfg::SW)) return 1;
    if (Cfg::PROJ) {
        if (encode_rows2(&maps.att, f.att, f.T, C, C, Cfg::BK, 128, Cfg::SW)) return 1;
        if (encode_cm(&maps.wp, f.wp_cm, C, C, Cfg::BK, C, Cfg::SW)) return 1;
    }
    if (encode_cm(&maps.w1, f.w1_cm, 2 * C, C, Cfg::BK, Cfg::HCH, Cfg::SW)) return 1;
    if (encode_rows2(&maps.w2, f.w2, C, 2 * C, 2 * C, Cfg::BK, C, Cfg::SW)) return 1;
    Fused2Params p;
    p.tiles = (int)((f.T + 127) / 128);
    p.bp = f.bp; p.b1 = f.b1; p.b2 = f.b2;
    if (ensure_dyn_smem((const void*)swin_mlp_fused2_kernel<C>, Cfg::SMEM)) return 1;
    int grid = device_sm_count();
    if (grid > (p.tiles + 1) / 2) grid = (p.tiles + 1) / 2;
    static const bool dbg = getenv("NB200_DEBUG") != nullptr;
    if (dbg) {
        int occ = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, swin_mlp_fused2_kernel<C>, Cfg::THREADS, Cfg::SMEM);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, swin_mlp_fused2_kernel<C>);
        fprintf(stderr, "nb200: swin_mlp_fused2<%d>: occupancy %d CTAs/SM, %d regs, %zu B local, smem %d, grid %d, tiles %d\n", C, occ, fa.numRegs,
                fa.localSizeBytes, Cfg::SMEM, grid, p.tiles);
    }
    const double Td = (double)f.T;
    ProfScope ps(st, PC_FUSED_MLP, 2.0 * Td * C * C * (Cfg::PROJ ? 5.0 : 4.0),
                 Td * C * 2.0 * (Cfg::PROJ ? 2.0 : 1.0) + (Cfg::PROJ ? 5.0 : 4.0) * C * C * 2.0, Td * C * 2.0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NB_CUDA(cudaLaunchKernelEx(&cfg, swin_mlp_fused2_kernel<C>, maps, p));
    NB_LAUNCHED();
    return 0;
}

// [rows][K] row-major -> [K/bk][rows][bk]
__global__ void pack_chunk_major_kernel(const __half* __restrict__ w, __half* __restrict__ out, int rows, int K, int bk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    const int r = i / K, k = i % K;
    out[((size_t)(k / bk) * rows + r) * bk + (k % bk)] = w[i];
}

}  // namespace

int swin_mlp_fused2(cudaStream_t st, const FusedMlp& f) {
    NB_CHECK(f.x && f.w1_cm && f.b1 && f.w2 && f.b2, "null pointer");
    NB_CHECK(f.T > 0, "empty input");
    NB_CHECK(f.C == 96 || f.C == 192, "fused Swin MLP supports C = 96 and C = 192");
    return f.C == 192 ? launch_mlp2<192>(st, f) : launch_mlp2<96>(st, f);
}

int pack_chunk_major(cudaStream_t st, const __half* w, __half* out, int rows, int K, int bk) {
    pack_chunk_major_kernel<<<(rows * K + 255) / 256, 256, 0, st>>>(w, out, rows, K, bk);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb200
