// Fused Swin block tail on tcgen05 (sm_100a):   x <- x1 + fc2(gelu(fc1(x1))),  x1 = x + proj(att)
//
// torchvision SwinTransformerBlock.forward (swin_transformer.py:452-455) with Identity norms (swin_unet.py:16-17):
//     x = x + attn(x)            the `proj` Linear of the attention (:228) + residual are fused here (PROJ)
//     x = x + mlp(x)             Linear(C, 2C) - GELU - Linear(2C, C)  (:444, mlp_ratio 2: swin_unet.py:31)
// One persistent CTA per SM walks 128-token tiles.  Per tile NOTHING but `att`, `x` (in) and `x` (out) touches HBM:
// x1 and the 2C-wide hidden activation live in shared memory / TMEM (round 1 wrote and re-read x1, and the hidden
// tensor twice: 18*C*2 bytes per token and block -> 5*C*2 with the attention kernel of swin_fused_attn.cu).
//
//   warps 0,18,19  THREE TMA producer warps.  Measured on this part (profiles/r2/tma_inflight.json): bulk-tensor loads issued
//               by ONE warp are executed strictly one after the other at ~0.34 us per op whatever their size (4 KB .. 72 KB),
//               while ops from different warps overlap perfectly.  So every load op of the CTA's sequence - the K-chunks
//               of the att / x tiles and the weight K-chunks of the three GEMMs (one ring) - is dealt round-robin to the
//               three producer warps, and the three output chunk stores go out from three different epilogue warps
//   warp 1      tcgen05.mma issuer (single thread):
//                 G0  D0[128 x C]      = att  . Wp^T                 (PROJ)
//                 G1  D1[j][128 x HCH] = x1   . W1[j]^T              hidden chunk j (two accumulators, ping-pong)
//                 G2  D2[128 x C]     += H[j] . W2[:, j]^T           H = gelu(D1 + b1) in shared memory
//               issue order G0, G1(0), G1(1), G2(0), G1(2), G2(1), ... so the GELU epilogue of chunk j overlaps
//               the MMAs of chunk j+1.  D0 aliases the D1 columns (it is dead before G1(0) is issued).
//   warps 2-17  16 epilogue warps (4 TMEM lane groups x 4 column quarters):
//                 E0  x1 = D0 + bp + x      -> fp16, written IN PLACE over the x tile (A operand of G1, residual of E2)
//                 E1  H  = gelu(D1 + b1)    -> fp16 [128][BK] K-major swizzled sub-chunks (a 2-deep ring), A operand of G2
//                 E2  x  = D2 + b2 + x1     -> fp16 in place, then one TMA tensor store per K-chunk
// Operand tiles are [rows][BK] fp16 with the TMA/UMMA 128B (BK = 64, C = 192) or 64B (BK = 32, C = 96) swizzle.
#include "gemm_tcgen05.cuh"
#include "swin_fused.h"
#include "tmap.h"

namespace nb200 {

extern unsigned long long* g_timeline;   // gemm.cu (nb200_debug_timeline)
extern int g_tune[16];                    // gemm.cu (nb200_tune_set)

namespace {

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <int C>
struct FmCfg {
    static constexpr int BK = (C % 64 == 0) ? 64 : 32;      // K-chunk = one swizzle span
    static constexpr int SW = BK * 2;
    static constexpr int KCH = C / BK;                      // K-chunks of a C-wide operand
    static constexpr int HID = 2 * C;
    static constexpr int HCH = (C == 192) ? 128 : 96;       // hidden units per D1 accumulator
    static constexpr int NCH = HID / HCH;
    static constexpr int HSUB = HCH / BK;                   // [128][BK] sub-chunks per hidden chunk
    static constexpr int XCH = 128 * BK * 2;                // bytes of one [128][BK] activation chunk
    static constexpr int XB = KCH * XCH;                    // bytes of a [128][C] tile
    static constexpr int WROWS = C > HCH ? C : HCH;
    static constexpr int WST = WROWS * BK * 2;              // bytes of one weight ring stage
    static constexpr int NHB = 2;                           // H sub-chunk ring
    static constexpr int STAGES = (C == 192) ? 3 : 8;       // weight ring depth
    static constexpr int D1COL = 0, D2COL = 2 * HCH;        // TMEM columns: D1[0], D1[1] | D2 ; D0 aliases D1
    static_assert(2 * HCH + C <= 512, "TMEM budget");
    static_assert(C <= 2 * HCH, "D0 must fit in the D1 columns");
    static_assert(HID % HCH == 0 && HCH % BK == 0 && C % BK == 0, "chunking");
    static_assert(KCH == 3, "one activation K-chunk per producer warp");
    static constexpr int NBIAS = 4 * C;                     // bp | b1 (2C) | b2
    static constexpr size_t smem_bytes(bool proj) {
        return 1024 + (size_t)XB * (proj ? 2 : 1) + (size_t)NHB * XCH + (size_t)STAGES * WST + NBIAS * 4 + 512;
    }
};

struct FusedMlpMaps {
    CUtensorMap x, att, wp, w1, w2;
};
struct FusedMlpParams {
    int tiles;
    const float *bp, *b1, *b2;
    unsigned long long* tl;   // optional debug timeline (nb200_debug_timeline): CTA 0 records (tag << 56 | aux << 40 | clock)
};

constexpr int FM_PRODUCERS = 3;                          // warp 0 and warps 18, 19 (640 threads keep 96 registers per thread)
constexpr int FM_THREADS = 64 + 32 * 16 + 32 * (FM_PRODUCERS - 1);

// one accumulator [128 x C] + bias + residual(in smem) -> fp16 written in place over the residual tile
template <int C>
__device__ __forceinline__ void epi_residual_inplace(uint32_t tcol0, uint8_t* sX, const float* sBias, int r, int q) {
    using Cfg = FmCfg<C>;
    constexpr int BK = Cfg::BK;
    constexpr int NPW = C / 32;   // 8-column pieces per warp (this warp's quarter of the row)
    uint32_t acc[NPW][8];
#pragma unroll
    for (int i = 0; i < NPW; ++i) tmem_ld8(tcol0 + (uint32_t)((q * NPW + i) * 8), acc[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NPW; ++i) {
        const int c0 = (q * NPW + i) * 8;
        const int ch = c0 / BK, jj = (c0 % BK) >> 3;
        uint4* ptr = reinterpret_cast<uint4*>(sX + ch * Cfg::XCH + stage_off<BK>(r, jj));
        const uint4 rv = *ptr;
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0), b1 = *reinterpret_cast<const float4*>(sBias + c0 + 4);
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

template <int C, bool PROJ>
__global__ void __launch_bounds__(FM_THREADS, 1) swin_mlp_fused_kernel(const __grid_constant__ FusedMlpMaps maps,
                                                                       const __grid_constant__ FusedMlpParams p) {
    using Cfg = FmCfg<C>;
    constexpr int BK = Cfg::BK, SW = Cfg::SW, KCH = Cfg::KCH, HCH = Cfg::HCH, NCH = Cfg::NCH, HSUB = Cfg::HSUB;
    constexpr int XCH = Cfg::XCH, XB = Cfg::XB, WST = Cfg::WST, NHB = Cfg::NHB, S = Cfg::STAGES;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* sX = smem;                                    // x tile -> x1 -> output (in place)
    uint8_t* sATT = sX + XB;                               // att tile (PROJ)
    uint8_t* sH = sATT + (PROJ ? XB : 0);                  // NHB hidden sub-chunks
    uint8_t* sW = sH + NHB * XCH;                          // weight ring
    float* sBias = reinterpret_cast<float*>(sW + S * WST); // bp[C] | b1[2C] | b2[C]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + Cfg::NBIAS);
    uint64_t* w_full = bars;            // [S]
    uint64_t* w_empty = w_full + S;     // [S]
    uint64_t* h_full = w_empty + S;     // [NHB]  16 epilogue warps -> MMA
    uint64_t* h_empty = h_full + NHB;   // [NHB]  MMA commit -> epilogue
    uint64_t* d1_full = h_empty + NHB;  // [2]
    uint64_t* att_full = d1_full + 2;
    uint64_t* att_empty = att_full + 1;
    uint64_t* x_full = att_empty + 1;
    uint64_t* x_empty = x_full + 1;
    uint64_t* d0_full = x_empty + 1;
    uint64_t* x1_ready = d0_full + 1;
    uint64_t* d2_full = x1_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // debug timeline: 8 role tracks of 2048 events (producers 0-2, MMA, epilogue warp 2)
    unsigned long long* tlb = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    int tli = 0;
#define FTL(track, tag, aux) do { if (tlb && tli < 2048) { tlb[(track) * 2048 + tli] = ((unsigned long long)(tag) << 56) | ((unsigned long long)((aux) & 0xffff) << 40) | (clock64() & 0xffffffffffull); ++tli; } } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.x);
        tma_prefetch_desc(&maps.w1);
        tma_prefetch_desc(&maps.w2);
        if (PROJ) { tma_prefetch_desc(&maps.att); tma_prefetch_desc(&maps.wp); }
        for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < NHB; ++s) { mbar_init(&h_full[s], 16); mbar_init(&h_empty[s], 1); }
        mbar_init(&d1_full[0], 1); mbar_init(&d1_full[1], 1);
        mbar_init(att_full, KCH); mbar_init(att_empty, 1);
        mbar_init(x_full, KCH); mbar_init(x_empty, KCH);
        mbar_init(d0_full, 1); mbar_init(x1_ready, 16); mbar_init(d2_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 0) NB_PDL_TRIGGER();
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < Cfg::NBIAS; i += FM_THREADS) {
        float v;
        if (i < C) v = PROJ ? __ldg(p.bp + i) : 0.f;
        else if (i < 3 * C) v = __ldg(p.b1 + (i - C));
        else v = __ldg(p.b2 + (i - 3 * C));
        sBias[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.x, stride = gridDim.x;

    if (warp == 0 || warp >= 18) {
        // ===================== TMA producers (one thread in each of 3 warps) =====================
        // Fixed ownership (mbarrier waits are 1-bit phase parities: a thread must see EVERY phase of a barrier it waits on, so a
        // ring slot is always refilled by the same producer): weight ring slot s belongs to producer s % 3, K-chunk kc of the
        // att / x tiles to producer kc (KCH == 3).  Every producer walks the whole op sequence to keep slots and phases.
        const int pid = warp == 0 ? 0 : warp - 17;
        if (elect_one() && first < p.tiles) {
            int ws = 0, wown = 0;            // ring slot and its owner (ws % FM_PRODUCERS, kept incrementally)
            uint32_t wph = 0;
            auto load_w = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
                if (wown == pid) {
                    FTL(pid, 1, ws);
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    FTL(pid, 2, ws);
                    mbar_expect_tx(&w_full[ws], bytes);
                    tma_load_2d(m, &w_full[ws], sW + ws * WST, c0, c1);
                }
                if (++wown == FM_PRODUCERS) wown = 0;
                if (++ws == S) { ws = 0; wown = 0; wph ^= 1; }
            };
            auto load_act = [&](const CUtensorMap* m, uint64_t* full, uint64_t* empty, uint8_t* dst, int row0, uint32_t par) {
                const int kc = pid;          // static_assert(KCH == FM_PRODUCERS)
                FTL(pid, 3, empty == x_empty ? 1 : 0);
                mbar_wait(empty, par ^ 1);
                FTL(pid, 4, empty == x_empty ? 1 : 0);
                mbar_expect_tx(full, XCH);
                tma_load_2d(m, full, dst + kc * XCH, kc * BK, row0);
            };
            auto g1w = [&](int j) {
                for (int kc = 0; kc < KCH; ++kc) load_w(&maps.w1, kc * BK, j * HCH, HCH * BK * 2);
            };
            auto g2w = [&](int j) {
                for (int s2 = 0; s2 < HSUB; ++s2) load_w(&maps.w2, (j * HSUB + s2) * BK, 0, C * BK * 2);
            };
            bool waited = false;
            uint32_t par = 0;
            for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
                const int row0 = tile * 128;
                // weights are launch constants: the first tile's proj chunks go out before the grid dependency resolves
                if (PROJ && !waited) for (int kc = 0; kc < KCH; ++kc) load_w(&maps.wp, kc * BK, 0, C * BK * 2);
                if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
                if (PROJ) {
                    load_act(&maps.att, att_full, att_empty, sATT, row0, par);
                    if (waited) for (int kc = 0; kc < KCH; ++kc) load_w(&maps.wp, kc * BK, 0, C * BK * 2);
                }
                waited = true;
                load_act(&maps.x, x_full, x_empty, sX, row0, par);
                g1w(0);
                if (NCH > 1) g1w(1);
                for (int j = 0; j < NCH; ++j) {
                    g2w(j);
                    if (j + 2 < NCH) g1w(j + 2);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_c = make_idesc_f16(C), idesc_h = make_idesc_f16(HCH);
        int ws = 0, hb = 0;
        uint32_t wph = 0, hph = 0, par = 0;
        const uint32_t aX = smem_u32(sX), aATT = smem_u32(sATT), aH = smem_u32(sH), aW = smem_u32(sW);
        for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
            if (PROJ) {
                if (lane == 0) FTL(3, 10, 0);
                mbar_wait(att_full, par);
                tc_fence_after();
                for (int kc = 0; kc < KCH; ++kc) {
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    if (lane == 0) FTL(3, 11, kc);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(tmem_base + Cfg::D1COL, make_kmajor_desc<SW>(aATT + kc * XCH + k * 32),
                                     make_kmajor_desc<SW>(aW + ws * WST + k * 32), idesc_c, (kc > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&w_empty[ws]);
                        if (kc == KCH - 1) { umma_commit(d0_full); umma_commit(att_empty); }
                    }
                    __syncwarp();
                    if (++ws == S) { ws = 0; wph ^= 1; }
                }
                if (lane == 0) FTL(3, 12, 0);
                mbar_wait(x1_ready, par);
                if (lane == 0) FTL(3, 13, 0);
            } else {
                mbar_wait(x_full, par);
            }
            tc_fence_after();
            auto g1 = [&](int j) {
                const uint32_t td = tmem_base + Cfg::D1COL + (uint32_t)((j & 1) * HCH);
                for (int kc = 0; kc < KCH; ++kc) {
                    if (lane == 0) FTL(3, 20, j);
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    if (lane == 0) FTL(3, 21, j);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(td, make_kmajor_desc<SW>(aX + kc * XCH + k * 32), make_kmajor_desc<SW>(aW + ws * WST + k * 32),
                                     idesc_h, (kc > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&w_empty[ws]);
                        if (kc == KCH - 1) umma_commit(&d1_full[j & 1]);
                    }
                    __syncwarp();
                    if (++ws == S) { ws = 0; wph ^= 1; }
                }
            };
            auto g2 = [&](int j) {
                for (int s2 = 0; s2 < HSUB; ++s2) {
                    if (lane == 0) FTL(3, 30, j);
                    mbar_wait(&h_full[hb], hph);
                    if (lane == 0) FTL(3, 31, j);
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    if (lane == 0) FTL(3, 32, j);
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
                    if (++ws == S) { ws = 0; wph ^= 1; }
                    if (++hb == NHB) { hb = 0; hph ^= 1; }
                }
            };
            g1(0);
            if (NCH > 1) g1(1);
            for (int j = 0; j < NCH; ++j) {
                g2(j);
                if (j + 2 < NCH) g1(j + 2);
            }
        }
    } else if (warp < 18) {
        // ===================== epilogue warps 2..17 =====================
        const int q = (warp - 2) >> 2;        // column quarter
        const int g = warp & 3;               // TMEM lane group this warp may read
        const int r = g * 32 + lane;          // accumulator row == token within the tile
        const uint32_t tlane = tmem_base + ((uint32_t)(g * 32) << 16);
        int hb = 0;
        uint32_t hph = 0, par = 0, d1use[2] = {0, 0};
        for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
            const bool tle = warp == 2 && lane == 0;
            if (tle) FTL(4, 40, 0);
            mbar_wait(x_full, par);
            if (tle) FTL(4, 41, 0);
            if (PROJ) {
                // ---- E0: x1 = att.Wp^T + bp + x, in place
                mbar_wait(d0_full, par);
                if (tle) FTL(4, 42, 0);
                tc_fence_after();
                epi_residual_inplace<C>(tlane + Cfg::D1COL, sX, sBias, r, q);
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(x1_ready);
                if (tle) FTL(4, 43, 0);
            }
            // ---- E1: hidden chunks
#pragma unroll 1
            for (int j = 0; j < NCH; ++j) {
                const int b = j & 1;
                if (tle) FTL(4, 50, j);
                mbar_wait(&d1_full[b], d1use[b] & 1);
                ++d1use[b];
                tc_fence_after();
                if (tle) FTL(4, 51, j);
#pragma unroll 1
                for (int s2 = 0; s2 < HSUB; ++s2) {
                    constexpr int CPW = BK / 4;     // columns per warp in a sub-chunk: 16 (BK=64) or 8 (BK=32)
                    const int cl = s2 * BK + q * CPW;   // column within the hidden chunk
                    const uint32_t tcol = tlane + Cfg::D1COL + (uint32_t)(b * HCH + cl);
                    const float* bia = sBias + C + j * HCH + cl;
                    uint8_t* hbuf = sH + hb * XCH;
                    float v[CPW];
                    if (CPW == 16) {
                        uint32_t acc[16];
                        tmem_ld16(tcol, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(acc[k]) + bia[k];
                    } else {
                        uint32_t acc[8];
                        tmem_ld8(tcol, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(acc[k]) + bia[k];
                    }
#pragma unroll
                    for (int k = 0; k < CPW; ++k) v[k] = gelu_erf(v[k]);
                    if (tle) FTL(4, 52, s2);
                    mbar_wait(&h_empty[hb], hph ^ 1);     // the G2 MMAs that read this ring slot have retired
                    if (tle) FTL(4, 53, s2);
#pragma unroll
                    for (int pc = 0; pc < CPW / 8; ++pc) {
                        __align__(16) __half2 o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = __floats2half2_rn(v[pc * 8 + 2 * k], v[pc * 8 + 2 * k + 1]);
                        *reinterpret_cast<uint4*>(hbuf + stage_off<BK>(r, (q * CPW) / 8 + pc)) = *reinterpret_cast<const uint4*>(o);
                    }
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_full[hb]);
                    if (++hb == NHB) { hb = 0; hph ^= 1; }
                }
            }
            // ---- E2: x = D2 + b2 + x1, in place, then TMA store
            if (tle) FTL(4, 60, 0);
            mbar_wait(d2_full, par);
            tc_fence_after();
            if (tle) FTL(4, 61, 0);
            epi_residual_inplace<C>(tlane + Cfg::D2COL, sX, sBias + 3 * C, r, q);
            tc_fence_before();
            fence_async_smem();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (warp < 2 + KCH && lane == 0) {
                // one K-chunk per storing warp (bulk ops of one warp run one after the other); rows >= T are clipped
                const int kc = warp - 2;
                tma_store_2d(&maps.x, sX + kc * XCH, kc * BK, tile * 128);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(x_empty);
                if (tle) FTL(4, 62, 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

static int encode_rows(CUtensorMap* m, const void* base, long long rows, int cols, int ld, int box_cols, int box_rows, int sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    return encode(m, base, 2, dims, strides, box, sw);
}

template <int C, bool PROJ>
static int launch_mlp(cudaStream_t st, const FusedMlp& f) {
    using Cfg = FmCfg<C>;
    FusedMlpMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (encode_rows(&maps.x, f.x, f.T, C, C, Cfg::BK, 128, Cfg::SW)) return 1;
    if (PROJ) {
        if (encode_rows(&maps.att, f.att, f.T, C, C, Cfg::BK, 128, Cfg::SW)) return 1;
        if (encode_rows(&maps.wp, f.wp, C, C, C, Cfg::BK, C, Cfg::SW)) return 1;
    }
    if (encode_rows(&maps.w1, f.w1, 2 * C, C, C, Cfg::BK, Cfg::HCH, Cfg::SW)) return 1;
    if (encode_rows(&maps.w2, f.w2, C, 2 * C, 2 * C, Cfg::BK, C, Cfg::SW)) return 1;
    FusedMlpParams p;
    p.tiles = (int)((f.T + 127) / 128);
    p.bp = f.bp; p.b1 = f.b1; p.b2 = f.b2;
    p.tl = g_timeline;
    const size_t smem = Cfg::smem_bytes(PROJ);
    if (ensure_dyn_smem((const void*)swin_mlp_fused_kernel<C, PROJ>, smem)) return 1;
    int grid = device_sm_count();
    if (grid > p.tiles) grid = p.tiles;
    const double Td = (double)f.T;
    ProfScope ps(st, PC_FUSED_MLP, 2.0 * Td * C * C * (PROJ ? 5.0 : 4.0), Td * C * 2.0 * (PROJ ? 2.0 : 1.0) + (PROJ ? 5.0 : 4.0) * C * C * 2.0,
                 Td * C * 2.0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(FM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NB_CUDA(cudaLaunchKernelEx(&cfg, swin_mlp_fused_kernel<C, PROJ>, maps, p));
    NB_LAUNCHED();
    return 0;
}

}  // namespace

int swin_mlp_fused(cudaStream_t st, const FusedMlp& f) {
    NB_CHECK(f.x && f.w1 && f.b1 && f.w2 && f.b2, "null pointer");
    NB_CHECK(f.T > 0, "empty input");
    NB_CHECK(f.C == 96 || f.C == 192, "fused Swin MLP supports C = 96 and C = 192");
    const bool proj = f.att != nullptr;
    if (proj) NB_CHECK(f.wp && f.bp, "proj weights missing");
    if (f.C == 192) return proj ? launch_mlp<192, true>(st, f) : launch_mlp<192, false>(st, f);
    return proj ? launch_mlp<96, true>(st, f) : launch_mlp<96, false>(st, f);
}

}  // namespace nb200

using namespace nb200;

// Low-level op for unit tests / micro-benchmarks (include/nunif_b200.h): x [T][C] fp16 is updated in place.
extern "C" int nb200_swin_mlp_fused_f16(void* x, const void* att, long long T, int C, const void* wp, const float* bp,
                                        const void* w1, const float* b1, const void* w2, const float* b2, void* stream) {
    FusedMlp f;
    f.x = (__half*)x; f.att = (const __half*)att; f.T = T; f.C = C;
    f.wp = (const __half*)wp; f.bp = bp; f.w1 = (const __half*)w1; f.b1 = b1; f.w2 = (const __half*)w2; f.b2 = b2;
    cudaStream_t st = (cudaStream_t)stream;
    // nb200_tune_set(11, 1): the one-CTA-per-SM kernel; default: the half-SM kernel (two CTAs per SM) where it applies
    // (C = 96 with or without att needs the proj operands: without att it falls back; C = 192 only without att)
    const bool half_sm = g_tune[11] == 0 && ((C == 192 && !att) || (C == 96 && att));
    if (!half_sm) return swin_mlp_fused(st, f);
    NB_CHECK(C == 96 || C == 192, "C must be 96 or 192");
    const int bk = C == 192 ? 64 : 32;
    __half *w1cm = nullptr, *wpcm = nullptr;
    NB_CUDA(cudaMallocAsync((void**)&w1cm, (size_t)2 * C * C * 2, st));
    int rc = pack_chunk_major(st, f.w1, w1cm, 2 * C, C, bk);
    if (!rc && att) {
        NB_CUDA(cudaMallocAsync((void**)&wpcm, (size_t)C * C * 2, st));
        rc = pack_chunk_major(st, f.wp, wpcm, C, C, bk);
    }
    f.w1_cm = w1cm; f.wp_cm = wpcm;
    if (!rc) rc = swin_mlp_fused2(st, f);
    cudaFreeAsync(w1cm, st);
    if (wpcm) cudaFreeAsync(wpcm, st);
    return rc;
}
