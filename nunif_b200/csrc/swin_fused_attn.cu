// Fused Swin block head (sm_100a):   att = shifted_window_attention_core( x . Wqkv^T + bqkv )
//
// torchvision shifted_window_attention (swin_transformer.py:166-221) up to, not including, the `proj` Linear (which is
// fused into swin_fused_mlp.cu): roll, 6x6 window partition, qkv Linear (:177), scale, QK^T, relative position bias
// (:190), shift mask (:193-209), softmax (:211), attn@v (:214), window reverse, un-roll.  q, k and v NEVER reach HBM:
// round 1 wrote them as three [T][C] planes and read them back (8*C*2 bytes per token); here the kernel reads x and
// writes att (2*C*2 bytes per token).
//
// One persistent CTA per SM walks tiles of 3 windows (108 tokens, padded to the 128 rows of one UMMA):
//   warpgroup 0 (24 regs/thread): warps 0 / 2 weight producers (even / odd ring slots), warp 1 tcgen05.mma issuer
//                (D[pair] (128 x 6d) = X (128 x C) . Wqkv[pair]^T, two TMEM accumulators, ping-pong), warp 3 activation
//                producer.  The window gather is done by the TMA unit - one (64 ch, 6, 6) box per window and K-chunk lands
//                the window's 36 tokens as 36 consecutive 128B-swizzled rows (windows that wrap around the rolled image use
//                3-wide single-row boxes).  Bulk-tensor ops of one warp execute one after the other (~0.34 us each,
//                profiles/r2/tma_inflight.json), hence one warp per stream.
//   warpgroups 1-2 (48 regs): 8 GEMM-epilogue warps: TMEM -> + bias -> fp16 -> q | k | v of the head pair as padded row-major
//                matrices in shared memory (double buffered)
//   warpgroups 3-6 (96 regs): 16 attention warps, one task each per head pair: (window, head, 16-row query tile); S = QK^T and
//                O = PV on mma.sync.m16n8k16 with the probabilities kept in registers (a 36x36xd problem per head is far
//                below a tcgen05 tile), base-2 softmax, bias table in shared memory, output rows staged over the dead q rows
//                and stored at their un-rolled token positions.  A clock64 timeline of the first version (9 attention warps,
//                profiles/r2/fused_timeline_attn_192.txt) showed these latency-bound warps as THE bottleneck; registers
//                for 16 of them come from setmaxnreg (the CTA launches at 72 registers x 896 threads).
#include "gemm_tcgen05.cuh"
#include "swin_fused.h"
#include "tmap.h"

namespace nb200 {

extern unsigned long long* g_timeline;   // gemm.cu (nb200_debug_timeline)

namespace {

constexpr int WS = 6, WTOK = 36, HEADS = 6, WPT = 3;   // 3 windows per tile
constexpr int BT_LD = 40;                              // bias table row: 36 keys + 4 masked pad columns
constexpr int BT_FLOATS = HEADS * WTOK * BT_LD;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x1_trans(uint32_t& r0, const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int C>
struct FaCfg {
    static constexpr int D = C / HEADS;                     // 32 or 16
    static constexpr int BK = 64, SW = 128;
    static constexpr int KCH = (C + 63) / 64;               // K-chunks; the last one may be partly out of bounds (zero filled)
    static constexpr int KLAST = (C - (KCH - 1) * 64) / 16; // UMMA k-steps in the last chunk
    static constexpr int NCHK = 6 * D;                      // GEMM N per head pair: q0 q1 k0 k1 v0 v1
    static constexpr int NPAIR = 3;
    static constexpr int LDH = 2 * D + 8;                   // padded row of the q/k/v matrices (conflict-free fragment loads)
    static constexpr int XCH = 128 * 128;                   // bytes of one [128][64] chunk
    static constexpr int XB = KCH * XCH;
    static constexpr int WST = NCHK * 128;                  // weight ring stage: [6d][64] fp16
    static constexpr int STAGES = (C == 192) ? 2 : 6;       // even: slot s is owned by producer s & 1
    static_assert(STAGES % 2 == 0, "ring slots are split between two producers");
    static constexpr int QKV_MAT = WPT * WTOK * LDH * 2;    // one matrix (q, k or v) of a head pair, all 3 windows
    static constexpr int QKV_BUF = 3 * QKV_MAT;
    static constexpr int TMEM_COLS = 2 * NCHK <= 256 ? 256 : 512;
    static constexpr int EPI_WARPS = 8, ATT_WARPS = 16;    // + warpgroup 0: 2 weight producers, MMA, activation producer
    static constexpr int THREADS = 32 * (4 + EPI_WARPS + ATT_WARPS);   // 896 = 7 warpgroups, launched at 72 registers/thread
    // setmaxnreg re-distributes the CTA's LAUNCH-TIME pool (72 registers x 896 threads = 64512), not the whole register file
    static constexpr int REG_LAUNCH = 72, REG_CTRL = 24, REG_EPI = 48, REG_ATT = 96;    // 128*24 + 256*48 + 512*96 = 64512
    static_assert(128 * REG_CTRL + 256 * REG_EPI + 512 * REG_ATT <= REG_LAUNCH * THREADS, "register pool of the CTA");
    static constexpr size_t SMEM = 1024 + (size_t)XB + 2 * QKV_BUF + BT_FLOATS * 4 + (size_t)STAGES * WST + 3 * C * 4 + 256;
};

struct FusedAttnMaps {
    CUtensorMap xw;    // x as (c, x, y, b), box (64, 6, 6, 1): one whole window
    CUtensorMap xh;    // same tensor, box (64, 3, 1, 1): half a window row (wrapped windows)
    CUtensorMap w;     // packed Wqkv [3C][C], box (64, 6d)
};
struct FusedAttnParams {
    unsigned long long* tl;   // optional debug timeline (nb200_debug_timeline)
    int B, H, W, shift;
    int nww, nwh, nwin, tiles;
    const float* bqkv;       // packed order
    const float* bias_tab;   // [6][36][40]
    __half* att;
};

// One 16-row query tile of one head of one window; q/k/v rows of the window start at sq/sk/sv (row pitch LD halfs), head
// columns at +hc.  Same arithmetic as window_attention_mma_kernel (swin_attention_mma.cu): 5 key tiles, row sums from a
// ones column in the PV product, LAST = rows 32..35 only.  Output -> so rows (staged over the q rows of this m-tile).
template <int D, int LD, bool LAST>
__device__ __forceinline__ void fa_mtile(const __half* sq, const __half* sk, const __half* sv, const float* bt, int mt, int hc,
                                         int lane, float scale, bool boundary, int reg_lo, int reg_hi) {
    constexpr int NKT = 5;
    constexpr uint32_t ONES = 0x3C003C00u;
    constexpr int HL = LAST ? 1 : 2;
    const int g = lane >> 2, t4 = lane & 3;
    const int row0 = min(mt * 16 + g, WTOK - 1), row1 = min(mt * 16 + g + 8, WTOK - 1);
    float s[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) s[nt][r] = 0.f;
#pragma unroll
    for (int kt = 0; kt < D / 16; ++kt) {
        uint32_t a[4];
        const __half* p0 = sq + row0 * LD + hc + kt * 16 + 2 * t4;
        const __half* p1 = sq + row1 * LD + hc + kt * 16 + 2 * t4;
        a[0] = *reinterpret_cast<const uint32_t*>(p0);
        a[1] = *reinterpret_cast<const uint32_t*>(p1);
        a[2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
            const __half* pk = sk + min(nt * 8 + g, WTOK - 1) * LD + hc + kt * 16 + 2 * t4;
            mma16816(s[nt], a, *reinterpret_cast<const uint32_t*>(pk), *reinterpret_cast<const uint32_t*>(pk + 8));
        }
    }
    // scale + relative position bias (log2e folded in; key columns 36..39 hold -1e30)
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
        const float2 b0 = *reinterpret_cast<const float2*>(bt + row0 * BT_LD + nt * 8 + 2 * t4);
        s[nt][0] = fmaf(s[nt][0], scale, b0.x);
        s[nt][1] = fmaf(s[nt][1], scale, b0.y);
        if (!LAST) {
            const float2 b1 = *reinterpret_cast<const float2*>(bt + row1 * BT_LD + nt * 8 + 2 * t4);
            s[nt][2] = fmaf(s[nt][2], scale, b1.x);
            s[nt][3] = fmaf(s[nt][3], scale, b1.y);
        }
    }
    if (boundary) {   // -100 across shift regions (:193-209); region of token t lives in lane t&31 (reg_lo: t<32, reg_hi: t>=32)
        const int q0 = LAST ? __shfl_sync(0xffffffffu, reg_hi, min(g, 3)) : __shfl_sync(0xffffffffu, reg_lo, row0);
        const int q1 = LAST ? 0 : __shfl_sync(0xffffffffu, reg_lo, row1 & 31);
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cr = nt < 4 ? __shfl_sync(0xffffffffu, reg_lo, nt * 8 + 2 * t4 + e)
                                      : __shfl_sync(0xffffffffu, reg_hi, min(2 * t4 + e, 3));
                if (cr != q0) s[nt][e] += -100.0f * 1.4426950408889634f;
                if (!LAST && cr != q1) s[nt][2 + e] += -100.0f * 1.4426950408889634f;
            }
    }
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
        const __half* vb = sv + (kt * 16 + (lane & 15)) * LD + hc;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            uint32_t b0, b1;
            ldmatrix_x2_trans(b0, b1, vb + nt * 8);
            mma16816(o[nt], a, b0, b1);
        }
    }
    {
        const uint32_t a0 = pack_half2(s[4][0], s[4][1]), a1 = pack_half2(s[4][2], s[4][3]);
        mma1688(osum, a0, a1, ONES);
        const __half* vb = sv + min(32 + (lane & 7), WTOK - 1) * LD + hc;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            uint32_t b0;
            ldmatrix_x1_trans(b0, vb + nt * 8);
            mma1688(o[nt], a0, a1, b0);
        }
    }
    const float inv0 = __fdividef(1.f, osum[0]);
    const float inv1 = LAST ? 0.f : __fdividef(1.f, osum[2]);
    __syncwarp();   // every lane has finished reading this head's q columns of rows [16 mt, 16 mt + 16)
    __half* so = const_cast<__half*>(sq);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        if (!LAST || r0 < WTOK) *reinterpret_cast<uint32_t*>(so + r0 * LD + hc + nt * 8 + 2 * t4) = pack_half2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (!LAST) *reinterpret_cast<uint32_t*>(so + r1 * LD + hc + nt * 8 + 2 * t4) = pack_half2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

// Register re-allocation between warpgroups (sm_90+): the CTA starts with 72 registers per thread (896 threads); the producer /
// MMA warpgroup and the two epilogue warpgroups give registers back, the four attention warpgroups take them (24 / 48 / 96).
// The pool is what the CTA was launched with (72 x 896 = 64512), so the three figures must add up to no more than that.
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <int C>
__global__ void __launch_bounds__(FaCfg<C>::THREADS, 1) swin_attn_fused_kernel(const __grid_constant__ FusedAttnMaps maps,
                                                                               const __grid_constant__ FusedAttnParams p) {
    using Cfg = FaCfg<C>;
    constexpr int D = Cfg::D, KCH = Cfg::KCH, NCHK = Cfg::NCHK, LDH = Cfg::LDH, XCH = Cfg::XCH, WST = Cfg::WST, S = Cfg::STAGES;
    constexpr int THREADS = Cfg::THREADS;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* sX = smem;                                        // [KCH][128][64] swizzled; rows 108..127 stay zero
    uint8_t* sW = sX + Cfg::XB;                                // weight ring
    uint8_t* sQKV = sW + S * WST;                              // 2 x { q | k | v } [108][LDH]
    float* sBT = reinterpret_cast<float*>(sQKV + 2 * Cfg::QKV_BUF);   // [6][36][40]
    float* sBias = sBT + BT_FLOATS;                            // [3C] packed order
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 3 * C);
    uint64_t* w_full = bars;            // [S]
    uint64_t* w_empty = w_full + S;     // [S]
    uint64_t* x_full = w_empty + S;
    uint64_t* x_empty = x_full + 1;
    uint64_t* d_full = x_empty + 1;     // [2]
    uint64_t* d_empty = d_full + 2;     // [2]  epilogue warps
    uint64_t* qkv_full = d_empty + 2;   // [2]  epilogue warps
    uint64_t* qkv_empty = qkv_full + 2; // [2]  attention warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qkv_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // debug timeline: tracks 0 weights producer, 1 activation producer, 2 MMA, 3 first epilogue warp, 4/5 attention warps 0/12
    unsigned long long* tlb = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    int tli = 0;
#define FTL(track, tag, aux) do { if (tlb && tli < 2048) { tlb[(track) * 2048 + tli] = ((unsigned long long)(tag) << 56) | ((unsigned long long)((aux) & 0xffff) << 40) | (clock64() & 0xffffffffffull); ++tli; } } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.xw);
        tma_prefetch_desc(&maps.xh);
        tma_prefetch_desc(&maps.w);
        for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], Cfg::EPI_WARPS);
            mbar_init(&qkv_full[s], Cfg::EPI_WARPS); mbar_init(&qkv_empty[s], Cfg::ATT_WARPS);
        }
        fence_barrier_init();
    }
    if (threadIdx.x == 0) NB_PDL_TRIGGER();
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    // launch constants: bias table, qkv bias; zero the whole activation tile once (its 20 pad rows are never written again)
    for (int i = threadIdx.x; i < BT_FLOATS; i += THREADS) sBT[i] = __ldg(p.bias_tab + i);
    for (int i = threadIdx.x; i < 3 * C; i += THREADS) sBias[i] = __ldg(p.bqkv + i);
    for (int i = threadIdx.x; i < Cfg::XB / 16; i += THREADS) reinterpret_cast<uint4*>(sX)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();     // the zeros are read by the tensor core through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int first = blockIdx.x, stride = gridDim.x;
    const int wpi = p.nww * p.nwh;   // windows per image

    if (warp < 4) {
        // ============ warpgroup 0: warp 0 / 2 weight producers (even / odd ring slots), warp 1 MMA, warp 3 activation producer
        reg_dec<Cfg::REG_CTRL>();
        if (warp == 3) {
            // ---- activation tile: 3 windows gathered by the TMA unit (bulk ops of one warp run one after the other, ~0.34 us each:
            //      profiles/r2/tma_inflight.json - 9 ops per tile stay well under the tile time, the weights have their own warps)
            if (elect_one() && first < p.tiles) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                uint32_t par = 0;
                for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
                    FTL(1, 1, 0);
                    mbar_wait(x_empty, par ^ 1);
                    FTL(1, 2, 0);
                    int nvalid = p.nwin - tile * WPT;
                    if (nvalid > WPT) nvalid = WPT;
                    mbar_expect_tx(x_full, (uint32_t)(nvalid * WTOK * 128 * KCH));
                    for (int wi = 0; wi < nvalid; ++wi) {
                        const int win = tile * WPT + wi;
                        const int b = win / wpi, rem = win - b * wpi;
                        const int wy = rem / p.nww, wx = rem - wy * p.nww;
                        const int y0 = wy * WS + p.shift, x0 = wx * WS + p.shift;   // torch.roll(-shift): window row r <- row (r + shift) % H
                        uint8_t* dst = sX + wi * WTOK * 128;
                        if (y0 + WS <= p.H && x0 + WS <= p.W) {
                            for (int kc = 0; kc < KCH; ++kc) tma_load_4d(&maps.xw, x_full, dst + kc * XCH, kc * 64, x0, y0, b);
                        } else {
                            for (int yy = 0; yy < WS; ++yy) {
                                int y = y0 + yy; if (y >= p.H) y -= p.H;
                                for (int hx = 0; hx < 2; ++hx) {
                                    int x = x0 + 3 * hx; if (x >= p.W) x -= p.W;
                                    for (int kc = 0; kc < KCH; ++kc)
                                        tma_load_4d(&maps.xh, x_full, dst + kc * XCH + (yy * WS + 3 * hx) * 128, kc * 64, x, y, b);
                                }
                            }
                        }
                    }
                }
            }
        } else if (warp != 1) {
            // ---- weights of the 3 head pairs, K-chunk by K-chunk; ring slot ws belongs to producer ws & 1 (fixed ownership:
            //      mbarrier waits are 1-bit phase parities, a thread must see every phase of a barrier it waits on)
            const int pid = warp >> 1;
            if (elect_one() && first < p.tiles) {
                int ws = 0;
                uint32_t wph = 0;
                for (int tile = first; tile < p.tiles; tile += stride)
                    for (int c = 0; c < Cfg::NPAIR; ++c)
                        for (int kc = 0; kc < KCH; ++kc) {
                            if ((ws & 1) == pid) {
                                if (pid == 0) FTL(0, 3, c);
                                mbar_wait(&w_empty[ws], wph ^ 1);
                                if (pid == 0) FTL(0, 4, c);
                                mbar_expect_tx(&w_full[ws], WST);
                                tma_load_2d(&maps.w, &w_full[ws], sW + ws * WST, kc * 64, c * NCHK);
                            }
                            if (++ws == S) { ws = 0; wph ^= 1; }
                        }
            }
        } else {
            // ---- MMA issuer
            const uint32_t idesc = make_idesc_f16(NCHK);
            const uint32_t aX = smem_u32(sX), aW = smem_u32(sW);
            int ws = 0;
            uint32_t wph = 0, par = 0, gc = 0;
            for (int tile = first; tile < p.tiles; tile += stride, par ^= 1) {
                if (lane == 0) FTL(2, 10, 0);
                mbar_wait(x_full, par);
                tc_fence_after();
                if (lane == 0) FTL(2, 11, 0);
                for (int c = 0; c < Cfg::NPAIR; ++c, ++gc) {
                    const uint32_t buf = gc & 1, ph = (gc >> 1) & 1;
                    mbar_wait(&d_empty[buf], ph ^ 1);
                    tc_fence_after();
                    if (lane == 0) FTL(2, 12, c);
                    const uint32_t td = tmem_base + buf * NCHK;
                    for (int kc = 0; kc < KCH; ++kc) {
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        if (lane == 0) FTL(2, 13, kc);
                        if (elect_one()) {
                            const int ksteps = kc == KCH - 1 ? Cfg::KLAST : 4;
                            for (int k = 0; k < ksteps; ++k)
                                umma_f16(td, make_kmajor_desc<128>(aX + kc * XCH + k * 32), make_kmajor_desc<128>(aW + ws * WST + k * 32),
                                         idesc, (kc > 0 || k > 0) ? 1u : 0u);
                            umma_commit(&w_empty[ws]);
                            if (kc == KCH - 1) {
                                umma_commit(&d_full[buf]);
                                if (c == Cfg::NPAIR - 1) umma_commit(x_empty);   // the activation tile may be overwritten
                            }
                        }
                        __syncwarp();
                        if (++ws == S) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp < 4 + Cfg::EPI_WARPS) {
        // ============ warpgroups 1-2: GEMM epilogue (warps 4..11): TMEM -> + bias -> fp16 -> q | k | v in shared memory
        reg_dec<Cfg::REG_EPI>();
        const int g = warp & 3, hq = (warp - 4) >> 2;   // TMEM lane group, column half
        const int r = g * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(g * 32) << 16);
        constexpr int NPW = NCHK / 16;                  // 8-column pieces per warp: 12 (C=192) or 6 (C=96)
        constexpr int GP = 2;                           // pieces per TMEM wait (48 registers per thread here)
        uint32_t gc = 0;
        for (int tile = first; tile < p.tiles; tile += stride) {
            for (int c = 0; c < Cfg::NPAIR; ++c, ++gc) {
                const uint32_t buf = gc & 1, ph = (gc >> 1) & 1;
                const bool tle = warp == 4 && lane == 0;
                if (tle) FTL(3, 20, c);
                mbar_wait(&d_full[buf], ph);
                tc_fence_after();
                if (tle) FTL(3, 21, c);
                mbar_wait(&qkv_empty[buf], ph ^ 1);
                if (tle) FTL(3, 22, c);
                uint8_t* qb = sQKV + buf * Cfg::QKV_BUF;
                const float* bia = sBias + c * NCHK;
#pragma unroll 1
                for (int grp = 0; grp < NPW / GP; ++grp) {
                    uint32_t acc[GP][8];
#pragma unroll
                    for (int i = 0; i < GP; ++i) tmem_ld8(tlane + buf * NCHK + (uint32_t)((hq * NPW + grp * GP + i) * 8), acc[i]);
                    tmem_ld_wait();
                    if (r < WPT * WTOK) {
#pragma unroll
                        for (int i = 0; i < GP; ++i) {
                            const int c0 = (hq * NPW + grp * GP + i) * 8;         // column within the pair chunk
                            const int m = c0 / (2 * D), cm = c0 - m * (2 * D);    // matrix (q,k,v), column within it
                            const float4 b0 = *reinterpret_cast<const float4*>(bia + c0), b1 = *reinterpret_cast<const float4*>(bia + c0 + 4);
                            __align__(16) __half2 o[4];
                            o[0] = __floats2half2_rn(__uint_as_float(acc[i][0]) + b0.x, __uint_as_float(acc[i][1]) + b0.y);
                            o[1] = __floats2half2_rn(__uint_as_float(acc[i][2]) + b0.z, __uint_as_float(acc[i][3]) + b0.w);
                            o[2] = __floats2half2_rn(__uint_as_float(acc[i][4]) + b1.x, __uint_as_float(acc[i][5]) + b1.y);
                            o[3] = __floats2half2_rn(__uint_as_float(acc[i][6]) + b1.z, __uint_as_float(acc[i][7]) + b1.w);
                            *reinterpret_cast<uint4*>(qb + m * Cfg::QKV_MAT + (r * LDH + cm) * 2) = *reinterpret_cast<const uint4*>(o);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&d_empty[buf]); mbar_arrive(&qkv_full[buf]); }
                if (tle) FTL(3, 23, c);
            }
        }
    } else {
        // ============ warpgroups 3-6: 16 attention warps.  Per head pair: 12 full tasks (window, head, 16-row query tile 0/1) on
        // warps 0..11, the 6 quarter-size last query tiles (rows 32..35, both heads of a window) on warps 12..14; warp 15 idles.
        reg_inc<Cfg::REG_ATT>();
        const int a = warp - (4 + Cfg::EPI_WARPS);
        const bool last_task = a >= 12;
        const int wi = last_task ? a - 12 : a >> 2;
        const int mt = last_task ? 2 : (a & 1);
        const int hh0 = last_task ? 0 : ((a >> 1) & 1), nh = last_task ? 2 : 1;   // heads of the pair handled by this warp
        const bool idle = a == 15;
        const float scale = ((D == 16) ? 0.25f : 0.17677669529663687f) * 1.4426950408889634f;   // d^-0.5 (:187) * log2(e)
        constexpr int PPH = (D * 2) / 16;               // 16-byte pieces per output row of ONE head: 4 or 2
        uint32_t gc = 0;
        for (int tile = first; tile < p.tiles; tile += stride) {
            const int win = tile * WPT + wi;
            const bool valid = !idle && win < p.nwin;
            int b = 0, wy = 0, wx = 0;
            if (valid) { b = win / wpi; const int rem = win - b * wpi; wy = rem / p.nww; wx = rem - wy * p.nww; }
            const bool boundary = p.shift > 0 && (wy == p.nwh - 1 || wx == p.nww - 1);   // only these windows mix mask regions
            // region id (:193-209) of window token t: lane t (t < 32) in reg_lo, lane t-32 in reg_hi
            int reg_lo = 0, reg_hi = 0;
            if (boundary) {
                auto region = [&](int t) {
                    const int ry = wy * WS + t / WS, rx = wx * WS + t % WS;
                    const int hr = ry < p.H - WS ? 0 : (ry < p.H - p.shift ? 1 : 2);
                    const int wr = rx < p.W - WS ? 0 : (rx < p.W - p.shift ? 1 : 2);
                    return hr * 3 + wr;
                };
                reg_lo = region(lane);
                reg_hi = region(32 + min(lane, 3));
            }
            for (int c = 0; c < Cfg::NPAIR; ++c, ++gc) {
                const uint32_t buf = gc & 1, ph = (gc >> 1) & 1;
                const bool tla = (a == 0 || a == 12) && lane == 0;
                const int ttr = a == 0 ? 4 : 5;
                if (tla) FTL(ttr, 30, c);
                mbar_wait(&qkv_full[buf], ph);
                if (tla) FTL(ttr, 31, c);
                if (valid) {
                    __half* mq = reinterpret_cast<__half*>(sQKV + buf * Cfg::QKV_BUF) + wi * WTOK * LDH;
                    const __half* mk = mq + Cfg::QKV_MAT / 2;
                    const __half* mv = mk + Cfg::QKV_MAT / 2;
#pragma unroll 1
                    for (int hi = 0; hi < nh; ++hi) {
                        const int hh = hh0 + hi;
                        const float* bt = sBT + (c * 2 + hh) * WTOK * BT_LD;
                        if (mt < 2) fa_mtile<D, LDH, false>(mq, mk, mv, bt, mt, hh * D, lane, scale, boundary, reg_lo, reg_hi);
                        else fa_mtile<D, LDH, true>(mq, mk, mv, bt, mt, hh * D, lane, scale, boundary, reg_lo, reg_hi);
                    }
                    __syncwarp();
                    if (tla) FTL(ttr, 32, c);
                    // rows [16 mt, 16 mt + 16) x this warp's head columns -> att at the un-rolled token positions
                    // (staging the rows in registers and releasing the buffer before the global stores was measured SLOWER:
                    //  626 -> 760 us on the 921 600-token launch, the extra live registers spill in all 16 warps)
                    const int nrows = mt < 2 ? 16 : WTOK - 32;
                    const int ppr = PPH * nh;                                     // pieces per row written by this warp
                    for (int idx = lane; idx < nrows * ppr; idx += 32) {
                        const int i = mt * 16 + idx / ppr, pc = hh0 * PPH + idx % ppr;
                        int y = wy * WS + i / WS + p.shift, x = wx * WS + i % WS + p.shift;
                        if (y >= p.H) y -= p.H;
                        if (x >= p.W) x -= p.W;
                        const size_t tok = ((size_t)b * p.H + y) * p.W + x;
                        *reinterpret_cast<uint4*>(p.att + tok * C + c * 2 * D + pc * 8) = *reinterpret_cast<const uint4*>(mq + i * LDH + pc * 8);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&qkv_empty[buf]);
                if (tla) FTL(ttr, 33, c);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int C>
static int launch_attn(cudaStream_t st, const FusedAttn& f) {
    using Cfg = FaCfg<C>;
    FusedAttnMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)f.W * C * 2, (cuuint64_t)f.H * f.W * C * 2};
        cuuint32_t boxw[4] = {64, WS, WS, 1}, boxh[4] = {64, 3, 1, 1};
        if (encode(&maps.xw, f.x, 4, dims, strides, boxw, 128)) return 1;
        if (encode(&maps.xh, f.x, 4, dims, strides, boxh, 128)) return 1;
        cuuint64_t wd[2] = {(cuuint64_t)C, (cuuint64_t)3 * C};
        cuuint64_t wst[1] = {(cuuint64_t)C * 2};
        cuuint32_t wb[2] = {64, (cuuint32_t)Cfg::NCHK};
        if (encode(&maps.w, f.wqkv, 2, wd, wst, wb, 128)) return 1;
    }
    FusedAttnParams p;
    p.B = f.B; p.H = f.H; p.W = f.W;
    p.shift = (WS >= f.H || WS >= f.W) ? 0 : f.shift;   // torchvision :151-155: no shift when the window covers the map
    p.nww = f.W / WS; p.nwh = f.H / WS;
    p.nwin = f.B * p.nww * p.nwh;
    p.tiles = (p.nwin + WPT - 1) / WPT;
    p.bqkv = f.bqkv; p.bias_tab = f.bias_tab; p.att = f.att;
    p.tl = g_timeline;
    if (ensure_dyn_smem((const void*)swin_attn_fused_kernel<C>, Cfg::SMEM)) return 1;
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
    NB_CUDA(cudaLaunchKernelEx(&cfg, swin_attn_fused_kernel<C>, maps, p));
    NB_LAUNCHED();
    return 0;
}

// reference row order (q | k | v, head-major inside each) -> packed (pair, {q,k,v}, head-in-pair, d)
__global__ void pack_qkv_kernel(const __half* __restrict__ w, const float* __restrict__ b, __half* __restrict__ wp, float* __restrict__ bp, int C) {
    const int D = C / HEADS;
    const int pr = blockIdx.x;                       // packed row
    const int c = pr / (6 * D), rem = pr % (6 * D);
    const int m = rem / (2 * D), hh = (rem % (2 * D)) / D, d = rem % D;
    const int src = m * C + (2 * c + hh) * D + d;
    for (int k = threadIdx.x; k < C; k += blockDim.x) wp[(size_t)pr * C + k] = w[(size_t)src * C + k];
    if (threadIdx.x == 0) bp[pr] = b[src];
}
// [6][36][40]: log2(e) * relative_position_bias_table[relative_position_index] (swin_transformer.py:267-279), pads = -1e30
__global__ void build_bias_tab_kernel(const float* __restrict__ table, float* __restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BT_FLOATS) return;
    const int col = i % BT_LD, row = (i / BT_LD) % WTOK, head = i / (BT_LD * WTOK);
    float v = -1e30f;
    if (col < WTOK) {
        const int qy = row / WS, qx = row % WS, ky = col / WS, kx = col % WS;
        v = 1.4426950408889634f * table[((qy - ky + WS - 1) * (2 * WS - 1) + (qx - kx + WS - 1)) * HEADS + head];
    }
    tab[i] = v;
}

}  // namespace

int swin_attn_fused(cudaStream_t st, const FusedAttn& f) {
    NB_CHECK(f.x && f.att && f.wqkv && f.bqkv && f.bias_tab, "null pointer");
    NB_CHECK(f.B > 0 && f.H > 0 && f.W > 0, "empty input");
    NB_CHECK(f.H % WS == 0 && f.W % WS == 0, "feature map must be a multiple of the 6x6 window");
    NB_CHECK(f.C == 96 || f.C == 192, "fused window attention supports C = 96 (d = 16) and C = 192 (d = 32)");
    NB_CHECK(f.shift == 0 || f.shift == 3, "shift must be 0 or window/2");
    return f.C == 192 ? launch_attn<192>(st, f) : launch_attn<96>(st, f);
}

int fused_attn_bias_tab_floats() { return BT_FLOATS; }

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_swin_attn_fused_f16(const void* x, const void* wqkv, const float* bqkv, const float* bias_table, void* att,
                                         int B, int H, int W, int C, int shift, void* stream) {
    NB_CHECK(x && wqkv && bqkv && bias_table && att, "null pointer");
    NB_CHECK(C == 96 || C == 192, "C must be 96 or 192");
    cudaStream_t st = (cudaStream_t)stream;
    __half* wp = nullptr;
    float *bp = nullptr, *tab = nullptr;
    NB_CUDA(cudaMallocAsync((void**)&wp, (size_t)3 * C * C * 2, st));
    NB_CUDA(cudaMallocAsync((void**)&bp, (size_t)3 * C * 4, st));
    NB_CUDA(cudaMallocAsync((void**)&tab, (size_t)BT_FLOATS * 4, st));
    pack_qkv_kernel<<<3 * C, 96, 0, st>>>((const __half*)wqkv, bqkv, wp, bp, C);
    build_bias_tab_kernel<<<(BT_FLOATS + 255) / 256, 256, 0, st>>>(bias_table, tab);
    FusedAttn f;
    f.x = (const __half*)x; f.att = (__half*)att; f.B = B; f.H = H; f.W = W; f.C = C; f.shift = shift;
    f.wqkv = wp; f.bqkv = bp; f.bias_tab = tab;
    const int rc = swin_attn_fused(st, f);
    cudaFreeAsync(wp, st); cudaFreeAsync(bp, st); cudaFreeAsync(tab, st);
    return rc;
}
