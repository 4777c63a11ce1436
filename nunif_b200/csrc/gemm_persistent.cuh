// Persistent tcgen05 implicit GEMM (sm_100a): the production kernel of path A.
//
// Same math and operand formats as gemm_tcgen05.cuh (which documents the PTX wrappers, descriptors and the
// tap/view scheme); this kernel changes the *schedule*:
//   * one CTA per SM, each CTA keeps ONE n-tile and walks the m-tiles with a static stride, so
//       - the weight tile (all K chunks of its BLOCK_N rows) is TMA-loaded ONCE and stays resident in shared
//         memory when it fits (RESIDENT_B; true for every Swin Linear and most convs) - per tile only the
//         activation chunk(s) move,
//       - the CTAs that share an activation tile (n_tiles of them) run in lock-step => the re-read hits L2;
//   * an A ring (TMA producer warp -> MMA warp) that runs ahead across tile boundaries;
//   * two TMEM accumulators (2 x BLOCK_N columns): the MMA of tile j+1 overlaps the epilogue of tile j;
//   * the epilogue is split into independent QUADS (4 warps = the 4 TMEM lane groups).  The CTA's output is a
//     sequence of [128 x CW] chunks; chunk q goes to quad q mod NQ, which owns one 128B-swizzled staging buffer,
//     converts the chunk (tcgen05.ld -> bias/activation/residual -> fp16), and issues its own TMA tensor store.
//     Quads only synchronise internally (128-thread named barriers), so up to NQ chunks are in flight and the
//     per-chunk latency chain (TMEM load, dependent math, fence, store issue) is overlapped NQ-fold;
//   * residual chunks are TMA-prefetched by the producer into per-quad buffers, NQ chunks ahead;
//   * no runtime integer division anywhere in the steady state: ring slots, phases and tile coordinates
//     advance incrementally (a clock64 timeline of the first version showed ~500-cycle dependent chains of
//     IDIV/IMOD sequences in every role - profiles/r1/timeline_*.txt).
// Every global access is a bulk tensor copy.  The Swin Linears (M ~ 1e6, K,N <= 576) are HBM-bound; the
// traffic-mix floor is max((R+W)/6.6, W/3.9, R/6.2 TB/s) (profiles/r1/hbm_microbench.json).
#pragma once
#include "gemm_tcgen05.cuh"

namespace nb200 {

constexpr int PG_MAX_STAGES = 8;   // A(/B) ring depth upper bound
constexpr int PG_MAX_QUADS = 4;    // epilogue quads (each 4 warps)
constexpr int PG_EPI_WARPS = 4 * PG_MAX_QUADS;
// TMA producers: warp 0 and the warps after the epilogue quads.  Bulk-tensor ops issued by ONE warp execute strictly one after
// the other on this part (~0.34 us per op whatever its size), ops of different warps overlap (profiles/r2/tma_inflight.json):
// the op sequence of a CTA is dealt round-robin to PG_PRODUCERS warps.  (3 producers = 640 threads keep 96 registers/thread.)
constexpr int PG_PRODUCERS = 3;
constexpr int PG_THREADS = 64 + 32 * PG_EPI_WARPS + 32 * (PG_PRODUCERS - 1);

struct PersistParams {
    GemmParams g;
    int m_tiles;        // tiles_x * tiles_y * B
    int grid_m;         // CTAs per n-tile (gridDim.x / n_tiles)
    int stages;         // ring depth chosen by the host for the shared-memory budget
    int k_iters;
    int nq;             // active epilogue quads == staging buffers (3 with a residual ring, else 4)
    unsigned long long* timeline;  // optional debug: CTA 0 records (event << 56 | clock) per role (profiles/gemm_timeline.py)
};

// (tx, ty, b) of a tile index, advanced by a fixed stride without divisions
struct TileWalk {
    int tx, ty, b, sx, sy, sb, nx, ny;
    __device__ __forceinline__ void init(int tile, int stride, int tiles_x, int tiles_y) {
        nx = tiles_x; ny = tiles_y;
        tx = tile % tiles_x; ty = (tile / tiles_x) % tiles_y; b = tile / (tiles_x * tiles_y);
        sx = stride % tiles_x; sy = (stride / tiles_x) % tiles_y; sb = stride / (tiles_x * tiles_y);
    }
    __device__ __forceinline__ void step() {
        tx += sx;
        int cy = 0;
        if (tx >= nx) { tx -= nx; cy = 1; }
        ty += sy + cy;
        int cb = 0;
        if (ty >= ny) { ty -= ny; cb = 1; }
        b += sb + cb;
    }
};

template <int BLOCK_N>
struct PgAcc {
    static constexpr int NACC = (4 * BLOCK_N <= 512) ? 4 : 2;
    static constexpr int SHIFT = NACC == 4 ? 2 : 1;
};

template <int BLOCK_N, int BK, bool RESIDENT_B>
__global__ void __launch_bounds__(PG_THREADS, 1) gemm_conv_persistent(const __grid_constant__ GemmMaps maps,
                                                                      const __grid_constant__ PersistParams pp) {
    using Cfg = GemmCfg<BLOCK_N, BK>;
    constexpr int CW = Cfg::CW, NCH = Cfg::NCH, CH_BYTES = Cfg::CH_BYTES, NSUB = CW / 16;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = ((Cfg::B_BYTES + 1023) / 1024) * 1024;
    constexpr int STAGE_BYTES = RESIDENT_B ? A_BYTES : A_BYTES + B_BYTES;
    // accumulator ring in TMEM: 4 tiles deep when they fit in the 512 columns (BLOCK_N <= 128), else 2
    constexpr int NACC = PgAcc<BLOCK_N>::NACC, ACC_SH = PgAcc<BLOCK_N>::SHIFT;
    constexpr int TMEM_COLS = NACC * BLOCK_N <= 32 ? 32 : (NACC * BLOCK_N <= 64 ? 64 : (NACC * BLOCK_N <= 128 ? 128 : (NACC * BLOCK_N <= 256 ? 256 : 512)));
    const GemmParams& p = pp.g;
    const int SA = pp.stages, k_iters = pp.k_iters, NQ = pp.nq;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                                               // resident weights: k_iters chunks of B_BYTES
    uint8_t* sRing = sB + (RESIDENT_B ? k_iters * B_BYTES : 0);       // SA stages
    uint8_t* sOut = sRing + SA * STAGE_BYTES;                         // NQ output staging chunks (one per quad)
    uint8_t* sRes = sOut + NQ * CH_BYTES;                             // NQ residual chunks (if has_res)
    float* sBias = reinterpret_cast<float*>(sRes + (p.has_res ? NQ * CH_BYTES : 0));  // this CTA's BLOCK_N biases
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + BLOCK_N);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + PG_MAX_STAGES;
    uint64_t* acc_full = a_empty + PG_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 4;
    uint64_t* res_full = acc_empty + 4;
    uint64_t* res_empty = res_full + PG_MAX_QUADS;
    uint64_t* b_full = res_empty + PG_MAX_QUADS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // timeline slots: producer [0,1024), MMA [1024,2048), quad-0 leader [2048,3072), quad-1 leader [3072,4096)
    unsigned long long* tl = (pp.timeline && blockIdx.x == 0) ? pp.timeline : nullptr;
    int tli = 0;
#define TL(base, ev) do { if (tl && tli < 1024) { tl[(base) + tli] = ((unsigned long long)(ev) << 56) | (clock64() & 0x00ffffffffffffffull); ++tli; } } while (0)
    const int n_tile = blockIdx.x % p.n_tiles;
    const int m_first = blockIdx.x / p.n_tiles;
    const int n0 = n_tile * BLOCK_N;
    const int my_tiles = m_first < pp.m_tiles ? (pp.m_tiles - m_first + pp.grid_m - 1) / pp.grid_m : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.o[0]);
        for (int s = 0; s < PG_MAX_STAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < NACC; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], NCH);   // one arrival per chunk of the tile (from the quad that converted it)
        }
        for (int s = 0; s < PG_MAX_QUADS; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 1);
        }
        mbar_init(b_full, (uint32_t)k_iters);          // one arrival per resident weight chunk op
        fence_barrier_init();
    }
    if (threadIdx.x == 0) NB_PDL_TRIGGER();   // the next GEMM of the stream may start its prologue while this grid drains
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + BLOCK_N) sBias[threadIdx.x - 64] = p.bias ? __ldg(p.bias + n0 + (threadIdx.x - 64)) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp >= 2 + PG_EPI_WARPS) {
        // ===================== TMA producers (one thread in each of PG_PRODUCERS warps) =====================
        // Fixed ownership: mbarrier waits are 1-bit phase parities, so a thread must see EVERY phase of a barrier it waits on -
        // activation ring slot s is always refilled by producer s % PG_PRODUCERS, residual slot rq by producer rq % PG_PRODUCERS.
        // All producers walk the same op sequence, so slots and phases need no communication; every op does its own
        // arrive.expect_tx on its barrier.
        const int pid = warp == 0 ? 0 : warp - (2 + PG_EPI_WARPS) + 1;
        if (elect_one() && my_tiles > 0) {
            if (RESIDENT_B) {
                for (int it = pid; it < k_iters; it += PG_PRODUCERS) {
                    mbar_expect_tx(b_full, (uint32_t)Cfg::B_BYTES);
                    tma_load_2d(&maps.b, b_full, sB + it * B_BYTES, it * BK, n0);
                }
            }
            // everything above touched only launch constants (weights, bias, tensor maps); activations and residuals are
            // produced by the preceding kernels of the stream
            asm volatile("griddepcontrol.wait;" ::: "memory");
            TileWalk tw;
            tw.init(m_first, pp.grid_m, p.tiles_x, p.tiles_y);
            int s = 0, sown = 0, rq = 0, rown = 0;
            uint32_t ph = 0, rph = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int x0 = tw.tx * p.TW, y0 = tw.ty * p.TH, b = tw.b;
                int tap = 0, ch = 0;
                for (int it = 0; it < k_iters; ++it) {
                    if (sown == pid) {
                        uint8_t* sa = sRing + s * STAGE_BYTES;
                        mbar_wait(&a_empty[s], ph ^ 1);
                        TL(0, 1);
                        mbar_expect_tx(&a_full[s], RESIDENT_B ? A_BYTES : A_BYTES + Cfg::B_BYTES);
                        tma_load_5d(&maps.a, &a_full[s], sa, ch * BK, x0 + p.tap_dx[tap], p.tap_dyi[tap], y0 + p.tap_dy[tap], b);
                        if (!RESIDENT_B) tma_load_2d(&maps.b, &a_full[s], sa + A_BYTES, it * BK, n0);
                    }
                    if (++ch == p.cpt) { ch = 0; ++tap; }
                    if (++sown == PG_PRODUCERS) sown = 0;
                    if (++s == SA) { s = 0; sown = 0; ph ^= 1; }
                }
                if (p.has_res) {
                    // residual chunks of this tile, prefetched into the consuming quad's buffer NQ chunks ahead
                    int n = n0;
                    for (int c = 0; c < NCH; ++c, n += CW) {
                        if (rown == pid) {
                            mbar_wait(&res_empty[rq], rph ^ 1);
                            int g = 0, co = n;
                            if (p.out_mode != OUT_NHWC) { g = n / p.cout; co = n - g * p.cout; }
                            mbar_expect_tx(&res_full[rq], CH_BYTES);
                            tma_load_4d(&maps.r[g], &res_full[rq], sRes + rq * CH_BYTES, co, x0 + p.res_cx, y0 + p.res_cy, b);
                        }
                        if (++rown == PG_PRODUCERS) rown = 0;
                        if (++rq == NQ) { rq = 0; rown = 0; rph ^= 1; }
                    }
                }
                tw.step();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BLOCK_N);
        if (my_tiles > 0 && RESIDENT_B) {
            mbar_wait(b_full, 0);
            tc_fence_after();
        }
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int a = j & (NACC - 1);
            mbar_wait(&acc_empty[a], ((j >> ACC_SH) & 1) ^ 1);  // every chunk of the tile that used this accumulator is converted
            tc_fence_after();
            if (lane == 0) TL(1024, 2);
            const uint32_t tacc = tmem_base + (uint32_t)(a * BLOCK_N);
            for (int it = 0; it < k_iters; ++it) {
                mbar_wait(&a_full[s], ph);
                tc_fence_after();
                if (lane == 0) TL(1024, 3);
                if (elect_one()) {
                    const uint32_t sa = smem_u32(sRing + s * STAGE_BYTES);
                    const uint32_t sb = RESIDENT_B ? smem_u32(sB + it * B_BYTES) : sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = make_kmajor_desc<Cfg::SWIZZLE>(sa + k * 32);
                        const uint64_t bd = make_kmajor_desc<Cfg::SWIZZLE>(sb + k * 32);
                        umma_f16(tacc, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&a_empty[s]);
                    if (it == k_iters - 1) umma_commit(&acc_full[a]);
                }
                __syncwarp();
                if (++s == SA) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp < 2 + PG_EPI_WARPS) {
        // ===================== epilogue quads (warps 2..17) =====================
        const int quad = (warp - 2) >> 2;     // 0..3
        const int lane_grp = warp & 3;        // TMEM lanes [32*lane_grp, +32)
        const int r = lane_grp * 32 + lane;   // row of the tile == staging row
        const bool qleader = (((warp - 2) & 3) == 0) && lane == 0;  // first warp of the quad
        const int act = p.act;
        const bool has_res = p.has_res != 0, res_first = p.res_before_act != 0;
        if (quad < NQ && my_tiles > 0) {
            uint8_t* bufp = sOut + quad * CH_BYTES;
            const uint8_t* resp = sRes + quad * CH_BYTES;
            const int tlb = quad == 0 ? 2048 : 3072;
            const bool tlq = qleader && quad < 2;
            // this quad converts chunks quad, quad+NQ, ... of the CTA's chunk sequence (tile-major, NCH chunks per tile)
            int j = quad / NCH, c = quad - j * NCH;
            TileWalk tw;
            tw.init(m_first, pp.grid_m, p.tiles_x, p.tiles_y);
            for (int jj = 0; jj < j; ++jj) tw.step();
            uint32_t rph = 0;
            bool first = true;
            while (j < my_tiles) {
                const int a = j & (NACC - 1);
                if (qleader && !first) tma_store_wait_read();   // my previous store has finished reading this quad's buffer
                first = false;
                mbar_wait(&acc_full[a], (j >> ACC_SH) & 1);
                if (has_res) mbar_wait(&res_full[quad], rph);
                tc_fence_after();
                asm volatile("bar.sync %0, 128;" ::"r"(quad + 1) : "memory");   // buffer free (leader waited) for all 4 warps
                if (tlq) TL(tlb, 5);
                const uint32_t tcol = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * BLOCK_N + c * CW);
                constexpr int PAIR = NSUB >= 2 ? 2 : 1;   // two 16-column TMEM loads in flight per wait
#pragma unroll
              for (int sp = 0; sp < NSUB; sp += PAIR) {
                uint32_t acc[PAIR][16];
#pragma unroll
                for (int u = 0; u < PAIR; ++u) tmem_ld16(tcol + (sp + u) * 16, acc[u]);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < PAIR; ++u) {
                    const int sub = sp + u;
                    const uint32_t o0 = stage_off<CW>(r, 2 * sub), o1 = stage_off<CW>(r, 2 * sub + 1);
                    float v[16];
                    const float4* bp = reinterpret_cast<const float4*>(sBias + c * CW + sub * 16);  // smem broadcast
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bq = bp[q];
                        v[4 * q] = __uint_as_float(acc[u][4 * q]) + bq.x;
                        v[4 * q + 1] = __uint_as_float(acc[u][4 * q + 1]) + bq.y;
                        v[4 * q + 2] = __uint_as_float(acc[u][4 * q + 2]) + bq.z;
                        v[4 * q + 3] = __uint_as_float(acc[u][4 * q + 3]) + bq.w;
                    }
                    if (has_res) {
                        float rv[16];
                        const uint4 r0 = *reinterpret_cast<const uint4*>(resp + o0), r1 = *reinterpret_cast<const uint4*>(resp + o1);
                        const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
                        const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 e0 = __half22float2(h0[q]), e1 = __half22float2(h1[q]);
                            rv[2 * q] = e0.x; rv[2 * q + 1] = e0.y; rv[8 + 2 * q] = e1.x; rv[8 + 2 * q + 1] = e1.y;
                        }
                        if (res_first) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) v[q] += rv[q];
                            apply_act16(v, act);
                        } else {
                            apply_act16(v, act);
#pragma unroll
                            for (int q = 0; q < 16; ++q) v[q] += rv[q];
                        }
                    } else {
                        apply_act16(v, act);
                    }
                    __align__(16) __half2 o[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                    *reinterpret_cast<uint4*>(bufp + o0) = reinterpret_cast<const uint4*>(o)[0];
                    *reinterpret_cast<uint4*>(bufp + o1) = reinterpret_cast<const uint4*>(o)[1];
                }
              }
                if (tlq) TL(tlb, 6);
                tc_fence_before();       // TMEM reads of this chunk are complete
                fence_async_smem();      // staging writes visible to the TMA (async proxy)
                asm volatile("bar.sync %0, 128;" ::"r"(quad + 1) : "memory");
                if (qleader) {
                    mbar_arrive(&acc_empty[a]);                    // 1 of NCH arrivals that hand the accumulator back
                    if (has_res) mbar_arrive(&res_empty[quad]);    // residual chunk consumed
                    const int n = n0 + c * CW;
                    int g = 0, co = n;
                    if (p.out_mode != OUT_NHWC) { g = n / p.cout; co = n - g * p.cout; }
                    // out-of-range rows/cols of edge tiles are clipped by the TMA unit
                    tma_store_4d(&maps.o[g], bufp, co, tw.tx * p.TW, tw.ty * p.TH, tw.b);
                    tma_store_commit();
                    if (tlq) TL(tlb, 9);
                }
                rph ^= 1;
                c += NQ;
                while (c >= NCH) { c -= NCH; ++j; tw.step(); }
            }
            if (qleader) tma_store_wait_read();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
#undef TL
}

}  // namespace nb200
