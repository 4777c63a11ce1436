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
//   * 16 epilogue warps; results go through a ring of 128B-swizzled staging chunks and leave with TMA
//     tensor stores (the leader only ever waits for the store issued PG_NOUT chunks earlier); residual tiles
//     are TMA-prefetched by the producer into their own ring, PG_NRES chunks ahead of the epilogue.
// Every global access is a bulk tensor copy; the kernel is designed to sit on the HBM roofline for the
// memory-bound Linears (M x {96,192} activations) and on the tensor roofline for the large-K convs.
#pragma once
#include "gemm_tcgen05.cuh"

namespace nb200 {

constexpr int PG_MAX_STAGES = 8;   // A(/B) ring depth upper bound
constexpr int PG_NOUT = 4;         // output staging ring capacity (chunks of [128][CW] fp16 waiting for their TMA store);
                                   // the host picks 4 (buffer hand-back off the critical path) or 3 (when a residual ring
                                   // also has to fit)
constexpr int PG_NRES = 3;         // residual prefetch ring (same chunk shape), only allocated when a residual exists
constexpr int PG_EPI_WARPS = 16;   // four warps per TMEM lane group: enough warps in flight to hide the tcgen05.ld / MUFU latency
constexpr int PG_THREADS = 64 + 32 * PG_EPI_WARPS;

struct PersistParams {
    GemmParams g;
    int m_tiles;        // tiles_x * tiles_y * B
    int grid_m;         // CTAs per n-tile (gridDim.x / n_tiles)
    int stages;         // ring depth chosen by the host for the shared-memory budget
    int k_iters;
    int nout;           // output staging buffers in use (3 or 4)
};

template <int BLOCK_N, int BK, bool RESIDENT_B>
__global__ void __launch_bounds__(PG_THREADS, 1) gemm_conv_persistent(const __grid_constant__ GemmMaps maps,
                                                                      const __grid_constant__ PersistParams pp) {
    using Cfg = GemmCfg<BLOCK_N, BK>;
    constexpr int CW = Cfg::CW, NCH = Cfg::NCH, CH_BYTES = Cfg::CH_BYTES;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = ((Cfg::B_BYTES + 1023) / 1024) * 1024;
    constexpr int STAGE_BYTES = RESIDENT_B ? A_BYTES : A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BLOCK_N <= 32 ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    const GemmParams& p = pp.g;
    const int SA = pp.stages, k_iters = pp.k_iters, NOUT = pp.nout;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                                               // resident weights: k_iters chunks of B_BYTES
    uint8_t* sRing = sB + (RESIDENT_B ? k_iters * B_BYTES : 0);       // SA stages
    uint8_t* sOut = sRing + SA * STAGE_BYTES;                         // PG_NOUT output staging chunks
    uint8_t* sRes = sOut + NOUT * CH_BYTES;                           // PG_NRES residual chunks (if has_res)
    float* sBias = reinterpret_cast<float*>(sRes + (p.has_res ? PG_NRES * CH_BYTES : 0));  // this CTA's BLOCK_N biases
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + BLOCK_N);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + PG_MAX_STAGES;
    uint64_t* acc_full = a_empty + PG_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* res_full = acc_empty + 2;
    uint64_t* res_empty = res_full + PG_NRES;
    uint64_t* out_empty = res_empty + PG_NRES;
    uint64_t* b_full = out_empty + PG_NOUT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x % p.n_tiles;
    const int m_first = blockIdx.x / p.n_tiles;
    const int n0 = n_tile * BLOCK_N;
    const int my_tiles = m_first < pp.m_tiles ? (pp.m_tiles - m_first + pp.grid_m - 1) / pp.grid_m : 0;
    const int tiles_xy = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.o[0]);
        for (int s = 0; s < PG_MAX_STAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 1);
        }
        for (int s = 0; s < PG_NRES; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 1);
        }
        for (int s = 0; s < PG_NOUT; ++s) mbar_init(&out_empty[s], 1);
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + BLOCK_N) sBias[threadIdx.x - 64] = p.bias ? __ldg(p.bias + n0 + (threadIdx.x - 64)) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (elect_one() && my_tiles > 0) {
            if (RESIDENT_B) {
                mbar_expect_tx(b_full, (uint32_t)(k_iters * Cfg::B_BYTES));
                for (int it = 0; it < k_iters; ++it) tma_load_2d(&maps.b, b_full, sB + it * B_BYTES, it * BK, n0);
            }
            uint32_t na = 0, cq = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int tile = m_first + j * pp.grid_m;
                const int tx_i = tile % p.tiles_x, ty_i = (tile / p.tiles_x) % p.tiles_y, b = tile / tiles_xy;
                const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
                for (int it = 0; it < k_iters; ++it, ++na) {
                    const int s = na % SA;
                    mbar_wait(&a_empty[s], ((na / SA) & 1) ^ 1);
                    const int tap = it / p.cpt, ch = it - tap * p.cpt;
                    uint8_t* sa = sRing + s * STAGE_BYTES;
                    mbar_expect_tx(&a_full[s], RESIDENT_B ? A_BYTES : A_BYTES + Cfg::B_BYTES);
                    tma_load_5d(&maps.a, &a_full[s], sa, ch * BK, x0 + p.tap_dx[tap], p.tap_dyi[tap], y0 + p.tap_dy[tap], b);
                    if (!RESIDENT_B) tma_load_2d(&maps.b, &a_full[s], sa + A_BYTES, it * BK, n0);
                }
                if (p.has_res) {
                    // residual chunks of this tile, prefetched PG_NRES chunks ahead of the epilogue
                    for (int c = 0; c < NCH; ++c, ++cq) {
                        const int buf = cq % PG_NRES;
                        mbar_wait(&res_empty[buf], ((cq / PG_NRES) & 1) ^ 1);
                        const int n = n0 + c * CW;
                        const int g = p.out_mode != OUT_NHWC ? n / p.cout : 0;
                        const int co = p.out_mode != OUT_NHWC ? n - g * p.cout : n;
                        mbar_expect_tx(&res_full[buf], CH_BYTES);
                        tma_load_4d(&maps.r[g], &res_full[buf], sRes + buf * CH_BYTES, co, x0 + p.res_cx, y0 + p.res_cy, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BLOCK_N);
        if (my_tiles > 0 && RESIDENT_B) {
            mbar_wait(b_full, 0);
            tc_fence_after();
        }
        uint32_t nm = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int a = j & 1;
            mbar_wait(&acc_empty[a], ((j >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(a * BLOCK_N);
            for (int it = 0; it < k_iters; ++it, ++nm) {
                const int s = nm % SA;
                mbar_wait(&a_full[s], (nm / SA) & 1);
                tc_fence_after();
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
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        const int lane_grp = warp & 3;
        const int part = (warp - 2) >> 2;   // which 16-column blocks of a chunk this warp converts
        const int r = lane_grp * 32 + lane;
        const bool leader = (warp == 2 && lane == 0);
        const int act = p.act;
        const bool has_res = p.has_res != 0, res_first = p.res_before_act != 0;
        uint32_t cq = 0, released = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = m_first + j * pp.grid_m;
            const int tx_i = tile % p.tiles_x, ty_i = (tile / p.tiles_x) % p.tiles_y, b = tile / tiles_xy;
            const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
            const int a = j & 1;
            mbar_wait(&acc_full[a], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++cq) {
                const int buf = cq % NOUT, rbuf = cq % PG_NRES;
                mbar_wait(&out_empty[buf], ((cq / NOUT) & 1) ^ 1);        // the store issued NOUT chunks ago has read this buffer
                if (has_res) mbar_wait(&res_full[rbuf], (cq / PG_NRES) & 1);  // residual chunk landed
                uint8_t* bufp = sOut + buf * CH_BYTES;
                const uint8_t* resp = sRes + rbuf * CH_BYTES;
#pragma unroll 1
                for (int sub = part; sub < CW / 16; sub += PG_EPI_WARPS / 4) {
                    uint32_t acc[16];
                    tmem_ld16(trow + c * CW + sub * 16, acc);
                    tmem_ld_wait();
                    const uint32_t o0 = stage_off<CW>(r, 2 * sub), o1 = stage_off<CW>(r, 2 * sub + 1);
                    uint4* s0 = reinterpret_cast<uint4*>(bufp + o0);
                    uint4* s1 = reinterpret_cast<uint4*>(bufp + o1);
                    float v[16];
                    {
                        const float4* bp = reinterpret_cast<const float4*>(sBias + c * CW + sub * 16);  // smem broadcast
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bq = bp[q];
                            v[4 * q] = __uint_as_float(acc[4 * q]) + bq.x;
                            v[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) + bq.y;
                            v[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) + bq.z;
                            v[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) + bq.w;
                        }
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
                    *s0 = reinterpret_cast<const uint4*>(o)[0];
                    *s1 = reinterpret_cast<const uint4*>(o)[1];
                }
                if (c == NCH - 1) tc_fence_before();  // this tile's TMEM reads are done before the barrier below
                fence_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * PG_EPI_WARPS) : "memory");
                if (leader) {
                    if (c == NCH - 1) mbar_arrive(&acc_empty[a]);  // hand the accumulator back to the MMA warp
                    if (has_res) mbar_arrive(&res_empty[rbuf]);    // residual chunk consumed by every epilogue thread
                    const int n = n0 + c * CW;
                    const int g = p.out_mode != OUT_NHWC ? n / p.cout : 0;
                    const int co = p.out_mode != OUT_NHWC ? n - g * p.cout : n;
                    tma_store_4d(&maps.o[g], bufp, co, x0, y0, b);
                    tma_store_commit();
                    // all but the 2 newest stores have read their staging buffers: hand those buffers back.  With 4
                    // buffers the next chunk's buffer was already handed back one chunk earlier (off the critical path).
                    asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                    while (released + 2 <= cq) {
                        mbar_arrive(&out_empty[released % NOUT]);
                        ++released;
                    }
                }
            }
        }
        if (leader) tma_store_wait_read();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace nb200
