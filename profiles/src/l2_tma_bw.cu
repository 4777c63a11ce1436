// How fast can an SM ingest L2-RESIDENT data through TMA, alone and with all 148 SMs active, and does cluster multicast
// raise the per-SM ingest rate?  (Round 2: the fused Swin kernels stream their weights from L2 per 128-token tile.)
//   W: [R][64] fp16 (128-byte rows, R*128 B = 360 KB, L2 resident); one load = a box of `rows` rows (rows*128 bytes)
//   unicast  : every CTA loads whole boxes into a ring of S stages
//   multicast: clusters of CS CTAs; each CTA loads rows/CS rows of the box and multicasts them to all CS CTAs, so every CTA
//              still receives the whole box (same bytes into each SM's shared memory, 1/CS of the L2 reads)
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/src/l2_tma_bw.cu -o profiles/_bin/l2_tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    return (EncodeTiledFn)fn;
}
static CUtensorMap make_map(void* base, int R, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, (cuuint64_t)R};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 26) && !ok; ++it)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
    if (!ok) { printf("timeout block %d\n", blockIdx.x); __trap(); }
}

// CS = cluster size (1 = unicast)
template <int CS>
__global__ void __launch_bounds__(64) ingest_kernel(const __grid_constant__ CUtensorMap map, int R, int rows, int S, int loads) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16], empty[16];
    const int stage_bytes = rows * 128;
    uint32_t rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(CS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CS > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        // producer and consumer in one thread: keep up to S loads in flight
        int issued = 0, waited = 0;
        int row = (blockIdx.x / CS * 37 * rows) % R;   // clusters start at different offsets
        while (waited < loads) {
            while (issued < loads && issued - waited < S) {
                const int slot = issued % S;
                const uint32_t ph = (uint32_t)((issued / S) & 1);
                mbar_wait(smem_u32(&empty[slot]), ph ^ 1);
                const uint32_t fb = smem_u32(&full[slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(stage_bytes) : "memory");
                if (CS == 1) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(smem + slot * stage_bytes)), "l"(&map), "r"(fb), "r"(0), "r"(row) : "memory");
                } else {
                    const int part = rows / CS;
                    const uint16_t mask = (uint16_t)((1u << CS) - 1);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                 ::"r"(smem_u32(smem + slot * stage_bytes + rank * part * 128)), "l"(&map), "r"(fb), "r"(0),
                                   "r"(row + (int)rank * part), "h"(mask) : "memory");
                }
                row += rows;
                if (row + rows > R) row = 0;
                ++issued;
            }
            const int slot = waited % S;
            mbar_wait(smem_u32(&full[slot]), (uint32_t)((waited / S) & 1));
            // consumed: release the slot in every CTA of the cluster
            if (CS == 1) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
            } else {
                for (int c = 0; c < CS; ++c) {
                    uint32_t ra;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&empty[slot])), "r"(c));
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
                }
            }
            ++waited;
        }
    }
    __syncthreads();
    if (CS > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

template <int CS>
static double run(const CUtensorMap& m, int R, int rows, int S, int loads, int grid, int iters = 5) {
    const size_t smem = (size_t)S * rows * 128 + 1024;
    cudaFuncSetAttribute(ingest_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(ingest_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = CS > 1 ? 1 : 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) cudaLaunchKernelEx(&cfg, ingest_kernel<CS>, m, R, rows, S, loads);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) cudaLaunchKernelEx(&cfg, ingest_kernel<CS>, m, R, rows, S, loads);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("\n# CUDA error (CS=%d rows=%d S=%d grid=%d): %s\n", CS, rows, S, grid, cudaGetErrorString(e)); return -1; }
    return ms / iters;
}

int main() {
    const int R = 2880;   // 360 KB
    void* w;
    cudaMalloc(&w, (size_t)R * 128);
    cudaMemset(w, 0, (size_t)R * 128);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int loads = 4000;
    printf("{\"clock_khz\": %d, \"note\": \"GB/s landed in EACH SM's shared memory (per_sm) and summed over the grid (total)\"", clk);
    for (int rows : {192, 96}) {
        const CUtensorMap m1 = make_map(w, R, rows), m2 = make_map(w, R, rows / 2), m4 = make_map(w, R, rows / 4);
        for (int S : {3, 8}) {
            for (int grid : {1, 4, 37, 74, 148}) {
                double ms = run<1>(m1, R, rows, S, loads, grid);
                const double per = (double)loads * rows * 128 / (ms * 1e-3) / 1e9;
                printf(",\n \"uni_rows%d_S%d_grid%d\": {\"per_sm\": %.1f, \"total\": %.0f}", rows, S, grid, per, per * grid);
                fflush(stdout);
            }
            for (int grid : {2, 74, 148}) {
                double ms = run<2>(m2, R, rows, S, loads, grid);
                const double per = (double)loads * rows * 128 / (ms * 1e-3) / 1e9;
                printf(",\n \"mc2_rows%d_S%d_grid%d\": {\"per_sm\": %.1f, \"total\": %.0f}", rows, S, grid, per, per * grid);
                fflush(stdout);
            }
            for (int grid : {4, 148}) {
                double ms = run<4>(m4, R, rows, S, loads, grid);
                const double per = (double)loads * rows * 128 / (ms * 1e-3) / 1e9;
                printf(",\n \"mc4_rows%d_S%d_grid%d\": {\"per_sm\": %.1f, \"total\": %.0f}", rows, S, grid, per, per * grid);
                fflush(stdout);
            }
        }
    }
    printf("\n}\n");
    return 0;
}
