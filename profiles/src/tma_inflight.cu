// Why does every TMA-fed kernel of this repo see ~0.4 us per box load per SM with ONE box in flight (profiles/r2/l2_tma_bw.json)?
// Discriminating experiments on an L2-resident matrix W [R][64] fp16 viewed as needed:
//   box     : rows in {32, 64, 128, 192, 256} x 128 B, ring 8                         -> per-op cost vs bytes
//   box3d   : (64, rows, nk) boxes of a [nk][R][64] view (one op = nk chunks)         -> do bigger ops amortise?
//   warps   : P producer warps (1, 2, 4), each with its own ring and barriers         -> is the limit per issuing warp?
//   ctas    : 1 or 2 CTAs per SM                                                      -> per CTA or per SM?
//   bulk1d  : cp.async.bulk (non-tensor, contiguous rows*128 B)                       -> tensor path specific?
//   noswz / promo : swizzle NONE, L2 promotion NONE / 128B
// Output: GB/s landed per SM.  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a ... -o profiles/_bin/tma_inflight
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
// 3-D view (64 cols, R rows, NK chunks) of a dense [NK][R][64] fp16 array
static CUtensorMap make_map3(void* base, int R, int NK, int box_rows, int box_k, CUtensorMapSwizzle sw, CUtensorMapL2promotion promo) {
    CUtensorMap m;
    cuuint64_t dims[3] = {64, (cuuint64_t)R, (cuuint64_t)NK};
    cuuint64_t str[2] = {128, (cuuint64_t)R * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_k}, es[3] = {1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("# encode failed %d\n", (int)r);
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

// P producer warps (blockDim = 32*P); warp p owns ring p (S slots of op_bytes) and issues `loads` ops.
// mode 0: 3-D tensor box load; mode 1: 1-D bulk copy of op_bytes contiguous bytes
__global__ void __launch_bounds__(128) inflight_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* base, int R, int NK,
                                                       int rows, int box_k, int S, int loads, int mode, int start_stride) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[4][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, P = blockDim.x >> 5;
    const int op_bytes = rows * 128 * box_k;
    if (threadIdx.x == 0) {
        for (int p = 0; p < P; ++p)
            for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[p][i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane == 0) {
        uint8_t* ring = smem + (size_t)warp * S * op_bytes;
        int issued = 0, waited = 0;
        int row = (int)(((long long)(blockIdx.x * P + warp) * start_stride) % R) / rows * rows, kk = 0;
        while (waited < loads) {
            while (issued < loads && issued - waited < S) {
                const int slot = issued % S;
                const uint32_t fb = smem_u32(&full[warp][slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(op_bytes) : "memory");
                if (mode == 0) {
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(smem_u32(ring + slot * op_bytes)), "l"(&map), "r"(fb), "r"(0), "r"(row), "r"(kk) : "memory");
                } else {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(ring + slot * op_bytes)), "l"(base + ((size_t)kk * R + row) * 128), "r"(op_bytes), "r"(fb) : "memory");
                }
                row += rows;
                if (row + rows > R) { row = 0; kk += box_k; if (kk + box_k > NK) kk = 0; }
                ++issued;
            }
            mbar_wait(smem_u32(&full[warp][waited % S]), (uint32_t)((waited / S) & 1));
            ++waited;
        }
    }
}

static double run(const CUtensorMap& m, const uint8_t* base, int R, int NK, int rows, int box_k, int S, int loads, int P, int ctas_per_sm,
                  int mode, int sms, int start_stride = 0) {
    if (!start_stride) start_stride = 37 * rows;
    const size_t smem = (size_t)P * S * rows * 128 * box_k + 1024;
    cudaFuncSetAttribute(inflight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) inflight_kernel<<<grid, 32 * P, smem>>>(m, base, R, NK, rows, box_k, S, loads, mode, start_stride);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int iters = 4;
    for (int i = 0; i < iters; ++i) inflight_kernel<<<grid, 32 * P, smem>>>(m, base, R, NK, rows, box_k, S, loads, mode, start_stride);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("\n# CUDA error: %s (rows=%d k=%d S=%d P=%d cps=%d mode=%d)\n", cudaGetErrorString(e), rows, box_k, S, P, ctas_per_sm, mode); return -1; }
    ms /= iters;
    // GB/s landed per SM
    return (double)loads * P * ctas_per_sm * rows * 128.0 * box_k / (ms * 1e-3) / 1e9;
}

int main() {
    const int R = 960, NK = 3;   // [3][960][64] fp16 = 360 KB, L2 resident
    uint8_t* w;
    cudaMalloc(&w, (size_t)R * NK * 128);
    cudaMemset(w, 0, (size_t)R * NK * 128);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const CUtensorMapSwizzle SW = CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapL2promotion PR = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    printf("{\"unit\": \"GB/s landed per SM (all %d SMs active unless sms=1)\"", sms);
    const int loads = 3000;
    for (int rows : {32, 64, 128, 192, 256}) {
        CUtensorMap m = make_map3(w, R, NK, rows, 1, SW, PR);
        const int S = rows * 128 * 8 <= 200 * 1024 ? 8 : 6;
        printf(",\n \"box_rows%d_S%d\": %.1f", rows, S, run(m, w, R, NK, rows, 1, S, loads, 1, 1, 0, sms));
        printf(", \"box_rows%d_S%d_1sm\": %.1f", rows, S, run(m, w, R, NK, rows, 1, S, loads, 1, 1, 0, 1));
        fflush(stdout);
    }
    for (int bk : {2, 3}) {
        CUtensorMap m = make_map3(w, R, NK, 192, bk, SW, PR);
        printf(",\n \"box3d_rows192_k%d_S2\": %.1f", bk, run(m, w, R, NK, 192, bk, 2, loads, 1, 1, 0, sms));
        CUtensorMap m2 = make_map3(w, R, NK, 64, bk, SW, PR);
        printf(", \"box3d_rows64_k%d_S8\": %.1f", bk, run(m2, w, R, NK, 64, bk, 8, loads, 1, 1, 0, sms));
        fflush(stdout);
    }
    {
        CUtensorMap m = make_map3(w, R, NK, 128, 1, SW, PR);
        for (int P : {1, 2, 4}) printf(",\n \"warps%d_rows128_S3\": %.1f", P, run(m, w, R, NK, 128, 1, 3, loads, P, 1, 0, sms));
        for (int c : {1, 2, 4}) printf(",\n \"ctas_per_sm%d_rows128_S3\": %.1f", c, run(m, w, R, NK, 128, 1, 3, loads, 1, c, 0, sms));
        fflush(stdout);
        for (int rows : {64, 192}) printf(",\n \"bulk1d_rows%d_S6\": %.1f", rows, run(m, w, R, NK, rows, 1, 6, loads, 1, 1, 1, sms));
        printf(",\n \"bulk1d_rows192_S6_4warps\": %.1f", run(m, w, R, NK, 192, 1, 2, loads, 4, 1, 1, sms));
        CUtensorMap mn = make_map3(w, R, NK, 128, 1, CU_TENSOR_MAP_SWIZZLE_NONE, PR);
        printf(",\n \"noswizzle_rows128_S6\": %.1f", run(mn, w, R, NK, 128, 1, 6, loads, 1, 1, 0, sms));
        CUtensorMap mp = make_map3(w, R, NK, 128, 1, SW, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        printf(",\n \"promo_none_rows128_S6\": %.1f", run(mp, w, R, NK, 128, 1, 6, loads, 1, 1, 0, sms));
        CUtensorMap mq = make_map3(w, R, NK, 128, 1, SW, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        printf(",\n \"promo_128_rows128_S6\": %.1f", run(mq, w, R, NK, 128, 1, 6, loads, 1, 1, 0, sms));
        fflush(stdout);
    }
    // a larger (HBM-resident) source for comparison: 1.5 GB
    {
        const int R2 = 4 << 20;   // 4M rows x 128 B = 512 MB per chunk
        uint8_t* big;
        if (cudaMalloc(&big, (size_t)R2 * 3 * 128) == cudaSuccess) {
            cudaMemset(big, 0, (size_t)R2 * 3 * 128);
            CUtensorMap m = make_map3(big, R2, 3, 128, 1, SW, PR);
            printf(",\n \"hbm_rows128_S8\": %.1f", run(m, big, R2, 3, 128, 1, 8, loads, 1, 1, 0, sms, R2 / sms));
            CUtensorMap m3 = make_map3(big, R2, 3, 256, 1, SW, PR);
            printf(", \"hbm_rows256_S6\": %.1f", run(m3, big, R2, 3, 256, 1, 6, loads, 1, 1, 0, sms, R2 / sms));
            printf(", \"hbm_bulk1d_rows256_S6\": %.1f", run(m3, big, R2, 3, 256, 1, 6, loads, 1, 1, 1, sms, R2 / sms));
        }
    }
    printf("\n}\n");
    return 0;
}
