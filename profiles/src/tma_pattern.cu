// What HBM throughput do the persistent GEMM's TMA access patterns reach without any math?
// [T][N] fp16 matrices, 128-row tiles, one persistent CTA per SM (like gemm_conv_persistent):
//   store: N/CW box stores of [128 rows][CW cols] per tile (CW*2 bytes contiguous per row), 4-deep staging ring
//   load : K/CW box loads of the same shape per tile into a ring, mbarrier-completed
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lcuda (driver entry point fetched at run time).
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
static CUtensorMap make_map(void* base, long long T, int N, int cw, int rows, int swz) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
    cuuint64_t str[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)rows}, es[2] = {1, 1};
    CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (N=%d cw=%d)\n", (int)r, N, cw); }
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// stores only: every tile = nch chunk stores
__global__ void __launch_bounds__(128) store_kernel(const __grid_constant__ CUtensorMap map, int m_tiles, int nch, int cw, int ring) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int chb = 128 * cw * 2;
    for (int i = threadIdx.x; i < ring * chb / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        int slot = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x)
            for (int c = 0; c < nch; ++c) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                             "r"(smem_u32(smem + slot * chb)), "r"(c * cw), "r"(t * 128) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (ring >= 8) asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                if (++slot == ring) slot = 0;
            }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// loads only: every tile = nch chunk loads into a ring; consumer = the same thread waiting on the mbarrier
__global__ void __launch_bounds__(128) load_kernel(const __grid_constant__ CUtensorMap map, int m_tiles, int nch, int cw, int ring) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[16];
    const int chb = 128 * cw * 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ring; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // issue `ring` loads ahead, then wait oldest / issue next
        long long total = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) total += nch;
        long long issued = 0, waited = 0;
        int t_i = blockIdx.x, c_i = 0;
        while (waited < total) {
            while (issued < total && issued - waited < ring) {
                const int slot = (int)(issued % ring);
                const uint32_t b = smem_u32(&bar[slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(chb) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(smem + slot * chb)), "l"(&map), "r"(b), "r"(c_i * cw), "r"(t_i * 128) : "memory");
                if (++c_i == nch) { c_i = 0; t_i += gridDim.x; }
                ++issued;
            }
            const int slot = (int)(waited % ring);
            const uint32_t ph = (uint32_t)((waited / ring) & 1);
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&bar[slot])), "r"(ph) : "memory");
            ++waited;
        }
    }
}

template <typename F>
static double time_ms(F launch, int iters = 10) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms / iters;
}

int main() {
    const long long T = 921600 * 2;   // rows (2 x one 16-tile batch, so every matrix is far larger than L2)
    const int m_tiles = (int)(T / 128);
    void* buf;
    cudaMalloc(&buf, (size_t)T * 576 * 2);
    cudaMemset(buf, 0, (size_t)T * 576 * 2);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("{");
    struct Cfg { const char* name; int N, cw, swz; };
    const Cfg cfgs[] = {{"N192_cw64", 192, 64, 128}, {"N576_cw64", 576, 64, 128}, {"N96_cw32", 96, 32, 64}, {"N64_cw64", 64, 64, 128},
                        {"N192_cw192_noswz", 192, 192, 0}, {"N384_cw64", 384, 64, 128}, {"N256_cw256_noswz", 256, 256, 0}};
    for (const Cfg& c : cfgs) {
        CUtensorMap m = make_map(buf, T, c.N, c.cw, 128, c.swz);
        const int nch = c.N / c.cw, chb = 128 * c.cw * 2;
        for (int ring : {4, 8}) {
            if ((size_t)ring * chb > 190 * 1024) continue;
            for (int mult : {1, 2}) {
                if ((size_t)ring * chb * mult > 200 * 1024) continue;
                const size_t smem = (size_t)ring * chb + 1024;
                double ms = time_ms([&] { store_kernel<<<sms * mult, 128, smem>>>(m, m_tiles, nch, c.cw, ring); });
                printf("\"store_%s_ring%d_x%d\": %.0f, ", c.name, ring, mult, (double)T * c.N * 2 / (ms * 1e-3) / 1e9);
                ms = time_ms([&] { load_kernel<<<sms * mult, 128, smem>>>(m, m_tiles, nch, c.cw, ring); });
                printf("\"load_%s_ring%d_x%d\": %.0f, ", c.name, ring, mult, (double)T * c.N * 2 / (ms * 1e-3) / 1e9);
                fflush(stdout);
            }
        }
    }
    printf("\"unit\": \"GB/s\"}\n");
    return 0;
}
