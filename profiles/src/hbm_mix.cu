// HBM traffic-mix microbenchmark (profiles/r1/hbm_mix.json): what a streaming kernel can reach for write-only and
// read:write mixes with different store instructions.  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE>
__device__ __forceinline__ void store4(float4* p, float4 v) {
    if (MODE == 0) *p = v;
    else if (MODE == 1) __stcs(p, v);
    else if (MODE == 2) __stwt(p, v);
    else __stcg(p, v);
}

// NR reads and NW writes of n float4 each (separate arrays), one-shot grid
template <int NR, int NW, int MODE>
__global__ void __launch_bounds__(256) mix_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        float4 t = __ldcs(in + (size_t)r * n + i);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) store4<MODE>(out + (size_t)w * n + i, v);
}

// persistent grid-stride variant, UNR float4 per thread per step
template <int NR, int NW, int MODE>
__global__ void __launch_bounds__(256) mix_persistent(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            float4 t = __ldcs(in + (size_t)r * n + i);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) store4<MODE>(out + (size_t)w * n + i, v);
    }
}

// write-only through the TMA unit: each CTA fills a 32 KB shared buffer once and bulk-stores it repeatedly
__global__ void __launch_bounds__(128) bulk_store_kernel(uint8_t* out, size_t bytes) {
    extern __shared__ __align__(128) uint8_t buf[];
    constexpr int CH = 32768;
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x) reinterpret_cast<float4*>(buf)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
        for (size_t off = (size_t)blockIdx.x * CH; off + CH <= bytes; off += (size_t)gridDim.x * CH) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off), "r"(s), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static double time_ms(F launch, int iters = 10) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters;
}

int main() {
    const size_t n = (size_t)1 << 26;  // float4 per array = 1 GiB
    float4 *in, *out;
    cudaMalloc(&in, n * 16 * 2);
    cudaMalloc(&out, n * 16 * 3);
    cudaMemset(in, 0, n * 16 * 2);
    const unsigned g = (unsigned)((n + 255) / 256);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{");
#define RUN(name, NR, NW, expr) { double ms = time_ms([&] { expr; }); \
    printf("\"%s\": %.1f, ", name, (double)(NR + NW) * n * 16 / (ms * 1e-3) / 1e9); fflush(stdout); }
    RUN("w1_st", 0, 1, (mix_kernel<0, 1, 0><<<g, 256>>>(in, out, n)))
    RUN("w1_stcs", 0, 1, (mix_kernel<0, 1, 1><<<g, 256>>>(in, out, n)))
    RUN("w1_stwt", 0, 1, (mix_kernel<0, 1, 2><<<g, 256>>>(in, out, n)))
    RUN("w1_stcg", 0, 1, (mix_kernel<0, 1, 3><<<g, 256>>>(in, out, n)))
    RUN("w1_persist8", 0, 1, (mix_persistent<0, 1, 0><<<sms * 8, 256>>>(in, out, n)))
    RUN("w2_st", 0, 2, (mix_kernel<0, 2, 0><<<g, 256>>>(in, out, n)))
    RUN("w3_st", 0, 3, (mix_kernel<0, 3, 0><<<g, 256>>>(in, out, n)))
    RUN("r1_w1_st", 1, 1, (mix_kernel<1, 1, 0><<<g, 256>>>(in, out, n)))
    RUN("r1_w1_stcs", 1, 1, (mix_kernel<1, 1, 1><<<g, 256>>>(in, out, n)))
    RUN("r1_w2_st", 1, 2, (mix_kernel<1, 2, 0><<<g, 256>>>(in, out, n)))
    RUN("r1_w2_stcs", 1, 2, (mix_kernel<1, 2, 1><<<g, 256>>>(in, out, n)))
    RUN("r1_w2_persist8", 1, 2, (mix_persistent<1, 2, 0><<<sms * 8, 256>>>(in, out, n)))
    RUN("r1_w3_st", 1, 3, (mix_kernel<1, 3, 0><<<g, 256>>>(in, out, n)))
    RUN("r2_w1_st", 2, 1, (mix_kernel<2, 1, 0><<<g, 256>>>(in, out, n)))
    cudaFuncSetAttribute(bulk_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    RUN("w1_tma_bulk_1cta", 0, 1, (bulk_store_kernel<<<sms, 128, 32768>>>((uint8_t*)out, n * 16)))
    RUN("w1_tma_bulk_4cta", 0, 1, (bulk_store_kernel<<<sms * 4, 128, 32768>>>((uint8_t*)out, n * 16)))
    printf("\"unit\": \"GB/s (reads+writes)\"}\n");
    return 0;
}
