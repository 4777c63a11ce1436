// How long does ONE tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) occupy the tensor pipe as a function of N, of the
// number of accumulators the issue stream rotates over, and of the operand swizzle?  The fused Swin kernels issue ~100 small
// UMMAs per 128-token tile; their timelines (profiles/r2/attn_tc_timeline_v*.txt) show ~200 cycles per instruction.
// One CTA per SM, one issuing thread, operands = zero-filled shared memory (no loads), `n` UMMAs then one commit.
// Output: cycles per UMMA (issue -> commit observed) and the issue-loop cycles per UMMA.
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/src/umma_rate.cu -o profiles/_bin/umma_rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 26) && !ok; ++it)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
    if (!ok) { printf("timeout block %d\n", blockIdx.x); __trap(); }
}
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, int swz) {
    const uint64_t layout = swz == 128 ? 2 : (swz == 64 ? 4 : 6);
    const uint64_t sbo = (8 * swz) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// n UMMAs (n % 4 == 0) issued by one elected lane of warp 0 from precomputed descriptors (4 K-steps of one 64-wide stage), rotating
// over NACC accumulators; COMMIT: a tcgen05.commit (to a barrier nobody waits on) after every group of 4.
template <int NACC, bool COMMIT>
__global__ void __launch_bounds__(128) rate_kernel(int N, int n, int swz, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    if (threadIdx.x < 32) {
        const uint32_t aA = smem_u32(smem), aB = aA + 128 * 128;
        const uint32_t id = idesc(N);
        uint64_t ad[4], bd[4];
        uint32_t td[4];
        for (int k = 0; k < 4; ++k) { ad[k] = kdesc(aA + k * 32, swz); bd[k] = kdesc(aB + k * 32, swz); td[k] = tb + (uint32_t)((k % NACC) * (512 / NACC)); }
        uint32_t ph = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            long long t1 = 0;
            if (pred) {
                for (int i = 0; i < n; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma(td[j], ad[j], bd[j], id, (i > 0 || j >= NACC) ? 1u : 0u);
                    if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
                }
                t1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
            }
            __syncwarp();
            mbar_wait(smem_u32(&bar[0]), ph);
            ph ^= 1;
            const long long t2 = clock64();
            if (blockIdx.x == 0 && rep == 2 && pred) { out[0] = t2 - t0; out[1] = t1 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int n = 1024;
    printf("{\"unit\": \"cycles per UMMA (M=128, K=16, kind::f16), total = issue..commit observed / n, issue = issue loop / n; n = %d\",\n", n);
    auto run = [&](const char* name, auto kern, int N, int swz, int grid, bool lastrow) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        kern<<<grid, 128, 98 * 1024>>>(N, n, swz, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) printf(" \"%s\": \"%s\"%s\n", name, cudaGetErrorString(e), lastrow ? "" : ",");
        else printf(" \"%s_N%d_sw%d_grid%d\": {\"total\": %.1f, \"issue\": %.1f}%s\n", name, N, swz, grid, (double)h[0] / n, (double)h[1] / n, lastrow ? "" : ",");
    };
    const int Ns[] = {16, 32, 48, 64, 96, 112, 128, 192, 256};
    for (int N : Ns) run("chain", rate_kernel<1, false>, N, 128, sms, false);
    for (int N : Ns) run("rot2", rate_kernel<2, false>, N, 128, sms, false);
    for (int N : Ns) if (N <= 128) run("rot4", rate_kernel<4, false>, N, 128, sms, false);
    run("chain_sw64", rate_kernel<1, false>, 112, 64, sms, false);
    run("chain_sw64", rate_kernel<1, false>, 48, 64, sms, false);
    for (int N : {48, 96, 192}) run("chain_commit4", rate_kernel<1, true>, N, 128, sms, false);
    run("rot2_commit4", rate_kernel<2, true>, 96, 128, sms, false);
    run("one_sm_chain", rate_kernel<1, false>, 96, 128, 1, false);
    run("one_sm_chain", rate_kernel<1, false>, 256, 128, 1, true);
    printf("}\n");
    return 0;
}
