// Issue/throughput microbenchmark: scalar FFMA (3-register form), FFMA with an immediate operand, packed FFMA2
// (fma.rn.f32x2) and MUFU.EX2 on sm_100a; clk per warp instruction per SMSP with 8 independent chains per thread.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, float a0, float b0, int iters) {
    float x[8], y[8];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a0 + i + threadIdx.x; y[i] = b0 + i; float2 t = make_float2(x[i], y[i]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
    const float a = a0 * 1.0001f, b = b0;
    float2 ab2 = make_float2(a, a), bb2 = make_float2(b, b);
    const unsigned long long A2 = *reinterpret_cast<unsigned long long*>(&ab2), B2 = *reinterpret_cast<unsigned long long*>(&bb2);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                         // 3-register FFMA
            else if (MODE == 1) x[i] = fmaf(x[i], a, 0.3333f);              // immediate operand
            else if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(A2), "l"(B2));
            else if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            else { x[i] = fmaf(x[i], a, b); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y[i])); }   // FFMA + MUFU interleaved
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += x[i] + y[i] + t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int MODE>
static void run(const char* name, int warps_per_smsp) {
    float* d;
    cudaMalloc(&d, 1 << 20);
    const int iters = 4096, threads = 32 * 4 * warps_per_smsp;
    k<MODE><<<148, threads>>>(d, 1.0f, 0.5f, iters);
    k<MODE><<<148, threads>>>(d, 1.0f, 0.5f, iters);
    cudaDeviceSynchronize();
    float clk = 0;
    cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
    const double ninst = (double)iters * 8 * (MODE == 4 ? 2 : 1) * warps_per_smsp;   // warp instructions per SMSP
    printf("\"%s_w%d\": %.3f, ", name, warps_per_smsp, clk / ninst);
    cudaFree(d);
}

int main() {
    printf("{");
    for (int w : {1, 2, 4}) {
        if (w == 1) { run<0>("ffma_reg", 1); run<1>("ffma_imm", 1); run<2>("ffma2", 1); run<3>("mufu_ex2", 1); run<4>("ffma+mufu", 1); }
        if (w == 2) { run<0>("ffma_reg", 2); run<1>("ffma_imm", 2); run<2>("ffma2", 2); run<3>("mufu_ex2", 2); run<4>("ffma+mufu", 2); }
        if (w == 4) { run<0>("ffma_reg", 4); run<1>("ffma_imm", 4); run<2>("ffma2", 4); run<3>("mufu_ex2", 4); run<4>("ffma+mufu", 4); }
    }
    printf("\"unit\": \"clk per warp instruction per SMSP\"}\n");
    return 0;
}
