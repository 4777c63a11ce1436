// Shared helpers for the nunif_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

namespace nb200 {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(const std::string& msg) {
    g_last_error = msg;
    return 1;
}

#define NB_CHECK(cond, msg)                                                     \
    do {                                                                        \
        if (!(cond)) return ::nb200::fail(std::string(__func__) + ": " + (msg)); \
    } while (0)

#define NB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::nb200::fail(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e)); \
    } while (0)

// call after every kernel launch: counts it and surfaces launch-config errors
#define NB_LAUNCHED()                                                                      \
    do {                                                                                   \
        ::nb200::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return ::nb200::fail(std::string(__func__) + ": launch -> " + cudaGetErrorString(_e)); \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

}  // namespace nb200
