// Library-wide state and the small utility entry points of the C ABI.
#include "common.cuh"
#include "../../include/nunif_b200.h"

namespace nb200 {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace nb200

using namespace nb200;

extern "C" const char* nb200_last_error(void) { return g_last_error.c_str(); }
extern "C" int nb200_abi_version(void) { return NB200_ABI_VERSION; }
extern "C" uint64_t nb200_launch_count(void) { return g_launches.load(); }

extern "C" int nb200_check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("no CUDA device: the nunif_b200 hot path has no CPU fallback");
    NB_CHECK(device >= 0 && device < n, "device index out of range");
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; this library is built for sm_100a (B200) only");
    return 0;
}
