// libshpl.so: error reporting and device queries behind the C ABI (include/shpl.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "shpl_common.cuh"

namespace shpl {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;  // B200
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace shpl

#ifdef SHPL_DEBUG_CHECKS
namespace shpl { __device__ unsigned long long g_debug_failures; }      // zero-initialised; one copy for all translation units (-rdc)
#endif

// Failed in-kernel index checks so far: always 0 in the product library (the checks are compiled out); the debug
// build (make debug) counts them on the device.  Synchronises the device.
extern "C" int64_t shpl_debug_check_failures(void) {
#ifdef SHPL_DEBUG_CHECKS
    unsigned long long n = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&n, shpl::g_debug_failures, sizeof(n)) != cudaSuccess) return -1;
    return (int64_t)n;
#else
    return 0;
#endif
}

// 1 when the library was built with the in-kernel checks (libshpl_debug.so), else 0.
extern "C" int shpl_debug_checks_enabled(void) {
#ifdef SHPL_DEBUG_CHECKS
    return 1;
#else
    return 0;
#endif
}

extern "C" int shpl_abi_version(void) { return SHPL_ABI_VERSION; }
extern "C" const char* shpl_last_error(void) { return shpl::g_error; }
namespace shpl { uint64_t launches(); }
extern "C" uint64_t shpl_kernel_launches(void) { return shpl::launches(); }
