// Shared helpers for libshpl.so (sm_100a).  Internal header; the public C ABI is include/shpl.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "shpl.h"

namespace shpl {

// thread-local error text behind shpl_last_error()
void set_error(const char* fmt, ...);
int sm_count();
void count_launches(int n);

#define SHPL_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            ::shpl::set_error(__VA_ARGS__);           \
            return (code);                            \
        }                                             \
    } while (0)

#define SHPL_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t e_ = (expr);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            ::shpl::set_error("%s failed: %s", #expr, cudaGetErrorString(e_));          \
            return SHPL_ERR_CUDA;                                                       \
        }                                                                               \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SHPL_ERR_CUDA;
    }
    return SHPL_OK;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }


#ifdef __CUDACC__
// ---- in-kernel index checks of the DEBUG build (make debug -> libshpl_debug.so; compiled out of the product library).
// compute-sanitizer is closed on this pool, so this is the memory-safety evidence: every gather index, entry range and
// output slot the kernels compute is checked against the caller's sizes; a failure is counted (and the first few are
// printed) instead of trapping, so that a whole test run reports a total.  tools/memsafety_run.py reads the counter.
#ifdef SHPL_DEBUG_CHECKS
extern __device__ unsigned long long g_debug_failures;
#define SHPL_DASSERT(cond)                                                                                        \
    do {                                                                                                          \
        if (!(cond)) {                                                                                            \
            if (atomicAdd(&::shpl::g_debug_failures, 1ull) < 8ull)                                                \
                printf("SHPL_DASSERT failed: %s (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
        }                                                                                                         \
    } while (0)
#else
#define SHPL_DASSERT(cond) ((void)0)
#endif

// ---- single-pass prefix over CTAs (decoupled look-back), shared by the builder and the feeder
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// Single-pass exclusive prefix over tiles (decoupled look-back), executed by warp 0.
// `agg` packs (count_clip << 31 | count_nnz) of this tile; returns the packed sum of all earlier tiles.
__device__ inline unsigned long long lookback(unsigned long long* status, int tile, unsigned long long agg, int lane) {
    if (tile == 0) {
        if (lane == 0) atomicExch(status, kFlagIncl | agg);
        return 0ull;
    }
    if (lane == 0) atomicExch(status + tile, kFlagAgg | agg);
    unsigned long long excl = 0ull;
    int t = tile - 1;
    while (true) {
        const int i = t - lane;
        unsigned long long sv = kFlagIncl;       // virtual tile before tile 0: inclusive prefix 0
        do {
            if (i >= 0) sv = ld_volatile_u64(status + i);
        } while (__any_sync(kFull, (sv >> 62) == 0ull));
        const unsigned incl = __ballot_sync(kFull, (sv >> 62) == 2ull);
        const int first = incl ? (__ffs(incl) - 1) : 32;
        unsigned long long c = (lane <= first) ? (sv & kValMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
        excl += c;
        if (incl) break;
        t -= 32;
    }
    if (lane == 0) atomicExch(status + tile, kFlagIncl | (excl + agg));
    return excl;
}

#endif  // __CUDACC__

}  // namespace shpl
