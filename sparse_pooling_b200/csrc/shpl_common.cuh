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

}  // namespace shpl
