// Floor of the layer-B forward's memory pattern on this GPU: fused[r] = concat(dense[r] (32 floats), zeros (32 floats))
// for R = 560000 cells, written as plain streaming kernels without any CSR logic, timed like tools/microbench.py
// (back to back, and after a 256 MB flush that leaves the L2 dirty).  Build: nvcc -O3 -arch=sm_100a -o build/pattern_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// V1: one output float4 per thread per step, output-contiguous (512 B per warp store), half the lanes load
template <int ILP>
__global__ void __launch_bounds__(256) k_rowmajor(const float4* __restrict__ in, float4* __restrict__ out, long long n_out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_out; i0 += stride * ILP) {
        float4 v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const long long i = i0 + u * stride;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_out && (i & 15) < 8) v[u] = __ldcs(in + (i >> 4) * 8 + (i & 7));
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const long long i = i0 + u * stride;
            if (i < n_out) __stcs(out + i, v[u]);
        }
    }
}

// V2: the warp-tile pattern of shpl_pool_narrow_kernel: 32 cells per warp, 8 contiguous loads per lane, 8 strided
// stores (dense halves), then 8 zero stores (pooled halves)
__global__ void __launch_bounds__(256) k_tiles(const float4* __restrict__ in, float4* __restrict__ out, int n_cells) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (n_cells + 31) / 32;
    for (int t = blockIdx.x * 8 + warp; t < tiles; t += gridDim.x * 8) {
        const float4* din = in + (size_t)t * 256;
        float4* o = out + (size_t)t * 512;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(din + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = j * 32 + lane;
            __stcs(o + (s >> 3) * 16 + (s & 7), v[j]);
        }
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = j * 32 + lane;
            __stcs(o + (s >> 3) * 16 + 8 + (s & 7), z);
        }
    }
}

// V5: V2 plus the CSR offsets of the tile (one lane per cell), a ballot, and zero stores only when no cell is busy
template <bool kEarlyZero>
__global__ void __launch_bounds__(256) k_tiles_ptr(const float4* __restrict__ in, float4* __restrict__ out, const int* __restrict__ ptr,
                                                   int n_cells, int* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (n_cells + 31) / 32;
    int busy_tiles = 0;
    for (int t = blockIdx.x * 8 + warp; t < tiles; t += gridDim.x * 8) {
        const int lo = __ldg(ptr + t * 32 + lane), hi = __ldg(ptr + t * 32 + lane + 1);
        const float4* din = in + (size_t)t * 256;
        float4* o = out + (size_t)t * 512;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(din + j * 32 + lane);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kEarlyZero) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = j * 32 + lane;
                __stcs(o + (s >> 3) * 16 + 8 + (s & 7), z);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = j * 32 + lane;
            __stcs(o + (s >> 3) * 16 + (s & 7), v[j]);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, hi > lo);
        if (busy == 0u) {
            if (!kEarlyZero) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int s = j * 32 + lane;
                    __stcs(o + (s >> 3) * 16 + 8 + (s & 7), z);
                }
            }
        } else {
            ++busy_tiles;
        }
    }
    if (busy_tiles == 12345) *sink = 1;
}

// V3: like V2 but each lane writes full 256-byte rows' worth contiguously: shuffle-free variant where lane pairs own a row
__global__ void __launch_bounds__(256) k_rows(const float4* __restrict__ in, float4* __restrict__ out, int n_cells) {
    // a half-warp (16 lanes) writes one 256-byte row per step: lanes 0-7 dense, 8-15 zeros; 2 rows per warp step
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_out = (long long)n_cells * 16;
    for (long long i0 = gt; i0 < n_out; i0 += stride * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + u * stride;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_out && (i & 15) < 8) v[u] = __ldg(in + (i >> 4) * 8 + (i & 7));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + u * stride;
            if (i < n_out) out[i] = v[u];
        }
    }
}

int main() {
    const int R = 560000;
    float4 *in, *out;
    char* flush;
    CK(cudaMalloc(&in, (size_t)R * 8 * 16));
    CK(cudaMalloc(&out, (size_t)R * 16 * 16));
    CK(cudaMalloc(&flush, 256u << 20));
    CK(cudaMemset(in, 1, (size_t)R * 8 * 16));
    int *ptr, *sink;
    CK(cudaMalloc(&ptr, (size_t)(R + 64) * 4));
    CK(cudaMemset(ptr, 0, (size_t)(R + 64) * 4));
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const long long n_out = (long long)R * 16;
    struct V { const char* name; int grid; int kind; };
    std::vector<V> vs = {{"rowmajor_ilp1_g1184", 1184, 0}, {"rowmajor_ilp4_g1184", 1184, 1}, {"rowmajor_ilp8_g1184", 1184, 2},
                         {"rowmajor_ilp4_g2368", 2368, 1}, {"rowmajor_ilp4_g4736", 4736, 1}, {"rowmajor_ilp1_full", (int)((n_out + 255) / 256), 0},
                         {"tiles_g1184", 1184, 3}, {"tiles_g2188", 2188, 3}, {"tiles_g592", 592, 3},
                         {"rows_default_policy_g1184", 1184, 4}, {"rows_default_policy_g4736", 4736, 4},
                         {"tiles_ptr_g1184", 1184, 5}, {"tiles_ptr_earlyzero_g1184", 1184, 6}, {"tiles_ptr_g2188", 2188, 5},
                         {"tiles_ptr_earlyzero_g2188", 2188, 6}};
    for (auto& v : vs) {
        for (int cold = 0; cold < 2; ++cold) {
            std::vector<float> ts;
            for (int it = 0; it < 25; ++it) {
                if (cold) CK(cudaMemsetAsync(flush, it, 256u << 20));
                CK(cudaEventRecord(e0));
                switch (v.kind) {
                    case 0: k_rowmajor<1><<<v.grid, 256>>>(in, out, n_out); break;
                    case 1: k_rowmajor<4><<<v.grid, 256>>>(in, out, n_out); break;
                    case 2: k_rowmajor<8><<<v.grid, 256>>>(in, out, n_out); break;
                    case 3: k_tiles<<<v.grid, 256>>>(in, out, R); break;
                    case 5: k_tiles_ptr<false><<<v.grid, 256>>>(in, out, ptr, R, sink); break;
                    case 6: k_tiles_ptr<true><<<v.grid, 256>>>(in, out, ptr, R, sink); break;
                    default: k_rows<<<v.grid, 256>>>(in, out, R); break;
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it >= 5) ts.push_back(ms * 1e3f);
            }
            std::sort(ts.begin(), ts.end());
            printf("%-28s %s median %.1f us  min %.1f us  (%.0f GB/s of 215 MB)\n", v.name, cold ? "dirtyL2" : "back2back", ts[ts.size() / 2], ts[0],
                   215.04e6 / ts[ts.size() / 2] / 1e3);
        }
    }
    // reference: cudaMemcpy D2D of 110 MB
    for (int cold = 0; cold < 2; ++cold) {
        std::vector<float> ts;
        for (int it = 0; it < 25; ++it) {
            if (cold) CK(cudaMemsetAsync(flush, it, 256u << 20));
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(out, in, (size_t)R * 8 * 16, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 5) ts.push_back(ms * 1e3f);
        }
        std::sort(ts.begin(), ts.end());
        printf("%-28s %s median %.1f us (71.7 MB read + 71.7 MB written)\n", "cudaMemcpyD2D_72MB", cold ? "dirtyL2" : "back2back", ts[ts.size() / 2]);
    }
    return 0;
}
