// Does a TMA (cp.async.bulk.tensor) data path beat the LDG.128 / STG.128 stream CTAs on layer B forward's dense traffic?
//   fused[r] = concat(dense[r] (32 floats), zeros (32 floats)),  R = 560 000 cells: 71.7 MB read, 143.4 MB written.
// Variant A: the warp-tile LDG/STG pattern of shpl_pool_sparse_kernel's stream CTAs (8 x 128-bit loads in flight per lane,
//            streaming cache hints).
// Variant B: one thread per CTA drives the copy engine: 2-D tensor-map loads of [rows x 128 B] boxes into a ring of
//            shared-memory stages (mbarrier complete_tx), 2-D tensor-map stores of each stage into the 256-byte-pitch
//            fused layout, plus a bulk store of a zeroed stage for the pooled half; stages are recycled with
//            cp.async.bulk.wait_group.read.  No thread ever touches the data.
// Both run back to back on rotating buffers larger than L2, timed with CUDA events.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/tma_stream_probe tools/probe/tma_stream_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_ldg_stg(const float4* __restrict__ in, float4* __restrict__ out, int n_cells) {
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}

template <int kRows, int kStages>
__global__ void __launch_bounds__(32, 1) k_tma(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, int n_cells) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kStageBytes = kRows * 128;
    // layout: [kStages][kStageBytes] data, [kStageBytes] zeros, barriers
    uint8_t* zeros = smem + kStages * kStageBytes;
    const uint32_t base = smem_u32(smem), zaddr = smem_u32(zeros), bar0 = zaddr + kStageBytes;
    for (int i = threadIdx.x; i < kStageBytes / 16; i += 32) reinterpret_cast<float4*>(zeros)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0)
        for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x != 0) return;
    const int tiles = (n_cells + kRows - 1) / kRows;
    const int n_mine = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto row0 = [&](int i) { return ((int)blockIdx.x + i * (int)gridDim.x) * kRows; };
    auto load = [&](int i) {
        const int s = i % kStages;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(kStageBytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(base + s * kStageBytes), "l"(&map_in), "r"(bar0 + 8 * s), "r"(0), "r"(row0(i)) : "memory");
    };
    const int ahead = kStages - 2;            // loads in flight; two stages may still be read by stores
    for (int i = 0; i < ahead && i < n_mine; ++i) load(i);
    for (int i = 0; i < n_mine; ++i) {
        const int s = i % kStages;
        mbar_wait(bar0 + 8 * s, (uint32_t)((i / kStages) & 1));
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(&map_out), "r"(base + s * kStageBytes), "r"(0), "r"(row0(i)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(&map_out), "r"(zaddr), "r"(32), "r"(row0(i)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (i + ahead < n_mine) {
            // the stage tile i+ahead lands in was last stored from by tile i+ahead-kStages = i-2: allow 1 group (tile i) pending... keep 1
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(i + ahead);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make2d(EncodeTiledFn fn, void* base, uint64_t cols, uint64_t rows, uint64_t pitch_floats, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_floats * 4};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

template <typename F>
static float time_us(F f, int warm, int reps) {
    for (int i = 0; i < warm; ++i) f(i);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f(i);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms * 1e3f / reps;
}

int main() {
    const int R = 560000;
    const int kSets = 3;                       // 3 x (71.7 + 143.4) MB rotating: larger than the 126 MB L2
    void* p;
    EncodeTiledFn fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    fn = reinterpret_cast<EncodeTiledFn>(p);
    std::vector<float*> in(kSets), out(kSets);
    for (int s = 0; s < kSets; ++s) {
        CK(cudaMalloc(&in[s], (size_t)R * 32 * 4));
        CK(cudaMalloc(&out[s], (size_t)R * 64 * 4));
        CK(cudaMemset(in[s], 0x3c, (size_t)R * 32 * 4));
    }
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double bytes = (double)R * (32 + 64) * 4;
    for (int cps : {4, 6, 8, 12}) {
        const float us = time_us([&](int i) { k_ldg_stg<<<sms * cps, 256>>>(reinterpret_cast<const float4*>(in[i % kSets]), reinterpret_cast<float4*>(out[i % kSets]), R); }, 5, 40);
        printf("A  LDG/STG stream tiles, %2d CTAs/SM: %7.2f us  %7.1f GB/s\n", cps, us, bytes / us / 1e3);
    }
    auto run_tma = [&](auto kern, int rows, int stages, const char* name) {
        std::vector<CUtensorMap> mi(kSets), mo(kSets);
        for (int s = 0; s < kSets; ++s) {
            mi[s] = make2d(fn, in[s], 32, R, 32, rows);
            mo[s] = make2d(fn, out[s], 64, R, 64, rows);
        }
        const int smem = (stages + 1) * rows * 128 + 8 * stages + 128;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int cps : {1, 2}) {
            if (cps * smem > 220 * 1024) continue;
            const float us = time_us([&](int i) { kern<<<sms * cps, 32, smem>>>(mi[i % kSets], mo[i % kSets], R); }, 5, 40);
            CK(cudaGetLastError());
            printf("B  TMA %s, %d CTA/SM: %7.2f us  %7.1f GB/s\n", name, cps, us, bytes / us / 1e3);
        }
    };
    run_tma(k_tma<64, 12>, 64, 12, "box  64 rows x 128 B, 12 stages");
    run_tma(k_tma<128, 8>, 128, 8, "box 128 rows x 128 B,  8 stages");
    run_tma(k_tma<256, 6>, 256, 6, "box 256 rows x 128 B,  6 stages");
    run_tma(k_tma<256, 3>, 256, 3, "box 256 rows x 128 B,  3 stages");
    // correctness of the TMA variant (one set)
    {
        CK(cudaMemset(out[0], 0xff, (size_t)R * 64 * 4));
        CUtensorMap mi = make2d(fn, in[0], 32, R, 32, 128), mo = make2d(fn, out[0], 64, R, 64, 128);
        k_tma<128, 8><<<sms, 32, 9 * 128 * 128 + 64 + 128>>>(mi, mo, R);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h((size_t)R * 64);
        CK(cudaMemcpy(h.data(), out[0], h.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t r = 0; r < (size_t)R; ++r)
            for (int c = 0; c < 64; ++c) bad += h[r * 64 + c] != (c < 32 ? 0x3c3c3c3cu : 0u);
        printf("TMA variant output check: %zu wrong words of %zu\n", bad, h.size());
    }
    return 0;
}
