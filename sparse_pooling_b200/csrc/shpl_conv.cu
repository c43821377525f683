// Post-fusion 3x3 convolution fused with the pooling, sm_100a.
//
// Replaces, for the `rpn_sparse_pooling_conv_after_fusion` switch,
//   /root/reference/avod/avod/core/models/rpn_model.py:335-354      sparse_pool_layer -> slim.conv2d(fused, C, [3,3])
//   /root/reference/avod/avod/core/models/retinanet_model.py:337-348
// without ever writing the fused (concat) map:
//   conv(concat(dst, pooled)) = conv(dst; W[:, :, :C_d, :]) + conv(pooled; W[:, :, C_d:, :])
// The first term is a dense implicit GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32
// accumulators in TMEM, halo tiles brought in by TMA with the SAME padding supplied by the TMA's out-of-bounds zero
// fill, output tiles written back by TMA); fp32 parity comes from the 3xTF32 split (x = hi + lo, both tf32:
// hi*hi + lo*hi + hi*lo, error ~2^-21 per product).  The second term only touches the 3x3 neighbourhoods of the few
// cells that receive pooled features (~2 % at KITTI stride 1): a bitmap of those neighbourhoods is made first, the
// dense kernel leaves the marked cells un-activated, and a small warp-per-cell kernel adds W_pooled . pooled there
// (the pooled sums formed in the reference's entry order, like the pooling kernels do) and applies the epilogue.
//
// Dense kernel, one persistent CTA per SM, 6 warps:
//   warp 4  TMA producer   halo tile [18 x 10 pixels x 32 ch] fp32 -> staging (2 stages)
//   warps 0-3 workers      staging -> hi / lo operand planes ([chunk of 4 ch][halo pixel][16 B]: the K-major,
//                          no-swizzle core-matrix layout; a tap (dy, dx) is just a different start address), and
//                          the epilogue TMEM -> registers -> swizzled smem -> TMA store
//   warp 5  MMA issuer     per tile 9 taps x 4 K-steps x {A_hi x [B_hi | B_lo] (N = 64), A_lo x B_hi (N = 32)}
// Tile = 16 rows x 8 columns of output pixels = the 128 rows (TMEM lanes) of one MMA.
#include <cuda.h>

#include "shpl_common.cuh"

namespace {

constexpr int kC = 32;                     // dense input channels = output channels of this kernel
constexpr int kTileY = 16, kTileX = 8;     // output pixels per tile (M = 128)
constexpr int kHaloY = kTileY + 2, kHaloX = kTileX + 2, kHaloPix = kHaloY * kHaloX;   // 18 x 10 = 180
constexpr int kChunks = kC / 4;            // 16-byte channel chunks per pixel
constexpr int kPlanePitch = (kHaloPix + 1) * 16;     // bytes between channel chunks (+16: bank spread)
constexpr int kPlaneBytes = kChunks * kPlanePitch;   // one hi or lo plane set
constexpr int kOpndBytes = 2 * kPlaneBytes;          // hi + lo
constexpr int kStageBytes = kHaloPix * kC * 4;       // 23040: one TMA box
constexpr int kWChunkBytes = 64 * 16;                // [n = 64 (hi 32 | lo 32)][4 ci]
constexpr int kWTapBytes = kChunks * kWChunkBytes;   // 8192
constexpr int kWBytes = 9 * kWTapBytes;              // 73728
constexpr int kOutBytes = kTileY * kTileX * kC * 4;  // 16384
constexpr int kConvThreads = 192;
constexpr int kWorkers = 128;
constexpr int kTmemCols = 128;             // 2 accumulator buffers x 64 columns

// dynamic shared memory map (byte offsets from a 1024-aligned base)
constexpr int kSmOut = 0;                                  // 16384, 1024-aligned (128B-swizzled TMA store source)
constexpr int kSmStage = kSmOut + kOutBytes;               // 2 x 23040
constexpr int kSmW = kSmStage + 2 * kStageBytes;           // 73728
constexpr int kSmOpnd = kSmW + kWBytes;                    // 2 x 46336
constexpr int kSmBar = kSmOpnd + 2 * kOpndBytes;           // mbarriers
constexpr int kSmScale = kSmBar + 128;                     // scale[32], shift[32]
constexpr int kSmEnd = kSmScale + 256;
constexpr int kConvSmem = kSmEnd + 1024;                   // + slack for the 1024-byte alignment of the base
static_assert(kConvSmem <= 227 * 1024, "dense conv kernel: shared memory budget");
static_assert(kSmStage % 128 == 0 && kSmW % 128 == 0 && kSmOpnd % 16 == 0 && kSmBar % 8 == 0, "smem alignment");

enum { BAR_STAGE_FULL = 0, BAR_STAGE_EMPTY = 2, BAR_OPND_FULL = 4, BAR_OPND_EMPTY = 6, BAR_ACC_FULL = 8, BAR_ACC_EMPTY = 10,
       BAR_W = 12, BAR_COUNT = 13 };

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol error traps (the launch fails) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 bytes apart; `sbo` = bytes between 8-row groups
// (M / N direction), `lbo` = bytes between the two 16-byte K chunks of one tf32 MMA (K = 8).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: fp32 accumulate, tf32 x tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

#define SHPL_TMEM_LD32(taddr, v)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),            \
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),            \
                   "=r"(v[30]), "=r"(v[31])                                                                              \
                 : "r"(taddr))

struct ConvArgs {
    const float* wprep;          // [9][8][64][4] hi | lo weights in the B-operand layout (shpl_conv_prep_kernel)
    const float* scale;          // [32] or NULL (1)
    const float* shift;          // [32] or NULL (0)
    const uint32_t* active;      // bitmap over cells, or NULL: marked cells are stored un-activated (the sparse kernel finishes them)
    int relu;
    int frames, H, W;
    int tiles_x, tiles_y, n_tiles;
};

// ------------------------------------------------------------------------------------ dense kernel
__global__ void __launch_bounds__(kConvThreads, 1)
shpl_conv3x3_dense_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, ConvArgs a) {
    extern __shared__ uint8_t conv_smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t raw = smem_u32(conv_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = conv_smem_raw + (base - raw);
    const uint32_t bar0 = base + kSmBar;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    float* sc = reinterpret_cast<float*>(gbase + kSmScale);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(BAR_STAGE_FULL + s), 1);
            mbar_init(bar(BAR_STAGE_EMPTY + s), 4);
            mbar_init(bar(BAR_OPND_FULL + s), 4);
            mbar_init(bar(BAR_OPND_EMPTY + s), 1);
            mbar_init(bar(BAR_ACC_FULL + s), 1);
            mbar_init(bar(BAR_ACC_EMPTY + s), 4);
        }
        mbar_init(bar(BAR_W), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) sc[tid] = tid < 32 ? (a.scale ? a.scale[tid] : 1.f) : (a.shift ? a.shift[tid - 32] : 0.f);
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int n_mine = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    auto tile_coord = [&](int i, int& f, int& y0, int& x0) {
        int t = (int)blockIdx.x + i * (int)gridDim.x;
        const int per_frame = a.tiles_x * a.tiles_y;
        f = t / per_frame;
        t -= f * per_frame;
        const int ty = t / a.tiles_x;
        y0 = ty * kTileY;
        x0 = (t - ty * a.tiles_x) * kTileX;
    };

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(bar(BAR_W), kWBytes);
            for (int t = 0; t < 9; ++t) bulk_load_1d(base + kSmW + t * kWTapBytes, reinterpret_cast<const uint8_t*>(a.wprep) + t * kWTapBytes, kWTapBytes, bar(BAR_W));
            for (int i = 0; i < n_mine; ++i) {
                const int s = i & 1, u = i >> 1;
                int f, y0, x0;
                tile_coord(i, f, y0, x0);
                mbar_wait(bar(BAR_STAGE_EMPTY + s), (u & 1) ^ 1);
                mbar_expect_tx(bar(BAR_STAGE_FULL + s), kStageBytes);
                tma_load_4d(base + kSmStage + s * kStageBytes, &map_in, bar(BAR_STAGE_FULL + s), 0, x0 - 1, y0 - 1, f);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(bar(BAR_W), 0);
            constexpr uint32_t idesc64 = instr_desc(64), idesc32 = instr_desc(32);
            for (int i = 0; i < n_mine; ++i) {
                const int s = i & 1, u = i >> 1;
                mbar_wait(bar(BAR_OPND_FULL + s), u & 1);
                mbar_wait(bar(BAR_ACC_EMPTY + s), (u & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(s * 64);
                const uint32_t a_hi = base + kSmOpnd + s * kOpndBytes, a_lo = a_hi + kPlaneBytes;
                const uint32_t w0 = base + kSmW;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap % 3;
#pragma unroll
                    for (int j = 0; j < kChunks / 2; ++j) {
                        const uint32_t aoff = (uint32_t)((dy * kHaloX + dx) * 16 + 2 * j * kPlanePitch);
                        const uint64_t bd = smem_desc(w0 + tap * kWTapBytes + 2 * j * kWChunkBytes, kWChunkBytes, 128);
                        // A_hi x [B_hi | B_lo] -> columns 0..63 (the tile's first MMA initialises all of them)
                        umma_tf32(d, smem_desc(a_hi + aoff, kPlanePitch, kHaloX * 16), bd, idesc64, (tap | j) != 0);
                        // A_lo x B_hi -> columns 0..31
                        umma_tf32(d, smem_desc(a_lo + aoff, kPlanePitch, kHaloX * 16), bd, idesc32, 1u);
                    }
                }
                umma_commit(bar(BAR_OPND_EMPTY + s));
                umma_commit(bar(BAR_ACC_FULL + s));
            }
        }
    } else {
        // ===== workers: conversion + epilogue =====
        auto convert = [&](int i) {
            const int s = i & 1, u = i >> 1;
            mbar_wait(bar(BAR_STAGE_FULL + s), u & 1);
            mbar_wait(bar(BAR_OPND_EMPTY + s), (u & 1) ^ 1);
            const uint8_t* st = gbase + kSmStage + s * kStageBytes;
            uint8_t* op = gbase + kSmOpnd + s * kOpndBytes;
#pragma unroll 4
            for (int item = tid; item < kHaloPix * kChunks; item += kWorkers) {
                const int px = item >> 3, ck = item & 7;
                const float4 v = *reinterpret_cast<const float4*>(st + px * (kC * 4) + ck * 16);
                float4 hi, lo;
                hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
                lo.x = to_tf32(v.x - hi.x); lo.y = to_tf32(v.y - hi.y); lo.z = to_tf32(v.z - hi.z); lo.w = to_tf32(v.w - hi.w);
                *reinterpret_cast<float4*>(op + ck * kPlanePitch + px * 16) = hi;
                *reinterpret_cast<float4*>(op + kPlaneBytes + ck * kPlanePitch + px * 16) = lo;
            }
            fence_proxy_async();       // the tensor core reads shared memory through the async proxy
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(BAR_OPND_FULL + s));
                mbar_arrive(bar(BAR_STAGE_EMPTY + s));
            }
        };
        auto epilogue = [&](int i) {
            const int s = i & 1, u = i >> 1;
            int f, y0, x0;
            tile_coord(i, f, y0, x0);
            const int m = tid, gy = y0 + (m >> 3), gx = x0 + (m & 7);
            bool raw_out = false;
            if (a.active != nullptr && gy < a.H && gx < a.W) {
                const long long cell = ((long long)f * a.H + gy) * a.W + gx;
                raw_out = (__ldg(a.active + (cell >> 5)) >> (cell & 31)) & 1u;
            }
            mbar_wait(bar(BAR_ACC_FULL + s), u & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 64);
            SHPL_TMEM_LD32(taddr, r0);
            SHPL_TMEM_LD32(taddr + 32, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY + s));
            float v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float x = __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                if (!raw_out) {
                    x = fmaf(x, sc[c], sc[32 + c]);
                    if (a.relu) x = fmaxf(x, 0.f);
                }
                v[c] = x;
            }
            if (tid == 0) bulk_wait_read0();       // the previous tile's store has read the staging buffer
            asm volatile("bar.sync 1, 128;" ::: "memory");
            uint8_t* orow = gbase + kSmOut + m * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(orow + ((j ^ (m & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) {
                tma_store_4d(&map_out, base + kSmOut, 0, x0, y0, f);
                bulk_commit();
            }
        };
        if (n_mine > 0) convert(0);
        for (int i = 0; i < n_mine; ++i) {
            if (i + 1 < n_mine) convert(i + 1);
            epilogue(i);
        }
        if (tid == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------ weight prep
// W is the slim.conv2d variable, HWIO [3][3][c_in_total][c_out_total].  forward: B[n = co][k = ci] = W[tap][ci_off + ci][co];
// transposed (input gradient, SAME padding, stride 1): B[n = ci][k = co] = W[8 - tap][ci_off + ci][co].
__global__ void shpl_conv_prep_kernel(const float* __restrict__ w, int c_in_total, int c_out_total, int ci_off, int transposed,
                                      float* __restrict__ wprep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // over [9][8][64][4]
    if (i >= 9 * kChunks * 64 * 4) return;
    const int e = i & 3, n = (i >> 2) & 63, ck = (i >> 8) & 7, tap = i >> 11;
    const int k = ck * 4 + e, nn = n & 31;
    float x;
    if (!transposed) x = w[((size_t)tap * c_in_total + ci_off + k) * c_out_total + nn];
    else x = w[((size_t)(8 - tap) * c_in_total + ci_off + nn) * c_out_total + k];
    const float hi = to_tf32(x);
    wprep[i] = n < 32 ? hi : to_tf32(x - hi);
}

// ------------------------------------------------------------------------------------ sparse half
// Marks the cells that receive pooled features (busy) and their 3x3 neighbourhoods (active) from the key-sorted
// entry list; one thread per entry, the first entry of a cell does the work.
__global__ void shpl_conv_mark_kernel(const int* __restrict__ ptr, const int* __restrict__ key, int n_rows, int nnz_max,
                                      int H, int W, uint32_t* __restrict__ busy, uint32_t* __restrict__ active) {
    const int e_begin = __ldg(ptr), e_end = __ldg(ptr + n_rows);
    const int e = e_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_end || e - e_begin >= nnz_max) return;
    const int r = __ldg(key + e);
    if (e > e_begin && __ldg(key + e - 1) == r) return;
    atomicOr(busy + (r >> 5), 1u << (r & 31));
    const int f = r / (H * W), rem = r - f * (H * W), y = rem / W, x = rem - y * W;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int c = (f * H + yy) * W + xx;
            atomicOr(active + (c >> 5), 1u << (c & 31));
        }
}

struct SparseConvArgs {
    const float* src;            // [n_src, C_s] gathered map
    const int* ptr;              // CSR by destination cell
    const int* idx;
    const float* val;
    const float* w;              // HWIO weights; the pooled channels start at ci_off
    int c_in_total, c_out_total, ci_off, C_s;
    const float* scale;
    const float* shift;
    int relu;
    const uint32_t* busy;
    const uint32_t* active;
    int frames, H, W, n_words;
    float* out;                  // [cells, 32]: holds the dense term (un-activated) for the marked cells
};

// One warp per 32-cell word of the `active` bitmap (lane = output channel while a cell is processed):
//   out = act(scale * (dense + sum_taps W_pooled[tap]^T . pooled[nbr]) + shift).
// pooled[nbr] is formed like the pooling kernels form it: entries in stored order, separately rounded multiply and add.
// The dependent loads of a cell (busy bits -> CSR offsets -> first entry -> gathered row) are issued side by side for
// all nine taps by lanes 0..8, so a cell costs three round trips whatever the number of busy neighbours.
constexpr int kSparseThreads = 512;

__global__ void __launch_bounds__(kSparseThreads) shpl_conv_sparse_kernel(SparseConvArgs a) {
    extern __shared__ float wsm[];           // [9][C_s][32]
    const int n_w = 9 * a.C_s * 32;
    for (int i = threadIdx.x; i < n_w; i += blockDim.x) {
        const int co = i & 31, ci = (i >> 5) % a.C_s, tap = (i >> 5) / a.C_s;
        wsm[i] = a.w[((size_t)tap * a.c_in_total + a.ci_off + ci) * a.c_out_total + co];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const float s_c = a.scale ? a.scale[lane] : 1.f, b_c = a.shift ? a.shift[lane] : 0.f;
    const int HW = a.H * a.W;
    const int t_dy = lane / 3 - 1, t_dx = lane % 3 - 1;          // lanes 0..8: the tap this lane looks after
    for (int word = blockIdx.x * warps + warp; word < a.n_words; word += gridDim.x * warps) {
        uint32_t bits = __ldg(a.active + word);
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int c = word * 32 + b;
            const int f = c / HW, rem = c - f * HW, y = rem / a.W, x = rem - y * a.W;
            const float o_in = a.out[(size_t)c * 32 + lane];
            // lanes 0..8: is the neighbour of tap `lane` busy, and where are its entries
            int beg = 0, end = 0, p0 = 0;
            float w0 = 0.f;
            if (lane < 9) {
                const int yy = y + t_dy, xx = x + t_dx;
                if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
                    const int nb = (f * a.H + yy) * a.W + xx;
                    if ((__ldg(a.busy + (nb >> 5)) >> (nb & 31)) & 1u) {
                        beg = __ldg(a.ptr + nb);
                        end = __ldg(a.ptr + nb + 1);
                        p0 = __ldg(a.idx + beg);
                        w0 = __ldg(a.val + beg);
                    }
                }
            }
            const uint32_t taps = __ballot_sync(0xffffffffu, end > beg) & 0x1ffu;
            float acc = 0.f;
            for (int c0 = 0; c0 < a.C_s; c0 += 32) {
                float xrow[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {          // the first entry of every busy tap: all gathers in flight together
                    const int p = __shfl_sync(0xffffffffu, p0, t);
                    xrow[t] = ((taps >> t) & 1u) ? __ldg(a.src + (size_t)p * a.C_s + c0 + lane) : 0.f;
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (!((taps >> t) & 1u)) continue;   // warp-uniform
                    const int tb = __shfl_sync(0xffffffffu, beg, t), te = __shfl_sync(0xffffffffu, end, t);
                    float p = __fadd_rn(0.f, __fmul_rn(__shfl_sync(0xffffffffu, w0, t), xrow[t]));
                    for (int k = tb + 1; k < te; ++k)
                        p = __fadd_rn(p, __fmul_rn(__ldg(a.val + k), __ldg(a.src + (size_t)__ldg(a.idx + k) * a.C_s + c0 + lane)));
                    const float* wt = wsm + ((size_t)t * a.C_s + c0) * 32 + lane;
#pragma unroll 8
                    for (int ci = 0; ci < 32; ++ci) acc = fmaf(__shfl_sync(0xffffffffu, p, ci), wt[ci * 32], acc);
                }
            }
            float o = fmaf(o_in + acc, s_c, b_c);
            if (a.relu) o = fmaxf(o, 0.f);
            a.out[(size_t)c * 32 + lane] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;         // resolved once per process: a pure function pointer, no device state
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// NHWC fp32 [frames][H][W][32 channels out of `pitch` floats per pixel], box [1][box_y][box_x][32]
int make_map(CUtensorMap* m, const float* base, int frames, int H, int W, int pitch, int box_y, int box_x, bool swizzle128) {
    EncodeTiledFn fn = encode_fn();
    SHPL_REQUIRE(fn != nullptr, SHPL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)W * pitch * 4, (cuuint64_t)H * W * pitch * 4};
    const cuuint32_t box[4] = {(cuuint32_t)kC, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SHPL_REQUIRE(r == CUDA_SUCCESS, SHPL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SHPL_OK;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ConvWorkspace {
    float* wprep;
    uint32_t* busy;
    uint32_t* active;
    size_t words;
    size_t bytes;
};

ConvWorkspace carve(void* ws, long long cells) {
    ConvWorkspace c;
    c.words = (size_t)((cells + 31) / 32);
    const size_t bm = align_up(c.words * 4, 256);
    uint8_t* p = static_cast<uint8_t*>(ws);
    c.wprep = reinterpret_cast<float*>(p);
    c.busy = reinterpret_cast<uint32_t*>(p + kWBytes);
    c.active = reinterpret_cast<uint32_t*>(p + kWBytes + bm);
    c.bytes = kWBytes + 2 * bm;
    return c;
}

int launch_dense(const float* in, int in_pitch, float* out, const float* wprep, const float* scale, const float* shift, int relu,
                 const uint32_t* active, int frames, int H, int W, cudaStream_t s) {
    CUtensorMap map_in, map_out;
    if (int rc = make_map(&map_in, in, frames, H, W, in_pitch, kHaloY, kHaloX, false)) return rc;
    if (int rc = make_map(&map_out, out, frames, H, W, kC, kTileY, kTileX, true)) return rc;
    ConvArgs a{};
    a.wprep = wprep;
    a.scale = scale;
    a.shift = shift;
    a.active = active;
    a.relu = relu;
    a.frames = frames;
    a.H = H;
    a.W = W;
    a.tiles_x = (W + kTileX - 1) / kTileX;
    a.tiles_y = (H + kTileY - 1) / kTileY;
    a.n_tiles = frames * a.tiles_x * a.tiles_y;
    static bool attr_set = false;      // idempotent function attribute, not per-call state
    if (!attr_set) {
        SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_conv3x3_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
        attr_set = true;
    }
    const int grid = a.n_tiles < shpl::sm_count() ? a.n_tiles : shpl::sm_count();
    shpl_conv3x3_dense_kernel<<<grid, kConvThreads, kConvSmem, s>>>(map_in, map_out, a);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_conv3x3_dense_kernel");
}

}  // namespace

extern "C" size_t shpl_conv3x3_workspace_bytes(int32_t frames, int32_t H, int32_t W) {
    if (frames <= 0 || H <= 0 || W <= 0) return 0;
    return carve(nullptr, (long long)frames * H * W).bytes;
}

extern "C" int shpl_pool_conv3x3_forward(const float* dst, const float* src, const int32_t* ptr, const int32_t* key,
                                         const int32_t* idx, const float* val, int32_t nnz_max, int32_t frames, int32_t H,
                                         int32_t W, int32_t C_d, int32_t n_src, int32_t C_s, const float* weight, int32_t C_out,
                                         const float* scale, const float* shift, int32_t relu, float* out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    SHPL_REQUIRE(frames > 0 && H > 0 && W > 0 && n_src >= 0 && C_s >= 0 && nnz_max >= 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_forward: bad sizes frames=%d H=%d W=%d n_src=%d C_s=%d", frames, H, W, n_src, C_s);
    SHPL_REQUIRE(dst && weight && out && workspace, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_conv3x3_forward: null pointer");
    SHPL_REQUIRE(C_s == 0 || (src && ptr && key && idx && val), SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_forward: pooled channels need src and the CSR arrays (with the key array)");
    SHPL_REQUIRE(C_d == kC && C_out == kC && C_s % 32 == 0 && C_s <= 64, SHPL_ERR_UNSUPPORTED,
                 "shpl_pool_conv3x3_forward: built for C_d = C_out = 32 and C_s in {0, 32, 64} (got %d, %d, %d)", C_d, C_out, C_s);
    SHPL_REQUIRE((long long)frames * H * W < (1ll << 31), SHPL_ERR_UNSUPPORTED, "shpl_pool_conv3x3_forward: map too large");
    SHPL_REQUIRE(shpl::aligned(dst, 16) && shpl::aligned(out, 16) && shpl::aligned(workspace, 256), SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_forward: dst / out must be 16-byte aligned, workspace 256-byte aligned");
    const long long cells = (long long)frames * H * W;
    const ConvWorkspace c = carve(workspace, cells);
    SHPL_REQUIRE(workspace_bytes >= c.bytes, SHPL_ERR_WORKSPACE_TOO_SMALL, "shpl_pool_conv3x3_forward: workspace %zu < %zu bytes",
                 workspace_bytes, c.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int c_in_total = C_d + C_s;
    shpl_conv_prep_kernel<<<(9 * kChunks * 64 * 4 + 255) / 256, 256, 0, s>>>(weight, c_in_total, C_out, 0, 0, c.wprep);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_conv_prep_kernel")) return rc;
    const bool sparse = C_s > 0 && nnz_max > 0;
    if (sparse) {
        SHPL_CUDA_OK(cudaMemsetAsync(c.busy, 0, (size_t)(reinterpret_cast<uint8_t*>(c.active) - reinterpret_cast<uint8_t*>(c.busy)) + c.words * 4, s));
        shpl_conv_mark_kernel<<<(nnz_max + 255) / 256, 256, 0, s>>>(ptr, key, (int)cells, nnz_max, H, W, c.busy, c.active);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_mark_kernel")) return rc;
    }
    if (int rc = launch_dense(dst, C_d, out, c.wprep, scale, shift, relu, sparse ? c.active : nullptr, frames, H, W, s)) return rc;
    if (sparse) {
        SparseConvArgs sa{};
        sa.src = src;
        sa.ptr = ptr;
        sa.idx = idx;
        sa.val = val;
        sa.w = weight;
        sa.c_in_total = c_in_total;
        sa.c_out_total = C_out;
        sa.ci_off = C_d;
        sa.C_s = C_s;
        sa.scale = scale;
        sa.shift = shift;
        sa.relu = relu;
        sa.busy = c.busy;
        sa.active = c.active;
        sa.frames = frames;
        sa.H = H;
        sa.W = W;
        sa.n_words = (int)c.words;
        sa.out = out;
        const size_t smem = (size_t)9 * C_s * 32 * sizeof(float);
        static bool attr_set = false;
        if (smem > 48 * 1024 && !attr_set) {
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_conv_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 64 * 32 * 4));
            attr_set = true;
        }
        constexpr int kW = kSparseThreads / 32;
        int grid = ((int)c.words + kW - 1) / kW;
        const int cap = shpl::sm_count() * 4;
        if (grid > cap) grid = cap;
        shpl_conv_sparse_kernel<<<grid, kSparseThreads, smem, s>>>(sa);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_sparse_kernel")) return rc;
    }
    return SHPL_OK;
}
