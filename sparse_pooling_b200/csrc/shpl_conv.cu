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
// cells that receive pooled features (~2 % at KITTI stride 1): a second tcgen05 kernel (shpl_conv_z_tc_kernel) forms,
// for every such cell, Z[cell][tap] = W_pooled[tap]^T . pooled[cell] (the pooled sums in the reference's entry order,
// like the pooling kernels do) and sets the cell's bit in a bitmap; four warps of the dense kernel then add, for every
// output pixel, the Z rows of its busy neighbours into the tile's output staging buffer one tile ahead of the epilogue.
//
// Dense kernel, one persistent CTA per SM, 14 warps, every stage double-buffered so that they all overlap:
//   warp 8      TMA producer   halo tile [10 x 16 pixels x 32 ch] fp32 -> staging
//   warps 4-7   conversion     staging -> hi / lo operand planes ([chunk of 4 ch][halo pixel][16 B]: the K-major,
//                              no-swizzle core-matrix layout, so a row shift dy is just a different start address)
//   warps 10-13 pooled half    thread = output pixel: busy bits of its nine neighbours -> their CSR offsets -> their Z rows,
//                              summed in registers in tap order, stored into the tile's output staging buffer
//   warps 0-3   epilogue       TMEM -> registers (the dx shuffle-sum) + the pooled half from the staging buffer, scale / shift /
//                              ReLU -> back into the (128B-swizzled) staging buffer -> TMA store
//   warp 9    MMA issuer     per tile 3 dy x 4 K-steps x {A_hi x [B_hi | B_lo] (N = 192), A_lo x B_hi (N = 96)}
// The 128 rows of an MMA are 8 image rows x 16 HALO columns; the N dimension carries the three dx taps side by side
// (and hi | lo of the weights), so one read of A serves three taps: the tensor core fetches its shared-memory operands
// at only ~64-75 B/clk (measured, profiles/r2_conv_*), which -- not the math -- bounds this small-N problem.  The
// epilogue adds the three dx partial sums of neighbouring halo columns (adjacent TMEM lanes = adjacent threads:
// two shuffles per channel); the two halo columns of a tile produce no output (tile = 8 x 14 output pixels).
#include <cuda.h>

#include "shpl_common.cuh"

namespace {

constexpr int kC = 32;                     // dense input channels = output channels of this kernel
constexpr int kTileY = 8, kTileX = 14;     // output pixels per tile
constexpr int kHaloY = kTileY + 2, kHaloX = kTileX + 2, kHaloPix = kHaloY * kHaloX;   // 10 x 16 = 160
static_assert(kHaloX == 16 && kTileY * kHaloX == 128, "M = 128 rows = 8 image rows x 16 halo columns");
constexpr int kChunks = kC / 4;            // 16-byte channel chunks per pixel
constexpr int kPlanePitch = (kHaloPix + 1) * 16;     // bytes between channel chunks (+16: bank spread)
constexpr int kPlaneBytes = kChunks * kPlanePitch;   // one hi or lo plane set
constexpr int kOpndBytes = 2 * kPlaneBytes;          // hi + lo
constexpr int kStageBytes = kHaloPix * kC * 4;       // 20480: one TMA box
constexpr int kNB = 192;                             // B rows per dy: [hi | lo][dx 0..2][co 32]
constexpr int kWChunkBytes = kNB * 16;               // [n = 192][4 ci]
constexpr int kWDyBytes = kChunks * kWChunkBytes;    // 24576
constexpr int kWBytes = 3 * kWDyBytes;               // 73728
constexpr int kWElems = kWBytes / 4;
constexpr int kOutRows = kTileY * kTileX;            // 112
constexpr int kOutBytes = kOutRows * 128;            // 14336: one output tile, 128B-swizzled rows
static_assert(kOutBytes % 1024 == 0, "output staging buffers stay 1024-byte aligned (swizzle atom)");
constexpr int kWorkers = 256;
constexpr int kSparseWarps = 4;                      // warps 10..13: the pooled half's contributions, one tile ahead
static_assert(kOutRows <= kSparseWarps * 32 && kHaloX == 16, "a sparse thread per output pixel of the tile");
constexpr int kConvThreads = kWorkers + 64 + kSparseWarps * 32;
constexpr int kAccCols = kNB;                        // columns of one accumulator buffer
constexpr int kTmemCols = 512;                       // 2 x 192 used

// dynamic shared memory map (byte offsets from a 1024-aligned base)
constexpr int kSmOut = 0;                                  // 2 x 14336, 1024-aligned (128B-swizzled TMA store source)
constexpr int kSmStage = kSmOut + 2 * kOutBytes;           // 2 x 20480
constexpr int kSmW = kSmStage + 2 * kStageBytes;           // 73728
constexpr int kSmOpnd = kSmW + kWBytes;                    // 2 x 41216
constexpr int kSmBar = kSmOpnd + 2 * kOpndBytes;           // mbarriers
constexpr int kSmScale = kSmBar + 256;                     // scale[32], shift[32]
constexpr int kSmEnd = kSmScale + 256;
constexpr int kConvSmem = kSmEnd + 1024;                   // + slack for the 1024-byte alignment of the base
static_assert(kConvSmem <= 227 * 1024, "dense conv kernel: shared memory budget");
static_assert(kSmStage % 128 == 0 && kSmW % 128 == 0 && kSmOpnd % 16 == 0 && kSmBar % 8 == 0, "smem alignment");

enum { BAR_STAGE_FULL = 0, BAR_STAGE_EMPTY = 2, BAR_OPND_FULL = 4, BAR_OPND_EMPTY = 6, BAR_ACC_FULL = 8, BAR_ACC_EMPTY = 10,
       BAR_W = 12, BAR_CONTRIB_FULL = 13, BAR_OUT_EMPTY = 15, BAR_COUNT = 17 };
static_assert(BAR_COUNT * 8 <= 256, "mbarrier region");

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

#define SHPL_TMEM_LD8(taddr, v)                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                  \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
                 : "r"(taddr))

#define SHPL_TMEM_LD16(taddr, v)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                        \
                 : "r"(taddr))

struct ConvArgs {
    const float* wprep;          // [3 dy][8 chunks][192 = (hi | lo) x dx x co][4 ci]: B-operand layout (shpl_conv_prep_kernel)
    const float* scale;          // [32] or NULL (1)
    const float* shift;          // [32] or NULL (0)
    const uint32_t* busy;        // bitmap of the cells that receive pooled features, or NULL (no pooled half)
    const int* ptr;              // CSR offsets by cell: ptr[c] - ptr[0] = the Z row of a busy cell
    const float* Z;              // [entries][9 taps][32]: W_pooled[tap]^T . pooled[cell], from the Z kernel
    int z_rows;                  // entries Z holds (debug build: every gathered Z row is checked against it)
    int relu;
    int frames, H, W;
    int tiles_x, tiles_y, n_tiles;
};

// ------------------------------------------------------------------------------------ dense kernel
__global__ void __launch_bounds__(kConvThreads, 1)
shpl_conv3x3_dense_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, ConvArgs a) {
    extern __shared__ uint8_t conv_smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    __shared__ uint32_t contrib_mask[2][4];      // per staging buffer and sparse warp: the staging rows that hold a pooled contribution
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
            mbar_init(bar(BAR_STAGE_EMPTY + s), 4);      // the four conversion warps
            mbar_init(bar(BAR_OPND_FULL + s), 4);
            mbar_init(bar(BAR_OPND_EMPTY + s), 1);
            mbar_init(bar(BAR_ACC_FULL + s), 1);
            mbar_init(bar(BAR_ACC_EMPTY + s), 4);        // the four epilogue warps
            mbar_init(bar(BAR_CONTRIB_FULL + s), kSparseWarps);
            mbar_init(bar(BAR_OUT_EMPTY + s), 1);        // the thread that commits the tile stores
        }
        mbar_init(bar(BAR_W), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) sc[tid] = tid < 32 ? (a.scale ? a.scale[tid] : 1.f) : (a.shift ? a.shift[tid - 32] : 0.f);
    if (warp == 9) {
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

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(bar(BAR_W), kWBytes);
            for (int t = 0; t < 3; ++t) bulk_load_1d(base + kSmW + t * kWDyBytes, reinterpret_cast<const uint8_t*>(a.wprep) + t * kWDyBytes, kWDyBytes, bar(BAR_W));
            for (int i = 0; i < n_mine; ++i) {
                const int s = i & 1, u = i >> 1;
                int f, y0, x0;
                tile_coord(i, f, y0, x0);
                mbar_wait(bar(BAR_STAGE_EMPTY + s), (u & 1) ^ 1);
                mbar_expect_tx(bar(BAR_STAGE_FULL + s), kStageBytes);
                tma_load_4d(base + kSmStage + s * kStageBytes, &map_in, bar(BAR_STAGE_FULL + s), 0, x0 - 1, y0 - 1, f);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(bar(BAR_W), 0);
            constexpr uint32_t idesc_hi = instr_desc(kNB), idesc_lo = instr_desc(kNB / 2);
            for (int i = 0; i < n_mine; ++i) {
                const int s = i & 1, u = i >> 1;
                mbar_wait(bar(BAR_OPND_FULL + s), u & 1);
                mbar_wait(bar(BAR_ACC_EMPTY + s), (u & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(s * kAccCols);
                const uint32_t a_hi = base + kSmOpnd + s * kOpndBytes, a_lo = a_hi + kPlaneBytes;
                const uint32_t w0 = base + kSmW;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int j = 0; j < kChunks / 2; ++j) {
                        const uint32_t aoff = (uint32_t)(dy * kHaloX * 16 + 2 * j * kPlanePitch);
                        const uint64_t bd = smem_desc(w0 + dy * kWDyBytes + 2 * j * kWChunkBytes, kWChunkBytes, 128);
                        // A_hi x [B_hi | B_lo], three dx taps side by side -> all 192 columns (the tile's first MMA initialises them)
                        umma_tf32(d, smem_desc(a_hi + aoff, kPlanePitch, 128), bd, idesc_hi, (dy | j) != 0);
                        // A_lo x B_hi -> columns 0..95
                        umma_tf32(d, smem_desc(a_lo + aoff, kPlanePitch, 128), bd, idesc_lo, 1u);
                    }
                }
                umma_commit(bar(BAR_OPND_EMPTY + s));
                umma_commit(bar(BAR_ACC_FULL + s));
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== conversion warps (4..7): staging -> hi / lo operand planes =====
        const int ctid = tid - 128;
        for (int i = 0; i < n_mine; ++i) {
            const int s = i & 1, u = i >> 1;
            mbar_wait(bar(BAR_STAGE_FULL + s), u & 1);
            mbar_wait(bar(BAR_OPND_EMPTY + s), (u & 1) ^ 1);
            const uint8_t* st = gbase + kSmStage + s * kStageBytes;
            uint8_t* op = gbase + kSmOpnd + s * kOpndBytes;
            static_assert(kHaloPix * kChunks == 10 * 128, "ten items per conversion thread");
#pragma unroll 5
            for (int it = 0; it < 10; ++it) {
                const int item = ctid + it * 128;
                const int px = item >> 3, ck = item & 7;
                const float4 v = *reinterpret_cast<const float4*>(st + px * (kC * 4) + ck * 16);
                // hi = the tf32 the tensor core would read anyway (top 19 bits); lo = the exact remainder
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                *reinterpret_cast<float4*>(op + ck * kPlanePitch + px * 16) = hi;
                *reinterpret_cast<float4*>(op + kPlaneBytes + ck * kPlanePitch + px * 16) = lo;
            }
            fence_proxy_async();       // the tensor core reads shared memory through the async proxy
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(BAR_OPND_FULL + s));
                mbar_arrive(bar(BAR_STAGE_EMPTY + s));
            }
        }
    } else if (warp >= 10) {
        // ===== sparse warps (10..13): the pooled half of the conv, one tile ahead of the epilogue =====
        // Thread = output pixel of the tile (112 of the 128 threads): the sum, in tap order (a fixed order: deterministic), of
        // the Z rows of its busy neighbours -- Z[cell][tap] = W_pooled[tap]^T . pooled[cell], from the Z kernel -- formed in
        // registers and left in the tile's output staging buffer, where the epilogue adds the dense half on top.  The cost
        // of a tile is bounded (as many rounds as the busiest pixel of a warp has busy neighbours, at most nine) however the
        // busy cells cluster, and only the final store needs the staging buffer.  A warp works on its own (no barrier between
        // the four): lanes 0..4 hold the busy bits of the five halo rows its pixels look at.  No global latency but the Z
        // loads (which hit L1) is exposed per tile:
        //   tile i+2: lanes 0..4 load their halo row's busy-bitmap words
        //   tile i+1: bits -> every thread tests its nine neighbours and issues the (few) CSR-offset loads, predicated
        //   end of tile i: offsets -> Z rows of tile i+1, each (one 128-byte line) prefetched into L1
        if (a.busy != nullptr && n_mine > 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");           // the Z kernel's rows and bitmap (no-op without PDL)
            const int sw = warp - 10;
            const int r = sw * 32 + lane;                                // staging row = output pixel of the tile
            const bool px_ok = r < kOutRows;
            const int yl = px_ok ? r / kTileX : 0, xl = px_ok ? r - yl * kTileX : 0;
            const int yb = (sw * 32) / kTileX;                           // first output row of this warp = first halo row it looks at
            const int e_begin = __ldg(a.ptr);
            uint32_t hw_lo = 0u, hw_hi = 0u;
            int hw_sh = 0, hw_nv = 0, hw_ls = 0;
            int hf = 0, hy0 = 0, hx0 = 0;                                // coordinates of the tile whose halo words are in flight
            auto issue_halo = [&](int j) {                               // lanes 0..4: raw bitmap words of halo row yb + lane of tile j
                hw_nv = 0;
                if (j >= n_mine) return;
                tile_coord(j, hf, hy0, hx0);                             // one coordinate computation per tile: issue_offsets reuses it
                if (lane >= 5 || yb + lane >= kHaloY) return;
                const int yy = hy0 - 1 + yb + lane;
                if (yy < 0 || yy >= a.H) return;
                const int xs = hx0 > 0 ? hx0 - 1 : 0;                    // first in-image halo column
                const long long c = ((long long)hf * a.H + yy) * a.W + xs;
                const long long last = ((long long)a.frames * a.H * a.W - 1) >> 5;
                const long long wi = c >> 5;
                hw_lo = __ldg(a.busy + wi);
                hw_hi = wi + 1 <= last ? __ldg(a.busy + wi + 1) : 0u;
                hw_sh = (int)(c & 31);
                hw_ls = xs - (hx0 - 1);
                hw_nv = min(a.W - xs, kHaloX - hw_ls);                   // columns of this row inside the image
            };
            int poff[9];                                                 // CSR offsets of the busy neighbours (loads in flight)
            uint32_t pnear = 0u;                                         // bit t: the neighbour at tap t is busy
            auto issue_offsets = [&]() {                                 // for the tile of the last issue_halo
                uint32_t bits = 0u;
                if (hw_nv > 0) {
                    bits = (uint32_t)(((((uint64_t)hw_hi) << 32) | hw_lo) >> hw_sh) & 0xffffu;
                    bits &= (1u << hw_nv) - 1u;
                    bits <<= hw_ls;
                }
                const int gy = hy0 + yl, gx = hx0 + xl;
                pnear = 0u;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
                    pnear |= ((__shfl_sync(0xffffffffu, bits, yl + dy - yb) >> xl) & 7u) << (3 * dy);
                if (!(px_ok && gy < a.H && gx < a.W)) pnear = 0u;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int nb = (hf * a.H + gy + t / 3 - 1) * a.W + gx + t % 3 - 1;
                    const bool on = (pnear >> t) & 1u;
                    SHPL_DASSERT(!on || (nb >= 0 && nb < a.frames * a.H * a.W));
                    poff[t] = on ? __ldg(a.ptr + nb) : 0;
                }
            };
            int zrow[9];
            uint32_t zmask = 0u;
            auto rows_of_offsets = [&]() {                               // offsets (arrived) -> Z rows + their L1 prefetch
                zmask = pnear;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    zrow[t] = ((zmask >> t) & 1u) ? (poff[t] - e_begin) * 9 + t : -1;
                    SHPL_DASSERT(zrow[t] < 0 || zrow[t] / 9 < a.z_rows);
                    if (zrow[t] >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.Z + (size_t)zrow[t] * 32));
                }
            };
            issue_halo(0);
            issue_offsets();
            issue_halo(1);
            rows_of_offsets();
            for (int i = 0; i < n_mine; ++i) {
                const int ob = i & 1, u = i >> 1;
                if (i + 1 < n_mine) {
                    issue_offsets();                                     // tile i+1 (its halo words were loaded during the last tile)
                    issue_halo(i + 2);
                }
                float o[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = 0.f;
                // Only the pixels with a busy neighbour (about one in five) write their staging row; the epilogue reads the
                // rows flagged in contrib_mask and takes the others as zero.
                const uint32_t has_rows = __ballot_sync(0xffffffffu, px_ok && zmask != 0u);
                const bool has = px_ok && zmask != 0u;
                // Round k handles the k-th busy tap of every lane together: a warp pays one latency per round, and there are as
                // many rounds as its busiest pixel has busy neighbours.
                while (__any_sync(0xffffffffu, zmask != 0u)) {
                    if (zmask != 0u) {
                        const int t = __ffs(zmask) - 1;
                        zmask &= zmask - 1u;
                        int zr = zrow[0];
#pragma unroll
                        for (int tt = 1; tt < 9; ++tt) zr = (t == tt) ? zrow[tt] : zr;
                        const float* z = a.Z + (size_t)zr * 32;
                        float zv[32];
#pragma unroll
                        for (int j = 0; j < 4; ++j)      // 256-bit loads: a 128-byte row in four requests
                            asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=f"(zv[8 * j]), "=f"(zv[8 * j + 1]), "=f"(zv[8 * j + 2]), "=f"(zv[8 * j + 3]),
                                           "=f"(zv[8 * j + 4]), "=f"(zv[8 * j + 5]), "=f"(zv[8 * j + 6]), "=f"(zv[8 * j + 7])
                                         : "l"(z + 8 * j));
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] += zv[c];
                    }
                }
                mbar_wait(bar(BAR_OUT_EMPTY + ob), (u & 1) ^ 1);         // the store of tile i-2 has read this buffer
                if (lane == 0) contrib_mask[ob][sw] = has_rows;
                if (has) {
                    uint8_t* orow = gbase + kSmOut + ob * kOutBytes + r * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(orow + ((j ^ (r & 7)) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_CONTRIB_FULL + ob));
                if (i + 1 < n_mine) rows_of_offsets();                   // tile i+1's rows: in L1 by the time they are read
            }
        }
    } else {
        // ===== epilogue warps (0..3): TMEM -> registers (+ the pooled half from the staging buffer) -> swizzled smem -> TMA store =====
        const int q = warp;                                  // TMEM lane quadrant
        const int m = q * 32 + lane, yl = m >> 4, xq = m & 15;
        const bool col_ok = xq >= 1 && xq <= kTileX;
        const bool pooled = a.busy != nullptr;
        for (int i = 0; i < n_mine; ++i) {
            const int s = i & 1, u = i >> 1;
            int f, y0, x0;
            tile_coord(i, f, y0, x0);
            mbar_wait(bar(BAR_ACC_FULL + s), u & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols);
            float o[32];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int c0 = g * 8;
                uint32_t h[3][8], l[3][8];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    SHPL_TMEM_LD8(taddr + dx * 32 + c0, h[dx]);
                    SHPL_TMEM_LD8(taddr + 96 + dx * 32 + c0, l[dx]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float t0 = __uint_as_float(h[0][c]) + __uint_as_float(l[0][c]);
                    const float t1 = __uint_as_float(h[1][c]) + __uint_as_float(l[1][c]);
                    const float t2 = __uint_as_float(h[2][c]) + __uint_as_float(l[2][c]);
                    // output at halo column xq = tap dx=0 of column xq-1 + dx=1 of xq + dx=2 of xq+1
                    o[c0 + c] = __shfl_up_sync(0xffffffffu, t0, 1) + t1 + __shfl_down_sync(0xffffffffu, t2, 1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY + s));
            if (tid == 0 && i > 0) {
                // the store of tile i-1 (committed at the end of the last iteration) has read its staging buffer by now: free
                // for tile i+1, which the sparse warps fill during the rest of this tile
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(bar(BAR_OUT_EMPTY + (s ^ 1)));
            }
            // the staging buffer of this tile: filled with the pooled half by the sparse warps, or just free again
            mbar_wait(bar((pooled ? BAR_CONTRIB_FULL : BAR_OUT_EMPTY) + s), pooled ? (u & 1) : ((u & 1) ^ 1));
            const int srow = yl * kTileX + xq - 1;           // staging row of this thread's pixel (col_ok)
            uint8_t* orow = gbase + kSmOut + s * kOutBytes + srow * 128;
            const int rsw = srow & 7;
            if (pooled && col_ok && ((contrib_mask[s][srow >> 5] >> (srow & 31)) & 1u)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 z = *reinterpret_cast<const float4*>(orow + ((j ^ rsw) << 4));
                    o[4 * j] += z.x; o[4 * j + 1] += z.y; o[4 * j + 2] += z.z; o[4 * j + 3] += z.w;
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float x = fmaf(o[c], sc[c], sc[32 + c]);
                if (a.relu) x = fmaxf(x, 0.f);
                o[c] = x;
            }
            if (col_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(orow + ((j ^ rsw) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) {
                tma_store_4d(&map_out, base + kSmOut + s * kOutBytes, 0, x0, y0, f);
                bulk_commit();
            }
        }
        if (tid == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------ weight prep
// W is the slim.conv2d variable, HWIO [3][3][c_in_total][c_out_total].  forward: B[n = co][k = ci] = W[tap][ci_off + ci][co];
// transposed (input gradient, SAME padding, stride 1): B[n = ci][k = co] = W[8 - tap][ci_off + ci][co].
// gridDim.y = 2 preps a second block of 32 input channels (ci_off2 -> wprep2) in the same launch.
// zero[0 .. zero_words) is cleared by the same launch (the forward's bitmap of cells that receive pooled features).
__global__ void shpl_conv_prep_kernel(const float* __restrict__ w, int c_in_total, int c_out_total, int ci_off, int transposed,
                                      float* __restrict__ wprep, int ci_off2 = 0, float* __restrict__ wprep2 = nullptr,
                                      uint32_t* __restrict__ zero = nullptr, int zero_words = 0) {
    asm volatile("griddepcontrol.launch_dependents;");      // the Z kernel may start its gather under this launch
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // over [3 dy][8 chunks][192][4]
    for (int z = blockIdx.y * gridDim.x * blockDim.x + i; z < zero_words; z += gridDim.x * gridDim.y * blockDim.x) zero[z] = 0u;
    if (i >= kWElems) return;
    if (blockIdx.y == 1) {
        ci_off = ci_off2;
        wprep = wprep2;
    }
    const int e = i & 3, n = (i >> 2) % kNB, ck = (i / (kNB * 4)) % kChunks, dy = i / (kNB * 4 * kChunks);
    const int part = n / 96, dx = (n % 96) / 32, nn = n & 31;
    const int k = ck * 4 + e, tap = dy * 3 + dx;
    float x;
    if (!transposed) x = w[((size_t)tap * c_in_total + ci_off + k) * c_out_total + nn];
    else x = w[((size_t)(8 - tap) * c_in_total + ci_off + nn) * c_out_total + k];
    const float hi = to_tf32(x);
    wprep[i] = part == 0 ? hi : to_tf32(x - hi);
}

// ------------------------------------------------------------------------------------ sparse half
// Marks the cells that receive pooled features in the `busy` bitmap, from the key-sorted entry list; one thread per
// entry, the first entry of a cell sets the bit (atomicOr: order-independent, deterministic).
__global__ void shpl_conv_mark_kernel(const int* __restrict__ ptr, const int* __restrict__ key, int n_rows, int nnz_max,
                                      uint32_t* __restrict__ busy) {
    const int e_begin = __ldg(ptr), e_end = __ldg(ptr + n_rows);
    const int e = e_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_end || e - e_begin >= nnz_max) return;
    const int r = __ldg(key + e);
    if (e > e_begin && __ldg(key + e - 1) == r) return;
    atomicOr(busy + (r >> 5), 1u << (r & 31));
}

// ---- Z[e][tap][co] = sum_ci pooled[cell(e)][ci] * W[tap][C_d + ci][co] for every cell that receives pooled features,
// e = index of the cell's first CSR entry: what the cell contributes to each of its nine neighbours.  A small GEMM
// [cells x C_s] . [C_s x 288]; the dense kernel's epilogue gathers the rows.
struct ZArgs {
    const float* src;
    const int* ptr;
    const int* key;
    const int* idx;
    const float* val;
    const float* w;              // FFMA form: the HWIO weights
    const float* wprep;          // tensor-core form: the pooled channels' weights in the B-operand layout
    int c_in_total, c_out_total, ci_off, C_s;
    int n_rows, nnz_max;
    float* Z;                    // [nnz_max][9][32]
    uint32_t* busy;              // tensor-core form: the bitmap of the cells that receive pooled features is set here too
    int rows_per_tile;           // tensor-core form: entries a CTA takes per tile (multiple of 8, 8 .. 128)
};

// Tensor-core form (C_s = 32): a CTA takes rows_per_tile <= 128 consecutive entries per tile -- the host sizes the tiles so
// that the entries spread over two CTAs on every SM; the per-entry work (gather, Z row store) is what costs, the MMA is
// 128 rows wide whatever the tile holds.  Entry r of the tile is row m(r) = (r & 3) * 32 + (r >> 2) of the MMA (the four
// TMEM lane quarters, and with them the eight epilogue warps, share the rows evenly), holding the pooled vector of its
// cell if the entry is the cell's first (entries in stored order, two roundings each -- the pooling kernels' sum);
// other rows are never stored.  3xTF32 like the dense kernel: A_hi.B_hi + A_hi.B_lo + A_lo.B_hi into 288 TMEM columns
// = [tap][co], the layout of a Z row.  The epilogue transposes 4 x 4 blocks of float4 inside lane quads so that a
// store instruction writes 64 contiguous bytes per quad instead of 16 bytes per row.
#ifndef SHPL_CONV_PDL
#define SHPL_CONV_PDL 1
#endif
constexpr bool kConvPdl = SHPL_CONV_PDL != 0;
constexpr int kZtcThreads = 256;
constexpr int kZtcPitch = 128 * 16 + 16;                 // bytes between channel chunks of the A planes
constexpr int kZtcPlane = kChunks * kZtcPitch;
constexpr int kZtcSmA = 0;                               // hi, lo planes
constexpr int kZtcSmW = 2 * kZtcPlane;                   // 33024: 128-byte aligned (bulk-copy destination)
constexpr int kZtcSmBar = kZtcSmW + kWBytes;
constexpr int kZtcSmFlag = kZtcSmBar + 64;               // first-entry flags, one byte per row
constexpr int kZtcSmem = kZtcSmFlag + 128 + 128;         // + slack for the 128-byte alignment of the base
static_assert(kZtcSmW % 128 == 0 && kZtcSmBar % 8 == 0, "Z kernel smem alignment");
constexpr int kZtcSmemRequest = kZtcSmem;            // ~107 KB: two CTAs per SM
constexpr int kZtcTmemCols = 256;                    // 192 columns used per pass: two CTAs share the SM's 512

__global__ void __launch_bounds__(kZtcThreads, 2) shpl_conv_z_tc_kernel(ZArgs a) {
    extern __shared__ uint8_t z_smem_raw[];
    __shared__ uint32_t tmem_slot;
    const int e_begin = __ldg(a.ptr), e_end = min(__ldg(a.ptr + a.n_rows), e_begin + a.nnz_max);
    const int rpt = a.rows_per_tile, rpw = rpt >> 3;       // entries per tile, per warp (<= 16)
    const int n_tiles = (e_end - e_begin + rpt - 1) / rpt;
    if ((int)blockIdx.x >= n_tiles) return;               // whole CTA: nnz_max is only an upper bound
    const uint32_t raw = smem_u32(z_smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    uint8_t* gbase = z_smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t bar_w = base + kZtcSmBar, bar_mma = bar_w + 8;
    uint8_t* flags = gbase + kZtcSmFlag;
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kZtcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    int it = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++it) {
        const int e0 = e_begin + tile * rpt;
        // ---- gather: warp w fills the rows of entries ew .. ew + rpw - 1; lane = channel
        {
            const int r0 = warp * rpw, ew = e0 + r0;
            int k_l = ew + lane, key_l = -1, idx_l = 0, end_l = 0;
            float val_l = 0.f;
            bool first_l = false;
            if (lane < rpw && k_l < e_end) {
                key_l = __ldg(a.key + k_l);
                idx_l = __ldg(a.idx + k_l);
                val_l = __ldg(a.val + k_l);
                first_l = (k_l == e_begin) || (__ldg(a.key + k_l - 1) != key_l);
                if (first_l) end_l = __ldg(a.ptr + key_l + 1);
            }
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {                    // the gathered rows in flight together
                const int p = __shfl_sync(0xffffffffu, idx_l, j);
                SHPL_DASSERT(j >= rpw || ew + j >= e_end || p >= 0);
                x[j] = (j < rpw && ew + j < e_end) ? __ldg(a.src + (size_t)p * 32 + lane) : 0.f;
            }
            if (it == 0) {
                // Everything above reads the caller's arrays only.  The prep launch (this kernel is its programmatic
                // dependent) wrote the weights and cleared the bitmap: wait for it here, then let the dense kernel queue up.
                asm volatile("griddepcontrol.wait;" ::: "memory");
                asm volatile("griddepcontrol.launch_dependents;");
                if (tid == 0) {
                    mbar_expect_tx(bar_w, kWBytes);
                    for (int t = 0; t < 3; ++t)
                        bulk_load_1d(base + kZtcSmW + t * kWDyBytes, reinterpret_cast<const uint8_t*>(a.wprep) + t * kWDyBytes, kWDyBytes, bar_w);
                }
            }
            if (first_l) atomicOr(a.busy + (key_l >> 5), 1u << (key_l & 31));      // order-independent: deterministic
            const unsigned firsts = __ballot_sync(0xffffffffu, first_l) & 0xffffu;
            uint8_t* hi_p = gbase + kZtcSmA + (lane >> 2) * kZtcPitch + (lane & 3) * 4;
            uint8_t* lo_p = hi_p + kZtcPlane;
            auto put = [&](int j, float acc) {                // the pooled vector of the cell whose first entry is ew + j
                const int r = r0 + j, m = (r & 3) * 32 + (r >> 2);
                const float h = __uint_as_float(__float_as_uint(acc) & 0xffffe000u);
                *reinterpret_cast<float*>(hi_p + m * 16) = h;
                *reinterpret_cast<float*>(lo_p + m * 16) = acc - h;
            };
            float acc = 0.f;
            int open = -1, open_end = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j < rpw) {
                    if ((firsts >> j) & 1u) {
                        if (open >= 0) put(open, acc);            // the previous cell ended inside the window
                        open = j;
                        open_end = __shfl_sync(0xffffffffu, end_l, j);
                        acc = 0.f;
                    }
                    if (open >= 0 && ew + j < open_end)
                        acc = __fadd_rn(acc, __fmul_rn(__shfl_sync(0xffffffffu, val_l, j), x[j]));
                }
            }
            if (open >= 0) {
                for (int k = ew + rpw; k < open_end; ++k)     // the last cell runs on past the window
                    acc = __fadd_rn(acc, __fmul_rn(__ldg(a.val + k), __ldg(a.src + (size_t)__ldg(a.idx + k) * 32 + lane)));
                put(open, acc);
            }
            if (lane < rpw) {
                const int r = r0 + lane;
                flags[(r & 3) * 32 + (r >> 2)] = first_l ? 1 : 0;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                                      // A planes + flags complete; the previous tile's TMEM reads are done
        tc_fence_after();
        // Two passes over the taps (dy = 0, 1 -> 192 columns; dy = 2 -> 96 columns) so that 256 TMEM columns suffice and two
        // CTAs share an SM: every tile of a KITTI frame is then resident at once (no second wave).
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const int dy0 = pass * 2, ndy = 2 - pass;
            if (tid == 0) {
                if (it == 0 && pass == 0) mbar_wait(bar_w, 0);
                constexpr uint32_t idesc = instr_desc(96);
                const uint32_t a_hi = base + kZtcSmA, a_lo = a_hi + kZtcPlane, w0 = base + kZtcSmW;
                for (int d = 0; d < ndy; ++d) {
                    const int dy = dy0 + d;
#pragma unroll
                    for (int j = 0; j < kChunks / 2; ++j) {
                        const uint32_t aoff = (uint32_t)(2 * j * kZtcPitch);
                        const uint32_t wb = w0 + dy * kWDyBytes + 2 * j * kWChunkBytes;
                        const uint64_t b_hi = smem_desc(wb, kWChunkBytes, 128), b_lo = smem_desc(wb + 96 * 16, kWChunkBytes, 128);
                        const uint32_t dcol = tmem + (uint32_t)(d * 96);
                        umma_tf32(dcol, smem_desc(a_hi + aoff, kZtcPitch, 128), b_hi, idesc, j != 0);
                        umma_tf32(dcol, smem_desc(a_hi + aoff, kZtcPitch, 128), b_lo, idesc, 1u);
                        umma_tf32(dcol, smem_desc(a_lo + aoff, kZtcPitch, 128), b_hi, idesc, 1u);
                    }
                }
                umma_commit(bar_mma);
            }
            // ---- epilogue of the pass: rows that are first entries go out as (part of) their Z row
            mbar_wait(bar_mma, (uint32_t)((2 * it + pass) & 1));
            tc_fence_after();
            {
                const int q = warp & 3, half = warp >> 2;
                const int sub = lane & 3, l0 = lane & ~3;     // lane quad: after the transpose this thread holds float4 `sub`
                // of the rows of lanes l0 .. l0 + 3 = entries (l0 + i) * 4 + q of the tile
                const uint32_t f4 = *reinterpret_cast<const uint32_t*>(flags + q * 32 + l0);
                bool on[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) on[i] = ((f4 >> (8 * i)) & 0xffu) != 0u && (l0 + i) * 4 + q < rpt;
                float* zq = a.Z + (size_t)(e0 - e_begin + l0 * 4 + q) * 288 + dy0 * 96 + sub * 4;
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
                const int n16 = ndy * 3;                      // 16-column chunks per half (96 or 48 columns)
                auto quad_transpose = [&](uint32_t (&v)[16]) {  // unit = 4 registers (16 columns = 4 units per lane)
#pragma unroll
                    for (int mask = 1; mask <= 2; mask <<= 1) {
                        const bool upper = (sub & mask) != 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (u & mask) continue;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t lo = v[4 * u + k], hi = v[4 * (u | mask) + k];
                                const uint32_t got = __shfl_xor_sync(0xffffffffu, upper ? lo : hi, mask);
                                if (upper) v[4 * u + k] = got; else v[4 * (u | mask) + k] = got;
                            }
                        }
                    }
                };
                auto put16 = [&](const uint32_t (&v)[16], int col) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (on[i])
                            *reinterpret_cast<float4*>(zq + (size_t)i * (4 * 288) + col) = make_float4(
                                __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                };
                for (int cc = 0; cc < n16; cc += 2) {         // two chunks (32 columns) per TMEM round trip
                    const int col = half * n16 * 16 + cc * 16;
                    const bool two = cc + 1 < n16;
                    uint32_t va[16], vb[16];
                    SHPL_TMEM_LD16(taddr + col, va);
                    if (two) SHPL_TMEM_LD16(taddr + col + 16, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    quad_transpose(va);
                    put16(va, col);
                    if (two) {
                        quad_transpose(vb);
                        put16(vb, col + 16);
                    }
                }
            }
            if (pass == 0) {
                tc_fence_before();
                __syncthreads();                              // the accumulators are free for the second pass
                tc_fence_after();
            }
        }
        tc_fence_before();
        __syncthreads();                                      // the A planes, the flags and the accumulators are free again
        tc_fence_after();
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kZtcTmemCols) : "memory");
    }
}

// FFMA form (any C_s that is a multiple of 32): a warp takes a window of 4 entries and owns the cells whose first
// entry lies in it; the four pooled vectors go through shared memory ([ci][4]: one broadcast LDS.128 per ci serves four
// cells and nine taps), lane = output channel, 36 accumulators; the weights are read through L1.
constexpr int kZThreads = 256;
constexpr int kZWindow = 4;

__global__ void __launch_bounds__(kZThreads, 3) shpl_conv_z_kernel(ZArgs a) {
    __shared__ float4 pbuf_all[kZThreads / 32][32];      // per warp [32 ci][4 cells] pooled values
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    float4* pbuf = pbuf_all[warp];
    const int e_begin = __ldg(a.ptr), e_end = min(__ldg(a.ptr + a.n_rows), e_begin + a.nnz_max);
    const int n_windows = (e_end - e_begin + kZWindow - 1) / kZWindow;
    const float* wbase = a.w + (size_t)a.ci_off * a.c_out_total + lane;
    const size_t tap_stride = (size_t)a.c_in_total * a.c_out_total;
    for (int win = blockIdx.x * warps + warp; win < n_windows; win += gridDim.x * warps) {
        const int e0 = e_begin + win * kZWindow;
        int my_row = -1, my_end = 0;
        bool first = false;
        if (lane < kZWindow && e0 + lane < e_end) {
            my_row = __ldg(a.key + e0 + lane);
            first = (e0 + lane == e_begin) || (__ldg(a.key + e0 + lane - 1) != my_row);
            if (first) my_end = __ldg(a.ptr + my_row + 1);
        }
        uint32_t firsts = __ballot_sync(0xffffffffu, first);
        if (!firsts) continue;
        int fe[4], fend[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            fe[c] = -1;
            fend[c] = 0;
            if (firsts) {
                const int l = __ffs(firsts) - 1;
                firsts &= firsts - 1;
                fe[c] = e0 + l;
                fend[c] = __shfl_sync(0xffffffffu, my_end, l);
            }
        }
        float acc[9][4];
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
        for (int c0 = 0; c0 < a.C_s; c0 += 32) {
            float p[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {    // pooled[cell c][c0 + lane]: entries in stored order, two roundings each
                p[c] = 0.f;
                if (fe[c] >= 0)
                    for (int k = fe[c]; k < fend[c]; ++k)
                        p[c] = __fadd_rn(p[c], __fmul_rn(__ldg(a.val + k), __ldg(a.src + (size_t)__ldg(a.idx + k) * a.C_s + c0 + lane)));
            }
            __syncwarp();
            pbuf[lane] = make_float4(p[0], p[1], p[2], p[3]);
            __syncwarp();
            const float* wt = wbase + (size_t)c0 * a.c_out_total;
#pragma unroll 2
            for (int ci = 0; ci < 32; ++ci) {
                const float4 pv = pbuf[ci];
                float wv[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) wv[t] = __ldg(wt + t * tap_stride + (size_t)ci * a.c_out_total);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    acc[t][0] = fmaf(pv.x, wv[t], acc[t][0]);
                    acc[t][1] = fmaf(pv.y, wv[t], acc[t][1]);
                    acc[t][2] = fmaf(pv.z, wv[t], acc[t][2]);
                    acc[t][3] = fmaf(pv.w, wv[t], acc[t][3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (fe[c] >= 0) {
                float* z = a.Z + (size_t)(fe[c] - e_begin) * 288 + lane;
#pragma unroll
                for (int t = 0; t < 9; ++t) z[t * 32] = acc[t][c];
            }
    }
}

// ------------------------------------------------------------------------------------ backward
// out = conv3x3(concat(dst, pooled), W)  (the linear part; the activation's gradient is the caller's).  Given g_out:
//   g_dst    = conv3x3(g_out, W_d flipped and transposed)            -> the dense tcgen05 kernel on re-prepped weights
//   g_pooled = the same for the pooled channels, only at the cells that receive pooled features (shpl_conv_gp_kernel),
//   g_src    = transpose-CSR gather of g_pooled                       -> shpl_pool_backward_from on a remapped index
//   g_W[t][ci][co] = sum_p x[p + off(t)][ci] * g_out[p][co]           -> dense channels: shpl_conv_gwd_kernel (a pixel
//                    reduction, FFMA, per-CTA partial sums combined in order); pooled channels: inside shpl_conv_gp_kernel
// All sums run in a fixed order: deterministic.
struct GpArgs {
    const float* src;
    const float* g_out;
    const int* ptr;
    const int* key;
    const int* idx;
    const float* val;
    const float* w;              // HWIO weights
    int c_in_total, ci_off;      // C_out = 32, C_s = 32
    int n_rows, nnz_max, H, W;
    float* GP;                   // [entries][32] gradient of the pooled vector of the cell whose first entry is e
    float* part_p;               // [9][gridDim.x][32 ci][32 co] per-CTA partial sums of the pooled weight gradient, or NULL
};

// A CTA takes batches of 32 entries; a warp takes a window of four of them and owns the cells whose first entry lies in
// it: lane = output channel loads the 9 x 4 rows of g_out the four cells were seen through (all in flight together) and
// leaves them in shared memory as [tap][co] float4s; lane = input channel then does, per (tap, co), one conflict-free LDS
// of the transposed weight and one broadcast LDS.128 for four FFMAs (the first version: a cell per warp, a shuffle per
// FFMA and one load latency per tap, 75 us).  Sums in (tap, co) order, taps outside the map contributing exact zeros.
// The pooled channels' weight gradient is formed from the same staged rows: the eight warps share the 288 (tap, co)
// pairs (36 sums per lane = input channel) and walk the batch's cells in entry order with the pooled vectors the
// owners left in shared memory; per-CTA partial sums, added in CTA order by the reduce kernel (the separate kernel that
// re-read g_out per tap took 72 us).
constexpr int kGpWarps = 8;
constexpr int kGpPairs = 288 / kGpWarps;     // (tap, co) pairs of the weight gradient per warp
constexpr int kGpWtFloats = 9 * 32 * 33;     // [tap][co][33]: filled with coalesced loads, read with ci at stride 1
constexpr int kGpSmem = (kGpWtFloats + kGpWarps * 9 * 32 * 4 + 4 * kGpWarps * 32) * 4;
constexpr int kGpCtasPerSm = 2;

__global__ void __launch_bounds__(32 * kGpWarps, kGpCtasPerSm) shpl_conv_gp_kernel(GpArgs a) {
    extern __shared__ __align__(16) float gp_smem[];
    __shared__ int has_cells[kGpWarps];
    float* wt = gp_smem;
    for (int i = threadIdx.x; i < 9 * 32 * 32; i += blockDim.x) {
        const int co = i & 31, ci = (i >> 5) & 31, t = i >> 10;
        wt[(t * 32 + co) * 33 + ci] = __ldg(a.w + ((size_t)t * a.c_in_total + a.ci_off + ci) * 32 + co);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* gs_all = reinterpret_cast<float4*>(gp_smem + kGpWtFloats);     // [warp][tap][co] -> the four entries' g_out values
    float4* gs = gs_all + warp * 288;
    float* ps_all = gp_smem + kGpWtFloats + kGpWarps * 288 * 4;            // [entry of the batch][ci] pooled vectors (0: not a first entry)
    const int e_begin = __ldg(a.ptr), e_end = min(__ldg(a.ptr + a.n_rows), e_begin + a.nnz_max);
    const int HW = a.H * a.W;
    const bool weights = a.part_p != nullptr;
    float wacc[kGpPairs];                                                 // lane = ci; pair q = kGpPairs * warp + i = (tap, co)
#pragma unroll
    for (int i = 0; i < kGpPairs; ++i) wacc[i] = 0.f;
    // a CTA owns a contiguous range of entries (every CTA the same number of batches: no uneven last round)
    const int per = (((e_end - e_begin + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
    const int c0 = min(e_begin + (int)blockIdx.x * per, e_end), c1 = min(c0 + per, e_end);
    for (int b0 = c0; b0 < c1; b0 += 4 * kGpWarps) {
        const int w0 = b0 + warp * 4;
        int key_l = 0, idx_l = 0, end_l = 0;
        float val_l = 0.f;
        bool first_l = false;
        if (lane < 4 && w0 + lane < c1) {
            const int k = w0 + lane;
            key_l = __ldg(a.key + k);
            idx_l = __ldg(a.idx + k);
            val_l = __ldg(a.val + k);
            first_l = (k == e_begin) || (__ldg(a.key + k - 1) != key_l);
            if (first_l) end_l = __ldg(a.ptr + key_l + 1);
        }
        const unsigned firsts = __ballot_sync(0xffffffffu, first_l) & 0xfu;   // warp-uniform; 0: no cell starts in this window
        float g[9][4];                                                    // lane = co
        float pv[4] = {0.f, 0.f, 0.f, 0.f};                               // lane = ci: pooled vectors of the cells that start here
        if (firsts != 0u) {
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = __shfl_sync(0xffffffffu, idx_l, j);
                SHPL_DASSERT(w0 + j >= c1 || p >= 0);
                x[j] = (w0 + j < c1) ? __ldg(a.src + (size_t)p * 32 + lane) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = __shfl_sync(0xffffffffu, key_l, j);
                const bool fj = (firsts >> j) & 1u;
                const int f = r / HW, rem = r - f * HW, y = rem / a.W, xq = rem - y * a.W;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y - (t / 3 - 1), xx = xq - (t % 3 - 1);    // the output pixel that saw this cell through tap t
                    const bool ok = fj && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
                    g[t][j] = ok ? __ldg(a.g_out + ((size_t)(f * a.H + yy) * a.W + xx) * 32 + lane) : 0.f;
                }
            }
            if (weights) {   // pooled[cell][lane]: entries in stored order
                float p = 0.f;
                int open = -1, open_end = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((firsts >> j) & 1u) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (jj == open) pv[jj] = p;
                        open = j;
                        open_end = __shfl_sync(0xffffffffu, end_l, j);
                        p = 0.f;
                    }
                    if (open >= 0 && w0 + j < open_end) p = __fadd_rn(p, __fmul_rn(__shfl_sync(0xffffffffu, val_l, j), x[j]));
                }
                for (int k = w0 + 4; k < open_end; ++k)                   // the last cell runs on past the window
                    p = __fadd_rn(p, __fmul_rn(__ldg(a.val + k), __ldg(a.src + (size_t)__ldg(a.idx + k) * 32 + lane)));
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (jj == open) pv[jj] = p;
            }
        }
        __syncthreads();                                                  // the last batch's reads of gs / ps (and the fill of wt) are done
        if (firsts != 0u) {
#pragma unroll
            for (int t = 0; t < 9; ++t) gs[t * 32 + lane] = make_float4(g[t][0], g[t][1], g[t][2], g[t][3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ps_all[(warp * 4 + j) * 32 + lane] = pv[j];
        }
        if (lane == 0) has_cells[warp] = firsts != 0u;
        __syncthreads();
        if (firsts != 0u) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};                          // g_pooled[cell j][ci = lane]
            const float* wl = wt + lane;
            for (int t = 0; t < 9; ++t) {
#pragma unroll 8
                for (int co = 0; co < 32; ++co) {
                    const float w = wl[(t * 32 + co) * 33];
                    const float4 gv = gs[t * 32 + co];
                    acc[0] = fmaf(gv.x, w, acc[0]);
                    acc[1] = fmaf(gv.y, w, acc[1]);
                    acc[2] = fmaf(gv.z, w, acc[2]);
                    acc[3] = fmaf(gv.w, w, acc[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((firsts >> j) & 1u) a.GP[(size_t)(w0 + j - e_begin) * 32 + lane] = acc[j];
        }
        if (weights) {
            for (int ww = 0; ww < kGpWarps; ++ww) {                       // the batch's cells in entry order
                if (!has_cells[ww]) continue;
                const float p0 = ps_all[(ww * 4) * 32 + lane], p1 = ps_all[(ww * 4 + 1) * 32 + lane];
                const float p2 = ps_all[(ww * 4 + 2) * 32 + lane], p3 = ps_all[(ww * 4 + 3) * 32 + lane];
                const float4* gq = gs_all + ww * 288 + kGpPairs * warp;
#pragma unroll
                for (int i = 0; i < kGpPairs; ++i) {
                    const float4 gv = gq[i];
                    wacc[i] = fmaf(p0, gv.x, wacc[i]);
                    wacc[i] = fmaf(p1, gv.y, wacc[i]);
                    wacc[i] = fmaf(p2, gv.z, wacc[i]);
                    wacc[i] = fmaf(p3, gv.w, wacc[i]);
                }
            }
        }
    }
    if (weights) {
#pragma unroll
        for (int i = 0; i < kGpPairs; ++i) {
            const int q = kGpPairs * warp + i, t = q >> 5, co = q & 31;
            a.part_p[(((size_t)t * gridDim.x + blockIdx.x) * 32 + lane) * 32 + co] = wacc[i];
        }
    }
}

// idx_remap[k] = first-entry index of the destination cell of transposed entry k (the row of GP to gather)
__global__ void shpl_conv_remap_kernel(const int* __restrict__ ptr, const int* __restrict__ ptrT, const int* __restrict__ idxT, int n_rows,
                                       int n_src, int nnz_max, int* __restrict__ remap) {
    const int e_begin = __ldg(ptr), t_begin = __ldg(ptrT), t_end = min(__ldg(ptrT + n_src), t_begin + nnz_max);
    const int k = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t_end) return;
    const int r = __ldg(idxT + k);
    SHPL_DASSERT(r >= 0 && r < n_rows);
    remap[k] = __ldg(ptr + r) - e_begin;
}

// dense channels: persistent CTAs of 256 threads (two per SM) over 8 x 16-pixel tiles with a shared-memory halo, double
// buffered with cp.async (the next tile lands while this one is summed); the two halves of a CTA take the upper and
// the lower four rows of the tile and keep their own partial sums: thread = (half, input-channel pair, output-channel
// quad) holds 9 taps x 2 x 4 sums in registers, per pixel one LDS.128 of g_out and nine LDS.64 of the input feed 72 FFMAs
// (the first version, one input channel per thread, fed 36 and was bound by the loads; the second, 128-thread CTAs that
// loaded and summed in turn, left the SM idle during the loads and ended on an uneven last round: 247 us).
constexpr int kGwTileY = 8, kGwTileX = 16, kGwThreads = 256, kGwCtasPerSm = 2;
constexpr int kGwPartsPerCta = 2;      // the halves of a CTA
constexpr int kGwHaloFloats = (kGwTileY + 2) * (kGwTileX + 2) * 32, kGwTileFloats = kGwTileY * kGwTileX * 32;
constexpr int kGwBufFloats = kGwHaloFloats + kGwTileFloats;
constexpr int kGwdSmem = 2 * kGwBufFloats * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {      // bytes = 16, or 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kGwThreads, kGwCtasPerSm) shpl_conv_gwd_kernel(const float* __restrict__ x, const float* __restrict__ g_out, int frames, int H, int W,
                                                                                  float* __restrict__ part_d) {
    extern __shared__ __align__(16) float gw_smem[];
    // buffer b: [10][18][32] halo of the input map, then [8][16][32] the output-gradient tile (zeros outside the image)
    const int tiles_x = (W + kGwTileX - 1) / kGwTileX, tiles_y = (H + kGwTileY - 1) / kGwTileY;
    const int n_tiles = frames * tiles_x * tiles_y;
    const int half = threadIdx.x >> 7, cp = (threadIdx.x & 127) >> 3, cq = threadIdx.x & 7;      // input channels 2 cp, 2 cp + 1; output channels 4 cq .. 4 cq + 3
    const uint32_t smem0 = smem_u32(gw_smem);
    auto issue = [&](int tile, int buf) {
        const int f = tile / (tiles_x * tiles_y), tt = tile - f * tiles_x * tiles_y;
        const int y0 = (tt / tiles_x) * kGwTileY, x0 = (tt % tiles_x) * kGwTileX;
        const uint32_t xs = smem0 + buf * (kGwBufFloats * 4), gs = xs + kGwHaloFloats * 4;
        for (int i = threadIdx.x; i < kGwHaloFloats / 4; i += kGwThreads) {      // float4 units: [10][18][8]
            const int q = i & 7, px = (i >> 3) % (kGwTileX + 2), py = (i >> 3) / (kGwTileX + 2);
            const int yy = y0 - 1 + py, xx = x0 - 1 + px;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            cp_async16(xs + i * 16, ok ? x + ((size_t)(f * H + yy) * W + xx) * 32 + q * 4 : x, ok ? 16 : 0);
        }
        for (int i = threadIdx.x; i < kGwTileFloats / 4; i += kGwThreads) {
            const int q = i & 7, px = (i >> 3) % kGwTileX, py = (i >> 3) / kGwTileX;
            const int yy = y0 + py, xx = x0 + px;
            const bool ok = yy < H && xx < W;
            cp_async16(gs + i * 16, ok ? g_out + ((size_t)(f * H + yy) * W + xx) * 32 + q * 4 : g_out, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float4 acc[9][2];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    int buf = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        __syncthreads();                                   // the other buffer's sums (the tile before this one) are done
        const int next = tile + gridDim.x;
        if (next < n_tiles) {
            issue(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                   // this tile has landed for every thread
        const float* xs = gw_smem + buf * kGwBufFloats;
        const float* gs = xs + kGwHaloFloats;
        for (int py = half * (kGwTileY / 2); py < (half + 1) * (kGwTileY / 2); ++py) {
#pragma unroll 2
            for (int px = 0; px < kGwTileX; ++px) {
                const float4 g = reinterpret_cast<const float4*>(gs + (py * kGwTileX + px) * 32)[cq];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float2 xv = reinterpret_cast<const float2*>(xs + ((py + t / 3) * (kGwTileX + 2) + px + t % 3) * 32)[cp];
                    acc[t][0].x = fmaf(xv.x, g.x, acc[t][0].x); acc[t][0].y = fmaf(xv.x, g.y, acc[t][0].y);
                    acc[t][0].z = fmaf(xv.x, g.z, acc[t][0].z); acc[t][0].w = fmaf(xv.x, g.w, acc[t][0].w);
                    acc[t][1].x = fmaf(xv.y, g.x, acc[t][1].x); acc[t][1].y = fmaf(xv.y, g.y, acc[t][1].y);
                    acc[t][1].z = fmaf(xv.y, g.z, acc[t][1].z); acc[t][1].w = fmaf(xv.y, g.w, acc[t][1].w);
                }
            }
        }
    }
    const size_t part = (size_t)blockIdx.x * kGwPartsPerCta + half;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        reinterpret_cast<float4*>(part_d + ((part * 9 + t) * 32 + 2 * cp) * 32)[cq] = acc[t][0];
        reinterpret_cast<float4*>(part_d + ((part * 9 + t) * 32 + 2 * cp + 1) * 32)[cq] = acc[t][1];
    }
}

// g_weight[t][ci][co] = the partial sums added in a fixed tree (deterministic): a block takes one (t, ci) row of 32 output
// channels, slice k of its eight sums the k-th eighth of the partial sums in CTA / chunk order (eight loads in flight),
// and the eight slice sums are added in slice order.
constexpr int kGwReduceSlices = 8;

__global__ void __launch_bounds__(32 * kGwReduceSlices) shpl_conv_gw_reduce_kernel(const float* __restrict__ part_d, int n_ctas,
                                                                                   const float* __restrict__ part_p, int n_chunks_p,
                                                                                   int c_in_total, float* __restrict__ g_weight) {
    __shared__ float sums[kGwReduceSlices][32];
    const int co = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int ci = blockIdx.x % c_in_total, t = blockIdx.x / c_in_total;
    const float* p = nullptr;
    size_t pitch = 0;
    int n = 0;
    if (ci < 32) {
        p = part_d + ((size_t)t * 32 + ci) * 32 + co;
        pitch = (size_t)9 * 1024;
        n = n_ctas;
    } else if (part_p != nullptr) {
        p = part_p + ((size_t)t * n_chunks_p * 32 + ci - 32) * 32 + co;
        pitch = 1024;
        n = n_chunks_p;
    }
    const int per = (n + kGwReduceSlices - 1) / kGwReduceSlices;
    const int c1 = min(n, (k + 1) * per);
    int c = k * per;
    float s = 0.f;
    for (; c + 8 <= c1; c += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = p[(size_t)(c + j) * pitch];
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
    }
    for (; c < c1; ++c) s += p[(size_t)c * pitch];
    sums[k][co] = s;
    __syncthreads();
    if (k == 0) {
        float r = sums[0][co];
#pragma unroll
        for (int j = 1; j < kGwReduceSlices; ++j) r += sums[j][co];
        g_weight[((size_t)t * c_in_total + ci) * 32 + co] = r;
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
    float* wprep;       // dense channels' weights, B-operand layout
    float* wprep_p;     // pooled channels' weights, B-operand layout (tensor-core Z kernel)
    uint32_t* busy;
    float* Z;           // [nnz_max][9][32]
    size_t words;
    size_t bytes;
};

ConvWorkspace carve(void* ws, long long cells, long long nnz_max) {
    ConvWorkspace c;
    c.words = (size_t)((cells + 31) / 32);
    const size_t bm = align_up(c.words * 4, 256);
    uint8_t* p = static_cast<uint8_t*>(ws);
    c.wprep = reinterpret_cast<float*>(p);
    c.wprep_p = reinterpret_cast<float*>(p + kWBytes);
    c.busy = reinterpret_cast<uint32_t*>(p + 2 * kWBytes);
    c.bytes = 2 * kWBytes + bm;
    c.Z = reinterpret_cast<float*>(p + c.bytes);
    c.bytes += align_up((size_t)(nnz_max > 0 ? nnz_max : 0) * 288 * sizeof(float), 256);
    return c;
}

int launch_dense(const float* in, int in_pitch, float* out, const float* wprep, const float* scale, const float* shift, int relu,
                 const uint32_t* busy, const int* ptr, const float* Z, int z_rows, int frames, int H, int W, cudaStream_t s) {
    CUtensorMap map_in, map_out;
    if (int rc = make_map(&map_in, in, frames, H, W, in_pitch, kHaloY, kHaloX, false)) return rc;
    if (int rc = make_map(&map_out, out, frames, H, W, kC, kTileY, kTileX, true)) return rc;
    ConvArgs a{};
    a.wprep = wprep;
    a.scale = scale;
    a.shift = shift;
    a.busy = busy;
    a.ptr = ptr;
    a.Z = Z;
    a.z_rows = z_rows;
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
    // After the Z kernel the launch is a programmatic dependent one: the CTAs start as the Z kernel's CTAs leave their SMs
    // and run the dense half at once; only the sparse warps wait (griddepcontrol.wait) for the Z rows and the bitmap.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = kConvSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (kConvPdl && busy != nullptr) ? 1 : 0;
    SHPL_CUDA_OK(cudaLaunchKernelEx(&cfg, shpl_conv3x3_dense_kernel, map_in, map_out, a));
    shpl::count_launches(1);
    return shpl::check_launch("shpl_conv3x3_dense_kernel");
}

}  // namespace

extern "C" size_t shpl_conv3x3_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t nnz_max) {
    if (frames <= 0 || H <= 0 || W <= 0 || nnz_max < 0) return 0;
    return carve(nullptr, (long long)frames * H * W, nnz_max).bytes;
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
    const ConvWorkspace c = carve(workspace, cells, C_s > 0 ? nnz_max : 0);
    SHPL_REQUIRE(workspace_bytes >= c.bytes, SHPL_ERR_WORKSPACE_TOO_SMALL, "shpl_pool_conv3x3_forward: workspace %zu < %zu bytes",
                 workspace_bytes, c.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int c_in_total = C_d + C_s;
    const bool sparse = C_s > 0 && nnz_max > 0;
    const bool z_tc = sparse && C_s == kC;        // the pooled half's weights are prepped by the same launch
    shpl_conv_prep_kernel<<<dim3((kWElems + 255) / 256, z_tc ? 2 : 1), 256, 0, s>>>(weight, c_in_total, C_out, 0, 0, c.wprep, C_d, c.wprep_p,
                                                                                   sparse ? c.busy : nullptr, sparse ? (int)c.words : 0);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_conv_prep_kernel")) return rc;
    if (sparse) {
        if (!z_tc) {      // the tensor-core Z kernel sets the bitmap itself
            shpl_conv_mark_kernel<<<(nnz_max + 255) / 256, 256, 0, s>>>(ptr, key, (int)cells, nnz_max, c.busy);
            shpl::count_launches(1);
            if (int rc = shpl::check_launch("shpl_conv_mark_kernel")) return rc;
        }
        ZArgs za{};
        za.src = src;
        za.ptr = ptr;
        za.key = key;
        za.idx = idx;
        za.val = val;
        za.w = weight;
        za.wprep = c.wprep_p;
        za.c_in_total = c_in_total;
        za.c_out_total = C_out;
        za.ci_off = C_d;
        za.C_s = C_s;
        za.n_rows = (int)cells;
        za.nnz_max = nnz_max;
        za.Z = c.Z;
        za.busy = c.busy;
        if (z_tc) {
            static bool z_attr = false;
            if (!z_attr) {
                SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_conv_z_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kZtcSmemRequest));
                z_attr = true;
            }
            // tiles sized so that the entries spread over two CTAs on every SM (the per-entry work is what costs)
            const int slots = 2 * shpl::sm_count();
            int rpt = ((nnz_max + slots - 1) / slots + 7) & ~7;
            rpt = rpt < 32 ? 32 : rpt > 128 ? 128 : rpt;
            za.rows_per_tile = rpt;
            const int ztiles = (nnz_max + rpt - 1) / rpt;
            // a programmatic dependent of the prep launch: entries and source rows are gathered under it
            cudaLaunchConfig_t zcfg = {};
            zcfg.gridDim = dim3(ztiles < slots ? ztiles : slots);
            zcfg.blockDim = dim3(kZtcThreads);
            zcfg.dynamicSmemBytes = kZtcSmemRequest;
            zcfg.stream = s;
            cudaLaunchAttribute zattr[1];
            zattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            zattr[0].val.programmaticStreamSerializationAllowed = 1;
            zcfg.attrs = zattr;
            zcfg.numAttrs = kConvPdl ? 1 : 0;
            SHPL_CUDA_OK(cudaLaunchKernelEx(&zcfg, shpl_conv_z_tc_kernel, za));
            shpl::count_launches(1);
            if (int rc = shpl::check_launch("shpl_conv_z_tc_kernel")) return rc;
        } else {
            const int windows = (nnz_max + kZWindow - 1) / kZWindow;
            int zgrid = (windows + kZThreads / 32 - 1) / (kZThreads / 32);
            const int zcap = shpl::sm_count() * 6;
            if (zgrid > zcap) zgrid = zcap;
            shpl_conv_z_kernel<<<zgrid, kZThreads, 0, s>>>(za);
            shpl::count_launches(1);
            if (int rc = shpl::check_launch("shpl_conv_z_kernel")) return rc;
        }
    }
    if (int rc = launch_dense(dst, C_d, out, c.wprep, scale, shift, relu, sparse ? c.busy : nullptr, ptr, c.Z, nnz_max, frames, H, W, s)) return rc;
    return SHPL_OK;
}

namespace {
struct BwdWorkspace {
    float* wprep;
    float* GP;
    int* remap;
    float* part_d;
    float* part_p;
    int n_ctas;
    size_t bytes;
};
BwdWorkspace carve_bwd(void* ws, long long nnz_max) {
    BwdWorkspace b;
    uint8_t* p = static_cast<uint8_t*>(ws);
    size_t off = 0;
    auto take = [&](size_t n) { uint8_t* q = p + off; off += align_up(n, 256); return q; };
    const size_t n = (size_t)(nnz_max > 0 ? nnz_max : 0);
    b.n_ctas = kGwPartsPerCta * kGwCtasPerSm * 148;   // partial sums of the dense weight-gradient kernel (fixed: part of the summation tree)
    b.wprep = reinterpret_cast<float*>(take(kWBytes));
    b.GP = reinterpret_cast<float*>(take(n * 32 * 4));
    b.remap = reinterpret_cast<int*>(take(n * 4));
    b.part_d = reinterpret_cast<float*>(take((size_t)b.n_ctas * 9 * 1024 * 4));
    b.part_p = reinterpret_cast<float*>(take((size_t)9 * kGpCtasPerSm * 148 * 1024 * 4));      // per-CTA partial sums of the gp kernel
    b.bytes = off;
    return b;
}
}  // namespace

extern "C" size_t shpl_conv3x3_backward_workspace_bytes(int32_t nnz_max) {
    if (nnz_max < 0) return 0;
    return carve_bwd(nullptr, nnz_max).bytes;
}

extern "C" int shpl_pool_conv3x3_backward(const float* g_out, const float* dst, const float* src, const int32_t* ptr, const int32_t* key,
                                          const int32_t* idx, const float* val, const int32_t* ptrT, const int32_t* keyT,
                                          const int32_t* idxT, const float* valT, int32_t nnz_max, int32_t frames, int32_t H, int32_t W,
                                          int32_t C_d, int32_t n_src, int32_t C_s, const float* weight, int32_t C_out, float* g_dst,
                                          float* g_src, float* g_weight, void* workspace, size_t workspace_bytes, void* stream) {
    SHPL_REQUIRE(frames > 0 && H > 0 && W > 0 && n_src >= 0 && nnz_max >= 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_backward: bad sizes frames=%d H=%d W=%d n_src=%d", frames, H, W, n_src);
    SHPL_REQUIRE(g_out && weight && workspace && (g_weight == nullptr || dst), SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_backward: null pointer");
    SHPL_REQUIRE(C_d == kC && C_out == kC && (C_s == 0 || C_s == 32), SHPL_ERR_UNSUPPORTED,
                 "shpl_pool_conv3x3_backward: built for C_d = C_out = 32 and C_s in {0, 32} (got %d, %d, %d)", C_d, C_out, C_s);
    SHPL_REQUIRE(C_s == 0 || (src && ptr && key && idx && val && ptrT && keyT && idxT && valT), SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_conv3x3_backward: pooled channels need src and both CSR forms");
    SHPL_REQUIRE((long long)frames * H * W < (1ll << 31), SHPL_ERR_UNSUPPORTED, "shpl_pool_conv3x3_backward: map too large");
    SHPL_REQUIRE(shpl::aligned(g_out, 16) && shpl::aligned(workspace, 256) && (!g_dst || shpl::aligned(g_dst, 16)) && (!dst || shpl::aligned(dst, 16)),
                 SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_conv3x3_backward: maps must be 16-byte aligned, workspace 256-byte aligned");
    const long long cells = (long long)frames * H * W;
    const bool sparse = C_s > 0 && nnz_max > 0;
    BwdWorkspace b = carve_bwd(workspace, sparse ? nnz_max : 0);
    b.n_ctas = kGwPartsPerCta * kGwCtasPerSm * shpl::sm_count() < b.n_ctas ? kGwPartsPerCta * kGwCtasPerSm * shpl::sm_count() : b.n_ctas;
    SHPL_REQUIRE(workspace_bytes >= b.bytes, SHPL_ERR_WORKSPACE_TOO_SMALL, "shpl_pool_conv3x3_backward: workspace %zu < %zu bytes",
                 workspace_bytes, b.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int c_in_total = C_d + C_s;
    if (g_dst != nullptr) {       // the dense kernel on spatially flipped, in/out-transposed weights
        shpl_conv_prep_kernel<<<(kWElems + 255) / 256, 256, 0, s>>>(weight, c_in_total, C_out, 0, 1, b.wprep);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_prep_kernel")) return rc;
        if (int rc = launch_dense(g_out, kC, g_dst, b.wprep, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, frames, H, W, s)) return rc;
    }
    int ggrid = 0;                 // CTAs of the gp kernel = partial sums of the pooled weight gradient
    if (sparse && (g_src != nullptr || g_weight != nullptr)) {
        GpArgs ga{};
        ga.src = src;
        ga.g_out = g_out;
        ga.ptr = ptr;
        ga.key = key;
        ga.idx = idx;
        ga.val = val;
        ga.w = weight;
        ga.c_in_total = c_in_total;
        ga.ci_off = C_d;
        ga.n_rows = (int)cells;
        ga.nnz_max = nnz_max;
        ga.H = H;
        ga.W = W;
        ga.GP = b.GP;
        ga.part_p = g_weight != nullptr ? b.part_p : nullptr;
        static bool gp_attr = false;
        if (!gp_attr) {
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_conv_gp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGpSmem));
            gp_attr = true;
        }
        ggrid = (nnz_max + 4 * kGpWarps - 1) / (4 * kGpWarps);            // a window of four entries per warp
        const int gcap = kGpCtasPerSm * (shpl::sm_count() < 148 ? shpl::sm_count() : 148);      // (the workspace holds 148 SMs' partial sums)
        if (ggrid > gcap) ggrid = gcap;
        shpl_conv_gp_kernel<<<ggrid, 32 * kGpWarps, kGpSmem, s>>>(ga);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_gp_kernel")) return rc;
        if (g_src != nullptr) {
            shpl_conv_remap_kernel<<<(nnz_max + 255) / 256, 256, 0, s>>>(ptr, ptrT, idxT, (int)cells, n_src, nnz_max, b.remap);
            shpl::count_launches(1);
            if (int rc = shpl::check_launch("shpl_conv_remap_kernel")) return rc;
            // g_src[p] = sum over the entries at pixel p, in stored order, of val * g_pooled[cell]: the transpose-CSR kernel
            if (int rc = shpl_pool_backward_from(b.GP, 32, 0, ptrT, keyT, b.remap, valT, nnz_max, 0, nnz_max, n_src, 32, g_src, stream)) return rc;
        }
    } else if (g_src != nullptr && C_s > 0) {
        SHPL_CUDA_OK(cudaMemsetAsync(g_src, 0, (size_t)n_src * C_s * sizeof(float), s));
    }
    if (g_weight != nullptr) {
        static bool attr_set = false;
        if (!attr_set) {
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_conv_gwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGwdSmem));
            attr_set = true;
        }
        shpl_conv_gwd_kernel<<<b.n_ctas / kGwPartsPerCta, kGwThreads, kGwdSmem, s>>>(dst, g_out, frames, H, W, b.part_d);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_gwd_kernel")) return rc;
        shpl_conv_gw_reduce_kernel<<<9 * c_in_total, 32 * kGwReduceSlices, 0, s>>>(b.part_d, b.n_ctas, sparse ? b.part_p : nullptr, ggrid, c_in_total, g_weight);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_conv_gw_reduce_kernel")) return rc;
    }
    return SHPL_OK;
}
