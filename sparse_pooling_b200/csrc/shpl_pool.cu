// SHPL pooling kernels for sm_100a: forward gather-SpMM fused with the channel
// concat, and its deterministic (atomic-free) transpose-CSR backward.
//
// Replaces the TF graph ops of /root/reference/avod/avod/utils/sparse_pool_utils.py
//   :96-103  _sparse_pool_op        gather_nd -> sparse_tensor_dense_matmul -> reshape
//   :105-117 _sparse_pool_trans_op  sparse_transpose + matmul -> scatter_nd
//   :72,:87  tf.concat(axis=3)
// and the gradients TF autodiff derives for them.
//
// The work is a streaming copy with rare gathers (2 % of BEV cells are hit at KITTI
// stride 1), so the design is a copy-class kernel: 128-bit coalesced accesses, 8
// independent loads in flight per lane, streaming cache hints on the dense traffic,
// a warp per tile of destination cells.  A tile is <= 32 cells so that one lane per
// cell holds the CSR offsets and the whole tile's emptiness is one ballot.
// Sums are accumulated in stored (ascending k) order with separately rounded
// multiply and add, which makes the result bit-identical to the sequential oracle.
#include "shpl_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 8;
constexpr unsigned kFull = 0xffffffffu;

template <int W> struct VecOf;
template <> struct VecOf<4> { using type = float4; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<1> { using type = float; };

__device__ __forceinline__ float4 vzero(float4*) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float2 vzero(float2*) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ float vzero(float*) { return 0.f; }

// acc + w*x with two roundings (no fma contraction): the oracle's order of operations
__device__ __forceinline__ void axpy(float4& a, float w, const float4& x) {
    a.x = __fadd_rn(a.x, __fmul_rn(w, x.x));
    a.y = __fadd_rn(a.y, __fmul_rn(w, x.y));
    a.z = __fadd_rn(a.z, __fmul_rn(w, x.z));
    a.w = __fadd_rn(a.w, __fmul_rn(w, x.w));
}
__device__ __forceinline__ void axpy(float2& a, float w, const float2& x) {
    a.x = __fadd_rn(a.x, __fmul_rn(w, x.x));
    a.y = __fadd_rn(a.y, __fmul_rn(w, x.y));
}
__device__ __forceinline__ void axpy(float& a, float w, const float& x) { a = __fadd_rn(a, __fmul_rn(w, x)); }

// slot s of a tile -> (cell r inside the tile, vector q inside the cell); nv = vectors per cell
__device__ __forceinline__ void split(int s, int nv, int shift, int& r, int& q) {
    if (shift >= 0) {
        r = s >> shift;
        q = s & (nv - 1);
    } else {
        r = s / nv;
        q = s - r * nv;
    }
}

// Dense part: `rows` cells of `nv` vectors, in[r*in_stride + q] -> out[r*out_stride + q].
template <typename V>
__device__ __forceinline__ void copy_tile(const V* __restrict__ in, int in_stride, V* __restrict__ out,
                                          int out_stride, int nv, int shift, int rows, int lane) {
    const int n = rows * nv;
    for (int s0 = 0; s0 < n; s0 += 32 * kUnroll) {
        V v[kUnroll];
        int o[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            if (s < n) {
                int r, q;
                split(s, nv, shift, r, q);
                v[j] = __ldcs(in + r * in_stride + q);
                o[j] = r * out_stride + q;
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            if (s < n) __stcs(out + o[j], v[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Wide cells (>= 32 vectors per cell, i.e. C_s >= 128 with float4): a WARP PER OUTPUT CELL.
// A CTA owns a tile of 32 cells.  All 8 warps stream the dense part of the tile; the cells that
// receive contributions are dealt round-robin (by rank among the busy cells) to the warps, so a
// run of adjacent crowded cells -- the near-range ground cells of a stride-8 BEV map -- is spread
// over the CTA instead of being walked by one warp.  Inside a cell the lanes own channel vectors
// q = lane + 32*a; the cell's (idx, val) entries are fetched 32 at a time with one coalesced load
// and handed round by shuffles, so kGatherUnroll entries x ACC vectors are in flight per lane
// while the adds still run in ascending k.
constexpr int kGatherUnroll = 8;
constexpr int kWideTile = 32;

template <typename V, int ACC>
__device__ __forceinline__ void pool_row_wide(const V* __restrict__ src, int src_stride, int beg, int end,
                                              const int* __restrict__ idx, const float* __restrict__ val,
                                              V* __restrict__ orow, int nv, int lane) {
    for (int q0 = 0; q0 < nv; q0 += 32 * ACC) {
        V acc[ACC];
#pragma unroll
        for (int a = 0; a < ACC; ++a) acc[a] = vzero((V*)nullptr);
        for (int c = beg; c < end; c += 32) {
            int my_p = 0;
            float my_w = 0.f;
            if (c + lane < end) {
                my_p = __ldg(idx + c + lane);
                my_w = __ldg(val + c + lane);
            }
            const int cnt = min(32, end - c);
            for (int e = 0; e < cnt; e += kGatherUnroll) {
                V x[kGatherUnroll][ACC];
                float w[kGatherUnroll];
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
                    const int p = __shfl_sync(kFull, my_p, (e + j) & 31);
                    w[j] = __shfl_sync(kFull, my_w, (e + j) & 31);
                    const V* row = src + (size_t)p * src_stride;
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        const int q = q0 + a * 32 + lane;
                        if (e + j < cnt && q < nv) x[j][a] = __ldg(row + q);
                    }
                }
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        const int q = q0 + a * 32 + lane;
                        if (e + j < cnt && q < nv) axpy(acc[a], w[j], x[j][a]);
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            const int q = q0 + a * 32 + lane;
            if (q < nv) __stcs(orow + q, acc[a]);
        }
    }
}

// CTA-cooperative dense copy of `rows` cells (flat over rows*nv vectors, kUnroll loads in flight per lane)
template <typename V>
__device__ __forceinline__ void cta_copy_tile(const V* __restrict__ in, int in_stride, V* __restrict__ out,
                                              int out_stride, int nv, int shift, int rows, int warp, int lane) {
    const int n = rows * nv;
    for (int s0 = warp * 32 * kUnroll; s0 < n; s0 += kWarps * 32 * kUnroll) {
        V v[kUnroll];
        int o[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            if (s < n) {
                int r, q;
                split(s, nv, shift, r, q);
                v[j] = __ldcs(in + r * in_stride + q);
                o[j] = r * out_stride + q;
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            if (s < n) __stcs(out + o[j], v[j]);
        }
    }
}

// CTA-cooperative sparse part of a tile of <= 32 wide cells
template <typename V, int ACC>
__device__ __forceinline__ void cta_pool_tile_wide(const V* __restrict__ src, int src_stride,
                                                   const int* __restrict__ ptr, const int* __restrict__ idx,
                                                   const float* __restrict__ val, V* __restrict__ out,
                                                   int out_stride, int nv, int rows, int warp, int lane) {
    int lo = 0, hi = 0;
    if (lane < rows) {
        lo = __ldg(ptr + lane);
        hi = __ldg(ptr + lane + 1);
    }
    const unsigned busy = __ballot_sync(kFull, hi > lo);
    // empty cells: zeros, one cell per warp at a time
    const V z = vzero((V*)nullptr);
    for (int r = warp; r < rows; r += kWarps) {
        if ((busy >> r) & 1u) continue;
        V* orow = out + r * out_stride;
        for (int q = lane; q < nv; q += 32) __stcs(orow + q, z);
    }
    // busy cells: rank j among the busy ones goes to warp j % kWarps
    unsigned m = busy;
    int j = 0;
    while (m) {
        const int r = __ffs(m) - 1;
        m &= m - 1;
        if ((j++ % kWarps) != warp) continue;
        const int beg = __shfl_sync(kFull, lo, r);
        const int end = __shfl_sync(kFull, hi, r);
        pool_row_wide<V, ACC>(src, src_stride, beg, end, idx, val, out + r * out_stride, nv, lane);
    }
}

// Entry-parallel gather for wide cells: one warp takes 32 consecutive entries of the key-sorted
// entry list and owns every cell whose FIRST entry lies in that chunk (a cell is never split, so
// its sum keeps the ascending-k order; the warp reads on past the chunk until the cell ends, and
// skips leading entries that continue a cell begun in the previous chunk).  Work per warp is
// 32 entries +- one cell, whatever the row-length skew: the crowded near-range cells of a
// stride-8 BEV map no longer serialise on one CTA.  Entries are streamed through a segmented sum:
// kGatherUnroll x ACC gathers in flight, the accumulator is flushed when the key changes.
constexpr int kEntryChunk = 32;

template <typename V, int ACC>
__device__ __forceinline__ void pool_entries_wide(const V* __restrict__ src, int src_stride,
                                                  const int* __restrict__ key, const int* __restrict__ idx,
                                                  const float* __restrict__ val, int e0, int e1, int e_begin,
                                                  int e_end, V* __restrict__ out, int out_stride, int nv, int lane) {
    const int prev_row = (e0 > e_begin) ? __ldg(key + e0 - 1) : -1;
    for (int q0 = 0; q0 < nv; q0 += 32 * ACC) {
        int base = e0;
        int my_row = -1, my_p = 0;
        float my_w = 0.f;
        if (base + lane < e_end) {
            my_row = __ldg(key + base + lane);
            my_p = __ldg(idx + base + lane);
            my_w = __ldg(val + base + lane);
        }
        int pos = 0;
        if (prev_row >= 0) {
            const unsigned fresh = __ballot_sync(kFull, base + lane < e_end && my_row != prev_row);
            pos = fresh ? __ffs(fresh) - 1 : 32;
        }
        if (base + pos >= e1) return;          // no cell starts in this chunk
        int cur_row = -1;
        bool finished = false;
        V acc[ACC];
#pragma unroll
        for (int a = 0; a < ACC; ++a) acc[a] = vzero((V*)nullptr);
        while (true) {
            const int cnt = min(32, e_end - base);
            while (pos < cnt) {
                V x[kGatherUnroll][ACC];
                float w[kGatherUnroll];
                int row[kGatherUnroll];
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
                    const int ej = pos + j;
                    row[j] = __shfl_sync(kFull, my_row, ej & 31);
                    const int p = __shfl_sync(kFull, my_p, ej & 31);
                    w[j] = __shfl_sync(kFull, my_w, ej & 31);
                    const V* srow = src + (size_t)p * src_stride;
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        const int q = q0 + a * 32 + lane;
                        if (ej < cnt && q < nv) x[j][a] = __ldg(srow + q);
                    }
                }
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
                    if (finished || pos + j >= cnt) continue;
                    if (row[j] != cur_row) {
                        if (cur_row >= 0) {
#pragma unroll
                            for (int a = 0; a < ACC; ++a) {
                                const int q = q0 + a * 32 + lane;
                                if (q < nv) __stcs(out + (size_t)cur_row * out_stride + q, acc[a]);
                                acc[a] = vzero((V*)nullptr);
                            }
                        }
                        if (base + pos + j >= e1) {   // the next cell belongs to the next warp
                            finished = true;
                            cur_row = -1;
                            continue;
                        }
                        cur_row = row[j];
                    }
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        const int q = q0 + a * 32 + lane;
                        if (q < nv) axpy(acc[a], w[j], x[j][a]);
                    }
                }
                if (finished) break;
                pos += kGatherUnroll;
            }
            if (finished) break;
            base += 32;
            if (base >= e_end) break;
            my_row = -1;
            if (base + lane < e_end) {
                my_row = __ldg(key + base + lane);
                my_p = __ldg(idx + base + lane);
                my_w = __ldg(val + base + lane);
            }
            pos = 0;
        }
        if (cur_row >= 0) {
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                const int q = q0 + a * 32 + lane;
                if (q < nv) __stcs(out + (size_t)cur_row * out_stride + q, acc[a]);
            }
        }
    }
}

// Zeros for the cells of a tile that receive nothing (the busy ones are written by pool_entries_wide)
template <typename V>
__device__ __forceinline__ void cta_zero_empty_cells(const int* __restrict__ ptr, V* __restrict__ out, int out_stride,
                                                     int nv, int rows, int warp, int lane) {
    int lo = 0, hi = 0;
    if (lane < rows) {
        lo = __ldg(ptr + lane);
        hi = __ldg(ptr + lane + 1);
    }
    const unsigned busy = __ballot_sync(kFull, hi > lo);
    const V z = vzero((V*)nullptr);
    for (int r = warp; r < rows; r += kWarps) {
        if ((busy >> r) & 1u) continue;
        V* orow = out + r * out_stride;
        for (int q = lane; q < nv; q += 32) __stcs(orow + q, z);
    }
}

// Sparse part: out[r*out_stride + q] = sum_k val[k] * src[idx[k]*src_stride + q], k in [ptr[r], ptr[r+1]).
// `src` and `out` already carry their channel offset.
template <typename V>
__device__ __forceinline__ void pool_tile(const V* __restrict__ src, int src_stride,
                                          const int* __restrict__ ptr, const int* __restrict__ idx,
                                          const float* __restrict__ val, V* __restrict__ out, int out_stride,
                                          int nv, int shift, int rows, int lane) {
    int lo = 0, hi = 0;
    if (lane < rows) {
        lo = __ldg(ptr + lane);
        hi = __ldg(ptr + lane + 1);
    }
    const unsigned busy = __ballot_sync(kFull, hi > lo);
    const int n = rows * nv;
    if (busy == 0u) {  // the common case: nothing projects into this tile
        const V z = vzero((V*)nullptr);
        for (int s0 = 0; s0 < n; s0 += 32 * kUnroll) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const int s = s0 + j * 32 + lane;
                if (s < n) {
                    int r, q;
                    split(s, nv, shift, r, q);
                    __stcs(out + r * out_stride + q, z);
                }
            }
        }
        return;
    }
    // narrow rows: nv lanes per cell, 32/nv cells side by side
    for (int s0 = 0; s0 < n; s0 += 32) {  // warp-uniform trip count
        const int s = s0 + lane;
        int r, q;
        split(s, nv, shift, r, q);
        int beg = __shfl_sync(kFull, lo, r & 31);
        int end = __shfl_sync(kFull, hi, r & 31);
        if (s >= n) end = beg;
        V acc = vzero((V*)nullptr);
        int k = beg;
        // four gathers in flight; the adds stay in ascending k
        for (; k + 4 <= end; k += 4) {
            int p[4];
            float w[4];
            V x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p[j] = __ldg(idx + k + j);
                w[j] = __ldg(val + k + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = __ldg(src + (size_t)p[j] * src_stride + q);
#pragma unroll
            for (int j = 0; j < 4; ++j) axpy(acc, w[j], x[j]);
        }
        for (; k < end; ++k) {
            const int p = __ldg(idx + k);
            const float w = __ldg(val + k);
            const V x = __ldg(src + (size_t)p * src_stride + q);
            axpy(acc, w, x);
        }
        if (s < n) __stcs(out + r * out_stride + q, acc);
    }
}

struct PoolArgs {
    const void* dense_in;   // forward: dst map            backward: g_fused
    const void* gather_in;  // forward: src map            backward: g_fused
    void* dense_out;        // forward: fused              backward: g_dst
    void* pool_out;         // forward: fused              backward: g_src
    const int* ptr;
    const int* key;         // destination cell of each entry (NULL: walk cells instead of entries)
    const int* idx;
    const float* val;
    int entry_ctas;         // leading CTAs that gather by entry (wide kernels, key != NULL)
    int n_dense;            // cells of the dense part (0 = skip)
    int n_pool;             // cells of the sparse part
    int vd, vs;             // vectors per cell: own channels, pooled channels
    int vd_shift, vs_shift; // log2 or -1
    int rows_dense, rows_pool;  // cells per warp tile
};

// Forward, narrow cells: a WARP owns a tile of `rows` destination cells; dense part then sparse part.
template <int W>
__global__ void __launch_bounds__(kThreads) shpl_forward_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int vf = a.vd + a.vs;
    const int rows_per = a.rows_pool;
    const int n_tiles = (a.n_pool + rows_per - 1) / rows_per;
    const V* dst = static_cast<const V*>(a.dense_in);
    const V* src = static_cast<const V*>(a.gather_in);
    V* fused = static_cast<V*>(a.pool_out);
    for (int t = blockIdx.x * kWarps + warp; t < n_tiles; t += gridDim.x * kWarps) {
        const int r0 = t * rows_per;
        const int rows = min(rows_per, a.n_pool - r0);
        V* out = fused + (size_t)r0 * vf;
        if (a.vd > 0) copy_tile<V>(dst + (size_t)r0 * a.vd, a.vd, out, vf, a.vd, a.vd_shift, rows, lane);
        pool_tile<V>(src, a.vs, a.ptr + r0, a.idx, a.val, out + a.vd, vf, a.vs, a.vs_shift, rows, lane);
    }
}

// Forward, wide cells: a CTA owns a tile of 32 destination cells (one tile per CTA: the hardware
// block scheduler balances tiles of very different cost).
template <int W, int ACC>
__global__ void __launch_bounds__(kThreads, 2) shpl_forward_wide_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int vf = a.vd + a.vs;
    const V* dst = static_cast<const V*>(a.dense_in);
    const V* src = static_cast<const V*>(a.gather_in);
    V* fused = static_cast<V*>(a.pool_out);
    if ((int)blockIdx.x < a.entry_ctas) {       // gather CTAs come first: they are the long pole
        const int e_begin = __ldg(a.ptr), e_end = __ldg(a.ptr + a.n_pool);
        const int e0 = e_begin + (blockIdx.x * kWarps + warp) * kEntryChunk;
        if (e0 >= e_end) return;
        pool_entries_wide<V, ACC>(src, a.vs, a.key, a.idx, a.val, e0, min(e0 + kEntryChunk, e_end), e_begin, e_end,
                                  fused + a.vd, vf, a.vs, lane);
        return;
    }
    const int r0 = (blockIdx.x - a.entry_ctas) * kWideTile;
    const int rows = min(kWideTile, a.n_pool - r0);
    V* out = fused + (size_t)r0 * vf;
    if (a.vd > 0) cta_copy_tile<V>(dst + (size_t)r0 * a.vd, a.vd, out, vf, a.vd, a.vd_shift, rows, warp, lane);
    if (a.key != nullptr) cta_zero_empty_cells<V>(a.ptr + r0, out + a.vd, vf, a.vs, rows, warp, lane);
    else cta_pool_tile_wide<V, ACC>(src, a.vs, a.ptr + r0, a.idx, a.val, out + a.vd, vf, a.vs, rows, warp, lane);
}

// Backward, narrow cells: warp tiles [0, tiles_dense) slice-copy g_fused[:, :C_d] -> g_dst; the rest
// gather g_fused[:, C_d:] rows through the transposed CSR into the dense g_src (zeros included).
template <int W>
__global__ void __launch_bounds__(kThreads) shpl_backward_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int vf = a.vd + a.vs;
    const int tiles_dense = a.n_dense > 0 ? (a.n_dense + a.rows_dense - 1) / a.rows_dense : 0;
    const int tiles_pool = (a.n_pool + a.rows_pool - 1) / a.rows_pool;
    const V* g_fused = static_cast<const V*>(a.dense_in);
    V* g_dst = static_cast<V*>(a.dense_out);
    V* g_src = static_cast<V*>(a.pool_out);
    for (int t = blockIdx.x * kWarps + warp; t < tiles_dense + tiles_pool; t += gridDim.x * kWarps) {
        if (t < tiles_dense) {
            const int r0 = t * a.rows_dense;
            const int rows = min(a.rows_dense, a.n_dense - r0);
            copy_tile<V>(g_fused + (size_t)r0 * vf, vf, g_dst + (size_t)r0 * a.vd, a.vd, a.vd, a.vd_shift, rows, lane);
        } else {
            const int p0 = (t - tiles_dense) * a.rows_pool;
            const int rows = min(a.rows_pool, a.n_pool - p0);
            pool_tile<V>(g_fused + a.vd, vf, a.ptr + p0, a.idx, a.val, g_src + (size_t)p0 * a.vs, a.vs, a.vs,
                         a.vs_shift, rows, lane);
        }
    }
}

// Backward, wide cells: CTA tiles of 32 cells; the first `tiles_pool` CTAs gather (they are the
// expensive ones and start first), the rest slice-copy.
template <int W, int ACC>
__global__ void __launch_bounds__(kThreads, 2) shpl_backward_wide_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int vf = a.vd + a.vs;
    const int tiles_pool = (a.n_pool + kWideTile - 1) / kWideTile;
    const V* g_fused = static_cast<const V*>(a.dense_in);
    V* g_dst = static_cast<V*>(a.dense_out);
    V* g_src = static_cast<V*>(a.pool_out);
    if ((int)blockIdx.x < a.entry_ctas) {
        const int e_begin = __ldg(a.ptr), e_end = __ldg(a.ptr + a.n_pool);
        const int e0 = e_begin + (blockIdx.x * kWarps + warp) * kEntryChunk;
        if (e0 >= e_end) return;
        pool_entries_wide<V, ACC>(g_fused + a.vd, vf, a.key, a.idx, a.val, e0, min(e0 + kEntryChunk, e_end), e_begin,
                                  e_end, g_src, a.vs, a.vs, lane);
        return;
    }
    const int t = blockIdx.x - a.entry_ctas;
    if (t < tiles_pool) {
        const int p0 = t * kWideTile;
        const int rows = min(kWideTile, a.n_pool - p0);
        if (a.key != nullptr) cta_zero_empty_cells<V>(a.ptr + p0, g_src + (size_t)p0 * a.vs, a.vs, a.vs, rows, warp, lane);
        else cta_pool_tile_wide<V, ACC>(g_fused + a.vd, vf, a.ptr + p0, a.idx, a.val, g_src + (size_t)p0 * a.vs, a.vs,
                                        a.vs, rows, warp, lane);
    } else {
        const int r0 = (t - tiles_pool) * kWideTile;
        const int rows = min(kWideTile, a.n_dense - r0);
        cta_copy_tile<V>(g_fused + (size_t)r0 * vf, vf, g_dst + (size_t)r0 * a.vd, a.vd, a.vd, a.vd_shift, rows, warp,
                         lane);
    }
}

int log2_or_neg(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

// cells per warp tile: about 1024 vectors of traffic per tile, at most 32 (one lane per cell)
int tile_rows(int vectors_per_cell) {
    const int budget = vectors_per_cell >= 64 ? 512 : 1024;   // wide cells are gathered a cell at a time: keep tiles short
    int r = 32;
    while (r > 1 && r * vectors_per_cell > budget) r >>= 1;
    return r;
}

int pick_width(int C_d, int C_s, std::initializer_list<const void*> ptrs) {
    int w = 4;
    while (w > 1 && ((C_d % w) || (C_s % w))) w >>= 1;
    for (const void* p : ptrs)
        while (w > 1 && p && !shpl::aligned(p, sizeof(float) * w)) w >>= 1;
    return w;
}

int grid_for(long long tiles) {
    const long long ctas = (tiles + kWarps - 1) / kWarps;
    const long long cap = (long long)shpl::sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

}  // namespace

extern "C" int shpl_pool_forward(const float* dst, const float* src, const int32_t* ptr, const int32_t* key,
                                 const int32_t* idx, const float* val, int32_t nnz_max, int32_t n_rows, int32_t C_d,
                                 int32_t n_src, int32_t C_s, float* fused, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_d >= 0 && C_s > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_forward: bad sizes n_rows=%d n_src=%d C_d=%d C_s=%d", n_rows, n_src, C_d, C_s);
    SHPL_REQUIRE(src && ptr && fused && (C_d == 0 || dst), SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward: null pointer");
    SHPL_REQUIRE(idx && val, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward: null idx/val");
    if (n_rows == 0) return SHPL_OK;
    const int w = pick_width(C_d, C_s, {dst, src, fused});
    PoolArgs a{};
    a.dense_in = dst;
    a.gather_in = src;
    a.dense_out = fused;
    a.pool_out = fused;
    a.ptr = ptr;
    a.key = (key && nnz_max > 0) ? key : nullptr;
    a.idx = idx;
    a.val = val;
    a.entry_ctas = a.key ? (nnz_max + kEntryChunk * kWarps - 1) / (kEntryChunk * kWarps) : 0;
    a.n_dense = n_rows;
    a.n_pool = n_rows;
    a.vd = C_d / w;
    a.vs = C_s / w;
    a.vd_shift = log2_or_neg(a.vd);
    a.vs_shift = log2_or_neg(a.vs);
    a.rows_pool = a.rows_dense = tile_rows((a.vd + a.vs) * w / 4 > 0 ? (a.vd + a.vs) * w / 4 : 1);
    const long long tiles = ((long long)n_rows + a.rows_pool - 1) / a.rows_pool;
    const int grid = grid_for(tiles);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a.vs >= 32) {   // wide cells: one CTA per tile of 32 cells
        const unsigned g = (unsigned)(a.entry_ctas + (n_rows + kWideTile - 1) / kWideTile);
        const bool one = a.vs <= 32;
        if (w == 4 && one) shpl_forward_wide_kernel<4, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 4) shpl_forward_wide_kernel<4, 2><<<g, kThreads, 0, s>>>(a);
        else if (w == 2 && one) shpl_forward_wide_kernel<2, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 2) shpl_forward_wide_kernel<2, 2><<<g, kThreads, 0, s>>>(a);
        else if (one) shpl_forward_wide_kernel<1, 1><<<g, kThreads, 0, s>>>(a);
        else shpl_forward_wide_kernel<1, 2><<<g, kThreads, 0, s>>>(a);
    } else if (w == 4) shpl_forward_kernel<4><<<grid, kThreads, 0, s>>>(a);
    else if (w == 2) shpl_forward_kernel<2><<<grid, kThreads, 0, s>>>(a);
    else shpl_forward_kernel<1><<<grid, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_forward_kernel");
}

extern "C" int shpl_pool_backward(const float* g_fused, const int32_t* ptrT, const int32_t* keyT, const int32_t* idxT,
                                  const float* valT, int32_t nnz_max, int32_t n_rows, int32_t C_d, int32_t n_src,
                                  int32_t C_s, float* g_dst, float* g_src, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_d >= 0 && C_s > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_backward: bad sizes n_rows=%d n_src=%d C_d=%d C_s=%d", n_rows, n_src, C_d, C_s);
    SHPL_REQUIRE(g_fused && ptrT && idxT && valT && g_src, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_backward: null pointer");
    if (n_rows == 0 && n_src == 0) return SHPL_OK;
    const bool dense = g_dst != nullptr && C_d > 0 && n_rows > 0;
    const int w = pick_width(C_d, C_s, {g_fused, g_dst, g_src});
    PoolArgs a{};
    a.dense_in = g_fused;
    a.gather_in = g_fused;
    a.dense_out = g_dst;
    a.pool_out = g_src;
    a.ptr = ptrT;
    a.key = (keyT && nnz_max > 0) ? keyT : nullptr;
    a.idx = idxT;
    a.val = valT;
    a.entry_ctas = a.key ? (nnz_max + kEntryChunk * kWarps - 1) / (kEntryChunk * kWarps) : 0;
    a.n_dense = dense ? n_rows : 0;
    a.n_pool = n_src;
    a.vd = C_d / w;
    a.vs = C_s / w;
    a.vd_shift = log2_or_neg(a.vd);
    a.vs_shift = log2_or_neg(a.vs);
    a.rows_dense = tile_rows(a.vd * w / 4 > 0 ? a.vd * w / 4 * 2 : 1);
    a.rows_pool = tile_rows(a.vs * w / 4 > 0 ? a.vs * w / 4 : 1);
    const long long tiles = (dense ? ((long long)n_rows + a.rows_dense - 1) / a.rows_dense : 0) +
                            ((long long)n_src + a.rows_pool - 1) / a.rows_pool;
    if (tiles == 0) return SHPL_OK;
    const int grid = grid_for(tiles);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a.vs >= 32) {
        const unsigned g = (unsigned)(a.entry_ctas + (n_src + kWideTile - 1) / kWideTile + (a.n_dense + kWideTile - 1) / kWideTile);
        const bool one = a.vs <= 32;
        if (w == 4 && one) shpl_backward_wide_kernel<4, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 4) shpl_backward_wide_kernel<4, 2><<<g, kThreads, 0, s>>>(a);
        else if (w == 2 && one) shpl_backward_wide_kernel<2, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 2) shpl_backward_wide_kernel<2, 2><<<g, kThreads, 0, s>>>(a);
        else if (one) shpl_backward_wide_kernel<1, 1><<<g, kThreads, 0, s>>>(a);
        else shpl_backward_wide_kernel<1, 2><<<g, kThreads, 0, s>>>(a);
    } else if (w == 4) shpl_backward_kernel<4><<<grid, kThreads, 0, s>>>(a);
    else if (w == 2) shpl_backward_kernel<2><<<grid, kThreads, 0, s>>>(a);
    else shpl_backward_kernel<1><<<grid, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_backward_kernel");
}
