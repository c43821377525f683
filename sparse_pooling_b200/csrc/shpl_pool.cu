// SHPL pooling kernels for sm_100a: forward gather-SpMM fused with the channel
// concat, and its deterministic (atomic-free) transpose-CSR backward.
//
// Replaces the TF graph ops of /root/reference/avod/avod/utils/sparse_pool_utils.py
//   :96-103  _sparse_pool_op        gather_nd -> sparse_tensor_dense_matmul -> reshape
//   :105-117 _sparse_pool_trans_op  sparse_transpose + matmul -> scatter_nd
//   :72,:87  tf.concat(axis=3)
// and the gradients TF autodiff derives for them.
//
// Every launch executes a short list of JOBS.  A job is "for each of n cells: a dense part
// (copy vd vectors) and/or a pooled part (sum over the cell's CSR entries of val * gathered
// row, vs vectors), written side by side (concat) or added together (AddN of two gradient
// paths)".  Forward of one direction = 1 job; backward of one direction = 2 jobs (slice copy
// over destination cells, gather over source cells); the dual-direction forms are the two
// directions' jobs in one launch, and the dual backward uses the add form so that no
// intermediate gradient is materialised.
//
// Kernels (dispatch in launch_jobs; DESIGN.md section 4.1 has the measurements behind every choice):
//   shpl_pool_sparse_kernel<W, kAdd, ACC>  the entry + stream kernel, used whenever the key arrays are there:
//       ENTRY CTAs walk the key-sorted entry list in chunks and own the busy cells (segmented sum, ascending k);
//       STREAM CTAs (a warp per tile of <= 32 cells, 8 x 128-bit ld/st.global.cs in flight per lane) copy the dense
//       parts and write the zeros of the cells that receive nothing -- one dependent load (the tile's CSR offsets,
//       fetched one tile ahead) away from a bare copy.
//         ACC = 1, few entries per cell (KITTI stride 1): plain entry walk, lanes = channel vectors
//         ACC = 1, many entries per cell, power-of-two vectors per cell <= 16: PACKED entry walk (lane groups gather
//                  different entries, products handed over one entry at a time)
//         ACC = 2, wide channel counts (C_s >= 128): lanes own two channel vectors each
//   shpl_pool_narrow_kernel<W, kAdd>       C_s < 128 without key arrays or with odd vector counts: gathers inside the
//       streaming warp, cells with more than 32 entries summed by the whole warp
//   shpl_pool_wide_kernel<W, ACC>          C_s >= 128 without key arrays: a warp per output cell, CTA-tiled dense copy
//   shpl_pool_heavy_exact_kernel<W>        listed cells (more than SHPL_HEAVY_LEN entries) up to SHPL_EXACT_LEN: a thread-block
//       cluster per cell gathers and multiplies in parallel, adder warps add in entry order (sequential sum, bit-exact)
//   shpl_pool_heavy_kernel<W, kGroups>     longer listed cells: a cluster per cell, fixed summation tree over distributed
//       shared memory
// Sums run in stored (ascending k) order with separately rounded multiply and add, which makes
// the result bit-identical to the sequential oracle.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "shpl_common.cuh"

namespace cg = cooperative_groups;

namespace {

#ifndef SHPL_NARROW_MIN_CTAS
#define SHPL_NARROW_MIN_CTAS 3
#endif
#ifndef SHPL_UNROLL
#define SHPL_UNROLL 8
#endif
#ifndef SHPL_WIDE_MIN_CTAS
#define SHPL_WIDE_MIN_CTAS 2
#endif
#ifndef SHPL_WIDE_GATHERS
#define SHPL_WIDE_GATHERS 8       // gathers in flight per warp in the entry CTAs of the wide kernel
#endif
#ifndef SHPL_SPARSE_MIN_CTAS
#define SHPL_SPARSE_MIN_CTAS 4
#endif
#ifndef SHPL_SPARSE_MIN_CTAS_WIDE
#define SHPL_SPARSE_MIN_CTAS_WIDE 3
#endif
#ifndef SHPL_SPARSE_GATHERS_WIDE
#define SHPL_SPARSE_GATHERS_WIDE 4
#endif
#ifndef SHPL_SPARSE_GATHERS
#define SHPL_SPARSE_GATHERS 4     // gathers in flight per warp in the entry CTAs of the sparse kernel
#endif
#ifndef SHPL_PIECE_LEN
#define SHPL_PIECE_LEN 256     // entries per piece of a split listed cell (a CTA sums one piece)
#endif
#ifndef SHPL_LD_POLICY
#define SHPL_LD_POLICY 1      // 0 = default, 1 = ld.global.cs (streaming)
#endif
#ifndef SHPL_ST_POLICY
#define SHPL_ST_POLICY 1      // 0 = default, 1 = st.global.cs (streaming)
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = SHPL_UNROLL;

// dense-stream accessors (the policy is a build-time experiment knob)
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) {
#if SHPL_LD_POLICY == 1
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename V> __device__ __forceinline__ void st_stream(V* p, const V& v) {
#if SHPL_ST_POLICY == 1
    __stcs(p, v);
#else
    *p = v;
#endif
}
// 256-bit forms (sm_100: LDG / STG .ENL2.256) for a lane that owns two adjacent float4 vectors of a row
__device__ __forceinline__ void ld_gather_pair(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ld_stream_pair(const float4* p, float4& a, float4& b) {
#if SHPL_LD_POLICY == 1
    asm volatile("ld.global.cs.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
#else
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
#endif
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_stream_pair(float4* p, const float4& a, const float4& b) {
#if SHPL_ST_POLICY == 1
    asm volatile("st.global.cs.v8.f32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
#else
    asm volatile("st.global.v8.f32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
#endif
                 ::"f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "l"(p) : "memory");
}
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxJobs = 4;
constexpr int kGatherUnroll = 8;
constexpr int kWideTile = 32;

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

__device__ __forceinline__ float4 vadd(const float4& a, const float4& b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float2 vadd(const float2& a, const float2& b) {
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__device__ __forceinline__ float vadd(const float& a, const float& b) { return __fadd_rn(a, b); }

// w * x with one rounding per component (the first half of axpy)
__device__ __forceinline__ float4 vscale(float w, const float4& x) {
    return make_float4(__fmul_rn(w, x.x), __fmul_rn(w, x.y), __fmul_rn(w, x.z), __fmul_rn(w, x.w));
}
__device__ __forceinline__ float2 vscale(float w, const float2& x) { return make_float2(__fmul_rn(w, x.x), __fmul_rn(w, x.y)); }
__device__ __forceinline__ float vscale(float w, const float& x) { return __fmul_rn(w, x); }
__device__ __forceinline__ float4 vshfl(const float4& v, int src) {
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                       __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}
__device__ __forceinline__ float2 vshfl(const float2& v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ float vshfl(const float& v, int src) { return __shfl_sync(0xffffffffu, v, src); }

struct Job {
    const void* dense_in;    // [n_cells, dense_in_stride] vectors; first vd of each cell are used
    const void* gather_in;   // rows gathered through idx, gather_stride vectors apart (channel offset applied)
    void* dense_out;         // concat form: where the dense part goes
    void* pool_out;          // where the pooled part goes (add form: dense + pooled)
    const int* ptr;          // [n_cells+1]
    const int* key;          // [nnz] cell of each entry, or NULL (then cells are walked one by one)
    const int* idx;
    const float* val;
    int dense_in_stride, dense_out_stride, gather_stride, pool_out_stride;
    int vd, vs;              // vectors per cell: dense part / pooled part (either may be 0)
    int vd_shift, vs_shift;  // log2 or -1
    int n_cells;
    int n_gather;            // rows of gather_in (debug build: every gathered index is checked against it)
    int add;                 // 1: pool_out[c] = dense_in[c] + sum (vd == vs); 0: concat form
    int heavy_len;           // > 0: cells with more entries are left to shpl_pool_heavy (treated as empty here)
    int rows_per_tile;       // narrow: cells per warp tile
    int entry_ctas;          // wide / sparse: leading CTAs of this job that gather by entry; narrow: CTAs serving the job
    int stream_ctas;         // sparse: CTAs of this job that stream the dense parts and zeros
    int packed;              // sparse: 1 = lane-group entry walk for few vectors per cell (SHPL_PACKED=0 switches it off)
    int long_len;            // sparse: cells with more entries are summed by the stream warps as a whole (kLongRow; 512 when packed)
    int staged;              // sparse, staged instantiation: 2 = every entry CTA takes the staged walk, 1 = only CTAs that meet a long cell
    int q_slices;            // sparse: warps a cell's channel vectors are spread over (1: a warp sums the whole row)
    int entry_chunk;         // wide: entries per warp
    int tiles;               // narrow: warp tiles; wide: CTA tiles of kWideTile cells
};

struct PoolArgs {
    Job job[kMaxJobs];
    int n_jobs;
    int begin[kMaxJobs + 1];   // narrow: first warp tile of each job; wide: first CTA of each job
};

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

// ------------------------------------------------------------------------------------ narrow
// Dense part by one warp: `rows` cells of `nv` vectors, in[r*in_stride + q] -> out[r*out_stride + q].
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
                v[j] = ld_stream(in + r * in_stride + q);
                o[j] = r * out_stride + q;
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            if (s < n) st_stream(out + o[j], v[j]);
        }
    }
}

constexpr int kLongRow = SHPL_LONG_LEN;      // narrow kernels: cells with more entries are summed by the whole warp
constexpr int kLongUnroll = 4;

// One long cell by the whole warp (narrow kernels).  The lanes form E = 32 / nv groups of nv lanes; group g
// gathers entry k0 + u*E + g, so E * kLongUnroll gathers of the cell are in flight at once instead of the 4 the
// cell's own nv lanes would have.  The products w*x are then handed to lanes 0..nv-1 by shuffles and added
// there in ascending k: the same roundings in the same order as the sequential walk, bit for bit.
template <typename V>
__device__ __forceinline__ V long_row_sum(const V* __restrict__ src, int src_stride, const int* __restrict__ idx,
                                          const float* __restrict__ val, int beg, int end, int nv, int lane, int n_gather) {
    const int E = 32 / nv;
    const int sub = lane / nv, q = lane - sub * nv;
    const bool active = sub < E;
    V acc = vzero((V*)nullptr);
    for (int k0 = beg; k0 < end; k0 += E * kLongUnroll) {
        V prod[kLongUnroll];
#pragma unroll
        for (int u = 0; u < kLongUnroll; ++u) {
            const int k = k0 + u * E + sub;
            prod[u] = vzero((V*)nullptr);
            if (active && k < end) {
                const int p = __ldg(idx + k);
                SHPL_DASSERT((unsigned)p < (unsigned)n_gather);
                const float w = __ldg(val + k);
                prod[u] = vscale(w, __ldg(src + (size_t)p * src_stride + q));
            }
        }
#pragma unroll
        for (int u = 0; u < kLongUnroll; ++u) {
            const int left = end - (k0 + u * E);              // warp-uniform
            for (int e = 0; e < E && e < left; ++e) acc = vadd(acc, vshfl(prod[u], e * nv + q));
        }
    }
    return acc;
}

// The long cells of a warp tile (bit mask `longs`), one after the other, each by the whole warp.  Out of line on
// purpose: the call sits on a rare path, and inlining it cost the streaming path registers (measured: +25 % time
// on the KITTI stride-1 forward).
template <typename V, bool kAdd>
__device__ __noinline__ void long_cells(const V* __restrict__ src, int src_stride, const int* __restrict__ ptr,
                                        const int* __restrict__ idx, const float* __restrict__ val, V* __restrict__ out,
                                        int out_stride, const V* __restrict__ addend, int add_stride, int nv,
                                        unsigned longs, int lane, int n_gather) {
    for (unsigned m = longs; m; m &= m - 1) {
        const int r = __ffs(m) - 1;
        const int beg = __ldg(ptr + r), end = __ldg(ptr + r + 1);
        SHPL_DASSERT(beg >= 0 && beg <= end);
        V acc = long_row_sum<V>(src, src_stride, idx, val, beg, end, nv, lane, n_gather);
        if (lane < nv) {
            if constexpr (kAdd) acc = vadd(ld_stream(addend + r * add_stride + lane), acc);
            st_stream(out + r * out_stride + lane, acc);
        }
    }
}

// Pooled part by one warp: out[r*out_stride + q] = (addend ? addend[r*add_stride + q] : 0) +
// sum_k val[k] * src[idx[k]*src_stride + q], k in [ptr[r], ptr[r+1]).  nv lanes per cell, 32/nv cells
// side by side.
template <typename V, bool kAdd>
__device__ __forceinline__ void pool_tile(const V* __restrict__ src, int src_stride, const int* __restrict__ ptr,
                                          const int* __restrict__ idx, const float* __restrict__ val,
                                          V* __restrict__ out, int out_stride, const V* __restrict__ addend,
                                          int add_stride, int nv, int shift, int rows, int heavy_len, int lane,
                                          int lo, int hi, int n_gather) {
    SHPL_DASSERT(lo >= 0 && lo <= hi);
    // lo, hi: this lane's cell offsets ptr[lane], ptr[lane+1] (0, 0 beyond `rows`), loaded by the caller
    // together with the dense loads so that the two latencies overlap
    // heavy cell: shpl_pool_heavy writes its pooled part (or dense + pooled in the add form); nothing is written for it
    // here, so that the heavy kernels may run CONCURRENTLY with this one on another stream
    const unsigned heavy_m = __ballot_sync(kFull, heavy_len > 0 && hi - lo > heavy_len);
    if (heavy_len > 0 && hi - lo > heavy_len) hi = lo;
    const unsigned busy = __ballot_sync(kFull, hi > lo) | heavy_m;
    // long cells: left empty by the lane-per-vector walk below, then summed by the whole warp
    const unsigned longs = __ballot_sync(kFull, hi - lo > kLongRow);
    if (hi - lo > kLongRow) hi = lo;
    const int n = rows * nv;
    if (busy == 0u) {  // the common case: nothing projects into this tile
        if constexpr (kAdd) {
            copy_tile<V>(addend, add_stride, out, out_stride, nv, shift, rows, lane);
            return;
        }
        const V z = vzero((V*)nullptr);
        for (int s0 = 0; s0 < n; s0 += 32 * kUnroll) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const int s = s0 + j * 32 + lane;
                if (s < n) {
                    int r, q;
                    split(s, nv, shift, r, q);
                    st_stream(out + r * out_stride + q, z);
                }
            }
        }
        return;
    }
    for (int s0 = 0; s0 < n; s0 += 32) {  // warp-uniform trip count
        const int s = s0 + lane;
        int r, q;
        split(s, nv, shift, r, q);
        int beg = __shfl_sync(kFull, lo, r & 31);
        int end = __shfl_sync(kFull, hi, r & 31);
        if (s >= n) end = beg;
        V base = vzero((V*)nullptr);
        const bool mine = s < n && !((heavy_m >> (r & 31)) & 1u);
        if constexpr (kAdd) {
            if (mine) base = ld_stream(addend + r * add_stride + q);
        }
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
            for (int j = 0; j < 4; ++j) {
                SHPL_DASSERT((unsigned)p[j] < (unsigned)n_gather);
                x[j] = __ldg(src + (size_t)p[j] * src_stride + q);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) axpy(acc, w[j], x[j]);
        }
        for (; k < end; ++k) {
            const int p = __ldg(idx + k);
            SHPL_DASSERT((unsigned)p < (unsigned)n_gather);
            const float w = __ldg(val + k);
            const V x = __ldg(src + (size_t)p * src_stride + q);
            axpy(acc, w, x);
        }
        if constexpr (kAdd) acc = vadd(base, acc);
        if (mine) st_stream(out + r * out_stride + q, acc);
    }
    if (longs != 0u) {
        __syncwarp();      // orders the stores above before the overwrites below (other lanes, same addresses)
        long_cells<V, kAdd>(src, src_stride, ptr, idx, val, out, out_stride, addend, add_stride, nv, longs, lane, n_gather);
    }
}

// job of this CTA: static indices only, so the job's fields stay in (uniform) registers instead of a
// local-memory copy of the parameter block
__device__ __forceinline__ Job select_job(const PoolArgs& a, int cta, int& first) {
    Job jb = a.job[0];
    first = a.begin[0];
#pragma unroll
    for (int i = 1; i < kMaxJobs; ++i)
        if (i < a.n_jobs && cta >= a.begin[i]) {
            jb = a.job[i];
            first = a.begin[i];
        }
    return jb;
}

// A CTA belongs to one job (begin[] partitions the grid); its warps stride over that job's tiles.
template <int W, bool kAdd>
__global__ void __launch_bounds__(kThreads, SHPL_NARROW_MIN_CTAS) shpl_pool_narrow_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int first;
    const Job jb = select_job(a, blockIdx.x, first);
    const int ctas = jb.entry_ctas;            // narrow: CTAs assigned to this job
    const V* dense_in = static_cast<const V*>(jb.dense_in);
    const V* gather_in = static_cast<const V*>(jb.gather_in);
    V* dense_out = static_cast<V*>(jb.dense_out);
    V* pool_out = static_cast<V*>(jb.pool_out);
    for (int t = (blockIdx.x - first) * kWarps + warp; t < jb.tiles; t += ctas * kWarps) {
        const int r0 = t * jb.rows_per_tile;
        const int rows = min(jb.rows_per_tile, jb.n_cells - r0);
        const V* din = dense_in + (size_t)r0 * jb.dense_in_stride;
        int lo = 0, hi = 0;
        if (jb.vs > 0 && lane < rows) {
            lo = __ldg(jb.ptr + r0 + lane);
            hi = __ldg(jb.ptr + r0 + lane + 1);
        }
        if constexpr (!kAdd) {
            if (jb.vd > 0)
                copy_tile<V>(din, jb.dense_in_stride, dense_out + (size_t)r0 * jb.dense_out_stride,
                             jb.dense_out_stride, jb.vd, jb.vd_shift, rows, lane);
        }
        if (jb.vs > 0)
            pool_tile<V, kAdd>(gather_in, jb.gather_stride, jb.ptr + r0, jb.idx, jb.val,
                               pool_out + (size_t)r0 * jb.pool_out_stride, jb.pool_out_stride, din,
                               jb.dense_in_stride, jb.vs, jb.vs_shift, rows, jb.heavy_len, lane, lo, hi, jb.n_gather);
    }
}

// -------------------------------------------------------------------------------------- wide
// Cell-serial gather (used when no key array is supplied): one cell by one warp.
template <typename V, int ACC>
__device__ __forceinline__ void pool_row_wide(const V* __restrict__ src, int src_stride, int beg, int end,
                                              const int* __restrict__ idx, const float* __restrict__ val,
                                              V* __restrict__ orow, const V* __restrict__ arow, int nv, int lane, int n_gather) {
    SHPL_DASSERT(beg >= 0 && beg <= end);
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
                    SHPL_DASSERT(e + j >= cnt || (unsigned)p < (unsigned)n_gather);
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
            if (q < nv) {
                if (arow != nullptr) acc[a] = vadd(ld_stream(arow + q), acc[a]);
                st_stream(orow + q, acc[a]);
            }
        }
    }
}

// Entry-parallel gather for FEW channel vectors per cell (nv a power of two <= 16): the lanes of a warp form
// G = 32 / nv groups and every group gathers a different entry, so kPackedUnroll * G gathers are in flight per warp
// instead of leaving 32 - nv lanes idle.  The products are then handed to lanes 0..nv-1 by shuffles ONE ENTRY AT A
// TIME, in ascending k, with a flush whenever the key changes: the same additions in the same order as the
// sequential walk (bit-identical).  The hand-over costs ~60 issue cycles per entry, so it pays where entries are many
// and rows short (the dense regime: 1 M pairs, C = 64: 201 -> 130 us); with a few entries per cell the plain walk of
// pool_entries_wide is faster, and long rows go to the whole-warp path.  Chunk ownership as in pool_entries_wide: a warp owns the cells whose first entry lies in its chunk.
constexpr int kPackedUnroll = 4;

template <typename V>
__device__ __forceinline__ void pool_entries_packed(const V* __restrict__ src, int src_stride,
                                                    const int* __restrict__ key, const int* __restrict__ idx,
                                                    const float* __restrict__ val, int e0, int e1, int e_begin,
                                                    int e_end, V* __restrict__ out, int out_stride,
                                                    const V* __restrict__ addend, int add_stride, int nv,
                                                    const int* __restrict__ ptr, int heavy_len, int lane, int n_gather, int n_cells) {
    const int G = 32 / nv;
    const int g = lane / nv, q = lane & (nv - 1);
    int base = e0;
    if (e0 > e_begin) {       // skip the entries that continue a cell begun in an earlier chunk
        const int prev_row = __ldg(key + e0 - 1);
        while (true) {
            const int k = base + lane;
            const bool fresh = k >= e_end || __ldg(key + k) != prev_row;
            const unsigned m = __ballot_sync(kFull, fresh);
            if (m) {
                base += __ffs(m) - 1;
                break;
            }
            base += 32;
            if (base >= e1) return;
        }
        if (base >= e1) return;          // no cell starts in this chunk
    }
    if (base >= e_end) return;
    int cur_row = -1, run_len = 0;
    V acc = vzero((V*)nullptr);
    bool done = false;
    while (!done) {
        V prod[kPackedUnroll];
        int row[kPackedUnroll];
#pragma unroll
        for (int u = 0; u < kPackedUnroll; ++u) {
            const int k = base + u * G + g;
            row[u] = -1;
            prod[u] = vzero((V*)nullptr);
            if (k < e_end && g < G) {
                row[u] = __ldg(key + k);
                const int p = __ldg(idx + k);
                SHPL_DASSERT((unsigned)p < (unsigned)n_gather && (unsigned)row[u] < (unsigned)n_cells);
                const float w = __ldg(val + k);
                prod[u] = vscale(w, __ldg(src + (size_t)p * src_stride + q));
            }
        }
        int next_base = base + kPackedUnroll * G;
        bool leave = false;
#pragma unroll
        for (int u = 0; u < kPackedUnroll; ++u) {
            if (done || leave) break;
            for (int gg = 0; gg < G; ++gg) {
                const int k = base + u * G + gg;
                if (k >= e_end) {
                    done = true;
                    break;
                }
                const int r = __shfl_sync(kFull, row[u], gg * nv);
                if (r != cur_row) {
                    if (cur_row >= 0 && lane < nv) {
                        V o = acc;
                        if (addend != nullptr) o = vadd(ld_stream(addend + (size_t)cur_row * add_stride + q), acc);
                        st_stream(out + (size_t)cur_row * out_stride + q, o);
                    }
                    cur_row = -1;
                    if (k >= e1) {           // the next cell belongs to a later warp
                        done = true;
                        break;
                    }
                    cur_row = r;
                    run_len = 0;
                    acc = vzero((V*)nullptr);
                }
                ++run_len;
                if (heavy_len > 0 && run_len > heavy_len) {
                    // a heavy cell: shpl_pool_heavy sums it; drop the partial sum and continue after the cell
                    const int cell_end = __ldg(ptr + cur_row + 1);
                    cur_row = -1;
                    if (cell_end >= e1 || cell_end >= e_end) done = true;
                    next_base = cell_end;
                    leave = true;            // leave both loops; the products already gathered are dropped
                    break;
                }
                acc = vadd(acc, vshfl(prod[u], gg * nv + q));
            }
        }
        base = next_base;
        if (base >= e_end) done = true;
    }
    if (cur_row >= 0 && lane < nv) {
        V o = acc;
        if (addend != nullptr) o = vadd(ld_stream(addend + (size_t)cur_row * add_stride + q), acc);
        st_stream(out + (size_t)cur_row * out_stride + q, o);
    }
}

// Entry-parallel gather: one warp takes `chunk` consecutive entries of the key-sorted entry list and
// owns every cell whose FIRST entry lies in that chunk (a cell is never split, so its sum keeps the
// ascending-k order; the warp reads on past the chunk until the cell ends, and skips leading entries
// that continue a cell begun in an earlier chunk).  Work per warp is `chunk` entries +- one cell,
// whatever the row-length skew.  Entries are streamed through a segmented sum: kGatherUnroll x ACC
// gathers in flight, the accumulator is flushed when the key changes.
template <typename V, int ACC, int GU = kGatherUnroll>
__device__ __forceinline__ void pool_entries_wide(const V* __restrict__ src, int src_stride,
                                                  const int* __restrict__ key, const int* __restrict__ idx,
                                                  const float* __restrict__ val, int e0, int e1, int e_begin,
                                                  int e_end, V* __restrict__ out, int out_stride,
                                                  const V* __restrict__ addend, int add_stride, int nv,
                                                  const int* __restrict__ ptr, int heavy_len, int lane, int n_gather, int n_cells,
                                                  int q_lo = 0, int q_hi = 0x7fffffff) {
    // [q_lo, q_hi): the channel vectors this warp sums (a slice of the cell when its row is spread over several warps)
    const int prev_row = (e0 > e_begin) ? __ldg(key + e0 - 1) : -1;
    // With two float4 accumulators per lane the lane owns an ADJACENT pair of vectors when every row it touches is 32-byte
    // aligned: one 256-bit request per gathered row, addend row and output row instead of two 128-bit ones (the channel a
    // sum belongs to does not change its entry order: bit-exact either way).
    constexpr bool kPairs = (ACC == 2 && sizeof(V) == 16);
    bool pairs = false;
    if constexpr (kPairs)
        pairs = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(addend)) & 31) == 0 &&
                ((src_stride | out_stride | (addend != nullptr ? add_stride : 0)) & 1) == 0;
    auto qof = [&](int q0, int a) { return pairs ? q0 + 2 * lane + a : q0 + a * 32 + lane; };
    auto flush = [&](int q0, int row, V (&acc)[ACC]) {      // acc (+ addend) -> out, for the lane's vectors of this block
        if constexpr (kPairs) {
            const int q = q0 + 2 * lane;
            if (pairs && q + 1 < nv) {
                float4 a0 = acc[0], a1 = acc[1];
                if (addend != nullptr) {
                    float4 t0, t1;
                    ld_stream_pair(reinterpret_cast<const float4*>(addend + (size_t)row * add_stride + q), t0, t1);
                    a0 = vadd(t0, a0);
                    a1 = vadd(t1, a1);
                }
                st_stream_pair(reinterpret_cast<float4*>(out + (size_t)row * out_stride + q), a0, a1);
                return;
            }
        }
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            const int q = qof(q0, a);
            if (q < nv) {
                V o = acc[a];
                if (addend != nullptr) o = vadd(ld_stream(addend + (size_t)row * add_stride + q), o);
                st_stream(out + (size_t)row * out_stride + q, o);
            }
        }
    };
    for (int q0 = q_lo; q0 < nv && q0 < q_hi; q0 += 32 * ACC) {
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
        int run_len = 0;               // entries of cur_row summed so far
        bool finished = false;
        V acc[ACC];
#pragma unroll
        for (int a = 0; a < ACC; ++a) acc[a] = vzero((V*)nullptr);
        while (true) {
            if (heavy_len > 0 && run_len > heavy_len) {
                // a heavy cell: shpl_pool_heavy sums it; drop the partial sum and jump to the cell's end
                const int cell_end = __ldg(ptr + cur_row + 1);
                cur_row = -1;
                run_len = 0;
#pragma unroll
                for (int a = 0; a < ACC; ++a) acc[a] = vzero((V*)nullptr);
                if (cell_end >= e1 || cell_end >= e_end) break;
                base = cell_end;
                my_row = -1;
                if (base + lane < e_end) {
                    my_row = __ldg(key + base + lane);
                    my_p = __ldg(idx + base + lane);
                    my_w = __ldg(val + base + lane);
                }
                pos = 0;
            }
            const int cnt = min(32, e_end - base);
            while (pos < cnt) {
                V x[GU][ACC];
                float w[GU];
                int row[GU];
#pragma unroll
                for (int j = 0; j < GU; ++j) {
                    const int ej = pos + j;
                    row[j] = __shfl_sync(kFull, my_row, ej & 31);
                    const int p = __shfl_sync(kFull, my_p, ej & 31);
                    SHPL_DASSERT(ej >= cnt || ((unsigned)p < (unsigned)n_gather && (unsigned)row[j] < (unsigned)n_cells));
                    w[j] = __shfl_sync(kFull, my_w, ej & 31);
                    const V* srow = src + (size_t)p * src_stride;
                    bool loaded = false;
                    if constexpr (kPairs) {
                        const int q = q0 + 2 * lane;
                        if (pairs && ej < cnt && q + 1 < nv) {
                            ld_gather_pair(reinterpret_cast<const float4*>(srow + q), x[j][0], x[j][1]);
                            loaded = true;
                        }
                    }
                    if (!loaded) {
#pragma unroll
                        for (int a = 0; a < ACC; ++a) {
                            const int q = qof(q0, a);
                            if (ej < cnt && q < nv) x[j][a] = __ldg(srow + q);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < GU; ++j) {
                    if (finished || pos + j >= cnt) continue;
                    if (row[j] != cur_row) {
                        if (cur_row >= 0) {
                            flush(q0, cur_row, acc);
#pragma unroll
                            for (int a = 0; a < ACC; ++a) acc[a] = vzero((V*)nullptr);
                        }
                        if (base + pos + j >= e1) {   // the next cell belongs to a later warp
                            finished = true;
                            cur_row = -1;
                            continue;
                        }
                        cur_row = row[j];
                        run_len = 0;
                    }
                    ++run_len;
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        const int q = qof(q0, a);
                        if (q < nv) axpy(acc[a], w[j], x[j][a]);
                    }
                }
                if (finished) break;
                pos += GU;
                if (heavy_len > 0 && run_len > heavy_len) break;
            }
            if (finished) break;
            if (heavy_len > 0 && run_len > heavy_len) continue;     // handled at the top of the loop
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
        if (cur_row >= 0 && !(heavy_len > 0 && run_len > heavy_len)) flush(q0, cur_row, acc);
    }
}

// CTA-cooperative dense copy of `rows` cells (flat over rows*nv vectors, kUnroll loads in flight per
// lane); cells whose bit is set in `skip` are left alone (the gather warps write them).
template <typename V>
__device__ __forceinline__ void cta_copy_tile(const V* __restrict__ in, int in_stride, V* __restrict__ out,
                                              int out_stride, int nv, int shift, int rows, unsigned skip, int warp,
                                              int lane) {
    const int n = rows * nv;
    for (int s0 = warp * 32 * kUnroll; s0 < n; s0 += kWarps * 32 * kUnroll) {
        V v[kUnroll];
        int o[kUnroll];
        bool on[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int s = s0 + j * 32 + lane;
            on[j] = false;
            if (s < n) {
                int r, q;
                split(s, nv, shift, r, q);
                on[j] = !((skip >> r) & 1u);
                if (on[j]) {
                    v[j] = ld_stream(in + r * in_stride + q);
                    o[j] = r * out_stride + q;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j)
            if (on[j]) st_stream(out + o[j], v[j]);
    }
}

template <int W, int ACC>
__global__ void __launch_bounds__(kThreads, SHPL_WIDE_MIN_CTAS) shpl_pool_wide_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int first;
    const Job jb = select_job(a, blockIdx.x, first);
    const int b = blockIdx.x - first;
    const V* din = static_cast<const V*>(jb.dense_in);
    const V* src = static_cast<const V*>(jb.gather_in);
    V* pout = static_cast<V*>(jb.pool_out);
    if (b < jb.entry_ctas) {       // gather CTAs come first: they are the long pole
        const int e_begin = __ldg(jb.ptr), e_end = __ldg(jb.ptr + jb.n_cells);
        const int e0 = e_begin + (b * kWarps + warp) * jb.entry_chunk;
        if (e0 >= e_end) return;
        pool_entries_wide<V, ACC, SHPL_WIDE_GATHERS>(src, jb.gather_stride, jb.key, jb.idx, jb.val, e0, min(e0 + jb.entry_chunk, e_end),
                                  e_begin, e_end, pout, jb.pool_out_stride, jb.add ? din : nullptr,
                                  jb.dense_in_stride, jb.vs, jb.ptr, jb.heavy_len, lane, jb.n_gather, jb.n_cells);
        return;
    }
    const int r0 = (b - jb.entry_ctas) * kWideTile;
    const int rows = min(kWideTile, jb.n_cells - r0);
    unsigned busy = 0u, heavy_m = 0u;
    int lo = 0, hi = 0;
    if (jb.vs > 0) {
        if (lane < rows) {
            lo = __ldg(jb.ptr + r0 + lane);
            hi = __ldg(jb.ptr + r0 + lane + 1);
        }
        // heavy cells: shpl_pool_heavy writes them (possibly concurrently, on another stream): counted busy so that
        // nothing is written for them here, but never summed here
        heavy_m = __ballot_sync(kFull, jb.heavy_len > 0 && hi - lo > jb.heavy_len);
        busy = __ballot_sync(kFull, hi > lo);
    }
    const bool by_entry = jb.key != nullptr;
    if (jb.vd > 0) {
        if (!jb.add)
            cta_copy_tile<V>(din + (size_t)r0 * jb.dense_in_stride, jb.dense_in_stride,
                             static_cast<V*>(jb.dense_out) + (size_t)r0 * jb.dense_out_stride, jb.dense_out_stride,
                             jb.vd, jb.vd_shift, rows, 0u, warp, lane);
        else   // add form: cells that receive nothing are a plain copy; the busy ones get dense + sum
            cta_copy_tile<V>(din + (size_t)r0 * jb.dense_in_stride, jb.dense_in_stride,
                             pout + (size_t)r0 * jb.pool_out_stride, jb.pool_out_stride, jb.vd, jb.vd_shift, rows,
                             busy, warp, lane);
    }
    if (jb.vs == 0) return;
    V* out = pout + (size_t)r0 * jb.pool_out_stride;
    if (!jb.add) {                 // zeros for the cells that receive nothing
        const V z = vzero((V*)nullptr);
        for (int r = warp; r < rows; r += kWarps) {
            if ((busy >> r) & 1u) continue;
            V* orow = out + r * jb.pool_out_stride;
            for (int q = lane; q < jb.vs; q += 32) st_stream(orow + q, z);
        }
    }
    if (by_entry) return;
    // no key array: busy cells dealt round-robin (by rank) to the warps of this CTA
    unsigned m = busy & ~heavy_m;
    int rank = 0;
    while (m) {
        const int r = __ffs(m) - 1;
        m &= m - 1;
        if ((rank++ % kWarps) != warp) continue;
        const int beg = __shfl_sync(kFull, lo, r);
        const int end = __shfl_sync(kFull, hi, r);
        pool_row_wide<V, ACC>(src, jb.gather_stride, beg, end, jb.idx, jb.val, out + r * jb.pool_out_stride,
                              jb.add ? din + (size_t)(r0 + r) * jb.dense_in_stride : nullptr, jb.vs, lane, jb.n_gather);
    }
}


// ------------------------------------------------------------------------------------ staged
// Entry walk of a whole CTA through shared memory, for many entries per cell (the dense regime, ground-plane and Zipf
// skew).  The serial part of an fp32 sum in a fixed order is only the chain of additions; everything else is parallel:
//   phase 0  a batch of B = min(512, 2048 / nv) consecutive entries (key, idx, val) -> shared memory, coalesced
//   phase A  thread -> (entry, channel vector): up to 8 gathers in flight per thread, the rounded products val * row
//            parked in shared memory in entry order
//   phase B  run heads of the batch (key changes) numbered by ballots + a prefix over the 8 warps
//   phase C  thread -> (run, channel vector): adds the run's products in stored order from shared memory (an LDS and
//            the 4-cycle add per entry instead of the ~60 issue cycles of a shuffle hand-over) and writes the cell;
//            the run left open at the end of the batch is carried (sum and length) into the next batch
// Same products, same additions, same order as the sequential walk: bit-identical.  Ownership as everywhere: the CTA owns
// the cells whose first entry lies in its chunk [E0, E1) and walks [end of the cell running into E0, end of the last
// owned cell).  Cells whose running length exceeds heavy_len are dropped at that point and skipped (the heavy kernels
// write them), so a 178 k-entry cell costs its owner heavy_len + B wasted entries.
constexpr int kStageVecs = 2048;          // product vectors per batch (32 KB of float4)
constexpr int kStageEntries = 512;        // entries per batch at most
constexpr int kStageMaxVecs = 64;         // channel vectors per cell the staged walk takes (batch >= 32 entries)
constexpr int kStageLongSplit = 64;       // long-run instantiation: runs of this many entries of a batch are added one thread per scalar

template <typename V> constexpr int stage_smem_bytes() {
    return (kStageVecs + 2 * kStageMaxVecs) * (int)sizeof(V) + (4 * kStageEntries + 2 + 32) * 4;
}

template <typename V, bool kAdd, bool kLongRuns>
__device__ __noinline__ void pool_entries_staged(const V* __restrict__ src, int src_stride, const int* __restrict__ key,
                                                 const int* __restrict__ idx, const float* __restrict__ val,
                                                 const int* __restrict__ ptr, int E0, int E1, int e_begin, int e_end,
                                                 V* __restrict__ out, int out_stride, const V* __restrict__ addend,
                                                 int add_stride, int nv, int shift, int heavy_len, int n_gather,
                                                 int n_cells, unsigned char* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    V* s_prod = reinterpret_cast<V*>(smem);
    V* s_carry = s_prod + kStageVecs;                                   // [2][kStageMaxVecs]
    int* s_key = reinterpret_cast<int*>(s_carry + 2 * kStageMaxVecs);   // [kStageEntries + 1]: + the key after the batch
    int* s_idx = s_key + kStageEntries + 1;
    float* s_val = reinterpret_cast<float*>(s_idx + kStageEntries);
    int* s_run = reinterpret_cast<int*>(s_val + kStageEntries);         // [kStageEntries + 1] run starts
    int* s_wtot = s_run + kStageEntries + 1;                            // [8] run heads per warp
    int* s_ctl = s_wtot + 8;                                            // [2][4]: carry_row, carry_len, next_pos, n_runs
    int* s_long = s_ctl + 8;                                            // [0]: count, [1..8]: long runs of the batch (kLongRuns)

    int start = E0;
    if (E0 > e_begin) start = __ldg(ptr + __ldg(key + E0 - 1) + 1);     // the end of the cell running into this chunk
    if (start >= E1) return;                                            // no cell starts here
    const int stop = E1 < e_end ? __ldg(ptr + __ldg(key + E1 - 1) + 1) : e_end;
    SHPL_DASSERT(start >= E0 && stop >= E1 && stop <= e_end);
    const int B = min(kStageEntries, kStageVecs / nv);
    int par = 0;
    if (tid == 0) {
        s_ctl[0] = -1;
        s_ctl[1] = 0;
    }
    int pos = start;
    while (pos < stop) {
        const int n = min(B, stop - pos);
        // ---- phase 0: the batch's entries
        for (int j = tid; j <= n; j += kThreads) {
            const int k = pos + j;
            if (j < n) {
                s_key[j] = __ldg(key + k);
                s_idx[j] = __ldg(idx + k);
                s_val[j] = __ldg(val + k);
            } else {
                s_key[n] = k < e_end ? __ldg(key + k) : -1;
            }
        }
        __syncthreads();
        // ---- phase A (loads): thread -> (entry, vector)
        const int total = n * nv;
        V x[8];
        int ej[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = u * kThreads + tid;
            ej[u] = -1;
            if (i < total) {
                const int j = shift >= 0 ? (i >> shift) : i / nv;
                const int q = i - j * nv;
                const int p = s_idx[j];
                SHPL_DASSERT((unsigned)p < (unsigned)n_gather && (unsigned)s_key[j] < (unsigned)n_cells);
                x[u] = __ldg(src + (size_t)p * src_stride + q);
                ej[u] = j;
            }
        }
        // ---- phase B: run heads (entries 2 tid, 2 tid + 1 of the batch)
        const int ja = 2 * tid, jb2 = 2 * tid + 1;
        const bool ha = ja < n && (ja == 0 || s_key[ja] != s_key[ja - 1]);
        const bool hb = jb2 < n && s_key[jb2] != s_key[jb2 - 1];
        const unsigned ba = __ballot_sync(kFull, ha), bb = __ballot_sync(kFull, hb);
        if (lane == 0) s_wtot[warp] = __popc(ba) + __popc(bb);
        __syncthreads();
        {
            int wbase = 0, all = 0;
#pragma unroll
            for (int w2 = 0; w2 < kWarps; ++w2) {
                const int c = s_wtot[w2];
                wbase += w2 < warp ? c : 0;
                all += c;
            }
            const unsigned lt = (1u << lane) - 1u;
            const int rank = wbase + __popc(ba & lt) + __popc(bb & lt);
            if (ha) s_run[rank] = ja;
            if (hb) s_run[rank + (ha ? 1 : 0)] = jb2;
            if (tid == 0) {
                s_run[all] = n;
                s_ctl[4 * par + 3] = all;
                if constexpr (kLongRuns) s_long[0] = 0;
            }
        }
        // ---- phase A (products, in entry order)
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (ej[u] >= 0) s_prod[u * kThreads + tid] = vscale(s_val[ej[u]], x[u]);
        __syncthreads();
        // ---- phase C: thread -> (run, vector); additions in stored order
        const int carry_row = s_ctl[4 * par], carry_len = s_ctl[4 * par + 1], n_runs = s_ctl[4 * par + 3];
        const V* carry_in = s_carry + par * kStageMaxVecs;
        V* carry_out = s_carry + (par ^ 1) * kStageMaxVecs;
        // (a thread-per-run mapping for four-vector cells -- the index work once per run -- was measured: it spills at the
        // 64 registers of the short-run instantiation and doubles the kernel's time, profiles/r2_staged_variants.txt)
        const int pairs = n_runs * nv;
        for (int i = tid; i < pairs; i += kThreads) {
            const int r = shift >= 0 ? (i >> shift) : i / nv;
            const int q = i - r * nv;
            const int j0 = s_run[r], j1 = s_run[r + 1];
            const int row = s_key[j0];
            const bool cont = r == 0 && row == carry_row;
            V acc = cont ? carry_in[q] : vzero((V*)nullptr);
            const int len = (cont ? carry_len : 0) + (j1 - j0);
            const V* pp = s_prod + j0 * nv + q;
            int j = j0;
            if constexpr (kLongRuns) {
                if (j1 - j0 >= kStageLongSplit) {          // summed below, one thread per scalar channel
                    if (q == 0) s_long[1 + atomicAdd(s_long, 1)] = r;
                    continue;
                }
                for (; j + 4 <= j1; j += 4, pp += 4 * nv) {
                    // four products loaded side by side, then added in order: the LDS latency is paid once per four entries
                    const V t0 = pp[0], t1 = pp[nv], t2 = pp[2 * nv], t3 = pp[3 * nv];
                    acc = vadd(vadd(vadd(vadd(acc, t0), t1), t2), t3);
                }
            }
            for (; j < j1; ++j, pp += nv) acc = vadd(acc, *pp);
            if (s_key[j1] != row) {                                     // the cell ends inside the batch
                if (!(heavy_len > 0 && len > heavy_len)) {
                    if constexpr (kAdd) acc = vadd(ld_stream(addend + (size_t)row * add_stride + q), acc);
                    st_stream(out + (size_t)row * out_stride + q, acc);
                }
            } else {
                carry_out[q] = acc;
            }
        }
        if constexpr (kLongRuns) {
            // Runs of kStageLongSplit entries or more: the chain of additions (4 cycles each) is their only serial part.  One
            // thread per SCALAR channel (four times the threads of the vector mapping, and eight registers of double buffer
            // instead of thirty-two): the next four products are loaded from shared memory while the current four are added,
            // so no LDS latency sits on the chain.  Same additions in the same order, component by component.
            __syncthreads();
            const int n_long = s_long[0];
            constexpr int WF = (int)(sizeof(V) / sizeof(float));
            const int ncs = nv * WF;                                    // scalar channels per cell
            const float* carry_in_f = reinterpret_cast<const float*>(carry_in);
            float* carry_out_f = reinterpret_cast<float*>(carry_out);
            for (int i = tid; i < n_long * ncs; i += kThreads) {
                const int li = i / ncs, c = i - li * ncs;
                const int r = s_long[1 + li];
                const int j0 = s_run[r], j1 = s_run[r + 1];
                const int row = s_key[j0];
                const bool cont = r == 0 && row == carry_row;
                float acc = cont ? carry_in_f[c] : 0.f;
                const int len = (cont ? carry_len : 0) + (j1 - j0);
                const float* pp = reinterpret_cast<const float*>(s_prod + j0 * nv) + c;
                int j = j0 + 4;                                         // j1 - j0 >= kStageLongSplit >= 4
                float a0 = pp[0], a1 = pp[ncs], a2 = pp[2 * ncs], a3 = pp[3 * ncs];
                pp += 4 * ncs;
                for (; j + 4 <= j1; j += 4, pp += 4 * ncs) {
                    const float b0 = pp[0], b1 = pp[ncs], b2 = pp[2 * ncs], b3 = pp[3 * ncs];
                    acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a0), a1), a2), a3);
                    a0 = b0; a1 = b1; a2 = b2; a3 = b3;
                }
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a0), a1), a2), a3);
                for (; j < j1; ++j, pp += ncs) acc = __fadd_rn(acc, *pp);
                if (s_key[j1] != row) {                                 // the cell ends inside the batch
                    if (!(heavy_len > 0 && len > heavy_len)) {
                        if constexpr (kAdd) acc = __fadd_rn(__ldcs(reinterpret_cast<const float*>(addend + (size_t)row * add_stride) + c), acc);
                        __stcs(reinterpret_cast<float*>(out + (size_t)row * out_stride) + c, acc);
                    }
                } else {
                    carry_out_f[c] = acc;
                }
            }
        }
        if (tid == 0) {
            const int j0 = s_run[n_runs - 1];
            const int row = s_key[j0];
            int nrow = -1, nlen = 0, npos = pos + n;
            if (s_key[n] == row) {                                      // open run: carried into the next batch
                const bool first_open = !(n_runs == 1 && row == carry_row);
                nlen = (first_open ? 0 : carry_len) + (n - j0);
                nrow = row;
                bool heavy = heavy_len > 0 && nlen > heavy_len;
                if constexpr (kLongRuns) {
                    // plans with listed cells: the first time a long run runs past a batch, one look at the cell's length --
                    // a heavy cell is skipped at once instead of after heavy_len + B gathered entries (five batches at C = 16)
                    if (!heavy && heavy_len > 0 && first_open && n - j0 >= kStageLongSplit)
                        heavy = __ldg(ptr + row + 1) - __ldg(ptr + row) > heavy_len;
                }
                if (heavy) {                                            // a heavy cell: dropped, skipped
                    nrow = -1;
                    npos = __ldg(ptr + row + 1);
                }
            }
            s_ctl[4 * (par ^ 1)] = nrow;
            s_ctl[4 * (par ^ 1) + 1] = nlen;
            s_ctl[4 * (par ^ 1) + 2] = npos;
        }
        __syncthreads();
        par ^= 1;
        pos = s_ctl[4 * par + 2];
    }
}

// ------------------------------------------------------------------------------------ sparse
// Narrow channel counts in the SPARSE regime (few entries per cell on average: KITTI / MV3D shapes, where ~2 % of
// the BEV cells receive anything).  shpl_pool_narrow_kernel walks a busy cell's entries inside the streaming warp:
// three dependent round trips (offsets -> entries -> gathered rows) that stall the stream whenever the CSR arrays
// and the source map are not in L2 -- measured 51 us against 41 us for a bare copy of the same bytes with a dirty
// L2 (tools/probe/pattern_probe.cu).  Here the roles are split like in the wide kernel: ENTRY CTAs gather by entry
// chunk and own the busy cells; STREAM CTAs copy the dense parts and write the zeros of the cells that receive
// nothing, one dependent load (the tile's offsets) away from a bare copy.  Same sums in the same order.
__host__ __device__ __forceinline__ bool packed_ok(const Job& jb) { return jb.vs_shift >= 0 && jb.vs <= 16 && jb.packed; }

// kStaged: 0 = the plain instantiation; 1 = staged walk tuned for short runs (the dense regime without listed cells: 64
// registers, 4 CTAs per SM, plain add loop); 2 = tuned for long runs (plans with listed cells: 80 registers, 3 CTAs per SM,
// the add loop unrolled by four).  Measured (profiles/r2_staged_variants.txt): 1 M uniform pairs C = 16 41.7 us with 1, 56.2 us
// with 2; 100 k Zipf pairs C = 16 66.9 us with 1, 41.4 us with 2.
template <int W, bool kAdd, int ACC, int kStaged = 0>
#ifndef SHPL_LONG_RUN_MIN_CTAS
#define SHPL_LONG_RUN_MIN_CTAS SHPL_SPARSE_MIN_CTAS_WIDE
#endif
__global__ void __launch_bounds__(kThreads, ACC == 1 ? (kStaged != 2 ? SHPL_SPARSE_MIN_CTAS : SHPL_LONG_RUN_MIN_CTAS) : SHPL_SPARSE_MIN_CTAS_WIDE) shpl_pool_sparse_kernel(PoolArgs a) {
    using V = typename VecOf<W>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int first;
    const Job jb = select_job(a, blockIdx.x, first);
    const int b = blockIdx.x - first;
    const V* din = static_cast<const V*>(jb.dense_in);
    V* pout = static_cast<V*>(jb.pool_out);
    // (entry and stream CTAs interleaved in grid order in the dense staged mode, so that the two halves overlap from the start,
    // were measured: 1 M uniform pairs C = 16 41.8 -> 52.2 us; entry CTAs LAST: 46.9 us. Entry CTAs first stays.)
    if (b < jb.entry_ctas) {       // gather CTAs first: they hold the dependent chains
        const int e_begin = __ldg(jb.ptr), e_end = __ldg(jb.ptr + jb.n_cells);
        if constexpr (kStaged) {
            // The staged instantiation (launched with dynamic shared memory): the CTA as a whole takes its chunk through
            // shared memory when the host asked for it (dense regime: jb.staged == 2) or when the chunk meets a cell of
            // more than kLongRow entries -- a run that long covers one of the sample points 32 entries apart or the
            // chunk's last entry; otherwise its warps walk their sub-chunks as in the plain instantiation (none of their
            // cells is long then).
            extern __shared__ __align__(16) unsigned char stage_smem[];
            const int cta_chunk = jb.entry_chunk * kWarps;
            const int E0 = e_begin + b * cta_chunk;
            if (E0 >= e_end) return;
            const int E1 = min(E0 + cta_chunk, e_end);
            bool staged = jb.staged == 2;
            if (!staged) {
                // keys 16 entries apart, from the chunk's first entry to 48 past its end: a cell of more than 32 entries
                // that starts in the chunk covers two neighbouring sample points (one load level; a false positive --
                // a cell of 17 ... 32 entries, or the clipped tail -- only means the staged walk takes a chunk it did not need to)
                bool hit = false;
                for (int sp = lane; 16 * sp < cta_chunk + 48; sp += 32) {
                    const int ka = E0 + 16 * sp, kb = ka + 16;
                    if (ka < e_end - 1) hit = hit || __ldg(jb.key + ka) == __ldg(jb.key + (kb < e_end ? kb : e_end - 1));
                }
                staged = __any_sync(kFull, hit);
            }
            if (staged) {
                pool_entries_staged<V, kAdd, kStaged == 2>(static_cast<const V*>(jb.gather_in), jb.gather_stride, jb.key, jb.idx, jb.val, jb.ptr,
                                             E0, E1, e_begin, e_end, pout, jb.pool_out_stride, kAdd ? din : nullptr,
                                             jb.dense_in_stride, jb.vs, jb.vs_shift, jb.heavy_len, jb.n_gather, jb.n_cells, stage_smem);
                return;
            }
        }
        // rows of more than 32 * ACC vectors (MV3D: 192) are spread over q_slices warps, each summing its own slice of
        // the channels over the same entry chunk: the dependent gather rounds of a crowded cell shrink by that factor
        // (Fewer entry CTAs striding over several groups of chunks, so that a grid a little over one wave -- layer A's dual
        // launch: 574 CTAs on 444 resident slots -- runs as one wave, were measured: forward 21.7 -> 29.9 us. Not adopted.)
        const int gw = b * kWarps + warp;
        const bool sliced = ACC == 2 && jb.q_slices > 1;              // (compile-time false in the narrow instantiations)
        const int chunk_id = sliced ? gw / jb.q_slices : gw;
        const int q_lo = sliced ? (gw - chunk_id * jb.q_slices) * 32 * ACC : 0;
        const int e0 = e_begin + chunk_id * jb.entry_chunk;
        if (e0 >= e_end) return;
        if (ACC == 1 && packed_ok(jb)) {      // few vectors per cell: lane groups gather different entries
            pool_entries_packed<V>(static_cast<const V*>(jb.gather_in), jb.gather_stride, jb.key, jb.idx, jb.val, e0,
                                   min(e0 + jb.entry_chunk, e_end), e_begin, e_end, pout, jb.pool_out_stride,
                                   kAdd ? din : nullptr, jb.dense_in_stride, jb.vs, jb.ptr, jb.long_len, lane, jb.n_gather, jb.n_cells);
            return;
        }
        // cells with more than kLongRow entries are left to the stream CTAs (whole-warp sum) or to shpl_pool_heavy:
        // the entry walk skips them exactly like it skips heavy cells
        // (with more than 32 vectors per cell there is no whole-warp path: the entry walk sums every length)
        pool_entries_wide<V, ACC, (ACC == 1 ? SHPL_SPARSE_GATHERS : SHPL_SPARSE_GATHERS_WIDE)>(static_cast<const V*>(jb.gather_in), jb.gather_stride, jb.key, jb.idx, jb.val, e0,
                                min(e0 + jb.entry_chunk, e_end), e_begin, e_end, pout, jb.pool_out_stride,
                                kAdd ? din : nullptr, jb.dense_in_stride, jb.vs, jb.ptr,
                                jb.vs <= 32 ? jb.long_len : jb.heavy_len, lane, jb.n_gather, jb.n_cells,
                                q_lo, sliced ? q_lo + 32 * ACC : 0x7fffffff);
        return;
    }
    const int stream_ctas = jb.stream_ctas;
    V* dout = static_cast<V*>(jb.dense_out);
    unsigned any_long = 0u;
    const int t_step = stream_ctas * kWarps;
    int t = (b - jb.entry_ctas) * kWarps + warp;
    // the CSR offsets of a tile are its only dependent load: they are fetched one tile ahead
    int lo_n = 0, hi_n = 0;
    if (jb.vs > 0 && t < jb.tiles && t * jb.rows_per_tile + lane < jb.n_cells && lane < jb.rows_per_tile) {
        lo_n = __ldg(jb.ptr + t * jb.rows_per_tile + lane);
        hi_n = __ldg(jb.ptr + t * jb.rows_per_tile + lane + 1);
    }
    for (; t < jb.tiles; t += t_step) {
        const int r0 = t * jb.rows_per_tile;
        const int rows = min(jb.rows_per_tile, jb.n_cells - r0);
        unsigned busy = 0u, longs = 0u;
        if (jb.vs > 0) {
            int lo = lo_n, hi = hi_n;
            lo_n = hi_n = 0;
            {
                const int tn = t + t_step;
                if (tn < jb.tiles && lane < jb.rows_per_tile && tn * jb.rows_per_tile + lane < jb.n_cells) {
                    lo_n = __ldg(jb.ptr + tn * jb.rows_per_tile + lane);
                    hi_n = __ldg(jb.ptr + tn * jb.rows_per_tile + lane + 1);
                }
            }
            // heavy cells (shpl_pool_heavy writes them, possibly concurrently on another stream) count as busy: nothing is
            // written for them here
            const bool heavy = lane < rows && jb.heavy_len > 0 && hi - lo > jb.heavy_len;
            busy = __ballot_sync(kFull, hi > lo);                          // cells somebody else writes (entry CTAs, heavy kernels) ...
            longs = (!kStaged && jb.vs <= 32) ? __ballot_sync(kFull, !heavy && hi - lo > jb.long_len) : 0u;   // ... or this warp sums as a whole, below
        }
        if (jb.vd > 0) {
            // concat form: the dense part of every cell; add form: a plain copy for the cells that receive nothing
            const V* in = din + (size_t)r0 * jb.dense_in_stride;
            V* out = kAdd ? pout + (size_t)r0 * jb.pool_out_stride : dout + (size_t)r0 * jb.dense_out_stride;
            const int ostride = kAdd ? jb.pool_out_stride : jb.dense_out_stride;
            const unsigned skip = kAdd ? busy : 0u;
            const int n = rows * jb.vd;
            for (int s0 = 0; s0 < n; s0 += 32 * kUnroll) {
                V v[kUnroll];
                int o[kUnroll];
#pragma unroll
                for (int j = 0; j < kUnroll; ++j) {
                    const int s = s0 + j * 32 + lane;
                    o[j] = -1;
                    if (s < n) {
                        int r, q;
                        split(s, jb.vd, jb.vd_shift, r, q);
                        if (!((skip >> r) & 1u)) {
                            v[j] = ld_stream(in + r * jb.dense_in_stride + q);
                            o[j] = r * ostride + q;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kUnroll; ++j)
                    if (o[j] >= 0) st_stream(out + o[j], v[j]);
            }
        }
        if (!kAdd && jb.vs > 0) {      // zeros for the cells that receive nothing
            V* out = pout + (size_t)r0 * jb.pool_out_stride;
            const V z = vzero((V*)nullptr);
            const int n = rows * jb.vs;
            for (int s0 = 0; s0 < n; s0 += 32 * kUnroll) {
#pragma unroll
                for (int j = 0; j < kUnroll; ++j) {
                    const int s = s0 + j * 32 + lane;
                    if (s < n) {
                        int r, q;
                        split(s, jb.vs, jb.vs_shift, r, q);
                        if (!((busy >> r) & 1u)) st_stream(out + r * jb.pool_out_stride + q, z);
                    }
                }
            }
        }
        any_long |= longs;
    }
    if (any_long == 0u) return;
    // Second walk, only for warps that met a long cell (none at KITTI / MV3D shapes): the whole-warp sum is kept out
    // of the streaming loop so that its registers and its call do not weigh on it.
    for (int t = (b - jb.entry_ctas) * kWarps + warp; t < jb.tiles; t += stream_ctas * kWarps) {
        const int r0 = t * jb.rows_per_tile;
        const int rows = min(jb.rows_per_tile, jb.n_cells - r0);
        int len = 0;
        if (lane < rows) len = __ldg(jb.ptr + r0 + lane + 1) - __ldg(jb.ptr + r0 + lane);
        if (jb.heavy_len > 0 && len > jb.heavy_len) len = 0;
        const unsigned longs = __ballot_sync(kFull, len > jb.long_len);
        if (longs != 0u)
            long_cells<V, kAdd>(static_cast<const V*>(jb.gather_in), jb.gather_stride, jb.ptr + r0, jb.idx, jb.val,
                                pout + (size_t)r0 * jb.pool_out_stride, jb.pool_out_stride,
                                kAdd ? din + (size_t)r0 * jb.dense_in_stride : nullptr, jb.dense_in_stride, jb.vs, longs, lane, jb.n_gather);
    }
}


// ------------------------------------------------------------------------------------- heavy
// One thread-block cluster (8 CTAs x 8 warps) per heavy cell: 64 contiguous pieces of the cell's entry
// range, each summed in stored order by one warp; warp sums added in order inside the CTA (shared
// memory), CTA sums added in order by CTA 0 through distributed shared memory.  A fixed tree.
constexpr int kClusterSize = 8;

struct HeavyArgs {
    const void* gather_in;
    const void* addend;
    void* out;
    const int* ptr;
    const int* idx;
    const float* val;
    const int* list;
    const int* count_dev;
    int list_cap;
    int gather_stride, addend_stride, out_stride;   // in vectors
    int nv;                                          // vectors per cell
    int skip_le;                                     // listed cells up to this many entries were summed by the main kernel
    int exact_len;                                   // listed cells up to this many entries: exact kernel; longer: tree kernel
    int eu;                                          // exact kernel: entries per warp per round
    int n_gather;                                    // rows of gather_in when the caller knows it (debug checks), else INT_MAX
};

// One warp sums the entries [pb, pe) of a cell in stored order into part[warp][0:nv] (shared memory).  kGroups (few
// channels, nv <= 16): lanes would idle if they mapped to channel vectors only, so E = 32 / nv lane groups take entries
// pb + g, pb + g + E, ... (kGatherUnroll gathers in flight each) and the group sums are then added in group order --
// still a fixed tree.
template <typename V, bool kGroups>
__device__ __forceinline__ void heavy_piece_sum(const HeavyArgs& a, const V* __restrict__ src, int pb, int pe, V* __restrict__ part,
                                                int warp, int lane) {
    if constexpr (kGroups) {
        // few channels (nv <= 16): lanes would idle if they mapped to channel vectors only.  E = 32 / nv lane groups
        // take entries pb + g, pb + g + E, ... (kGatherUnroll gathers in flight each); the group sums are
        // then added in group order -- still a fixed tree.
        const int nv = a.nv, E = 32 / nv;
        const int g = lane / nv, q = lane - g * nv;
        V acc = vzero((V*)nullptr);
        if (g < E) {
            for (int c = pb + g; c < pe; c += E * kGatherUnroll) {
                V x[kGatherUnroll];
                float w[kGatherUnroll];
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
                    const int k = c + j * E;
                    w[j] = 0.f;
                    x[j] = vzero((V*)nullptr);
                    if (k < pe) {
                        w[j] = __ldg(a.val + k);
                        SHPL_DASSERT((unsigned)__ldg(a.idx + k) < (unsigned)a.n_gather);
                        x[j] = __ldg(src + (size_t)__ldg(a.idx + k) * a.gather_stride + q);
                    }
                }
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j)
                    if (c + j * E < pe) axpy(acc, w[j], x[j]);
            }
        }
        V tot = vshfl(acc, q);                                  // group 0
        for (int e = 1; e < E; ++e) tot = vadd(tot, vshfl(acc, e * nv + q));
        if (lane < nv) part[warp * nv + lane] = tot;
    } else {
    for (int q0 = 0; q0 < a.nv; q0 += 64) {
        V acc[2];
        acc[0] = vzero((V*)nullptr);
        acc[1] = vzero((V*)nullptr);
        for (int c = pb; c < pe; c += 32) {
            int my_p = 0;
            float my_w = 0.f;
            if (c + lane < pe) {
                my_p = __ldg(a.idx + c + lane);
                my_w = __ldg(a.val + c + lane);
            }
            const int cnt = min(32, pe - c);
            for (int e = 0; e < cnt; e += kGatherUnroll) {
                V x[kGatherUnroll][2];
                float w[kGatherUnroll];
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
                    const int p = __shfl_sync(kFull, my_p, (e + j) & 31);
                    w[j] = __shfl_sync(kFull, my_w, (e + j) & 31);
                    const V* row = src + (size_t)p * a.gather_stride;
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int q = q0 + b * 32 + lane;
                        if (e + j < cnt && q < a.nv) x[j][b] = __ldg(row + q);
                    }
                }
#pragma unroll
                for (int j = 0; j < kGatherUnroll; ++j) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int q = q0 + b * 32 + lane;
                        if (e + j < cnt && q < a.nv) axpy(acc[b], w[j], x[j][b]);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int q = q0 + b * 32 + lane;
            if (q < a.nv) part[warp * a.nv + q] = acc[b];
        }
    }
    }
}

template <int W, bool kGroups>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kThreads) shpl_pool_heavy_kernel(HeavyArgs a) {
    using V = typename VecOf<W>::type;
    extern __shared__ float4 heavy_smem[];
    V* part = reinterpret_cast<V*>(heavy_smem);        // [kWarps][nv] warp sums
    V* cta_part = part + kWarps * a.nv;                // [nv] this CTA's sum
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_clusters = gridDim.x / kClusterSize;
    const int n_heavy = min(__ldg(a.count_dev), a.list_cap);
    const V* src = static_cast<const V*>(a.gather_in);
    const V* addend = static_cast<const V*>(a.addend);
    V* out = static_cast<V*>(a.out);
    for (int h = blockIdx.x / kClusterSize; h < n_heavy; h += n_clusters) {
        const int cell = __ldg(a.list + h);
        const int beg = __ldg(a.ptr + cell), end = __ldg(a.ptr + cell + 1);
        const long long L = end - beg;
        if (L <= a.exact_len || L <= a.skip_le) continue;   // the exact kernel's / the main kernel's
        const int piece = crank * kWarps + warp;
        const int pb = beg + (int)(L * piece / (kClusterSize * kWarps));
        const int pe = beg + (int)(L * (piece + 1) / (kClusterSize * kWarps));
        heavy_piece_sum<V, kGroups>(a, src, pb, pe, part, warp, lane);
        __syncthreads();
        for (int q = threadIdx.x; q < a.nv; q += kThreads) {
            V s = part[q];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s = vadd(s, part[w * a.nv + q]);
            cta_part[q] = s;
        }
        cluster.sync();
        if (crank == 0) {
            for (int q = threadIdx.x; q < a.nv; q += kThreads) {
                V s = cta_part[q];
                for (int r = 1; r < kClusterSize; ++r) s = vadd(s, *cluster.map_shared_rank(cta_part + q, r));
                if (addend != nullptr) s = vadd(addend[(size_t)cell * a.addend_stride + q], s);
                out[(size_t)cell * a.out_stride + q] = s;
            }
        }
        cluster.sync();      // the peers' shared memory stays valid until CTA 0 has read it
    }
}

// ---- Long listed cells split over MANY CTAs (the Zipf stress case has one cell with 178 000 entries: on a single
// cluster it kept 8 of 148 SMs busy for 600 us).  A cell of L > exact_len entries is cut into P = ceil(L / kPieceLen)
// contiguous pieces; a CTA sums one piece (its 8 warps take sub-pieces in stored order, the warp sums are added in
// order) into partial[piece][0:C] in the caller's workspace; shpl_pool_heavy_combine_kernel then adds the P partials in
// order.  A fixed tree that depends only on L: deterministic, within fp32 rounding of the sequential sum.
constexpr int kPieceLen = SHPL_PIECE_LEN;
constexpr int kMaxListed = 4096;           // listed cells whose piece counts fit the shared-memory prefix array

struct SplitArgs {
    int* prefix;                 // workspace: [kMaxListed + 1] exclusive prefix of the piece counts (written by CTA 0)
    void* partial;               // workspace: [total pieces][nv] vectors
    int max_pieces;              // capacity of `partial`
};

__device__ __forceinline__ int heavy_pieces(const HeavyArgs& a, int cell) {
    const long long L = (long long)__ldg(a.ptr + cell + 1) - __ldg(a.ptr + cell);
    if (L <= a.exact_len || L <= a.skip_le) return 0;          // the exact kernel's / the main kernel's
    return (int)((L + kPieceLen - 1) / kPieceLen);
}

// exclusive prefix of the piece counts of the listed cells into pre[0 .. n] (shared memory); all threads call it
__device__ __forceinline__ void heavy_prefix(const HeavyArgs& a, int n, int* pre, int* warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + kThreads - 1) / kThreads;              // contiguous chunk of cells per thread
    const int h0 = threadIdx.x * per, h1 = min(h0 + per, n);
    int sum = 0;
    for (int h = h0; h < h1; ++h) sum += heavy_pieces(a, __ldg(a.list + h));
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    int run = base + incl - sum;
    for (int h = h0; h < h1; ++h) {
        pre[h] = run;
        run += heavy_pieces(a, __ldg(a.list + h));
    }
    if (threadIdx.x == kThreads - 1) pre[n] = run;
    __syncthreads();
}

template <int W, bool kGroups>
__global__ void __launch_bounds__(kThreads) shpl_pool_heavy_split_kernel(HeavyArgs a, SplitArgs sa) {
    using V = typename VecOf<W>::type;
    extern __shared__ float4 heavy_smem[];
    V* part = reinterpret_cast<V*>(heavy_smem);                 // [kWarps][nv] warp sums
    __shared__ int pre[kMaxListed + 1];
    __shared__ int warp_tot[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_heavy = min(min(__ldg(a.count_dev), a.list_cap), kMaxListed);
    heavy_prefix(a, n_heavy, pre, warp_tot);
    if (blockIdx.x == 0)
        for (int h = threadIdx.x; h <= n_heavy; h += kThreads) sa.prefix[h] = pre[h];
    const int total = min(pre[n_heavy], sa.max_pieces);
    const V* src = static_cast<const V*>(a.gather_in);
    V* partial = static_cast<V*>(sa.partial);
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        int lo = 0, hi = n_heavy;                               // the cell h with pre[h] <= g < pre[h + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] <= g) lo = mid; else hi = mid;
        }
        const int h = lo, p = g - pre[h], P = pre[h + 1] - pre[h];
        const int cell = __ldg(a.list + h);
        const int beg = __ldg(a.ptr + cell);
        const long long L = (long long)__ldg(a.ptr + cell + 1) - beg;
        const int qb = beg + (int)(L * p / P), qe = beg + (int)(L * (p + 1) / P);      // this CTA's piece
        const long long Lp = qe - qb;
        const int pb = qb + (int)(Lp * warp / kWarps), pe = qb + (int)(Lp * (warp + 1) / kWarps);
        heavy_piece_sum<V, kGroups>(a, src, pb, pe, part, warp, lane);
        __syncthreads();
        for (int q = threadIdx.x; q < a.nv; q += kThreads) {
            V s = part[q];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s = vadd(s, part[w * a.nv + q]);
            SHPL_DASSERT(g < sa.max_pieces && pb >= beg && pe <= beg + (int)L);
            partial[(size_t)g * a.nv + q] = s;
        }
        __syncthreads();
    }
}

// A cell's P partial sums are added as a two-level fixed tree: kCombineGroups contiguous groups of ceil(P / groups)
// partials are summed in order side by side (thread = (group, channel vector), eight loads in flight), then the group sums
// in order.  Depends only on P: deterministic.
constexpr int kCombineGroups = 32;

template <int W>
__global__ void __launch_bounds__(kThreads) shpl_pool_heavy_combine_kernel(HeavyArgs a, SplitArgs sa) {
    using V = typename VecOf<W>::type;
    extern __shared__ float4 heavy_smem[];
    V* gsum = reinterpret_cast<V*>(heavy_smem);                 // [kCombineGroups][nv]
    const int n_heavy = min(min(__ldg(a.count_dev), a.list_cap), kMaxListed);
    const V* partial = static_cast<const V*>(sa.partial);
    const V* addend = static_cast<const V*>(a.addend);
    V* out = static_cast<V*>(a.out);
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int g0 = __ldg(sa.prefix + h), g1 = min(__ldg(sa.prefix + h + 1), sa.max_pieces);
        if (g1 <= g0) continue;                                  // block-uniform
        const int cell = __ldg(a.list + h);
        const int P = g1 - g0, per = (P + kCombineGroups - 1) / kCombineGroups;
        for (int i = threadIdx.x; i < kCombineGroups * a.nv; i += kThreads) {
            const int grp = i / a.nv, q = i - grp * a.nv;
            const int b = g0 + grp * per, e = min(b + per, g1);
            V s = vzero((V*)nullptr);
            if (b < e) {
                s = partial[(size_t)b * a.nv + q];
                int g = b + 1;
                for (; g + 8 <= e; g += 8) {
                    V t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = partial[(size_t)(g + j) * a.nv + q];
#pragma unroll
                    for (int j = 0; j < 8; ++j) s = vadd(s, t[j]);
                }
                for (; g < e; ++g) s = vadd(s, partial[(size_t)g * a.nv + q]);
            }
            gsum[i] = s;
        }
        __syncthreads();
        const int groups = (P + per - 1) / per;                  // non-empty groups
        for (int q = threadIdx.x; q < a.nv; q += kThreads) {
            V s = gsum[q];
            for (int grp = 1; grp < groups; ++grp) s = vadd(s, gsum[grp * a.nv + q]);
            if (addend != nullptr) s = vadd(addend[(size_t)cell * a.addend_stride + q], s);
            out[(size_t)cell * a.out_stride + q] = s;
        }
        __syncthreads();
    }
}

// Listed cells of up to SHPL_EXACT_LEN entries, in the REFERENCE'S order (bit-exact like the short cells): the
// serial part of a sequential fp32 sum is only the chain of additions (4 cycles each); the gathers and the products
// w*x are independent.  So the 64 warps of the cluster gather a round of entries in parallel and park the rounded
// products, in entry order, in the shared memory of CTA 0 (remote stores through distributed shared memory: fire and
// forget, no latency on anybody's critical path); then the adder warps of CTA 0 (one per 32 channel vectors) walk the
// round out of their own shared memory, adding product after product to the running sums.
// CS = CTAs per cluster: 8 when a single call serves few cells (all 64 warps gather for one cell), 2 when many cells are
// listed (four times as many cells in flight; the round length is set by the shared-memory budget either way).
template <int W, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kThreads) shpl_pool_heavy_exact_kernel(HeavyArgs a) {
    using V = typename VecOf<W>::type;
    extern __shared__ float4 heavy_smem[];
    V* stage = reinterpret_cast<V*>(heavy_smem);        // CTA 0's copy: [CS * kWarps * eu][nv] products of a round
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    V* stage0 = cluster.map_shared_rank(stage, 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = a.nv, eu = a.eu;
    const int E = nv >= 32 ? 1 : 32 / nv;               // entries one warp-wide load covers
    const int sub = nv >= 32 ? 0 : lane / nv;           // which of them this lane works on
    const int q0 = nv >= 32 ? lane : lane - sub * nv;   // first channel vector of this lane
    const bool lane_on = sub < E;
    const int U = eu / E;                               // warp-wide loads per round (eu is a multiple of E, eu <= 32)
    const int per_round = CS * kWarps * eu;
    const int n_clusters = gridDim.x / CS;
    const int n_heavy = min(__ldg(a.count_dev), a.list_cap);
    const V* src = static_cast<const V*>(a.gather_in);
    const V* addend = static_cast<const V*>(a.addend);
    V* out = static_cast<V*>(a.out);
    const int aq = warp * 32 + lane;                    // adder warps: channel vector of this lane
    const bool adder = crank == 0 && aq < nv;
    constexpr int kG = 4;                               // warp-wide loads in flight
    constexpr int kA = 8;                               // products an adder lane fetches ahead of its additions
    for (int h = blockIdx.x / CS; h < n_heavy; h += n_clusters) {
        const int cell = __ldg(a.list + h);
        const int beg = __ldg(a.ptr + cell), end = __ldg(a.ptr + cell + 1);
        if (end - beg > a.exact_len || end - beg <= a.skip_le) continue;   // the tree kernel's / the main kernel's (the whole cluster agrees)
        V acc = vzero((V*)nullptr);
        for (int rb = beg; rb < end; rb += per_round) {
            // ---- gather: warp (crank, warp) owns entries [wb, wb + eu) of the round
            const int w_first = (crank * kWarps + warp) * eu;
            const int wb = rb + w_first;
            int my_p = 0;
            float my_w = 0.f;
            if (lane < eu && wb + lane < end) {
                my_p = __ldg(a.idx + wb + lane);
                my_w = __ldg(a.val + wb + lane);
            }
            V* mine = stage0 + (size_t)w_first * nv;
            for (int u0 = 0; u0 < U; u0 += kG) {
#pragma unroll
                for (int j = 0; j < kG; ++j) {
                    const int slot = (u0 + j) * E + sub;
                    const int p = __shfl_sync(kFull, my_p, slot & 31);
                    const float w = __shfl_sync(kFull, my_w, slot & 31);
                    if (lane_on && u0 + j < U && wb + slot < end) {
                        const V* row = src + (size_t)p * a.gather_stride;
                        for (int q = q0; q < nv; q += 32) mine[slot * nv + q] = vscale(w, __ldg(row + q));
                    }
                }
            }
            cluster.sync();
            // ---- add, in entry order, out of CTA 0's own shared memory
            if (adder) {
                const int n_round = min(per_round, end - rb);
                const V* base = stage + aq;
                int s0 = 0;
                if (n_round >= kA) {
                    // full batches, no predicates on the chain: batch b+1 is fetched while batch b is added
                    V x[kA];
#pragma unroll
                    for (int j = 0; j < kA; ++j) x[j] = base[(size_t)j * nv];
                    for (; s0 + 2 * kA <= n_round; s0 += kA) {
                        V y[kA];
#pragma unroll
                        for (int j = 0; j < kA; ++j) y[j] = base[(size_t)(s0 + kA + j) * nv];
#pragma unroll
                        for (int j = 0; j < kA; ++j) acc = vadd(acc, x[j]);
#pragma unroll
                        for (int j = 0; j < kA; ++j) x[j] = y[j];
                    }
#pragma unroll
                    for (int j = 0; j < kA; ++j) acc = vadd(acc, x[j]);
                    s0 += kA;
                }
                for (; s0 < n_round; ++s0) acc = vadd(acc, base[(size_t)s0 * nv]);
            }
            cluster.sync();      // the round's products have been consumed: the buffer may be refilled
        }
        if (adder) {
            if (addend != nullptr) acc = vadd(addend[(size_t)cell * a.addend_stride + aq], acc);
            out[(size_t)cell * a.out_stride + aq] = acc;
        }
    }
}

// --------------------------------------------------------------------------------------- host
// Tuning constants.  The product library is stateless: every value below is a compile-time constant (the measured
// optimum).  Only a build with -DSHPL_EXPERIMENT (make exp -> libshpl_exp.so, used by tools/ for A/B
// runs through SHPL_LIB) reads the environment variable of the same name instead, once per process.
#ifdef SHPL_EXPERIMENT
#define SHPL_KNOB(name, dflt)                                   \
    ([]() -> int {                                              \
        static const int v = []() {                             \
            const char* e = getenv(name);                       \
            return e ? atoi(e) : (dflt);                        \
        }();                                                    \
        return v;                                               \
    }())
#else
#define SHPL_KNOB(name, dflt) (dflt)
#endif

// Listed cells of up to SHPL_EXACT_LEN entries stay in the MAIN kernels for C <= 128 (at C = 256 a batch of the staged walk
// holds 32 entries and a 2048-entry cell would take 64 of them on one CTA): the staged entry walk sums them in
// entry order on one CTA (bit-exact, like the exact cluster kernel, without its extra launches: the exact kernel was
// 91 us of the 1 M Zipf pairs forward at C = 64 on the side stream's critical path); the heavy entry points skip them
// (HeavyArgs::skip_le) and only split the longer ones.  The rule is a function of C alone so that both sides agree.
// SHPL_MAIN_KEEP=0 (experiment builds) brings the round-1 rule back (everything above heavy_len goes to the heavy kernels).
constexpr int kMainKeep = SHPL_EXACT_LEN, kMainKeepMaxC = 128;
int main_kernel_heavy_len(int heavy_len, int c_pool) {
    const int keep = SHPL_KNOB("SHPL_MAIN_KEEP", kMainKeep);
    return heavy_len > 0 && c_pool <= kMainKeepMaxC && heavy_len < keep ? keep : heavy_len;
}

int log2_or_neg(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

// cells per warp tile (narrow): about 1024 float4 of traffic per tile, at most 32 (one lane per cell)
int tile_rows(int float4_per_cell) {
    int r = 32;
    while (r > 1 && r * float4_per_cell > 1024) r >>= 1;
    return r;
}

int pick_width(std::initializer_list<int> channels, std::initializer_list<const void*> ptrs) {
    int w = 4;
    for (int c : channels)
        while (w > 1 && (c % w)) w >>= 1;
    for (const void* p : ptrs)
        while (w > 1 && p && !shpl::aligned(p, sizeof(float) * w)) w >>= 1;
    return w;
}

struct JobSpec {            // in floats / cells, before the vector width is chosen
    const float* dense_in = nullptr;
    int dense_in_stride = 0;
    int c_dense = 0;
    const float* gather_in = nullptr;
    int gather_stride = 0;
    int c_pool = 0;
    float* dense_out = nullptr;
    int dense_out_stride = 0;
    float* pool_out = nullptr;
    int pool_out_stride = 0;
    const int* ptr = nullptr;
    const int* key = nullptr;
    const int* idx = nullptr;
    const float* val = nullptr;
    int nnz_max = 0;
    int n_cells = 0;
    int n_gather = 0x7fffffff;   // rows of gather_in; entry points set it (debug build checks gathered indices against it)
    int add = 0;
    int heavy_len = 0;
};

// grid of the narrow kernel: 8 CTAs per SM although only 3 are resident (see launch_jobs)
int narrow_ctas_per_sm() { const int v = SHPL_KNOB("SHPL_NARROW_CTAS_PER_SM", 8); return v > 0 ? v : 8; }
// entries per warp in the entry CTAs (0: the per-width default)
int entry_chunk_knob() { const int v = SHPL_KNOB("SHPL_ENTRY_CHUNK", 0); return v > 0 ? v : 0; }
int packed_knob() { return SHPL_KNOB("SHPL_PACKED", 1); }
// entries per warp with the packed entry walk
int packed_chunk_knob() { const int v = SHPL_KNOB("SHPL_PACKED_CHUNK", 64); return v > 0 ? v : 64; }
int stream_ctas_per_sm() { const int v = SHPL_KNOB("SHPL_STREAM_CTAS_PER_SM", 6); return v > 0 ? v : 6; }
// 0: the CTA-tiled wide kernel; 1: entry + stream kernel up to 64 vectors per cell; 2: always
int wide_stream_knob() { return SHPL_KNOB("SHPL_WIDE_STREAM", 2); }
// staged entry walk: 0 = never (packed / plain walks as in round 1), 1 = dense regime + long cells, 2 = whenever eligible
int staged_knob() { return SHPL_KNOB("SHPL_STAGED", 1); }
// batches per entry CTA in the dense regime
// staged = 2 (every entry CTA) up to this many vectors per cell, when entries * density > cells
// -1: by the caller's heavy_len; 1 / 2: force the short-run / long-run staged instantiation (experiments)
int staged_variant_knob() { return SHPL_KNOB("SHPL_STAGED_VARIANT", -1); }
int q_slices_knob() { return SHPL_KNOB("SHPL_Q_SLICES", 1); }
int staged_all_vecs_knob() { return SHPL_KNOB("SHPL_STAGED_ALL_VECS", 8); }
// measured (profiles/r2_staged_ab.txt): 100 k pairs on 560 k cells gain 10 % with every entry CTA staging, 20 k pairs lose 8 %
int staged_density_knob() { const int v = SHPL_KNOB("SHPL_STAGED_DENSITY", 16); return v > 0 ? v : 16; }
int staged_batches_knob() { const int v = SHPL_KNOB("SHPL_STAGED_BATCHES", 4); return v > 0 ? v : 4; }

// Sparse regime of the narrow channel counts: every pooled job comes with its key array and the entries are few
// next to the cells (KITTI stride 1: 20 k entries for 560 k cells).
bool sparse_regime(const PoolArgs& a, const JobSpec* const* spec) {
    const int mode = SHPL_KNOB("SHPL_SPARSE", 1);      // 0: never, 1: by density, 2: always (experiments)
    if (!mode) return false;
    long long nnz = 0, cells = 0;
    for (int i = 0; i < a.n_jobs; ++i) {
        const Job& o = a.job[i];
        if (o.vs > 0) {
            if (o.key == nullptr) return false;
            nnz += spec[i]->nnz_max;
        }
        cells += o.n_cells;
    }
    return mode > 1 || nnz * 4 <= cells;
}

int launch_jobs(const JobSpec* specs, int n_specs, cudaStream_t s, const char* who) {
    int w = 4, max_vs = 0;
    for (int i = 0; i < n_specs; ++i) {
        const JobSpec& j = specs[i];
        if (j.n_cells <= 0) continue;
        const int wi = pick_width({j.c_dense, j.c_pool, j.dense_in_stride, j.dense_out_stride, j.gather_stride,
                                   j.pool_out_stride},
                                  {j.dense_in, j.gather_in, j.dense_out, j.pool_out});
        w = wi < w ? wi : w;
    }
    PoolArgs a{};
    const JobSpec* src_spec[kMaxJobs];
    a.n_jobs = 0;
    for (int i = 0; i < n_specs && a.n_jobs < kMaxJobs; ++i) {
        const JobSpec& j = specs[i];
        if (j.n_cells <= 0) continue;
        src_spec[a.n_jobs] = &j;
        Job& o = a.job[a.n_jobs++];
        o.dense_in = j.dense_in;
        o.gather_in = j.gather_in;
        o.dense_out = j.dense_out;
        o.pool_out = j.pool_out;
        o.ptr = j.ptr;
        o.key = (j.key && j.nnz_max > 0) ? j.key : nullptr;
        o.idx = j.idx;
        o.val = j.val;
        o.dense_in_stride = j.dense_in_stride / w;
        o.dense_out_stride = j.dense_out_stride / w;
        o.gather_stride = j.gather_stride / w;
        o.pool_out_stride = j.pool_out_stride / w;
        o.vd = j.c_dense / w;
        o.vs = j.c_pool / w;
        o.vd_shift = log2_or_neg(o.vd);
        o.vs_shift = log2_or_neg(o.vs);
        o.n_cells = j.n_cells;
        o.n_gather = j.n_gather;
        o.add = j.add;
        o.heavy_len = main_kernel_heavy_len(j.heavy_len, j.c_pool);
        max_vs = o.vs > max_vs ? o.vs : max_vs;
    }
    if (a.n_jobs == 0) return SHPL_OK;
    bool wide = max_vs >= 32;
    bool stream_split = !wide && sparse_regime(a, src_spec);
    // The staged entry walk (pool_entries_staged): every pooled job has its key array and at most kStageMaxVecs vectors per
    // cell.  Dense regime (many entries per cell on average): every entry CTA takes it (staged = 2).  Otherwise, when the
    // caller handles heavy cells (heavy_len > 0: the drop-in layer, the sweep), the staged INSTANTIATION is launched and
    // an entry CTA takes the staged walk only if its chunk meets a cell of more than 32 entries (staged = 1); callers that
    // pass heavy_len = 0 (FramePipeline, the custom op: KITTI / MV3D shapes) keep the plain instantiation.
    int staged = 0;
    bool listed = false;       // a caller's heavy_len <= SHPL_HEAVY_LEN says its plan has listed cells (see shpl.h): long runs
    for (int i = 0; i < n_specs; ++i) listed = listed || (specs[i].heavy_len > 0 && specs[i].heavy_len <= SHPL_HEAVY_LEN);
    if (staged_variant_knob() >= 0) listed = staged_variant_knob() == 2;
    {
        const int mode = staged_knob();
        bool keys = true, heavy_all = true;
        long long nnz = 0, cells = 0;
        for (int i = 0; i < a.n_jobs; ++i) {
            const Job& o = a.job[i];
            if (o.vs > 0) {
                keys = keys && o.key != nullptr && o.vs <= kStageMaxVecs;
                heavy_all = heavy_all && o.heavy_len > 0;
                nnz += src_spec[i]->nnz_max;
            }
            cells += o.n_cells;
        }
        if (mode > 0 && keys && nnz > 0) {
            // every entry CTA stages when the cells are narrow (<= 8 vectors: the plain walk leaves 24+ lanes idle per
            // entry) and the entries many; wider cells only where a long cell is met, and only for callers that say cells
            // may be long (heavy_len > 0)
            if (mode > 1 || (max_vs <= staged_all_vecs_knob() && nnz * staged_density_knob() > cells)) staged = 2;
            else if (heavy_all) staged = 1;
        }
        if (staged) {
            wide = false;
            stream_split = true;
        }
    }
    bool packed = false;
    if (staged != 2 && !wide && (staged == 1 ? !sparse_regime(a, src_spec) : !stream_split) && packed_knob()) {
        // dense regime, the CTAs that do not stage: the entry + stream kernel with the PACKED entry walk, when every
        // pooled job has its key array and a power-of-two number of vectors per cell <= 16
        bool ok = true;
        for (int i = 0; i < a.n_jobs; ++i) {
            const Job& o = a.job[i];
            if (o.vs > 0 && (o.key == nullptr || o.vs_shift < 0 || o.vs > 16)) ok = false;
        }
        packed = ok;
        stream_split = stream_split || ok;
    }
    // Wide jobs take the entry + stream kernel too whenever their key arrays are there (measured on B200: full scan
    // C = 128 forward 218 -> 165 us, RetinaNet P2 24.0 -> 20.4 us, the bench step 142 -> 131 us); the CTA-tiled
    // shpl_pool_wide_kernel stays as the path without key arrays.
    if (wide && wide_stream_knob() && (max_vs <= 64 || wide_stream_knob() > 1)) {
        bool keys = true;
        for (int i = 0; i < a.n_jobs; ++i) keys = keys && (a.job[i].vs == 0 || a.job[i].key != nullptr);
        if (keys) {
            wide = false;
            stream_split = true;
        }
    }
    a.begin[0] = 0;
    for (int i = 0; i < a.n_jobs; ++i) {
        Job& o = a.job[i];
        const int nnz_max = src_spec[i]->nnz_max;
        if (wide) {
            o.entry_chunk = nnz_max > (1 << 18) ? 32 : 8;   // small problems are latency-bound: more, shorter chains
            o.entry_ctas = (o.key && o.vs > 0) ? (nnz_max + o.entry_chunk * kWarps - 1) / (o.entry_chunk * kWarps) : 0;
            o.tiles = (o.n_cells + kWideTile - 1) / kWideTile;
            a.begin[i + 1] = a.begin[i] + o.entry_ctas + o.tiles;
        } else {
            const int f4 = ((o.add ? o.vs : o.vd + o.vs) * w + 3) / 4;
            o.rows_per_tile = tile_rows(f4 > 0 ? f4 : 1);
            o.tiles = (o.n_cells + o.rows_per_tile - 1) / o.rows_per_tile;
            a.begin[i + 1] = a.begin[i] + o.tiles;
        }
    }
    if (wide) {
        const unsigned g = (unsigned)a.begin[a.n_jobs];
        const bool one = max_vs <= 32;
        if (w == 4 && one) shpl_pool_wide_kernel<4, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 4) shpl_pool_wide_kernel<4, 2><<<g, kThreads, 0, s>>>(a);
        else if (w == 2 && one) shpl_pool_wide_kernel<2, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 2) shpl_pool_wide_kernel<2, 2><<<g, kThreads, 0, s>>>(a);
        else if (one) shpl_pool_wide_kernel<1, 1><<<g, kThreads, 0, s>>>(a);
        else shpl_pool_wide_kernel<1, 2><<<g, kThreads, 0, s>>>(a);
    } else if (stream_split) {
        // sparse regime: entry CTAs (gathers) + stream CTAs (dense parts, zeros) in one launch
        long long total_tiles = 0;
        for (int i = 0; i < a.n_jobs; ++i) total_tiles += a.job[i].tiles;
        // stream CTAs per SM: 6.  Measured in the bench step (4 streams): 4 -> 8479 frames/s but layer B forward alone
        // 45.1 us; 6 -> 8353 / 42.7 us; 8 -> 8172 / 42.2 us; 12 -> 7836 / 41.1 us.  Fewer CTAs leave room for the kernels
        // of the other streams, more CTAs make the kernel itself faster.
        const long long cap = (long long)shpl::sm_count() * stream_ctas_per_sm();
        long long want = (total_tiles + kWarps - 1) / kWarps;
        if (want > cap) want = cap;
        a.begin[0] = 0;
        for (int i = 0; i < a.n_jobs; ++i) {
            Job& o = a.job[i];
            long long c = (want * o.tiles + total_tiles - 1) / (total_tiles > 0 ? total_tiles : 1);
            const long long need = ((long long)o.tiles + kWarps - 1) / kWarps;
            if (c > need) c = need;
            if (c < 1) c = 1;
            o.stream_ctas = (int)c;
            // entries per warp: 16 measured best for rows up to 1 KB (bench step 130 -> 123 us against 8; 24 gains 2 % more
            // there but hurts skewed maps); 3 KB rows (MV3D, C = 768) want the shorter chains of 8
            o.entry_chunk = entry_chunk_knob() > 0 ? entry_chunk_knob() : (o.vs * w > 256 ? 8 : 16);
            o.packed = packed ? 1 : 0;
            o.long_len = packed ? 512 : kLongRow;     // the packed walk hands a row over at ~60 cycles per entry: fine up to 512
            if (packed && o.vs > 0) o.entry_chunk = packed_chunk_knob();
            o.staged = staged;
            if (staged) o.long_len = o.heavy_len > 0 ? o.heavy_len : 0x7fffffff;   // no cell is left to the stream warps
            if (staged == 2 && o.vs > 0) {      // CTA chunk = a whole number of batches
                const int batch = kStageVecs / o.vs < kStageEntries ? kStageVecs / o.vs : kStageEntries;
                // batches per entry CTA: enough CTAs to fill the machine twice over, at most staged_batches_knob()
                long long nb = (long long)src_spec[i]->nnz_max / ((long long)batch * 2 * shpl::sm_count());
                if (nb > staged_batches_knob()) nb = staged_batches_knob();
                if (nb < 1) nb = 1;
                o.entry_chunk = batch * (int)nb / kWarps;
                if (o.entry_chunk < 4) o.entry_chunk = 4;
            }
            o.q_slices = 1;
            if (!staged && !packed && o.vs > 64 && q_slices_knob()) {
                o.q_slices = (o.vs + 63) / 64;                        // 32 lanes x ACC = 2 vectors per slice
                // the slices multiply the warps: longer chunks keep their number (measured at MV3D's C = 768, three slices:
                // 8 entries per warp 32.3 / 25.8 us forward / backward, 24 entries 27.9 / 23.4 us)
                if (entry_chunk_knob() == 0) o.entry_chunk *= o.q_slices;
            }
            const long long chunks = ((long long)src_spec[i]->nnz_max + o.entry_chunk - 1) / o.entry_chunk;
            o.entry_ctas = o.vs > 0 ? (int)((chunks * o.q_slices + kWarps - 1) / kWarps) : 0;
            a.begin[i + 1] = a.begin[i] + o.entry_ctas + o.stream_ctas;
        }
        const unsigned g = (unsigned)a.begin[a.n_jobs];
        const bool add = a.job[0].add != 0;
        if (staged) {
#define SHPL_LAUNCH_STAGED(WW, ACC_)                                                                                   \
    do {                                                                                                               \
        using VV = VecOf<WW>::type;                                                                                    \
        if (listed && add) shpl_pool_sparse_kernel<WW, true, ACC_, 2><<<g, kThreads, stage_smem_bytes<VV>(), s>>>(a);  \
        else if (listed) shpl_pool_sparse_kernel<WW, false, ACC_, 2><<<g, kThreads, stage_smem_bytes<VV>(), s>>>(a);   \
        else if (add) shpl_pool_sparse_kernel<WW, true, ACC_, 1><<<g, kThreads, stage_smem_bytes<VV>(), s>>>(a);       \
        else shpl_pool_sparse_kernel<WW, false, ACC_, 1><<<g, kThreads, stage_smem_bytes<VV>(), s>>>(a);               \
    } while (0)
            if (max_vs > 32) {
                if (w == 4) SHPL_LAUNCH_STAGED(4, 2);
                else if (w == 2) SHPL_LAUNCH_STAGED(2, 2);
                else SHPL_LAUNCH_STAGED(1, 2);
            } else if (w == 4) SHPL_LAUNCH_STAGED(4, 1);
            else if (w == 2) SHPL_LAUNCH_STAGED(2, 1);
            else SHPL_LAUNCH_STAGED(1, 1);
#undef SHPL_LAUNCH_STAGED
        } else if (max_vs > 32) {
            if (w == 4 && add) shpl_pool_sparse_kernel<4, true, 2><<<g, kThreads, 0, s>>>(a);
            else if (w == 4) shpl_pool_sparse_kernel<4, false, 2><<<g, kThreads, 0, s>>>(a);
            else if (w == 2 && add) shpl_pool_sparse_kernel<2, true, 2><<<g, kThreads, 0, s>>>(a);
            else if (w == 2) shpl_pool_sparse_kernel<2, false, 2><<<g, kThreads, 0, s>>>(a);
            else if (add) shpl_pool_sparse_kernel<1, true, 2><<<g, kThreads, 0, s>>>(a);
            else shpl_pool_sparse_kernel<1, false, 2><<<g, kThreads, 0, s>>>(a);
        } else if (w == 4 && add) shpl_pool_sparse_kernel<4, true, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 4) shpl_pool_sparse_kernel<4, false, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 2 && add) shpl_pool_sparse_kernel<2, true, 1><<<g, kThreads, 0, s>>>(a);
        else if (w == 2) shpl_pool_sparse_kernel<2, false, 1><<<g, kThreads, 0, s>>>(a);
        else if (add) shpl_pool_sparse_kernel<1, true, 1><<<g, kThreads, 0, s>>>(a);
        else shpl_pool_sparse_kernel<1, false, 1><<<g, kThreads, 0, s>>>(a);
    } else {
        // partition the resident grid between the jobs in proportion to their tiles (a CTA serves one job)
        long long total_tiles = 0;
        for (int i = 0; i < a.n_jobs; ++i) total_tiles += a.job[i].tiles;
        // 8 CTAs per SM although only 3 are resident (80 registers): measured on B200 (profiles/README.md), a grid of
        // exactly the resident CTAs is 15 % slower -- the late waves start on SMs whose first CTAs have drained and
        // keep the memory pipes fed through the tail.
        const long long cap = (long long)shpl::sm_count() * narrow_ctas_per_sm();
        long long want = (total_tiles + kWarps - 1) / kWarps;
        if (want > cap) want = cap;
        a.begin[0] = 0;
        for (int i = 0; i < a.n_jobs; ++i) {
            long long c = (want * a.job[i].tiles + total_tiles - 1) / (total_tiles > 0 ? total_tiles : 1);
            const long long need = ((long long)a.job[i].tiles + kWarps - 1) / kWarps;
            if (c > need) c = need;
            if (c < 1) c = 1;
            a.job[i].entry_ctas = (int)c;          // narrow kernels: number of CTAs serving this job
            a.begin[i + 1] = a.begin[i] + (int)c;
        }
        const unsigned g = (unsigned)a.begin[a.n_jobs];
        const bool add = a.job[0].add != 0;      // the jobs of one launch share the form
        if (w == 4 && add) shpl_pool_narrow_kernel<4, true><<<g, kThreads, 0, s>>>(a);
        else if (w == 4) shpl_pool_narrow_kernel<4, false><<<g, kThreads, 0, s>>>(a);
        else if (w == 2 && add) shpl_pool_narrow_kernel<2, true><<<g, kThreads, 0, s>>>(a);
        else if (w == 2) shpl_pool_narrow_kernel<2, false><<<g, kThreads, 0, s>>>(a);
        else if (add) shpl_pool_narrow_kernel<1, true><<<g, kThreads, 0, s>>>(a);
        else shpl_pool_narrow_kernel<1, false><<<g, kThreads, 0, s>>>(a);
    }
    shpl::count_launches(1);
    return shpl::check_launch(who);
}

JobSpec forward_job(const float* dst, const float* src, const int32_t* ptr, const int32_t* key, const int32_t* idx,
                    const float* val, int nnz_max, int heavy_len, int n_rows, int C_d, int C_s, float* fused) {
    JobSpec j;
    j.dense_in = dst;
    j.dense_in_stride = C_d;
    j.c_dense = C_d;
    j.gather_in = src;
    j.gather_stride = C_s;
    j.c_pool = C_s;
    j.dense_out = fused;
    j.dense_out_stride = C_d + C_s;
    j.pool_out = fused + C_d;
    j.pool_out_stride = C_d + C_s;
    j.ptr = ptr;
    j.key = key;
    j.idx = idx;
    j.val = val;
    j.nnz_max = nnz_max;
    j.heavy_len = heavy_len;
    j.n_cells = n_rows;
    return j;
}

}  // namespace

extern "C" int shpl_pool_forward(const float* dst, const float* src, const int32_t* ptr, const int32_t* key,
                                 const int32_t* idx, const float* val, int32_t nnz_max, int32_t heavy_len,
                                 int32_t n_rows, int32_t C_d, int32_t n_src, int32_t C_s, float* fused, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_d >= 0 && C_s > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_forward: bad sizes n_rows=%d n_src=%d C_d=%d C_s=%d", n_rows, n_src, C_d, C_s);
    SHPL_REQUIRE(src && ptr && fused && (C_d == 0 || dst), SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward: null pointer");
    SHPL_REQUIRE(idx && val, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward: null idx/val");
    if (n_rows == 0) return SHPL_OK;
    JobSpec j = forward_job(dst, src, ptr, key, idx, val, nnz_max, heavy_len, n_rows, C_d, C_s, fused);
    j.n_gather = n_src;
    return launch_jobs(&j, 1, static_cast<cudaStream_t>(stream), "shpl_pool_forward");
}

extern "C" int shpl_pool_backward(const float* g_fused, const int32_t* ptrT, const int32_t* keyT, const int32_t* idxT,
                                  const float* valT, int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_d,
                                  int32_t n_src, int32_t C_s, float* g_dst, float* g_src, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_d >= 0 && C_s > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_backward: bad sizes n_rows=%d n_src=%d C_d=%d C_s=%d", n_rows, n_src, C_d, C_s);
    SHPL_REQUIRE(g_fused && ptrT && idxT && valT && g_src, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_backward: null pointer");
    JobSpec js[2];
    // job 0: g_src[p] = sum over the entries of source p of valT * g_fused[idxT, C_d:]
    js[0].gather_in = g_fused + C_d;
    js[0].gather_stride = C_d + C_s;
    js[0].c_pool = C_s;
    js[0].pool_out = g_src;
    js[0].pool_out_stride = C_s;
    js[0].ptr = ptrT;
    js[0].key = keyT;
    js[0].idx = idxT;
    js[0].val = valT;
    js[0].nnz_max = nnz_max;
    js[0].heavy_len = heavy_len;
    js[0].n_cells = n_src;
    js[0].n_gather = n_rows;
    // job 1: g_dst = g_fused[:, :C_d]
    if (g_dst != nullptr && C_d > 0) {
        js[1].dense_in = g_fused;
        js[1].dense_in_stride = C_d + C_s;
        js[1].c_dense = C_d;
        js[1].dense_out = g_dst;
        js[1].dense_out_stride = C_d;
        js[1].n_cells = n_rows;
    }
    return launch_jobs(js, 2, static_cast<cudaStream_t>(stream), "shpl_pool_backward");
}

extern "C" int shpl_pool_forward_dual(const float* bev, const float* img, const int32_t* row_ptr,
                                      const int32_t* csr_row, const int32_t* csr_src, const float* csr_val,
                                      const int32_t* pix_ptr, const int32_t* csrT_pix, const int32_t* csrT_dst,
                                      const float* csrT_val, int32_t nnz_max, int32_t heavy_len, int32_t n_rows,
                                      int32_t C_b, int32_t n_src, int32_t C_i, float* fused_bev, float* fused_img,
                                      void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_b > 0 && C_i > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_forward_dual: bad sizes n_rows=%d n_src=%d C_b=%d C_i=%d", n_rows, n_src, C_b, C_i);
    SHPL_REQUIRE(bev && img && row_ptr && csr_src && csr_val && pix_ptr && csrT_dst && csrT_val && fused_bev && fused_img,
                 SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward_dual: null pointer");
    JobSpec js[2];
    js[0] = forward_job(bev, img, row_ptr, csr_row, csr_src, csr_val, nnz_max, heavy_len, n_rows, C_b, C_i, fused_bev);
    js[1] = forward_job(img, bev, pix_ptr, csrT_pix, csrT_dst, csrT_val, nnz_max, heavy_len, n_src, C_i, C_b, fused_img);
    js[0].n_gather = n_src;
    js[1].n_gather = n_rows;
    return launch_jobs(js, 2, static_cast<cudaStream_t>(stream), "shpl_pool_forward_dual");
}

extern "C" int shpl_pool_backward_dual(const float* g_fused_bev, const float* g_fused_img, const int32_t* row_ptr,
                                       const int32_t* csr_row, const int32_t* csr_src, const float* csr_val,
                                       const int32_t* pix_ptr, const int32_t* csrT_pix, const int32_t* csrT_dst,
                                       const float* csrT_val, int32_t nnz_max, int32_t heavy_len, int32_t n_rows,
                                       int32_t C_b, int32_t n_src, int32_t C_i, float* g_bev, float* g_img,
                                       void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_b > 0 && C_i > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_backward_dual: bad sizes n_rows=%d n_src=%d C_b=%d C_i=%d", n_rows, n_src, C_b, C_i);
    SHPL_REQUIRE(g_fused_bev && g_fused_img && row_ptr && csr_src && csr_val && pix_ptr && csrT_dst && csrT_val &&
                     g_bev && g_img, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_backward_dual: null pointer");
    const int Fb = C_b + C_i, Fi = C_i + C_b;
    JobSpec js[2];
    // g_bev[r] = g_fused_bev[r, :C_b] + sum_{k in row r} val * g_fused_img[pix_k, C_i:]
    js[0].dense_in = g_fused_bev;
    js[0].dense_in_stride = Fb;
    js[0].c_dense = C_b;
    js[0].gather_in = g_fused_img + C_i;
    js[0].gather_stride = Fi;
    js[0].c_pool = C_b;
    js[0].pool_out = g_bev;
    js[0].pool_out_stride = C_b;
    js[0].ptr = row_ptr;
    js[0].key = csr_row;
    js[0].idx = csr_src;
    js[0].val = csr_val;
    js[0].nnz_max = nnz_max;
    js[0].n_cells = n_rows;
    js[0].n_gather = n_src;
    js[0].add = 1;
    js[0].heavy_len = heavy_len;
    // g_img[p] = g_fused_img[p, :C_i] + sum_{k at pixel p} val * g_fused_bev[row_k, C_b:]
    js[1].dense_in = g_fused_img;
    js[1].dense_in_stride = Fi;
    js[1].c_dense = C_i;
    js[1].gather_in = g_fused_bev + C_b;
    js[1].gather_stride = Fb;
    js[1].c_pool = C_i;
    js[1].pool_out = g_img;
    js[1].pool_out_stride = C_i;
    js[1].ptr = pix_ptr;
    js[1].key = csrT_pix;
    js[1].idx = csrT_dst;
    js[1].val = csrT_val;
    js[1].nnz_max = nnz_max;
    js[1].n_cells = n_src;
    js[1].n_gather = n_rows;
    js[1].add = 1;
    js[1].heavy_len = heavy_len;
    return launch_jobs(js, 2, static_cast<cudaStream_t>(stream), "shpl_pool_backward_dual");
}

// ---- no-concat ("sparse-only") forms, SURVEY.md 8(d): the producer of the destination map writes its channels straight
// into the fused buffer, the op writes only the pooled channels (zeros for cells that receive nothing); the backward
// reads the pooled channels of g_fused in place and g_dst is simply the view g_fused[:, :C_d] -- no slice copy.
namespace {
JobSpec into_job(const float* src, const int32_t* ptr, const int32_t* key, const int32_t* idx, const float* val, int nnz_max,
                 int heavy_len, int n_rows, int C_s, float* fused, int fused_stride, int chan_off) {
    JobSpec j;
    j.gather_in = src;
    j.gather_stride = C_s;
    j.c_pool = C_s;
    j.pool_out = fused + chan_off;
    j.pool_out_stride = fused_stride;
    j.ptr = ptr;
    j.key = key;
    j.idx = idx;
    j.val = val;
    j.nnz_max = nnz_max;
    j.heavy_len = heavy_len;
    j.n_cells = n_rows;
    return j;
}
}  // namespace

extern "C" int shpl_pool_forward_into(const float* src, const int32_t* ptr, const int32_t* key, const int32_t* idx,
                                      const float* val, int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t n_src,
                                      int32_t C_s, float* fused, int32_t fused_stride, int32_t chan_off, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_s > 0 && chan_off >= 0 && fused_stride >= chan_off + C_s, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_forward_into: bad sizes n_rows=%d n_src=%d C_s=%d fused_stride=%d chan_off=%d", n_rows, n_src, C_s,
                 fused_stride, chan_off);
    SHPL_REQUIRE(src && ptr && idx && val && fused, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward_into: null pointer");
    if (n_rows == 0) return SHPL_OK;
    JobSpec j = into_job(src, ptr, key, idx, val, nnz_max, heavy_len, n_rows, C_s, fused, fused_stride, chan_off);
    j.n_gather = n_src;
    return launch_jobs(&j, 1, static_cast<cudaStream_t>(stream), "shpl_pool_forward_into");
}

extern "C" int shpl_pool_forward_into_dual(const float* bev, const float* img, const int32_t* row_ptr, const int32_t* csr_row,
                                           const int32_t* csr_src, const float* csr_val, const int32_t* pix_ptr,
                                           const int32_t* csrT_pix, const int32_t* csrT_dst, const float* csrT_val,
                                           int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_b, int32_t n_src,
                                           int32_t C_i, float* fused_bev, float* fused_img, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_b > 0 && C_i > 0, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_forward_into_dual: bad sizes n_rows=%d n_src=%d C_b=%d C_i=%d", n_rows, n_src, C_b, C_i);
    SHPL_REQUIRE(bev && img && row_ptr && csr_src && csr_val && pix_ptr && csrT_dst && csrT_val && fused_bev && fused_img,
                 SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_forward_into_dual: null pointer");
    JobSpec js[2];
    js[0] = into_job(img, row_ptr, csr_row, csr_src, csr_val, nnz_max, heavy_len, n_rows, C_i, fused_bev, C_b + C_i, C_b);
    js[1] = into_job(bev, pix_ptr, csrT_pix, csrT_dst, csrT_val, nnz_max, heavy_len, n_src, C_b, fused_img, C_i + C_b, C_i);
    js[0].n_gather = n_src;
    js[1].n_gather = n_rows;
    return launch_jobs(js, 2, static_cast<cudaStream_t>(stream), "shpl_pool_forward_into_dual");
}

extern "C" int shpl_pool_backward_from(const float* g_fused, int32_t g_stride, int32_t chan_off, const int32_t* ptrT,
                                       const int32_t* keyT, const int32_t* idxT, const float* valT, int32_t nnz_max,
                                       int32_t heavy_len, int32_t n_rows, int32_t n_src, int32_t C_s, float* g_src, void* stream) {
    SHPL_REQUIRE(n_rows >= 0 && n_src >= 0 && C_s > 0 && chan_off >= 0 && g_stride >= chan_off + C_s, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_backward_from: bad sizes n_rows=%d n_src=%d C_s=%d g_stride=%d chan_off=%d", n_rows, n_src, C_s, g_stride,
                 chan_off);
    SHPL_REQUIRE(g_fused && ptrT && idxT && valT && g_src, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_backward_from: null pointer");
    if (n_src == 0) return SHPL_OK;
    JobSpec j;
    j.gather_in = g_fused + chan_off;
    j.gather_stride = g_stride;
    j.c_pool = C_s;
    j.pool_out = g_src;
    j.pool_out_stride = C_s;
    j.ptr = ptrT;
    j.key = keyT;
    j.idx = idxT;
    j.val = valT;
    j.nnz_max = nnz_max;
    j.heavy_len = heavy_len;
    j.n_cells = n_src;
    j.n_gather = n_rows;
    return launch_jobs(&j, 1, static_cast<cudaStream_t>(stream), "shpl_pool_backward_from");
}

namespace {
size_t heavy_ws_pieces(int64_t nnz_max, int32_t list_cap) {
    // ceil(L / kPieceLen) summed over the listed cells <= nnz_max / kPieceLen + number of listed cells
    const int listed = list_cap < kMaxListed ? (list_cap > 0 ? list_cap : 0) : kMaxListed;
    return (size_t)(nnz_max > 0 ? nnz_max : 0) / kPieceLen + (size_t)listed + 1;
}
}  // namespace

extern "C" size_t shpl_pool_heavy_workspace_bytes(int32_t C, int64_t nnz_max, int32_t list_cap) {
    if (C <= 0) return 0;
    return 256 + (size_t)(kMaxListed + 64) * sizeof(int) + heavy_ws_pieces(nnz_max, list_cap) * (size_t)C * sizeof(float);
}

static int pool_heavy_impl(const float* gather_in, int32_t gather_stride, int32_t C, const int32_t* ptr,
                           const int32_t* idx, const float* val, const int32_t* list, const int32_t* count_dev,
                           int32_t list_cap, const float* addend, int32_t addend_stride, float* out,
                           int32_t out_stride, int64_t nnz_max, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int shpl_pool_heavy(const float* gather_in, int32_t gather_stride, int32_t C, const int32_t* ptr,
                               const int32_t* idx, const float* val, const int32_t* list, const int32_t* count_dev,
                               int32_t list_cap, const float* addend, int32_t addend_stride, float* out,
                               int32_t out_stride, void* stream) {
    return pool_heavy_impl(gather_in, gather_stride, C, ptr, idx, val, list, count_dev, list_cap, addend, addend_stride, out,
                           out_stride, 0, nullptr, 0, stream);
}

extern "C" int shpl_pool_heavy_split(const float* gather_in, int32_t gather_stride, int32_t C, const int32_t* ptr,
                                     const int32_t* idx, const float* val, const int32_t* list, const int32_t* count_dev,
                                     int32_t list_cap, const float* addend, int32_t addend_stride, float* out,
                                     int32_t out_stride, int64_t nnz_max, void* workspace, size_t workspace_bytes, void* stream) {
    SHPL_REQUIRE(workspace != nullptr && shpl::aligned(workspace, 16), SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_heavy_split: workspace must be a 16-byte aligned device buffer");
    SHPL_REQUIRE(workspace_bytes >= shpl_pool_heavy_workspace_bytes(C, nnz_max, list_cap), SHPL_ERR_WORKSPACE_TOO_SMALL,
                 "shpl_pool_heavy_split: workspace %zu < %zu bytes", workspace_bytes, shpl_pool_heavy_workspace_bytes(C, nnz_max, list_cap));
    return pool_heavy_impl(gather_in, gather_stride, C, ptr, idx, val, list, count_dev, list_cap, addend, addend_stride, out,
                           out_stride, nnz_max, workspace, workspace_bytes, stream);
}

static int pool_heavy_impl(const float* gather_in, int32_t gather_stride, int32_t C, const int32_t* ptr,
                           const int32_t* idx, const float* val, const int32_t* list, const int32_t* count_dev,
                           int32_t list_cap, const float* addend, int32_t addend_stride, float* out,
                           int32_t out_stride, int64_t nnz_max, void* workspace, size_t workspace_bytes, void* stream) {
    SHPL_REQUIRE(gather_in && ptr && idx && val && list && count_dev && out, SHPL_ERR_INVALID_ARGUMENT,
                 "shpl_pool_heavy: null pointer");
    SHPL_REQUIRE(C > 0 && list_cap >= 0, SHPL_ERR_INVALID_ARGUMENT, "shpl_pool_heavy: bad sizes C=%d list_cap=%d", C,
                 list_cap);
    if (list_cap == 0) return SHPL_OK;
    const int w = pick_width({C, gather_stride, addend ? addend_stride : 0, out_stride}, {gather_in, addend, out});
    HeavyArgs a{};
    a.gather_in = gather_in;
    a.addend = addend;
    a.out = out;
    a.ptr = ptr;
    a.idx = idx;
    a.val = val;
    a.list = list;
    a.count_dev = count_dev;
    a.list_cap = list_cap;
    a.gather_stride = gather_stride / w;
    a.addend_stride = addend_stride / w;
    a.out_stride = out_stride / w;
    a.nv = C / w;
    a.n_gather = 0x7fffffff;
    const size_t smem = (size_t)(kWarps + 1) * a.nv * sizeof(float) * w;
    SHPL_REQUIRE(smem <= 200 * 1024, SHPL_ERR_UNSUPPORTED, "shpl_pool_heavy: C=%d needs %zu bytes of shared memory", C, smem);
    const int clusters = list_cap < 64 ? list_cap : 64;
    const unsigned grid = (unsigned)(clusters * kClusterSize);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // exact kernel: 8 adder warps cover 256 channel vectors; a round of products (64 warps x eu entries) is parked in
    // CTA 0's shared memory: up to 128 KB of it, 196 KB when one entry per warp already needs that much (C = 768)
    a.exact_len = a.nv <= kWarps * 32 && (size_t)kClusterSize * kWarps * C * sizeof(float) <= 200 * 1024 ? SHPL_EXACT_LEN : 0;
    a.skip_le = main_kernel_heavy_len(SHPL_HEAVY_LEN, C) > SHPL_HEAVY_LEN ? main_kernel_heavy_len(SHPL_HEAVY_LEN, C) : 0;
    if (a.exact_len > a.skip_le) {
        // with a workspace (the split path below) many cells may be listed: 2-CTA clusters, one per SM pair of CTAs, four
        // times as many cells in flight; without one, the original 8-CTA clusters
        const int cs = (workspace != nullptr) ? 2 : kClusterSize;
        const int E = a.nv >= 32 ? 1 : 32 / a.nv;
        int eu = (int)((128 * 1024) / ((size_t)cs * kWarps * C * sizeof(float)));
        if (eu > 32) eu = 32;
        eu = eu / E * E;
        if (eu < E) eu = E;
        a.eu = eu;
        const size_t smem_x = (size_t)cs * kWarps * eu * C * sizeof(float);
        const int max_clusters = cs == 2 ? shpl::sm_count() / 2 * 2 : 128;
        const int clusters_x = list_cap < max_clusters ? list_cap : max_clusters;
        const unsigned grid_x = (unsigned)(clusters_x * cs);
#define SHPL_LAUNCH_EXACT(WW, CC)                                                                                       \
    do {                                                                                                                \
        if (smem_x > 48 * 1024)                                                                                         \
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_pool_heavy_exact_kernel<WW, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)); \
        shpl_pool_heavy_exact_kernel<WW, CC><<<grid_x, kThreads, smem_x, s>>>(a);                                       \
    } while (0)
        if (cs == 2) {
            if (w == 4) SHPL_LAUNCH_EXACT(4, 2);
            else if (w == 2) SHPL_LAUNCH_EXACT(2, 2);
            else SHPL_LAUNCH_EXACT(1, 2);
        } else {
            if (w == 4) SHPL_LAUNCH_EXACT(4, kClusterSize);
            else if (w == 2) SHPL_LAUNCH_EXACT(2, kClusterSize);
            else SHPL_LAUNCH_EXACT(1, kClusterSize);
        }
#undef SHPL_LAUNCH_EXACT
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_pool_heavy_exact_kernel")) return rc;
    }
    const bool groups = a.nv * 2 <= 32;       // few channels: lane groups over entries (see the kernel)
    if (workspace != nullptr && list_cap <= kMaxListed) {
        // long cells split over many CTAs + in-order combine (two launches); partials in the caller's workspace
        (void)workspace_bytes;
        SplitArgs sa{};
        uint8_t* wsb = static_cast<uint8_t*>(workspace);
        sa.prefix = reinterpret_cast<int*>(wsb);
        sa.partial = wsb + 256 + (size_t)(kMaxListed + 64) * sizeof(int) - ((size_t)(kMaxListed + 64) * sizeof(int)) % 16;
        sa.max_pieces = (int)heavy_ws_pieces(nnz_max, list_cap);
        const size_t smem_s = (size_t)kWarps * a.nv * sizeof(float) * w;
        const unsigned grid_s = (unsigned)(shpl::sm_count() * 4);      // pieces are gather-latency bound: many CTAs in flight
#define SHPL_LAUNCH_SPLIT(WW, GG)                                                                                       \
    do {                                                                                                                \
        if (smem_s > 48 * 1024)                                                                                         \
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_pool_heavy_split_kernel<WW, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s)); \
        shpl_pool_heavy_split_kernel<WW, GG><<<grid_s, kThreads, smem_s, s>>>(a, sa);                                   \
    } while (0)
        if (w == 4 && groups) SHPL_LAUNCH_SPLIT(4, true);
        else if (w == 4) SHPL_LAUNCH_SPLIT(4, false);
        else if (w == 2 && groups) SHPL_LAUNCH_SPLIT(2, true);
        else if (w == 2) SHPL_LAUNCH_SPLIT(2, false);
        else if (groups) SHPL_LAUNCH_SPLIT(1, true);
        else SHPL_LAUNCH_SPLIT(1, false);
#undef SHPL_LAUNCH_SPLIT
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_pool_heavy_split_kernel")) return rc;
        const unsigned grid_c = (unsigned)(list_cap < 256 ? list_cap : 256);
        const size_t smem_c = (size_t)kCombineGroups * a.nv * sizeof(float) * w;
#define SHPL_LAUNCH_COMBINE(WW)                                                                                         \
    do {                                                                                                                \
        if (smem_c > 48 * 1024)                                                                                         \
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_pool_heavy_combine_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c)); \
        shpl_pool_heavy_combine_kernel<WW><<<grid_c, kThreads, smem_c, s>>>(a, sa);                                     \
    } while (0)
        if (w == 4) SHPL_LAUNCH_COMBINE(4);
        else if (w == 2) SHPL_LAUNCH_COMBINE(2);
        else SHPL_LAUNCH_COMBINE(1);
#undef SHPL_LAUNCH_COMBINE
        shpl::count_launches(1);
        return shpl::check_launch("shpl_pool_heavy_combine_kernel");
    }
#define SHPL_LAUNCH_HEAVY(WW, GG)                                                                                       \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            SHPL_CUDA_OK(cudaFuncSetAttribute(shpl_pool_heavy_kernel<WW, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        shpl_pool_heavy_kernel<WW, GG><<<grid, kThreads, smem, s>>>(a);                                                 \
    } while (0)
    if (w == 4 && groups) SHPL_LAUNCH_HEAVY(4, true);
    else if (w == 4) SHPL_LAUNCH_HEAVY(4, false);
    else if (w == 2 && groups) SHPL_LAUNCH_HEAVY(2, true);
    else if (w == 2) SHPL_LAUNCH_HEAVY(2, false);
    else if (groups) SHPL_LAUNCH_HEAVY(1, true);
    else SHPL_LAUNCH_HEAVY(1, false);
#undef SHPL_LAUNCH_HEAVY
    shpl::count_launches(1);
    return shpl::check_launch("shpl_pool_heavy_kernel");
}
