// SHPL correspondence builder for sm_100a.
//
// Replaces the numpy host half of the reference:
//   /root/reference/avod/avod/utils/transform.py:3-40          projectToImage / clip3DwithinImage
//   /root/reference/avod/avod/utils/sparse_pool_utils.py:6-20  gen_sparse_pooling_input_avod
//   /root/reference/avod/avod/utils/sparse_pool_utils.py:22-58 produce_sparse_pooling_input
// and turns the COO the reference hands to tf.SparseTensor (rpn_model.py:292-293) into the
// canonical CSR (by destination cell) and CSR^T (by source pixel) the pooling kernels read.
//
// Pipeline (all on the caller's stream, no host synchronisation):
//   memset   header (tile tickets, look-back words, digit histograms)
//   K1       shpl_pairs_kernel: fp64 projection (the fma chain the reference's BLAS call rounds
//            with), clip, rint, stride floor/clamp, row id, row filter; STABLE compaction of
//            the survivors by a single-pass decoupled look-back scan; emits Mij_pool,
//            img_index_flip_pool, M_val and the (key,k) sort items of both sorts, and counts
//            the first radix digit per tile
//   S x P    shpl_radix_pass_kernel: stable LSD radix pass by destination cell (grid.y = 0) and
//            by source pixel (grid.y = 1); ranks come from warp match-any over in-order chunks,
//            tile bases from the per-tile digit counts of the previous kernel
//   F        shpl_finalize_kernel: row offsets by lower-bound over the sorted keys, payload
//            gather into csr_src/csr_val and csrT_dst/csrT_val
// Entries whose row / pixel is out of range get the sentinel key n (sorted last) and never
// enter the CSRs; they are counted in counts[2].
#include "shpl_common.cuh"
#include "shpl_sort.cuh"

namespace {
using shpl::lookback;

enum Mode { kModeAvod = 0, kModePairs = 1, kModeCoo = 2, kModeGenOnly = 3, kModeVoxel = 4 };

struct PairsArgs {
    int mode;
    long long n;                 // candidates
    const int* n_dev;            // optional device count: candidates i >= *n_dev do not exist
    // kModeAvod / kModeGenOnly
    const double* points;
    const long long* vox;
    double P[12];
    // kModePairs
    double* img_u;
    double* img_v;
    const long long* bv_index;
    // kModeCoo
    const long long* coo;
    const float* coo_val;
    const void* src_index;
    int index_is_i64;
    long long ncol;
    // kModeVoxel: src_index holds the [K,4] (batch, d, h, w) coordinates, index_is_i64 their width
    int grid[4];                 // B, D, H, W of the dense voxel grid
    // common
    const double* m_val;
    int im_w, im_h;              // raw image size (clip)
    int s_img, s_bv;
    int Wp, Hp, Wb, Hb;          // strided sizes
    int n_rows, src_h, src_w;
    int row_base, pix_base;
    // outputs
    long long* gen_bv;
    double* gen_u;
    double* gen_v;
    long long* mij;
    long long* flip;
    float* mval_out;
    long long* msize_out;
    int* counts;
    int* heavy_count;            // zeroed by the first frame of a plan (row_base == pix_base == 0)
    Workspace ws;
    SortPlan sp[2];
};

// floor(a / s) for s > 0; cell indices fit 32 bits in practice, 64-bit division is the slow path
__device__ __forceinline__ long long floordiv_ll(long long a, int s) {
    if (a == (long long)(int)a) {
        const int ai = (int)a;
        int q = ai / s;
        if (ai - q * s < 0) --q;
        return q;
    }
    long long q = a / s;
    if (a - q * s < 0) --q;
    return q;
}

// transform.py:17-24 -- one row of P times [x y z 1], rounded like the reference's dgemm
__device__ __forceinline__ double prow(const double* p, double x, double y, double z) {
    double t = __dmul_rn(p[0], x);
    t = __fma_rn(p[1], y, t);
    t = __fma_rn(p[2], z, t);
    t = __fma_rn(p[3], 1.0, t);
    return t;
}

// What one candidate pair contributes (recomputed in both passes of shpl_pairs_kernel: the
// kernel is latency-bound and straight-line code this size costs more in instruction fetch
// than the arithmetic does, so the chunk loops are kept rolled and nothing is cached per chunk).
struct Cand {
    bool clip, keep;
    long long row;
    int up, vp;
    double gu, gv;      // rounded (u, v) before the stride is applied: the gen dict's img_index
    double us, vs;      // after floor(/stride) and clamp: what the reference writes back in place
    long long bx, bz;
};

__device__ __forceinline__ Cand eval_candidate(const PairsArgs& a, long long i, long long R) {
    Cand c;
    c.clip = c.keep = false;
    c.row = 0; c.up = 0; c.vp = 0; c.gu = 0; c.gv = 0; c.us = 0; c.vs = 0; c.bx = 0; c.bz = 0;
    if (i >= a.n || (a.n_dev != nullptr && i >= (long long)*a.n_dev)) return c;
    if (a.mode == kModeCoo) {
        const long long r = a.coo[2 * i], col = a.coo[2 * i + 1];
        c.clip = c.keep = true;
        c.row = r;
        long long b = -1, v = -1, u = -1;
        if (col >= 0 && col < a.ncol) {
            if (a.index_is_i64) {
                const long long* q = static_cast<const long long*>(a.src_index) + 3 * col;
                b = q[0]; v = q[1]; u = q[2];
            } else {
                const int* q = static_cast<const int*>(a.src_index) + 3 * col;
                b = q[0]; v = q[1]; u = q[2];
            }
        }
        const bool ok = (b == 0) && v >= 0 && v < a.src_h && u >= 0 && u < a.src_w;
        c.vp = ok ? (int)v : -1;
        c.up = ok ? (int)u : -1;
        return c;
    }
    if (a.mode == kModeVoxel) {
        // group_pointcloud.py:84-85: row k of the voxel-wise features lands in grid cell coordinate[k]
        long long q[4];
        if (a.index_is_i64) {
            const longlong2* p = reinterpret_cast<const longlong2*>(static_cast<const long long*>(a.src_index) + 4 * i);
            const longlong2 lo = p[0], hi = p[1];
            q[0] = lo.x; q[1] = lo.y; q[2] = hi.x; q[3] = hi.y;
        } else {
            const int4 t = *reinterpret_cast<const int4*>(static_cast<const int*>(a.src_index) + 4 * i);
            q[0] = t.x; q[1] = t.y; q[2] = t.z; q[3] = t.w;
        }
        bool in = true;
        long long r = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            in = in && q[d] >= 0 && q[d] < (long long)a.grid[d];
            r = r * a.grid[d] + q[d];
        }
        c.clip = c.keep = true;
        c.row = in ? r : -1;
        c.vp = (int)i;       // the source "pixel" is the feature row itself
        c.up = 0;
        return c;
    }
    double u, v;
    if (a.mode == kModePairs) {
        u = a.img_u[i];
        v = a.img_v[i];
        c.bx = a.bv_index[2 * i];
        c.bz = a.bv_index[2 * i + 1];
        c.clip = true;
    } else {
        const double x = a.points[3 * i], y = a.points[3 * i + 1], z = a.points[3 * i + 2];
        const double w = prow(a.P + 8, x, y, z);
        u = __ddiv_rn(prow(a.P + 0, x, y, z), w);
        v = __ddiv_rn(prow(a.P + 4, x, y, z), w);
        // transform.py:36-38 (NaN compares false)
        c.clip = (u < (double)(a.im_w - 1)) && (u >= 0.0) && (v >= 0.0) && (v < (double)(a.im_h - 1));
        u = rint(u);   // sparse_pool_utils.py:18, ties to even
        v = rint(v);
        c.bx = a.vox[2 * i];
        c.bz = a.vox[2 * i + 1];
    }
    c.gu = u;
    c.gv = v;
    if (c.clip && a.mode != kModeGenOnly) {
        // sparse_pool_utils.py:30-34
        double us = floor(u / (double)a.s_img), vs = floor(v / (double)a.s_img);
        if (us >= (double)a.Wp) us = (double)(a.Wp - 1);
        if (vs >= (double)a.Hp) vs = (double)(a.Hp - 1);
        c.us = us;
        c.vs = vs;
        const long long ul = (long long)floor(us), vl = (long long)floor(vs);
        // :38-44
        const long long xs = floordiv_ll(c.bx, a.s_bv), zs = floordiv_ll(c.bz, a.s_bv);
        c.row = zs * (long long)a.Wb + xs;
        c.keep = c.row < R;
        // exact integers for the COO output (clamped to int32 range only)
        c.up = (int)max(min(ul, (long long)INT32_MAX), (long long)INT32_MIN);
        c.vp = (int)max(min(vl, (long long)INT32_MAX), (long long)INT32_MIN);
    }
    return c;
}

// One candidate pair per thread: evaluate, count survivors (warp ballots), single-pass prefix over
// the CTAs (decoupled look-back), write the compacted outputs.  Many small CTAs on purpose: the
// kernel is a chain of dependent latencies (loads -> fp64 -> look-back -> stores), so the only
// lever is to run every candidate's chain concurrently.
__global__ void __launch_bounds__(kThreads) shpl_pairs_kernel(PairsArgs a, int use_ticket) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp_clip[kWarps], s_warp_keep[kWarps];
    __shared__ unsigned long long s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int tile = blockIdx.x;
    if (use_ticket) {          // more CTAs than can be resident: order the look-back by arrival
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ws.ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    const long long i = (long long)tile * kPairsTile + threadIdx.x;
    const long long R = (long long)a.Hb * a.Wb;
    const Cand cd = eval_candidate(a, i, R);
    const unsigned mc = __ballot_sync(kFull, cd.clip);
    const unsigned mk = __ballot_sync(kFull, cd.keep);
    if (lane == 0) {
        s_warp_clip[warp] = __popc(mc);
        s_warp_keep[warp] = __popc(mk);
    }
    __syncthreads();
    unsigned wb_clip = 0, wb_keep = 0, tot_clip = 0, tot_keep = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        if (w < warp) { wb_clip += s_warp_clip[w]; wb_keep += s_warp_keep[w]; }
        tot_clip += s_warp_clip[w];
        tot_keep += s_warp_keep[w];
    }
    if (warp == 0) {
        unsigned long long excl = 0ull;
        if (a.mode == kModeCoo) excl = ((unsigned long long)tile * kPairsTile) * ((1ull << 31) + 1ull);
        else excl = lookback(a.ws.status, tile, ((unsigned long long)tot_clip << 31) | tot_keep, lane);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    const long long j = (long long)(s_excl >> 31) + wb_clip + __popc(mc & lt);
    const long long k = (long long)(s_excl & ((1ull << 31) - 1)) + wb_keep + __popc(mk & lt);

    const long long n_tiles = (a.n + kPairsTile - 1) / kPairsTile > 0 ? (a.n + kPairsTile - 1) / kPairsTile : 1;
    if (tile == 0 && threadIdx.x == 0 && a.heavy_count != nullptr && a.row_base == 0 && a.pix_base == 0) {
        a.heavy_count[0] = 0;
        a.heavy_count[1] = 0;
    }
    if (tile == n_tiles - 1 && threadIdx.x == 0) {   // inclusive prefix of the last tile = totals
        const long long n_clip = (long long)(s_excl >> 31) + tot_clip;
        const long long nnz = (long long)(s_excl & ((1ull << 31) - 1)) + tot_keep;
        a.counts[0] = (int)n_clip;
        a.counts[1] = (a.mode == kModeGenOnly) ? 0 : (int)nnz;
        a.counts[5] = 0;                          // set by the finalize kernel when a stacked frame overruns the plan
        a.counts[6] = 0;                          // rows / pixels with more than SHPL_LONG_LEN entries (finalize kernel)
        a.counts[7] = 0;
        if (a.msize_out) { a.msize_out[0] = R; a.msize_out[1] = nnz; }
    }

    if (a.mode == kModePairs && cd.clip) {      // the reference mutates img_index in place (:30-34)
        a.img_u[i] = cd.us;
        a.img_v[i] = cd.vs;
    }
    if (cd.clip && a.gen_bv) {
        a.gen_bv[2 * j] = cd.bx;
        a.gen_bv[2 * j + 1] = cd.bz;
        a.gen_u[j] = cd.gu;
        a.gen_v[j] = cd.gv;
    }
    if (!cd.keep) return;
    const long long r = cd.row;
    float w = 1.0f;
    if (a.mode == kModeCoo) w = a.coo_val[k];
    else if (a.m_val) w = (float)a.m_val[k];      // indexed by output column (:56-57), f64 -> f32 like the feed
    const bool pix_ok = cd.vp >= 0 && cd.vp < a.src_h && cd.up >= 0 && cd.up < a.src_w;
    const bool ok = pix_ok && r >= 0 && r < (long long)a.n_rows;
    const int pix = cd.vp * a.src_w + cd.up;
    SHPL_DASSERT(k >= 0 && k < a.n && j >= 0 && j <= i);          // compacted slots never run ahead of the candidate index
    if (a.mij) { a.mij[2 * k] = r; a.mij[2 * k + 1] = k; }
    if (a.flip) { a.flip[3 * k] = 0; a.flip[3 * k + 1] = cd.vp; a.flip[3 * k + 2] = cd.up; }
    if (a.mval_out) a.mval_out[k] = w;
    a.ws.valk[k] = w;
    a.ws.rowk[k] = ok ? (int)r + a.row_base : -1;
    a.ws.pixk[k] = ok ? pix + a.pix_base : -1;
    const unsigned key_r = ok ? (unsigned)r : (unsigned)a.sp[0].n_keys;
    const unsigned key_p = ok ? (unsigned)pix : (unsigned)a.sp[1].n_keys;
    a.ws.items[0][0][k] = ((unsigned long long)key_r << 32) | (unsigned)k;
    a.ws.items[1][0][k] = ((unsigned long long)key_p << 32) | (unsigned)k;
    const int t0 = (int)(k / kTile);
    atomicAdd(a.ws.hist[0][0] + ((size_t)t0 << a.sp[0].bits[0]) + (key_r & ((1u << a.sp[0].bits[0]) - 1u)), 1u);
    atomicAdd(a.ws.hist[1][0] + ((size_t)t0 << a.sp[1].bits[0]) + (key_p & ((1u << a.sp[1].bits[0]) - 1u)), 1u);
}

struct FinalArgs {
    int row_base, pix_base;
    int* counts;
    Workspace ws;
    SortPlan sp[2];
    shpl_plan plan;
    const int* entry_base_dev;
};

__device__ __forceinline__ unsigned key_of(const unsigned long long* items, int i) {
    return (unsigned)(__ldg(items + i) >> 32);
}

// lower bound of `key` in items[lo, hi) by plain binary search (window already narrowed)
__device__ __forceinline__ int lower_bound_key(const unsigned long long* items, int lo, int hi, unsigned key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key_of(items, mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// lower bound over the whole array by one warp: 32 probes per round (3 rounds for 32k items)
__device__ int warp_lower_bound(const unsigned long long* items, int n, unsigned key, int lane) {
    int lo = 0, hi = n;                      // answer in [lo, hi]
    while (hi - lo > 32) {
        const int step = (hi - lo + 32) / 33;          // probes at lo + step*(lane+1) - 1
        const int pos = min(lo + step * (lane + 1) - 1, hi - 1);
        const bool less = key_of(items, pos) < key;    // monotone: true for a prefix of the lanes
        const unsigned m = __ballot_sync(kFull, less);
        const int c = __popc(m);                        // probes strictly below the key
        const int new_lo = c == 0 ? lo : min(lo + step * c, hi);   // items up to probe c-1 are < key
        const int new_hi = c == 32 ? hi : min(lo + step * (c + 1) - 1, hi - 1);
        lo = new_lo;
        hi = max(new_hi, new_lo);
    }
    // final: each lane tests one position
    const int pos = lo + lane;
    const bool less = pos < hi && key_of(items, pos) < key;
    return lo + __popc(__ballot_sync(kFull, less));
}

constexpr int kFinalKeys = 256;    // offsets written per CTA (one window search per thread)

// Roles by blockIdx.x: [0, nb_row) row_ptr chunks, [nb_row, nb_row+nb_pix) pix_ptr chunks, the rest
// gather the payloads of kFinalKeys sorted entries each.
__global__ void __launch_bounds__(kThreads) shpl_finalize_kernel(FinalArgs a, int nb_row, int nb_pix) {
    __shared__ int s_lb[2];
    __shared__ int s_off[kThreads];
    static_assert(kFinalKeys == kThreads, "one offset per thread");
    const int n = a.counts[1];
    const int ebase = a.entry_base_dev ? *a.entry_base_dev : 0;
    const unsigned long long* by_row = a.ws.items[0][a.sp[0].passes & 1];
    const unsigned long long* by_pix = a.ws.items[1][a.sp[1].passes & 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = blockIdx.x;
    if (b < nb_row + nb_pix) {
        const bool is_row = b < nb_row;
        if (!is_row) b -= nb_row;
        const unsigned long long* items = is_row ? by_row : by_pix;
        const int n_keys = is_row ? a.plan.n_rows : a.plan.n_src;
        int* ptr = is_row ? a.plan.row_ptr : a.plan.pix_ptr;
        const int j0 = b * kFinalKeys;
        const int j1 = min(j0 + kFinalKeys, n_keys + 1);          // offsets [j0, j1) belong to this CTA
        if (warp < 2) {
            const int lb = warp_lower_bound(items, n, (unsigned)(warp == 0 ? j0 : j1), lane);
            if (lane == 0) s_lb[warp] = lb;
        }
        __syncthreads();
        const int w0 = s_lb[0], w1 = s_lb[1];
        const int j = j0 + threadIdx.x;          // kFinalKeys == kThreads: one offset per thread
        int lb = w1;
        if (j < j1) {
            lb = lower_bound_key(items, w0, w1, (unsigned)j);
            SHPL_DASSERT(lb >= 0 && lb <= n);
            ptr[j] = ebase + lb;
            if (is_row && j == n_keys) {
                a.counts[2] = n - lb;
                a.counts[3] = lb;
                a.counts[4] = ebase + lb;     // entry base of the next stacked frame
            }
        }
        // heavy cells (more than SHPL_HEAVY_LEN entries) are listed for shpl_pool_heavy
        s_off[threadIdx.x] = lb;
        __syncthreads();
        {   // cells with more than SHPL_LONG_LEN entries, counted: a caller that reads 0 back may tell the pooling entry points
            // (heavy_len = 0) that no cell needs the long-cell paths
            const int next = (threadIdx.x + 1 < kThreads && j + 1 < j1) ? s_off[threadIdx.x + 1] : w1;
            const bool is_long = j < j1 && j < n_keys && next - lb > SHPL_LONG_LEN;
            const int n_long = __syncthreads_count(is_long);
            if (threadIdx.x == 0 && n_long > 0) atomicAdd(a.counts + (is_row ? 6 : 7), n_long);
        }
        if (a.plan.heavy_cap > 0 && j < j1 && j < n_keys) {
            const int next = (threadIdx.x + 1 < kThreads && j + 1 < j1) ? s_off[threadIdx.x + 1] : w1;
            if (next - lb > SHPL_HEAVY_LEN) {
                const int slot = atomicAdd(a.plan.heavy_count + (is_row ? 0 : 1), 1);
                if (slot < a.plan.heavy_cap) (is_row ? a.plan.heavy_row : a.plan.heavy_pix)[slot] = j + (is_row ? a.row_base : a.pix_base);
            }
        }
        return;
    }
    b -= nb_row + nb_pix;
    for (int j = b * kFinalKeys + threadIdx.x; j < min((b + 1) * kFinalKeys, n); j += kThreads) {
        if (ebase + j >= a.plan.capacity) {       // stacked frames: the host only knows capacity >= this frame's n
            a.counts[5] = 1;                      // nothing is written out of bounds; the caller sees the flag
            continue;
        }
        const unsigned long long ir = by_row[j];
        if ((unsigned)(ir >> 32) < (unsigned)a.plan.n_rows) {
            const unsigned k = (unsigned)ir;
            SHPL_DASSERT((int)k < n && a.ws.rowk[k] >= a.row_base && a.ws.rowk[k] < a.row_base + a.plan.n_rows &&
                         a.ws.pixk[k] >= a.pix_base && a.ws.pixk[k] < a.pix_base + a.plan.n_src);
            a.plan.csr_row[ebase + j] = a.ws.rowk[k];
            a.plan.csr_src[ebase + j] = a.ws.pixk[k];
            a.plan.csr_val[ebase + j] = a.ws.valk[k];
        }
        const unsigned long long ip = by_pix[j];
        if ((unsigned)(ip >> 32) < (unsigned)a.plan.n_src) {
            const unsigned k = (unsigned)ip;
            a.plan.csrT_pix[ebase + j] = a.ws.pixk[k];
            a.plan.csrT_dst[ebase + j] = a.ws.rowk[k];
            a.plan.csrT_val[ebase + j] = a.ws.valk[k];
        }
    }
}

int floordiv_host(int a, int s) { return (int)floor((double)a / (double)s); }

int run_build(PairsArgs& pa, const shpl_plan* plan, const int32_t* entry_base_dev, void* workspace,
              size_t workspace_bytes, void* stream, const char* who) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool sorting = pa.mode != kModeGenOnly;
    SHPL_REQUIRE(pa.n >= 0 && pa.n < (1ll << 30), SHPL_ERR_INVALID_ARGUMENT, "%s: n=%lld out of range", who, pa.n);
    SHPL_REQUIRE(workspace != nullptr && shpl::aligned(workspace, 64), SHPL_ERR_INVALID_ARGUMENT,
                 "%s: workspace must be a 64-byte aligned device buffer", who);
    SHPL_REQUIRE(pa.counts != nullptr, SHPL_ERR_INVALID_ARGUMENT, "%s: counts is null", who);
    if (sorting) {
        SHPL_REQUIRE(plan && plan->row_ptr && plan->pix_ptr && plan->csr_src && plan->csr_val && plan->csrT_dst &&
                         plan->csrT_val && plan->csr_row && plan->csrT_pix, SHPL_ERR_INVALID_ARGUMENT, "%s: plan has a null array", who);
        SHPL_REQUIRE(plan->n_rows >= 0 && plan->n_src >= 0 && plan->capacity >= pa.n, SHPL_ERR_INVALID_ARGUMENT,
                     "%s: plan capacity %d < %lld candidates", who, plan->capacity, pa.n);
        pa.sp[0] = make_sort_plan(plan->n_rows);
        pa.sp[1] = make_sort_plan(plan->n_src);
        SHPL_REQUIRE(plan->heavy_cap == 0 || (plan->heavy_row && plan->heavy_pix && plan->heavy_count),
                     SHPL_ERR_INVALID_ARGUMENT, "%s: heavy_cap > 0 but a heavy array is null", who);
        pa.heavy_count = plan->heavy_cap > 0 ? plan->heavy_count : nullptr;
    } else {
        pa.sp[0] = make_sort_plan(1);
        pa.sp[1] = make_sort_plan(1);
    }
    pa.ws = carve(workspace, pa.n, pa.sp);
    SHPL_REQUIRE(pa.ws.total_bytes <= workspace_bytes, SHPL_ERR_WORKSPACE_TOO_SMALL,
                 "%s: workspace %zu bytes < %zu needed", who, workspace_bytes, pa.ws.total_bytes);
    SHPL_CUDA_OK(cudaMemsetAsync(workspace, 0, pa.ws.header_bytes, s));
    const long long tiles = (pa.n + kTile - 1) / kTile > 0 ? (pa.n + kTile - 1) / kTile : 1;
    const long long ptiles = (pa.n + kPairsTile - 1) / kPairsTile > 0 ? (pa.n + kPairsTile - 1) / kPairsTile : 1;
    // The look-back spins on earlier tiles, so tile ids are ALWAYS handed out by arrival (one atomicAdd per CTA on a
    // header word the memset above cleared): CUDA gives no blockIdx-order dispatch guarantee, and these kernels run on
    // side streams beside pooling kernels that fill the SMs -- a spinning high tile must never wait for an unscheduled low one.
    const int use_ticket = 1;
    shpl_pairs_kernel<<<(unsigned)ptiles, kThreads, 0, s>>>(pa, use_ticket);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_pairs_kernel")) return rc;
    if (!sorting) return SHPL_OK;
    RadixArgs ra{};
    ra.n_dev = pa.counts + 1;
    ra.ws = pa.ws;
    ra.sp[0] = pa.sp[0];
    ra.sp[1] = pa.sp[1];
    const int passes = pa.sp[0].passes > pa.sp[1].passes ? pa.sp[0].passes : pa.sp[1].passes;
    for (int q = 0; q < passes && pa.n > 0; ++q) {
        ra.pass = q;
        shpl_radix_pass_kernel<<<dim3((unsigned)tiles, 2), kThreads, 0, s>>>(ra);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_radix_pass_kernel")) return rc;
    }
    FinalArgs fa{};
    fa.counts = pa.counts;
    fa.ws = pa.ws;
    fa.sp[0] = pa.sp[0];
    fa.sp[1] = pa.sp[1];
    fa.plan = *plan;
    fa.entry_base_dev = entry_base_dev;
    fa.row_base = pa.row_base;
    fa.pix_base = pa.pix_base;
    const int nb_row = (plan->n_rows + 1 + kFinalKeys - 1) / kFinalKeys;
    const int nb_pix = (plan->n_src + 1 + kFinalKeys - 1) / kFinalKeys;
    const int nb_ent = (int)((pa.n + kFinalKeys - 1) / kFinalKeys);
    shpl_finalize_kernel<<<(unsigned)(nb_row + nb_pix + nb_ent), kThreads, 0, s>>>(fa, nb_row, nb_pix);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_finalize_kernel");
}

int fill_geometry(PairsArgs& pa, int im_w, int im_h, int bv_h, int bv_w, int s_img, int s_bv, int src_h, int src_w,
                  const shpl_plan* plan, const char* who) {
    SHPL_REQUIRE(s_img > 0 && s_bv > 0, SHPL_ERR_INVALID_ARGUMENT, "%s: strides must be positive integers", who);
    pa.im_w = im_w;
    pa.im_h = im_h;
    pa.s_img = s_img;
    pa.s_bv = s_bv;
    pa.Wp = floordiv_host(im_w, s_img);
    pa.Hp = floordiv_host(im_h, s_img);
    pa.Hb = floordiv_host(bv_h, s_bv);
    pa.Wb = floordiv_host(bv_w, s_bv);
    pa.src_h = src_h > 0 ? src_h : pa.Hp;
    pa.src_w = src_w > 0 ? src_w : pa.Wp;
    if (plan) {
        SHPL_REQUIRE((long long)pa.Hb * pa.Wb == plan->n_rows, SHPL_ERR_INVALID_ARGUMENT,
                     "%s: plan.n_rows=%d but floor(bv/stride) gives %d x %d cells", who, plan->n_rows, pa.Hb, pa.Wb);
        SHPL_REQUIRE((long long)pa.src_h * pa.src_w == plan->n_src, SHPL_ERR_INVALID_ARGUMENT,
                     "%s: plan.n_src=%d but the source map is %d x %d", who, plan->n_src, pa.src_h, pa.src_w);
        pa.n_rows = plan->n_rows;
    }
    return SHPL_OK;
}

}  // namespace

extern "C" size_t shpl_build_workspace_bytes(int64_t n_max) {
    if (n_max < 0) n_max = 0;
    return carve(nullptr, n_max, nullptr).total_bytes;
}

extern "C" int shpl_gen_input_avod(const double* points, const int64_t* voxel_indices, int64_t N, const int32_t* N_dev,
                                   const double* P_host,
                                   int32_t im_w, int32_t im_h, int64_t* bv_index_out, double* img_u_out,
                                   double* img_v_out, int32_t* counts, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    const char* who = "shpl_gen_input_avod";
    SHPL_REQUIRE(P_host && (N == 0 || (points && voxel_indices)) && bv_index_out && img_u_out && img_v_out,
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    PairsArgs pa{};
    pa.mode = kModeGenOnly;
    pa.n = N;
    pa.n_dev = N_dev;
    pa.points = points;
    pa.vox = reinterpret_cast<const long long*>(voxel_indices);
    for (int i = 0; i < 12; ++i) pa.P[i] = P_host[i];
    pa.im_w = im_w;
    pa.im_h = im_h;
    pa.s_img = pa.s_bv = 1;
    pa.gen_bv = reinterpret_cast<long long*>(bv_index_out);
    pa.gen_u = img_u_out;
    pa.gen_v = img_v_out;
    pa.counts = counts;
    return run_build(pa, nullptr, nullptr, workspace, workspace_bytes, stream, who);
}

extern "C" int shpl_produce_input(double* img_u, double* img_v, const int64_t* bv_index, int64_t n, int32_t im_w,
                                  int32_t im_h, int32_t bv_h, int32_t bv_w, int32_t stride_img, int32_t stride_bv,
                                  const double* m_val, int32_t src_h, int32_t src_w, int64_t* Mij_pool,
                                  int64_t* img_index_flip_pool, float* M_val_out, int64_t* M_size_out,
                                  const shpl_plan* plan, int32_t row_base, int32_t pix_base,
                                  const int32_t* entry_base_dev, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    const char* who = "shpl_produce_input";
    SHPL_REQUIRE(plan && (n == 0 || (img_u && img_v && bv_index)), SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    PairsArgs pa{};
    pa.mode = kModePairs;
    pa.n = n;
    pa.img_u = img_u;
    pa.img_v = img_v;
    pa.bv_index = reinterpret_cast<const long long*>(bv_index);
    pa.m_val = m_val;
    if (int rc = fill_geometry(pa, im_w, im_h, bv_h, bv_w, stride_img, stride_bv, src_h, src_w, plan, who)) return rc;
    pa.row_base = row_base;
    pa.pix_base = pix_base;
    pa.mij = reinterpret_cast<long long*>(Mij_pool);
    pa.flip = reinterpret_cast<long long*>(img_index_flip_pool);
    pa.mval_out = M_val_out;
    pa.msize_out = reinterpret_cast<long long*>(M_size_out);
    pa.counts = plan->counts;
    return run_build(pa, plan, entry_base_dev, workspace, workspace_bytes, stream, who);
}

extern "C" int shpl_build_avod(const double* points, const int64_t* voxel_indices, int64_t N, const int32_t* N_dev,
                               const double* P_host,
                               int32_t im_w, int32_t im_h, int32_t bv_h, int32_t bv_w, int32_t stride_img,
                               int32_t stride_bv, const double* m_val, int32_t src_h, int32_t src_w, int64_t* Mij_pool,
                               int64_t* img_index_flip_pool, float* M_val_out, int64_t* M_size_out,
                               const shpl_plan* plan, int32_t row_base, int32_t pix_base,
                               const int32_t* entry_base_dev, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "shpl_build_avod";
    SHPL_REQUIRE(plan && P_host && (N == 0 || (points && voxel_indices)), SHPL_ERR_INVALID_ARGUMENT,
                 "%s: null pointer", who);
    PairsArgs pa{};
    pa.mode = kModeAvod;
    pa.n = N;
    pa.n_dev = N_dev;
    pa.points = points;
    pa.vox = reinterpret_cast<const long long*>(voxel_indices);
    for (int i = 0; i < 12; ++i) pa.P[i] = P_host[i];
    pa.m_val = m_val;
    if (int rc = fill_geometry(pa, im_w, im_h, bv_h, bv_w, stride_img, stride_bv, src_h, src_w, plan, who)) return rc;
    pa.row_base = row_base;
    pa.pix_base = pix_base;
    pa.mij = reinterpret_cast<long long*>(Mij_pool);
    pa.flip = reinterpret_cast<long long*>(img_index_flip_pool);
    pa.mval_out = M_val_out;
    pa.msize_out = reinterpret_cast<long long*>(M_size_out);
    pa.counts = plan->counts;
    return run_build(pa, plan, entry_base_dev, workspace, workspace_bytes, stream, who);
}

extern "C" int shpl_plan_from_coo(const int64_t* Mij, const float* val, int64_t m, const void* source_index,
                                  int32_t index_is_i64, int64_t ncol, int32_t src_h, int32_t src_w,
                                  const shpl_plan* plan, int32_t row_base, int32_t pix_base,
                                  const int32_t* entry_base_dev, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    const char* who = "shpl_plan_from_coo";
    SHPL_REQUIRE(plan && (m == 0 || (Mij && val)) && (ncol == 0 || source_index) && ncol >= 0,
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE((long long)src_h * src_w == plan->n_src, SHPL_ERR_INVALID_ARGUMENT,
                 "%s: plan.n_src=%d but the source map is %d x %d", who, plan->n_src, src_h, src_w);
    PairsArgs pa{};
    pa.mode = kModeCoo;
    pa.n = m;
    pa.coo = reinterpret_cast<const long long*>(Mij);
    pa.coo_val = val;
    pa.src_index = source_index;
    pa.index_is_i64 = index_is_i64;
    pa.ncol = ncol;
    pa.s_img = pa.s_bv = 1;
    pa.Hb = 1;
    pa.Wb = plan->n_rows;
    pa.n_rows = plan->n_rows;
    pa.src_h = src_h;
    pa.src_w = src_w;
    pa.row_base = row_base;
    pa.pix_base = pix_base;
    pa.counts = plan->counts;
    return run_build(pa, plan, entry_base_dev, workspace, workspace_bytes, stream, who);
}

extern "C" int shpl_plan_from_voxel_coords(const void* coordinate, int32_t index_is_i64, int64_t K, const int32_t* K_dev,
                                           int32_t batch, int32_t depth, int32_t height, int32_t width,
                                           const shpl_plan* plan, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "shpl_plan_from_voxel_coords";
    SHPL_REQUIRE(plan && (K == 0 || coordinate), SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(shpl::aligned(coordinate, 16), SHPL_ERR_INVALID_ARGUMENT, "%s: coordinate must be 16-byte aligned", who);
    SHPL_REQUIRE(batch > 0 && depth > 0 && height > 0 && width > 0 &&
                     (long long)batch * depth * height * width == plan->n_rows, SHPL_ERR_INVALID_ARGUMENT,
                 "%s: plan.n_rows=%d but the grid is %d x %d x %d x %d", who, plan->n_rows, batch, depth, height, width);
    SHPL_REQUIRE(plan->n_src >= K, SHPL_ERR_INVALID_ARGUMENT, "%s: plan.n_src=%d < %lld feature rows", who, plan->n_src,
                 (long long)K);
    PairsArgs pa{};
    pa.mode = kModeVoxel;
    pa.n = K;
    pa.n_dev = K_dev;
    pa.src_index = coordinate;
    pa.index_is_i64 = index_is_i64;
    pa.grid[0] = batch;
    pa.grid[1] = depth;
    pa.grid[2] = height;
    pa.grid[3] = width;
    pa.s_img = pa.s_bv = 1;
    pa.Hb = 1;
    pa.Wb = plan->n_rows;
    pa.n_rows = plan->n_rows;
    pa.src_h = plan->n_src;
    pa.src_w = 1;
    pa.counts = plan->counts;
    return run_build(pa, plan, nullptr, workspace, workspace_bytes, stream, who);
}
