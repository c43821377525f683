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

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunks = 8;                       // 32-wide chunks per warp
constexpr int kTile = kThreads * kChunks;        // 2048 candidates (or sort items) per CTA
constexpr int kMaxRadixBits = 10;
constexpr int kMaxRadix = 1 << kMaxRadixBits;
constexpr int kMaxPasses = 4;
constexpr unsigned kFull = 0xffffffffu;

enum Mode { kModeAvod = 0, kModePairs = 1, kModeCoo = 2, kModeGenOnly = 3 };

struct SortPlan {
    int n_keys;               // keys are in [0, n_keys]; n_keys itself is the sentinel
    int passes;
    int shift[kMaxPasses];
    int bits[kMaxPasses];
};

struct Workspace {           // device pointers carved out of the caller's scratch
    unsigned* ticket;                         // [1]
    unsigned long long* status;               // [tiles]
    unsigned* hist[2][kMaxPasses];            // [tiles][1<<bits]
    unsigned long long* items[2][2];          // [sort][ping/pong][n]
    int* rowk;                                // [n] destination row (+row_base) of column k, -1 if invalid
    int* pixk;                                // [n] source pixel (+pix_base) of column k, -1 if invalid
    float* valk;                              // [n]
    size_t header_bytes;                      // zeroed at the start of every call
    size_t total_bytes;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Workspace carve(void* base, long long n, const SortPlan* sp /* [2] or null for worst case */) {
    Workspace w{};
    const long long tiles = (n + kTile - 1) / kTile > 0 ? (n + kTile - 1) / kTile : 1;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned*>(p + off);
    off += 64;
    w.status = reinterpret_cast<unsigned long long*>(p + off);
    off = align_up(off + sizeof(unsigned long long) * tiles, 64);
    for (int s = 0; s < 2; ++s)
        for (int q = 0; q < kMaxPasses; ++q) {
            const int radix = sp ? (q < sp[s].passes ? (1 << sp[s].bits[q]) : 0) : kMaxRadix;
            w.hist[s][q] = reinterpret_cast<unsigned*>(p + off);
            off = align_up(off + sizeof(unsigned) * tiles * radix, 64);
        }
    w.header_bytes = off;
    for (int s = 0; s < 2; ++s)
        for (int b = 0; b < 2; ++b) {
            w.items[s][b] = reinterpret_cast<unsigned long long*>(p + off);
            off = align_up(off + sizeof(unsigned long long) * (n > 0 ? n : 1), 64);
        }
    w.rowk = reinterpret_cast<int*>(p + off);
    off = align_up(off + sizeof(int) * (n > 0 ? n : 1), 64);
    w.pixk = reinterpret_cast<int*>(p + off);
    off = align_up(off + sizeof(int) * (n > 0 ? n : 1), 64);
    w.valk = reinterpret_cast<float*>(p + off);
    off = align_up(off + sizeof(float) * (n > 0 ? n : 1), 64);
    w.total_bytes = off;
    return w;
}

SortPlan make_sort_plan(int n_keys) {
    SortPlan sp{};
    sp.n_keys = n_keys;
    int total_bits = 1;
    while ((1ll << total_bits) <= (long long)n_keys) ++total_bits;   // keys up to and including n_keys
    sp.passes = (total_bits + kMaxRadixBits - 1) / kMaxRadixBits;
    const int per = (total_bits + sp.passes - 1) / sp.passes;
    int shift = 0;
    for (int q = 0; q < sp.passes; ++q) {
        sp.shift[q] = shift;
        sp.bits[q] = (total_bits - shift) < per ? (total_bits - shift) : per;
        shift += sp.bits[q];
    }
    return sp;
}

struct PairsArgs {
    int mode;
    long long n;                 // candidates
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
    Workspace ws;
    SortPlan sp[2];
};

__device__ __forceinline__ long long floordiv_ll(long long a, long long s) {
    long long q = a / s;
    if ((a % s != 0) && ((a < 0) != (s < 0))) --q;
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

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// Single-pass exclusive prefix over tiles (decoupled look-back), executed by warp 0.
// `agg` packs (count_clip << 31 | count_nnz) of this tile; returns the packed sum of all earlier tiles.
__device__ unsigned long long lookback(unsigned long long* status, int tile, unsigned long long agg, int lane) {
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

__global__ void __launch_bounds__(kThreads) shpl_pairs_kernel(PairsArgs a) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp_clip[kWarps], s_warp_keep[kWarps];
    __shared__ unsigned long long s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ws.ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const long long base = (long long)tile * kTile + warp * (kChunks * 32);
    const long long R = (long long)a.Hb * a.Wb;

    // per candidate: flags and the values that survive to the outputs
    bool clip[kChunks], keep[kChunks];
    long long row[kChunks];
    int up[kChunks], vp[kChunks];
    double gu[kChunks], gv[kChunks];
    long long bx[kChunks], bz[kChunks];
    unsigned pre_clip[kChunks], pre_keep[kChunks];
    unsigned run_clip = 0, run_keep = 0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const long long i = base + c * 32 + lane;
        clip[c] = false; keep[c] = false; row[c] = 0; up[c] = 0; vp[c] = 0; gu[c] = 0; gv[c] = 0; bx[c] = 0; bz[c] = 0;
        if (i < a.n) {
            if (a.mode == kModeCoo) {
                const long long r = a.coo[2 * i], col = a.coo[2 * i + 1];
                clip[c] = keep[c] = true;
                row[c] = r;
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
                vp[c] = ok ? (int)v : -1;
                up[c] = ok ? (int)u : -1;
            } else {
                double u, v;
                if (a.mode == kModePairs) {
                    u = a.img_u[i];
                    v = a.img_v[i];
                    bx[c] = a.bv_index[2 * i];
                    bz[c] = a.bv_index[2 * i + 1];
                    clip[c] = true;
                } else {
                    const double x = a.points[3 * i], y = a.points[3 * i + 1], z = a.points[3 * i + 2];
                    const double w = prow(a.P + 8, x, y, z);
                    u = __ddiv_rn(prow(a.P + 0, x, y, z), w);
                    v = __ddiv_rn(prow(a.P + 4, x, y, z), w);
                    // transform.py:36-38 (NaN compares false)
                    clip[c] = (u < (double)(a.im_w - 1)) && (u >= 0.0) && (v >= 0.0) && (v < (double)(a.im_h - 1));
                    u = rint(u);   // sparse_pool_utils.py:18, ties to even
                    v = rint(v);
                    bx[c] = a.vox[2 * i];
                    bz[c] = a.vox[2 * i + 1];
                }
                gu[c] = u;
                gv[c] = v;
                if (clip[c] && a.mode != kModeGenOnly) {
                    // sparse_pool_utils.py:30-34
                    double us = floor(u / (double)a.s_img), vs = floor(v / (double)a.s_img);
                    if (us >= (double)a.Wp) us = (double)(a.Wp - 1);
                    if (vs >= (double)a.Hp) vs = (double)(a.Hp - 1);
                    if (a.mode == kModePairs) {     // the reference mutates img_index in place (:30)
                        a.img_u[i] = us;
                        a.img_v[i] = vs;
                    }
                    const long long ul = (long long)floor(us), vl = (long long)floor(vs);
                    // :38-44
                    const long long xs = floordiv_ll(bx[c], a.s_bv), zs = floordiv_ll(bz[c], a.s_bv);
                    row[c] = zs * (long long)a.Wb + xs;
                    keep[c] = row[c] < R;
                    const bool ok = vl >= 0 && vl < a.src_h && ul >= 0 && ul < a.src_w;
                    // keep the exact integers for the COO output, clamp what feeds the CSR
                    up[c] = (int)max(min(ul, (long long)INT32_MAX), (long long)INT32_MIN);
                    vp[c] = (int)max(min(vl, (long long)INT32_MAX), (long long)INT32_MIN);
                    if (!ok) { /* marked below through pix = -1 */ }
                }
            }
        }
        const unsigned mc = __ballot_sync(kFull, clip[c]);
        const unsigned mk = __ballot_sync(kFull, keep[c]);
        const unsigned lt = (1u << lane) - 1u;
        pre_clip[c] = run_clip + __popc(mc & lt);
        pre_keep[c] = run_keep + __popc(mk & lt);
        run_clip += __popc(mc);
        run_keep += __popc(mk);
    }
    if (lane == 0) {
        s_warp_clip[warp] = run_clip;
        s_warp_keep[warp] = run_keep;
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
        if (a.mode == kModeCoo) excl = ((unsigned long long)tile * kTile) * ((1ull << 31) + 1ull);
        else excl = lookback(a.ws.status, tile, ((unsigned long long)tot_clip << 31) | tot_keep, lane);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const long long g_clip = (long long)(s_excl >> 31) + wb_clip;
    const long long g_keep = (long long)(s_excl & ((1ull << 31) - 1)) + wb_keep;

    const long long n_tiles = (a.n + kTile - 1) / kTile > 0 ? (a.n + kTile - 1) / kTile : 1;
    if (tile == n_tiles - 1 && threadIdx.x == 0) {   // inclusive prefix of the last tile = totals
        const long long n_clip = (long long)(s_excl >> 31) + tot_clip;
        const long long nnz = (long long)(s_excl & ((1ull << 31) - 1)) + tot_keep;
        a.counts[0] = (int)n_clip;
        a.counts[1] = (a.mode == kModeGenOnly) ? 0 : (int)nnz;
        if (a.msize_out) { a.msize_out[0] = R; a.msize_out[1] = nnz; }
    }

#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        if (clip[c] && a.gen_bv) {
            const long long j = g_clip + pre_clip[c];
            a.gen_bv[2 * j] = bx[c];
            a.gen_bv[2 * j + 1] = bz[c];
            a.gen_u[j] = gu[c];
            a.gen_v[j] = gv[c];
        }
        if (!keep[c]) continue;
        const long long k = g_keep + pre_keep[c];
        const long long r = row[c];
        float w = 1.0f;
        if (a.mode == kModeCoo) w = a.coo_val[k];
        else if (a.m_val) w = (float)a.m_val[k];          // indexed by output column (:56-57), f64 -> f32 like the feed
        const bool pix_ok = vp[c] >= 0 && vp[c] < a.src_h && up[c] >= 0 && up[c] < a.src_w;
        const bool ok = pix_ok && r >= 0 && r < (long long)a.n_rows;
        const int pix = vp[c] * a.src_w + up[c];
        if (a.mij) { a.mij[2 * k] = r; a.mij[2 * k + 1] = k; }
        if (a.flip) { a.flip[3 * k] = 0; a.flip[3 * k + 1] = vp[c]; a.flip[3 * k + 2] = up[c]; }
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
}

struct RadixArgs {
    const int* counts;            // counts[1] = number of items
    Workspace ws;
    SortPlan sp[2];
    int pass;
};

// One stable LSD pass over (key<<32 | k) items.  grid.y selects the sort (0 = by cell, 1 = by pixel).
__global__ void __launch_bounds__(kThreads) shpl_radix_pass_kernel(RadixArgs a) {
    __shared__ unsigned short s_wh[kWarps][kMaxRadix];   // per-warp digit counts, then exclusive warp prefix
    __shared__ unsigned s_gb[kMaxRadix];                 // global base of each digit for this tile
    __shared__ unsigned s_scan[kWarps];
    const int sort = blockIdx.y;
    const SortPlan& sp = a.sp[sort];
    const int pass = a.pass;
    if (pass >= sp.passes) return;
    const int n = a.counts[1];
    const int tile = blockIdx.x;
    const int tile_base = tile * kTile;
    if (tile_base >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bits = sp.bits[pass], shift = sp.shift[pass];
    const int radix = 1 << bits;
    const unsigned mask = radix - 1;
    const unsigned long long* in = a.ws.items[sort][pass & 1];
    unsigned long long* out = a.ws.items[sort][(pass + 1) & 1];
    const unsigned* hist = a.ws.hist[sort][pass];
    const int n_tiles = (n + kTile - 1) / kTile;

    for (int d = threadIdx.x; d < radix * kWarps; d += kThreads) (&s_wh[0][0])[(d / radix) * kMaxRadix + (d % radix)] = 0;
    __syncthreads();

    // phase A: rank of every item inside (warp, digit), chunks taken in order
    unsigned long long item[kChunks];
    unsigned short rank[kChunks];
    const int wbase = tile_base + warp * (kChunks * 32);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int i = wbase + c * 32 + lane;
        const bool live = i < n;
        item[c] = live ? in[i] : 0ull;
        const unsigned digit = live ? (unsigned)((item[c] >> (32 + shift)) & mask) : (unsigned)radix + lane;
        const unsigned peers = __match_any_sync(kFull, digit);
        const int leader = __ffs(peers) - 1;
        unsigned prev = 0;
        if (live && lane == leader) {
            prev = s_wh[warp][digit];
            s_wh[warp][digit] = (unsigned short)(prev + __popc(peers));
        }
        prev = __shfl_sync(kFull, prev, leader);
        rank[c] = (unsigned short)(prev + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // phase B: exclusive prefix over warps per digit; tile base of each digit from the global per-tile counts
    unsigned tot4[kMaxRadix / kThreads];
    unsigned local_sum = 0;
#pragma unroll
    for (int j = 0; j < kMaxRadix / kThreads; ++j) {
        const int d = threadIdx.x * (kMaxRadix / kThreads) + j;   // consecutive digits per thread
        tot4[j] = 0;
        if (d < radix) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned t = s_wh[w][d];
                s_wh[w][d] = (unsigned short)run;
                run += t;
            }
            unsigned below = 0, all = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const unsigned h = hist[((size_t)t << bits) + d];
                all += h;
                if (t < tile) below += h;
            }
            s_gb[d] = below;
            tot4[j] = all;
        }
        local_sum += tot4[j];
    }
    // block-wide exclusive scan of the digit totals (digits are blocked 4 per thread)
    unsigned incl = local_sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned wpre = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if (w < warp) wpre += s_scan[w];
    unsigned run = wpre + incl - local_sum;
#pragma unroll
    for (int j = 0; j < kMaxRadix / kThreads; ++j) {
        const int d = threadIdx.x * (kMaxRadix / kThreads) + j;
        if (d < radix) s_gb[d] += run;
        run += tot4[j];
    }
    __syncthreads();

    // phase C: scatter, and count the next pass's digit per destination tile
    const bool more = pass + 1 < sp.passes;
    const int nbits = more ? sp.bits[pass + 1] : 0, nshift = more ? sp.shift[pass + 1] : 0;
    unsigned* nhist = more ? a.ws.hist[sort][pass + 1] : nullptr;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int i = wbase + c * 32 + lane;
        if (i < n) {
            const unsigned digit = (unsigned)((item[c] >> (32 + shift)) & mask);
            const unsigned dest = s_gb[digit] + s_wh[warp][digit] + rank[c];
            out[dest] = item[c];
            if (more) {
                const unsigned nd = (unsigned)((item[c] >> (32 + nshift)) & ((1u << nbits) - 1u));
                atomicAdd(nhist + ((size_t)(dest / kTile) << nbits) + nd, 1u);
            }
        }
    }
}

struct FinalArgs {
    int* counts;
    Workspace ws;
    SortPlan sp[2];
    shpl_plan plan;
    const int* entry_base_dev;
};

__device__ __forceinline__ int lower_bound_key(const unsigned long long* items, int n, unsigned key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((unsigned)(__ldg(items + mid) >> 32) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads) shpl_finalize_kernel(FinalArgs a) {
    const int j = blockIdx.x * kThreads + threadIdx.x;
    const int n = a.counts[1];
    const int ebase = a.entry_base_dev ? *a.entry_base_dev : 0;
    const unsigned long long* by_row = a.ws.items[0][a.sp[0].passes & 1];
    const unsigned long long* by_pix = a.ws.items[1][a.sp[1].passes & 1];
    if (j <= a.plan.n_rows) {
        const int lb = lower_bound_key(by_row, n, (unsigned)j);
        a.plan.row_ptr[j] = ebase + lb;
        if (j == a.plan.n_rows) {
            a.counts[2] = n - lb;
            a.counts[3] = lb;
            a.counts[4] = ebase + lb;     // entry base of the next stacked frame
        }
    }
    if (j <= a.plan.n_src) a.plan.pix_ptr[j] = ebase + lower_bound_key(by_pix, n, (unsigned)j);
    if (j < n) {
        const unsigned long long ir = by_row[j];
        if ((unsigned)(ir >> 32) < (unsigned)a.plan.n_rows) {
            const unsigned k = (unsigned)ir;
            a.plan.csr_src[ebase + j] = a.ws.pixk[k];
            a.plan.csr_val[ebase + j] = a.ws.valk[k];
        }
        const unsigned long long ip = by_pix[j];
        if ((unsigned)(ip >> 32) < (unsigned)a.plan.n_src) {
            const unsigned k = (unsigned)ip;
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
                         plan->csrT_val, SHPL_ERR_INVALID_ARGUMENT, "%s: plan has a null array", who);
        SHPL_REQUIRE(plan->n_rows >= 0 && plan->n_src >= 0 && plan->capacity >= pa.n, SHPL_ERR_INVALID_ARGUMENT,
                     "%s: plan capacity %d < %lld candidates", who, plan->capacity, pa.n);
        pa.sp[0] = make_sort_plan(plan->n_rows);
        pa.sp[1] = make_sort_plan(plan->n_src);
    } else {
        pa.sp[0] = make_sort_plan(1);
        pa.sp[1] = make_sort_plan(1);
    }
    pa.ws = carve(workspace, pa.n, pa.sp);
    SHPL_REQUIRE(pa.ws.total_bytes <= workspace_bytes, SHPL_ERR_WORKSPACE_TOO_SMALL,
                 "%s: workspace %zu bytes < %zu needed", who, workspace_bytes, pa.ws.total_bytes);
    SHPL_CUDA_OK(cudaMemsetAsync(workspace, 0, pa.ws.header_bytes, s));
    const long long tiles = (pa.n + kTile - 1) / kTile > 0 ? (pa.n + kTile - 1) / kTile : 1;
    shpl_pairs_kernel<<<(unsigned)tiles, kThreads, 0, s>>>(pa);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_pairs_kernel")) return rc;
    if (!sorting) return SHPL_OK;
    RadixArgs ra{};
    ra.counts = pa.counts;
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
    long long span = plan->n_rows + 1;
    if (plan->n_src + 1 > span) span = plan->n_src + 1;
    if (pa.n > span) span = pa.n;
    shpl_finalize_kernel<<<(unsigned)((span + kThreads - 1) / kThreads), kThreads, 0, s>>>(fa);
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

extern "C" int shpl_gen_input_avod(const double* points, const int64_t* voxel_indices, int64_t N, const double* P_host,
                                   int32_t im_w, int32_t im_h, int64_t* bv_index_out, double* img_u_out,
                                   double* img_v_out, int32_t* counts, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    const char* who = "shpl_gen_input_avod";
    SHPL_REQUIRE(P_host && (N == 0 || (points && voxel_indices)) && bv_index_out && img_u_out && img_v_out,
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    PairsArgs pa{};
    pa.mode = kModeGenOnly;
    pa.n = N;
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

extern "C" int shpl_build_avod(const double* points, const int64_t* voxel_indices, int64_t N, const double* P_host,
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
