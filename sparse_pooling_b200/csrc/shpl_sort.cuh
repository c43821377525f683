// Stable LSD radix sort of (key << 32 | index) items for libshpl.so (sm_100a): the sort plan, the scratch layout
// and the pass kernel shared by the correspondence builder (shpl_build.cu: by destination cell and by source
// pixel) and the MV3D voxel feeder (shpl_mv3d.cu: by voxel).  Internal header.
#pragma once
#include "shpl_common.cuh"

namespace {
using shpl::kFull;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
#ifndef SHPL_SORT_CHUNKS
#define SHPL_SORT_CHUNKS 8
#endif
constexpr int kChunks = SHPL_SORT_CHUNKS;        // 32-wide chunks per warp in a radix tile
constexpr int kTile = kThreads * kChunks;        // 2048 sort items per radix CTA
constexpr int kPairsTile = kThreads;             // one candidate pair per thread in shpl_pairs_kernel
constexpr int kMaxRadixBits = 10;
constexpr int kMaxRadix = 1 << kMaxRadixBits;
constexpr int kMaxPasses = 4;

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
    const long long ptiles = (n + kPairsTile - 1) / kPairsTile > 0 ? (n + kPairsTile - 1) / kPairsTile : 1;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned*>(p + off);
    off += 64;
    w.status = reinterpret_cast<unsigned long long*>(p + off);
    off = align_up(off + sizeof(unsigned long long) * ptiles, 64);
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

struct RadixArgs {
    const int* n_dev;             // device: number of items
    Workspace ws;
    SortPlan sp[2];
    int pass;
};

// One stable LSD pass over (key<<32 | k) items.  grid.y selects the sort (0 = by cell, 1 = by pixel).
// The tile is staged in shared memory with every load in flight at once; ranks come from warp
// match-any over 32-item chunks taken in order (stability); the tile's base per digit comes from the
// per-tile digit counts the previous kernel accumulated.
__global__ void __launch_bounds__(kThreads) shpl_radix_pass_kernel(RadixArgs a) {
    __shared__ unsigned long long s_items[kTile];
    __shared__ __align__(16) unsigned short s_wh[kWarps][kMaxRadix];   // per-warp digit counts, then exclusive warp prefix
    __shared__ unsigned short s_rank[kTile];
    __shared__ unsigned s_gb[kMaxRadix];                 // global base of each digit for this tile
    __shared__ unsigned s_scan[kWarps];
    const int sort = blockIdx.y;
    const SortPlan& sp = a.sp[sort];
    const int pass = a.pass;
    if (pass >= sp.passes) return;
    const int n = *a.n_dev;
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
    constexpr int kDigitsPerThread = kMaxRadix / kThreads;

    // stage the tile (coalesced, all loads in flight) and clear the warp histograms
#pragma unroll
    for (int j = 0; j < kChunks; ++j) {
        const int i = tile_base + j * kThreads + threadIdx.x;
        s_items[j * kThreads + threadIdx.x] = i < n ? in[i] : ~0ull;
    }
    {   // clear the per-warp digit counters (16 KB) with 128-bit stores
        uint4* z = reinterpret_cast<uint4*>(&s_wh[0][0]);
        constexpr int kVec = (int)(sizeof(s_wh) / sizeof(uint4));
        for (int d = threadIdx.x; d < kVec; d += kThreads) z[d] = make_uint4(0u, 0u, 0u, 0u);
    }

    // this thread's digits: how many such items sit in earlier tiles / in all tiles (loads overlap phase A)
    unsigned below[kDigitsPerThread], all[kDigitsPerThread];
#pragma unroll
    for (int j = 0; j < kDigitsPerThread; ++j) {
        const int d = threadIdx.x * kDigitsPerThread + j;   // consecutive digits per thread (for the scan below)
        below[j] = all[j] = 0;
        if (d < radix) {
            for (int t0 = 0; t0 < n_tiles; t0 += 8) {      // 8 independent loads in flight
                unsigned h[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) h[u] = (t0 + u < n_tiles) ? __ldg(hist + ((size_t)(t0 + u) << bits) + d) : 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    all[j] += h[u];
                    if (t0 + u < tile) below[j] += h[u];
                }
            }
        }
    }
    __syncthreads();

    // phase A: rank of every item inside (warp, digit); a warp owns 256 consecutive items, in order
    const int wofs = warp * (kChunks * 32);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
        const int li = wofs + c * 32 + lane;
        const bool live = tile_base + li < n;
        const unsigned long long it = s_items[li];
        const unsigned digit = live ? (unsigned)((it >> (32 + shift)) & mask) : (unsigned)radix + lane;
        const unsigned peers = __match_any_sync(kFull, digit);
        const int leader = __ffs(peers) - 1;
        unsigned prev = 0;
        if (live && lane == leader) {
            prev = s_wh[warp][digit];
            s_wh[warp][digit] = (unsigned short)(prev + __popc(peers));
        }
        prev = __shfl_sync(kFull, prev, leader);
        s_rank[li] = (unsigned short)(prev + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // phase B: exclusive prefix over warps per digit, exclusive scan of the digit totals over the block
    unsigned local_sum = 0;
#pragma unroll
    for (int j = 0; j < kDigitsPerThread; ++j) {
        const int d = threadIdx.x * kDigitsPerThread + j;
        if (d < radix) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned t = s_wh[w][d];
                s_wh[w][d] = (unsigned short)run;
                run += t;
            }
        }
        local_sum += all[j];
    }
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
    for (int j = 0; j < kDigitsPerThread; ++j) {
        const int d = threadIdx.x * kDigitsPerThread + j;
        if (d < radix) s_gb[d] = run + below[j];
        run += all[j];
    }
    __syncthreads();

    // phase C: scatter, and count the next pass's digit per destination tile
    const bool more = pass + 1 < sp.passes;
    const int nbits = more ? sp.bits[pass + 1] : 0, nshift = more ? sp.shift[pass + 1] : 0;
    unsigned* nhist = more ? a.ws.hist[sort][pass + 1] : nullptr;
#pragma unroll 2
    for (int c = 0; c < kChunks; ++c) {
        const int li = wofs + c * 32 + lane;
        if (tile_base + li < n) {
            const unsigned long long it = s_items[li];
            const unsigned digit = (unsigned)((it >> (32 + shift)) & mask);
            const unsigned dest = s_gb[digit] + s_wh[warp][digit] + s_rank[li];
            out[dest] = it;
            if (more) {
                const unsigned nd = (unsigned)((it >> (32 + nshift)) & ((1u << nbits) - 1u));
                atomicAdd(nhist + ((size_t)(dest / kTile) << nbits) + nd, 1u);
            }
        }
    }
}

}  // namespace
