// BEV slicing / voxelisation feeder for sm_100a.
//
// Replaces the numpy step that produces the SHPL builder's inputs:
//   /root/reference/avod/avod/core/bev_generators/bev_slices.py:33-156        BevSlices.generate_bev(output_indices=True)
//   /root/reference/avod/avod/core/bev_generators/bev_generator.py:23-41      _create_density_map
//   /root/reference/avod/wavedata/wavedata/tools/core/voxel_grid_2d.py:43-162 VoxelGrid2D.voxelize_2d
//   /root/reference/avod/avod/datasets/kitti/kitti_utils.py:79-107            create_slice_filter
//   /root/reference/avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:444-491  get_point_filter
//   /root/reference/avod/wavedata/wavedata/tools/core/geometry_utils.py:25-40 dist_to_plane
//
// The reference sorts every slice (np.lexsort by x, z, y; stable) and keeps the first point of every
// (x, z) run.  "First" = smallest (floor(y/voxel), original index) of the cell, so no sort is needed:
//   memset   header + occupancy bitmaps + density counts = 0 (one memset), winner grid = ~0, maps = 0
//   K1       shpl_bev_scatter_kernel: one point per thread -- extents + plane filters (fp64, the reference's
//            comparisons), floor(p / voxel), atomicMin of (y_disc, index) into the slice's cell of the winner
//            grid, atomicOr of the cell's bit in the slice's occupancy bitmap, integer atomicAdd for the
//            density band (all order-independent: deterministic)
//   K2       shpl_bev_emit_kernel: walks the BITMAPS (420 KB instead of the 22 MB grid) in the reference's
//            output order (slice, x, z): popc + block scan + decoupled look-back give every occupied cell its
//            output row (stable compaction); the tile's occupied cells are listed in shared memory and then
//            spread over the CTA's threads, so crowded near-range words do not serialise one thread.
//            Emits voxel_indices (x, Z - z), unique_pts and the non-zeros of the height / density maps.
// The winner grid uses 32-bit words (y bin | index) whenever the y extent and the point count fit, else 64-bit.
#include "shpl_common.cuh"

namespace {
using shpl::kFull;
using shpl::lookback;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTileCells = kThreads * 32;     // one bitmap word per thread

struct BevGeom {
    double plane[4];              // ground plane a, b, c, d
    double ext[6];                // x_lo, x_hi, y_lo, y_hi, z_lo, z_hi
    double voxel;
    double d_lo[SHPL_BEV_MAX_SLICES], d_hi[SHPL_BEV_MAX_SLICES];   // d - offset of the two planes of each slice (obj_utils.py:479)
    double d_band_lo, d_band_hi;                                   // the density band [height_lo, height_hi]
    double lo[SHPL_BEV_MAX_SLICES];                                // height_lo of each slice (bev_slices.py:66)
    double hpd;                                                    // height_per_division (:30-31)
    double norm;                                                   // sqrt(a^2 + b^2 + c^2) (geometry_utils.py:40)
    double log_norm;                                               // NORM_VALUES[source] (bev_slices.py:12-14)
    int S, X, Z;
    int min_x, min_z;
    int yd_min, ybits;            // y bins the extents allow: floor(y/voxel) - yd_min fits ybits bits
    int W, tiles_per_plane;       // bitmap words per plane (ceil(X*Z/32)) and emit tiles per plane
};

// internal counters, in the zeroed workspace header (the caller's counts are written once, at the end)
enum { kCtrFlags = 1, kCtrBand = 2, kCtrSlice = 16 };

struct FeederWs {
    unsigned* ticket;             // [1]
    int* ctr;                     // [32]
    unsigned long long* status;   // [tiles]
    unsigned* bitmap;             // [S+1][W]: slices, then the density band
    int* dcount;                  // [X*Z]
    size_t zero_bytes;            // everything above: zeroed by one memset
    void* grid;                   // [S][X*Z] winner words, memset to ~0
    size_t grid_bytes;
    size_t total_bytes;
};

struct ScatterArgs {
    const double* pts;
    long long coord_stride, point_stride;
    long long P;
    const int* P_dev;             // optional device count: points i >= *P_dev do not exist
    BevGeom g;
    FeederWs ws;
};

// get_point_filter's plane test: np.dot(offset_plane, [x y z 1]) < 0.  The reference's BLAS (dgemv, OpenBLAS 0.3.30
// Haswell kernel, probed on this image) rounds the 4-term dot as b*y, fma(a,x,.), fma(c,z,.), + d'.
__device__ __forceinline__ double plane_base(const BevGeom& g, double x, double y, double z) {
    double t = __dmul_rn(g.plane[1], y);
    t = __fma_rn(g.plane[0], x, t);
    t = __fma_rn(g.plane[2], z, t);
    return t;
}

template <typename Word> struct WordOps;
template <> struct WordOps<unsigned> {
    static __device__ __forceinline__ unsigned make(const BevGeom& g, int yd, long long i) {
        return g.ybits ? (((unsigned)(yd - g.yd_min) << (32 - g.ybits)) | (unsigned)i) : (unsigned)i;
    }
    static __device__ __forceinline__ long long index(const BevGeom& g, unsigned w) {
        return (long long)(g.ybits ? (w & (0xffffffffu >> g.ybits)) : w);
    }
};
template <> struct WordOps<unsigned long long> {
    static __device__ __forceinline__ unsigned long long make(const BevGeom&, int yd, long long i) {
        return ((unsigned long long)((unsigned)yd ^ 0x80000000u) << 32) | (unsigned)i;
    }
    static __device__ __forceinline__ long long index(const BevGeom&, unsigned long long w) { return (long long)(unsigned)w; }
};

template <typename Word>
__global__ void __launch_bounds__(kThreads) shpl_bev_scatter_kernel(ScatterArgs a) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const BevGeom& g = a.g;
    bool inside = false;
    double x = 0, y = 0, z = 0;
    if (i < a.P && (a.P_dev == nullptr || i < (long long)*a.P_dev)) {
        const double* p = a.pts + i * a.point_stride;
        x = p[0];
        y = p[a.coord_stride];
        z = p[2 * a.coord_stride];
        // obj_utils.py:467-472: strict on both sides
        inside = (x > g.ext[0]) && (x < g.ext[1]) && (y > g.ext[2]) && (y < g.ext[3]) && (z > g.ext[4]) && (z < g.ext[5]);
    }
    const double base = plane_base(g, x, y, z);
    // voxel_grid_2d.py:67
    const int xd = (int)floor(__ddiv_rn(x, g.voxel)), yd = (int)floor(__ddiv_rn(y, g.voxel)), zd = (int)floor(__ddiv_rn(z, g.voxel));
    const int xi = xd - g.min_x, zi = zd - g.min_z;
    const bool in_grid = xi >= 0 && xi < g.X && zi >= 0 && zi < g.Z;
    const long long XZ = (long long)g.X * g.Z;
    const long long cell = (long long)xi * g.Z + zi;
    const Word word = WordOps<Word>::make(g, yd, i);
    const unsigned bit = 1u << (cell & 31);
    Word* grid = static_cast<Word*>(a.ws.grid);
    bool bad = false;
    // per-slice point counters: summed per CTA in shared memory first (one global atomic per CTA and counter
    // instead of one per warp on the same six addresses)
    __shared__ int s_cnt[SHPL_BEV_MAX_SLICES + 1];
    if (threadIdx.x <= SHPL_BEV_MAX_SLICES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int s = 0; s < g.S; ++s) {
        // kitti_utils.py:97-107: xor of the two point filters
        const bool in_s = inside && ((__dadd_rn(base, g.d_hi[s]) < 0.0) != (__dadd_rn(base, g.d_lo[s]) < 0.0));
        const unsigned m = __ballot_sync(kFull, in_s);
        if (m && lane == 0) atomicAdd(s_cnt + s, __popc(m));
        if (in_s) {
            if (in_grid) {
                // the first point to reach a cell (it finds the empty word) sets the cell's occupancy bit
                const Word old = atomicMin(grid + (size_t)s * XZ + cell, word);
                if (old == (Word)~(Word)0) atomicOr(a.ws.bitmap + (size_t)s * g.W + (cell >> 5), bit);
            } else {
                bad = true;       // voxel_grid_2d.py:133-138 raises ValueError
            }
        }
    }
    const bool in_band = inside && ((__dadd_rn(base, g.d_band_hi) < 0.0) != (__dadd_rn(base, g.d_band_lo) < 0.0));
    const unsigned mb = __ballot_sync(kFull, in_band);
    if (mb && lane == 0) atomicAdd(s_cnt + SHPL_BEV_MAX_SLICES, __popc(mb));
    if (in_band) {
        if (in_grid) {
            if (atomicAdd(a.ws.dcount + cell, 1) == 0) atomicOr(a.ws.bitmap + (size_t)g.S * g.W + (cell >> 5), bit);
        } else {
            bad = true;
        }
    }
    if (bad) atomicOr(a.ws.ctr + kCtrFlags, SHPL_BEV_ERR_EXTENTS);
    __syncthreads();
    if (threadIdx.x < g.S && s_cnt[threadIdx.x]) atomicAdd(a.ws.ctr + kCtrSlice + threadIdx.x, s_cnt[threadIdx.x]);
    if (threadIdx.x == SHPL_BEV_MAX_SLICES && s_cnt[SHPL_BEV_MAX_SLICES]) atomicAdd(a.ws.ctr + kCtrBand, s_cnt[SHPL_BEV_MAX_SLICES]);
}

struct EmitArgs {
    const double* pts;
    long long coord_stride, point_stride;
    BevGeom g;
    FeederWs ws;
    int use_ticket;
    long long cap;
    long long* vox_out;           // [cap,2]
    double* pts_out;              // [cap,3]
    double* maps;                 // [S+1][Z][X] or null
    const double* lut;            // density value for n points, n < lut_len (device) or null
    int lut_len;
    int* counts;                  // the caller's [SHPL_BEV_COUNTS]
};

template <typename Word>
__global__ void __launch_bounds__(kThreads) shpl_bev_emit_kernel(EmitArgs a) {
    __shared__ int s_tile;
    __shared__ unsigned s_wtot[kWarps];
    __shared__ unsigned long long s_excl;
    __shared__ int s_src[SHPL_BEV_MAX_SLICES];
    __shared__ unsigned short s_list[kTileCells];     // occupied cells of the tile (bit offset inside the tile), in order
    const BevGeom& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int tile = blockIdx.x;
    if (a.use_ticket) {          // more CTAs than can be resident: order the look-back by arrival
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ws.ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    // bev_slices.py:79: a slice with <= 1 point re-uses the previous slice's grid (quirk A.4-7)
    if (threadIdx.x == 0) {
        int src = -1;
        for (int s = 0; s < g.S; ++s) {
            if (a.ws.ctr[kCtrSlice + s] > 1) src = s;
            s_src[s] = src;
        }
    }
    __syncthreads();
    const int plane = tile / g.tiles_per_plane;              // < S: a height slice; == S: the density band
    const int tip = tile - plane * g.tiles_per_plane;
    const int n_slice_tiles = g.S * g.tiles_per_plane;
    const long long XZ = (long long)g.X * g.Z;
    const int w = tip * kThreads + threadIdx.x;              // bitmap word of this thread inside the plane
    const int src = plane < g.S ? s_src[plane] : g.S;
    unsigned bits = 0;
    if (w < g.W && src >= 0) bits = a.ws.bitmap[(size_t)src * g.W + w];
    const unsigned cnt = __popc(bits);
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    unsigned before = incl - cnt, tot = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        if (q < warp) before += s_wtot[q];
        tot += s_wtot[q];
    }
    for (unsigned b = bits, r = before; b; b &= b - 1, ++r) s_list[r] = (unsigned short)(threadIdx.x * 32 + (__ffs(b) - 1));
    if (warp == 0 && plane < g.S) {
        const unsigned long long excl = lookback(a.ws.status, tile, (unsigned long long)tot, lane);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const long long base = plane < g.S ? (long long)s_excl : 0;
    if (threadIdx.x == 0 && plane < g.S) {
        if (tip == 0) a.counts[8 + plane] = (int)base;                 // where slice `plane` starts in the output
        if (tile == n_slice_tiles - 1) {                               // inclusive prefix of the last slice tile = totals
            const long long n = base + tot;
            int flags = a.ws.ctr[kCtrFlags];
            if (s_src[0] < 0) flags |= SHPL_BEV_ERR_FIRST_SLICE_EMPTY;           // NameError in the reference
            if (a.ws.ctr[kCtrBand] == 0) flags |= SHPL_BEV_ERR_NO_POINTS;         // voxelize_2d of nothing raises
            if (n > a.cap) flags |= SHPL_BEV_ERR_CAPACITY;
            a.counts[0] = (int)n;
            a.counts[1] = flags;
            a.counts[2] = a.ws.ctr[kCtrBand];
            for (int q = 3; q < 8; ++q) a.counts[q] = 0;
            for (int q = 8 + g.S; q < 16; ++q) a.counts[q] = 0;
            for (int q = 0; q < 16; ++q) a.counts[16 + q] = q < g.S ? a.ws.ctr[kCtrSlice + q] : 0;
        }
    }
    const Word* grid = static_cast<const Word*>(a.ws.grid);
    if (plane == g.S) {          // the density band: the non-zeros of the density map
        if (a.maps == nullptr) return;
        for (unsigned it = threadIdx.x; it < tot; it += kThreads) {
            const long long cell = (long long)tip * kTileCells + s_list[it];
            const int xi = (int)(cell / g.Z), zi = (int)(cell - (long long)xi * g.Z);
            // bev_generator.py:35-36: min(1, log(n + 1) / norm)
            const int n = a.ws.dcount[cell];
            double dv;
            if (a.lut) dv = n < a.lut_len ? a.lut[n] : 1.0;
            else dv = fmin(1.0, __ddiv_rn(log((double)n + 1.0), g.log_norm));
            a.maps[(size_t)g.S * XZ + (size_t)(g.Z - 1 - zi) * g.X + xi] = dv;      // np.flip(map.transpose(), axis=0)
        }
        return;
    }
    // A slice: kItems occupied cells per thread and round, so that the dependent loads (winner word -> point)
    // of several cells are in flight together; crowded near-range tiles hold > 1000 cells.
    constexpr int kItems = 4;
    for (unsigned it0 = threadIdx.x; it0 < tot; it0 += kItems * kThreads) {
        long long cell[kItems];
        Word win[kItems];
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
            const unsigned it = it0 + u * kThreads;
            cell[u] = -1;
            win[u] = 0;
            if (it < tot) {
                cell[u] = (long long)tip * kTileCells + s_list[it];
                win[u] = grid[(size_t)src * XZ + cell[u]];
            }
        }
        double px[kItems], py[kItems], pz[kItems];
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
            px[u] = py[u] = pz[u] = 0.0;
            if (cell[u] >= 0) {
                const double* p = a.pts + WordOps<Word>::index(g, win[u]) * a.point_stride;
                px[u] = p[0];
                py[u] = p[a.coord_stride];
                pz[u] = p[2 * a.coord_stride];
            }
        }
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
            if (cell[u] < 0) continue;
            const long long pos = base + it0 + u * kThreads;
            const int xi = (int)(cell[u] / g.Z), zi = (int)(cell[u] - (long long)xi * g.Z);
            const double x = px[u], y = py[u], z = pz[u];
            if (pos < a.cap) {
                a.vox_out[2 * pos] = xi;
                a.vox_out[2 * pos + 1] = g.Z - zi;                     // bev_slices.py:106-108 (num_divisions - z)
                a.pts_out[3 * pos] = x;
                a.pts_out[3 * pos + 1] = y;
                a.pts_out[3 * pos + 2] = z;
            }
            if (a.maps) {
                // geometry_utils.py:40: (a*x + b*y + c*z + d) / norm, numpy elementwise (no contraction)
                double h = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.plane[0], x), __dmul_rn(g.plane[1], y)), __dmul_rn(g.plane[2], z)), g.plane[3]);
                h = __ddiv_rn(h, g.norm);
                // bev_slices.py:100: heights -= height_lo, once per slice the grid is used for
                for (int t = src; t <= plane; ++t) h = __dsub_rn(h, g.lo[t]);
                // np.flip(map.transpose(), axis=0)  (:116-118)
                a.maps[(size_t)plane * XZ + (size_t)(g.Z - 1 - zi) * g.X + xi] = __ddiv_rn(h, g.hpd);
            }
        }
    }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

FeederWs carve(void* base, const BevGeom& g, bool wide) {
    FeederWs w{};
    char* p = static_cast<char*>(base);
    const long long XZ = (long long)g.X * g.Z;
    const long long tiles = (long long)(g.S + 1) * g.tiles_per_plane;
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned*>(p + off);
    off += 64;
    w.ctr = reinterpret_cast<int*>(p + off);
    off += 192;
    w.status = reinterpret_cast<unsigned long long*>(p + off);
    off = align_up(off + sizeof(unsigned long long) * tiles, 64);
    w.bitmap = reinterpret_cast<unsigned*>(p + off);
    off = align_up(off + sizeof(unsigned) * (size_t)(g.S + 1) * g.W, 64);
    w.dcount = reinterpret_cast<int*>(p + off);
    off = align_up(off + sizeof(int) * XZ, 256);
    w.zero_bytes = off;
    w.grid = p + off;
    w.grid_bytes = (wide ? sizeof(unsigned long long) : sizeof(unsigned)) * (size_t)g.S * XZ;
    off = align_up(off + w.grid_bytes, 256);
    w.total_bytes = off;
    return w;
}

int make_geometry(BevGeom& g, const double* plane, const double* ext, double voxel, double height_lo, double height_hi,
                  int S, double log_norm, const char* who) {
    SHPL_REQUIRE(plane && ext, SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(S >= 1 && S <= SHPL_BEV_MAX_SLICES, SHPL_ERR_INVALID_ARGUMENT, "%s: num_slices=%d not in [1, %d]", who, S,
                 SHPL_BEV_MAX_SLICES);
    SHPL_REQUIRE(voxel > 0.0, SHPL_ERR_INVALID_ARGUMENT, "%s: voxel_size must be positive", who);
    for (int i = 0; i < 4; ++i) g.plane[i] = plane[i];
    for (int i = 0; i < 6; ++i) g.ext[i] = ext[i];
    g.voxel = voxel;
    g.S = S;
    // voxel_grid_2d.py:126-149
    const double min_x = floor(ext[0] / voxel), max_x = ceil(ext[1] / voxel - 1.0);
    const double min_z = floor(ext[4] / voxel), max_z = ceil(ext[5] / voxel - 1.0);
    const double nx = max_x - min_x + 1.0, nz = max_z - min_z + 1.0;
    SHPL_REQUIRE(nx >= 1.0 && nz >= 1.0 && nx * nz * (S + 1) < 2147483647.0 && fabs(min_x) < 1e9 && fabs(min_z) < 1e9,
                 SHPL_ERR_INVALID_ARGUMENT, "%s: extents / voxel_size give a %g x %g grid", who, nx, nz);
    g.X = (int)nx;
    g.Z = (int)nz;
    g.min_x = (int)min_x;
    g.min_z = (int)min_z;
    g.W = (int)(((long long)g.X * g.Z + 31) / 32);
    g.tiles_per_plane = (g.W + kThreads - 1) / kThreads;
    // y strictly inside the extents => floor(y / voxel) inside [floor(y_lo / voxel), floor(y_hi / voxel)] (division is monotone)
    const double yd_lo = floor(ext[2] / voxel), yd_hi = floor(ext[3] / voxel);
    g.yd_min = 0;
    g.ybits = 32;                                     // 32: the 32-bit winner words cannot be used
    if (yd_hi >= yd_lo && fabs(yd_lo) < 1e9 && yd_hi - yd_lo < 1e9) {
        g.yd_min = (int)yd_lo;
        const unsigned span = (unsigned)(yd_hi - yd_lo);
        g.ybits = 0;
        while (g.ybits < 32 && (span >> g.ybits) != 0u) ++g.ybits;
    }
    // bev_slices.py:29-31, :66-67
    g.hpd = (height_hi - height_lo) / (double)S;
    for (int s = 0; s < S; ++s) {
        volatile double lo = height_lo + (double)s * g.hpd;     // volatile: each step rounded to double like Python's floats
        volatile double hi = lo + g.hpd;
        g.lo[s] = lo;
        g.d_lo[s] = plane[3] + (-lo);                           // obj_utils.py:479: ground_plane + [0, 0, 0, -offset]
        g.d_hi[s] = plane[3] + (-hi);
    }
    g.d_band_lo = plane[3] + (-height_lo);
    g.d_band_hi = plane[3] + (-height_hi);
    g.norm = sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);
    g.log_norm = log_norm;
    return SHPL_OK;
}

}  // namespace

extern "C" int shpl_bev_grid_dims(const double* extents_host, double voxel_size, int32_t* nx, int32_t* nz) {
    const double plane[4] = {0, -1, 0, 0};
    BevGeom g{};
    if (int rc = make_geometry(g, plane, extents_host, voxel_size, 0.0, 1.0, 1, 1.0, "shpl_bev_grid_dims")) return rc;
    if (nx) *nx = g.X;
    if (nz) *nz = g.Z;
    return SHPL_OK;
}

extern "C" size_t shpl_bev_workspace_bytes(const double* extents_host, double voxel_size, int32_t num_slices) {
    const double plane[4] = {0, -1, 0, 0};
    BevGeom g{};
    if (make_geometry(g, plane, extents_host, voxel_size, 0.0, 1.0, num_slices, 1.0, "shpl_bev_workspace_bytes")) return 0;
    return carve(nullptr, g, true).total_bytes;       // sized for the 64-bit winner words
}

extern "C" int shpl_bev_slices(const double* points, int64_t coord_stride, int64_t point_stride, int64_t P,
                               const int32_t* P_dev, const double* ground_plane_host, const double* extents_host, double voxel_size,
                               double height_lo, double height_hi, int32_t num_slices, double log_norm,
                               const double* density_lut, int32_t lut_len,
                               int64_t* voxel_indices_out, double* unique_pts_out, int64_t capacity,
                               double* bev_maps_out, int32_t* counts, void* workspace, size_t workspace_bytes,
                               void* stream) {
    const char* who = "shpl_bev_slices";
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHPL_REQUIRE(P >= 0 && P < (1ll << 31), SHPL_ERR_INVALID_ARGUMENT, "%s: P=%lld out of range", who, (long long)P);
    SHPL_REQUIRE((P == 0 || points) && counts && workspace && capacity >= 0 && (capacity == 0 || (voxel_indices_out && unique_pts_out)),
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(shpl::aligned(workspace, 256), SHPL_ERR_INVALID_ARGUMENT, "%s: workspace must be 256-byte aligned", who);
    SHPL_REQUIRE(log_norm > 0.0 && (density_lut == nullptr || lut_len > 0), SHPL_ERR_INVALID_ARGUMENT, "%s: bad density normalisation", who);
    BevGeom g{};
    if (int rc = make_geometry(g, ground_plane_host, extents_host, voxel_size, height_lo, height_hi, num_slices, log_norm, who)) return rc;
    const long long XZ = (long long)g.X * g.Z;
    // 32-bit winner words (y bin | point index) when both fit; the all-ones word must stay unused ("empty")
    const bool wide = !(g.ybits < 32 && (unsigned long long)P < (1ull << (32 - g.ybits)));
    FeederWs w = carve(workspace, g, wide);
    SHPL_REQUIRE(w.total_bytes <= workspace_bytes, SHPL_ERR_WORKSPACE_TOO_SMALL, "%s: workspace %zu bytes < %zu needed", who,
                 workspace_bytes, w.total_bytes);
    SHPL_CUDA_OK(cudaMemsetAsync(workspace, 0, w.zero_bytes, s));
    SHPL_CUDA_OK(cudaMemsetAsync(w.grid, 0xff, w.grid_bytes, s));
    if (bev_maps_out) SHPL_CUDA_OK(cudaMemsetAsync(bev_maps_out, 0, sizeof(double) * (size_t)(g.S + 1) * XZ, s));

    if (P > 0) {
        ScatterArgs sa{};
        sa.pts = points;
        sa.coord_stride = coord_stride;
        sa.point_stride = point_stride;
        sa.P = P;
        sa.P_dev = P_dev;
        sa.g = g;
        sa.ws = w;
        const unsigned blocks = (unsigned)((P + kThreads - 1) / kThreads);
        if (wide) shpl_bev_scatter_kernel<unsigned long long><<<blocks, kThreads, 0, s>>>(sa);
        else shpl_bev_scatter_kernel<unsigned><<<blocks, kThreads, 0, s>>>(sa);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_bev_scatter_kernel")) return rc;
    }
    EmitArgs ea{};
    ea.pts = points;
    ea.coord_stride = coord_stride;
    ea.point_stride = point_stride;
    ea.g = g;
    ea.ws = w;
    const int n_tiles = (g.S + 1) * g.tiles_per_plane;
    ea.use_ticket = 1;      // always order the look-back by arrival (no blockIdx-order dispatch guarantee; see shpl_build.cu)
    ea.cap = capacity;
    ea.vox_out = reinterpret_cast<long long*>(voxel_indices_out);
    ea.pts_out = unique_pts_out;
    ea.maps = bev_maps_out;
    ea.lut = density_lut;
    ea.lut_len = lut_len;
    ea.counts = counts;
    if (wide) shpl_bev_emit_kernel<unsigned long long><<<(unsigned)n_tiles, kThreads, 0, s>>>(ea);
    else shpl_bev_emit_kernel<unsigned><<<(unsigned)n_tiles, kThreads, 0, s>>>(ea);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_bev_emit_kernel");
}

// ------------------------------------------------------------------------------------------------ ingest
// Point-cloud ingest: velodyne scan -> camera frame -> points in front of the camera that
// project inside the image.  Replaces the arithmetic of
//   obj_utils.get_lidar_point_cloud   /root/reference/avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:220-268
//   calib_utils.lidar_to_cam_frame    /root/reference/avod/wavedata/wavedata/tools/core/calib_utils.py:371-410
//   calib_utils.project_to_image      /root/reference/avod/wavedata/wavedata/tools/core/calib_utils.py:281-297
// (the two file reads stay on the host).  One point per thread, fp64 in the rounding order of the reference's BLAS
// calls (a 4-term dot = mul, fma, fma, fma: probed on this image for both dgemm shapes), STABLE compaction by a
// decoupled look-back, output coordinate-major [3, capacity] so that shpl_bev_slices reads it through its strides.
namespace {

struct IngestArgs {
    const float* velo;            // [N,4] x, y, z, intensity
    long long N;
    double R[12];                 // rows 0..2 of R0_rect(4x4) . Tr_velo_to_cam(4x4)
    double P[12];                 // p2
    int filter;                   // 0: every point (im_size = None)
    double im_w, im_h;
    int use_intensity;
    float min_intensity;
    unsigned* ticket;
    unsigned long long* status;
    int use_ticket;
    double* out;                  // [3,cap]
    long long cap;
    int* counts;
};

__device__ __forceinline__ double dot4(const double* m, double x, double y, double z) {
    double t = __dmul_rn(m[0], x);
    t = __fma_rn(m[1], y, t);
    t = __fma_rn(m[2], z, t);
    t = __fma_rn(m[3], 1.0, t);
    return t;
}

__global__ void __launch_bounds__(kThreads) shpl_lidar_to_cam_kernel(IngestArgs a) {
    __shared__ int s_tile;
    __shared__ unsigned s_w[kWarps], s_f[kWarps];
    __shared__ unsigned long long s_ex;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int tile = blockIdx.x;
    if (a.use_ticket) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    const long long i = (long long)tile * kThreads + threadIdx.x;
    bool keep = false, front = false;
    double cx = 0, cy = 0, cz = 0;
    if (i < a.N) {
        const float4 p = reinterpret_cast<const float4*>(a.velo)[i];
        const double x = (double)p.x, y = (double)p.y, z = (double)p.z;      // np.append(float32, float64 ones) -> float64
        cx = dot4(a.R + 0, x, y, z);                                         // calib_utils.py:407
        cy = dot4(a.R + 4, x, y, z);
        cz = dot4(a.R + 8, x, y, z);
        keep = true;
        front = true;
        if (a.filter) {
            front = cz > 0.0;                                                // obj_utils.py:251
            const double w = dot4(a.P + 8, cx, cy, cz);
            const double u = __ddiv_rn(dot4(a.P + 0, cx, cy, cz), w);        // calib_utils.py:290-295
            const double v = __ddiv_rn(dot4(a.P + 4, cx, cy, cz), w);
            keep = front && (u > 0.0) && (u < a.im_w) && (v > 0.0) && (v < a.im_h);   // obj_utils.py:258-261 (strict)
            if (a.use_intensity) keep = keep && (p.w > a.min_intensity);     // :266
        }
    }
    const unsigned m = __ballot_sync(kFull, keep), mf = __ballot_sync(kFull, front);
    if (lane == 0) {
        s_w[warp] = __popc(m);
        s_f[warp] = __popc(mf);
    }
    __syncthreads();
    unsigned before = 0, tot = 0, totf = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        if (q < warp) before += s_w[q];
        tot += s_w[q];
        totf += s_f[q];
    }
    if (warp == 0) {
        const unsigned long long ex = lookback(a.status, tile, ((unsigned long long)totf << 31) | tot, lane);
        if (lane == 0) s_ex = ex;
    }
    __syncthreads();
    const long long base = (long long)(s_ex & ((1ull << 31) - 1));
    const long long n_tiles = (a.N + kThreads - 1) / kThreads > 0 ? (a.N + kThreads - 1) / kThreads : 1;
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        a.counts[0] = (int)(base + tot);
        a.counts[1] = (int)((long long)(s_ex >> 31) + totf);
        a.counts[2] = (base + tot > a.cap) ? 1 : 0;
    }
    const long long j = base + before + __popc(m & ((1u << lane) - 1u));
    if (keep && j < a.cap) {
        a.out[j] = cx;
        a.out[a.cap + j] = cy;
        a.out[2 * a.cap + j] = cz;
    }
}

}  // namespace

extern "C" size_t shpl_lidar_workspace_bytes(int64_t n_max) {
    if (n_max < 0) n_max = 0;
    const size_t tiles = (size_t)((n_max + kThreads - 1) / kThreads) + 1;
    return 64 + align_up(sizeof(unsigned long long) * tiles, 64);
}

extern "C" int shpl_lidar_to_cam(const float* velo_xyzi, int64_t N, const double* rectified_host, const double* p2_host,
                                 int32_t im_w, int32_t im_h, int32_t use_min_intensity, float min_intensity,
                                 double* cam_out, int64_t capacity, int32_t* counts, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    const char* who = "shpl_lidar_to_cam";
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHPL_REQUIRE(N >= 0 && N < (1ll << 30), SHPL_ERR_INVALID_ARGUMENT, "%s: N=%lld out of range", who, (long long)N);
    SHPL_REQUIRE((N == 0 || velo_xyzi) && rectified_host && counts && workspace && capacity >= 0 && (capacity == 0 || cam_out),
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(shpl::aligned(velo_xyzi, 16) && shpl::aligned(workspace, 64), SHPL_ERR_INVALID_ARGUMENT,
                 "%s: the scan must be 16-byte aligned ([N,4] float32) and the workspace 64-byte aligned", who);
    const bool filter = im_w > 0 && im_h > 0;
    SHPL_REQUIRE(!filter || p2_host, SHPL_ERR_INVALID_ARGUMENT, "%s: an image size needs p2", who);
    SHPL_REQUIRE(workspace_bytes >= shpl_lidar_workspace_bytes(N), SHPL_ERR_WORKSPACE_TOO_SMALL, "%s: workspace %zu bytes < %zu needed",
                 who, workspace_bytes, shpl_lidar_workspace_bytes(N));
    IngestArgs a{};
    a.velo = velo_xyzi;
    a.N = N;
    for (int i = 0; i < 12; ++i) a.R[i] = rectified_host[i];
    for (int i = 0; i < 12; ++i) a.P[i] = p2_host ? p2_host[i] : 0.0;
    a.filter = filter ? 1 : 0;
    a.im_w = (double)im_w;
    a.im_h = (double)im_h;
    a.use_intensity = (filter && use_min_intensity) ? 1 : 0;
    a.min_intensity = min_intensity;
    a.ticket = static_cast<unsigned*>(workspace);
    a.status = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + 64);
    a.out = cam_out;
    a.cap = capacity;
    a.counts = counts;
    SHPL_CUDA_OK(cudaMemsetAsync(workspace, 0, shpl_lidar_workspace_bytes(N), s));
    const long long tiles = (N + kThreads - 1) / kThreads > 0 ? (N + kThreads - 1) / kThreads : 1;
    a.use_ticket = 1;       // always order the look-back by arrival
    shpl_lidar_to_cam_kernel<<<(unsigned)tiles, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_lidar_to_cam_kernel");
}
