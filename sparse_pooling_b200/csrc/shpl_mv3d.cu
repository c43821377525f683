// MV3D voxel feeder for sm_100a: the source of SHPL's
// non-homogeneous weights M_val = 1 / (points in the pair's 3-D voxel).
//
// Replaces  point_cloud_2_top_sparse(points, ..., points_in_cam=True, img_index2=...)
//   /root/reference/MV3D_TF_release/lib/utils/construct_voxel.py:37-162
// called once per training sample from
//   /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:93
//
// The reference finds the 3-D voxels with np.unique(axis=0) and then walks the points in a Python
// loop, keeping the first MAX_NUM_POINTS points of every voxel.  Here:
//   K1  shpl_mv3d_cells_kernel   range filter (strict, :89-96), cell indices by truncation (:116-118),
//                                STABLE compaction of the in-range points (decoupled look-back)
//   Kr  shpl_radix_pass_kernel   stable LSD radix sort by voxel key (x, y, z lexicographic = the row order
//                                np.unique(axis=0) returns); inside a voxel the points keep input order
//   K3  shpl_mv3d_runs_kernel    voxel id of every sorted point (scan of run heads), run starts
//   K4  shpl_mv3d_pairs_kernel   slot of every point inside its voxel, cap (:135-140), STABLE compaction of
//                                the survivors in input order -> img_index, bv_index, M_val (:156-160)
//   K5  shpl_mv3d_voxels_kernel  a warp per voxel: feature_buffer [V,T,7] with the offsets from the voxel
//                                mean (summed slot by slot like np.sum(axis=1), :143), coordinate_buffer,
//                                number_buffer (:146-148)
// Everything is asynchronous on the caller's stream; nothing is read back.
#include "shpl_common.cuh"
#include "shpl_sort.cuh"

namespace {
using shpl::lookback;

struct Mv3dGeom {
    double side_lo, side_hi, fwd_lo, fwd_hi, h_lo, h_hi;
    double res, zres;
    int nx, ny, nz;      // x_max+1 (side), y_max+1 (fwd), z_max+1 (height)   (:81-84)
    int T;               // MAX_NUM_POINTS
};

struct Mv3dWs {
    unsigned* ticket[3];
    unsigned long long* status[3];
    size_t zero_bytes;
    Workspace sort;              // items[0][*], hist[0][*] are used
    int* src_i;                  // [n] original index of in-range point j
    int* pos_j;                  // [n] sorted position of j
    int* vid_t;                  // [n] voxel of sorted position t
    int* run_start;              // [n+1]
    size_t total_bytes;
};

size_t align_up2(size_t x, size_t a) { return (x + a - 1) / a * a; }

Mv3dWs carve_mv3d(void* base, long long n, const SortPlan* sp) {
    Mv3dWs w{};
    char* p = static_cast<char*>(base);
    const long long tiles = (n + kPairsTile - 1) / kPairsTile > 0 ? (n + kPairsTile - 1) / kPairsTile : 1;
    size_t off = 0;
    for (int q = 0; q < 3; ++q) {
        w.ticket[q] = reinterpret_cast<unsigned*>(p + off);
        off += 64;
    }
    for (int q = 0; q < 3; ++q) {
        w.status[q] = reinterpret_cast<unsigned long long*>(p + off);
        off = align_up2(off + sizeof(unsigned long long) * tiles, 64);
    }
    w.zero_bytes = off;
    w.sort = carve(p + off, n, sp);
    off = align_up2(off + w.sort.total_bytes, 64);
    const size_t cnt = (size_t)(n > 0 ? n : 1);
    w.src_i = reinterpret_cast<int*>(p + off);
    off = align_up2(off + sizeof(int) * cnt, 64);
    w.pos_j = reinterpret_cast<int*>(p + off);
    off = align_up2(off + sizeof(int) * cnt, 64);
    w.vid_t = reinterpret_cast<int*>(p + off);
    off = align_up2(off + sizeof(int) * cnt, 64);
    w.run_start = reinterpret_cast<int*>(p + off);
    off = align_up2(off + sizeof(int) * (cnt + 1), 64);
    w.total_bytes = off;
    return w;
}

struct Mv3dArgs {
    const double* points;        // [n,4] camera frame: x (side), y (height), z (forward), reflectance
    const long long* img2;       // [2,n]
    long long n;
    Mv3dGeom g;
    Mv3dWs ws;
    SortPlan sp;
    int use_ticket;
    long long cap, vcap;
    long long* img_index_out;    // [3,cap]
    long long* bv_index_out;     // [cap,2]
    double* m_val_out;           // [cap]
    double* feature;             // [vcap,T,7] or null
    long long* coordinate;       // [vcap,4] or null
    long long* number;           // [vcap] or null
    int* counts;                 // [0] in-range points, [1] pairs kept, [2] voxels, [3] error bits
};

__device__ __forceinline__ int tile_of(unsigned* ticket, int use_ticket, int* s_tile) {
    int tile = blockIdx.x;
    if (use_ticket) {
        if (threadIdx.x == 0) *s_tile = (int)atomicAdd(ticket, 1u);
        __syncthreads();
        tile = *s_tile;
    }
    return tile;
}

// block-wide exclusive position of a flag + prefix over earlier tiles; returns the position, sets `total`
// to the inclusive prefix of this tile (earlier tiles + this tile)
__device__ __forceinline__ long long block_compact(bool flag, unsigned long long* status, int tile, long long& total) {
    __shared__ unsigned s_w[kWarps];
    __shared__ unsigned long long s_ex;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(kFull, flag);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    unsigned before = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        if (q < warp) before += s_w[q];
        tot += s_w[q];
    }
    if (warp == 0) {
        const unsigned long long ex = lookback(status, tile, (unsigned long long)tot, lane);
        if (lane == 0) s_ex = ex;
    }
    __syncthreads();
    total = (long long)s_ex + tot;
    return (long long)s_ex + before + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kThreads) shpl_mv3d_cells_kernel(Mv3dArgs a) {
    __shared__ int s_tile;
    const int tile = tile_of(a.ws.ticket[0], a.use_ticket, &s_tile);
    const long long i = (long long)tile * kPairsTile + threadIdx.x;
    const Mv3dGeom& g = a.g;
    bool in = false;
    unsigned key = 0;
    if (i < a.n) {
        const double side = a.points[4 * i], h = a.points[4 * i + 1], fwd = a.points[4 * i + 2];   // :59 column swap
        in = (fwd > g.fwd_lo) && (fwd < g.fwd_hi) && (side > g.side_lo) && (side < g.side_hi) && (h > g.h_lo) && (h < g.h_hi);
        if (in) {
            const int xi = (int)__ddiv_rn(__dsub_rn(side, g.side_lo), g.res);      // :116-118 astype(np.int32) truncates
            const int yi = (int)__ddiv_rn(__dsub_rn(fwd, g.fwd_lo), g.res);
            const int zi = (int)__ddiv_rn(__dsub_rn(h, g.h_lo), g.zres);
            const bool ok = xi >= 0 && xi < g.nx && yi >= 0 && yi < g.ny && zi >= 0 && zi < g.nz;
            if (!ok) atomicOr(a.counts + 3, 1);
            key = ok ? (unsigned)((xi * g.ny + yi) * g.nz + zi) : 0u;
        }
    }
    long long total;
    const long long j = block_compact(in, a.ws.status[0], tile, total);
    const long long n_tiles = (a.n + kPairsTile - 1) / kPairsTile > 0 ? (a.n + kPairsTile - 1) / kPairsTile : 1;
    if (tile == n_tiles - 1 && threadIdx.x == 0) a.counts[0] = (int)total;
    if (!in) return;
    a.ws.src_i[j] = (int)i;
    a.ws.sort.items[0][0][j] = ((unsigned long long)key << 32) | (unsigned)j;
    atomicAdd(a.ws.sort.hist[0][0] + ((size_t)(j / kTile) << a.sp.bits[0]) + (key & ((1u << a.sp.bits[0]) - 1u)), 1u);
}

__global__ void __launch_bounds__(kThreads) shpl_mv3d_runs_kernel(Mv3dArgs a) {
    __shared__ int s_tile;
    const int tile = tile_of(a.ws.ticket[1], a.use_ticket, &s_tile);
    const int n = a.counts[0];
    if ((long long)tile * kPairsTile >= (long long)n && tile > 0) return;      // empty tiles follow non-empty ones
    const unsigned long long* items = a.ws.sort.items[0][a.sp.passes & 1];
    const long long t = (long long)tile * kPairsTile + threadIdx.x;
    bool head = false;
    unsigned long long it = 0;
    if (t < n) {
        it = items[t];
        head = t == 0 || (unsigned)(items[t - 1] >> 32) != (unsigned)(it >> 32);
    }
    // voxel id = (heads at or before t) - 1: compaction position of a head, or of the last head before t
    long long total;
    const long long hpos = block_compact(head, a.ws.status[1], tile, total);   // heads strictly before t
    if (t < n) {
        const int v = (int)(head ? hpos : hpos - 1);
        a.ws.vid_t[t] = v;
        a.ws.pos_j[(unsigned)it] = (int)t;
        if (head) a.ws.run_start[v] = (int)t;
        if (t == n - 1) {
            a.ws.run_start[v + 1] = n;
            a.counts[2] = v + 1;
            if ((long long)v + 1 > a.vcap) atomicOr(a.counts + 3, 4);
        }
    }
    if (n == 0 && tile == 0 && threadIdx.x == 0) {
        a.ws.run_start[0] = 0;
        a.counts[2] = 0;
    }
}

__global__ void __launch_bounds__(kThreads) shpl_mv3d_pairs_kernel(Mv3dArgs a) {
    __shared__ int s_tile;
    const int tile = tile_of(a.ws.ticket[2], a.use_ticket, &s_tile);
    const int n = a.counts[0];
    if ((long long)tile * kPairsTile >= (long long)n && tile > 0) return;
    const Mv3dGeom& g = a.g;
    const long long j = (long long)tile * kPairsTile + threadIdx.x;
    bool kept = false;
    int count = 1;
    unsigned key = 0;
    if (j < n) {
        const int t = a.ws.pos_j[j];
        const int v = a.ws.vid_t[t];
        const int s0 = a.ws.run_start[v];
        const int len = a.ws.run_start[v + 1] - s0;
        kept = t - s0 < g.T;                                   // :135-140 first MAX_NUM_POINTS points, input order
        count = len < g.T ? len : g.T;
        key = (unsigned)(a.ws.sort.items[0][a.sp.passes & 1][t] >> 32);
    }
    long long total;
    const long long k = block_compact(kept, a.ws.status[2], tile, total);
    const long long n_tiles = ((long long)n + kPairsTile - 1) / kPairsTile > 0 ? ((long long)n + kPairsTile - 1) / kPairsTile : 1;
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        a.counts[1] = (int)total;
        if (total > a.cap) atomicOr(a.counts + 3, 2);
    }
    if (!kept || k >= a.cap) return;
    const long long i = a.ws.src_i[j];
    const int zi = (int)(key % (unsigned)g.nz);
    const unsigned xy = key / (unsigned)g.nz;
    const int yi = (int)(xy % (unsigned)g.ny), xi = (int)(xy / (unsigned)g.ny);
    (void)zi;
    a.img_index_out[k] = a.img2[i];                            // :156-158
    a.img_index_out[a.cap + k] = a.img2[a.n + i];
    a.img_index_out[2 * a.cap + k] = 0;
    a.bv_index_out[2 * k] = yi;                                // :159  xyz_img[:, [1, 0]] = (fwd cell, side cell)
    a.bv_index_out[2 * k + 1] = xi;
    a.m_val_out[k] = __ddiv_rn(1.0, (double)count);            // :160
}

// A warp per voxel.
__global__ void __launch_bounds__(kThreads) shpl_mv3d_voxels_kernel(Mv3dArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long v = (long long)blockIdx.x * kWarps + warp;
    const int V = a.counts[2];
    if (v >= V || v >= a.vcap) return;
    const Mv3dGeom& g = a.g;
    const unsigned long long* items = a.ws.sort.items[0][a.sp.passes & 1];
    const int s0 = a.ws.run_start[v];
    const int len = a.ws.run_start[v + 1] - s0;
    const int count = len < g.T ? len : g.T;
    if (lane == 0) {
        const unsigned key = (unsigned)(items[s0] >> 32);
        const int zi = (int)(key % (unsigned)g.nz);
        const unsigned xy = key / (unsigned)g.nz;
        if (a.coordinate) {                                    // :128 [0, z, x, y]
            a.coordinate[4 * v] = 0;
            a.coordinate[4 * v + 1] = zi;
            a.coordinate[4 * v + 2] = (int)(xy / (unsigned)g.ny);
            a.coordinate[4 * v + 3] = (int)(xy % (unsigned)g.ny);
        }
        if (a.number) a.number[v] = count;
    }
    if (a.feature == nullptr) return;
    // pass 1: sum of the voxel's points, slot by slot (np.sum(axis=1) adds the slots in order; empty slots add 0)
    double sf = 0.0, ss = 0.0, sh = 0.0;
    for (int r0 = 0; r0 < count; r0 += 32) {
        double f = 0.0, s = 0.0, h = 0.0;
        if (r0 + lane < count) {
            const long long i = a.ws.src_i[(unsigned)items[s0 + r0 + lane]];
            s = a.points[4 * i];
            h = a.points[4 * i + 1];
            f = a.points[4 * i + 2];
        }
        const int m = count - r0 < 32 ? count - r0 : 32;
        for (int q = 0; q < m; ++q) {
            sf = __dadd_rn(sf, __shfl_sync(kFull, f, q));
            ss = __dadd_rn(ss, __shfl_sync(kFull, s, q));
            sh = __dadd_rn(sh, __shfl_sync(kFull, h, q));
        }
    }
    const double mf = __ddiv_rn(sf, (double)count), ms = __ddiv_rn(ss, (double)count), mh = __ddiv_rn(sh, (double)count);
    // pass 2: [slot, 0:4] = (fwd, side, height, reflectance), [slot, 4:7] = xyz - mean (also for empty slots: 0 - mean, :143)
    double* out = a.feature + (size_t)v * g.T * 7;
    for (int slot = lane; slot < g.T; slot += 32) {
        double f = 0.0, s = 0.0, h = 0.0, r = 0.0;
        if (slot < count) {
            const long long i = a.ws.src_i[(unsigned)items[s0 + slot]];
            s = a.points[4 * i];
            h = a.points[4 * i + 1];
            f = a.points[4 * i + 2];
            r = a.points[4 * i + 3];
        }
        double* o = out + (size_t)slot * 7;
        o[0] = f;
        o[1] = s;
        o[2] = h;
        o[3] = r;
        o[4] = __dsub_rn(f, mf);
        o[5] = __dsub_rn(s, ms);
        o[6] = __dsub_rn(h, mh);
    }
}

int make_mv3d_geometry(Mv3dGeom& g, const double* ranges, double res, double zres, int max_points, const char* who) {
    SHPL_REQUIRE(ranges != nullptr, SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(res > 0.0 && zres > 0.0 && max_points >= 1, SHPL_ERR_INVALID_ARGUMENT, "%s: res, zres and max_points must be positive", who);
    g.side_lo = ranges[0];
    g.side_hi = ranges[1];
    g.fwd_lo = ranges[2];
    g.fwd_hi = ranges[3];
    g.h_lo = ranges[4];
    g.h_hi = ranges[5];
    g.res = res;
    g.zres = zres;
    g.T = max_points;
    // construct_voxel.py:81-84: int((hi - lo) / res) + 1 cells per axis
    const double nx = trunc((g.side_hi - g.side_lo) / res) + 1.0, ny = trunc((g.fwd_hi - g.fwd_lo) / res) + 1.0,
                 nz = trunc((g.h_hi - g.h_lo) / zres) + 1.0;
    SHPL_REQUIRE(nx >= 1.0 && ny >= 1.0 && nz >= 1.0 && nx * ny * nz < 2147483647.0, SHPL_ERR_INVALID_ARGUMENT,
                 "%s: ranges / resolution give a %g x %g x %g voxel grid", who, nx, ny, nz);
    g.nx = (int)nx;
    g.ny = (int)ny;
    g.nz = (int)nz;
    return SHPL_OK;
}

}  // namespace

extern "C" size_t shpl_mv3d_workspace_bytes(int64_t n_max) {
    if (n_max < 0) n_max = 0;
    return carve_mv3d(nullptr, n_max, nullptr).total_bytes;
}

extern "C" int shpl_mv3d_voxelize(const double* points, const int64_t* img_index2, int64_t n, double res, double zres,
                                  const double* ranges_host, int32_t max_points, int32_t* voxel_full_size_host,
                                  int64_t* img_index_out, int64_t* bv_index_out, double* m_val_out, int64_t capacity,
                                  double* feature_buffer, int64_t* coordinate_buffer, int64_t* number_buffer,
                                  int64_t voxel_capacity, int32_t* counts, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    const char* who = "shpl_mv3d_voxelize";
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHPL_REQUIRE(n >= 0 && n < (1ll << 30), SHPL_ERR_INVALID_ARGUMENT, "%s: n=%lld out of range", who, (long long)n);
    SHPL_REQUIRE((n == 0 || (points && img_index2)) && counts && workspace && capacity >= 0 && voxel_capacity >= 0 &&
                     (capacity == 0 || (img_index_out && bv_index_out && m_val_out)),
                 SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(shpl::aligned(workspace, 64), SHPL_ERR_INVALID_ARGUMENT, "%s: workspace must be 64-byte aligned", who);
    Mv3dArgs a{};
    if (int rc = make_mv3d_geometry(a.g, ranges_host, res, zres, max_points, who)) return rc;
    if (voxel_full_size_host) {      // :84 [z_max+1, x_max+1, y_max+1]
        voxel_full_size_host[0] = a.g.nz;
        voxel_full_size_host[1] = a.g.nx;
        voxel_full_size_host[2] = a.g.ny;
    }
    SortPlan sp[2];
    sp[0] = make_sort_plan(a.g.nx * a.g.ny * a.g.nz - 1);
    sp[1] = make_sort_plan(1);
    sp[1].passes = 0;
    a.sp = sp[0];
    a.ws = carve_mv3d(workspace, n, sp);
    SHPL_REQUIRE(a.ws.total_bytes <= workspace_bytes, SHPL_ERR_WORKSPACE_TOO_SMALL, "%s: workspace %zu bytes < %zu needed", who,
                 workspace_bytes, a.ws.total_bytes);
    a.points = points;
    a.img2 = reinterpret_cast<const long long*>(img_index2);
    a.n = n;
    a.cap = capacity;
    a.vcap = voxel_capacity;
    a.img_index_out = reinterpret_cast<long long*>(img_index_out);
    a.bv_index_out = reinterpret_cast<long long*>(bv_index_out);
    a.m_val_out = m_val_out;
    a.feature = feature_buffer;
    a.coordinate = reinterpret_cast<long long*>(coordinate_buffer);
    a.number = reinterpret_cast<long long*>(number_buffer);
    a.counts = counts;
    SHPL_CUDA_OK(cudaMemsetAsync(workspace, 0, a.ws.zero_bytes, s));
    SHPL_CUDA_OK(cudaMemsetAsync(a.ws.sort.ticket, 0, a.ws.sort.header_bytes, s));
    SHPL_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 8, s));
    const long long ptiles = (n + kPairsTile - 1) / kPairsTile > 0 ? (n + kPairsTile - 1) / kPairsTile : 1;
    const long long tiles = (n + kTile - 1) / kTile > 0 ? (n + kTile - 1) / kTile : 1;
    a.use_ticket = 1;       // always order the look-back by arrival
    shpl_mv3d_cells_kernel<<<(unsigned)ptiles, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_mv3d_cells_kernel")) return rc;
    RadixArgs ra{};
    ra.n_dev = counts;
    ra.ws = a.ws.sort;
    ra.sp[0] = sp[0];
    ra.sp[1] = sp[1];
    for (int q = 0; q < sp[0].passes && n > 0; ++q) {
        ra.pass = q;
        shpl_radix_pass_kernel<<<dim3((unsigned)tiles, 1), kThreads, 0, s>>>(ra);
        shpl::count_launches(1);
        if (int rc = shpl::check_launch("shpl_radix_pass_kernel")) return rc;
    }
    shpl_mv3d_runs_kernel<<<(unsigned)ptiles, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_mv3d_runs_kernel")) return rc;
    shpl_mv3d_pairs_kernel<<<(unsigned)ptiles, kThreads, 0, s>>>(a);
    shpl::count_launches(1);
    if (int rc = shpl::check_launch("shpl_mv3d_pairs_kernel")) return rc;
    if (feature_buffer || coordinate_buffer || number_buffer) {
        const long long vmax = n < voxel_capacity ? n : voxel_capacity;
        const long long blocks = (vmax + kWarps - 1) / kWarps;
        if (blocks > 0) {
            shpl_mv3d_voxels_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(a);
            shpl::count_launches(1);
            if (int rc = shpl::check_launch("shpl_mv3d_voxels_kernel")) return rc;
        }
    }
    return SHPL_OK;
}
