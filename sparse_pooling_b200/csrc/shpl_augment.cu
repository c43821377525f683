// Augmentation hooks that must keep the point <-> pixel correspondences of SHPL valid (sm_100a).
//
// Replaces, on the device, the numpy lines of the reference's data layers that touch the arrays the
// correspondence builder reads:
//   /root/reference/avod/avod/datasets/kitti/kitti_aug.py:24-29                flip_point_cloud
//   /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:172-180, :183-185
//                                  img_index2 = round(projectToImage(lidar_pc)) and augment_voxel's shift /
//                                  expansion / rotation of the point cloud (projection FIRST: the pixels a point
//                                  falls on do not move with the BEV augmentation)
//   /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:205-206
//                                  augment_fv: img_index = (img_index * expansion_ratio + shift).astype(int)
// One element-wise kernel each: a few hundred KB of traffic, launch-latency bound; they exist so that the chain
// scan -> ingest -> (augment) -> feeder -> builder never leaves the device.  fp64 arithmetic is spelled out with
// __dmul_rn / __dadd_rn / __fma_rn in the order numpy (and its BLAS for the matrix products) rounds.
#include "shpl_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) shpl_flip_kernel(double* points, long long point_stride, long long n,
                                                             const int* n_dev) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n || (n_dev != nullptr && i >= (long long)*n_dev)) return;
    points[i * point_stride] = -points[i * point_stride];      // kitti_aug.py:28
}

struct Mv3dAugArgs {
    double* pc;                  // [n,4] (x, y, z, reflectance), camera frame; transformed in place when augmenting
    long long n;
    const int* n_dev;
    double P[12];
    int augment;
    double sx, sz, ratio;
    double rot[4];               // rot_mat row-major: [[cos, sin], [-sin, cos]] as the host evaluated it
    long long* img_index2;       // [2,n]
};

// transform.py:429-452: one row of P times [x y z 1], rounded like the reference's dgemm (fma chain in k order)
__device__ __forceinline__ double prow(const double* p, double x, double y, double z) {
    double t = __dmul_rn(p[0], x);
    t = __fma_rn(p[1], y, t);
    t = __fma_rn(p[2], z, t);
    t = __fma_rn(p[3], 1.0, t);
    return t;
}

// np.round(v).astype(int) on x86-64: rint, then cvttsd2si -- "integer indefinite" (INT64_MIN) for NaN, infinities
// and anything outside the int64 range
__device__ __forceinline__ long long round_to_int64(double v) {
    const double r = rint(v);
    if (!(r >= -9223372036854775808.0 && r < 9223372036854775808.0)) return (long long)0x8000000000000000ull;
    return (long long)r;
}

__global__ void __launch_bounds__(kThreads) shpl_mv3d_project_augment_kernel(Mv3dAugArgs a) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.n || (a.n_dev != nullptr && i >= (long long)*a.n_dev)) return;
    double2* row = reinterpret_cast<double2*>(a.pc + 4 * i);
    const double2 xy = row[0];
    double2 zr = row[1];
    double x = xy.x, y = xy.y, z = zr.x;
    if (a.img_index2) {          // minibatch_mv3d_img.py:172-174 / :183-185
        const double w = prow(a.P + 8, x, y, z);
        a.img_index2[i] = round_to_int64(__ddiv_rn(prow(a.P + 0, x, y, z), w));
        a.img_index2[a.n + i] = round_to_int64(__ddiv_rn(prow(a.P + 4, x, y, z), w));
    }
    if (!a.augment) return;
    x = __dadd_rn(x, a.sx);                      // :176-177
    z = __dadd_rn(z, a.sz);
    x = __dmul_rn(x, a.ratio);                   // :179
    y = __dmul_rn(y, a.ratio);
    z = __dmul_rn(z, a.ratio);
    // :181  np.dot(rot_mat, pc[:, [0,2]].T).T -- dgemm with k = 2: product of the first term, fma of the second
    const double xr = __fma_rn(a.rot[1], z, __dmul_rn(a.rot[0], x));
    const double zr2 = __fma_rn(a.rot[3], z, __dmul_rn(a.rot[2], x));
    row[0] = make_double2(xr, y);
    zr.x = zr2;
    row[1] = zr;
}

__global__ void __launch_bounds__(kThreads) shpl_augment_fv_kernel(long long* img_index, long long ld, long long n,
                                                                   const int* n_dev, double ratio, double sx, double sy) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n || (n_dev != nullptr && i >= (long long)*n_dev)) return;
    // (int64 * float64 + float64).astype(int): two roundings, then truncation toward zero
    const double u = __dadd_rn(__dmul_rn((double)img_index[i], ratio), sx);
    const double v = __dadd_rn(__dmul_rn((double)img_index[ld + i], ratio), sy);
    const bool uok = u >= -9223372036854775808.0 && u < 9223372036854775808.0;
    const bool vok = v >= -9223372036854775808.0 && v < 9223372036854775808.0;
    img_index[i] = uok ? (long long)u : (long long)0x8000000000000000ull;
    img_index[ld + i] = vok ? (long long)v : (long long)0x8000000000000000ull;
}

unsigned blocks_for(long long n) { return (unsigned)((n + kThreads - 1) / kThreads > 0 ? (n + kThreads - 1) / kThreads : 1); }

}  // namespace

extern "C" int shpl_flip_point_cloud(double* points, int64_t point_stride, int64_t n, const int32_t* n_dev, void* stream) {
    const char* who = "shpl_flip_point_cloud";
    SHPL_REQUIRE(n >= 0 && n < (1ll << 31) && point_stride > 0, SHPL_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    SHPL_REQUIRE(n == 0 || points, SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    if (n == 0) return SHPL_OK;
    shpl_flip_kernel<<<blocks_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(points, point_stride, n, n_dev);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_flip_kernel");
}

extern "C" int shpl_mv3d_project_augment(double* lidar_pc, int64_t n, const int32_t* n_dev, const double* P_host,
                                         int32_t augment, double sx, double sz, double expansion_ratio,
                                         const double* rot_host, int64_t* img_index2_out, void* stream) {
    const char* who = "shpl_mv3d_project_augment";
    SHPL_REQUIRE(n >= 0 && n < (1ll << 31), SHPL_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    SHPL_REQUIRE(n == 0 || lidar_pc, SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    SHPL_REQUIRE(shpl::aligned(lidar_pc, 16), SHPL_ERR_INVALID_ARGUMENT, "%s: lidar_pc must be 16-byte aligned", who);
    SHPL_REQUIRE(!img_index2_out || P_host, SHPL_ERR_INVALID_ARGUMENT, "%s: img_index2 wanted but P is null", who);
    SHPL_REQUIRE(!augment || rot_host, SHPL_ERR_INVALID_ARGUMENT, "%s: augmenting but rot_mat is null", who);
    if (n == 0) return SHPL_OK;
    Mv3dAugArgs a{};
    a.pc = lidar_pc;
    a.n = n;
    a.n_dev = n_dev;
    if (P_host)
        for (int i = 0; i < 12; ++i) a.P[i] = P_host[i];
    a.augment = augment;
    a.sx = sx;
    a.sz = sz;
    a.ratio = expansion_ratio;
    if (rot_host)
        for (int i = 0; i < 4; ++i) a.rot[i] = rot_host[i];
    a.img_index2 = reinterpret_cast<long long*>(img_index2_out);
    shpl_mv3d_project_augment_kernel<<<blocks_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_mv3d_project_augment_kernel");
}

extern "C" int shpl_augment_fv_index(int64_t* img_index, int64_t ld, int64_t n, const int32_t* n_dev,
                                     double expansion_ratio, double sx, double sy, void* stream) {
    const char* who = "shpl_augment_fv_index";
    SHPL_REQUIRE(n >= 0 && n < (1ll << 31) && ld >= n, SHPL_ERR_INVALID_ARGUMENT, "%s: bad sizes", who);
    SHPL_REQUIRE(n == 0 || img_index, SHPL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
    if (n == 0) return SHPL_OK;
    shpl_augment_fv_kernel<<<blocks_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<long long*>(img_index), ld, n, n_dev, expansion_ratio, sx, sy);
    shpl::count_launches(1);
    return shpl::check_launch("shpl_augment_fv_kernel");
}
