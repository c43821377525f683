// A host program on the bare C ABI of include/shpl.h -- no Python, no PyTorch: what a C / C++ caller of libshpl.so
// writes.  It runs the avod-FPN two-layer SHPL step of bench.py (layer A: stride 8, dual, 88x100x256 <-> 45x150x256;
// layer B: stride 1, 700x800x32 <- 360x1200x32) on synthetic camera-frame points, end to end from pinned HOST buffers:
//   per step: H2D of the points / voxel indices, shpl_build_avod for both layers, forward + backward of both layers,
//             D2H of the plan counters and a few gradient values, read on the host one step later (double-buffered).
// It checks the result against a host recomputation from the plan it reads back (the pooled half of one listed set of
// cells: sum_k val_k * img[pix_k] in stored order, bit for bit) and prints frames/s.
//
//   make -C sparse_pooling_b200/csrc && make -C examples && examples/cabi_step [steps]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "shpl.h"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
            std::exit(2);                                                                       \
        }                                                                                       \
    } while (0)
#define SH(x)                                                                                   \
    do {                                                                                        \
        int rc_ = (x);                                                                          \
        if (rc_ != SHPL_OK) {                                                                   \
            std::fprintf(stderr, "%s:%d: shpl error %d: %s\n", __FILE__, __LINE__, rc_, shpl_last_error()); \
            std::exit(3);                                                                       \
        }                                                                                       \
    } while (0)

namespace {

template <typename T>
T* dev_alloc(size_t n) {
    T* p = nullptr;
    CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    return p;
}

struct Layer {
    int bev_h, bev_w, img_h, img_w, c_bev, c_img, s_img, s_bv, bv_h, bv_w;
    bool dual;
    shpl_plan plan{};
    void* ws = nullptr;
    size_t ws_bytes = 0;
    float *bev, *img, *fused_bev, *fused_img, *g_fused_bev, *g_fused_img, *g_bev, *g_img;
    int R() const { return bev_h * bev_w; }
    int Q() const { return img_h * img_w; }
};

uint64_t rng_state = 0x9e3779b97f4a7c15ull;
double uniform() {      // xorshift64*: the example needs plausible inputs, not a particular stream
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}

void fill_random(float* dev, size_t n) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(2.0 * uniform() - 1.0);
    CK(cudaMemcpy(dev, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
}

void make_layer(Layer& L, int n_max) {
    const int R = L.R(), Q = L.Q();
    shpl_plan& p = L.plan;
    p.n_rows = R;
    p.n_src = Q;
    p.capacity = n_max;
    p.row_ptr = dev_alloc<int32_t>(R + 1);
    p.pix_ptr = dev_alloc<int32_t>(Q + 1);
    p.csr_row = dev_alloc<int32_t>(n_max);
    p.csr_src = dev_alloc<int32_t>(n_max);
    p.csr_val = dev_alloc<float>(n_max);
    p.csrT_pix = dev_alloc<int32_t>(n_max);
    p.csrT_dst = dev_alloc<int32_t>(n_max);
    p.csrT_val = dev_alloc<float>(n_max);
    p.heavy_cap = 0;      // heavy_len = 0 below: every cell is summed sequentially in the main kernels
    p.counts = dev_alloc<int32_t>(8);
    CK(cudaMemset(p.counts, 0, 8 * sizeof(int32_t)));
    L.ws_bytes = shpl_build_workspace_bytes(n_max);
    CK(cudaMalloc(&L.ws, L.ws_bytes));
    const size_t nb = (size_t)R * L.c_bev, ni = (size_t)Q * L.c_img;
    L.bev = dev_alloc<float>(nb);
    L.img = dev_alloc<float>(ni);
    L.fused_bev = dev_alloc<float>((size_t)R * (L.c_bev + L.c_img));
    L.g_fused_bev = dev_alloc<float>((size_t)R * (L.c_bev + L.c_img));
    L.g_bev = dev_alloc<float>(nb);
    L.g_img = dev_alloc<float>(ni);
    L.fused_img = L.dual ? dev_alloc<float>((size_t)Q * (L.c_bev + L.c_img)) : nullptr;
    L.g_fused_img = L.dual ? dev_alloc<float>((size_t)Q * (L.c_bev + L.c_img)) : nullptr;
    fill_random(L.bev, nb);
    fill_random(L.img, ni);
    fill_random(L.g_fused_bev, (size_t)R * (L.c_bev + L.c_img));
    if (L.dual) fill_random(L.g_fused_img, (size_t)Q * (L.c_bev + L.c_img));
}

// one layer of one step: plan, forward, backward, all asynchronous on `s`
void run_layer(Layer& L, const double* pts, const int64_t* vox, int n, const double* P, cudaStream_t s) {
    const shpl_plan& p = L.plan;
    SH(shpl_build_avod(pts, vox, n, nullptr, P, 1200, 360, L.bv_h, L.bv_w, L.s_img, L.s_bv, nullptr, L.img_h, L.img_w, nullptr,
                       nullptr, nullptr, nullptr, &p, 0, 0, nullptr, L.ws, L.ws_bytes, s));
    if (L.dual) {
        SH(shpl_pool_forward_dual(L.bev, L.img, p.row_ptr, p.csr_row, p.csr_src, p.csr_val, p.pix_ptr, p.csrT_pix, p.csrT_dst,
                                  p.csrT_val, n, 0, L.R(), L.c_bev, L.Q(), L.c_img, L.fused_bev, L.fused_img, s));
        SH(shpl_pool_backward_dual(L.g_fused_bev, L.g_fused_img, p.row_ptr, p.csr_row, p.csr_src, p.csr_val, p.pix_ptr,
                                   p.csrT_pix, p.csrT_dst, p.csrT_val, n, 0, L.R(), L.c_bev, L.Q(), L.c_img, L.g_bev, L.g_img, s));
    } else {
        SH(shpl_pool_forward(L.bev, L.img, p.row_ptr, p.csr_row, p.csr_src, p.csr_val, n, 0, L.R(), L.c_bev, L.Q(), L.c_img,
                             L.fused_bev, s));
        SH(shpl_pool_backward(L.g_fused_bev, p.pix_ptr, p.csrT_pix, p.csrT_dst, p.csrT_val, n, 0, L.R(), L.c_bev, L.Q(),
                              L.c_img, L.g_bev, L.g_img, s));
    }
}

// host recomputation of the pooled half of the busiest cells of layer L from the plan read back: the same products,
// the same order of additions -> the same bits
int verify(const Layer& L) {
    const int R = L.R(), Q = L.Q(), Cb = L.c_bev, Ci = L.c_img;
    std::vector<int32_t> ptr(R + 1), counts(8);
    CK(cudaMemcpy(ptr.data(), L.plan.row_ptr, (R + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(counts.data(), L.plan.counts, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int nnz = counts[3];
    if (nnz <= 0 || ptr[R] != nnz || counts[2] != 0) {
        std::fprintf(stderr, "plan counters: clip %d nnz %d oob %d csr %d, row_ptr[R] %d\n", counts[0], counts[1], counts[2], counts[3], ptr[R]);
        return 1;
    }
    std::vector<int32_t> src(nnz);
    std::vector<float> val(nnz), img((size_t)Q * Ci), fused((size_t)R * (Cb + Ci)), bev((size_t)R * Cb);
    CK(cudaMemcpy(src.data(), L.plan.csr_src, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(val.data(), L.plan.csr_val, nnz * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(img.data(), L.img, img.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bev.data(), L.bev, bev.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fused.data(), L.fused_bev, fused.size() * sizeof(float), cudaMemcpyDeviceToHost));
    long long bad = 0, busy = 0;
    for (int r = 0; r < R; ++r) {
        const float* out = fused.data() + (size_t)r * (Cb + Ci);
        if (std::memcmp(out, bev.data() + (size_t)r * Cb, Cb * sizeof(float)) != 0) ++bad;      // the dense half
        busy += ptr[r + 1] > ptr[r];
        for (int c = 0; c < Ci; ++c) {
            volatile float acc = 0.f;                // volatile: no fma contraction, no reassociation
            for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
                volatile float prod = val[k] * img[(size_t)src[k] * Ci + c];
                acc = acc + prod;
            }
            const float a = acc;
            if (std::memcmp(&a, out + Cb + c, sizeof(float)) != 0) ++bad;
        }
    }
    std::printf("verify: %d x %d cells, %lld busy, nnz %d, mismatching values: %lld\n", L.bev_h, L.bev_w, busy, nnz, bad);
    return bad != 0;
}

}  // namespace

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 400;
    if (shpl_abi_version() != SHPL_ABI_VERSION) {
        std::fprintf(stderr, "libshpl.so ABI %d, header %d\n", shpl_abi_version(), SHPL_ABI_VERSION);
        return 1;
    }
    // synthetic frame: points on a ground plane and on obstacles in front of the camera, one point per BEV cell they hit
    const double P[12] = {721.5377, 0.0, 609.5593, 44.85728, 0.0, 721.5377, 172.854, 0.2163791, 0.0, 0.0, 1.0, 0.002745884};
    const int n_max = 24000, n_frames = 4;
    std::vector<std::vector<double>> pts(n_frames);
    std::vector<std::vector<int64_t>> vox(n_frames);
    for (int f = 0; f < n_frames; ++f) {
        while ((int)vox[f].size() / 2 < 21000) {
            const double z = 5.0 + 60.0 * uniform() * uniform(), half = 0.8 * z < 39.9 ? 0.8 * z : 39.9;
            const double x = (2.0 * uniform() - 1.0) * half, y = 1.65 - 2.0 * uniform() * (uniform() < 0.35);
            pts[f].insert(pts[f].end(), {x, y, z});
            vox[f].push_back((int64_t)std::floor((x + 40.0) / 0.1));
            vox[f].push_back((int64_t)(700 - (int64_t)std::floor(z / 0.1)));
        }
    }
    double* pin_pts[2];
    int64_t* pin_vox[2];
    int32_t* pin_res[2];
    double* d_pts[2];
    int64_t* d_vox[2];
    for (int b = 0; b < 2; ++b) {
        CK(cudaMallocHost(&pin_pts[b], n_max * 3 * sizeof(double)));
        CK(cudaMallocHost(&pin_vox[b], n_max * 2 * sizeof(int64_t)));
        CK(cudaMallocHost(&pin_res[b], 64 * sizeof(int32_t)));
        d_pts[b] = dev_alloc<double>(n_max * 3);
        d_vox[b] = dev_alloc<int64_t>(n_max * 2);
    }
    // two buffer sets of two layers: two frames in flight
    Layer layers[2][2];
    for (int b = 0; b < 2; ++b) {
        layers[b][0] = Layer{88, 100, 45, 150, 256, 256, 8, 8, 704, 800, true};
        layers[b][1] = Layer{700, 800, 360, 1200, 32, 32, 1, 1, 700, 800, false};
        for (int l = 0; l < 2; ++l) make_layer(layers[b][l], n_max);
    }
    cudaStream_t lane[2], side[2];
    cudaEvent_t fork[2], join[2], done[2], t0, t1;
    for (int b = 0; b < 2; ++b) {
        CK(cudaStreamCreate(&lane[b]));
        CK(cudaStreamCreate(&side[b]));
        CK(cudaEventCreateWithFlags(&fork[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&join[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    }
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));

    long long seen = 0;
    auto step = [&](int k) {
        const int b = k & 1, f = k % n_frames;
        const int n = (int)vox[f].size() / 2;
        if (k >= 2) CK(cudaEventSynchronize(done[b]));          // the pinned buffers of set b are free again
        std::memcpy(pin_pts[b], pts[f].data(), n * 3 * sizeof(double));           // "the loader" refills them
        std::memcpy(pin_vox[b], vox[f].data(), n * 2 * sizeof(int64_t));
        CK(cudaMemcpyAsync(d_pts[b], pin_pts[b], n * 3 * sizeof(double), cudaMemcpyHostToDevice, lane[b]));
        CK(cudaMemcpyAsync(d_vox[b], pin_vox[b], n * 2 * sizeof(int64_t), cudaMemcpyHostToDevice, lane[b]));
        CK(cudaEventRecord(fork[b], lane[b]));
        CK(cudaStreamWaitEvent(side[b], fork[b], 0));
        run_layer(layers[b][0], d_pts[b], d_vox[b], n, P, side[b]);              // layer A beside layer B
        run_layer(layers[b][1], d_pts[b], d_vox[b], n, P, lane[b]);
        CK(cudaEventRecord(join[b], side[b]));
        CK(cudaStreamWaitEvent(lane[b], join[b], 0));
        for (int l = 0; l < 2; ++l) {
            CK(cudaMemcpyAsync(pin_res[b] + 16 * l, layers[b][l].plan.counts, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, lane[b]));
            CK(cudaMemcpyAsync(pin_res[b] + 16 * l + 8, layers[b][l].g_img, 8 * sizeof(float), cudaMemcpyDeviceToHost, lane[b]));
        }
        CK(cudaEventRecord(done[b], lane[b]));
        if (k >= 1) {                                           // the host read: step k-1's result
            CK(cudaEventSynchronize(done[1 - b]));
            seen += pin_res[1 - b][1] > 0 && pin_res[1 - b][17] > 0;
        }
    };
    for (int k = 0; k < 8; ++k) step(k);
    CK(cudaDeviceSynchronize());
    int rc = verify(layers[1][0]) | verify(layers[1][1]);
    seen = 0;
    const uint64_t launches0 = shpl_kernel_launches();
    CK(cudaEventRecord(t0, lane[0]));
    for (int k = 0; k < steps; ++k) step(k + 8);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(t1, lane[0]));
    CK(cudaEventSynchronize(t1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    if (seen != steps) {
        std::fprintf(stderr, "only %lld of %d results were read back with sane counters\n", seen, steps);
        rc = 1;
    }
    std::printf("{\"what\": \"C ABI from a C++ host, pinned host buffers in, results read back every step, 2 frames in flight\", "
                "\"steps\": %d, \"us_per_frame\": %.1f, \"frames_per_s\": %.0f, \"kernel_launches_per_frame\": %.1f, \"ok\": %s}\n",
                steps, ms * 1e3 / steps, steps / (ms * 1e-3), (double)(shpl_kernel_launches() - launches0) / steps, rc ? "false" : "true");
    return rc;
}
