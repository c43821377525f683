/* CPU oracle (plain C) for the SHPL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restates, as sequential loops, the algorithm of
 *   /root/reference/avod/avod/utils/sparse_pool_utils.py:6-58   (index path)
 *   /root/reference/avod/avod/utils/transform.py:3-40           (projection / clip)
 *   /root/reference/avod/avod/utils/sparse_pool_utils.py:61-117 (value path: the TF ops it calls)
 * following SURVEY.md Appendix A.  Used (a) by tests/ as the full-size checker of
 * the CUDA kernels, (b) by bench.py's cpu_baseline / --impl reference legs.  The
 * product (sparse_pooling_b200/) never links or loads this file.
 *
 * Index path: PINNED against the reference's numpy output (tests/golden, KAT-1/2).
 * Value path: PARITY UNPINNED at the TensorFlow 1.8 boundary (no reference golden
 * exists) -- it follows TF's documented op semantics: zero-initialised
 * fp32 output, contributions added in COO order, product rounded before the add.
 *
 * Build:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile).
 * -ffp-contract=off keeps  acc + w*x  as two roundings; the projection spells its
 * fused chain out with fma() because that is what the reference's BLAS call does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int shpl_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* floor division for a positive divisor (np.floor(a / s) on integer-valued doubles) */
static int64_t floordiv(int64_t a, int64_t s) {
    int64_t q = a / s;
    if ((a % s != 0) && ((a < 0) != (s < 0))) --q;
    return q;
}

/* transform.py:17-24 -- one row of P times [x y z 1], rounded the way the
 * reference's np.dot (OpenBLAS dgemm) rounds it: product, then three fma. */
static double prow(const double* p, double x, double y, double z) {
    double t = p[0] * x;
    t = fma(p[1], y, t);
    t = fma(p[2], z, t);
    t = fma(p[3], 1.0, t);
    return t;
}

/* gen_sparse_pooling_input_avod (:6-20) + produce_sparse_pooling_input (:22-58),
 * fused over one frame.  Outputs sized for N entries by the caller.
 *   points  f64 [N,3], vox i64 [N,2] (x, zflip), P f64 [12] row-major
 *   counts[0] = n  (pairs that survive the image clip, :14)
 *   counts[1] = nnz (pairs that survive row < R', :44)
 *   gen_bv   i64 [n,2], gen_img f64 [3,n]  -- the dict of :20, BEFORE the in-place floor
 *   mij      i64 [nnz,2], flip i64 [nnz,3]
 *   msize    i64 [2]
 */
void shpl_oracle_build_avod(const double* points, const int64_t* vox, int64_t N, const double* P,
                            int64_t im_w, int64_t im_h, int64_t bv_h, int64_t bv_w,
                            int64_t s_img, int64_t s_bv,
                            int64_t* counts, int64_t* gen_bv, double* gen_img_u, double* gen_img_v,
                            int64_t* mij, int64_t* flip, int64_t* msize) {
    const int64_t Wp = floordiv(im_w, s_img), Hp = floordiv(im_h, s_img);
    const int64_t Hb = floordiv(bv_h, s_bv), Wb = floordiv(bv_w, s_bv);
    const int64_t R = Hb * Wb;
    int64_t n = 0, nnz = 0;
    for (int64_t i = 0; i < N; ++i) {
        const double x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
        const double w = prow(P + 8, x, y, z);
        const double u = prow(P + 0, x, y, z) / w;
        const double v = prow(P + 4, x, y, z) / w;
        /* transform.py:36-38 */
        if (!(u < (double)(im_w - 1) && u >= 0.0 && v >= 0.0 && v < (double)(im_h - 1))) continue;
        const int64_t ui = (int64_t)rint(u), vi = (int64_t)rint(v);   /* :18 ties-to-even */
        if (gen_bv) { gen_bv[2 * n] = vox[2 * i]; gen_bv[2 * n + 1] = vox[2 * i + 1]; }
        if (gen_img_u) { gen_img_u[n] = (double)ui; gen_img_v[n] = (double)vi; }
        ++n;
        /* :30-34 */
        int64_t up = floordiv(ui, s_img), vp = floordiv(vi, s_img);
        if (up >= Wp) up = Wp - 1;
        if (vp >= Hp) vp = Hp - 1;
        /* :38-44 */
        const int64_t xp = floordiv(vox[2 * i], s_bv), zp = floordiv(vox[2 * i + 1], s_bv);
        const int64_t row = zp * Wb + xp;
        if (!(row < R)) continue;
        mij[2 * nnz] = row; mij[2 * nnz + 1] = nnz;                   /* :50 */
        flip[3 * nnz] = 0; flip[3 * nnz + 1] = vp; flip[3 * nnz + 2] = up;   /* :36 */
        ++nnz;
    }
    counts[0] = n; counts[1] = nnz;
    msize[0] = R; msize[1] = nnz;                                      /* :52 */
}

/* tf.concat([a, b], axis=3) on [R, Ca] and [R, Cb] */
static void concat_rows(const float* a, int64_t Ca, const float* b, int64_t Cb, int64_t R, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; ++r) {
        memcpy(out + r * (Ca + Cb), a + r * Ca, sizeof(float) * Ca);
        memcpy(out + r * (Ca + Cb) + Ca, b + r * Cb, sizeof(float) * Cb);
    }
}

static void zero_rows(float* p, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (n + 4095) / 4096; ++i) {
        int64_t lo = i * 4096, hi = lo + 4096 < n ? lo + 4096 : n;
        memset(p + lo, 0, sizeof(float) * (hi - lo));
    }
}

/* Forward, one direction (sparse_pool_utils.py:65-72 with :96-103):
 *   G = gather_nd(src, pix)         [m, Cs]
 *   Y = SpMM(M, G)                  [R, Cs]  zero-init, COO order e = 0..m-1
 *   fused = concat(dst, Y)          [R, Cd+Cs]
 * rows/cols are M's COO indices; pix[c] is the linear source pixel of column c.
 * scratch: G [ncols*Cs] and Y [R*Cs] floats (caller-provided so timing excludes malloc).
 * Entries with row/col/pix out of range are skipped (TF-CPU would raise). */
void shpl_oracle_forward(const float* dst, const float* src, int64_t m, const int64_t* rows, const int64_t* cols,
                         const float* vals, int64_t ncols, const int64_t* pix,
                         int64_t R, int64_t Cd, int64_t Q, int64_t Cs,
                         float* G, float* Y, float* fused) {
    for (int64_t c = 0; c < ncols; ++c) {
        if (pix[c] >= 0 && pix[c] < Q) memcpy(G + c * Cs, src + pix[c] * Cs, sizeof(float) * Cs);
        else memset(G + c * Cs, 0, sizeof(float) * Cs);
    }
    zero_rows(Y, R * Cs);
    for (int64_t e = 0; e < m; ++e) {
        const int64_t r = rows[e], c = cols[e];
        if (r < 0 || r >= R || c < 0 || c >= ncols || pix[c] < 0 || pix[c] >= Q) continue;
        const float w = vals[e];
        float* y = Y + r * Cs;
        const float* g = G + c * Cs;
        for (int64_t j = 0; j < Cs; ++j) y[j] = y[j] + w * g[j];
    }
    concat_rows(dst, Cd, Y, Cs, R, fused);
}

/* Backward of the above:
 *   g_dst = g_fused[:, :Cd]                          (concat grad)
 *   gG[c] = sum_e{col_e = c} w_e * g_fused[row_e, Cd:]   (A^T g, COO order)
 *   g_src = scatter_nd(pix, gG)  zero-init, column order c = 0..ncols-1 */
void shpl_oracle_backward(const float* g_fused, int64_t m, const int64_t* rows, const int64_t* cols,
                          const float* vals, int64_t ncols, const int64_t* pix,
                          int64_t R, int64_t Cd, int64_t Q, int64_t Cs,
                          float* gG, float* g_dst, float* g_src) {
    const int64_t F = Cd + Cs;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; ++r) memcpy(g_dst + r * Cd, g_fused + r * F, sizeof(float) * Cd);
    zero_rows(gG, ncols * Cs);
    for (int64_t e = 0; e < m; ++e) {
        const int64_t r = rows[e], c = cols[e];
        if (r < 0 || r >= R || c < 0 || c >= ncols || pix[c] < 0 || pix[c] >= Q) continue;
        const float w = vals[e];
        const float* g = g_fused + r * F + Cd;
        float* o = gG + c * Cs;
        for (int64_t j = 0; j < Cs; ++j) o[j] = o[j] + w * g[j];
    }
    zero_rows(g_src, Q * Cs);
    for (int64_t c = 0; c < ncols; ++c) {
        if (pix[c] < 0 || pix[c] >= Q) continue;
        float* o = g_src + pix[c] * Cs;
        const float* g = gG + c * Cs;
        for (int64_t j = 0; j < Cs; ++j) o[j] = o[j] + g[j];
    }
}

/* Forward, reverse direction (sparse_pool_utils.py:79-87 with :105-117):
 *   S[c] = sum_e{col_e=c} w_e * bev[row_e]      (sparse_transpose + SpMM)
 *   P    = scatter_nd(pix, S)  [Q, Cb]          duplicates summed in column order
 *   fused_i = concat(img, P)
 * scratch S [ncols*Cb], Pm [Q*Cb]. */
void shpl_oracle_forward_trans(const float* img, const float* bev, int64_t m, const int64_t* rows,
                               const int64_t* cols, const float* vals, int64_t ncols, const int64_t* pix,
                               int64_t R, int64_t Cb, int64_t Q, int64_t Ci,
                               float* S, float* Pm, float* fused_i) {
    zero_rows(S, ncols * Cb);
    for (int64_t e = 0; e < m; ++e) {
        const int64_t r = rows[e], c = cols[e];
        if (r < 0 || r >= R || c < 0 || c >= ncols || pix[c] < 0 || pix[c] >= Q) continue;
        const float w = vals[e];
        const float* b = bev + r * Cb;
        float* o = S + c * Cb;
        for (int64_t j = 0; j < Cb; ++j) o[j] = o[j] + w * b[j];
    }
    zero_rows(Pm, Q * Cb);
    for (int64_t c = 0; c < ncols; ++c) {
        if (pix[c] < 0 || pix[c] >= Q) continue;
        float* o = Pm + pix[c] * Cb;
        const float* s = S + c * Cb;
        for (int64_t j = 0; j < Cb; ++j) o[j] = o[j] + s[j];
    }
    concat_rows(img, Ci, Pm, Cb, Q, fused_i);
}

/* Backward of the reverse direction:
 *   g_img = g_fused_i[:, :Ci]
 *   gS[c] = g_fused_i[pix_c, Ci:]                  (scatter_nd grad = gather_nd)
 *   g_bev[row_e] += w_e * gS[col_e]  zero-init, COO order   ((A^T)^T g) */
void shpl_oracle_backward_trans(const float* g_fused_i, int64_t m, const int64_t* rows, const int64_t* cols,
                                const float* vals, int64_t ncols, const int64_t* pix,
                                int64_t R, int64_t Cb, int64_t Q, int64_t Ci,
                                float* g_img, float* g_bev) {
    const int64_t F = Ci + Cb;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < Q; ++q) memcpy(g_img + q * Ci, g_fused_i + q * F, sizeof(float) * Ci);
    zero_rows(g_bev, R * Cb);
    for (int64_t e = 0; e < m; ++e) {
        const int64_t r = rows[e], c = cols[e];
        if (r < 0 || r >= R || c < 0 || c >= ncols || pix[c] < 0 || pix[c] >= Q) continue;
        const float w = vals[e];
        const float* g = g_fused_i + pix[c] * F + Ci;
        float* o = g_bev + r * Cb;
        for (int64_t j = 0; j < Cb; ++j) o[j] = o[j] + w * g[j];
    }
}
