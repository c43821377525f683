/* shpl.h -- C ABI of libshpl.so: the B200 (sm_100a) implementation of the Sparse
 * Non-homogeneous Pooling Layer (SHPL) hot path of YeungLy/Sparse_Pooling.
 *
 * The reference has no native boundary for this path (it is numpy + TensorFlow
 * graph ops); each entry point below names the reference interface it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (PyTorch in this repo); the library keeps no
 *     state between calls and allocates nothing;
 *   - every function takes the cudaStream_t (as void*) to launch on, is
 *     asynchronous and never synchronises;
 *   - return value: 0 on success, negative shpl_status on failure, with a
 *     thread-local message available from shpl_last_error();
 *   - feature maps are NHWC fp32, contiguous, channel counts multiples of 4 and
 *     base addresses 16-byte aligned (128-bit vector access).
 */
#ifndef SHPL_H_
#define SHPL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHPL_ABI_VERSION 10

/* Cells with more entries than this are "heavy": the builder lists them in the plan and their sum
 * is formed by shpl_pool_heavy (a thread-block cluster per cell) instead of one warp walking the cell.
 * Listed cells of up to SHPL_EXACT_LEN entries are still summed in the reference's sequential order
 * (bit-exact: the cluster gathers and multiplies in parallel, one warp per 32 channel vectors adds in
 * entry order); longer ones by a fixed summation tree (deterministic, within 1e-5 of the sum of |terms|).
 */
#define SHPL_HEAVY_LEN 512
#define SHPL_EXACT_LEN 2048
/* Cells with more entries than this are "long": the builder counts them (counts[6], counts[7]); the pooling kernels sum them
 * through shared memory (a whole CTA gathers in parallel, the additions stay in entry order) instead of one warp walking them. */
#define SHPL_LONG_LEN 32

typedef enum shpl_status {
    SHPL_OK = 0,
    SHPL_ERR_INVALID_ARGUMENT = -1, /* reference: Python assert / TF InvalidArgumentError */
    SHPL_ERR_CUDA = -2,
    SHPL_ERR_WORKSPACE_TOO_SMALL = -3,
    SHPL_ERR_UNSUPPORTED = -4
} shpl_status;

/* Canonical sparse structure of one M (one frame, or several frames stacked).
 * It is the stable sort of the reference's COO  tf.SparseTensor(Mij_pool, M_val,
 * M_size)  (avod/avod/core/models/rpn_model.py:292-293, :330-331) by destination
 * row (CSR) and by source pixel (CSR^T); inside a row / pixel the entries keep the
 * reference's column order k, which is TF-CPU's accumulation order.
 * All arrays are caller-allocated device memory of the stated capacity. */
typedef struct shpl_plan {
    int32_t  n_rows;    /* R  = H_b' * W_b' * frames   (destination cells)            */
    int32_t  n_src;     /* Q  = H_i' * W_i' * frames   (source pixels)                */
    int32_t  capacity;  /* entries the arrays below can hold (>= candidate pairs)     */
    int32_t* row_ptr;   /* [n_rows+1]  CSR offsets by destination row                 */
    int32_t* csr_row;   /* [capacity]  destination row of each entry (the sort key):
                           lets the kernels split work by ENTRY, not by row          */
    int32_t* csr_src;   /* [capacity]  linear source pixel of each entry              */
    float*   csr_val;   /* [capacity]  non-homogeneous weight of each entry           */
    int32_t* pix_ptr;   /* [n_src+1]   CSR^T offsets by source pixel                  */
    int32_t* csrT_pix;  /* [capacity]  source pixel of each entry (the sort key)      */
    int32_t* csrT_dst;  /* [capacity]  destination row of each entry                  */
    float*   csrT_val;  /* [capacity]                                                 */
    int32_t  heavy_cap; /* capacity of heavy_row / heavy_pix (0: do not list heavy cells)  */
    int32_t* heavy_row; /* [heavy_cap] destination rows with more than SHPL_HEAVY_LEN entries */
    int32_t* heavy_pix; /* [heavy_cap] source pixels with more than SHPL_HEAVY_LEN entries    */
    int32_t* heavy_count; /* [2] device: number of listed rows, pixels (all stacked frames)  */
    int32_t* counts;    /* [8] device: [0]=n after image clip, [1]=nnz (columns of M),
                           [2]=entries left out of the CSRs because an index is out of
                           range (TF-CPU raises InvalidArgumentError for those),
                           [3]=entries in the CSRs; [4]=running entry offset after this
                           frame (the next stacked frame starts there); [5]=1 if this
                           frame's entries did not fit the plan behind the frames stacked
                           before it (capacity < entry base + n: nothing is written out
                           of bounds, the entries are dropped); [6], [7] = destination
                           rows / source pixels of this frame with more than SHPL_LONG_LEN
                           entries: a caller that reads 0 may pass heavy_len = 0 to the
                           pooling entry points of that direction                       */
} shpl_plan;

int         shpl_abi_version(void);
const char* shpl_last_error(void);

/* Number of CUDA kernels this library has launched in this process so far (all
 * threads).  bench.py differences it around the timed region ("gpu_launches"). */
uint64_t shpl_kernel_launches(void);

/* Memory-safety evidence without compute-sanitizer: `make -C sparse_pooling_b200/csrc debug` builds libshpl_debug.so with
 * in-kernel checks of every gather index, entry range and output slot against the caller's sizes.
 * shpl_debug_checks_enabled(): 1 in that build, 0 in the product library (where the checks are compiled out).
 * shpl_debug_check_failures(): failed checks so far (synchronises the device; always 0 in the product library). */
int     shpl_debug_checks_enabled(void);
int64_t shpl_debug_check_failures(void);

/* Bytes of scratch the shpl_build_* / shpl_plan_from_coo calls need for up to
 * n_max candidate pairs. */
size_t shpl_build_workspace_bytes(int64_t n_max);

/* gen_sparse_pooling_input_avod  (avod/avod/utils/sparse_pool_utils.py:6-20) with
 * projectToImage / clip3DwithinImage (avod/avod/utils/transform.py:3-40) inlined.
 *   points f64 [N,3] camera frame, voxel_indices i64 [N,2] = (x, zflip),
 *   P_host f64 [12] = stereo_calib.p2 row-major, image size (W, H) in pixels.
 *   N_dev (device int32*, may be NULL): when given, only the first min(N, *N_dev) candidates exist --
 *   lets the feeder's device-side pair count (shpl_bev_slices counts[0]) flow in without a host read.
 * Writes the n surviving pairs, input order kept:
 *   bv_index_out i64 [N,2] (first n rows), img_u_out / img_v_out f64 [N] (first n;
 *   rows 0 and 1 of the reference's [3,n] img_index; row 2 is all zero),
 *   counts[0] = n. */
int shpl_gen_input_avod(const double* points, const int64_t* voxel_indices, int64_t N, const int32_t* N_dev,
                        const double* P_host, int32_t im_w, int32_t im_h,
                        int64_t* bv_index_out, double* img_u_out, double* img_v_out,
                        int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* produce_sparse_pooling_input  (avod/avod/utils/sparse_pool_utils.py:22-58; twin
 * MV3D_TF_release/lib/utils/sparse_pool_utils.py:22-54; MV3D entry
 * MV3D_TF_release/lib/networks/MV3D_voxel_train.py:89-91), plus the canonical
 * CSR / CSR^T of the resulting M.
 *   img_u, img_v f64 [n]  rows 0,1 of img_index -- FLOORED AND CLAMPED IN PLACE like
 *                         the reference does (:30-34);
 *   bv_index i64 [n,2]; image size (W,H); BEV size (H,W); stride_img = stride[0],
 *   stride_bv = stride[1] (the reference's comment has them swapped, :23 vs :30,:38);
 *   m_val f64 [>=nnz] or NULL (-> ones), indexed by output column k like the
 *   reference (:56-57; it is NOT filtered by the row test);
 *   src_h, src_w: height/width of the feature map that will be gathered from; pass
 *   0,0 to use floor(H/stride_img), floor(W/stride_img).
 * Outputs (any of the first four may be NULL):
 *   Mij_pool i64 [n,2], img_index_flip_pool i64 [n,3], M_val_out f32 [n]
 *   (first nnz rows valid), M_size_out i64 [2] (device), plan (see shpl_plan).
 * row_base / pix_base are added to every destination row / source pixel and
 * entry_base_dev (device int32*, may be NULL = 0) to every CSR offset, so that the
 * frames of a batch can be stacked into one plan by consecutive calls;
 * plan->row_ptr then points at the sub-array of this frame. */
int shpl_produce_input(double* img_u, double* img_v, const int64_t* bv_index, int64_t n,
                       int32_t im_w, int32_t im_h, int32_t bv_h, int32_t bv_w,
                       int32_t stride_img, int32_t stride_bv, const double* m_val,
                       int32_t src_h, int32_t src_w,
                       int64_t* Mij_pool, int64_t* img_index_flip_pool, float* M_val_out, int64_t* M_size_out,
                       const shpl_plan* plan, int32_t row_base, int32_t pix_base, const int32_t* entry_base_dev,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The two functions above fused (the call kitti_dataset.py:376-378 makes per
 * sample): no intermediate dict is materialised. */
int shpl_build_avod(const double* points, const int64_t* voxel_indices, int64_t N, const int32_t* N_dev,
                    const double* P_host, int32_t im_w, int32_t im_h, int32_t bv_h, int32_t bv_w,
                    int32_t stride_img, int32_t stride_bv, const double* m_val,
                    int32_t src_h, int32_t src_w,
                    int64_t* Mij_pool, int64_t* img_index_flip_pool, float* M_val_out, int64_t* M_size_out,
                    const shpl_plan* plan, int32_t row_base, int32_t pix_base, const int32_t* entry_base_dev,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Plan from an arbitrary COO -- the tf.SparseTensor + gather index that crosses the
 * reference's host->device boundary (placeholders avod/avod/core/models/rpn_model.py:219-242):
 *   Mij i64 [m,2] = (row, col), val f32 [m], source_index [ncol,3] = (0, v, u) as int32
 *   (index_is_i64 = 0) or int64 (= 1).  Entry e gathers pixel v*src_w+u of column col_e.
 * counts[1] = m, counts[2] = entries with an out-of-range row / col / pixel. */
int shpl_plan_from_coo(const int64_t* Mij, const float* val, int64_t m,
                       const void* source_index, int32_t index_is_i64, int64_t ncol,
                       int32_t src_h, int32_t src_w,
                       const shpl_plan* plan, int32_t row_base, int32_t pix_base, const int32_t* entry_base_dev,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Plan of the VFE scatter: tf.scatter_nd(coordinate, voxelwise, [B, 10, H, W, 128])
 * (MV3D_TF_release/lib/networks/group_pointcloud.py:84-85; coordinate rows are
 * (batch, d, h, w), built by build_input :88-105 from the feeder's coordinate_buffer).
 *   coordinate [K,4] int32 (index_is_i64 = 0) or int64 (= 1), 16-byte aligned;
 *   K_dev (device int32*, may be NULL): only the first min(K, *K_dev) rows exist
 *   (the voxel count shpl_mv3d_voxelize leaves on the device).
 * Row k of the feature matrix becomes the single entry (cell(coordinate[k]), k) with
 * weight 1; plan->n_rows must equal batch*depth*height*width and plan->n_src >= K.
 * The scatter itself is shpl_pool_forward with C_d = 0 on the CSR arrays (duplicate
 * coordinates are summed in row order k, like TF-CPU), its gradient -- a gather of
 * the grid gradient at the coordinates -- shpl_pool_backward on the CSR^T arrays.
 * counts[1] = K, counts[2] = rows whose coordinate is outside the grid. */
int shpl_plan_from_voxel_coords(const void* coordinate, int32_t index_is_i64, int64_t K, const int32_t* K_dev,
                                int32_t batch, int32_t depth, int32_t height, int32_t width,
                                const shpl_plan* plan, void* workspace, size_t workspace_bytes, void* stream);

/* Forward of one direction: _sparse_pool_op + tf.concat
 * (avod/avod/utils/sparse_pool_utils.py:96-103 with :67-72; for the reverse
 * direction _sparse_pool_trans_op :105-117 with :82-87 -- pass the CSR^T arrays):
 *   fused[r, 0:C_d]       = dst[r, :]
 *   fused[r, C_d:C_d+C_s] = sum over the entries of row r, in stored order, of
 *                           val * src[idx, :]      (0 for an empty row)
 * dst [n_rows, C_d], src [n_src, C_s], fused [n_rows, C_d+C_s], all fp32.
 * dst may be NULL with C_d = 0 (pooled map only: _sparse_pool_op without concat).
 * key [nnz] = destination row of each entry (plan.csr_row / plan.csrT_pix) and
 * nnz_max >= number of entries (e.g. plan.capacity) let the kernel balance the
 * gathers by entry; key may be NULL (then rows are walked cell by cell).
 * heavy_len > 0: cells with more than heavy_len entries are NOT summed here and their pooled part (in the
 * add forms: their whole output) is left UNWRITTEN; the caller runs shpl_pool_heavy / shpl_pool_heavy_split on
 * the listed cells -- after this call, or concurrently on another stream (the two write disjoint cells).
 * Pass SHPL_HEAVY_LEN for a plan with listed cells.  (For C_s <= 128 the kernels keep listed cells of up to
 * SHPL_EXACT_LEN entries themselves -- summed in entry order through shared memory -- and the heavy entry points skip
 * exactly those: the rule depends on C_s alone, so the two sides always agree.)  A value above SHPL_HEAVY_LEN (e.g.
 * SHPL_EXACT_LEN) says: some cells have more than SHPL_LONG_LEN entries but none is listed -- nothing is left out, the
 * long-cell paths stay on.
 * heavy_len = 0: every cell is summed here, strictly sequentially, and the kernels skip their long-cell paths: right
 * for plans whose counts[6] / counts[7] read 0 (every KITTI / MV3D plan). */
int shpl_pool_forward(const float* dst, const float* src,
                      const int32_t* ptr, const int32_t* key, const int32_t* idx, const float* val,
                      int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_d, int32_t n_src, int32_t C_s,
                      float* fused, void* stream);

/* Backward of shpl_pool_forward (what TF autodiff derives, SURVEY.md row a13;
 * trainer.py:91-95 builds it): deterministic, no atomics.
 *   g_dst[r, :] = g_fused[r, 0:C_d]
 *   g_src[p, :] = sum over the entries of source p (transposed arrays: ptrT, idxT =
 *                 destination row, valT), in stored order, of valT * g_fused[idxT, C_d:]
 * g_dst may be NULL (no slice copy).  keyT / nnz_max as in shpl_pool_forward. */
int shpl_pool_backward(const float* g_fused,
                       const int32_t* ptrT, const int32_t* keyT, const int32_t* idxT, const float* valT,
                       int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_d, int32_t n_src, int32_t C_s,
                       float* g_dst, float* g_src, void* stream);

/* Both directions of sparse_pool_layer in ONE launch (the `bv_index is not None` branch,
 * avod/avod/utils/sparse_pool_utils.py:79-87, on top of :65-72):
 *   fused_bev[r] = concat(bev[r], sum_{k in row r}   val_k * img[pix_k])     [n_rows, C_b+C_i]
 *   fused_img[p] = concat(img[p], sum_{k at pixel p} val_k * bev[row_k])     [n_src,  C_i+C_b]
 * The eight index arrays are the fields of shpl_plan. */
int shpl_pool_forward_dual(const float* bev, const float* img,
                           const int32_t* row_ptr, const int32_t* csr_row, const int32_t* csr_src, const float* csr_val,
                           const int32_t* pix_ptr, const int32_t* csrT_pix, const int32_t* csrT_dst, const float* csrT_val,
                           int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_b, int32_t n_src, int32_t C_i,
                           float* fused_bev, float* fused_img, void* stream);

/* Backward of shpl_pool_forward_dual in one launch.  Each input of the layer feeds two
 * consumers, so TF adds two partial gradients (AddN); here the sum is formed in registers:
 *   g_bev[r] = g_fused_bev[r, :C_b] + sum_{k in row r}   val_k * g_fused_img[pix_k, C_i:]
 *   g_img[p] = g_fused_img[p, :C_i] + sum_{k at pixel p} val_k * g_fused_bev[row_k, C_b:]
 * (the pooled sum is accumulated from zero in stored order, then added to the slice: the same
 * roundings as slice + pooled computed separately). */
int shpl_pool_backward_dual(const float* g_fused_bev, const float* g_fused_img,
                            const int32_t* row_ptr, const int32_t* csr_row, const int32_t* csr_src, const float* csr_val,
                            const int32_t* pix_ptr, const int32_t* csrT_pix, const int32_t* csrT_dst, const float* csrT_val,
                            int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_b, int32_t n_src, int32_t C_i,
                            float* g_bev, float* g_img, void* stream);

/* shpl_pool_heavy with the LONG listed cells (more than SHPL_EXACT_LEN entries) split over many CTAs instead of one
 * cluster per cell: a cell of L entries is cut into ceil(L / 256) contiguous pieces, a CTA sums one piece in stored
 * order into the workspace, and a second kernel adds a cell's partial sums as a two-level tree (32 groups in order side by
 * side, then the group sums in order) -- a fixed tree that depends only on L (deterministic, within 1e-5 of the sum of
 * |terms|; not bit-identical to the sequential sum).  The Zipf stress case's
 * 178 000-entry cell runs on ~90 SMs instead of 8.  Cells up to SHPL_EXACT_LEN entries take the exact cluster kernel as
 * in shpl_pool_heavy -- for C > 128; for C <= 128 the main kernels have summed them already (see shpl_pool_forward) and
 * both heavy entry points skip them.  nnz_max >= total entries of the plan; workspace: shpl_pool_heavy_workspace_bytes(C, nnz_max,
 * list_cap) bytes, 16-byte aligned.  (More than 4096 listed cells: falls back to shpl_pool_heavy's cluster tree.) */
size_t shpl_pool_heavy_workspace_bytes(int32_t C, int64_t nnz_max, int32_t list_cap);
int shpl_pool_heavy_split(const float* gather_in, int32_t gather_stride, int32_t C,
                          const int32_t* ptr, const int32_t* idx, const float* val,
                          const int32_t* list, const int32_t* count_dev, int32_t list_cap,
                          const float* addend, int32_t addend_stride,
                          float* out, int32_t out_stride, int64_t nnz_max,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- No-concat ("sparse-only") forms of the three calls above: for callers whose producer of the
 * destination map writes its C_d channels straight into the fused buffer (what tf.concat would otherwise copy).
 * Two thirds of the bytes of the KITTI pre-RPN forward are that copy (220 MB -> 74 MB).
 *   shpl_pool_forward_into:   fused[r*fused_stride + chan_off + 0:C_s] = sum over the entries of row r, in stored order,
 *                             of val * src[idx, :]  (0 for an empty row); every other channel of `fused` is left alone.
 *   shpl_pool_forward_into_dual: both directions of the layer in one launch, into fused_bev [n_rows, C_b+C_i] at channel
 *                             C_b and fused_img [n_src, C_i+C_b] at channel C_i (bev / img are only gathered from).
 *   shpl_pool_backward_from:  g_src[p, :] = sum over the entries of source p of valT * g_fused[idxT*g_stride + chan_off + :]
 *                             -- the gradient of the gathered map; the gradient of the destination map is the VIEW
 *                             g_fused[:, 0:chan_off] (no copy is made, nothing to compute).
 * Same values, bit for bit, as the concat forms; key / nnz_max / heavy_len as in shpl_pool_forward (for heavy cells follow
 * with shpl_pool_heavy, whose strides are general). */
int shpl_pool_forward_into(const float* src,
                           const int32_t* ptr, const int32_t* key, const int32_t* idx, const float* val,
                           int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t n_src, int32_t C_s,
                           float* fused, int32_t fused_stride, int32_t chan_off, void* stream);
int shpl_pool_forward_into_dual(const float* bev, const float* img,
                                const int32_t* row_ptr, const int32_t* csr_row, const int32_t* csr_src, const float* csr_val,
                                const int32_t* pix_ptr, const int32_t* csrT_pix, const int32_t* csrT_dst, const float* csrT_val,
                                int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t C_b, int32_t n_src, int32_t C_i,
                                float* fused_bev, float* fused_img, void* stream);
int shpl_pool_backward_from(const float* g_fused, int32_t g_stride, int32_t chan_off,
                            const int32_t* ptrT, const int32_t* keyT, const int32_t* idxT, const float* valT,
                            int32_t nnz_max, int32_t heavy_len, int32_t n_rows, int32_t n_src, int32_t C_s,
                            float* g_src, void* stream);

/* Listed ("heavy") cells (more than SHPL_HEAVY_LEN entries; none at KITTI / MV3D shapes, the Zipf stress case has
 * a 178 000-entry cell): one thread-block CLUSTER of 8 CTAs per listed cell, in two kernels.
 *   up to SHPL_EXACT_LEN entries: the sequential sum, bit-identical to what the main kernels produce -- the 64
 *     warps of the cluster gather a round of entries and park the rounded products w*x, in entry order, in CTA 0's
 *     shared memory (remote stores through distributed shared memory); adder warps of CTA 0 (one per 32 channel
 *     vectors) then add them in entry order;
 *   longer cells: the entries are cut into 64 contiguous pieces (8 CTAs x 8 warps); every warp sums its piece in
 *     stored order, the 8 warp sums of a CTA are added in order in shared memory, and CTA 0 adds the 8 CTA sums in
 *     order through distributed shared memory: a fixed tree, deterministic, within fp32 rounding of the sequential
 *     sum (not bit-identical to it).
 * Overwrites what the main kernel left for those cells:
 *   out[c*out_stride + 0:C] = (addend ? addend[c*addend_stride + 0:C] : 0) + sum_k val[k] * gather_in[idx[k]*gather_stride + 0:C]
 * for every c in list[0:*count_dev].  All strides in floats; the channel offsets are folded into the
 * pointers (e.g. out = fused + C_d, out_stride = C_d + C_s for the forward). */
int shpl_pool_heavy(const float* gather_in, int32_t gather_stride, int32_t C,
                    const int32_t* ptr, const int32_t* idx, const float* val,
                    const int32_t* list, const int32_t* count_dev, int32_t list_cap,
                    const float* addend, int32_t addend_stride,
                    float* out, int32_t out_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Feeder: BEV slicing / voxelisation.  Replaces
 *   BevSlices.generate_bev(..., output_indices=True)   avod/avod/core/bev_generators/bev_slices.py:33-156
 *   VoxelGrid2D.voxelize_2d                             avod/wavedata/wavedata/tools/core/voxel_grid_2d.py:43-162
 *   KittiUtils.create_slice_filter                      avod/avod/datasets/kitti/kitti_utils.py:79-107
 *   obj_utils.get_point_filter                          avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:444-491
 *   BevGenerator._create_density_map                    avod/avod/core/bev_generators/bev_generator.py:23-41
 * whose outputs (voxel_indices, unique_pts) are the inputs of shpl_build_avod. */
#define SHPL_BEV_MAX_SLICES 8
#define SHPL_BEV_COUNTS 32
/* bits of counts[1]: conditions under which the reference raises */
#define SHPL_BEV_ERR_FIRST_SLICE_EMPTY 1 /* first slice has <= 1 point: NameError (bev_slices.py:79,93)    */
#define SHPL_BEV_ERR_EXTENTS 2           /* a voxel coordinate outside the extents: ValueError (voxel_grid_2d.py:133-138) */
#define SHPL_BEV_ERR_CAPACITY 4          /* more (slice, cell) pairs than `capacity` (not a reference condition) */
#define SHPL_BEV_ERR_NO_POINTS 8         /* no point in the density band: voxelize_2d of nothing raises     */

/* Grid the extents imply: nx = num_divisions[0] (800 at KITTI), nz = num_divisions[2] (700)
 * (voxel_grid_2d.py:126-149).  extents_host f64 [6] = x_lo, x_hi, y_lo, y_hi, z_lo, z_hi. */
int shpl_bev_grid_dims(const double* extents_host, double voxel_size, int32_t* nx, int32_t* nz);

/* Scratch shpl_bev_slices needs (0 on invalid arguments). */
size_t shpl_bev_workspace_bytes(const double* extents_host, double voxel_size, int32_t num_slices);

/* generate_bev(output_indices=True).
 *   points f64: coordinate c of point i at points[c*coord_stride + i*point_stride] -- the reference's
 *   point_cloud [3,P] is (coord_stride=P, point_stride=1), an [P,3] array (1, 3);
 *   P_dev (device int32*, may be NULL): when given, only the first min(P, *P_dev) points exist -- lets the ingest's
 *   device-side count (shpl_lidar_to_cam counts[0]) flow in without a host read;
 *   ground_plane_host f64 [4]; extents_host f64 [6]; slices of (height_hi-height_lo)/num_slices above the plane;
 *   log_norm = NORM_VALUES[source] (log 16 for lidar, bev_slices.py:12-14);
 *   density_lut (DEVICE f64 [lut_len], may be NULL): the density value of a cell holding n points for
 *   n < lut_len, 1.0 beyond -- lets the host tabulate min(1, log(n+1)/log_norm) with the reference's own
 *   libm so the map is bit-identical; with NULL the kernel evaluates the formula (CUDA log, <= 1 ulp).
 * Outputs, in the reference's order (slice, then x, then z ascending):
 *   voxel_indices_out i64 [capacity,2] = (x, nz - z)  (the reference's flipped row index, :106-108),
 *   unique_pts_out f64 [capacity,3] = first point of each cell in the lexsort (x, z, y) order (:97-98),
 *   bev_maps_out f64 [num_slices+1, nz, nx] (may be NULL): the height maps (:116-118) then the density map,
 *   counts i32 [SHPL_BEV_COUNTS] (device): [0] = number of pairs N, [1] = SHPL_BEV_ERR_* bits,
 *   [2] = points in the density band, [8+s] = output offset where slice s starts, [16+s] = points in slice s.
 * A slice with <= 1 point repeats the previous slice's cells and compounds its heights, like the
 * reference does (:79-112). */
int shpl_bev_slices(const double* points, int64_t coord_stride, int64_t point_stride, int64_t P,
                    const int32_t* P_dev, const double* ground_plane_host, const double* extents_host, double voxel_size,
                    double height_lo, double height_hi, int32_t num_slices, double log_norm,
                    const double* density_lut, int32_t lut_len,
                    int64_t* voxel_indices_out, double* unique_pts_out, int64_t capacity,
                    double* bev_maps_out, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MV3D voxel feeder: the source of the non-homogeneous weights.  Replaces
 *   point_cloud_2_top_sparse(points, ..., points_in_cam=True, img_index2=...)
 *                                        MV3D_TF_release/lib/utils/construct_voxel.py:37-162
 * (call site MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:93).
 *   points f64 [n,4] camera frame (x, y, z, reflectance); img_index2 i64 [2,n] rounded pixel (u; v) of every point;
 *   ranges_host f64 [6] = side_lo, side_hi, fwd_lo, fwd_hi, height_lo, height_hi (construct_voxel.py:11-13);
 *   res / zres: horizontal / vertical voxel size; max_points = cfg.VOXEL_POINT_COUNT (:17).
 * Outputs (n_in = points strictly inside the ranges, m = points kept after the per-voxel cap, V = voxels):
 *   voxel_full_size_host i32 [3] (HOST, may be NULL) = (z_max+1, x_max+1, y_max+1)  (:84);
 *   img_index_out i64 [3,capacity] rows (u, v, 0), first m columns valid, row stride = capacity  (:156-158);
 *   bv_index_out i64 [capacity,2] = (forward cell, side cell)  (:159);
 *   m_val_out f64 [capacity] = 1 / (points kept in the pair's voxel)  (:160);
 *   feature_buffer f64 [voxel_capacity, max_points, 7], coordinate_buffer i64 [voxel_capacity,4] = (0, z, x, y),
 *   number_buffer i64 [voxel_capacity]: the VoxelNet buffers of voxel_dict (:126-148), voxels in np.unique's
 *   lexicographic (x, y, z) order; any of the three may be NULL;
 *   counts i32 [8] (device): [0] = n_in, [1] = m, [2] = V, [3] = error bits (1: cell outside the grid,
 *   2: m > capacity, 4: V > voxel_capacity). */
size_t shpl_mv3d_workspace_bytes(int64_t n_max);
int shpl_mv3d_voxelize(const double* points, const int64_t* img_index2, int64_t n, double res, double zres,
                       const double* ranges_host, int32_t max_points, int32_t* voxel_full_size_host,
                       int64_t* img_index_out, int64_t* bv_index_out, double* m_val_out, int64_t capacity,
                       double* feature_buffer, int64_t* coordinate_buffer, int64_t* number_buffer,
                       int64_t voxel_capacity, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point-cloud ingest.  Replaces the arithmetic of
 *   obj_utils.get_lidar_point_cloud   avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:220-268
 *   calib_utils.lidar_to_cam_frame    avod/wavedata/wavedata/tools/core/calib_utils.py:371-410
 *   calib_utils.project_to_image      avod/wavedata/wavedata/tools/core/calib_utils.py:281-297
 * (reading the calibration .txt and the velodyne .bin stays on the host).
 *   velo_xyzi f32 [N,4] (DEVICE): the raw scan as the KITTI .bin holds it (x, y, z, intensity);
 *   rectified_host f64 [12]: rows 0..2 of R0_rect(4x4) . Tr_velo_to_cam(4x4)  (calib_utils.py:406);
 *   p2_host f64 [12]: the camera matrix; im_w, im_h: the image size to filter with, or 0, 0 for every point
 *   (im_size = None); use_min_intensity / min_intensity: the reference's optional intensity filter (:264-268).
 * Output cam_out f64 [3, capacity], coordinate-major (row stride = capacity; first counts[0] columns valid): the
 * camera-frame points with z > 0 that project strictly inside the image, input order kept -- what
 * shpl_bev_slices takes with coord_stride = capacity, point_stride = 1.
 * counts i32 [4] (device): [0] = points kept, [1] = points with z > 0, [2] = 1 if capacity was exceeded. */
size_t shpl_lidar_workspace_bytes(int64_t n_max);
int shpl_lidar_to_cam(const float* velo_xyzi, int64_t N, const double* rectified_host, const double* p2_host,
                      int32_t im_w, int32_t im_h, int32_t use_min_intensity, float min_intensity,
                      double* cam_out, int64_t capacity, int32_t* counts, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Augmentation hooks that touch the arrays the correspondence builder reads.
 * Element-wise, in place, asynchronous; n_dev (device int32*, may be NULL) = device-side element count as above. */

/* kitti_aug.flip_point_cloud (avod/avod/datasets/kitti/kitti_aug.py:24-29): x -> -x.
 *   points: address of the first point's x; point_stride: doubles between consecutive points' x
 *   (1 for the reference's [3,N] point cloud, 3 for an [N,3] array). */
int shpl_flip_point_cloud(double* points, int64_t point_stride, int64_t n, const int32_t* n_dev, void* stream);

/* MV3D sample preparation (MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py):
 *   img_index2 = np.round(projectToImage(lidar_pc[:, 0:3].T, P)).astype(int)   (:172-174, :183-185;
 *   projectToImage = lib/utils/transform.py:429-452) into img_index2_out i64 [2,n] (NULL: not wanted), computed
 *   from the points BEFORE the augmentation, then, when augment != 0, augment_voxel's point transforms (:176-181):
 *   x += sx, z += sz; (x, y, z) *= expansion_ratio; (x, z) = rot_mat . (x, z) with rot_host f64 [4] =
 *   [[cos a, sin a], [-sin a, cos a]] row-major as the host evaluated it.
 *   lidar_pc f64 [n,4] (x, y, z, reflectance), camera frame, 16-byte aligned, transformed in place.
 * The outputs are the (points, img_index2) inputs of shpl_mv3d_voxelize. */
int shpl_mv3d_project_augment(double* lidar_pc, int64_t n, const int32_t* n_dev, const double* P_host,
                              int32_t augment, double sx, double sz, double expansion_ratio,
                              const double* rot_host, int64_t* img_index2_out, void* stream);

/* augment_fv's index update (minibatch_mv3d_img.py:205-206):
 *   img_index[0,:] = (img_index[0,:]*expansion_ratio + sx).astype(int), row 1 with sy (truncation toward zero).
 *   img_index i64 [3, ld] row-major (ld >= n: row stride). */
int shpl_augment_fv_index(int64_t* img_index, int64_t ld, int64_t n, const int32_t* n_dev,
                          double expansion_ratio, double sx, double sy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Post-fusion 3x3 convolution fused with the pooling: the fused (concat) map is never
 * written.  Replaces, for the rpn_sparse_pooling_conv_after_fusion switch (model.proto:89, default true),
 *   sparse_pool_layer(...) -> slim.conv2d(bev_fused, C, [3, 3], ...)   avod/avod/core/models/rpn_model.py:335-346
 *   the RetinaNet form                                                  avod/avod/core/models/retinanet_model.py:337-348
 * with  conv(concat(dst, pooled)) = conv(dst; W[:, :, :C_d, :]) + conv(pooled; W[:, :, C_d:, :]):
 * the first term is a dense implicit GEMM on the tcgen05 tensor cores (3xTF32 split: fp32-level accuracy, TMA halo
 * loads with the SAME zero padding, TMEM accumulators), the second a sparse update of the 3x3 neighbourhoods of the
 * cells that receive pooled features.
 *   dst [frames, H, W, C_d], src [n_src, C_s] (the map gathered from), CSR by destination cell (ptr, key, idx, val,
 *   nnz_max as in shpl_pool_forward; plan rows of frame f are offset by f*H*W; key is required),
 *   weight: the slim.conv2d variable, HWIO [3, 3, C_d + C_s, C_out], stride 1, padding SAME;
 *   scale / shift [C_out] (either may be NULL = 1 / 0): bias or folded inference batch norm; relu != 0: ReLU;
 *   out [frames, H, W, C_out] = act(scale * conv(concat(dst, pooled)) + shift).
 * Built for C_d = C_out = 32 and C_s in {0, 32, 64} (the KITTI pre-RPN layer: 32 + 32 -> 32); other shapes return
 * SHPL_ERR_UNSUPPORTED.  workspace: shpl_conv3x3_workspace_bytes(frames, H, W, nnz_max) bytes, 256-byte aligned
 * (weights in the tensor-core operand layout, two cell bitmaps, and 1152 bytes per entry for the sparse half).
 * Accuracy: |error| <= 1e-5 * sum |terms| per output (3xTF32 products, fp32 accumulation); not bit-reproducible
 * against a sequential fp32 loop (neither is cuDNN / TF).
 * The call enqueues three kernels on `stream`; the second and third are programmatic dependent launches of the one before
 * (they start early and wait inside the kernel for what they read).  To the caller the call is stream-ordered like any
 * other, and it can be captured in a CUDA graph. */
size_t shpl_conv3x3_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t nnz_max);
int shpl_pool_conv3x3_forward(const float* dst, const float* src,
                              const int32_t* ptr, const int32_t* key, const int32_t* idx, const float* val,
                              int32_t nnz_max, int32_t frames, int32_t H, int32_t W, int32_t C_d, int32_t n_src, int32_t C_s,
                              const float* weight, int32_t C_out, const float* scale, const float* shift, int32_t relu,
                              float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of shpl_pool_conv3x3_forward's linear part, out = conv3x3(concat(dst, pooled), weight) (what TF autodiff derives
 * for slim.conv2d + sparse_pool_layer; the gradient of scale / shift / ReLU is the caller's: pass the gradient with
 * respect to the conv output).  Deterministic: every sum runs in a fixed order.
 *   g_dst    [frames, H, W, C_d]  = conv3x3(g_out; W[:, :, :C_d, :] flipped and transposed): the tcgen05 kernel again;
 *   g_src    [n_src, C_s]         = the pooled channels' input gradient at the cells that receive pooled features, pushed
 *                                   through the transposed CSR like shpl_pool_backward does;
 *   g_weight [3, 3, C_d+C_s, C_out] (HWIO) = sum over pixels of x[p + tap] (x) g_out[p].
 * Any of the three outputs may be NULL.  (ptr, key, idx, val): CSR by destination cell; (ptrT, keyT, idxT, valT): CSR by
 * source pixel, both of the same plan.  Built for C_d = C_out = 32 and C_s in {0, 32}.
 * workspace: shpl_conv3x3_backward_workspace_bytes(nnz_max) bytes, 256-byte aligned.
 * Accuracy: g_dst as the forward (3xTF32); g_src and g_weight fp32 FMA sums, |error| <= 1e-5 * sum |terms|. */
size_t shpl_conv3x3_backward_workspace_bytes(int32_t nnz_max);
int shpl_pool_conv3x3_backward(const float* g_out, const float* dst, const float* src,
                               const int32_t* ptr, const int32_t* key, const int32_t* idx, const float* val,
                               const int32_t* ptrT, const int32_t* keyT, const int32_t* idxT, const float* valT,
                               int32_t nnz_max, int32_t frames, int32_t H, int32_t W, int32_t C_d, int32_t n_src, int32_t C_s,
                               const float* weight, int32_t C_out, float* g_dst, float* g_src, float* g_weight,
                               void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHPL_H_ */
