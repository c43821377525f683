"""Index oracle: CPU restatement of the SHPL correspondence builder (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, step by step:

  * projectToImage            /root/reference/avod/avod/utils/transform.py:3-26
                              (twin MV3D_TF_release/lib/utils/transform.py:429-452)
  * clip3DwithinImage         /root/reference/avod/avod/utils/transform.py:28-40
  * gen_sparse_pooling_input_avod
                              /root/reference/avod/avod/utils/sparse_pool_utils.py:6-20
  * produce_sparse_pooling_input
                              /root/reference/avod/avod/utils/sparse_pool_utils.py:22-58
                              (twin MV3D_TF_release/lib/utils/sparse_pool_utils.py:22-54)

It is written from the behavioural description in SURVEY.md Appendix A.1, one
numbered step per block, not from the reference's statements.  Pinned against
the reference's own output by tests/test_oracle_golden.py (fixtures made by
oracle/gen_goldens.py) and against KAT-1 / KAT-2.

The canonical CSR / transposed-CSR forms (A.1 step 9) are NOT in the reference
(TF keeps COO); they are defined here as the stable sort of the reference's COO
by destination row / by source pixel, which is the order TF-CPU accumulates in.
"""
import numpy as np


# ---------------------------------------------------------------- projection
def project_to_image(pts_3d, P):
    """A.1 step 1.  pts_3d f64 [3,n], P f64 [3,4] -> (u,v) f64 [2,n].

    transform.py:17-24: homogeneous product through BLAS (np.dot), rows 0,1
    divided by row 2.  np.dot is kept on purpose: on this image OpenBLAS rounds
    each row as  t=P0*x; t=fma(P1,y,t); t=fma(P2,z,t); t=fma(P3,1,t)  (probed,
    DESIGN.md), which is the chain the C oracle and the CUDA builder spell out.
    """
    pts_3d = np.asarray(pts_3d, dtype=np.float64)
    n = pts_3d.shape[1]
    hom = np.empty((4, n), dtype=np.float64)
    hom[:3] = pts_3d
    hom[3] = 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        uvw = np.dot(np.asarray(P, dtype=np.float64), hom)
        return np.stack((uvw[0] / uvw[2], uvw[1] / uvw[2]))


def clip_within_image(pts_3d, P, image_size):
    """A.1 step 2 (transform.py:36-38).  image_size = [W, H].

    keep <=> 0 <= u < W-1  and  0 <= v < H-1  (NaN / inf compare false).
    """
    uv = project_to_image(pts_3d, P)
    W, H = image_size[0], image_size[1]
    with np.errstate(invalid="ignore"):
        return (uv[0] < W - 1) & (uv[0] >= 0) & (uv[1] >= 0) & (uv[1] < H - 1)


# ------------------------------------------------------------------- builder
def gen_sparse_pooling_input_avod(points, voxel_indices, P, im_size, bv_size):
    """sparse_pool_utils.py:6-20.  `P` is stereo_calib.p2 (:12).

    points f64 [N,3] (camera frame), voxel_indices int [N,>=2] (x, zflip).
    Returns the reference's dict: bv_index int [n,2], img_index f64 [3,n] =
    (round_half_even(u), round_half_even(v), 0), bv_size [H_b,W_b], img_size [W,H].
    """
    points = np.asarray(points, dtype=np.float64)
    voxel_indices = np.asarray(voxel_indices)
    keep = clip_within_image(points.T, P, im_size)                     # :14
    kept_pts = points[keep]                                            # :16
    uv = project_to_image(kept_pts.T, P)                               # :17
    uv_int = np.rint(uv).astype(int)                                   # :18 (np.round == rint, ties to even)
    img_index = np.zeros((3, uv_int.shape[1]), dtype=np.float64)       # :19 (vstack with a zero row -> f64)
    img_index[:2] = uv_int
    return {
        "bv_index": voxel_indices[keep][:, :2].copy(),                 # :11, :15
        "img_index": img_index,
        "bv_size": np.array([bv_size[0], bv_size[1]]),
        "img_size": np.array(im_size),
    }


def produce_sparse_pooling_input(input_dict, M_val=None, stride=(1, 1)):
    """sparse_pool_utils.py:22-58.  stride[0] scales the IMAGE, stride[1] the BEV
    (quirk A.4-1).  Mutates input_dict['img_index'] in place like the reference
    (:30, quirk A.4-6)."""
    bv_index = input_dict["bv_index"]
    img_index = input_dict["img_index"]
    if img_index.shape[0] != 3:                                        # :29 (assert)
        raise AssertionError("wrong img_index shape, should be 3xN instead " + str(img_index.shape))
    s_img, s_bv = stride[0], stride[1]

    # step 4: image side -- floor-divide, then clamp the upper bound only (:30-34)
    img_index[0:2, :] = np.floor(img_index[0:2, :] / s_img)
    im_size = np.floor(np.asarray(input_dict["img_size"]) / s_img)     # [W', H']
    for axis in (0, 1):
        over = img_index[axis, :] >= im_size[axis]
        img_index[axis, over] = im_size[axis] - 1
    # step 5: gather triples [0, v', u'] (:36)
    flip = np.floor(img_index.T[:, ::-1]).astype(int)

    # step 6-7: BEV side (:38-44)
    bv_size = np.floor(np.asarray(input_dict["bv_size"]) / s_bv)       # [H_b', W_b']
    cell = np.floor(np.asarray(bv_index) / s_bv)                       # (x', z')
    row = (cell[:, 1] * bv_size[1] + cell[:, 0]).astype(int)
    n_rows = int(bv_size[0] * bv_size[1])
    inside = row < n_rows

    # step 8: renumber survivors (:46-57)
    flip = flip[inside]
    row = row[inside]
    nnz = row.shape[0]
    Mij = np.stack((row, np.arange(nnz, dtype=row.dtype)), axis=1) if nnz else np.zeros((0, 2), dtype=int)
    M_size = np.array([bv_size[0] * bv_size[1], nnz]).astype(int)
    if M_val is None:
        M_val = np.ones(nnz)
    return {
        "Mij_pool": Mij,
        "M_val": M_val,
        "M_size": M_size,
        "img_index_flip_pool": flip,
        "bev_index_flip_pool": np.zeros((0, 3)),
    }


# -------------------------------------------------- canonical CSR (A.1 step 9)
def coo_to_csr(keys, n_keys):
    """Stable sort of entry ids by key -> (ptr int32 [n_keys+1], order int64 [m]).

    Entries whose key is outside [0, n_keys) are left out (they are the indices
    TF-CPU rejects with InvalidArgumentError; the CUDA kernels never read them).
    Returns also the number left out.
    """
    keys = np.asarray(keys, dtype=np.int64)
    ok = (keys >= 0) & (keys < n_keys)
    ids = np.nonzero(ok)[0]
    order = ids[np.argsort(keys[ids], kind="stable")]
    counts = np.bincount(keys[ids], minlength=n_keys) if n_keys > 0 else np.zeros(0, dtype=np.int64)
    ptr = np.zeros(n_keys + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr.astype(np.int32), order, int(keys.shape[0] - ids.shape[0])


def build_plan(Mij, M_val, img_index_flip, n_rows, src_h, src_w, batch_stride_rows=None):
    """Canonical CSR (by destination row) and CSR^T (by source pixel) of the
    reference's COO, as the CUDA builder must emit them.

    Mij int [m,2] = (row, col); M_val [m]; img_index_flip int [ncol,3] = (b, v, u)
    with b always 0 in the reference (sparse_pool_utils.py:98).  Entry e reads
    source pixel  pix_e = v[col_e]*src_w + u[col_e];  an entry is *valid* when
    0<=row<n_rows, 0<=col<ncol, 0<=v<src_h, 0<=u<src_w.  Invalid entries are left
    out of both CSRs and counted in 'n_oob'.
    """
    Mij = np.asarray(Mij, dtype=np.int64).reshape(-1, 2)
    flip = np.asarray(img_index_flip, dtype=np.int64).reshape(-1, 3)
    val = np.asarray(M_val, dtype=np.float32).reshape(-1)
    m = Mij.shape[0]
    row, col = Mij[:, 0], Mij[:, 1]
    col_ok = (col >= 0) & (col < flip.shape[0])
    safe_col = np.where(col_ok, col, 0)
    if flip.shape[0] == 0:
        v = np.zeros(m, dtype=np.int64)
        u = np.zeros(m, dtype=np.int64)
        col_ok = np.zeros(m, dtype=bool)
    else:
        v = flip[safe_col, 1]
        u = flip[safe_col, 2]
    b = flip[safe_col, 0] if flip.shape[0] else np.zeros(m, dtype=np.int64)
    valid = col_ok & (row >= 0) & (row < n_rows) & (v >= 0) & (v < src_h) & (u >= 0) & (u < src_w) & (b == 0)
    pix = v * src_w + u
    row_key = np.where(valid, row, -1)
    pix_key = np.where(valid, pix, -1)
    row_ptr, order_r, _ = coo_to_csr(row_key, n_rows)
    pix_ptr, order_p, _ = coo_to_csr(pix_key, src_h * src_w)
    return {
        "row_ptr": row_ptr,
        "csr_row": row[order_r].astype(np.int32),      # destination row of each entry (the sort key)
        "csr_src": pix[order_r].astype(np.int32),      # source pixel of each entry, row-major order
        "csr_val": val[order_r].astype(np.float32),
        "csr_ent": order_r.astype(np.int32),           # COO entry id (k) of each CSR slot
        "pix_ptr": pix_ptr,
        "csrT_pix": pix[order_p].astype(np.int32),     # source pixel of each entry (the sort key)
        "csrT_dst": row[order_p].astype(np.int32),     # destination row of each entry, pixel-major order
        "csrT_val": val[order_p].astype(np.float32),
        "csrT_ent": order_p.astype(np.int32),
        "n_oob": int(m - valid.sum()),
    }


# ------------------------------------------------------ MV3D weights (a7)
def mv3d_voxel_weights(points_fsh, res, zres, side_range, fwd_range, height_range, max_points):
    """The part of point_cloud_2_top_sparse that feeds SHPL
    (MV3D_TF_release/lib/utils/construct_voxel.py:89-100, :116-122, :133-140, :159-160).

    points_fsh f64 [n,>=3] = (forward, side, height) after the :59 column swap.
    Returns (filter mask over the input [n], kept ids among the filtered points,
    bv_index int [m,2] = (fwd_cell, side_cell), M_val f64 [m] = 1/count of the
    point's 3-D voxel after the per-voxel cap).
    """
    p = np.asarray(points_fsh, dtype=np.float64)
    f, s, h = p[:, 0], p[:, 1], p[:, 2]
    inrange = ((f > fwd_range[0]) & (f < fwd_range[1]) & (s > side_range[0]) & (s < side_range[1])
               & (h > height_range[0]) & (h < height_range[1]))                 # :89-96 (strict both sides)
    q = p[inrange]
    cell = np.empty((q.shape[0], 3), dtype=np.int64)
    cell[:, 0] = ((q[:, 1] - side_range[0]) / res).astype(np.int32)             # :116 truncation
    cell[:, 1] = ((q[:, 0] - fwd_range[0]) / res).astype(np.int32)              # :117
    cell[:, 2] = ((q[:, 2] - height_range[0]) / zres).astype(np.int32)          # :118
    _, inv = np.unique(cell, axis=0, return_inverse=True)                       # :122
    inv = np.asarray(inv).reshape(-1)
    # :133-140 -- first `max_points` points of each voxel, in input order, survive
    order = np.argsort(inv, kind="stable")
    sorted_inv = inv[order]
    first = np.r_[True, sorted_inv[1:] != sorted_inv[:-1]] if inv.size else np.zeros(0, dtype=bool)
    start = np.maximum.accumulate(np.where(first, np.arange(inv.size), 0)) if inv.size else np.zeros(0, dtype=np.int64)
    rank_sorted = np.arange(inv.size) - start
    rank = np.empty(inv.size, dtype=np.int64)
    rank[order] = rank_sorted
    kept = np.nonzero(rank < max_points)[0]
    counts = np.minimum(np.bincount(inv, minlength=(inv.max() + 1 if inv.size else 0)), max_points)
    bv_index = cell[kept][:, [1, 0]]                                            # :159
    m_val = 1.0 / counts[inv[kept]]                                             # :160
    return inrange, kept, bv_index, m_val
