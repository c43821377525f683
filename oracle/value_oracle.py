"""Value oracle: CPU restatement (numpy, fp32) of the SHPL device half.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows

  * sparse_pool_layer       /root/reference/avod/avod/utils/sparse_pool_utils.py:61-92
  * _sparse_pool_op         .../sparse_pool_utils.py:96-103
  * _sparse_pool_trans_op   .../sparse_pool_utils.py:105-117
  * concat_bn_op            .../sparse_pool_utils.py:120-124 (BN is delegated, see below)
  * the gradients TF autodiff derives for them

The arithmetic of those functions lives in TensorFlow 1.8.0 (third-party, pinned
by avod/README.md:31, not vendored and not installable here).  The op semantics
restated here are TF's published ones:

  tf.gather_nd(x[1,H,W,C], idx[n,3])          -> x[idx[:,0], idx[:,1], idx[:,2], :]
  tf.sparse_tensor_dense_matmul(A, B)         -> out zero-initialised fp32;
                                                 out[row_e] += val_e * B[col_e] in COO order e=0..m-1
  tf.sparse_transpose + matmul                -> S[col_e] += val_e * X[row_e]
  tf.scatter_nd(idx, upd, shape)              -> zeros; out[idx_k] += upd[k] in k order (duplicates summed)
  tf.concat(axis=3)                           -> channel concatenation
  gradients: SpMM wrt dense = A^T g ; gather_nd -> scatter_nd ; scatter_nd -> gather_nd ; concat -> slices

PARITY UNPINNED at this boundary: the reference holds no golden vector for it
(SURVEY.md 8c); tests cross-check this file against torch CPU autograd instead.

Products are rounded to fp32 before they are added (numpy has no fused
multiply-add), which is the order of operations the CUDA kernels spell out with
__fmul_rn / __fadd_rn.
"""
import numpy as np

F32 = np.float32


def _rank_within_group(keys):
    """rank[e] = number of earlier entries e' < e with keys[e'] == keys[e]."""
    keys = np.asarray(keys, dtype=np.int64)
    m = keys.shape[0]
    if m == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    first = np.r_[True, sk[1:] != sk[:-1]]
    start = np.maximum.accumulate(np.where(first, np.arange(m), 0))
    rank = np.empty(m, dtype=np.int64)
    rank[order] = np.arange(m) - start
    return rank


def segment_accumulate(out, keys, contrib):
    """out[keys[e]] += contrib[e] for e = 0..m-1 IN THAT ORDER, fp32.

    Vectorised over 'rank within key' so that every destination still receives
    its contributions strictly in ascending e (what a sequential loop does)."""
    keys = np.asarray(keys, dtype=np.int64)
    rank = _rank_within_group(keys)
    for j in range(int(rank.max()) + 1 if rank.size else 0):
        sel = np.nonzero(rank == j)[0]
        out[keys[sel]] = out[keys[sel]] + contrib[sel]
    return out


def gather_nd(x, idx):
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size and ((idx < 0).any() or (idx >= np.array(x.shape[:3])).any()):
        raise IndexError("gather_nd: index out of range (TF-CPU: InvalidArgumentError)")
    return x[idx[:, 0], idx[:, 1], idx[:, 2], :]


def spmm(indices, values, dense_shape, B):
    """tf.sparse_tensor_dense_matmul(SparseTensor(indices, values, dense_shape), B)."""
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    values = np.asarray(values, dtype=F32)
    out = np.zeros((int(dense_shape[0]), B.shape[1]), dtype=F32)
    if indices.shape[0] == 0:
        return out
    if (indices[:, 0] < 0).any() or (indices[:, 0] >= dense_shape[0]).any() or \
       (indices[:, 1] < 0).any() or (indices[:, 1] >= B.shape[0]).any():
        raise IndexError("spmm: index out of range (TF-CPU: InvalidArgumentError)")
    contrib = (values[:, None] * B[indices[:, 1]].astype(F32)).astype(F32)
    return segment_accumulate(out, indices[:, 0], contrib)


def spmm_transposed(indices, values, dense_shape, X):
    """tf.sparse_tensor_dense_matmul(tf.sparse_transpose(A), X) -> [ncols, C]."""
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    values = np.asarray(values, dtype=F32)
    out = np.zeros((int(dense_shape[1]), X.shape[1]), dtype=F32)
    if indices.shape[0] == 0:
        return out
    contrib = (values[:, None] * X[indices[:, 0]].astype(F32)).astype(F32)
    return segment_accumulate(out, indices[:, 1], contrib)


def scatter_nd(idx, updates, shape):
    idx = np.asarray(idx, dtype=np.int64)
    out = np.zeros((shape[0] * shape[1] * shape[2], shape[3]), dtype=F32)
    if idx.shape[0] == 0:
        return out.reshape(shape)
    if (idx < 0).any() or (idx >= np.array(shape[:3])).any():
        raise IndexError("scatter_nd: index out of range (TF-CPU: InvalidArgumentError)")
    lin = (idx[:, 0] * shape[1] + idx[:, 1]) * shape[2] + idx[:, 2]
    return segment_accumulate(out, lin, updates.astype(F32)).reshape(shape)


# ------------------------------------------------------------- the two ops
def sparse_pool_op(M, x, source_index, pooled_size):
    """_sparse_pool_op (:96-103).  M = (indices, values, dense_shape)."""
    G = gather_nd(x, source_index)                                    # :101
    Y = spmm(M[0], M[1], M[2], G)                                     # :102
    return Y.reshape(pooled_size)                                     # :103


def sparse_pool_trans_op(M, x, source_index, pooled_size):
    """_sparse_pool_trans_op (:105-117)."""
    S = spmm_transposed(M[0], M[1], M[2], x.reshape(-1, x.shape[3]))  # :110-111
    return scatter_nd(source_index, S, pooled_size)                   # :116


def sparse_pool_layer(inputs, feature_depths, M, img_index_flip=None, bv_index=None):
    """sparse_pool_layer (:61-92), use_bn=False branch."""
    input_bv, input_img = inputs
    if img_index_flip is not None:
        size = [1, input_bv.shape[1], input_bv.shape[2], feature_depths[0]]      # :66-67
        bv_fused = np.concatenate([input_bv, sparse_pool_op(M, input_img, img_index_flip, size)], axis=3)
    else:
        bv_fused = input_bv
    if bv_index is not None:                                                     # :79 content ignored
        size = [1, input_img.shape[1], input_img.shape[2], feature_depths[1]]    # :81-82
        img_fused = np.concatenate([input_img, sparse_pool_trans_op(M, input_bv, img_index_flip, size)], axis=3)
    else:
        img_fused = input_img
    return bv_fused, img_fused


# ----------------------------------------------------------------- gradients
def sparse_pool_layer_grad(inputs, feature_depths, M, img_index_flip, bv_index, g_bv_fused, g_img_fused):
    """Gradients of sparse_pool_layer wrt (input_bv, input_img) as TF autodiff
    composes them.  When an input feeds two consumers (dual
    direction) TF adds the two partial gradients (AddN): slice first, pooled
    path second."""
    input_bv, input_img = inputs
    Cb, Ci = input_bv.shape[3], input_img.shape[3]
    ind, val, shp = M
    ind = np.asarray(ind, dtype=np.int64).reshape(-1, 2)
    val = np.asarray(val, dtype=F32)
    g_bv = None
    g_img = None
    if img_index_flip is not None:
        g_bv = g_bv_fused[..., :Cb].copy()                                        # concat grad = slice
        gY = g_bv_fused[..., Cb:].reshape(-1, Ci)
        gG = spmm_transposed(ind, val, shp, gY)                                   # A^T g
        g_img_pool = scatter_nd(img_index_flip, gG, input_img.shape)              # gather_nd grad
    else:
        g_bv = g_bv_fused.copy()
        g_img_pool = None
    if bv_index is not None:
        g_img = g_img_fused[..., :Ci].copy()
        gS = gather_nd(g_img_fused[..., Ci:], img_index_flip)                     # scatter_nd grad
        g_bv_pool = spmm(ind, val, shp, gS).reshape(input_bv.shape)               # (A^T)^T g
        g_bv = (g_bv + g_bv_pool).astype(F32)
    else:
        g_img = g_img_fused.copy()
    if g_img_pool is not None:
        g_img = (g_img + g_img_pool).astype(F32)          # AddN of the direct path and the pooled path
    return g_bv, g_img


# ------------------------------------------------------------- the VFE scatter
def voxel_scatter(coordinate, voxelwise, shape, strict=True):
    """tf.scatter_nd(coordinate, voxelwise, [B, 10, H, W, C])
    (/root/reference/MV3D_TF_release/lib/networks/group_pointcloud.py:84-85): zeros, then
    out[coordinate[k]] += voxelwise[k] in k order.  strict: TF-CPU's InvalidArgumentError for a coordinate
    outside the grid; otherwise such rows are dropped (TF-GPU)."""
    coordinate = np.asarray(coordinate, dtype=np.int64).reshape(-1, 4)
    grid = np.array(shape[:4], dtype=np.int64)
    C = int(shape[4])
    out = np.zeros((int(np.prod(grid)), C), dtype=F32)
    inside = ((coordinate >= 0) & (coordinate < grid)).all(axis=1)
    if strict and not inside.all():
        raise IndexError("scatter_nd: index out of range (TF-CPU: InvalidArgumentError)")
    c = coordinate[inside]
    lin = ((c[:, 0] * grid[1] + c[:, 1]) * grid[2] + c[:, 2]) * grid[3] + c[:, 3]
    if lin.shape[0]:
        segment_accumulate(out, lin, np.asarray(voxelwise, dtype=F32)[inside])
    return out.reshape(tuple(int(g) for g in grid) + (C,))


def voxel_scatter_grad(coordinate, g_out):
    """Gradient of the scatter wrt voxelwise: gather_nd(g_out, coordinate); rows dropped by the forward get zeros."""
    coordinate = np.asarray(coordinate, dtype=np.int64).reshape(-1, 4)
    grid = np.array(g_out.shape[:4], dtype=np.int64)
    inside = ((coordinate >= 0) & (coordinate < grid)).all(axis=1)
    g = np.zeros((coordinate.shape[0], g_out.shape[4]), dtype=F32)
    c = coordinate[inside]
    g[inside] = g_out[c[:, 0], c[:, 1], c[:, 2], c[:, 3]]
    return g


# ---------------------------------------------------------------------------------------------
# Post-fusion 3x3 convolution: slim.conv2d(fused, C, [3, 3]) right after sparse_pool_layer,
# /root/reference/avod/avod/core/models/rpn_model.py:338-354, retinanet_model.py:344-348.  slim.conv2d defaults:
# stride 1, padding SAME, weights HWIO [3, 3, C_in, C_out]; then biases (or the normalizer) and ReLU.  The
# arithmetic again lives in TensorFlow (cuDNN / Eigen): PARITY UNPINNED, cross-checked against torch conv2d in the
# tests.  Accumulated in float64 here, so that the comparison bound (1e-5 * sum |terms|) is about the CUDA kernel
# alone.
def conv3x3_same(x, w):
    """x [B, H, W, Cin], w [3, 3, Cin, Cout] -> (y [B, H, W, Cout] float64, sum of |terms| per output, float64)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    B, H, W, _ = x.shape
    xp = np.zeros((B, H + 2, W + 2, x.shape[3]))
    xp[:, 1:H + 1, 1:W + 1] = x
    y = np.zeros((B, H, W, w.shape[3]))
    mag = np.zeros_like(y)
    for dy in range(3):
        for dx in range(3):
            win = xp[:, dy:dy + H, dx:dx + W]
            y += win @ w[dy, dx]
            mag += np.abs(win) @ np.abs(w[dy, dx])
    return y, mag


def conv3x3_after_fusion(fused, w, scale=None, shift=None, relu=False):
    """act(scale * conv3x3_same(fused, w) + shift); returns (y float64, magnitude of the un-activated sum)."""
    y, mag = conv3x3_same(fused, w)
    if scale is not None:
        y = y * np.asarray(scale, dtype=np.float64)
        mag = mag * np.abs(np.asarray(scale, dtype=np.float64))
    if shift is not None:
        y = y + np.asarray(shift, dtype=np.float64)
        mag = mag + np.abs(np.asarray(shift, dtype=np.float64))
    if relu:
        y = np.maximum(y, 0.0)
    return y, mag


def conv3x3_same_grad(x, w, g_out):
    """Gradients of y = conv3x3_same(x, w) (float64): (g_x, g_w) and the sums of |terms| behind each of their elements."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    g = np.asarray(g_out, dtype=np.float64)
    B, H, W, Ci = x.shape
    gp = np.zeros((B, H + 2, W + 2, g.shape[3]))
    gp[:, 1:H + 1, 1:W + 1] = g
    xp = np.zeros((B, H + 2, W + 2, Ci))
    xp[:, 1:H + 1, 1:W + 1] = x
    g_x, mag_x = np.zeros_like(x), np.zeros_like(x)
    g_w, mag_w = np.zeros_like(w), np.zeros_like(w)
    for dy in range(3):
        for dx in range(3):
            # y[p] += x[p + (dy-1, dx-1)] . w[dy, dx]  =>  g_x[q] += g[q - (dy-1, dx-1)] . w[dy, dx]^T
            win = gp[:, 2 - dy:2 - dy + H, 2 - dx:2 - dx + W]
            g_x += win @ w[dy, dx].T
            mag_x += np.abs(win) @ np.abs(w[dy, dx]).T
            xwin = xp[:, dy:dy + H, dx:dx + W].reshape(-1, Ci)
            g_w[dy, dx] = xwin.T @ g.reshape(-1, g.shape[3])
            mag_w[dy, dx] = np.abs(xwin).T @ np.abs(g.reshape(-1, g.shape[3]))
    return g_x, g_w, mag_x, mag_w
