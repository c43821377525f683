"""Drop-in mirror of the reference's SHPL module, running on B200 CUDA kernels.

Same function names, positional order, return structures and error behaviour as
  /root/reference/avod/avod/utils/sparse_pool_utils.py          (:6, :22, :61, :96, :105, :120)
  /root/reference/MV3D_TF_release/lib/utils/sparse_pool_utils.py (identical twin)
so that the call sites
  kitti_dataset.py:376-379, rpn_model.py:335-336, fusion_vgg_pyramid.py:59-60,
  retinanet_model.py:337-340, MV3D_voxel_train.py:89-91, network.py:242-246
keep their shape.  numpy in -> numpy out (the reference's host half works on numpy);
torch CUDA tensors in -> torch CUDA tensors out (no host round trip).  Feature maps
are torch CUDA tensors, NHWC float32, where the reference has TF tensors.

All arithmetic runs in libshpl.so (include/shpl.h); there is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _cabi, ops
from .ops import SparsePoolPlan, _ptr, _stream

_lib = _cabi.lib

# TF-CPU rejects out-of-range gather / SpMM indices with InvalidArgumentError; the
# kernels never read them.  With STRICT_INDEX_CHECK the layer raises ValueError when
# the plan holds such entries; without it they contribute nothing (TF-GPU behaviour).
STRICT_INDEX_CHECK = True

PLAN_KEY = "shpl_plan"      # extra key of produce_sparse_pooling_input's dict: the cached CSR plan


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype, device):
    if isinstance(a, torch.Tensor):
        if not a.is_cuda and a.dtype == dtype and a.is_contiguous() and a.is_pinned():
            # pinned host tensor: asynchronous upload, no stream synchronise.  Like torch's non_blocking=True the
            # caller keeps the buffer unchanged until the next call that synchronises (produce_sparse_pooling_input
            # reads the counters back, so: until it returns).
            return a.to(device=device, non_blocking=True)
        return a.to(device=device, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a), dtype={torch.float64: np.float64, torch.int64: np.int64,
                                                     torch.float32: np.float32, torch.int32: np.int32}[dtype])
    return torch.from_numpy(arr).to(device)


def _as_int(x, what):
    xi = int(x)
    if xi != x:
        raise ValueError("%s must be an integer for the CUDA builder, got %r" % (what, x))
    return xi


class SparseTensor:
    """Mirror of tf.SparseTensor(indices, values, dense_shape) -- the type `M` has at the
    reference's call sites (rpn_model.py:292-293, :330-331; MV3D_voxel_train.py:41-46).
    A CSR plan built for it is cached on the object."""

    def __init__(self, indices, values, dense_shape, plan=None):
        self.indices = indices
        self.values = values
        self.dense_shape = dense_shape
        self._plan = plan
        self._plan_key = None

    @classmethod
    def from_sparse_pooling_input(cls, d):
        """From the dict produce_sparse_pooling_input returns (what create_feed_dict feeds,
        rpn_model.py:842-847)."""
        return cls(d["Mij_pool"], d["M_val"], d["M_size"], plan=d.get(PLAN_KEY))


# --------------------------------------------------------------------------- host half
class SparsePoolingInput(dict):
    """The dict gen_sparse_pooling_input_avod returns (sparse_pool_utils.py:20), evaluated lazily.

    'bv_size' and 'img_size' are there at once.  'bv_index' [n,2] and 'img_index' [3,n] need the
    number of pairs that survive the image clip, i.e. a device->host read; they are computed (by
    shpl_gen_input_avod) the first time anything looks at them.  When the dict goes straight into
    produce_sparse_pooling_input -- the reference's own call pattern, kitti_dataset.py:376-378 -- nobody
    looks, and produce runs the fused builder (shpl_build_avod) instead: one pipeline, one read-back.
    Looking later still shows what the reference would show, including produce's in-place floor/clamp
    of img_index (:30-34)."""

    _LAZY = ("bv_index", "img_index")

    def __init__(self, pts, vox, P, im_size, bv_size, as_numpy, device, n_dev=None):
        super().__init__(bv_size=np.array([bv_size[0], bv_size[1]]), img_size=np.array(im_size))
        self._pts, self._vox, self._P = pts, vox, P
        self._n_dev = n_dev           # device int32 count of valid candidates (feeder output not yet read back)
        self._as_numpy, self._device = as_numpy, device
        self._done = False
        self._mutations = []          # image strides applied in place by produce calls made before materialisation

    def _materialize(self):
        if self._done:
            return
        self._done = True
        dev, pts, vox = self._device, self._pts, self._vox
        im_size = dict.__getitem__(self, "img_size")
        N = pts.shape[0]
        bv_out = torch.empty((max(N, 1), 2), dtype=torch.int64, device=dev)
        uv_out = torch.empty((2, max(N, 1)), dtype=torch.float64, device=dev)
        counts = torch.zeros(8, dtype=torch.int32, device=dev)
        ws = ops.workspace(dev, N)
        rc = _lib.shpl_gen_input_avod(_ptr(pts), _ptr(vox), N, self._n_dev, self._P.ctypes.data_as(ctypes.c_void_p),
                                      _as_int(im_size[0], "im_size"), _as_int(im_size[1], "im_size"),
                                      _ptr(bv_out), _ptr(uv_out[0]), _ptr(uv_out[1]), _ptr(counts), _ptr(ws),
                                      ws.numel(), _stream())
        _cabi.check(rc, "shpl_gen_input_avod")
        n = int(counts[0].item())
        img_index = torch.zeros((3, n), dtype=torch.float64, device=dev)
        img_index[0:2] = uv_out[:, :n]
        for s_img in self._mutations:          # what an earlier produce call did to the reference's array (:30-34)
            for axis in (0, 1):
                lim = float(int(im_size[axis]) // s_img)
                row = torch.floor(img_index[axis] / s_img)
                img_index[axis] = torch.where(row >= lim, torch.full_like(row, lim - 1), row)
        bv_index = bv_out[:n]
        if self._as_numpy:
            bv_index, img_index = bv_index.cpu().numpy(), img_index.cpu().numpy()
        dict.__setitem__(self, "bv_index", bv_index)
        dict.__setitem__(self, "img_index", img_index)

    def __missing__(self, key):
        if key in self._LAZY:
            self._materialize()
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def get(self, key, default=None):
        if key in self._LAZY:
            self._materialize()
        return dict.get(self, key, default)

    def __contains__(self, key):
        return key in self._LAZY or dict.__contains__(self, key)

    def __iter__(self):
        self._materialize()
        return dict.__iter__(self)

    def __len__(self):
        return 4

    def keys(self):
        self._materialize()
        return dict.keys(self)

    def items(self):
        self._materialize()
        return dict.items(self)

    def values(self):
        self._materialize()
        return dict.values(self)

    def copy(self):
        self._materialize()
        return dict(self)


def gen_sparse_pooling_input_avod(points, voxel_indices, stereo_calib, im_size, bv_size):
    """sparse_pool_utils.py:6-20 on the GPU (shpl_gen_input_avod).

    points [N,3] f64 camera frame, voxel_indices [N,>=2] (x, zflip), stereo_calib.p2 [3,4],
    im_size [W,H], bv_size [H_b,W_b].  Returns the reference's dict (a lazily evaluated
    SparsePoolingInput: see there)."""
    as_numpy = not isinstance(points, torch.Tensor)
    dev = points.device if (not as_numpy and points.is_cuda) else _device()
    pts = _to_dev(points, torch.float64, dev).reshape(-1, 3)
    vox = voxel_indices if isinstance(voxel_indices, torch.Tensor) else np.asarray(voxel_indices)
    vox = _to_dev(vox if vox.shape[1] == 2 else vox[:, :2], torch.int64, dev)
    P = np.ascontiguousarray(np.asarray(stereo_calib.p2, dtype=np.float64).reshape(12))
    _as_int(im_size[0], "im_size"), _as_int(im_size[1], "im_size")
    return SparsePoolingInput(pts, vox, P, im_size, bv_size, as_numpy, dev)


def _produce_outputs(plan, Mij, flip, R, M_val, n_mval, as_numpy, dev):
    plan.read_counts()
    nnz = plan.nnz[0]
    if M_val is not None and n_mval != nnz:
        raise ValueError("M_val has %d entries but M has %d columns (tf.SparseTensor would reject it)" % (n_mval, nnz))
    plan.flip = out_flip = flip[:nnz]
    plan.Mij = out_Mij = Mij[:nnz]
    M_size = np.array([R, nnz]).astype(int)
    if as_numpy:
        out_Mij, out_flip = out_Mij.cpu().numpy(), out_flip.cpu().numpy()
        out_val = np.ones(nnz) if M_val is None else M_val
        bev_flip = np.zeros((0, 3))
    else:
        out_val = _ones(nnz, dev) if M_val is None else M_val
        bev_flip = torch.zeros((0, 3), device=dev)
    return {"Mij_pool": out_Mij, "M_val": out_val, "M_size": M_size, "img_index_flip_pool": out_flip,
            "bev_index_flip_pool": bev_flip, PLAN_KEY: plan}


def _ones(n, dev):
    """np.ones(n) of the reference (:52) as a device tensor.  A FRESH tensor every call, like the reference's: a
    caller that scales or normalises M_val in place must not change any other frame's weights."""
    return torch.ones(n, dtype=torch.float64, device=dev)


def _mval_to_dev(M_val, n, dev):
    mval_dev = _to_dev(M_val, torch.float64, dev).reshape(-1)
    n_mval = mval_dev.shape[0]
    if n_mval < n:      # the kernel indexes it by output column k < nnz <= n: keep every read in bounds
        mval_dev = torch.cat([mval_dev, torch.zeros(n - n_mval, dtype=torch.float64, device=dev)])
    return mval_dev, n_mval


def produce_sparse_pooling_input(input_dict, M_val=None, stride=[1, 1]):
    """sparse_pool_utils.py:22-58 on the GPU (shpl_produce_input), plus the CSR plan.

    stride[0] scales the image, stride[1] the BEV (the reference's comment has them
    swapped).  Mutates input_dict['img_index'] in place like the reference (:30).
    The returned dict has the reference's five keys and one extra, 'shpl_plan'."""
    s_img, s_bv = _as_int(stride[0], "stride[0]"), _as_int(stride[1], "stride[1]")
    bv_size = input_dict["bv_size"]
    im_size = input_dict["img_size"]
    im_w, im_h = _as_int(im_size[0], "img_size"), _as_int(im_size[1], "img_size")
    bv_h, bv_w = _as_int(bv_size[0], "bv_size"), _as_int(bv_size[1], "bv_size")
    Hp, Wp = im_h // s_img, im_w // s_img
    R = (bv_h // s_bv) * (bv_w // s_bv)

    if isinstance(input_dict, SparsePoolingInput) and not input_dict._done and not input_dict._mutations:
        # straight from gen_sparse_pooling_input_avod: fused builder, the intermediate dict is never made
        dev, pts, vox = input_dict._device, input_dict._pts, input_dict._vox
        N = int(pts.shape[0])
        mval_dev, n_mval = (None, 0) if M_val is None else _mval_to_dev(M_val, N, dev)
        plan = SparsePoolPlan(R, (Hp, Wp), N, dev, zero_meta=False)
        plan.entry_bound = max(N, 1)
        cap = max(N, 1)
        Mij = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        flip = torch.empty((cap, 3), dtype=torch.int64, device=dev)
        ws = ops.workspace(dev, N)
        st = plan.frame_struct(0)
        rc = _lib.shpl_build_avod(_ptr(pts), _ptr(vox), N, input_dict._n_dev, input_dict._P.ctypes.data_as(ctypes.c_void_p),
                                  im_w, im_h, bv_h, bv_w, s_img, s_bv, _ptr(mval_dev), 0, 0, _ptr(Mij), _ptr(flip),
                                  None, None, ctypes.byref(st), 0, 0, None, _ptr(ws), ws.numel(), _stream())
        _cabi.check(rc, "shpl_build_avod")
        input_dict._mutations.append(s_img)
        return _produce_outputs(plan, Mij, flip, R, M_val, n_mval, input_dict._as_numpy, dev)

    bv_index = input_dict["bv_index"]
    img_index = input_dict["img_index"]
    assert img_index.shape[0] == 3, 'wrong img_index shape, should be 3xN instead ' + str(tuple(img_index.shape))
    as_numpy = not isinstance(img_index, torch.Tensor)
    dev = img_index.device if (not as_numpy and img_index.is_cuda) else _device()
    n = int(img_index.shape[1])

    uv = _to_dev(img_index[0:2], torch.float64, dev)
    bv = _to_dev(bv_index, torch.int64, dev).reshape(-1, 2)
    if bv.shape[0] != n:
        raise ValueError("bv_index has %d rows, img_index %d columns" % (bv.shape[0], n))
    mval_dev, n_mval = (None, 0) if M_val is None else _mval_to_dev(M_val, n, dev)

    plan = SparsePoolPlan(R, (Hp, Wp), n, dev, zero_meta=False)
    plan.entry_bound = max(n, 1)
    cap = max(n, 1)
    Mij = torch.empty((cap, 2), dtype=torch.int64, device=dev)
    flip = torch.empty((cap, 3), dtype=torch.int64, device=dev)
    ws = ops.workspace(dev, n)
    st = plan.frame_struct(0)
    rc = _lib.shpl_produce_input(_ptr(uv[0]), _ptr(uv[1]), _ptr(bv), n, im_w, im_h, bv_h, bv_w, s_img, s_bv,
                                 _ptr(mval_dev), 0, 0, _ptr(Mij), _ptr(flip), None, None,
                                 ctypes.byref(st), 0, 0, None, _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "shpl_produce_input")
    out = _produce_outputs(plan, Mij, flip, R, M_val, n_mval, as_numpy, dev)
    # the in-place mutation of the caller's img_index (:30-34)
    if as_numpy:
        img_index[0:2, :] = uv.cpu().numpy()
    elif uv.data_ptr() != img_index.data_ptr():
        img_index[0:2] = uv.to(img_index.device)
    return out


# ------------------------------------------------------------------------- device half
def _resolve_plan(M, source_index, n_rows, src_hw, device):
    """The CSR plan of M for a destination map of n_rows cells and a source map of src_hw."""
    if isinstance(M, dict):
        M = SparseTensor.from_sparse_pooling_input(M)
    plan = getattr(M, "_plan", None)
    if isinstance(M, SparsePoolPlan):
        plan = M
    if plan is not None and plan.rows_per_frame == n_rows and plan.src_hw == tuple(int(x) for x in src_hw):
        return plan
    if source_index is None:
        raise ValueError("no cached plan matches the feature maps and no source_index was given")
    key = (n_rows, tuple(src_hw))
    cached = getattr(M, "_coo_plans", None)
    if cached is not None and key in cached:
        return cached[key]
    dense_shape = [int(x) for x in np.asarray(M.dense_shape.cpu() if isinstance(M.dense_shape, torch.Tensor) else M.dense_shape)]
    if dense_shape[0] != n_rows:
        raise ValueError("M has %d rows but the pooled map has %d cells (tf.reshape would fail, "
                         "sparse_pool_utils.py:103)" % (dense_shape[0], n_rows))
    ind = M.indices if isinstance(M.indices, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(M.indices))
    val = M.values if isinstance(M.values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(M.values))
    sidx = source_index if isinstance(source_index, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(source_index))
    plan = ops.plan_from_coo(ind.to(device), val.to(device), sidx.to(device), n_rows, src_hw)
    plan.read_counts()
    try:
        if cached is None:
            M._coo_plans = {}
        M._coo_plans[key] = plan
    except AttributeError:
        pass
    return plan


_dual_announced = False


def _announce_dual():
    """The reference prints this while it builds the TF graph, i.e. once (:80); an eager layer would
    print it every step, so it is printed on first use only."""
    global _dual_announced
    if not _dual_announced:
        _dual_announced = True
        print('using dual sparse pooling')


def _check_oob(plan):
    if STRICT_INDEX_CHECK and plan.n_oob is not None and sum(plan.n_oob) > 0:
        raise ValueError("SHPL: %d entries of M index outside the feature maps "
                         "(TF-CPU raises InvalidArgumentError here); set "
                         "sparse_pool_utils.STRICT_INDEX_CHECK = False to drop them" % sum(plan.n_oob))


def _sparse_pool_op(M, input, source_index, pooled_size):
    """sparse_pool_utils.py:96-103: gather_nd -> sparse_tensor_dense_matmul -> reshape.
    input [1,H_i,W_i,C] CUDA float32; returns the pooled map of shape pooled_size."""
    ops.require_cuda(input, "input")
    n_rows = int(pooled_size[1]) * int(pooled_size[2])
    plan = _resolve_plan(M, source_index, n_rows, (input.shape[1], input.shape[2]), input.device)
    _check_oob(plan)
    if int(pooled_size[3]) != input.shape[3]:
        raise ValueError("pooled_size depth %d != input depth %d" % (int(pooled_size[3]), input.shape[3]))
    out = ops.sparse_pool(None, input, plan, transposed=False)
    return out.reshape([int(x) for x in pooled_size])


def _sparse_pool_trans_op(M, input, source_index, pooled_size):
    """sparse_pool_utils.py:105-117: sparse_transpose + matmul -> scatter_nd (duplicates summed).
    input [1,H_b,W_b,C] CUDA float32; returns [1,H_i,W_i,C] = pooled_size."""
    ops.require_cuda(input, "input")
    n_rows = input.shape[1] * input.shape[2]
    plan = _resolve_plan(M, source_index, n_rows, (int(pooled_size[1]), int(pooled_size[2])), input.device)
    _check_oob(plan)
    out = ops.sparse_pool(None, input, plan, transposed=True)
    return out.reshape([int(x) for x in pooled_size])


def concat_bn_op(inputs, axis, training, bn_modules=None):
    """sparse_pool_utils.py:120-124: batch-norm each map on its own, then concat.
    The reference creates slim.batch_norm variables in the graph; here the two
    normalisers are passed in (SparsePoolLayer owns them).  Without modules the
    fresh-variable behaviour of slim is reproduced: batch statistics, beta = 0, no gamma."""
    outs = []
    for i, x in enumerate(inputs):
        xc = x.permute(0, 3, 1, 2)
        if bn_modules is not None:
            bn_modules[i].train(bool(training))
            y = bn_modules[i](xc)
        else:
            y = torch.nn.functional.batch_norm(xc, None, None, training=True, eps=1e-3)
        outs.append(y.permute(0, 2, 3, 1))
    return torch.cat(outs, dim=axis)


def sparse_pool_layer(inputs, feature_depths, M, img_index_flip=None, bv_index=None, use_bn=False, training=True,
                      bn_modules=None, out=None):
    """sparse_pool_utils.py:61-92.

    inputs = [input_bv, input_img] NHWC float32 CUDA tensors; feature_depths =
    [C_img, C_bev] (depths of the pooled maps, :62); M a SparseTensor (or the dict
    from produce_sparse_pooling_input); img->bev runs when img_index_flip is given,
    bev->img when bv_index is not None (its content is ignored, :79).
    Returns (bv_fused, img_fused).

    out (extension, SURVEY.md 8(d) "sparse-only"): (bv_fused_buffer, img_fused_buffer or None) -- preallocated fused maps
    whose leading channels ARE inputs[0] / inputs[1] (their producers wrote straight into them: inputs[k] must be the
    view out[k][..., :C]).  Then no concat copy is made: only the pooled channels are written, and the buffers are
    returned."""
    input_bv, input_img = inputs[0], inputs[1]
    if out is not None:
        if use_bn:
            raise ValueError("out= (no-concat form) cannot be combined with use_bn")
        return _sparse_pool_layer_into(input_bv, input_img, feature_depths, M, img_index_flip, bv_index, out)
    ops.require_cuda(input_bv, "inputs[0]")
    ops.require_cuda(input_img, "inputs[1]")
    plan = None
    if img_index_flip is not None or bv_index is not None:
        plan = _resolve_plan(M, img_index_flip, input_bv.shape[1] * input_bv.shape[2],
                             (input_img.shape[1], input_img.shape[2]), input_bv.device)
        _check_oob(plan)
    if img_index_flip is not None and bv_index is not None and not use_bn:
        if int(feature_depths[0]) != input_img.shape[3] or int(feature_depths[1]) != input_bv.shape[3]:
            raise ValueError("feature_depths %r do not match the maps (%d image, %d BEV channels)"
                             % (list(feature_depths), input_img.shape[3], input_bv.shape[3]))
        _announce_dual()
        return ops.sparse_pool_dual(input_bv, input_img, plan)
    if img_index_flip is not None:
        if int(feature_depths[0]) != input_img.shape[3]:
            raise ValueError("feature_depths[0]=%d but the image map has %d channels" % (feature_depths[0], input_img.shape[3]))
        if use_bn:
            img_pooled = ops.sparse_pool(None, input_img, plan, transposed=False).reshape(
                input_bv.shape[0], input_bv.shape[1], input_bv.shape[2], input_img.shape[3])
            bv_fused = concat_bn_op([input_bv, img_pooled], axis=3, training=training,
                                    bn_modules=None if bn_modules is None else bn_modules[0])
        else:
            bv_fused = ops.sparse_pool(input_bv, input_img, plan, transposed=False)
    else:
        bv_fused = input_bv

    if bv_index is not None:
        if int(feature_depths[1]) != input_bv.shape[3]:
            raise ValueError("feature_depths[1]=%d but the BEV map has %d channels" % (feature_depths[1], input_bv.shape[3]))
        if use_bn:
            bv_pooled = ops.sparse_pool(None, input_bv, plan, transposed=True).reshape(
                input_img.shape[0], input_img.shape[1], input_img.shape[2], input_bv.shape[3])
            img_fused = concat_bn_op([input_img, bv_pooled], axis=3, training=training,
                                     bn_modules=None if bn_modules is None else bn_modules[1])
        else:
            img_fused = ops.sparse_pool(input_img, input_bv, plan, transposed=True)
    else:
        img_fused = input_img
    return bv_fused, img_fused


def _sparse_pool_layer_into(input_bv, input_img, feature_depths, M, img_index_flip, bv_index, out):
    ops.require_cuda(input_bv, "inputs[0]")
    ops.require_cuda(input_img, "inputs[1]")
    bv_buf, img_buf = out[0], out[1]
    plan = _resolve_plan(M, img_index_flip, input_bv.shape[1] * input_bv.shape[2], (input_img.shape[1], input_img.shape[2]),
                         input_bv.device)
    _check_oob(plan)

    def leading_view(buf, x, what):
        if buf.data_ptr() != x.data_ptr() or tuple(buf.shape[:3]) != tuple(x.shape[:3]) or x.stride(2) != buf.shape[3]:
            raise ValueError("%s must be the view out[...][..., :C] of its fused buffer (no-concat form)" % what)
    bv_fused, img_fused = input_bv, input_img
    if img_index_flip is not None:
        leading_view(bv_buf, input_bv, "inputs[0]")
        if bv_buf.shape[3] != input_bv.shape[3] + input_img.shape[3] or int(feature_depths[0]) != input_img.shape[3]:
            raise ValueError("fused BEV buffer must have C_bev + C_img channels")
        bv_fused = ops.sparse_pool_into(bv_buf, input_img, plan, transposed=False)
    if bv_index is not None:
        leading_view(img_buf, input_img, "inputs[1]")
        if img_buf.shape[3] != input_img.shape[3] + input_bv.shape[3] or int(feature_depths[1]) != input_bv.shape[3]:
            raise ValueError("fused image buffer must have C_img + C_bev channels")
        _announce_dual()
        img_fused = ops.sparse_pool_into(img_buf, input_bv, plan, transposed=True)
    return bv_fused, img_fused


class SparsePoolLayer(torch.nn.Module):
    """Module form of sparse_pool_layer for callers that want use_bn=True with persistent
    batch-norm state (slim.batch_norm defaults: decay 0.999, epsilon 1e-3, beta only)."""

    def __init__(self, c_bev, c_img, dual=False, use_bn=False):
        super().__init__()
        self.c_bev, self.c_img, self.dual, self.use_bn = c_bev, c_img, dual, use_bn
        if use_bn:
            def bn(c):
                m = torch.nn.BatchNorm2d(c, eps=1e-3, momentum=1e-3, affine=True)
                m.weight.requires_grad_(False)      # slim default scale=False
                return m
            self.bn_bev = torch.nn.ModuleList([bn(c_bev), bn(c_img)])
            self.bn_img = torch.nn.ModuleList([bn(c_img), bn(c_bev)]) if dual else None

    def forward(self, input_bv, input_img, M, img_index_flip=None):
        mods = [self.bn_bev, self.bn_img] if self.use_bn else None
        return sparse_pool_layer([input_bv, input_img], [self.c_img, self.c_bev], M, img_index_flip=img_index_flip,
                                 bv_index=(np.zeros((1, 3)) if self.dual else None), use_bn=self.use_bn,
                                 training=self.training, bn_modules=mods)
