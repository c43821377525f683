"""Device-resident correspondence builder for a batch of frames (no reference
counterpart: the reference builds one frame per step on the host,
kitti_dataset.py:374-379, and is batch-1 only, sparse_pool_utils.py:98).

Each frame goes through shpl_build_avod (gen_sparse_pooling_input_avod +
produce_sparse_pooling_input fused); consecutive calls stack the frames into ONE
plan (rows / pixels offset by the frame index) so that a [B,H,W,C] batch is pooled
by one launch.  Nothing is copied to the host unless `read_counts` is asked for.
"""
import ctypes

import numpy as np
import torch

from . import _cabi, ops
from .ops import SparsePoolPlan, _ptr, _stream

_lib = _cabi.lib


def _dev_f64(p, dev):
    if isinstance(p, torch.Tensor):
        return p.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(p, dtype=np.float64)).to(dev)


def _dev_vox(v, dev):
    if isinstance(v, torch.Tensor):
        return v[:, :2].to(device=dev, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(v)[:, :2], dtype=np.int64)).to(dev)


def build_avod_plan(points, voxel_indices, P, im_size, bv_size, stride=(1, 1), src_hw=None, m_val=None,
                    want_coo=False, read_counts=True, plan=None):
    """points / voxel_indices / P: one frame (arrays) or a list of frames.

    points f64 [N,3] CUDA (or numpy, copied), voxel_indices i64 [N,2], P [3,4] numpy.
    Returns the stacked SparsePoolPlan; with want_coo also per-frame (Mij_pool,
    img_index_flip_pool, M_val f32, M_size) device tensors sized for N (first nnz rows valid)."""
    if not isinstance(points, (list, tuple)):
        points, voxel_indices, P = [points], [voxel_indices], [P]
        m_val = [m_val]
    elif m_val is None:
        m_val = [None] * len(points)
    frames = len(points)
    dev = points[0].device if isinstance(points[0], torch.Tensor) and points[0].is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    s_img, s_bv = int(stride[0]), int(stride[1])
    im_w, im_h = int(im_size[0]), int(im_size[1])
    bv_h, bv_w = int(bv_size[0]), int(bv_size[1])
    if src_hw is None:
        src_hw = (im_h // s_img, im_w // s_img)
    R = (bv_h // s_bv) * (bv_w // s_bv)
    pts = [_dev_f64(p, dev).reshape(-1, 3) for p in points]
    vox = [_dev_vox(v, dev) for v in voxel_indices]
    n_total = sum(int(p.shape[0]) for p in pts)
    if plan is None:
        plan = SparsePoolPlan(R, src_hw, n_total, dev, frames=frames)
    elif plan.capacity < n_total or plan.frames != frames or plan.rows_per_frame != R:
        raise ValueError("the plan passed in does not fit this batch")
    n_max = max([int(p.shape[0]) for p in pts] + [1])
    ws = ops.workspace(dev, n_max)
    coo = []
    keep = []
    for f in range(frames):
        N = int(pts[f].shape[0])
        Pf = np.ascontiguousarray(np.asarray(P[f], dtype=np.float64).reshape(12))
        mv = None
        if m_val[f] is not None:
            mv = _dev_f64(m_val[f], dev).reshape(-1)
            if mv.shape[0] < N:
                mv = torch.cat([mv, torch.zeros(N - mv.shape[0], dtype=torch.float64, device=dev)])
            keep.append(mv)
        if want_coo:
            Mij = torch.empty((max(N, 1), 2), dtype=torch.int64, device=dev)
            flip = torch.empty((max(N, 1), 3), dtype=torch.int64, device=dev)
            val = torch.empty(max(N, 1), dtype=torch.float32, device=dev)
            msize = torch.zeros(2, dtype=torch.int64, device=dev)
            coo.append((Mij, flip, val, msize))
        else:
            Mij = flip = val = msize = None
        st = plan.frame_struct(f)
        rc = _lib.shpl_build_avod(_ptr(pts[f]), _ptr(vox[f]), N, None, Pf.ctypes.data_as(ctypes.c_void_p),
                                  im_w, im_h, bv_h, bv_w, s_img, s_bv, _ptr(mv), int(src_hw[0]), int(src_hw[1]),
                                  _ptr(Mij), _ptr(flip), _ptr(val), _ptr(msize), ctypes.byref(st),
                                  f * plan.rows_per_frame, f * plan.src_per_frame, plan.entry_base(f),
                                  _ptr(ws), ws.numel(), _stream())
        _cabi.check(rc, "shpl_build_avod")
    plan.entry_bound = max(n_total, 1)
    plan._keep = (pts, vox, keep)
    if read_counts:
        plan.read_counts()
    return (plan, coo) if want_coo else plan


def build_pairs_plan(img_index, bv_index, im_size, bv_size, stride=(1, 1), src_hw=None, m_val=None, read_counts=True,
                     plan=None):
    """Batched form of produce_sparse_pooling_input for callers that already hold (img_index, bv_index) pairs -- the
    MV3D path (minibatch_mv3d_img.py:93 -> train_mv_voxel.py:325-326), whose reference is batch-1 only.

    img_index: one [3,n] (or [2,n]) array per frame, bv_index: one [n,2] array per frame, m_val: one [n] array per
    frame or None (numpy or CUDA tensors).  The frames are stacked into ONE plan (rows / pixels offset by the frame
    index), so a [B,H,W,C] batch is pooled by one launch.  Unlike produce_sparse_pooling_input the callers' img_index
    arrays are NOT modified in place (the floor / clamp is applied to device copies)."""
    if not isinstance(img_index, (list, tuple)):
        img_index, bv_index = [img_index], [bv_index]
        m_val = [m_val]
    elif m_val is None:
        m_val = [None] * len(img_index)
    frames = len(img_index)
    first = img_index[0]
    dev = first.device if isinstance(first, torch.Tensor) and first.is_cuda else torch.device("cuda", torch.cuda.current_device())
    s_img, s_bv = int(stride[0]), int(stride[1])
    im_w, im_h = int(im_size[0]), int(im_size[1])
    bv_h, bv_w = int(bv_size[0]), int(bv_size[1])
    if src_hw is None:
        src_hw = (im_h // s_img, im_w // s_img)
    R = (bv_h // s_bv) * (bv_w // s_bv)
    uv = [_dev_f64(np.asarray(a)[0:2] if not isinstance(a, torch.Tensor) else a[0:2], dev).contiguous() for a in img_index]
    bv = [(b.to(device=dev, dtype=torch.int64) if isinstance(b, torch.Tensor)
           else torch.from_numpy(np.ascontiguousarray(np.asarray(b), dtype=np.int64)).to(dev)).reshape(-1, 2).contiguous() for b in bv_index]
    ns = [int(u.shape[1]) for u in uv]
    for f in range(frames):
        if bv[f].shape[0] != ns[f]:
            raise ValueError("frame %d: bv_index has %d rows, img_index %d columns" % (f, bv[f].shape[0], ns[f]))
    n_total = sum(ns)
    if plan is None:
        plan = SparsePoolPlan(R, src_hw, n_total, dev, frames=frames)
    elif plan.capacity < n_total or plan.frames != frames or plan.rows_per_frame != R:
        raise ValueError("the plan passed in does not fit this batch")
    ws = ops.workspace(dev, max(ns + [1]))
    keep = []
    for f in range(frames):
        mv = None
        if m_val[f] is not None:
            mv = _dev_f64(m_val[f], dev).reshape(-1)
            if mv.shape[0] < ns[f]:
                mv = torch.cat([mv, torch.zeros(ns[f] - mv.shape[0], dtype=torch.float64, device=dev)])
            keep.append(mv)
        st = plan.frame_struct(f)
        rc = _lib.shpl_produce_input(_ptr(uv[f][0]), _ptr(uv[f][1]), _ptr(bv[f]), ns[f], im_w, im_h, bv_h, bv_w, s_img, s_bv,
                                     _ptr(mv), int(src_hw[0]), int(src_hw[1]), None, None, None, None, ctypes.byref(st),
                                     f * plan.rows_per_frame, f * plan.src_per_frame, plan.entry_base(f),
                                     _ptr(ws), ws.numel(), _stream())
        _cabi.check(rc, "shpl_produce_input")
    plan.entry_bound = max(n_total, 1)
    plan._keep = (uv, bv, keep)
    if read_counts:
        plan.read_counts()
    return plan
