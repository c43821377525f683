"""Lean, allocation-free SHPL frame pipeline: correspondence build + forward +
backward of every SHPL layer of a model, on preallocated buffers, ready to be
captured in a CUDA graph.

The reference runs this path once per training step for one frame
(trainer.py:207-235 -> kitti_dataset.py:374-379 -> rpn_model.py:291-361); the
layers of one model share the frame's points, so one FramePipeline owns one plan
per layer.  Every call below is a direct ctypes call into libshpl.so with
arguments bound once in __init__ -- no tensor is created per step, nothing is
copied to the host, and the only thing that changes between steps is the content
of the buffers.
"""
import ctypes
import numpy as np
import torch

from . import _cabi, ops
from .layer_spec import LayerSpec  # noqa: F401  (re-exported: pipeline.LayerSpec)
from .ops import SparsePoolPlan

_lib = _cabi.lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class _LayerState:
    def __init__(self, spec, n_max, device):
        self.spec = spec
        f32 = dict(dtype=torch.float32, device=device)
        self.plan = SparsePoolPlan(spec.R, spec.img_hw, n_max, device)
        self.plan_struct = self.plan.frame_struct(0)
        self.plan_struct.heavy_cap = 0      # lean path: every cell summed sequentially in the main kernels (heavy_len = 0)
        need = int(_lib.shpl_build_workspace_bytes(int(n_max)))
        self.ws = torch.empty(need, dtype=torch.uint8, device=device)     # own scratch: layers build concurrently
        self.fused_bev = torch.empty((1,) + spec.bev_hw + (spec.c_bev + spec.c_img,), **f32)
        self.g_bev = torch.empty((1,) + spec.bev_hw + (spec.c_bev,), **f32)
        self.g_img = torch.empty((1,) + spec.img_hw + (spec.c_img,), **f32)
        if spec.dual:
            self.fused_img = torch.empty((1,) + spec.img_hw + (spec.c_img + spec.c_bev,), **f32)
        pl = self.plan
        self.plan_ptrs = pl.ptrs8()


class FramePipeline:
    """no_concat=True runs the layers in the "sparse-only" form of SURVEY.md 8(d): the producer of a destination map is
    taken to have written its channels straight into the fused buffer (L.fused_bev[..., :C_b], L.fused_img[..., :C_i]),
    so the forward writes only the pooled channels (shpl_pool_forward_into) and the single-direction backward computes
    only the gradient of the gathered map (shpl_pool_backward_from): the gradient of the destination map is the view
    g_fused[..., :C_d].  The `bev` argument of forward_layer is then unused for single-direction layers."""

    def __init__(self, layers, n_points_max, device, no_concat=False):
        self.device = device
        self.n_max = int(n_points_max)
        self.no_concat = bool(no_concat)
        self.layers = [_LayerState(s, self.n_max, device) for s in layers]

    # -- correspondence build: shpl_build_avod per layer (different strides, same points) --
    def build(self, points, voxel_indices, P, n_points, stream, n_dev=None):
        P = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(12))
        for L in self.layers:
            s = L.spec
            rc = _lib.shpl_build_avod(_p(points), _p(voxel_indices), int(n_points), n_dev, P.ctypes.data_as(ctypes.c_void_p),
                                      s.im_size[0], s.im_size[1], s.bv_size[0], s.bv_size[1], s.stride[0], s.stride[1],
                                      None, s.img_hw[0], s.img_hw[1], None, None, None, None,
                                      ctypes.byref(L.plan_struct), 0, 0, None, _p(L.ws), L.ws.numel(), stream)
            _cabi.check(rc, "shpl_build_avod")

    def build_layer(self, i, points, voxel_indices, P, n_points, stream, n_dev=None):
        L = self.layers[i]
        s = L.spec
        P = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(12))
        rc = _lib.shpl_build_avod(_p(points), _p(voxel_indices), int(n_points), n_dev, P.ctypes.data_as(ctypes.c_void_p),
                                  s.im_size[0], s.im_size[1], s.bv_size[0], s.bv_size[1], s.stride[0], s.stride[1],
                                  None, s.img_hw[0], s.img_hw[1], None, None, None, None,
                                  ctypes.byref(L.plan_struct), 0, 0, None, _p(L.ws), L.ws.numel(), stream)
        _cabi.check(rc, "shpl_build_avod")

    # -- forward of layer i: img -> bev (and bev -> img when dual, same launch), concat fused in --
    def forward_layer(self, i, bev, img, stream, nnz_max=None):
        L = self.layers[i]
        s, pl = L.spec, L.plan
        bound = int(nnz_max) if nnz_max else pl.capacity
        if self.no_concat:
            if s.dual:
                rc = _lib.shpl_pool_forward_into_dual(_p(bev), _p(img), *L.plan_ptrs, bound, 0, s.R, s.c_bev, s.Q, s.c_img,
                                                      _p(L.fused_bev), _p(L.fused_img), stream)
                _cabi.check(rc, "shpl_pool_forward_into_dual")
                return
            rc = _lib.shpl_pool_forward_into(_p(img), _p(pl.row_ptr), _p(pl.csr_row), _p(pl.csr_src), _p(pl.csr_val),
                                             bound, 0, s.R, s.Q, s.c_img, _p(L.fused_bev), s.c_bev + s.c_img, s.c_bev, stream)
            _cabi.check(rc, "shpl_pool_forward_into")
            return
        if s.dual:
            rc = _lib.shpl_pool_forward_dual(_p(bev), _p(img), *L.plan_ptrs, bound, 0, s.R, s.c_bev, s.Q, s.c_img,
                                             _p(L.fused_bev), _p(L.fused_img), stream)
            _cabi.check(rc, "shpl_pool_forward_dual")
            return
        rc = _lib.shpl_pool_forward(_p(bev), _p(img), _p(pl.row_ptr), _p(pl.csr_row), _p(pl.csr_src), _p(pl.csr_val),
                                    bound, 0, s.R, s.c_bev, s.Q, s.c_img, _p(L.fused_bev), stream)
        _cabi.check(rc, "shpl_pool_forward")

    # -- backward of layer i from the upstream gradients of the fused maps --
    def backward_layer(self, i, g_fused_bev, g_fused_img, stream, nnz_max=None):
        L = self.layers[i]
        s, pl = L.spec, L.plan
        bound = int(nnz_max) if nnz_max else pl.capacity
        if s.dual:   # AddN of the two partial gradients of each input formed in the kernel
            rc = _lib.shpl_pool_backward_dual(_p(g_fused_bev), _p(g_fused_img), *L.plan_ptrs, bound, 0, s.R, s.c_bev,
                                              s.Q, s.c_img, _p(L.g_bev), _p(L.g_img), stream)
            _cabi.check(rc, "shpl_pool_backward_dual")
            return
        if self.no_concat:                                   # g_bev is the view g_fused_bev[..., :C_b]: nothing to do for it
            rc = _lib.shpl_pool_backward_from(_p(g_fused_bev), s.c_bev + s.c_img, s.c_bev, _p(pl.pix_ptr), _p(pl.csrT_pix),
                                              _p(pl.csrT_dst), _p(pl.csrT_val), bound, 0, s.R, s.Q, s.c_img, _p(L.g_img), stream)
            _cabi.check(rc, "shpl_pool_backward_from")
            return
        rc = _lib.shpl_pool_backward(_p(g_fused_bev), _p(pl.pix_ptr), _p(pl.csrT_pix), _p(pl.csrT_dst), _p(pl.csrT_val),
                                     bound, 0, s.R, s.c_bev, s.Q, s.c_img, _p(L.g_bev), _p(L.g_img), stream)
        _cabi.check(rc, "shpl_pool_backward")


class PairsPipeline:
    """The same lean pipeline for callers that already hold (img_index, bv_index, M_val) pairs -- the MV3D path
    (minibatch_mv3d_img.py:93 -> train_mv_voxel.py:325-326 -> network.py:242-246) and the direct-pair configs of
    BASELINE.json -- for a BATCH of frames stacked into one plan (rows / pixels offset by the frame index), so that a
    [B,H,W,C] batch is pooled by one launch each way.  The reference is batch-1 (sparse_pool_utils.py:98); the batch
    is this repo's sharding unit.  Single direction (img -> bev), like every MV3D / pre-RPN call site."""

    def __init__(self, spec, frames, n_max_per_frame, device):
        if spec.dual:
            raise ValueError("PairsPipeline pools img -> bev only")
        self.spec, self.frames, self.n_max, self.device = spec, int(frames), int(n_max_per_frame), device
        f32 = dict(dtype=torch.float32, device=device)
        self.plan = SparsePoolPlan(spec.R, spec.img_hw, self.frames * self.n_max, device, frames=self.frames)
        self.structs = []
        for f in range(self.frames):
            st = self.plan.frame_struct(f)
            st.heavy_cap = 0                  # every cell summed sequentially in the main kernels (heavy_len = 0)
            self.structs.append(st)
        self.ws = torch.empty(int(_lib.shpl_build_workspace_bytes(self.n_max)), dtype=torch.uint8, device=device)
        self.uv = torch.empty((self.frames, 2, self.n_max), dtype=torch.float64, device=device)   # floored in place by the builder
        self.fused_bev = torch.empty((self.frames,) + spec.bev_hw + (spec.c_bev + spec.c_img,), **f32)
        self.g_bev = torch.empty((self.frames,) + spec.bev_hw + (spec.c_bev,), **f32)
        self.g_img = torch.empty((self.frames,) + spec.img_hw + (spec.c_img,), **f32)

    def build_frame(self, f, uv, bv_index, m_val, n, stream):
        """produce_sparse_pooling_input for frame f of the batch: uv f64 [2, >=n] (rows 0, 1 of img_index; copied, the
        caller's array is not touched), bv_index i64 [n,2], m_val f64 [n] or None.  Call with f = 0, 1, ... in order
        (the entry offsets chain through the plan counters on the device).  torch's current stream must be `stream`."""
        s, n = self.spec, int(n)
        if n > self.n_max:
            raise ValueError("frame of %d pairs, pipeline sized for %d" % (n, self.n_max))
        self.uv[f, :, :n].copy_(uv[:, :n], non_blocking=True)
        u, v = self.uv[f, 0], self.uv[f, 1]
        rc = _lib.shpl_produce_input(_p(u), _p(v), _p(bv_index), n, s.im_size[0], s.im_size[1], s.bv_size[0], s.bv_size[1],
                                     s.stride[0], s.stride[1], None if m_val is None else _p(m_val), s.img_hw[0], s.img_hw[1],
                                     None, None, None, None, ctypes.byref(self.structs[f]), f * s.R, f * s.Q,
                                     self.plan.entry_base(f), _p(self.ws), self.ws.numel(), stream)
        _cabi.check(rc, "shpl_produce_input")

    def forward(self, bev, img, stream, nnz_max=None):
        s, pl = self.spec, self.plan
        bound = int(nnz_max) if nnz_max else pl.capacity
        rc = _lib.shpl_pool_forward(_p(bev), _p(img), _p(pl.row_ptr), _p(pl.csr_row), _p(pl.csr_src), _p(pl.csr_val),
                                    bound, 0, self.frames * s.R, s.c_bev, self.frames * s.Q, s.c_img, _p(self.fused_bev), stream)
        _cabi.check(rc, "shpl_pool_forward")

    def backward(self, g_fused_bev, stream, nnz_max=None):
        s, pl = self.spec, self.plan
        bound = int(nnz_max) if nnz_max else pl.capacity
        rc = _lib.shpl_pool_backward(_p(g_fused_bev), _p(pl.pix_ptr), _p(pl.csrT_pix), _p(pl.csrT_dst), _p(pl.csrT_val),
                                     bound, 0, self.frames * s.R, s.c_bev, self.frames * s.Q, s.c_img, _p(self.g_bev),
                                     _p(self.g_img), stream)
        _cabi.check(rc, "shpl_pool_backward")
