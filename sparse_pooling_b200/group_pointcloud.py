"""The sparse half of the reference's VFE feature net, running on B200 CUDA kernels.

  /root/reference/MV3D_TF_release/lib/networks/group_pointcloud.py:84-85
      self.outputs = tf.scatter_nd(self.coordinate, voxelwise, [batch_size, 10, INPUT_HEIGHT, INPUT_WIDTH, 128])
  /root/reference/MV3D_TF_release/lib/networks/group_pointcloud.py:88-105   build_input

`voxel_scatter` is that scatter_nd (and, through autograd, its gradient -- a gather of the grid gradient at the
coordinates): the [K,128] voxel-wise features of the non-empty voxels land in a dense, zero-filled voxel grid.  It is
the same "few rows into a big zero map" pattern as SHPL's pooled map, and runs on the same kernels: a plan with one
entry (cell(coordinate[k]), k) per feature row (shpl_plan_from_voxel_coords), then shpl_pool_forward with no dense
part.  Duplicate coordinates are summed in row order like TF-CPU's scatter_nd.  The dense layers in front of the
scatter (VFELayer, Dense, BatchNormalization, :17-81) are ordinary network layers and are not part of this package.

No CPU fallback: everything runs in libshpl.so.
"""
import ctypes

import numpy as np
import torch

from . import _cabi, ops
from .ops import SparsePoolPlan, _ptr, _stream

_lib = _cabi.lib

# utils/config_voxels.py:50-64 (cfg.DETECT_OBJ = 'Pedestrian' / 'Cyclist'); the Car grid is 400 x 352 (:33-48)
INPUT_HEIGHT = 200
INPUT_WIDTH = 240
GRID_DEPTH = 10

# TF-CPU's scatter_nd raises InvalidArgumentError for a coordinate outside the grid, TF-GPU drops the row.
# With STRICT_INDEX_CHECK voxel_scatter reads the plan counters back and raises ValueError; without it such
# rows contribute nothing and nothing is read back.
STRICT_INDEX_CHECK = True


def build_input(voxel_dict_list, fix_coordinate_columns=False):
    """group_pointcloud.py:88-105: concatenate the per-sample buffers and put the sample number in front of each
    coordinate row with np.pad -- ALWAYS a new first column, exactly like the reference (default).

    The reference's pad was written for [K,3] (d,h,w) buffers, but this fork's feeder emits [K,4] = (0, d, h, w)
    (construct_voxel.py:128, :147), for which the pad gives FIVE columns that tf.scatter_nd into the rank-5 grid then
    reads as (batch, 0, d, h, w) -- a reference inconsistency (reproduce, do not fix silently).
    fix_coordinate_columns=True is the explicit opt-in repair: for 4-column buffers the sample number overwrites the
    leading zero instead, giving the (batch, d, h, w) rows voxel_scatter takes.
    numpy in -> numpy out, CUDA tensors in -> CUDA tensors out."""
    batch_size = len(voxel_dict_list)
    features, numbers, coords = [], [], []
    for i, vd in enumerate(voxel_dict_list):
        features.append(vd['feature_buffer'])
        numbers.append(vd['number_buffer'])
        c = vd['coordinate_buffer']
        if isinstance(c, torch.Tensor):
            if fix_coordinate_columns and c.shape[1] == 4:
                c = c.clone()
                c[:, 0] = i
            else:
                c = torch.nn.functional.pad(c, (1, 0), value=i)
        else:
            c = np.asarray(c)
            if fix_coordinate_columns and c.shape[1] == 4:
                c = c.copy()
                c[:, 0] = i
            else:
                c = np.pad(c, ((0, 0), (1, 0)), mode='constant', constant_values=i)
        coords.append(c)
    if isinstance(features[0], torch.Tensor):
        return batch_size, torch.cat(features), torch.cat(numbers), torch.cat(coords)
    return batch_size, np.concatenate(features), np.concatenate(numbers), np.concatenate(coords)


def voxel_scatter_plan(coordinate, batch_size, grid=(GRID_DEPTH, INPUT_HEIGHT, INPUT_WIDTH), k_dev=None, read_counts=None):
    """The plan of one scatter: coordinate [K,4] int32 / int64 CUDA tensor (numpy is copied), rows (batch, d, h, w).
    k_dev: optional device int32 tensor -- only the first min(K, k_dev[0]) rows exist (the voxel count the feeder left
    on the device).  read_counts (default: STRICT_INDEX_CHECK) synchronises and fills plan.nnz / plan.n_oob."""
    if not isinstance(coordinate, torch.Tensor):
        if not torch.cuda.is_available():
            raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        coordinate = torch.from_numpy(np.ascontiguousarray(np.asarray(coordinate), dtype=np.int64)).cuda()
    ops.require_cuda(coordinate, "coordinate")
    if coordinate.dim() != 2 or coordinate.shape[1] != 4:
        raise ValueError("coordinate must have shape (K, 4) = (batch, d, h, w), got %s" % (tuple(coordinate.shape),))
    if coordinate.dtype not in (torch.int32, torch.int64):
        coordinate = coordinate.to(torch.int64)
    coordinate = coordinate.contiguous()
    dev = coordinate.device
    K = int(coordinate.shape[0])
    B, (D, H, W) = int(batch_size), (int(g) for g in grid)
    plan = SparsePoolPlan(B * D * H * W, (max(K, 1), 1), K, dev, zero_meta=False)
    plan.entry_bound = max(K, 1)
    plan.grid = (B, D, H, W)
    plan.n_coords = K
    ws = ops.workspace(dev, K)
    st = plan.frame_struct(0)
    rc = _lib.shpl_plan_from_voxel_coords(_ptr(coordinate), int(coordinate.dtype == torch.int64), K, _ptr(k_dev), B, D, H, W,
                                          ctypes.byref(st), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "shpl_plan_from_voxel_coords")
    plan._keep = (coordinate, k_dev)
    if STRICT_INDEX_CHECK if read_counts is None else read_counts:
        plan.read_counts()
    return plan


class VoxelScatterFunction(torch.autograd.Function):
    """tf.scatter_nd(coordinate, voxelwise, grid shape) and its registered gradient (gather_nd of the grid gradient);
    no gradient flows to the coordinates."""

    @staticmethod
    def forward(ctx, voxelwise, plan):
        ops.require_cuda(voxelwise, "voxelwise")
        if voxelwise.dtype != torch.float32:
            raise ValueError("voxelwise must be float32 (the reference's dtype)")
        if voxelwise.dim() != 2 or voxelwise.shape[0] != plan.n_coords:
            raise ValueError("voxelwise %s does not match the %d coordinates of the plan" % (tuple(voxelwise.shape), plan.n_coords))
        if voxelwise.shape[1] % 4:
            raise ValueError("the feature width must be a multiple of 4 (128-bit vector access), got %d" % voxelwise.shape[1])
        x = voxelwise.contiguous()
        ctx.plan, ctx.K = plan, x.shape[0]
        if x.shape[0] == 0:       # nothing to scatter: the kernels still write the zeros
            x = torch.zeros((1, x.shape[1]), dtype=torch.float32, device=x.device)
        out = ops.pool_forward(None, x, plan.by_row(), plan.n_rows, plan.n_src)
        return out.view(*plan.grid, x.shape[1])

    @staticmethod
    def backward(ctx, g_out):
        plan = ctx.plan
        C = g_out.shape[-1]
        g = g_out.contiguous().view(plan.n_rows, C)
        _, g_src = ops.pool_backward(g, plan.by_pixel(), plan.n_rows, 0, plan.n_src, C, want_dst=False)
        return g_src[:ctx.K], None


def voxel_scatter(coordinate, voxelwise, batch_size, grid=(GRID_DEPTH, INPUT_HEIGHT, INPUT_WIDTH), k_dev=None, plan=None):
    """group_pointcloud.py:84-85.  coordinate [K,4] (batch, d, h, w); voxelwise [K,C] float32 CUDA tensor, C a
    multiple of 4 (128 in the reference) -> [batch_size, *grid, C] float32, zero where no voxel lands.
    A plan from voxel_scatter_plan can be passed to reuse it (e.g. for several feature tensors of one sample)."""
    if plan is None:
        plan = voxel_scatter_plan(coordinate, batch_size, grid, k_dev=k_dev)
    if STRICT_INDEX_CHECK and plan.n_oob is not None and sum(plan.n_oob) > 0:
        raise ValueError("voxel_scatter: %d coordinates outside the %s grid (TF-CPU raises InvalidArgumentError here); set "
                         "group_pointcloud.STRICT_INDEX_CHECK = False to drop them" % (sum(plan.n_oob), plan.grid))
    return VoxelScatterFunction.apply(voxelwise, plan)
