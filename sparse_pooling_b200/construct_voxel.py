"""Drop-in mirror of the reference's MV3D voxel feeder, running on B200 CUDA kernels.

Same function name, argument order and return tuple as
  /root/reference/MV3D_TF_release/lib/utils/construct_voxel.py:37-162   (point_cloud_2_top_sparse)
called per sample from MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:93.  Its outputs
img_index / bv_index / M_val go straight into produce_sparse_pooling_input
(MV3D_TF_release/lib/fast_rcnn/train_mv_voxel.py:325-326) -- M_val = 1/count is SHPL's
"non-homogeneous" weight.

numpy in -> numpy out; torch CUDA tensors in -> torch CUDA tensors out.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .ops import _ptr, _stream

_lib = _cabi.lib

# construct_voxel.py:11-17 with the Pedestrian/Cyclist values of utils/config_voxels.py:50-59
# (cfg.DETECT_OBJ = 'Pedestrian'); the Car set is config_voxels.py:33-43.
side_range = (-20, 20 - 0.01)
fwd_range = (0, 48 - 0.01)
height_range = (-1, 3 - 0.01)
res = 0.2
zres = 0.4
NUM_VOXEL_FEATURES = 7
MAX_NUM_POINTS = 45

def _workspace(device, n):
    """scratch of shpl_mv3d_voxelize for the current stream (ops.scratch: never shared between streams / threads)"""
    from . import ops
    return ops.scratch("mv3d", device, int(_lib.shpl_mv3d_workspace_bytes(int(n))))


def voxelize_raw(points, img_index2, n, res, zres, side_range, fwd_range, height_range, max_points, out, stream=None):
    """One asynchronous shpl_mv3d_voxelize call into the preallocated buffers of `out` (dict with img_index
    [3,cap] i64, bv_index [cap,2] i64, m_val [cap] f64, counts [8] i32, ws, and optionally feature /
    coordinate / number); nothing is read back.  Returns voxel_full_size (host ints)."""
    if not (points.is_contiguous() and img_index2.is_contiguous()):
        raise ValueError("voxelize_raw: points [n,4] and img_index2 [2,n] must be contiguous")
    ranges = np.array([side_range[0], side_range[1], fwd_range[0], fwd_range[1], height_range[0], height_range[1]],
                      dtype=np.float64)
    vfs = (ctypes.c_int32 * 3)()
    feat = out.get("feature")
    rc = _lib.shpl_mv3d_voxelize(_ptr(points), _ptr(img_index2), int(n), float(res), float(zres),
                                 ranges.ctypes.data_as(ctypes.c_void_p), int(max_points), vfs,
                                 _ptr(out["img_index"]), _ptr(out["bv_index"]), _ptr(out["m_val"]), int(out["m_val"].numel()),
                                 _ptr(feat), _ptr(out.get("coordinate")), _ptr(out.get("number")),
                                 0 if feat is None else int(feat.shape[0]),
                                 _ptr(out["counts"]), _ptr(out["ws"]), out["ws"].numel(),
                                 _stream() if stream is None else stream)
    _cabi.check(rc, "shpl_mv3d_voxelize")
    return np.array([vfs[0], vfs[1], vfs[2]])


def point_cloud_2_top_sparse(points, res=res, zres=zres, side_range=side_range, fwd_range=fwd_range,
                             height_range=height_range, top_count=None, to_camera_frame=False, points_in_cam=False,
                             calib=None, img_size=[0, 0], augmentation=False, img_index2=None):
    """construct_voxel.py:37-162.  points [n,4] camera frame (x, y, z, reflectance) with points_in_cam=True,
    img_index2 int [2,n].  Returns (voxel_dict, voxel_full_size, img_index [3,m], bv_index [m,2], M_val [m])."""
    if to_camera_frame:
        # construct_voxel.py:56 uses an undefined `P` on this branch: the reference raises NameError
        raise NameError("name 'P' is not defined")
    assert points_in_cam, 'Wrong, cannot process LIDAR coordinate points'      # :69
    as_numpy = not isinstance(points, torch.Tensor)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        pts = torch.from_numpy(np.ascontiguousarray(np.asarray(points, dtype=np.float64))).to(dev)
        img2 = torch.from_numpy(np.ascontiguousarray(np.asarray(img_index2), dtype=np.int64)).to(dev)
    else:
        if not points.is_cuda:
            raise RuntimeError("points must be a CUDA tensor (or a numpy array): there is no CPU fallback")
        dev = points.device
        pts = points.to(torch.float64).contiguous()
        img2 = (img_index2 if isinstance(img_index2, torch.Tensor) else torch.from_numpy(np.asarray(img_index2))).to(
            device=dev, dtype=torch.int64).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 4:
        raise ValueError("points must have shape (n, 4), got %s" % (tuple(pts.shape),))
    n = int(pts.shape[0])
    if img2.dim() != 2 or img2.shape[0] != 2 or img2.shape[1] != n:
        raise IndexError("img_index2 must have shape (2, %d), got %s" % (n, tuple(img2.shape)))
    T = int(MAX_NUM_POINTS)
    cap = max(n, 1)
    out = dict(img_index=torch.empty((3, cap), dtype=torch.int64, device=dev),
               bv_index=torch.empty((cap, 2), dtype=torch.int64, device=dev),
               m_val=torch.empty(cap, dtype=torch.float64, device=dev),
               feature=torch.empty((cap, T, NUM_VOXEL_FEATURES), dtype=torch.float64, device=dev),
               coordinate=torch.empty((cap, 4), dtype=torch.int64, device=dev),
               number=torch.empty(cap, dtype=torch.int64, device=dev),
               counts=torch.zeros(8, dtype=torch.int32, device=dev), ws=_workspace(dev, n))
    vfs = voxelize_raw(pts, img2, n, res, zres, side_range, fwd_range, height_range, T, out)
    c = out["counts"].cpu()
    if int(c[3]) != 0:
        raise RuntimeError("shpl_mv3d_voxelize: error bits %d" % int(c[3]))
    m, V = int(c[1]), int(c[2])
    feature, coordinate, number = out["feature"][:V], out["coordinate"][:V], out["number"][:V]
    img_index, bv_index, m_val = out["img_index"][:, :m].contiguous(), out["bv_index"][:m], out["m_val"][:m]
    if as_numpy:
        feature, coordinate, number = feature.cpu().numpy(), coordinate.cpu().numpy(), number.cpu().numpy()
        img_index, bv_index, m_val = img_index.cpu().numpy(), bv_index.cpu().numpy(), m_val.cpu().numpy()
    voxel_dict = {'feature_buffer': feature, 'coordinate_buffer': coordinate, 'number_buffer': number}
    return voxel_dict, vfs, img_index, bv_index, m_val
