"""Drop-in mirror of the reference's point-cloud ingest, running on B200 CUDA kernels.

Same function names and argument order as
  obj_utils.get_lidar_point_cloud   /root/reference/avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:220-268
  calib_utils.read_calibration      /root/reference/avod/wavedata/wavedata/tools/core/calib_utils.py:55-112
  calib_utils.read_lidar            /root/reference/avod/wavedata/wavedata/tools/core/calib_utils.py:328-368
called per sample from KittiUtils.get_point_cloud (avod/avod/datasets/kitti/kitti_utils.py:136-163).  The two file
reads stay on the host (numpy); the transform into the camera frame, the projection and the field-of-view filter
run in shpl_lidar_to_cam (include/shpl.h).  Output: the (3, M) float64 camera-frame cloud that
BevSlices.generate_bev takes.  No CPU fallback for the arithmetic.
"""
import csv
import ctypes
import os

import numpy as np
import torch

from . import _cabi
from .ops import _ptr, _stream

_lib = _cabi.lib


class FrameCalibrationData:
    """calib_utils.py:8-52: p0..p3 (3x4), r0_rect (3x3), tr_velodyne_to_cam (3x4)."""

    def __init__(self):
        self.p0 = self.p1 = self.p2 = self.p3 = []
        self.r0_rect = []
        self.tr_velodyne_to_cam = []


def read_calibration(calib_dir, img_idx):
    """calib_utils.py:55-112 (host: a six-line text file)."""
    cal = FrameCalibrationData()
    with open(calib_dir + "/%06d.txt" % img_idx, 'r') as f:
        data = [row for row in csv.reader(f, delimiter=' ')]
    p_all = [np.reshape([float(v) for v in data[i][1:]], (3, 4)) for i in range(4)]
    cal.p0, cal.p1, cal.p2, cal.p3 = p_all
    cal.r0_rect = np.reshape([float(v) for v in data[4][1:]], (3, 3))
    cal.tr_velodyne_to_cam = np.reshape([float(v) for v in data[5][1:]], (3, 4))
    return cal


def read_lidar(velo_dir, img_idx):
    """calib_utils.py:328-368: (x, y, z, i) float32 arrays of the .bin file, or [] when it does not exist."""
    path = velo_dir + "/%06d.bin" % img_idx
    if not os.path.exists(path):
        return []
    with open(path, 'rb') as fid:
        data = np.fromfile(fid, np.single)
    xyzi = data.reshape(-1, 4)
    return xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]


def rectified_matrix(frame_calib):
    """R0_rect padded to 4x4 times Tr_velo_to_cam padded to 4x4 (calib_utils.py:389-406), as numpy computes it."""
    r0 = np.pad(np.asarray(frame_calib.r0_rect, dtype=np.float64), ((0, 1), (0, 1)), 'constant', constant_values=0)
    r0[3, 3] = 1
    tf = np.pad(np.asarray(frame_calib.tr_velodyne_to_cam, dtype=np.float64), ((0, 1), (0, 0)), 'constant', constant_values=0)
    tf[3, 3] = 1
    return np.dot(r0, tf)


def _workspace(device, n):
    """scratch of shpl_lidar_to_cam for the current stream (ops.scratch: never shared between streams / threads)"""
    from . import ops
    return ops.scratch("lidar", device, int(_lib.shpl_lidar_workspace_bytes(int(n))))


def lidar_to_cam_raw(velo, n, frame_calib, im_size, out, counts, min_intensity=None, stream=None, ws=None):
    """One asynchronous shpl_lidar_to_cam call: velo f32 [N,4] CUDA (contiguous), out f64 [3,cap] CUDA, counts i32 [4].
    ws: scratch of shpl_lidar_workspace_bytes(n) bytes; by default the one cached for the current stream."""
    if not velo.is_contiguous() or velo.dtype != torch.float32:
        raise ValueError("lidar_to_cam_raw: the scan must be a contiguous float32 [N,4] CUDA tensor")
    R = np.ascontiguousarray(rectified_matrix(frame_calib)[0:3].reshape(12))
    P = np.ascontiguousarray(np.asarray(frame_calib.p2, dtype=np.float64).reshape(12))
    w, h = (int(im_size[0]), int(im_size[1])) if im_size else (0, 0)
    if ws is None:
        ws = _workspace(velo.device, n)
    rc = _lib.shpl_lidar_to_cam(_ptr(velo), int(n), R.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p), w, h,
                                1 if min_intensity else 0, float(min_intensity or 0.0), _ptr(out), int(out.shape[1]),
                                _ptr(counts), _ptr(ws), ws.numel(), _stream() if stream is None else stream)
    _cabi.check(rc, "shpl_lidar_to_cam")


def lidar_to_cam_fov(velo_xyzi, frame_calib, im_size=None, min_intensity=None):
    """The arithmetic of get_lidar_point_cloud after the file reads: velo_xyzi float32 [N,4] (numpy or CUDA tensor).
    Returns the (3, M) float64 cloud (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out)."""
    as_numpy = not isinstance(velo_xyzi, torch.Tensor)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        velo = torch.from_numpy(np.ascontiguousarray(np.asarray(velo_xyzi, dtype=np.float32).reshape(-1, 4))).to(dev)
    else:
        if not velo_xyzi.is_cuda:
            raise RuntimeError("the scan must be a CUDA tensor (or a numpy array): there is no CPU fallback")
        dev = velo_xyzi.device
        velo = velo_xyzi.to(torch.float32).reshape(-1, 4).contiguous()
    n = int(velo.shape[0])
    out = torch.empty((3, max(n, 1)), dtype=torch.float64, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    lidar_to_cam_raw(velo, n, frame_calib, im_size, out, counts, min_intensity=min_intensity)
    c = counts.cpu()
    if im_size and min_intensity and int(c[1]) != n:
        # obj_utils.py:266: the intensity mask has N entries, the image mask only those with z > 0
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % (int(c[1]), n))
    m = int(c[0])
    pc = out[:, :m].contiguous()
    return pc.cpu().numpy() if as_numpy else pc


def get_lidar_point_cloud(img_idx, calib_dir, velo_dir, im_size=None, min_intensity=None):
    """obj_utils.py:220-268: the frame's LIDAR cloud in the camera frame, optionally cut to the image's field of view."""
    frame_calib = read_calibration(calib_dir, img_idx)
    x, y, z, i = read_lidar(velo_dir=velo_dir, img_idx=img_idx)
    velo = np.ascontiguousarray(np.stack((x, y, z, i), axis=1))
    return lidar_to_cam_fov(velo, frame_calib, im_size=im_size, min_intensity=min_intensity)
