"""Drop-in mirror of the reference's BEV slicing feeder, running on B200 CUDA kernels.

Same class name, constructor and method signature as
  /root/reference/avod/avod/core/bev_generators/bev_slices.py:8-156   (BevSlices.generate_bev)
with VoxelGrid2D.voxelize_2d (avod/wavedata/wavedata/tools/core/voxel_grid_2d.py:43-162) and
KittiUtils.create_slice_filter (avod/avod/datasets/kitti/kitti_utils.py:79-107) inlined in the
kernels (shpl_bev_slices, include/shpl.h).  The call site is KittiUtils.create_bev_maps
(kitti_utils.py:109-127), reached from kitti_dataset.py:356-357; its outputs voxel_indices /
unique_pts feed gen_sparse_pooling_input_avod (kitti_dataset.py:376-377).

numpy in -> numpy out; torch CUDA tensor in -> torch CUDA tensors out.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .ops import _ptr, _stream

_lib = _cabi.lib

ERR_FIRST_SLICE_EMPTY, ERR_EXTENTS, ERR_CAPACITY, ERR_NO_POINTS = 1, 2, 4, 8
N_COUNTS = 32


def _f64_array(x, n):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if a.shape[0] != n:
        raise ValueError("expected %d values, got %d" % (n, a.shape[0]))
    return a


def density_lut(norm_value):
    """min(1, log(n+1)/norm) for n = 0, 1, ... until it saturates at 1.0 (bev_generator.py:35-36), tabulated with
    numpy -- the reference's own log -- so the density map is bit-identical; beyond the table the value is 1.0."""
    n_sat = int(np.ceil(np.exp(norm_value))) + 2
    n = np.arange(n_sat)
    return np.minimum(1.0, np.log(n + 1) / norm_value)


class BevWorkspace:
    """Scratch + outputs of shpl_bev_slices for one grid geometry, allocated once and reused."""

    def __init__(self, extents, voxel_size, num_slices, capacity, device, with_maps=True):
        ext = _f64_array(np.asarray(extents, dtype=np.float64), 6)
        nx, nz = ctypes.c_int32(0), ctypes.c_int32(0)
        _cabi.check(_lib.shpl_bev_grid_dims(ext.ctypes.data_as(ctypes.c_void_p), float(voxel_size), ctypes.byref(nx),
                                            ctypes.byref(nz)), "shpl_bev_grid_dims")
        self.nx, self.nz, self.num_slices = nx.value, nz.value, int(num_slices)
        need = int(_lib.shpl_bev_workspace_bytes(ext.ctypes.data_as(ctypes.c_void_p), float(voxel_size), int(num_slices)))
        if need == 0:
            raise ValueError("shpl_bev_workspace_bytes: %s" % _lib.shpl_last_error().decode())
        self.device = device
        self.capacity = int(max(capacity, 1))
        self.ws = torch.empty(need, dtype=torch.uint8, device=device)
        self.counts = torch.zeros(N_COUNTS, dtype=torch.int32, device=device)
        self.voxel_indices = torch.empty((self.capacity, 2), dtype=torch.int64, device=device)
        self.unique_pts = torch.empty((self.capacity, 3), dtype=torch.float64, device=device)
        self.maps = torch.empty((self.num_slices + 1, self.nz, self.nx), dtype=torch.float64, device=device) if with_maps else None


def bev_slices_raw(points, coord_stride, point_stride, P, ground_plane, extents, voxel_size, height_lo, height_hi,
                   num_slices, norm_value, work, lut=None, stream=None, p_dev=None):
    """One asynchronous shpl_bev_slices call into `work` (a BevWorkspace); nothing is read back.  p_dev: optional
    device pointer (ctypes.c_void_p) to the int32 number of valid points (P is then the capacity)."""
    gp = _f64_array(ground_plane, 4)
    ext = _f64_array(np.asarray(extents, dtype=np.float64), 6)
    rc = _lib.shpl_bev_slices(_ptr(points), int(coord_stride), int(point_stride), int(P), p_dev,
                              gp.ctypes.data_as(ctypes.c_void_p), ext.ctypes.data_as(ctypes.c_void_p), float(voxel_size),
                              float(height_lo), float(height_hi), int(num_slices), float(norm_value),
                              _ptr(lut), 0 if lut is None else int(lut.numel()),
                              _ptr(work.voxel_indices), _ptr(work.unique_pts), work.capacity, _ptr(work.maps),
                              _ptr(work.counts), _ptr(work.ws), work.ws.numel(), _stream() if stream is None else stream)
    _cabi.check(rc, "shpl_bev_slices")


class BevSlices:
    """bev_slices.py:8-31: BEV maps created using slices of the point cloud."""

    NORM_VALUES = {
        'lidar': np.log(16),
    }

    def __init__(self, config, kitti_utils=None):
        """config: object with height_lo, height_hi, num_slices (the bev_generator protobuf config);
        kitti_utils: accepted for signature compatibility -- its create_slice_filter is inlined in the kernel."""
        self.height_lo = config.height_lo
        self.height_hi = config.height_hi
        self.num_slices = config.num_slices
        self.kitti_utils = kitti_utils
        self.height_per_division = (self.height_hi - self.height_lo) / self.num_slices
        self._work = {}
        self._luts = {}

    def _lut(self, source, device):
        key = (source, str(device))
        if key not in self._luts:
            self._luts[key] = torch.from_numpy(density_lut(self.NORM_VALUES[source])).to(device)
        return self._luts[key]

    def generate_bev(self, source, point_cloud, ground_plane, area_extents, voxel_size, output_indices=False):
        """bev_slices.py:33-156.  point_cloud (3, N) camera frame; returns the bev_maps dict
        {'height_maps': [num_slices maps (nz, nx)], 'density_map': (nz, nx)} and, with output_indices,
        (bev_maps, voxel_indices [n,2], unique_pts [n,3])."""
        as_numpy = not isinstance(point_cloud, torch.Tensor)
        if as_numpy:
            if not torch.cuda.is_available():
                raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
            pc = torch.from_numpy(np.ascontiguousarray(np.asarray(point_cloud, dtype=np.float64))).to(dev)
        else:
            if not point_cloud.is_cuda:
                raise RuntimeError("point_cloud must be a CUDA tensor (or a numpy array): there is no CPU fallback")
            dev = point_cloud.device
            pc = point_cloud.to(torch.float64)
        if pc.dim() != 2 or pc.shape[0] != 3:
            raise ValueError("point_cloud must have shape (3, N), got %s" % (tuple(pc.shape),))
        P = int(pc.shape[1])
        key = (tuple(np.asarray(area_extents, dtype=np.float64).reshape(-1)), float(voxel_size), str(dev))
        work = self._work.get(key)
        cap = max(P, 1) * self.num_slices
        if work is None or work.capacity < min(cap, work.nx * work.nz * self.num_slices):
            work = BevWorkspace(area_extents, voxel_size, self.num_slices, cap, dev)
            work.capacity = min(work.capacity, work.nx * work.nz * self.num_slices)
            self._work[key] = work
        bev_slices_raw(pc, pc.stride(0), pc.stride(1), P, ground_plane, area_extents, voxel_size, self.height_lo,
                       self.height_hi, self.num_slices, self.NORM_VALUES[source], work, lut=self._lut(source, dev))
        counts = work.counts.cpu()
        flags = int(counts[1])
        if flags & ERR_FIRST_SLICE_EMPTY:
            raise NameError("name 'voxel_grid_2d' is not defined")         # bev_slices.py:93 with an empty first slice
        if flags & ERR_EXTENTS:
            raise ValueError("Extents are smaller than the voxel coordinates")   # voxel_grid_2d.py:133-138
        if flags & ERR_NO_POINTS:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        if flags & ERR_CAPACITY:
            raise RuntimeError("shpl_bev_slices: output capacity exceeded")
        n = int(counts[0])
        maps = work.maps.clone()
        vox, upts = work.voxel_indices[:n].clone(), work.unique_pts[:n].clone()
        if as_numpy:
            maps, vox, upts = maps.cpu().numpy(), vox.cpu().numpy(), upts.cpu().numpy()
        bev_maps = dict()
        bev_maps['height_maps'] = [maps[i] for i in range(self.num_slices)]
        bev_maps['density_map'] = maps[self.num_slices]
        if output_indices:
            return bev_maps, vox, upts
        return bev_maps
