"""Augmentation hooks that touch the arrays SHPL's correspondence builder reads, running on B200 CUDA kernels.

Mirrors (same names / argument order where the reference has a function of its own):
  /root/reference/avod/avod/datasets/kitti/kitti_aug.py:24-29, :85-97, :100-118
        flip_point_cloud, flip_ground_plane, flip_stereo_calib_p2      (called from kitti_dataset.py:304-311)
  /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:191-209   augment_fv
  /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:165-185   the point-cloud half of
        augment_voxel: img_index2 from the un-augmented points, then shift / expansion / rotation

Two reference quirks are kept, not fixed:
  * avod flips the point cloud and a COPY of P2 (kitti_dataset.py:304-311) but builds the correspondences from the
    unflipped stereo_calib.p2 (:376): gen_sparse_pooling_input_avod here reads whatever `stereo_calib.p2` holds, like
    the reference; pass a calib object carrying flip_stereo_calib_p2's matrix to get correspondences that match the
    flipped image.
  * augment_fv moves img_index without clipping it to the image: entries can leave the padded image; the builder
    counts them (plan counters) and the layer's STRICT_INDEX_CHECK decides what happens.

The point arrays go through libshpl.so (numpy in -> numpy out, CUDA tensors are transformed in place and returned);
the 12 / 4-number matrices are host arithmetic like in the reference.  No CPU fallback for the point arrays.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .ops import _ptr, _stream

_lib = _cabi.lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("sparse_pooling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------ avod: flipping
def flip_point_cloud(point_cloud, n_dev=None):
    """kitti_aug.py:24-29.  point_cloud [3,N] f64: numpy -> a flipped copy (like the reference); a CUDA tensor is
    copied too unless it is passed through flip_point_cloud_(...)."""
    if isinstance(point_cloud, torch.Tensor):
        return flip_point_cloud_(point_cloud.clone(), n_dev=n_dev)
    pc = torch.from_numpy(np.array(point_cloud, dtype=np.float64, order="C")).to(_device())
    return flip_point_cloud_(pc).cpu().numpy()


def flip_point_cloud_(point_cloud, n_dev=None):
    """In-place form for device-resident chains (e.g. on lidar_ingest.lidar_to_cam_raw's [3,cap] output with its
    device-side count)."""
    if not (isinstance(point_cloud, torch.Tensor) and point_cloud.is_cuda and point_cloud.dtype == torch.float64):
        raise RuntimeError("flip_point_cloud_: a float64 CUDA tensor is required (no CPU fallback)")
    if point_cloud.dim() != 2 or point_cloud.shape[0] != 3:
        raise ValueError("flip_point_cloud_: expected a [3,N] tensor, got %s" % (tuple(point_cloud.shape),))
    if point_cloud.shape[1] == 0:
        return point_cloud
    # row 0 through its stride: a [3,N] array as well as the transposed view of an [N,3] one
    rc = _lib.shpl_flip_point_cloud(_ptr(point_cloud), int(point_cloud.stride(1)), int(point_cloud.shape[1]), _ptr(n_dev), _stream())
    _cabi.check(rc, "shpl_flip_point_cloud")
    return point_cloud


def flip_ground_plane(ground_plane):
    """kitti_aug.py:85-97: the x coefficient of ax + by + cz + d = 0 changes sign."""
    flipped = np.copy(ground_plane)
    flipped[0] = -ground_plane[0]
    return flipped


def flip_stereo_calib_p2(calib_p2, image_shape):
    """kitti_aug.py:100-118: x0 mirrored about the image width (image_shape = (h, w)), t1 negated."""
    flipped = np.copy(calib_p2)
    flipped[0, 2] = image_shape[1] - calib_p2[0, 2]
    flipped[0, 3] = -calib_p2[0, 3]
    return flipped


# ------------------------------------------------------------------------------------ MV3D
def project_and_augment_points(lidar_pc, P, sx=0.0, sz=0.0, expansion_ratio=1.0, rotation_angle=None, n_dev=None):
    """minibatch_mv3d_img.py:165-185.  lidar_pc [n,4] f64 camera-frame points, P [3,4].
    img_index2 = np.round(projectToImage(lidar_pc[:, 0:3].T, P)).astype(int) from the points as given; then, when
    rotation_angle is not None, augment_voxel's transforms: x += sx, z += sz, xyz *= expansion_ratio, (x, z) rotated by
    rot_mat = [[cos, sin], [-sin, cos]].  Returns (lidar_pc, img_index2 [2,n] int64): numpy in -> numpy out (the input
    array is modified in place like the reference's); a CUDA tensor is transformed in place."""
    as_numpy = not isinstance(lidar_pc, torch.Tensor)
    if as_numpy:
        pc = torch.from_numpy(np.ascontiguousarray(lidar_pc, dtype=np.float64)).to(_device())
    else:
        pc = lidar_pc
        if not (pc.is_cuda and pc.dtype == torch.float64 and pc.is_contiguous()):
            raise RuntimeError("lidar_pc must be a contiguous float64 CUDA tensor (or a numpy array): no CPU fallback")
    if pc.dim() != 2 or pc.shape[1] != 4:
        raise ValueError("lidar_pc must have shape (n, 4), got %s" % (tuple(pc.shape),))
    n = int(pc.shape[0])
    Ph = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(12))
    img2 = torch.empty((2, max(n, 1)), dtype=torch.int64, device=pc.device)
    augment = rotation_angle is not None
    rot = None
    if augment:
        a = np.asarray(rotation_angle, dtype=np.float64).reshape(-1)[:1]
        rot = np.ascontiguousarray(np.array([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]]).reshape(4))     # :148
    ratio = float(np.asarray(expansion_ratio, dtype=np.float64).reshape(-1)[0])
    rc = _lib.shpl_mv3d_project_augment(_ptr(pc), n, _ptr(n_dev), Ph.ctypes.data_as(ctypes.c_void_p), int(augment), float(sx),
                                        float(sz), ratio, None if rot is None else rot.ctypes.data_as(ctypes.c_void_p),
                                        _ptr(img2), _stream())
    _cabi.check(rc, "shpl_mv3d_project_augment")
    img2 = img2[:, :n]
    if as_numpy:
        out = pc.cpu().numpy()
        if isinstance(lidar_pc, np.ndarray) and lidar_pc.dtype == np.float64 and lidar_pc.shape == out.shape:
            lidar_pc[...] = out          # the reference mutates the array it was given (:176-181)
            out = lidar_pc
        return out, img2.cpu().numpy()
    return pc, img2


def augment_fv_index(img_index, sx, sy, expansion_ratio, n_dev=None):
    """minibatch_mv3d_img.py:205-206 on img_index [3,n] (int64), in place:
    row 0 = (row 0 * expansion_ratio + sx).astype(int), row 1 with sy."""
    ratio = float(np.asarray(expansion_ratio, dtype=np.float64).reshape(-1)[0])
    if isinstance(img_index, torch.Tensor):
        if not (img_index.is_cuda and img_index.dtype == torch.int64 and img_index.dim() == 2 and img_index.stride(1) == 1):
            raise RuntimeError("img_index must be an int64 CUDA tensor [3,n] with contiguous rows (or a numpy array)")
        rc = _lib.shpl_augment_fv_index(_ptr(img_index), int(img_index.stride(0)), int(img_index.shape[1]), _ptr(n_dev), ratio,
                                        float(sx), float(sy), _stream())
        _cabi.check(rc, "shpl_augment_fv_index")
        return img_index
    dev_idx = torch.from_numpy(np.ascontiguousarray(img_index, dtype=np.int64)).to(_device())
    augment_fv_index(dev_idx, sx, sy, ratio)
    img_index[0:2, :] = dev_idx[0:2].cpu().numpy()
    return img_index


def augment_fv(blobs, scale=10):
    """minibatch_mv3d_img.py:191-209.  Draws (sx, sy) and expansion_ratio from np.random in the reference's order,
    resizes / shifts / crops blobs['image_data'] on the host (OpenCV, as the reference does; skipped when the blob
    holds no image), moves blobs['gt_boxes'] and -- on the GPU -- blobs['img_index'].
    Returns (blobs, [sx, sy], expansion_ratio)."""
    sx, sy = np.random.uniform(0, scale, 2)
    expansion_ratio = np.random.uniform(0.95, 1.05, 1)
    if blobs.get('image_data') is not None:
        import cv2
        from .config import PAD_IMAGE_TO
        img = cv2.resize(blobs['image_data'], None, fx=float(expansion_ratio[0]), fy=float(expansion_ratio[0]))
        rows, cols = img.shape[0:2]
        img = cv2.warpAffine(img, np.float32([[1, 0, sx], [0, 1, sy]]), (cols, rows))
        blobs['image_data'] = img[0:PAD_IMAGE_TO[1], 0:PAD_IMAGE_TO[0]]            # clip to the padded size (:200-203)
    if blobs.get('gt_boxes') is not None:
        blobs['gt_boxes'][:, [0, 2]] = blobs['gt_boxes'][:, [0, 2]] * expansion_ratio + sx
        blobs['gt_boxes'][:, [1, 3]] = blobs['gt_boxes'][:, [1, 3]] * expansion_ratio + sy
    blobs['img_index'] = augment_fv_index(blobs['img_index'], sx, sy, expansion_ratio)
    return blobs, [sx, sy], expansion_ratio
