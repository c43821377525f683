"""Feeder oracle: CPU restatement (numpy) of the avod BEV slicing / voxelisation step that
produces the SHPL builder's inputs (rows a1/a2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows

  * BevSlices.generate_bev(..., output_indices=True)
        /root/reference/avod/avod/core/bev_generators/bev_slices.py:33-156
  * BevGenerator._create_density_map
        /root/reference/avod/avod/core/bev_generators/bev_generator.py:23-41
  * VoxelGrid2D.voxelize_2d
        /root/reference/avod/wavedata/wavedata/tools/core/voxel_grid_2d.py:43-162
  * KittiUtils.create_slice_filter   /root/reference/avod/avod/datasets/kitti/kitti_utils.py:79-107
  * obj_utils.get_point_filter       /root/reference/avod/wavedata/wavedata/tools/obj_detection/obj_utils.py:444-491
  * geometry_utils.dist_to_plane     /root/reference/avod/wavedata/wavedata/tools/core/geometry_utils.py:25-40

Pinned against the reference's own output (tests/golden/bev_slices_seed*.npz, made by
oracle/gen_goldens.py, which imports those files where they lie).
"""
import numpy as np


def point_filter(point_cloud, extents, ground_plane, offset_dist):
    """obj_utils.py:444-491: strict extents, and below the plane shifted up by offset_dist."""
    pc = np.asarray(point_cloud)
    inside = ((pc[0] > extents[0][0]) & (pc[0] < extents[0][1]) & (pc[1] > extents[1][0]) & (pc[1] < extents[1][1])
              & (pc[2] > extents[2][0]) & (pc[2] < extents[2][1]))
    plane = np.array(ground_plane) + [0, 0, 0, -offset_dist]                     # :479
    hom = np.vstack([pc, np.ones(pc.shape[1])])
    return inside & (np.dot(plane, hom) < 0)                                     # :482-486


def slice_filter(point_cloud, extents, ground_plane, lo, hi):
    """kitti_utils.py:97-107: between the planes at lo and hi above the ground."""
    return np.logical_xor(point_filter(point_cloud, extents, ground_plane, hi),
                          point_filter(point_cloud, extents, ground_plane, lo))


def grid_geometry(extents, voxel_size):
    """voxel_grid_2d.py:119-149: min/max voxel coordinate and divisions from the extents."""
    ext = np.array(extents).transpose()
    lo = np.floor(ext[0] / voxel_size)
    hi = np.ceil((ext[1] / voxel_size) - 1)
    lo[1] = 0
    hi[1] = 0
    return lo, hi, ((hi - lo) + 1).astype(np.int32)


def voxelize_2d(pts, voxel_size, extents, ground_plane):
    """voxel_grid_2d.py:59-152.  Returns (voxel_indices int [m,3], unique_pts f64 [m,3],
    heights f64 [m], num_pts int [m], num_divisions int32 [3])."""
    disc = np.floor(pts / voxel_size).astype(np.int32)                           # :67
    order = np.lexsort((disc[:, 1], disc[:, 2], disc[:, 0]))                     # :70-73 x, then z, then y
    spts, sdisc = pts[order], disc[order]
    first = np.ones(len(spts), dtype=bool)
    if len(spts) > 1:
        first[1:] = (sdisc[1:, 0] != sdisc[:-1, 0]) | (sdisc[1:, 2] != sdisc[:-1, 2])   # :86-94 unique on (x, 0, z)
    starts = np.nonzero(first)[0]
    coords = sdisc[starts].copy()
    coords[:, 1] = 0
    unique_pts = spts[starts]                                                    # :97-98
    num_pts = np.diff(np.r_[starts, len(spts)])                                  # :101-104
    a, b, c, d = ground_plane
    heights = (a * unique_pts[:, 0] + b * unique_pts[:, 1] + c * unique_pts[:, 2] + d) / np.sqrt(a ** 2 + b ** 2 + c ** 2)
    lo, hi, ndiv = grid_geometry(extents, voxel_size)
    if not (lo <= np.amin(coords, axis=0)).all() or not (hi >= np.amax(coords, axis=0)).all():
        raise ValueError("Extents are smaller than the voxel coordinates")       # :133-138
    return (coords - lo).astype(int), unique_pts, heights, num_pts, ndiv         # :152


def generate_bev(point_cloud, ground_plane, area_extents, voxel_size, height_lo, height_hi, num_slices,
                 norm_value=np.log(16)):
    """bev_slices.py:33-156 with output_indices=True.  point_cloud f64 [3,P].
    Returns (height_maps list of [Z,X] f64, density_map [Z,X] f64, voxel_indices int [N,2], unique_pts f64 [N,3])."""
    all_points = np.transpose(point_cloud)
    hpd = (height_hi - height_lo) / num_slices                                   # :29-31
    height_maps, idx_stack, pts_stack = [], [], []
    grid = None
    for s in range(num_slices):
        lo = height_lo + s * hpd                                                 # :66-67
        hi = lo + hpd
        pts = all_points[slice_filter(point_cloud, area_extents, ground_plane, lo, hi)]
        if len(pts) > 1:                                                         # :79
            vi, upts, heights, _, ndiv = voxelize_2d(pts, voxel_size, area_extents, ground_plane)
            grid = dict(vi=vi[:, [0, 2]], upts=upts, heights=heights, ndiv=ndiv)
        elif grid is None:
            raise NameError("voxel_grid_2d")                                     # quirk A.4-7: first slice empty
        hm = np.zeros((grid["ndiv"][0], grid["ndiv"][2]))
        grid["heights"] = grid["heights"] - lo                                   # :100 (compounds when a grid is reused)
        hm[grid["vi"][:, 0], grid["vi"][:, 1]] = np.asarray(grid["heights"]) / hpd
        height_maps.append(hm)
        rot = np.copy(grid["vi"])
        rot[:, 1] = grid["ndiv"][2] - rot[:, 1]                                  # :106-108
        idx_stack.append(rot)
        pts_stack.append(grid["upts"])
    height_maps = [np.flip(h.transpose(), axis=0) for h in height_maps]          # :116-118
    dens_pts = all_points[slice_filter(point_cloud, area_extents, ground_plane, height_lo, height_hi)]
    vi, _, _, num_pts, ndiv = voxelize_2d(dens_pts, voxel_size, area_extents, ground_plane)
    dm = np.zeros((ndiv[0], ndiv[2]))
    dm[vi[:, 0], vi[:, 2]] = np.minimum(1.0, np.log(num_pts + 1) / norm_value)   # bev_generator.py:35-36
    dm = np.flip(dm.transpose(), axis=0)
    return height_maps, dm, np.vstack(idx_stack), np.vstack(pts_stack)


# ------------------------------------------------------------------ MV3D voxel feeder
def point_cloud_2_top_sparse(points, img_index2, res, zres, side_range, fwd_range, height_range, max_points):
    """CPU restatement of point_cloud_2_top_sparse(points, ..., points_in_cam=True, img_index2=...)
        /root/reference/MV3D_TF_release/lib/utils/construct_voxel.py:37-162
    points f64 [n,4] camera frame (x, y, z, reflectance); img_index2 int [2,n] (rounded pixel of every point).
    Returns (voxel_dict, voxel_full_size, img_index int [3,m], bv_index int [m,2], M_val f64 [m]) like the
    reference.  The per-voxel cap loop (:133-140) is restated with a stable sort instead of the Python loop.
    Pinned against the reference's own output (tests/golden/mv3d_*.npz)."""
    points = np.asarray(points, dtype=np.float64)[:, [2, 0, 1, 3]]                # :59 forward, side, height
    x_max = int((side_range[1] - side_range[0]) / res)                            # :81-84
    y_max = int((fwd_range[1] - fwd_range[0]) / res)
    z_max = int((height_range[1] - height_range[0]) / zres)
    voxel_full_size = np.array([z_max + 1, x_max + 1, y_max + 1])
    filt = ((points[:, 0] > fwd_range[0]) & (points[:, 0] < fwd_range[1]) & (points[:, 1] > side_range[0])
            & (points[:, 1] < side_range[1]) & (points[:, 2] > height_range[0]) & (points[:, 2] < height_range[1]))   # :89-96
    pf = points[filt]
    img2 = np.asarray(img_index2)[:, filt]
    xyz = np.zeros((pf.shape[0], 3), dtype=int)
    xyz[:, 0] = ((pf[:, 1] - side_range[0]) / res).astype(np.int32)                # :116-118
    xyz[:, 1] = ((pf[:, 0] - fwd_range[0]) / res).astype(np.int32)
    xyz[:, 2] = ((pf[:, 2] - height_range[0]) / zres).astype(np.int32)
    uniq, inv = np.unique(xyz, axis=0, return_inverse=True)                         # :122
    inv = np.asarray(inv).reshape(-1)
    V = uniq.shape[0]
    order = np.argsort(inv, kind="stable")                                          # voxel by voxel, input order inside
    sinv = inv[order]
    start = np.searchsorted(sinv, np.arange(V))
    slot_sorted = np.arange(inv.size) - start[sinv]
    slot = np.empty(inv.size, dtype=np.int64)
    slot[order] = slot_sorted
    keep = slot < max_points                                                        # :135-140
    filt_indices = np.nonzero(keep)[0]
    number = np.minimum(np.bincount(inv, minlength=V), max_points)
    top = np.zeros((V, max_points, 7))
    top[inv[filt_indices], slot[filt_indices], 0:4] = pf[filt_indices]
    top[:, :, 4:7] = top[:, :, 0:3] - np.expand_dims(np.sum(top[:, :, 0:3], axis=1) / number[:, None], 1)   # :143
    coord = np.zeros((V, 4), dtype=int)
    coord[:, 1:4] = uniq[:, [2, 0, 1]]                                              # :128
    voxel_dict = {'feature_buffer': top, 'coordinate_buffer': coord, 'number_buffer': number}
    img2 = img2[:, filt_indices]
    img_index = np.vstack((img2, np.zeros((1, img2.shape[1])).astype(int)))         # :156-158
    bv_index = xyz[filt_indices, :][:, [1, 0]]                                      # :159
    M_val = 1.0 / number[inv[filt_indices]]                                         # :160
    return voxel_dict, voxel_full_size, img_index, bv_index, M_val


# ------------------------------------------------------------------ point-cloud ingest
def lidar_to_cam_frame(xyz_lidar, r0_rect, tr_velodyne_to_cam):
    """calib_utils.py:371-410: p_cam = R0_rect(4x4) . Tr_velo_to_cam(4x4) . [x y z 1]."""
    r0 = np.pad(np.asarray(r0_rect, dtype=np.float64), ((0, 1), (0, 1)), 'constant', constant_values=0)
    r0[3, 3] = 1
    tf = np.pad(np.asarray(tr_velodyne_to_cam, dtype=np.float64), ((0, 1), (0, 0)), 'constant', constant_values=0)
    tf[3, 3] = 1
    one_pad = np.ones(xyz_lidar.shape[0]).reshape(-1, 1)
    xyz = np.append(xyz_lidar, one_pad, axis=1)
    rectified = np.dot(r0, tf)                                                   # :406
    return np.dot(rectified, xyz.T)[0:3].T                                       # :407-410


def project_to_image(point_cloud, p):
    """calib_utils.py:281-297: (3, N) camera-frame points -> (2, N) pixels."""
    pts_2d = np.dot(p, np.append(point_cloud, np.ones((1, point_cloud.shape[1])), axis=0))
    pts_2d[0, :] = pts_2d[0, :] / pts_2d[2, :]
    pts_2d[1, :] = pts_2d[1, :] / pts_2d[2, :]
    return np.delete(pts_2d, 2, 0)


def get_lidar_point_cloud(velo_xyzi, p2, r0_rect, tr_velodyne_to_cam, im_size=None, min_intensity=None):
    """obj_utils.get_lidar_point_cloud (obj_utils.py:220-268) after the two file reads: velo_xyzi float32 [N,4] as
    calib_utils.read_lidar returns it (x, y, z, i).  Returns the (3, M) float64 camera-frame cloud."""
    x, y, z, i = (velo_xyzi[:, c] for c in range(4))
    pts = np.vstack((x, y, z)).T                                                 # float32 [N,3]  (:241)
    pts = lidar_to_cam_frame(pts, r0_rect, tr_velodyne_to_cam)                   # float64
    if not im_size:
        return pts.T                                                             # :245-247
    pts = pts[pts[:, 2] > 0]                                                     # :251
    point_cloud = pts.T
    point_in_im = project_to_image(point_cloud, p=p2).T                          # :255
    image_filter = ((point_in_im[:, 0] > 0) & (point_in_im[:, 0] < im_size[0]) &
                    (point_in_im[:, 1] > 0) & (point_in_im[:, 1] < im_size[1]))  # :258-261
    if not min_intensity:
        return pts[image_filter].T
    intensity_filter = i > min_intensity                                         # :266: `i` was NOT cut by z > 0 above
    return pts[np.logical_and(image_filter, intensity_filter)].T                 # raises when a point had z <= 0


# ------------------------------------------------------------------ augmentation hooks
def flip_point_cloud(point_cloud):
    """/root/reference/avod/avod/datasets/kitti/kitti_aug.py:24-29."""
    flipped = np.copy(point_cloud)
    flipped[0] = -point_cloud[0]
    return flipped


def mv3d_project_round(lidar_pc, P):
    """img_index2 of /root/reference/MV3D_TF_release/lib/roi_data_layer/minibatch_mv3d_img.py:172-174 (:183-185):
    np.round(projectToImage(lidar_pc[:, 0:3].T, P)).astype(int), projectToImage = lib/utils/transform.py:429-452."""
    pts = lidar_pc[:, 0:3].transpose()
    mat = np.vstack((pts, np.ones((pts.shape[1]))))
    uvw = np.dot(np.asarray(P, dtype=np.float64), mat)
    uv = np.stack((uvw[0] / uvw[2], uvw[1] / uvw[2]))
    with np.errstate(invalid="ignore"):
        return np.round(uv).astype(int)


def mv3d_augment_points(lidar_pc, sx, sz, expansion_ratio, rotation_angle):
    """The point-cloud half of augment_voxel (minibatch_mv3d_img.py:148, :176-181), on a copy."""
    pc = np.array(lidar_pc, dtype=np.float64)
    rot_mat = np.array([[np.cos(rotation_angle), np.sin(rotation_angle)],
                        [-np.sin(rotation_angle), np.cos(rotation_angle)]]).reshape(2, 2)
    pc[:, 0] += sx
    pc[:, 2] += sz
    pc[:, 0:3] *= expansion_ratio
    pc[:, [0, 2]] = np.dot(rot_mat, pc[:, [0, 2]].transpose()).transpose()
    return pc


def augment_fv_index(img_index, sx, sy, expansion_ratio):
    """minibatch_mv3d_img.py:205-206, on a copy."""
    out = np.array(img_index)
    out[0, :] = (out[0, :] * expansion_ratio + sx).astype(int)
    out[1, :] = (out[1, :] * expansion_ratio + sy).astype(int)
    return out
