"""Seeded synthetic inputs for the SURVEY.md 8(d) configurations.

Input generation only (no reference algorithm in here): shared by tests/, smoke(), bench.py and the tools so that
every arm (CUDA, oracle, reference) sees byte-identical input.
Nothing here comes from KITTI files; geometry constants are the ones the
reference's configs state.
"""
import numpy as np

# KITTI P2 used throughout the survey's probes (Appendix A.3)
P2_KITTI = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float64)

# avod area extents / voxel size (pyramid_people_with_2NHSP_example_train.config:160-172)
AVOD_EXTENTS = np.array([[-40.0, 40.0], [-5.0, 3.0], [0.0, 70.0]])
AVOD_VOXEL = 0.1
AVOD_BEV_HW = (700, 800)
AVOD_IMG_WH = (1200, 360)      # image tower input 360x1200 (config :14-15)

# name -> dict(bev H,W ; img H,W ; C_bev, C_img ; stride (img, bev) ; dual)
CONFIGS = {
    # config 1: avod pre-RPN SHPL, stride 1, single direction
    "kitti_s1_c32": dict(bev_hw=(700, 800), img_hw=(360, 1200), c_bev=32, c_img=32, stride=(1, 1), dual=False),
    # config 2 layer A: after VGG conv4, stride 8, dual (code-as-written 87x100 BEV map)
    "kitti_s8_c256_dual": dict(bev_hw=(87, 100), img_hw=(45, 150), c_bev=256, c_img=256, stride=(8, 8), dual=True),
    # config 2': RetinaNet P2, stride 4
    "kitti_s4_c256": dict(bev_hw=(175, 200), img_hw=(90, 300), c_bev=256, c_img=256, stride=(4, 4), dual=False),
    # config 3: MV3D ped/cyc middle-stage fusion, strides img 8 / bev 2
    "mv3d_ped_c768": dict(bev_hw=(100, 120), img_hw=(48, 160), c_bev=768, c_img=768, stride=(8, 2), dual=False),
    # config 4: full scans, C=128
    "kitti_s1_c128": dict(bev_hw=(700, 800), img_hw=(360, 1200), c_bev=128, c_img=128, stride=(1, 1), dual=False),
}


def lidar_scan(seed, az_step_deg=0.09, n_beams=64, obstacle_p=0.35):
    """Synthetic 64-beam scan in the camera frame:
    x right, y down, z forward; ground plane 1.65 m below the sensor.
    Returns f64 [n,3] of points with z>0 that project inside the raw 1242x375 image."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.deg2rad(np.arange(-45.0, 45.0, az_step_deg))
    E, A = np.meshgrid(elev, az, indexing="ij")
    down = -np.sin(E)                                   # >0 for beams that point below the horizon
    with np.errstate(divide="ignore", invalid="ignore"):
        ground = np.where(down > 1e-6, 1.65 / down, np.inf)
    rng_m = ground.copy()
    hit = rng.random(E.shape) < obstacle_p
    rng_m = np.where(hit, np.minimum(rng.uniform(5.0, 75.0, E.shape), ground), rng_m)
    ok = np.isfinite(rng_m) & (rng_m < 80.0)
    r, e, a = rng_m[ok], E[ok], A[ok]
    pts = np.stack((r * np.cos(e) * np.sin(a), -r * np.sin(e), r * np.cos(e) * np.cos(a)), axis=1)
    pts += rng.normal(0.0, 0.01, pts.shape)
    hom = np.c_[pts, np.ones(len(pts))] @ P2_KITTI.T
    u, v = hom[:, 0] / hom[:, 2], hom[:, 1] / hom[:, 2]
    keep = (pts[:, 2] > 0) & (u > 0) & (u < 1242) & (v > 0) & (v < 375)
    return np.ascontiguousarray(pts[keep])


def lidar_scan_gappy(seed, az_step_deg=0.3):
    """lidar_scan with two height slices (of the 5 avod slices, -0.2..2.3 m) thinned out: slice 2 keeps exactly
    one point and slice 4 none, so BevSlices re-uses the previous slice's grid there."""
    pts = lidar_scan(seed, az_step_deg=az_step_deg)
    h = 1.65 - pts[:, 1]
    in2 = (h > 0.75) & (h < 1.35)
    in4 = (h > 1.75) & (h < 2.35)
    keep = ~(in2 | in4)
    one = np.nonzero((h > 0.9) & (h < 1.2))[0]
    if one.size:
        keep[one[0]] = True
    return np.ascontiguousarray(pts[keep])


def one_point_per_cell(points, n_slices=5, h_lo=-0.2, h_hi=2.3, ground_y=1.65):
    """A light stand-in for the avod feeder (BevSlices + VoxelGrid2D):
    per height slice, the first point (in x,z,y lexicographic cell order) of every
    occupied 0.1 m cell, with the reference's flipped z index (num_divisions - z).
    Returns (unique_pts f64 [N,3], voxel_indices int64 [N,2] = (x_idx, zflip))."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    inside = (x >= AVOD_EXTENTS[0, 0]) & (x < AVOD_EXTENTS[0, 1]) & (z >= AVOD_EXTENTS[2, 0]) & (z < AVOD_EXTENTS[2, 1])
    height = ground_y - y                                # height above the ground plane
    step = (h_hi - h_lo) / n_slices
    out_p, out_i = [], []
    for s in range(n_slices):
        sel = np.nonzero(inside & (height > h_lo + s * step) & (height < h_lo + (s + 1) * step))[0]
        if sel.size < 2:
            continue
        d = np.floor(points[sel] / AVOD_VOXEL).astype(np.int32)
        order = np.lexsort((d[:, 1], d[:, 2], d[:, 0]))
        d, sel = d[order], sel[order]
        first = np.r_[True, (d[1:, 0] != d[:-1, 0]) | (d[1:, 2] != d[:-1, 2])]
        xi = d[first, 0].astype(np.int64) + 400
        zi = d[first, 2].astype(np.int64)
        out_p.append(points[sel[first]])
        out_i.append(np.stack((xi, 700 - zi), axis=1))
    return np.vstack(out_p), np.vstack(out_i)


def avod_frame(seed, az_step_deg=0.05):
    """points + voxel_indices as kitti_dataset.py:376-377 hands them to
    gen_sparse_pooling_input_avod; ~20k pairs at the default azimuth step."""
    pts, idx = one_point_per_cell(lidar_scan(seed, az_step_deg=az_step_deg))
    return dict(points=pts, voxel_indices=idx, P=P2_KITTI.copy(),
                im_size=list(AVOD_IMG_WH), bv_size=AVOD_BEV_HW)


def direct_pairs(seed, n, bev_hw=(700, 800), img_wh=(1200, 360), skew="uniform"):
    """SURVEY.md 8(d) 'direct' variant: the dict gen_sparse_pooling_input_avod would
    return, drawn directly.  skew in {uniform, zipf, ground}."""
    rng = np.random.default_rng(seed)
    Hb, Wb = bev_hw
    if skew == "uniform":
        x = rng.integers(0, Wb, n)
        z = rng.integers(1, Hb + 1, n)
    elif skew == "zipf":
        cell = (rng.zipf(1.2, n) - 1) % (Hb * Wb)
        cell = (cell * 2654435761) % (Hb * Wb)          # spread the hot cells over the map
        x, z = cell % Wb, cell // Wb
    elif skew == "ground":                               # 80 % of pairs in 2 % of the rows
        hot = rng.choice(Hb * Wb, size=max(1, (Hb * Wb) // 50), replace=False)
        pick_hot = rng.random(n) < 0.8
        cell = np.where(pick_hot, hot[rng.integers(0, hot.size, n)], rng.integers(0, Hb * Wb, n))
        x, z = cell % Wb, cell // Wb
    else:
        raise ValueError(skew)
    u = rng.integers(0, img_wh[0], n)
    v = rng.integers(0, img_wh[1], n)
    img_index = np.zeros((3, n), dtype=np.float64)
    img_index[0], img_index[1] = u, v
    return dict(bv_index=np.stack((x, z), axis=1).astype(np.int64), img_index=img_index,
                bv_size=np.array(bev_hw), img_size=np.array(img_wh))


def mv3d_cam4(frame):
    """camera-frame [n,4] = (x, y, z, reflectance) of a mv3d_frame, as point_cloud_2_top_sparse(points_in_cam=True) takes it."""
    fsh = frame["points_fsh"]
    refl = (np.arange(len(fsh)) % 97) / 97.0
    return np.ascontiguousarray(np.c_[fsh[:, [1, 2, 0]], refl])


def mv3d_frame(seed, n_points=20000, max_points=45, car=False):
    """MV3D ped/cyc feeder inputs: camera-frame points in
    fwd (0,48) x side (-20,20) x height (-1,3) at 0.2/0.2/0.4 m, a dense cluster so
    the 45-point cap bites, image padded to 1280x384 (config.py:229)."""
    rng = np.random.default_rng(seed)
    n_cluster = n_points // 5
    fwd = np.r_[rng.uniform(0.5, 47.5, n_points - n_cluster), rng.normal(12.0, 0.15, n_cluster)]
    side = np.r_[rng.uniform(-19.5, 19.5, n_points - n_cluster), rng.normal(1.0, 0.15, n_cluster)]
    hgt = np.r_[rng.uniform(-0.9, 2.9, n_points - n_cluster), rng.normal(1.0, 0.1, n_cluster)]
    perm = rng.permutation(n_points)
    fsh = np.stack((fwd, side, hgt), axis=1)[perm]      # (forward, side, height) == cam (z, x, y)
    cam = fsh[:, [1, 2, 0]]
    hom = np.c_[cam, np.ones(n_points)] @ P2_KITTI.T
    img_index2 = np.rint(np.stack((hom[:, 0] / hom[:, 2], hom[:, 1] / hom[:, 2]))).astype(int)
    if car:      # config_voxels.py:33-48: side (-40,40), fwd (0,70.4), height (-3,1), 35 points per voxel
        fsh = fsh * np.array([70.4 / 48.0, 2.0, 1.0]) - np.array([0.0, 0.0, 2.0])
        cam = fsh[:, [1, 2, 0]]
        hom = np.c_[cam, np.ones(n_points)] @ P2_KITTI.T
        img_index2 = np.rint(np.stack((hom[:, 0] / hom[:, 2], hom[:, 1] / hom[:, 2]))).astype(int)
        return dict(points_fsh=fsh, img_index2=img_index2, res=0.2, zres=0.4,
                    side_range=(-40, 40 - 0.01), fwd_range=(0, 70.4 - 0.01), height_range=(-3, 1 - 0.01), max_points=35,
                    bv_size=[int((40 - 0.01 - -40) / 0.2) + 1, int((70.4 - 0.01 - 0) / 0.2) + 1],
                    img_size=np.array([1280, 384]), stride=[8, 2])
    return dict(points_fsh=fsh, img_index2=img_index2, res=0.2, zres=0.4,
                # construct_voxel.py:11-13: ranges are (MIN, MAX-0.01) of config_voxels.py:50-57
                side_range=(-20, 20 - 0.01), fwd_range=(0, 48 - 0.01), height_range=(-1, 3 - 0.01),
                max_points=max_points,
                bv_size=[int((20 - 0.01 - -20) / 0.2) + 1, int((48 - 0.01 - 0) / 0.2) + 1],   # construct_voxel.py:81-84 -> [200, 240]
                img_size=np.array([1280, 384]), stride=[8, 2])


def features(seed, shape):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


# ------------------------------------------------------------------ raw velodyne scans
# KITTI object calibration 000000-like values (R0_rect, Tr_velo_to_cam; P2 as above)
R0_RECT_KITTI = np.array([[9.999239e-01, 9.837760e-03, -7.445048e-03],
                          [-9.869795e-03, 9.999421e-01, -4.278459e-03],
                          [7.402527e-03, 4.351614e-03, 9.999631e-01]])
TR_VELO_TO_CAM_KITTI = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                                 [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                                 [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01]])


def velodyne_scan(seed, az_step_deg=0.2, n_beams=64, obstacle_p=0.35):
    """Raw 360-degree velodyne scan, float32 [N,4] = (x forward, y left, z up, intensity) in the LIDAR frame, as a KITTI
    .bin file holds it (calib_utils.read_lidar): ground 1.73 m below the sensor, random obstacles."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.deg2rad(np.arange(-180.0, 180.0, az_step_deg))
    E, A = np.meshgrid(elev, az, indexing="ij")
    down = -np.sin(E)
    with np.errstate(divide="ignore", invalid="ignore"):
        ground = np.where(down > 1e-6, 1.73 / down, np.inf)
    hit = rng.random(E.shape) < obstacle_p
    rng_m = np.where(hit, np.minimum(rng.uniform(3.0, 75.0, E.shape), ground), ground)
    ok = np.isfinite(rng_m) & (rng_m < 80.0)
    r, e, a = rng_m[ok], E[ok], A[ok]
    xyz = np.stack((r * np.cos(e) * np.cos(a), r * np.cos(e) * np.sin(a), r * np.sin(e)), axis=1)
    xyz += rng.normal(0.0, 0.01, xyz.shape)
    inten = rng.random(len(xyz))
    return np.ascontiguousarray(np.c_[xyz, inten].astype(np.float32))


def kitti_calib_text(p2=None, r0=None, tr=None):
    """A KITTI object calib file (7 lines: P0..P3, R0_rect, Tr_velo_to_cam, Tr_imu_to_velo) as read_calibration parses it."""
    p2 = P2_KITTI if p2 is None else p2
    r0 = R0_RECT_KITTI if r0 is None else r0
    tr = TR_VELO_TO_CAM_KITTI if tr is None else tr

    def line(name, m):
        return name + ": " + " ".join("%.12e" % v for v in np.asarray(m).reshape(-1))
    p0 = p2.copy()
    p0[:, 3] = 0.0
    return "\n".join([line("P0", p0), line("P1", p0), line("P2", p2), line("P3", p2), line("R0_rect", r0),
                      line("Tr_velo_to_cam", tr), line("Tr_imu_to_velo", tr)]) + "\n"
