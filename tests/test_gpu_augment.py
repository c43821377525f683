"""GPU parity tests of the augmentation hooks (sparse_pooling_b200.augment: shpl_flip_point_cloud,
shpl_mv3d_project_augment, shpl_augment_fv_index) against the fixtures the REFERENCE's own functions produced
(tests/golden/augment_hooks.npz) and against the numpy restatements in oracle/feeder_oracle.py.  Correctly rounded
fp64 in the reference's order and integer work: every comparison is bit-exact.  Run with `pytest -m gpu` on a B200."""
import hashlib
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import feeder_oracle as fo, index_oracle as io, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "augment_hooks.npz"), allow_pickle=False)


def test_flip_matches_the_reference_fixture(shpl, golden):
    aug = shpl.augment
    pts = synth.lidar_scan(1, az_step_deg=0.09).T
    assert digest(pts) == str(golden["flip_input_sha"])
    before = pts.copy()
    fl = aug.flip_point_cloud(pts)
    assert fl.dtype == np.float64 and digest(fl) == str(golden["flip_sha"])
    np.testing.assert_array_equal(pts, before)                               # a copy, like kitti_aug.py:27
    np.testing.assert_array_equal(aug.flip_ground_plane(np.array([0.01, -1.0, 0.02, 1.65])), golden["flip_plane"])
    np.testing.assert_array_equal(aug.flip_stereo_calib_p2(synth.P2_KITTI, (375, 1242)), golden["flip_p2"])
    # device-resident form with a device-side count: points past the count are left alone
    t = torch.from_numpy(pts).cuda()
    n_dev = torch.tensor([1000], dtype=torch.int32, device="cuda")
    aug.flip_point_cloud_(t, n_dev=n_dev)
    got = t.cpu().numpy()
    np.testing.assert_array_equal(got[0, :1000], -pts[0, :1000])
    np.testing.assert_array_equal(got[0, 1000:], pts[0, 1000:])
    np.testing.assert_array_equal(got[1:], pts[1:])


def test_mv3d_project_and_augment_matches_the_reference_fixture(shpl, golden):
    aug = shpl.augment
    f = synth.mv3d_frame(seed=7, n_points=5000)
    pc = synth.mv3d_cam4(f)
    assert digest(pc) == str(golden["voxel_input_sha"])
    sx, sz, ratio, angle = golden["voxel_params"]
    out, img2 = aug.project_and_augment_points(pc.copy(), synth.P2_KITTI, sx, sz, np.array([ratio]), np.array([angle]))
    np.testing.assert_array_equal(img2, golden["voxel_img_index2"])
    assert digest(out) == str(golden["voxel_pc_sha"])
    # projection only (minibatch_mv3d_img.py:183-185): the points stay as they are
    same, img2b = aug.project_and_augment_points(pc.copy(), synth.P2_KITTI)
    np.testing.assert_array_equal(same, pc)
    np.testing.assert_array_equal(img2b, golden["voxel_img_index2"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mv3d_project_and_augment_matches_oracle(shpl, seed):
    """Random draws in augment_voxel's ranges (:133-136), 20 000 points, some behind / on the camera plane."""
    aug = shpl.augment
    rng = np.random.default_rng(seed)
    f = synth.mv3d_frame(seed=40 + seed, n_points=20000)
    pc = synth.mv3d_cam4(f)
    pc[:50, 2] = -pc[:50, 2]                       # behind the camera
    pc[50:60, 2] = 0.0                             # on the camera plane: w = 0 -> inf / nan -> INT64_MIN after astype(int)
    pc[60, :3] = 0.0
    sx, sz = rng.uniform(-0.8, 0.8, 2)
    ratio, angle = rng.uniform(0.95, 1.05, 1), rng.uniform(-np.pi / 10, np.pi / 10, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref_img2 = fo.mv3d_project_round(pc, synth.P2_KITTI)
    ref_pc = fo.mv3d_augment_points(pc, sx, sz, ratio, angle)
    t = torch.from_numpy(pc).cuda()
    out, img2 = aug.project_and_augment_points(t, synth.P2_KITTI, sx, sz, ratio, angle)
    assert out.data_ptr() == t.data_ptr() and img2.is_cuda
    np.testing.assert_array_equal(img2.cpu().numpy(), ref_img2)
    np.testing.assert_array_equal(out.cpu().numpy(), ref_pc)


def test_augment_fv_index_matches_fixture_and_oracle(shpl, golden):
    aug = shpl.augment
    f = synth.mv3d_frame(seed=7, n_points=5000)
    img_index = np.vstack((f["img_index2"], np.zeros((1, f["img_index2"].shape[1]), dtype=int)))
    assert digest(img_index) == str(golden["fv_input_sha"])
    fsx, fsy, fratio = golden["fv_params"]
    got = aug.augment_fv_index(img_index.copy(), fsx, fsy, np.array([fratio]))
    np.testing.assert_array_equal(got, golden["fv_img_index"])
    # negative indices truncate toward zero, like astype(int)
    rng = np.random.default_rng(4)
    wild = np.vstack((rng.integers(-3000, 3000, (2, 4000)), np.zeros((1, 4000), dtype=int)))
    for sx, sy, r in ((0.3, 9.99, 0.95), (7.5, 0.0, 1.05), (2.25, 3.75, 1.0)):
        ref = fo.augment_fv_index(wild, sx, sy, np.array([r]))
        t = torch.from_numpy(wild.copy()).cuda()
        aug.augment_fv_index(t, sx, sy, r)
        np.testing.assert_array_equal(t.cpu().numpy(), ref)
    # augment_fv itself: the reference's random draws, in its order
    np.random.seed(99)
    blobs, shift, ratio = aug.augment_fv(dict(image_data=np.zeros((375, 1242, 3), np.float32), gt_boxes=np.zeros((1, 5)),
                                              img_index=img_index.copy()), scale=10)
    np.testing.assert_array_equal(np.array([shift[0], shift[1], ratio[0]]), golden["fv_params"])
    np.testing.assert_array_equal(blobs["img_index"], golden["fv_img_index"])
    np.testing.assert_array_equal(np.array(blobs["image_data"].shape), golden["fv_image_shape"])


def test_flipped_chain_scan_to_plan_on_the_device(shpl):
    """kitti_dataset.py:304-311 + :376-378 on the device: flip the point cloud and the ground plane, slice, build the
    correspondences with the flipped P2 -- against the same chain through the oracles."""
    aug = shpl.augment
    pts = synth.lidar_scan(5, az_step_deg=0.2).T
    gp = np.array([0.0, -1.0, 0.0, 1.65])
    p2f = aug.flip_stereo_calib_p2(synth.P2_KITTI, (360, 1200))
    _, _, ref_idx, ref_pts = fo.generate_bev(fo.flip_point_cloud(pts), aug.flip_ground_plane(gp), synth.AVOD_EXTENTS,
                                             synth.AVOD_VOXEL, -0.2, 2.3, 5)
    d_ref = io.gen_sparse_pooling_input_avod(ref_pts, ref_idx, p2f, [1200, 360], (700, 800))
    o_ref = io.produce_sparse_pooling_input(d_ref, stride=[4, 4])

    class Calib:
        p2 = p2f
    import types
    cfg = types.SimpleNamespace(height_lo=-0.2, height_hi=2.3, num_slices=5)
    t = aug.flip_point_cloud_(torch.from_numpy(pts).cuda())
    maps, idx, upts = shpl.BevSlices(cfg, None).generate_bev("lidar", t, aug.flip_ground_plane(gp), synth.AVOD_EXTENTS,
                                                              synth.AVOD_VOXEL, output_indices=True)
    d = shpl.gen_sparse_pooling_input_avod(upts, idx, Calib, [1200, 360], (700, 800))
    o = shpl.produce_sparse_pooling_input(d, stride=[4, 4])
    assert len(o_ref["Mij_pool"]) > 3000
    np.testing.assert_array_equal(o["Mij_pool"].cpu().numpy(), o_ref["Mij_pool"])
    np.testing.assert_array_equal(o["img_index_flip_pool"].cpu().numpy(), o_ref["img_index_flip_pool"])
