"""GPU parity tests of the point-cloud ingest (shpl_lidar_to_cam through the drop-in lidar_ingest module) against the
ingest oracle and the fixtures the REFERENCE's obj_utils.get_lidar_point_cloud produced from KITTI-format files
(tests/golden/lidar_ingest_seed*.npz).  fp64 arithmetic in the reference's rounding order, selection, stable
compaction: bit-exact.  Run with `pytest -m gpu` on a B200."""
import hashlib
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import feeder_oracle as fo, synth  # noqa: E402

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


def calib_of(g):
    return types.SimpleNamespace(p2=g["p2"], r0_rect=g["r0_rect"], tr_velodyne_to_cam=g["tr_velodyne_to_cam"])


@pytest.mark.parametrize("seed,az", [(1, 0.4), (2, 0.15)])
def test_ingest_matches_reference_fixture_through_files(shpl, golden_dir, tmp_path, seed, az):
    """The whole drop-in call: KITTI-format files on disk -> (3, M) cloud, against what the reference returned."""
    g = np.load(os.path.join(golden_dir, "lidar_ingest_seed%d.npz" % seed), allow_pickle=False)
    scan = synth.velodyne_scan(seed, az_step_deg=az)
    (tmp_path / "calib").mkdir()
    (tmp_path / "velodyne").mkdir()
    (tmp_path / "calib" / ("%06d.txt" % seed)).write_text(synth.kitti_calib_text())
    scan.tofile(str(tmp_path / "velodyne" / ("%06d.bin" % seed)))
    li = shpl.lidar_ingest
    pc = li.get_lidar_point_cloud(seed, str(tmp_path / "calib"), str(tmp_path / "velodyne"), im_size=list(g["im_size"]))
    assert isinstance(pc, np.ndarray) and pc.dtype == np.float64 and pc.shape == (3, int(g["n_fov"]))
    assert digest(pc) == str(g["fov_sha"])
    if "fov_points" in g:
        np.testing.assert_array_equal(pc, g["fov_points"])
    pc_all = li.get_lidar_point_cloud(seed, str(tmp_path / "calib"), str(tmp_path / "velodyne"))
    assert pc_all.shape == (3, len(scan)) and digest(pc_all) == str(g["all_sha"])
    cal = li.read_calibration(str(tmp_path / "calib"), seed)
    np.testing.assert_array_equal(cal.p2, g["p2"])
    np.testing.assert_array_equal(cal.r0_rect, g["r0_rect"])
    np.testing.assert_array_equal(cal.tr_velodyne_to_cam, g["tr_velodyne_to_cam"])


@pytest.mark.parametrize("seed,az,im_size", [(3, 0.3, [1242, 375]), (4, 0.05, [1200, 360]), (5, 1.0, [64, 48])])
def test_ingest_matches_oracle(shpl, golden_dir, seed, az, im_size):
    g = np.load(os.path.join(golden_dir, "lidar_ingest_seed1.npz"), allow_pickle=False)
    cal = calib_of(g)
    scan = synth.velodyne_scan(seed, az_step_deg=az)
    ref = fo.get_lidar_point_cloud(scan, cal.p2, cal.r0_rect, cal.tr_velodyne_to_cam, im_size=im_size)
    got = shpl.lidar_ingest.lidar_to_cam_fov(scan, cal, im_size=im_size)
    np.testing.assert_array_equal(got, ref)
    # CUDA tensor in -> CUDA tensor out
    t = shpl.lidar_ingest.lidar_to_cam_fov(torch.from_numpy(scan).cuda(), cal, im_size=im_size)
    assert t.is_cuda and t.dtype == torch.float64
    np.testing.assert_array_equal(t.cpu().numpy(), ref)


def test_ingest_min_intensity_branch_behaves_like_the_reference(shpl, golden_dir):
    g = np.load(os.path.join(golden_dir, "lidar_ingest_seed1.npz"), allow_pickle=False)
    cal = calib_of(g)
    scan = synth.velodyne_scan(6, az_step_deg=0.5)
    # points behind the camera exist: the reference's masks have different lengths and numpy raises (obj_utils.py:266)
    with pytest.raises((ValueError, IndexError)):
        fo.get_lidar_point_cloud(scan, cal.p2, cal.r0_rect, cal.tr_velodyne_to_cam, im_size=[1242, 375], min_intensity=0.5)
    with pytest.raises(ValueError):
        shpl.lidar_ingest.lidar_to_cam_fov(scan, cal, im_size=[1242, 375], min_intensity=0.5)
    # with every point in front of the camera the branch works: image filter AND intensity > threshold
    front = scan[scan[:, 0] > 1.0]
    ref = fo.get_lidar_point_cloud(front, cal.p2, cal.r0_rect, cal.tr_velodyne_to_cam, im_size=[1242, 375], min_intensity=0.5)
    got = shpl.lidar_ingest.lidar_to_cam_fov(front, cal, im_size=[1242, 375], min_intensity=0.5)
    assert 0 < ref.shape[1] < front.shape[0]
    np.testing.assert_array_equal(got, ref)


def test_raw_scan_to_plans_without_host_reads(shpl, golden_dir):
    """scan -> shpl_lidar_to_cam -> shpl_bev_slices -> shpl_build_avod with every count handed on as a DEVICE pointer
    (P_dev, N_dev): no host read between the raw scan and the CSR plan.  Same plan as the host-side chain through
    the oracles; stale data beyond the counts is ignored."""
    import ctypes
    from oracle import index_oracle as io
    from sparse_pooling_b200 import bev_slices as bs
    from sparse_pooling_b200.pipeline import FramePipeline, LayerSpec
    g = np.load(os.path.join(golden_dir, "lidar_ingest_seed1.npz"), allow_pickle=False)
    cal = calib_of(g)
    scan = synth.velodyne_scan(7, az_step_deg=0.1)
    dev = torch.device("cuda", 0)
    GP = np.array([0.0, -1.0, 0.0, 1.65])
    velo = torch.from_numpy(scan).to(dev)
    n = scan.shape[0]
    cam = torch.full((3, n), 3.0, dtype=torch.float64, device=dev)          # stale points inside the extents
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    shpl.lidar_ingest.lidar_to_cam_raw(velo, n, cal, [1242, 375], cam, counts)
    cap = 65536
    work = bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, cap, dev, with_maps=False)
    bs.bev_slices_raw(cam, cam.stride(0), cam.stride(1), n, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5, np.log(16), work,
                      p_dev=ctypes.c_void_p(counts.data_ptr()))
    spec = LayerSpec("s1", (700, 800), (360, 1200), 8, 8, (1, 1), False, (1200, 360), (700, 800))
    pipe = FramePipeline([spec], cap, dev)
    pipe.build_layer(0, work.unique_pts, work.voxel_indices, cal.p2, cap, stream, n_dev=ctypes.c_void_p(work.counts.data_ptr()))
    torch.cuda.synchronize()
    # host chain through the oracles
    pc = fo.get_lidar_point_cloud(scan, cal.p2, cal.r0_rect, cal.tr_velodyne_to_cam, im_size=[1242, 375])
    assert int(counts[0].item()) == pc.shape[1]
    _, _, idx, upts = fo.generate_bev(pc, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    assert int(work.counts[0].item()) == len(idx) < cap
    d = io.gen_sparse_pooling_input_avod(upts, idx, cal.p2, [1200, 360], (700, 800))
    o = io.produce_sparse_pooling_input(d, stride=[1, 1])
    ref = io.build_plan(o["Mij_pool"], np.ones(len(o["Mij_pool"]), np.float32), o["img_index_flip_pool"], 560000, 360, 1200)
    plan = pipe.layers[0].plan
    nnz = int(plan.counts.cpu()[0, 3])
    assert nnz == len(ref["csr_src"]) > 1000
    np.testing.assert_array_equal(plan.row_ptr.cpu().numpy(), ref["row_ptr"])
    np.testing.assert_array_equal(plan.csr_src.cpu().numpy()[:nnz], ref["csr_src"])
    np.testing.assert_array_equal(plan.csrT_dst.cpu().numpy()[:nnz], ref["csrT_dst"])
