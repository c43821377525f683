"""GPU parity tests of the MV3D voxel feeder (shpl_mv3d_voxelize through the drop-in
construct_voxel.point_cloud_2_top_sparse) against the feeder oracle and the fixtures the REFERENCE's
construct_voxel.py produced (tests/golden/mv3d_*.npz).  Index work and correctly rounded fp64 arithmetic in
the reference's order: every comparison is bit-exact.  Run with `pytest -m gpu` on a B200."""
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


def run_gpu(shpl, f, points=None, img2=None):
    cv = shpl.construct_voxel
    cv.MAX_NUM_POINTS = f["max_points"]
    try:
        return cv.point_cloud_2_top_sparse(synth.mv3d_cam4(f) if points is None else points, res=f["res"], zres=f["zres"],
                                           side_range=f["side_range"], fwd_range=f["fwd_range"], height_range=f["height_range"],
                                           points_in_cam=True, img_index2=f["img_index2"] if img2 is None else img2)
    finally:
        cv.MAX_NUM_POINTS = 45


def run_oracle(f, points=None, img2=None):
    return fo.point_cloud_2_top_sparse(synth.mv3d_cam4(f) if points is None else points,
                                       f["img_index2"] if img2 is None else img2, f["res"], f["zres"], f["side_range"],
                                       f["fwd_range"], f["height_range"], f["max_points"])


def assert_same(got, ref):
    vd, vfs, img_index, bv_index, m_val = got
    rvd, rvfs, rimg, rbv, rmv = ref
    np.testing.assert_array_equal(np.asarray(vfs), rvfs)
    np.testing.assert_array_equal(img_index, rimg)
    np.testing.assert_array_equal(bv_index, rbv)
    np.testing.assert_array_equal(m_val, rmv)
    np.testing.assert_array_equal(vd["coordinate_buffer"], rvd["coordinate_buffer"])
    np.testing.assert_array_equal(vd["number_buffer"], rvd["number_buffer"])
    np.testing.assert_array_equal(vd["feature_buffer"], rvd["feature_buffer"])


@pytest.mark.parametrize("name,seed,n,kw", [("mv3d_seed5", 5, 6000, {}), ("mv3d_car_seed6", 6, 9000, dict(car=True))])
def test_mv3d_feeder_matches_reference_fixture(shpl, golden_dir, name, seed, n, kw):
    g = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    f = synth.mv3d_frame(seed=seed, n_points=n, **kw)
    vd, vfs, img_index, bv_index, m_val = run_gpu(shpl, f)
    assert isinstance(m_val, np.ndarray) and m_val.dtype == np.float64 and bv_index.dtype == np.int64
    np.testing.assert_array_equal(np.asarray(vfs), g["voxel_full_size"])
    np.testing.assert_array_equal(img_index, g["img_index"])
    np.testing.assert_array_equal(bv_index, g["bv_index"])
    np.testing.assert_array_equal(m_val, g["M_val"])                          # 1/count, bit-exact
    np.testing.assert_array_equal(vd["coordinate_buffer"], g["coordinate_buffer"])
    np.testing.assert_array_equal(vd["number_buffer"], g["number_buffer"])
    np.testing.assert_array_equal(vd["feature_buffer"][:64], g["feature_buffer_head"])
    assert vd["feature_buffer"].shape == (len(g["number_buffer"]), f["max_points"], 7)
    assert digest(vd["feature_buffer"]) == str(g["feature_buffer_sha"])      # the whole [V,T,7] buffer, bit for bit
    # straight into produce_sparse_pooling_input, as train_mv_voxel.py:325-326 does
    out = shpl.produce_sparse_pooling_input(dict(img_index=np.array(img_index, dtype=np.float64), img_size=f["img_size"],
                                                 bv_index=bv_index, bv_size=[vfs[1], vfs[2]]), M_val=m_val, stride=[8, 2])
    np.testing.assert_array_equal(out["Mij_pool"][:, 0], g["row"])
    np.testing.assert_array_equal(out["M_size"], g["M_size"])
    np.testing.assert_array_equal(out["img_index_flip_pool"], g["flip"])


@pytest.mark.parametrize("seed,n,kw", [(7, 20000, {}), (8, 120000, {}), (9, 60000, dict(car=True)), (10, 300, {})])
def test_mv3d_feeder_matches_oracle(shpl, seed, n, kw):
    f = synth.mv3d_frame(seed=seed, n_points=n, **kw)
    assert_same(run_gpu(shpl, f), run_oracle(f))


def test_mv3d_feeder_edge_cases(shpl):
    f = synth.mv3d_frame(seed=11, n_points=500)
    pts = synth.mv3d_cam4(f)
    # nothing in range
    far = pts.copy()
    far[:, 2] += 1000.0
    vd, vfs, img_index, bv_index, m_val = run_gpu(shpl, f, points=far)
    assert img_index.shape == (3, 0) and bv_index.shape == (0, 2) and m_val.shape == (0,)
    assert vd["feature_buffer"].shape == (0, 45, 7) and vd["number_buffer"].shape == (0,)
    # one point; and a point exactly on a range boundary is dropped (strict comparisons, :89-96)
    one = pts[:2].copy()
    one[1, 2] = 0.0                                       # forward == fwd_range[0]
    assert_same(run_gpu(shpl, f, points=one, img2=f["img_index2"][:, :2]), run_oracle(f, points=one, img2=f["img_index2"][:, :2]))
    # every point in ONE voxel: only the first 45 survive, all with weight 1/45
    same = np.tile(pts[:1], (200, 1))
    same[:, 3] = np.arange(200)
    img2 = np.stack((np.arange(200), np.arange(200) % 7))
    got = run_gpu(shpl, f, points=same, img2=img2)
    assert_same(got, run_oracle(f, points=same, img2=img2))
    assert got[4].shape == (45,) and np.all(got[4] == 1.0 / 45) and np.array_equal(got[2][0], np.arange(45))
    # the branches the reference cannot run
    with pytest.raises(AssertionError):
        shpl.construct_voxel.point_cloud_2_top_sparse(pts, img_index2=f["img_index2"])
    with pytest.raises(NameError):
        shpl.construct_voxel.point_cloud_2_top_sparse(pts, to_camera_frame=True, img_index2=f["img_index2"])


def test_mv3d_feeder_cuda_tensors_stay_on_device(shpl):
    f = synth.mv3d_frame(seed=12, n_points=20000)
    dev = torch.device("cuda", 0)
    pts = torch.from_numpy(synth.mv3d_cam4(f)).to(dev)
    img2 = torch.from_numpy(f["img_index2"]).to(dev)
    vd, vfs, img_index, bv_index, m_val = run_gpu(shpl, f, points=pts, img2=img2)
    assert img_index.is_cuda and bv_index.is_cuda and m_val.is_cuda and vd["feature_buffer"].is_cuda
    ref = run_oracle(f)
    got = ({k: v.cpu().numpy() for k, v in vd.items()}, vfs, img_index.cpu().numpy(), bv_index.cpu().numpy(), m_val.cpu().numpy())
    assert_same(got, ref)
    # the weights-only restatement agrees too
    _, _, bv2, mv2 = io.mv3d_voxel_weights(f["points_fsh"], f["res"], f["zres"], f["side_range"], f["fwd_range"],
                                           f["height_range"], f["max_points"])
    np.testing.assert_array_equal(bv2, got[3])
    np.testing.assert_array_equal(mv2, got[4])


def test_mv3d_config3_chain_full_shape_two_frames(shpl):
    """BASELINE config 3 at full MV3D shape: feeder -> produce_sparse_pooling_input(stride [8, 2], M_val = 1/count)
    -> _sparse_pool_op-style pooling of a 48x160x768 image map into the 100x120x768 BEV map (ped/cyc ranges,
    MV3D_voxel_train.py:134-150), two frames stacked in one launch, forward and backward against the C oracle."""
    from oracle import cref
    frames = [synth.mv3d_frame(seed=s, n_points=20000) for s in (31, 32)]
    C = 768
    rng = np.random.default_rng(5)
    plans, refs = [], []
    bev = rng.standard_normal((2, 100, 120, 8), dtype=np.float32)            # lidar_features stand-in (narrow dense part)
    img = rng.standard_normal((2, 48, 160, C), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    outs = []
    g = rng.standard_normal((2, 100, 120, 8 + C), dtype=np.float32)
    for k, f in enumerate(frames):
        # keep the points that project inside the padded 1280x384 image, as the MV3D pipeline does before this call
        # (minibatch_mv3d_img.py clips the cloud to the camera view); TF-CPU rejects out-of-range gather indices
        u, v = f["img_index2"]
        inside = (u >= 0) & (u < 1280) & (v >= 0) & (v < 384)
        cam4, img2 = np.ascontiguousarray(synth.mv3d_cam4(f)[inside]), np.ascontiguousarray(f["img_index2"][:, inside])
        vd, vfs, img_index, bv_index, m_val = run_gpu(shpl, f, points=cam4, img2=img2)
        assert list(vfs[1:]) == [200, 240] and len(m_val) > 3000
        d = dict(img_index=np.array(img_index, dtype=np.float64), img_size=f["img_size"], bv_index=bv_index, bv_size=[vfs[1], vfs[2]])
        o = shpl.produce_sparse_pooling_input(d, M_val=m_val, stride=[8, 2])
        assert o["M_size"].tolist() == [12000, len(m_val)]
        rvd, _, rimg, rbv, rmv = run_oracle(f, points=cam4, img2=img2)
        o_ref = io.produce_sparse_pooling_input(dict(img_index=np.array(rimg, dtype=np.float64), img_size=f["img_size"], bv_index=rbv,
                                                     bv_size=[200, 240]), M_val=rmv, stride=[8, 2])
        np.testing.assert_array_equal(o["Mij_pool"], o_ref["Mij_pool"])
        np.testing.assert_array_equal(o["img_index_flip_pool"], o_ref["img_index_flip_pool"])
        M = shpl.SparseTensor.from_sparse_pooling_input(o)
        fused, _ = shpl.sparse_pool_layer([tb[k:k + 1], ti[k:k + 1]], [C, 8], M, img_index_flip=o["img_index_flip_pool"])
        val = np.asarray(rmv, dtype=np.float32)                              # f64 -> f32 like the placeholder feed
        ref = cref.forward(bev[k], img[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"])
        np.testing.assert_array_equal(fused[0].detach().cpu().numpy(), ref)
        outs.append(fused)
        refs.append((o_ref, val))
    torch.autograd.backward(outs, [torch.from_numpy(g[k:k + 1]).cuda() for k in range(2)])
    for k, (o_ref, val) in enumerate(refs):
        gd, gs = cref.backward(g[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"], 8, (48, 160, C))
        np.testing.assert_array_equal(tb.grad[k].cpu().numpy(), gd)
        np.testing.assert_array_equal(ti.grad[k].cpu().numpy(), gs)


def test_mv3d_feeder_large_cloud_ticketed_scans(shpl):
    """200 k points: more CTAs than can be resident, so the three look-back scans order their tiles by arrival
    ticket, and the radix sort runs ~60 tiles per pass.  Still bit-identical to the oracle."""
    f = synth.mv3d_frame(seed=41, n_points=200000)
    got, ref = run_gpu(shpl, f), run_oracle(f)
    assert ref[4].shape[0] > 150000
    assert_same(got, ref)


def test_mv3d_config3_batch_of_8_frames_one_launch(shpl):
    """BASELINE config 3, batch 8: eight MV3D frames (feeder -> pairs with 1/count weights) stacked into ONE plan by
    build_pairs_plan and pooled by one launch each way, against the per-frame oracle chain.  (The reference is batch-1:
    minibatch_mv3d_img.py:49-50; the batch is this repo's sharding unit.)"""
    from oracle import cref
    C, B = 64, 8
    rng = np.random.default_rng(17)
    frames = [synth.mv3d_frame(seed=50 + k, n_points=6000) for k in range(B)]
    img_index, bv_index, m_val, refs = [], [], [], []
    for f in frames:
        u, v = f["img_index2"]
        inside = (u >= 0) & (u < 1280) & (v >= 0) & (v < 384)
        cam4, img2 = np.ascontiguousarray(synth.mv3d_cam4(f)[inside]), np.ascontiguousarray(f["img_index2"][:, inside])
        _, vfs, ii, bi, mv = run_gpu(shpl, f, points=cam4, img2=img2)
        img_index.append(ii)
        bv_index.append(bi)
        m_val.append(mv)
        _, _, rimg, rbv, rmv = run_oracle(f, points=cam4, img2=img2)
        o_ref = io.produce_sparse_pooling_input(dict(img_index=np.array(rimg, dtype=np.float64), img_size=f["img_size"], bv_index=rbv,
                                                     bv_size=[200, 240]), M_val=rmv, stride=[8, 2])
        refs.append((o_ref, np.asarray(rmv, dtype=np.float32)))
    plan = shpl.build_pairs_plan(img_index, bv_index, frames[0]["img_size"], [200, 240], stride=(8, 2), m_val=m_val)
    assert plan.frames == B and plan.nnz == [len(r[1]) for r in refs] and sum(plan.n_oob) == 0
    bev = rng.standard_normal((B, 100, 120, 16), dtype=np.float32)
    img = rng.standard_normal((B, 48, 160, C), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused = shpl.sparse_pool(tb, ti, plan)
    g = rng.standard_normal(tuple(fused.shape), dtype=np.float32)
    fused.backward(torch.from_numpy(g).cuda())
    for k, (o_ref, val) in enumerate(refs):
        np.testing.assert_array_equal(fused[k].detach().cpu().numpy(),
                                      cref.forward(bev[k], img[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"]))
        gd, gs = cref.backward(g[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"], 16, (48, 160, C))
        np.testing.assert_array_equal(tb.grad[k].cpu().numpy(), gd)
        np.testing.assert_array_equal(ti.grad[k].cpu().numpy(), gs)
