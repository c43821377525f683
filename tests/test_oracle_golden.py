"""Pins the CPU oracle (numpy restatement AND the plain-C port) against what the
REFERENCE's own code produced (tests/golden/*.npz, made by oracle/gen_goldens.py
from /root/reference) and against the known-answer vectors of SURVEY.md
Appendix B.  No GPU needed."""
import hashlib
import os

import numpy as np
import pytest

from oracle import cref, index_oracle as io, synth, value_oracle as vo


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ------------------------------------------------------------------ KAT-1 / 2
def test_kat1_expected_literals(golden_dir):
    """SURVEY.md Appendix B KAT-1, literal expectations + the reference-run fixture."""
    g = load(golden_dir, "kat1.npz")
    d = io.gen_sparse_pooling_input_avod(g["points"], g["voxel_indices"], g["P"], list(g["im_size"]), tuple(g["bv_size"]))
    assert d["img_index"][0].tolist() == [0, 2, 2, 0, 9, 8]          # ties to even: 0.5->0 1.5->2 2.5->2
    assert d["img_index"][1].tolist() == [0, 2, 2, 0, 3, 5]
    assert d["bv_index"].tolist() == [[0, 1], [1, 2], [2, 3], [3, 4], [5, 6], [7, 8]]
    np.testing.assert_array_equal(d["bv_index"], g["gen_bv_index"])
    np.testing.assert_array_equal(d["img_index"], g["gen_img_index"])
    assert d["img_index"].dtype == np.float64 and d["img_index"].shape[0] == 3
    o = io.produce_sparse_pooling_input(d, stride=[2, 2])
    assert o["M_size"].tolist() == [32, 5]
    assert o["Mij_pool"].tolist() == [[0, 0], [8, 1], [9, 2], [17, 3], [26, 4]]
    assert o["img_index_flip_pool"].tolist() == [[0, 0, 0], [0, 1, 1], [0, 1, 1], [0, 0, 0], [0, 1, 4]]
    for k in ("Mij_pool", "M_val", "M_size", "img_index_flip_pool"):
        np.testing.assert_array_equal(o[k], g[k])
        assert o[k].dtype == g[k].dtype, k
    np.testing.assert_array_equal(d["img_index"], g["img_index_after"])   # in-place mutation (quirk A.4-6)
    assert o["bev_index_flip_pool"].shape == (0, 3)


def test_kat1_c_oracle(golden_dir):
    g = load(golden_dir, "kat1.npz")
    c = cref.build_avod(g["points"], g["voxel_indices"], g["P"], g["im_size"], g["bv_size"], g["stride"])
    np.testing.assert_array_equal(c["Mij_pool"], g["Mij_pool"])
    np.testing.assert_array_equal(c["img_index_flip_pool"], g["img_index_flip_pool"])
    np.testing.assert_array_equal(c["M_size"], g["M_size"])
    np.testing.assert_array_equal(c["gen"]["bv_index"], g["gen_bv_index"])
    np.testing.assert_array_equal(c["gen"]["img_index"], g["gen_img_index"])


def test_kat2_row_filter_and_x_wrap(golden_dir):
    g = load(golden_dir, "kat2.npz")
    d = {k[3:]: np.array(g[k]) for k in g.files if k.startswith("in_")}
    o = io.produce_sparse_pooling_input(d, stride=[1, 1])
    assert o["Mij_pool"].tolist() == [[115, 0], [127, 1], [112, 2]]   # row 131 dropped; x=16 wraps
    for k in ("Mij_pool", "M_val", "M_size", "img_index_flip_pool"):
        np.testing.assert_array_equal(o[k], g[k])


def test_wrong_img_index_shape_asserts():
    d = synth.direct_pairs(0, 10)
    d["img_index"] = d["img_index"][:2]
    with pytest.raises(AssertionError):
        io.produce_sparse_pooling_input(d)


# ------------------------------------------------- reference-run frame fixtures
@pytest.mark.parametrize("seed,az", [(1, 0.09), (2, 0.05)])
def test_avod_frame_matches_reference(golden_dir, seed, az):
    g = load(golden_dir, "avod_frame_seed%d.npz" % seed)
    frame = synth.avod_frame(seed, az_step_deg=az)
    assert digest(frame["points"], frame["voxel_indices"]) == str(g["input_sha"]), "synthetic input drifted"
    for s in (1, 4, 8):
        d = io.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], frame["P"], frame["im_size"], frame["bv_size"])
        if s == 1:
            np.testing.assert_array_equal(d["bv_index"], g["gen_bv_index"])
            np.testing.assert_array_equal(d["img_index"], g["gen_img_index"].astype(np.float64))
        o = io.produce_sparse_pooling_input(d, stride=[s, s])
        np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["Mij_pool_s%d" % s])
        np.testing.assert_array_equal(o["Mij_pool"][:, 1], np.arange(len(o["Mij_pool"])))
        np.testing.assert_array_equal(o["M_size"], g["M_size_s%d" % s])
        np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip_s%d" % s])
        assert o["Mij_pool"].dtype == np.int64 and o["img_index_flip_pool"].dtype == np.int64
        # plain-C port (fma-chain projection) agrees bit for bit as well
        c = cref.build_avod(frame["points"], frame["voxel_indices"], frame["P"], frame["im_size"], frame["bv_size"], (s, s))
        np.testing.assert_array_equal(c["Mij_pool"][:, 0], g["Mij_pool_s%d" % s])
        np.testing.assert_array_equal(c["img_index_flip_pool"], g["flip_s%d" % s])
        np.testing.assert_array_equal(c["M_size"], g["M_size_s%d" % s])


@pytest.mark.parametrize("name,kw", [("direct_uniform", dict(seed=0, n=20000)),
                                     ("direct_ground", dict(seed=3, n=50000, skew="ground")),
                                     ("direct_zipf", dict(seed=4, n=30000, skew="zipf"))])
def test_direct_pairs_match_reference(golden_dir, name, kw):
    g = load(golden_dir, name + ".npz")
    d0 = synth.direct_pairs(**kw)
    assert digest(d0["bv_index"], d0["img_index"]) == str(g["input_sha"])
    for s in ((1, 1), (8, 8), (8, 2)):
        d = {k: np.array(v, copy=True) for k, v in d0.items()}
        o = io.produce_sparse_pooling_input(d, stride=list(s))
        tag = "s%d_%d" % s
        np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["row_" + tag])
        np.testing.assert_array_equal(o["M_size"], g["M_size_" + tag])
        np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip_" + tag])


def test_mv3d_feeder_weights_match_reference(golden_dir):
    g = load(golden_dir, "mv3d_seed5.npz")
    f = synth.mv3d_frame(seed=5, n_points=6000)
    assert digest(f["points_fsh"], f["img_index2"]) == str(g["input_sha"])
    inrange, kept, bv_index, m_val = io.mv3d_voxel_weights(f["points_fsh"], f["res"], f["zres"], f["side_range"],
                                                          f["fwd_range"], f["height_range"], f["max_points"])
    np.testing.assert_array_equal(bv_index, g["bv_index"])
    np.testing.assert_array_equal(m_val, g["M_val"])                    # 1/count, float64, bit-exact
    img2 = f["img_index2"][:, inrange][:, kept]
    np.testing.assert_array_equal(img2, g["img_index"][:2])
    assert f["bv_size"] == [int(g["voxel_full_size"][1]), int(g["voxel_full_size"][2])]
    img_index = np.vstack((img2, np.zeros((1, img2.shape[1])))).astype(np.float64)
    o = io.produce_sparse_pooling_input(dict(img_index=img_index, img_size=f["img_size"], bv_index=bv_index,
                                             bv_size=f["bv_size"]), M_val=m_val, stride=f["stride"])
    np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["row"])
    np.testing.assert_array_equal(o["M_size"], g["M_size"])
    np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip"])
    assert m_val.min() == 1.0 / 45


# ------------------------------------------------------------ canonical CSR
def test_plan_is_stable_sort_of_coo():
    d = synth.direct_pairs(7, 5000, bev_hw=(40, 50), img_wh=(60, 30))
    o = io.produce_sparse_pooling_input(d, stride=[1, 1])
    R, (Hs, Ws) = int(o["M_size"][0]), (30, 60)
    val = np.random.default_rng(0).random(len(o["Mij_pool"])).astype(np.float32)
    p = io.build_plan(o["Mij_pool"], val, o["img_index_flip_pool"], R, Hs, Ws)
    rows = o["Mij_pool"][:, 0]
    assert p["n_oob"] == 0 and p["row_ptr"][-1] == len(rows) == p["pix_ptr"][-1]
    for r in np.unique(rows)[:200]:
        seg = slice(p["row_ptr"][r], p["row_ptr"][r + 1])
        ks = p["csr_ent"][seg]
        np.testing.assert_array_equal(ks, np.nonzero(rows == r)[0])      # ascending k inside a row
        np.testing.assert_array_equal(p["csr_val"][seg], val[ks])
    pix = o["img_index_flip_pool"][:, 1] * Ws + o["img_index_flip_pool"][:, 2]
    np.testing.assert_array_equal(pix[p["csr_ent"]], p["csr_src"])
    for q in np.unique(pix)[:200]:
        seg = slice(p["pix_ptr"][q], p["pix_ptr"][q + 1])
        np.testing.assert_array_equal(p["csrT_ent"][seg], np.nonzero(pix == q)[0])
        np.testing.assert_array_equal(p["csrT_dst"][seg], rows[p["csrT_ent"][seg]])


def test_plan_drops_out_of_range_entries():
    Mij = np.array([[0, 0], [5, 1], [-1, 2], [2, 3], [1, 4]])
    flip = np.array([[0, 0, 0], [0, 1, 1], [0, 0, 1], [0, -1, 0], [0, 1, 3]])
    p = io.build_plan(Mij, np.ones(5), flip, n_rows=4, src_h=2, src_w=3)
    assert p["n_oob"] == 4                       # row 5>=4, row -1, v=-1, u=3>=W
    assert p["row_ptr"].tolist() == [0, 1, 1, 1, 1] and p["csr_src"].tolist() == [0]


# ---------------------------------------------------------------- value path
def _small_case(seed=0, dual=True):
    rng = np.random.default_rng(seed)
    Hb, Wb, Hi, Wi, Cb, Ci = 9, 11, 7, 13, 8, 12
    d = synth.direct_pairs(seed, 400, bev_hw=(Hb, Wb), img_wh=(Wi, Hi))
    o = io.produce_sparse_pooling_input(d, stride=[1, 1])
    val = rng.random(len(o["Mij_pool"])).astype(np.float32) if seed % 2 else np.ones(len(o["Mij_pool"]), np.float32)
    bev = rng.standard_normal((1, Hb, Wb, Cb), dtype=np.float32)
    img = rng.standard_normal((1, Hi, Wi, Ci), dtype=np.float32)
    return o, val, bev, img


@pytest.mark.parametrize("seed", [0, 1])
def test_value_oracle_numpy_vs_c_bitexact(seed):
    o, val, bev, img = _small_case(seed)
    M = (o["Mij_pool"], val, o["M_size"])
    flip = o["img_index_flip_pool"]
    bv_fused, img_fused = vo.sparse_pool_layer([bev, img], [img.shape[3], bev.shape[3]], M, flip, np.zeros((1, 3)))
    c_bv = cref.forward(bev[0], img[0], o["Mij_pool"], val, flip)
    c_img = cref.forward_trans(img[0], bev[0], o["Mij_pool"], val, flip)
    np.testing.assert_array_equal(bv_fused[0], c_bv)
    np.testing.assert_array_equal(img_fused[0], c_img)
    rng = np.random.default_rng(seed + 10)
    g_bv = rng.standard_normal(bv_fused.shape, dtype=np.float32)
    g_im = rng.standard_normal(img_fused.shape, dtype=np.float32)
    gb, gi = vo.sparse_pool_layer_grad([bev, img], None, M, flip, np.zeros((1, 3)), g_bv, g_im)
    gd1, gs1 = cref.backward(g_bv[0], o["Mij_pool"], val, flip, bev.shape[3], img.shape[1:])
    gi2, gb2 = cref.backward_trans(g_im[0], o["Mij_pool"], val, flip, img.shape[3], bev.shape[1:])
    np.testing.assert_array_equal(gb[0], gd1 + gb2)
    np.testing.assert_array_equal(gi[0], gi2 + gs1)


@pytest.mark.parametrize("seed", [0, 1])
def test_value_oracle_vs_torch_autograd(seed):
    """Cross-check of the restated TF semantics against an independent engine:
    torch CPU index_select / index_add / autograd (fp32 tolerance 1e-5 relative,
    the north_star's bound; summation order differs)."""
    torch = pytest.importorskip("torch")
    o, val, bev, img = _small_case(seed)
    flip = o["img_index_flip_pool"]
    rows = torch.from_numpy(o["Mij_pool"][:, 0].copy())
    Hb, Wb, Cb = bev.shape[1:]
    Hi, Wi, Ci = img.shape[1:]
    pix = torch.from_numpy(flip[:, 1] * Wi + flip[:, 2])
    w = torch.from_numpy(val)[:, None]
    tb = torch.from_numpy(bev).clone().requires_grad_(True)
    ti = torch.from_numpy(img).clone().requires_grad_(True)
    G = ti.reshape(-1, Ci).index_select(0, pix) * w
    Y = torch.zeros(Hb * Wb, Ci).index_add(0, rows, G)
    bv_fused = torch.cat([tb, Y.reshape(1, Hb, Wb, Ci)], dim=3)
    S = tb.reshape(-1, Cb).index_select(0, rows) * w
    Pm = torch.zeros(Hi * Wi, Cb).index_add(0, pix, S)
    img_fused = torch.cat([ti, Pm.reshape(1, Hi, Wi, Cb)], dim=3)
    M = (o["Mij_pool"], val, o["M_size"])
    o_bv, o_img = vo.sparse_pool_layer([bev, img], [Ci, Cb], M, flip, np.zeros((1, 3)))
    np.testing.assert_allclose(o_bv, bv_fused.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(o_img, img_fused.detach().numpy(), rtol=1e-5, atol=1e-5)
    rng = np.random.default_rng(seed + 10)
    g_bv = rng.standard_normal(o_bv.shape, dtype=np.float32)
    g_im = rng.standard_normal(o_img.shape, dtype=np.float32)
    (bv_fused * torch.from_numpy(g_bv)).sum().add((img_fused * torch.from_numpy(g_im)).sum()).backward()
    gb, gi = vo.sparse_pool_layer_grad([bev, img], None, M, flip, np.zeros((1, 3)), g_bv, g_im)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(gi, ti.grad.numpy(), rtol=1e-5, atol=2e-5)


def test_value_oracle_rejects_out_of_range_like_tf_cpu():
    x = np.zeros((1, 4, 5, 2), np.float32)
    with pytest.raises(IndexError):
        vo.gather_nd(x, np.array([[0, 4, 0]]))
    with pytest.raises(IndexError):
        vo.spmm(np.array([[9, 0]]), np.ones(1), [4, 1], np.zeros((1, 2), np.float32))


# ------------------------------------------------------------ feeder
@pytest.mark.parametrize("seed,az", [(1, 0.09), (2, 0.05), (3, None)])
def test_feeder_oracle_matches_reference_bev_slices(golden_dir, seed, az):
    """oracle/feeder_oracle.py against BevSlices.generate_bev(output_indices=True) of the reference."""
    from oracle import feeder_oracle as fo
    g = load(golden_dir, "bev_slices_seed%d.npz" % seed)
    pts = synth.lidar_scan_gappy(seed) if az is None else synth.lidar_scan(seed, az_step_deg=az)   # seed 3: slice re-use quirk
    assert digest(pts) == str(g["input_sha"])
    hms, dm, idx, upts = fo.generate_bev(pts.T, np.array([0.0, -1.0, 0.0, 1.65]), synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                         -0.2, 2.3, 5)
    np.testing.assert_array_equal(idx, g["voxel_indices"])
    np.testing.assert_array_equal(upts, g["unique_pts"])
    for i, hm in enumerate(hms):
        assert hm.shape == (700, 800)
        nz = np.nonzero(hm)
        np.testing.assert_array_equal(np.stack(nz, axis=1), g["hm%d_idx" % i])
        np.testing.assert_array_equal(hm[nz], g["hm%d_val" % i])
    nz = np.nonzero(dm)
    np.testing.assert_array_equal(np.stack(nz, axis=1), g["dm_idx"])
    np.testing.assert_array_equal(dm[nz], g["dm_val"])
    if az is None:
        return
    # the stand-in the other fixtures use (synth.one_point_per_cell) is the same feeder for this scan
    p2, i2 = synth.one_point_per_cell(pts)
    np.testing.assert_array_equal(i2, idx)
    np.testing.assert_array_equal(p2, upts)


@pytest.mark.parametrize("name,seed,n,kw", [("mv3d_seed5", 5, 6000, {}), ("mv3d_car_seed6", 6, 9000, dict(car=True))])
def test_mv3d_feeder_oracle_matches_reference(golden_dir, name, seed, n, kw):
    """oracle/feeder_oracle.point_cloud_2_top_sparse against the reference's construct_voxel.py output:
    pairs, weights, voxel coordinates / counts, and the [V, T, 7] feature buffer (by digest), bit for bit."""
    from oracle import feeder_oracle as fo
    g = load(golden_dir, name + ".npz")
    f = synth.mv3d_frame(seed=seed, n_points=n, **kw)
    assert digest(f["points_fsh"], f["img_index2"]) == str(g["input_sha"])
    vd, vfs, img_index, bv_index, m_val = fo.point_cloud_2_top_sparse(
        synth.mv3d_cam4(f), f["img_index2"], f["res"], f["zres"], f["side_range"], f["fwd_range"], f["height_range"],
        f["max_points"])
    np.testing.assert_array_equal(vfs, g["voxel_full_size"])
    np.testing.assert_array_equal(img_index, g["img_index"])
    np.testing.assert_array_equal(bv_index, g["bv_index"])
    np.testing.assert_array_equal(m_val, g["M_val"])
    np.testing.assert_array_equal(vd["coordinate_buffer"], g["coordinate_buffer"])
    np.testing.assert_array_equal(vd["number_buffer"], g["number_buffer"])
    np.testing.assert_array_equal(vd["feature_buffer"][:64], g["feature_buffer_head"])
    assert digest(vd["feature_buffer"]) == str(g["feature_buffer_sha"])
    assert vd["number_buffer"].max() == f["max_points"]          # the cap bites
    # and it agrees with the weights-only restatement the other tests use
    inrange, kept, bv2, mv2 = io.mv3d_voxel_weights(f["points_fsh"], f["res"], f["zres"], f["side_range"],
                                                    f["fwd_range"], f["height_range"], f["max_points"])
    np.testing.assert_array_equal(bv2, bv_index)
    np.testing.assert_array_equal(mv2, m_val)


# ------------------------------------------------------------ point-cloud ingest
@pytest.mark.parametrize("seed,az", [(1, 0.4), (2, 0.15)])
def test_ingest_oracle_matches_reference_get_lidar_point_cloud(golden_dir, seed, az):
    """oracle/feeder_oracle.get_lidar_point_cloud against obj_utils.get_lidar_point_cloud of the reference, which read
    the same scan and calibration from KITTI-format files (oracle/gen_goldens.ingest_goldens)."""
    from oracle import feeder_oracle as fo
    g = load(golden_dir, "lidar_ingest_seed%d.npz" % seed)
    scan = synth.velodyne_scan(seed, az_step_deg=az)
    assert digest(scan) == str(g["input_sha"])
    # the calibration went through the text file: use the parsed values
    pc = fo.get_lidar_point_cloud(scan, g["p2"], g["r0_rect"], g["tr_velodyne_to_cam"], im_size=list(g["im_size"]))
    assert pc.shape == (3, int(g["n_fov"])) and pc.dtype == np.float64
    assert digest(pc) == str(g["fov_sha"])
    if "fov_points" in g:
        np.testing.assert_array_equal(pc, g["fov_points"])
    pc_all = fo.get_lidar_point_cloud(scan, g["p2"], g["r0_rect"], g["tr_velodyne_to_cam"])
    assert pc_all.shape == (3, len(scan)) and digest(pc_all) == str(g["all_sha"])
    # the reference's min_intensity branch indexes the z-filtered cloud with an unfiltered intensity mask (obj_utils.py:266)
    with pytest.raises((ValueError, IndexError)):
        fo.get_lidar_point_cloud(scan, g["p2"], g["r0_rect"], g["tr_velodyne_to_cam"], im_size=list(g["im_size"]), min_intensity=0.5)


def test_voxel_scatter_oracle_vs_torch_autograd():
    """The VFE scatter oracle (tf.scatter_nd, group_pointcloud.py:84-85) against torch CPU: index_put_ with
    accumulate adds duplicates in index order too, and autograd's gradient of it is the gather."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    B, grid, C, K = 2, (3, 4, 5), 8, 300
    coord = np.stack([rng.integers(0, B, K)] + [rng.integers(0, g, K) for g in grid], axis=1)
    vw = rng.standard_normal((K, C)).astype(np.float32)
    out = vo.voxel_scatter(coord, vw, (B, *grid, C))
    x = torch.from_numpy(vw).requires_grad_(True)
    ref = torch.zeros((B, *grid, C)).index_put(tuple(torch.from_numpy(coord[:, j]) for j in range(4)), x, accumulate=True)
    np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-6, atol=1e-6)
    g = rng.standard_normal(out.shape).astype(np.float32)
    ref.backward(torch.from_numpy(g))
    np.testing.assert_array_equal(vo.voxel_scatter_grad(coord, g), x.grad.numpy())
    # known answer: two rows on one cell, one row outside the grid
    c = np.array([[0, 1, 2, 3], [0, 1, 2, 3], [0, 3, 0, 0]])
    v = np.array([[1, 2, 3, 4], [10, 20, 30, 40], [5, 5, 5, 5]], dtype=np.float32)
    with pytest.raises(IndexError):
        vo.voxel_scatter(c, v, (1, *grid, 4))
    o = vo.voxel_scatter(c, v, (1, *grid, 4), strict=False)
    assert o[0, 1, 2, 3].tolist() == [11, 22, 33, 44] and o.sum() == 110
    assert vo.voxel_scatter_grad(c, np.ones((1, *grid, 4), np.float32)).tolist() == [[1] * 4, [1] * 4, [0] * 4]


def test_augment_oracle_matches_reference_hooks(golden_dir):
    """The augmentation-hook restatements against what the reference's own functions returned here
    (oracle/gen_goldens.py augment_goldens: kitti_aug.flip_* imported, augment_voxel / augment_fv executed from source)."""
    from oracle import feeder_oracle as fo
    g = load(golden_dir, "augment_hooks.npz")
    pts = synth.lidar_scan(1, az_step_deg=0.09).T
    assert digest(pts) == str(g["flip_input_sha"])
    fl = fo.flip_point_cloud(pts)
    assert digest(fl) == str(g["flip_sha"])
    np.testing.assert_array_equal(fl[:, :16], g["flip_head"])
    f = synth.mv3d_frame(seed=7, n_points=5000)
    pc = synth.mv3d_cam4(f)
    assert digest(pc) == str(g["voxel_input_sha"])
    sx, sz, ratio, angle = g["voxel_params"]
    np.testing.assert_array_equal(fo.mv3d_project_round(pc, synth.P2_KITTI), g["voxel_img_index2"])
    aug = fo.mv3d_augment_points(pc, sx, sz, np.array([ratio]), np.array([angle]))
    assert digest(aug) == str(g["voxel_pc_sha"])
    np.testing.assert_array_equal(aug[:16], g["voxel_pc_head"])
    img_index = np.vstack((f["img_index2"], np.zeros((1, f["img_index2"].shape[1]), dtype=int)))
    assert digest(img_index) == str(g["fv_input_sha"])
    fsx, fsy, fratio = g["fv_params"]
    np.testing.assert_array_equal(fo.augment_fv_index(img_index, fsx, fsy, np.array([fratio])), g["fv_img_index"])


@pytest.mark.parametrize("seed", [0, 1])
def test_value_oracle_vs_scipy_sparse_matmul(seed):
    """A third engine for the restated TF semantics (SURVEY.md 8c suggests it): M as a scipy COO matrix,
    pooled = M @ gather(img) and, transposed, M^T @ bev scattered to the pixels -- fp64 accumulation in scipy, so the
    fp32 oracle must agree to the north_star's 1e-5 of the term sum."""
    sp = pytest.importorskip("scipy.sparse")
    o, val, bev, img = _small_case(seed)
    flip, Mij = o["img_index_flip_pool"], o["Mij_pool"]
    Hb, Wb, Cb = bev.shape[1:]
    Hi, Wi, Ci = img.shape[1:]
    R, n = int(o["M_size"][0]), int(o["M_size"][1])
    M = sp.coo_matrix((val.astype(np.float64), (Mij[:, 0], Mij[:, 1])), shape=(R, n)).tocsr()
    pix = flip[:, 1] * Wi + flip[:, 2]
    G = img.reshape(-1, Ci).astype(np.float64)[pix]                        # gather_nd
    Y = M @ G                                                             # sparse_tensor_dense_matmul
    S = M.T @ bev.reshape(-1, Cb).astype(np.float64)                      # sparse_transpose + matmul: [n, Cb]
    Pm = np.zeros((Hi * Wi, Cb))
    np.add.at(Pm, pix, S)                                                 # scatter_nd sums duplicates
    o_bv, o_img = vo.sparse_pool_layer([bev, img], [Ci, Cb], (Mij, val, o["M_size"]), flip, np.zeros((1, 3)))
    absM = abs(M)
    tol_bv = 1e-5 * (absM @ np.abs(G)).max() + 1e-12
    tol_img = 1e-5 * np.abs(Pm).max() + 1e-6
    assert np.abs(o_bv[0, ..., Cb:].reshape(-1, Ci) - Y).max() <= tol_bv
    assert np.abs(o_img[0, ..., Ci:].reshape(-1, Cb) - Pm).max() <= max(tol_img, tol_bv)
    np.testing.assert_array_equal(o_bv[0, ..., :Cb], bev[0])              # the concatenated halves are the inputs themselves
    np.testing.assert_array_equal(o_img[0, ..., :Ci], img[0])


def test_conv_oracle_agrees_with_torch_cpu_conv2d():
    """oracle.value_oracle.conv3x3_same (the f3 oracle: slim.conv2d's SAME-padded, stride-1 3x3 conv in HWIO layout,
    rpn_model.py:338-346) against an independent engine, torch CPU conv2d in float64."""
    import torch
    from oracle import value_oracle as vo
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 11, 6))
    w = rng.standard_normal((3, 3, 6, 5))
    y, mag = vo.conv3x3_same(x, w)
    t = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(w).permute(3, 2, 0, 1), padding=1)
    np.testing.assert_allclose(y, t.permute(0, 2, 3, 1).numpy(), rtol=1e-12, atol=1e-12)
    assert (mag >= np.abs(y) - 1e-12).all()
    # a hand-checkable case: all-ones 3x3x1x1 kernel = count of in-image neighbours times the value
    y1, _ = vo.conv3x3_same(np.ones((1, 3, 3, 1)), np.ones((3, 3, 1, 1)))
    assert y1[0, :, :, 0].tolist() == [[4, 6, 4], [6, 9, 6], [4, 6, 4]]
    ya, _ = vo.conv3x3_after_fusion(-np.ones((1, 3, 3, 1)), np.ones((3, 3, 1, 1)), scale=[2.0], shift=[1.0], relu=True)
    assert (ya == 0).all()


def test_value_path_exact_arithmetic_kat_numpy_and_c_oracles():
    """tests/kat_value.py: hand-derived expected values on inputs where every summation order gives the same bits --
    the limit of what can be pinned for the value path without TensorFlow.  Both oracles must reproduce the literals."""
    from tests import kat_value as K
    M = (K.MIJ, K.VAL, K.M_SIZE)
    bv, im = vo.sparse_pool_layer([K.BEV, K.IMG], [2, 2], M, img_index_flip=K.FLIP, bv_index=np.zeros((1, 3)))
    np.testing.assert_array_equal(bv, K.FUSED_BEV)
    np.testing.assert_array_equal(im, K.FUSED_IMG)
    bv1, im1 = vo.sparse_pool_layer([K.BEV, K.IMG], [2, 2], M, img_index_flip=K.FLIP, bv_index=None)
    np.testing.assert_array_equal(bv1, K.FUSED_BEV)
    assert im1 is K.IMG
    gb, gi = vo.sparse_pool_layer_grad([K.BEV, K.IMG], [2, 2], M, K.FLIP, None, K.G_FUSED_BEV, np.zeros_like(K.IMG))   # img passes through
    np.testing.assert_array_equal(gb, K.G_BEV_SINGLE)
    np.testing.assert_array_equal(gi, K.G_IMG_SINGLE)
    gb, gi = vo.sparse_pool_layer_grad([K.BEV, K.IMG], [2, 2], M, K.FLIP, np.zeros((1, 3)), K.G_FUSED_BEV, K.G_FUSED_IMG)
    np.testing.assert_array_equal(gb, K.G_BEV_DUAL)
    np.testing.assert_array_equal(gi, K.G_IMG_DUAL)
    # the plain-C oracle
    np.testing.assert_array_equal(cref.forward(K.BEV[0], K.IMG[0], K.MIJ, K.VAL, K.FLIP), K.FUSED_BEV[0])
    np.testing.assert_array_equal(cref.forward_trans(K.IMG[0], K.BEV[0], K.MIJ, K.VAL, K.FLIP), K.FUSED_IMG[0])
    gd, gs = cref.backward(K.G_FUSED_BEV[0], K.MIJ, K.VAL, K.FLIP, 2, (2, 3, 2))
    np.testing.assert_array_equal(gd, K.G_BEV_SINGLE[0])
    np.testing.assert_array_equal(gs, K.G_IMG_SINGLE[0])
    gi2, gb2 = cref.backward_trans(K.G_FUSED_IMG[0], K.MIJ, K.VAL, K.FLIP, 2, (2, 2, 2))
    np.testing.assert_array_equal(gd + gb2, K.G_BEV_DUAL[0])
    np.testing.assert_array_equal(gi2 + gs, K.G_IMG_DUAL[0])


def test_conv_gradient_oracle_agrees_with_torch_autograd():
    import torch
    from oracle import value_oracle as vo
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 7, 9, 5))
    w = rng.standard_normal((3, 3, 5, 4))
    g = rng.standard_normal((2, 7, 9, 4))
    g_x, g_w, mag_x, mag_w = vo.conv3x3_same_grad(x, w, g)
    tx = torch.from_numpy(x).permute(0, 3, 1, 2).requires_grad_(True)
    tw = torch.from_numpy(w).permute(3, 2, 0, 1).requires_grad_(True)
    y = torch.nn.functional.conv2d(tx, tw, padding=1)
    y.backward(torch.from_numpy(g).permute(0, 3, 1, 2))
    np.testing.assert_allclose(g_x, tx.grad.permute(0, 2, 3, 1).numpy(), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(g_w, tw.grad.permute(2, 3, 1, 0).numpy(), rtol=1e-11, atol=1e-11)
    assert (mag_x >= np.abs(g_x) - 1e-9).all() and (mag_w >= np.abs(g_w) - 1e-9).all()
