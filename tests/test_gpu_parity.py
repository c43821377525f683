"""GPU parity tests: the CUDA path (through the ctypes C ABI of libshpl.so) against
the CPU oracle and the reference-run golden fixtures.  Index / CSR work must be
bit-exact; pooled features and gradients are compared bit-exact too where the
summation order equals the oracle's, and within 1e-5 relative (the north_star's
fp32 bound) otherwise.  Run with `pytest -m gpu` on a B200."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import cref, index_oracle as io, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


class Calib:
    def __init__(self, p2):
        self.p2 = p2


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def plan_arrays(plan, f=0):
    """Host copies of one frame's CSR / CSR^T (entry offsets made frame-local)."""
    c = plan.counts.cpu().numpy()
    R, Q = plan.rows_per_frame, plan.src_per_frame
    row_ptr = plan.row_ptr.cpu().numpy()[f * R:(f + 1) * R + 1]
    pix_ptr = plan.pix_ptr.cpu().numpy()[f * Q:(f + 1) * Q + 1]
    lo, hi = int(row_ptr[0]), int(row_ptr[-1])
    return dict(row_ptr=row_ptr - lo, pix_ptr=pix_ptr - lo,
                csr_src=plan.csr_src.cpu().numpy()[lo:hi] - f * Q, csr_val=plan.csr_val.cpu().numpy()[lo:hi],
                csr_row=plan.csr_row.cpu().numpy()[lo:hi] - f * R, csrT_pix=plan.csrT_pix.cpu().numpy()[lo:hi] - f * Q,
                csrT_dst=plan.csrT_dst.cpu().numpy()[lo:hi] - f * R, csrT_val=plan.csrT_val.cpu().numpy()[lo:hi],
                counts=c[f])


def assert_plan_equals_oracle(plan, Mij, val, flip, n_rows, src_hw, f=0):
    ref = io.build_plan(Mij, val, flip, n_rows, src_hw[0], src_hw[1])
    got = plan_arrays(plan, f)
    for k in ("row_ptr", "pix_ptr", "csr_row", "csr_src", "csr_val", "csrT_pix", "csrT_dst", "csrT_val"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    assert int(got["counts"][2]) == ref["n_oob"]
    assert int(got["counts"][3]) == len(ref["csr_src"])


# ------------------------------------------------------------------ builder: KATs
def test_kat1_through_the_dropin_api(shpl, golden_dir):
    g = load(golden_dir, "kat1.npz")
    d = shpl.gen_sparse_pooling_input_avod(g["points"], g["voxel_indices"], Calib(g["P"]), list(g["im_size"]), tuple(g["bv_size"]))
    assert isinstance(d["img_index"], np.ndarray) and d["img_index"].dtype == np.float64
    np.testing.assert_array_equal(d["bv_index"], g["gen_bv_index"])
    np.testing.assert_array_equal(d["img_index"], g["gen_img_index"])
    assert d["img_index"][0].tolist() == [0, 2, 2, 0, 9, 8]       # 0.5->0, 1.5->2, 2.5->2, -0.0 kept, -1e-9 dropped
    o = shpl.produce_sparse_pooling_input(d, stride=[2, 2])
    for k in ("Mij_pool", "M_val", "M_size", "img_index_flip_pool"):
        np.testing.assert_array_equal(o[k], g[k], err_msg=k)
        assert o[k].dtype == g[k].dtype, k
    np.testing.assert_array_equal(d["img_index"], g["img_index_after"])      # in-place mutation reproduced
    assert o["bev_index_flip_pool"].shape == (0, 3)
    assert_plan_equals_oracle(o["shpl_plan"], g["Mij_pool"], np.ones(5), g["img_index_flip_pool"], 32, (3, 5))


def test_kat2_row_filter_and_x_wrap(shpl, golden_dir):
    g = load(golden_dir, "kat2.npz")
    d = {k[3:]: np.array(g[k]) for k in g.files if k.startswith("in_")}
    o = shpl.produce_sparse_pooling_input(d, stride=[1, 1])
    assert o["Mij_pool"].tolist() == [[115, 0], [127, 1], [112, 2]]
    for k in ("Mij_pool", "M_val", "M_size", "img_index_flip_pool"):
        np.testing.assert_array_equal(o[k], g[k], err_msg=k)


def test_wrong_img_index_shape_asserts(shpl):
    d = synth.direct_pairs(0, 10)
    d["img_index"] = d["img_index"][:2]
    with pytest.raises(AssertionError):
        shpl.produce_sparse_pooling_input(d)


def test_cpu_feature_maps_are_rejected(shpl):
    d = synth.direct_pairs(0, 50, bev_hw=(8, 8), img_wh=(8, 8))
    o = shpl.produce_sparse_pooling_input(d)
    with pytest.raises(RuntimeError):
        shpl.sparse_pool_layer([torch.zeros(1, 8, 8, 4), torch.zeros(1, 8, 8, 4)], [4, 4], o,
                               img_index_flip=o["img_index_flip_pool"])


# ------------------------------------------- builder: reference-run frame fixtures
@pytest.mark.parametrize("seed,az", [(1, 0.09), (2, 0.05)])
def test_avod_frame_matches_reference_fixture(shpl, golden_dir, seed, az):
    g = load(golden_dir, "avod_frame_seed%d.npz" % seed)
    frame = synth.avod_frame(seed, az_step_deg=az)
    for s in (1, 4, 8):
        d = shpl.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], Calib(frame["P"]),
                                               frame["im_size"], frame["bv_size"])
        if s == 1:
            np.testing.assert_array_equal(d["bv_index"], g["gen_bv_index"])
            np.testing.assert_array_equal(d["img_index"], g["gen_img_index"].astype(np.float64))
        o = shpl.produce_sparse_pooling_input(d, stride=[s, s])
        np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["Mij_pool_s%d" % s])
        np.testing.assert_array_equal(o["Mij_pool"][:, 1], np.arange(len(o["Mij_pool"])))
        np.testing.assert_array_equal(o["M_size"], g["M_size_s%d" % s])
        np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip_s%d" % s])
        assert o["Mij_pool"].dtype == np.int64 and o["img_index_flip_pool"].dtype == np.int64
        R = int(o["M_size"][0])
        assert_plan_equals_oracle(o["shpl_plan"], o["Mij_pool"], np.ones(len(o["Mij_pool"])), o["img_index_flip_pool"],
                                  R, (360 // s, 1200 // s))
        # fused device-resident builder gives the same COO and plan
        plan, coo = shpl.build_avod_plan(frame["points"], frame["voxel_indices"], frame["P"], frame["im_size"],
                                         frame["bv_size"], stride=(s, s), want_coo=True)
        nnz = plan.nnz[0]
        np.testing.assert_array_equal(coo[0][0][:nnz].cpu().numpy(), o["Mij_pool"])
        np.testing.assert_array_equal(coo[0][1][:nnz].cpu().numpy(), o["img_index_flip_pool"])
        np.testing.assert_array_equal(coo[0][3].cpu().numpy(), o["M_size"])
        assert_plan_equals_oracle(plan, o["Mij_pool"], np.ones(nnz), o["img_index_flip_pool"], R, (360 // s, 1200 // s))


@pytest.mark.parametrize("name,kw", [("direct_uniform", dict(seed=0, n=20000)),
                                     ("direct_ground", dict(seed=3, n=50000, skew="ground")),
                                     ("direct_zipf", dict(seed=4, n=30000, skew="zipf"))])
def test_direct_pairs_match_reference_fixture(shpl, golden_dir, name, kw):
    g = load(golden_dir, name + ".npz")
    d0 = synth.direct_pairs(**kw)
    for s in ((1, 1), (8, 8), (8, 2)):
        d = {k: np.array(v, copy=True) for k, v in d0.items()}
        o = shpl.produce_sparse_pooling_input(d, stride=list(s))
        tag = "s%d_%d" % s
        np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["row_" + tag])
        np.testing.assert_array_equal(o["M_size"], g["M_size_" + tag])
        np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip_" + tag])
        R = int(o["M_size"][0])
        assert_plan_equals_oracle(o["shpl_plan"], o["Mij_pool"], np.ones(len(o["Mij_pool"])), o["img_index_flip_pool"],
                                  R, (360 // s[0], 1200 // s[0]))


def test_mv3d_weights_and_strides(shpl, golden_dir):
    """MV3D entry (MV3D_voxel_train.py:89-91): external non-homogeneous M_val = 1/count, stride [8,2]."""
    g = load(golden_dir, "mv3d_seed5.npz")
    img_index = np.vstack((g["img_index"][:2], np.zeros((1, g["img_index"].shape[1])))).astype(np.float64)
    d = dict(img_index=img_index, img_size=np.array([1280, 384]), bv_index=g["bv_index"].astype(np.int64), bv_size=[200, 240])
    o = shpl.produce_sparse_pooling_input(d, M_val=g["M_val"], stride=[8, 2])
    np.testing.assert_array_equal(o["Mij_pool"][:, 0], g["row"])
    np.testing.assert_array_equal(o["M_size"], g["M_size"])
    np.testing.assert_array_equal(o["img_index_flip_pool"], g["flip"])
    assert o["M_val"] is not None and np.array_equal(o["M_val"], g["M_val"])
    assert_plan_equals_oracle(o["shpl_plan"], o["Mij_pool"], g["M_val"].astype(np.float32), o["img_index_flip_pool"],
                              12000, (48, 160))


def test_torch_cuda_inputs_stay_on_device(shpl):
    d0 = synth.direct_pairs(11, 3000, bev_hw=(64, 80), img_wh=(96, 40))
    ref = io.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d0.items()}, stride=[2, 2])
    d = dict(bv_index=torch.from_numpy(d0["bv_index"]).cuda(), img_index=torch.from_numpy(d0["img_index"]).cuda(),
             bv_size=d0["bv_size"], img_size=d0["img_size"])
    o = shpl.produce_sparse_pooling_input(d, stride=[2, 2])
    assert o["Mij_pool"].is_cuda and o["img_index_flip_pool"].is_cuda
    np.testing.assert_array_equal(o["Mij_pool"].cpu().numpy(), ref["Mij_pool"])
    np.testing.assert_array_equal(o["img_index_flip_pool"].cpu().numpy(), ref["img_index_flip_pool"])
    ref_after = {k: np.array(v, copy=True) for k, v in d0.items()}
    io.produce_sparse_pooling_input(ref_after, stride=[2, 2])
    np.testing.assert_array_equal(d["img_index"].cpu().numpy(), ref_after["img_index"])   # mutated in place on the device


# --------------------------------------------------------------- builder: edge cases
def test_empty_frame(shpl):
    d = dict(bv_index=np.zeros((0, 2), dtype=np.int64), img_index=np.zeros((3, 0)), bv_size=np.array([8, 8]), img_size=np.array([8, 8]))
    o = shpl.produce_sparse_pooling_input(d)
    assert o["Mij_pool"].shape == (0, 2) and o["M_size"].tolist() == [64, 0]
    p = plan_arrays(o["shpl_plan"])
    assert (p["row_ptr"] == 0).all() and (p["pix_ptr"] == 0).all()
    bev = torch.randn(1, 8, 8, 8, device="cuda")
    img = torch.randn(1, 8, 8, 4, device="cuda")
    fused, _ = shpl.sparse_pool_layer([bev, img], [4, 8], o, img_index_flip=o["img_index_flip_pool"])
    assert torch.equal(fused[..., :8], bev) and (fused[..., 8:] == 0).all()


def test_all_points_outside_image(shpl):
    pts = np.array([[100.0, 0.0, 1.0], [0.0, 100.0, 1.0], [-100.0, 0.0, 1.0]])
    d = shpl.gen_sparse_pooling_input_avod(pts, np.zeros((3, 2), dtype=np.int64), Calib(synth.P2_KITTI), [1200, 360], (700, 800))
    assert d["img_index"].shape == (3, 0) and d["bv_index"].shape == (0, 2)


def test_out_of_range_indices_are_counted_not_read(shpl):
    """Negative pixels / rows (possible after MV3D's augment_fv, quirk A.4-10): TF-CPU raises;
    here the COO is reproduced bit-exact, the CSR leaves the entries out and the layer raises."""
    d = dict(bv_index=np.array([[0, 0], [1, 0], [-5, 0], [2, 1]], dtype=np.int64),
             img_index=np.array([[1.0, -3.0, 2.0, 7.0], [1.0, 1.0, 1.0, -1.0], [0, 0, 0, 0]]),
             bv_size=np.array([4, 4]), img_size=np.array([8, 8]))
    ref = io.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()})
    o = shpl.produce_sparse_pooling_input(d)
    np.testing.assert_array_equal(o["Mij_pool"], ref["Mij_pool"])
    np.testing.assert_array_equal(o["img_index_flip_pool"], ref["img_index_flip_pool"])
    plan = o["shpl_plan"]
    assert plan.n_oob == [3]
    assert_plan_equals_oracle(plan, ref["Mij_pool"], np.ones(4), ref["img_index_flip_pool"], 16, (8, 8))
    bev = torch.randn(1, 4, 4, 4, device="cuda")
    img = torch.randn(1, 8, 8, 4, device="cuda")
    with pytest.raises(ValueError):
        shpl.sparse_pool_layer([bev, img], [4, 4], o, img_index_flip=o["img_index_flip_pool"])


def test_plan_from_arbitrary_coo(shpl):
    """M handed over as a bare tf.SparseTensor-like triple + int32 gather index (rpn_model.py:219-242)."""
    rng = np.random.default_rng(5)
    m, R, Hs, Ws = 7000, 900, 20, 30
    Mij = np.stack((rng.integers(0, R, m), rng.permutation(m)), axis=1).astype(np.int64)
    flip = np.stack((np.zeros(m, np.int64), rng.integers(0, Hs, m), rng.integers(0, Ws, m)), axis=1)
    val = rng.random(m).astype(np.float32)
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), np.array([R, m]))
    from sparse_pooling_b200 import sparse_pool_utils as spu
    plan = spu._resolve_plan(M, torch.from_numpy(flip.astype(np.int32)).cuda(), R, (Hs, Ws), torch.device("cuda"))
    assert_plan_equals_oracle(plan, Mij, val, flip, R, (Hs, Ws))


def test_stacked_batch_plan(shpl):
    frames = [synth.avod_frame(s, az_step_deg=0.4) for s in (3, 4, 5)]
    plan = shpl.build_avod_plan([f["points"] for f in frames], [f["voxel_indices"] for f in frames],
                                [f["P"] for f in frames], [1200, 360], (700, 800), stride=(4, 4))
    for i, f in enumerate(frames):
        d = io.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], f["P"], [1200, 360], (700, 800))
        o = io.produce_sparse_pooling_input(d, stride=[4, 4])
        assert plan.nnz[i] == len(o["Mij_pool"])
        assert_plan_equals_oracle(plan, o["Mij_pool"], np.ones(len(o["Mij_pool"])), o["img_index_flip_pool"], 175 * 200, (90, 300), f=i)


# ----------------------------------------------------------------- pooling kernels
def _case(seed, n, bev_hw, img_hw, cb, ci, skew="uniform", weights=False):
    d = synth.direct_pairs(seed, n, bev_hw=bev_hw, img_wh=(img_hw[1], img_hw[0]), skew=skew)
    o = io.produce_sparse_pooling_input(d, stride=[1, 1])
    rng = np.random.default_rng(seed + 100)
    nnz = len(o["Mij_pool"])
    val = (1.0 / rng.integers(1, 46, nnz)).astype(np.float32) if weights else np.ones(nnz, np.float32)
    bev = rng.standard_normal((1,) + tuple(bev_hw) + (cb,), dtype=np.float32)
    img = rng.standard_normal((1,) + tuple(img_hw) + (ci,), dtype=np.float32)
    return o, val, bev, img


POOL_CASES = [
    # seed, pairs, bev HxW, img HxW, C_bev, C_img, skew, weights
    (0, 400, (9, 11), (7, 13), 8, 12, "uniform", True),
    (1, 3000, (40, 50), (30, 60), 32, 32, "uniform", False),
    (2, 3000, (40, 50), (30, 60), 16, 64, "ground", True),
    (3, 2000, (33, 47), (21, 35), 3, 3, "uniform", True),       # raw RGB pooling (MV3D tests/test_sparse_pooling.py:55-69)
    (4, 2000, (33, 47), (21, 35), 6, 10, "zipf", True),         # float2 path
    (5, 4000, (25, 30), (12, 40), 256, 256, "uniform", False),  # stride-8 VGG conv4 depth
    (6, 5000, (20, 24), (12, 40), 768, 768, "ground", True),    # MV3D depth, long rows
    # few entries per cell: the sparse-regime kernel (entry CTAs + stream CTAs), every vector width, dual = add form
    (7, 200, (40, 50), (30, 60), 32, 32, "uniform", True),
    (8, 150, (33, 47), (21, 35), 3, 3, "uniform", True),
    (9, 300, (40, 50), (30, 60), 6, 10, "zipf", True),
    (10, 400, (60, 70), (30, 60), 64, 16, "ground", False),
]


@pytest.mark.parametrize("case", POOL_CASES, ids=lambda c: "seed%d_C%dx%d_%s" % (c[0], c[4], c[5], c[6]))
def test_layer_forward_backward_dual_bitexact(shpl, case):
    seed, n, bev_hw, img_hw, cb, ci, skew, weights = case
    o, val, bev, img = _case(seed, n, bev_hw, img_hw, cb, ci, skew, weights)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb = torch.from_numpy(bev).cuda().requires_grad_(True)
    ti = torch.from_numpy(img).cuda().requires_grad_(True)
    bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [ci, cb], M, img_index_flip=torch.from_numpy(flip.astype(np.int32)).cuda(),
                                                 bv_index=np.zeros((1, 3)))
    ref_bv = cref.forward(bev[0], img[0], Mij, val, flip)
    ref_img = cref.forward_trans(img[0], bev[0], Mij, val, flip)
    np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), ref_bv)
    np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), ref_img)
    rng = np.random.default_rng(seed + 200)
    g1 = rng.standard_normal(ref_bv.shape, dtype=np.float32)
    g2 = rng.standard_normal(ref_img.shape, dtype=np.float32)
    torch.autograd.backward([bv_fused, img_fused], [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
    gd, gs = cref.backward(g1, Mij, val, flip, cb, img.shape[1:])
    gi, gb = cref.backward_trans(g2, Mij, val, flip, ci, bev.shape[1:])
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd + gb)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gi + gs)


def test_bare_ops_match_reference_shapes(shpl):
    """_sparse_pool_op / _sparse_pool_trans_op as MV3D's Network.sparse_pool calls them (network.py:242-246)."""
    o, val, bev, img = _case(7, 1500, (20, 24), (12, 40), 8, 16, "uniform", True)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    idx = torch.from_numpy(flip.astype(np.int32)).cuda()
    pooled = shpl._sparse_pool_op(M, torch.from_numpy(img).cuda(), idx, [1, 20, 24, 16])
    assert tuple(pooled.shape) == (1, 20, 24, 16)
    ref = cref.forward(bev[0], img[0], Mij, val, flip)[..., 8:]
    np.testing.assert_array_equal(pooled[0].cpu().numpy(), ref)
    pooled_t = shpl._sparse_pool_trans_op(M, torch.from_numpy(bev).cuda(), idx, [1, 12, 40, 8])
    ref_t = cref.forward_trans(img[0], bev[0], Mij, val, flip)[..., 16:]
    np.testing.assert_array_equal(pooled_t[0].cpu().numpy(), ref_t)


def test_batched_frames_one_launch(shpl):
    frames = [synth.avod_frame(s, az_step_deg=0.4) for s in (6, 7)]
    plan = shpl.build_avod_plan([f["points"] for f in frames], [f["voxel_indices"] for f in frames],
                                [f["P"] for f in frames], [1200, 360], (700, 800), stride=(8, 8))
    rng = np.random.default_rng(0)
    bev = rng.standard_normal((2, 87, 100, 16), dtype=np.float32)
    img = rng.standard_normal((2, 45, 150, 8), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused = shpl.sparse_pool(tb, ti, plan)
    g = rng.standard_normal(tuple(fused.shape), dtype=np.float32)
    fused.backward(torch.from_numpy(g).cuda())
    for i, f in enumerate(frames):
        d = io.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], f["P"], [1200, 360], (700, 800))
        o = io.produce_sparse_pooling_input(d, stride=[8, 8])
        val = np.ones(len(o["Mij_pool"]), np.float32)
        np.testing.assert_array_equal(fused[i].detach().cpu().numpy(), cref.forward(bev[i], img[i], o["Mij_pool"], val, o["img_index_flip_pool"]))
        gd, gs = cref.backward(g[i], o["Mij_pool"], val, o["img_index_flip_pool"], 16, (45, 150, 8))
        np.testing.assert_array_equal(tb.grad[i].cpu().numpy(), gd)
        np.testing.assert_array_equal(ti.grad[i].cpu().numpy(), gs)


def test_batch_norm_variant(shpl):
    """use_bn=True (concat_bn_op, sparse_pool_utils.py:120-124): each map normalised on its own, then concat."""
    o, val, bev, img = _case(8, 1500, (20, 24), (12, 40), 8, 16, "uniform", False)
    tb, ti = torch.from_numpy(bev).cuda(), torch.from_numpy(img).cuda()
    fused, _ = shpl.sparse_pool_layer([tb, ti], [16, 8], o | {"M_val": val}, img_index_flip=o["img_index_flip_pool"], use_bn=True)
    ref = cref.forward(bev[0], img[0], o["Mij_pool"], val, o["img_index_flip_pool"])

    def bn(x):
        x = x.astype(np.float64)
        return (x - x.mean(axis=(0, 1))) / np.sqrt(x.var(axis=(0, 1)) + 1e-3)
    want = np.concatenate([bn(ref[..., :8]), bn(ref[..., 8:])], axis=2)
    np.testing.assert_allclose(fused[0].cpu().numpy(), want, rtol=1e-4, atol=1e-5)


# ------------------------------------------------- full KITTI size (BASELINE configs)
def test_kitti_stride1_full_size_against_c_oracle(shpl):
    """Config 1: BEV 700x800x32 <- image 360x1200x32, ~20k pairs, forward + backward, bit-exact
    against the plain-C oracle at the full size."""
    frame = synth.avod_frame(2, az_step_deg=0.05)
    d = shpl.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], Calib(frame["P"]), frame["im_size"], frame["bv_size"])
    o = shpl.produce_sparse_pooling_input(d, stride=[1, 1])
    rng = np.random.default_rng(1)
    bev = rng.standard_normal((1, 700, 800, 32), dtype=np.float32)
    img = rng.standard_normal((1, 360, 1200, 32), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    M = shpl.SparseTensor.from_sparse_pooling_input(o)
    fused, img_out = shpl.sparse_pool_layer([tb, ti], [32, 32], M, img_index_flip=o["img_index_flip_pool"], bv_index=None)
    assert img_out is ti
    nnz = len(o["Mij_pool"])
    val = np.ones(nnz, np.float32)
    ref = cref.forward(bev[0], img[0], o["Mij_pool"], val, o["img_index_flip_pool"])
    np.testing.assert_array_equal(fused[0].detach().cpu().numpy(), ref)
    g = rng.standard_normal(ref.shape, dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    gd, gs = cref.backward(g, o["Mij_pool"], val, o["img_index_flip_pool"], 32, (360, 1200, 32))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)
    # size-independent properties: the dense half is a copy, the pooled half is linear in the source map
    assert torch.equal(fused[..., :32].detach(), tb.detach())
    fused2, _ = shpl.sparse_pool_layer([tb.detach(), 2.0 * ti.detach()], [32, 32], M, img_index_flip=o["img_index_flip_pool"])
    assert torch.equal(fused2[..., 32:], 2.0 * fused[..., 32:].detach())
    # checksum of checksums: sum of the pooled half equals the sum of the gathered rows (fp64 on the host)
    pix = o["img_index_flip_pool"][:, 1] * 1200 + o["img_index_flip_pool"][:, 2]
    want = img[0].reshape(-1, 32)[pix].astype(np.float64).sum()
    got = fused[0, :, :, 32:].detach().double().sum().item()
    assert abs(got - want) <= 1e-5 * np.abs(img[0].reshape(-1, 32)[pix]).astype(np.float64).sum()


def test_avod_fpn_two_layer_config_full_size(shpl):
    """Config 2 (the bench workload) and 2' (RetinaNet P2) at full size from ONE frame's points: layer A after VGG conv4
    (stride 8, dual, 88x100x256 <-> 45x150x256 on the padded 704-row BEV, kitti_dataset.py:384-388) and the P2 layer
    (stride 4, 175x200x256 <- 90x300x256, retinanet_model.py:337-340): indices against the numpy oracle, forward and
    backward bit-exact against the plain-C oracle."""
    frame = synth.avod_frame(3, az_step_deg=0.04)
    rng = np.random.default_rng(9)
    for stride, bv_size, bev_hw, img_hw, dual in ((8, (704, 800), (88, 100), (45, 150), True), (4, (700, 800), (175, 200), (90, 300), False)):
        C = 256
        d = shpl.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], Calib(frame["P"]), frame["im_size"], bv_size)
        o = shpl.produce_sparse_pooling_input(d, stride=[stride, stride])
        d_ref = io.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], frame["P"], frame["im_size"], bv_size)
        o_ref = io.produce_sparse_pooling_input(d_ref, stride=[stride, stride])
        for k in ("Mij_pool", "M_size", "img_index_flip_pool"):
            np.testing.assert_array_equal(np.asarray(o[k]), o_ref[k], err_msg=k)
        Mij, flip = o_ref["Mij_pool"], o_ref["img_index_flip_pool"]
        val = np.ones(len(Mij), np.float32)
        bev = rng.standard_normal((1,) + bev_hw + (C,), dtype=np.float32)
        img = rng.standard_normal((1,) + img_hw + (C,), dtype=np.float32)
        tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
        M = shpl.SparseTensor.from_sparse_pooling_input(o)
        bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=o["img_index_flip_pool"],
                                                     bv_index=(np.zeros((1, 3)) if dual else None))
        np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
        g1 = rng.standard_normal(bev_hw + (2 * C,), dtype=np.float32)
        gd, gs = cref.backward(g1, Mij, val, flip, C, img_hw + (C,))
        if dual:
            np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), cref.forward_trans(img[0], bev[0], Mij, val, flip))
            g2 = rng.standard_normal(img_hw + (2 * C,), dtype=np.float32)
            torch.autograd.backward([bv_fused, img_fused], [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
            gi, gb = cref.backward_trans(g2, Mij, val, flip, C, bev_hw + (C,))
            np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd + gb)
            np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gi + gs)
        else:
            assert img_fused is ti
            bv_fused.backward(torch.from_numpy(g1[None]).cuda())
            np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)
            np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)


def test_full_scan_c128_skewed_rows(shpl):
    """Config 4 / 5 shape: 120k pairs, C=128, ground-plane-skewed rows, non-homogeneous weights."""
    o, val, bev, img = _case(9, 120000, (700, 800), (360, 1200), 128, 128, "ground", True)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused, _ = shpl.sparse_pool_layer([tb, ti], [128, 128], M, img_index_flip=torch.from_numpy(flip).cuda())
    ref = cref.forward(bev[0], img[0], Mij, val, flip)
    np.testing.assert_array_equal(fused[0].detach().cpu().numpy(), ref)
    del ref
    g = np.random.default_rng(3).standard_normal((700, 800, 256), dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    gd, gs = cref.backward(g, Mij, val, flip, 128, (360, 1200, 128))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)


@pytest.mark.parametrize("C,skew", [(16, "uniform"), (32, "ground"), (16, "zipf")])
def test_one_million_pairs_at_the_kitti_map_size(shpl, C, skew):
    """BASELINE config 5 at its largest: 1 M pairs on the 700 x 800 / 360 x 1200 maps -- the dense regime, where every entry
    CTA takes the staged walk (uniform: short runs; ground-plane: runs of ~100 entries; Zipf: listed cells, the long-run
    instantiation beside the split tree).  Forward and backward bit-exact against the plain-C oracle, except the cells above
    SHPL_EXACT_LEN = 2048 entries (Zipf), which the split tree sums within 1e-5 of the sum of |terms|."""
    o, val, bev, img = _case(21, 1000000, (700, 800), (360, 1200), C, C, skew, True)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused, _ = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda())
    got = fused[0].detach().cpu().numpy().reshape(-1, 2 * C)
    ref = cref.forward(bev[0], img[0], Mij, val, flip).reshape(-1, 2 * C)
    rows = np.bincount(Mij[:, 0], minlength=700 * 800)
    tree = rows > 2048
    np.testing.assert_array_equal(got[~tree], ref[~tree])
    if tree.any():
        mag = np.zeros((700 * 800, C))
        src = img[0].reshape(-1, C)[flip[:, 1] * 1200 + flip[:, 2]]
        np.add.at(mag, Mij[:, 0], np.abs(val[:, None].astype(np.float64) * src))
        err = np.abs(got[tree, C:].astype(np.float64) - ref[tree, C:])
        assert float((err / np.maximum(mag[tree], 1e-30)).max()) <= 1e-5
        np.testing.assert_array_equal(got[tree, :C], ref[tree, :C])
    del ref, got
    g = np.random.default_rng(4).standard_normal((700, 800, 2 * C), dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    gd, gs = cref.backward(g, Mij, val, flip, C, (360, 1200, C))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)
    pix = np.bincount(flip[:, 1] * 1200 + flip[:, 2], minlength=360 * 1200)
    assert pix.max() <= 2048                                    # the transposed direction has no tree cell: bit-exact
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)


def _long_cell_case(n, n_same_pixel, seed=12):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 64, n)
    u[:n_same_pixel] = 5
    v = rng.integers(0, 32, n)
    v[:n_same_pixel] = 7
    d = dict(bv_index=np.stack((np.full(n, 3), np.full(n, 2)), axis=1).astype(np.int64),
             img_index=np.stack((u, v, np.zeros(n))).astype(np.float64), bv_size=np.array([16, 16]), img_size=np.array([64, 32]))
    val = (1.0 / rng.integers(1, 46, n)).astype(np.float32)
    bev = rng.standard_normal((1, 16, 16, 32), dtype=np.float32)
    img = rng.standard_normal((1, 32, 64, 32), dtype=np.float32)
    return d, val, bev, img


@pytest.mark.parametrize("n_cell,n_pix,listed", [(512, 300, (0, 0)), (513, 513, (1, 1)), (1300, 700, (1, 1)), (2048, 2048, (1, 1))])
def test_long_and_listed_cells_up_to_exact_len_are_bit_exact(shpl, n_cell, n_pix, listed):
    """Every pair lands in one BEV cell and n_pix pairs share one pixel.  Up to SHPL_HEAVY_LEN = 512 entries the cell is
    summed by one warp in the main kernel; above, the builder lists it and shpl_pool_heavy's exact kernel (a cluster
    gathers and multiplies in parallel, the adds stay in entry order) forms the sum, up to SHPL_EXACT_LEN = 2048
    entries: identical to the sequential oracle, bit for bit, on both sides of the listing threshold and at the limit."""
    d, val, bev, img = _long_cell_case(n_cell, n_pix)
    o = shpl.produce_sparse_pooling_input(d)
    assert o["shpl_plan"].n_heavy == listed
    M = shpl.SparseTensor(torch.from_numpy(o["Mij_pool"]).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused, _ = shpl.sparse_pool_layer([tb, ti], [32, 32], M, img_index_flip=torch.from_numpy(o["img_index_flip_pool"]).cuda())
    ref = cref.forward(bev[0], img[0], o["Mij_pool"], val, o["img_index_flip_pool"])
    np.testing.assert_array_equal(fused[0].detach().cpu().numpy(), ref)
    g = np.random.default_rng(1).standard_normal(ref.shape, dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    gd, gs = cref.backward(g, o["Mij_pool"], val, o["img_index_flip_pool"], 32, (32, 64, 32))
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)


@pytest.mark.parametrize("C,n,bev_hw", [(4, 700, (16, 16)), (12, 333, (16, 16)), (16, 920, (16, 16)), (24, 100, (16, 16)),
                                        (64, 1500, (16, 16)), (100, 257, (16, 16)), (2, 90, (16, 16)), (3, 65, (16, 16)),
                                        # the same on a map with few entries per cell: the sparse-regime kernel, whose
                                        # stream warps take the long cells over from the entry CTAs
                                        (16, 920, (120, 110)), (4, 700, (120, 110)), (24, 100, (120, 110)), (100, 257, (120, 110)),
                                        (6, 333, (120, 110)),
                                        # listed cells (> 512 entries) through the exact cluster kernel: one / two / six
                                        # adder warps, three lane groups (C = 36), and the sparse-regime map
                                        (128, 1500, (16, 16)), (256, 700, (16, 16)), (768, 600, (16, 16)), (36, 800, (16, 16)),
                                        (128, 1100, (120, 110)), (8, 2000, (16, 16))])
def test_long_rows_are_summed_by_the_whole_warp_in_k_order(shpl, C, n, bev_hw):
    """Narrow kernels hand cells with more than 32 entries to the whole warp (lane groups gather in parallel,
    the adds stay in ascending k): bit-identical to the sequential oracle for every vector layout --
    C=4 (one float4 per cell, 32 entries in parallel), C=24 (6 vectors, 5 groups), C=100 (25 vectors, 1 group),
    C=2 / C=3 (float2 / scalar instantiations) -- next to short cells and a long pixel for the backward."""
    rng = np.random.default_rng(C * 1000 + n)
    n_bg = 400
    u = np.r_[rng.integers(0, 64, n), rng.integers(0, 64, n_bg)]
    v = np.r_[rng.integers(0, 32, n), rng.integers(0, 32, n_bg)]
    u[: n // 2] = 9
    v[: n // 2] = 4                                              # a long pixel (n/2 entries) for the transposed direction
    bx = np.r_[np.full(n, 3), rng.integers(0, bev_hw[1], n_bg)]
    bz = np.r_[np.full(n, 2), rng.integers(0, bev_hw[0], n_bg)]
    perm = rng.permutation(n + n_bg)                             # long-cell entries interleaved with the others in k
    d = dict(bv_index=np.stack((bx, bz), axis=1)[perm].astype(np.int64),
             img_index=np.stack((u, v, np.zeros(n + n_bg)))[:, perm].astype(np.float64),
             bv_size=np.array(bev_hw), img_size=np.array([64, 32]))
    val = (1.0 / rng.integers(1, 46, n + n_bg)).astype(np.float32)
    bev = rng.standard_normal((1,) + tuple(bev_hw) + (C,), dtype=np.float32)
    img = rng.standard_normal((1, 32, 64, C), dtype=np.float32)
    o = shpl.produce_sparse_pooling_input(d)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    pixel = flip[:, 1] * 64 + flip[:, 2]
    assert np.bincount(Mij[:, 0]).max() >= n
    assert o["shpl_plan"].n_heavy == (int((np.bincount(Mij[:, 0]) > 512).sum()), int((np.bincount(pixel) > 512).sum()))
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda(),
                                                 bv_index=np.zeros((1, 3)))
    np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
    np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), cref.forward_trans(img[0], bev[0], Mij, val, flip))
    g1 = rng.standard_normal(tuple(bev_hw) + (2 * C,), dtype=np.float32)
    g2 = rng.standard_normal((32, 64, 2 * C), dtype=np.float32)
    torch.autograd.backward([bv_fused, img_fused], [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
    gd, gs = cref.backward(g1, Mij, val, flip, C, (32, 64, C))
    gi, gb = cref.backward_trans(g2, Mij, val, flip, C, tuple(bev_hw) + (C,))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd + gb)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gi + gs)
    # single direction too (concat form of the narrow kernel, no AddN)
    tb2, ti2 = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    only, _ = shpl.sparse_pool_layer([tb2, ti2], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda())
    np.testing.assert_array_equal(only[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
    only.backward(torch.from_numpy(g1[None]).cuda())
    np.testing.assert_array_equal(ti2.grad[0].cpu().numpy(), gs)
    np.testing.assert_array_equal(tb2.grad[0].cpu().numpy(), gd)


@pytest.mark.parametrize("C,max_len,bev_hw", [(4, 700, (12, 12)), (16, 500, (12, 12)), (20, 300, (12, 12)), (64, 400, (12, 12)),
                                              (128, 300, (12, 12)), (200, 150, (10, 10)), (256, 200, (10, 10)), (6, 700, (12, 12)),
                                              (3, 900, (12, 12)), (32, 1500, (12, 12)),
                                              # few entries per cell on average: only the entry CTAs that meet a long cell stage it
                                              (32, 400, (200, 180)), (8, 120, (200, 180)), (128, 90, (150, 160))])
def test_staged_entry_walk_is_bit_exact(shpl, C, max_len, bev_hw):
    """Cells of 1 ... max_len entries side by side (runs that cross the batches and the CTA chunks of the staged entry walk,
    pool_entries_staged), pixels likewise for the transposed direction, every vector layout (float4 / float2 / scalar,
    power-of-two and odd vector counts, one and two vectors per lane), dual = the add form: forward and backward of both
    directions bit-identical to the sequential oracle.  Cells above 512 entries take the exact cluster kernel."""
    rng = np.random.default_rng(C * 7919 + max_len)
    n_long = 40
    lens = rng.integers(1, max_len + 1, n_long)
    cells = rng.choice(bev_hw[0] * bev_hw[1], n_long, replace=False)
    rows = np.concatenate([np.full(int(l), c) for l, c in zip(lens, cells)] + [rng.integers(0, bev_hw[0] * bev_hw[1], 600)])
    n = len(rows)
    u, v = rng.integers(0, 64, n), rng.integers(0, 32, n)
    k_pix = min(n // 2, max_len)
    u[:k_pix], v[:k_pix] = 11, 5                                  # one long pixel for the transposed direction
    perm = rng.permutation(n)
    rows, u, v = rows[perm], u[perm], v[perm]
    d = dict(bv_index=np.stack((rows % bev_hw[1], bev_hw[0] - 1 - rows // bev_hw[1]), axis=1).astype(np.int64),
             img_index=np.stack((u, v, np.zeros(n))).astype(np.float64), bv_size=np.array(bev_hw), img_size=np.array([64, 32]))
    val = (1.0 / rng.integers(1, 46, n)).astype(np.float32)
    bev = rng.standard_normal((1,) + tuple(bev_hw) + (C,), dtype=np.float32)
    img = rng.standard_normal((1, 32, 64, C), dtype=np.float32)
    o = shpl.produce_sparse_pooling_input(d)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    assert len(Mij) == n and np.bincount(Mij[:, 0]).max() >= lens.max()
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda(), bv_index=np.zeros((1, 3)))
    np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
    np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), cref.forward_trans(img[0], bev[0], Mij, val, flip))
    g1 = rng.standard_normal(tuple(bev_hw) + (2 * C,), dtype=np.float32)
    g2 = rng.standard_normal((32, 64, 2 * C), dtype=np.float32)
    torch.autograd.backward([bv_fused, img_fused], [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
    gd, gs = cref.backward(g1, Mij, val, flip, C, (32, 64, C))
    gi, gb = cref.backward_trans(g2, Mij, val, flip, C, tuple(bev_hw) + (C,))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd + gb)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gi + gs)
    tb2, ti2 = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    only, _ = shpl.sparse_pool_layer([tb2, ti2], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda())
    np.testing.assert_array_equal(only[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
    only.backward(torch.from_numpy(g1[None]).cuda())
    np.testing.assert_array_equal(ti2.grad[0].cpu().numpy(), gs)
    np.testing.assert_array_equal(tb2.grad[0].cpu().numpy(), gd)


@pytest.mark.parametrize("n_cell,n_pix", [(3, 2), (33, 10), (100, 40), (32, 600), (513, 513)])
def test_long_cell_counters_pick_the_heavy_len(shpl, n_cell, n_pix):
    """The builder counts the cells / pixels with more than SHPL_LONG_LEN = 32 entries next to the listed ones (counts[6],
    counts[7]); the drop-in layer turns the counters it reads back into the heavy_len it passes to the pooling entry points:
    0 (no long cell: the kernels skip every long-cell path), SHPL_EXACT_LEN (long but nothing listed), SHPL_HEAVY_LEN (listed
    cells).  Whatever the level, the result is the sequential oracle's, bit for bit."""
    from sparse_pooling_b200 import _cabi
    d, val, bev, img = _long_cell_case(max(n_cell, n_pix) + 200, n_pix)
    d["bv_index"][n_cell:, 0] = np.arange(len(d["bv_index"]) - n_cell) % 16        # only the first n_cell pairs share a cell
    d["bv_index"][n_cell:, 1] = 5 + (np.arange(len(d["bv_index"]) - n_cell) // 16) % 11
    o = shpl.produce_sparse_pooling_input(d)
    plan, Mij, flip = o["shpl_plan"], o["Mij_pool"], o["img_index_flip_pool"]
    rows, pix = np.bincount(Mij[:, 0]), np.bincount(flip[:, 1] * 64 + flip[:, 2])
    assert plan.n_long == (int((rows > 32).sum()), int((pix > 32).sum()))
    assert plan.n_heavy == (int((rows > 512).sum()), int((pix > 512).sum()))

    def level(c):
        return _cabi.HEAVY_LEN if (c > 512).any() else (_cabi.EXACT_LEN if (c > 32).any() else 0)
    assert (plan.heavy_len(False), plan.heavy_len(True)) == (level(rows), level(pix))
    assert plan.heavy_len() == level(np.concatenate((rows, pix)))          # the dual entry points: either direction
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [32, 32], M, img_index_flip=torch.from_numpy(flip).cuda(), bv_index=np.zeros((1, 3)))
    np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip))
    np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), cref.forward_trans(img[0], bev[0], Mij, val, flip))


@pytest.mark.parametrize("C", [32, 128, 256])
def test_main_and_heavy_entry_points_agree_on_who_sums_a_listed_cell(shpl, C):
    """Bare C ABI, the un-split heavy entry point: a listed cell of 1300 entries is summed by the main kernel (staged walk)
    for C <= 128 and by shpl_pool_heavy's exact cluster kernel above -- never by both, never by neither: the rule depends
    on C alone.  Either way the sequential sum, bit for bit."""
    import ctypes
    from sparse_pooling_b200 import _cabi, ops
    d, val, bev, img = _long_cell_case(1300, 40)
    o = shpl.produce_sparse_pooling_input(d)
    plan, Mij, flip = o["shpl_plan"], o["Mij_pool"], o["img_index_flip_pool"]
    assert plan.n_heavy[0] == 1
    rng = np.random.default_rng(C)
    bev = rng.standard_normal((16 * 16, C), dtype=np.float32)
    img = rng.standard_normal((32 * 64, C), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda(), torch.from_numpy(img).cuda()
    fused = torch.full((16 * 16, 2 * C), float("nan"), device="cuda")
    ptr, key, idx, v, nnz_max, heavy, _ = plan.by_row()
    P = ops._ptr
    rc = _cabi.lib.shpl_pool_forward(P(tb), P(ti), ptr, key, idx, v, int(nnz_max), _cabi.HEAVY_LEN, 16 * 16, C, 32 * 64, C, P(fused), ops._stream())
    _cabi.check(rc, "shpl_pool_forward")
    lst, count_dev, cap = heavy[0], heavy[1], heavy[2]
    rc = _cabi.lib.shpl_pool_heavy(P(ti), C, C, ptr, idx, v, lst, count_dev, int(cap), None, 0,
                                   ctypes.c_void_p(fused.data_ptr() + 4 * C), 2 * C, ops._stream())
    _cabi.check(rc, "shpl_pool_heavy")
    ref = cref.forward(bev.reshape(16, 16, C), img.reshape(32, 64, C), Mij, np.ones(len(Mij), np.float32), flip).reshape(-1, 2 * C)
    np.testing.assert_array_equal(fused.cpu().numpy(), ref)


@pytest.mark.parametrize("dual", [False, True])
def test_heavy_cells_use_the_cluster_tree(shpl, dual):
    """Stress (BASELINE config 5, Zipf-like skew): one BEV cell with 30k entries and one pixel with 5k.
    Cells above SHPL_HEAVY_LEN are summed by a fixed tree over a thread-block cluster: deterministic,
    and within the north_star's 1e-5 (relative to the sum of |terms|, the bound of fp32 summation) of
    the sequential oracle -- not bit-identical to it.  Everything else stays bit-exact."""
    d, val, bev, img = _long_cell_case(30000, 5000)
    o = shpl.produce_sparse_pooling_input(d)
    plan = o["shpl_plan"]
    assert plan.n_heavy == (1, 1)
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"], plan=plan)
    plan.csr_val.copy_(torch.from_numpy(val[io.build_plan(Mij, val, flip, 256, 32, 64)["csr_ent"]]).cuda())
    plan.csrT_val.copy_(torch.from_numpy(val[io.build_plan(Mij, val, flip, 256, 32, 64)["csrT_ent"]]).cuda())
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    outs = shpl.sparse_pool_layer([tb, ti], [32, 32], M, img_index_flip=torch.from_numpy(flip).cuda(),
                                  bv_index=(np.zeros((1, 3)) if dual else None))
    ref_bv = cref.forward(bev[0], img[0], Mij, val, flip)
    pix = flip[:, 1] * 64 + flip[:, 2]
    heavy_row, heavy_pix = 2 * 16 + 3, 7 * 64 + 5

    def close(got, want, scale):     # |diff| <= 1e-5 * sum|terms|
        assert np.abs(got - want).max() <= 1e-5 * scale, (np.abs(got - want).max(), scale)

    got_bv = outs[0][0].detach().cpu().numpy().reshape(256, 64)
    want_bv = ref_bv.reshape(256, 64)
    mask = np.ones(256, bool)
    mask[heavy_row] = False
    np.testing.assert_array_equal(got_bv[mask], want_bv[mask])
    np.testing.assert_array_equal(got_bv[heavy_row, :32], want_bv[heavy_row, :32])
    close(got_bv[heavy_row, 32:], want_bv[heavy_row, 32:], np.abs(val[:, None] * img[0].reshape(-1, 32)[pix]).sum(0).max())
    again = shpl.sparse_pool_layer([tb, ti], [32, 32], M, img_index_flip=torch.from_numpy(flip).cuda(),
                                   bv_index=(np.zeros((1, 3)) if dual else None))
    assert torch.equal(again[0], outs[0])                              # deterministic
    rng = np.random.default_rng(3)
    g1 = rng.standard_normal((16, 16, 64), dtype=np.float32)
    gd, gs = cref.backward(g1, Mij, val, flip, 32, (32, 64, 32))
    if dual:
        ref_img = cref.forward_trans(img[0], bev[0], Mij, val, flip).reshape(2048, 64)
        got_img = outs[1][0].detach().cpu().numpy().reshape(2048, 64)
        pm = np.ones(2048, bool)
        pm[heavy_pix] = False
        np.testing.assert_array_equal(got_img[pm], ref_img[pm])
        close(got_img[heavy_pix, 32:], ref_img[heavy_pix, 32:], np.abs(val[:, None] * bev[0].reshape(-1, 32)[Mij[:, 0]]).sum(0).max())
        g2 = rng.standard_normal((32, 64, 64), dtype=np.float32)
        torch.autograd.backward(list(outs), [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
        gi, gb = cref.backward_trans(g2, Mij, val, flip, 32, (16, 16, 32))
        want_b, want_i = (gd + gb).reshape(256, 32), (gi + gs).reshape(2048, 32)
    else:
        outs[0].backward(torch.from_numpy(g1[None]).cuda())
        want_b, want_i = gd.reshape(256, 32), gs.reshape(2048, 32)
        pm = np.ones(2048, bool)
        pm[heavy_pix] = False
    got_b, got_i = tb.grad[0].cpu().numpy().reshape(256, 32), ti.grad[0].cpu().numpy().reshape(2048, 32)
    np.testing.assert_array_equal(got_i[pm], want_i[pm])
    close(got_i[heavy_pix], want_i[heavy_pix], np.abs(val[:, None] * g1.reshape(256, 64)[Mij[:, 0], 32:]).sum(0).max() + 10)
    if dual:
        np.testing.assert_array_equal(got_b[mask], want_b[mask])
        close(got_b[heavy_row], want_b[heavy_row], 1e3)
    else:
        np.testing.assert_array_equal(got_b, want_b)


@pytest.mark.parametrize("C,n_heavy", [(32, 2600), (128, 2600), (32, 1900), (128, 1900), (16, 1900)])
def test_sparse_regime_with_a_heavy_cell(shpl, C, n_heavy):
    """(C = 32: narrow channel count, sparse regime; C = 128: the wide variant of the same kernel.)
    Few entries next to the cells (sparse-regime kernel) but n_heavy of them in ONE BEV cell (> SHPL_HEAVY_LEN):
    the entry CTAs skip the listed cell, the stream CTAs write it as empty, shpl_pool_heavy fills it in.  Everything
    but the listed cell is bit-exact; the listed cell too when it has at most SHPL_EXACT_LEN = 2048 entries (the exact
    kernel keeps the sequential order; with C = 16 the main kernel keeps the cell itself), otherwise it is within 1e-5
    of the sum of |terms| (the tree)."""
    rng = np.random.default_rng(21)
    n_bg = 300
    n = n_heavy + n_bg
    bx = np.r_[np.full(n_heavy, 7), rng.integers(0, 110, n_bg)]
    bz = np.r_[np.full(n_heavy, 5), rng.integers(0, 120, n_bg)]
    u, v = rng.integers(0, 64, n), rng.integers(0, 32, n)
    perm = rng.permutation(n)
    d = dict(bv_index=np.stack((bx, bz), axis=1)[perm].astype(np.int64), img_index=np.stack((u, v, np.zeros(n)))[:, perm].astype(np.float64),
             bv_size=np.array([120, 110]), img_size=np.array([64, 32]))
    val = (1.0 / rng.integers(1, 46, n)).astype(np.float32)
    bev = rng.standard_normal((1, 120, 110, C), dtype=np.float32)
    img = rng.standard_normal((1, 32, 64, C), dtype=np.float32)
    o = shpl.produce_sparse_pooling_input(d, M_val=val.astype(np.float64))
    plan = o["shpl_plan"]
    assert plan.n_heavy == (1, 0) and 4 * n <= 120 * 110
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    M = shpl.SparseTensor.from_sparse_pooling_input(o)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused, _ = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=flip)
    ref = cref.forward(bev[0], img[0], Mij, val, flip).reshape(-1, 2 * C)
    got = fused[0].detach().cpu().numpy().reshape(-1, 2 * C)
    heavy_row = 5 * 110 + 7
    mask = np.ones(len(ref), bool)
    mask[heavy_row] = False
    np.testing.assert_array_equal(got[mask], ref[mask])
    np.testing.assert_array_equal(got[heavy_row, :C], ref[heavy_row, :C])
    pix = flip[:, 1] * 64 + flip[:, 2]
    rows = Mij[:, 0]
    if n_heavy <= 2048:
        np.testing.assert_array_equal(got[heavy_row, C:], ref[heavy_row, C:])      # sequential order kept
    else:
        scale = np.abs(val[rows == heavy_row, None] * img[0].reshape(-1, C)[pix[rows == heavy_row]]).sum(0).max()
        assert np.abs(got[heavy_row, C:] - ref[heavy_row, C:]).max() <= 1e-5 * scale
    g = rng.standard_normal((120, 110, 2 * C), dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    gd, gs = cref.backward(g, Mij, val, flip, C, (32, 64, C))
    np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd)
    np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs)       # no pixel is heavy: the backward stays bit-exact


def test_lazy_gen_dict_behaves_like_the_reference_dict(shpl, golden_dir):
    """gen's dict is evaluated lazily so that gen -> produce runs the fused builder; looking at it
    before or after produce must still show what the reference shows (incl. the in-place mutation
    and the compounding of a second produce call, quirk A.4-6)."""
    frame = synth.avod_frame(1, az_step_deg=0.09)
    ref = io.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], frame["P"], frame["im_size"], frame["bv_size"])
    ref_o1 = io.produce_sparse_pooling_input(ref, stride=[4, 4])
    ref_after1 = ref["img_index"].copy()
    ref_o2 = io.produce_sparse_pooling_input(ref, stride=[2, 2])
    # fused path: produce before anyone looks
    d = shpl.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], Calib(frame["P"]), frame["im_size"], frame["bv_size"])
    assert "img_index" in d and len(d) == 4 and not d._done
    o1 = shpl.produce_sparse_pooling_input(d, stride=[4, 4])
    assert not d._done                                             # nothing was read back for the gen dict
    np.testing.assert_array_equal(o1["Mij_pool"], ref_o1["Mij_pool"])
    np.testing.assert_array_equal(o1["img_index_flip_pool"], ref_o1["img_index_flip_pool"])
    np.testing.assert_array_equal(d["img_index"], ref_after1)      # materialised now, mutation included
    o2 = shpl.produce_sparse_pooling_input(d, stride=[2, 2])       # compounds, like the reference
    np.testing.assert_array_equal(o2["Mij_pool"], ref_o2["Mij_pool"])
    np.testing.assert_array_equal(o2["img_index_flip_pool"], ref_o2["img_index_flip_pool"])
    np.testing.assert_array_equal(d["img_index"], ref["img_index"])
    assert sorted(d.keys()) == ["bv_index", "bv_size", "img_index", "img_size"]


def test_randomized_shapes_every_dispatch_path(shpl):
    """Forty seeded random layers -- map sizes, channel counts (every vector width, odd counts, narrow and wide), pair
    counts from empty to dense, uniform / ground-plane / Zipf row skew, single and dual direction, unit and 1/count
    weights -- each forward + backward bit-exact against the plain-C oracle.  Sweeps the dispatch rules of
    launch_jobs (entry + stream kernel in its three forms, narrow kernel, long-row paths) with shapes nobody hand-picked."""
    rng = np.random.default_rng(20241018)
    channels = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 100, 128, 160, 256]
    for case in range(40):
        bev_hw = (int(rng.integers(3, 48)), int(rng.integers(3, 56)))
        img_hw = (int(rng.integers(3, 40)), int(rng.integers(3, 64)))
        cb, ci = int(rng.choice(channels)), int(rng.choice(channels))
        n = int(rng.choice([0, 1, 7, 60, 300, 1500, 4000]))
        skew = str(rng.choice(["uniform", "ground", "zipf"]))
        dual = bool(rng.integers(0, 2))
        weights = bool(rng.integers(0, 2))
        tag = "case %d: bev %s img %s C %dx%d n %d %s dual %s" % (case, bev_hw, img_hw, cb, ci, n, skew, dual)
        d = synth.direct_pairs(1000 + case, max(n, 1), bev_hw=bev_hw, img_wh=(img_hw[1], img_hw[0]), skew=skew)
        if n == 0:
            d = dict(bv_index=d["bv_index"][:0], img_index=d["img_index"][:, :0], bv_size=d["bv_size"], img_size=d["img_size"])
        o = io.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()}, stride=[1, 1])
        Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
        nnz = len(Mij)
        val = (1.0 / rng.integers(1, 46, nnz)).astype(np.float32) if weights else np.ones(nnz, np.float32)
        bev = rng.standard_normal((1,) + bev_hw + (cb,), dtype=np.float32)
        img = rng.standard_normal((1,) + img_hw + (ci,), dtype=np.float32)
        M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), o["M_size"])
        tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
        bv_fused, img_fused = shpl.sparse_pool_layer([tb, ti], [ci, cb], M, img_index_flip=torch.from_numpy(flip).cuda(),
                                                     bv_index=(np.zeros((1, 3)) if dual else None))
        np.testing.assert_array_equal(bv_fused[0].detach().cpu().numpy(), cref.forward(bev[0], img[0], Mij, val, flip), err_msg=tag)
        g1 = rng.standard_normal(bev_hw + (cb + ci,), dtype=np.float32)
        gd, gs = cref.backward(g1, Mij, val, flip, cb, img.shape[1:])
        if dual:
            np.testing.assert_array_equal(img_fused[0].detach().cpu().numpy(), cref.forward_trans(img[0], bev[0], Mij, val, flip), err_msg=tag)
            g2 = rng.standard_normal(img_hw + (ci + cb,), dtype=np.float32)
            torch.autograd.backward([bv_fused, img_fused], [torch.from_numpy(g1[None]).cuda(), torch.from_numpy(g2[None]).cuda()])
            gi, gb = cref.backward_trans(g2, Mij, val, flip, ci, bev.shape[1:])
            np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd + gb, err_msg=tag)
            np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gi + gs, err_msg=tag)
        else:
            bv_fused.backward(torch.from_numpy(g1[None]).cuda())
            np.testing.assert_array_equal(tb.grad[0].cpu().numpy(), gd, err_msg=tag)
            np.testing.assert_array_equal(ti.grad[0].cpu().numpy(), gs, err_msg=tag)


def test_randomized_builder_against_the_index_oracle(shpl):
    """Thirty seeded random calls of produce_sparse_pooling_input -- map sizes, strides 1..8 on either side, pixel and
    cell indices that leave the maps on both sides (MV3D's augment_fv does that, minibatch_mv3d_img.py:205-206), with
    and without weights -- COO outputs, in-place mutation and the CSR / CSR^T plan against the numpy oracle, bit for bit."""
    rng = np.random.default_rng(777)
    for case in range(30):
        im_w, im_h = int(rng.integers(16, 400)), int(rng.integers(8, 200))
        bv_h, bv_w = int(rng.integers(8, 300)), int(rng.integers(8, 300))
        s_img, s_bv = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        n = int(rng.choice([0, 1, 5, 200, 3000]))
        wild = bool(rng.integers(0, 2))                     # indices outside the maps
        lo, hi = (-0.2, 1.3) if wild else (0.0, 1.0)
        u = np.floor(rng.uniform(lo, hi, n) * im_w)
        v = np.floor(rng.uniform(lo, hi, n) * im_h)
        img_index = np.stack((u, v, np.zeros(n))).astype(np.float64)
        bx = np.floor(rng.uniform(lo, hi, n) * bv_w).astype(np.int64)
        bz = np.floor(rng.uniform(lo, hi, n) * bv_h).astype(np.int64)
        d = dict(img_index=img_index, bv_index=np.stack((bx, bz), axis=1), bv_size=np.array([bv_h, bv_w]), img_size=np.array([im_w, im_h]))
        d_ref = {k: np.array(vv, copy=True) for k, vv in d.items()}
        o_ref = io.produce_sparse_pooling_input(d_ref, stride=[s_img, s_bv])
        nnz = len(o_ref["Mij_pool"])
        m_val = (1.0 / rng.integers(1, 46, nnz)) if (rng.integers(0, 2) and nnz) else None
        tag = "case %d: img %dx%d bev %dx%d stride (%d,%d) n %d wild %s" % (case, im_w, im_h, bv_h, bv_w, s_img, s_bv, n, wild)
        o = shpl.produce_sparse_pooling_input(d, M_val=m_val, stride=[s_img, s_bv])
        for k in ("Mij_pool", "M_size", "img_index_flip_pool"):
            np.testing.assert_array_equal(np.asarray(o[k]), o_ref[k], err_msg=tag + " " + k)
        np.testing.assert_array_equal(d["img_index"], d_ref["img_index"], err_msg=tag + " in-place floor/clamp")
        Hp, Wp = im_h // s_img, im_w // s_img
        R = (bv_h // s_bv) * (bv_w // s_bv)
        if R == 0 or Hp == 0 or Wp == 0:
            continue
        val = np.ones(nnz, np.float32) if m_val is None else m_val.astype(np.float32)
        assert_plan_equals_oracle(o["shpl_plan"], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"], R, (Hp, Wp))


@pytest.mark.parametrize("C", [32, 128])
def test_many_listed_cells_share_the_clusters(shpl, C):
    """More listed cells than clusters in the launch (200 cells of 520..900 entries, 128 clusters for the exact kernel, 64 for
    the tree) plus three cells above SHPL_EXACT_LEN: every cluster walks several cells one after the other, re-using its
    staging buffer.  Exact-kernel cells and everything else bit-identical to the sequential oracle, tree cells within 1e-5
    of the sum of |terms|."""
    rng = np.random.default_rng(100 + C)
    Hb, Wb, Hi, Wi = 120, 110, 32, 64
    cells = rng.choice(Hb * Wb, 203, replace=False)
    lens = np.r_[rng.integers(520, 901, 200), [2100, 3000, 5000]]
    rows = np.r_[np.repeat(cells, lens), rng.integers(0, Hb * Wb, 3000)]
    n = len(rows)
    perm = rng.permutation(n)
    rows = rows[perm]
    u, v = rng.integers(0, Wi, n), rng.integers(0, Hi, n)
    d = dict(bv_index=np.stack((rows % Wb, rows // Wb), axis=1).astype(np.int64), img_index=np.stack((u, v, np.zeros(n))).astype(np.float64),
             bv_size=np.array([Hb, Wb]), img_size=np.array([Wi, Hi]))
    val = (1.0 / rng.integers(1, 46, n)).astype(np.float32)
    bev = rng.standard_normal((1, Hb, Wb, C), dtype=np.float32)
    img = rng.standard_normal((1, Hi, Wi, C), dtype=np.float32)
    o = shpl.produce_sparse_pooling_input(d, M_val=val.astype(np.float64))
    Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
    counts = np.bincount(Mij[:, 0], minlength=Hb * Wb)
    assert o["shpl_plan"].n_heavy[0] == int((counts > 512).sum()) >= 203
    M = shpl.SparseTensor.from_sparse_pooling_input(o)
    tb, ti = torch.from_numpy(bev).cuda(), torch.from_numpy(img).cuda()
    fused, _ = shpl.sparse_pool_layer([tb, ti], [C, C], M, img_index_flip=torch.from_numpy(flip).cuda())
    got = fused[0].cpu().numpy().reshape(Hb * Wb, 2 * C)
    ref = cref.forward(bev[0], img[0], Mij, val, flip).reshape(Hb * Wb, 2 * C)
    tree = counts > 2048
    assert tree.sum() == 3
    np.testing.assert_array_equal(got[~tree], ref[~tree])
    pix = flip[:, 1] * Wi + flip[:, 2]
    for r in np.nonzero(tree)[0]:
        sel = Mij[:, 0] == r
        scale = np.abs(val[sel, None] * img[0].reshape(-1, C)[pix[sel]]).sum(0).max()
        np.testing.assert_array_equal(got[r, :C], ref[r, :C])
        assert np.abs(got[r, C:] - ref[r, C:]).max() <= 1e-5 * scale


def test_stacked_frames_with_listed_cells(shpl):
    """Three frames stacked into one plan (build_pairs_plan), each with its own listed cell / pixel (600..1900 entries) at a
    different place: the heavy lists carry the stacked row / pixel numbers and the entry offsets of the later frames, the
    exact kernel serves all of them -- forward and backward of every frame bit-identical to the per-frame oracle."""
    rng = np.random.default_rng(77)
    B, C, Hb, Wb, Hi, Wi = 3, 32, 40, 50, 24, 64
    img_index, bv_index, m_val, refs = [], [], [], []
    for f in range(B):
        L, Lp = int(rng.integers(600, 1900)), int(rng.integers(520, 900))
        n = L + 1500
        bx = np.r_[np.full(L, 3 + f), rng.integers(0, Wb, 1500)]
        bz = np.r_[np.full(L, 7 + 2 * f), rng.integers(0, Hb, 1500)]
        u, v = rng.integers(0, Wi, n), rng.integers(0, Hi, n)
        u[L:L + Lp], v[L:L + Lp] = 5 + f, 9                                  # a listed pixel for the transposed arrays
        perm = rng.permutation(n)
        ii = np.stack((u, v, np.zeros(n)))[:, perm].astype(np.float64)
        bi = np.stack((bx, bz), axis=1)[perm].astype(np.int64)
        mv = 1.0 / rng.integers(1, 46, n)
        o_ref = io.produce_sparse_pooling_input(dict(img_index=ii.copy(), img_size=np.array([Wi, Hi]), bv_index=bi, bv_size=np.array([Hb, Wb])),
                                                M_val=mv)
        img_index.append(ii)
        bv_index.append(bi)
        m_val.append(mv)
        refs.append((o_ref, mv.astype(np.float32)))
    plan = shpl.build_pairs_plan(img_index, bv_index, [Wi, Hi], [Hb, Wb], m_val=m_val)
    assert plan.frames == B and plan.n_heavy == (B, B)
    bev = rng.standard_normal((B, Hb, Wb, C), dtype=np.float32)
    img = rng.standard_normal((B, Hi, Wi, C), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda().requires_grad_(True), torch.from_numpy(img).cuda().requires_grad_(True)
    fused = shpl.sparse_pool(tb, ti, plan)
    g = rng.standard_normal(tuple(fused.shape), dtype=np.float32)
    fused.backward(torch.from_numpy(g).cuda())
    for k, (o_ref, val) in enumerate(refs):
        np.testing.assert_array_equal(fused[k].detach().cpu().numpy(),
                                      cref.forward(bev[k], img[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"]))
        gd, gs = cref.backward(g[k], o_ref["Mij_pool"], val, o_ref["img_index_flip_pool"], C, (Hi, Wi, C))
        np.testing.assert_array_equal(tb.grad[k].cpu().numpy(), gd)
        np.testing.assert_array_equal(ti.grad[k].cpu().numpy(), gs)


# ------------------------------------------------------------------ no-concat ("sparse-only") forms, SURVEY.md 8(d)
@pytest.mark.parametrize("C_b,C_i,n,skew", [(32, 32, 3000, "uniform"), (16, 64, 5000, "ground"), (256, 256, 2000, "uniform"), (4, 12, 900, "zipf")])
def test_no_concat_forms_equal_the_concat_forms_bit_for_bit(shpl, C_b, C_i, n, skew):
    """shpl_pool_forward_into writes only the pooled channels of a fused buffer whose destination channels its producer
    already wrote; shpl_pool_backward_from reads the pooled channels of g_fused in place.  Same values as the concat
    forms (and therefore as the oracle), the untouched channels stay untouched."""
    from sparse_pooling_b200 import ops
    bev_hw, img_hw = (60, 70), (30, 50)
    d = synth.direct_pairs(11, n, bev_hw, (img_hw[1], img_hw[0]), skew=skew)
    o = shpl.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()})
    plan = o["shpl_plan"]
    rng = np.random.default_rng(5)
    bev = rng.standard_normal((1,) + bev_hw + (C_b,), dtype=np.float32)
    img = rng.standard_normal((1,) + img_hw + (C_i,), dtype=np.float32)
    tb, ti = torch.from_numpy(bev).cuda(), torch.from_numpy(img).cuda()
    ref = shpl.sparse_pool(tb, ti, plan)                                   # concat form (bit-exact vs the oracle elsewhere)
    buf = torch.full((1,) + bev_hw + (C_b + C_i,), 7.0, device="cuda")
    buf[..., :C_b] = tb                                                    # "the producer wrote its channels there"
    out = ops.sparse_pool_into(buf, ti, plan)
    assert out.data_ptr() == buf.data_ptr()
    np.testing.assert_array_equal(out.cpu().numpy(), ref.cpu().numpy())
    # backward: gradient of the gathered map from g_fused in place; the destination gradient is a view
    g = torch.from_numpy(rng.standard_normal(tuple(ref.shape), dtype=np.float32)).cuda()
    R, Q = bev_hw[0] * bev_hw[1], img_hw[0] * img_hw[1]
    gd_ref, gs_ref = ops.pool_backward(g.reshape(R, -1), plan.by_pixel(), R, C_b, Q, C_i)
    gs = ops.pool_backward_from(g.reshape(R, -1), plan.by_pixel(), R, Q, C_i, C_b)
    np.testing.assert_array_equal(gs.cpu().numpy(), gs_ref.cpu().numpy())
    np.testing.assert_array_equal(g.reshape(R, -1)[:, :C_b].cpu().numpy(), gd_ref.cpu().numpy())
    # through the drop-in layer with out= and autograd
    buf2 = torch.zeros((1,) + bev_hw + (C_b + C_i,), device="cuda")
    buf2[..., :C_b] = tb
    ti2 = ti.clone().requires_grad_(True)
    fused, same_img = shpl.sparse_pool_layer([buf2[..., :C_b], ti2], [C_i, C_b], o, img_index_flip=o["img_index_flip_pool"], out=(buf2, None))
    assert same_img is ti2
    np.testing.assert_array_equal(fused.detach().cpu().numpy(), ref.cpu().numpy())
    fused.backward(g)
    np.testing.assert_array_equal(ti2.grad.reshape(Q, -1).cpu().numpy(), gs_ref.cpu().numpy())


def test_no_concat_dual_forward_in_one_launch(shpl):
    from sparse_pooling_b200 import _cabi, ops
    bev_hw, img_hw, C = (44, 50), (23, 75), 64
    d = synth.direct_pairs(2, 4000, bev_hw, (img_hw[1], img_hw[0]))
    o = shpl.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()})
    plan = o["shpl_plan"]
    rng = np.random.default_rng(6)
    tb = torch.from_numpy(rng.standard_normal((1,) + bev_hw + (C,), dtype=np.float32)).cuda()
    ti = torch.from_numpy(rng.standard_normal((1,) + img_hw + (C,), dtype=np.float32)).cuda()
    ref_b, ref_i = ops.sparse_pool_dual(tb, ti, plan)
    fb = torch.zeros_like(ref_b)
    fi = torch.zeros_like(ref_i)
    fb[..., :C] = tb
    fi[..., :C] = ti
    rc = _cabi.lib.shpl_pool_forward_into_dual(ops._ptr(tb), ops._ptr(ti), *plan.ptrs8(), int(plan.entry_bound), 0, plan.n_rows, C,
                                               plan.n_src, C, ops._ptr(fb), ops._ptr(fi), ops._stream())
    assert rc == 0
    np.testing.assert_array_equal(fb.cpu().numpy(), ref_b.cpu().numpy())
    np.testing.assert_array_equal(fi.cpu().numpy(), ref_i.cpu().numpy())


def test_value_path_exact_arithmetic_kat_on_the_gpu(shpl):
    """tests/kat_value.py: the hand-derived literals (duplicate rows, duplicate pixels, an empty row, both directions,
    both gradients; integer features and power-of-two weights, so no summation order can change a bit) through the
    drop-in layer and autograd, the no-concat form and the registered custom op."""
    from tests import kat_value as K
    from sparse_pooling_b200 import ops
    M = shpl.SparseTensor(torch.from_numpy(K.MIJ).cuda(), torch.from_numpy(K.VAL).cuda(), K.M_SIZE)
    flip = torch.from_numpy(K.FLIP).cuda()
    for dual in (False, True):
        tb, ti = torch.from_numpy(K.BEV).cuda().requires_grad_(True), torch.from_numpy(K.IMG).cuda().requires_grad_(True)
        bv, im = shpl.sparse_pool_layer([tb, ti], [2, 2], M, img_index_flip=flip, bv_index=(np.zeros((1, 3)) if dual else None))
        np.testing.assert_array_equal(bv.detach().cpu().numpy(), K.FUSED_BEV)
        if dual:
            np.testing.assert_array_equal(im.detach().cpu().numpy(), K.FUSED_IMG)
            torch.autograd.backward([bv, im], [torch.from_numpy(K.G_FUSED_BEV).cuda(), torch.from_numpy(K.G_FUSED_IMG).cuda()])
            np.testing.assert_array_equal(tb.grad.cpu().numpy(), K.G_BEV_DUAL)
            np.testing.assert_array_equal(ti.grad.cpu().numpy(), K.G_IMG_DUAL)
        else:
            bv.backward(torch.from_numpy(K.G_FUSED_BEV).cuda())
            np.testing.assert_array_equal(tb.grad.cpu().numpy(), K.G_BEV_SINGLE)
            np.testing.assert_array_equal(ti.grad.cpu().numpy(), K.G_IMG_SINGLE)
    # bare ops
    np.testing.assert_array_equal(shpl._sparse_pool_op(M, torch.from_numpy(K.IMG).cuda(), flip, [1, 2, 2, 2]).cpu().numpy(), K.FUSED_BEV[..., 2:])
    np.testing.assert_array_equal(shpl._sparse_pool_trans_op(M, torch.from_numpy(K.BEV).cuda(), flip, [1, 2, 3, 2]).cpu().numpy(), K.FUSED_IMG[..., 2:])
    # the no-concat form
    plan = M._coo_plans[(4, (2, 3))]
    buf = torch.zeros((1, 2, 2, 4), device="cuda")
    buf[..., :2] = torch.from_numpy(K.BEV).cuda()
    np.testing.assert_array_equal(ops.sparse_pool_into(buf, torch.from_numpy(K.IMG).cuda(), plan).cpu().numpy(), K.FUSED_BEV)
