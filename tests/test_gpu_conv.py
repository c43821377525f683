"""GPU parity tests of the fused post-fusion 3x3 convolution: shpl_pool_conv3x3_forward through the
C ABI against the CPU oracle -- oracle.value_oracle.conv3x3_after_fusion (float64) over the value oracle's fused map, the
restatement of  sparse_pool_layer -> slim.conv2d(fused, C, [3, 3])  (rpn_model.py:335-346).

The arithmetic of that conv lives in TensorFlow / cuDNN (parity unpinned, like the rest of the value path); the bound
asserted here is the north star's 1e-5, relative to the sum of |terms| of each output (what an fp32 accumulation's
error is proportional to): |out - ref| <= 1e-5 * sum |x_i w_i|.  Measured: <= 4e-7."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import cref, index_oracle as io, synth, value_oracle as vo  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


def _pairs(seed, n, H, W, Hi, Wi, dup_rows=False):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, H * W, n)
    if dup_rows and n:                       # several pairs per cell and neighbouring busy cells
        rows[: n // 2] = rows[n // 2: 2 * (n // 2)]
        rows[: n // 4] = np.clip(rows[n // 4: 2 * (n // 4)] + 1, 0, H * W - 1)
    flip = np.stack([np.zeros(n, np.int64), rng.integers(0, Hi, n), rng.integers(0, Wi, n)], 1)
    Mij = np.stack([rows, np.arange(n)], 1).astype(np.int64)
    val = (rng.random(n) + 0.5).astype(np.float32)
    return Mij, val, flip


def _run(shpl, B, H, W, Hi, Wi, n, seed, relu, affine, pooled=True, dup_rows=False, Ci=32):
    from sparse_pooling_b200 import conv_fusion
    rng = np.random.default_rng(seed)
    bev = rng.standard_normal((B, H, W, 32), dtype=np.float32)
    img = rng.standard_normal((B, Hi, Wi, Ci), dtype=np.float32)
    w = (rng.standard_normal((3, 3, 32 + Ci if pooled else 32, 32)) * 0.1).astype(np.float32)
    scale = (rng.random(32) + 0.5).astype(np.float32) if affine else None
    shift = rng.standard_normal(32).astype(np.float32) if affine else None
    tb, ti, tw = (torch.from_numpy(x).cuda() for x in (bev, img, w))
    ts = None if scale is None else torch.from_numpy(scale).cuda()
    th = None if shift is None else torch.from_numpy(shift).cuda()
    if pooled:
        assert B == 1
        Mij, val, flip = _pairs(seed + 1, n, H, W, Hi, Wi, dup_rows)
        M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), np.array([H * W, n]))
        out = conv_fusion.sparse_pool_conv3x3([tb, ti], M, torch.from_numpy(flip).cuda(), tw, ts, th, relu)
        fused = (cref.forward(bev[0], img[0], Mij, val, flip)[None] if n
                 else np.concatenate([bev, np.zeros(bev.shape[:3] + (Ci,), np.float32)], axis=3))
    else:
        out = conv_fusion.sparse_pool_conv3x3([tb, None], None, None, tw, ts, th, relu)
        fused = bev
    ref, mag = vo.conv3x3_after_fusion(fused, w, scale, shift, relu)
    err = np.abs(out.cpu().numpy().astype(np.float64) - ref)
    rel = float((err / np.maximum(mag, 1e-30)).max())
    assert rel <= TOL, "max |err| / sum|terms| = %.3e" % rel
    return rel


@pytest.mark.parametrize("B,H,W", [(1, 8, 14), (1, 16, 8), (1, 9, 15), (2, 50, 44), (3, 7, 100), (1, 1, 1)])
def test_dense_half_alone_matches_the_oracle(shpl, B, H, W):
    """No pooled channels: a plain SAME-padded 3x3 conv of the BEV map, tile-aligned and ragged sizes, batches."""
    _run(shpl, B, H, W, 4, 4, 0, 10 + H, relu=False, affine=False, pooled=False)
    _run(shpl, B, H, W, 4, 4, 0, 20 + H, relu=True, affine=True, pooled=False)


@pytest.mark.parametrize("n,dup", [(0, False), (1, False), (300, False), (3000, True), (129, True)])
def test_fused_conv_matches_pool_then_conv(shpl, n, dup):
    """conv(concat(bev, pooled)) with the fused map never written == the oracle's pool -> concat -> conv."""
    _run(shpl, 1, 50, 44, 20, 30, n, 30 + n, relu=False, affine=False, dup_rows=dup)
    _run(shpl, 1, 50, 44, 20, 30, n, 40 + n, relu=True, affine=True, dup_rows=dup)


def test_fused_conv_at_the_kitti_size(shpl):
    """700x800x(32+32 -> 32) <- 360x1200x32, 20 000 pairs: the pre-RPN layer of the pyramid people config."""
    rel = _run(shpl, 1, 700, 800, 360, 1200, 20000, 5, relu=True, affine=True)
    print("KITTI size: max |err| / sum|terms| = %.2e" % rel)


@pytest.mark.parametrize("H,W,n,dup", [(200, 176, 30000, False), (200, 176, 60000, True), (120, 96, 45000, True)])
def test_fused_conv_with_many_pairs(shpl, H, W, n, dup):
    """More pairs than one wave of the Z kernel's tiles holds (rows_per_tile at its cap of 128, several tiles per CTA) and
    the tile sizes in between (30 000 pairs: 104 entries per tile), on maps where most cells receive several pairs."""
    _run(shpl, 1, H, W, 60, 90, n, 7 + n, relu=False, affine=True, dup_rows=dup)


@pytest.mark.parametrize("n,dup", [(0, False), (700, False), (5000, True)])
def test_fused_conv_with_64_pooled_channels(shpl, n, dup):
    """C_s = 64 (the FFMA form of the Z kernel and the separate marking kernel)."""
    _run(shpl, 1, 50, 44, 20, 30, n, 90 + n, relu=True, affine=True, dup_rows=dup, Ci=64)


def test_fused_conv_through_the_builder_plan(shpl):
    """The plan the correspondence builder leaves (produce_sparse_pooling_input's dict) goes straight in."""
    from sparse_pooling_b200 import conv_fusion
    d = synth.direct_pairs(3, 5000, (88, 100), (150, 45))
    o = shpl.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()}, stride=[1, 1])
    rng = np.random.default_rng(2)
    bev = rng.standard_normal((1, 88, 100, 32), dtype=np.float32)
    img = rng.standard_normal((1, 45, 150, 32), dtype=np.float32)
    w = (rng.standard_normal((3, 3, 64, 32)) * 0.1).astype(np.float32)
    out = conv_fusion.sparse_pool_conv3x3([torch.from_numpy(bev).cuda(), torch.from_numpy(img).cuda()], o,
                                          o["img_index_flip_pool"], torch.from_numpy(w).cuda())
    oref = io.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()}, stride=(1, 1))
    fused = cref.forward(bev[0], img[0], oref["Mij_pool"], np.ones(len(oref["Mij_pool"]), np.float32), oref["img_index_flip_pool"])[None]
    ref, mag = vo.conv3x3_after_fusion(fused, w)
    assert float((np.abs(out.cpu().numpy() - ref) / np.maximum(mag, 1e-30)).max()) <= TOL


def test_unsupported_shapes_are_refused_not_miscomputed(shpl):
    from sparse_pooling_b200 import conv_fusion
    bev = torch.zeros((1, 8, 8, 16), device="cuda")
    w = torch.zeros((3, 3, 16, 16), device="cuda")
    with pytest.raises(ValueError):
        conv_fusion.sparse_pool_conv3x3([bev, None], None, None, w)


@pytest.mark.parametrize("H,W,n,dup", [(50, 44, 3000, True), (23, 30, 0, False), (96, 112, 6000, False), (64, 80, 20000, True)])
def test_fused_conv_backward_matches_the_oracle(shpl, H, W, n, dup):
    """g_bev, g_img and g_weight of conv3x3(concat(bev, pooled(img)), W) through autograd against the float64 oracle
    (conv gradients of the concat form, then the pooling gradient of SURVEY.md a13), each within 1e-5 of its sum of |terms|."""
    from sparse_pooling_b200 import conv_fusion
    Hi, Wi = 20, 30
    rng = np.random.default_rng(7 + n)
    bev = rng.standard_normal((1, H, W, 32), dtype=np.float32)
    img = rng.standard_normal((1, Hi, Wi, 32), dtype=np.float32)
    w = (rng.standard_normal((3, 3, 64, 32)) * 0.1).astype(np.float32)
    g = rng.standard_normal((1, H, W, 32), dtype=np.float32)
    Mij, val, flip = _pairs(3, max(n, 1), H, W, Hi, Wi, dup)
    if n == 0:
        Mij, val, flip = Mij[:0], val[:0], flip[:0]
    tb, ti, tw = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (bev, img, w))
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), np.array([H * W, len(val)]))
    out = conv_fusion.sparse_pool_conv3x3_autograd([tb, ti], M, torch.from_numpy(flip).cuda(), tw)
    out.backward(torch.from_numpy(g).cuda())
    fused = cref.forward(bev[0], img[0], Mij, val, flip)[None] if n else np.concatenate([bev, np.zeros_like(bev)], axis=3)
    ref, mag = vo.conv3x3_same(fused, w)
    assert float((np.abs(out.detach().cpu().numpy() - ref) / np.maximum(mag, 1e-30)).max()) <= TOL
    g_x, g_w, mag_x, mag_w = vo.conv3x3_same_grad(fused, w, g)
    assert float((np.abs(tb.grad.cpu().numpy() - g_x[..., :32]) / np.maximum(mag_x[..., :32], 1e-30)).max()) <= TOL
    assert float((np.abs(tw.grad.cpu().numpy() - g_w) / np.maximum(mag_w, 1e-30)).max()) <= TOL
    # pooled-path gradient -> image map: g_img[p] = sum_k val_k * g_pooled[row_k]  (float64, then compared at 1e-5 of the term sum)
    gp, mp = g_x[0, :, :, 32:].reshape(-1, 32), mag_x[0, :, :, 32:].reshape(-1, 32)
    g_img = np.zeros((Hi * Wi, 32))
    m_img = np.zeros((Hi * Wi, 32))
    pix = flip[:, 1] * Wi + flip[:, 2]
    np.add.at(g_img, pix, val[:, None].astype(np.float64) * gp[Mij[:, 0]])
    np.add.at(m_img, pix, np.abs(val[:, None].astype(np.float64)) * mp[Mij[:, 0]])
    err = np.abs(ti.grad.cpu().numpy().reshape(-1, 32) - g_img)
    assert float((err / np.maximum(m_img, 1e-30))[m_img > 0].max() if (m_img > 0).any() else 0.0) <= TOL
    assert float(err[m_img == 0].max() if (m_img == 0).any() else 0.0) == 0.0


def test_fused_conv_two_frames_stacked_plan(shpl):
    """A batch of two frames behind one stacked plan (build_avod_plan: rows / pixels of frame f offset by f * H * W and
    f * H_i * W_i): forward and all three gradients against the per-frame oracle."""
    from sparse_pooling_b200 import conv_fusion
    frames = [synth.avod_frame(s, az_step_deg=0.4) for s in (11, 12)]
    H, W, Hi, Wi = 175, 200, 90, 300
    plan = shpl.build_avod_plan([f["points"] for f in frames], [f["voxel_indices"] for f in frames],
                                [f["P"] for f in frames], [1200, 360], (700, 800), stride=(4, 4))
    rng = np.random.default_rng(21)
    bev = rng.standard_normal((2, H, W, 32), dtype=np.float32)
    img = rng.standard_normal((2, Hi, Wi, 32), dtype=np.float32)
    w = (rng.standard_normal((3, 3, 64, 32)) * 0.1).astype(np.float32)
    g = rng.standard_normal((2, H, W, 32), dtype=np.float32)
    tb, ti, tw = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (bev, img, w))
    out = conv_fusion.sparse_pool_conv3x3_autograd([tb, ti], plan, None, tw)
    out.backward(torch.from_numpy(g).cuda())
    fused, pairs = [], []
    for f, fr in enumerate(frames):
        d = io.gen_sparse_pooling_input_avod(fr["points"], fr["voxel_indices"], fr["P"], [1200, 360], (700, 800))
        o = io.produce_sparse_pooling_input(d, stride=[4, 4])
        Mij, flip = np.asarray(o["Mij_pool"]), np.asarray(o["img_index_flip_pool"])
        val = np.ones(len(Mij), np.float32)
        assert len(Mij) > 100
        fused.append(cref.forward(bev[f], img[f], Mij, val, flip))
        pairs.append((Mij, val, flip))
    fused = np.stack(fused)
    ref, mag = vo.conv3x3_same(fused, w)
    assert float((np.abs(out.detach().cpu().numpy() - ref) / np.maximum(mag, 1e-30)).max()) <= TOL
    g_x, g_w, mag_x, mag_w = vo.conv3x3_same_grad(fused, w, g)
    assert float((np.abs(tb.grad.cpu().numpy() - g_x[..., :32]) / np.maximum(mag_x[..., :32], 1e-30)).max()) <= TOL
    assert float((np.abs(tw.grad.cpu().numpy() - g_w) / np.maximum(mag_w, 1e-30)).max()) <= TOL
    for f, (Mij, val, flip) in enumerate(pairs):
        gp, mp = g_x[f, :, :, 32:].reshape(-1, 32), mag_x[f, :, :, 32:].reshape(-1, 32)
        g_img = np.zeros((Hi * Wi, 32))
        m_img = np.zeros((Hi * Wi, 32))
        pix = flip[:, 1] * Wi + flip[:, 2]
        np.add.at(g_img, pix, val[:, None].astype(np.float64) * gp[Mij[:, 0]])
        np.add.at(m_img, pix, np.abs(val[:, None].astype(np.float64)) * mp[Mij[:, 0]])
        err = np.abs(ti.grad[f].cpu().numpy().reshape(-1, 32) - g_img)
        assert float((err / np.maximum(m_img, 1e-30))[m_img > 0].max()) <= TOL
        assert float(err[m_img == 0].max()) == 0.0


def test_fused_conv_backward_with_outputs_left_out(shpl):
    """Any of the three gradients may be left out (NULL in the C ABI); the ones asked for do not change by a bit."""
    from sparse_pooling_b200 import conv_fusion
    from sparse_pooling_b200.sparse_pool_utils import _resolve_plan
    H, W, Hi, Wi, n = 64, 80, 20, 30, 6000
    rng = np.random.default_rng(5)
    tb = torch.from_numpy(rng.standard_normal((1, H, W, 32), dtype=np.float32)).cuda()
    ti = torch.from_numpy(rng.standard_normal((1, Hi, Wi, 32), dtype=np.float32)).cuda()
    tw = torch.from_numpy((rng.standard_normal((3, 3, 64, 32)) * 0.1).astype(np.float32)).cuda()
    tg = torch.from_numpy(rng.standard_normal((1, H, W, 32), dtype=np.float32)).cuda()
    Mij, val, flip = _pairs(9, n, H, W, Hi, Wi, True)
    M = shpl.SparseTensor(torch.from_numpy(Mij).cuda(), torch.from_numpy(val).cuda(), np.array([H * W, n]))
    plan = _resolve_plan(M, torch.from_numpy(flip).cuda(), H * W, (Hi, Wi), tb.device)
    full = conv_fusion.sparse_pool_conv3x3_backward(tg, [tb, ti], plan, tw)
    for need in ((True, False, False), (False, True, False), (False, False, True), (True, True, False), (False, True, True)):
        part = conv_fusion.sparse_pool_conv3x3_backward(tg, [tb, ti], plan, tw, *need)
        for want, a, b in zip(need, part, full):
            assert (a is None) == (not want)
            if want:
                assert torch.equal(a, b)
