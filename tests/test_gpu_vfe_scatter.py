"""GPU parity tests of the VFE scatter (group_pointcloud.voxel_scatter: shpl_plan_from_voxel_coords + the pooling
kernels) against the numpy oracle of tf.scatter_nd / its gradient (oracle/value_oracle.py).  fp32 sums in the
reference's order: every comparison is bit-exact.  Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import feeder_oracle as fo, synth, value_oracle as vo  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


def feeder_frame(seed, n=20000, car=False):
    f = synth.mv3d_frame(seed=seed, n_points=n, car=car)
    vd, vfs, _, _, _ = fo.point_cloud_2_top_sparse(synth.mv3d_cam4(f), f["img_index2"], f["res"], f["zres"], f["side_range"],
                                                   f["fwd_range"], f["height_range"], f["max_points"])
    return vd, [int(x) for x in vfs]


def scatter_both_ways(shpl, coordinate, voxelwise, batch, grid, seed, **kw):
    """forward + backward on the GPU; returns (grid tensor, gradient wrt voxelwise, upstream gradient) as numpy"""
    gp = shpl.group_pointcloud
    x = torch.from_numpy(voxelwise).cuda().requires_grad_(True)
    out = gp.voxel_scatter(coordinate, x, batch, grid=grid, **kw)
    g = torch.from_numpy(np.random.default_rng(seed).standard_normal(tuple(out.shape)).astype(np.float32)).cuda()
    out.backward(g)
    return out.detach().cpu().numpy(), x.grad.cpu().numpy(), g.cpu().numpy()


def test_scatter_of_a_feeder_frame_matches_the_oracle(shpl):
    """One ped/cyc sample: the feeder's coordinate_buffer [K,4] = (0, z, x, y) (construct_voxel.py:128, :147) into the
    [1, 10, 200, 240, 128] grid (group_pointcloud.py:81-85), forward and gradient."""
    vd, vfs = feeder_frame(61)
    assert vfs == [10, 200, 240]
    coord = vd["coordinate_buffer"]
    K = coord.shape[0]
    assert K > 10000
    vw = np.random.default_rng(1).standard_normal((K, 128)).astype(np.float32)
    out, gx, g = scatter_both_ways(shpl, coord, vw, 1, vfs, seed=2)
    ref = vo.voxel_scatter(coord, vw, (1, *vfs, 128))
    assert out.shape == (1, 10, 200, 240, 128)
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(gx, vo.voxel_scatter_grad(coord, g))
    assert np.count_nonzero(np.abs(out).sum(axis=-1)) == K          # one cell per voxel, nothing else written


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("C", [8, 36, 128])
def test_duplicates_are_summed_in_row_order_and_strays_are_reported(shpl, dtype, C):
    """tf.scatter_nd sums duplicate indices (k order); a coordinate outside the grid is TF-CPU's InvalidArgumentError
    (ValueError here) or, without the strict check, a dropped row (TF-GPU)."""
    gp = shpl.group_pointcloud
    rng = np.random.default_rng(5 + C)
    grid = (3, 5, 7)
    B, K = 2, 400
    coord = np.stack([rng.integers(0, B, K), rng.integers(0, 3, K), rng.integers(0, 5, K), rng.integers(0, 7, K)], axis=1).astype(dtype)
    vw = rng.standard_normal((K, C)).astype(np.float32)
    cdev = torch.from_numpy(coord).cuda()
    out, gx, g = scatter_both_ways(shpl, cdev, vw, B, grid, seed=3)
    np.testing.assert_array_equal(out, vo.voxel_scatter(coord, vw, (B, *grid, C)))
    np.testing.assert_array_equal(gx, vo.voxel_scatter_grad(coord, g))
    stray = coord.copy()
    stray[7] = (0, 3, 0, 0)
    stray[100] = (-1, 0, 0, 0)
    stray[399] = (B, 2, 4, 6)
    with pytest.raises(ValueError, match="3 coordinates outside"):
        gp.voxel_scatter(torch.from_numpy(stray).cuda(), torch.from_numpy(vw).cuda(), B, grid=grid)
    gp.STRICT_INDEX_CHECK = False
    try:
        out, gx, g = scatter_both_ways(shpl, torch.from_numpy(stray).cuda(), vw, B, grid, seed=4)
    finally:
        gp.STRICT_INDEX_CHECK = True
    np.testing.assert_array_equal(out, vo.voxel_scatter(stray, vw, (B, *grid, C), strict=False))
    np.testing.assert_array_equal(gx, vo.voxel_scatter_grad(stray, g))


def test_batch_from_build_input_and_device_side_count(shpl):
    """build_input (group_pointcloud.py:88-105) over two samples, then the scatter with the voxel count left on the
    device (k_dev): rows past the count do not exist, nothing is read back."""
    gp = shpl.group_pointcloud
    dicts, grids = zip(*(feeder_frame(s, n=6000) for s in (71, 72)))
    B, feature, number, coord = gp.build_input(list(dicts), fix_coordinate_columns=True)   # the feeder emits [K,4]: see build_input
    assert gp.build_input(list(dicts))[3].shape[1] == 5      # the reference's pad, reproduced by default
    assert B == 2 and coord.shape[1] == 4 and feature.shape[0] == number.shape[0] == coord.shape[0]
    K0 = dicts[0]["coordinate_buffer"].shape[0]
    assert (coord[:K0, 0] == 0).all() and (coord[K0:, 0] == 1).all()
    np.testing.assert_array_equal(coord[:, 1:], np.concatenate([d["coordinate_buffer"][:, 1:] for d in dicts]))
    # torch in -> torch out
    tdicts = [{k: torch.from_numpy(v).cuda() for k, v in d.items()} for d in dicts]
    tB, tf_, tn, tc = gp.build_input(tdicts, fix_coordinate_columns=True)
    assert tc.is_cuda and tB == 2
    np.testing.assert_array_equal(tc.cpu().numpy(), coord)
    K = coord.shape[0]
    vw = np.random.default_rng(9).standard_normal((K, 128)).astype(np.float32)
    out, gx, g = scatter_both_ways(shpl, tc, vw, B, grids[0], seed=5)
    np.testing.assert_array_equal(out, vo.voxel_scatter(coord, vw, (B, *grids[0], 128)))
    np.testing.assert_array_equal(gx, vo.voxel_scatter_grad(coord, g))
    # capacity-sized buffers with garbage past the device-side count
    cap = K + 1000
    cbig = torch.full((cap, 4), 3, dtype=torch.int64, device="cuda")
    cbig[:K] = tc
    vbig = np.random.default_rng(10).standard_normal((cap, 128)).astype(np.float32)
    vbig[:K] = vw
    k_dev = torch.tensor([K], dtype=torch.int32, device="cuda")
    gp.STRICT_INDEX_CHECK = False          # no read-back at all
    try:
        out2, gx2, g2 = scatter_both_ways(shpl, cbig, vbig, B, grids[0], seed=5, k_dev=k_dev)
    finally:
        gp.STRICT_INDEX_CHECK = True
    np.testing.assert_array_equal(out2, out)
    np.testing.assert_array_equal(gx2[:K], gx)
    assert not gx2[K:].any()


def test_empty_sample_gives_a_zero_grid(shpl):
    gp = shpl.group_pointcloud
    x = torch.zeros((0, 128), device="cuda", requires_grad=True)
    out = gp.voxel_scatter(torch.zeros((0, 4), dtype=torch.int64, device="cuda"), x, 1, grid=(2, 20, 24))
    assert out.shape == (1, 2, 20, 24, 128) and not out.any()
    out.sum().backward()
    assert x.grad.shape == (0, 128)


def test_car_grid_full_size_properties(shpl):
    """The Car grid [1, 10, 400, 352, 128] (config_voxels.py:33-48; 721 MB): too large for the numpy oracle to be
    quick, so size-independent properties -- every feature row is found at its coordinate, everything else is zero,
    and the gradient is the gather."""
    gp = shpl.group_pointcloud
    vd, vfs = feeder_frame(81, n=20000, car=True)
    assert vfs == [10, 400, 352]
    coord = torch.from_numpy(vd["coordinate_buffer"]).cuda()
    K = coord.shape[0]
    x = torch.randn((K, 128), device="cuda", requires_grad=True)
    out = gp.voxel_scatter(coord, x, 1, grid=vfs)
    assert out.shape == (1, 10, 400, 352, 128)
    got = out[coord[:, 0], coord[:, 1], coord[:, 2], coord[:, 3]]
    assert torch.equal(got, x.detach())                         # unique coordinates: the rows themselves
    assert int((out != 0).any(dim=-1).sum()) == K
    assert float(out.detach().double().sum()) == pytest.approx(float(x.detach().double().sum()), rel=1e-12, abs=1e-9)
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.equal(x.grad, g[coord[:, 0], coord[:, 1], coord[:, 2], coord[:, 3]])
