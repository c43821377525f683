"""GPU parity tests of the BEV slicing feeder (shpl_bev_slices, through the drop-in BevSlices class and the
raw C ABI) against the feeder oracle and the fixtures the REFERENCE's BevSlices.generate_bev produced
(tests/golden/bev_slices_seed*.npz).  Everything here is index / selection work or exactly rounded fp64
arithmetic, so every comparison is bit-exact.  Run with `pytest -m gpu` on a B200."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import feeder_oracle as fo, index_oracle as io, synth  # noqa: E402

pytestmark = pytest.mark.gpu

GP = np.array([0.0, -1.0, 0.0, 1.65])
CFG = types.SimpleNamespace(height_lo=-0.2, height_hi=2.3, num_slices=5)


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def scan(seed, az):
    return synth.lidar_scan_gappy(seed) if az is None else synth.lidar_scan(seed, az_step_deg=az)


def assert_maps_equal(got_maps, ref_hms, ref_dm):
    for i, hm in enumerate(ref_hms):
        np.testing.assert_array_equal(got_maps['height_maps'][i], hm, err_msg="height map %d" % i)
    np.testing.assert_array_equal(got_maps['density_map'], ref_dm, err_msg="density map")


# ------------------------------------------------------------------ against the reference-run fixtures
@pytest.mark.parametrize("seed,az", [(1, 0.09), (2, 0.05), (3, None)])
def test_bev_slices_matches_reference_fixture(shpl, golden_dir, seed, az):
    g = load(golden_dir, "bev_slices_seed%d.npz" % seed)
    pts = scan(seed, az)
    maps, idx, upts = shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                            output_indices=True)
    assert isinstance(idx, np.ndarray) and idx.dtype == np.int64 and upts.dtype == np.float64
    np.testing.assert_array_equal(idx, g["voxel_indices"])
    np.testing.assert_array_equal(upts, g["unique_pts"])
    assert len(maps['height_maps']) == 5
    for i, hm in enumerate(maps['height_maps']):
        assert hm.shape == (700, 800) and hm.dtype == np.float64
        nz = np.nonzero(hm)
        np.testing.assert_array_equal(np.stack(nz, axis=1), g["hm%d_idx" % i])
        np.testing.assert_array_equal(hm[nz], g["hm%d_val" % i])
    dm = maps['density_map']
    nz = np.nonzero(dm)
    np.testing.assert_array_equal(np.stack(nz, axis=1), g["dm_idx"])
    np.testing.assert_array_equal(dm[nz], g["dm_val"])            # bit-exact through the host-tabulated log
    # without output_indices only the maps come back (bev_slices.py:155-156)
    only = shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL)
    assert set(only) == {"height_maps", "density_map"}


# ------------------------------------------------------------------ against the oracle, other shapes
@pytest.mark.parametrize("case", ["tilted_plane", "coarse_voxels", "dense_ties", "three_slices", "wide_words", "flat_y"])
def test_bev_slices_matches_oracle(shpl, case):
    rng = np.random.default_rng(11)
    gp, ext, vox, cfg = GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, CFG
    if case == "tilted_plane":
        pts = synth.lidar_scan(5, az_step_deg=0.2)
        gp = np.array([0.012, -0.9995, 0.021, 1.61])
    elif case == "coarse_voxels":
        pts = synth.lidar_scan(6, az_step_deg=0.2)
        vox = 0.4
    elif case == "dense_ties":
        # 60k points in a 3 m x 3 m patch on a 0.05 m lattice: many points share (cell, y bin); the lowest index must win
        pts = np.stack((rng.integers(-30, 30, 60000) * 0.05 + 0.013, 1.65 - rng.integers(0, 40, 60000) * 0.05 - 0.011,
                        rng.integers(200, 260, 60000) * 0.05 + 0.017), axis=1)
    elif case == "wide_words":
        # a y extent of 10^7 m needs 27 bits of y bin: the kernels fall back to 64-bit winner words
        pts = synth.lidar_scan(10, az_step_deg=0.2)
        ext = np.array([[-40.0, 40.0], [-1.0e7, 3.0], [0.0, 70.0]])
    elif case == "flat_y":
        # every point in one y bin: the winner is decided by the point index alone
        pts = synth.lidar_scan(10, az_step_deg=0.2)
        pts = pts[(pts[:, 1] > 1.601) & (pts[:, 1] < 1.699)]
        ext = np.array([[-40.0, 40.0], [1.6001, 1.6999], [0.0, 70.0]])
        cfg = types.SimpleNamespace(height_lo=-0.2, height_hi=0.3, num_slices=2)
    else:
        pts = synth.lidar_scan(7, az_step_deg=0.2)
        cfg = types.SimpleNamespace(height_lo=0.0, height_hi=1.5, num_slices=3)
        ext = np.array([[-20.0, 20.0], [-4.0, 2.0], [5.0, 45.0]])
    hms, dm, idx, upts = fo.generate_bev(pts.T, gp, ext, vox, cfg.height_lo, cfg.height_hi, cfg.num_slices)
    maps, gidx, gupts = shpl.BevSlices(cfg, None).generate_bev("lidar", pts.T, gp, ext, vox, output_indices=True)
    np.testing.assert_array_equal(gidx, idx)
    np.testing.assert_array_equal(gupts, upts)
    assert_maps_equal(maps, hms, dm)


def test_bev_slices_cuda_tensors_and_strided_input(shpl):
    """torch CUDA in -> torch CUDA out; a (3,N) view of an [N,3] buffer is read through its strides."""
    pts = synth.lidar_scan(8, az_step_deg=0.15)
    hms, dm, idx, upts = fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    dev = torch.device("cuda", 0)
    t = torch.from_numpy(pts).to(dev)                     # [N,3] contiguous
    maps, gidx, gupts = shpl.BevSlices(CFG, None).generate_bev("lidar", t.t(), GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                              output_indices=True)
    assert gidx.is_cuda and gupts.is_cuda and maps['density_map'].is_cuda
    np.testing.assert_array_equal(gidx.cpu().numpy(), idx)
    np.testing.assert_array_equal(gupts.cpu().numpy(), upts)
    t2 = t.t().contiguous()                               # [3,N] contiguous
    _, gidx2, gupts2 = shpl.BevSlices(CFG, None).generate_bev("lidar", t2, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                             output_indices=True)
    assert torch.equal(gidx, gidx2) and torch.equal(gupts, gupts2)
    np.testing.assert_array_equal(maps['height_maps'][2].cpu().numpy(), hms[2])
    np.testing.assert_array_equal(maps['density_map'].cpu().numpy(), dm)


def test_bev_slices_density_formula_without_table(shpl):
    """density_lut = NULL: the kernel evaluates min(1, log(n+1)/norm) itself (CUDA log: within 1 ulp of numpy's)."""
    from sparse_pooling_b200 import bev_slices as bs
    pts = synth.lidar_scan(9, az_step_deg=0.2)
    _, dm, idx, _ = fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    dev = torch.device("cuda", 0)
    t = torch.from_numpy(np.ascontiguousarray(pts.T)).to(dev)
    work = bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, 5 * pts.shape[0], dev)
    bs.bev_slices_raw(t, t.stride(0), t.stride(1), pts.shape[0], GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5,
                      np.log(16), work, lut=None)
    got = work.maps[5].cpu().numpy()
    assert int(work.counts[0].item()) == len(idx)
    np.testing.assert_array_equal(got != 0, dm != 0)
    np.testing.assert_allclose(got, dm, rtol=4e-16, atol=0)       # 1 ulp of fp64 (stated tolerance), exact where it saturates
    assert np.array_equal(got[dm == 1.0], dm[dm == 1.0])


# ------------------------------------------------------------------ error behaviour of the reference
def test_bev_slices_first_slice_empty_raises_nameerror(shpl):
    """bev_slices.py:79,93: with <= 1 point in the first slice `voxel_grid_2d` is unbound -> NameError."""
    pts = synth.lidar_scan(1, az_step_deg=0.3)
    pts = pts[(1.65 - pts[:, 1]) > 0.5]
    with pytest.raises(NameError):
        fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    with pytest.raises(NameError):
        shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, output_indices=True)


def test_bev_slices_rejects_bad_arguments(shpl):
    pts = synth.lidar_scan(1, az_step_deg=0.5)
    with pytest.raises(ValueError):                      # (N,3) instead of (3,N)
        shpl.BevSlices(CFG, None).generate_bev("lidar", pts, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL)
    with pytest.raises(ValueError):
        shpl.BevSlices(types.SimpleNamespace(height_lo=0.0, height_hi=1.0, num_slices=9), None).generate_bev(
            "lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL)
    with pytest.raises(RuntimeError):                    # CPU torch tensor: no fallback
        shpl.BevSlices(CFG, None).generate_bev("lidar", torch.from_numpy(np.ascontiguousarray(pts.T)), GP,
                                               synth.AVOD_EXTENTS, synth.AVOD_VOXEL)


# ------------------------------------------------------------------ feeder -> builder chain (kitti_dataset.py:356-378)
def test_feeder_into_builder_matches_the_oracle_chain(shpl):
    pts = synth.lidar_scan(2, az_step_deg=0.05)

    class Calib:
        p2 = synth.P2_KITTI
    maps, vox, upts = shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                            output_indices=True)
    d = shpl.gen_sparse_pooling_input_avod(upts, vox, Calib, [1200, 360], maps['density_map'].shape[0:2])
    out = shpl.produce_sparse_pooling_input(d, stride=[4, 4])
    _, _, idx, rupts = fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    d_ref = io.gen_sparse_pooling_input_avod(rupts, idx, synth.P2_KITTI, [1200, 360], (700, 800))
    o_ref = io.produce_sparse_pooling_input(d_ref, stride=[4, 4])
    for k in ("Mij_pool", "M_size", "img_index_flip_pool"):
        np.testing.assert_array_equal(np.asarray(out[k]), o_ref[k], err_msg=k)


def test_feeder_device_count_flows_into_the_builder_without_a_host_read(shpl):
    """shpl_bev_slices counts[0] handed to shpl_build_avod as N_dev: same plan as with the host-known count."""
    from sparse_pooling_b200 import bev_slices as bs
    from sparse_pooling_b200.pipeline import FramePipeline, LayerSpec
    pts = synth.lidar_scan(4, az_step_deg=0.1)
    dev = torch.device("cuda", 0)
    t = torch.from_numpy(np.ascontiguousarray(pts.T)).to(dev)
    cap = 5 * pts.shape[0]
    work = bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, cap, dev, with_maps=False)
    work.voxel_indices.fill_(123456)                     # stale rows beyond the count must be ignored
    work.unique_pts.fill_(7.0)
    bs.bev_slices_raw(t, t.stride(0), t.stride(1), pts.shape[0], GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5,
                      np.log(16), work)
    spec = LayerSpec("s4", (175, 200), (90, 300), 8, 8, (4, 4), False, (1200, 360), (700, 800))
    stream = torch.cuda.current_stream().cuda_stream
    import ctypes
    pipe_dev = FramePipeline([spec], cap, dev)
    pipe_dev.build_layer(0, work.unique_pts, work.voxel_indices, synth.P2_KITTI, cap, stream,
                         n_dev=ctypes.c_void_p(work.counts.data_ptr()))
    n = int(work.counts[0].item())
    assert 0 < n < cap
    pipe_host = FramePipeline([spec], cap, dev)
    pipe_host.build_layer(0, work.unique_pts, work.voxel_indices, synth.P2_KITTI, n, stream)
    torch.cuda.synchronize()
    a, b = pipe_dev.layers[0].plan, pipe_host.layers[0].plan
    ca, cb = a.counts.cpu().numpy()[0], b.counts.cpu().numpy()[0]
    np.testing.assert_array_equal(ca[:4], cb[:4])
    m = int(ca[3])
    assert m > 0
    for name in ("row_ptr", "pix_ptr"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for name in ("csr_row", "csr_src", "csr_val", "csrT_pix", "csrT_dst", "csrT_val"):
        assert torch.equal(getattr(a, name)[:m], getattr(b, name)[:m]), name


# ------------------------------------------------------------------ full-size properties (120k-point scan)
def test_bev_slices_full_scan_properties(shpl):
    """A 64-beam scan too slow to push through every oracle path at test time is checked through invariants the
    domain offers: per-slice cells are unique and (x, z)-sorted, every representative point lies in its cell and
    slice, and the density map's support is the union of the slices' cells."""
    pts = synth.lidar_scan(12, az_step_deg=0.02)
    assert pts.shape[0] > 100000
    maps, idx, upts = shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                            output_indices=True)
    x, zf = idx[:, 0], idx[:, 1]
    z = 700 - zf
    key = x * 700 + z
    starts = np.nonzero(np.diff(key) <= 0)[0] + 1         # a new slice starts where the (x, z) order restarts
    assert len(starts) == 4
    bounds = np.r_[0, starts, len(key)]
    h = 1.65 - upts[:, 1]
    support = np.zeros((700, 800), bool)
    for s in range(5):
        k = key[bounds[s]:bounds[s + 1]]
        assert np.all(np.diff(k) > 0)
        hs = h[bounds[s]:bounds[s + 1]]
        assert np.all(hs > -0.2 + 0.5 * s - 1e-9) and np.all(hs < -0.2 + 0.5 * (s + 1) + 1e-9)
        hm = maps['height_maps'][s]
        assert np.count_nonzero(hm) <= len(k)
        support[699 - z[bounds[s]:bounds[s + 1]], x[bounds[s]:bounds[s + 1]]] = True
    np.testing.assert_array_equal(np.floor(upts[:, 0] / 0.1).astype(np.int64) + 400, x)
    np.testing.assert_array_equal(np.floor(upts[:, 2] / 0.1).astype(np.int64), z)
    np.testing.assert_array_equal(maps['density_map'] > 0, support)
    # and the oracle agrees on this size too (it finishes in about a second)
    _, dm, ridx, rupts = fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(upts, rupts)
    np.testing.assert_array_equal(maps['density_map'], dm)


def test_bev_slices_million_points_uniform_cloud(shpl):
    """1.2 M points spread over the whole grid (975 k occupied (slice, cell) pairs, 28 % of all cells): every emit
    tile is crowded, the look-back runs with arrival tickets and the scatter has heavy same-cell contention."""
    rng = np.random.default_rng(5)
    n = 1200000
    pts = np.stack((rng.uniform(-39.9, 39.9, n), 1.65 - rng.uniform(-0.15, 2.25, n), rng.uniform(0.05, 69.9, n)), axis=1)
    hms, dm, idx, upts = fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    assert len(idx) > 900000
    maps, gidx, gupts = shpl.BevSlices(CFG, None).generate_bev("lidar", pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                              output_indices=True)
    np.testing.assert_array_equal(gidx, idx)
    np.testing.assert_array_equal(gupts, upts)
    assert_maps_equal(maps, hms, dm)
