"""Times the MV3D voxel feeder (shpl_mv3d_voxelize) on cuda:0 as CUDA-graph replays with CUDA events, next to
the numpy oracle.  Prints one JSON object.  Usage: python tools/mv3d_bench.py [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import feeder_oracle as fo  # noqa: E402  (CPU timing beside the GPU number only)
from tools import synth  # noqa: E402
from sparse_pooling_b200 import construct_voxel as cv, group_pointcloud as gp, ops  # noqa: E402


def time_graph(fn, reps=50):
    """us per call of fn() as CUDA-graph replays, CUDA events on the replay stream"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def vfe_scatter_bench(dev, peak_gbs):
    """tf.scatter_nd of the voxel-wise features into the dense grid (group_pointcloud.py:84-85): plan, forward,
    gradient.  The forward writes the whole grid (245 MB ped, 721 MB car: larger than L2, no flush needed)."""
    out = {}
    for name, n, kw in (("ped_20k", 20000, {}), ("car_60k", 60000, dict(car=True))):
        f = synth.mv3d_frame(seed=21, n_points=n, **kw)
        cv.MAX_NUM_POINTS = f["max_points"]          # 45 ped / cyc, 35 car (config_voxels.py:43, :59)
        try:
            vd, vfs, _, _, _ = cv.point_cloud_2_top_sparse(torch.from_numpy(synth.mv3d_cam4(f)).to(dev), res=f["res"], zres=f["zres"],
                                                           side_range=f["side_range"], fwd_range=f["fwd_range"],
                                                           height_range=f["height_range"], points_in_cam=True,
                                                           img_index2=torch.from_numpy(f["img_index2"]).to(dev))
        finally:
            cv.MAX_NUM_POINTS = 45
        coord = vd["coordinate_buffer"].contiguous()
        K, C = int(coord.shape[0]), 128
        grid = tuple(int(x) for x in vfs)
        R = grid[0] * grid[1] * grid[2]
        x = torch.randn((K, C), device=dev)
        plan = gp.voxel_scatter_plan(coord, 1, grid, read_counts=True)
        res = {"voxels": K, "grid": [1, *grid, C], "grid_mbytes": R * C * 4 / 1e6}
        res["us_plan"] = time_graph(lambda: gp.voxel_scatter_plan(coord, 1, grid, read_counts=False))
        res["us_forward"] = time_graph(lambda: ops.pool_forward(None, x, plan.by_row(), R, K))
        g = torch.randn((R, C), device=dev)
        res["us_backward"] = time_graph(lambda: ops.pool_backward(g, plan.by_pixel(), R, 0, K, C, want_dst=False))
        fwd_bytes = 4 * (R * C + K * (C + 3) + R + 1)
        res["forward_algorithmic_mbytes"] = fwd_bytes / 1e6
        res["forward_gbs"] = fwd_bytes / res["us_forward"] / 1e3
        res["forward_frac_of_measured_peak"] = res["forward_gbs"] / peak_gbs
        res["us_bare_memset_of_the_grid"] = time_graph(lambda: torch.zeros((R, C), device=dev))
        out[name] = res
    return out


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for name, n, kw in (("ped_20k", 20000, {}), ("ped_120k", 120000, {}), ("car_60k", 60000, dict(car=True))):
        f = synth.mv3d_frame(seed=21, n_points=n, **kw)
        T = f["max_points"]
        pts = torch.from_numpy(synth.mv3d_cam4(f)).to(dev)
        img2 = torch.from_numpy(f["img_index2"]).to(dev)
        res = {"points": n, "max_points": T}
        for with_buffers in (True, False):
            bufs = dict(img_index=torch.empty((3, n), dtype=torch.int64, device=dev),
                        bv_index=torch.empty((n, 2), dtype=torch.int64, device=dev),
                        m_val=torch.empty(n, dtype=torch.float64, device=dev),
                        counts=torch.zeros(8, dtype=torch.int32, device=dev), ws=cv._workspace(dev, n))
            if with_buffers:
                bufs.update(feature=torch.empty((n, T, 7), dtype=torch.float64, device=dev),
                            coordinate=torch.empty((n, 4), dtype=torch.int64, device=dev),
                            number=torch.empty(n, dtype=torch.int64, device=dev))

            def call():
                cv.voxelize_raw(pts, img2, n, f["res"], f["zres"], f["side_range"], f["fwd_range"], f["height_range"], T, bufs)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                call()
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res["us_with_voxel_buffers" if with_buffers else "us_pairs_only"] = e0.elapsed_time(e1) * 20.0
            c = bufs["counts"].cpu().tolist()
            res["in_range"], res["pairs"], res["voxels"] = c[0], c[1], c[2]
        t0 = time.perf_counter()
        fo.point_cloud_2_top_sparse(synth.mv3d_cam4(f), f["img_index2"], f["res"], f["zres"], f["side_range"], f["fwd_range"],
                                    f["height_range"], T)
        res["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        out[name] = res
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out["vfe_scatter"] = vfe_scatter_bench(dev, peak)
    print(json.dumps(out))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
