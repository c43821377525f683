"""Times the BEV slicing feeder (shpl_bev_slices) on cuda:0: CUDA-event time per call, back to back,
for a few synthetic scans.  Prints one JSON object.  Usage: python tools/feeder_bench.py [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import feeder_oracle as fo  # noqa: E402  (CPU timing beside the GPU number only)
from tools import synth  # noqa: E402
from sparse_pooling_b200 import bev_slices as bs  # noqa: E402

GP = np.array([0.0, -1.0, 0.0, 1.65])


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for name, seed, az in (("scan_30k", 1, 0.09), ("scan_55k", 2, 0.05), ("scan_137k", 12, 0.02)):
        pts = synth.lidar_scan(seed, az_step_deg=az)
        P = pts.shape[0]
        t = torch.from_numpy(np.ascontiguousarray(pts.T)).to(dev)
        lut = torch.from_numpy(bs.density_lut(np.log(16))).to(dev)
        res = {"points": int(P)}
        for with_maps in (True, False):
            work = bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, 5 * P, dev, with_maps=with_maps)

            def call():
                bs.bev_slices_raw(t, t.stride(0), t.stride(1), P, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5,
                                  np.log(16), work, lut=lut)
            for _ in range(5):
                call()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                call()
            for _ in range(5):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(100):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res["us_with_maps" if with_maps else "us_indices_only"] = e0.elapsed_time(e1) * 10.0
            res["pairs"] = int(work.counts[0].item())
        t0 = time.perf_counter()
        fo.generate_bev(pts.T, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
        res["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        out[name] = res
    # the ingest in front of the feeder (shpl_lidar_to_cam): raw velodyne scans
    import types
    from sparse_pooling_b200 import lidar_ingest as li
    cal = types.SimpleNamespace(p2=synth.P2_KITTI, r0_rect=synth.R0_RECT_KITTI, tr_velodyne_to_cam=synth.TR_VELO_TO_CAM_KITTI)
    for name, seed, az in (("velodyne_118k", 1, 0.18), ("velodyne_235k", 2, 0.09)):
        scan = synth.velodyne_scan(seed, az_step_deg=az)
        n = scan.shape[0]
        velo = torch.from_numpy(scan).to(dev)
        cam = torch.empty((3, n), dtype=torch.float64, device=dev)
        counts = torch.zeros(4, dtype=torch.int32, device=dev)

        def call():
            li.lidar_to_cam_raw(velo, n, cal, [1242, 375], cam, counts)
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            call()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fo.get_lidar_point_cloud(scan, cal.p2, cal.r0_rect, cal.tr_velodyne_to_cam, im_size=[1242, 375])
        out[name] = {"points": int(n), "fov_points": int(counts[0].item()), "us_ingest": e0.elapsed_time(e1) * 10.0,
                     "cpu_oracle_ms": (time.perf_counter() - t0) * 1e3}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
