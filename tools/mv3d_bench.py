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
from sparse_pooling_b200 import construct_voxel as cv  # noqa: E402


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
    print(json.dumps(out))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
