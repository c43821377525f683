#!/usr/bin/env python
"""Per-piece device timings (CUDA events, one stream, each piece alone in a loop) of the SHPL
path at the bench workload: which kernel the time goes to, without profiler distortion.
    python tools/microbench.py [--iters 50] [--config bench|mv3d|c128]
Prints one JSON object; bench.py remains the contract benchmark."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402
from sparse_pooling_b200.pipeline import FramePipeline, LayerSpec  # noqa: E402


def specs_for(name):
    if name == "bench":
        return [LayerSpec("A_s8_dual_c256", (88, 100), (45, 150), 256, 256, (8, 8), True, (1200, 360), (704, 800)),
                LayerSpec("B_s1_c32", (700, 800), (360, 1200), 32, 32, (1, 1), False, (1200, 360), (700, 800))]
    if name == "b":
        return [LayerSpec("B_s1_c32", (700, 800), (360, 1200), 32, 32, (1, 1), False, (1200, 360), (700, 800))]
    if name == "retina":
        return [LayerSpec("P2_s4_c256", (175, 200), (90, 300), 256, 256, (4, 4), False, (1200, 360), (700, 800))]
    if name == "c128":
        return [LayerSpec("s1_c128", (700, 800), (360, 1200), 128, 128, (1, 1), False, (1200, 360), (700, 800))]
    if name == "c64":
        return [LayerSpec("s1_c64", (700, 800), (360, 1200), 64, 64, (1, 1), False, (1200, 360), (700, 800))]
    raise SystemExit("unknown config " + name)


def timeit(fn, iters, flush=None, reps=4):
    """Median device time of fn() from CUDA-graph replays (no host launch cost in the number);
    with `flush`, an L2-sized buffer is rewritten before every replay (one fn() per replay)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if flush is not None:
        reps = 1
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    ts = np.array(ts)
    return {"us_median": float(np.median(ts)), "us_min": float(ts.min()), "us_mean": float(ts.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--config", default="bench")
    ap.add_argument("--az", type=float, default=0.028)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    specs = specs_for(args.config)
    f = synth.avod_frame(100, az_step_deg=args.az)
    pts = torch.from_numpy(f["points"]).to(dev)
    vox = torch.from_numpy(np.ascontiguousarray(f["voxel_indices"][:, :2])).to(dev)
    n = int(pts.shape[0])
    pipe = FramePipeline(specs, 1 << 17, dev)
    def st():
        return torch.cuda.current_stream().cuda_stream      # the capture stream inside torch.cuda.graph
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    out = {"config": args.config, "candidates": n, "pieces": {}}
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for i, s in enumerate(specs):
        bev = torch.randn(1, *s.bev_hw, s.c_bev, device=dev)
        img = torch.randn(1, *s.img_hw, s.c_img, device=dev)
        g_bev = torch.randn(1, *s.bev_hw, s.c_bev + s.c_img, device=dev)
        g_img = torch.randn(1, *s.img_hw, s.c_img + s.c_bev, device=dev) if s.dual else None
        r = timeit(lambda: pipe.build_layer(i, pts, vox, synth.P2_KITTI, n, st()), args.iters)
        out["pieces"][s.name + ".build"] = r
        torch.cuda.synchronize()
        nnz = int(pipe.layers[i].plan.counts[0, 3].item())
        out["pieces"][s.name + ".nnz"] = nnz
        for label, fl in (("warmL2", None), ("coldL2", flush)):
            r = timeit(lambda: pipe.forward_layer(i, bev, img, st(), n), args.iters, fl)
            r["GBs"] = s.bytes_forward(nnz) / r["us_median"] / 1e3
            r["frac_of_peak"] = r["GBs"] / peak
            out["pieces"][s.name + ".forward." + label] = r
            r = timeit(lambda: pipe.backward_layer(i, g_bev, g_img, st(), n), args.iters, fl)
            r["GBs"] = s.bytes_backward(nnz) / r["us_median"] / 1e3
            r["frac_of_peak"] = r["GBs"] / peak
            out["pieces"][s.name + ".backward." + label] = r
    # a plain device copy of the same size as layer B's forward traffic, for reference
    a = torch.empty(110 << 20, dtype=torch.uint8, device=dev)
    b = torch.empty(110 << 20, dtype=torch.uint8, device=dev)
    r = timeit(lambda: b.copy_(a), args.iters, flush)
    r["GBs"] = 2 * a.numel() / r["us_median"] / 1e3
    out["pieces"]["torch_copy_220MB.coldL2"] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
