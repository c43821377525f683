#!/usr/bin/env python
"""Stress sweep (BASELINE.json configs 2', 3, 4, 5): forward / backward device time of the pooling
kernels over nnz, channel count and row-length skew, as CUDA-graph replays timed with CUDA events.
    python tools/sweep.py [--quick]
Prints one JSON object (a table); bench.py remains the contract benchmark."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sparse_pooling_b200 as shpl  # noqa: E402
from sparse_pooling_b200 import ops  # noqa: E402
from tools import synth  # noqa: E402


def timeit(fn, iters=10, reps=2):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return float(np.median(ts))


def case(name, bev_hw, img_hw, C, n_pairs, skew, stride=(1, 1), weights=False, peak=6544.0, seed=11):
    dev = torch.device("cuda", 0)
    d = synth.direct_pairs(seed, n_pairs, bev_hw=(bev_hw[0] * stride[1], bev_hw[1] * stride[1]),
                           img_wh=(img_hw[1] * stride[0], img_hw[0] * stride[0]), skew=skew)
    dd = dict(bv_index=torch.from_numpy(d["bv_index"]).to(dev), img_index=torch.from_numpy(d["img_index"]).to(dev),
              bv_size=d["bv_size"], img_size=d["img_size"])
    mval = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    o = shpl.produce_sparse_pooling_input(dd, M_val=mval, stride=list(stride))
    e1.record()
    torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    plan = o["shpl_plan"]
    nnz = plan.nnz[0]
    if weights:      # perf only: non-homogeneous weights written straight into the plan
        w = (1.0 / torch.randint(1, 46, (plan.capacity,), device=dev)).float()
        plan.csr_val.copy_(w)
        plan.csrT_val.copy_(w)
    R, Q = bev_hw[0] * bev_hw[1], img_hw[0] * img_hw[1]
    rp = plan.row_ptr.cpu().numpy()
    max_row = int(np.diff(rp).max()) if nnz else 0
    bev = torch.randn(R, C, device=dev)
    img = torch.randn(Q, C, device=dev)
    g = torch.randn(R, 2 * C, device=dev)
    fused = torch.empty(R, 2 * C, device=dev)
    g_dst = torch.empty(R, C, device=dev)
    g_src = torch.empty(Q, C, device=dev)
    lib = shpl._cabi.lib
    P = ops._ptr

    def fwd():
        ops.pool_forward(bev, img, plan.by_row(), R, Q)

    def bwd():
        ops.pool_backward(g, plan.by_pixel(), R, C, Q, C)

    tf, tb = timeit(fwd), timeit(bwd)
    bf = 4 * (R * C + R * 2 * C + nnz * (C + 2) + R + 1)
    bb = 4 * (2 * R * C + nnz * (C + 2) + Q * C + Q + 1)
    return dict(case=name, seed=seed, R=R, Q=Q, C=C, nnz=nnz, max_row=max_row, skew=skew, build_api_ms=round(build_ms, 3),
                fwd_us=round(tf, 1), fwd_GBs=round(bf / tf / 1e3), fwd_frac=round(bf / tf / 1e3 / peak, 3),
                bwd_us=round(tb, 1), bwd_GBs=round(bb / tb / 1e3), bwd_frac=round(bb / tb / 1e3 / peak, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--full", action="store_true", help="the whole BASELINE config-5 grid: nnz {5k,20k,100k,1M} x C {16,32,64,128,256} x "
                                                        "{uniform, zipf, ground} x seeds 0-4, plus a per-cell summary")
    ap.add_argument("--only", default=None, help="with --full: restrict to one row distribution (uniform | zipf | ground)")
    args = ap.parse_args()
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    K = ((700, 800), (360, 1200))
    rows = []
    rows.append(case("cfg1 kitti s1 C32", *K, 32, 20000, "uniform", peak=peak))
    rows.append(case("cfg2' retinanet P2 s4 C256", (175, 200), (90, 300), 256, 20000, "uniform", stride=(4, 4), peak=peak))
    rows.append(case("cfg2A vgg conv4 s8 C256", (88, 100), (45, 150), 256, 20000, "uniform", stride=(8, 8), peak=peak))
    rows.append(case("cfg3 mv3d ped C768 (8,2)", (100, 120), (48, 160), 768, 20000, "ground", stride=(8, 2), weights=True, peak=peak))
    rows.append(case("cfg4 full scan C128", *K, 128, 120000, "ground", weights=True, peak=peak))
    if args.full:
        grid = []
        for nnz in (5000, 20000, 100000, 1000000):
            for C in (16, 32, 64, 128, 256):
                for skew in ("uniform", "zipf", "ground"):
                    if args.only and skew != args.only:
                        continue
                    per_seed = []
                    for seed in range(5):
                        r = case("cfg5 nnz%d C%d %s" % (nnz, C, skew), *K, C, nnz, skew, weights=True, peak=peak, seed=seed)
                        per_seed.append(r)
                        rows.append(r)
                        torch.cuda.empty_cache()
                    grid.append(dict(nnz=nnz, C=C, skew=skew, seeds=5, max_row=max(r["max_row"] for r in per_seed),
                                     fwd_us_median=float(np.median([r["fwd_us"] for r in per_seed])),
                                     fwd_frac_median=float(np.median([r["fwd_frac"] for r in per_seed])),
                                     fwd_frac_min=min(r["fwd_frac"] for r in per_seed),
                                     bwd_us_median=float(np.median([r["bwd_us"] for r in per_seed])),
                                     bwd_frac_median=float(np.median([r["bwd_frac"] for r in per_seed])),
                                     bwd_frac_min=min(r["bwd_frac"] for r in per_seed)))
        print(json.dumps({"peak_GBs": peak, "grid": grid, "rows": rows}, indent=1))
        return
    if not args.quick:
        for nnz in (5000, 100000, 1000000):
            for C in (16, 64, 256):
                for skew in ("uniform", "zipf", "ground"):
                    if C == 256 and nnz == 1000000 and skew == "zipf":
                        pass
                    rows.append(case("cfg5 nnz%d C%d %s" % (nnz, C, skew), *K, C, nnz, skew, weights=True, peak=peak))
                    torch.cuda.empty_cache()
    print(json.dumps({"peak_GBs": peak, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
