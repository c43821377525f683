#!/usr/bin/env python
"""Calibration: what plain torch kernels reach for the read/write mixes the SHPL kernels have
(timed as CUDA-graph replays with CUDA events, L2 left dirty by a 256 MB rewrite before each replay,
and back to back).  Context for roofline fractions only."""
import json
import numpy as np
import torch

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, cold=True, reps=1):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return float(np.median(ts))


MB = 1 << 20
out = {}
a72 = torch.empty(72 * MB // 4, dtype=torch.float32, device=dev)
b72 = torch.empty(72 * MB // 4, dtype=torch.float32, device=dev)
w143 = torch.empty(143 * MB // 4, dtype=torch.float32, device=dev)
a110 = torch.empty(110 * MB // 4, dtype=torch.float32, device=dev)
b110 = torch.empty(110 * MB // 4, dtype=torch.float32, device=dev)
big_a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
big_b = torch.empty(1 << 28, dtype=torch.float32, device=dev)
for cold in (True, False):
    tag = "dirtyL2" if cold else "back2back"
    reps = 1 if cold else 4
    t = timeit(lambda: w143.zero_(), cold=cold, reps=reps); out["fill_143MB." + tag] = {"us": t, "GBs": 143 * MB / t / 1e3}
    t = timeit(lambda: b72.copy_(a72), cold=cold, reps=reps); out["copy_72MB_to_72MB." + tag] = {"us": t, "GBs": 144 * MB / t / 1e3}
    t = timeit(lambda: b110.copy_(a110), cold=cold, reps=reps); out["copy_110MB." + tag] = {"us": t, "GBs": 220 * MB / t / 1e3}
    t = timeit(lambda: a72.sum(), cold=cold, reps=reps); out["read_72MB." + tag] = {"us": t, "GBs": 72 * MB / t / 1e3}
t = timeit(lambda: big_b.copy_(big_a), cold=False, reps=1); out["copy_1GiB_to_1GiB"] = {"us": t, "GBs": 2 * (1 << 30) / t / 1e3}
t = timeit(lambda: big_b.zero_(), cold=False, reps=1); out["fill_1GiB"] = {"us": t, "GBs": (1 << 30) / t / 1e3}
print(json.dumps(out, indent=1))
