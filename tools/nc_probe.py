"""A/B of the no-concat kernels (shpl_pool_forward_into / shpl_pool_backward_from) at the KITTI pre-RPN shape against the
grid-size knob of the experiment build:  SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_STREAM_CTAS_PER_SM=12 python tools/nc_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_pooling_b200 as shpl  # noqa: E402
from sparse_pooling_b200 import ops  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda", 0)
R, Q, C = 560000, 432000, 32
d = synth.direct_pairs(0, 20000, (700, 800), (1200, 360))
o = shpl.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()})
plan = o["shpl_plan"]
sets = [dict(img=torch.randn(Q, C, device=dev), fused=torch.empty(R, 2 * C, device=dev), g=torch.randn(R, 2 * C, device=dev)) for _ in range(3)]


def timeit(fn, n=30):
    for k in range(3):
        fn(k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(3):
            fn(k)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * n)


nnz = plan.nnz[0]
tf = timeit(lambda k: ops.pool_forward_into(sets[k % 3]["fused"], sets[k % 3]["img"], plan.by_row(), R, Q, C))
tb = timeit(lambda k: ops.pool_backward_from(sets[k % 3]["g"], plan.by_pixel(), R, Q, C, C))
bf = 4 * (R * C + nnz * (C + 2) + R + 1)
bb = 4 * (nnz * (C + 2) + Q * C + Q + 1)
print("knob SHPL_STREAM_CTAS_PER_SM=%s: forward_into %.1f us (%.0f GB/s)  backward_from %.1f us (%.0f GB/s)"
      % (os.environ.get("SHPL_STREAM_CTAS_PER_SM", "default"), tf, bf / tf / 1e3, tb, bb / tb / 1e3))
