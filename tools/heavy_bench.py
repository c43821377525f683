"""Times shpl_pool_heavy alone (exact kernel + tree kernel) for ONE listed cell of L entries, per channel count.
CUDA events around 20 back-to-back calls.  Usage: python tools/heavy_bench.py [C ...]   (default: 16 64 256)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_pooling_b200 as shpl  # noqa: E402
from sparse_pooling_b200 import _cabi, ops  # noqa: E402
from sparse_pooling_b200.ops import _ptr, _stream  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
out = []
for C in ([int(x) for x in sys.argv[1:]] or [16, 64, 256]):
    for L in (600, 2048, 8192, 16384, 40000):
        n = L + 500
        bx = np.r_[np.full(L, 3), rng.integers(0, 110, 500)]
        bz = np.r_[np.full(L, 2), rng.integers(0, 120, 500)]
        u, v = rng.integers(0, 1200, n), rng.integers(0, 360, n)
        d = dict(bv_index=np.stack((bx, bz), axis=1).astype(np.int64), img_index=np.stack((u, v, np.zeros(n))).astype(np.float64),
                 bv_size=np.array([120, 110]), img_size=np.array([1200, 360]))
        o = shpl.produce_sparse_pooling_input(d)
        plan = o["shpl_plan"]
        img = torch.randn((360 * 1200, C), device=dev)
        fused = torch.zeros((120 * 110, 2 * C), device=dev)
        P8 = plan.ptrs8()
        lst, cnt, cap, expected = plan.heavy(False)

        def call():
            rc = _cabi.lib.shpl_pool_heavy(_ptr(img), C, C, P8[0], P8[2], P8[3], lst, cnt, int(cap), None, 0,
                                           ops._off(fused, C), 2 * C, _stream())
            assert rc == 0
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        out.append(dict(C=C, L=L, listed=expected, us=round(e0.elapsed_time(e1) * 50.0, 1)))
        print(out[-1], flush=True)
print(json.dumps(out))
