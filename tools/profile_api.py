#!/usr/bin/env python
"""cProfile of the public drop-in API on the bench workload (where does the host time go?)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_pooling_b200 as shpl  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda", 0)
f = synth.avod_frame(100, az_step_deg=0.028)
pts = torch.from_numpy(f["points"]).pin_memory()
vox = torch.from_numpy(np.ascontiguousarray(f["voxel_indices"][:, :2])).pin_memory()


class Calib:
    p2 = synth.P2_KITTI


layers = [((88, 100), (45, 150), 256, (8, 8), True, (704, 800)), ((700, 800), (360, 1200), 32, (1, 1), False, (700, 800))]
maps = []
for bev_hw, img_hw, C, stride, dual, bv in layers:
    maps.append((torch.randn(1, *bev_hw, C, device=dev, requires_grad=True), torch.randn(1, *img_hw, C, device=dev, requires_grad=True),
                 torch.randn(1, *bev_hw, 2 * C, device=dev), torch.randn(1, *img_hw, 2 * C, device=dev)))


def step():
    for (bev_hw, img_hw, C, stride, dual, bv), (bev, img, gb, gi) in zip(layers, maps):
        d = shpl.gen_sparse_pooling_input_avod(pts, vox, Calib, [1200, 360], bv)
        o = shpl.produce_sparse_pooling_input(d, stride=list(stride))
        M = shpl.SparseTensor.from_sparse_pooling_input(o)
        bev.grad = None
        img.grad = None
        a, b = shpl.sparse_pool_layer([bev, img], [C, C], M, img_index_flip=o["img_index_flip_pool"],
                                      bv_index=(np.zeros((1, 3)) if dual else None))
        if dual:
            torch.autograd.backward([a, b], [gb, gi])
        else:
            torch.autograd.backward([a], [gb])
    return float(maps[0][0].grad.reshape(-1)[0])


for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
print("per step ms", (time.perf_counter() - t0) / 50 * 1e3)
# host time per phase (perf_counter around each call; the syncs inside produce / the final read are part of their phase)
ph = {}


def timed(name, fn, *a, **k):
    t = time.perf_counter()
    r = fn(*a, **k)
    ph[name] = ph.get(name, 0.0) + time.perf_counter() - t
    return r


def step_phases():
    roots, grads = [], []
    for (bev_hw, img_hw, C, stride, dual, bv), (bev, img, gb, gi) in zip(layers, maps):
        d = timed("gen", shpl.gen_sparse_pooling_input_avod, pts, vox, Calib, [1200, 360], bv)
        o = timed("produce", shpl.produce_sparse_pooling_input, d, stride=list(stride))
        M = timed("SparseTensor", shpl.SparseTensor.from_sparse_pooling_input, o)
        bev.grad = None
        img.grad = None
        a, b = timed("layer", shpl.sparse_pool_layer, [bev, img], [C, C], M, img_index_flip=o["img_index_flip_pool"],
                     bv_index=(np.zeros((1, 3)) if dual else None))
        roots += [a, b] if dual else [a]
        grads += [gb, gi] if dual else [gb]
    timed("backward", torch.autograd.backward, roots, grads)
    return timed("read", lambda: float(maps[0][0].grad.reshape(-1)[0]))


for _ in range(5):
    step_phases()
ph.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step_phases()
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / 200 * 1e6
print("phases, us per step (2 layers, one backward): total %.0f  " % tot + "  ".join("%s %.0f" % (k, v / 200 * 1e6) for k, v in ph.items()))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
