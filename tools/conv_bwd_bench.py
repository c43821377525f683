#!/usr/bin/env python
"""Backward of the fused post-fusion conv (shpl_pool_conv3x3_backward) at the KITTI size: time per call (CUDA-graph replays,
CUDA events, three rotating input sets) -- run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_pooling_b200 as shpl  # noqa: E402
from sparse_pooling_b200 import conv_fusion  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda", 0)
H, W, Hi, Wi = 700, 800, 360, 1200
sets = []
for i in range(3):
    f = synth.avod_frame(100 + i, az_step_deg=0.028)

    class Calib:
        p2 = f["P"]
    g = shpl.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], Calib, f["im_size"], [H, W])
    o = shpl.produce_sparse_pooling_input(g, stride=[1, 1])
    sets.append((torch.randn(1, H, W, 32, device=dev), torch.randn(1, Hi, Wi, 32, device=dev), torch.randn(1, H, W, 32, device=dev), o["shpl_plan"]))
w = torch.randn(3, 3, 64, 32, device=dev) * 0.1
outs = (torch.empty(1, H, W, 32, device=dev), torch.empty(1, Hi, Wi, 32, device=dev), torch.empty_like(w))
need = int(shpl._cabi.lib.shpl_conv3x3_backward_workspace_bytes(40000))
ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)


def run(k):
    bev, img, g_out, plan = sets[k % 3]
    conv_fusion.sparse_pool_conv3x3_backward(g_out, [bev, img], plan, w, out=outs, workspace=ws)


for k in range(4):
    run(k)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for k in range(3):
        run(k)
for _ in range(2):
    gr.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    gr.replay()
e1.record()
torch.cuda.synchronize()
print("conv3x3 backward at 700x800, 32(+32)->32, %s pairs: %.1f us per call (g_bev + g_img + g_weight)"
      % ([int(s[3].nnz[0]) for s in sets], e0.elapsed_time(e1) * 1e3 / (3 * n)))
