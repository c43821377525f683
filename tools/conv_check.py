"""Bring-up check of the fused post-fusion 3x3 conv (shpl_pool_conv3x3_forward) on a GPU:
errors against torch conv2d (float64) at a few sizes, then timings at the KITTI size."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_pooling_b200 as shpl  # noqa: E402
from sparse_pooling_b200 import conv_fusion  # noqa: E402
from tools import synth  # noqa: E402


def ref_conv(fused, w, scale, shift, relu):
    x = fused.double().permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(x, w.double().permute(3, 2, 0, 1), padding=1)
    mag = torch.nn.functional.conv2d(x.abs(), w.double().abs().permute(3, 2, 0, 1), padding=1)
    y, mag = y.permute(0, 2, 3, 1), mag.permute(0, 2, 3, 1)
    if scale is not None:
        y, mag = y * scale.double(), mag * scale.double().abs()
    if shift is not None:
        y, mag = y + shift.double(), mag + shift.double().abs()
    if relu:
        y = y.clamp_min(0)
    return y, mag


def check(B, H, W, Hi, Wi, n_pairs, relu, affine, seed, pooled=True):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    bev = torch.randn(B, H, W, 32, device=dev, generator=g)
    img = torch.randn(B, Hi, Wi, 32, device=dev, generator=g)
    w = torch.randn(3, 3, 64 if pooled else 32, 32, device=dev, generator=g) * 0.1
    scale = torch.rand(32, device=dev, generator=g) + 0.5 if affine else None
    shift = torch.randn(32, device=dev, generator=g) if affine else None
    if pooled:
        rng = np.random.default_rng(seed)
        n = n_pairs
        rows = rng.integers(0, H * W, n)
        flip = np.stack([np.zeros(n, np.int64), rng.integers(0, Hi, n), rng.integers(0, Wi, n)], 1)
        Mij = np.stack([rows, np.arange(n)], 1).astype(np.int64)
        val = rng.random(n).astype(np.float32) + 0.5
        M = shpl.SparseTensor(torch.from_numpy(Mij).to(dev), torch.from_numpy(val).to(dev), np.array([H * W, n]))
        flip_t = torch.from_numpy(flip).to(dev)
        assert B == 1
        fused, _ = shpl.sparse_pool_layer([bev, img], [32, 32], M, img_index_flip=flip_t)
        out = conv_fusion.sparse_pool_conv3x3([bev, img], M, flip_t, w, scale, shift, relu)
    else:
        fused = bev
        out = conv_fusion.sparse_pool_conv3x3([bev, None], None, None, w, scale, shift, relu)
    torch.cuda.synchronize()
    y, mag = ref_conv(fused, w, scale, shift, relu)
    err = (out.double() - y).abs()
    rel = (err / mag.clamp_min(1e-30)).max().item()
    print("B=%d %dx%d pooled=%s pairs=%d relu=%d affine=%d: max|err|=%.3e  max err/sum|terms|=%.3e  %s"
          % (B, H, W, pooled, n_pairs, relu, affine, err.max().item(), rel, "OK" if rel < 1e-5 else "FAIL"), flush=True)
    return rel < 1e-5


def timing():
    dev = torch.device("cuda", 0)
    H, W, Hi, Wi = 700, 800, 360, 1200
    sets = []
    rng = np.random.default_rng(0)
    for i in range(3):
        bev = torch.randn(1, H, W, 32, device=dev)
        img = torch.randn(1, Hi, Wi, 32, device=dev)
        d = synth.direct_pairs(i, 20000, (H, W), (Wi, Hi))
        sets.append((bev, img, d))
    w = torch.randn(3, 3, 64, 32, device=dev) * 0.1
    outs = [torch.empty(1, H, W, 32, device=dev) for _ in range(3)]
    ws = conv_fusion.conv_workspace(1, H, W, dev, 40000)
    Ms = []
    scan = "--scan" in sys.argv      # the bench's pairs: a synthetic 64-beam scan (busy cells clustered along the rings)
    for i, (bev, img, d) in enumerate(sets):
        if scan:
            f = synth.avod_frame(100 + i, az_step_deg=0.028)

            class Calib:
                p2 = f["P"]
            g = shpl.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], Calib, f["im_size"], [H, W])
            o = shpl.produce_sparse_pooling_input(g, stride=[1, 1])
        else:
            o = shpl.produce_sparse_pooling_input(dict(d), stride=[1, 1])
        Ms.append(o)
    print("pairs per frame:", [int(o["M_size"][1]) for o in Ms], "(scan pattern)" if scan else "(uniform)")
    def run(k, pooled=True):
        bev, img, d = sets[k % 3]
        o = Ms[k % 3]
        if pooled:
            conv_fusion.sparse_pool_conv3x3([bev, img], o, o["img_index_flip_pool"], w, out=outs[k % 3], workspace=ws)
        else:
            conv_fusion.sparse_pool_conv3x3([bev, None], None, None, w[:, :, :32].contiguous(), out=outs[k % 3], workspace=ws)
    for pooled in (False, True):
        for k in range(5):
            run(k, pooled)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for k in range(3):
                run(k, pooled)
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        torch.cuda.synchronize()
        e0.record()
        for k in range(n):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (3 * n)
        flops = 2.0 * H * W * 32 * 288
        print("KITTI 700x800 32(+32)->32 pooled=%s: %.1f us per call (dense half %.1f GFLOP -> %.1f TFLOP/s fp32-equivalent; "
              "in+out 143.4 MB -> %.0f GB/s)" % (pooled, us, flops / 1e9, flops / us / 1e6, 143.4e6 / us / 1e3), flush=True)


if __name__ == "__main__":
    ok = True
    ok &= check(1, 16, 8, 10, 10, 0, False, False, 0, pooled=False)
    ok &= check(1, 32, 16, 10, 10, 0, False, False, 1, pooled=False)
    ok &= check(2, 50, 44, 10, 10, 0, True, True, 2, pooled=False)
    ok &= check(1, 50, 44, 20, 30, 300, False, False, 3)
    ok &= check(1, 50, 44, 20, 30, 3000, True, True, 4)
    ok &= check(1, 700, 800, 360, 1200, 20000, True, True, 5)
    if "--time" in sys.argv:
        timing()
    print("conv_check:", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
