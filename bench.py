#!/usr/bin/env python
"""bench.py -- SHPL forward+backward (+ correspondence build) frames/sec at KITTI shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 2|1|2p|3|4]

--config selects the BASELINE.json configuration (default 2 = configs[1], the one the metric is quoted on):
  1   avod SHPL layer alone: BEV 700x800x32 <- image 360x1200x32, fwd+bwd+build, one frame per step
  2   avod-FPN pyramid people, 2 NHSP layers (A: stride-8 DUAL 88x100x256 <-> 45x150x256; B: config 1's), batch 1
  2p  RetinaNet P2 (stride 4): 175x200x256 <- 90x300x256
  3   MV3D VoxelNet+MSCNN middle-stage fusion (image stride 8 / BEV stride 2, C = 768, 1/count weights), batch 8 per GPU
  4   full 64-beam scans (120 k pairs, C = 128, KITTI stride 1), a FIXED batch of 32 frames sharded across the GPUs
      (strong scaling: sharding.frames_for_rank)
One STEP = one batch of synthetic frames: correspondence build + forward + backward of every layer.
`value`   : frames/s with every input already resident in HBM (CUDA-graph replay of the step)
`e2e`     : frames/s from HOST buffers: the frame's inputs in pinned host memory (H2D inside the timed region), a D2H
            read of every step's result
`roofline`: the dominant kernel (the largest layer's forward) timed with CUDA events, algorithmic bytes
`cpu_baseline` / --impl reference: the CPU oracle (port of the reference's algorithm; TensorFlow is not installable) on
            ALL host cores of the box (the thread count is set explicitly: torchrun exports OMP_NUM_THREADS=1)
N > 1: frames are independent -> each rank runs its own frames, no collective on the hot path; NCCL is used only to take
       the max time over ranks.
Before anything is timed, one step of the timed path is compared with the CPU oracle, bit for bit; a mismatch aborts.
"""
import argparse
import ctypes
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SHPL fwd+bwd+build frames/sec at KITTI shape (avod-FPN 2 NHSP layers, batch 1)"
UNIT = "frames/s"
AZ_STEP = 0.028         # azimuth step of the synthetic 64-beam scan: ~20k correspondence pairs per frame
N_FRAMES = 4             # distinct synthetic frames rotated through
N_MAX = 32768
MIN_TIMED_S = 0.5        # the K-step block is repeated until the timed region lasts at least this long


def _layer_spec_module():
    """sparse_pooling_b200/layer_spec.py loaded BY PATH: pure Python, so the reference arm gets the shapes and byte
    formulas without importing the package (whose __init__ maps libshpl.so)."""
    spec = importlib.util.spec_from_file_location("shpl_layer_spec", os.path.join(ROOT, "sparse_pooling_b200", "layer_spec.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault("shpl_layer_spec", mod)
    spec.loader.exec_module(mod)
    return mod


def configs():
    LS = _layer_spec_module().LayerSpec
    A = LS("A_vgg_conv4_s8_dual", (88, 100), (45, 150), 256, 256, (8, 8), True, (1200, 360), (704, 800))
    B = LS("B_pre_rpn_s1", (700, 800), (360, 1200), 32, 32, (1, 1), False, (1200, 360), (700, 800))
    P2 = LS("P2_retinanet_s4", (175, 200), (90, 300), 256, 256, (4, 4), False, (1200, 360), (700, 800))
    MV = LS("MV3D_ped_img8_bev2", (100, 120), (48, 160), 768, 768, (8, 2), False, (1280, 384), (200, 240))
    FS = LS("full_scan_c128_s1", (700, 800), (360, 1200), 128, 128, (1, 1), False, (1200, 360), (700, 800))
    return {
        "2": dict(kind="avod", layers=[A, B], metric=METRIC, scaling="weak", batch=1,
                  workload="avod-FPN pyramid people, 2 NHSP layers (A: stride-8 dual 88x100x256<->45x150x256; B: stride-1 "
                           "700x800x32<-360x1200x32), fwd+bwd+build, batch 1"),
        "1": dict(kind="avod", layers=[B], metric="SHPL fwd+bwd+build frames/sec at KITTI shape (avod SHPL layer, one frame)",
                  scaling="weak", batch=1, workload="avod SHPL pre-RPN layer, stride-1 700x800x32<-360x1200x32, fwd+bwd+build, batch 1"),
        "2p": dict(kind="avod", layers=[P2], metric="SHPL fwd+bwd+build frames/sec (RetinaNet P2, stride 4)", scaling="weak", batch=1,
                   workload="RetinaNet P2 SHPL, stride-4 175x200x256<-90x300x256, fwd+bwd+build, batch 1"),
        "3": dict(kind="pairs", layers=[MV], metric="SHPL fwd+bwd+build frames/sec (MV3D middle-stage fusion, batch 8)", scaling="weak",
                  batch=8, pairs=20000, gen="mv3d",
                  workload="MV3D VoxelNet+MSCNN middle-stage fusion: image stride 8 / BEV stride 2, 100x120x768<-48x160x768, "
                           "~20k points per frame with 1/count weights, fwd+bwd+build, batch 8 per GPU in one launch"),
        "4": dict(kind="pairs", layers=[FS], metric="SHPL fwd+bwd+build frames/sec (full 64-beam scans, C=128, batch 32 sharded)",
                  scaling="strong", batch=32, pairs=120000, gen="ground",
                  workload="full 64-beam scans: 120k pairs per frame (ground-plane-skewed cells, 1/row-count weights), "
                           "700x800x128<-360x1200x128, fwd+bwd+build, a fixed batch of 32 frames sharded by frame across the GPUs"),
    }


def config_dict(name, cfg, world):
    """The `config` object of the JSON line: what the workload IS.  Both arms print exactly this."""
    per_gpu = cfg["batch"] if cfg["scaling"] == "weak" else None
    return {"workload": cfg["workload"], "baseline_config": name, "frames_per_step": cfg["batch"] * (world if cfg["scaling"] == "weak" else 1),
            "frames_per_step_per_gpu": per_gpu if per_gpu is not None else "%d / n_gpus" % cfg["batch"],
            "sharding": "frames by rank, no data-path collective",
            "l2": "inputs larger than L2 (126 MB): %d rotating input/buffer sets, every step touches more than L2 holds" % 2}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get("shpl_forward_kernel_layerB_dram_bytes")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)      # first NVML query is slow: pay for it here
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if bit and (r & bit):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------- synthetic frames (numpy, both arms)
def avod_frame_inputs(seed):
    from tools import synth
    return synth.avod_frame(seed, az_step_deg=AZ_STEP)


def pairs_frame_inputs(cfg, seed):
    """One frame of a `pairs` configuration: dict(img_index f64 [3,n], bv_index i64 [n,2], m_val f64 [n], img_size, bv_size)."""
    from tools import synth
    s = cfg["layers"][0]
    if cfg["gen"] == "mv3d":      # the MV3D feeder's own outputs: pairs + 1/count weights (construct_voxel.py:116-160)
        from oracle import index_oracle as io
        f = synth.mv3d_frame(seed, n_points=cfg["pairs"])
        u, v = f["img_index2"]
        inside = (u >= 0) & (u < s.im_size[0]) & (v >= 0) & (v < s.im_size[1])
        fsh, img2 = f["points_fsh"][inside], f["img_index2"][:, inside]
        inrange, kept, bv_index, m_val = io.mv3d_voxel_weights(fsh, f["res"], f["zres"], f["side_range"], f["fwd_range"],
                                                               f["height_range"], f["max_points"])
        uv = img2[:, inrange][:, kept]
        img_index = np.zeros((3, uv.shape[1]), dtype=np.float64)
        img_index[0:2] = uv
        return dict(img_index=img_index, bv_index=np.ascontiguousarray(bv_index, dtype=np.int64),
                    m_val=np.ascontiguousarray(m_val, dtype=np.float64), img_size=np.array(s.im_size), bv_size=np.array(s.bv_size))
    d = synth.direct_pairs(seed, cfg["pairs"], tuple(s.bv_size), tuple(s.im_size), skew=cfg["gen"])
    # non-homogeneous weights 1 / (pairs in the cell) -- SURVEY.md 8(d) config 4; every row is < R here, so M_val needs no filtering
    cell = d["bv_index"][:, 1] * s.bv_size[1] + d["bv_index"][:, 0]
    _, inv, cnt = np.unique(cell, return_inverse=True, return_counts=True)
    d["m_val"] = 1.0 / cnt[inv]
    return d


def set_host_threads():
    """The oracle's OpenMP loops use every core this process may run on, whatever OMP_NUM_THREADS the launcher exported
    (torch.distributed.run sets it to 1).  Must run before the oracle library is loaded."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


# ------------------------------------------------------------------------------- CPU legs
class CpuWorkload:
    """The same step on the host: numpy index oracle + plain-C value oracle (port of the
    reference's algorithm; TensorFlow is not installable, so kind = "port")."""

    def __init__(self, cfg):
        self.cores_wanted = set_host_threads()
        from oracle import cref
        cref.build()
        self.cfg = cfg
        self.specs = cfg["layers"]
        rng = np.random.default_rng(0)
        self.maps = []
        for s in self.specs:
            bev = rng.standard_normal(s.bev_hw + (s.c_bev,), dtype=np.float32)
            img = rng.standard_normal(s.img_hw + (s.c_img,), dtype=np.float32)
            g_bev = rng.standard_normal(s.bev_hw + (s.c_bev + s.c_img,), dtype=np.float32)
            g_img = rng.standard_normal(s.img_hw + (s.c_img + s.c_bev,), dtype=np.float32) if s.dual else None
            self.maps.append((bev, img, g_bev, g_img))
        if cfg["kind"] == "avod":
            self.frames = [avod_frame_inputs(100 + i) for i in range(N_FRAMES)]
        else:
            self.frames = [pairs_frame_inputs(cfg, 100 + i) for i in range(2)]
        self.threads = cref.threads()

    def frame(self, k):
        """build + forward + backward of every layer for one frame"""
        from oracle import cref, index_oracle as io
        f = self.frames[k % len(self.frames)]
        out = []
        for s, (bev, img, g_bev, g_img) in zip(self.specs, self.maps):
            if self.cfg["kind"] == "avod":
                d = io.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], f["P"], list(s.im_size), s.bv_size)
                o = io.produce_sparse_pooling_input(d, stride=list(s.stride))
                val = np.ones(len(o["Mij_pool"]), np.float32)
            else:
                d = dict(img_index=f["img_index"].copy(), bv_index=f["bv_index"], img_size=f["img_size"], bv_size=f["bv_size"])
                o = io.produce_sparse_pooling_input(d, M_val=f["m_val"], stride=list(s.stride))
                val = np.asarray(o["M_val"], dtype=np.float32)
            Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
            fused = cref.forward(bev, img, Mij, val, flip)
            gd, gs = cref.backward(g_bev, Mij, val, flip, s.c_bev, img.shape)
            if s.dual:
                fused_i = cref.forward_trans(img, bev, Mij, val, flip)
                gi, gb = cref.backward_trans(g_img, Mij, val, flip, s.c_img, bev.shape)
                gd += gb
                gs += gi
                out.append(float(fused_i[0, 0, 0]))
            out.append(float(fused[0, 0, 0]) + float(gd[0, 0, 0]) + float(gs[0, 0, 0]))
        return out


def run_reference(args, name, cfg):
    """--impl reference: the reference's own CPU path for this workload.  The reference is numpy + TensorFlow 1.8 (not
    installable here), so its algorithm is timed through the CPU oracle (numpy builder + plain-C restatement of the TF
    ops), on all host threads.  A step is a BOUNDED SAMPLE of the workload's batch (whole frames), so that the run ends
    within minutes; frames/s does not depend on the batch size."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    wl = CpuWorkload(cfg)
    sample = 1 if cfg["batch"] == 1 else min(cfg["batch"], 2)
    for k in range(min(args.warmup, 2)):
        wl.frame(k)
    t0 = time.perf_counter()
    n = 0
    for k in range(args.steps):
        for j in range(sample):
            wl.frame(n)
            n += 1
    dt = time.perf_counter() - t0
    fps = n / dt
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, max(world, args.gpus)),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": wl.threads, "cores_available": wl.cores_wanted, "kind": "port",
                         "sample": "%d frames (%d per step; numpy correspondence builder + plain-C gather/SpMM/concat "
                                   "and gradients, OpenMP on the dense loops, %d threads)" % (n, sample, wl.threads)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if wl.threads < wl.cores_wanted:
        line["rejected"] = "the oracle runs on %d of %d available host threads" % (wl.threads, wl.cores_wanted)
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(cfg):
    wl = CpuWorkload(cfg)
    wl.frame(0)
    n_cpu, t0 = 0, time.perf_counter()
    while n_cpu < 2 or (time.perf_counter() - t0 < 10.0 and n_cpu < 40):
        wl.frame(n_cpu)
        n_cpu += 1
    dt = time.perf_counter() - t0
    return {"value": n_cpu / dt, "unit": UNIT, "cores": wl.threads, "kind": "port",
            "sample": "%d frames of the same workload (numpy correspondence builder + plain-C oracle of the TF ops, %d OpenMP threads)"
                      % (n_cpu, wl.threads)}


# ------------------------------------------------------------------------------- shared GPU helpers
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the SHPL path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_blocks(est_step_s, K):
    """How many times the K-step block is repeated so that the timed region lasts >= MIN_TIMED_S."""
    return max(1, int(math.ceil(MIN_TIMED_S / max(est_step_s * K, 1e-9))))


def assert_equal_bits(got, ref, what):
    got = got.detach().cpu().numpy()
    if got.shape != ref.shape or not np.array_equal(got, ref):
        bad = int(np.sum(got != ref)) if got.shape == ref.shape else -1
        raise SystemExit("bench.py: PARITY CHECK FAILED before timing: %s differs from the CPU oracle (%d elements)" % (what, bad))


# ------------------------------------------------------------------------------- GPU arm, avod-type configurations
def run_gpu_avod(args, name, cfg):
    D = Dist()
    torch, dist = D.torch, D.dist
    world, rank, local_rank, dev = D.world, D.rank, D.local_rank, D.dev

    import sparse_pooling_b200 as shpl
    from sparse_pooling_b200 import _cabi
    from sparse_pooling_b200.pipeline import FramePipeline
    from tools import synth

    lib = _cabi.lib
    specs = cfg["layers"]
    nL = len(specs)
    K, W = args.steps, max(args.warmup, 3)
    barrier = D.barrier

    # ---- synthetic inputs, resident in HBM (rank-dependent seeds: every rank has its own frames)
    frames_host = [synth.avod_frame(100 + rank * N_FRAMES + i, az_step_deg=AZ_STEP) for i in range(N_FRAMES)]
    n_pts = [int(f["points"].shape[0]) for f in frames_host]
    assert max(n_pts) <= N_MAX
    pts_dev = [torch.from_numpy(f["points"]).to(dev) for f in frames_host]
    vox_dev = [torch.from_numpy(np.ascontiguousarray(f["voxel_indices"][:, :2])).to(dev) for f in frames_host]
    P = synth.P2_KITTI
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n_sets = 2

    def randn(*shape):
        return torch.randn(*shape, device=dev, dtype=torch.float32, generator=g)
    maps = []
    for _ in range(n_sets):
        per_layer = []
        for s in specs:
            per_layer.append(dict(
                bev=randn(1, *s.bev_hw, s.c_bev), img=randn(1, *s.img_hw, s.c_img),
                g_bev=randn(1, *s.bev_hw, s.c_bev + s.c_img),
                g_img=randn(1, *s.img_hw, s.c_img + s.c_bev) if s.dual else None))
        maps.append(per_layer)
    pipes = [FramePipeline(specs, N_MAX, dev) for _ in range(n_sets)]

    # Streams: pooling of the LAST layer (the big one) on `side`, the others on the main stream; one builder stream per
    # layer at high priority (short latency-bound kernels scheduled ahead of the bandwidth-bound pooling CTAs: measured
    # 8454 -> 8600 frames/s in round 1)
    side = torch.cuda.Stream(device=dev)
    build_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(nL)]
    dom = nL - 1                                                 # the dominant layer: timed for the roofline

    def pool_stream(li, main):
        return side if (li == nL - 1 and nL > 1) else main

    def lean_step(k, timing_events=None, overlap=True, PP=None):
        """build + fwd + bwd of every layer on preallocated buffers.  overlap=True: layers on their own streams and
        the builders of frame k+1 beside the pooling of frame k; overlap=False keeps one stream so that an event pair
        brackets exactly one kernel.  PP: the pipelines to run (default: the drop-in concat form)."""
        PP = pipes if PP is None else PP
        fi, si = k % N_FRAMES, k % n_sets
        pipe, mp = PP[si], maps[si]
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        if not overlap:
            for li in range(nL):
                pipe.build_layer(li, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)
            for li in range(nL):
                if li != dom:
                    pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], ms, n_pts[fi])
            if timing_events is not None:
                timing_events[0].record(main)
            pipe.forward_layer(dom, mp[dom]["bev"], mp[dom]["img"], ms, n_pts[fi])
            if timing_events is not None:
                timing_events[1].record(main)
            pipe.backward_layer(dom, mp[dom]["g_bev"], mp[dom]["g_img"], ms, n_pts[fi])
            if timing_events is not None:
                timing_events[2].record(main)
            for li in range(nL):
                if li != dom:
                    pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], ms, n_pts[fi])
            return
        # Software-pipelined across frames: while frame k is pooled (plans built during step k-1), the plans of frame
        # k+1 are built into the other buffer set on the builder streams.  The builder is a chain of small
        # latency-bound kernels, so it hides under the bandwidth-bound pooling.  Every step still performs one frame's
        # build + forward + backward.
        nf, ns = (k + 1) % N_FRAMES, (k + 1) % n_sets
        others = [side] + build_streams
        for st_ in others:
            st_.wait_stream(main)
        for li in range(nL):
            with torch.cuda.stream(build_streams[li]):
                PP[ns].build_layer(li, pts_dev[nf], vox_dev[nf], P, n_pts[nf], build_streams[li].cuda_stream)
        for li in range(nL):
            pst = pool_stream(li, main)
            with torch.cuda.stream(pst):
                pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], pst.cuda_stream, n_pts[fi])
                pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], pst.cuda_stream, n_pts[fi])
        for st_ in others:
            main.wait_stream(st_)

    def prologue_build(k, PP=None):
        """plans of frame k (the first frame of a timed region has no previous step to build them)"""
        PP = pipes if PP is None else PP
        fi, si = k % N_FRAMES, k % n_sets
        ms = torch.cuda.current_stream().cuda_stream
        for li in range(nL):
            PP[si].build_layer(li, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)

    # ---- PARITY CHECK before timing: one step of the timed path (FramePipeline -> C ABI), every layer, against the CPU
    #      oracle (oracle/cref: the plain-C restatement), bit for bit; a mismatch aborts the run
    prologue_build(0)
    lean_step(0)
    torch.cuda.synchronize()
    nnz = [int(L.plan.counts[0, 3].item()) for L in pipes[0].layers]
    parity = {"checked": False}
    if not args.no_parity_check:
        from oracle import cref
        cref.build()
        f0 = frames_host[0]
        for li, s in enumerate(specs):
            o = cref.build_avod(f0["points"], f0["voxel_indices"], P, list(s.im_size), s.bv_size, list(s.stride), want_gen=False)
            Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
            val = np.ones(o["nnz"], np.float32)
            assert o["nnz"] == int(pipes[0].layers[li].plan.counts[0, 1].item()), "nnz differs from the oracle's"
            mp = maps[0][li]
            bev, img, gb_, gi_ = (None if t is None else t[0].cpu().numpy() for t in (mp["bev"], mp["img"], mp["g_bev"], mp["g_img"]))
            L = pipes[0].layers[li]
            assert_equal_bits(L.fused_bev[0], cref.forward(bev, img, Mij, val, flip), "%s: fused BEV map" % s.name)
            gd, gs = cref.backward(gb_, Mij, val, flip, s.c_bev, img.shape)
            if s.dual:
                assert_equal_bits(L.fused_img[0], cref.forward_trans(img, bev, Mij, val, flip), "%s: fused image map" % s.name)
                gi, gb = cref.backward_trans(gi_, Mij, val, flip, s.c_img, bev.shape)
                gd, gs = gd + gb, gs + gi
            assert_equal_bits(L.g_bev[0], gd, "%s: gradient of the BEV map" % s.name)
            assert_equal_bits(L.g_img[0], gs, "%s: gradient of the image map" % s.name)
        parity = {"checked": True, "what": "one FramePipeline step (build + forward + backward) of every layer == oracle/cref, bit for bit",
                  "layers": [s.name for s in specs]}

    # ---- capture the step in CUDA graphs (one per frame/buffer-set combination)
    graphs = []
    use_graph = not args.no_graph
    if use_graph:
        try:
            n_variants = N_FRAMES if N_FRAMES % n_sets == 0 else N_FRAMES * n_sets
            for v in range(n_variants):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    lean_step(v)
                graphs.append(gr)
        except Exception as e:  # pragma: no cover
            print("graph capture failed (%r); timing eager launches" % (e,), file=sys.stderr)
            graphs, use_graph = [], False
            torch.cuda.synchronize()

    # ---- G consecutive steps in ONE graph: the same kernels, but the per-step fork/join and the gap between two
    #      graph launches are paid once per G steps, and the dependencies are per layer (the plans of frame k+1 are
    #      built as soon as the pooling that last used their buffer set has finished; the pooling of frame k starts as
    #      soon as its plans are there) instead of a barrier at every step boundary.
    G = max(4, int(args.graph_steps) // 4 * 4)
    multi = None

    def multi_step(k0, n_steps):
        main = torch.cuda.current_stream()
        others = [side] + build_streams
        for st_ in others:
            st_.wait_stream(main)
        pool_done, build_done = {}, {}
        for j in range(n_steps):
            k = k0 + j
            fi, si = k % N_FRAMES, k % n_sets
            nf, ns = (k + 1) % N_FRAMES, (k + 1) % n_sets
            pipe, mp = pipes[si], maps[si]
            for li in range(nL):                                # plans of frame k+1 into buffer set ns
                bst = build_streams[li]
                if (li, k - 1) in pool_done:
                    bst.wait_event(pool_done[(li, k - 1)])      # the pooling of step k-1 read buffer set ns
                with torch.cuda.stream(bst):
                    pipes[ns].build_layer(li, pts_dev[nf], vox_dev[nf], P, n_pts[nf], bst.cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(bst)
                    build_done[(li, k + 1)] = ev
            for li in range(nL):                                # pooling of frame k
                pst = pool_stream(li, main)
                if (li, k) in build_done:
                    pst.wait_event(build_done[(li, k)])
                with torch.cuda.stream(pst):
                    ps = pst.cuda_stream
                    pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], ps, n_pts[fi])
                    pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], ps, n_pts[fi])
                    ev = torch.cuda.Event()
                    ev.record(pst)
                    pool_done[(li, k)] = ev
        for st_ in others:
            main.wait_stream(st_)

    if use_graph and not args.single_step_graphs:
        try:
            assert G % N_FRAMES == 0 and G % n_sets == 0
            multi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(multi):
                multi_step(0, G)
        except Exception as e:  # pragma: no cover
            print("multi-step graph capture failed (%r); one graph per step" % (e,), file=sys.stderr)
            multi = None
            torch.cuda.synchronize()

    def step(k):
        if use_graph:
            graphs[k % len(graphs)].replay()
        else:
            lean_step(k)

    def run_steps(k_begin, k_end):
        """steps k_begin .. k_end-1, G at a time where the multi-step graph lines up"""
        k = k_begin
        while k < k_end:
            if multi is not None and k % G == 0 and k + G <= k_end:
                multi.replay()
                k += G
            else:
                step(k)
                k += 1

    prologue_build(0)
    evw0, evw1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_steps(0, 1)
    evw0.record()
    run_steps(1, W)
    evw1.record()
    torch.cuda.synchronize()
    # after W steps the plans of frame W are built; the timed region continues the same sequence
    est = D.max(evw0.elapsed_time(evw1) * 1e-3 / max(W - 1, 1))
    reps = timed_blocks(est, K)
    KT = K * reps

    # ---- timed region: the K-step block, `reps` times back to back (>= MIN_TIMED_S), device-timed, barrier +
    #      synchronize on both sides
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    run_steps(W, W + KT)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms_total = D.max(ev0.elapsed_time(ev1))
    # launches per step: counted on eager launches (graph replays re-run the captured ones)
    l0 = int(lib.shpl_kernel_launches())
    lean_step(0)
    launches_per_step = int(lib.shpl_kernel_launches()) - l0
    torch.cuda.synchronize()
    gpu_launches = launches_per_step * KT
    value = world * KT / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (the largest layer's forward), CUDA events on its launching stream
    sB = specs[dom]
    fwd_ms, bwd_ms = [], []
    roof_how = "CUDA events recorded inside CUDA-graph replays of the single-stream step (no host launch gaps)"
    KR = max(K, 20)
    try:
        tgraphs, tevs = [], []
        for v in range(N_FRAMES):
            ev3 = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)]
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                lean_step(v, ev3, overlap=False)
            tgraphs.append(gr)
            tevs.append(ev3)
        for k in range(max(W, 3)):
            tgraphs[k % N_FRAMES].replay()
        torch.cuda.synchronize()
        for k in range(KR):
            tgraphs[k % N_FRAMES].replay()
            torch.cuda.synchronize()
            e = tevs[k % N_FRAMES]
            fwd_ms.append(e[0].elapsed_time(e[1]))
            bwd_ms.append(e[1].elapsed_time(e[2]))
    except Exception as ex:  # pragma: no cover
        print("event-in-graph timing unavailable (%r); timing eager launches" % (ex,), file=sys.stderr)
        roof_how = "CUDA events around eager launches of the single-stream step"
        torch.cuda.synchronize()
        fwd_ms, bwd_ms = [], []
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KR)]
        for k in range(KR):
            lean_step(k, evs[k], overlap=False)      # the same step, one stream: the event pair brackets one kernel
        torch.cuda.synchronize()
        for e in evs:
            fwd_ms.append(e[0].elapsed_time(e[1]))
            bwd_ms.append(e[1].elapsed_time(e[2]))
    peak, peak_src = peaks()
    bytes_fwd_B = sB.bytes_forward(nnz[dom])
    bytes_bwd_B = sB.bytes_backward(nnz[dom])
    fwd_avg = float(np.mean(fwd_ms)) * 1e-3
    bwd_avg = float(np.mean(bwd_ms)) * 1e-3
    achieved = bytes_fwd_B / fwd_avg / 1e9
    roofline = {"bound": "hbm", "kernel": "shpl_pool_sparse_kernel as the forward of layer %s" % sB.name,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic() if name in ("1", "2") else None, "bytes_per_launch": bytes_fwd_B, "us_per_launch": fwd_avg * 1e6,
                "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0, "timing": roof_how,
                "backward_kernel": {"achieved": bytes_bwd_B / bwd_avg / 1e9, "frac": bytes_bwd_B / bwd_avg / 1e9 / peak,
                                    "bytes_per_launch": bytes_bwd_B, "us_per_launch": bwd_avg * 1e6}}
    bytes_step = sum(s.bytes_forward(n) + s.bytes_backward(n) for s, n in zip(specs, nnz))
    step_gbs = bytes_step * KT / (ms_total * 1e-3) / 1e9

    # ---- the no-concat ("sparse-only") form of SURVEY.md 8(d): the producer of a destination map writes its channels
    #      straight into the fused buffer, so the forward writes only the pooled channels and the single-direction
    #      backward only the gradient of the gathered map (g_dst is a view).  Same step, same plans, own byte formulas.
    no_concat = None
    if not args.no_no_concat:
        try:
            pipes_nc = [FramePipeline(specs, N_MAX, dev, no_concat=True) for _ in range(n_sets)]
            for si_ in range(n_sets):
                for li in range(nL):          # the "producer": the destination channels are already in the fused buffers
                    pipes_nc[si_].layers[li].fused_bev[..., :specs[li].c_bev].copy_(maps[si_][li]["bev"])
                    if specs[li].dual:
                        pipes_nc[si_].layers[li].fused_img[..., :specs[li].c_img].copy_(maps[si_][li]["img"])
            prologue_build(0, pipes_nc)
            lean_step(0, PP=pipes_nc)
            torch.cuda.synchronize()
            if not args.no_parity_check:       # same values as the concat form, bit for bit (which equals the oracle, above)
                prologue_build(0)
                lean_step(0)
                torch.cuda.synchronize()
                for li in range(nL):
                    a_, b_ = pipes_nc[0].layers[li], pipes[0].layers[li]
                    if not (torch.equal(a_.fused_bev, b_.fused_bev) and torch.equal(a_.g_img, b_.g_img)):
                        raise SystemExit("bench.py: PARITY CHECK FAILED: no-concat form of layer %s differs from the concat form" % specs[li].name)
            ngr = []
            for v in range(N_FRAMES if N_FRAMES % n_sets == 0 else N_FRAMES * n_sets):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    lean_step(v, PP=pipes_nc)
                ngr.append(gr)
            prologue_build(0, pipes_nc)
            for k in range(W):
                ngr[k % len(ngr)].replay()
            evn0, evn1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            KN = max(K, int(math.ceil(0.2 / max(est, 1e-9))))
            torch.cuda.synchronize()
            barrier()
            evn0.record()
            for k in range(W, W + KN):
                ngr[k % len(ngr)].replay()
            evn1.record()
            torch.cuda.synchronize()
            ms_nc = D.max(evn0.elapsed_time(evn1))
            nf_ms, nb_ms = [], []
            tg2, te2 = [], []
            for v in range(N_FRAMES):
                ev3 = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)]
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    lean_step(v, ev3, overlap=False, PP=pipes_nc)
                tg2.append(gr)
                te2.append(ev3)
            for k in range(3):
                tg2[k % N_FRAMES].replay()
            torch.cuda.synchronize()
            for k in range(KR):
                tg2[k % N_FRAMES].replay()
                torch.cuda.synchronize()
                e = te2[k % N_FRAMES]
                nf_ms.append(e[0].elapsed_time(e[1]))
                nb_ms.append(e[1].elapsed_time(e[2]))
            # ... and alone: back-to-back graph replays on the rotating buffer sets (what tools/sweep.py measures for the
            # concat forms), without the other kernels of the step around them
            def alone(fn):
                for k in range(n_sets):
                    fn(k)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for k in range(2 * n_sets):
                        fn(k)
                gr.replay()
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                ea.record()
                for _ in range(20):
                    gr.replay()
                eb.record()
                torch.cuda.synchronize()
                return ea.elapsed_time(eb) * 1e-3 / (20 * 2 * n_sets)
            cs_ = torch.cuda.current_stream().cuda_stream
            tf_alone = alone(lambda k: pipes_nc[k % n_sets].forward_layer(dom, maps[k % n_sets][dom]["bev"], maps[k % n_sets][dom]["img"], cs_, n_pts[0]))
            tb_alone = alone(lambda k: pipes_nc[k % n_sets].backward_layer(dom, maps[k % n_sets][dom]["g_bev"], maps[k % n_sets][dom]["g_img"], cs_, n_pts[0]))
            bf, bb = sB.bytes_forward_sparse_only(nnz[dom]), sB.bytes_backward_sparse_only(nnz[dom])
            tf_, tb_ = float(np.mean(nf_ms)) * 1e-3, float(np.mean(nb_ms)) * 1e-3
            bytes_step_nc = sum((s.bytes_forward_sparse_only(n) + (s.bytes_backward(n) if s.dual else s.bytes_backward_sparse_only(n)))
                                for s, n in zip(specs, nnz))
            no_concat = {"what": "the same step with every layer in the no-concat form (shpl_pool_forward_into / _into_dual, "
                                 "shpl_pool_backward_from; the dual layer's backward keeps the AddN form): frames/s, and layer %s's "
                                 "kernels against the sparse-only byte formulas of SURVEY.md 8(d)" % sB.name,
                         "value": world * KN / (ms_nc * 1e-3), "unit": UNIT, "steps": KN, "ms_per_step": ms_nc / KN,
                         "algorithmic_bytes_per_step": bytes_step_nc,
                         "forward_kernel": {"bytes_per_launch": bf, "us_per_launch": tf_ * 1e6, "achieved": bf / tf_ / 1e9, "frac": bf / tf_ / 1e9 / peak,
                                            "alone_us_per_launch": tf_alone * 1e6, "alone_frac": bf / tf_alone / 1e9 / peak},
                         "backward_kernel": {"bytes_per_launch": bb, "us_per_launch": tb_ * 1e6, "achieved": bb / tb_ / 1e9, "frac": bb / tb_ / 1e9 / peak,
                                             "alone_us_per_launch": tb_alone * 1e6, "alone_frac": bb / tb_alone / 1e9 / peak},
                         "timing": "us_per_launch / frac: CUDA events around the kernel inside graph replays of the whole single-stream step "
                                   "(launch gaps and the other kernels' dirty L2 lines weigh on a 15 us kernel); alone_*: the kernel back to back "
                                   "on rotating buffers"}
            del pipes_nc, ngr, tg2
        except SystemExit:
            raise
        except Exception as ex:  # pragma: no cover
            print("no-concat leg failed: %r" % (ex,), file=sys.stderr)
            torch.cuda.synchronize()

    # ---- the post-fusion 3x3 conv fused with the pooling (rpn_model.py:335-346, the
    #      rpn_sparse_pooling_conv_after_fusion switch): conv(concat(bev, pooled)) without ever writing the fused map.
    #      Dense half on the tcgen05 tensor cores (3xTF32), pooled half as a Z gather in the epilogue.
    conv = None
    if (sB.c_bev, sB.c_img) == (32, 32) and not sB.dual and not args.no_conv:
        try:
            from sparse_pooling_b200 import conv_fusion
            wt = randn(3, 3, 64, 32) * 0.1
            wt_d = wt[:, :, :32].contiguous()
            outs_c = [torch.empty(1, *sB.bev_hw, 32, device=dev) for _ in range(n_sets)]
            prologue_build(0)
            prologue_build(1)
            torch.cuda.synchronize()
            plans_c = [pipes[si_].layers[dom].plan for si_ in range(n_sets)]
            for pl_ in plans_c:
                pl_.entry_bound = N_MAX
            ws_c = conv_fusion.conv_workspace(1, sB.bev_hw[0], sB.bev_hw[1], dev, N_MAX)

            def conv_call(k, pooled_=True):
                si_ = k % n_sets
                mp_ = maps[si_][dom]
                if pooled_:
                    conv_fusion.sparse_pool_conv3x3([mp_["bev"], mp_["img"]], plans_c[si_], None, wt, None, None, True, out=outs_c[si_], workspace=ws_c)
                else:
                    conv_fusion.sparse_pool_conv3x3([mp_["bev"], None], None, None, wt_d, None, None, True, out=outs_c[si_], workspace=ws_c)

            def time_conv(pooled_):
                for k in range(4):
                    conv_call(k, pooled_)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for k in range(4):
                        conv_call(k, pooled_)
                for _ in range(3):
                    gr.replay()
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_ = 25
                torch.cuda.synchronize()
                e0_.record()
                for _ in range(n_):
                    gr.replay()
                e1_.record()
                torch.cuda.synchronize()
                return e0_.elapsed_time(e1_) * 1e3 / (4 * n_)
            us_fused, us_dense = time_conv(True), time_conv(False)
            # backward of the fused conv (g_bev on the tensor cores, g_img through the transposed CSR, g_weight)
            g_outs_c = [randn(1, *sB.bev_hw, 32) for _ in range(n_sets)]
            grads_c = (torch.empty(1, *sB.bev_hw, 32, device=dev), torch.empty(1, *sB.img_hw, 32, device=dev), torch.empty_like(wt))
            need_b_ = int(shpl._cabi.lib.shpl_conv3x3_backward_workspace_bytes(int(N_MAX)))
            ws_b = torch.empty(need_b_ + 256, dtype=torch.uint8, device=dev)
            _fwd_call = conv_call

            def conv_call(k, pooled_=True):                   # noqa: F811  (time_conv times whatever conv_call is)
                si_ = k % n_sets
                mp_ = maps[si_][dom]
                conv_fusion.sparse_pool_conv3x3_backward(g_outs_c[si_], [mp_["bev"], mp_["img"]], plans_c[si_], wt, out=grads_c, workspace=ws_b)
            us_bwd = time_conv(True)
            conv_call = _fwd_call
            # spot check of the fused result against the concat form + a float64 conv on a crop of the map
            conv_call(0, True)
            fused_ref = pipes[0].layers[dom].fused_bev
            pipes[0].forward_layer(dom, maps[0][dom]["bev"], maps[0][dom]["img"], torch.cuda.current_stream().cuda_stream, N_MAX)
            crop = fused_ref[:, 100:164, 200:328].double().permute(0, 3, 1, 2)
            ref_c = torch.nn.functional.conv2d(crop, wt.double().permute(3, 2, 0, 1)).clamp_min(0).permute(0, 2, 3, 1)
            mag_c = torch.nn.functional.conv2d(crop.abs(), wt.double().abs().permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
            got_c = outs_c[0][:, 101:163, 201:327].double()
            rel_c = float(((got_c - ref_c).abs() / mag_c.clamp_min(1e-30)).max().item())
            if not rel_c <= 1e-5:
                raise SystemExit("bench.py: PARITY CHECK FAILED: fused conv differs from pool -> concat -> conv by %.3e of sum|terms|" % rel_c)
            R_ = sB.R
            flops_alg = 2.0 * R_ * 9 * 64 * 32                      # what slim.conv2d(fused 64 -> 32) computes
            flops_tc = 3 * 2.0 * R_ * 9 * 32 * 32                   # executed on the tensor cores: the dense half, three TF32 products
            try:
                bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
                tf_src = "measured bf16 burst peak / 2 (MEASURED_PEAKS.json bf16_tflops)"
            except Exception:
                bf16_peak, tf_src = 1590.0, "fallback bf16 1.59 PFLOP/s / 2 (B200_PROFILING.md)"
            tf32_peak = bf16_peak / 2.0
            io_bytes = 4.0 * R_ * (32 + 32) + 4.0 * nnz[dom] * 34
            conv = {"what": "shpl_pool_conv3x3_forward on layer %s: relu(conv3x3(concat(bev, pooled(img)), W[3,3,64,32])) with the fused map never "
                            "written; CUDA-graph replays, CUDA events; includes the weight prep, the busy-cell bitmap and the Z kernel" % sB.name,
                    "us_per_call": us_fused, "us_dense_half_only": us_dense, "us_backward_call": us_bwd,
                    "max_err_over_sum_abs_terms_on_a_64x128_crop": rel_c,
                    "roofline": {"bound": "tensor", "unit": "TFLOP/s",
                                 "achieved": flops_tc / us_dense / 1e6, "peak": tf32_peak, "frac": flops_tc / us_dense / 1e6 / tf32_peak,
                                 "peak_source": tf_src + "; TF32 dense = half the bf16 rate",
                                 "note": "dense-half kernel alone (C_s = 0 call): executed tensor flops = 3 TF32 products per fp32 product (3xTF32 "
                                         "split for 1e-5 parity). The kernel is bound by the tensor core's shared-memory operand fetch at N = 192 "
                                         "(profiles/r2_conv_*), not by this peak",
                                 "fp32_equivalent_tflops_of_the_whole_conv": flops_alg / us_fused / 1e6},
                    "hbm": {"bytes_per_call": io_bytes, "achieved_GBs": io_bytes / us_fused / 1e3, "frac_of_peak": io_bytes / us_fused / 1e3 / peak},
                    "unfused_bytes": {"pool_forward_concat": sB.bytes_forward(nnz[dom]), "conv_reads_fused_writes_out": 4.0 * R_ * (64 + 32),
                                      "fused_path": io_bytes}}
        except SystemExit:
            raise
        except Exception as ex:  # pragma: no cover
            print("conv leg failed: %r" % (ex,), file=sys.stderr)
            torch.cuda.synchronize()

    # ---- e2e: the public drop-in API, frame inputs in pinned host memory, result read back
    class Calib:
        p2 = P
    pts_pin = [torch.from_numpy(f["points"]).pin_memory() for f in frames_host]
    vox_pin = [torch.from_numpy(np.ascontiguousarray(f["voxel_indices"][:, :2])).pin_memory() for f in frames_host]
    result_pin = torch.empty(2 * nL * 256, dtype=torch.float32).pin_memory()
    h2d = d2h = 0

    def e2e_step(k):
        nonlocal h2d, d2h
        fi, si = k % N_FRAMES, k % n_sets
        mp = maps[si]
        outs, roots, grads, leaves = [], [], [], []
        h2d = d2h = 0
        for li, s in enumerate(specs):
            d = shpl.gen_sparse_pooling_input_avod(pts_pin[fi], vox_pin[fi], Calib, list(s.im_size), s.bv_size)
            h2d += pts_pin[fi].numel() * 8 + vox_pin[fi].numel() * 8
            d2h += 4                                                    # n read-back
            o = shpl.produce_sparse_pooling_input(d, stride=list(s.stride))
            d2h += 32                                                   # counts read-back
            M = shpl.SparseTensor.from_sparse_pooling_input(o)
            bev = mp[li]["bev"].requires_grad_(True)
            img = mp[li]["img"].requires_grad_(True)
            bev.grad = None
            img.grad = None
            bv_fused, img_fused = shpl.sparse_pool_layer([bev, img], [s.c_img, s.c_bev], M,
                                                         img_index_flip=o["img_index_flip_pool"],
                                                         bv_index=(np.zeros((1, 3)) if s.dual else None))
            roots.append(bv_fused)
            grads.append(mp[li]["g_bev"])
            if s.dual:
                roots.append(img_fused)
                grads.append(mp[li]["g_img"])
            leaves.append((bev, img))
        torch.autograd.backward(roots, grads)      # one backward over all layers, like one sess.run(train_op)
        for bev, img in leaves:
            outs.append(bev.grad.reshape(-1)[:256])
            outs.append(img.grad.reshape(-1)[:256])
        result_pin.copy_(torch.cat(outs), non_blocking=False)          # D2H read of the step's result
        d2h += result_pin.numel() * 4

    K_e2e = max(3, min(K, 30))
    for k in range(3):
        e2e_step(k)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        e2e_step(k)
    torch.cuda.synchronize()
    dt_e2e = D.max(time.perf_counter() - t0)
    for mp in maps:
        for m in mp:
            m["bev"].requires_grad_(False)
            m["img"].requires_grad_(False)
    e2e = {"value": world * K_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": K_e2e,
           "what": "gen_sparse_pooling_input_avod + produce_sparse_pooling_input + sparse_pool_layer per layer, then one autograd "
                   "backward over all layers, through the public API; points/voxel indices copied from pinned host memory every step, "
                   "feature maps device-resident (they are device-resident TF tensors in the reference), the heads of the "
                   "gradients read back to pinned host memory every step"}

    # ---- e2e through the bare C ABI (the lean ctypes pipeline): frame inputs copied from pinned host memory and a
    #      result read back every step, no cross-frame pipelining (the step ends with a host read)
    stage_pts = torch.empty((N_MAX, 3), dtype=torch.float64, device=dev)
    stage_vox = torch.empty((N_MAX, 2), dtype=torch.int64, device=dev)
    res_dev = torch.empty(2 * nL * 256 + 8 * nL, dtype=torch.float32, device=dev)
    res_pin = torch.empty_like(res_dev, device="cpu").pin_memory()

    def pool_all(pipe, mp, pts, vox, n, n_dev=None):
        """build + forward + backward of every layer: the last layer on the current stream, the others on `side`; every
        layer's builder on its own high-priority stream when the step is being captured into a CUDA graph (its short
        latency-bound kernels are then scheduled ahead of the pooling CTAs of the step before, which shares the GPU with
        it: e2e 7 740 -> 7 855 frames/s; with eager launches the extra stream operations cost more than they gain)"""
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        hp = torch.cuda.is_current_stream_capturing()
        if hp:
            for li in range(nL):
                bst = build_streams[li]
                bst.wait_stream(main)
                with torch.cuda.stream(bst):
                    pipe.build_layer(li, pts, vox, P, n, bst.cuda_stream, n_dev=n_dev)
        if nL > 1:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ss = side.cuda_stream
                for li in range(nL - 1):
                    if hp:
                        side.wait_stream(build_streams[li])
                    else:
                        pipe.build_layer(li, pts, vox, P, n, ss, n_dev=n_dev)
                    pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], ss, n)
                    pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], ss, n)
        li = nL - 1
        if hp:
            main.wait_stream(build_streams[li])
        else:
            pipe.build_layer(li, pts, vox, P, n, ms, n_dev=n_dev)
        pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], ms, n)
        pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], ms, n)
        if nL > 1:
            main.wait_stream(side)

    def gather_result(pipe):
        off = 0
        for li in range(nL):
            res_dev[off:off + 256].copy_(pipe.layers[li].g_bev.reshape(-1)[:256])
            res_dev[off + 256:off + 512].copy_(pipe.layers[li].g_img.reshape(-1)[:256])
            off += 512
        res_dev[off:off + 8 * nL].copy_(torch.cat([L.plan.counts.reshape(-1)[:8] for L in pipe.layers]).float())
        res_pin.copy_(res_dev, non_blocking=False)

    def cabi_step(k):
        fi, si = k % N_FRAMES, k % n_sets
        n = n_pts[fi]
        stage_pts[:n].copy_(pts_pin[fi], non_blocking=True)
        stage_vox[:n].copy_(vox_pin[fi], non_blocking=True)
        pool_all(pipes[si], maps[si], stage_pts, stage_vox, n)
        gather_result(pipes[si])

    for k in range(3):
        cabi_step(k)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        cabi_step(k)
    torch.cuda.synchronize()
    dt_cabi = D.max(time.perf_counter() - t0)
    e2e["c_abi_pipeline"] = {"value": world * K_e2e / dt_cabi, "unit": UNIT,
                             "h2d_bytes_per_step": int(n_pts[0] * 40), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                             "what": "same step through the ctypes C-ABI calls on preallocated buffers: points/voxel indices "
                                     "copied from pinned host memory, every plan built, forward+backward of every layer, "
                                     "heads of the gradients + the plan counters read back to pinned host memory, every step"}

    # ---- the same C-ABI step, software-pipelined by one frame: every step still uploads its inputs from pinned host
    #      memory and its result is still read back inside the timed region, but the host waits for the result of step
    #      k-1 (event) after it has enqueued step k, so enqueueing overlaps the GPU work (double-buffered staging / result
    #      buffers; what a data-loader thread in front of a training loop does)
    stage2 = [(torch.empty((N_MAX, 3), dtype=torch.float64, device=dev), torch.empty((N_MAX, 2), dtype=torch.int64, device=dev))
              for _ in range(2)]
    n_res = 2 * nL * 256 + 8 * nL
    res_dev2 = [torch.empty(n_res, dtype=torch.int32, device=dev) for _ in range(2)]
    res_pin2 = [torch.empty(n_res, dtype=torch.int32).pin_memory() for _ in range(2)]
    res_done = [torch.cuda.Event() for _ in range(2)]
    # the words read back per buffer set: heads of the gradients (fp32 bits) and every plan's counters
    res_views = []
    for pipe in pipes:
        v = []
        for L in pipe.layers:
            v += [L.g_bev.reshape(-1)[:256].view(torch.int32), L.g_img.reshape(-1)[:256].view(torch.int32)]
        v += [L.plan.counts.reshape(-1)[:8] for L in pipe.layers]
        res_views.append(v)
    seen = []
    nnz_word = 2 * nL * 256 + 1          # counts[1] of the first layer's plan, inside the words read back

    def cabi_enqueue(k):
        fi, si, b = k % N_FRAMES, k % n_sets, k % 2
        n = n_pts[fi]
        sp, sv = stage2[b]
        sp[:n].copy_(pts_pin[fi], non_blocking=True)
        sv[:n].copy_(vox_pin[fi], non_blocking=True)
        pool_all(pipes[si], maps[si], sp, sv, n)
        torch.cat(res_views[si], out=res_dev2[b])
        res_pin2[b].copy_(res_dev2[b], non_blocking=True)

    def read_results(k, last):
        b = k % 2
        res_done[b].record(torch.cuda.current_stream())
        for bb in ((1 - b,) if not last else (1 - b, b)):        # the host read: step k-1's result (and k's at the end)
            res_done[bb].synchronize()
            seen.append(int(res_pin2[bb][nnz_word]))              # nnz of the first layer's plan, out of the words just read

    def cabi_step_pipelined(k, last=False):
        cabi_enqueue(k)
        read_results(k, last)

    K_fast = max(K_e2e, min(KT, 2000))          # the pipelined legs are cheap: time as many steps as the device-resident leg
    for k in range(3):
        cabi_step_pipelined(k, last=(k == 2))
    torch.cuda.synchronize()
    barrier()
    seen.clear()
    t0 = time.perf_counter()
    for k in range(K_fast):
        cabi_step_pipelined(k, last=(k == K_fast - 1))
    torch.cuda.synchronize()
    dt_pipe = D.max(time.perf_counter() - t0)
    assert len(seen) == K_fast + 1 and all(x > 0 for x in seen[1:]), "pipelined C-ABI leg: a result was not read back"
    e2e["c_abi_pipelined"] = {"value": world * K_fast / dt_pipe, "unit": UNIT, "steps": K_fast,
                              "h2d_bytes_per_step": int(n_pts[0] * 40), "d2h_bytes_per_step": int(res_pin2[0].numel() * 4),
                              "what": "the C-ABI step software-pipelined by one frame: inputs uploaded from pinned host memory and "
                                      "the gradient heads + plan counters read back for EVERY step inside the timed region, the host "
                                      "waiting for step k-1's result after enqueueing step k (double-buffered staging)"}

    # ---- and with the enqueueing itself captured: one CUDA graph per (frame, buffer set) holding the two uploads from
    #      the pinned host buffers, the builds, forward + backward of every layer and the D2H of the result; a step is one
    #      graph launch + the wait for the previous step's result
    lanes = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    n_combo = max(N_FRAMES, n_sets, 2)          # smallest k-period of (frame, buffer set, staging buffer)
    while n_combo % N_FRAMES or n_combo % n_sets or n_combo % 2:
        n_combo += 1
    try:
        e2e_graphs = []
        for k in range(n_combo):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                cabi_enqueue(k)
            e2e_graphs.append(gr)

        def cabi_step_graph(k, last=False):
            # consecutive steps use disjoint staging / plan / result buffers (k % 2): launched on alternating streams
            # the uploads and builds of step k+1 overlap the pooling of step k; steps k and k+2 share a stream, so the
            # buffers they share are reused in order
            with torch.cuda.stream(lanes[k % 2]):
                e2e_graphs[k % n_combo].replay()
                read_results(k, last)

        for k in range(4):
            cabi_step_graph(k, last=(k == 3))
        torch.cuda.synchronize()
        barrier()
        seen.clear()
        t0 = time.perf_counter()
        for k in range(K_fast):
            cabi_step_graph(k, last=(k == K_fast - 1))
        torch.cuda.synchronize()
        dt_g = D.max(time.perf_counter() - t0)
        assert len(seen) == K_fast + 1 and all(x > 0 for x in seen[1:]), "graph C-ABI leg: a result was not read back"
        e2e["c_abi_graph"] = {"value": world * K_fast / dt_g, "unit": UNIT, "steps": K_fast,
                              "h2d_bytes_per_step": int(n_pts[0] * 40), "d2h_bytes_per_step": int(res_pin2[0].numel() * 4),
                              "what": "the pipelined C-ABI step captured in CUDA graphs (uploads from the pinned host buffers, builds, "
                                      "forward + backward, D2H of the result all inside the graph): one graph launch per step on "
                                      "alternating streams (double-buffered), every step's result read on the host one step later"}
    except Exception as ex:  # pragma: no cover
        print("graph C-ABI leg failed: %r" % (ex,), file=sys.stderr)
        torch.cuda.synchronize()

    # ---- the headline e2e: the package's batched pipeline API (pipeline.FramePipeline = the C-ABI calls) fed from host
    #      buffers, in the form a deployment runs it (graph launch per step, double-buffered); the reference-signature
    #      Python functions, which synchronise inside every call like the numpy code they mirror, are reported beside it
    e2e["python_dropin_api"] = {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps", "what")}
    head = "c_abi_graph" if "c_abi_graph" in e2e else "c_abi_pipelined"
    for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps"):
        e2e[k] = e2e[head][k]
    e2e["leg"] = head
    e2e["what"] = ("FramePipeline (ctypes -> C ABI of include/shpl.h) fed from pinned HOST buffers: per step the frame's points / "
                   "voxel indices are uploaded, every plan built, every layer run forward + backward and the gradient heads + the "
                   "plan counters copied back and read on the host -- " + e2e[head]["what"] + ".  Other legs: python_dropin_api "
                   "(the reference-signature functions, one frame at a time, synchronising like the reference), c_abi_pipeline "
                   "(no pipelining: every step ends with its own host read), c_abi_pipelined (eager launches)")

    # ---- the feeder in front of the path: BevSlices.generate_bev of the raw scan on the
    #      GPU (shpl_bev_slices), its pair count handed to the builders on the device (no host read in between)
    feeder = None
    if name == "2" and not args.no_feeder:
        try:
            feeder = feeder_legs(D, K, K_e2e, K_fast, specs, pipes, maps, n_sets, n_pts, P, pool_all, gather_result, res_pin, res_views,
                                 res_dev2, res_pin2, read_results, seen, lanes, n_combo, nnz_word)
        except Exception as ex:  # pragma: no cover
            print("feeder leg failed: %r" % (ex,), file=sys.stderr)
            torch.cuda.synchronize()

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(cfg)

    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / KT, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(name, cfg, world),
            "run": {"timed_blocks": reps, "timed_steps": KT, "timed_ms": ms_total, "min_timed_s": MIN_TIMED_S,
                    "candidate_pairs_per_frame": n_pts, "nnz_per_layer": nnz,
                    "launch": ("CUDA graph replay, %d steps per graph where the step index lines up" % G if multi is not None
                               else "CUDA graph replay, one step per graph") if use_graph else "eager launches",
                    "algorithmic_bytes_per_step": bytes_step, "step_gbs_per_gpu": step_gbs, "step_frac_of_peak": step_gbs / peak,
                    "parity_check": parity},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline,
        }
        if no_concat is not None:
            line["no_concat"] = no_concat
        if conv is not None:
            line["conv_after_fusion"] = conv
        if feeder is not None:
            line["feeder"] = feeder
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    D.close()
    return 0


def feeder_legs(D, K, K_e2e, K_fast, specs, pipes, maps, n_sets, n_pts, P, pool_all, gather_result, res_pin, res_views, res_dev2,
                res_pin2, read_results, seen, lanes, n_combo, nnz_word):
    """Config 2 only: the BEV slicing feeder, and the whole chain from the camera-frame scan / from a raw velodyne scan."""
    torch = D.torch
    dev, rank, world = D.dev, D.rank, D.world
    from sparse_pooling_b200 import _cabi, bev_slices as bs
    from oracle import feeder_oracle as fo
    from tools import synth
    nL = len(specs)
    GP = np.array([0.0, -1.0, 0.0, 1.65])
    scans = [np.ascontiguousarray(synth.lidar_scan(100 + rank * N_FRAMES + i, az_step_deg=AZ_STEP).T) for i in range(N_FRAMES)]
    scan_pin = [torch.from_numpy(sc).pin_memory() for sc in scans]
    p_max = max(sc.shape[1] for sc in scans)
    stage_scan = torch.empty((3, p_max), dtype=torch.float64, device=dev)
    work = bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, N_MAX, dev, with_maps=True)
    lut = torch.from_numpy(bs.density_lut(np.log(16))).to(dev)
    n_dev = ctypes.c_void_p(work.counts.data_ptr())

    def feeder_call(fi, src=None):
        sc = stage_scan if src is None else src
        Pn = scans[fi].shape[1]
        bs.bev_slices_raw(sc, sc.stride(0), sc.stride(1), Pn, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5,
                          np.log(16), work, lut=lut)

    scan_dev = [torch.from_numpy(sc).to(dev) for sc in scans]
    for fi in range(N_FRAMES):
        feeder_call(fi, scan_dev[fi])
    torch.cuda.synchronize()
    assert int(work.counts[0].item()) == n_pts[N_FRAMES - 1], "feeder pair count differs from the frame's"
    fg = []
    for fi in range(N_FRAMES):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            feeder_call(fi, scan_dev[fi])
        fg.append(gr)
    evf0, evf1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KF = max(K, 50)
    torch.cuda.synchronize()
    evf0.record()
    for k in range(KF):
        fg[k % N_FRAMES].replay()
    evf1.record()
    torch.cuda.synchronize()
    feeder_us = evf0.elapsed_time(evf1) * 1e3 / KF

    def scan_step(k):
        fi, si = k % N_FRAMES, k % n_sets
        Pn = scans[fi].shape[1]
        stage_scan[:, :Pn].copy_(scan_pin[fi], non_blocking=True)
        feeder_call(fi)
        pool_all(pipes[si], maps[si], work.unique_pts, work.voxel_indices, N_MAX, n_dev=n_dev)
        gather_result(pipes[si])

    for k in range(3):
        scan_step(k)
    torch.cuda.synchronize()
    ref_counts = [int(x) for x in res_pin[-8 * nL:].tolist()]
    D.barrier()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        scan_step(k)
    torch.cuda.synchronize()
    dt_scan = D.max(time.perf_counter() - t0)
    t0 = time.perf_counter()
    for fi in range(N_FRAMES):
        fo.generate_bev(scans[fi], GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / N_FRAMES

    # ---- the whole widened chain from a RAW velodyne scan: float32 [N,4] in pinned host
    #      memory -> H2D -> ingest (camera frame + FOV filter) -> feeder -> every plan -> forward -> backward ->
    #      D2H, every intermediate count handed on as a device pointer (no host read inside the step)
    from sparse_pooling_b200 import lidar_ingest as li
    import types as _types
    cal = _types.SimpleNamespace(p2=P, r0_rect=synth.R0_RECT_KITTI, tr_velodyne_to_cam=synth.TR_VELO_TO_CAM_KITTI)
    velos = [synth.velodyne_scan(300 + rank * N_FRAMES + i, az_step_deg=0.09) for i in range(N_FRAMES)]
    velo_pin = [torch.from_numpy(v).pin_memory() for v in velos]
    v_max = max(v.shape[0] for v in velos)
    stage_velo = torch.empty((v_max, 4), dtype=torch.float32, device=dev)
    cam_buf = torch.empty((3, v_max), dtype=torch.float64, device=dev)
    ing_counts = torch.zeros(4, dtype=torch.int32, device=dev)
    p_dev = ctypes.c_void_p(ing_counts.data_ptr())

    def velo_step(k):
        fi, si = k % N_FRAMES, k % n_sets
        nv = velos[fi].shape[0]
        stage_velo[:nv].copy_(velo_pin[fi], non_blocking=True)
        li.lidar_to_cam_raw(stage_velo, nv, cal, [1242, 375], cam_buf, ing_counts)
        bs.bev_slices_raw(cam_buf, cam_buf.stride(0), cam_buf.stride(1), nv, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                          -0.2, 2.3, 5, np.log(16), work, lut=lut, p_dev=p_dev)
        pool_all(pipes[si], maps[si], work.unique_pts, work.voxel_indices, N_MAX, n_dev=n_dev)
        gather_result(pipes[si])

    for k in range(3):
        velo_step(k)
    torch.cuda.synchronize()
    velo_counts = [int(x) for x in res_pin[-8 * nL:].tolist()][:4] + [int(ing_counts[0].item()), int(work.counts[0].item())]
    D.barrier()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        velo_step(k)
    torch.cuda.synchronize()
    dt_velo = D.max(time.perf_counter() - t0)
    # ---- the same chain as the headline e2e leg runs it: captured in CUDA graphs (upload of the raw scan, ingest,
    #      feeder, builds, pooling, D2H), launched on alternating streams with every buffer of the chain doubled,
    #      each step's result read on the host one step later
    velo_graph = None
    try:
        vsets = []
        for b in range(2):
            vsets.append(dict(stage=torch.empty((v_max, 4), dtype=torch.float32, device=dev),
                              cam=torch.empty((3, v_max), dtype=torch.float64, device=dev),
                              cnt=torch.zeros(4, dtype=torch.int32, device=dev),
                              work=bs.BevWorkspace(synth.AVOD_EXTENTS, synth.AVOD_VOXEL, 5, N_MAX, dev, with_maps=True),
                              ws=torch.empty(int(_cabi.lib.shpl_lidar_workspace_bytes(int(v_max))) + 64, dtype=torch.uint8, device=dev)))

        def velo_enqueue(k):
            fi, si, b = k % N_FRAMES, k % n_sets, k % 2
            vs = vsets[b]
            nv = velos[fi].shape[0]
            vs["stage"][:nv].copy_(velo_pin[fi], non_blocking=True)
            li.lidar_to_cam_raw(vs["stage"], nv, cal, [1242, 375], vs["cam"], vs["cnt"], ws=vs["ws"])
            bs.bev_slices_raw(vs["cam"], vs["cam"].stride(0), vs["cam"].stride(1), nv, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                              -0.2, 2.3, 5, np.log(16), vs["work"], lut=lut, p_dev=ctypes.c_void_p(vs["cnt"].data_ptr()))
            nd = ctypes.c_void_p(vs["work"].counts.data_ptr())
            pool_all(pipes[si], maps[si], vs["work"].unique_pts, vs["work"].voxel_indices, N_MAX, n_dev=nd)
            torch.cat(res_views[si], out=res_dev2[b])
            res_pin2[b].copy_(res_dev2[b], non_blocking=True)

        for k in range(n_combo):
            velo_enqueue(k)
        torch.cuda.synchronize()
        vgraphs = []
        for k in range(n_combo):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                velo_enqueue(k)
            vgraphs.append(gr)

        def velo_step_graph(k, last=False):
            with torch.cuda.stream(lanes[k % 2]):
                vgraphs[k % n_combo].replay()
                read_results(k, last)

        for k in range(4):
            velo_step_graph(k, last=(k == 3))
        torch.cuda.synchronize()
        vg_counts = [int(x) for x in res_pin2[0][-8 * nL:].tolist()][:4]        # step k = 2: the frame the eager check read
        D.barrier()
        seen.clear()
        t0 = time.perf_counter()
        for k in range(K_fast):
            velo_step_graph(k, last=(k == K_fast - 1))
        torch.cuda.synchronize()
        dt_vg = D.max(time.perf_counter() - t0)
        assert len(seen) == K_fast + 1 and all(x > 0 for x in seen[1:]), "velodyne graph leg: a result was not read back"
        assert vg_counts == velo_counts[:4], "velodyne graph leg: plan counters differ from the eager chain's"
        velo_graph = {"value": world * K_fast / dt_vg, "unit": UNIT, "steps": K_fast,
                      "h2d_bytes_per_step": int(velos[0].shape[0] * 16), "d2h_bytes_per_step": int(res_pin2[0].numel() * 4),
                      "what": "the velodyne chain captured in CUDA graphs (upload of the raw scan, ingest, feeder, every plan, "
                              "forward + backward, D2H), one graph launch per step on alternating streams with every buffer "
                              "doubled, every step's result read on the host one step later"}
    except Exception as ex:  # pragma: no cover
        print("velodyne graph leg failed: %r" % (ex,), file=sys.stderr)
        torch.cuda.synchronize()

    return {"what": "BevSlices.generate_bev(output_indices=True) on the GPU: 5 height maps + density map [6,700,800] f64, "
                    "voxel_indices, unique_pts (shpl_bev_slices, CUDA-graph replays, CUDA events)",
            "us_per_frame": feeder_us, "points_per_scan": [int(sc.shape[1]) for sc in scans],
            "cpu_oracle_ms_per_frame": cpu_ms,
            "e2e_from_scan": {"value": world * K_e2e / dt_scan, "unit": UNIT,
                              "h2d_bytes_per_step": int(scans[0].shape[1] * 24), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                              "what": "raw scan [3,P] f64 copied from pinned host memory, feeder, every plan built from the "
                                      "feeder's device-side pair count, forward+backward of every layer, gradients + plan "
                                      "counters read back, every step (ctypes C-ABI calls)"},
            "plan_counts_check": ref_counts[:4],
            "e2e_from_velodyne": {"value": world * K_e2e / dt_velo, "unit": UNIT,
                                  "h2d_bytes_per_step": int(velos[0].shape[0] * 16), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                                  "points_per_scan": [int(v.shape[0]) for v in velos],
                                  "what": "raw 360-degree velodyne scan float32 [N,4] from pinned host memory, ingest (camera "
                                          "frame, FOV filter), feeder, every plan, forward+backward of every layer, read-back; "
                                          "all intermediate counts stay on the device",
                                  "counts_check(nclip,nnz,oob,csr,fov_points,pairs)": velo_counts},
            "e2e_from_velodyne_graph": velo_graph}


# ------------------------------------------------------------------------------- GPU arm, pair-type configurations (3, 4)
def run_gpu_pairs(args, name, cfg):
    D = Dist()
    torch = D.torch
    world, rank, local_rank, dev = D.world, D.rank, D.local_rank, D.dev
    from sparse_pooling_b200 import _cabi
    from sparse_pooling_b200.pipeline import PairsPipeline
    from sparse_pooling_b200.sharding import frames_for_rank
    lib = _cabi.lib
    spec = cfg["layers"][0]
    K, W = args.steps, max(args.warmup, 3)
    strong = cfg["scaling"] == "strong"
    # frames this rank pools per step: config 4 shards ONE fixed batch of 32 by frame (strong scaling); config 3 gives
    # every rank its own batch of 8 (weak scaling)
    my_frames = list(frames_for_rank(cfg["batch"], rank, world)) if strong else list(range(cfg["batch"]))
    F = len(my_frames)
    stacked = not strong                  # config 3: the batch is ONE plan and one launch each way; config 4: frame by frame
    n_distinct = 4
    host = [pairs_frame_inputs(cfg, 1000 + 17 * rank + i) for i in range(n_distinct)]
    n_pairs = [int(h["img_index"].shape[1]) for h in host]
    n_max = max(n_pairs)
    uv_dev = [torch.from_numpy(np.ascontiguousarray(h["img_index"][0:2])).to(dev) for h in host]
    bv_dev = [torch.from_numpy(h["bv_index"]).to(dev) for h in host]
    mv_dev = [torch.from_numpy(h["m_val"]).to(dev) for h in host]
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + rank)
    Bp = F if stacked else 1              # frames per pipeline / per launch
    n_sets = 2

    def randn(*shape):
        return torch.randn(*shape, device=dev, dtype=torch.float32, generator=g)
    maps = [dict(bev=randn(Bp, *spec.bev_hw, spec.c_bev), img=randn(Bp, *spec.img_hw, spec.c_img),
                 g_bev=randn(Bp, *spec.bev_hw, spec.c_bev + spec.c_img)) for _ in range(n_sets)]
    pipes = [PairsPipeline(spec, Bp, n_max, dev) for _ in range(n_sets)]
    side = torch.cuda.Stream(device=dev, priority=-1)

    def frame_of(k, j):
        """which distinct synthetic frame slot j of step k holds"""
        return (k * F + j) % n_distinct

    def build_unit(pipe, k, u, stream):
        """plans of launch unit u of step k (stacked: the whole batch; else: frame u)"""
        if stacked:
            for f in range(F):
                d = frame_of(k, f)
                pipe.build_frame(f, uv_dev[d], bv_dev[d], mv_dev[d], n_pairs[d], stream)
        else:
            d = frame_of(k, u)
            pipe.build_frame(0, uv_dev[d], bv_dev[d], mv_dev[d], n_pairs[d], stream)

    units = 1 if stacked else F
    bound = (F if stacked else 1) * n_max

    def lean_step(k, timing_events=None, overlap=True):
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        for u in range(units):
            t = k * units + u                       # running unit index: buffer sets alternate per unit
            pipe, mp = pipes[t % n_sets], maps[t % n_sets]
            if not overlap:
                build_unit(pipe, k, u, ms)
            else:
                # the plans of the NEXT unit are built on the side stream while this unit is pooled
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    kn, un = (k, u + 1) if u + 1 < units else (k + 1, 0)
                    build_unit(pipes[(t + 1) % n_sets], kn, un, side.cuda_stream)
            if timing_events is not None and u == 0:
                timing_events[0].record(main)
            pipe.forward(mp["bev"], mp["img"], ms, bound)
            if timing_events is not None and u == 0:
                timing_events[1].record(main)
            pipe.backward(mp["g_bev"], ms, bound)
            if timing_events is not None and u == 0:
                timing_events[2].record(main)
            if overlap:
                main.wait_stream(side)

    def prologue_build(k):
        build_unit(pipes[(k * units) % n_sets], k, 0, torch.cuda.current_stream().cuda_stream)

    # ---- PARITY CHECK before timing: the first launch unit of step 0 (build + forward + backward through PairsPipeline ->
    #      C ABI) against the CPU oracle, bit for bit
    ms0 = torch.cuda.current_stream().cuda_stream
    build_unit(pipes[0], 0, 0, ms0)
    pipes[0].forward(maps[0]["bev"], maps[0]["img"], ms0, bound)
    pipes[0].backward(maps[0]["g_bev"], ms0, bound)
    torch.cuda.synchronize()
    c0 = pipes[0].plan.counts.cpu().numpy()
    nnz = [int(c0[f, 1]) for f in range(Bp)]
    n_csr = int(c0[:, 3].sum())
    parity = {"checked": False}
    if not args.no_parity_check:
        from oracle import cref, index_oracle as io
        cref.build()
        check_frames = range(min(Bp, 2))
        for f in check_frames:
            h = host[frame_of(0, f if stacked else 0)]
            o = io.produce_sparse_pooling_input(dict(img_index=h["img_index"].copy(), bv_index=h["bv_index"], img_size=h["img_size"],
                                                     bv_size=h["bv_size"]), M_val=h["m_val"], stride=list(spec.stride))
            val = np.asarray(o["M_val"], dtype=np.float32)
            Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
            assert len(val) == nnz[f], "nnz differs from the oracle's"
            bev, img, gb_ = (maps[0][k_][f].cpu().numpy() for k_ in ("bev", "img", "g_bev"))
            assert_equal_bits(pipes[0].fused_bev[f], cref.forward(bev, img, Mij, val, flip), "frame %d: fused BEV map" % f)
            gd, gs = cref.backward(gb_, Mij, val, flip, spec.c_bev, img.shape)
            assert_equal_bits(pipes[0].g_bev[f], gd, "frame %d: gradient of the BEV map" % f)
            assert_equal_bits(pipes[0].g_img[f], gs, "frame %d: gradient of the image map" % f)
        parity = {"checked": True, "frames": len(list(check_frames)),
                  "what": "PairsPipeline (build + forward + backward) == oracle/index_oracle + oracle/cref, bit for bit"}
    prologue_build(0)
    lean_step(0)
    torch.cuda.synchronize()

    # ---- CUDA graphs of the step
    use_graph = not args.no_graph
    graphs = []
    period = n_distinct * n_sets            # k-period of (frames, buffer sets)
    if use_graph:
        try:
            for v in range(period):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    lean_step(v)
                graphs.append(gr)
        except Exception as e:  # pragma: no cover
            print("graph capture failed (%r); timing eager launches" % (e,), file=sys.stderr)
            graphs, use_graph = [], False
            torch.cuda.synchronize()

    def run_steps(k0, k1):
        for k in range(k0, k1):
            if use_graph:
                graphs[k % period].replay()
            else:
                lean_step(k)

    prologue_build(0)
    evw0, evw1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_steps(0, 1)
    evw0.record()
    run_steps(1, W)
    evw1.record()
    torch.cuda.synchronize()
    est = D.max(evw0.elapsed_time(evw1) * 1e-3 / max(W - 1, 1))
    reps = timed_blocks(est, K)
    KT = K * reps
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    run_steps(W, W + KT)
    ev1.record()
    torch.cuda.synchronize()
    D.barrier()
    clocks = sampler.stop()
    ms_total = D.max(ev0.elapsed_time(ev1))
    l0 = int(lib.shpl_kernel_launches())
    lean_step(0)
    launches_per_step = int(lib.shpl_kernel_launches()) - l0
    torch.cuda.synchronize()
    frames_per_step = cfg["batch"] * (1 if strong else world)
    value = frames_per_step * KT / (ms_total * 1e-3)

    # ---- roofline: the forward launch (one frame at config 4, the stacked batch at config 3)
    peak, peak_src = peaks()
    fwd_ms, bwd_ms = [], []
    KR = max(K, 10)
    tg, tev = [], []
    for v in range(period):
        ev3 = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            lean_step(v, ev3, overlap=False)
        tg.append(gr)
        tev.append(ev3)
    for k in range(3):
        tg[k % period].replay()
    torch.cuda.synchronize()
    for k in range(KR):
        tg[k % period].replay()
        torch.cuda.synchronize()
        e = tev[k % period]
        fwd_ms.append(e[0].elapsed_time(e[1]))
        bwd_ms.append(e[1].elapsed_time(e[2]))
    nnz_launch = n_csr
    bytes_fwd = Bp * spec.bytes_forward(0) + 4 * nnz_launch * (spec.c_img + 2)
    bytes_bwd = Bp * spec.bytes_backward(0) + 4 * nnz_launch * (spec.c_img + 2)
    fwd_avg, bwd_avg = float(np.mean(fwd_ms)) * 1e-3, float(np.mean(bwd_ms)) * 1e-3
    achieved = bytes_fwd / fwd_avg / 1e9
    roofline = {"bound": "hbm", "kernel": "shpl_pool_sparse_kernel as the forward of %s (%d frame%s per launch)" % (spec.name, Bp, "s" if Bp > 1 else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "bytes_per_launch": bytes_fwd, "us_per_launch": fwd_avg * 1e6, "peak_source": peak_src,
                "frac_of_8TBs_nominal": achieved / 8000.0,
                "timing": "CUDA events recorded inside CUDA-graph replays of the single-stream step",
                "backward_kernel": {"achieved": bytes_bwd / bwd_avg / 1e9, "frac": bytes_bwd / bwd_avg / 1e9 / peak,
                                    "bytes_per_launch": bytes_bwd, "us_per_launch": bwd_avg * 1e6}}
    bytes_step = units * (bytes_fwd + bytes_bwd)
    step_gbs = bytes_step * KT / (ms_total * 1e-3) / 1e9

    # ---- e2e: every frame's pairs (img_index rows, bv_index, M_val) in pinned HOST memory, uploaded inside the timed
    #      region; the heads of the gradients + the plan counters read back for every step, one step later
    uv_pin = [torch.from_numpy(np.ascontiguousarray(h["img_index"][0:2])).pin_memory() for h in host]
    bv_pin = [torch.from_numpy(h["bv_index"]).pin_memory() for h in host]
    mv_pin = [torch.from_numpy(h["m_val"]).pin_memory() for h in host]
    n_slots = 2 * max(units, 1) if not stacked else 2 * F
    st_uv = [torch.empty((2, n_max), dtype=torch.float64, device=dev) for _ in range(n_slots)]
    st_bv = [torch.empty((n_max, 2), dtype=torch.int64, device=dev) for _ in range(n_slots)]
    st_mv = [torch.empty(n_max, dtype=torch.float64, device=dev) for _ in range(n_slots)]
    n_res = 512 + 8
    res_dev2 = [torch.empty(n_res, dtype=torch.int32, device=dev) for _ in range(2)]
    res_pin2 = [torch.empty(n_res, dtype=torch.int32).pin_memory() for _ in range(2)]
    res_done = [torch.cuda.Event() for _ in range(2)]
    seen = []

    def e2e_enqueue(k):
        b = k % 2
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        for u in range(units):
            t = k * units + u
            pipe, mp = pipes[t % n_sets], maps[t % n_sets]
            for f in (range(F) if stacked else (0,)):
                d = frame_of(k, f if stacked else u)
                slot = b * (n_slots // 2) + (f if stacked else u)
                n = n_pairs[d]
                st_uv[slot][:, :n].copy_(uv_pin[d], non_blocking=True)
                st_bv[slot][:n].copy_(bv_pin[d], non_blocking=True)
                st_mv[slot][:n].copy_(mv_pin[d], non_blocking=True)
                pipe.build_frame(f, st_uv[slot], st_bv[slot], st_mv[slot], n, ms)
            pipe.forward(mp["bev"], mp["img"], ms, bound)
            pipe.backward(mp["g_bev"], ms, bound)
        last = pipes[(k * units + units - 1) % n_sets]
        torch.cat([last.g_bev.reshape(-1)[:256].view(torch.int32), last.g_img.reshape(-1)[:256].view(torch.int32),
                   last.plan.counts.reshape(-1)[:8]], out=res_dev2[b])
        res_pin2[b].copy_(res_dev2[b], non_blocking=True)

    def e2e_step(k, last=False):
        b = k % 2
        e2e_enqueue(k)
        res_done[b].record(torch.cuda.current_stream())
        for bb in ((1 - b,) if not last else (1 - b, b)):
            res_done[bb].synchronize()
            seen.append(int(res_pin2[bb][513]))

    K_e2e = max(3, min(K, 20))
    for k in range(3):
        e2e_step(k, last=(k == 2))
    torch.cuda.synchronize()
    D.barrier()
    seen.clear()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        e2e_step(k, last=(k == K_e2e - 1))
    torch.cuda.synchronize()
    dt = D.max(time.perf_counter() - t0)
    assert len(seen) == K_e2e + 1 and all(x > 0 for x in seen[1:]), "e2e leg: a result was not read back"
    h2d = F * n_pairs[0] * (16 + 16 + 8)
    e2e = {"value": frames_per_step * K_e2e / dt, "unit": UNIT, "steps": K_e2e, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(n_res * 4), "leg": "c_abi_pipelined",
           "what": "PairsPipeline (ctypes -> C ABI) fed from pinned HOST buffers: per step every frame's img_index rows / bv_index / "
                   "M_val are uploaded, the plans built, the batch pooled forward + backward, and the gradient heads + plan "
                   "counters copied back and read on the host one step later (eager launches, double-buffered staging)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(cfg)
    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / KT, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(name, cfg, world),
            "run": {"timed_blocks": reps, "timed_steps": KT, "timed_ms": ms_total, "min_timed_s": MIN_TIMED_S,
                    "frames_per_step_this_rank": F, "frames_per_launch": Bp, "pairs_per_frame": n_pairs, "nnz_first_unit": nnz,
                    "launch": "CUDA graph replay, one step per graph" if use_graph else "eager launches",
                    "algorithmic_bytes_per_step_per_gpu": bytes_step, "step_gbs_per_gpu": step_gbs, "step_frac_of_peak": step_gbs / peak,
                    "parity_check": parity},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * KT), "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=["1", "2", "2p", "3", "4"], help="BASELINE.json configuration (default 2 = configs[1])")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison before timing")
    ap.add_argument("--no-feeder", action="store_true", help="skip the feeder / velodyne chain legs (config 2)")
    ap.add_argument("--no-no-concat", action="store_true", help="skip the no-concat (sparse-only) leg")
    ap.add_argument("--no-conv", action="store_true", help="skip the fused post-fusion conv leg")
    ap.add_argument("--single-step-graphs", action="store_true", help="one CUDA graph per step (no multi-step graph)")
    ap.add_argument("--graph-steps", type=int, default=8, help="consecutive steps captured in one CUDA graph (multiple of 4)")
    args = ap.parse_args()
    cfg = configs()[args.config]
    if args.impl == "reference":
        return run_reference(args, args.config, cfg)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if cfg["kind"] == "avod":
        return run_gpu_avod(args, args.config, cfg)
    return run_gpu_pairs(args, args.config, cfg)


if __name__ == "__main__":
    sys.exit(main())
