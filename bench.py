#!/usr/bin/env python
"""bench.py -- SHPL forward+backward (+ correspondence build) frames/sec at KITTI shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): avod-FPN pyramid people config with 2 NHSP
layers, forward + backward, batch 1.  One STEP = one synthetic KITTI frame:
  * correspondence build for both layers from the frame's points (shpl_build_avod x2)
  * layer A (after VGG conv4, stride 8, DUAL):  BEV 88x100x256 <-> image 45x150x256
  * layer B (pre-RPN, stride 1, single):        BEV 700x800x32 <-  image 360x1200x32
  * backward of both layers from upstream gradients of the fused maps
`value`  : frames/s with every input already resident in HBM (CUDA-graph replay of the step)
`e2e`    : frames/s from HOST buffers: the frame's points / voxel indices in pinned host memory (H2D inside the
           timed region), a D2H read of every step's result; headline = the C-ABI pipeline (FramePipeline) as a
           deployment runs it, the reference-signature Python API and the unpipelined forms beside it
`roofline`: the dominant kernel (layer B forward) timed with CUDA events, algorithmic bytes
`cpu_baseline`: the CPU oracle (port of the reference's algorithm) on the box's host cores
N > 1: frames are independent -> each rank runs its own frames, no collective on the hot path;
       NCCL is used only to take the max time over ranks.
"""
import argparse
import ctypes
import json
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
WORKLOAD = "avod-FPN pyramid people, 2 NHSP layers (A: stride-8 dual 88x100x256<->45x150x256; B: stride-1 700x800x32<-360x1200x32), fwd+bwd+build, batch 1"
AZ_STEP = 0.028         # azimuth step of the synthetic 64-beam scan: ~20k correspondence pairs per frame
N_FRAMES = 4             # distinct synthetic frames rotated through
N_MAX = 32768


def layer_specs():
    from sparse_pooling_b200.pipeline import LayerSpec
    return [
        LayerSpec("A_vgg_conv4_s8_dual", (88, 100), (45, 150), 256, 256, (8, 8), True, (1200, 360), (704, 800)),
        LayerSpec("B_pre_rpn_s1", (700, 800), (360, 1200), 32, 32, (1, 1), False, (1200, 360), (700, 800)),
    ]


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


# ------------------------------------------------------------------------------- CPU legs
def cpu_frame_inputs(seed):
    from tools import synth
    return synth.avod_frame(seed, az_step_deg=AZ_STEP)


class CpuWorkload:
    """The same step on the host: numpy index oracle + plain-C value oracle (port of the
    reference's algorithm; TensorFlow is not installable, so kind = "port")."""

    def __init__(self):
        from oracle import cref
        cref.build()
        self.specs = layer_specs()
        rng = np.random.default_rng(0)
        self.maps = []
        for s in self.specs:
            bev = rng.standard_normal(s.bev_hw + (s.c_bev,), dtype=np.float32)
            img = rng.standard_normal(s.img_hw + (s.c_img,), dtype=np.float32)
            g_bev = rng.standard_normal(s.bev_hw + (s.c_bev + s.c_img,), dtype=np.float32)
            g_img = rng.standard_normal(s.img_hw + (s.c_img + s.c_bev,), dtype=np.float32) if s.dual else None
            self.maps.append((bev, img, g_bev, g_img))
        self.frames = [cpu_frame_inputs(100 + i) for i in range(N_FRAMES)]
        self.threads = cref.threads()

    def step(self, k):
        from oracle import cref, index_oracle as io
        f = self.frames[k % N_FRAMES]
        out = []
        for s, (bev, img, g_bev, g_img) in zip(self.specs, self.maps):
            d = io.gen_sparse_pooling_input_avod(f["points"], f["voxel_indices"], f["P"], list(s.im_size), s.bv_size)
            o = io.produce_sparse_pooling_input(d, stride=list(s.stride))
            val = np.ones(len(o["Mij_pool"]), np.float32)
            Mij, flip = o["Mij_pool"], o["img_index_flip_pool"]
            fused = cref.forward(bev, img, Mij, val, flip)
            gd, gs = cref.backward(g_bev, Mij, val, flip, s.c_bev, img.shape)
            if s.dual:
                fused_i = cref.forward_trans(img, bev, Mij, val, flip)
                gi, gb = cref.backward_trans(g_img, Mij, val, flip, s.c_img, bev.shape)
                gd += gb
                gs += gi
            out.append(float(fused[0, 0, 0]) + float(gd[0, 0, 0]) + float(gs[0, 0, 0]))
        return out


def run_reference(args):
    """--impl reference: the reference's own CPU path for this workload.  The reference is
    numpy + TensorFlow 1.8 (not installable here), so its algorithm is timed through the
    CPU oracle (numpy builder + plain-C restatement of the TF ops), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = CpuWorkload()
    for k in range(args.warmup):
        wl.step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        wl.step(k)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": wl.threads, "kind": "port",
                         "sample": "%d frames, one per step (numpy correspondence builder + plain-C gather/SpMM/concat "
                                   "and gradients, OpenMP on the dense loops)" % args.steps},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SHPL path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import sparse_pooling_b200 as shpl
    from sparse_pooling_b200 import _cabi
    from sparse_pooling_b200.pipeline import FramePipeline
    from tools import synth

    lib = _cabi.lib
    specs = layer_specs()
    K, W = args.steps, max(args.warmup, 3)

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

    # The two builder streams run at high priority: their kernels are short and latency-bound, and scheduled ahead of
    # the bandwidth-bound pooling CTAs they have the next frame's plans ready earlier (measured: 8454 -> 8600 frames/s;
    # layer B's stream at high priority instead: 7850).  SHPL_BENCH_PRIO: 0 = none, 1 = layer B, 2 = builders (default),
    # 3 = builders + layer A on a high-priority stream of its own
    knob = os.environ.get("SHPL_BENCH_PRIO", "2")
    side = torch.cuda.Stream(device=dev, priority=-1 if knob == "1" else 0)
    side2 = torch.cuda.Stream(device=dev, priority=-1 if knob in ("2", "3") else 0)
    side3 = torch.cuda.Stream(device=dev, priority=-1 if knob in ("2", "3") else 0)
    side_a = torch.cuda.Stream(device=dev, priority=-1) if knob == "3" else None

    def lean_step(k, timing_events=None, overlap=True):
        """build(A,B) + fwd(A,B) + bwd(B,A) on preallocated buffers.  overlap=True puts layer B on a side
        stream so the latency-bound builder kernels overlap the bandwidth-bound pooling kernels;
        overlap=False keeps one stream so that an event pair brackets exactly one kernel."""
        fi, si = k % N_FRAMES, k % n_sets
        pipe, mp = pipes[si], maps[si]
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        if not overlap:
            pipe.build_layer(0, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)
            pipe.build_layer(1, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)
            pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ms, n_pts[fi])
            if timing_events is not None:
                timing_events[0].record(main)
            pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, n_pts[fi])
            if timing_events is not None:
                timing_events[1].record(main)
            pipe.backward_layer(1, mp[1]["g_bev"], None, ms, n_pts[fi])
            if timing_events is not None:
                timing_events[2].record(main)
            pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ms, n_pts[fi])
            return
        # Software-pipelined across frames: while frame k is pooled (plans built during step k-1), the
        # plans of frame k+1 are built into the other buffer set on two more streams.  The builder is a
        # chain of small latency-bound kernels, so it hides under the bandwidth-bound pooling.  Every
        # step still performs one frame's build + forward + backward.
        nf, ns = (k + 1) % N_FRAMES, (k + 1) % n_sets
        for st_ in (side, side2, side3):
            st_.wait_stream(main)
        with torch.cuda.stream(side2):
            pipes[ns].build_layer(0, pts_dev[nf], vox_dev[nf], P, n_pts[nf], side2.cuda_stream)
        with torch.cuda.stream(side3):
            pipes[ns].build_layer(1, pts_dev[nf], vox_dev[nf], P, n_pts[nf], side3.cuda_stream)
        with torch.cuda.stream(side):
            ss = side.cuda_stream
            pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ss, n_pts[fi])
            pipe.backward_layer(1, mp[1]["g_bev"], None, ss, n_pts[fi])
        pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ms, n_pts[fi])
        pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ms, n_pts[fi])
        for st_ in (side, side2, side3):
            main.wait_stream(st_)

    def prologue_build(k):
        """plans of frame k (the first frame of a timed region has no previous step to build them)"""
        fi, si = k % N_FRAMES, k % n_sets
        ms = torch.cuda.current_stream().cuda_stream
        pipes[si].build_layer(0, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)
        pipes[si].build_layer(1, pts_dev[fi], vox_dev[fi], P, n_pts[fi], ms)

    # ---- parity spot check before timing: one lean step against the public API
    prologue_build(0)
    lean_step(0)
    torch.cuda.synchronize()
    nnz = [int(L.plan.counts[0, 3].item()) for L in pipes[0].layers]

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
        lane_a = main if side_a is None else side_a
        for st_ in (side, side2, side3) + (() if side_a is None else (side_a,)):
            st_.wait_stream(main)
        pool_done, build_done = {}, {}
        for j in range(n_steps):
            k = k0 + j
            fi, si = k % N_FRAMES, k % n_sets
            nf, ns = (k + 1) % N_FRAMES, (k + 1) % n_sets
            pipe, mp = pipes[si], maps[si]
            for li, bst in ((0, side2), (1, side3)):           # plans of frame k+1 into buffer set ns
                if (li, k - 1) in pool_done:
                    bst.wait_event(pool_done[(li, k - 1)])      # the pooling of step k-1 read buffer set ns
                with torch.cuda.stream(bst):
                    pipes[ns].build_layer(li, pts_dev[nf], vox_dev[nf], P, n_pts[nf], bst.cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(bst)
                    build_done[(li, k + 1)] = ev
            for li, pst in ((0, lane_a), (1, side)):           # pooling of frame k
                if (li, k) in build_done:
                    pst.wait_event(build_done[(li, k)])
                with torch.cuda.stream(pst):
                    ps = pst.cuda_stream
                    pipe.forward_layer(li, mp[li]["bev"], mp[li]["img"], ps, n_pts[fi])
                    pipe.backward_layer(li, mp[li]["g_bev"], mp[li]["g_img"], ps, n_pts[fi])
                    ev = torch.cuda.Event()
                    ev.record(pst)
                    pool_done[(li, k)] = ev
        for st_ in (side, side2, side3) + (() if side_a is None else (side_a,)):
            main.wait_stream(st_)

    if use_graph and K >= 2 * G and not args.single_step_graphs:
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

    def barrier():
        if world > 1:
            dist.barrier()

    prologue_build(0)
    run_steps(0, W)
    torch.cuda.synchronize()
    # after W steps the plans of frame W are built; the timed region continues the same sequence

    # ---- timed region: EXACTLY K steps, device-timed, barrier + synchronize on both sides
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(lib.shpl_kernel_launches())
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    run_steps(W, W + K)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # launches per step: counted on eager launches (graph replays re-run the captured ones)
    l0 = int(lib.shpl_kernel_launches())
    lean_step(0)
    launches_per_step = int(lib.shpl_kernel_launches()) - l0
    torch.cuda.synchronize()
    gpu_launches = launches_per_step * K if use_graph else int(lib.shpl_kernel_launches()) - launches0 - launches_per_step
    value = world * K / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (layer B forward), CUDA events on its launching stream
    sB = specs[1]
    fwd_ms, bwd_ms = [], []
    roof_how = "CUDA events recorded inside CUDA-graph replays of the single-stream step (no host launch gaps)"
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
        for k in range(K):
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
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        for k in range(K):
            lean_step(k, evs[k], overlap=False)      # the same step, one stream: the event pair brackets one kernel
        torch.cuda.synchronize()
        for e in evs:
            fwd_ms.append(e[0].elapsed_time(e[1]))
            bwd_ms.append(e[1].elapsed_time(e[2]))
    peak, peak_src = peaks()
    bytes_fwd_B = sB.bytes_forward(nnz[1])
    bytes_bwd_B = sB.bytes_backward(nnz[1])
    fwd_avg = float(np.mean(fwd_ms)) * 1e-3
    bwd_avg = float(np.mean(bwd_ms)) * 1e-3
    achieved = bytes_fwd_B / fwd_avg / 1e9
    roofline = {"bound": "hbm", "kernel": "shpl_pool_sparse_kernel<4,false> as layer B forward (700x800x32 <- 360x1200x32)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "bytes_per_launch": bytes_fwd_B, "us_per_launch": fwd_avg * 1e6,
                "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0, "timing": roof_how,
                "backward_kernel": {"achieved": bytes_bwd_B / bwd_avg / 1e9, "frac": bytes_bwd_B / bwd_avg / 1e9 / peak,
                                    "bytes_per_launch": bytes_bwd_B, "us_per_launch": bwd_avg * 1e6}}
    bytes_step = sum(s.bytes_forward(n) + s.bytes_backward(n) for s, n in zip(specs, nnz))
    step_gbs = bytes_step * K / (ms_total * 1e-3) / 1e9

    # ---- e2e: the public drop-in API, frame inputs in pinned host memory, result read back
    class Calib:
        p2 = P
    pts_pin = [torch.from_numpy(f["points"]).pin_memory() for f in frames_host]
    vox_pin = [torch.from_numpy(np.ascontiguousarray(f["voxel_indices"][:, :2])).pin_memory() for f in frames_host]
    result_pin = torch.empty(2 * len(specs) * 256, dtype=torch.float32).pin_memory()
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
        torch.autograd.backward(roots, grads)      # one backward over both layers, like one sess.run(train_op)
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
    dt_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_e2e = float(t.item())
    for mp in maps:
        for m in mp:
            m["bev"].requires_grad_(False)
            m["img"].requires_grad_(False)
    e2e = {"value": world * K_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": K_e2e,
           "what": "gen_sparse_pooling_input_avod + produce_sparse_pooling_input + sparse_pool_layer per layer, then one autograd "
                   "backward over both layers, through the public API; points/voxel indices copied from pinned host memory every step, "
                   "feature maps device-resident (they are device-resident TF tensors in the reference), 2 KB of the "
                   "gradients read back to pinned host memory every step"}

    # ---- e2e through the bare C ABI (the lean ctypes pipeline): frame inputs copied from pinned host memory and a
    #      result read back every step, no cross-frame pipelining (the step ends with a host read)
    stage_pts = torch.empty((N_MAX, 3), dtype=torch.float64, device=dev)
    stage_vox = torch.empty((N_MAX, 2), dtype=torch.int64, device=dev)
    res_dev = torch.empty(2 * len(specs) * 256 + 16, dtype=torch.float32, device=dev)
    res_pin = torch.empty_like(res_dev, device="cpu").pin_memory()

    def cabi_step(k):
        fi, si = k % N_FRAMES, k % n_sets
        pipe, mp = pipes[si], maps[si]
        n = n_pts[fi]
        stage_pts[:n].copy_(pts_pin[fi], non_blocking=True)
        stage_vox[:n].copy_(vox_pin[fi], non_blocking=True)
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        # the two layers are independent: layer A (build, forward, backward) on a second stream beside layer B
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ss = side.cuda_stream
            pipe.build_layer(0, stage_pts, stage_vox, P, n, ss)
            pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ss, n)
            pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ss, n)
        pipe.build_layer(1, stage_pts, stage_vox, P, n, ms)
        pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, n)
        pipe.backward_layer(1, mp[1]["g_bev"], mp[1]["g_img"], ms, n)
        main.wait_stream(side)
        off = 0
        for li in range(len(specs)):
            res_dev[off:off + 256].copy_(pipe.layers[li].g_bev.reshape(-1)[:256])
            res_dev[off + 256:off + 512].copy_(pipe.layers[li].g_img.reshape(-1)[:256])
            off += 512
        res_dev[off:off + 16].copy_(torch.cat([L.plan.counts.reshape(-1)[:8] for L in pipe.layers]).float())
        res_pin.copy_(res_dev, non_blocking=False)

    for k in range(3):
        cabi_step(k)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for k in range(K_e2e):
        cabi_step(k)
    torch.cuda.synchronize()
    dt_cabi = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt_cabi], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_cabi = float(t.item())
    e2e["c_abi_pipeline"] = {"value": world * K_e2e / dt_cabi, "unit": UNIT,
                             "h2d_bytes_per_step": int(n_pts[0] * 40), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                             "what": "same step through the ctypes C-ABI calls on preallocated buffers, the two layers on two streams: points/voxel indices "
                                     "copied from pinned host memory, both plans built, forward+backward of both layers, "
                                     "4 KB of gradients + the plan counters read back to pinned host memory, every step"}

    # ---- the same C-ABI step, software-pipelined by one frame: every step still uploads its inputs from pinned host
    #      memory and its result is still read back inside the timed region, but the host waits for the result of step
    #      k-1 (event) after it has enqueued step k, so enqueueing overlaps the GPU work (double-buffered staging / result
    #      buffers; what a data-loader thread in front of a training loop does)
    stage2 = [(torch.empty((N_MAX, 3), dtype=torch.float64, device=dev), torch.empty((N_MAX, 2), dtype=torch.int64, device=dev))
              for _ in range(2)]
    res_dev2 = [torch.empty(2 * len(specs) * 256 + 16, dtype=torch.int32, device=dev) for _ in range(2)]
    res_pin2 = [torch.empty(2 * len(specs) * 256 + 16, dtype=torch.int32).pin_memory() for _ in range(2)]
    res_done = [torch.cuda.Event() for _ in range(2)]
    # the words read back per buffer set: heads of the four gradients (fp32 bits) and both plans' counters
    res_views = []
    for pipe in pipes:
        v = []
        for L in pipe.layers:
            v += [L.g_bev.reshape(-1)[:256].view(torch.int32), L.g_img.reshape(-1)[:256].view(torch.int32)]
        v += [L.plan.counts.reshape(-1)[:8] for L in pipe.layers]
        res_views.append(v)
    seen = []

    def cabi_enqueue(k):
        fi, si, b = k % N_FRAMES, k % n_sets, k % 2
        pipe, mp = pipes[si], maps[si]
        n = n_pts[fi]
        sp, sv = stage2[b]
        sp[:n].copy_(pts_pin[fi], non_blocking=True)
        sv[:n].copy_(vox_pin[fi], non_blocking=True)
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ss = side.cuda_stream
            pipe.build_layer(0, sp, sv, P, n, ss)
            pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ss, n)
            pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ss, n)
        pipe.build_layer(1, sp, sv, P, n, ms)
        pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, n)
        pipe.backward_layer(1, mp[1]["g_bev"], mp[1]["g_img"], ms, n)
        main.wait_stream(side)
        torch.cat(res_views[si], out=res_dev2[b])
        res_pin2[b].copy_(res_dev2[b], non_blocking=True)

    def read_results(k, last):
        b = k % 2
        res_done[b].record(torch.cuda.current_stream())
        for bb in ((1 - b,) if not last else (1 - b, b)):        # the host read: step k-1's result (and k's at the end)
            res_done[bb].synchronize()
            seen.append(int(res_pin2[bb][-15]))                   # nnz of layer A's plan, out of the words just read

    def cabi_step_pipelined(k, last=False):
        cabi_enqueue(k)
        read_results(k, last)

    K_fast = max(K_e2e, min(K, 400))          # the pipelined legs are cheap: time as many steps as the device-resident leg
    for k in range(3):
        cabi_step_pipelined(k, last=(k == 2))
    torch.cuda.synchronize()
    barrier()
    seen.clear()
    t0 = time.perf_counter()
    for k in range(K_fast):
        cabi_step_pipelined(k, last=(k == K_fast - 1))
    torch.cuda.synchronize()
    dt_pipe = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt_pipe], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_pipe = float(t.item())
    assert len(seen) == K_fast + 1 and all(x > 0 for x in seen[1:]), "pipelined C-ABI leg: a result was not read back"
    e2e["c_abi_pipelined"] = {"value": world * K_fast / dt_pipe, "unit": UNIT, "steps": K_fast,
                              "h2d_bytes_per_step": int(n_pts[0] * 40), "d2h_bytes_per_step": int(res_pin2[0].numel() * 4),
                              "what": "the C-ABI step software-pipelined by one frame: inputs uploaded from pinned host memory and "
                                      "4 KB of gradients + plan counters read back for EVERY step inside the timed region, the host "
                                      "waiting for step k-1's result after enqueueing step k (double-buffered staging)"}

    # ---- and with the enqueueing itself captured: one CUDA graph per (frame, buffer set) holding the two uploads from
    #      the pinned host buffers, both builds, forward + backward of both layers and the D2H of the result; a step is one
    #      graph launch + the wait for the previous step's result
    try:
        n_combo = max(N_FRAMES, n_sets, 2)          # smallest k-period of (frame, buffer set, staging buffer)
        while n_combo % N_FRAMES or n_combo % n_sets or n_combo % 2:
            n_combo += 1
        e2e_graphs = []
        for k in range(n_combo):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                cabi_enqueue(k)
            e2e_graphs.append(gr)

        lanes = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

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
        dt_g = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_g], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_g = float(t.item())
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
                   "voxel indices are uploaded, both plans built, both layers run forward + backward and 4 KB of gradients + the "
                   "plan counters copied back and read on the host -- " + e2e[head]["what"] + ".  Other legs: python_dropin_api "
                   "(the reference-signature functions, one frame at a time, synchronising like the reference), c_abi_pipeline "
                   "(no pipelining: every step ends with its own host read), c_abi_pipelined (eager launches)")

    # ---- the feeder in front of the path: BevSlices.generate_bev of the raw scan on the
    #      GPU (shpl_bev_slices), its pair count handed to the builders on the device (no host read in between)
    feeder = None
    try:
        from sparse_pooling_b200 import bev_slices as bs
        from oracle import feeder_oracle as fo
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
        torch.cuda.synchronize()
        evf0.record()
        for k in range(K):
            fg[k % N_FRAMES].replay()
        evf1.record()
        torch.cuda.synchronize()
        feeder_us = evf0.elapsed_time(evf1) * 1e3 / K

        def scan_step(k):
            fi, si = k % N_FRAMES, k % n_sets
            pipe, mp = pipes[si], maps[si]
            Pn = scans[fi].shape[1]
            stage_scan[:, :Pn].copy_(scan_pin[fi], non_blocking=True)
            main = torch.cuda.current_stream()
            ms = main.cuda_stream
            feeder_call(fi)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ss = side.cuda_stream
                pipe.build_layer(0, work.unique_pts, work.voxel_indices, P, N_MAX, ss, n_dev=n_dev)
                pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ss, N_MAX)
                pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ss, N_MAX)
            pipe.build_layer(1, work.unique_pts, work.voxel_indices, P, N_MAX, ms, n_dev=n_dev)
            pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, N_MAX)
            pipe.backward_layer(1, mp[1]["g_bev"], mp[1]["g_img"], ms, N_MAX)
            main.wait_stream(side)
            off = 0
            for li in range(len(specs)):
                res_dev[off:off + 256].copy_(pipe.layers[li].g_bev.reshape(-1)[:256])
                res_dev[off + 256:off + 512].copy_(pipe.layers[li].g_img.reshape(-1)[:256])
                off += 512
            res_dev[off:off + 16].copy_(torch.cat([L.plan.counts.reshape(-1)[:8] for L in pipe.layers]).float())
            res_pin.copy_(res_dev, non_blocking=False)

        for k in range(3):
            scan_step(k)
        torch.cuda.synchronize()
        ref_counts = [int(x) for x in res_pin[-16:].tolist()]
        barrier()
        t0 = time.perf_counter()
        for k in range(K_e2e):
            scan_step(k)
        torch.cuda.synchronize()
        dt_scan = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_scan], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_scan = float(t.item())
        t0 = time.perf_counter()
        for fi in range(N_FRAMES):
            fo.generate_bev(scans[fi], GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL, -0.2, 2.3, 5)
        cpu_ms = (time.perf_counter() - t0) * 1e3 / N_FRAMES

        # ---- the whole widened chain from a RAW velodyne scan: float32 [N,4] in pinned host
        #      memory -> H2D -> ingest (camera frame + FOV filter) -> feeder -> both plans -> forward -> backward ->
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
            pipe, mp = pipes[si], maps[si]
            nv = velos[fi].shape[0]
            stage_velo[:nv].copy_(velo_pin[fi], non_blocking=True)
            main = torch.cuda.current_stream()
            ms = main.cuda_stream
            li.lidar_to_cam_raw(stage_velo, nv, cal, [1242, 375], cam_buf, ing_counts)
            bs.bev_slices_raw(cam_buf, cam_buf.stride(0), cam_buf.stride(1), nv, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                              -0.2, 2.3, 5, np.log(16), work, lut=lut, p_dev=p_dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ss = side.cuda_stream
                pipe.build_layer(0, work.unique_pts, work.voxel_indices, P, N_MAX, ss, n_dev=n_dev)
                pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ss, N_MAX)
                pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ss, N_MAX)
            pipe.build_layer(1, work.unique_pts, work.voxel_indices, P, N_MAX, ms, n_dev=n_dev)
            pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, N_MAX)
            pipe.backward_layer(1, mp[1]["g_bev"], mp[1]["g_img"], ms, N_MAX)
            main.wait_stream(side)
            off = 0
            for li_ in range(len(specs)):
                res_dev[off:off + 256].copy_(pipe.layers[li_].g_bev.reshape(-1)[:256])
                res_dev[off + 256:off + 512].copy_(pipe.layers[li_].g_img.reshape(-1)[:256])
                off += 512
            res_dev[off:off + 16].copy_(torch.cat([L.plan.counts.reshape(-1)[:8] for L in pipe.layers]).float())
            res_pin.copy_(res_dev, non_blocking=False)

        for k in range(3):
            velo_step(k)
        torch.cuda.synchronize()
        velo_counts = [int(x) for x in res_pin[-16:].tolist()][:4] + [int(ing_counts[0].item()), int(work.counts[0].item())]
        barrier()
        t0 = time.perf_counter()
        for k in range(K_e2e):
            velo_step(k)
        torch.cuda.synchronize()
        dt_velo = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_velo], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_velo = float(t.item())
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
                pipe, mp, vs = pipes[si], maps[si], vsets[b]
                nv = velos[fi].shape[0]
                vs["stage"][:nv].copy_(velo_pin[fi], non_blocking=True)
                main = torch.cuda.current_stream()
                ms = main.cuda_stream
                li.lidar_to_cam_raw(vs["stage"], nv, cal, [1242, 375], vs["cam"], vs["cnt"], ws=vs["ws"])
                bs.bev_slices_raw(vs["cam"], vs["cam"].stride(0), vs["cam"].stride(1), nv, GP, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                  -0.2, 2.3, 5, np.log(16), vs["work"], lut=lut, p_dev=ctypes.c_void_p(vs["cnt"].data_ptr()))
                nd = ctypes.c_void_p(vs["work"].counts.data_ptr())
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ss = side.cuda_stream
                    pipe.build_layer(0, vs["work"].unique_pts, vs["work"].voxel_indices, P, N_MAX, ss, n_dev=nd)
                    pipe.forward_layer(0, mp[0]["bev"], mp[0]["img"], ss, N_MAX)
                    pipe.backward_layer(0, mp[0]["g_bev"], mp[0]["g_img"], ss, N_MAX)
                pipe.build_layer(1, vs["work"].unique_pts, vs["work"].voxel_indices, P, N_MAX, ms, n_dev=nd)
                pipe.forward_layer(1, mp[1]["bev"], mp[1]["img"], ms, N_MAX)
                pipe.backward_layer(1, mp[1]["g_bev"], mp[1]["g_img"], ms, N_MAX)
                main.wait_stream(side)
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
            vg_counts = [int(x) for x in res_pin2[0][-16:].tolist()][:4]        # step k = 2: the frame the eager check read
            barrier()
            seen.clear()
            t0 = time.perf_counter()
            for k in range(K_fast):
                velo_step_graph(k, last=(k == K_fast - 1))
            torch.cuda.synchronize()
            dt_vg = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt_vg], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt_vg = float(t.item())
            assert len(seen) == K_fast + 1 and all(x > 0 for x in seen[1:]), "velodyne graph leg: a result was not read back"
            assert vg_counts == velo_counts[:4], "velodyne graph leg: plan counters differ from the eager chain's"
            velo_graph = {"value": world * K_fast / dt_vg, "unit": UNIT, "steps": K_fast,
                          "h2d_bytes_per_step": int(velos[0].shape[0] * 16), "d2h_bytes_per_step": int(res_pin2[0].numel() * 4),
                          "what": "the velodyne chain captured in CUDA graphs (upload of the raw scan, ingest, feeder, both plans, "
                                  "forward + backward, D2H), one graph launch per step on alternating streams with every buffer "
                                  "doubled, every step's result read on the host one step later"}
        except Exception as ex:  # pragma: no cover
            print("velodyne graph leg failed: %r" % (ex,), file=sys.stderr)
            torch.cuda.synchronize()

        feeder = {"what": "BevSlices.generate_bev(output_indices=True) on the GPU: 5 height maps + density map [6,700,800] f64, "
                          "voxel_indices, unique_pts (shpl_bev_slices, CUDA-graph replays, CUDA events)",
                  "us_per_frame": feeder_us, "points_per_scan": [int(sc.shape[1]) for sc in scans],
                  "cpu_oracle_ms_per_frame": cpu_ms,
                  "e2e_from_scan": {"value": world * K_e2e / dt_scan, "unit": UNIT,
                                    "h2d_bytes_per_step": int(scans[0].shape[1] * 24), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                                    "what": "raw scan [3,P] f64 copied from pinned host memory, feeder, both plans built from the "
                                            "feeder's device-side pair count, forward+backward of both layers, gradients + plan "
                                            "counters read back, every step (ctypes C-ABI calls)"},
                  "plan_counts_check": ref_counts[:4],
                  "e2e_from_velodyne": {"value": world * K_e2e / dt_velo, "unit": UNIT,
                                        "h2d_bytes_per_step": int(velos[0].shape[0] * 16), "d2h_bytes_per_step": int(res_pin.numel() * 4),
                                        "points_per_scan": [int(v.shape[0]) for v in velos],
                                        "what": "raw 360-degree velodyne scan float32 [N,4] from pinned host memory, ingest (camera "
                                                "frame, FOV filter), feeder, both plans, forward+backward of both layers, read-back; "
                                                "all intermediate counts stay on the device",
                                        "counts_check(nclip,nnz,oob,csr,fov_points,pairs)": velo_counts},
                  "e2e_from_velodyne_graph": velo_graph}
    except Exception as ex:  # pragma: no cover
        print("feeder leg failed: %r" % (ex,), file=sys.stderr)
        torch.cuda.synchronize()

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        wl = CpuWorkload()
        wl.step(0)
        n_cpu, t0 = 0, time.perf_counter()
        while n_cpu < 3 or (time.perf_counter() - t0 < 10.0 and n_cpu < 40):
            wl.step(n_cpu)
            n_cpu += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": wl.threads, "kind": "port",
               "sample": "%d frames of the same workload (numpy correspondence builder + plain-C oracle of the TF ops)" % n_cpu}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "candidate_pairs_per_frame": n_pts,
                       "nnz_per_layer": nnz, "sharding": "frames by rank, no data-path collective",
                       "l2": "inputs larger than L2: %.0f MB touched per step, %d rotating buffer sets" % (bytes_step / 1e6, n_sets),
                       "launch": ("CUDA graph replay, %d steps per graph where the step index lines up" % G if multi is not None
                                  else "CUDA graph replay, one step per graph") if use_graph else "eager launches",
                       "algorithmic_bytes_per_step": bytes_step, "step_gbs_per_gpu": step_gbs,
                       "step_frac_of_peak": step_gbs / peak},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline,
        }
        if feeder is not None:
            line["feeder"] = feeder
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-step-graphs", action="store_true", help="one CUDA graph per step (no multi-step graph)")
    ap.add_argument("--graph-steps", type=int, default=8, help="consecutive steps captured in one CUDA graph (multiple of 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
