"""Memory-safety run (SURVEY.md section 5 "race detection"; compute-sanitizer is closed on this pool).

Runs the builder, the pooling kernels (sparse / packed / narrow / wide regimes, both no-concat forms), the heavy-cell
kernels (exact cluster kernel + split tree) and the fused conv through libshpl_debug.so -- the build with in-kernel
index checks (make debug) -- on CANARY-PADDED buffers: every array handed to the library sits between two guard
regions filled with a pattern; after the kernels the guards must be intact (no out-of-bounds write) and the library's
check counter must be zero (no out-of-range gather index, entry range or output slot).

    SHPL_LIB=sparse_pooling_b200/libshpl_debug.so python tools/memsafety_run.py      (tests/test_gpu_memsafety.py does this)
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SHPL_LIB", os.path.join(ROOT, "sparse_pooling_b200", "libshpl_debug.so"))
from sparse_pooling_b200 import _cabi  # noqa: E402
from tools import synth  # noqa: E402

lib = _cabi.lib
GUARD = 4096          # bytes on either side
PATTERN = 0xA5


class Guarded:
    """a device array of n elements between two guard regions"""

    def __init__(self, n, dtype, device, fill=None):
        self.esize = torch.empty((), dtype=dtype).element_size()
        nbytes = (int(n) * self.esize + 255) // 256 * 256
        self.raw = torch.full((GUARD + nbytes + GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.nbytes_used = int(n) * self.esize
        self.view = self.raw[GUARD:GUARD + int(n) * self.esize].view(dtype)
        if fill is not None:
            self.view.copy_(fill.reshape(-1).to(device=device, dtype=dtype))

    def ptr(self, offset_elems=0):
        return ctypes.c_void_p(self.view.data_ptr() + offset_elems * self.esize)

    def intact(self):
        head = self.raw[:GUARD]
        tail = self.raw[GUARD + self.nbytes_used:]          # includes the alignment slack
        return bool((head == PATTERN).all().item()) and bool((tail == PATTERN).all().item())


ALL = []


def G(n, dtype, dev, fill=None):
    g = Guarded(n, dtype, dev, fill)
    ALL.append(g)
    return g


def check(rc, what):
    _cabi.check(rc, what)


def guarded_plan(R, Q, cap, dev, heavy=True):
    i32, f32 = torch.int32, torch.float32
    arr = dict(row_ptr=G(R + 1, i32, dev), csr_row=G(cap, i32, dev), csr_src=G(cap, i32, dev), csr_val=G(cap, f32, dev),
               pix_ptr=G(Q + 1, i32, dev), csrT_pix=G(cap, i32, dev), csrT_dst=G(cap, i32, dev), csrT_val=G(cap, f32, dev),
               counts=G(8, i32, dev, torch.zeros(8, dtype=i32)), heavy_count=G(2, i32, dev, torch.zeros(2, dtype=i32)))
    hc = cap // _cabi.HEAVY_LEN + 1 if heavy else 0
    arr["heavy_row"], arr["heavy_pix"] = G(max(hc, 1), i32, dev), G(max(hc, 1), i32, dev)
    st = _cabi.ShplPlan()
    st.n_rows, st.n_src, st.capacity, st.heavy_cap = R, Q, cap, hc
    for k in ("row_ptr", "csr_row", "csr_src", "csr_val", "pix_ptr", "csrT_pix", "csrT_dst", "csrT_val", "heavy_row", "heavy_pix",
              "heavy_count", "counts"):
        setattr(st, k, arr[k].ptr().value)
    return st, arr, hc


def run_case(name, bev_hw, img_hw, C_b, C_i, n, skew, dev, stride=(1, 1), dual=False, one_cell=0):
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    R, Q = (bev_hw[0] // stride[1]) * (bev_hw[1] // stride[1]), (img_hw[0] // stride[0]) * (img_hw[1] // stride[0])
    d = synth.direct_pairs(7, n, bev_hw, (img_hw[1], img_hw[0]), skew=skew)
    if one_cell:                       # a listed (heavy) cell: the first `one_cell` pairs land in one BEV cell
        d["bv_index"][:one_cell] = (3, 2)
    u = G(n, torch.float64, dev, torch.from_numpy(d["img_index"][0]))
    v = G(n, torch.float64, dev, torch.from_numpy(d["img_index"][1]))
    bv = G(2 * n, torch.int64, dev, torch.from_numpy(d["bv_index"]))
    mval = G(n, torch.float64, dev, torch.rand(n, dtype=torch.float64) + 0.5)
    st, arr, hc = guarded_plan(R, Q, n, dev)
    ws_bytes = int(lib.shpl_build_workspace_bytes(n))
    ws = G(ws_bytes, torch.uint8, dev)
    mij, flip, mv32, msz = G(2 * n, torch.int64, dev), G(3 * n, torch.int64, dev), G(n, torch.float32, dev), G(2, torch.int64, dev)
    check(lib.shpl_produce_input(u.ptr(), v.ptr(), bv.ptr(), n, img_hw[1], img_hw[0], bev_hw[0], bev_hw[1], stride[0], stride[1],
                                 mval.ptr(), 0, 0, mij.ptr(), flip.ptr(), mv32.ptr(), msz.ptr(), ctypes.byref(st), 0, 0, None,
                                 ws.ptr(), ws_bytes, stream), "shpl_produce_input")
    bev = G(R * C_b, torch.float32, dev, torch.randn(R * C_b))
    img = G(Q * C_i, torch.float32, dev, torch.randn(Q * C_i))
    fused = G(R * (C_b + C_i), torch.float32, dev)
    g_fused = G(R * (C_b + C_i), torch.float32, dev, torch.randn(R * (C_b + C_i)))
    g_bev, g_img = G(R * C_b, torch.float32, dev), G(Q * C_i, torch.float32, dev)
    P = [arr[k].ptr() for k in ("row_ptr", "csr_row", "csr_src", "csr_val", "pix_ptr", "csrT_pix", "csrT_dst", "csrT_val")]
    H = _cabi.HEAVY_LEN
    check(lib.shpl_pool_forward(bev.ptr(), img.ptr(), P[0], P[1], P[2], P[3], n, H, R, C_b, Q, C_i, fused.ptr(), stream), "forward")
    hws_b = int(lib.shpl_pool_heavy_workspace_bytes(max(C_b, C_i), n, hc))
    hws = G(hws_b, torch.uint8, dev)
    if hc:
        check(lib.shpl_pool_heavy_split(img.ptr(), C_i, C_i, P[0], P[2], P[3], arr["heavy_row"].ptr(), arr["heavy_count"].ptr(), hc,
                                        None, 0, fused.ptr(C_b), C_b + C_i, n, hws.ptr(), hws_b, stream), "heavy_split")
        check(lib.shpl_pool_heavy(img.ptr(), C_i, C_i, P[0], P[2], P[3], arr["heavy_row"].ptr(), arr["heavy_count"].ptr(), hc,
                                  None, 0, fused.ptr(C_b), C_b + C_i, stream), "heavy")
    check(lib.shpl_pool_backward(g_fused.ptr(), P[4], P[5], P[6], P[7], n, H, R, C_b, Q, C_i, g_bev.ptr(), g_img.ptr(), stream), "backward")
    if hc:
        check(lib.shpl_pool_heavy_split(g_fused.ptr(C_b), C_b + C_i, C_i, P[4], P[6], P[7], arr["heavy_pix"].ptr(),
                                        ctypes.c_void_p(arr["heavy_count"].ptr().value + 4), hc, None, 0, g_img.ptr(), C_i, n,
                                        hws.ptr(), hws_b, stream), "heavy_split (backward)")
    # no-concat forms
    check(lib.shpl_pool_forward_into(img.ptr(), P[0], P[1], P[2], P[3], n, 0, R, Q, C_i, fused.ptr(), C_b + C_i, C_b, stream), "forward_into")
    check(lib.shpl_pool_backward_from(g_fused.ptr(), C_b + C_i, C_b, P[4], P[5], P[6], P[7], n, 0, R, Q, C_i, g_img.ptr(), stream), "backward_from")
    if dual:
        fused_i = G(Q * (C_i + C_b), torch.float32, dev)
        g_fi = G(Q * (C_i + C_b), torch.float32, dev, torch.randn(Q * (C_i + C_b)))
        check(lib.shpl_pool_forward_dual(bev.ptr(), img.ptr(), *P, n, 0, R, C_b, Q, C_i, fused.ptr(), fused_i.ptr(), stream), "forward_dual")
        check(lib.shpl_pool_backward_dual(g_fused.ptr(), g_fi.ptr(), *P, n, 0, R, C_b, Q, C_i, g_bev.ptr(), g_img.ptr(), stream), "backward_dual")
        check(lib.shpl_pool_forward_into_dual(bev.ptr(), img.ptr(), *P, n, 0, R, C_b, Q, C_i, fused.ptr(), fused_i.ptr(), stream), "forward_into_dual")
    if C_b == 32 and C_i == 32 and stride == (1, 1):
        Hh, Ww = bev_hw
        cws_b = int(lib.shpl_conv3x3_workspace_bytes(1, Hh, Ww, n))
        cws = G(cws_b + 256, torch.uint8, dev)
        off = (-cws.view.data_ptr()) % 256
        w = G(9 * 64 * 32, torch.float32, dev, torch.randn(9 * 64 * 32) * 0.1)
        out = G(R * 32, torch.float32, dev)
        check(lib.shpl_pool_conv3x3_forward(bev.ptr(), img.ptr(), P[0], P[1], P[2], P[3], n, 1, Hh, Ww, 32, Q, 32, w.ptr(), 32, None, None, 1,
                                            out.ptr(), ctypes.c_void_p(cws.view.data_ptr() + off), cws_b, stream), "conv3x3")
        bws_b = int(lib.shpl_conv3x3_backward_workspace_bytes(n))
        bws = G(bws_b + 256, torch.uint8, dev)
        boff = (-bws.view.data_ptr()) % 256
        g_out = G(R * 32, torch.float32, dev, torch.randn(R * 32))
        gc_bev, gc_img, gc_w = G(R * 32, torch.float32, dev), G(Q * 32, torch.float32, dev), G(9 * 64 * 32, torch.float32, dev)
        check(lib.shpl_pool_conv3x3_backward(g_out.ptr(), bev.ptr(), img.ptr(), *P, n, 1, Hh, Ww, 32, Q, 32, w.ptr(), 32,
                                             gc_bev.ptr(), gc_img.ptr(), gc_w.ptr(), ctypes.c_void_p(bws.view.data_ptr() + boff), bws_b, stream),
              "conv3x3_backward")
    torch.cuda.synchronize()
    print("case %-28s ran" % name, flush=True)


def main():
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    assert lib.shpl_debug_checks_enabled() == 1, "memsafety_run needs libshpl_debug.so (make -C sparse_pooling_b200/csrc debug; SHPL_LIB=...)"
    run_case("kitti-like sparse C=32 + conv", (96, 112), (45, 150), 32, 32, 3000, "uniform", dev)
    run_case("ragged conv size", (50, 44), (20, 30), 32, 32, 700, "ground", dev)
    run_case("dense regime packed C=16", (16, 16), (32, 64), 16, 16, 6000, "uniform", dev)
    run_case("odd widths narrow C=12/20", (30, 20), (25, 35), 12, 20, 2500, "zipf", dev)
    run_case("wide dual C=256 stride 8", (88, 104), (48, 152), 256, 256, 4000, "ground", dev, stride=(8, 8), dual=True)
    run_case("heavy exact cell (1500)", (16, 16), (32, 64), 32, 32, 4000, "uniform", dev, one_cell=1500)
    run_case("heavy split cell (9000) C=64", (16, 16), (32, 64), 64, 64, 12000, "uniform", dev, one_cell=9000)
    run_case("heavy split cell (9000) C=8", (16, 16), (32, 64), 8, 8, 12000, "uniform", dev, one_cell=9000)
    torch.cuda.synchronize()
    failures = int(lib.shpl_debug_check_failures())
    broken = [i for i, g in enumerate(ALL) if not g.intact()]
    print("buffers: %d, guard regions overwritten: %d, in-kernel check failures: %d" % (len(ALL), len(broken), failures))
    ok = failures == 0 and not broken
    print("memsafety:", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
