"""ctypes wrapper around oracle/_build/libshpl_oracle.so (the plain-C oracle).

TEST INFRASTRUCTURE (see oracle/__init__.py): imported only by tests/, smoke()
and bench.py's CPU-baseline legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libshpl_oracle.so")
_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64


def build(force=False):
    src = os.path.join(_HERE, "shpl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.shpl_oracle_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def threads():
    return int(lib().shpl_oracle_threads())


def build_avod(points, vox, P, im_size, bv_size, stride, want_gen=True):
    """C restatement of gen_sparse_pooling_input_avod + produce_sparse_pooling_input."""
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    vox = np.ascontiguousarray(np.asarray(vox)[:, :2], dtype=np.int64)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
    N = points.shape[0]
    counts = np.zeros(2, dtype=np.int64)
    gen_bv = np.zeros((N, 2), dtype=np.int64) if want_gen else None
    gu = np.zeros(N, dtype=np.float64) if want_gen else None
    gv = np.zeros(N, dtype=np.float64) if want_gen else None
    mij = np.zeros((N, 2), dtype=np.int64)
    flip = np.zeros((N, 3), dtype=np.int64)
    msize = np.zeros(2, dtype=np.int64)
    lib().shpl_oracle_build_avod(_p(points), _p(vox), _i64(N), _p(P), _i64(int(im_size[0])), _i64(int(im_size[1])),
                                 _i64(int(bv_size[0])), _i64(int(bv_size[1])), _i64(int(stride[0])), _i64(int(stride[1])),
                                 _p(counts), _p(gen_bv), _p(gu), _p(gv), _p(mij), _p(flip), _p(msize))
    n, nnz = int(counts[0]), int(counts[1])
    out = {"Mij_pool": mij[:nnz].copy(), "M_size": msize, "img_index_flip_pool": flip[:nnz].copy(),
           "M_val": np.ones(nnz), "n": n, "nnz": nnz}
    if want_gen:
        img_index = np.zeros((3, n), dtype=np.float64)
        img_index[0] = gu[:n]
        img_index[1] = gv[:n]
        out["gen"] = {"bv_index": gen_bv[:n].copy(), "img_index": img_index}
    return out


def _coo(Mij, val, flip, src_w):
    Mij = np.ascontiguousarray(Mij, dtype=np.int64).reshape(-1, 2)
    rows = np.ascontiguousarray(Mij[:, 0])
    cols = np.ascontiguousarray(Mij[:, 1])
    val = np.ascontiguousarray(val, dtype=np.float32)
    flip = np.asarray(flip, dtype=np.int64).reshape(-1, 3)
    u, v = flip[:, 2], flip[:, 1]
    pix = np.where((u >= 0) & (u < src_w) & (v >= 0) & (flip[:, 0] == 0), v * src_w + u, -1)
    return rows, cols, val, np.ascontiguousarray(pix, dtype=np.int64)


def forward(dst, src, Mij, val, flip, scratch=None):
    """dst [Hd,Wd,Cd] f32, src [Hs,Ws,Cs] f32 -> fused [Hd,Wd,Cd+Cs] (img->bev direction)."""
    Hd, Wd, Cd = dst.shape
    Hs, Ws, Cs = src.shape
    R, Q = Hd * Wd, Hs * Ws
    rows, cols, val, pix = _coo(Mij, val, flip, Ws)
    ncols = pix.shape[0]
    if scratch is None:
        scratch = (np.empty(max(ncols, 1) * Cs, np.float32), np.empty(R * Cs, np.float32))
    fused = np.empty((Hd, Wd, Cd + Cs), dtype=np.float32)
    lib().shpl_oracle_forward(_p(dst), _p(src), _i64(rows.shape[0]), _p(rows), _p(cols), _p(val), _i64(ncols), _p(pix),
                              _i64(R), _i64(Cd), _i64(Q), _i64(Cs), _p(scratch[0]), _p(scratch[1]), _p(fused))
    return fused


def backward(g_fused, Mij, val, flip, Cd, src_shape, scratch=None):
    Hd, Wd, F = g_fused.shape
    Hs, Ws, Cs = src_shape
    assert F == Cd + Cs
    R, Q = Hd * Wd, Hs * Ws
    rows, cols, val, pix = _coo(Mij, val, flip, Ws)
    ncols = pix.shape[0]
    if scratch is None:
        scratch = np.empty(max(ncols, 1) * Cs, np.float32)
    g_dst = np.empty((Hd, Wd, Cd), dtype=np.float32)
    g_src = np.empty((Hs, Ws, Cs), dtype=np.float32)
    lib().shpl_oracle_backward(_p(g_fused), _i64(rows.shape[0]), _p(rows), _p(cols), _p(val), _i64(ncols), _p(pix),
                               _i64(R), _i64(Cd), _i64(Q), _i64(Cs), _p(scratch), _p(g_dst), _p(g_src))
    return g_dst, g_src


def forward_trans(img, bev, Mij, val, flip):
    """img [Hi,Wi,Ci], bev [Hb,Wb,Cb] -> fused_i [Hi,Wi,Ci+Cb] (bev->img direction)."""
    Hi, Wi, Ci = img.shape
    Hb, Wb, Cb = bev.shape
    R, Q = Hb * Wb, Hi * Wi
    rows, cols, val, pix = _coo(Mij, val, flip, Wi)
    ncols = pix.shape[0]
    S = np.empty(max(ncols, 1) * Cb, np.float32)
    Pm = np.empty(Q * Cb, np.float32)
    fused = np.empty((Hi, Wi, Ci + Cb), dtype=np.float32)
    lib().shpl_oracle_forward_trans(_p(img), _p(bev), _i64(rows.shape[0]), _p(rows), _p(cols), _p(val), _i64(ncols),
                                    _p(pix), _i64(R), _i64(Cb), _i64(Q), _i64(Ci), _p(S), _p(Pm), _p(fused))
    return fused


def backward_trans(g_fused_i, Mij, val, flip, Ci, bev_shape):
    Hi, Wi, F = g_fused_i.shape
    Hb, Wb, Cb = bev_shape
    assert F == Ci + Cb
    R, Q = Hb * Wb, Hi * Wi
    rows, cols, val, pix = _coo(Mij, val, flip, Wi)
    g_img = np.empty((Hi, Wi, Ci), dtype=np.float32)
    g_bev = np.empty((Hb, Wb, Cb), dtype=np.float32)
    lib().shpl_oracle_backward_trans(_p(g_fused_i), _i64(rows.shape[0]), _p(rows), _p(cols), _p(val),
                                     _i64(pix.shape[0]), _p(pix), _i64(R), _i64(Cb), _i64(Q), _i64(Ci),
                                     _p(g_img), _p(g_bev))
    return g_img, g_bev
