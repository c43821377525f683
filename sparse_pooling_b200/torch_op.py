"""The SHPL pooling kernels as registered PyTorch custom ops (``torch.ops.shpl.pool`` / ``torch.ops.shpl.pool_backward``)
with autograd and fake-tensor (meta) support, next to the ``torch.autograd.Function`` wrappers of ``ops.py``.

``ops.SparsePoolFunction`` stays the path ``sparse_pool_layer`` takes (least host overhead); the registered ops are for
callers that want ``torch.ops`` semantics -- ``torch.compile`` / ``torch.export`` tracing through the layer, ``opcheck``,
serialisable graphs.  Both call the same entry points of libshpl.so (``shpl_pool_forward`` / ``shpl_pool_backward``);
cells are always summed sequentially here (``heavy_len = 0``), so the result is bit-identical to the oracle for every
row length.

One direction of the layer (sparse_pool_utils.py:96-103 with :72, or :105-117 with :87 on the transposed arrays):
    fused[r] = concat(dst[r], sum_{k in cell r} val_k * src[idx_k])
"""
from typing import Optional, Tuple

import torch

from . import _cabi
from .ops import _ptr, _stream

_lib = _cabi.lib


def _check(t, name, dtype):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the SHPL kernels run on the GPU only (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _check_csr(ptr, key, idx, val, nnz_max, device, what):
    """The index arrays are handed to the kernels as raw addresses: wrong dtype / device / strides / lengths would be
    read out of bounds instead of raising, so they are checked here."""
    for t, name, dtype in ((ptr, "ptr", torch.int32), (key, "key", torch.int32), (idx, "idx", torch.int32), (val, "val", torch.float32)):
        _check(t, "%s %s" % (what, name), dtype)
        if t.device != device:
            raise ValueError("%s %s is on %s, the feature maps on %s" % (what, name, t.device, device))
        if t.dim() != 1 or not t.is_contiguous():
            raise ValueError("%s %s must be a contiguous 1-D tensor" % (what, name))
    if ptr.shape[0] < 1:
        raise ValueError("%s ptr must hold n_rows + 1 offsets" % what)
    if nnz_max < 0 or min(key.shape[0], idx.shape[0], val.shape[0]) < nnz_max:
        raise ValueError("%s key / idx / val hold %d / %d / %d entries, nnz_max = %d"
                         % (what, key.shape[0], idx.shape[0], val.shape[0], nnz_max))


@torch.library.custom_op("shpl::pool", mutates_args=(), device_types="cuda")
def pool(dst: Optional[torch.Tensor], src: torch.Tensor, ptr: torch.Tensor, key: torch.Tensor, idx: torch.Tensor,
         val: torch.Tensor, ptrT: torch.Tensor, keyT: torch.Tensor, idxT: torch.Tensor, valT: torch.Tensor,
         nnz_max: int) -> torch.Tensor:
    """dst [n_rows, C_d] or None, src [n_src, C_s] -> fused [n_rows, C_d + C_s].  (ptr, key, idx, val): CSR by
    destination cell; (ptrT, keyT, idxT, valT): CSR by source cell, carried for the backward."""
    _check(src, "src", torch.float32)
    if src.dim() != 2:
        raise ValueError("src must be [n_src, C_s]")
    _check_csr(ptr, key, idx, val, nnz_max, src.device, "forward CSR")
    _check_csr(ptrT, keyT, idxT, valT, nnz_max, src.device, "transposed CSR")
    n_src, C_s = src.shape
    n_rows = ptr.shape[0] - 1
    if ptrT.shape[0] - 1 != n_src:
        raise ValueError("transposed CSR has %d rows, src %d" % (ptrT.shape[0] - 1, n_src))
    if dst is not None:
        _check(dst, "dst", torch.float32)
        if dst.dim() != 2 or dst.shape[0] != n_rows or dst.device != src.device:
            raise ValueError("dst must be [n_rows = %d, C_d] on %s, got %s on %s" % (n_rows, src.device, tuple(dst.shape), dst.device))
    C_d = 0 if dst is None else dst.shape[1]
    s = src.contiguous()
    d = None if dst is None else dst.contiguous()
    fused = torch.empty((n_rows, C_d + C_s), dtype=torch.float32, device=src.device)
    rc = _lib.shpl_pool_forward(_ptr(d), _ptr(s), _ptr(ptr), _ptr(key), _ptr(idx), _ptr(val), int(nnz_max), 0, n_rows, C_d,
                                n_src, C_s, _ptr(fused), _stream())
    _cabi.check(rc, "shpl_pool_forward")
    return fused


@pool.register_fake
def _(dst, src, ptr, key, idx, val, ptrT, keyT, idxT, valT, nnz_max):
    C_d = 0 if dst is None else dst.shape[1]
    return src.new_empty((ptr.shape[0] - 1, C_d + src.shape[1]))


@torch.library.custom_op("shpl::pool_backward", mutates_args=(), device_types="cuda")
def pool_backward(g_fused: torch.Tensor, ptrT: torch.Tensor, keyT: torch.Tensor, idxT: torch.Tensor, valT: torch.Tensor,
                  nnz_max: int, C_d: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """g_fused [n_rows, C_d + C_s] -> (g_dst [n_rows, C_d], g_src [n_src, C_s]); deterministic, no atomics."""
    _check(g_fused, "g_fused", torch.float32)
    if g_fused.dim() != 2 or not (0 <= C_d < g_fused.shape[1]):
        raise ValueError("g_fused must be [n_rows, C_d + C_s] with C_s > 0")
    _check_csr(ptrT, keyT, idxT, valT, nnz_max, g_fused.device, "transposed CSR")
    g = g_fused.contiguous()
    n_rows, C = g.shape
    C_s = C - C_d
    n_src = ptrT.shape[0] - 1
    g_dst = torch.empty((n_rows, C_d), dtype=torch.float32, device=g.device)
    g_src = torch.empty((n_src, C_s), dtype=torch.float32, device=g.device)
    rc = _lib.shpl_pool_backward(_ptr(g), _ptr(ptrT), _ptr(keyT), _ptr(idxT), _ptr(valT), int(nnz_max), 0, n_rows, C_d, n_src,
                                 C_s, _ptr(g_dst) if C_d else None, _ptr(g_src), _stream())
    _cabi.check(rc, "shpl_pool_backward")
    return g_dst, g_src


@pool_backward.register_fake
def _(g_fused, ptrT, keyT, idxT, valT, nnz_max, C_d):
    return (g_fused.new_empty((g_fused.shape[0], C_d)), g_fused.new_empty((ptrT.shape[0] - 1, g_fused.shape[1] - C_d)))


def _setup_context(ctx, inputs, output):
    dst, src, ptr, key, idx, val, ptrT, keyT, idxT, valT, nnz_max = inputs
    ctx.save_for_backward(ptrT, keyT, idxT, valT)
    ctx.nnz_max = nnz_max
    ctx.C_d = 0 if dst is None else dst.shape[1]
    ctx.has_dst = dst is not None


def _backward(ctx, g_fused):
    ptrT, keyT, idxT, valT = ctx.saved_tensors
    g_dst, g_src = torch.ops.shpl.pool_backward(g_fused, ptrT, keyT, idxT, valT, ctx.nnz_max, ctx.C_d)
    return (g_dst if ctx.has_dst else None, g_src) + (None,) * 9


torch.library.register_autograd("shpl::pool", _backward, setup_context=_setup_context)


def sparse_pool(dst, src, plan, transposed=False):
    """[B,Hd,Wd,Cd] (or None), [B,Hs,Ws,Cs] -> [B,Hd,Wd,Cd+Cs] through the registered op (same result as
    ops.sparse_pool; `plan` is an ops.SparsePoolPlan)."""
    a = (plan.row_ptr, plan.csr_row, plan.csr_src, plan.csr_val)
    b = (plan.pix_ptr, plan.csrT_pix, plan.csrT_dst, plan.csrT_val)
    fwd, bwd = (b, a) if transposed else (a, b)
    d2 = None if dst is None else dst.reshape(-1, dst.shape[-1])
    fused = torch.ops.shpl.pool(d2, src.reshape(-1, src.shape[-1]), *fwd, *bwd, int(plan.entry_bound))
    if dst is not None:
        return fused.reshape(dst.shape[0], dst.shape[1], dst.shape[2], -1)
    return fused
