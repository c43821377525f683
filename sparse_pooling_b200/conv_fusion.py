"""sparse_pool_layer + the post-fusion 3x3 convolution in one call, with the fused (concat) map never written
(SURVEY.md 8(f) rank 3).

The reference's call sites are
    bev_fused, _ = sparse_pool_layer([bev, img], feature_depths, M, img_index_flip=..., bv_index=None, ...)
    bev = slim.conv2d(bev_fused, feature_depths[0], [3, 3], ...)      rpn_model.py:335-346, retinanet_model.py:337-348
(the `rpn_sparse_pooling_conv_after_fusion` switch, model.proto:89).  `sparse_pool_conv3x3` takes what those two calls
take -- the two maps, M, the gather index, the conv weights in slim's HWIO layout -- and returns what the second one
returns.  All arithmetic runs in libshpl.so (shpl_pool_conv3x3_forward): the dense half on the tcgen05 tensor cores
(3xTF32), the pooled half as a sparse update.  There is no CPU fallback.
"""
import torch

from . import _cabi, ops
from .ops import _ptr, _stream
from .sparse_pool_utils import _check_oob, _resolve_plan

_lib = _cabi.lib


def conv_workspace(frames, H, W, device, nnz_max=0):
    need = int(_lib.shpl_conv3x3_workspace_bytes(int(frames), int(H), int(W), int(nnz_max)))
    return torch.empty(need + 256, dtype=torch.uint8, device=device)


def sparse_pool_conv3x3(inputs, M, img_index_flip, weight, scale=None, shift=None, relu=False, out=None, workspace=None):
    """act(scale * conv3x3(concat(input_bv, pooled(input_img)), weight) + shift).

    inputs = [input_bv [B,H,W,C_b], input_img [B,H_i,W_i,C_i]] NHWC float32 CUDA tensors; M / img_index_flip as in
    sparse_pool_layer (None, None: plain conv of input_bv with weight [3,3,C_b,C_out]); weight [3,3,C_b+C_i,C_out]
    (slim.conv2d's variable, padding SAME, stride 1); scale / shift [C_out]: bias or folded inference batch norm."""
    bev, img = inputs[0], inputs[1]
    ops.require_cuda(bev, "inputs[0]")
    if bev.dtype != torch.float32 or weight.dtype != torch.float32:
        raise ValueError("SHPL feature maps and weights must be float32 (the reference's dtype)")
    B, H, W, Cb = bev.shape
    pooled = M is not None
    if pooled:
        ops.require_cuda(img, "inputs[1]")
        plan = _resolve_plan(M, img_index_flip, H * W, (img.shape[1], img.shape[2]), bev.device)
        _check_oob(plan)
        if plan.frames != B:
            raise ValueError("feature batch %d != frames in the plan %d" % (B, plan.frames))
        Ci, n_src = img.shape[3], plan.n_src
        ptr, key, idx, val, nnz_max, _, _ = plan.by_row()
        src = img.contiguous()
    else:
        Ci, n_src, ptr, key, idx, val, nnz_max, src = 0, 0, None, None, None, None, 0, None
    if tuple(weight.shape[:3]) != (3, 3, Cb + Ci):
        raise ValueError("weight must be [3, 3, %d, C_out] (HWIO), got %s" % (Cb + Ci, tuple(weight.shape)))
    Cout = weight.shape[3]
    w = weight.contiguous()
    b = bev.contiguous()
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=bev.device)
    ws = conv_workspace(B, H, W, bev.device, nnz_max) if workspace is None else workspace
    off = (-ws.data_ptr()) % 256
    sc = None if scale is None else scale.to(device=bev.device, dtype=torch.float32).contiguous()
    sh = None if shift is None else shift.to(device=bev.device, dtype=torch.float32).contiguous()
    import ctypes
    rc = _lib.shpl_pool_conv3x3_forward(_ptr(b), _ptr(src), _ptr(ptr), _ptr(key), _ptr(idx), _ptr(val), int(nnz_max),
                                        B, H, W, Cb, int(n_src), Ci, _ptr(w), Cout, _ptr(sc), _ptr(sh), int(bool(relu)),
                                        _ptr(out), ctypes.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream())
    _cabi.check(rc, "shpl_pool_conv3x3_forward")
    return out


class SparsePoolConv3x3Function(torch.autograd.Function):
    """conv3x3(concat(bev, pooled(img)), weight) with autograd: forward = shpl_pool_conv3x3_forward without epilogue,
    backward = shpl_pool_conv3x3_backward (g_bev on the tensor cores, g_img through the transposed CSR, g_weight).
    Bias / batch norm / ReLU follow as ordinary torch ops, like slim's normalizer and activation follow the conv."""

    @staticmethod
    def forward(ctx, bev, img, weight, plan):
        out = sparse_pool_conv3x3([bev, img], plan, None, weight)
        ctx.save_for_backward(bev, img, weight)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, g_out):
        bev, img, weight = ctx.saved_tensors
        need_b, need_i, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g_bev, g_img, g_w = sparse_pool_conv3x3_backward(g_out, [bev, img], ctx.plan, weight, need_b, need_i, need_w)
        return g_bev, g_img, g_w, None


def sparse_pool_conv3x3_backward(g_out, inputs, plan, weight, need_bev=True, need_img=True, need_weight=True, out=None,
                                 workspace=None):
    """Gradients of conv3x3(concat(bev, pooled(img)), weight) with respect to bev, img and weight
    (shpl_pool_conv3x3_backward); `out` = preallocated (g_bev, g_img, g_weight) or None."""
    import ctypes
    bev, img = inputs[0], inputs[1]
    B, H, W, Cb = bev.shape
    Ci = img.shape[3]
    g = g_out.contiguous()
    g_bev = (out[0] if out is not None else torch.empty_like(bev)) if need_bev else None
    g_img = (out[1] if out is not None else torch.empty_like(img)) if need_img else None
    g_w = (out[2] if out is not None else torch.empty_like(weight)) if need_weight else None
    ptr, key, idx, val, nnz_max, _, _ = plan.by_row()
    ptrT, keyT, idxT, valT, _, _, _ = plan.by_pixel()
    if workspace is None:
        need = int(_lib.shpl_conv3x3_backward_workspace_bytes(int(nnz_max)))
        workspace = ops.scratch("conv_bwd", bev.device, need + 256)
    off = (-workspace.data_ptr()) % 256
    rc = _lib.shpl_pool_conv3x3_backward(_ptr(g), _ptr(bev.contiguous()), _ptr(img.contiguous()), _ptr(ptr), _ptr(key), _ptr(idx),
                                         _ptr(val), _ptr(ptrT), _ptr(keyT), _ptr(idxT), _ptr(valT), int(nnz_max), B, H, W, Cb,
                                         int(plan.n_src), Ci, _ptr(weight.contiguous()), weight.shape[3], _ptr(g_bev), _ptr(g_img),
                                         _ptr(g_w), ctypes.c_void_p(workspace.data_ptr() + off), workspace.numel() - off, _stream())
    _cabi.check(rc, "shpl_pool_conv3x3_backward")
    return g_bev, g_img, g_w


def sparse_pool_conv3x3_autograd(inputs, M, img_index_flip, weight):
    """The differentiable form: [bev, img] NHWC float32 CUDA tensors, M / img_index_flip as in sparse_pool_layer,
    weight [3, 3, C_b + C_i, C_out] -> conv3x3(concat(bev, pooled(img)), weight) (no bias / activation)."""
    bev, img = inputs[0], inputs[1]
    plan = _resolve_plan(M, img_index_flip, bev.shape[1] * bev.shape[2], (img.shape[1], img.shape[2]), bev.device)
    _check_oob(plan)
    return SparsePoolConv3x3Function.apply(bev, img, weight, plan)
