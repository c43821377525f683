"""The registered PyTorch custom ops (torch.ops.shpl.pool / pool_backward): same results as the autograd.Function
path and the oracle, torch.library.opcheck (schema, fake tensors, autograd registration), and tracing through
torch.compile(backend="eager").  Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import cref, index_oracle as io, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shpl():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sparse_pooling_b200 as m
    return m


def _case(shpl, seed=3, n=1500, bev_hw=(40, 50), img_hw=(30, 60), cb=16, ci=24):
    d = synth.direct_pairs(seed, n, bev_hw=bev_hw, img_wh=(img_hw[1], img_hw[0]), skew="ground")
    o_ref = io.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()}, stride=[1, 1])
    rng = np.random.default_rng(seed)
    val = (1.0 / rng.integers(1, 46, len(o_ref["Mij_pool"]))).astype(np.float32)
    o = shpl.produce_sparse_pooling_input(d, M_val=val.astype(np.float64))
    bev = rng.standard_normal((1,) + bev_hw + (cb,), dtype=np.float32)
    img = rng.standard_normal((1,) + img_hw + (ci,), dtype=np.float32)
    return o, o_ref, val, bev, img


@pytest.mark.parametrize("transposed", [False, True])
def test_custom_op_matches_function_path_and_oracle(shpl, transposed):
    o, o_ref, val, bev, img = _case(shpl)
    plan = o["shpl_plan"]
    Mij, flip = o_ref["Mij_pool"], o_ref["img_index_flip_pool"]
    dst_np, src_np = (img, bev) if transposed else (bev, img)
    dst = torch.from_numpy(dst_np).cuda().requires_grad_(True)
    src = torch.from_numpy(src_np).cuda().requires_grad_(True)
    fused = shpl.torch_op.sparse_pool(dst, src, plan, transposed=transposed)
    ref = (cref.forward_trans(img[0], bev[0], Mij, val, flip) if transposed else cref.forward(bev[0], img[0], Mij, val, flip))
    np.testing.assert_array_equal(fused[0].detach().cpu().numpy(), ref)
    g = np.random.default_rng(5).standard_normal(ref.shape, dtype=np.float32)
    fused.backward(torch.from_numpy(g[None]).cuda())
    if transposed:
        gd, gs = cref.backward_trans(g, Mij, val, flip, img.shape[-1], bev.shape[1:])
    else:
        gd, gs = cref.backward(g, Mij, val, flip, bev.shape[-1], img.shape[1:])
    np.testing.assert_array_equal(dst.grad[0].cpu().numpy(), gd)
    np.testing.assert_array_equal(src.grad[0].cpu().numpy(), gs)
    # the autograd.Function path gives the same bits
    d2 = torch.from_numpy(dst_np).cuda().requires_grad_(True)
    s2 = torch.from_numpy(src_np).cuda().requires_grad_(True)
    f2 = shpl.sparse_pool(d2, s2, plan, transposed=transposed)
    assert torch.equal(f2, fused)
    f2.backward(torch.from_numpy(g[None]).cuda())
    assert torch.equal(d2.grad, dst.grad) and torch.equal(s2.grad, src.grad)


def test_custom_op_opcheck_and_compile(shpl):
    o, o_ref, val, bev, img = _case(shpl, seed=4, n=600, cb=8, ci=8)
    plan = o["shpl_plan"]
    dst = torch.from_numpy(bev).cuda().reshape(-1, 8).requires_grad_(True)
    src = torch.from_numpy(img).cuda().reshape(-1, 8).requires_grad_(True)
    args = (dst, src, plan.row_ptr, plan.csr_row, plan.csr_src, plan.csr_val, plan.pix_ptr, plan.csrT_pix, plan.csrT_dst,
            plan.csrT_val, int(plan.entry_bound))
    torch.library.opcheck(torch.ops.shpl.pool.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    g = torch.randn(dst.shape[0], 16, device="cuda")
    torch.library.opcheck(torch.ops.shpl.pool_backward.default,
                          (g, plan.pix_ptr, plan.csrT_pix, plan.csrT_dst, plan.csrT_val, int(plan.entry_bound), 8),
                          test_utils=("test_schema", "test_faketensor"))
    # pooled map only (dst = None): the bare _sparse_pool_op
    pooled = torch.ops.shpl.pool(None, src, *args[2:])
    assert pooled.shape == (dst.shape[0], 8)
    full = torch.ops.shpl.pool(*args)
    assert torch.equal(full[:, 8:], pooled) and torch.equal(full[:, :8], dst.detach())

    @torch.compile(backend="eager", fullgraph=True)
    def layer(d, s):
        return torch.ops.shpl.pool(d, s, *args[2:]) * 2.0

    out = layer(dst, src)
    assert torch.equal(out, full * 2.0)
    out.sum().backward()
    assert dst.grad is not None and torch.equal(dst.grad, torch.full_like(dst, 2.0))
