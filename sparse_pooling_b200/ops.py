"""Device-side objects of the SHPL path: the CSR plan, the workspace, and the
PyTorch autograd op around the two pooling kernels of libshpl.so.

PyTorch is used only for device memory, streams and autograd bookkeeping; all
arithmetic happens in the CUDA kernels reached through the ctypes C ABI
(``_cabi.py`` / ``include/shpl.h``).  Nothing here computes on the CPU.
"""
import ctypes
import threading

import torch

from . import _cabi

_lib = _cabi.lib


def _ptr(t):
    if t is None or isinstance(t, ctypes.c_void_p):
        return t
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: torch.cuda.current_stream()
    costs ~20 us of Python per call, this costs well under 1 us)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the SHPL kernels run on the GPU only (no CPU fallback)" % name)


_scratch = {}
_scratch_lock = threading.Lock()


def scratch(tag, device, nbytes):
    """Scratch bytes for one asynchronous library call, cached per (purpose, device, CUDA STREAM, host thread).

    Two calls in flight on different streams (or threads) therefore never share a buffer -- the library's scratch
    holds tile tickets, look-back words and sort items, which concurrent builds would corrupt.  A buffer is only ever
    used on the stream it was allocated under, so when it is replaced by a larger one torch's caching allocator (which
    is stream-ordered) cannot hand the old block to anybody before the kernels queued on it have run."""
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch._C._cuda_getCurrentRawStream(dev_index)
    key = (tag, dev_index, stream, threading.get_ident())
    ws = _scratch.get(key)
    if ws is None or ws.numel() < nbytes:
        with torch.cuda.device(dev_index):
            ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        with _scratch_lock:
            _scratch[key] = ws
    return ws


def workspace(device, n_max):
    """Scratch for the correspondence builder (shpl_build_* / shpl_plan_from_*), for the current stream."""
    return scratch("build", device, int(_lib.shpl_build_workspace_bytes(int(n_max))))


_staging = threading.local()


def _count_staging(n):
    """(pinned int32 tensor, its numpy view) of at least n elements, one per host thread."""
    st = getattr(_staging, "buf", None)
    if st is None or st[0].shape[0] < n:
        t = torch.empty(max(int(n), 64), dtype=torch.int32).pin_memory()
        st = _staging.buf = (t, t.numpy())
    return st


class SparsePoolPlan:
    """Canonical CSR (by destination BEV cell) + CSR^T (by source pixel) of one M,
    or of the Ms of several frames stacked (struct shpl_plan of include/shpl.h).

    rows_per_frame = H_b'*W_b', src_per_frame = H_i'*W_i'; with `frames` > 1 the
    rows / pixels of frame f are offset by f*rows_per_frame / f*src_per_frame so a
    [B,H,W,C] batch is pooled by one launch."""

    _INT_FIELDS = ("row_ptr", "pix_ptr", "csr_row", "csr_src", "csrT_pix", "csrT_dst", "heavy_row", "heavy_pix")

    def __init__(self, rows_per_frame, src_hw, capacity, device, frames=1, zero_meta=True):
        """zero_meta=False skips clearing the counters: for callers that run a builder on the plan at once (every
        builder writes the counters of its frame, and the first frame's clears the heavy-cell counters)."""
        self.frames = int(frames)
        self.rows_per_frame = int(rows_per_frame)
        self.src_hw = (int(src_hw[0]), int(src_hw[1]))
        self.src_per_frame = self.src_hw[0] * self.src_hw[1]
        self.capacity = int(max(capacity, 1))
        self.device = device
        R, Q = self.rows_per_frame * self.frames, self.src_per_frame * self.frames
        if max(R, Q, self.capacity) >= 2 ** 31 - 1:
            raise ValueError("SHPL plan too large for 32-bit indices")
        # one allocation for every int32 array, one for the weights (16-byte aligned sub-arrays).  The tensor
        # views of the sub-arrays are made on first use only (a view costs microseconds of host time and the
        # hot path needs nothing but addresses).
        cap = self.capacity

        def up(x):
            return (x + 3) // 4 * 4
        self.heavy_cap = cap // _cabi.HEAVY_LEN + 1      # cells that can hold more than HEAVY_LEN of `cap` entries
        hc = up(self.heavy_cap)
        lens = [R + 1, Q + 1, cap, cap, cap, cap, self.heavy_cap, self.heavy_cap]
        sizes = [up(R + 1), up(Q + 1), up(cap), up(cap), up(cap), up(cap), hc, hc, up(8 * self.frames) + 4]
        self._ints = torch.empty(sum(sizes), dtype=torch.int32, device=device)
        self._span = {}
        off = 0
        for name, n, sz in zip(self._INT_FIELDS, lens, sizes):
            self._span[name] = (off, n)
            off += sz
        self._meta_off = off                               # counts of every frame, then the 2 heavy counters
        self._heavy_count_off = off + up(8 * self.frames)
        self._views = {}
        self._meta = self._ints[off:]
        if zero_meta:
            self._meta.zero_()
        self._vals = torch.empty(2 * up(cap), dtype=torch.float32, device=device)
        self._val_span = {"csr_val": (0, cap), "csrT_val": (up(cap), cap)}
        self._ibase = self._ints.data_ptr()
        self._vbase = self._vals.data_ptr()
        self.n_heavy = None   # (rows, pixels) with more than HEAVY_LEN entries; host ints after read_counts
        self.n_long = None    # (rows, pixels) with more than SHPL_LONG_LEN entries; host ints after read_counts
        self._heavy_len = None
        self.entry_bound = self.capacity   # host-known upper bound on the entries (builders tighten it)
        self.nnz = None       # columns of M per frame (host ints), known after the builder's read-back
        self.n_oob = None     # entries TF-CPU would reject, per frame
        self._ptr8 = None

    def __getattr__(self, name):
        # lazily made tensor views of the sub-arrays (row_ptr, csr_src, csr_val, counts, ...)
        d = self.__dict__
        views = d.get("_views")
        if views is None:
            raise AttributeError(name)
        v = views.get(name)
        if v is not None:
            return v
        if name in d.get("_span", ()):
            off, n = d["_span"][name]
            v = d["_ints"][off:off + n]
        elif name in d.get("_val_span", ()):
            off, n = d["_val_span"][name]
            v = d["_vals"][off:off + n]
        elif name == "counts":
            v = d["_ints"][d["_meta_off"]:d["_meta_off"] + 8 * d["frames"]].view(d["frames"], 8)
        elif name == "heavy_count":
            v = d["_ints"][d["_heavy_count_off"]:d["_heavy_count_off"] + 2]
        else:
            raise AttributeError(name)
        views[name] = v
        return v

    def addr(self, name):
        """Device address of a sub-array (no tensor view is created)."""
        if name in self._span:
            return self._ibase + 4 * self._span[name][0]
        if name in self._val_span:
            return self._vbase + 4 * self._val_span[name][0]
        if name == "counts":
            return self._ibase + 4 * self._meta_off
        if name == "heavy_count":
            return self._ibase + 4 * self._heavy_count_off
        raise KeyError(name)

    def ptrs8(self):
        """ctypes pointers of the eight index arrays, in the argument order of the dual-direction entry points."""
        if self._ptr8 is None:
            self._ptr8 = [ctypes.c_void_p(self.addr(n)) for n in ("row_ptr", "csr_row", "csr_src", "csr_val",
                                                                  "pix_ptr", "csrT_pix", "csrT_dst", "csrT_val")]
        return self._ptr8

    @property
    def n_rows(self):
        return self.rows_per_frame * self.frames

    @property
    def n_src(self):
        return self.src_per_frame * self.frames

    def _by(self, by_pixel):
        # the argument tuple of one direction, cached until the entry bound or the counters change (a hot-path call per
        # forward and per backward: building it costs ~5 us of ctypes objects)
        state = (self.entry_bound, self.n_heavy, self.n_long)
        c = self.__dict__.get("_by_cache")
        if c is None or c[0] != state:
            c = self.__dict__["_by_cache"] = (state, {})
        t = c[1].get(by_pixel)
        if t is None:
            p = self.ptrs8()
            o = 4 if by_pixel else 0
            t = c[1][by_pixel] = (p[o], p[o + 1], p[o + 2], p[o + 3], self.entry_bound, self.heavy(by_pixel), self.heavy_len(by_pixel))
        return t

    def by_row(self):
        """(ptr, key, idx, val, nnz_max, heavy list, heavy_len) of the CSR keyed by destination BEV cell."""
        return self._by(False)

    def by_pixel(self):
        """(ptr, key, idx, val, nnz_max, heavy list, heavy_len) of the CSR^T keyed by source pixel."""
        return self._by(True)

    def frame_struct(self, f):
        """shpl_plan for frame f: ptr arrays point at the frame's sub-array; entry arrays are shared."""
        s = _cabi.ShplPlan()
        a = self.addr
        s.n_rows = self.rows_per_frame
        s.n_src = self.src_per_frame
        s.capacity = self.capacity
        s.row_ptr = a("row_ptr") + 4 * f * self.rows_per_frame
        s.pix_ptr = a("pix_ptr") + 4 * f * self.src_per_frame
        s.csr_row = a("csr_row")
        s.csr_src = a("csr_src")
        s.csrT_pix = a("csrT_pix")
        s.csr_val = a("csr_val")
        s.csrT_dst = a("csrT_dst")
        s.csrT_val = a("csrT_val")
        s.heavy_cap = self.heavy_cap
        s.heavy_row = a("heavy_row")
        s.heavy_pix = a("heavy_pix")
        s.heavy_count = a("heavy_count")
        s.counts = a("counts") + 32 * f
        return s

    def entry_base(self, f):
        """Device pointer to the running entry offset frame f must start at (None for frame 0)."""
        if f == 0:
            return None
        return ctypes.c_void_p(self.addr("counts") + 32 * (f - 1) + 16)   # counts[f-1][4]

    def read_counts(self):
        """One small device->host copy into pinned staging memory: fills nnz / n_oob / n_heavy (synchronises the
        stream).  Returns the [frames, 8] counters as a numpy array."""
        n = self._meta.shape[0]
        stage = _count_staging(n)
        stage[0][:n].copy_(self._meta)               # pinned destination: one async copy + a stream synchronise
        m = stage[1][:n].copy()
        c = m[:8 * self.frames].reshape(self.frames, 8)
        self.nnz = c[:, 1].tolist()
        self.n_oob = c[:, 2].tolist()
        h = (8 * self.frames + 3) // 4 * 4
        self.n_heavy = (min(int(m[h]), self.heavy_cap), min(int(m[h + 1]), self.heavy_cap))
        self.n_long = (int(c[:, 6].sum()), int(c[:, 7].sum()))
        self._heavy_len = None
        return c

    def heavy_len(self, by_pixel=None):
        """The heavy_len argument of the pooling entry points for this plan, from the counters read back (direction: rows,
        pixels, None = either): SHPL_HEAVY_LEN when cells are listed for shpl_pool_heavy; SHPL_EXACT_LEN when some cell has
        more than SHPL_LONG_LEN entries but none is listed (no cell is left out; the kernels take their long-cell paths);
        0 when no cell is long (the kernels skip every long-cell path: every KITTI / MV3D plan)."""
        if self.n_long is None or self.n_heavy is None:
            return _cabi.HEAVY_LEN
        hl = self._heavy_len
        if hl is None:
            def level(n_heavy, n_long):
                if n_heavy > 0:
                    return _cabi.HEAVY_LEN      # listed cells: left to shpl_pool_heavy; the kernels tune for long runs
                return _cabi.EXACT_LEN if n_long > 0 else 0     # long but not listed: nothing is left out
            h, g = self.n_heavy, self.n_long
            hl = self._heavy_len = {False: level(h[0], g[0]), True: level(h[1], g[1]), None: level(h[0] + h[1], g[0] + g[1])}
        return hl[by_pixel]

    def heavy(self, by_pixel):
        """(list pointer, device counter pointer, list capacity, how many to expect or None when the counters have
        not been read back)."""
        k = 1 if by_pixel else 0
        return (ctypes.c_void_p(self.addr("heavy_pix" if by_pixel else "heavy_row")),
                ctypes.c_void_p(self.addr("heavy_count") + 4 * k), self.heavy_cap,
                None if self.n_heavy is None else self.n_heavy[k], self.entry_bound, self.device)


def plan_from_coo(indices, values, source_index, n_rows, src_hw, device=None):
    """Plan of an arbitrary tf.SparseTensor-like COO (shpl_plan_from_coo)."""
    require_cuda(indices, "M.indices")
    device = indices.device
    Mij = indices.to(torch.int64).contiguous().reshape(-1, 2)
    val = values.to(device=device, dtype=torch.float32).contiguous().reshape(-1)
    src_index = source_index.to(device=device)
    if src_index.dtype not in (torch.int32, torch.int64):
        src_index = src_index.to(torch.int64)
    src_index = src_index.contiguous().reshape(-1, 3)
    m = Mij.shape[0]
    if val.shape[0] != m:
        raise ValueError("M.values has %d entries, M.indices %d" % (val.shape[0], m))
    plan = SparsePoolPlan(n_rows, src_hw, m, device)
    ws = workspace(device, m)
    st = plan.frame_struct(0)
    rc = _lib.shpl_plan_from_coo(_ptr(Mij), _ptr(val), m, _ptr(src_index), int(src_index.dtype == torch.int64),
                                 src_index.shape[0], plan.src_hw[0], plan.src_hw[1], ctypes.byref(st), 0, 0, None,
                                 _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "shpl_plan_from_coo")
    plan._keep = (Mij, val, src_index)
    return plan


def _run_heavy(heavy, gather_in, gather_stride, C, ptr, idx, val, addend, addend_stride, out, out_stride):
    """shpl_pool_heavy on the listed heavy cells, unless the host already knows there are none."""
    lst, count_dev, cap, expected, nnz_max, device = heavy
    if expected == 0:
        return
    # long cells are split over many CTAs (shpl_pool_heavy_split); the partial sums live in per-stream scratch
    need = int(_lib.shpl_pool_heavy_workspace_bytes(int(C), int(nnz_max), int(cap)))
    ws = scratch("heavy", device, need)
    rc = _lib.shpl_pool_heavy_split(gather_in, gather_stride, C, _ptr(ptr), _ptr(idx), _ptr(val), lst, count_dev,
                                    int(cap), addend, addend_stride, out, out_stride, int(nnz_max), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "shpl_pool_heavy_split")


_heavy_streams = {}


def _launch_with_heavy(device, main, heavy_jobs):
    """Launches `main()` (the pooling kernel) on the current stream and the heavy-cell kernels of `heavy_jobs` (argument
    tuples of _run_heavy) on a side stream BESIDE it: the main kernel writes nothing for listed cells, the heavy kernels
    nothing else, so the two are independent -- the long gather chains of a few giant cells hide under the bandwidth-bound
    main kernel.  When the plan's counters say there is no heavy cell (the KITTI / MV3D case) only main() runs."""
    jobs = [j for j in heavy_jobs if j[0][3] != 0]
    if not jobs:
        main()
        return
    cur = torch.cuda.current_stream()
    key = device.index if device.index is not None else torch.cuda.current_device()
    side = _heavy_streams.get(key)
    if side is None:
        side = _heavy_streams[key] = torch.cuda.Stream(device=device)
    side.wait_stream(cur)                  # the plan and the inputs are ready
    main()
    with torch.cuda.stream(side):
        for j in jobs:
            _run_heavy(*j)
    cur.wait_stream(side)


def _off(t, n_floats):
    return ctypes.c_void_p(t.data_ptr() + 4 * n_floats)


def pool_forward(dst, src, csr, n_rows, n_src):
    """fused[r] = concat(dst[r], sum_k val_k * src[idx_k])  (shpl_pool_forward, then shpl_pool_heavy for
    cells with more than HEAVY_LEN entries).  csr = (ptr, key, idx, val, nnz_max, heavy)."""
    ptr, key, idx, val, nnz_max, heavy, heavy_len = csr
    require_cuda(src, "source feature map")
    C_s = src.shape[-1]
    C_d = 0 if dst is None else dst.shape[-1]
    fused = torch.empty((n_rows, C_d + C_s), dtype=torch.float32, device=src.device)

    def main():
        rc = _lib.shpl_pool_forward(_ptr(dst), _ptr(src), _ptr(ptr), _ptr(key), _ptr(idx), _ptr(val), int(nnz_max),
                                    heavy_len, n_rows, C_d, n_src, C_s, _ptr(fused), _stream())
        _cabi.check(rc, "shpl_pool_forward")
    _launch_with_heavy(src.device, main, [(heavy, _ptr(src), C_s, C_s, ptr, idx, val, None, 0, _off(fused, C_d), C_d + C_s)])
    return fused


def pool_backward(g_fused, csrT, n_rows, C_d, n_src, C_s, want_dst=True):
    ptrT, keyT, idxT, valT, nnz_max, heavy, heavy_len = csrT
    g_dst = torch.empty((n_rows, C_d), dtype=torch.float32, device=g_fused.device) if (want_dst and C_d) else None
    g_src = torch.empty((n_src, C_s), dtype=torch.float32, device=g_fused.device)
    def main():
        rc = _lib.shpl_pool_backward(_ptr(g_fused), _ptr(ptrT), _ptr(keyT), _ptr(idxT), _ptr(valT), int(nnz_max),
                                     heavy_len, n_rows, C_d, n_src, C_s, _ptr(g_dst), _ptr(g_src), _stream())
        _cabi.check(rc, "shpl_pool_backward")
    _launch_with_heavy(g_fused.device, main, [(heavy, _off(g_fused, C_d), C_d + C_s, C_s, ptrT, idxT, valT, None, 0, _ptr(g_src), C_s)])
    return g_dst, g_src


def pool_forward_into(fused, src, csr, n_rows, n_src, chan_off):
    """No-concat forward (shpl_pool_forward_into): fused [n_rows, F] gets its channels chan_off : chan_off + C_s
    overwritten with the pooled sums (zeros for cells that receive nothing); the other channels are left alone."""
    ptr, key, idx, val, nnz_max, heavy, heavy_len = csr
    C_s, F = src.shape[-1], fused.shape[-1]
    def main():
        rc = _lib.shpl_pool_forward_into(_ptr(src), _ptr(ptr), _ptr(key), _ptr(idx), _ptr(val), int(nnz_max), heavy_len,
                                         n_rows, n_src, C_s, _ptr(fused), F, int(chan_off), _stream())
        _cabi.check(rc, "shpl_pool_forward_into")
    _launch_with_heavy(src.device, main, [(heavy, _ptr(src), C_s, C_s, ptr, idx, val, None, 0, _off(fused, chan_off), F)])
    return fused


def pool_backward_from(g_fused, csrT, n_rows, n_src, C_s, chan_off):
    """No-concat backward (shpl_pool_backward_from): the gradient of the gathered map from the pooled channels of
    g_fused, read in place; the gradient of the destination map is the view g_fused[:, :chan_off]."""
    ptrT, keyT, idxT, valT, nnz_max, heavy, heavy_len = csrT
    F = g_fused.shape[-1]
    g_src = torch.empty((n_src, C_s), dtype=torch.float32, device=g_fused.device)
    def main():
        rc = _lib.shpl_pool_backward_from(_ptr(g_fused), F, int(chan_off), _ptr(ptrT), _ptr(keyT), _ptr(idxT), _ptr(valT),
                                          int(nnz_max), heavy_len, n_rows, n_src, C_s, _ptr(g_src), _stream())
        _cabi.check(rc, "shpl_pool_backward_from")
    _launch_with_heavy(g_fused.device, main, [(heavy, _off(g_fused, chan_off), F, C_s, ptrT, idxT, valT, None, 0, _ptr(g_src), C_s)])
    return g_src


class SparsePoolIntoFunction(torch.autograd.Function):
    """One direction of SHPL in the no-concat form (SURVEY.md 8(d) "sparse-only"): `fused` [B,H,W,C_d+C_s] already
    holds the destination map in its first C_d channels (its producer wrote there); the pooled channels are written in
    place and the same tensor is returned.  Backward: the gradient handed back for `fused` is g_fused itself (no slice
    copy; its pooled channels mean nothing to a producer that wrote only the first C_d), the gradient of `src` comes
    from the deterministic transpose-CSR kernel."""

    @staticmethod
    def forward(ctx, fused, src, plan, transposed):
        require_cuda(fused, "fused buffer")
        require_cuda(src, "source feature map")
        if fused.dtype != torch.float32 or src.dtype != torch.float32:
            raise ValueError("SHPL feature maps must be float32 (the reference's dtype)")
        if not fused.is_contiguous():
            raise ValueError("the fused buffer must be contiguous (NHWC)")
        if transposed:
            csr, n_rows, n_src = plan.by_pixel(), plan.n_src, plan.n_rows
        else:
            csr, n_rows, n_src = plan.by_row(), plan.n_rows, plan.n_src
        src_c = src.contiguous()
        C_s, F = src_c.shape[-1], fused.shape[-1]
        if src_c.numel() != n_src * C_s or fused.numel() != n_rows * F or F <= C_s:
            raise ValueError("maps %s / %s do not match the plan (%d cells <- %d cells)" % (tuple(fused.shape), tuple(src.shape), n_rows, n_src))
        pool_forward_into(fused, src_c, csr, n_rows, n_src, F - C_s)
        ctx.mark_dirty(fused)
        ctx.plan, ctx.transposed, ctx.src_shape, ctx.chan_off = plan, transposed, tuple(src.shape), F - C_s
        return fused

    @staticmethod
    def backward(ctx, g_fused):
        plan = ctx.plan
        if ctx.transposed:
            csrT, n_rows, n_src = plan.by_row(), plan.n_src, plan.n_rows
        else:
            csrT, n_rows, n_src = plan.by_pixel(), plan.n_rows, plan.n_src
        g = g_fused.contiguous()
        g_src = pool_backward_from(g, csrT, n_rows, n_src, ctx.src_shape[-1], ctx.chan_off)
        return g, g_src.reshape(ctx.src_shape), None, None


def sparse_pool_into(fused, src, plan, transposed=False):
    """fused [B,Hd,Wd,Cd+Cs] (first Cd channels already written), src [B,Hs,Ws,Cs] -> fused, pooled channels filled in."""
    return SparsePoolIntoFunction.apply(fused, src, plan, transposed)


class SparsePoolFunction(torch.autograd.Function):
    """One direction of SHPL with the channel concat fused in.

    transposed=False: img -> bev  (_sparse_pool_op + tf.concat, sparse_pool_utils.py:65-72)
    transposed=True : bev -> img  (_sparse_pool_trans_op + tf.concat, :79-87)
    dst may be None: the pooled map alone (the bare _sparse_pool_op / _sparse_pool_trans_op).
    Backward is the deterministic transpose-CSR kernel; no gradient
    flows to M's values (they are a placeholder in the reference, rpn_model.py:219-242)."""

    @staticmethod
    def forward(ctx, dst, src, plan, transposed):
        require_cuda(src, "source feature map")
        B = src.shape[0]
        if B != plan.frames:
            raise ValueError("feature batch %d != frames in the plan %d" % (B, plan.frames))
        if transposed:
            csr, n_rows, n_src = plan.by_pixel(), plan.n_src, plan.n_rows
        else:
            csr, n_rows, n_src = plan.by_row(), plan.n_rows, plan.n_src
        src_c = src.contiguous()
        if src_c.dtype != torch.float32:
            raise ValueError("SHPL feature maps must be float32 (the reference's dtype)")
        if src_c.shape[0] * src_c.shape[1] * src_c.shape[2] != n_src:
            raise ValueError("source map %s does not match the plan (%d cells)" % (tuple(src.shape), n_src))
        dst_c = None
        if dst is not None:
            require_cuda(dst, "destination feature map")
            dst_c = dst.contiguous()
            if dst_c.shape[0] * dst_c.shape[1] * dst_c.shape[2] != n_rows:
                raise ValueError("destination map %s does not match the plan (%d cells)" % (tuple(dst.shape), n_rows))
        fused = pool_forward(dst_c, src_c, csr, n_rows, n_src)
        ctx.plan = plan
        ctx.transposed = transposed
        ctx.src_shape = tuple(src.shape)
        ctx.dst_shape = None if dst is None else tuple(dst.shape)
        return fused

    @staticmethod
    def backward(ctx, g_fused):
        plan = ctx.plan
        if ctx.transposed:   # entries grouped by what was the *source* of the forward
            csrT, n_rows, n_src = plan.by_row(), plan.n_src, plan.n_rows
        else:
            csrT, n_rows, n_src = plan.by_pixel(), plan.n_rows, plan.n_src
        C_s = ctx.src_shape[-1]
        C_d = 0 if ctx.dst_shape is None else ctx.dst_shape[-1]
        g = g_fused.contiguous()
        g_dst, g_src = pool_backward(g, csrT, n_rows, C_d, n_src, C_s,
                                     want_dst=ctx.needs_input_grad[0] and C_d > 0)
        if g_dst is not None:
            g_dst = g_dst.reshape(ctx.dst_shape)
        return g_dst, g_src.reshape(ctx.src_shape), None, None


def _plan_ptrs(plan):
    return plan.ptrs8()


class SparsePoolDualFunction(torch.autograd.Function):
    """Both directions of sparse_pool_layer (bv_index is not None, sparse_pool_utils.py:65-87) in one
    launch each way: shpl_pool_forward_dual / shpl_pool_backward_dual.  The backward forms TF's AddN
    of the two partial gradients of each input in registers (no intermediate tensors)."""

    @staticmethod
    def forward(ctx, bev, img, plan):
        require_cuda(bev, "inputs[0]")
        require_cuda(img, "inputs[1]")
        if bev.dtype != torch.float32 or img.dtype != torch.float32:
            raise ValueError("SHPL feature maps must be float32 (the reference's dtype)")
        if bev.shape[0] != plan.frames or img.shape[0] != plan.frames:
            raise ValueError("feature batch does not match the frames in the plan")
        b, i = bev.contiguous(), img.contiguous()
        R, Q, Cb, Ci = plan.n_rows, plan.n_src, bev.shape[-1], img.shape[-1]
        if b.numel() != R * Cb or i.numel() != Q * Ci:
            raise ValueError("feature maps %s / %s do not match the plan (%d cells, %d pixels)"
                             % (tuple(bev.shape), tuple(img.shape), R, Q))
        fused_bev = torch.empty(tuple(bev.shape[:3]) + (Cb + Ci,), dtype=torch.float32, device=bev.device)
        fused_img = torch.empty(tuple(img.shape[:3]) + (Ci + Cb,), dtype=torch.float32, device=bev.device)
        P8 = plan.ptrs8()
        def main():
            rc = _lib.shpl_pool_forward_dual(_ptr(b), _ptr(i), *P8, int(plan.entry_bound), plan.heavy_len(),
                                             R, Cb, Q, Ci, _ptr(fused_bev), _ptr(fused_img), _stream())
            _cabi.check(rc, "shpl_pool_forward_dual")
        _launch_with_heavy(bev.device, main,
                           [(plan.heavy(False), _ptr(i), Ci, Ci, P8[0], P8[2], P8[3], None, 0, _off(fused_bev, Cb), Cb + Ci),
                            (plan.heavy(True), _ptr(b), Cb, Cb, P8[4], P8[6], P8[7], None, 0, _off(fused_img, Ci), Ci + Cb)])
        ctx.plan = plan
        ctx.shapes = (tuple(bev.shape), tuple(img.shape))
        return fused_bev, fused_img

    @staticmethod
    def backward(ctx, g_fused_bev, g_fused_img):
        plan = ctx.plan
        (sb, si) = ctx.shapes
        R, Q, Cb, Ci = plan.n_rows, plan.n_src, sb[-1], si[-1]
        gb = g_fused_bev.contiguous()
        gi = g_fused_img.contiguous()
        g_bev = torch.empty(sb, dtype=torch.float32, device=gb.device)
        g_img = torch.empty(si, dtype=torch.float32, device=gb.device)
        P8 = plan.ptrs8()
        rc = _lib.shpl_pool_backward_dual(_ptr(gb), _ptr(gi), *P8, int(plan.entry_bound), plan.heavy_len(),
                                          R, Cb, Q, Ci, _ptr(g_bev), _ptr(g_img), _stream())
        _cabi.check(rc, "shpl_pool_backward_dual")
        # g_bev[r] = g_fused_bev[r,:Cb] + sum_{k in row r} val * g_fused_img[pix_k, Ci:]   (heavy rows)
        _run_heavy(plan.heavy(False), _off(gi, Ci), Ci + Cb, Cb, P8[0], P8[2], P8[3],
                   _ptr(gb), Cb + Ci, _ptr(g_bev), Cb)
        _run_heavy(plan.heavy(True), _off(gb, Cb), Cb + Ci, Ci, P8[4], P8[6], P8[7],
                   _ptr(gi), Ci + Cb, _ptr(g_img), Ci)
        return g_bev, g_img, None


def sparse_pool_dual(bev, img, plan):
    """(bv_fused, img_fused) of the dual-direction layer."""
    return SparsePoolDualFunction.apply(bev, img, plan)


def sparse_pool(dst, src, plan, transposed=False):
    """[B,Hd,Wd,Cd] (or None), [B,Hs,Ws,Cs] -> [B,Hd,Wd,Cd+Cs]."""
    fused = SparsePoolFunction.apply(dst, src, plan, transposed)
    if dst is not None:
        return fused.reshape(dst.shape[0], dst.shape[1], dst.shape[2], -1)
    return fused
