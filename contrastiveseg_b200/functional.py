"""Autograd entry points of the engine: thin Python over the C ABI (include/pcl.h).

PyTorch is plumbing here (device memory, streams, autograd hand-off); all arithmetic of the path runs in
the CUDA kernels of libpcl_b200.so.  No fallback: tensors must be CUDA tensors and the library must load.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import weakref
from typing import Callable, Optional, Tuple

import torch

from . import _abi
from . import rng as _rng


@dataclasses.dataclass
class ContrastOptions:
    temperature: float = 0.1
    base_temperature: float = 0.07
    max_samples: int = 1024
    max_views: int = 100
    ignore_label: int = -1
    num_classes: Optional[int] = None      # class ids tracked; default: seg planes / bank classes / 256
    normalize: bool = False                # True: `embed` is the raw projection, the engine normalises the sampled columns
    nan_safe: bool = False                 # False reproduces the reference's NaN for rows without positives (Q8)
    rng: str = "device"                    # "device" (no host sync) | "torch_cpu" (reference RNG stream)
    perm_fn: Optional[Callable[[int], torch.Tensor]] = None   # injected permutations (parity tests)
    seed: int = 304
    precision: str = "fp32"                # "fp32" exact SIMT sweep | "bf16" tcgen05 sweep (bank / large problems)
    contrast_norm_bound: float = 1.0       # bound of the contrast rows' L2 norm (tensor path stabiliser; 1 = normalised)
    topk_negatives: Optional[int] = None   # a10 extension: keep only the k hardest negatives per anchor (None = reference)


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """torch.cuda.device(...) only when the tensor's device is not already current (saves ~10 us per call)."""

    __slots__ = ("ctx",)

    def __init__(self, device: torch.device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _abi.PclError(f"{name} must be a CUDA tensor: the engine has no CPU path")


class ContrastWorkspace:
    """Caller-owned scratch for one geometry (torch allocator owns the memory; the C side keeps no state)."""

    def __init__(self, device: torch.device, geom: _abi.Geom, mode: int, bank_K: int, bank_M0: int, bank_M1: int):
        lib = _abi.load()
        self.device = device
        self.geom = geom
        sizes = _abi.SelectSizes()
        _abi.check(lib.pcl_select_sizes(C.byref(geom), C.byref(sizes)), "pcl_select_sizes")
        self.sizes = sizes
        ms, D = geom.max_samples, geom.D
        sw = _abi.SweepDesc()
        sw.a_rows, sw.D, sw.mode = ms, D, mode
        sw.bank_K, sw.bank_M0, sw.bank_M1 = bank_K, bank_M0, bank_M1
        sw.temperature, sw.base_temperature = 1.0, 1.0
        ss = _abi.SweepSizes()
        _abi.check(lib.pcl_sweep_sizes(C.byref(sw), C.byref(ss)), "pcl_sweep_sizes")
        self.sweep_sizes = ss
        n_partial, n_dpartial = ss.partial_f32, ss.dpartial_f32
        self.tc_ok = D == 256
        if self.tc_ok:                       # the tensor sweep uses its own split count: size for the larger of the two
            td = _abi.TcDesc()
            td.a_rows, td.D, td.mode = ms, D, mode
            # the step passes the device-side plan (live anchor count unknown on the host): the partial-slot count
            # depends on that, so size with a non-null plan exactly like the runtime descriptor
            self.plan = torch.zeros(sizes.plan_i32, dtype=torch.int32, device=device)
            td.plan = self.plan.data_ptr()
            td.bank_K, td.bank_R = bank_K, bank_M0 + bank_M1
            td.temperature, td.base_temperature = 1.0, 1.0
            ts = _abi.SweepSizes()
            _abi.check(lib.pcl_tc_sizes(C.byref(td), C.byref(ts)), "pcl_tc_sizes")
            n_partial, n_dpartial = max(n_partial, ts.partial_f32), max(n_dpartial, ts.dpartial_f32)
        i32 = dict(dtype=torch.int32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.keys = torch.empty(sizes.keys_u16, dtype=torch.int16, device=device)
        self.chunk_pref = torch.empty(sizes.chunk_pref_i32, **i32)
        self.counts = torch.empty(sizes.counts_i32, **i32)
        if not hasattr(self, "plan"):
            self.plan = torch.zeros(sizes.plan_i32, **i32)
        self.anchor_meta = torch.empty(sizes.anchor_meta_i32, **i32)
        self.anchors_f32 = torch.empty((ms, D), **f32)
        self.anchors_bf16 = torch.empty((-(-ms // 128) * 128, D), dtype=torch.bfloat16, device=device)
        self.inv_norm = torch.empty(ms, **f32)
        self.norm_max = torch.zeros(1, **f32)
        self.partials = torch.empty(5 * n_partial, **f32)
        self.rowstats = torch.empty(6 * ss.rowstat_f32, **f32)
        self.dpartials = torch.empty(n_dpartial, **f32)
        self.row_m2 = torch.empty(-(-ms // 256) * 256 + 512, **f32)
        self.dA = torch.empty((ms, D), **f32)
        self.loss = torch.zeros(1, **f32)
        self.ranks = torch.zeros(ms, **i32)
        self.ranks_host = torch.zeros(ms, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None
        self.sync = torch.zeros(1024, **i32)      # inter-CTA counters of the fused kernels (re-armed by the kernels themselves)
        self.busy = False
        self.token = 0               # generation counter: a stale autograd node must not release a re-used workspace
        self.bank_key = None         # (device index, segment_queue address) while a backward that re-reads that bank is pending
        self.deferred = []           # bank writes held back until that backward has run (see bank.dequeue_and_enqueue)
        d = _abi.StepDesc()
        d.g = geom
        d.mode, d.bank_K, d.bank_M0, d.bank_M1 = mode, bank_K, bank_M0, bank_M1
        for name in ("keys", "chunk_pref", "counts", "plan", "anchor_meta", "anchors_f32", "anchors_bf16", "inv_norm",
                     "partials", "rowstats", "dpartials", "dA", "loss", "row_m2", "sync"):   # norm_max: optional, unused
            setattr(d, name, getattr(self, name).data_ptr())
        self.desc = d

    # views used by tests / diagnostics
    def plan_header(self):
        return self.plan[:_abi.PLAN_HEADER].tolist()


_WS_CACHE = {}                   # (device index, geometry key) -> [workspaces]; insertion order = age of the geometry
_WS_CACHE_MAX = int(os.environ.get("PCL_WS_CACHE_MAX", "8"))    # geometries kept (multi-scale / random-crop training)


_LAST_WS = {}


def _get_workspace(device, key, geom, mode, bank_K, bank_M0, bank_M1) -> ContrastWorkspace:
    ck = (device.index, key)
    lst = _WS_CACHE.get(ck)
    if lst is None:
        # a new geometry: forget the oldest idle ones first (their scratch goes back to the torch allocator; a
        # workspace still referenced by a pending autograd graph stays alive through that graph)
        if len(_WS_CACHE) >= _WS_CACHE_MAX:
            for old in list(_WS_CACHE):
                if len(_WS_CACHE) < _WS_CACHE_MAX:
                    break
                if not any(w.busy for w in _WS_CACHE[old]):
                    del _WS_CACHE[old]
        lst = _WS_CACHE[ck] = []
    for ws in lst:
        if not ws.busy:
            _LAST_WS[device.index] = ws
            return ws
    ws = ContrastWorkspace(device, geom, mode, bank_K, bank_M0, bank_M1)
    lst.append(ws)
    _LAST_WS[device.index] = ws
    return ws


def clear_workspaces() -> None:
    for key in list(_BANK_READERS):          # bank writes still held back go in before the bookkeeping is dropped
        _flush_bank(*key)
    _WS_CACHE.clear()
    _LAST_WS.clear()


_step_counter = [0]


# ---- banks with a pending reader -------------------------------------------------------------------------------
# The reference's autograd keeps its own copy of the bank for the backward (torch.cat in loss_contrast_mem.py:221), so
# the trainer may overwrite bank rows between the loss and loss.backward() (trainer_contrastive.py:241-255).  The engine
# keeps no copy: its backward sweep re-reads the bank.  A bank write that arrives while such a backward is pending is
# therefore held back and applied right after it (same final bank, same gradient as the reference).
_BANK_READERS = {}               # (device index, segment_queue address) -> [workspaces with a pending backward]


def bank_reader(device_index, segq_ptr):
    """The most recent loss whose pending backward will re-read the bank at this address, or None."""
    lst = _BANK_READERS.get((device_index, segq_ptr))
    return lst[-1] if lst else None


def _run_deferred(ws) -> None:
    calls, ws.deferred = ws.deferred, []
    for fn in calls:
        fn()


def _drop_bank_reader(ws) -> None:
    key, ws.bank_key = ws.bank_key, None
    if key is not None:
        lst = _BANK_READERS.get(key)
        if lst is not None:
            if ws in lst:
                lst.remove(ws)
            if not lst:
                del _BANK_READERS[key]
    _run_deferred(ws)


def _flush_bank(device_index, segq_ptr) -> None:
    """A new forward is about to read this bank: writes still held back for older pending losses go in first."""
    lst = _BANK_READERS.pop((device_index, segq_ptr), None)
    if lst:
        for ws in lst:
            ws.bank_key = None
            _run_deferred(ws)


def _release_workspace(ws, token) -> None:
    if ws.token == token:
        if ws.bank_key is not None or ws.deferred:
            _drop_bank_reader(ws)
        ws.busy = False


def _step_sweep_desc(ws, d) -> "_abi.SweepDesc":
    """pcl_sweep_desc of the step's exact sweep (what pcl_step.cu:fill_sweep builds on the C side)."""
    ms = ws.geom.max_samples
    sw = _abi.SweepDesc()
    sw.anchors = ws.anchors_f32.data_ptr()
    sw.anchor_cls = ws.anchor_meta.data_ptr() + 4 * 2 * ms
    sw.diag_col = ws.anchor_meta.data_ptr() + 4 * 3 * ms
    sw.plan = ws.plan.data_ptr()
    sw.a_rows, sw.D, sw.mode = ms, ws.geom.D, d.mode
    sw.segment_queue, sw.pixel_queue = d.segment_queue, d.pixel_queue
    sw.bank_K, sw.bank_M0, sw.bank_M1 = d.bank_K, d.bank_M0, d.bank_M1
    sw.temperature, sw.base_temperature, sw.nan_safe = d.temperature, d.base_temperature, d.nan_safe
    return sw


def _step_tc_desc(ws, d) -> "_abi.TcDesc":
    """pcl_tc_desc of the step's tensor sweep (what pcl_step.cu:fill_tc builds on the C side)."""
    ms = ws.geom.max_samples
    t = _abi.TcDesc()
    t.anchors_bf16 = ws.anchors_bf16.data_ptr()
    t.anchor_cls = ws.anchor_meta.data_ptr() + 4 * 2 * ms
    t.diag_col = ws.anchor_meta.data_ptr() + 4 * 3 * ms
    t.plan = ws.plan.data_ptr()
    t.a_rows, t.D, t.mode = ms, ws.geom.D, d.mode
    t.contrast_bf16, t.contrast_rows_alloc = d.shadow_bf16, d.shadow_rows
    t.bank_K, t.bank_R, t.sorted = d.bank_K, d.bank_M0 + d.bank_M1, 1
    t.contrast_norm_bound = d.contrast_norm_bound
    t.temperature, t.base_temperature, t.nan_safe = d.temperature, d.base_temperature, d.nan_safe
    return t


def _topk_step_forward(lib, ws, d, opts, stream):
    """a10: selection + gather as usual, then the top-k InfoNCE sweep (exact fp32 path, or the tensor path: the radix
    select then runs in the tcgen05 sweep's epilogue on the fp32-accumulated logits of the bf16 operands)."""
    k = int(opts.topk_negatives)
    if k < 1:
        raise _abi.PclError("topk_negatives must be >= 1 (None disables the selection)")
    g = ws.geom
    if opts.precision == "bf16":
        _abi.check(lib.pcl_select_gather(C.byref(g), d.embed, ws.keys.data_ptr(), ws.chunk_pref.data_ptr(),
                                         ws.plan.data_ptr(), d.ranks, d.seed, d.normalize, ws.anchor_meta.data_ptr(),
                                         ws.anchors_f32.data_ptr(), ws.anchors_bf16.data_ptr(), ws.inv_norm.data_ptr(),
                                         ws.norm_max.data_ptr(), stream), "pcl_select_gather")
        t = _step_tc_desc(ws, d)
        n = lib.pcl_tc_topk_scratch_u32(C.byref(t))
        if n < 0:
            _abi.check(int(n), "pcl_tc_topk_scratch_u32")
        scratch = getattr(ws, "topk_scratch", None)
        if scratch is None or scratch.numel() < n:
            scratch = ws.topk_scratch = torch.empty(n, dtype=torch.int32, device=ws.device)
        _abi.check(lib.pcl_infonce_tc_topk_fwd(C.byref(t), k, scratch.data_ptr(), ws.row_m2.data_ptr(), ws.partials.data_ptr(),
                                               ws.rowstats.data_ptr(), d.loss, stream), "pcl_infonce_tc_topk_fwd")
        return t, k, scratch
    _abi.check(lib.pcl_select_gather(C.byref(g), d.embed, ws.keys.data_ptr(), ws.chunk_pref.data_ptr(),
                                     ws.plan.data_ptr(), d.ranks, d.seed, d.normalize, ws.anchor_meta.data_ptr(),
                                     ws.anchors_f32.data_ptr(), ws.anchors_bf16.data_ptr(), ws.inv_norm.data_ptr(),
                                     ws.norm_max.data_ptr(), stream), "pcl_select_gather")
    sw = _step_sweep_desc(ws, d)
    n = lib.pcl_topk_scratch_u32(C.byref(sw))
    if n < 0:
        _abi.check(int(n), "pcl_topk_scratch_u32")
    scratch = getattr(ws, "topk_scratch", None)
    if scratch is None or scratch.numel() < n:
        scratch = ws.topk_scratch = torch.empty(n, dtype=torch.int32, device=ws.device)
    _abi.check(lib.pcl_infonce_topk_fwd(C.byref(sw), k, scratch.data_ptr(), ws.partials.data_ptr(),
                                        ws.rowstats.data_ptr(), d.loss, stream), "pcl_infonce_topk_fwd")
    return sw, k, scratch


def _topk_step_backward(lib, ws, d, topk, go, stream):
    sw, k, scratch = topk
    if isinstance(sw, _abi.TcDesc):
        _abi.check(lib.pcl_infonce_tc_topk_bwd(C.byref(sw), k, scratch.data_ptr(), ws.row_m2.data_ptr(), ws.rowstats.data_ptr(),
                                               go.data_ptr(), ws.dpartials.data_ptr(), ws.dA.data_ptr(), stream),
                   "pcl_infonce_tc_topk_bwd")
        _abi.check(lib.pcl_scatter_grad(C.byref(ws.geom), ws.plan.data_ptr(), ws.anchor_meta.data_ptr(), ws.dA.data_ptr(),
                                        ws.anchors_f32.data_ptr(), ws.inv_norm.data_ptr(), d.normalize, d.grad_embed,
                                        stream), "pcl_scatter_grad")
        return
    _abi.check(lib.pcl_infonce_topk_bwd(C.byref(sw), k, scratch.data_ptr(), ws.rowstats.data_ptr(), go.data_ptr(),
                                        ws.dpartials.data_ptr(), ws.dA.data_ptr(), stream), "pcl_infonce_topk_bwd")
    _abi.check(lib.pcl_scatter_grad(C.byref(ws.geom), ws.plan.data_ptr(), ws.anchor_meta.data_ptr(), ws.dA.data_ptr(),
                                    ws.anchors_f32.data_ptr(), ws.inv_norm.data_ptr(), d.normalize, d.grad_embed,
                                    stream), "pcl_scatter_grad")


class _PixelContrastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embed, labels, seg, predict, segq, pixq, shadow, opts: ContrastOptions):
        lib = _abi.load()
        _require_cuda(embed, "embed")
        device = embed.device
        if embed.dtype != torch.float32:
            raise _abi.PclError("embed must be float32 (the reference trainer runs without AMP)")
        embed_c = embed if embed.is_contiguous() else embed.contiguous()
        labels_c = labels
        if labels.device != device or labels.dtype != torch.int64 or not labels.is_contiguous():
            labels_c = labels.to(device=device, dtype=torch.int64).contiguous()
        B, D, h, w = embed_c.shape
        if labels_c.dim() != 3 or labels_c.shape[0] != B:
            raise _abi.PclError("labels must be (B, Himg, Wimg)")
        seg_c = pred_c = None
        if seg is not None:
            seg_c = seg.detach()
            if seg_c.dtype != torch.float32 or not seg_c.is_contiguous():
                seg_c = seg_c.to(torch.float32).contiguous()
            if seg_c.shape[0] != B or tuple(seg_c.shape[2:]) != (h, w):
                raise _abi.PclError("seg must be (B, K, h, w) at the embedding resolution")
        elif predict is not None:
            pred_c = predict.to(device=device, dtype=torch.int64).contiguous()
        else:
            raise _abi.PclError("either seg or predict is required")
        mode, bank_K, M0, M1 = 0, 0, 0, 0
        segq_c = pixq_c = None
        if segq is not None:
            mode = 1
            segq_c = segq.detach()
            if segq_c.dtype != torch.float32 or not segq_c.is_contiguous():
                segq_c = segq_c.to(torch.float32).contiguous()
            bank_K, M0 = segq_c.shape[0], segq_c.shape[1]
            if pixq is not None:
                pixq_c = pixq.detach()
                if pixq_c.dtype != torch.float32 or not pixq_c.is_contiguous():
                    pixq_c = pixq_c.to(torch.float32).contiguous()
                M1 = pixq_c.shape[1]
            if segq_c.shape[2] != D:
                raise _abi.PclError("bank feature dim differs from the embedding dim")
            if _BANK_READERS:
                _flush_bank(device.index, segq_c.data_ptr())
        K = opts.num_classes or (seg_c.shape[1] if seg_c is not None else (bank_K if mode == 1 else _abi.MAX_CLASSES))
        if seg_c is not None and seg_c.shape[1] != K:
            raise _abi.PclError("num_classes differs from the number of seg planes")
        key = (B, D, h, w, labels_c.shape[1], labels_c.shape[2], K, opts.max_samples, opts.max_views,
               opts.ignore_label, mode, bank_K, M0, M1)
        lst = _WS_CACHE.get((device.index, key))
        ws = None
        if lst:
            for cand in lst:
                if not cand.busy:
                    ws = cand
                    _LAST_WS[device.index] = ws
                    break
        if ws is None:
            geom = _abi.Geom(B, D, h, w, labels_c.shape[1], labels_c.shape[2], K, opts.max_samples, opts.max_views,
                             opts.ignore_label)
            ws = _get_workspace(device, key, geom, mode, bank_K, M0, M1)
        d = ws.desc
        d.embed, d.labels = embed_c.data_ptr(), labels_c.data_ptr()
        d.seg = _abi.ptr(seg_c)
        d.predict = _abi.ptr(pred_c)
        d.segment_queue, d.pixel_queue = _abi.ptr(segq_c), _abi.ptr(pixq_c)
        d.temperature, d.base_temperature = opts.temperature, opts.base_temperature
        d.nan_safe = int(opts.nan_safe)
        d.normalize = int(opts.normalize)
        d.precision = 0
        shadow_c = None
        if opts.precision == "bf16":
            if not ws.tc_ok:
                raise _abi.PclError("precision='bf16' (tcgen05 sweep) needs proj_dim == 256")
            d.precision = 1
            d.contrast_norm_bound = float(opts.contrast_norm_bound)
            if mode == 1:
                if pixq_c is None or M0 != M1:
                    raise _abi.PclError("the tensor sweep reads the bank as (segment_queue, pixel_queue) of equal size")
                if shadow is None:           # no maintained shadow: rebuild it from the fp32 queues (one extra pass)
                    from .bank import shadow_rows
                    shadow = torch.empty((shadow_rows(bank_K, M0), D), dtype=torch.bfloat16, device=device)
                    with torch.cuda.device(device):
                        _abi.check(lib.pcl_bank_shadow_rebuild(segq_c.data_ptr(), pixq_c.data_ptr(), bank_K, M0, D,
                                                               shadow.data_ptr(), _stream_ptr(device)),
                                   "pcl_bank_shadow_rebuild")
                shadow_c = shadow
                d.shadow_bf16, d.shadow_rows = shadow_c.data_ptr(), shadow_c.shape[0]
        elif opts.precision != "fp32":
            raise _abi.PclError(f"unknown precision {opts.precision!r}")
        _step_counter[0] += 1
        d.seed = (int(opts.seed) * 0x9E3779B97F4A7C15 + _step_counter[0]) & 0xFFFFFFFFFFFFFFFF
        out = torch.empty((), dtype=torch.float32, device=device)     # the loss is written here directly
        d.loss = out.data_ptr()
        with _on_device(device):
            stream = _stream_ptr(device)
            _abi.check(lib.pcl_step_stats(C.byref(d), stream), "pcl_step_stats")
            if opts.perm_fn is not None or opts.rng == "torch_cpu":
                # reference RNG stream: one small D2H copy (B*2K int32), then randperm in the reference's order
                counts = ws.counts.view(B, 2 * K).cpu().numpy()
                pairs, TC, V = _rng.host_plan(counts, opts.max_samples, opts.max_views)
                table = _rng.anchor_rank_table(pairs, V, opts.perm_fn or (lambda n: torch.randperm(n)))
                flat = table[:max(TC, 1), :max(V, 1)].reshape(-1)
                n = min(flat.numel(), ws.ranks.numel())
                ws.ranks_host[:n].copy_(flat[:n])
                ws.ranks[:n].copy_(ws.ranks_host[:n], non_blocking=True)
                d.ranks = ws.ranks.data_ptr()
            elif opts.rng == "device":
                d.ranks = None
            else:
                raise _abi.PclError(f"unknown rng mode {opts.rng!r}")
            if opts.topk_negatives:
                ctx.topk = _topk_step_forward(lib, ws, d, opts, stream)
            else:
                ctx.topk = None
                _abi.check(lib.pcl_step_forward(C.byref(d), stream), "pcl_step_forward")
        ctx.ws = ws
        ctx.embed_shape = tuple(embed_c.shape)
        ctx.keep = (embed_c, labels_c, seg_c, pred_c, segq_c, pixq_c, shadow_c)   # keep inputs alive until kernels ran
        if ctx.needs_input_grad[0]:          # (grad mode is always off inside forward: ask the ctx, not torch)
            ws.busy = True
            ws.token += 1
            ctx.token = ws.token
            if mode == 1:
                ws.bank_key = (device.index, segq_c.data_ptr())
                _BANK_READERS.setdefault(ws.bank_key, []).append(ws)
            # if the graph is dropped without a backward (e.g. a loss that is only logged) give the workspace back
            weakref.finalize(ctx, _release_workspace, ws, ws.token).atexit = False
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _abi.load()
        ws = ctx.ws
        device = ws.device
        # The backward sweep re-reads the workspace scratch and the input tensors of THIS forward.  Both are handed back
        # after the first backward (another forward may have overwritten them since), so a second backward through the
        # same node (retain_graph=True) cannot be served: fail loudly instead of returning a gradient of something else.
        if ctx.keep is None or ws.token != getattr(ctx, "token", None):
            raise _abi.PclError("pixel_contrast_loss: backward ran twice on the same graph (retain_graph=True is not "
                                "supported: the engine keeps no copy of the forward's scratch); call the loss again")
        grad = torch.empty(ctx.embed_shape, dtype=torch.float32, device=device)
        go = grad_out
        if go.device != device or go.dtype != torch.float32 or not go.is_contiguous():
            go = go.detach().to(device=device, dtype=torch.float32).contiguous()
        d = ws.desc
        d.grad_embed = grad.data_ptr()
        with _on_device(device):
            if ctx.topk is not None:
                _topk_step_backward(lib, ws, d, ctx.topk, go, _stream_ptr(device))
            else:
                _abi.check(lib.pcl_step_backward(C.byref(d), go.data_ptr(), _stream_ptr(device)), "pcl_step_backward")
        _release_workspace(ws, getattr(ctx, "token", ws.token))     # also applies bank writes held back for this backward
        ctx.keep = None
        return grad, None, None, None, None, None, None, None


def pixel_contrast_loss(embed: torch.Tensor, labels: torch.Tensor, *, seg: Optional[torch.Tensor] = None,
                        predict: Optional[torch.Tensor] = None, segment_queue: Optional[torch.Tensor] = None,
                        pixel_queue: Optional[torch.Tensor] = None, bank_shadow: Optional[torch.Tensor] = None,
                        options: Optional[ContrastOptions] = None):
    """Pixel-contrast loss (lib/loss/loss_contrast.py:130-147 / loss_contrast_mem.py:154-171) on the GPU engine.

    embed (B,D,h,w) fp32 CUDA; labels (B,Himg,Wimg) int64; either seg (B,K,h,w) logits (argmax fused) or
    predict (B,h,w) int64; optional bank queues (K,M,D).  Returns a 0-dim tensor with autograd to embed."""
    return _PixelContrastFn.apply(embed, labels, seg, predict, segment_queue, pixel_queue, bank_shadow,
                                  options or ContrastOptions())


def last_workspace(embed_device: torch.device):
    """Workspace used by the most recent loss call on a device (diagnostics / tests)."""
    return _LAST_WS.get(embed_device.index)


# ------------------------------------------------------------------------------------------------
# direct InfoNCE on explicit anchors / contrast rows (S4 sweeps, kernel tests)
# ------------------------------------------------------------------------------------------------
def infonce_forward(anchors: torch.Tensor, anchor_cls: torch.Tensor, *, contrast: Optional[torch.Tensor] = None,
                    contrast_cls: Optional[torch.Tensor] = None, queues: Optional[Tuple[torch.Tensor, ...]] = None,
                    diag_col: Optional[torch.Tensor] = None, temperature: float = 0.1, base_temperature: float = 0.07,
                    nan_safe: bool = False, topk: Optional[int] = None):
    """Returns (loss (1,), rowstats (6, A), desc-state) for the exact fp32 sweep.
    Modes: self-contrast (contrast None, queues None), explicit matrix (contrast given), bank (queues given).
    Bank mode expects the anchors grouped by class in the engine's row order (class rank 1, 2, ..., K-1, 0 — what the
    selection kernel produces): the positive sweep only visits the column range of a row tile's classes.
    topk=k keeps only the k hardest negatives of every anchor (a10 extension; None = all, the reference)."""
    lib = _abi.load()
    _require_cuda(anchors, "anchors")
    dev = anchors.device
    a = anchors.detach().to(torch.float32).contiguous()
    A, D = a.shape
    cls = anchor_cls.to(device=dev, dtype=torch.int32).contiguous()
    sw = _abi.SweepDesc()
    sw.anchors, sw.anchor_cls = a.data_ptr(), cls.data_ptr()
    dg = None
    if diag_col is not None:
        dg = diag_col.to(device=dev, dtype=torch.int32).contiguous()
        sw.diag_col = dg.data_ptr()
    sw.a_rows, sw.D = A, D
    keep = [a, cls, dg]
    if queues is not None:
        sw.mode = 1
        sq = queues[0].detach().to(torch.float32).contiguous()
        pq = queues[1].detach().to(torch.float32).contiguous() if len(queues) > 1 and queues[1] is not None else None
        sw.segment_queue, sw.pixel_queue = sq.data_ptr(), _abi.ptr(pq)
        sw.bank_K, sw.bank_M0, sw.bank_M1 = sq.shape[0], sq.shape[1], (pq.shape[1] if pq is not None else 0)
        keep += [sq, pq]
    elif contrast is not None:
        sw.mode = 2
        c = contrast.detach().to(torch.float32).contiguous()
        cc = contrast_cls.to(device=dev, dtype=torch.int32).contiguous()
        sw.contrast, sw.contrast_cls, sw.n_cols = c.data_ptr(), cc.data_ptr(), c.shape[0]
        keep += [c, cc]
    else:
        sw.mode = 0
    sw.temperature, sw.base_temperature, sw.nan_safe = temperature, base_temperature, int(nan_safe)
    ss = _abi.SweepSizes()
    _abi.check(lib.pcl_sweep_sizes(C.byref(sw), C.byref(ss)), "pcl_sweep_sizes")
    partials = torch.empty(5 * ss.partial_f32, dtype=torch.float32, device=dev)
    rowstats = torch.empty(6 * ss.rowstat_f32, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    if topk is not None:
        n = lib.pcl_topk_scratch_u32(C.byref(sw))
        if n < 0:
            _abi.check(int(n), "pcl_topk_scratch_u32")
        scratch = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.pcl_infonce_topk_fwd(C.byref(sw), int(topk), scratch.data_ptr(), partials.data_ptr(),
                                                rowstats.data_ptr(), loss.data_ptr(), _stream_ptr(dev)),
                       "pcl_infonce_topk_fwd")
        return loss, rowstats.view(6, A), (sw, ss, keep, int(topk), scratch)
    with torch.cuda.device(dev):
        _abi.check(lib.pcl_infonce_fwd(C.byref(sw), partials.data_ptr(), rowstats.data_ptr(), loss.data_ptr(),
                                       _stream_ptr(dev)), "pcl_infonce_fwd")
    return loss, rowstats.view(6, A), (sw, ss, keep)


def topk_selection(state, a_rows: int):
    """Selection result of a topk forward (diagnostics/tests): (tau_key uint32 as int64, tie_weight f32, n_above, n_ties)."""
    scratch = state[-1]                      # (exact sweep: 5-tuple, tensor sweep: 6-tuple; the scratch comes last)
    sel = scratch[scratch.numel() - 4 * a_rows:].view(4, a_rows)
    key = sel[0].to(torch.int64) & 0xFFFFFFFF
    return key, sel[1].view(torch.float32), sel[2].to(torch.int64) & 0xFFFFFFFF, sel[3].to(torch.int64) & 0xFFFFFFFF


def infonce_backward(state, rowstats: torch.Tensor, grad_loss: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _abi.load()
    sw, ss, keep = state[:3]
    dev = keep[0].device
    dpart = torch.empty(ss.dpartial_f32, dtype=torch.float32, device=dev)
    dA = torch.empty((sw.a_rows, sw.D), dtype=torch.float32, device=dev)
    if len(state) == 5:                          # forward ran with topk: same scratch (selection result) in the backward
        k, scratch = state[3], state[4]
        with torch.cuda.device(dev):
            _abi.check(lib.pcl_infonce_topk_bwd(C.byref(sw), k, scratch.data_ptr(), rowstats.contiguous().data_ptr(),
                                                _abi.ptr(grad_loss), dpart.data_ptr(), dA.data_ptr(), _stream_ptr(dev)),
                       "pcl_infonce_topk_bwd")
        return dA
    with torch.cuda.device(dev):
        _abi.check(lib.pcl_infonce_bwd(C.byref(sw), rowstats.contiguous().data_ptr(), _abi.ptr(grad_loss),
                                       dpart.data_ptr(), dA.data_ptr(), _stream_ptr(dev)), "pcl_infonce_bwd")
    return dA


# ------------------------------------------------------------------------------------------------
# a1: projection-head normalise (lib/models/modules/projection.py:24)
# ------------------------------------------------------------------------------------------------
class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _abi.load()
        _require_cuda(x, "x")
        xc = x.to(torch.float32).contiguous()
        B, D = xc.shape[0], xc.shape[1]
        HW = xc.numel() // (B * D)
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _abi.check(lib.pcl_l2norm_fwd(xc.data_ptr(), y.data_ptr(), B, D, HW, _stream_ptr(xc.device)), "pcl_l2norm_fwd")
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        lib = _abi.load()
        (xc,) = ctx.saved_tensors
        B, D = xc.shape[0], xc.shape[1]
        HW = xc.numel() // (B * D)
        g = gy.to(torch.float32).contiguous()
        gx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _abi.check(lib.pcl_l2norm_bwd(xc.data_ptr(), g.data_ptr(), gx.data_ptr(), B, D, HW, _stream_ptr(xc.device)),
                       "pcl_l2norm_bwd")
        return gx


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, p=2, dim=1) for (B, D, ...) CUDA tensors on the engine kernel."""
    return _L2NormFn.apply(x)


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) InfoNCE on explicit operands — D must be 256
# ------------------------------------------------------------------------------------------------
def to_bf16_rows(x: torch.Tensor, rows_alloc: int) -> torch.Tensor:
    """fp32 (n, 256) -> bf16 (rows_alloc, 256) zero-padded copy on the engine's conversion kernel."""
    lib = _abi.load()
    xc = x.detach().to(torch.float32).contiguous()
    out = torch.empty((rows_alloc, xc.shape[1]), dtype=torch.bfloat16, device=xc.device)
    with torch.cuda.device(xc.device):
        _abi.check(lib.pcl_to_bf16(xc.data_ptr(), out.data_ptr(), xc.numel(), out.numel(), _stream_ptr(xc.device)),
                   "pcl_to_bf16")
    return out


def _tc_desc(anchors, anchor_cls, contrast_bf16, contrast_cls, n_cols, bank, diag_col, temperature, base_temperature,
             nan_safe, sorted_cols, norm_bound):
    dev = anchors.device
    a = anchors.detach().to(torch.float32).contiguous()
    A, D = a.shape
    a_pad = -(-A // 256) * 256                      # row tiles of the forward sweep
    a16 = torch.empty((-(-A // 128) * 128, D), dtype=torch.bfloat16, device=dev)
    cls = anchor_cls.to(device=dev, dtype=torch.int32).contiguous()
    d = _abi.TcDesc()
    d.anchors_f32, d.anchors_bf16, d.anchor_cls = a.data_ptr(), a16.data_ptr(), cls.data_ptr()
    keep = [a, a16, cls]
    if diag_col is not None:
        dg = diag_col.to(device=dev, dtype=torch.int32).contiguous()
        d.diag_col = dg.data_ptr()
        keep.append(dg)
    d.a_rows, d.D = A, D
    if bank is not None:                       # (shadow bf16, K, R)
        shadow, K, R = bank
        d.mode, d.contrast_bf16, d.bank_K, d.bank_R = 1, shadow.data_ptr(), K, R
        d.contrast_rows_alloc = shadow.shape[0]
        keep.append(shadow)
    elif contrast_bf16 is not None:
        cc = contrast_cls.to(device=dev, dtype=torch.int32).contiguous()
        d.mode, d.contrast_bf16, d.contrast_cls = 2, contrast_bf16.data_ptr(), cc.data_ptr()
        d.n_cols, d.contrast_rows_alloc, d.sorted = n_cols, contrast_bf16.shape[0], int(sorted_cols)
        keep += [contrast_bf16, cc]
    else:
        d.mode = 0
    d.contrast_norm_bound = norm_bound
    d.temperature, d.base_temperature, d.nan_safe = temperature, base_temperature, int(nan_safe)
    return d, keep, a_pad


def infonce_tc_forward(anchors: torch.Tensor, anchor_cls: torch.Tensor, *, contrast_bf16: Optional[torch.Tensor] = None,
                       contrast_cls: Optional[torch.Tensor] = None, n_cols: int = 0, bank=None,
                       diag_col: Optional[torch.Tensor] = None, temperature: float = 0.1, base_temperature: float = 0.07,
                       nan_safe: bool = False, sorted_cols: bool = True, norm_bound: float = 1.0, neg_only: bool = False,
                       topk: Optional[int] = None):
    """bf16 tcgen05 sweep.  Returns (loss (1,), rowstats (6, A), state).  neg_only=True runs just the similarity +
    negative-sum sweep (roofline measurement of the dense contraction).  topk=k: a10, the k hardest negatives per anchor
    (radix select in the sweep's epilogue)."""
    lib = _abi.load()
    _require_cuda(anchors, "anchors")
    dev = anchors.device
    d, keep, a_pad = _tc_desc(anchors, anchor_cls, contrast_bf16, contrast_cls, n_cols, bank, diag_col, temperature,
                              base_temperature, nan_safe, sorted_cols, norm_bound)
    d.neg_only = int(neg_only)
    ss = _abi.SweepSizes()
    _abi.check(lib.pcl_tc_sizes(C.byref(d), C.byref(ss)), "pcl_tc_sizes")
    row_m2 = torch.empty(a_pad + 512, dtype=torch.float32, device=dev)
    partials = torch.empty(5 * ss.partial_f32, dtype=torch.float32, device=dev)
    rowstats = torch.empty(6 * ss.rowstat_f32, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    if topk is not None:
        n = lib.pcl_tc_topk_scratch_u32(C.byref(d))
        if n < 0:
            _abi.check(int(n), "pcl_tc_topk_scratch_u32")
        scratch = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.pcl_infonce_tc_topk_fwd(C.byref(d), int(topk), scratch.data_ptr(), row_m2.data_ptr(),
                                                   partials.data_ptr(), rowstats.data_ptr(), loss.data_ptr(),
                                                   _stream_ptr(dev)), "pcl_infonce_tc_topk_fwd")
        return loss, rowstats.view(6, d.a_rows), (d, ss, keep, row_m2, int(topk), scratch)
    with torch.cuda.device(dev):
        _abi.check(lib.pcl_infonce_tc_fwd(C.byref(d), row_m2.data_ptr(), partials.data_ptr(), rowstats.data_ptr(),
                                          loss.data_ptr(), _stream_ptr(dev)), "pcl_infonce_tc_fwd")
    return loss, rowstats.view(6, d.a_rows), (d, ss, keep, row_m2)


def tc_dump_logits(anchors: torch.Tensor, contrast_bf16: Optional[torch.Tensor], n_cols: int) -> torch.Tensor:
    """Raw S = A.C^T from the TMA/tcgen05 pipeline (self-test of descriptors and barriers)."""
    lib = _abi.load()
    dev = anchors.device
    A = anchors.shape[0]
    cls = torch.zeros(A, dtype=torch.int32, device=dev)
    ccls = torch.zeros(max(n_cols, 1), dtype=torch.int32, device=dev)
    d, keep, a_pad = _tc_desc(anchors, cls, contrast_bf16, ccls if contrast_bf16 is not None else None, n_cols, None,
                              None, 1.0, 1.0, False, False, 1.0)
    ncols = n_cols if contrast_bf16 is not None else A
    ld = -(-ncols // 256) * 256
    dump = torch.zeros((a_pad, ld), dtype=torch.float32, device=dev)
    row_m2 = torch.empty(a_pad + 512, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(lib.pcl_tc_dump_logits(C.byref(d), row_m2.data_ptr(), dump.data_ptr(), _stream_ptr(dev)),
                   "pcl_tc_dump_logits")
    return dump[:A, :ncols]


def infonce_tc_backward(state, rowstats: torch.Tensor, grad_loss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dA (A, 256) fp32 from the tcgen05 backward sweep (recompute S, G tile in shared memory, dA += G.C)."""
    lib = _abi.load()
    d, ss, keep, row_m2 = state[:4]
    dev = row_m2.device
    dpart = torch.empty(ss.dpartial_f32, dtype=torch.float32, device=dev)
    dA = torch.empty((d.a_rows, d.D), dtype=torch.float32, device=dev)
    if len(state) == 6:                          # forward ran with topk: same selection state in the backward
        with torch.cuda.device(dev):
            _abi.check(lib.pcl_infonce_tc_topk_bwd(C.byref(d), state[4], state[5].data_ptr(), row_m2.data_ptr(),
                                                   rowstats.contiguous().data_ptr(), _abi.ptr(grad_loss), dpart.data_ptr(),
                                                   dA.data_ptr(), _stream_ptr(dev)), "pcl_infonce_tc_topk_bwd")
        return dA
    with torch.cuda.device(dev):
        _abi.check(lib.pcl_infonce_tc_bwd(C.byref(d), row_m2.data_ptr(), rowstats.contiguous().data_ptr(),
                                          _abi.ptr(grad_loss), dpart.data_ptr(), dA.data_ptr(), _stream_ptr(dev)),
                   "pcl_infonce_tc_bwd")
    return dA


# ------------------------------------------------------------------------------------------------
# §8f row 1: fused bilinear up-sampling (align_corners=True) + weighted cross-entropy with ignore_index
# (lib/loss/loss_contrast.py:180-181 + lib/loss/loss_helper.py:169-212)
# ------------------------------------------------------------------------------------------------
class _SegCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seg, target, weight, ignore_index):
        lib = _abi.load()
        _require_cuda(seg, "seg")
        dev = seg.device
        seg_c = seg if (seg.dtype == torch.float32 and seg.is_contiguous()) else seg.to(torch.float32).contiguous()
        tgt = target
        if tgt.device != dev or tgt.dtype != torch.int64 or not tgt.is_contiguous():
            tgt = tgt.to(device=dev, dtype=torch.int64).contiguous()
        B, K, h, w = seg_c.shape
        H, W = tgt.shape[1], tgt.shape[2]
        wt = None
        if weight is not None:
            wt = weight.to(device=dev, dtype=torch.float32).contiguous()
        scratch = torch.empty(lib.pcl_seg_ce_scratch_floats(B, H, W), dtype=torch.float32, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with _on_device(dev):
            _abi.check(lib.pcl_seg_ce_fwd(seg_c.data_ptr(), tgt.data_ptr(), _abi.ptr(wt), B, K, h, w, H, W, int(ignore_index),
                                          scratch.data_ptr(), out.data_ptr(), _stream_ptr(dev)), "pcl_seg_ce_fwd")
        ctx.keep = (seg_c, tgt, wt, scratch)
        ctx.dims = (B, K, h, w, H, W, int(ignore_index))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _abi.load()
        seg_c, tgt, wt, scratch = ctx.keep
        B, K, h, w, H, W, ign = ctx.dims
        dev = seg_c.device
        go = grad_out
        if go.device != dev or go.dtype != torch.float32 or not go.is_contiguous():
            go = go.detach().to(device=dev, dtype=torch.float32).contiguous()
        dseg = torch.empty_like(seg_c)
        with _on_device(dev):
            _abi.check(lib.pcl_seg_ce_bwd(seg_c.data_ptr(), tgt.data_ptr(), _abi.ptr(wt), B, K, h, w, H, W, ign,
                                          scratch.data_ptr(), go.data_ptr(), dseg.data_ptr(), _stream_ptr(dev)),
                       "pcl_seg_ce_bwd")
        ctx.keep = None
        return dseg, None, None, None


def upsample_cross_entropy(seg: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor] = None,
                           ignore_index: int = -1) -> torch.Tensor:
    """mean CE of bilinearly (align_corners=True) up-sampled logits against a full-resolution label map, fused:
    == F.cross_entropy(F.interpolate(seg, target.shape[1:], mode='bilinear', align_corners=True), target,
                       weight, ignore_index=ignore_index)  without the (B,K,Himg,Wimg) intermediate."""
    return _SegCeFn.apply(seg, target, weight, ignore_index)
