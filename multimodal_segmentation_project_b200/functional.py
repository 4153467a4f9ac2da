"""torch-facing wrappers and autograd Functions over the libb200unet C ABI.

Activations inside the network are channels-last ``[N, D, H, W, C]`` tensors (fp32 or bf16);
the reference's NCDHW tensors exist only at the input (``in_channels == 1`` makes the two
layouts identical) and at the logits produced by :func:`final_conv1x1`.

Every function requires CUDA tensors: there is no CPU path (the oracle under ``oracle/`` is
test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref

import torch

from . import _lib
from ._lib import check

_VP = ctypes.c_void_p


def _ptr(t):
    return None if t is None else _VP(t.data_ptr())


def _stream():
    return _VP(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.B200_F32
    if t.dtype == torch.bfloat16:
        return _lib.B200_BF16
    raise TypeError(f"libb200unet supports float32 and bfloat16 activations, got {t.dtype}")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "multimodal_segmentation_project_b200 runs on CUDA (sm_100a) tensors only; got a "
                f"{t.device} tensor. There is no CPU fallback."
            )


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


# --------------------------------------------------------------------------- plain wrappers
def to_channels_last(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """NCDHW fp32 -> NDHWC ``dtype``."""
    _require_cuda(x)
    L = _lib.load()
    if x.dtype == torch.bfloat16 and dtype == torch.bfloat16 and x.shape[1] == 1:
        # a bf16 single-channel volume already IS the channels-last bf16 tensor (e.g. batches shipped over PCIe as bf16)
        return x.detach().contiguous().reshape(x.shape[0], *x.shape[2:], 1)
    x = x.detach().float().contiguous()
    N, C = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:])
    S = 1
    for d in sp:
        S *= d
    y = torch.empty((N, *sp, C), dtype=dtype, device=x.device)
    if C == 1:
        if dtype == torch.float32:
            return x.reshape(N, *sp, 1)
        check(L.b200_cast_from_f32(_dt(y), _ptr(x), _ptr(y), x.numel(), _stream()), "cast_from_f32")
        return y
    check(L.b200_ncdhw_to_ndhwc(_dt(y), _ptr(x), _ptr(y), N, C, S, _stream()), "ncdhw_to_ndhwc")
    return y


def to_channels_first_f32(x: torch.Tensor) -> torch.Tensor:
    """NDHWC (fp32/bf16) -> NCDHW fp32."""
    _require_cuda(x)
    L = _lib.load()
    x = x.detach().contiguous()
    N, C = x.shape[0], x.shape[-1]
    sp = tuple(x.shape[1:-1])
    S = 1
    for d in sp:
        S *= d
    y = torch.empty((N, C, *sp), dtype=torch.float32, device=x.device)
    check(L.b200_ndhwc_to_ncdhw(_dt(x), _ptr(x), _ptr(y), N, C, S, _stream()), "ndhwc_to_ncdhw")
    return y


class _ToChannelsLast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        return to_channels_last(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return to_channels_first_f32(g), None


# ---- packed tcgen05 weight layouts: cached per parameter, refreshed in ONE launch per optimiser step -------------------------
# The conv kernels read weights in a bf16 UMMA layout derived from the fp32 nn.Parameter.  Deriving it per call costs one small
# dependent kernel per layer and direction (34 per train step).  Cache entries are keyed on the parameter OBJECT and validated
# against (data_ptr, torch's in-place version counter, this module's weights epoch): torch-side mutation (optimizer.step of a
# torch optimizer, load_state_dict, ...) bumps _version; kernels of this library that write parameters through raw pointers
# (FlatAdamW) bump the epoch with weights_changed().  A trainer then calls repack_cached_weights() once: one batched kernel
# rewrites every cached layout in place (stable addresses: safe inside a captured CUDA graph).
_weights_epoch = 0
_pack_cache = {}          # (id(param), mode) -> dict(ref, packed, ptr, version, epoch, Cout, Cin)
_pack_job_tables = {}     # subset tag -> (key tuple, device job table, njobs, total_groups)


def weights_changed() -> None:
    """Call after a kernel of this library wrote parameters through raw pointers (no torch version bump)."""
    global _weights_epoch
    _weights_epoch += 1


def _pack_now(L, w, mode, dt, out, Cout, Cin):
    check(L.b200_pack_conv3_weights(mode, dt, _ptr(w), _ptr(out), Cout, Cin, _stream()), "pack_conv3_weights")


def pack_conv3_weights(weight: torch.Tensor, mode: int, dtype: torch.dtype) -> torch.Tensor:
    L = _lib.load()
    Cout, Cin = weight.shape[0], weight.shape[1]
    dt = _lib.B200_F32 if dtype == torch.float32 else _lib.B200_BF16
    cacheable = (mode in (_lib.PACK_FPROP_TC, _lib.PACK_DGRAD_TC) and isinstance(weight, torch.nn.Parameter) and weight.dtype == torch.float32
                 and weight.is_contiguous() and _cache_packed_weights)
    if not cacheable:
        w = _f32(weight)
        out = torch.empty(L.b200_pack_conv3_bytes(mode, dt, Cout, Cin), dtype=torch.uint8, device=w.device)
        _pack_now(L, w, mode, dt, out, Cout, Cin)
        return out
    key = (id(weight), mode)
    e = _pack_cache.get(key)
    if e is not None and (e["ref"]() is not weight or e["ptr"] != weight.data_ptr()):
        e = None                                   # a different tensor behind the same id, or the parameter was re-homed
    if e is None:
        out = torch.empty(L.b200_pack_conv3_bytes(mode, dt, Cout, Cin), dtype=torch.uint8, device=weight.device)
        e = {"ref": weakref.ref(weight, lambda _r, k=key: _pack_cache.pop(k, None)), "packed": out, "ptr": weight.data_ptr(), "version": None,
             "epoch": None, "Cout": Cout, "Cin": Cin, "mode": mode}
        _pack_cache[key] = e
        _pack_job_tables.clear()
    if e["version"] != weight._version or e["epoch"] != _weights_epoch:
        _pack_now(L, weight.detach(), mode, dt, e["packed"], Cout, Cin)
        e["version"], e["epoch"] = weight._version, _weights_epoch
    return e["packed"]


def repack_cached_weights(ptr_range=None, inside=True) -> int:
    """Rewrites every cached packed layout from its parameter's CURRENT values in one kernel launch and marks the entries fresh.
    ptr_range = (lo, hi): only the parameters whose storage starts inside (inside=True) or outside (inside=False) that address
    range — a trainer repacks the buckets it has already stepped while the last bucket's all-reduce is in flight.
    Returns the number of layouts written (0 = nothing to do, no launch)."""
    L = _lib.load()
    live = [(k, e) for k, e in _pack_cache.items() if e["ref"]() is not None]
    if ptr_range is not None:
        lo, hi = ptr_range
        live = [(k, e) for k, e in live if (lo <= e["ptr"] < hi) == bool(inside)]
    if not live:
        return 0
    tag = None if ptr_range is None else (ptr_range[0], ptr_range[1], bool(inside))
    sig = tuple((k, e["ptr"], e["packed"].data_ptr()) for k, e in live)
    _pack_jobs = _pack_job_tables.get(tag)
    if _pack_jobs is None or _pack_jobs[0] != sig:
        import struct
        buf, g = bytearray(), 0
        for _, e in live:
            buf += struct.pack("<QQiiiiq", e["ptr"], e["packed"].data_ptr(), e["Cout"], e["Cin"], int(e["mode"] == _lib.PACK_DGRAD_TC), 0, g)
            g += 27 * e["Cin"] * e["Cout"] // 8
        buf += struct.pack("<QQiiiiq", 0, 0, 0, 0, 0, 0, g)
        dev = live[0][1]["packed"].device
        table = torch.frombuffer(buf, dtype=torch.uint8).clone().to(dev)
        _pack_jobs = _pack_job_tables[tag] = (sig, table, len(live), g)
    _, table, n, total = _pack_jobs
    check(L.b200_pack_conv3_batched(_ptr(table), n, total, _stream()), "pack_conv3_batched")
    for _, e in live:
        w = e["ref"]()
        e["version"], e["epoch"] = w._version, _weights_epoch
    return n


def cached_weights_stale() -> bool:
    """True if a cached layout no longer matches its parameter as far as the host can tell (torch version counter, epoch).  A
    captured graph does not contain per-layer pack kernels any more, so a trainer checks this before every replay and repacks."""
    for e in _pack_cache.values():
        w = e["ref"]()
        if w is not None and (e["version"] != w._version or e["epoch"] != _weights_epoch or e["ptr"] != w.data_ptr()):
            return True
    return False


_cache_packed_weights = os.environ.get("B200_CACHE_PACKED_WEIGHTS", "1") != "0"


def set_cache_packed_weights(on: bool) -> None:
    global _cache_packed_weights
    _cache_packed_weights = bool(on)


def conv3d_select_impl(x0, x1, co0, co1, impl=0) -> int:
    """1 = CUDA-core implicit GEMM, 2 = tcgen05; resolves 0 (auto) for this problem."""
    L = _lib.load()
    N, D, H, W, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    return L.b200_conv3d_k3_select(_dt(x0), impl, c0, c1, co0, co1, N, D, H, W)


def pack_mode(impl: int, dgrad: bool) -> int:
    if impl == 2:
        return _lib.PACK_DGRAD_TC if dgrad else _lib.PACK_FPROP_TC
    return _lib.PACK_DGRAD if dgrad else _lib.PACK_FPROP


def conv3d_k3_raw(x0, x1, wpack, bias, co0, co1=0, impl=1):
    """3x3x3 'same' convolution of the virtual concat [x0|x1]; returns (y0, y1) with co0 / co1 channels."""
    L = _lib.load()
    N, D, H, W, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    y0 = torch.empty((N, D, H, W, co0), dtype=x0.dtype, device=x0.device)
    y1 = torch.empty((N, D, H, W, co1), dtype=x0.dtype, device=x0.device) if co1 else None
    check(
        L.b200_conv3d_k3(_dt(x0), impl, _ptr(x0), c0, _ptr(x1), c1, _ptr(wpack), _ptr(bias), _ptr(y0), co0, _ptr(y1), co1,
                         N, D, H, W, _stream()),
        "conv3d_k3",
    )
    return y0, y1


# Gradient sinks (train_unet.py:226): a trainer that keeps all gradients in one flat fp32 buffer (dp.FlatParams)
# registers, per parameter, the slice that belongs to it.  The backward functions below then hand that slice to their kernels as
# the destination of the parameter gradient and return None for the input, so autograd neither stores nor copies anything: no
# per-parameter gradient tensors, no multi-tensor gather, no zero fills in the step.  Keys are parameter data pointers (the
# tensors autograd hands back in ctx.saved_tensors are the parameters' storage, whatever Python object wraps them).
_grad_sinks = {}       # param.data_ptr() -> fp32 view (parameter shape) of the flat gradient buffer
_sunk = set()          # data pointers whose gradient was written to its sink since the last begin_grad_sinks()
_bwd_callbacks = {}    # param.data_ptr() -> callable, run when the backward of the node owning that parameter starts


def begin_grad_sinks(sinks, callbacks=None) -> None:
    """Activates `sinks` ({param.data_ptr(): flat-gradient view}) for the backward passes that follow; `callbacks` maps a
    parameter's data pointer to a function called at the START of the backward of the node that owns it (every node that ran
    before it has launched its gradient kernels: what DDP's bucket-ready hooks observe, train_unet.py:384)."""
    global _grad_sinks, _bwd_callbacks
    _grad_sinks = sinks
    _bwd_callbacks = callbacks or {}
    _sunk.clear()


def end_grad_sinks() -> None:
    global _grad_sinks, _bwd_callbacks
    _grad_sinks = {}
    _bwd_callbacks = {}


def grad_was_sunk(ptr: int) -> bool:
    return ptr in _sunk


def _grad_dst(ptr, shape, device):
    """(fp32 tensor the kernel writes a parameter gradient to, True if it is the registered sink of the parameter at `ptr`)."""
    v = _grad_sinks.get(ptr) if ptr else None
    if v is not None and v.numel() == math.prod(shape) and v.data_ptr() % 16 == 0:
        _sunk.add(ptr)
        return v.view(shape), True
    return torch.empty(shape, dtype=torch.float32, device=device), False


def _run_bwd_callback(ptr) -> None:
    cb = _bwd_callbacks.get(ptr) if _bwd_callbacks else None
    if cb is not None:
        cb()


def conv3d_wgrad_raw(x0, x1, dy, want_bias=True, side=None, dw_out=None, db_out=None):
    """side: a stream already ordered after the producers of x / dy (see fork_side); the launch then goes there and the
    call returns (dw, db, workspace) — the caller keeps `workspace` alive until it has joined the side stream.
    dw_out / db_out: destinations (fp32, torch layout) instead of fresh tensors."""
    L = _lib.load()
    N, D, H, W, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    Cout = dy.shape[-1]
    ws_bytes = L.b200_conv3d_wgrad_workspace(c0, c1, Cout, N, D, H, W)
    # all buffers come from the CURRENT stream's pool, whichever stream the kernels run on
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x0.device)
    dw = dw_out if dw_out is not None else torch.empty((Cout, c0 + c1, 3, 3, 3), dtype=torch.float32, device=x0.device)
    db = (db_out if db_out is not None else torch.empty(Cout, dtype=torch.float32, device=x0.device)) if want_bias else None
    check(
        L.b200_conv3d_wgrad(_dt(x0), _ptr(x0), c0, _ptr(x1), c1, _ptr(dy), Cout, _ptr(dw), _ptr(db), _ptr(ws), ws_bytes,
                            N, D, H, W, _stream() if side is None else side.cuda_stream),
        "conv3d_wgrad",
    )
    return (dw, db) if side is None else (dw, db, ws)


# The weight gradient and the data gradient of a layer read the same dY and are independent: the weight gradient runs on
# a side stream (fork after dY is produced, join before backward() returns).  The deep layers (16^3, 8^3 voxels) launch
# fewer CTAs than the GPU has SMs, so the two kernels genuinely run side by side; under CUDA-graph capture the fork/join
# becomes two parallel branches of the graph.
_side_streams = {}
_overlap_wgrad = os.environ.get("B200_OVERLAP_WGRAD", "1") != "0"


def fork_side(device):
    """Returns a side stream that waits for everything enqueued so far on the current stream (None if disabled)."""
    if not _overlap_wgrad:
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device())
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    return side


def join_side(side, device, *keep):
    """Joins the side stream now — or, in deferred mode, at the end of the running backward pass (the weight gradients
    then also overlap the following layers' HBM-bound BatchNorm kernels); `keep` = tensors the side-stream kernels still
    read or write, held until the join so that their memory is not handed out again on the main stream."""
    if side is None:
        return
    if not _defer_join:
        torch.cuda.current_stream(device).wait_stream(side)
        return
    _pending.append((side, device, keep))
    torch.autograd.Variable._execution_engine.queue_callback(join_pending)


# Deferred mode is only safe when nothing consumes a parameter gradient on the main stream before the end of backward
# (p.grad is None, so autograd's AccumulateGrad just stores the tensor): DataParallelTrainer's detached-gradient mode.
_defer_join = False
_pending = []


def set_deferred_wgrad_join(on: bool) -> None:
    global _defer_join
    join_pending()
    _defer_join = bool(on)


def order_after_pending(stream) -> None:
    """Orders `stream` after every outstanding side-stream weight gradient WITHOUT joining them into the current stream (the held
    tensors stay held): a communication stream that must see a bucket's gradients while the main chain runs on."""
    seen = set()
    for side, _, _ in _pending:
        if id(side) not in seen:
            seen.add(id(side))
            stream.wait_stream(side)


def join_pending() -> None:
    """Orders the current stream after every outstanding side-stream weight gradient and releases the held tensors."""
    seen = set()
    for side, device, _ in _pending:
        if id(side) not in seen:
            seen.add(id(side))
            torch.cuda.current_stream(device).wait_stream(side)
    _pending.clear()


def set_wgrad_impl(impl: int) -> None:
    """0 auto, 1 CUDA-core split-K, 2 tcgen05 (process-wide; tests and benchmarks)."""
    check(_lib.load().b200_set_wgrad_impl(int(impl)), "set_wgrad_impl")


def set_wgrad_pair(on: bool, dseg: int = 0) -> None:
    """Voxel-pair tcgen05 weight gradient for the 16-channel layers (default on); dseg > 0 forces the d-run per CTA (tests, A/B timing)."""
    check(_lib.load().b200_set_wgrad_pair(int(bool(on)), int(dseg)), "set_wgrad_pair")


def set_conv_persistent(mode: int) -> None:
    """0 never, 1 auto (every layer with enough tiles), 2 same as 1, 3 only 16->16 layers (process-wide; tests and benchmarks)."""
    check(_lib.load().b200_set_conv_persistent(int(mode)), "set_conv_persistent")


_fuse_bn_stats = os.environ.get("B200_FUSE_BN_STATS", "1") != "0"


def set_fuse_bn_stats(on: bool) -> None:
    """Batch statistics from the convolution epilogue (default) or from the separate bn_stats pass (tests / A-B timing)."""
    global _fuse_bn_stats
    _fuse_bn_stats = bool(on)


def set_conv_rowstream(on: bool) -> None:
    """Row-streaming tcgen05 convolution for full-resolution 16 / 32-channel layers (default on; tests and A/B timing)."""
    check(_lib.load().b200_set_conv_rowstream(int(bool(on))), "set_conv_rowstream")


def _bn_partials(C, device, rows=0):
    """Partial-sum buffer [rows][2][C] of the statistics kernels (at least the b200_bn_stats row count; the first layer's fused
    statistics write one row per CTA, more than that)."""
    L = _lib.load()
    return torch.empty(max(L.b200_bn_partials_bytes(C) // 4, rows * 2 * C), dtype=torch.float32, device=device)


# --------------------------------------------------------------------------- conv + BN + ReLU + Dropout3d
class _ConvBNAct(torch.autograd.Function):
    """One half of the reference's DoubleConv (models/unet.py:11-14 / :15-18):
    Conv3d(k3,p1)+bias -> BatchNorm3d -> ReLU -> Dropout3d (channel mask supplied by the caller).
    The input may be the virtual concat of two tensors (models/unet.py:84)."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, gamma, beta, running_mean, running_var, nbt, dropmask, training, eps,
                momentum, impl, prev_conv_out=None, prev_stats=None):
        # prev_*: x0 == relu(bn(prev_conv_out)) of the previous _ConvBNAct (no dropout) — lets backward fold that layer's
        # BatchNorm-backward reduction into this layer's data-gradient kernel
        _require_cuda(x0, x1, weight)
        L = _lib.load()
        ctx.impl_req = impl
        x0 = x0.contiguous()
        x1 = None if x1 is None else x1.contiguous()
        N, D, H, W, _ = x0.shape
        Cout = weight.shape[0]
        dev = x0.device
        # the statistics kernels write fp32 / int64 through raw pointers: anything else (model.half(), .to(bfloat16)) would
        # overflow the buffers, so it is an error instead of a cast
        for name, buf, want in (("running_mean", running_mean, torch.float32), ("running_var", running_var, torch.float32),
                                ("num_batches_tracked", nbt, torch.int64)):
            if buf is not None and (buf.dtype != want or not buf.is_contiguous() or not buf.is_cuda):
                raise TypeError(f"BatchNorm3d.{name} must be a contiguous CUDA {want} tensor (got {buf.dtype} on {buf.device}): keep "
                                "the module in float32 and select bf16 compute with torch.autocast or model.compute_dtype")
        impl = conv3d_select_impl(x0, x1, Cout, 0, impl)
        wpack = pack_conv3_weights(weight, pack_mode(impl, False), x0.dtype)
        M, S = N * D * H * W, D * H * W
        if training and M <= 1:  # same contract as torch.nn.functional.batch_norm
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {[N, Cout, D, H, W]}")
        stats = torch.empty((4, Cout), dtype=torch.float32, device=dev)  # scale, shift, mean, invstd
        g32, b32, bias32 = _f32(gamma), _f32(beta), _f32(bias)
        c0, c1 = x0.shape[-1], (0 if x1 is None else x1.shape[-1])
        rows = L.b200_conv3d_k3_bnstats_blocks(_dt(x0), impl, c0, c1, Cout, 0, N, D, H, W) if (training and _fuse_bn_stats) else 0
        if rows > 0:
            # convolution and batch statistics in ONE kernel: the conv epilogue emits per-CTA (sum, sum of squares) of (y - bias)
            partials = _bn_partials(Cout, dev, rows)
            conv_out = torch.empty((N, D, H, W, Cout), dtype=x0.dtype, device=dev)
            check(L.b200_conv3d_k3_bnstats(_dt(x0), impl, _ptr(x0), c0, _ptr(x1), c1, _ptr(wpack), _ptr(bias32), _ptr(conv_out), Cout,
                                           N, D, H, W, _ptr(partials), _stream()), "conv3d_k3_bnstats")
            check(L.b200_bn_finalize_ex(_ptr(partials), rows, _ptr(bias32), M, Cout, _ptr(g32), _ptr(b32), float(eps), float(momentum),
                                        _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                        _ptr(stats[3]), _stream()), "bn_finalize_ex")
        else:
            conv_out, _ = conv3d_k3_raw(x0, x1, wpack, bias32, Cout, 0, impl)
            partials = None
            if training:
                partials = _bn_partials(Cout, dev)
                check(L.b200_bn_stats(_dt(conv_out), _ptr(conv_out), M, Cout, _ptr(partials), _stream()), "bn_stats")
            check(
                L.b200_bn_finalize(_dt(conv_out), _ptr(conv_out), _ptr(partials), M, Cout, _ptr(g32), _ptr(b32), float(eps), float(momentum), int(training),
                                   _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(stats[0]), _ptr(stats[1]),
                                   _ptr(stats[2]), _ptr(stats[3]), _stream()),
                "bn_finalize",
            )
        y = torch.empty_like(conv_out)
        if getattr(ctx, "_b200_want_pool", False):
            # _ConvBNActSkipPool: the apply pass also writes MaxPool3d(2,2)(y) (one read of conv_out, no re-read of y)
            pooled = torch.empty((N, D // 2, H // 2, W // 2, Cout), dtype=y.dtype, device=dev)
            check(L.b200_bn_act_pool_fwd(_dt(conv_out), _ptr(conv_out), _ptr(y), _ptr(pooled), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                         N, D, H, W, Cout, _stream()), "bn_act_pool_fwd")
            ctx._b200_pooled = pooled
        else:
            check(
                L.b200_bn_act_fwd(_dt(conv_out), _ptr(conv_out), _ptr(y), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(dropmask), 1, N, S,
                                  Cout, _stream()),
                "bn_act_fwd",
            )
        ctx.save_for_backward(x0, x1, weight, conv_out, stats, dropmask, prev_conv_out, prev_stats)
        ctx.training = bool(training)
        ctx.has_bias = bias is not None
        ctx.bias_ptr = bias.data_ptr() if bias is not None else 0
        ctx.gamma_ptr, ctx.beta_ptr = gamma.data_ptr(), beta.data_ptr()
        ctx.mark_non_differentiable(conv_out, stats)
        ctx.set_materialize_grads(False)      # no zero-filled gradient tensors for the two pass-through outputs
        return y, conv_out, stats

    @staticmethod
    def backward(ctx, gy, _gconv=None, _gstats=None):
        L = _lib.load()
        x0, x1, weight, conv_out, stats, dropmask, prev_conv_out, prev_stats = ctx.saved_tensors[:8]
        if gy is None:                        # (materialize_grads is off) nothing flows back through this block
            return (None,) * 16
        _run_bwd_callback(weight.data_ptr())
        gy = gy.contiguous()
        N, D, H, W, Cout = conv_out.shape
        M, S = N * D * H * W, D * H * W
        dev = gy.device
        dt = _dt(conv_out)
        dgamma, g_sunk = _grad_dst(ctx.gamma_ptr, (Cout,), dev)
        dbeta, b_sunk = _grad_dst(ctx.beta_ptr, (Cout,), dev)
        sums = torch.empty(2 * Cout, dtype=torch.float32, device=dev)
        hit = _bnbwd_handoff.pop(conv_out.data_ptr(), None)
        if hit is not None and hit[2] == gy.data_ptr() and dropmask is None:
            # the kernel that produced gy (the next layer's data gradient) already reduced it against conv_out
            check(L.b200_bn_bwd_finalize_ex(_ptr(hit[0]), hit[1], M, Cout, _ptr(dgamma), _ptr(dbeta), _ptr(sums), _stream()), "bn_bwd_finalize_ex")
        else:
            partials = _bn_partials(Cout, dev)
            check(
                L.b200_bn_act_bwd_reduce(dt, _ptr(gy), _ptr(conv_out), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]),
                                         _ptr(dropmask), 1, N, S, Cout, _ptr(partials), _stream()),
                "bn_act_bwd_reduce",
            )
            check(L.b200_bn_bwd_finalize(_ptr(partials), M, Cout, _ptr(dgamma), _ptr(dbeta), _ptr(sums), _stream()), "bn_bwd_finalize")
        dconv = torch.empty_like(conv_out)
        check(
            L.b200_bn_act_bwd_apply(dt, _ptr(gy), _ptr(conv_out), _ptr(dconv), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                    _ptr(stats[3]), _ptr(dropmask), 1, _ptr(sums), int(ctx.training), N, S, Cout, _stream()),
            "bn_act_bwd_apply",
        )
        dx0, dx1, dw, db = _conv_bwd_from_dconv(ctx, x0, x1, weight, dconv, zero_bias_grad=ctx.training, prev=(prev_conv_out, prev_stats))
        return (dx0, dx1, dw, db, None if g_sunk else dgamma, None if b_sunk else dbeta, None, None, None, None, None, None, None, None, None, None)


# (partials, rows, gy.data_ptr()) of a BatchNorm-backward reduction computed by the kernel that produced gy, keyed by the data
# pointer of the conv_out it was taken against; written by the consumer layer's backward, popped by the producer layer's backward
# within the same backward pass.
_bnbwd_handoff = {}
_fuse_bn_bwd = os.environ.get("B200_FUSE_BN_BWD", "1") != "0"


def set_fuse_bn_bwd(on: bool) -> None:
    """BatchNorm-backward reduction inside the data-gradient kernel of the following conv (default) or as its own pass."""
    global _fuse_bn_bwd
    _fuse_bn_bwd = bool(on)


def _conv_bwd_from_dconv(ctx, x0, x1, weight, dconv, zero_bias_grad, prev=(None, None)):
    """Weight / bias / data gradients of the 3x3x3 convolution given d loss / d conv_out (shared by _ConvBNAct and _ConvStats).
    zero_bias_grad: the bias feeds a batch-statistics BatchNorm, its gradient is exactly zero.
    prev = (conv_out, stats) of the layer whose relu(bn(.)) is x0: its BatchNorm-backward sums ride on the data-gradient kernel."""
    dev = dconv.device
    Cout = dconv.shape[-1]
    dw = db = None
    need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
    need_x = ctx.needs_input_grad[0] or (x1 is not None and ctx.needs_input_grad[1])
    side = keep = None
    if need_w:
        side = fork_side(dev) if need_x else None
        dw_dst, w_sunk = _grad_dst(weight.data_ptr(), tuple(weight.shape), dev)
        bias_ptr = getattr(ctx, "bias_ptr", 0)
        b_sunk = False
        db_dst = None
        if not zero_bias_grad and ctx.has_bias:
            db_dst, b_sunk = _grad_dst(bias_ptr, (Cout,), dev)
        if side is None:
            dw, db = conv3d_wgrad_raw(x0, x1, dconv, want_bias=not zero_bias_grad, dw_out=dw_dst, db_out=db_dst)
        else:
            dw, db, keep = conv3d_wgrad_raw(x0, x1, dconv, want_bias=not zero_bias_grad, side=side, dw_out=dw_dst, db_out=db_dst)
        if w_sunk:
            dw = None
        if not ctx.has_bias or b_sunk:
            db = None
        elif db is None:
            # the bias feeds a batch-statistics BatchNorm: its gradient is exactly zero.  With a registered sink the flat slice is
            # simply left at zero (the trainer zeroes it once); otherwise autograd gets a zero tensor, as the reference's does
            db = None if bias_ptr in _grad_sinks else torch.zeros(Cout, dtype=torch.float32, device=dev)
    dx0 = dx1 = None
    if need_x:
        c0 = x0.shape[-1]
        c1 = 0 if x1 is None else x1.shape[-1]
        impl = conv3d_select_impl(dconv, None, c0, c1, ctx.impl_req)
        wpack = pack_conv3_weights(weight, pack_mode(impl, True), dconv.dtype)
        L = _lib.load()
        N, D, H, W, _ = dconv.shape
        pco, pst = prev
        rows = 0
        if pco is not None and c1 == 0 and _fuse_bn_bwd and pco.shape == x0.shape and pco.dtype == dconv.dtype:
            rows = L.b200_conv3d_k3_bnbwd_blocks(_dt(dconv), impl, Cout, c0, N, D, H, W)
        if rows > 0:
            dx0 = torch.empty_like(x0)
            partials = _bn_partials(c0, dev)
            check(L.b200_conv3d_k3_bnbwd(_dt(dconv), impl, _ptr(dconv), Cout, _ptr(wpack), _ptr(dx0), c0, N, D, H, W, _ptr(pco), _ptr(pst[0]),
                                         _ptr(pst[1]), _ptr(pst[2]), _ptr(pst[3]), _ptr(partials), _stream()), "conv3d_k3_bnbwd")
            _bnbwd_handoff[pco.data_ptr()] = (partials, rows, dx0.data_ptr())
        else:
            dx0, dx1 = conv3d_k3_raw(dconv, None, wpack, None, c0, c1, impl)
    join_side(side, dev, x0, x1, dconv, keep)
    del keep
    return dx0, dx1, dw, db


def _conv_and_batch_stats(L, x0, x1, weight, bias, gamma, beta, running_mean, running_var, nbt, eps, momentum, impl):
    """conv3x3x3 + training-mode BatchNorm statistics (fused into the conv epilogue where a kernel offers it) + finalize.
    Returns (conv_out, stats[4, C] = scale, shift, mean, invstd)."""
    N, D, H, W, _ = x0.shape
    Cout = weight.shape[0]
    dev = x0.device
    M = N * D * H * W
    wpack = pack_conv3_weights(weight, pack_mode(impl, False), x0.dtype)
    stats = torch.empty((4, Cout), dtype=torch.float32, device=dev)
    g32, b32, bias32 = _f32(gamma), _f32(beta), _f32(bias)
    c0, c1 = x0.shape[-1], (0 if x1 is None else x1.shape[-1])
    rows = L.b200_conv3d_k3_bnstats_blocks(_dt(x0), impl, c0, c1, Cout, 0, N, D, H, W) if _fuse_bn_stats else 0
    partials = _bn_partials(Cout, dev, rows)
    if rows > 0:
        conv_out = torch.empty((N, D, H, W, Cout), dtype=x0.dtype, device=dev)
        check(L.b200_conv3d_k3_bnstats(_dt(x0), impl, _ptr(x0), c0, _ptr(x1), c1, _ptr(wpack), _ptr(bias32), _ptr(conv_out), Cout,
                                       N, D, H, W, _ptr(partials), _stream()), "conv3d_k3_bnstats")
        check(L.b200_bn_finalize_ex(_ptr(partials), rows, _ptr(bias32), M, Cout, _ptr(g32), _ptr(b32), float(eps), float(momentum),
                                    _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                    _ptr(stats[3]), _stream()), "bn_finalize_ex")
    else:
        conv_out, _ = conv3d_k3_raw(x0, x1, wpack, bias32, Cout, 0, impl)
        check(L.b200_bn_stats(_dt(conv_out), _ptr(conv_out), M, Cout, _ptr(partials), _stream()), "bn_stats")
        check(L.b200_bn_finalize(_dt(conv_out), _ptr(conv_out), _ptr(partials), M, Cout, _ptr(g32), _ptr(b32), float(eps), float(momentum), 1,
                                 _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                 _ptr(stats[3]), _stream()), "bn_finalize")
    return conv_out, stats


class _ConvStats(torch.autograd.Function):
    """Conv3d(k3,p1)+bias and the batch statistics of the BatchNorm3d that follows (models/unet.py:15-16), WITHOUT applying it:
    the fused head (_FusedHead) normalises on the fly.  Returns (conv_out, stats); the gradient arrives w.r.t. conv_out."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, gamma, beta, running_mean, running_var, nbt, eps, momentum, impl, prev_conv_out=None, prev_stats=None):
        _require_cuda(x0, x1, weight)
        L = _lib.load()
        ctx.impl_req = impl
        x0 = x0.contiguous()
        x1 = None if x1 is None else x1.contiguous()
        impl = conv3d_select_impl(x0, x1, weight.shape[0], 0, impl)
        conv_out, stats = _conv_and_batch_stats(L, x0, x1, weight, bias, gamma.detach(), beta.detach(), running_mean, running_var, nbt, eps,
                                                momentum, impl)
        ctx.save_for_backward(x0, x1, weight, prev_conv_out, prev_stats)
        ctx.has_bias = bias is not None
        ctx.bias_ptr = bias.data_ptr() if bias is not None else 0
        ctx.mark_non_differentiable(stats)
        ctx.set_materialize_grads(False)
        return conv_out, stats

    @staticmethod
    def backward(ctx, dconv, _gstats):
        x0, x1, weight, prev_conv_out, prev_stats = ctx.saved_tensors
        if dconv is None:
            return (None,) * 14
        _run_bwd_callback(weight.data_ptr())
        dx0, dx1, dw, db = _conv_bwd_from_dconv(ctx, x0, x1, weight, dconv.contiguous(), zero_bias_grad=True, prev=(prev_conv_out, prev_stats))
        return (dx0, dx1, dw, db, None, None, None, None, None, None, None, None, None, None)


class _FusedHead(torch.autograd.Function):
    """Last BatchNorm3d + ReLU, final 1x1x1 conv, segmentation loss and confusion counts in one pass each way
    (models/unet.py:16-18, 62, 87; utils/metrics.py:14-40, 65-167) — csrc/head_fused.cu."""

    @staticmethod
    def forward(ctx, conv_out, stats, gamma, beta, fw, fb, target, mode, alpha, beta_t, want_conf, round_bf16):
        _require_cuda(conv_out, target)
        L = _lib.load()
        N, D, H, W, Cin = conv_out.shape
        C = fw.shape[0]
        S = D * H * W
        if target.dtype not in (torch.int64, torch.uint8):
            raise RuntimeError(f"expected int64 (or uint8) class-index target, got {target.dtype}")
        if target.numel() != N * S:
            raise ValueError(f"target shape {tuple(target.shape)} does not match the volume {(N, D, H, W)}")
        y = target.detach().contiguous()
        lb = 1 if y.dtype == torch.uint8 else 8
        dev = conv_out.device
        w32, b32 = _f32(fw).reshape(C, Cin), _f32(fb)
        logits = torch.empty((N, C, D, H, W), dtype=torch.float32, device=dev)
        sums = torch.empty(4 + 4 * C, dtype=torch.float64, device=dev)
        conf = torch.empty((C, C), dtype=torch.int64, device=dev) if want_conf else None
        check(L.b200_head_fwd(_ptr(conv_out), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(w32), _ptr(b32), int(round_bf16), _ptr(y), lb,
                              N, S, Cin, C, _ptr(logits), _ptr(sums), _ptr(conf), _stream()), "head_fwd")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty(2 + 2 * C, dtype=torch.float32, device=dev)
        check(L.b200_seg_loss_finalize(_ptr(sums), mode, float(alpha), float(beta_t), 1.0, 1.0, 0, N, C, S, _ptr(loss), _ptr(coef), _stream()),
              "seg_loss_finalize")
        ctx.save_for_backward(conv_out, stats, w32, logits, y, coef)
        ctx.fw_shape = tuple(fw.shape)
        ctx.has_fb = fb is not None
        ctx.ptrs = (fw.data_ptr(), fb.data_ptr() if fb is not None else 0, gamma.data_ptr(), beta.data_ptr())
        ctx.label_bytes = lb
        ctx.mark_non_differentiable(logits)
        if conf is not None:
            ctx.mark_non_differentiable(conf)
        ctx.set_materialize_grads(False)
        return loss, logits, conf

    @staticmethod
    def backward(ctx, gloss, _glogits, _gconf):
        L = _lib.load()
        conv_out, stats, w32, logits, y, coef = ctx.saved_tensors
        if gloss is None:
            return (None,) * 12
        N, D, H, W, Cin = conv_out.shape
        C = w32.shape[0]
        S = D * H * W
        dev = conv_out.device
        go = gloss.detach().float().contiguous().reshape(1)
        nb = L.b200_head_blocks(N, S)
        gy = torch.empty_like(conv_out)
        wpart = torch.empty(nb * 4 * (Cin + 1), dtype=torch.float32, device=dev)
        bnpart = torch.empty(nb * 2 * Cin, dtype=torch.float32, device=dev)
        fw_ptr, fb_ptr, gamma_ptr, beta_ptr = ctx.ptrs
        dw, w_sunk = _grad_dst(fw_ptr, (C, Cin), dev)
        db, fb_sunk = _grad_dst(fb_ptr, (C,), dev)
        check(L.b200_head_bwd(_ptr(logits), _ptr(y), ctx.label_bytes, _ptr(coef), _ptr(go), _ptr(conv_out), _ptr(stats[0]), _ptr(stats[1]),
                              _ptr(stats[2]), _ptr(stats[3]), _ptr(w32), N, S, Cin, C, _ptr(gy), _ptr(wpart), _ptr(bnpart), _ptr(dw), _ptr(db),
                              _stream()), "head_bwd")
        dgamma, g_sunk = _grad_dst(gamma_ptr, (Cin,), dev)
        dbeta, b_sunk = _grad_dst(beta_ptr, (Cin,), dev)
        sums = torch.empty(2 * Cin, dtype=torch.float32, device=dev)
        check(L.b200_bn_bwd_finalize_ex(_ptr(bnpart), nb, N * S, Cin, _ptr(dgamma), _ptr(dbeta), _ptr(sums), _stream()), "bn_bwd_finalize_ex")
        dconv = torch.empty_like(conv_out)
        check(L.b200_bn_act_bwd_apply(_dt(conv_out), _ptr(gy), _ptr(conv_out), _ptr(dconv), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]),
                                      _ptr(stats[3]), None, 1, _ptr(sums), 1, N, S, Cin, _stream()), "bn_act_bwd_apply")
        return (dconv, None, None if g_sunk else dgamma, None if b_sunk else dbeta, None if w_sunk else dw.reshape(ctx.fw_shape),
                (db if ctx.has_fb and not fb_sunk else None), None, None, None, None, None, None)


def conv_batch_stats(x0, x1, conv, bn, impl=0, prev=None):
    """(conv_out, stats) of conv -> BatchNorm3d in training mode, the normalisation itself left to fused_head()."""
    if bn.momentum is None:
        raise ValueError("BatchNorm3d(momentum=None) (cumulative moving average) is not supported by libb200unet")
    for name, buf, want in (("running_mean", bn.running_mean, torch.float32), ("running_var", bn.running_var, torch.float32),
                            ("num_batches_tracked", bn.num_batches_tracked, torch.int64)):
        if buf is not None and (buf.dtype != want or not buf.is_contiguous() or not buf.is_cuda):
            raise TypeError(f"BatchNorm3d.{name} must be a contiguous CUDA {want} tensor (got {buf.dtype} on {buf.device})")
    pco, pst = prev if prev is not None else (None, None)
    return _ConvStats.apply(x0, x1, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                            bn.eps, bn.momentum, impl, pco, pst)


def fused_head(conv_out, stats, bn, final_conv, target, mode, alpha=0.5, beta=0.5, want_confusion=False, round_bf16=True):
    """(loss, logits NCDHW fp32, confusion int64 [C, C] or None)"""
    return _FusedHead.apply(conv_out, stats, bn.weight, bn.bias, final_conv.weight, final_conv.bias, target, mode, alpha, beta,
                            want_confusion, round_bf16)


def conv_bn_act(x0, x1, conv, bn, dropmask, training, impl=0, prev=None, return_ctx=False):
    """One half of a DoubleConv.  prev = (conv_out, stats) returned (return_ctx=True) by the conv_bn_act whose output is x0."""
    if bn.momentum is None:
        # torch switches to a cumulative moving average (factor 1 / num_batches_tracked) here; the reference never does
        # (models/unet.py:12,16 use the default 0.1) and the fused finalize kernel takes a launch-constant factor
        raise ValueError("BatchNorm3d(momentum=None) (cumulative moving average) is not supported by libb200unet")
    pco, pst = prev if prev is not None else (None, None)
    y, conv_out, stats = _ConvBNAct.apply(x0, x1, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                          bn.num_batches_tracked, dropmask, training, bn.eps, bn.momentum, impl, pco, pst)
    return (y, (conv_out, stats)) if return_ctx else y


# --------------------------------------------------------------------------- MaxPool3d(2,2)
class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        L = _lib.load()
        x = x.contiguous()
        N, D, H, W, C = x.shape
        y = torch.empty((N, D // 2, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        check(L.b200_maxpool2_fwd(_dt(x), _ptr(x), _ptr(y), N, D, H, W, C, _stream()), "maxpool2_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.load()
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        N, D, H, W, C = x.shape
        gx = torch.empty_like(x)
        check(L.b200_maxpool2_bwd(_dt(x), _ptr(x), _ptr(gy), _ptr(gx), N, D, H, W, C, _stream()), "maxpool2_bwd")
        return gx


def maxpool2(x):
    return _MaxPool2.apply(x)


class _SkipPool(torch.autograd.Function):
    """Encoder hand-off (models/unet.py:69-71): returns (skip, MaxPool3d(2,2)(x)).  Owning both uses of x lets the backward
    pass add the skip connection's gradient inside the pool-backward kernel instead of in a separate accumulation pass."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        L = _lib.load()
        x = x.contiguous()
        N, D, H, W, C = x.shape
        y = torch.empty((N, D // 2, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        check(L.b200_maxpool2_fwd(_dt(x), _ptr(x), _ptr(y), N, D, H, W, C, _stream()), "maxpool2_fwd")
        ctx.save_for_backward(x)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, g_skip, g_pool):
        L = _lib.load()
        (x,) = ctx.saved_tensors
        if g_pool is None:
            return g_skip
        N, D, H, W, C = x.shape
        g_pool = g_pool.contiguous()
        gx = torch.empty_like(x)
        if g_skip is None:
            check(L.b200_maxpool2_bwd(_dt(x), _ptr(x), _ptr(g_pool), _ptr(gx), N, D, H, W, C, _stream()), "maxpool2_bwd")
        else:
            g_skip = g_skip.contiguous()
            check(L.b200_maxpool2_bwd_add(_dt(x), _ptr(x), _ptr(g_pool), _ptr(g_skip), _ptr(gx), N, D, H, W, C, _stream()), "maxpool2_bwd_add")
        return gx


def skip_and_pool(x):
    """(skip, pooled) = (x, maxpool2(x)) with a fused backward."""
    return _SkipPool.apply(x)


_fuse_pool_fwd = os.environ.get("B200_FUSE_POOL_FWD", "1") != "0"


def set_fuse_pool_fwd(on: bool) -> None:
    """Encoder hand-off forward: BatchNorm apply + MaxPool3d in one pass (default) or bn_act_fwd followed by maxpool2_fwd."""
    global _fuse_pool_fwd
    _fuse_pool_fwd = bool(on)


def bn_pool_fusable(x0, out_channels: int, dropmask) -> bool:
    """True if conv_bn_act_skip_pool serves a layer with input x0 [N, D, H, W, C] (even sizes, no Dropout3d mask)."""
    if not _fuse_pool_fwd or dropmask is not None or not x0.is_cuda or x0.dim() != 5:
        return False
    _, D, H, W, _ = x0.shape
    return bool(_lib.load().b200_bn_act_pool_fwd_supported(D, H, W, out_channels))


class _ConvBNActSkipPool(torch.autograd.Function):
    """_ConvBNAct followed by the encoder hand-off (skip, MaxPool3d(2,2)) — models/unet.py:15-18 feeding :69-71 — as ONE autograd
    node: the BatchNorm apply pass writes the pooled tensor as well (csrc/elementwise_kernels.cu, bn_act_pool_fwd_kernel), so the
    activation is not re-read by a separate pool pass.  Backward: maxpool2_bwd_add joins the two incoming gradients (as _SkipPool
    does), then _ConvBNAct's backward.  Same arithmetic as the two-node sequence, bit for bit."""

    @staticmethod
    def forward(ctx, *args):
        ctx._b200_want_pool = True
        y, conv_out, stats = _ConvBNAct.forward(ctx, *args)
        pooled = ctx._b200_pooled
        ctx._b200_pooled = None
        # the pool backward routes by the arg-max of y: saved as a ninth tensor (an output, alive as the skip connection anyway).
        # ctx.to_save is what _ConvBNAct.forward handed to save_for_backward a moment ago (torch.autograd.function.FunctionCtx)
        saved = getattr(ctx, "to_save", None)
        if saved is None or len(saved) != 8:
            raise RuntimeError("_ConvBNActSkipPool: expected the eight tensors _ConvBNAct.forward saves for backward")
        ctx.save_for_backward(*saved, y)
        return y, pooled, conv_out, stats

    @staticmethod
    def backward(ctx, g_skip, g_pool, _gconv=None, _gstats=None):
        if g_pool is None:
            return _ConvBNAct.backward(ctx, g_skip)
        L = _lib.load()
        y = ctx.saved_tensors[8]
        N, D, H, W, C = y.shape
        g_pool = g_pool.contiguous()
        gy = torch.empty_like(y)
        if g_skip is None:
            check(L.b200_maxpool2_bwd(_dt(y), _ptr(y), _ptr(g_pool), _ptr(gy), N, D, H, W, C, _stream()), "maxpool2_bwd")
        else:
            g_skip = g_skip.contiguous()
            check(L.b200_maxpool2_bwd_add(_dt(y), _ptr(y), _ptr(g_pool), _ptr(g_skip), _ptr(gy), N, D, H, W, C, _stream()), "maxpool2_bwd_add")
        return _ConvBNAct.backward(ctx, gy)


def conv_bn_act_skip_pool(x0, x1, conv, bn, training, impl=0, prev=None):
    """(skip, pooled) = skip_and_pool(conv_bn_act(...)) with the pooled tensor written by the BatchNorm apply pass; call only when
    bn_pool_fusable() holds."""
    if bn.momentum is None:
        raise ValueError("BatchNorm3d(momentum=None) (cumulative moving average) is not supported by libb200unet")
    pco, pst = prev if prev is not None else (None, None)
    y, pooled, _, _ = _ConvBNActSkipPool.apply(x0, x1, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                               bn.num_batches_tracked, None, training, bn.eps, bn.momentum, impl, pco, pst)
    return y, pooled


# --------------------------------------------------------------------------- ConvTranspose3d(k2,s2)
class _ConvT2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda(x, weight)
        L = _lib.load()
        x = x.contiguous()
        N, D, H, W, Cin = x.shape
        Cout = weight.shape[1]
        w32, b32 = _f32(weight), _f32(bias)
        y = torch.empty((N, 2 * D, 2 * H, 2 * W, Cout), dtype=x.dtype, device=x.device)
        check(L.b200_convt2_fwd(_dt(x), _ptr(x), _ptr(w32), _ptr(b32), _ptr(y), N, D, H, W, Cin, Cout, _stream()), "convt2_fwd")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias_ptr = bias.data_ptr() if bias is not None else 0
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.load()
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        N, D, H, W, Cin = x.shape
        Cout = weight.shape[1]
        w32 = _f32(weight)
        gx = dw = db = ws = side = None
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if need_w:  # on a side stream next to the data gradient (see fork_side)
            side = fork_side(x.device) if ctx.needs_input_grad[0] else None
            ws_bytes = L.b200_convt2_wgrad_workspace(Cin, Cout, N, D, H, W)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            dw, w_sunk = _grad_dst(weight.data_ptr(), tuple(w32.shape), x.device)
            db, b_sunk = _grad_dst(ctx.bias_ptr, (Cout,), x.device)
            check(
                L.b200_convt2_bwd_weight(_dt(x), _ptr(x), _ptr(gy), _ptr(dw), _ptr(db), _ptr(ws), ws_bytes, N, D, H, W, Cin, Cout,
                                         _stream() if side is None else side.cuda_stream),
                "convt2_bwd_weight",
            )
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            check(L.b200_convt2_bwd_data(_dt(x), _ptr(gy), _ptr(w32), _ptr(gx), N, D, H, W, Cin, Cout, _stream()), "convt2_bwd_data")
        join_side(side, x.device, x, gy, ws)
        del ws
        if need_w and w_sunk:
            dw = None
        if need_w and b_sunk:
            db = None
        return gx, dw, (db if ctx.has_bias else None)


def conv_transpose2(x, weight, bias):
    return _ConvT2.apply(x, weight, bias)


class _NearestResize(torch.autograd.Function):
    """F.interpolate(x, size=...) with the default mode='nearest' (models/unet.py:81-83)."""

    @staticmethod
    def forward(ctx, x, size):
        _require_cuda(x)
        L = _lib.load()
        x = x.contiguous()
        N, D, H, W, C = x.shape
        OD, OH, OW = size
        y = torch.empty((N, OD, OH, OW, C), dtype=x.dtype, device=x.device)
        check(L.b200_nearest_resize_fwd(_dt(x), _ptr(x), _ptr(y), N, D, H, W, OD, OH, OW, C, _stream()), "nearest_resize_fwd")
        ctx.in_shape = (N, D, H, W, C)
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.load()
        gy = gy.contiguous()
        N, D, H, W, C = ctx.in_shape
        OD, OH, OW = gy.shape[1:4]
        gx = torch.empty(ctx.in_shape, dtype=gy.dtype, device=gy.device)
        check(L.b200_nearest_resize_bwd(_dt(gy), _ptr(gy), _ptr(gx), N, D, H, W, OD, OH, OW, C, _stream()), "nearest_resize_bwd")
        return gx, None


def nearest_resize(x, size):
    return _NearestResize.apply(x, tuple(int(s) for s in size))


# --------------------------------------------------------------------------- final 1x1x1 conv
class _FinalConv1x1(torch.autograd.Function):
    """nn.Conv3d(features[0], out_channels, 1) (models/unet.py:62,87): NDHWC -> NCDHW fp32 logits."""

    @staticmethod
    def forward(ctx, x, weight, bias, round_bf16):
        _require_cuda(x, weight)
        L = _lib.load()
        x = x.contiguous()
        N, D, H, W, Cin = x.shape
        Cout = weight.shape[0]
        S = D * H * W
        w32 = _f32(weight).reshape(Cout, Cin)
        b32 = _f32(bias)
        y = torch.empty((N, Cout, D, H, W), dtype=torch.float32, device=x.device)
        check(L.b200_conv1x1_fwd(_dt(x), _ptr(x), _ptr(w32), _ptr(b32), _ptr(y), N, S, Cin, Cout, int(round_bf16), _stream()), "conv1x1_fwd")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias_ptr = bias.data_ptr() if bias is not None else 0
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.load()
        x, weight = ctx.saved_tensors
        gy = gy.float().contiguous()
        N, D, H, W, Cin = x.shape
        Cout = weight.shape[0]
        S = D * H * W
        w32 = _f32(weight).reshape(Cout, Cin)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, w_sunk = _grad_dst(weight.data_ptr(), (Cout, Cin), x.device)
        db, b_sunk = _grad_dst(ctx.bias_ptr, (Cout,), x.device)
        partials = torch.empty(L.b200_conv1x1_partials_bytes(Cin, Cout) // 4, dtype=torch.float32, device=x.device)
        check(
            L.b200_conv1x1_bwd(_dt(x), _ptr(x), _ptr(w32), _ptr(gy), _ptr(gx), _ptr(dw), _ptr(db), _ptr(partials), N, S, Cin, Cout,
                               _stream()),
            "conv1x1_bwd",
        )
        return gx, (None if w_sunk else dw.reshape(weight.shape)), (db if ctx.has_bias and not b_sunk else None), None


def final_conv1x1(x, weight, bias, round_bf16=False):
    return _FinalConv1x1.apply(x, weight, bias, round_bf16)


# --------------------------------------------------------------------------- global average pool
class _GlobalAvgPool(torch.autograd.Function):
    """torch.mean(bottleneck, dim=[2,3,4]) (models/unet_dann.py:79) on an NDHWC tensor -> [N, C] fp32."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        L = _lib.load()
        x = x.contiguous()
        N, C = x.shape[0], x.shape[-1]
        S = x.numel() // (N * C)
        out = torch.empty((N, C), dtype=torch.float32, device=x.device)
        check(L.b200_gap_fwd(_dt(x), _ptr(x), _ptr(out), N, S, C, _stream()), "gap_fwd")
        ctx.shape = tuple(x.shape)
        ctx.dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.load()
        g = g.float().contiguous()
        N, C = ctx.shape[0], ctx.shape[-1]
        gx = torch.empty(ctx.shape, dtype=ctx.dtype, device=g.device)
        S = gx.numel() // (N * C)
        check(L.b200_gap_bwd(_dt(gx), _ptr(g), _ptr(gx), 0, N, S, C, _stream()), "gap_bwd")
        return gx


def global_avg_pool(x):
    return _GlobalAvgPool.apply(x)


# --------------------------------------------------------------------------- losses
class _SegLoss(torch.autograd.Function):
    """Fused softmax + CE + Dice/Tversky (+ KD-KL) — utils/metrics.py:14-40, 137-190."""

    @staticmethod
    def forward(ctx, logits, target, teacher, mode, alpha, beta, kd_alpha, temperature):
        _require_cuda(logits, target, teacher)
        L = _lib.load()
        if logits.dim() < 3:
            raise ValueError(f"expected logits [B, C, ...], got {tuple(logits.shape)}")
        z = logits.detach().float().contiguous()
        N, C = z.shape[0], z.shape[1]
        S = z.numel() // (N * C)
        if target.dtype != torch.int64:
            raise RuntimeError(f"expected int64 class-index target (the reference's CrossEntropyLoss contract), got {target.dtype}")
        if target.numel() != N * S:
            raise ValueError(f"target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
        y = target.detach().contiguous()
        t = None
        if teacher is not None:
            if teacher.shape != logits.shape:
                raise ValueError("teacher and student logits must have the same shape")
            t = teacher.detach().float().contiguous()
        dev = z.device
        sums = torch.empty(4 + 4 * C, dtype=torch.float64, device=dev)
        check(L.b200_kd_loss_fwd(_ptr(z), _ptr(t), _ptr(y), float(temperature), N, C, S, _ptr(sums), _stream()), "seg_loss_fwd")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty(2 + 2 * C, dtype=torch.float32, device=dev)
        check(
            L.b200_seg_loss_finalize(_ptr(sums), mode, float(alpha), float(beta), float(kd_alpha), float(temperature),
                                     int(t is not None), N, C, S, _ptr(loss), _ptr(coef), _stream()),
            "seg_loss_finalize",
        )
        ctx.save_for_backward(z, y, t, coef)
        ctx.temperature = float(temperature)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(sums)
        return loss, sums

    @staticmethod
    def backward(ctx, gout, _gsums):
        L = _lib.load()
        z, y, t, coef = ctx.saved_tensors
        N, C = z.shape[0], z.shape[1]
        S = z.numel() // (N * C)
        go = gout.detach().float().contiguous().reshape(1)
        dz = torch.empty_like(z)
        check(
            L.b200_seg_loss_bwd(_ptr(z), _ptr(t), _ptr(y), _ptr(coef), _ptr(go), ctx.temperature, N, C, S, _ptr(dz), _stream()),
            "seg_loss_bwd",
        )
        if ctx.in_dtype != torch.float32:
            dz = dz.to(ctx.in_dtype)
        return dz, None, None, None, None, None, None, None


def seg_loss(logits, target, mode, alpha=0.5, beta=0.5, teacher=None, kd_alpha=1.0, temperature=1.0):
    loss, _ = _SegLoss.apply(logits, target, teacher, mode, alpha, beta, kd_alpha, temperature)
    return loss


def seg_loss_backward_raw(z, target, coef, gout, teacher=None, temperature=1.0):
    """dlogits of the fused loss from explicit gradient coefficients (csrc/loss_kernels.cu coef layout); z fp32 contiguous."""
    L = _lib.load()
    N, C = z.shape[0], z.shape[1]
    S = z.numel() // (N * C)
    go = gout.detach().float().contiguous().reshape(1)
    dz = torch.empty_like(z)
    check(L.b200_seg_loss_bwd(_ptr(z), _ptr(teacher), _ptr(target.contiguous()), _ptr(coef), _ptr(go), float(temperature), N, C, S, _ptr(dz),
                              _stream()), "seg_loss_bwd")
    return dz


def seg_loss_sums(logits, target):
    """(CE_sum, per-class I, P, T) as float64 — exposed for tests."""
    _, sums = _SegLoss.apply(logits, target, None, _lib.LOSS_DICE_CE, 0.5, 0.5, 1.0, 1.0)
    return sums


# --------------------------------------------------------------------------- metrics
def confusion_counts(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """int64 [C, C] confusion matrix conf[target, argmax] in one pass (device tensor, no sync)."""
    _require_cuda(pred, target)
    L = _lib.load()
    z = pred.detach().float().contiguous()
    N, C = z.shape[0], z.shape[1]
    S = z.numel() // max(N * C, 1)
    if target.dtype != torch.int64:
        target = target.long()
    y = target.detach().contiguous()
    if y.numel() != N * S:
        raise ValueError(f"target shape {tuple(target.shape)} does not match pred {tuple(pred.shape)}")
    conf = torch.empty((C, C), dtype=torch.int64, device=z.device)
    check(L.b200_confusion(_ptr(z), _ptr(y), N, C, S, _ptr(conf), _stream()), "confusion")
    return conf


def argmax_mask(pred: torch.Tensor) -> torch.Tensor:
    """torch.argmax(pred, dim=1) as uint8 [N, ...]."""
    _require_cuda(pred)
    L = _lib.load()
    z = pred.detach().float().contiguous()
    N, C = z.shape[0], z.shape[1]
    S = z.numel() // max(N * C, 1)
    out = torch.empty((N, *z.shape[2:]), dtype=torch.uint8, device=z.device)
    check(L.b200_argmax(_ptr(z), N, C, S, _ptr(out), _stream()), "argmax")
    return out


# --------------------------------------------------------------------------- DANN pieces
def scale(x: torch.Tensor, alpha: float) -> torch.Tensor:
    """alpha * x through the library (used by the gradient-reversal backward)."""
    _require_cuda(x)
    L = _lib.load()
    x32 = x.detach().float().contiguous()
    out = torch.empty_like(x32)
    check(L.b200_scale_f32(_ptr(x32), _ptr(out), float(alpha), x32.numel(), _stream()), "scale_f32")
    return out.to(x.dtype)


class _LinearAct(torch.autograd.Function):
    """nn.Linear (+ReLU) (+Dropout mask) — train_dann.py:37-46."""

    @staticmethod
    def forward(ctx, x, weight, bias, dropmask, relu):
        _require_cuda(x, weight)
        L = _lib.load()
        x32 = x.detach().float().contiguous()
        w32, b32 = _f32(weight), _f32(bias)
        B, I = x32.shape
        O = w32.shape[0]
        y = torch.empty((B, O), dtype=torch.float32, device=x.device)
        check(L.b200_linear_fwd(_ptr(x32), _ptr(w32), _ptr(b32), _ptr(dropmask), int(relu), _ptr(y), B, I, O, _stream()), "linear_fwd")
        ctx.save_for_backward(x32, w32, y, dropmask)
        ctx.relu = int(relu)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.load()
        x32, w32, y, dropmask = ctx.saved_tensors
        gy = gy.float().contiguous()
        B, I = x32.shape
        O = w32.shape[0]
        gx = torch.empty_like(x32) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w32)
        db = torch.empty(O, dtype=torch.float32, device=gy.device)
        check(
            L.b200_linear_bwd(_ptr(x32), _ptr(w32), _ptr(y), _ptr(gy), _ptr(dropmask), ctx.relu, _ptr(gx), _ptr(dw), _ptr(db), B, I, O,
                              _stream()),
            "linear_bwd",
        )
        return gx, dw, (db if ctx.has_bias else None), None, None


def linear_act(x, weight, bias, dropmask=None, relu=False):
    return _LinearAct.apply(x, weight, bias, dropmask, relu)


class _CERows(torch.autograd.Function):
    """nn.CrossEntropyLoss() on [B, C] logits (train_dann.py:258)."""

    @staticmethod
    def forward(ctx, logits, labels):
        _require_cuda(logits, labels)
        L = _lib.load()
        z = logits.detach().float().contiguous()
        y = labels.detach().long().contiguous()
        B, C = z.shape
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        check(L.b200_ce_rows(_ptr(z), _ptr(y), B, C, _ptr(loss), _ptr(dz), _stream()), "ce_rows")
        ctx.save_for_backward(dz)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


def cross_entropy_rows(logits, labels):
    return _CERows.apply(logits, labels)


# --------------------------------------------------------------------------- fused AdamW on flat buffers
class FlatAdamW(torch.optim.Optimizer):
    """AdamW (train_unet.py:378) over one flat fp32 parameter/gradient buffer; graph-capturable.

    A ``torch.optim.Optimizer`` so that the reference's ``ReduceLROnPlateau(optimizer, mode='max', ...)``
    (train_unet.py:381, 442) drives it unchanged, and with the wire format of ``torch.optim.AdamW.state_dict()`` so that
    the reference's checkpoints (train_unet.py:477-486) load into it and vice versa: pass ``named_params`` (the model's
    trainable parameters in ``model.parameters()`` order) and ``offsets`` {name: (offset, numel)} into the flat buffer.
    The learning rate lives in device memory (``hyper[0]``): changing ``param_groups[0]['lr']`` takes effect at the next
    ``step()`` / ``sync_hyper()`` even inside a captured graph; betas / eps / weight_decay are launch constants."""

    def __init__(self, flat_param, flat_grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, named_params=None, offsets=None):
        _require_cuda(flat_param, flat_grad)
        self.named = list(named_params) if named_params is not None else None
        self.offsets = offsets
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__([q for _, q in self.named] if self.named else [flat_param], defaults)
        self.p, self.g = flat_param, flat_grad
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.step_count = torch.zeros((), dtype=torch.int64, device=flat_param.device)
        self.hyper = torch.tensor([lr, 1.0, 1.0, 0.0], dtype=torch.float32, device=flat_param.device)
        self._lr_on_device = float(lr)

    # kept for callers that predate param_groups
    @property
    def betas(self):
        return self.param_groups[0]["betas"]

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    @property
    def weight_decay(self):
        return self.param_groups[0]["weight_decay"]

    def set_lr(self, lr: float):
        self.param_groups[0]["lr"] = float(lr)
        self.sync_hyper()

    def sync_hyper(self):
        """Pushes param_groups[0]['lr'] to the device word the (possibly graph-captured) kernels read."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.hyper[0:1].fill_(lr)
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        self.begin_step()
        self.apply_range(0, self.p.numel(), grad_scale)

    @torch.no_grad()
    def begin_step(self):
        """Advances the step counter and the bias corrections once per optimiser step; follow with apply_range() calls that
        together cover the flat buffer (a trainer steps the buckets whose all-reduce has finished while the last is in flight)."""
        L = _lib.load()
        self.sync_hyper()
        g = self.param_groups[0]
        check(L.b200_adamw_prepare(_ptr(self.step_count), g["betas"][0], g["betas"][1], _ptr(self.hyper), _stream()), "adamw_prepare")
        weights_changed()       # parameters are about to be written through raw pointers: cached packed layouts go stale ONCE per step

    @torch.no_grad()
    def apply_range(self, lo: int, hi: int, grad_scale: float = 1.0):
        """AdamW update of flat elements [lo, hi); lo must be a multiple of 4 (16-byte aligned slices)."""
        if hi <= lo:
            return
        if lo % 4:
            raise ValueError(f"FlatAdamW.apply_range: lo={lo} must be a multiple of 4")
        L = _lib.load()
        g = self.param_groups[0]
        check(
            L.b200_adamw_flat(_ptr(self.p[lo:hi]), _ptr(self.g[lo:hi]), _ptr(self.m[lo:hi]), _ptr(self.v[lo:hi]), hi - lo, _ptr(self.hyper),
                              g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], float(grad_scale), _stream()),
            "adamw_flat",
        )

    # ---- torch.optim.AdamW wire format ------------------------------------------------------------------------------
    def _slices(self):
        if self.named is None or self.offsets is None:
            raise RuntimeError("FlatAdamW: state_dict()/load_state_dict() need named_params and offsets")
        return [(q, *self.offsets[n]) for n, q in self.named]

    def state_dict(self):
        step = float(self.step_count.item())
        state = {}
        if step > 0:
            for i, (q, off, k) in enumerate(self._slices()):
                state[i] = {"step": torch.tensor(step), "exp_avg": self.m[off:off + k].view_as(q).clone(),
                            "exp_avg_sq": self.v[off:off + k].view_as(q).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.named)))
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        slices = self._slices()
        groups = state_dict["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(slices):
            raise ValueError(f"loaded state dict has {len(ids)} parameters, this optimizer has {len(slices)}")
        self.m.zero_()
        self.v.zero_()
        step = 0.0
        for (q, off, k), i in zip(slices, ids):
            st = state_dict["state"].get(i)
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(q.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} does not match parameter {tuple(q.shape)}")
            self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, float(st["step"]))
        self.step_count.fill_(int(step))
        g0 = groups[0]
        for key in ("lr", "betas", "eps", "weight_decay"):
            if key in g0:
                self.param_groups[0][key] = tuple(g0[key]) if key == "betas" else g0[key]
        self._lr_on_device = None
        self.sync_hyper()
