"""Batch-sharded data-parallel training step (one process per GPU), the B200-native counterpart of
what HF Accelerate/DDP does for the reference (train_unet.py:309-312, 384-386, 221-226):

* all parameters live in ONE flat fp32 buffer and all gradients in another (``p.data`` / ``p.grad``
  are views), so the gradient exchange is a single in-place NCCL all-reduce per bucket instead of
  DDP's per-bucket copies, and AdamW is one fused kernel over the flat buffer;
* BatchNorm statistics stay per replica (the reference does not use SyncBN), so forward/backward
  need no communication;
* the bucket holding the parameter-heavy, cheap layers (decoder.*, upconvs, bottleneck = 83 % of the
  parameters) is reduced on a side stream as soon as its last gradient is produced, overlapping
  the remaining full-resolution encoder backward;
* the whole step (zero-grad, forward, loss, backward, all-reduce, AdamW, confusion counts) can be
  captured in one CUDA graph: every per-step scalar lives on the device.

The flat-buffer / bucket logic is device agnostic and is exercised on CPU with the gloo backend in
tests/test_dp_cpu.py; the fused optimiser and the graph capture need CUDA.
"""
from __future__ import annotations

from typing import Callable, Optional

import os

import torch
import torch.distributed as dist


class FlatParams:
    """Re-homes a module's parameters and gradients into two flat fp32 buffers.

    Parameters are laid out bucket by bucket in the order their gradients COMPLETE during backward: bucket 0 = the parameter-heavy
    tail of the forward pass (final conv, decoder, up-convs, bottleneck: 83 % of the U-Net's parameters), then groups of encoder
    levels from the deepest up.  Bucket i is final when backward reaches the first layer of the next group: that layer's
    second-conv weight is the bucket's SENTINEL (a post-accumulate-grad hook on it starts the bucket's all-reduce on a side
    stream).  Only the last, small bucket (the top encoder level: 30 KB for the default U-Net) is reduced after backward, and
    the trainer hides even that under the optimiser step of the other buckets.  Buckets start on 16-byte boundaries."""

    def __init__(self, module: torch.nn.Module, early_prefixes=("final_conv", "decoder", "upconvs", "bottleneck"), late_levels_per_bucket=3,
                 tail_levels=1):
        params = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        early = [(n, p) for n, p in params if n.startswith(tuple(early_prefixes))]
        rest = [(n, p) for n, p in params if not n.startswith(tuple(early_prefixes))]
        enc_ids = sorted({int(n.split(".")[1]) for n, _ in rest if n.startswith("encoder.") and n.split(".")[1].isdigit()}, reverse=True)
        groups = [early] if early else []
        # encoder levels, deepest first, `late_levels_per_bucket` levels per bucket; anything else trainable goes last
        covered = set()
        # the LAST bucket is the only one whose all-reduce cannot hide under backward: it holds just the top `tail_levels` encoder
        # level(s) (30 KB for the default U-Net), and the trainer hides it under the optimiser step of the other buckets
        tail = max(0, min(int(tail_levels), len(enc_ids) - 1)) if len(enc_ids) > 1 else 0
        body_ids = enc_ids[:len(enc_ids) - tail] if tail else enc_ids
        chunks = [body_ids[i:i + max(1, late_levels_per_bucket)] for i in range(0, len(body_ids), max(1, late_levels_per_bucket))]
        if tail:
            chunks.append(enc_ids[len(enc_ids) - tail:])
        for ids in chunks:
            grp = [(n, p) for n, p in rest if any(n.startswith(f"encoder.{j}.") for j in ids)]
            covered |= {n for n, _ in grp}
            if grp:
                groups.append(grp)
        other = [(n, p) for n, p in rest if n not in covered]
        if other:
            if len(groups) > (1 if early else 0):
                groups[-1] = groups[-1] + other
            else:
                groups.append(other)
        if not groups:
            raise ValueError("FlatParams: the module has no trainable parameters")
        self.groups = groups
        self.order = [np_ for g in groups for np_ in g]
        # every bucket starts on a 16-byte boundary (the fused optimiser steps bucket ranges with 128-bit accesses); the padding
        # elements are zeros with zero gradients, which AdamW leaves at zero
        starts, off = [], 0
        for g in groups:
            off = (off + 3) // 4 * 4
            starts.append(off)
            off += sum(p.numel() for _, p in g)
        self.total = (off + 3) // 4 * 4
        dev = self.order[0][1].device
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.offsets = {}
        for g, off in zip(groups, starts):
            for n, p in g:
                k = p.numel()
                self.flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                self.offsets[n] = (off, k)
                off += k
        # element / parameter-count boundaries of the buckets (a bucket's trailing padding rides with it)
        self.bucket_elems, self.bucket_params = starts + [self.total], [0]
        for g in groups:
            self.bucket_params.append(self.bucket_params[-1] + len(g))
        # Sentinel of bucket i (i < last) = a parameter of the autograd node that runs right AFTER the bucket is complete: the second
        # conv of the deepest encoder level of the NEXT group.  AccumulateGrad nodes run with top priority, so when it fires every
        # gradient of the earlier buckets is final.
        self.sentinels = []
        for i in range(len(groups) - 1):
            nxt = groups[i + 1]
            ids = sorted({int(n.split(".")[1]) for n, _ in nxt if n.startswith("encoder.") and n.split(".")[1].isdigit()}, reverse=True)
            cand = [p for n, p in nxt if ids and n == f"encoder.{ids[0]}.double_conv.4.weight"]
            self.sentinels.append(cand[0] if cand else None)
        if any(s_ is None for s_ in self.sentinels):      # unknown layout: fall back to a single bucket reduced after backward
            self.sentinels = []
            self.bucket_elems, self.bucket_params = [0, self.total], [0, len(self.order)]
        self.n_early = self.bucket_elems[1] if len(self.bucket_elems) > 2 else (self.bucket_elems[1] if early and rest else (self.total if early else 0))
        self.early_sentinel = self.sentinels[0] if self.sentinels else None
        self._views = [self.grad[o:o + k].view_as(p) for (n, p), (o, k) in zip(self.order, (self.offsets[n] for n, _ in self.order))]
        # gradient sinks (functional.begin_grad_sinks): parameter data pointer -> its slice of the flat gradient buffer
        self.sinks = {p.data_ptr(): v for (_, p), v in zip(self.order, self._views)}
        self._known_zero = set()     # indices of slices that hold zeros and that nobody writes (parameters without a gradient)
        self.n_early_params = self.bucket_params[1] if len(self.bucket_params) > 2 else len(self.order)

    @property
    def n_buckets(self):
        return len(self.bucket_elems) - 1

    def bucket(self, i):
        return self.grad[self.bucket_elems[i]:self.bucket_elems[i + 1]]

    def buckets(self):
        return [self.bucket(i) for i in range(self.n_buckets)]

    def zero_grad(self):
        self.grad.zero_()
        self._known_zero = set(range(len(self.order)))

    # ---- "detached gradient" mode: autograd stores every parameter gradient as its own tensor (p.grad starts as
    # None, so AccumulateGrad keeps the tensor the kernels produced instead of launching one add per parameter), and the
    # gradients are gathered into the flat buffer by one multi-tensor copy per bucket.
    def detach_grads(self):
        for _, p in self.order:
            p.grad = None

    def attach_grads(self):
        """after a step: parameters whose gradient lives only in the flat buffer show it as .grad again"""
        for (_, p), v in zip(self.order, self._views):
            if p.grad is None:
                p.grad = v

    def gather_bucket(self, i, accumulate=False):
        """copies the detached gradients of bucket i into the flat buffer"""
        self._gather_range(self.bucket_params[i], self.bucket_params[i + 1], accumulate)

    def gather_grads(self, which="all", accumulate=False):
        # A parameter that received no gradient this step gets a ZERO slice: FlatAdamW then still applies weight decay and
        # decays its moments, whereas torch.optim.AdamW (the reference's optimiser) skips `grad is None` parameters.  Every
        # parameter of UNet3D / the DANN variant receives a gradient on every step, so the two agree on the path this
        # library covers; modules with conditionally unused parameters must not rely on the skip.
        lo, hi = {"all": (0, len(self.order)), "early": (0, self.n_early_params), "late": (self.n_early_params, len(self.order))}[which]
        self._gather_range(lo, hi, accumulate)

    def _gather_range(self, lo, hi, accumulate=False):
        from .functional import grad_was_sunk
        dst, src = [], []
        for i, ((n, p), v) in enumerate(zip(self.order[lo:hi], self._views[lo:hi]), start=lo):
            if p.grad is None:
                if grad_was_sunk(p.data_ptr()):
                    self._known_zero.discard(i)       # the backward kernels wrote this slice in place
                elif not accumulate and i not in self._known_zero:
                    # no gradient this step (e.g. a conv bias in front of a batch-statistics BatchNorm): the slice is zeroed once
                    # and stays zero — all-reduce sums zeros, the optimiser only reads
                    v.zero_()
                    self._known_zero.add(i)
            else:
                self._known_zero.discard(i)
                dst.append(v)
                src.append(p.grad)
        if dst:
            if accumulate:
                torch._foreach_add_(dst, src)
            else:
                torch._foreach_copy_(dst, src)


class _LossHandle:
    """Loss of one step on its way to the host (pinned slot + completion event)."""

    def __init__(self, host, event):
        self._host, self._event = host, event

    def value(self) -> float:
        self._event.synchronize()
        return float(self._host)


class DataParallelTrainer:
    """step(x, y) = zero_grad -> forward -> loss -> backward -> all-reduce(sum) -> AdamW(grad/world)."""

    def __init__(self, model, loss_fn: Callable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None, overlap=True,
                 metrics_fn: Optional[Callable] = None, optimizer_factory=None, accumulation_steps: int = 1, fused_head: bool = True):
        # metrics_fn: a callable (logits, target) -> anything, or the string "confusion" = int64 [C, C] counts conf[target, argmax]
        # (computed inside the fused head when the model offers it: UNet3D.forward_with_loss)
        self.model, self.loss_fn, self.metrics_fn = model, loss_fn, metrics_fn
        self.fused_head = fused_head and os.environ.get("B200_FUSED_HEAD", "1") != "0"
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.autocast_dtype = autocast_dtype
        self.fp = FlatParams(model)
        dev = self.fp.flat.device
        if self.world > 1:
            self.broadcast_state(src=0)
        if optimizer_factory is not None:
            self.opt = optimizer_factory(self.fp)
        else:
            from .functional import FlatAdamW
            named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]  # model.parameters() order = torch.optim's
            self.opt = FlatAdamW(self.fp.flat, self.fp.grad, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                 named_params=named, offsets=self.fp.offsets)
        self.overlap = overlap and self.world > 1 and dev.type == "cuda" and len(self.fp.sentinels) > 0
        # the communication stream outranks the main chain (priority -1, see capture()): an all-reduce kernel needs a handful of CTAs
        # and sits on the step's critical path once backward is over — at default priority the tail all-reduce waited for every
        # block of the optimiser kernel to be dispatched first (timeline: profiles/r02_timeline_2gpu_graph_replay.txt)
        self._side = torch.cuda.Stream(device=dev, priority=int(os.environ.get("B200_COMM_STREAM_PRIORITY", "-2"))) if self.overlap else None
        # weight gradients run on a side stream and are joined at the end of backward / before a bucket is gathered:
        # safe here because step() detaches the gradients first (autograd stores, never accumulates)
        self._defer_wgrad = dev.type == "cuda" and os.environ.get("B200_DEFER_WGRAD_JOIN", "1") != "0"
        self._grad_sinks_on = os.environ.get("B200_GRAD_SINKS", "1") != "0"
        self._one = None
        self._buckets_done = 0      # buckets whose all-reduce has been launched during the running backward pass
        if self.overlap:
            for i, sentinel in enumerate(self.fp.sentinels):
                sentinel.register_post_accumulate_grad_hook(lambda _p, i=i: self._bucket_hook(i))
        # gradient accumulation (accelerator.accumulate, train_unet.py:221): micro-steps 1..k-1 only add their gradients
        # (of loss / k) into the flat buffer; the k-th also all-reduces and steps the optimiser
        self.accum = max(1, int(accumulation_steps))
        self._micro = 0
        self.graph = None
        self.static_x = self.static_y = None
        self._copy_stream = None
        self._stage = None
        self.loss = None
        self.metrics = None

    # -- communication ---------------------------------------------------------------------------
    def broadcast_state(self, src: int = 0):
        """Every replica starts from rank `src`'s parameters and buffers — what DDP / Accelerate's prepare() does at
        construction (train_unet.py:384-386); the reference seeds only when --seed is given, so ranks may have built
        different initial weights.  One broadcast for the flat parameter buffer, one per dtype for the module buffers
        (BatchNorm running_mean / running_var / num_batches_tracked)."""
        dist.broadcast(self.fp.flat, src=src, group=self.pg)
        by_dtype = {}
        for b in self.model.buffers():
            by_dtype.setdefault(b.dtype, []).append(b)
        for dtype, bufs in by_dtype.items():
            flat = torch.cat([b.detach().reshape(-1) for b in bufs])
            dist.broadcast(flat, src=src, group=self.pg)
            off = 0
            with torch.no_grad():
                for b in bufs:
                    b.copy_(flat[off:off + b.numel()].view_as(b))
                    off += b.numel()

    def _bucket_hook(self, i):
        # called by autograd right after sentinel i's gradient was accumulated: every gradient of bucket i is final -> reduce it on
        # the side stream while the rest of the backward pass continues
        if self.accum > 1 or i != self._buckets_done:
            return  # accumulating: buckets are reduced once, after the last micro-step
        lo, hi = self.fp.bucket_params[i], self.fp.bucket_params[i + 1]
        if self._defer_wgrad and all(p.grad is None for _, p in self.fp.order[lo:hi]):
            # every gradient of the bucket was written into the flat buffer by its kernel (gradient sinks): nothing to copy, so only
            # the COMMUNICATION stream has to wait for the side-stream weight gradients — the main chain keeps running under them
            from . import functional as F
            F.order_after_pending(self._side)
        else:
            self._join_wgrads()
        self.fp.gather_bucket(i)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.fp.bucket(i), op=dist.ReduceOp.SUM, group=self.pg)
        self._buckets_done = i + 1

    def _join_wgrads(self):
        if self._defer_wgrad:
            from . import functional as F
            F.join_pending()

    def allreduce_grads(self, defer_last=False):
        """Reduces the buckets that backward has not launched yet.  defer_last: when only the LAST bucket is left, start its
        all-reduce on the side stream and return True WITHOUT waiting for it — the current stream has then waited for the earlier
        buckets only, and the caller must wait for the side stream before it touches the last bucket (see _step_body)."""
        self._join_wgrads()
        n = self.fp.n_buckets
        done = self._buckets_done if self.overlap else 0
        for i in range(done, n):
            self.fp.gather_bucket(i)
        deferred = False
        if self.world > 1:
            if defer_last and self.overlap and n >= 2 and done == n - 1:
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(self._side)                 # the earlier buckets' all-reduces
                self._side.wait_stream(cur)           # the last weight gradients
                with torch.cuda.stream(self._side):
                    dist.all_reduce(self.fp.bucket(n - 1), op=dist.ReduceOp.SUM, group=self.pg)
                cur.wait_event(ev)
                deferred = True
            else:
                for i in range(done, n):
                    dist.all_reduce(self.fp.bucket(i), op=dist.ReduceOp.SUM, group=self.pg)
                if done:
                    torch.cuda.current_stream().wait_stream(self._side)
        self._buckets_done = 0
        return deferred

    # -- one training step -------------------------------------------------------------------------
    def _step_impl(self, x, y):
        if self._defer_wgrad:
            from . import functional as F
            F.set_deferred_wgrad_join(True)
        try:
            return self._step_body(x, y)
        finally:
            if self._defer_wgrad:
                F.set_deferred_wgrad_join(False)

    def _step_body(self, x, y):
        self.fp.detach_grads()
        want_conf = isinstance(self.metrics_fn, str) and self.metrics_fn == "confusion"
        fwl = getattr(self.model, "forward_with_loss", None) if self.fused_head else None
        conf = None
        if fwl is not None and getattr(self.loss_fn, "_b200_spec", None) is not None:
            # model + loss (+ confusion counts) through one call: the head runs fused when the configuration allows
            if self.autocast_dtype is not None and x.is_cuda:
                with torch.autocast("cuda", dtype=self.autocast_dtype):
                    logits, loss, conf = fwl(x, y, self.loss_fn, want_confusion=want_conf)
            else:
                logits, loss, conf = fwl(x, y, self.loss_fn, want_confusion=want_conf)
        else:
            if self.autocast_dtype is not None and x.is_cuda:
                with torch.autocast("cuda", dtype=self.autocast_dtype):
                    out = self.model(x)
            else:
                out = self.model(x)
            logits = out[0] if isinstance(out, tuple) else out
            loss = self.loss_fn(logits.float(), y if y.dtype == torch.int64 else y.long())
        if self.accum == 1:
            from . import functional as F
            use_sinks = self.fp.flat.is_cuda and self._grad_sinks_on
            if use_sinks:
                # parameter gradients are written straight into the flat buffer by the backward kernels; the bucket all-reduces
                # are triggered when the backward of the first node AFTER a bucket starts
                cbs = {s_.data_ptr(): (lambda i=i: self._bucket_hook(i)) for i, s_ in enumerate(self.fp.sentinels)} if self.overlap else None
                F.begin_grad_sinks(self.fp.sinks, cbs)
            try:
                # a persistent seed gradient: autograd's own ones_like(loss) would be one more fill kernel per step
                if self._one is None or self._one.device != loss.device or self._one.dtype != loss.dtype:
                    self._one = torch.ones_like(loss)
                loss.backward(gradient=self._one)
                split = hasattr(self.opt, "apply_range") and hasattr(self.opt, "begin_step")
                deferred = self.allreduce_grads(defer_last=split)
            finally:
                if use_sinks:
                    F.end_grad_sinks()
            if deferred:
                # the last bucket (top encoder level) is still being reduced on the side stream: step and repack everything else
                # under it, then the few kilobytes that were waiting — the tail all-reduce is off the critical path
                lo = self.fp.bucket_elems[-2]
                rng = (self.fp.flat.data_ptr() + 4 * lo, self.fp.flat.data_ptr() + 4 * self.fp.total)
                self.opt.begin_step()
                self.opt.apply_range(0, lo, 1.0 / self.world)
                self._repack(rng, inside=False)
                torch.cuda.current_stream().wait_stream(self._side)
                self.opt.apply_range(lo, self.fp.total, 1.0 / self.world)
                self._repack(rng, inside=True)
            else:
                self.opt.step(grad_scale=1.0 / self.world)
                self._repack()
            self.fp.attach_grads()
        else:
            (loss / self.accum).backward()
            self._join_wgrads()
            self.fp.gather_grads("all", accumulate=self._micro % self.accum != 0)
            self._micro += 1
            if self._micro % self.accum == 0:
                if self.world > 1:
                    for t in self.fp.buckets():
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
                self.opt.step(grad_scale=1.0 / self.world)
                self._repack()
        if want_conf:
            if conf is None:
                from .functional import confusion_counts
                conf = confusion_counts(logits.detach(), y if y.dtype == torch.int64 else y.long())
            metrics = conf
        else:
            metrics = self.metrics_fn(logits.detach(), y) if self.metrics_fn is not None else None
        return loss.detach(), metrics

    def _repack(self, ptr_range=None, inside=True):
        # the conv kernels' bf16 weight layouts follow the new fp32 parameters: one batched kernel per step
        # (functional.repack_cached_weights) instead of one pack kernel per layer and direction
        if self.fp.flat.is_cuda:
            from . import functional as F
            F.repack_cached_weights(ptr_range, inside)

    def step(self, x, y):
        """Eager step on device tensors."""
        self.loss, self.metrics = self._step_impl(x, y)
        return self.loss

    # -- CUDA-graph path ---------------------------------------------------------------------------
    def capture(self, x_example, y_example, warmup=3):
        """Captures the whole step in one CUDA graph (static input buffers)."""
        if self.accum > 1:
            raise RuntimeError("capture() with gradient accumulation is not supported: use step() (eager launches)")
        self.static_x = x_example.clone()
        self.static_y = y_example.clone()
        # the step's main chain is captured on a HIGH-priority stream: the side-stream weight gradients (default priority,
        # needed only at the end of backward) then fill SMs the main chain leaves idle instead of delaying it
        prio = int(os.environ.get("B200_MAIN_STREAM_PRIORITY", "-1"))
        s = torch.cuda.Stream(priority=prio)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_impl(self.static_x, self.static_y)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            self.loss, self.metrics = self._step_impl(self.static_x, self.static_y)
        self.graph = g
        return g

    def replay(self, x=None, y=None):
        if x is not None:
            self.static_x.copy_(x, non_blocking=True)
        if y is not None:
            self.static_y.copy_(y, non_blocking=True)
        self._sync_hyper()
        self._refresh_derived()
        self.graph.replay()
        return self.loss

    def replay_prefetched_async(self):
        """replay_prefetched() + an asynchronous device -> host copy of the step's loss into pinned memory.  Returns a handle
        whose .value() blocks until THAT step's loss has landed: read it while the next step is already running."""
        loss = self.replay_prefetched()
        if getattr(self, "_loss_ring", None) is None:
            self._loss_ring = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._loss_events = [torch.cuda.Event() for _ in range(4)]
            self._loss_slot = 0
        i = self._loss_slot
        self._loss_slot = (i + 1) % len(self._loss_ring)
        self._loss_ring[i].copy_(loss, non_blocking=True)
        self._loss_events[i].record(torch.cuda.current_stream())
        return _LossHandle(self._loss_ring[i], self._loss_events[i])

    def load_flat(self, flat: torch.Tensor):
        """Overwrites all parameters from a flat fp32 vector laid out like ``self.fp.flat`` (e.g. to rewind to a saved point) and
        refreshes everything derived from them.  Writing ``self.fp.flat`` directly bypasses torch's version counters: call
        ``functional.weights_changed()`` afterwards if you do."""
        self.fp.flat.copy_(flat)
        if self.fp.flat.is_cuda:
            from . import functional as F
            F.weights_changed()
            F.repack_cached_weights()

    def _refresh_derived(self):
        # parameters changed outside the captured step (load_state_dict, load_flat, ...): the graph no longer repacks per layer
        if self.fp.flat.is_cuda:
            from . import functional as F
            if F.cached_weights_stale():
                F.repack_cached_weights()

    def _sync_hyper(self):
        # a scheduler (ReduceLROnPlateau, train_unet.py:381,442) edits param_groups between steps: push lr to the device word
        sync = getattr(self.opt, "sync_hyper", None)
        if sync is not None:
            sync()

    # -- input pipeline: the next batch crosses PCIe while the current step computes -----------------
    def prefetch(self, x_host, y_host):
        """Starts the host -> device copy of the NEXT step's batch (pinned host tensors) on a copy stream.
        The batch lands in a staging pair; replay_prefetched() consumes it.  One batch in flight at a time."""
        if self.graph is None:
            raise RuntimeError("prefetch() needs a captured step: call capture() first")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.static_x.device)
            self._stage = (torch.empty_like(self.static_x), torch.empty_like(self.static_y))
            self._stage_ready, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._stage_free)  # the previous consumer has read the staging pair
        with torch.cuda.stream(cs):
            self._stage[0].copy_(x_host, non_blocking=True)
            self._stage[1].copy_(y_host, non_blocking=True)
            self._stage_ready.record(cs)

    def replay_prefetched(self):
        """Runs the captured step on the batch staged by the last prefetch()."""
        if self._stage is None:
            raise RuntimeError("replay_prefetched() without a prefetch()")
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready)
        self.static_x.copy_(self._stage[0])
        self.static_y.copy_(self._stage[1])
        self._stage_free.record(cur)
        self._sync_hyper()
        self._refresh_derived()
        self.graph.replay()
        return self.loss
