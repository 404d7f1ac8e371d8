"""Step runners: the calls a user makes to train or screen with the B200 hot path.

`TrainStep` = one optimisation step of a GLAM model (forward + loss + backward + optimizer, and for
world_size > 1 one NCCL all-reduce over a single flat fp32 gradient bucket), captured
once into a CUDA graph and replayed: at the reference's batch sizes a step is ~100 small kernels, so
launch overhead — not the kernels — would otherwise set the pace.  `ScreenStep` = eval-mode forward
for virtual screening (no collective: graphs shard by molecule across ranks).

Host batches (`glam_b200.synth.GraphBatch`, i.e. the fields of a PyG Batch) are copied into fixed device
buffers, so a captured step has one shape (N, E, B).  `ScreenStep.step` pads smaller batches to it with dummy
graphs behind the real ones (`synth.pad_graph_batch`) — one capture serves a loader with varying molecule sizes;
`TrainStep` does the same when captured on a padded example with `masked_loss` (dummy graphs weigh zero).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import graph as G
from .synth import GraphBatch


def _on_device(model: torch.nn.Module, device: torch.device) -> torch.nn.Module:
    """Move the model only if it is not there yet.  `module.to()` is NOT a no-op for a model that already lives on the device:
    nn.GRU / nn.LSTM re-flatten their weights into a freshly allocated buffer on every `_apply`, i.e. the parameters MOVE —
    which would leave every previously captured CUDA graph (and FlatAdam's parameter views) pointing at freed memory."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else None)
    there = all(t.device.type == dev.type and (idx is None or t.device.index == idx)
                for t in list(model.parameters()) + list(model.buffers()))
    return model if there else model.to(dev)


def _tup(b):
    """A step's input is one GraphBatch (GLAM-GP) or a tuple of them (the two towers of GLAM-DDI / GLAM-DTI: the model is
    called as model(*batches), the target is the first batch's y)."""
    return tuple(b) if isinstance(b, (tuple, list)) else (b,)


_GRAPH_FIELDS = ("x", "edge_index", "edge_attr", "batch", "y", "mask")


def _fields(g):
    from .packed import FIELDS, PackedBatch
    return FIELDS if isinstance(g, PackedBatch) else _GRAPH_FIELDS


def _static_like(b, device):
    from .packed import PackedBatch
    z = lambda t: None if t is None else torch.empty_like(t, device=device)
    return tuple(g._map(z) if isinstance(g, PackedBatch) else
                 GraphBatch(z(g.x), z(g.edge_index), z(g.edge_attr), z(g.batch), z(g.y), g.num_graphs, z(g.mask)) for g in _tup(b))


def _model_args(static):
    """What the model is called with: packed batches (glam_b200/packed.py) are unpacked on the device first — inside the
    captured graph, so a replay consumes whatever packed bytes were copied into the static buffers."""
    from .packed import PackedBatch
    return tuple(g.unpack() if isinstance(g, PackedBatch) else g for g in static)


def _copy_into(dst, src) -> int:
    """Async copies of every field (H2D from pinned memory or D2D); returns bytes moved."""
    n = 0
    dst, src = _tup(dst), _tup(src)
    if len(dst) != len(src):
        raise ValueError(f"the captured step takes {len(dst)} graph batches per call, got {len(src)}")
    for dg, sg in zip(dst, src):
        for name in _fields(sg):
            s, d = getattr(sg, name), getattr(dg, name)
            if s is None:
                continue
            if s.shape != d.shape:
                raise ValueError(f"batch field {name} has shape {tuple(s.shape)}, the captured step expects {tuple(d.shape)}")
            d.copy_(s, non_blocking=True)
            n += s.numel() * s.element_size()
    return n


def batch_nbytes(b) -> int:
    return sum(g.nbytes() for g in _tup(b))


def masked_loss(fn: Callable) -> Callable:
    """`fn(out, y, reduction="none")` (e.g. torch.nn.functional.mse_loss) -> loss(out, y, mask): the mean over the REAL graphs of a
    padded batch (mask = 1 real / 0 dummy, synth.pad_graph_batch).  With it a TrainStep captured on a padded example takes
    batches of varying size: the dummy graphs carry zero weight, so they contribute neither loss nor gradient."""
    def loss(out, y, mask):
        per = fn(out, y, reduction="none")
        per = per.reshape(per.shape[0], -1).mean(1)
        return (per * mask).sum() / mask.sum()
    return loss


class FlatGrads:
    """All parameter gradients in one flat fp32 buffer (one all-reduce per step).

    gather=False: every p.grad is a view into the buffer and autograd accumulates in place (one add kernel per
    parameter per step).  gather=True: p.grad is cleared before backward, autograd hands over each gradient tensor
    as is, and collect() packs them into the buffer with one batched copy — ~20 launches fewer per step."""

    def __init__(self, params, gather: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.gather = gather
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        if not gather:
            off = 0
            for p in self.params:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()

    def zero(self):
        if self.gather:
            for p in self.params:
                p.grad = None
            if self._early_hook is not None:
                self._early_pending[0] = self._early_n          # re-arm the early bucket's countdown
        else:
            self.flat.zero_()

    def collect(self):
        if self.gather:
            late = self.params[:self._early_at] if self._early_work is not None else self.params
            if late:
                parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in late]
                torch.cat(parts, out=self.flat[:sum(p.numel() for p in late)])

    def all_reduce_mean(self, world: int):
        self.all_reduce_sum()
        self.flat.mul_(1.0 / world)

    def all_reduce_sum(self):
        """Sum over ranks.  With an early bucket (enable_early_bucket) the tail of the buffer — the parameters behind the message
        stack, whose gradients are final before the stack's backward starts — is already being reduced on the collective's
        own stream; only the head goes now, then both are joined."""
        import torch.distributed as dist
        if self._early_work is not None:
            if self._early_off > 0:
                dist.all_reduce(self.flat[:self._early_off], op=dist.ReduceOp.SUM)
            self._early_work.wait()
            self._early_work = None
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    # ---- overlap of the collective with backward ---------------------------------------------------------------
    _early_at = 0            # index of the first early parameter
    _early_off = 0           # its offset in the flat buffer
    _early_work = None       # the in-flight reduction of flat[_early_off:]
    _early_hook = None
    stream = None            # the stream the step runs on (set by the step runner before backward)

    def enable_early_bucket(self, model: torch.nn.Module, block_types: tuple) -> bool:
        """Split the bucket at the last parameter of the last `block_types` module (the message stack): everything registered
        after it (readout, flat / output LinearBlocks: ~85 % of the bytes of a GLAM model) has its gradient BEFORE the stack's
        backward — the bulk of the step — starts, so its all-reduce is issued from a gradient hook and runs on the collective's
        stream under the stack's backward.  Needs gather=True.  Returns False (and changes nothing) when there is no such split.
        The caller checks the readiness order once with `check_early_order()` before relying on it."""
        import torch.distributed as dist
        if not self.gather or self._early_hook is not None:
            return self._early_hook is not None
        blocked = {id(p) for m in model.modules() if isinstance(m, block_types) for p in m.parameters()}
        last = max((i for i, p in enumerate(self.params) if id(p) in blocked), default=-1)
        if last < 0 or last + 1 >= len(self.params):
            return False
        self._early_at = last + 1
        self._early_off = sum(p.numel() for p in self.params[:self._early_at])
        early = self.params[self._early_at:]

        pending = [len(early)]

        def fire():
            # (hooks run on the autograd thread: name the step's stream explicitly — a CUDA-graph capture refuses anything
            # that lands on the legacy stream)
            import contextlib
            ctx = torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
            with ctx:
                parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in early]
                torch.cat(parts, out=self.flat[self._early_off:])
                self._early_work = dist.all_reduce(self.flat[self._early_off:], op=dist.ReduceOp.SUM, async_op=True)

        def on_ready(_p):
            pending[0] -= 1
            if pending[0] == 0:
                fire()

        self._early_pending = pending
        self._early_n = len(early)
        self._early_hook = [p.register_post_accumulate_grad_hook(on_ready) for p in early]
        return True

    def disable_early_bucket(self):
        for h in (self._early_hook or ()):
            h.remove()
        self._early_hook, self._early_work, self._early_at, self._early_off = None, None, 0, 0

    def check_early_order(self, backward: Callable) -> bool:
        """Run `backward()` once with a readiness probe on every parameter: True iff every early parameter's gradient was
        final before the first gradient of a late parameter."""
        order, handles = [], []
        for i, p in enumerate(self.params):
            handles.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: order.append(i)))
        try:
            backward()
        finally:
            for h in handles:
                h.remove()
        first_late = next((k for k, i in enumerate(order) if i < self._early_at), len(order))
        return len(set(order[:first_late])) == len(self.params) - self._early_at


class FlatAdam:
    """torch.optim.Adam(params, lr) (the reference trainer's optimizer, src_1gp/trainer.py:49-50) on flat buffers:
    the parameters are re-pointed at views of one fp32 buffer, and a step is one library call over
    (parameters, gradient bucket, exp_avg, exp_avg_sq).  `lr` is a device scalar: set_lr() works between CUDA-graph
    replays, which is what a ReduceLROnPlateau scheduler needs."""

    def __init__(self, grads: FlatGrads, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        from . import ops
        self._ops = ops
        self.grads = grads
        dev = grads.flat.device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on the CUDA library only (no CPU path)")
        self.flat = torch.empty_like(grads.flat)
        off = 0
        with torch.no_grad():
            for p in grads.params:
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view_as(p)
                off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.state = torch.zeros(3, dtype=torch.float32, device=dev)
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay

    def set_lr(self, lr: float):
        self.lr.fill_(float(lr))

    def step(self, grad_scale: float = 1.0):
        self._ops.adam_step(self.flat, self.grads.flat, self.exp_avg, self.exp_avg_sq, self.lr, self.state, self.betas[0],
                            self.betas[1], self.eps, self.weight_decay, grad_scale)


class TrainStep:
    """One optimisation step (forward + loss + backward [+ all-reduce] + Adam) captured in a CUDA graph.

    double_buffer=True keeps TWO sets of static input buffers, each with its own captured graph (sharing one memory
    pool: they never run concurrently).  `step(batch, prefetch=next_batch)` then enqueues the host->device copy of the
    NEXT batch on a copy stream while the current step computes, so the PCIe transfer disappears from the step time —
    what a DataLoader with pinned memory and `non_blocking=True` does for the reference's `data.to(device)`
    (src_1gp/trainer.py:175), which the reference itself leaves synchronous."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example: GraphBatch, lr: float = 1e-3,
                 device="cuda", world_size: int = 1, use_cuda_graph: bool = True, warmup: int = 3,
                 double_buffer: bool = False, overlap_allreduce: bool = False):
        self.model = _on_device(model, device)
        self.loss_fn = loss_fn
        self.device = torch.device(device)
        self.world = world_size
        self.statics = [_static_like(example, self.device) for _ in range(2 if double_buffer else 1)]
        for st in self.statics:
            _copy_into(st, example)
        self.static = self.statics[0]
        self.grads = FlatGrads(self.model.parameters(), gather=True)
        self.opt = FlatAdam(self.grads, lr=lr)
        # overlap_allreduce (opt-in): reduce the gradients of everything behind the message stack from a gradient hook, under the
        # stack's backward.  Measured on 8 B200 it LOSES: 1.85 ms per step against 1.57 ms — the fused kernels are persistent
        # grids of one CTA per SM with nearly all of its shared memory, so the collective's CTAs either wait for SMs or push part
        # of the grid into a second wave.  The single bucket after backward (0.43 MB, latency-bound) stays the default.
        self.overlap_allreduce = False
        if world_size > 1 and overlap_allreduce:
            from .layer import MessageBlock
            self.overlap_allreduce = self.grads.enable_early_bucket(self.model, (MessageBlock,))
        self.loss = torch.zeros((), device=self.device)
        self.graphs = []
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_cuda_graph = use_cuda_graph
        self._slot = 0
        self._prefetched = None                                   # host batch whose copy into slot `_slot` is in flight
        if double_buffer:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [torch.cuda.Event() for _ in self.statics]
            self._consumed = [torch.cuda.Event() for _ in self.statics]
        if use_cuda_graph:
            self._capture(warmup)

    def _body(self, static=None):
        static = self.static if static is None else static
        self.grads.zero()
        if self.overlap_allreduce:
            self.grads.stream = torch.cuda.current_stream(self.device)
        out = self.model(*static)
        m = getattr(static[0], "mask", None)                    # padded batches: the loss weights the real graphs (masked_loss)
        loss = self.loss_fn(out, static[0].y) if m is None else self.loss_fn(out, static[0].y, m)
        loss.backward()
        self.grads.collect()
        if self.world > 1:
            self.grads.all_reduce_sum()
        self.opt.step(grad_scale=1.0 / self.world)              # mean over ranks folded into the gradient load
        self.loss.copy_(loss.detach())

    def _capture(self, warmup: int):
        # the warm-up iterations (allocator / cuBLAS / NCCL initialisation outside capture) are real optimisation steps:
        # snapshot parameters, both Adam moments and the step counter first and put them back afterwards, so that
        # constructing a TrainStep leaves the model and the optimizer exactly as the caller built them
        # (torch.optim.Adam starts from step 0, src_1gp/trainer.py:49-50)
        snap = [t.clone() for t in (self.opt.flat, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.state)]
        bufs = [(b, b.clone()) for b in self.model.buffers()]
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            if self.overlap_allreduce:
                # one probe step: the early bucket may only be used if its gradients really are final before the first late one
                def probe():
                    self.grads.zero()
                    self.grads.stream = torch.cuda.current_stream(self.device)
                    self.loss_fn(self.model(*self.static), self.static[0].y).backward()
                    self.grads.collect()
                    self.grads.all_reduce_sum()
                if not self.grads.check_early_order(probe):
                    self.grads.disable_early_bucket()
                    self.overlap_allreduce = False
            for _ in range(max(warmup, 1)):
                G.clear_caches()
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        with torch.no_grad():
            for dst, src in zip((self.opt.flat, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.state), snap):
                dst.copy_(src)
            for b, saved in bufs:
                b.copy_(saved)
        pool = None
        for st in self.statics:
            G.clear_caches()                  # the index build must be part of the captured step
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._body(st)
            pool = g.pool()
            self.graphs.append(g)
        self.graph = self.graphs[0]
        G.clear_caches()

    def run_resident(self):
        """One step on whatever is in the static buffers of the current slot (inputs already in HBM)."""
        if self.graphs:
            self.graphs[self._slot].replay()
        else:
            G.clear_caches()
            self._body(self.statics[self._slot])
        return self.loss

    def _fit(self, batch):
        """A TrainStep captured on a PADDED example (its static batch has a mask) takes any batch that fits: it is padded with dummy
        graphs of zero loss weight (synth.pad_graph_batch; the loss must be a masked one, engine.masked_loss)."""
        st = self.static[0]
        if isinstance(batch, GraphBatch) and isinstance(st, GraphBatch) and st.mask is not None:
            from .synth import pad_graph_batch
            if batch.mask is None or (batch.num_nodes, batch.num_edges, batch.num_graphs) != (st.num_nodes, st.num_edges, st.num_graphs):
                return pad_graph_batch(batch, st.num_nodes, st.num_edges, st.num_graphs, with_mask=True)
        return batch

    def load(self, batch: GraphBatch) -> int:
        return _copy_into(self.statics[self._slot], self._fit(batch))

    def step(self, batch: GraphBatch, prefetch: Optional[GraphBatch] = None) -> torch.Tensor:
        """Public API: host (pinned) or device batch in, loss tensor (device scalar) out.  `prefetch` (double_buffer
        only) names the batch of the NEXT call: its copy overlaps this step's compute."""
        if len(self.statics) == 1:
            self.load(batch)
            return self.run_resident()
        main = torch.cuda.current_stream(self.device)
        slot = self._slot
        if self._prefetched is not None:
            main.wait_event(self._copied[slot])                   # the copy enqueued during the previous step (also when the
        if self._prefetched is not batch:                         # caller then passes a different batch: it overwrites it)
            _copy_into(self.statics[slot], self._fit(batch))
        self._prefetched = None
        self.run_resident()
        self._consumed[slot].record(main)
        if prefetch is not None:
            nxt = slot ^ 1
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._consumed[nxt])  # the last step that read those buffers is done
                _copy_into(self.statics[nxt], self._fit(prefetch))
                self._copied[nxt].record(self._copy_stream)
            self._prefetched = prefetch
            self._slot = nxt
        return self.loss


class ScreenStep:
    """Eval-mode forward for virtual screening; returns scores [B, out_dim] (device).

    double_buffer=True: two sets of static input buffers and two captured graphs; `step(batch, prefetch=next_batch)`
    copies the next host batch on a copy stream while the current one computes (same scheme as TrainStep).  Each slot
    has its own output tensor, so the scores of step i stay valid while step i+1 runs."""

    def __init__(self, model: torch.nn.Module, example: GraphBatch, device="cuda", use_cuda_graph: bool = True,
                 warmup: int = 3, double_buffer: bool = False):
        self.model = _on_device(model, device).eval()
        self.device = torch.device(device)
        self.statics = [_static_like(example, self.device) for _ in range(2 if double_buffer else 1)]
        for st in self.statics:
            _copy_into(st, example)
        self.static = self.statics[0]
        self.outs = [None] * len(self.statics)
        self.graphs = []
        self._slot = 0
        self._prefetched = None
        if double_buffer:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [torch.cuda.Event() for _ in self.statics]
            self._consumed = [torch.cuda.Event() for _ in self.statics]
        with torch.no_grad():
            if use_cuda_graph:
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    for _ in range(max(warmup, 1)):
                        G.clear_caches()
                        self.model(*_model_args(self.static))
                torch.cuda.current_stream(self.device).wait_stream(s)
                torch.cuda.synchronize(self.device)
                pool = None
                for i, st in enumerate(self.statics):
                    G.clear_caches()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        self.outs[i] = self.model(*_model_args(st))
                    pool = g.pool()
                    self.graphs.append(g)
                G.clear_caches()
        self.graph = self.graphs[0] if self.graphs else None

    @property
    def out(self):
        return self.outs[self._slot]

    def run_resident(self):
        if self.graphs:
            self.graphs[self._slot].replay()
        else:
            with torch.no_grad():
                G.clear_caches()
                self.outs[self._slot] = self.model(*_model_args(self.statics[self._slot]))
        return self.outs[self._slot]

    def _fit(self, batch):
        """Batches smaller than the captured shape are padded with dummy graphs behind the real ones (synth.pad_graph_batch), so a
        loader with varying molecule sizes / a short last batch replays the same captured step; returns (batch, real graphs)."""
        from .synth import pad_graph_batch
        if isinstance(batch, GraphBatch):
            st = self.static[0]
            if isinstance(st, GraphBatch) and (batch.num_nodes, batch.num_edges, batch.num_graphs) != (st.num_nodes, st.num_edges, st.num_graphs):
                return pad_graph_batch(batch, st.num_nodes, st.num_edges, st.num_graphs), batch.num_graphs
        return batch, None

    def step(self, batch: GraphBatch, prefetch: Optional[GraphBatch] = None) -> torch.Tensor:
        """Scores of `batch` (device tensor; valid until the slot is reused).  A single GraphBatch smaller than the captured
        example is padded to it and the scores of its real graphs are returned."""
        fitted, real = self._fit(batch)
        if len(self.statics) == 1:
            _copy_into(self.static, fitted)
            out = self.run_resident()
            return out if real is None else out[:real]
        main = torch.cuda.current_stream(self.device)
        slot = self._slot
        if self._prefetched is not None:
            main.wait_event(self._copied[slot])
        if self._prefetched is not batch:
            _copy_into(self.statics[slot], fitted)
        self._prefetched = None
        out = self.run_resident()
        self._consumed[slot].record(main)
        if prefetch is not None:
            nxt = slot ^ 1
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._consumed[nxt])
                _copy_into(self.statics[nxt], self._fit(prefetch)[0])
                self._copied[nxt].record(self._copy_stream)
            self._prefetched = prefetch
            self._slot = nxt
        return out if real is None else out[:real]
