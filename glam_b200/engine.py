"""Step runners: the calls a user makes to train or screen with the B200 hot path.

`TrainStep` = one optimisation step of a GLAM model (forward + loss + backward + optimizer, and for
world_size > 1 one NCCL all-reduce over a single flat fp32 gradient bucket), captured
once into a CUDA graph and replayed: at the reference's batch sizes a step is ~100 small kernels, so
launch overhead — not the kernels — would otherwise set the pace.  `ScreenStep` = eval-mode forward
for virtual screening (no collective: graphs shard by molecule across ranks).

Host batches (`glam_b200.synth.GraphBatch`, i.e. the fields of a PyG Batch) are copied into fixed device
buffers, so every batch of a run must have the same (N, E, B) — the synthetic generator can emit
fixed-size batches; a real loader would bucket by size and keep one captured graph per bucket.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import graph as G
from .synth import GraphBatch


def _static_like(b: GraphBatch, device) -> GraphBatch:
    z = lambda t: None if t is None else torch.empty_like(t, device=device)
    return GraphBatch(z(b.x), z(b.edge_index), z(b.edge_attr), z(b.batch), z(b.y), b.num_graphs)


def _copy_into(dst: GraphBatch, src: GraphBatch) -> int:
    """Async copies of every field (H2D from pinned memory or D2D); returns bytes moved."""
    n = 0
    for name in ("x", "edge_index", "edge_attr", "batch", "y"):
        s, d = getattr(src, name), getattr(dst, name)
        if s is None:
            continue
        if s.shape != d.shape:
            raise ValueError(f"batch field {name} has shape {tuple(s.shape)}, the captured step expects {tuple(d.shape)}")
        d.copy_(s, non_blocking=True)
        n += s.numel() * s.element_size()
    return n


class FlatGrads:
    """All parameter gradients as views into one flat fp32 buffer (one all-reduce per step)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, world: int):
        import torch.distributed as dist
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / world)


class TrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example: GraphBatch, lr: float = 1e-3,
                 device="cuda", world_size: int = 1, use_cuda_graph: bool = True, warmup: int = 3):
        self.model = model.to(device)
        self.loss_fn = loss_fn
        self.device = torch.device(device)
        self.world = world_size
        self.static = _static_like(example, self.device)
        _copy_into(self.static, example)
        self.grads = FlatGrads(self.model.parameters())
        self.opt = torch.optim.Adam(self.grads.params, lr=lr, capturable=use_cuda_graph, fused=True)
        self.loss = torch.zeros((), device=self.device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_cuda_graph = use_cuda_graph
        if use_cuda_graph:
            self._capture(warmup)

    def _body(self):
        self.grads.zero()
        out = self.model(self.static)
        loss = self.loss_fn(out, self.static.y)
        loss.backward()
        if self.world > 1:
            self.grads.all_reduce_mean(self.world)
        self.opt.step()
        self.loss.copy_(loss.detach())

    def _capture(self, warmup: int):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(max(warmup, 1)):
                G.clear_caches()
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        G.clear_caches()                      # the index build must be part of the captured step
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        G.clear_caches()

    def run_resident(self):
        """One step on whatever is in the static buffers (inputs already in HBM)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            G.clear_caches()
            self._body()
        return self.loss

    def load(self, batch: GraphBatch) -> int:
        return _copy_into(self.static, batch)

    def step(self, batch: GraphBatch) -> torch.Tensor:
        """Public API: host (pinned) or device batch in, loss tensor (device scalar) out."""
        self.load(batch)
        return self.run_resident()


class ScreenStep:
    """Eval-mode forward for virtual screening; returns scores [B, out_dim] (device)."""

    def __init__(self, model: torch.nn.Module, example: GraphBatch, device="cuda", use_cuda_graph: bool = True,
                 warmup: int = 3):
        self.model = model.to(device).eval()
        self.device = torch.device(device)
        self.static = _static_like(example, self.device)
        _copy_into(self.static, example)
        self.out = None
        self.graph = None
        with torch.no_grad():
            if use_cuda_graph:
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    for _ in range(max(warmup, 1)):
                        G.clear_caches()
                        self.model(self.static)
                torch.cuda.current_stream(self.device).wait_stream(s)
                torch.cuda.synchronize(self.device)
                G.clear_caches()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out = self.model(self.static)
                G.clear_caches()

    def run_resident(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                G.clear_caches()
                self.out = self.model(self.static)
        return self.out

    def step(self, batch: GraphBatch) -> torch.Tensor:
        _copy_into(self.static, batch)
        return self.run_resident()
