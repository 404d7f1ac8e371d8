"""Packed graph store: the screening input format.

The reference featurises a molecule once (`src_1gp/dataset.py:60-97`) and then ships, for every batch, fp32 features,
an int64 `edge_index`, fp32 one-hot `edge_attr` and an int64 `batch` vector to the device (`src_1gp/trainer.py:175`),
where PyG re-derives gather indices on every layer call.  A molecule's graph never changes, so `pack_batch` fixes at pack
time what the kernels read — the destination-sorted in-edge lists (`argsort(edge_index[1], stable=True)` order, the same
order `glam_build_csr` produces), bond types, degrees — and stores it in the narrowest integer types that hold it:

    n_g uint8 [B] | e_g uint16 [B] | deg uint8 [N] | nbr uint8 [E] (graph-local source) | etype uint8 [E] | xq uint8 [N,F]

~360 B per 25-atom molecule instead of ~2.6 KB.  `PackedBatch.unpack()` (device) is two scans + one warp per graph
(`csrc/packed.cu`) and returns an object with the PyG field names whose `edge_index` / `edge_attr` / `batch` carry the
prebuilt index, so `glam_b200.model` / `glam_b200.layer` take it unchanged; no CSR build, no edge_attr gather, no type scan.
Packing validates its preconditions (<= 255 atoms per graph, one-hot bond features, small-integer atom features, every edge
inside its graph) and raises otherwise — such batches go through the regular fields.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import graph as G
from . import ops

FIELDS = ("n_g", "e_g", "deg", "nbr", "etype", "xq", "y")


@dataclass
class PackedBatch:
    n_g: torch.Tensor            # uint8  [B]
    e_g: torch.Tensor            # int16  [B] (bit pattern of uint16)
    deg: torch.Tensor            # uint8  [N]
    nbr: torch.Tensor            # uint8  [E]
    etype: torch.Tensor          # uint8  [E]
    xq: torch.Tensor             # uint8  [N, F]
    y: Optional[torch.Tensor]
    num_graphs: int
    edge_dim: int

    def _map(self, fn):
        return PackedBatch(*[None if getattr(self, f) is None else fn(getattr(self, f)) for f in FIELDS], self.num_graphs, self.edge_dim)

    def to(self, device, non_blocking: bool = False) -> "PackedBatch":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "PackedBatch":
        return self._map(lambda t: t.pin_memory())

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in FIELDS if getattr(self, f) is not None)

    @property
    def num_nodes(self) -> int:
        return self.deg.shape[0]

    @property
    def num_edges(self) -> int:
        return self.nbr.shape[0]

    def unpack(self) -> "UnpackedBatch":
        """Device side: the index and fp32 features the kernels read (csrc/packed.cu); safe inside CUDA-graph capture."""
        dev = self.deg.device
        if dev.type != "cuda":
            raise ops._lib.GlamError("PackedBatch.unpack needs device tensors (there is no CPU path)")
        B, N, E, F = self.num_graphs, self.num_nodes, self.num_edges, self.xq.shape[1]
        i32 = dict(dtype=torch.int32, device=dev)
        gptr, eptr = torch.empty(B + 1, **i32), torch.empty(B + 1, **i32)
        rowptr, dst_src = torch.empty(N + 1, **i32), torch.empty(max(E, 1), **i32)
        x = torch.empty((N, F), dtype=torch.float32, device=dev)
        p = ops._p
        ops._call("glam_unpack_graphs", p(self.n_g), p(self.e_g), p(self.deg), p(self.nbr), p(self.xq), B, N, E, F, p(gptr), p(eptr),
                  p(rowptr), p(dst_src), p(x), ops._stream(x))
        index = PrebuiltIndex(N, E, B, rowptr, dst_src[:E], self.etype, gptr, self.edge_dim)
        return UnpackedBatch(x, index, PrebuiltEdgeAttr(index), PrebuiltBatchVector(index), self.y, B)


class PrebuiltIndex:
    """The part of graph.GraphIndex the forward (eval) kernels read, delivered by the packed store instead of being rebuilt
    from edge_index: dst_rowptr / dst_src (dst-sorted CSR), bond types, graph offsets, and the fused kernel's tile table."""

    def __init__(self, N, E, B, rowptr, dst_src, etype, gptr, edge_dim):
        self.num_nodes, self.num_edges, self.num_graphs, self.edge_dim = int(N), int(E), int(B), int(edge_dim)
        self.dst_rowptr, self.dst_src, self.etype, self.gptr = rowptr, dst_src, etype, gptr
        self.dst_tiles = self.src_tiles = None
        self._fused = None
        self._ea = None

    def __getattr__(self, name):
        if name in ("dst_perm", "dst_dst", "src_rowptr", "src_pos", "src_dst"):
            raise ops._lib.GlamError(f"packed batches carry the forward index only ({name} is needed by backward): "
                                     "train from the regular PyG fields")
        raise AttributeError(name)

    def fused_index(self, gptr=None, num_graphs=None, edge_attr=None):
        if self._fused is None:
            meta = torch.zeros(4, dtype=torch.int32, device=self.dst_rowptr.device)
            # pack_batch has verified on the host what the tile builder's check kernel would (edges stay inside their graph,
            # bond features one-hot); graphs larger than a tile still set meta[1] and poison the kernel's outputs
            tiles = ops.build_graph_tiles(self.gptr, self.num_graphs, self, meta, check_edges=False)
            self._fused = G.FusedIndex(tiles, meta, self.etype, self.edge_dim)
        return self._fused

    def sorted_edge_attr(self, edge_attr=None):
        """fp32 one-hot rows in dst order, materialised only for the per-op kernels (the fused kernel reads the types)."""
        if self._ea is None:
            self._ea = torch.nn.functional.one_hot(self.etype[:self.num_edges].long(), self.edge_dim).to(torch.float32)
        return self._ea


class PrebuiltEdgeAttr:
    """Stands in for `edge_attr` [E, De]: carries the prebuilt index; only its shape is ever looked at on the packed path."""

    def __init__(self, index: PrebuiltIndex):
        self.index = index
        self.shape = (index.num_edges, index.edge_dim)

    def dim(self):
        return 2


class PrebuiltBatchVector:
    """Stands in for the PyG `batch` vector: graph.graph_ptr() answers from the prebuilt offsets."""

    def __init__(self, index: PrebuiltIndex):
        self.index = index


@dataclass
class UnpackedBatch:
    x: torch.Tensor
    edge_index: PrebuiltIndex
    edge_attr: PrebuiltEdgeAttr
    batch: PrebuiltBatchVector
    y: Optional[torch.Tensor]
    num_graphs: int


def pack_batch(b, max_type: int = 4) -> PackedBatch:
    """Host side, numpy: GraphBatch / PyG-Batch-like (x, edge_index, edge_attr, batch[, y, num_graphs]) -> PackedBatch."""
    x = b.x.detach().cpu().numpy()
    ei = b.edge_index.detach().cpu().numpy()
    ea = b.edge_attr.detach().cpu().numpy()
    bt = b.batch.detach().cpu().numpy()
    N, E = x.shape[0], ei.shape[1]
    B = int(getattr(b, "num_graphs", 0) or (bt.max() + 1 if N else 0))
    if ea.ndim != 2:
        raise ValueError("pack_batch: edge_attr must be [E, De]")
    if np.any(np.diff(bt) < 0):
        raise ValueError("pack_batch: `batch` must be non-decreasing (PyG Batch.from_data_list order)")
    n_g = np.bincount(bt, minlength=B).astype(np.int64)
    if n_g.max(initial=0) > 255:
        raise ValueError("pack_batch: a graph has more than 255 atoms")
    gstart = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(n_g, out=gstart[1:])
    src, dst = ei[0], ei[1]
    if E and (np.any(bt[src] != bt[dst])):
        raise ValueError("pack_batch: an edge connects two different graphs")
    perm = np.argsort(dst, kind="stable")                       # = glam_build_csr's dst_perm
    src_s, dst_s, ea_s = src[perm], dst[perm], ea[perm]
    deg = np.bincount(dst, minlength=N)
    if deg.max(initial=0) > 255:
        raise ValueError("pack_batch: an atom has more than 255 in-edges")
    e_g = np.bincount(bt[dst], minlength=B)
    if e_g.max(initial=0) > 65535:
        raise ValueError("pack_batch: a graph has more than 65535 directed edges")
    onehot = (ea_s.sum(1) == 1) & (ea_s.max(1) == 1) & (ea_s.min(1) == 0) if E else np.ones(0, bool)
    if not onehot.all() or ea.shape[1] > max_type:
        raise ValueError("pack_batch: edge_attr rows must be exact one-hot bond types (src_1gp/dataset.py:82) with <= 4 types")
    etype = ea_s.argmax(1).astype(np.uint8) if E else np.zeros(0, np.uint8)
    nbr = (src_s - gstart[bt[dst_s]]).astype(np.uint8) if E else np.zeros(0, np.uint8)
    xr = np.rint(x)
    if not (np.array_equal(xr, x) and x.min(initial=0) >= 0 and x.max(initial=0) <= 255):
        raise ValueError("pack_batch: atom features must be integers in [0, 255] (one-hot / count features, src_1gp/dataset.py:92-95)")
    t = torch.from_numpy
    y = None if getattr(b, "y", None) is None else b.y.detach().cpu().clone()
    return PackedBatch(t(n_g.astype(np.uint8)), t(e_g.astype(np.uint16).view(np.int16)), t(deg.astype(np.uint8)), t(nbr), t(etype),
                       t(x.astype(np.uint8)), y, B, int(ea.shape[1]))
