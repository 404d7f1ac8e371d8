"""Per-batch graph index: the dst-/src-sorted CSR and per-graph node ranges.

The reference re-derives gather indices and atomics-scatters on every layer call (PyG
`MessagePassing.propagate`, reached from src_1gp/layer.py:40,86).  Here the index is built once per
batch on the device and shared by all `message_steps` applications of the block and by backward
(SURVEY.md §8b "Ownership").  Lookups are keyed on the identity *and version* of the caller's
`edge_index` tensor, so the drop-in `forward(x, edge_index, edge_attr)` signature stays unchanged;
the cache holds a reference to the key tensor, so its storage cannot be recycled for another batch
while the entry is alive, and an in-place refill (`copy_`) bumps the version and rebuilds.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from . import ops

_CACHE_SIZE = 16


class GraphIndex:
    """Device-resident index of one batch (all int32; see include/glam_b200.h (1))."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        self.num_nodes = int(num_nodes)
        self.num_edges = int(edge_index.shape[1])
        csr = ops.build_csr(edge_index, self.num_nodes)
        self.dst_rowptr = csr["dst_rowptr"]
        self.dst_src = csr["dst_src"]
        self.dst_perm = csr["dst_perm"]
        self.dst_dst = csr["dst_dst"]
        self.src_rowptr = csr["src_rowptr"]
        self.src_pos = csr["src_pos"]
        self.src_dst = csr["src_dst"]
        self._csr = csr
        self._edge_tiles = None
        self._edge_attr_key = None
        self._edge_attr_sorted = None
        self._gcn = None
        self._nn_key = None
        self._nn = None
        self._fused_key = None
        self._fused = None

    def _tiles(self):
        # tile descriptors of the windowed per-op edge kernels: built on first use (the one-launch kernels never read them)
        if self._edge_tiles is None:
            self._edge_tiles = ops.build_edge_tiles(self._csr, self.num_nodes)
        return self._edge_tiles

    @property
    def dst_tiles(self):
        return self._tiles()[0]

    @property
    def src_tiles(self):
        return self._tiles()[1]

    def fused_index(self, gptr: torch.Tensor, num_graphs: int, edge_attr: torch.Tensor):
        """Index of the fused message kernel (csrc/mp_fused.cu) for this batch, or None when the batch does not meet its
        preconditions (a graph with more rows / in-edges than a tile, an edge between graphs, edge_attr not one-hot).
        Built once per batch.  The check reads one int back outside CUDA-graph capture; during capture the verdict of
        the last eager batch of the same shape is assumed, and the kernel poisons its outputs with NaN if it is wrong."""
        key = (gptr.data_ptr(), gptr._version, int(num_graphs), edge_attr.data_ptr(), edge_attr._version, tuple(edge_attr.shape))
        if self._fused_key == key:
            return self._fused
        # the kernels only need the bond TYPE per dst-ordered edge: read it through dst_perm from the caller's rows (fp32,
        # contiguous) — the dst-ordered copy of edge_attr is made only for callers that want the rows themselves
        ea = edge_attr if edge_attr.dim() == 2 else edge_attr.view(edge_attr.shape[0], -1)
        meta = torch.zeros(4, dtype=torch.int32, device=ea.device)
        if ea.dtype == torch.float32 and ea.is_contiguous():
            etype = ops.edge_types(ea, meta, perm=self.dst_perm)
        else:
            ea = self.sorted_edge_attr(edge_attr)
            etype = ops.edge_types(ea, meta)
        tiles = ops.build_graph_tiles(gptr, num_graphs, self, meta)
        shape_key = (self.num_nodes, self.num_edges, int(num_graphs), ea.shape[1], ea.device)
        if _capturing():
            ok = _fused_ok.get(shape_key, False)
        else:
            ok = int(meta[1].item()) == 0
            _fused_ok[shape_key] = ok
        self._fused = FusedIndex(tiles, meta, etype, ea.shape[1]) if ok else None
        self._fused_key = key
        self._fused_refs = (gptr, edge_attr)
        return self._fused

    def gcn_norm(self):
        """(dinv^2 [N], w_dst [E], w_src [E]) of PyG GCNConv's normalisation, computed once per batch."""
        if self._gcn is None:
            dinv, w_dst, w_src = ops.gcn_norm(self)
            self._gcn = (dinv * dinv, w_dst, w_src)
        return self._gcn

    def nn_index(self, edge_attr: torch.Tensor):
        """Per-batch index of the typed NNConv path (bond features are one-hot, src_1gp/dataset.py:82): with t(e) the bond
        type of edge e, rows r = src*De + t of the grouped projection Y [N*De, C] are what the messages read.  Returns
        (col_fwd [E] = r per dst-ordered edge, inv_deg [N] = 1/max(in-degree,1), rowptr_r [N*De+1], col_r [E] = destination
        of the edges grouped by r, w_r [E] = inv_deg of those destinations).  Outside CUDA-graph capture the rows are
        checked to be exact unit one-hot (the general case would need a per-edge [C,C] matrix)."""
        key = (edge_attr.data_ptr(), edge_attr._version, tuple(edge_attr.shape))
        if self._nn_key == key:
            return self._nn
        ea = self.sorted_edge_attr(edge_attr)
        De = ea.shape[1]
        if not _capturing():
            ok = bool(((ea.sum(1) == 1) & (ea.max(1).values == 1) & (ea.min(1).values == 0)).all()) if ea.numel() else True
            if not ok:
                raise ops._lib.GlamError("_NNConv kernel path needs exact one-hot edge_attr rows (bond types, "
                                         "src_1gp/dataset.py:82); general rows are not supported")
        t = ea.argmax(1).to(torch.int32)
        col_fwd = self.dst_src * De + t
        deg = (self.dst_rowptr[1:] - self.dst_rowptr[:-1]).clamp(min=1).to(torch.float32)
        inv_deg = 1.0 / deg
        fake = torch.stack([self.dst_dst.long(), col_fwd.long()])            # "edges" destination -> typed source row
        csr = ops.build_csr(fake, self.num_nodes * De)
        col_r = csr["dst_src"]
        w_r = inv_deg[col_r.long()]
        self._nn = (col_fwd, inv_deg, csr["dst_rowptr"], col_r, w_r, De)
        self._nn_key = key
        self._nn_ref = edge_attr
        return self._nn

    def sorted_edge_attr(self, edge_attr: torch.Tensor) -> torch.Tensor:
        """edge_attr rows permuted into destination order (done once per batch, reused by every step)."""
        key = (edge_attr.data_ptr(), edge_attr._version, tuple(edge_attr.shape))
        if self._edge_attr_key != key:
            ea = edge_attr if edge_attr.dim() == 2 else edge_attr.view(edge_attr.shape[0], -1)
            self._edge_attr_sorted = ops.gather_rows(ea, self.dst_perm)
            self._edge_attr_key = key
            self._edge_attr_ref = edge_attr
        return self._edge_attr_sorted


class FusedIndex:
    """tiles int32 [B,4] {n0,n1,e0,e1}, meta int32 [4] (count, violation flags), etype uint8 [E] (dst order)."""

    def __init__(self, tiles, meta, etype, edge_dim):
        self.tiles, self.meta, self.etype, self.edge_dim = tiles, meta, etype, int(edge_dim)


_fused_ok = {}                      # (N, E, B, De, device) -> verdict of the last host-side precondition check
_graph_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
_ptr_cache: "OrderedDict[tuple, tuple]" = OrderedDict()


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def graph_index(edge_index, num_nodes: int):
    if not torch.is_tensor(edge_index):                  # packed store: the index came prebuilt (glam_b200/packed.py)
        return edge_index
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(num_nodes), edge_index.device)
    hit = _graph_cache.get(key)
    if hit is not None:
        _graph_cache.move_to_end(key)
        return hit[1]
    gi = GraphIndex(edge_index, num_nodes)
    _graph_cache[key] = (edge_index, gi)
    while len(_graph_cache) > _CACHE_SIZE:
        _graph_cache.popitem(last=False)
    return gi


def graph_ptr(batch: torch.Tensor, num_graphs: Optional[int] = None):
    """(graph_ptr int32 [B+1], B).  `num_graphs=None` reads batch[-1] back to the host once per batch tensor
    (the reference does `batch.max().item()` on every readout call)."""
    if not torch.is_tensor(batch):                       # packed store: graph offsets came prebuilt
        return batch.index.gptr, batch.index.num_graphs
    key = (batch.data_ptr(), batch._version, tuple(batch.shape), num_graphs, batch.device)
    hit = _ptr_cache.get(key)
    if hit is not None:
        _ptr_cache.move_to_end(key)
        return hit[1], hit[2]
    if num_graphs is None:
        if _capturing():
            raise RuntimeError("num_graphs must be passed explicitly while a CUDA graph is being captured")
        num_graphs = int(batch[-1].item()) + 1 if batch.numel() > 0 else 0
    ptr = ops.graph_ptr(batch, num_graphs)
    _ptr_cache[key] = (batch, ptr, int(num_graphs))
    while len(_ptr_cache) > _CACHE_SIZE:
        _ptr_cache.popitem(last=False)
    return ptr, int(num_graphs)


def clear_caches() -> None:
    _graph_cache.clear()
    _ptr_cache.clear()
