"""Synthetic PyG-style batches shaped like the reference's data (no RDKit / datasets here).

Follows the edge-ordering convention of the reference featuriser (`src_1gp/dataset.py:75-87`): every
bond appears in both directions, the edges of a molecule are sorted by key ``src*n+dst`` and
`Batch.from_data_list` concatenates molecules with node-index offsets, so the batch-global edge
list is (src,dst)-lexicographic, node/edge ranges per graph are contiguous and `batch` is
non-decreasing.  Molecule sizes follow `src_1gp/demo/raw/demo.csv` (mean ~25 heavy atoms, long
tail); protein graphs follow `src_2gi_dti_scr/dataset.py:67-103` (main chain + symmetric contacts,
8 edge features).

Everything is vectorised numpy so 64k-graph batches build in well under a second on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class GraphBatch:
    """Plain-tensor stand-in for a `torch_geometric.data.Batch` (fields x, edge_index, edge_attr, batch, y)."""

    x: torch.Tensor            # [N, Din] fp32
    edge_index: torch.Tensor   # [2, E] int64, row 0 = source, row 1 = target
    edge_attr: torch.Tensor    # [E, De] fp32
    batch: torch.Tensor        # [N] int64, non-decreasing
    y: Optional[torch.Tensor] = None
    num_graphs: int = 0
    mask: Optional[torch.Tensor] = None     # [num_graphs] fp32, 1 = real graph, 0 = padding (pad_graph_batch); None = all real

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return GraphBatch(mv(self.x), mv(self.edge_index), mv(self.edge_attr), mv(self.batch), mv(self.y),
                          self.num_graphs, mv(self.mask))

    def pin_memory(self) -> "GraphBatch":
        pm = lambda t: None if t is None else t.pin_memory()
        return GraphBatch(pm(self.x), pm(self.edge_index), pm(self.edge_attr), pm(self.batch), pm(self.y),
                          self.num_graphs, pm(self.mask))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.x, self.edge_index, self.edge_attr, self.batch, self.y, self.mask) if t is not None)

    @property
    def num_nodes(self) -> int:
        return self.x.shape[0]

    @property
    def num_edges(self) -> int:
        return self.edge_index.shape[1]


def molecule_sizes(rng: np.random.Generator, num_graphs: int, mean_atoms: float = 25.0,
                   sigma: float = 0.35, lo: int = 4, hi: int = 128) -> np.ndarray:
    mu = np.log(mean_atoms) - 0.5 * sigma * sigma
    n = np.rint(rng.lognormal(mu, sigma, size=num_graphs)).astype(np.int64)
    return np.clip(n, lo, hi)


def _molecule_edges(rng: np.random.Generator, sizes: np.ndarray, n_rings: Optional[int] = None):
    """Undirected bonds: a chain-like random tree (each atom bonds to one of its 3 predecessors, so at
    most 3 children + 1 parent) plus ring closures (a, a+5|6) on ~8 % of atoms -> ~1.08 n bonds
    (or exactly `n_rings` closures)."""
    B = sizes.shape[0]
    offs = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    N = int(offs[-1])
    gid = np.repeat(np.arange(B, dtype=np.int64), sizes)
    local = np.arange(N, dtype=np.int64) - offs[gid]
    # tree edges
    child = np.nonzero(local > 0)[0]
    back = rng.choice(3, size=child.shape[0], p=[0.7, 0.2, 0.1])
    parent_local = np.maximum(local[child] - 1 - back, 0)
    parent = offs[gid[child]] + parent_local
    # ring closures
    span = 5 + (rng.random(N) < 0.5).astype(np.int64)
    fits = local + span < sizes[gid]
    if n_rings is None:
        ra = np.nonzero(fits & (rng.random(N) < 0.085))[0]
    else:                                   # exactly n_rings closures (fixed edge count, e.g. for CUDA-graph replay)
        cand = np.nonzero(fits)[0]
        assert 0 <= n_rings <= cand.shape[0], "total_edges not reachable for these molecule sizes"
        ra = np.sort(rng.choice(cand, size=n_rings, replace=False))
    rb = ra + span[ra]
    a = np.concatenate([child, ra])
    b = np.concatenate([parent, rb])
    return a, b, offs, gid, N


def _fit_total(rng: np.random.Generator, sizes: np.ndarray, total: int, lo: int = 4, hi: int = 128) -> np.ndarray:
    """Nudge random molecules by +-1 atom until the batch has exactly `total` atoms."""
    sizes = sizes.copy()
    diff = int(total - sizes.sum())
    while diff != 0:
        step = 1 if diff > 0 else -1
        ok = np.nonzero((sizes + step >= lo) & (sizes + step <= hi))[0]
        pick = rng.choice(ok, size=min(abs(diff), ok.shape[0]), replace=False)
        sizes[pick] += step
        diff = int(total - sizes.sum())
    return sizes


def tile_order(nodes, edges=None, cap_nodes: int = 128, cap_edges: int = 768, impl: str = "auto") -> np.ndarray:
    """Order of the graphs of one batch in which CONSECUTIVE graphs fill the fused kernels' tiles (whole graphs, <= cap_nodes
    rows and <= cap_edges in-edges per tile, include/glam_b200.h (8)): first-fit decreasing bin packing, the bins laid out one
    after the other.  The kernels' time per tile does not depend on how full it is, and a batch has no order of its own (the
    reference's DataLoader shuffles, src_1gp/trainer.py:95-101), so this is free throughput: 918 -> ~805 tiles for 4096
    MoleculeNet-shaped graphs.  Returns a permutation `perm` (new position -> old graph index); graphs over the caps keep a
    bin of their own.  `edges` (in-edges per graph) is optional: molecules never reach the edge cap before the node cap."""
    nodes = np.ascontiguousarray(np.asarray(nodes, dtype=np.int64))
    B = nodes.shape[0]
    if impl != "numpy":
        # the library's host routine (glam_tile_order: O(B), ~1 ms per 65 536 graphs; this numpy form takes 0.2 s there) —
        # the same permutation, tested; impl="auto" falls back to numpy only when the library cannot be loaded at all
        try:
            from . import _lib
            lib = _lib.load()
        except Exception:
            if impl == "native":
                raise
            lib = None
        if lib is not None:
            import ctypes
            perm = np.empty(B, dtype=np.int64)
            e64 = None if edges is None else np.ascontiguousarray(np.asarray(edges, dtype=np.int64))
            rc = lib.glam_tile_order(ctypes.c_void_p(nodes.ctypes.data), None if e64 is None else ctypes.c_void_p(e64.ctypes.data), B,
                                     int(cap_nodes), int(cap_edges), ctypes.c_void_p(perm.ctypes.data))
            if rc != 0:
                raise _lib.GlamError(f"glam_tile_order failed: {lib.glam_last_error().decode()}")
            return perm
    edges = np.zeros(B, dtype=np.int64) if edges is None else np.asarray(edges, dtype=np.int64)
    big = (nodes > cap_nodes) | (edges > cap_edges)          # over either cap: a tile of its own, at the end
    order_big = np.nonzero(big)[0]
    small = np.nonzero(~big)[0]
    # stacks of graph ids per size
    by_size = [[] for _ in range(cap_nodes + 1)]
    for gid in small[np.argsort(nodes[small], kind="stable")]:
        by_size[int(nodes[gid])].append(int(gid))
    have = np.array([len(v) for v in by_size], dtype=np.int64)
    have[0] = 0
    out = []
    for gid in by_size[0]:                                  # empty graphs take no rows: anywhere
        out.append(gid)
    remaining = int(have.sum())
    while remaining:
        room, eroom = cap_nodes, cap_edges
        while True:
            nz = np.nonzero(have[:room + 1])[0]
            if nz.shape[0] == 0:
                break
            placed = False
            for sz in nz[::-1]:                             # largest size that fits (and whose edges fit)
                gid = by_size[int(sz)][-1]
                if edges[gid] <= eroom:
                    by_size[int(sz)].pop(); have[sz] -= 1; remaining -= 1
                    out.append(gid); room -= int(sz); eroom -= int(edges[gid]); placed = True
                    break
            if not placed:
                break
    return np.concatenate([np.asarray(out, dtype=np.int64), order_big]) if B else np.zeros(0, dtype=np.int64)


def permute_graphs(b: "GraphBatch", perm) -> "GraphBatch":
    """The same batch with its graphs in the order `perm` (new position -> old graph index): nodes, edges, edge_attr, batch
    vector, y and mask follow.  Host tensors (collate time)."""
    perm = torch.as_tensor(np.asarray(perm), dtype=torch.int64)
    B = int(b.num_graphs)
    assert perm.shape[0] == B
    batch = b.batch
    counts = torch.bincount(batch, minlength=B)
    old_off = torch.zeros(B + 1, dtype=torch.int64)
    old_off[1:] = torch.cumsum(counts, 0)
    new_counts = counts[perm]
    new_off = torch.zeros(B + 1, dtype=torch.int64)
    new_off[1:] = torch.cumsum(new_counts, 0)
    inv = torch.empty(B, dtype=torch.int64)
    inv[perm] = torch.arange(B, dtype=torch.int64)
    # new index of every old node: position inside its graph + the graph's new offset
    node_new = torch.arange(batch.shape[0], dtype=torch.int64) - old_off[batch] + new_off[inv[batch]]
    order_nodes = torch.argsort(node_new)
    ei = node_new[b.edge_index]
    order_edges = torch.argsort(ei[0] * max(int(batch.shape[0]), 1) + ei[1], stable=True)
    kw = {}
    if getattr(b, "mask", None) is not None:
        kw["mask"] = b.mask[perm]
    y = b.y[perm] if (b.y is not None and b.y.shape[0] == B) else b.y
    return GraphBatch(b.x[order_nodes], ei[:, order_edges].contiguous(), b.edge_attr[order_edges], inv[batch][order_nodes], y, B, **kw)


def make_molecule_batch(num_graphs: int, node_dim: int = 9, edge_dim: int = 3, seed: int = 1234,
                        features: str = "chem", sizes: Optional[np.ndarray] = None,
                        targets: str = "regression", total_nodes: Optional[int] = None,
                        total_edges: Optional[int] = None, tile_pack: bool = False) -> GraphBatch:
    """One PyG-shaped batch of `num_graphs` synthetic molecules.

    features="chem": one-hot atom-type block + small non-negative integer columns (like
    `src_1gp/dataset.py:92-95`) and one-hot bond types; features="normal": N(0,1) nodes (layer-level numerics).
    """
    rng = np.random.default_rng(seed)
    if sizes is None:
        sizes = molecule_sizes(rng, num_graphs)
    sizes = np.asarray(sizes, dtype=np.int64)
    if total_nodes is not None:
        sizes = _fit_total(rng, sizes, total_nodes)
    if tile_pack:                                            # the same molecules, in the order that fills the kernels' tiles
        sizes = sizes[tile_order(sizes)]
    n_rings = None
    if total_edges is not None:
        assert total_edges % 2 == 0
        n_rings = total_edges // 2 - int(sizes.sum() - num_graphs)
    a, b, offs, gid, N = _molecule_edges(rng, sizes, n_rings)
    bond_type = rng.choice(edge_dim, size=a.shape[0], p=_bond_probs(edge_dim))
    src = np.concatenate([a, b])
    dst = np.concatenate([b, a])
    bt = np.concatenate([bond_type, bond_type])
    order = np.argsort(src * N + dst, kind="stable")
    src, dst, bt = src[order], dst[order], bt[order]
    edge_attr = np.zeros((src.shape[0], edge_dim), dtype=np.float32)
    edge_attr[np.arange(src.shape[0]), bt] = 1.0
    if features == "chem":
        x = np.zeros((N, node_dim), dtype=np.float32)
        n_onehot = max(node_dim - 3, 1)
        kind = rng.choice(n_onehot, size=N, p=_atom_probs(n_onehot))
        x[np.arange(N), kind] = 1.0
        for c in range(n_onehot, node_dim):
            x[:, c] = rng.integers(0, 4, size=N).astype(np.float32)
    else:
        x = rng.standard_normal((N, node_dim)).astype(np.float32)
    if targets == "regression":
        y = rng.standard_normal((num_graphs, 1)).astype(np.float32)
    else:
        y = (rng.random((num_graphs, 1)) < 0.5).astype(np.float32)
    return GraphBatch(torch.from_numpy(x), torch.from_numpy(np.stack([src, dst])), torch.from_numpy(edge_attr),
                      torch.from_numpy(gid), torch.from_numpy(y), int(num_graphs))


def _bond_probs(k: int) -> np.ndarray:
    p = np.array([0.62, 0.12, 0.02, 0.24, 0.05, 0.05, 0.05, 0.05][:k], dtype=np.float64)
    return p / p.sum()


def _atom_probs(k: int) -> np.ndarray:
    p = np.array([0.02, 0.68, 0.10, 0.12, 0.02, 0.02, 0.02, 0.01, 0.01] + [0.005] * 64, dtype=np.float64)[:k]
    return p / p.sum()


def make_protein_batch(num_graphs: int, node_dim: int = 49, edge_dim: int = 8, seed: int = 1234,
                       min_len: int = 300, max_len: int = 700, contacts_per_residue: float = 3.3,
                       same_protein: bool = False) -> GraphBatch:
    """Protein contact-map graphs as built by `src_2gi_dti_scr/dataset.py:67-103`: main-chain pairs
    (i,i+1),(i+1,i) first, then the symmetric contact list in row-major order (duplicates of main-chain
    pairs possible, as in the reference).  edge_attr = [is_chain, p, 1-p, l1..l5]."""
    rng = np.random.default_rng(seed)
    if same_protein:
        L = np.full(num_graphs, int(rng.integers(min_len, max_len + 1)), dtype=np.int64)
    else:
        L = rng.integers(min_len, max_len + 1, size=num_graphs).astype(np.int64)
    xs, srcs, dsts, eas, gids = [], [], [], [], []
    off = 0
    for g in range(num_graphs):
        if same_protein and g > 0:
            # LIT-PCBA style: every pair shares one protein (src_2gi_dti_scr/dataset.py:297,301)
            xs.append(xs[0]); eas.append(eas[0])
            srcs.append(srcs[0] + off); dsts.append(dsts[0] + off)
            gids.append(np.full(L[g], g, dtype=np.int64))
            off += int(L[g])
            continue
        n = int(L[g])
        x = np.zeros((n, node_dim), dtype=np.float32)
        aa = rng.integers(0, min(20, node_dim), size=n)
        x[np.arange(n), aa] = 1.0
        if node_dim > 20:
            nb = min(5, node_dim - 20)
            x[:, 20:20 + nb] = (rng.random((n, nb)) < 0.3)
            if node_dim > 25:
                x[:, 25:] = rng.random((n, node_dim - 25)).astype(np.float32)
        i = np.arange(n - 1)
        ch_s = np.stack([i, i + 1], 1).reshape(-1)
        ch_d = np.stack([i + 1, i], 1).reshape(-1)
        ch_ea = np.zeros((ch_s.shape[0], edge_dim), dtype=np.float32)
        ch_ea[:, 0] = 1.0
        m = int(contacts_per_residue * n / 2)
        ca = rng.integers(0, n, size=m)
        sep = np.maximum(1, np.rint(np.abs(rng.standard_cauchy(m)) * 6 + 1)).astype(np.int64)
        cb = ca + sep
        ok = cb < n
        ca, cb = ca[ok], cb[ok]
        key = np.unique(ca * n + cb)
        ca, cb = key // n, key % n
        p = rng.uniform(0.1, 1.0, size=ca.shape[0]).astype(np.float32)
        cs = np.concatenate([ca, cb]); cd = np.concatenate([cb, ca]); pp = np.concatenate([p, p])
        o = np.argsort(cs * n + cd, kind="stable")
        cs, cd, pp = cs[o], cd[o], pp[o]
        c_ea = np.zeros((cs.shape[0], edge_dim), dtype=np.float32)
        if edge_dim >= 3:
            c_ea[:, 1] = pp; c_ea[:, 2] = 1 - pp
        for k, (lo, hi) in enumerate([(0.1, 0.3), (0.3, 0.5), (0.5, 0.7), (0.5, 0.9), (0.9, 1.01)]):
            if 3 + k < edge_dim:
                c_ea[:, 3 + k] = ((pp >= lo) & (pp < hi))
        xs.append(x)
        srcs.append(np.concatenate([ch_s, cs]) + off)
        dsts.append(np.concatenate([ch_d, cd]) + off)
        eas.append(np.concatenate([ch_ea, c_ea]))
        gids.append(np.full(n, g, dtype=np.int64))
        off += n
    x = np.concatenate(xs); src = np.concatenate(srcs); dst = np.concatenate(dsts)
    return GraphBatch(torch.from_numpy(x), torch.from_numpy(np.stack([src, dst]).astype(np.int64)),
                      torch.from_numpy(np.concatenate(eas)), torch.from_numpy(np.concatenate(gids)),
                      None, int(num_graphs))


def shard_by_graph(batch: GraphBatch, rank: int, world: int) -> GraphBatch:
    """Rank `rank`'s contiguous slice of graphs (partition by graph, no edge crosses graphs)."""
    B = batch.num_graphs
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    nmask = (batch.batch >= lo) & (batch.batch < hi)
    nodes = torch.nonzero(nmask).flatten()
    if nodes.numel() == 0:
        n0, n1 = 0, 0
    else:
        n0, n1 = int(nodes[0]), int(nodes[-1]) + 1
    emask = (batch.edge_index[1] >= n0) & (batch.edge_index[1] < n1)
    return GraphBatch(batch.x[n0:n1], batch.edge_index[:, emask] - n0, batch.edge_attr[emask],
                      batch.batch[n0:n1] - lo, None if batch.y is None else batch.y[lo:hi], hi - lo)


def pad_graph_batch(b: GraphBatch, nodes: int, edges: int, graphs: int, max_nodes: int = 128, max_edges: int = 768,
                    with_mask: bool = False) -> GraphBatch:
    """`b` padded to exactly (`nodes`, `edges`, `graphs`) with dummy graphs appended BEHIND the real ones, so that batches of
    varying size can be replayed through ONE captured step (engine.ScreenStep pads with this): the padding nodes / edges are
    spread evenly over the `graphs - b.num_graphs` dummy graphs (zero features, bond type 0, ring edges i -> i+1; each at most
    `max_nodes` nodes / `max_edges` edges — the fused kernels' per-graph caps).  The scores of the real graphs are unchanged
    (graphs are independent); rows `b.num_graphs:` of the output belong to the dummies.  `mask` [graphs] marks the real graphs
    (always attached when padding happens; `with_mask` attaches an all-ones mask to a batch that already fits) — what a masked
    training loss weights by (engine.masked_loss)."""
    N, E, B = b.num_nodes, b.num_edges, b.num_graphs
    pn, pe, d = nodes - N, edges - E, graphs - B
    if min(pn, pe, d) < 0:
        raise ValueError(f"batch ({N} nodes, {E} edges, {B} graphs) exceeds the captured capacity ({nodes}, {edges}, {graphs})")
    if pn == 0 and pe == 0 and d == 0:
        if with_mask and b.mask is None:
            return GraphBatch(b.x, b.edge_index, b.edge_attr, b.batch, b.y, B, torch.ones(B, dtype=torch.float32, device=b.x.device))
        return b
    if d == 0 or pn < d or pn > max_nodes * d or pe > max_edges * d:
        raise ValueError(f"cannot pad ({N}, {E}, {B}) to ({nodes}, {edges}, {graphs}): {pn} nodes / {pe} edges do not fit {d} dummy graphs "
                         f"of 1..{max_nodes} nodes and <= {max_edges} edges — capture with more spare graphs")
    k = torch.arange(d)
    n_g = pn // d + (k < pn % d).long()                              # nodes per dummy graph (>= 1)
    e_g = pe // d + (k < pe % d).long()
    n_off = torch.cumsum(n_g, 0) - n_g + N
    gid = torch.repeat_interleave(k, e_g)                            # dummy graph of every padding edge
    j = torch.arange(pe) - torch.repeat_interleave(torch.cumsum(e_g, 0) - e_g, e_g)   # running index inside its graph
    src = n_off[gid] + j % n_g[gid]
    dst = n_off[gid] + (j + 1) % n_g[gid]
    dev = b.x.device
    ea = torch.zeros((pe, b.edge_attr.shape[1]), dtype=b.edge_attr.dtype)
    ea[:, 0] = 1.0
    y = None if b.y is None else torch.cat([b.y, torch.zeros((d,) + tuple(b.y.shape[1:]), dtype=b.y.dtype, device=b.y.device)])
    return GraphBatch(torch.cat([b.x, torch.zeros((pn, b.x.shape[1]), dtype=b.x.dtype, device=dev)]),
                      torch.cat([b.edge_index, torch.stack([src, dst]).to(dev)], dim=1),
                      torch.cat([b.edge_attr, ea.to(dev)]),
                      torch.cat([b.batch, (B + torch.repeat_interleave(k, n_g)).to(dev)]), y, graphs,
                      torch.cat([torch.ones(B), torch.zeros(d)]).to(dev))
