"""Drop-in mirror of the reference's `layer.py` namespace (src_1gp/layer.py, identical copies in
src_2gi_ddi/ and src_2gi_dti_scr/) for the message-passing hot path.

Same class names, constructor signatures, forward signatures, parameter names / shapes /
registration order / init, so `state_dict`s interchange with the reference and the reference's
`model.py` can build its blocks from the same name strings (`mol_block='_TripletMessage'`,
`mol_readout='Set2Set'`, `graph_act='RReLU'`, `graph_do='Dropout(0.2)'` ...).  What is different is
underneath: every node-/edge-sized computation of TripletMessage, TripletMessageLight, the GRU update,
GlobalLAPool / Set2Set and dot_and_global_pool2 runs in hand-written sm_100a kernels behind the C ABI
(include/glam_b200.h).  There is no CPU path: CPU tensors raise.

Out of the hot path and therefore plain torch here: the norm wrappers, dropout,
activations and the wide graph-level `LinearBlock`s.  The reference's defaults `_NNConv`, `_GCNConv` and
`GlobalPool5` run on the library too; `_GATConv` is a third-party PyG layer outside the path
and raises at construction.
"""
from __future__ import annotations

import re
from typing import Optional

import torch
from torch import nn
from torch.nn import Parameter
from torch.nn.init import kaiming_uniform_, zeros_

from . import functional as Fn
from . import graph as G
from . import ops


# The fused message kernel (csrc/mp_fused.cu) is used whenever its preconditions hold; False keeps the per-op kernels
# (A/B timing, and the path tests compare it against).
USE_FUSED_STACK = True


def _fused_index(g, x, edge_attr, batch, num_graphs, heads, channels):
    """FusedIndex of this batch for the fused message kernel, or None (unsupported shape / math mode, no `batch`, batch
    violating the kernel's preconditions)."""
    if not (USE_FUSED_STACK and batch is not None and x.is_cuda and edge_attr.dim() == 2):
        return None
    if not ops.message_stack_supported(channels, heads, edge_attr.shape[1]):
        return None
    if num_graphs is None and G._capturing():
        return None
    gptr, B = G.graph_ptr(batch, num_graphs)
    return g.fused_index(gptr, B, edge_attr)


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _ldxp(heads: int, channels: int) -> int:
    """Row pitch of the extended projection xp | s_i | s_j, padded to 16 bytes."""
    return (heads * channels + 2 * heads + 3) // 4 * 4


# --------------------------------------------------------------------------------------------------
# message layers
# --------------------------------------------------------------------------------------------------
class TripletMessage(nn.Module):
    """Multi-head attention over (target node, edge, source node) triplets — src_1gp/layer.py:15-64.

    forward(x [N,C], edge_index int64 [2,E] (row 0 source, row 1 target), edge_attr [E,De]) -> [N,C]
    """

    def __init__(self, node_channels, edge_channels, heads=3, negative_slope=0.2, **kwargs):
        super().__init__()
        self.node_channels = node_channels
        self.edge_channels = edge_channels
        self.heads = heads
        self.negative_slope = negative_slope
        self.weight_node = Parameter(torch.empty(node_channels, heads * node_channels))
        self.weight_edge = Parameter(torch.empty(edge_channels, heads * node_channels))
        self.weight_triplet_att = Parameter(torch.empty(1, heads, 3 * node_channels))
        self.weight_scale = Parameter(torch.empty(heads * node_channels, node_channels))
        self.bias = Parameter(torch.empty(node_channels))
        self.reset_parameters()

    def reset_parameters(self):
        for w in (self.weight_node, self.weight_edge, self.weight_triplet_att, self.weight_scale):
            kaiming_uniform_(w)
        zeros_(self.bias)

    def derived(self):
        """(w_ext, att_edge) for the current parameters: one tiny kernel, differentiable."""
        C, H = self.node_channels, self.heads
        return Fn.TripletPrepFn.apply(self.weight_node, self.weight_edge, self.weight_triplet_att.view(H, 3 * C), C, H,
                                      self.edge_channels, _ldxp(H, C))

    def forward(self, x, edge_index, edge_attr, size=None, batch=None, num_graphs=None):
        """`batch` (optional, not in the reference signature): the PyG batch vector; with it the layer runs as one fused
        kernel on graph-aligned tiles (projection -> edge phase -> scale projection, csrc/mp_fused.cu)."""
        g = G.graph_index(edge_index, x.shape[0])
        ea = g.sorted_edge_attr(edge_attr)
        w_ext, att_edge = self.derived()
        fi = _fused_index(g, x, edge_attr, batch, num_graphs, self.heads, self.node_channels)
        if fi is not None and not _wants_grad(x, *self.parameters()):
            out, _ = ops.message_stack_fwd(x.contiguous(), None, w_ext, self.weight_edge, att_edge, self.weight_scale, self.bias,
                                           None, None, None, None, g, fi, self.heads, self.node_channels, 1,
                                           self.negative_slope, ops.ACT_NONE, 0.0, False, conv_only=True)
            return out[0]
        return Fn.TripletConvFn.apply(x, w_ext, self.weight_edge, att_edge, self.weight_scale, self.bias, ea, g, self.heads,
                                      self.node_channels, self.negative_slope, fi)

    def extra_repr(self):
        return f"{self.node_channels}, {self.node_channels}, heads={self.heads}"


class TripletMessageLight(nn.Module):
    """Single-head variant without edge/output projections — src_1gp/layer.py:67-104."""

    def __init__(self, node_channels, edge_channels, negative_slope=0.2, **kwargs):
        super().__init__()
        self.node_channels = node_channels
        self.edge_channels = edge_channels
        self.negative_slope = negative_slope
        self.weight_node = Parameter(torch.empty(node_channels, node_channels))
        self.weight_triplet_att = Parameter(torch.empty(1, 2 * node_channels + edge_channels))
        self.bias = Parameter(torch.empty(node_channels))
        self.reset_parameters()

    def reset_parameters(self):
        kaiming_uniform_(self.weight_node)
        kaiming_uniform_(self.weight_triplet_att)
        zeros_(self.bias)

    def derived(self):
        C = self.node_channels
        return Fn.TripletPrepFn.apply(self.weight_node, None, self.weight_triplet_att.view(-1), C, 1, self.edge_channels,
                                      _ldxp(1, C))

    def forward(self, x, edge_index, edge_attr, size=None):
        g = G.graph_index(edge_index, x.shape[0])
        ea = g.sorted_edge_attr(edge_attr)
        w_ext, att_edge = self.derived()
        agg = Fn.TripletConvFn.apply(x, w_ext, None, att_edge, None, None, ea, g, 1, self.node_channels, self.negative_slope, None)
        return agg + self.bias

    def extra_repr(self):
        return f"{self.node_channels}, {self.node_channels}"


class _None(nn.Module):
    """Placeholder for "no norm / dropout / activation" (src_1gp/layer.py:107-112)."""

    def __init__(self, **params):
        super().__init__()

    def forward(self, x, batch=None, num_graphs=None):
        return x


class _TripletMessage(nn.Module):
    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        self.conv = TripletMessage(in_dim, edge_in_dim)      # out_dim ignored, as in the reference (:128)

    def forward(self, x, edge_index, edge_attr):
        return self.conv(x, edge_index, edge_attr)


class _TripletMessageLight(nn.Module):
    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        self.conv = TripletMessageLight(in_dim, edge_in_dim)

    def forward(self, x, edge_index, edge_attr):
        return self.conv(x, edge_index, edge_attr)


def _third_party(name):
    class _Missing(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(
                f"{name} wraps a torch_geometric layer that is outside the triplet hot path (SURVEY.md §2.1); "
                "use '_TripletMessage' or '_TripletMessageLight'")
    _Missing.__name__ = name
    return _Missing


class GCNConv(nn.Module):
    """PyG GCNConv(in_channels, out_channels) @1.7.2 defaults (self loops, symmetric normalisation, bias): parameters
    `weight [in,out]` (glorot) and `bias [out]` (zeros), as `_GCNConv` builds it (src_1gp/layer.py:143-149)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        self.bias = Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = (6.0 / (self.in_channels + self.out_channels)) ** 0.5
        nn.init.uniform_(self.weight, -bound, bound)
        zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        if edge_weight is not None:
            raise NotImplementedError("GCNConv: edge weights are not used by the reference (layer.py:149 passes none)")
        g = G.graph_index(edge_index, x.shape[0])
        return Fn.GCNConvFn.apply(x, self.weight, self.bias, g, g.gcn_norm())


class _GCNConv(nn.Module):
    """src_1gp/layer.py:143-149 (default protein block of the drug-target model, src_2gi_dti_scr/run.py:19)."""

    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        self.conv = GCNConv(in_dim, out_dim)

    def forward(self, x, edge_index, edge_attr):
        return self.conv(x, edge_index)


class NNConv(nn.Module):
    """PyG NNConv(in_channels, out_channels, nn, aggr='mean') @1.7.2 as `_NNConv` builds it (src_1gp/layer.py:115-122):
    parameters `nn.{0,2}.{weight,bias}` (torch Linear init), `root [in,out]` (uniform(1/sqrt(in))), `bias [out]` (zeros).
    Bond features are one-hot (src_1gp/dataset.py:82), so nn(edge_attr) has edge_dim distinct values: they are evaluated
    in parameter space (tiny torch ops, differentiable) and the node-sized work runs in the library (Fn.NNConvFn)."""

    def __init__(self, in_channels, out_channels, nn_module, aggr="mean"):
        super().__init__()
        if aggr != "mean":
            raise NotImplementedError("NNConv: the reference only builds aggr='mean' (layer.py:119)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.nn = nn_module
        self.root = Parameter(torch.empty(in_channels, out_channels))
        self.bias = Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.nn.modules():                          # PyG: reset(self.nn) first, then root and bias (same RNG order)
            if m is not self.nn and hasattr(m, "reset_parameters"):
                m.reset_parameters()
        bound = 1.0 / (self.in_channels ** 0.5)
        nn.init.uniform_(self.root, -bound, bound)
        zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr):
        g = G.graph_index(edge_index, x.shape[0])
        idx = g.nn_index(edge_attr)
        De = idx[5]
        eye = torch.eye(De, dtype=x.dtype, device=x.device)
        theta = self.nn(eye).view(De, self.in_channels, self.out_channels)       # Theta_t = nn(e_t)
        theta_cat = theta.permute(1, 0, 2).reshape(self.in_channels, De * self.out_channels)
        return Fn.NNConvFn.apply(x, theta_cat, self.root, self.bias, g, idx)


class _NNConv(nn.Module):
    """src_1gp/layer.py:115-122 (the reference's default mol_block, src_1gp/run.py:21)."""

    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        net = nn.Sequential(nn.Linear(edge_in_dim, 32), nn.ReLU(), nn.Linear(32, in_dim * out_dim))
        self.conv = NNConv(in_dim, out_dim, net, aggr="mean")

    def forward(self, x, edge_index, edge_attr):
        return self.conv(x, edge_index, edge_attr)


_GATConv = _third_party("_GATConv")


# --------------------------------------------------------------------------------------------------
# norm wrappers (torch; PyG-1.7.2 semantics) — not on the hot path
# --------------------------------------------------------------------------------------------------
def _seg_mean(x, batch, B):
    out = x.new_zeros((B,) + tuple(x.shape[1:])).index_add_(0, batch, x)
    cnt = x.new_zeros(B).index_add_(0, batch, x.new_ones(x.shape[0])).clamp_(min=1)
    return out / cnt.view(-1, *([1] * (x.dim() - 1)))


class _BatchNorm(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.norm = nn.Module()
        self.norm.module = nn.BatchNorm1d(in_channels)       # PyG BatchNorm keeps it under `.module`

    def forward(self, x, batch=None, num_graphs=None):
        return self.norm.module(x)


class _LayerNorm(nn.Module):
    """PyG LayerNorm: statistics over all nodes AND channels of each graph."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.norm = nn.Module()
        self.norm.weight = Parameter(torch.ones(in_channels))
        self.norm.bias = Parameter(torch.zeros(in_channels))
        self.eps = eps

    def forward(self, x, batch=None, num_graphs=None):
        if batch is None:
            x = x - x.mean()
            out = x / (x.std(unbiased=False) + self.eps)
        else:
            B = int(batch[-1]) + 1 if num_graphs is None else num_graphs
            mean = _seg_mean(x, batch, B).mean(dim=-1, keepdim=True)
            x = x - mean[batch]
            var = _seg_mean(x * x, batch, B).mean(dim=-1, keepdim=True)
            out = x / (var + self.eps).sqrt()[batch]
        return out * self.norm.weight + self.norm.bias


class _PairNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.eps = eps

    def forward(self, x, batch=None, num_graphs=None):
        if batch is None:
            x = x - x.mean(dim=0, keepdim=True)
            return x / (self.eps + x.pow(2).sum(-1).mean()).sqrt()
        if x.is_cuda and x.dim() == 2 and x.shape[1] <= 128:
            # one warp per graph, fixed-order reductions (csrc/norm.cu): reproducible run to run, no index_add_ atomics
            gptr, B = G.graph_ptr(batch, num_graphs)
            return Fn.PairNormFn.apply(x, gptr, B, self.eps)
        return self.composed_forward(x, batch, num_graphs)

    def composed_forward(self, x, batch, num_graphs=None):
        """The same in torch ops (CPU tensors; kept as an independent cross-check of the kernel)."""
        B = int(batch[-1]) + 1 if num_graphs is None else num_graphs
        x = x - _seg_mean(x, batch, B)[batch]
        return x / torch.sqrt(self.eps + _seg_mean(x.pow(2).sum(-1, keepdim=True), batch, B)[batch])


class _GraphSizeNorm(nn.Module):
    def __init__(self, in_channels):
        super().__init__()

    def forward(self, x, batch=None, num_graphs=None):
        return x / (x.shape[0] ** 0.5)                       # called without batch in the reference (:194)


# --------------------------------------------------------------------------------------------------
# readouts
# --------------------------------------------------------------------------------------------------
class GlobalAttention(nn.Module):
    """PyG GlobalAttention(gate_nn=Linear(C,1), nn=Linear(C,2C)) @1.7.2.

    Uses linearity of `nn`: sum_n a_n (W x_n + b) = W (sum_n a_n x_n) + b sum_n a_n, so the [N,2C] projection is
    never formed; gate, softmax and pooling are one kernel."""

    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn

    def forward(self, x, batch, size=None, num_graphs=None):
        if not (isinstance(self.gate_nn, nn.Linear) and self.gate_nn.out_features == 1
                and (self.nn is None or isinstance(self.nn, nn.Linear))):
            raise NotImplementedError("GlobalAttention kernel path needs gate_nn=Linear(C,1) and nn=Linear|None")
        gptr, B = G.graph_ptr(batch, size if size is not None else num_graphs)
        pooled, asum = Fn.SegAttnPoolFn.apply(x, self.gate_nn.weight, self.gate_nn.bias, gptr, B)
        if self.nn is None:
            return pooled
        return Fn.LinearFn.apply(pooled, self.nn.weight, None) + asum.unsqueeze(1) * self.nn.bias


class GlobalLAPool(nn.Module):
    """Global linear-attention pool, [N,C] -> [B,2C] — src_1gp/layer.py:206-220."""

    def __init__(self, in_channels, **params):
        super().__init__()
        self.pool = GlobalAttention(gate_nn=nn.Linear(in_channels, 1), nn=nn.Linear(in_channels, 2 * in_channels))

    def forward(self, x, batch, num_graphs=None):
        return self.pool(x, batch, num_graphs=num_graphs)


class Set2Set(nn.Module):
    """PyG Set2Set(in_channels, processing_steps, num_layers=1) @1.7.2 (imported by src_1gp/model.py:2,
    constructed at :41).  The LSTM parameters live in a torch.nn.LSTM so names/init match (`lstm.weight_ih_l0` ...);
    its cell math runs through the library's GEMM + gate kernels, the attention + pooling through one kernel."""

    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("Set2Set: only num_layers=1 (the reference never passes another value)")
        self.in_channels = in_channels
        self.out_channels = 2 * in_channels
        self.processing_steps = processing_steps
        self.num_layers = num_layers
        self.lstm = nn.LSTM(self.out_channels, in_channels, num_layers)

    fused = True     # one kernel per direction for all rounds; False = the composed GEMM + gates + pooling path

    def forward(self, x, batch, num_graphs=None):
        gptr, B = G.graph_ptr(batch, num_graphs)
        C, l = self.in_channels, self.lstm
        if self.fused and C <= 128 and self.processing_steps <= 8:
            return Fn.Set2SetFn.apply(x, l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0, gptr, B,
                                      self.processing_steps)
        h = x.new_zeros((B, C))
        c = x.new_zeros((B, C))
        q_star = x.new_zeros((B, 2 * C))
        for _ in range(self.processing_steps):
            gates = Fn.LinearFn.apply(q_star, l.weight_ih_l0, l.bias_ih_l0) + Fn.LinearFn.apply(h, l.weight_hh_l0, l.bias_hh_l0)
            h, c = Fn.LSTMGatesFn.apply(gates, c)
            r, _ = Fn.SegAttnPoolFn.apply(x, h, None, gptr, B)
            q_star = torch.cat([h, r], dim=-1)
        return q_star


class GlobalPool5(nn.Module):
    """mean | sum | sort-pool(k=3) readout (src_1gp/layer.py:197-203; the reference's default mol_readout): one warp per
    graph (csrc/pool5.cu).  `composed_forward` is the same thing in torch ops, kept as an independent cross-check."""

    def __init__(self, **params):
        super().__init__()

    def forward(self, x, batch, num_graphs=None):
        gptr, B = G.graph_ptr(batch, num_graphs)
        return Fn.Pool5Fn.apply(x, gptr, B)

    def composed_forward(self, x, batch, num_graphs=None):
        B = int(batch[-1]) + 1 if num_graphs is None else num_graphs
        C = x.shape[1]
        total = x.new_zeros((B, C)).index_add_(0, batch, x)
        cnt = torch.bincount(batch, minlength=B)
        mean = total / cnt.clamp(min=1).view(-1, 1).to(x.dtype)
        # global_sort_pool: per graph, nodes sorted by last channel (descending), first 3 kept, zero padded
        key = x[:, -1]
        order = torch.argsort(key, descending=True, stable=True)
        order = order[torch.argsort(batch[order], stable=True)]
        start = torch.cumsum(cnt, 0) - cnt
        rank = torch.arange(x.shape[0], device=x.device) - start[batch[order]]
        keep = rank < 3
        top = x.new_zeros((B, 3, C))
        top[batch[order][keep], rank[keep]] = x[order][keep]
        return torch.cat([mean, total, top.view(B, 3 * C)], dim=-1)


# --------------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------------
_NORMS = {"_None": _None, "_BatchNorm": _BatchNorm, "_LayerNorm": _LayerNorm, "_PairNorm": _PairNorm,
          "_GraphSizeNorm": _GraphSizeNorm}
_CONVS = {"_TripletMessage": _TripletMessage, "_TripletMessageLight": _TripletMessageLight, "_NNConv": _NNConv,
          "_GCNConv": _GCNConv, "_GATConv": _GATConv}


def _build_dropout(spec) -> nn.Module:
    """The reference passes dropouts as constructor strings: '_None()', 'Dropout(0.2)' (run.py:31-34)."""
    if isinstance(spec, nn.Module):
        return spec
    m = re.fullmatch(r"\s*(_None|Dropout)\s*\(\s*([0-9.eE+-]*)\s*\)\s*", spec)
    if m is None:
        raise ValueError(f"unknown dropout spec {spec!r}")
    return _None() if m.group(1) == "_None" else nn.Dropout(float(m.group(2) or 0.5))


def _build_act(name) -> nn.Module:
    """Activations are passed by class name without parentheses: '_None', 'ReLU', 'RReLU', 'CELU' ... (run.py:35-37)."""
    if isinstance(name, nn.Module):
        return name
    if name == "_None":
        return _None()
    cls = getattr(nn, name.replace("nn.", "").rstrip("()"), None)
    if cls is None or not issubclass(cls, nn.Module):
        raise ValueError(f"unknown activation {name!r}")
    return cls()


def _fusable_act(act: nn.Module, training: bool):
    """(code, param) if the activation can run inside the GRU-update kernel, else None."""
    if isinstance(act, _None):
        return ops.ACT_NONE, 0.0
    if type(act) is nn.ReLU:
        return ops.ACT_RELU, 0.0
    if type(act) is nn.LeakyReLU:
        return ops.ACT_LEAKY, float(act.negative_slope)
    if type(act) is nn.CELU and act.alpha == 1.0:
        return ops.ACT_CELU, 1.0
    if type(act) is nn.RReLU and not training:                # eval-mode RReLU is leaky_relu((lower+upper)/2)
        return ops.ACT_LEAKY, float((act.lower + act.upper) / 2)
    return None


class LinearBlock(nn.Module):
    """norm -> dropout -> Linear -> act (src_1gp/layer.py:223-237)."""

    def __init__(self, in_dim=32, out_dim=64, norm="_None", dropout="_None()", act="ReLU"):
        super().__init__()
        self.norm = _NORMS[norm](in_channels=in_dim)
        self.dropout = _build_dropout(dropout)
        self.linear = nn.Linear(in_dim, out_dim)
        self.act = _build_act(act)

    def fusable_into_stack(self, x):
        """(weight, bias, act code, act param) when this block can be applied inside the fused message kernel while it loads
        its input rows (no norm, dropout inactive, kernel-side activation, <= 16 raw features), else None."""
        act = _fusable_act(self.act, self.training)
        drop_off = isinstance(self.dropout, _None) or not self.training or getattr(self.dropout, "p", 1.0) == 0.0
        if act is None or not drop_off or not isinstance(self.norm, _None) or x.dim() != 2 or x.shape[1] > 16 or x.dtype != torch.float32:
            return None
        return self.linear.weight, self.linear.bias, act[0], act[1]

    def forward(self, x, batch=None):
        z = self.dropout(self.norm(x, batch))
        lin = self.linear
        # node-level projections (many rows, narrow) go through the library's GEMMs: the weight gradient is a long
        # fixed-order A^T B there; wide graph-level layers ([B, e_dim]) are plain library GEMMs (cuBLAS via torch)
        if z.is_cuda and z.dim() == 2 and lin.out_features <= 256 and lin.in_features <= 288:
            return self.act(Fn.LinearFn.apply(z, lin.weight, lin.bias))
        return self.act(lin(z))


class MessageBlock(nn.Module):
    """One message-passing step: norm -> dropout -> conv -> CELU -> GRU -> (+identity) -> act
    (src_1gp/layer.py:240-267).  forward(x, edge_index, edge_attr, h=None, batch=None) -> (x, h)."""

    def __init__(self, in_dim=32, out_dim=64, in_edge_dim=13, norm="_None", dropout="Dropout(0.2)", conv="_NNConv",
                 act="ReLU", res=True):
        super().__init__()
        self.norm = _NORMS[norm](in_channels=in_dim)
        self.dropout = _build_dropout(dropout)
        self.conv = _CONVS[conv](in_dim, out_dim, in_edge_dim)
        self.gru = nn.GRU(in_dim, out_dim)
        if conv in ("_GCNConv", "_GATConv"):
            self.gru = None
        self.act = _build_act(act)
        self.res = res

    def run_steps(self, x, edge_index, edge_attr, steps, batch=None, num_graphs=None, keep="all", pre=None):
        """`steps` applications of this block starting from h=None (the loop of src_1gp/model.py:60-62), returning
        ([x_1 .. x_steps], h) — or ([x_steps], h) with keep="last".  With the triplet layer, no norm and a fusable
        activation the whole loop is one autograd node (functional.MessageStackFn) and, when the batch meets the
        preconditions of csrc/mp_fused.cu, ONE kernel launch; otherwise it is the plain loop over forward().
        `pre`: the LinearBlock that produces this block's input (src_1gp/model.py:49) — `x` is then ITS input (the raw
        features); in evaluation it is applied inside the fused kernel, otherwise simply called first."""
        inner = getattr(self.conv, "conv", None)
        pre_fused = pre_train = pre_block = None
        if pre is not None:
            can = pre.fusable_into_stack(x) if (x.is_cuda and not _wants_grad(x)) else None
            if can is not None and _wants_grad(*self.parameters(), *pre.parameters()):
                pre_train, pre_block = can, pre                  # decided below, once the batch's tile index is known
            else:
                pre_fused = can
            if can is None:
                x, pre = pre(x, batch=batch), None
        width = pre.linear.out_features if pre is not None else x.shape[1]
        fused = _fusable_act(self.act, self.training)
        drop = self.dropout
        pairnorm = isinstance(self.norm, _PairNorm) and batch is not None and x.shape[1] <= 128
        stackable = (self.gru is not None and isinstance(inner, TripletMessage) and (isinstance(self.norm, _None) or pairnorm)
                     and fused is not None and isinstance(drop, (_None, nn.Dropout)) and x.is_cuda and steps >= 1
                     and width == inner.node_channels and self.gru.hidden_size == width)
        if not stackable:
            if pre is not None:
                x, pre = pre(x, batch=batch), None
            xs, h = [], None
            for _ in range(steps):
                x, h = self.forward(x, edge_index, edge_attr, h=h, batch=batch, num_graphs=num_graphs)
                xs.append(x)
            return (xs if keep == "all" else xs[-1:]), h
        p_drop = float(drop.p) if (isinstance(drop, nn.Dropout) and self.training) else 0.0
        g = G.graph_index(edge_index, x.shape[0])
        w_ext, att_edge = inner.derived()
        gru = self.gru
        pn = None
        if pairnorm:
            gptr_pn, B_pn = G.graph_ptr(batch, num_graphs)
            pn = (gptr_pn, B_pn, float(self.norm.eps))
        no_grad = not _wants_grad(x, *self.parameters())
        # PairNorm runs inside the fused kernel in evaluation (tile-local statistics: tiles are whole graphs); training with
        # PairNorm keeps the per-op stacked node (deterministic norm kernels)
        pn_in_kernel = pn is not None and no_grad and torch.is_tensor(batch) and batch.dtype == torch.int64 and batch.is_contiguous()
        fi = (_fused_index(g, x, edge_attr, batch, num_graphs, inner.heads, inner.node_channels)
              if (p_drop == 0.0 and (pn is None or pn_in_kernel)) else None)
        if fi is not None and no_grad:
            # screening / evaluation: nothing is kept for backward, only the outputs leave the SM
            x_out, h_out = ops.message_stack_fwd(
                x.contiguous(), None, w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias,
                gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, g, fi, inner.heads, inner.node_channels,
                int(steps), inner.negative_slope, fused[0], fused[1], bool(self.res), keep_all=(keep == "all"), pre=pre_fused,
                pn=(batch, float(self.norm.eps)) if pn_in_kernel else None)
            return list(x_out.unbind(0)), h_out.unsqueeze(0)
        if pn is not None:
            fi = None                                            # (the training-mode fused pair has no PairNorm)
        if pre_fused is not None:
            x = pre(x, batch=batch)
        ea2 = edge_attr if edge_attr.dim() == 2 else edge_attr.view(edge_attr.shape[0], -1)
        both = (fi is not None and Fn.USE_FUSED_BWD and g.src_rowptr is not None
                and ops.message_stack_bwd_supported(inner.node_channels, inner.heads, ea2.shape[1], int(steps)))
        # the one-launch pair reads bond types, not edge_attr rows: the node then only needs edge_attr's SHAPE, and the
        # dst-ordered copy (a gather launch per batch) is made only for the per-op kernels
        ea = ea2 if both else g.sorted_edge_attr(edge_attr)
        pre_args = (None, None, None)
        if pre_train is not None:
            # training: the input LinearBlock runs inside the one-launch pair too when both kernels take this batch (its
            # weight / bias gradients come back from the same autograd node); otherwise it is simply called first
            if both:
                pre_args = (pre_train[0], pre_train[1], (pre_train[2], pre_train[3]))
            else:
                x = pre_block(x, batch=batch)
        out = Fn.MessageStackFn.apply(
            x, w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias,
            gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, ea, g,
            inner.heads, inner.node_channels, inner.negative_slope, fused[0], fused[1], bool(self.res), int(steps), p_drop, fi, pn,
            *pre_args)
        xs = list(out[:steps])
        return (xs if keep == "all" else xs[-1:]), out[steps].unsqueeze(0)

    def forward(self, x, edge_index, edge_attr, h=None, batch=None, num_graphs=None):
        identity = x
        if h is None:
            h = x.unsqueeze(0)
        x = self.dropout(self.norm(x, batch, num_graphs))
        if self.gru is None:
            x = self.conv(x, edge_index, edge_attr)
            x = x + identity if self.res else x
            return self.act(x), h
        gru = self.gru
        fused = _fusable_act(self.act, self.training)
        act_code, act_param = fused if fused is not None else (ops.ACT_NONE, 0.0)
        ident = identity if self.res else None
        inner = getattr(self.conv, "conv", None)
        if isinstance(inner, TripletMessage):
            g = G.graph_index(edge_index, x.shape[0])
            ea = g.sorted_edge_attr(edge_attr)
            w_ext, att_edge = inner.derived()
            fi = None
            if x is identity and gru.hidden_size == x.shape[1] == inner.node_channels:      # no norm / dropout in front of the conv
                fi = _fused_index(g, x, edge_attr, batch, num_graphs, inner.heads, inner.node_channels)
            if fi is not None and not _wants_grad(x, h, *self.parameters()):
                x_o, h_o = ops.message_stack_fwd(
                    x.contiguous(), h[0].contiguous(), w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias,
                    gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, g, fi, inner.heads, inner.node_channels,
                    1, inner.negative_slope, act_code, act_param, bool(self.res))
                x, h_new = x_o[0], h_o
            else:
                x, h_new = Fn.MessageBlockFn.apply(
                    x, ident, h[0], w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias,
                    gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, ea, g,
                    inner.heads, inner.node_channels, inner.negative_slope, act_code, act_param, fi)
        else:
            m = torch.celu(self.conv(x, edge_index, edge_attr))
            x, h_new = Fn.GRUUpdateFn.apply(m, h[0], ident, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0,
                                            gru.bias_hh_l0, act_code, act_param)
        if fused is None:
            x = self.act(x)
        return x, h_new.unsqueeze(0)


# --------------------------------------------------------------------------------------------------
# cross-graph interaction pool
# --------------------------------------------------------------------------------------------------
def dot_and_global_pool2(mol_out, pro_out, mol_batch, pro_batch, num_graphs: Optional[int] = None, pro_index=None,
                         num_pro_graphs: Optional[int] = None):
    """Per pair [max, mean] of X_mol X_pro^T — src_2gi_ddi/layer.py:270-283 — as one kernel launch, no host loop.
    `pro_index` (int [num_pairs], evaluation only): the protein side holds every DISTINCT graph once (`num_pro_graphs` of them)
    and pair g reads graph pro_index[g] — the reference collates one protein copy per pair (src_2gi_dti_scr/dataset.py:329-335)
    although a screening set has one protein per target (dataset.py:297)."""
    ptr_a, B = G.graph_ptr(mol_batch, num_graphs)
    if pro_index is None:
        ptr_b, _ = G.graph_ptr(pro_batch, B)
        return Fn.PairDotPoolFn.apply(mol_out, pro_out, ptr_a, ptr_b, B)
    ptr_b, _ = G.graph_ptr(pro_batch, num_pro_graphs)
    idx = pro_index if pro_index.dtype == torch.int32 else pro_index.to(torch.int32)
    return Fn.PairDotPoolFn.apply(mol_out, pro_out, ptr_a, ptr_b, B, idx.contiguous())
