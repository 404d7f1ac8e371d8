"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-PyTorch (CPU, fp32 or fp64) restatement of the reference's message-passing hot path, written
op-for-op in the reference's own UNFUSED formulation so that it can serve both as the parity
checker and as the "port" CPU baseline.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; `glam_b200/` never does.

Parity status: PINNED against the reference's own `layer.py` executed in this container over a
restated PyG-1.7.2 shim (`oracle/pyg_shim`, `tests/golden/make_golden.py` -> `tests/golden/*.pt`,
checked by `tests/test_oracle_golden.py`).  The third-party arithmetic itself (torch-geometric
1.7.2 / torch-scatter, absent from this image) is restated from its published algorithm
 and is therefore "unpinned" at that level: the reference ships no tests or
golden vectors for this path.

Each function cites the reference lines it follows (paths under /root/reference).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

Tensor = torch.Tensor

# Optional operand pre-processing for every dense projection (tests use it to emulate TF32 tensor-core
# operands: fp32 words with the 13 low mantissa bits ignored).  None = exact.
MM_OPERAND_HOOK = None


def tf32_truncate(t: Tensor) -> Tensor:
    """Value of `t` as a TF32 tensor-core operand (straight-through for autograd)."""
    if t.dtype == torch.float32:
        q = (t.detach().contiguous().view(torch.int32) & -8192).view(torch.float32)
    else:
        q = (t.detach().float().contiguous().view(torch.int32) & -8192).view(torch.float32).to(t.dtype)
    return t + (q - t).detach()


def _mm(a: Tensor, b: Tensor) -> Tensor:
    if MM_OPERAND_HOOK is not None:
        a, b = MM_OPERAND_HOOK(a), MM_OPERAND_HOOK(b)
    return a @ b


# --------------------------------------------------------------------------------------------------
# third-party primitives (torch_scatter / torch_geometric.utils @1.7.2)
# --------------------------------------------------------------------------------------------------
def seg_sum(src: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_scatter.scatter(reduce='sum'): zero-initialised, accumulates in edge order on CPU."""
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def seg_max(src: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_scatter.scatter(reduce='max'): empty segments stay 0."""
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return out.scatter_reduce(0, idx, src, "amax", include_self=False)


def seg_softmax(a: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_geometric.utils.softmax @1.7.2 (called at src_1gp/layer.py:51,95):
    exp(a - max_seg) / (sum_seg + 1e-16), independently per trailing column."""
    m = seg_max(a, index, n).index_select(0, index)
    e = (a - m).exp()
    s = seg_sum(e, index, n).index_select(0, index)
    return e / (s + 1e-16)


# --------------------------------------------------------------------------------------------------
# a1 / a2: the triplet attention message layers
# --------------------------------------------------------------------------------------------------
def triplet_message(x, edge_index, edge_attr, weight_node, weight_edge, weight_triplet_att, weight_scale, bias,
                    heads: int = 3, negative_slope: float = 0.2, return_alpha: bool = False):
    """TripletMessage.forward/message/update — src_1gp/layer.py:36-61 (+ PyG propagate: x_j =
    x[edge_index[0]], x_i = x[edge_index[1]], scatter-add over edge_index[1])."""
    N, C = x.shape[0], weight_node.shape[0]
    xp = _mm(x, weight_node)                                # :37
    ep = edge_attr @ weight_edge                            # :38 (K = De: not a tensor-core contraction)
    src, dst = edge_index[0], edge_index[1]
    x_j = xp.index_select(0, src).view(-1, heads, C)        # :44
    x_i = xp.index_select(0, dst).view(-1, heads, C)        # :45
    e_ij = ep.view(-1, heads, C)                            # :46
    triplet = torch.cat([x_i, e_ij, x_j], dim=-1)           # :48
    alpha = (triplet * weight_triplet_att).sum(dim=-1)      # :49
    alpha = F.leaky_relu(alpha, negative_slope)             # :50
    alpha = seg_softmax(alpha, dst, N)                      # :51
    msg = alpha.view(-1, heads, 1) * e_ij * x_j             # :55
    agg = seg_sum(msg, dst, N).view(N, heads * C)           # aggr='add', node_dim=0 (:17), :58
    out = _mm(agg, weight_scale) + bias                     # :59-60
    return (out, alpha) if return_alpha else out


def triplet_message_light(x, edge_index, edge_attr, weight_node, weight_triplet_att, bias,
                          negative_slope: float = 0.2):
    """TripletMessageLight — src_1gp/layer.py:83-101 (single head, raw edge_attr in the logit only)."""
    N = x.shape[0]
    xp = _mm(x, weight_node)                                # :84
    src, dst = edge_index[0], edge_index[1]
    x_j, x_i = xp.index_select(0, src), xp.index_select(0, dst)
    triplet = torch.cat([x_i, edge_attr, x_j], dim=-1)      # :92
    alpha = (triplet * weight_triplet_att).sum(dim=-1)      # :93
    alpha = F.leaky_relu(alpha, negative_slope)             # :94
    alpha = seg_softmax(alpha, dst, N)                      # :95
    return seg_sum(alpha.view(-1, 1) * x_j, dst, N) + bias  # :97, :100


def gru_cell(m, h, weight_ih, weight_hh, bias_ih, bias_hh):
    """torch.nn.GRU, one layer, seq_len 1 (src_1gp/layer.py:247,262); gate order r, z, n."""
    gi = _mm(m, weight_ih.t()) + bias_ih
    gh = _mm(h, weight_hh.t()) + bias_hh
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def pair_norm(x, batch, eps: float = 1e-5, scale: float = 1.0):
    """torch_geometric.nn.PairNorm @1.7.2 defaults, with `batch` (wrapped at src_1gp/layer.py:179-185)."""
    B = int(batch.max()) + 1
    cnt = seg_sum(torch.ones_like(x[:, :1]), batch, B).clamp(min=1)
    x = x - (seg_sum(x, batch, B) / cnt).index_select(0, batch)
    ms = seg_sum(x.pow(2).sum(-1, keepdim=True), batch, B) / cnt
    return scale * x / torch.sqrt(eps + ms.index_select(0, batch))


# --------------------------------------------------------------------------------------------------
# a4 / a5: readouts, a6: cross-graph pool
# --------------------------------------------------------------------------------------------------
def global_attention(x, batch, num_graphs, gate_w, gate_b, nn_w, nn_b):
    """GlobalLAPool = PyG GlobalAttention(gate_nn=Linear(C,1), nn=Linear(C,2C)) — src_1gp/layer.py:206-220."""
    gate = (x @ gate_w.t() + gate_b).view(-1, 1)
    v = x @ nn_w.t() + nn_b
    gate = seg_softmax(gate, batch, num_graphs)
    return seg_sum(gate * v, batch, num_graphs)


def lstm_cell(inp, h, c, weight_ih, weight_hh, bias_ih, bias_hh):
    """torch.nn.LSTM one layer, one step; gate order i, f, g, o."""
    g = _mm(inp, weight_ih.t()) + bias_ih + _mm(h, weight_hh.t()) + bias_hh
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def set2set(x, batch, num_graphs, weight_ih, weight_hh, bias_ih, bias_hh, processing_steps: int = 3):
    """PyG Set2Set(in_channels=C, processing_steps=3) @1.7.2, constructed at src_1gp/model.py:41."""
    C = x.shape[1]
    h = x.new_zeros(num_graphs, C)
    c = x.new_zeros(num_graphs, C)
    q_star = x.new_zeros(num_graphs, 2 * C)
    for _ in range(processing_steps):
        h, c = lstm_cell(q_star, h, c, weight_ih, weight_hh, bias_ih, bias_hh)
        q = h
        e = (x * q.index_select(0, batch)).sum(dim=-1, keepdim=True)
        a = seg_softmax(e, batch, num_graphs)
        r = seg_sum(a * x, batch, num_graphs)
        q_star = torch.cat([q, r], dim=-1)
    return q_star


def dot_and_global_pool2(mol_out, pro_out, mol_batch, pro_batch):
    """src_2gi_ddi/layer.py:270-283 (identical in src_2gi_dti_scr): per pair [max, mean] of Xa @ Xb^T."""
    B = int(mol_batch.max()) + 1
    ma = torch.bincount(mol_batch, minlength=B).cumsum(0).tolist()
    pa = torch.bincount(pro_batch, minlength=B).cumsum(0).tolist()
    rows = []
    for i in range(B):
        a0, p0 = (ma[i - 1], pa[i - 1]) if i else (0, 0)
        item = mol_out[a0:ma[i]] @ pro_out[p0:pa[i]].t()
        rows.append(torch.stack([item.max(), item.mean()]))
    return torch.stack(rows)


# --------------------------------------------------------------------------------------------------
# module mirrors (same parameter names / registration order / init as the reference, so that
# state_dicts interchange with the reference and with glam_b200.layer)
# --------------------------------------------------------------------------------------------------
def global_pool5(x, batch, num_graphs):
    """GlobalPool5 (src_1gp/layer.py:197-203): [global_mean_pool | global_add_pool | global_sort_pool(k=3)]; the sort
    pool keeps each graph's first 3 rows by last channel, descending, ties in node order, zero padded
    (PyG global_sort_pool @1.7.2)."""
    C = x.shape[1]
    total = seg_sum(x, batch, num_graphs)
    cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype).view(-1, 1)
    top = x.new_zeros((num_graphs, 3, C))
    for g in range(num_graphs):
        idx = (batch == g).nonzero().view(-1)
        if idx.numel() == 0:
            continue
        order = torch.argsort(x[idx, -1], descending=True, stable=True)[:3]
        top[g, :order.numel()] = x[idx[order]]
    return torch.cat([total / cnt, total, top.view(num_graphs, 3 * C)], dim=-1)


def gcn_conv(x, edge_index, weight, bias):
    """PyG GCNConv(in,out) @1.7.2 defaults as `_GCNConv` builds it (src_1gp/layer.py:143-149): existing self edges are
    replaced by one unit self loop per node, symmetric normalisation by in-degree, linear, sum aggregation, bias."""
    N = x.shape[0]
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(N)
    src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
    deg = torch.zeros(N, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype))
    dis = deg.pow(-0.5)
    norm = dis[src] * dis[dst]
    xw = _mm(x, weight)
    return seg_sum(norm.view(-1, 1) * xw[src], dst, N) + bias


class TripletMessage(nn.Module):
    """src_1gp/layer.py:15-64."""

    def __init__(self, node_channels, edge_channels, heads=3, negative_slope=0.2):
        super().__init__()
        self.node_channels, self.heads, self.negative_slope = node_channels, heads, negative_slope
        self.weight_node = nn.Parameter(torch.empty(node_channels, heads * node_channels))
        self.weight_edge = nn.Parameter(torch.empty(edge_channels, heads * node_channels))
        self.weight_triplet_att = nn.Parameter(torch.empty(1, heads, 3 * node_channels))
        self.weight_scale = nn.Parameter(torch.empty(heads * node_channels, node_channels))
        self.bias = nn.Parameter(torch.empty(node_channels))
        for p in (self.weight_node, self.weight_edge, self.weight_triplet_att, self.weight_scale):
            nn.init.kaiming_uniform_(p)                      # :29-33
        nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr, size=None):
        return triplet_message(x, edge_index, edge_attr, self.weight_node, self.weight_edge,
                               self.weight_triplet_att, self.weight_scale, self.bias, self.heads,
                               self.negative_slope)


class TripletMessageLight(nn.Module):
    """src_1gp/layer.py:67-104."""

    def __init__(self, node_channels, edge_channels, negative_slope=0.2):
        super().__init__()
        self.node_channels, self.negative_slope = node_channels, negative_slope
        self.weight_node = nn.Parameter(torch.empty(node_channels, node_channels))
        self.weight_triplet_att = nn.Parameter(torch.empty(1, 2 * node_channels + edge_channels))
        self.bias = nn.Parameter(torch.empty(node_channels))
        nn.init.kaiming_uniform_(self.weight_node)
        nn.init.kaiming_uniform_(self.weight_triplet_att)
        nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr, size=None):
        return triplet_message_light(x, edge_index, edge_attr, self.weight_node, self.weight_triplet_att,
                                     self.bias, self.negative_slope)


class _Wrap(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, x, edge_index, edge_attr):
        return self.conv(x, edge_index, edge_attr)


def _TripletMessage(in_dim, out_dim, edge_in_dim):          # src_1gp/layer.py:125-131
    return _Wrap(TripletMessage(in_dim, edge_in_dim))


def _TripletMessageLight(in_dim, out_dim, edge_in_dim):     # src_1gp/layer.py:134-140
    return _Wrap(TripletMessageLight(in_dim, edge_in_dim))


class _None(nn.Module):
    def __init__(self, **params):
        super().__init__()

    def forward(self, x, batch=None):
        return x


class _PairNorm(nn.Module):
    def __init__(self, in_channels):
        super().__init__()

    def forward(self, x, batch=None):
        return pair_norm(x, batch)


class _GCNInner(nn.Module):
    """PyG GCNConv parameters: weight [in,out] (glorot), bias [out] (zeros)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -bound, bound)


class _GCNConv(nn.Module):
    """src_1gp/layer.py:143-149."""

    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        self.conv = _GCNInner(in_dim, out_dim)

    def forward(self, x, edge_index, edge_attr):
        return gcn_conv(x, edge_index, self.conv.weight, self.conv.bias)


def nn_conv(x, edge_index, edge_attr, lin1_w, lin1_b, lin2_w, lin2_b, root, bias):
    """PyG NNConv(in, out, nn=Sequential(Linear(De,32), ReLU(), Linear(32, in*out)), aggr='mean') @1.7.2 as `_NNConv` builds
    it (src_1gp/layer.py:115-122): message = x_j @ nn(edge_attr).view(-1,in,out), MEAN over in-edges (nodes without
    in-edges get 0), + x @ root + bias."""
    N, cin = x.shape
    cout = root.shape[1]
    src, dst = edge_index[0], edge_index[1]
    hid = torch.relu(_mm(edge_attr, lin1_w.t()) + lin1_b)
    theta = (_mm(hid, lin2_w.t()) + lin2_b).view(-1, cin, cout)
    msg = torch.matmul(x[src].unsqueeze(1), theta).squeeze(1)
    cnt = torch.zeros(N, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype)).clamp(min=1)
    return seg_sum(msg, dst, N) / cnt.view(-1, 1) + _mm(x, root) + bias


class _NNInner(nn.Module):
    """Parameter container with PyG NNConv's names: nn.{0,2}.{weight,bias}, root, bias."""

    def __init__(self, in_channels, out_channels, edge_dim):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.nn = nn.Sequential(nn.Linear(edge_dim, 32), nn.ReLU(), nn.Linear(32, in_channels * out_channels))
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels)
        nn.init.uniform_(self.root, -bound, bound)


class _NNConv(nn.Module):
    """src_1gp/layer.py:115-122 (the reference's default mol_block, src_1gp/run.py:21)."""

    def __init__(self, in_dim, out_dim, edge_in_dim):
        super().__init__()
        self.conv = _NNInner(in_dim, out_dim, edge_in_dim)

    def forward(self, x, edge_index, edge_attr):
        c = self.conv
        return nn_conv(x, edge_index, edge_attr, c.nn[0].weight, c.nn[0].bias, c.nn[2].weight, c.nn[2].bias, c.root, c.bias)


_CONVS = {"_TripletMessage": _TripletMessage, "_TripletMessageLight": _TripletMessageLight, "_GCNConv": _GCNConv,
          "_NNConv": _NNConv}
_NORMS = {"_None": _None, "_PairNorm": _PairNorm}
_ACTS = {"_None": _None, "ReLU": nn.ReLU, "CELU": nn.CELU, "LeakyReLU": nn.LeakyReLU, "RReLU": nn.RReLU}


def _make_dropout(spec: str) -> nn.Module:
    if spec.startswith("_None"):
        return _None()
    assert spec.startswith("Dropout(") and spec.endswith(")")
    return nn.Dropout(float(spec[len("Dropout("):-1]))


class MessageBlock(nn.Module):
    """src_1gp/layer.py:240-267."""

    def __init__(self, in_dim=32, out_dim=64, in_edge_dim=13, norm="_None", dropout="Dropout(0.2)",
                 conv="_TripletMessage", act="ReLU", res=True):
        super().__init__()
        self.norm = _NORMS[norm](in_channels=in_dim)
        self.dropout = _make_dropout(dropout)
        self.conv = _CONVS[conv](in_dim, out_dim, in_edge_dim)
        self.gru = nn.GRU(in_dim, out_dim)
        if conv in ("_GCNConv", "_GATConv"):
            self.gru = None                                  # :248
        self.act = _ACTS[act]()
        self.res = res

    def forward(self, x, edge_index, edge_attr, h=None, batch=None):
        identity = x                                         # :253
        if h is None:
            h = x.unsqueeze(0)                               # :254 (pre-norm x)
        x = self.dropout(self.norm(x, batch))                # :255-256
        x = self.conv(x, edge_index, edge_attr)              # :259
        if self.gru is None:                                 # :260 (GCN / GAT blocks have no GRU)
            x = x + identity if self.res else x
            return self.act(x), h
        x = F.celu(x)                                        # :261
        hn = gru_cell(x, h[0], self.gru.weight_ih_l0, self.gru.weight_hh_l0,
                      self.gru.bias_ih_l0, self.gru.bias_hh_l0)  # :262
        x = hn + identity if self.res else hn                # :265
        return self.act(x), hn.unsqueeze(0)                  # :266-267


class _GlobalAttention(nn.Module):
    def __init__(self, gate_nn, nn_):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn_


class GlobalLAPool(nn.Module):
    """src_1gp/layer.py:206-220; parameter names pool.gate_nn.*, pool.nn.*."""

    def __init__(self, in_channels, **params):
        super().__init__()
        self.pool = _GlobalAttention(nn.Linear(in_channels, 1), nn.Linear(in_channels, 2 * in_channels))

    def forward(self, x, batch):
        B = int(batch[-1]) + 1
        return global_attention(x, batch, B, self.pool.gate_nn.weight, self.pool.gate_nn.bias,
                                self.pool.nn.weight, self.pool.nn.bias)


class Set2Set(nn.Module):
    """PyG Set2Set @1.7.2 (imported by src_1gp/model.py:2); parameter names lstm.*."""

    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()
        assert num_layers == 1
        self.in_channels, self.out_channels, self.processing_steps = in_channels, 2 * in_channels, processing_steps
        self.lstm = nn.LSTM(self.out_channels, in_channels, num_layers)

    def forward(self, x, batch):
        B = int(batch.max()) + 1
        l = self.lstm
        return set2set(x, batch, B, l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0,
                       self.processing_steps)


class GlobalPool5(nn.Module):
    """src_1gp/layer.py:197-203."""

    def __init__(self, **params):
        super().__init__()

    def forward(self, x, batch):
        return global_pool5(x, batch, int(batch[-1]) + 1)


_READOUTS = {"Set2Set": Set2Set, "GlobalLAPool": GlobalLAPool, "GlobalPool5": GlobalPool5}
_READOUT_WIDTH = {"Set2Set": 2, "GlobalLAPool": 2, "GlobalPool5": 5}


class LinearBlock(nn.Module):
    """src_1gp/layer.py:223-237."""

    def __init__(self, in_dim=32, out_dim=64, norm="_None", dropout="_None()", act="ReLU"):
        super().__init__()
        self.norm = _NORMS[norm](in_channels=in_dim)
        self.dropout = _make_dropout(dropout)
        self.linear = nn.Linear(in_dim, out_dim)
        self.act = _ACTS[act]()

    def forward(self, x, batch=None):
        return self.act(self.linear(self.dropout(self.norm(x, batch))))


class ArchitectureGP(nn.Module):
    """src_1gp/model.py:23-62 (GLAM-GP)."""

    def __init__(self, mol_in_dim=15, mol_edge_in_dim=4, hid_dim_alpha=4, e_dim=1024, out_dim=1,
                 mol_block="_TripletMessage", message_steps=3, mol_readout="Set2Set",
                 pre_norm="_None", graph_norm="_None", flat_norm="_None", end_norm="_None",
                 pre_do="_None()", graph_do="_None()", flat_do="_None()", end_do="_None()",
                 pre_act="ReLU", graph_act="ReLU", flat_act="ReLU", graph_res=True):
        super().__init__()
        hid = mol_in_dim * hid_dim_alpha
        self.mol_lin0 = LinearBlock(mol_in_dim, hid, norm=pre_norm, dropout=pre_do, act=pre_act)
        self.mol_conv = MessageBlock(hid, hid, mol_edge_in_dim, norm=graph_norm, dropout=graph_do,
                                     conv=mol_block, act=graph_act, res=graph_res)
        self.message_steps = message_steps
        self.mol_readout = _READOUTS[mol_readout](in_channels=hid, processing_steps=3)
        self.mol_flat = LinearBlock(_READOUT_WIDTH[mol_readout] * hid, e_dim, norm=flat_norm, dropout=flat_do, act=flat_act)
        self.lin_out1 = LinearBlock(e_dim, out_dim, norm=end_norm, dropout=end_do, act="_None")

    def forward(self, data_mol):
        xm = self.mol_lin0(data_mol.x, batch=data_mol.batch)
        hm = None
        for _ in range(self.message_steps):
            xm, hm = self.mol_conv(xm, data_mol.edge_index, data_mol.edge_attr, h=hm, batch=data_mol.batch)
        outm = self.mol_flat(self.mol_readout(xm, data_mol.batch))
        return self.lin_out1(outm)


class ArchitecturePair(nn.Module):
    """Two-tower models: src_2gi_ddi/model.py:9-61 (prefixes mol1/mol2) and
    src_2gi_dti_scr/model.py:14-68 (prefixes mol/pro)."""

    def __init__(self, a_in_dim=15, b_in_dim=15, a_edge_in_dim=4, b_edge_in_dim=4, prefixes=("mol1", "mol2"),
                 hid_dim_alpha=4, e_dim=1024, out_dim=1, a_block="_TripletMessage", b_block="_TripletMessage",
                 message_steps=3, a_readout="Set2Set", b_readout="Set2Set",
                 graph_norm="_None", graph_act="ReLU", pre_act="ReLU", flat_act="ReLU", end_act="ReLU",
                 graph_res=True):
        super().__init__()
        hid = a_in_dim * hid_dim_alpha
        self.pa, self.pb = prefixes
        mods = {}
        for p, din, de, blk, ro in ((self.pa, a_in_dim, a_edge_in_dim, a_block, a_readout),
                                    (self.pb, b_in_dim, b_edge_in_dim, b_block, b_readout)):
            mods[p + "_lin0"] = LinearBlock(din, hid, act=pre_act)
        for p, din, de, blk, ro in ((self.pa, a_in_dim, a_edge_in_dim, a_block, a_readout),
                                    (self.pb, b_in_dim, b_edge_in_dim, b_block, b_readout)):
            mods[p + "_conv"] = MessageBlock(hid, hid, de, norm=graph_norm, dropout="_None()", conv=blk,
                                             act=graph_act, res=graph_res)
        self.message_steps = message_steps
        for p, din, de, blk, ro in ((self.pa, a_in_dim, a_edge_in_dim, a_block, a_readout),
                                    (self.pb, b_in_dim, b_edge_in_dim, b_block, b_readout)):
            mods[p + "_readout"] = _READOUTS[ro](in_channels=hid, processing_steps=3)
        for p, ro in ((self.pa, a_readout), (self.pb, b_readout)):
            mods[p + "_flat"] = LinearBlock(_READOUT_WIDTH[ro] * hid, hid, act=flat_act)
        for k, v in mods.items():
            self.add_module(k, v)
        self.lin_out0 = LinearBlock(hid * 2 + message_steps * 2, e_dim, act=end_act)
        self.lin_out1 = LinearBlock(e_dim, out_dim, act="_None")

    def forward(self, da, db):
        g = lambda n: getattr(self, n)
        xa = g(self.pa + "_lin0")(da.x, batch=da.batch)
        xb = g(self.pb + "_lin0")(db.x, batch=db.batch)
        ha = hb = None
        fusion = []
        for _ in range(self.message_steps):
            xa, ha = g(self.pa + "_conv")(xa, da.edge_index, da.edge_attr, h=ha, batch=da.batch)
            xb, hb = g(self.pb + "_conv")(xb, db.edge_index, db.edge_attr, h=hb, batch=db.batch)
            fusion.append(dot_and_global_pool2(xa, xb, da.batch, db.batch))
        oa = g(self.pa + "_flat")(g(self.pa + "_readout")(xa, da.batch))
        ob = g(self.pb + "_flat")(g(self.pb + "_readout")(xb, db.batch))
        out = self.lin_out0(torch.cat([oa, ob, torch.cat(fusion, dim=-1)], dim=-1))
        return self.lin_out1(out)
