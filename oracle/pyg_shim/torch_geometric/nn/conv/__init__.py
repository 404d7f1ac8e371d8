"""torch_geometric.nn.conv (1.7.2) — restated subset: MessagePassing, GCNConv, NNConv."""
import inspect
import math
import torch
from torch.nn import Parameter
from torch_scatter import scatter
from torch_geometric.utils import add_remaining_self_loops


class MessagePassing(torch.nn.Module):
    """conv/message_passing.py @1.7.2, flow='source_to_target' only.

    propagate(): `<arg>_j` = arg.index_select(node_dim, edge_index[0]) (sources),
    `<arg>_i` = arg.index_select(node_dim, edge_index[1]) (targets), `edge_index_i` = edge_index[1],
    `size_i` = number of target nodes; message() -> aggregate() = scatter over edge_index[1] along
    node_dim with dim_size=N -> update().
    """

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]
        self._upd_params = [p for p in inspect.signature(self.update).parameters][1:]

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() > 0 and n is None and v.size(self.node_dim) != edge_index.size(1):
                n = v.size(self.node_dim)
        if size is not None:
            n = size[1] if isinstance(size, (tuple, list)) else size
        if n is None:
            n = kwargs["x"].size(self.node_dim)
        coll = {}
        for name in self._msg_params:
            if name.endswith("_j") or name.endswith("_i"):
                base, which = name[:-2], name[-2:]
                if base == "edge_index":
                    coll[name] = j if which == "_j" else i
                elif base == "size":
                    coll[name] = n
                else:
                    data = kwargs[base]
                    coll[name] = data.index_select(self.node_dim, j if which == "_j" else i)
            else:
                coll[name] = kwargs.get(name)
        out = self.message(**coll)
        out = scatter(out, i, dim=self.node_dim, dim_size=n, reduce=self.aggr)
        return self.update(out, **{k: kwargs.get(k) for k in self._upd_params})

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


def _glorot(t):
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-stdv, stdv)


class GCNConv(MessagePassing):
    """conv/gcn_conv.py @1.7.2 defaults: self-loops, sym norm, weight [in,out] glorot, bias zeros."""

    def __init__(self, in_channels, out_channels):
        super().__init__(aggr="add")
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = Parameter(torch.Tensor(out_channels))
        _glorot(self.weight)
        self.bias.data.zero_()

    def forward(self, x, edge_index, edge_weight=None):
        N = x.size(0)
        ew = torch.ones((edge_index.size(1),), dtype=x.dtype)
        edge_index, ew = add_remaining_self_loops(edge_index, ew, 1.0, N)
        row, col = edge_index[0], edge_index[1]
        deg = scatter(ew, col, dim=0, dim_size=N, reduce="sum")
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        norm = dis[row] * ew * dis[col]
        x = x @ self.weight
        out = scatter(norm.view(-1, 1) * x.index_select(0, row), col, dim=0, dim_size=N, reduce="sum")
        return out + self.bias


class NNConv(MessagePassing):
    """conv/nn_conv.py @1.7.2: msg = x_j @ nn(edge_attr).view(-1,in,out); + x @ root + bias."""

    def __init__(self, in_channels, out_channels, nn, aggr="add", root_weight=True, bias=True):
        super().__init__(aggr=aggr)
        self.in_channels, self.out_channels, self.nn = in_channels, out_channels, nn
        self.root = Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = Parameter(torch.Tensor(out_channels))
        for m in nn.modules():                               # reset_parameters @1.7.2: reset(nn); uniform(in, root); zeros(bias)
            if m is not nn and hasattr(m, "reset_parameters"):
                m.reset_parameters()
        bound = 1.0 / math.sqrt(in_channels)
        self.root.data.uniform_(-bound, bound)
        self.bias.data.zero_()

    def forward(self, x, edge_index, edge_attr=None, size=None):
        out = self.propagate(edge_index, x=x, edge_attr=edge_attr, size=size)
        return out + x @ self.root + self.bias

    def message(self, x_j, edge_attr):
        weight = self.nn(edge_attr).view(-1, self.in_channels, self.out_channels)
        return torch.matmul(x_j.unsqueeze(1), weight).squeeze(1)


class GATConv(MessagePassing):
    def __init__(self, *a, **k):
        raise NotImplementedError("GATConv is outside the hot path (SURVEY.md §2.1) and not restated")
