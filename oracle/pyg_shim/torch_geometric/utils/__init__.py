"""torch_geometric.utils (1.7.2) — restated subset."""
import torch
from torch_scatter import scatter


def maybe_num_nodes(index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(index.max()) + 1 if index.numel() > 0 else 0


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    """utils/softmax.py @1.7.2: exp(src - segmax) / (segsum + 1e-16), per trailing column."""
    assert ptr is None and index is not None
    N = maybe_num_nodes(index, num_nodes)
    src_max = scatter(src, index, dim, dim_size=N, reduce="max").index_select(dim, index)
    out = (src - src_max).exp()
    out_sum = scatter(out, index, dim, dim_size=N, reduce="sum").index_select(dim, index)
    return out / (out_sum + 1e-16)


def degree(index, num_nodes=None, dtype=None):
    N = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((N,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, torch.ones((index.size(0),), dtype=out.dtype))


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(0, N, dtype=row.dtype).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        inv = ~mask
        loop_weight = torch.full((N,), fill_value, dtype=edge_weight.dtype)
        remaining = edge_weight[inv]
        if remaining.numel() > 0:
            loop_weight[row[inv]] = remaining
        edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_weight


def to_dense_batch(x, batch, fill_value=0.0):
    B = int(batch.max()) + 1
    num = degree(batch, B, dtype=torch.long)
    cum = torch.cat([num.new_zeros(1), num.cumsum(0)])
    M = int(num.max())
    idx = torch.arange(batch.size(0)) - cum[batch] + batch * M
    out = x.new_full((B * M, x.size(-1)), fill_value)
    out[idx] = x
    mask = torch.zeros(B * M, dtype=torch.bool)
    mask[idx] = True
    return out.view(B, M, -1), mask.view(B, M)
