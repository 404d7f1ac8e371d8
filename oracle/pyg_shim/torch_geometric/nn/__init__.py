"""torch_geometric.nn (1.7.2) — restated subset used by the reference's layer.py / model.py."""
import torch
from torch_scatter import scatter, scatter_add
from torch_geometric.utils import softmax, degree, to_dense_batch
from torch_geometric.nn.conv import MessagePassing, GCNConv, NNConv, GATConv


def global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="add")


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="mean")


def global_max_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="max")


def global_sort_pool(x, batch, k):
    """glob/sort.py @1.7.2: sort nodes per graph by last channel (descending), keep k, zero-pad."""
    fill_value = x.min().item() - 1
    batch_x, _ = to_dense_batch(x, batch, fill_value)
    B, N, D = batch_x.size()
    _, perm = batch_x[:, :, -1].sort(dim=-1, descending=True)
    arange = torch.arange(B, dtype=torch.long) * N
    perm = perm + arange.view(-1, 1)
    batch_x = batch_x.view(B * N, D)[perm.view(-1)].view(B, N, D)
    if N >= k:
        batch_x = batch_x[:, :k].contiguous()
    else:
        batch_x = torch.cat([batch_x, batch_x.new_full((B, k - N, D), fill_value)], dim=1)
    batch_x[batch_x == fill_value] = 0
    return batch_x.view(B, k * D)


class Set2Set(torch.nn.Module):
    """glob/set2set.py @1.7.2."""

    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, 2 * in_channels
        self.processing_steps, self.num_layers = processing_steps, num_layers
        self.lstm = torch.nn.LSTM(self.out_channels, self.in_channels, num_layers)
        self.lstm.reset_parameters()

    def forward(self, x, batch):
        batch_size = batch.max().item() + 1
        h = (x.new_zeros((self.num_layers, batch_size, self.in_channels)),
             x.new_zeros((self.num_layers, batch_size, self.in_channels)))
        q_star = x.new_zeros(batch_size, self.out_channels)
        for _ in range(self.processing_steps):
            q, h = self.lstm(q_star.unsqueeze(0), h)
            q = q.view(batch_size, self.in_channels)
            e = (x * q[batch]).sum(dim=-1, keepdim=True)
            a = softmax(e, batch, num_nodes=batch_size)
            r = scatter_add(a * x, batch, dim=0, dim_size=batch_size)
            q_star = torch.cat([q, r], dim=-1)
        return q_star


class GlobalAttention(torch.nn.Module):
    """glob/attention.py @1.7.2."""

    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn

    def forward(self, x, batch, size=None):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        size = batch[-1].item() + 1 if size is None else size
        gate = self.gate_nn(x).view(-1, 1)
        x = self.nn(x) if self.nn is not None else x
        gate = softmax(gate, batch, num_nodes=size)
        return scatter_add(gate * x, batch, dim=0, dim_size=size)


class BatchNorm(torch.nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class LayerNorm(torch.nn.Module):
    """norm/layer_norm.py @1.7.2: graph-wise statistics over nodes AND channels."""

    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(in_channels)) if affine else None
        self.bias = torch.nn.Parameter(torch.zeros(in_channels)) if affine else None

    def forward(self, x, batch=None):
        if batch is None:
            x = x - x.mean()
            out = x / (x.std(unbiased=False) + self.eps)
        else:
            B = int(batch.max()) + 1
            norm = degree(batch, B, dtype=x.dtype).clamp_(min=1).mul_(x.size(-1)).view(-1, 1)
            mean = scatter(x, batch, dim=0, dim_size=B, reduce="add").sum(dim=-1, keepdim=True) / norm
            x = x - mean[batch]
            var = scatter(x * x, batch, dim=0, dim_size=B, reduce="add").sum(dim=-1, keepdim=True) / norm
            out = x / (var + self.eps).sqrt()[batch]
        if self.weight is not None:
            out = out * self.weight + self.bias
        return out


class PairNorm(torch.nn.Module):
    """norm/pair_norm.py @1.7.2."""

    def __init__(self, scale=1.0, scale_individually=False, eps=1e-5):
        super().__init__()
        self.scale, self.scale_individually, self.eps = scale, scale_individually, eps

    def forward(self, x, batch=None):
        if batch is None:
            x = x - x.mean(dim=0, keepdim=True)
            if not self.scale_individually:
                return self.scale * x / (self.eps + x.pow(2).sum(-1).mean()).sqrt()
            return self.scale * x / (self.eps + x.norm(2, -1, keepdim=True))
        x = x - scatter(x, batch, dim=0, reduce="mean")[batch]
        if not self.scale_individually:
            return self.scale * x / torch.sqrt(
                self.eps + scatter(x.pow(2).sum(-1, keepdim=True), batch, dim=0, reduce="mean")[batch])
        return self.scale * x / (self.eps + x.norm(2, -1, keepdim=True))


class GraphSizeNorm(torch.nn.Module):
    def forward(self, x, batch=None):
        if batch is None:
            batch = torch.zeros(x.size(0), dtype=torch.long)
        inv_sqrt_deg = degree(batch, dtype=x.dtype).pow(-0.5)
        return x * inv_sqrt_deg[batch].view(-1, 1)


class InstanceNorm(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not used by the reference (imported only)")


class MessageNorm(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not used by the reference (imported only)")
