"""Restatement of torch_scatter.scatter semantics (test infrastructure; see ../README.md).

out is zero-initialised; empty segments stay 0 (also for max/min); CPU accumulates in index order.
"""
import torch


def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src), dim


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    idx, dim = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = src.new_zeros(shape)
    if reduce in ("sum", "add"):
        return res.scatter_add_(dim, idx, src)
    if reduce == "mean":
        res = res.scatter_add_(dim, idx, src)
        cnt = src.new_zeros(shape).scatter_add_(dim, idx, torch.ones_like(src)).clamp_(min=1)
        return res / cnt
    if reduce in ("max", "min"):
        red = "amax" if reduce == "max" else "amin"
        return res.scatter_reduce(dim, idx, src, red, include_self=False)
    raise ValueError(reduce)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "max"), None
