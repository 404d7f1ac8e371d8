"""Shared helpers for the parity tests."""
import types

import torch


def f64_case(fx, name):
    """Golden f64 cases store only outputs/grads; inputs, state and cotangents are the f32 ones upcast."""
    base = dict(fx[name.replace("_f64", "_f32")])
    up = lambda v: v.double() if torch.is_tensor(v) and v.is_floating_point() else v
    out = {k: up(v) for k, v in base.items() if k not in ("state", "grad_params")}
    out["state"] = {k: up(v) for k, v in base.get("state", {}).items()}
    out.update(fx[name])
    return out


def case(fx, name):
    return f64_case(fx, name) if name.endswith("_f64") else fx[name]


def ns(x, edge_index, edge_attr, batch):
    return types.SimpleNamespace(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=batch)


def tol_check(ours, ref32, ref64, name="", rtol=1e-4, atol=1e-5):
    """SURVEY.md §8c tolerance: |ours - ref64| <= max(2*|ref32 - ref64|, atol + rtol*|ref64|), evaluated with a
    tensor-level scale so that near-zero entries are judged against the magnitude of the tensor."""
    o, r32, r64 = ours.double().cpu(), ref32.double().cpu(), ref64.double().cpu()
    scale = r64.abs().max().clamp(min=1e-30)
    err = (o - r64).abs().max()
    base = (r32 - r64).abs().max()
    bound = torch.maximum(2 * base, atol + rtol * scale)
    assert err <= bound, f"{name}: err {err:.3e} > bound {bound:.3e} (fp32-ref err {base:.3e}, scale {scale:.3e})"
