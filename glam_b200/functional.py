"""Autograd bindings of the hot-path kernels.

Every Function below is a fixed sequence of C-ABI calls (glam_b200.ops); nothing here computes with
torch ops on node- or edge-sized tensors.  Backward formulas follow SURVEY.md Appendix C.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops
from .ops import (ACT_NONE, EPI_ACCUM, EPI_CELU, EPI_MUL_CELU_GRAD, EPI_NONE)


def _c(t):
    return None if t is None else (t if t.is_contiguous() else t.contiguous())


# --------------------------------------------------------------------------------------------------
# side stream for weight gradients: they are off the critical path of backward (nothing downstream in the same
# step consumes them), and each is a latency-bound streaming kernel, so they overlap with the dgrad chain.
# Every fork starts with side.wait_stream(main) and backward joins before returning, which also keeps the caching
# allocator's stream-ordered reuse valid.  Works under CUDA-graph capture (fork/join become graph branches).
# --------------------------------------------------------------------------------------------------
_side_streams = {}
# Off by default: measured gain was ~3 % (two tensor-core CTAs cannot share an SM's shared memory), and tensors
# allocated on the main stream but consumed on the side stream would need record_stream()/a join before every free.
# When enabled, every producer below joins before its operands go out of scope.
OVERLAP_WGRAD = False


class _Side:
    """with _Side(t): ... runs the body on the device's side stream after everything enqueued so far on the
    current stream."""

    def __init__(self, ref: torch.Tensor):
        self.dev = ref.device
        self.on = OVERLAP_WGRAD and ref.is_cuda

    def __enter__(self):
        if not self.on:
            return self
        self.main = torch.cuda.current_stream(self.dev)
        side = _side_streams.get(self.dev)
        if side is None:
            side = _side_streams[self.dev] = torch.cuda.Stream(device=self.dev)
        self.side = side
        side.wait_stream(self.main)
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.on:
            self.ctx.__exit__(*a)
        return False


def _join(ref: torch.Tensor):
    if OVERLAP_WGRAD and ref.is_cuda:
        side = _side_streams.get(ref.device)
        if side is not None:
            torch.cuda.current_stream(ref.device).wait_stream(side)


# --------------------------------------------------------------------------------------------------
# shared forward/backward pieces
# --------------------------------------------------------------------------------------------------
def _conv_fwd(x, w_ext, w_edge, att_edge, w_scale, bias, ea, g, heads, channels, slope, epilogue):
    hc = heads * channels                                                      # [N, ldxp]: xp | s_i | s_j | 0
    xpe = ops.gemm(x, w_ext, exact_cols=(hc, hc + 2 * heads))                  # logit columns always exact fp32
    agg, alpha = ops.triplet_edge_fwd(xpe, ea, w_edge, att_edge, g, heads, channels, slope)
    out = agg if w_scale is None else ops.gemm(agg, w_scale, bias=bias, epilogue=epilogue)
    return xpe, agg, alpha, out


def _conv_bwd(g_pre, x, w_ext, w_edge, att_edge, w_scale, xpe, agg, alpha, ea, g, heads, channels, slope):
    """g_pre: gradient w.r.t. the layer output before any activation ([N,C], or [N,HC] for the Light layer)."""
    if w_scale is not None:
        with _Side(g_pre):
            g_w_scale, g_bias = ops.gemm_tn_ex(agg, g_pre, want_colsum=True)
        g_agg = ops.gemm(g_pre, w_scale, transpose_w=True)                     # [N,HC]
    else:
        g_agg, g_w_scale, g_bias = g_pre, None, None
    g_xpe, g_logit, g_w_edge = ops.triplet_edge_bwd(xpe, ea, w_edge, att_edge, alpha, g_agg, g, heads, channels, slope)
    with _Side(g_xpe):
        g_att_edge, _ = ops.gemm_tn_ex(ea, g_logit)                            # [De,H]
        g_w_ext, _ = ops.gemm_tn_ex(x, g_xpe)                                  # [C,ldxp]
        # the 2H logit columns are near-total cancellations (softmax gradients are zero-sum per destination): exact fp32
        hc = heads * channels
        ops.gemm_tn_ex(x, g_xpe[:, hc:hc + 2 * heads], out=g_w_ext[:, hc:hc + 2 * heads])
    g_x = ops.gemm(g_xpe, w_ext, transpose_w=True)                             # [N,C]
    _join(g_x)
    return g_x, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias


def _gru_fwd(m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param):
    if ops.gru_fused_supported(m, h, h.shape[1]):
        return ops.gru_fused_fwd(m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param)   # rzn, gh (n-part), h_new, x_out
    gi = ops.gemm(m, w_ih, transpose_w=True, bias=b_ih)                        # [N,3C]
    gh = ops.gemm(h, w_hh, transpose_w=True, bias=b_hh)
    h_new, x_out = ops.gru_gates_fwd(gi, gh, h, identity, act, act_param)      # gi now holds r|z|n
    return gi, gh, h_new, x_out


def _gru_bwd(g_x_out, g_h_new, rzn, gh, h, m, x_out, w_ih, w_hh, act, act_param, want_identity, celu_aux):
    g_gi, g_gh, g_h_prev, g_id = ops.gru_gates_bwd(rzn, gh, h, x_out, g_x_out, g_h_new, act, act_param, want_identity)
    with _Side(g_gi):
        g_w_ih, g_b_ih = ops.gemm_tn_ex(m, g_gi, transpose_out=True, want_colsum=True)     # (m^T g_gi)^T = [3C,C]
        g_w_hh, g_b_hh = ops.gemm_tn_ex(h, g_gh, transpose_out=True, want_colsum=True)
    if celu_aux is not None:                                                   # m = celu(pre): return d/d pre
        g_m = ops.gemm(g_gi, w_ih, epilogue=EPI_MUL_CELU_GRAD, aux=celu_aux)
    else:
        g_m = ops.gemm(g_gi, w_ih)
    ops.gemm(g_gh, w_hh, epilogue=EPI_ACCUM, out=g_h_prev)
    _join(g_m)                         # g_gi / g_gh die with this frame: the side-stream readers must be ordered before that
    return g_m, g_h_prev, g_id, g_w_ih, g_w_hh, g_b_ih, g_b_hh


# --------------------------------------------------------------------------------------------------
# TripletMessage / TripletMessageLight propagate (src_1gp/layer.py:36-61, :83-101)
# --------------------------------------------------------------------------------------------------
class TripletConvFn(Function):
    """out = (segment-softmax attention aggregate) @ w_scale + bias; for the Light layer (w_edge and w_scale
    None) returns the aggregate itself and the caller adds the bias."""

    @staticmethod
    def forward(ctx, x, w_ext, w_edge, att_edge, w_scale, bias, ea, g, heads, channels, slope, fi):
        x, w_ext, w_edge, att_edge, w_scale, bias = map(_c, (x, w_ext, w_edge, att_edge, w_scale, bias))
        ops._need_cuda(x, w_ext)
        if fi is not None and w_scale is not None:
            N, dev = x.shape[0], x.device
            new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            sv = dict(XPE=new(1, N, w_ext.shape[1]), AGG=new(1, N, heads * channels), ALPHA=new(1, ea.shape[0], heads))
            out, _ = ops.message_stack_fwd(x, None, w_ext, w_edge, att_edge, w_scale, bias, None, None, None, None, g, fi, heads,
                                           channels, 1, slope, ACT_NONE, 0.0, False, conv_only=True, save=sv)
            xpe, agg, alpha, out = sv["XPE"][0], sv["AGG"][0], sv["ALPHA"][0], out[0]
        else:
            xpe, agg, alpha, out = _conv_fwd(x, w_ext, w_edge, att_edge, w_scale, bias, ea, g, heads, channels, slope, EPI_NONE)
        ctx.save_for_backward(x, w_ext, w_edge, att_edge, w_scale, xpe, agg, alpha, ea)
        ctx.g, ctx.cfg = g, (heads, channels, slope)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, w_ext, w_edge, att_edge, w_scale, xpe, agg, alpha, ea = ctx.saved_tensors
        heads, channels, slope = ctx.cfg
        g_x, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias = _conv_bwd(
            _c(g_out), x, w_ext, w_edge, att_edge, w_scale, xpe, agg, alpha, ea, ctx.g, heads, channels, slope)
        return g_x, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------------
# GRU node update (src_1gp/layer.py:260-266) on a given message m
# --------------------------------------------------------------------------------------------------
class GRUUpdateFn(Function):
    @staticmethod
    def forward(ctx, m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param):
        m, h, identity, w_ih, w_hh, b_ih, b_hh = map(_c, (m, h, identity, w_ih, w_hh, b_ih, b_hh))
        ops._need_cuda(m, h)
        rzn, gh, h_new, x_out = _gru_fwd(m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param)
        ctx.save_for_backward(m, h, w_ih, w_hh, rzn, gh, x_out)
        ctx.cfg = (act, act_param, identity is not None)
        ctx.set_materialize_grads(False)
        return x_out, h_new

    @staticmethod
    def backward(ctx, g_x_out, g_h_new):
        m, h, w_ih, w_hh, rzn, gh, x_out = ctx.saved_tensors
        act, act_param, has_id = ctx.cfg
        g_m, g_h, g_id, g_w_ih, g_w_hh, g_b_ih, g_b_hh = _gru_bwd(
            _c(g_x_out), _c(g_h_new), rzn, gh, h, m, x_out, w_ih, w_hh, act, act_param, has_id, None)
        _join(g_m)
        return g_m, g_h, g_id, g_w_ih, g_w_hh, g_b_ih, g_b_hh, None, None


# --------------------------------------------------------------------------------------------------
# fused MessageBlock core: conv -> CELU -> GRU -> (+identity) -> act   (src_1gp/layer.py:259-266)
# --------------------------------------------------------------------------------------------------
def _stack_buffers(x0, S, H, C, ld, E, tiled_gates=False):
    """The stacked activations of `S` message steps that backward consumes (see MessageStackFn).  tiled_gates: the gate-side
    tensors in the tile-blocked layout the one-launch backward reads (GT [S,N,7C]) instead of row-major RZN / GH."""
    N, dev = x0.shape[0], x0.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    sv = dict(X=new(S + 1, N, C), HH=new(S + 1, N, C), XPE=new(S, N, ld), AGG=new(S, N, H * C), ALPHA=new(S, E, H))
    if tiled_gates:
        # GT: gate side, tile-blocked; MH: rows m | h_in | 1 — with the backward's G4 rows the one operand pair of both GRU
        # weight gradients (M alone is not needed then: the backward kernel reads m from GT)
        sv.update(GT=new(S, N, 7 * C), MH=new(S, N, 2 * C + 4))
    else:
        sv.update(M=new(S, N, C), RZN=new(S, N, 3 * C), GH=new(S, N, C))
    return sv


def _gru_wgrads(MH, G4, C):
    """Both GRU weight gradients and both bias gradients from ONE contraction: [m | h | 1]^T [g_r | g_z | g_n | g_n r] over all
    rows ([2C+4] x [4C]; the constant column of MH makes row 2C the column sums of G4).  G_GI = G4[:, :3C], G_GH = G4[:, :2C] |
    G4[:, 3C:] (the r and z gradients are shared by the input- and the hidden-side products)."""
    Dt, _ = ops.gemm_tn_ex(MH, G4, transpose_out=True)                         # [4C, 2C+4]
    g_w_ih = Dt[:3 * C, :C]
    g_w_hh = torch.cat([Dt[:2 * C, C:2 * C], Dt[3 * C:, C:2 * C]])
    g_b_ih = Dt[:3 * C, 2 * C]
    g_b_hh = torch.cat([Dt[:2 * C, 2 * C], Dt[3 * C:, 2 * C]])
    return g_w_ih, g_w_hh, g_b_ih, g_b_hh


USE_FUSED_BWD = True       # one-launch backward of the message stack (tests flip it to compare with the per-op path)


class MessageBlockFn(Function):
    @staticmethod
    def forward(ctx, x, identity, h, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh, ea, g,
                heads, channels, slope, act, act_param, fi):
        (x, identity, h, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh) = map(
            _c, (x, identity, h, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh))
        ops._need_cuda(x, h)
        ctx.fused = None
        if fi is not None and (identity is None or identity is x):
            # one launch for the whole block (csrc/mp_fused.cu); the tensors backward reads come out as side outputs — for the
            # one-launch backward (csrc/mp_fused_bwd.cu, steps = 1, its own h0) with the gate side in the tile-blocked layout
            fused_bwd = (USE_FUSED_BWD and g.src_rowptr is not None and ops.message_stack_bwd_supported(channels, heads, ea.shape[1], 1))
            sv = _stack_buffers(x, 1, heads, channels, w_ext.shape[1], ea.shape[0], tiled_gates=fused_bwd)
            ops.message_stack_fwd(x, h, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh, g, fi, heads, channels, 1,
                                  slope, act, act_param, identity is not None, save=sv)
            xpe, agg, alpha, x_out, h_new = sv["XPE"][0], sv["AGG"][0], sv["ALPHA"][0], sv["X"][1], sv["HH"][1]
            m, rzn, gh = (None, None, None) if fused_bwd else (sv["M"][0], sv["RZN"][0], sv["GH"][0])
            if fused_bwd:
                ctx.fused = (sv, fi)
        else:
            xpe, agg, alpha, m = _conv_fwd(x, w_ext, w_edge, att_edge, w_scale, bias, ea, g, heads, channels, slope, EPI_CELU)
            rzn, gh, h_new, x_out = _gru_fwd(m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param)
        ctx.save_for_backward(x, h, w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, xpe, agg, alpha, m, rzn, gh, x_out, ea)
        ctx.g, ctx.cfg = g, (heads, channels, slope, act, act_param, identity is not None)
        ctx.set_materialize_grads(False)
        return x_out, h_new

    @staticmethod
    def backward(ctx, g_x_out, g_h_new):
        (x, h, w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, xpe, agg, alpha, m, rzn, gh, x_out, ea) = ctx.saved_tensors
        heads, channels, slope, act, act_param, has_id = ctx.cfg
        if ctx.fused is not None:
            # the whole reverse chain of the block in one launch; the residual's gradient is part of g_x (identity IS x)
            sv, fi = ctx.fused
            N, C, HC, ld = x.shape[0], channels, heads * channels, xpe.shape[1]
            new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=x.device)
            G4, G_PRE, G_XPE = new(1, N, 4 * C), new(1, N, C), new(1, N, ld)
            (g_x, g_h), g_w_edge, g_att_edge = ops.message_stack_bwd(
                sv, [_c(g_x_out)], _c(g_h_new), w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, ctx.g, fi, heads, C, 1, slope, act,
                act_param, has_id, None, None, G_PRE, G_XPE, separate_h0=True, G4=G4)
            g_w_ih, g_w_hh, g_b_ih, g_b_hh = _gru_wgrads(sv["MH"].view(N, 2 * C + 4), G4.view(N, 4 * C), C)
            g_w_scale, g_bias = ops.gemm_tn_ex(agg, G_PRE[0], want_colsum=True)
            g_w_ext, _ = ops.gemm_tn_ex(x, G_XPE[0])
            ops.gemm_tn_ex(x, G_XPE[0][:, HC:HC + 2 * heads], out=g_w_ext[:, HC:HC + 2 * heads])      # logit columns: exact fp32
            return (g_x, None, g_h, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias, g_w_ih, g_w_hh, g_b_ih, g_b_hh,
                    None, None, None, None, None, None, None, None)
        g_pre, g_h, g_id, g_w_ih, g_w_hh, g_b_ih, g_b_hh = _gru_bwd(
            _c(g_x_out), _c(g_h_new), rzn, gh, h, m, x_out, w_ih, w_hh, act, act_param, has_id, m)
        g_x, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias = _conv_bwd(
            g_pre, x, w_ext, w_edge, att_edge, w_scale, xpe, agg, alpha, ea, ctx.g, heads, channels, slope)
        return (g_x, g_id, g_h, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias, g_w_ih, g_w_hh, g_b_ih, g_b_hh,
                None, None, None, None, None, None, None, None)


# --------------------------------------------------------------------------------------------------
# `message_steps` applications of ONE MessageBlock (shared weights; src_1gp/model.py:60-62 loops the same block) as a
# single autograd node.  Activations of all steps live in stacked [S, N, .] buffers, so every weight gradient is ONE
# fixed-order A^T B over S*N rows instead of S contractions that autograd then adds up, and the derived weights
# (w_ext, att_edge) are prepared once.  Dropout on the block input (graph_do) stays torch's (philox, graph-safe).
# --------------------------------------------------------------------------------------------------
class MessageStackFn(Function):
    @staticmethod
    def forward(ctx, x0, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh, ea, g,
                heads, channels, slope, act, act_param, res, steps, p_drop, fi, pn=None, w_pre=None, b_pre=None, pre_act=None):
        """pn = (graph_ptr, num_graphs, eps): PairNorm on every step's block input (src_1gp/layer.py:255) inside the node.
        w_pre / b_pre / pre_act = (code, param): the model's input LinearBlock (src_1gp/model.py:49) applied inside the one-launch
        kernels — x0 is then the RAW feature matrix [N, raw_dim] and the node also returns that block's weight / bias gradients
        (only offered by run_steps when both one-launch kernels take the batch)."""
        (x0, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh) = map(
            _c, (x0, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh))
        ops._need_cuda(x0, w_ext)
        N, C = x0.shape[0], channels
        S, H, HC, ld, E, dev = steps, heads, heads * channels, w_ext.shape[1], ea.shape[0], x0.device
        ctx.pre = None
        if w_pre is not None:
            assert fi is not None and p_drop == 0.0 and pn is None
            w_pre, b_pre = _c(w_pre), (None if b_pre is None else _c(b_pre))
            ctx.pre = (x0, b_pre is not None, pre_act)
        if fi is not None and p_drop == 0.0 and pn is None:
            # ONE launch for all steps (csrc/mp_fused.cu): x and h stay in shared memory from step to step; what backward
            # reads leaves the SM as tile-sized contiguous copies
            fused_bwd = (USE_FUSED_BWD and g.src_rowptr is not None and ops.message_stack_bwd_supported(channels, H, ea.shape[1], S))
            sv = _stack_buffers(x0, S, H, channels, ld, E, tiled_gates=fused_bwd)
            assert fused_bwd or w_pre is None
            ops.message_stack_fwd(x0, None, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh, g, fi, H, channels, S,
                                  slope, act, act_param, res, save=sv,
                                  pre=None if w_pre is None else (w_pre, b_pre, pre_act[0], pre_act[1]))
            X, HH = sv["X"], sv["HH"]
            ctx.save_for_backward(w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, ea, X, HH, X, None, sv["XPE"], sv["AGG"],
                                  sv["ALPHA"], sv.get("M"), sv.get("RZN"), sv.get("GH"), None)
            ctx.gt, ctx.mh = sv.get("GT"), sv.get("MH")
            ctx.g, ctx.cfg = g, (H, channels, slope, act, act_param, res, S, p_drop, None)
            ctx.fi = fi                                          # the same tile table serves the one-launch backward
            ctx.set_materialize_grads(False)
            return tuple(X[s + 1] for s in range(S)) + (HH[S],)
        new = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
        X, HH = new(S + 1, N, C), new(S + 1, N, C)               # block inputs / GRU states; [s+1] = outputs of step s
        X[0].copy_(x0)
        HH[0].copy_(x0)                                          # h = x.unsqueeze(0) on the first step (layer.py:254)
        drop = p_drop > 0.0
        XN = new(S, N, C) if pn is not None else X               # block inputs after PairNorm
        XD = new(S, N, C) if drop else XN                        # conv inputs (after dropout)
        MASK = new(S, N, C, dtype=torch.bool) if drop else None
        XPE, AGG, ALPHA = new(S, N, ld), new(S, N, HC), new(S, E, H)
        fused_gru = ops.gru_fused_supported(XPE[0], HH[0], C)        # then only the n-gate part of gh is kept: [N,C]
        M, RZN, GH = new(S, N, C), new(S, N, 3 * C), new(S, N, C if fused_gru else 3 * C)
        for s in range(S):
            if pn is not None:
                ops.pair_norm_fwd(X[s], pn[0], pn[1], pn[2], out=XN[s])
            if drop:
                torch.ops.aten.native_dropout.out(XN[s], p_drop, True, out0=XD[s], out1=MASK[s])
            ops.gemm(XD[s], w_ext, exact_cols=(HC, HC + 2 * H), out=XPE[s])
            ops.triplet_edge_fwd(XPE[s], ea, w_edge, att_edge, g, H, channels, slope, agg=AGG[s], alpha=ALPHA[s])
            ops.gemm(AGG[s], w_scale, bias=bias, epilogue=EPI_CELU, out=M[s])
            if fused_gru:
                ops.gru_fused_fwd(M[s], HH[s], X[s] if res else None, w_ih, w_hh, b_ih, b_hh, act, act_param,
                                  rzn=RZN[s], gh=GH[s], h_new=HH[s + 1], x_out=X[s + 1])
            else:
                ops.gemm(M[s], w_ih, transpose_w=True, bias=b_ih, out=RZN[s])
                ops.gemm(HH[s], w_hh, transpose_w=True, bias=b_hh, out=GH[s])
                ops.gru_gates_fwd(RZN[s], GH[s], HH[s], X[s] if res else None, act, act_param, h_new=HH[s + 1], x_out=X[s + 1])
        ctx.save_for_backward(w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, ea, X, HH, XD, MASK, XPE, AGG, ALPHA, M, RZN, GH,
                              None if pn is None else pn[0])
        ctx.g, ctx.cfg = g, (H, channels, slope, act, act_param, res, S, p_drop, None if pn is None else (pn[1], pn[2]))
        ctx.set_materialize_grads(False)
        return tuple(X[s + 1] for s in range(S)) + (HH[S],)      # every step's output (the pair models pool them) + final h

    @staticmethod
    def backward(ctx, *grads):
        (w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, ea, X, HH, XD, MASK, XPE, AGG, ALPHA, M, RZN, GH, pn_ptr) = ctx.saved_tensors
        H, C, slope, act, act_param, res, S, p_drop, pn = ctx.cfg
        g = ctx.g
        N, HC, ld, E, De, dev = X.shape[1], H * C, XPE.shape[2], ea.shape[0], ea.shape[1], X.device
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        G_PRE, G_XPE = new(S, N, C), new(S, N, ld)
        fi = getattr(ctx, "fi", None)
        gt = getattr(ctx, "gt", None)                            # forward saved the gate side tile-blocked: only the one-launch backward reads it
        fused_bwd = gt is not None or (USE_FUSED_BWD and fi is not None and p_drop == 0.0 and pn is None and g.src_rowptr is not None
                                       and ops.message_stack_bwd_supported(C, H, De, S))
        if fused_bwd:
            # ONE launch for the reverse loop (csrc/mp_fused_bwd.cu): gate backward, input-gradient projections and both edge
            # passes per tile, the carried gradients resident on the SM; what the weight-gradient contractions read comes out
            mh = getattr(ctx, "mh", None)
            G4 = new(S, N, 4 * C) if mh is not None else None
            G_GI, G_GH = (None, None) if mh is not None else (new(S, N, 3 * C), new(S, N, 3 * C))
            sv = dict(X=X, HH=HH, XPE=XPE, ALPHA=ALPHA, M=M, RZN=RZN, GH=GH, GT=gt)
            g_x0, g_w_edge, g_att_edge = ops.message_stack_bwd(sv, list(grads[:S]), _c(grads[S]), w_ext, w_edge, att_edge, w_scale,
                                                               w_ih, w_hh, g, fi, H, C, S, slope, act, act_param, res,
                                                               G_GI, G_GH, G_PRE, G_XPE, G4=G4,
                                                               pre_act=(ops.ACT_NONE, 0.0) if ctx.pre is None else ctx.pre[2])
        else:
            G_GI, G_GH, G4, mh = new(S, N, 3 * C), new(S, N, 3 * C), None, None
            G_LOGIT, G_WE = new(S, E, H), new(S, De, HC)
        g_ext, g_h, g_x = grads[:S], _c(grads[S]), None
        for s in (() if fused_bwd else range(S - 1, -1, -1)):
            if g_ext[s] is not None:                             # gradient arriving at this step's output from outside
                g_x = _c(g_ext[s]) if g_x is None else g_x.add_(g_ext[s])
            if g_x is None:
                g_x = torch.zeros((N, C), dtype=torch.float32, device=dev)
            _, _, g_h_prev, g_id = ops.gru_gates_bwd(RZN[s], GH[s], HH[s], X[s + 1], g_x, g_h, act, act_param, res,
                                                     g_gi=G_GI[s], g_gh=G_GH[s])
            ops.gemm(G_GI[s], w_ih, epilogue=EPI_MUL_CELU_GRAD, aux=M[s], out=G_PRE[s])       # d/d(pre-CELU message)
            ops.gemm(G_GH[s], w_hh, epilogue=EPI_ACCUM, out=g_h_prev)
            g_agg = ops.gemm(G_PRE[s], w_scale, transpose_w=True)                             # [N,HC]
            ops.triplet_edge_bwd(XPE[s], ea, w_edge, att_edge, ALPHA[s], g_agg, g, H, C, slope,
                                 g_xpe=G_XPE[s], g_logit=G_LOGIT[s], g_we=G_WE[s])
            if p_drop > 0.0 or pn is not None:
                g_x = ops.gemm(G_XPE[s], w_ext, transpose_w=True)                                 # d/d(conv input)
                if p_drop > 0.0:
                    g_x = torch.ops.aten.native_dropout_backward(g_x, MASK[s], 1.0 / (1.0 - p_drop))
                if pn is not None:                                                               # through PairNorm, onto the residual
                    g_x = ops.pair_norm_bwd(X[s], g_x, pn_ptr, pn[0], pn[1], out=g_id if res else None, accumulate=res)
                elif res:
                    g_x.add_(g_id)
            elif res:
                g_x = ops.gemm(G_XPE[s], w_ext, transpose_w=True, epilogue=EPI_ACCUM, out=g_id)
            else:
                g_x = ops.gemm(G_XPE[s], w_ext, transpose_w=True)
            g_h = g_h_prev
        if not fused_bwd:
            g_x0 = g_x.add_(g_h)                                                              # X[0] and HH[0] are both x0
        SN = S * N
        if G4 is not None:
            g_w_ih, g_w_hh, g_b_ih, g_b_hh = _gru_wgrads(mh.view(SN, 2 * C + 4), G4.view(SN, 4 * C), C)
        else:
            g_w_ih, g_b_ih = ops.gemm_tn_ex(M.view(SN, C), G_GI.view(SN, 3 * C), transpose_out=True, want_colsum=True)
            g_w_hh, g_b_hh = ops.gemm_tn_ex(HH[:S].view(SN, C), G_GH.view(SN, 3 * C), transpose_out=True, want_colsum=True)
        g_w_scale, g_bias = ops.gemm_tn_ex(AGG.view(SN, HC), G_PRE.view(SN, C), want_colsum=True)
        xd = XD[:S].view(SN, C)
        gxpe = G_XPE.view(SN, ld)
        g_w_ext, _ = ops.gemm_tn_ex(xd, gxpe)
        # the 2H logit columns are near-total cancellations (softmax gradients are zero-sum per destination): exact fp32
        ops.gemm_tn_ex(xd, gxpe[:, HC:HC + 2 * H], out=g_w_ext[:, HC:HC + 2 * H])
        if not fused_bwd:
            g_att_edge, _ = ops.gemm_tn_ex(ea, G_LOGIT.sum(0) if S > 1 else G_LOGIT[0])
            g_w_edge = G_WE.sum(0) if S > 1 else G_WE[0]
        g_w_pre = g_b_pre = None
        if getattr(ctx, "pre", None) is not None:
            # g_x0 left the kernel as the gradient of the input LinearBlock's pre-activation rows: its weight / bias gradients
            # are one skinny exact contraction with the raw features, which themselves need no gradient
            x_raw, has_bias, _ = ctx.pre
            g_w_pre, g_b_pre = ops.gemm_tn_ex(x_raw, g_x0, transpose_out=True, want_colsum=has_bias)
            g_x0 = None
        return (g_x0, g_w_ext, g_w_edge, g_att_edge, g_w_scale, g_bias, g_w_ih, g_w_hh, g_b_ih, g_b_hh,
                None, None, None, None, None, None, None, None, None, None, None, None, g_w_pre, g_b_pre, None)


# --------------------------------------------------------------------------------------------------
# PairNorm per graph (the reference's default graph_norm): deterministic warp-per-graph kernels
# --------------------------------------------------------------------------------------------------
class PairNormFn(Function):
    @staticmethod
    def forward(ctx, x, gptr, num_graphs, eps):
        x = _c(x)
        ops._need_cuda(x)
        ctx.save_for_backward(x, gptr)
        ctx.cfg = (num_graphs, eps)
        return ops.pair_norm_fwd(x, gptr, num_graphs, eps)

    @staticmethod
    def backward(ctx, g_y):
        x, gptr = ctx.saved_tensors
        num_graphs, eps = ctx.cfg
        return ops.pair_norm_bwd(x, _c(g_y), gptr, num_graphs, eps), None, None, None


# --------------------------------------------------------------------------------------------------
# small dense layer on graph-level rows (LSTM gates of Set2Set, nn of GlobalAttention)
# --------------------------------------------------------------------------------------------------
class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = map(_c, (x, weight, bias))
        ops._need_cuda(x, weight)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return ops.gemm(x, weight, transpose_w=True, bias=bias)

    @staticmethod
    def backward(ctx, g_y):
        x, weight = ctx.saved_tensors
        g_y = _c(g_y)
        g_w, g_b = ops.gemm_tn_ex(x, g_y, transpose_out=True, want_colsum=ctx.has_bias)       # (x^T g_y)^T = [N,K]
        g_x = ops.gemm(g_y, weight) if ctx.needs_input_grad[0] else None                      # raw features need none
        return g_x, g_w, g_b


class LSTMGatesFn(Function):
    """(gates_pre [B,4C], c_prev) -> (h_new, c_new); torch.nn.LSTM gate order i,f,g,o."""

    @staticmethod
    def forward(ctx, gates, c_prev):
        gates = gates.contiguous().clone()       # activated in place by the kernel
        c_prev = _c(c_prev)
        h_new, c_new = ops.lstm_gates_fwd(gates, c_prev)
        ctx.save_for_backward(gates, c_prev, c_new)
        ctx.set_materialize_grads(False)
        return h_new, c_new

    @staticmethod
    def backward(ctx, g_h, g_c):
        gates, c_prev, c_new = ctx.saved_tensors
        g_gates, g_c_prev = ops.lstm_gates_bwd(gates, c_prev, c_new, _c(g_h), _c(g_c))
        return g_gates, g_c_prev


# --------------------------------------------------------------------------------------------------
# per-graph attention pooling (GlobalAttention gate + pool; one Set2Set step)
# --------------------------------------------------------------------------------------------------
class SegAttnPoolFn(Function):
    """e[n] = <x[n], q[g]> (+ bias); a = softmax per graph (PyG form); returns (r[g] = sum a x, asum[g] = sum a).
    `q` is [B,C] (per graph) or [1,C] (shared, GlobalAttention's gate_nn weight)."""

    @staticmethod
    def forward(ctx, x, q, q_bias, gptr, num_graphs):
        x, q, q_bias = map(_c, (x, q, q_bias))
        ops._need_cuda(x, q)
        shared = q.shape[0] == 1 and num_graphs != 1
        stride = 0 if (shared or q.shape[0] == 1) else q.shape[1]
        a, r, asum = ops.seg_attn_pool_fwd(x, q, stride, q_bias, gptr, num_graphs)
        ctx.save_for_backward(x, q, a, gptr)
        ctx.cfg = (num_graphs, stride, q.shape[0], q_bias is not None)
        ctx.set_materialize_grads(False)
        return r, asum

    @staticmethod
    def backward(ctx, g_r, g_asum):
        x, q, a, gptr = ctx.saved_tensors
        num_graphs, stride, q_rows, has_bias = ctx.cfg
        if g_r is None:
            g_r = torch.zeros((num_graphs, x.shape[1]), dtype=x.dtype, device=x.device)
        g_x = torch.empty_like(x)
        g_q, g_e = ops.seg_attn_pool_bwd(x, q, stride, a, _c(g_r), _c(g_asum), gptr, num_graphs, g_x, False)
        if q_rows == 1 and num_graphs != 1:
            g_q = ops.colsum(g_q).view(1, -1)
        g_b = ops.colsum(g_e.view(-1, 1)) if has_bias else None
        return g_x, g_q, g_b, None, None


class Set2SetFn(Function):
    """PyG Set2Set: (x, lstm weights) -> q* [B,2C].  Per round one tensor-core GEMM for the gates of all graphs and one
    warp-per-graph kernel for cell + attention + pooled read (csrc/set2set.cu)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, gptr, num_graphs, steps):
        x, w_ih, w_hh, b_ih, b_hh = map(_c, (x, w_ih, w_hh, b_ih, b_hh))
        ops._need_cuda(x, w_ih)
        N, C = x.shape
        B, S, dev = num_graphs, steps, x.device
        w_cat = torch.cat([w_ih, w_hh], 1)                                  # [4C,3C] against u = [h | r | h]
        b_sum = b_ih + b_hh
        U = torch.empty((S + 1, B, 3 * C), dtype=torch.float32, device=dev)
        cs = torch.empty((S + 1, B, C), dtype=torch.float32, device=dev)
        gates = torch.empty((S, B, 4 * C), dtype=torch.float32, device=dev)
        att = torch.empty((S, N), dtype=torch.float32, device=dev)
        out = torch.empty((B, 2 * C), dtype=torch.float32, device=dev)
        U[0].zero_()
        cs[0].zero_()
        for s in range(S):
            if s == 0:
                gates[0].copy_(b_sum)                                       # u_0 = 0: the gate pre-activations of the first round ARE the bias
            else:
                ops.gemm(U[s], w_cat, transpose_w=True, bias=b_sum, out=gates[s])
            ops.set2set_round_fwd(x, gates[s], cs[s], cs[s + 1], gptr, B, att[s], U[s + 1], out if s == S - 1 else None)
        ctx.save_for_backward(x, w_cat, gptr, U, cs, gates, att)
        ctx.cfg = (B, S)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, w_cat, gptr, U, cs, gates, att = ctx.saved_tensors
        B, S = ctx.cfg
        N, C = x.shape
        dev = x.device
        G = torch.empty((S, B, 4 * C), dtype=torch.float32, device=dev)
        g_c = torch.zeros((B, C), dtype=torch.float32, device=dev)
        g_x = torch.empty((N, C), dtype=torch.float32, device=dev)
        g_u = _c(g_out)
        for s in range(S - 1, -1, -1):
            ops.set2set_round_bwd(x, gates[s], cs[s], cs[s + 1], att[s], gptr, B, g_u, g_c, g_x, s != S - 1, G[s])
            if s > 0:
                g_u = ops.gemm(G[s], w_cat)                                 # [B,3C]
        g_w, g_b = ops.gemm_tn_ex(U[:S].reshape(S * B, 3 * C), G.view(S * B, 4 * C), transpose_out=True, want_colsum=True)
        return g_x, g_w[:, :2 * C], g_w[:, 2 * C:], g_b, g_b, None, None, None


# --------------------------------------------------------------------------------------------------
# cross-graph dot pool (src_2gi_ddi/layer.py:270-283)
# --------------------------------------------------------------------------------------------------
class PairDotPoolFn(Function):
    @staticmethod
    def forward(ctx, xa, xb, ptr_a, ptr_b, num_pairs, idx_b=None):
        xa, xb = _c(xa), _c(xb)
        ops._need_cuda(xa, xb)
        if idx_b is not None and (xa.requires_grad or xb.requires_grad):
            raise ops._lib.GlamError("dot_and_global_pool2 with a shared protein index (pro_index) is forward-only (evaluation / screening)")
        out, argmax, sa, sb = ops.pair_dot_pool_fwd(xa, xb, ptr_a, ptr_b, num_pairs, idx_b)
        ctx.save_for_backward(xa, xb, ptr_a, ptr_b, argmax, sa, sb)
        ctx.num_pairs = num_pairs
        return out

    @staticmethod
    def backward(ctx, g_out):
        xa, xb, ptr_a, ptr_b, argmax, sa, sb = ctx.saved_tensors
        g_xa, g_xb = ops.pair_dot_pool_bwd(xa, xb, ptr_a, ptr_b, _c(g_out), argmax, sa, sb, ctx.num_pairs)
        return g_xa, g_xb, None, None, None, None


# --------------------------------------------------------------------------------------------------
# parameter-space preparation (derived weights of the triplet layers)
# --------------------------------------------------------------------------------------------------
class TripletPrepFn(Function):
    """(weight_node, weight_edge|None, weight_triplet_att) -> (w_ext [C,ldxp], att_edge [De,H])."""

    @staticmethod
    def forward(ctx, weight_node, weight_edge, att, channels, heads, edge_dim, ldxp):
        light = weight_edge is None
        weight_node, weight_edge, att = map(_c, (weight_node, weight_edge, att))
        ops._need_cuda(weight_node, att)
        w_ext, att_edge = ops.triplet_prep_fwd(weight_node, weight_edge, att, channels, heads, edge_dim, light, ldxp)
        ctx.save_for_backward(weight_node, weight_edge, att)
        ctx.cfg = (channels, heads, edge_dim, light, ldxp)
        ctx.set_materialize_grads(False)
        return w_ext, att_edge

    @staticmethod
    def backward(ctx, g_w_ext, g_att_edge):
        weight_node, weight_edge, att = ctx.saved_tensors
        channels, heads, edge_dim, light, ldxp = ctx.cfg
        if g_w_ext is None:
            g_w_ext = torch.zeros((channels, ldxp), dtype=torch.float32, device=att.device)
        if g_att_edge is None:
            g_att_edge = torch.zeros((edge_dim, heads), dtype=torch.float32, device=att.device)
        g_wn, g_we, g_att = ops.triplet_prep_bwd(weight_node, weight_edge, att, _c(g_w_ext), _c(g_att_edge), None,
                                                 channels, heads, edge_dim, light, ldxp)
        return g_wn, g_we, g_att, None, None, None, None


# --------------------------------------------------------------------------------------------------
# "next" rows: GlobalPool5 readout, GCN tower
# --------------------------------------------------------------------------------------------------
class Pool5Fn(Function):
    """[mean | sum | sort-pool(k=3)] per graph (src_1gp/layer.py:197-203)."""

    @staticmethod
    def forward(ctx, x, gptr, num_graphs):
        x = _c(x)
        ops._need_cuda(x)
        out, top = ops.pool5_fwd(x, gptr, num_graphs)
        ctx.save_for_backward(gptr, top)
        ctx.cfg = (num_graphs, x.shape[0], x.shape[1])
        return out

    @staticmethod
    def backward(ctx, g_out):
        gptr, top = ctx.saved_tensors
        B, N, C = ctx.cfg
        return ops.pool5_bwd(_c(g_out), gptr, top, B, N, C), None, None


class GCNConvFn(Function):
    """PyG GCNConv(in,out) @1.7.2: D^-1/2 (A + I) D^-1/2 (x W) + b over the cached graph index; backward aggregates over the
    source-sorted CSR (the transposed normalised adjacency) — no atomics."""

    @staticmethod
    def forward(ctx, x, weight, bias, g, norm):
        x, weight, bias = map(_c, (x, weight, bias))
        ops._need_cuda(x, weight)
        dinv2, w_dst, w_src = norm
        xw = ops.gemm(x, weight)                                                   # [N,out]
        out = ops.csr_aggregate(xw, g.dst_rowptr, g.dst_src, edge_w=w_dst, self_w=dinv2, bias=bias)
        ctx.save_for_backward(x, weight, dinv2, w_src)
        ctx.g = g
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, weight, dinv2, w_src = ctx.saved_tensors
        g = ctx.g
        g_out = _c(g_out)
        g_xw = ops.csr_aggregate(g_out, g.src_rowptr, g.src_dst, edge_w=w_src, self_w=dinv2)
        g_w, _ = ops.gemm_tn_ex(x, g_xw)                                           # [in,out]
        g_b = ops.colsum(g_out) if ctx.has_bias else None
        g_x = ops.gemm(g_xw, weight, transpose_w=True) if ctx.needs_input_grad[0] else None
        return g_x, g_w, g_b, None, None


class NNConvFn(Function):
    """PyG NNConv(aggr='mean') for one-hot bond features (`_NNConv`, src_1gp/layer.py:115-122): the per-edge [C,C] matrix
    nn(edge_attr) takes only edge_dim distinct values, so the messages are rows of ONE grouped projection
    Y = x @ [Theta_0 | .. | Theta_{De-1}] (a [N,C]x[C,De*C] tensor-core GEMM) and the layer is a typed gather-mean over the
    dst CSR plus the root term; backward scatters through a CSR grouped by (source, type).  No [E,C,C] tensor, no atomics."""

    @staticmethod
    def forward(ctx, x, theta_cat, root, bias, g, idx):
        x, theta_cat, root, bias = map(_c, (x, theta_cat, root, bias))
        ops._need_cuda(x, theta_cat)
        col_fwd, inv_deg, rowptr_r, col_r, w_r, De = idx
        N, C = x.shape
        Co = root.shape[1]
        y = ops.gemm(x, theta_cat)                                                 # [N, De*Co]
        out = ops.csr_aggregate(y.view(N * De, Co), g.dst_rowptr, col_fwd, row_scale=inv_deg, num_rows=N)
        ops.gemm(x, root, bias=bias, epilogue=EPI_ACCUM, out=out)
        ctx.save_for_backward(x, theta_cat, root, rowptr_r, col_r, w_r)
        ctx.cfg = (De, Co, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, theta_cat, root, rowptr_r, col_r, w_r = ctx.saved_tensors
        De, Co, has_bias = ctx.cfg
        g_out = _c(g_out)
        N = x.shape[0]
        g_y = ops.csr_aggregate(g_out, rowptr_r, col_r, edge_w=w_r, num_rows=N * De).view(N, De * Co)
        g_theta, _ = ops.gemm_tn_ex(x, g_y)                                        # [C, De*Co]
        g_root, g_bias = ops.gemm_tn_ex(x, g_out, want_colsum=has_bias)
        g_x = None
        if ctx.needs_input_grad[0]:
            g_x = ops.gemm(g_y, theta_cat, transpose_w=True)
            ops.gemm(g_out, root, transpose_w=True, epilogue=EPI_ACCUM, out=g_x)
        return g_x, g_theta, g_root, g_bias, None, None
