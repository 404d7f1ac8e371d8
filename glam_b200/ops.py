"""Tensor-level wrappers over the C ABI (include/glam_b200.h).

PyTorch is used here only for device memory (torch.empty through the caching allocator) and for the
current CUDA stream; every computation below is a call into libglam_b200.so.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_NONE, EPI_CELU, EPI_MUL_CELU_GRAD, EPI_ACCUM = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_CELU = 0, 1, 2, 3


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_call_dev = None     # device of the stream handed to the next library call (see _call)


def _stream(t: torch.Tensor):
    global _call_dev
    _call_dev = t.device
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.GlamError("glam_b200 kernels need CUDA tensors: there is no CPU fallback "
                                 f"(got a tensor on {t.device})")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise _lib.GlamError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_profile = None     # when a list: (name, start_event, end_event) per C-ABI call (bench.py's per-kernel timing)


def set_profile(sink):
    """Enable (sink = list) or disable (None) CUDA-event timing of every library call on the current stream."""
    global _profile
    _profile = sink


def _call(name: str, *args, label: str = ""):
    fn = getattr(_lib.load(), name)
    dev = _call_dev
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        # tensors on a device that is not the current one (one process driving several GPUs): the launch, the per-device
        # function attributes and the workspace queries must see THAT device
        with torch.cuda.device(dev):
            return _call_on_device(fn, name, args, label)
    return _call_on_device(fn, name, args, label)


def _call_on_device(fn, name, args, label):
    if _profile is None:
        _lib.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args), name)
    e1.record()
    _profile.append((name + label, e0, e1))


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------ CSR
def build_csr(edge_index: torch.Tensor, num_nodes: int):
    _need_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise _lib.GlamError(f"edge_index must be int64 [2,E], got {edge_index.dtype} {tuple(edge_index.shape)}")
    edge_index = edge_index.contiguous()
    lib = _lib.load()
    E, N, dev = edge_index.shape[1], int(num_nodes), edge_index.device
    i32 = dict(dtype=torch.int32, device=dev)
    out = {
        "dst_rowptr": torch.empty(N + 1, **i32), "dst_src": torch.empty(E, **i32), "dst_perm": torch.empty(E, **i32),
        "dst_dst": torch.empty(E, **i32),
        "src_rowptr": torch.empty(N + 1, **i32), "src_pos": torch.empty(E, **i32), "src_dst": torch.empty(E, **i32),
    }
    nbytes = lib.glam_csr_workspace_bytes(N, E)
    ws = _ws(nbytes, dev)
    _call("glam_build_csr", _p(edge_index), E, N, _p(out["dst_rowptr"]), _p(out["dst_src"]), _p(out["dst_perm"]),
          _p(out["dst_dst"]), _p(out["src_rowptr"]), _p(out["src_pos"]), _p(out["src_dst"]), _p(ws), ws.numel(),
                                  _stream(edge_index))
    return out


# windowed (bulk-copy staged) edge kernels; False = per-edge gather kernels (kept for A/B timing and as the tested fallback)
USE_EDGE_TILES = True


def build_edge_tiles(csr: dict, num_nodes: int):
    """(dst_tiles, src_tiles) int32 [T,4]: tile descriptors of the windowed edge kernels (include/glam_b200.h)."""
    lib = _lib.load()
    N = int(num_nodes)
    dev = csr["dst_rowptr"].device
    T = int(lib.glam_edge_tile_count(N))
    dst_tiles = torch.empty((max(T, 1), 4), dtype=torch.int32, device=dev)
    src_tiles = torch.empty((max(T, 1), 4), dtype=torch.int32, device=dev)
    E = csr["dst_src"].shape[0]
    _call("glam_build_edge_tiles", _p(csr["dst_rowptr"]), _p(csr["dst_src"]), _p(csr["src_rowptr"]), _p(csr["src_dst"]), N, E,
          _p(dst_tiles), _p(src_tiles), _stream(dst_tiles))
    return dst_tiles, src_tiles


def graph_ptr(batch: torch.Tensor, num_graphs: int) -> torch.Tensor:
    _need_cuda(batch)
    if batch.dtype != torch.int64 or batch.dim() != 1:
        raise _lib.GlamError("batch must be int64 [N]")
    batch = batch.contiguous()
    out = torch.empty(int(num_graphs) + 1, dtype=torch.int32, device=batch.device)
    _call("glam_graph_ptr", _p(batch), batch.shape[0], int(num_graphs), _p(out), _stream(batch))
    return out


def gather_rows(src: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    _need_cuda(src, perm)
    src = _f32c(src, "gather_rows")
    src2 = src.view(src.shape[0], -1) if src.dim() != 2 else src
    out = torch.empty((perm.shape[0], src2.shape[1]), dtype=torch.float32, device=src.device)
    _call("glam_gather_rows", _p(src2), _p(perm), perm.shape[0], src2.shape[1], _p(out), _stream(src))
    return out


# ------------------------------------------------------------------------------------------------ GEMMs
def gemm(X: torch.Tensor, W: torch.Tensor, transpose_w: bool = False, bias: Optional[torch.Tensor] = None,
         epilogue: int = EPI_NONE, aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         ldy: Optional[int] = None, exact_cols=(0, 0)) -> torch.Tensor:
    """Y = epilogue(X @ W + bias) (W is [K,N]) or X @ W.T (W is [N,K], transpose_w=True).
    X may be a row-strided 2-D view (stride(1) == 1)."""
    _need_cuda(X, W)
    assert X.dim() == 2 and W.dim() == 2 and X.stride(1) == 1 and W.is_contiguous()
    M, K = X.shape
    if transpose_w:
        N, Kw = W.shape
        w_sk, w_sn = 1, W.stride(0)
    else:
        Kw, N = W.shape
        w_sk, w_sn = W.stride(0), 1
    assert Kw == K, f"gemm: K mismatch {Kw} vs {K}"
    if out is None:
        ld = N if ldy is None else ldy
        out = torch.empty((M, ld), dtype=torch.float32, device=X.device)
    ld = out.stride(0)
    if aux is not None:
        assert aux.stride(1) == 1
    _call("glam_gemm_ex", _p(X), X.stride(0), _p(W), w_sk, w_sn, _p(bias), _p(aux),
          0 if aux is None else aux.stride(0), _p(out), ld, M, N, K, epilogue, int(exact_cols[0]), int(exact_cols[1]), _stream(X),
          label=f"[M={M},N={N},K={K},{'nt' if transpose_w else 'nn'},epi={epilogue}]")
    return out


def gemm_tn(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """A^T @ B with a fixed-order reduction over the (long) row dimension: [Ka,Kb]."""
    _need_cuda(A, B)
    assert A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0] and A.stride(1) == 1 and B.stride(1) == 1
    M, Ka = A.shape
    Kb = B.shape[1]
    lib = _lib.load()
    out = torch.empty((Ka, Kb), dtype=torch.float32, device=A.device)
    ws = _ws(lib.glam_gemm_tn_workspace_bytes(M, Ka, Kb), A.device)
    _call("glam_gemm_tn", _p(A), A.stride(0), _p(B), B.stride(0), M, Ka, Kb, _p(out), Kb, _p(ws), ws.numel(),
          _stream(A), label=f"[M={M},Ka={Ka},Kb={Kb}]")
    return out


def gemm_tn_ex(A: torch.Tensor, B: torch.Tensor, transpose_out: bool = False, want_colsum: bool = False,
               out: Optional[torch.Tensor] = None):
    """(A^T @ B  [Ka,Kb] — or its transpose [Kb,Ka] —, colsum(B) [Kb] | None), reduction over the long row
    dimension in a fixed order; tensor cores in TF32 mode."""
    _need_cuda(A, B)
    assert A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0] and A.stride(1) == 1 and B.stride(1) == 1
    M, Ka = A.shape
    Kb = B.shape[1]
    lib = _lib.load()
    if out is None:
        out = torch.empty((Kb, Ka) if transpose_out else (Ka, Kb), dtype=torch.float32, device=A.device)
    assert out.stride(1) == 1 and tuple(out.shape) == ((Kb, Ka) if transpose_out else (Ka, Kb))
    cs = torch.empty((Kb,), dtype=torch.float32, device=A.device) if want_colsum else None
    ws = _ws(lib.glam_gemm_tn_ex_workspace_bytes(M, Ka, Kb, 1 if want_colsum else 0), A.device)
    _call("glam_gemm_tn_ex", _p(A), A.stride(0), _p(B), B.stride(0), M, Ka, Kb, _p(out), out.stride(0),
          1 if transpose_out else 0, _p(cs), _p(ws), ws.numel(), _stream(A), label=f"[M={M},Ka={Ka},Kb={Kb}]")
    return out, cs


def colsum(G: torch.Tensor) -> torch.Tensor:
    _need_cuda(G)
    assert G.dim() == 2 and G.stride(1) == 1
    M, N = G.shape
    lib = _lib.load()
    out = torch.empty((N,), dtype=torch.float32, device=G.device)
    ws = _ws(lib.glam_colsum_workspace_bytes(M, N), G.device)
    _call("glam_colsum", _p(G), G.stride(0), M, N, _p(out), _p(ws), ws.numel(), _stream(G), label=f"[M={M},N={N}]")
    return out


# ------------------------------------------------------------------------------------------------ edge phase
def triplet_edge_fwd(xpe, ea_sorted, w_edge, att_edge, g, heads, channels, slope, agg=None, alpha=None):
    N, E, dev = xpe.shape[0], ea_sorted.shape[0], xpe.device
    HC = heads * channels
    agg = torch.empty((N, HC), dtype=torch.float32, device=dev) if agg is None else agg
    alpha = torch.empty((E, heads), dtype=torch.float32, device=dev) if alpha is None else alpha
    assert agg.is_contiguous() and alpha.is_contiguous()
    _call("glam_triplet_edge_fwd", _p(xpe), xpe.stride(0), _p(ea_sorted), _p(w_edge), _p(att_edge),
                                                 _p(g.dst_rowptr), _p(g.dst_src), _p(g.dst_tiles if USE_EDGE_TILES else None),
                                                 N, E, heads, channels,
                                                 ea_sorted.shape[1], float(slope), _p(agg), _p(alpha), _stream(xpe))
    return agg, alpha


def triplet_edge_bwd(xpe, ea_sorted, w_edge, att_edge, alpha, g_agg, g, heads, channels, slope,
                     g_xpe=None, g_logit=None, g_we=None):
    """Returns g_xpe [N, ldxp], g_logit [E,H] (dst order), g_w_edge [De,HC] or None."""
    lib = _lib.load()
    N, E, dev = xpe.shape[0], ea_sorted.shape[0], xpe.device
    De, HC, ld = ea_sorted.shape[1], heads * channels, xpe.stride(0)
    g_xpe = torch.empty((N, ld), dtype=torch.float32, device=dev) if g_xpe is None else g_xpe
    g_logit = torch.empty((E, heads), dtype=torch.float32, device=dev) if g_logit is None else g_logit
    assert g_xpe.is_contiguous() and g_logit.is_contiguous() and g_xpe.shape[1] == ld
    ws = None
    if w_edge is not None:
        g_we = torch.empty((De, HC), dtype=torch.float32, device=dev) if g_we is None else g_we
        assert g_we.is_contiguous()
        ws = _ws(lib.glam_triplet_bwd_workspace_bytes(heads, channels, De), dev)
    st = _stream(xpe)
    _call("glam_triplet_edge_bwd_dst", _p(xpe), ld, _p(ea_sorted), _p(w_edge), _p(att_edge), _p(alpha), _p(g_agg),
                                             _p(g.dst_rowptr), _p(g.dst_src), _p(g.dst_dst),
                                             _p(g.dst_tiles if USE_EDGE_TILES else None), N, E, heads, channels, De, float(slope),
                                             _p(g_logit), _p(g_xpe), _p(g_we), _p(ws), 0 if ws is None else ws.numel(),
                                             st)
    _call("glam_triplet_edge_bwd_src", _p(ea_sorted), _p(w_edge), _p(alpha), _p(g_agg), _p(g_logit),
                                             _p(g.src_rowptr), _p(g.src_pos), _p(g.src_dst),
                                             _p(g.src_tiles if USE_EDGE_TILES else None), N, E, heads, channels, De,
                                             _p(g_xpe), ld, st)
    return g_xpe, g_logit, g_we


def triplet_prep_fwd(weight_node, weight_edge, att, channels, heads, edge_dim, light, ldxp):
    dev = weight_node.device
    w_ext = torch.empty((channels, ldxp), dtype=torch.float32, device=dev)
    att_edge = torch.empty((edge_dim, heads), dtype=torch.float32, device=dev)
    _call("glam_triplet_prep_fwd", _p(weight_node), _p(weight_edge), _p(att), channels, heads, edge_dim,
                                                 1 if light else 0, ldxp, _p(w_ext), _p(att_edge), _stream(weight_node))
    return w_ext, att_edge


def triplet_prep_bwd(weight_node, weight_edge, att, g_w_ext, g_att_edge, g_w_edge_direct, channels, heads, edge_dim,
                     light, ldxp):
    g_wn = torch.empty_like(weight_node)
    g_we = None if light else torch.empty_like(weight_edge)
    g_att = torch.empty_like(att)
    _call("glam_triplet_prep_bwd", _p(weight_node), _p(weight_edge), _p(att), _p(g_w_ext), _p(g_att_edge),
                                                 _p(g_w_edge_direct), channels, heads, edge_dim, 1 if light else 0, ldxp,
                                                 _p(g_wn), _p(g_we), _p(g_att), _stream(weight_node))
    return g_wn, g_we, g_att


# ------------------------------------------------------------------------------------------------ fused message stack
def graph_tile_caps():
    import ctypes as C
    a, b = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().glam_graph_tile_caps(C.addressof(a), C.addressof(b)), "glam_graph_tile_caps")
    return a.value, b.value


def build_graph_tiles(gptr: torch.Tensor, num_graphs: int, g, meta: torch.Tensor, check_edges: bool = True) -> torch.Tensor:
    """Graph-aligned tiles {n0, n1, e0, e1} (int32 [B,4]; meta[0] of them are valid) for the fused message kernel."""
    _need_cuda(gptr)
    B = int(num_graphs)
    tiles = torch.empty((max((B + 63) // 64 * 64, 64), 4), dtype=torch.int32, device=gptr.device)
    ws = _ws(_lib.load().glam_graph_tiles_workspace_bytes(B), gptr.device)
    _call("glam_build_graph_tiles", _p(gptr), B, _p(g.dst_rowptr), _p(g.dst_src if check_edges else None), g.num_nodes, g.num_edges,
          _p(tiles), _p(meta), _p(ws), ws.numel(), _stream(gptr))
    return tiles


def edge_types(ea: torch.Tensor, meta: torch.Tensor, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Bond type per dst-ordered edge.  `ea` is edge_attr already in dst order, or — with perm = the index's dst_perm — in the
    caller's edge order (the permuted copy is then never made)."""
    E, De = ea.shape
    assert ea.is_contiguous() and ea.dtype == torch.float32 and (perm is None or (perm.dtype == torch.int32 and perm.numel() == E))
    et = torch.empty((max(E, 1),), dtype=torch.uint8, device=ea.device)
    _call("glam_edge_types", _p(ea), _p(perm), E, De, _p(et), _p(meta), _stream(ea))
    return et


def message_stack_supported(channels: int, heads: int, edge_dim: int) -> bool:
    return bool(_lib.load().glam_message_stack_supported(int(channels), int(heads), int(edge_dim)))


def message_stack_fwd(x0, h0, w_ext, w_edge, att_edge, w_scale, bias, w_ih, w_hh, b_ih, b_hh, g, fi, heads, channels, steps,
                      slope, act, act_param, res, conv_only=False, keep_all=False, save=None, pre=None, pn=None):
    """The whole message stack in one launch (csrc/mp_fused.cu).  Eval (save=None): returns (x_out [S|1,N,C], h_out [N,C]|None).
    Training: `save` = dict of preallocated stacked tensors X, HH, XPE, AGG, ALPHA, M, RZN, GH (conv_only: XPE, AGG, ALPHA) that
    the kernel fills; returns (x_out|None, None)."""
    _need_cuda(x0, w_ext)
    N, C = x0.shape[0], channels
    E, dev = g.num_edges, x0.device
    assert x0.is_contiguous() and (h0 is None or h0.is_contiguous()) and w_ext.is_contiguous()
    # pre = (weight [C, raw_dim], bias [C] | None, act code, act param): the model's input LinearBlock, applied in the kernel
    # to x0 = the RAW features [N, raw_dim]
    x_in, x_raw, raw_dim, w_pre, b_pre, pre_act, pre_par = x0, None, 0, None, None, 0, 0.0
    if pre is not None:
        w_pre, b_pre, pre_act, pre_par = pre
        assert w_pre.is_contiguous() and w_pre.shape == (C, x0.shape[1]) and (b_pre is None or b_pre.is_contiguous())
        x_in, x_raw, raw_dim = None, x0, x0.shape[1]
    else:
        assert x0.shape[1] == C
    # pn = (batch int64 [N], eps): PairNorm on every step's block input inside the kernel (evaluation only)
    pn_batch, pn_eps = (None, 0.0) if pn is None else pn
    if pn_batch is not None:
        assert save is None and not conv_only and h0 is None and pn_batch.dtype == torch.int64 and pn_batch.is_contiguous() and pn_batch.numel() == N
    x_out = h_out = None
    if save is None or conv_only:
        x_out = torch.empty(((steps if keep_all else 1), N, C), dtype=torch.float32, device=dev)
    if save is None and not conv_only:
        h_out = torch.empty((N, C), dtype=torch.float32, device=dev)
    sv = save or {}
    for k, t in sv.items():
        assert t.is_contiguous(), k
    _call("glam_message_stack_fwd", _p(x_in), _p(h0), _p(x_raw), raw_dim, _p(w_pre), _p(b_pre), pre_act, float(pre_par), _p(w_ext), w_ext.stride(0), _p(w_edge), _p(att_edge), _p(w_scale), _p(bias),
          _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh), _p(fi.tiles), _p(fi.meta), _p(g.dst_rowptr), _p(g.dst_src), _p(fi.etype),
          N, E, channels, heads, fi.edge_dim, steps, float(slope), act, float(act_param), 1 if res else 0,
          1 if conv_only else 0, 1 if keep_all else 0, _p(x_out), _p(h_out), _p(sv.get("X")), _p(sv.get("HH")), _p(sv.get("XPE")),
          _p(sv.get("AGG")), _p(sv.get("ALPHA")), _p(sv.get("M")), _p(sv.get("RZN")), _p(sv.get("GH")), _p(sv.get("GT")),
          _p(sv.get("MH")), _p(pn_batch), float(pn_eps), _stream(x0),
          label=f"[N={N},S={steps},{'save' if save is not None else 'eval'}{',conv' if conv_only else ''}]")
    return x_out, h_out


def message_stack_bwd_supported(channels: int, heads: int, edge_dim: int, steps: int) -> bool:
    return bool(_lib.load().glam_message_stack_bwd_supported(int(channels), int(heads), int(edge_dim), int(steps)))


def message_stack_bwd(sv, g_ext, g_h_final, w_ext, w_edge, att_edge, w_scale, w_ih, w_hh, g, fi, heads, channels, steps, slope, act,
                      act_param, res, G_GI, G_GH, G_PRE, G_XPE, separate_h0=False, G4=None, pre_act=(ACT_NONE, 0.0)):
    """Backward of the whole message stack in one launch (csrc/mp_fused_bwd.cu).  `sv` = the tensors message_stack_fwd saved;
    `g_ext` = per-step gradients of the step outputs (None entries allowed), `g_h_final` = gradient of the final GRU state.
    Fills G_GI, G_GH [S,N,3C], G_PRE [S,N,C], G_XPE [S,N,ld]; returns (g_x0 [N,C], g_w_edge [De,HC], g_att_edge [De,H]) — with
    separate_h0 (the forward had its own h0 tensor) g_x0 is the pair (g_x0, g_h0).  pre_act = (code, param) of the input
    LinearBlock the forward applied in the kernel: g_x0 is then the gradient of that block's pre-activation rows."""
    import ctypes
    X = sv["X"]
    _need_cuda(X, w_ext)
    N, C, E, dev = X.shape[1], channels, g.num_edges, X.device
    De, HC = fi.edge_dim, heads * channels
    for t in (X, sv.get("HH"), sv["XPE"], sv["ALPHA"], sv.get("M"), sv.get("RZN"), sv.get("GH"), sv.get("GT"), G_GI, G_GH, G4, G_PRE, G_XPE,
              w_ext, w_edge, att_edge, w_scale, w_ih, w_hh):
        assert t is None or t.is_contiguous()
    g_ext = [None if t is None else (t if t.is_contiguous() else t.contiguous()) for t in g_ext]
    assert len(g_ext) == steps
    ptrs = (ctypes.c_void_p * steps)(*[_p(t) for t in g_ext])
    g_x0 = torch.empty((N, C), dtype=torch.float32, device=dev)
    g_h0 = torch.empty((N, C), dtype=torch.float32, device=dev) if separate_h0 else None
    g_we = torch.empty((De, HC), dtype=torch.float32, device=dev)
    g_ae = torch.empty((De, heads), dtype=torch.float32, device=dev)
    ws = _ws(_lib.load().glam_message_stack_bwd_workspace_bytes(channels, heads, De), dev)
    _call("glam_message_stack_bwd", _p(X), _p(sv.get("HH")), _p(sv["XPE"]), _p(sv["ALPHA"]), _p(sv.get("M")), _p(sv.get("RZN")),
          _p(sv.get("GH")), _p(sv.get("GT")), ctypes.cast(ptrs, ctypes.c_void_p), _p(g_h_final), _p(w_ext), w_ext.stride(0), _p(w_edge), _p(att_edge), _p(w_scale),
          _p(w_ih), _p(w_hh), _p(fi.tiles), _p(fi.meta), _p(g.dst_rowptr), _p(g.dst_src), _p(fi.etype), _p(g.src_rowptr),
          _p(g.src_pos), _p(g.src_dst), N, E, channels, heads, De, steps, float(slope), act, float(act_param), 1 if res else 0,
          int(pre_act[0]), float(pre_act[1]), _p(G_GI), _p(G_GH), _p(G4), _p(G_PRE), _p(G_XPE), _p(g_x0), _p(g_h0), _p(g_we), _p(g_ae), _p(ws), ws.numel(), _stream(X),
          label=f"[N={N},S={steps}]")
    return ((g_x0, g_h0) if separate_h0 else g_x0), g_we, g_ae


# ------------------------------------------------------------------------------------------------ cells
def gru_gates_fwd(gi, gh, h, identity, act, act_param, h_new=None, x_out=None):
    N, C = h.shape
    h_new = torch.empty_like(h) if h_new is None else h_new
    x_out = torch.empty_like(h) if x_out is None else x_out
    assert h_new.is_contiguous() and x_out.is_contiguous()
    _call("glam_gru_gates_fwd", _p(gi), _p(gh), _p(h), _p(identity), N, C, act, float(act_param),
                                              _p(h_new), _p(x_out), _stream(h))
    return h_new, x_out


def gru_fused_supported(m, h, channels) -> bool:
    return (bool(_lib.load().glam_gru_fused_supported(int(channels))) and m.stride(0) % 4 == 0 and h.stride(0) % 4 == 0
            and m.data_ptr() % 16 == 0 and h.data_ptr() % 16 == 0)


def gru_fused_fwd(m, h, identity, w_ih, w_hh, b_ih, b_hh, act, act_param, rzn=None, gh=None, h_new=None, x_out=None):
    """gi/gh GEMMs + gates in one tensor-core kernel; returns (rzn [N,3C], gh_n [N,C], h_new, x_out)."""
    _need_cuda(m, h)
    N, C = h.shape
    dev = h.device
    rzn = torch.empty((N, 3 * C), dtype=torch.float32, device=dev) if rzn is None else rzn
    gh = torch.empty((N, C), dtype=torch.float32, device=dev) if gh is None else gh
    assert tuple(gh.shape) == (N, C)
    h_new = torch.empty((N, C), dtype=torch.float32, device=dev) if h_new is None else h_new
    x_out = torch.empty((N, C), dtype=torch.float32, device=dev) if x_out is None else x_out
    assert rzn.is_contiguous() and gh.is_contiguous() and h_new.is_contiguous() and x_out.is_contiguous()
    assert w_ih.is_contiguous() and w_hh.is_contiguous() and m.stride(1) == 1 and h.stride(1) == 1
    assert identity is None or identity.is_contiguous()
    _call("glam_gru_fused_fwd", _p(m), m.stride(0), _p(h), h.stride(0), _p(identity), _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh),
          N, C, act, float(act_param), _p(rzn), _p(gh), _p(h_new), _p(x_out), _stream(h))
    return rzn, gh, h_new, x_out


def gru_gates_bwd(rzn, gh, h, x_out, g_x_out, g_h_carry, act, act_param, want_identity, g_gi=None, g_gh=None):
    """gh: either the full hidden-side pre-activations [N,3C] (unfused forward) or the n-gate part alone [N,C]
    (glam_gru_fused_fwd)."""
    N, C = h.shape
    gh_n = gh[:, 2 * C:] if gh.shape[1] == 3 * C else gh
    assert gh_n.shape[1] == C and gh_n.stride(1) == 1
    g_gi = torch.empty_like(rzn) if g_gi is None else g_gi
    g_gh = torch.empty_like(rzn) if g_gh is None else g_gh
    assert g_gi.is_contiguous() and g_gh.is_contiguous()
    g_h_prev = torch.empty_like(h)
    g_id = torch.empty_like(h) if want_identity else None
    _call("glam_gru_gates_bwd_ex", _p(rzn), _p(gh_n), gh_n.stride(0), _p(h), _p(x_out), _p(g_x_out), _p(g_h_carry), N, C, act,
                                              float(act_param), _p(g_gi), _p(g_gh), _p(g_h_prev), _p(g_id), _stream(h))
    return g_gi, g_gh, g_h_prev, g_id


def lstm_gates_fwd(gates, c_prev):
    R, C = c_prev.shape
    c_new = torch.empty_like(c_prev)
    h_new = torch.empty_like(c_prev)
    _call("glam_lstm_gates_fwd", _p(gates), _p(c_prev), R, C, _p(c_new), _p(h_new), _stream(gates))
    return h_new, c_new


def lstm_gates_bwd(gates_act, c_prev, c_new, g_h, g_c):
    R, C = c_prev.shape
    g_gates = torch.empty_like(gates_act)
    g_c_prev = torch.empty_like(c_prev)
    _call("glam_lstm_gates_bwd", _p(gates_act), _p(c_prev), _p(c_new), _p(g_h), _p(g_c), R, C, _p(g_gates),
                                               _p(g_c_prev), _stream(gates_act))
    return g_gates, g_c_prev


# ------------------------------------------------------------------------------------------------ pooling
def seg_attn_pool_fwd(x, q, q_stride, q_bias, gptr, num_graphs, r_out=None):
    N, C = x.shape
    a = torch.empty((N,), dtype=torch.float32, device=x.device)
    r = torch.empty((num_graphs, C), dtype=torch.float32, device=x.device) if r_out is None else r_out
    asum = torch.empty((num_graphs,), dtype=torch.float32, device=x.device)
    _call("glam_seg_attn_pool_fwd", _p(x), x.stride(0), _p(q), q_stride, _p(q_bias), _p(gptr), num_graphs,
                                                  C, _p(a), _p(r), r.stride(0), _p(asum), _stream(x))
    return a, r, asum


def seg_attn_pool_bwd(x, q, q_stride, a, g_r, g_asum, gptr, num_graphs, g_x, accumulate):
    N, C = x.shape
    g_q = torch.empty((num_graphs, C), dtype=torch.float32, device=x.device)
    g_e = torch.empty((N,), dtype=torch.float32, device=x.device)
    _call("glam_seg_attn_pool_bwd", _p(x), x.stride(0), _p(q), q_stride, _p(a), _p(g_r), g_r.stride(0),
                                                  _p(g_asum), _p(gptr), num_graphs, C, 1 if accumulate else 0, _p(g_x),
                                                  g_x.stride(0), _p(g_q), _p(g_e), _stream(x))
    return g_q, g_e


def set2set_round_fwd(x, gates, c_prev, c_new, gptr, num_graphs, att, u_next, q_star=None):
    _call("glam_set2set_round_fwd", _p(x), x.stride(0), _p(gptr), num_graphs, x.shape[1], _p(gates), _p(c_prev), _p(c_new),
          _p(att), _p(u_next), _p(q_star), _stream(x))


def set2set_round_bwd(x, gates, c_prev, c_new, att, gptr, num_graphs, g_u, g_c, g_x, accumulate, G):
    assert g_u.stride(1) == 1
    _call("glam_set2set_round_bwd", _p(x), x.stride(0), _p(gptr), num_graphs, x.shape[1], _p(gates), _p(c_prev), _p(c_new),
          _p(att), _p(g_u), g_u.stride(0), g_u.shape[1], _p(g_c), _p(g_x), 1 if accumulate else 0, _p(G), _stream(x))


TC_DOT_POOL_MIN_ROWS = 96    # average rows per graph of the second side from which the tensor-core dot-pool is used
SMALL_DOT_POOL_MAX_ROWS = 40     # average rows per graph on BOTH sides up to which the warp-per-pair dot-pool kernel is used


def pair_dot_pool_fwd(xa, xb, ptr_a, ptr_b, num_pairs, idx_b=None):
    """idx_b (int32 [num_pairs], optional): pair g reads graph idx_b[g] of the b side (distinct graphs stored once)."""
    C, dev = xa.shape[1], xa.device
    out = torch.empty((num_pairs, 2), dtype=torch.float32, device=dev)
    argmax = torch.empty((num_pairs, 2), dtype=torch.int32, device=dev)
    sa = torch.empty((num_pairs, C), dtype=torch.float32, device=dev)
    sb = torch.empty((num_pairs, C), dtype=torch.float32, device=dev)
    # second side large (drug-target pairs: ~500 residues): S = Xa Xb^T is a real GEMM -> tensor cores (tf32 math mode); small
    # second sides (drug-drug: 25 x 25) stay on the exact CUDA-core kernel, which is faster there
    rows_b = xb.shape[0] / max(int(ptr_b.numel()) - 1, 1)
    if rows_b >= TC_DOT_POOL_MIN_ROWS and xa.is_contiguous() and xb.is_contiguous() and _lib.load().glam_pair_dot_pool_tc_supported(C):
        if idx_b is not None:
            assert idx_b.dtype == torch.int32 and idx_b.is_contiguous() and idx_b.numel() == num_pairs
        _call("glam_pair_dot_pool_fwd_tc", _p(xa), _p(xb), _p(ptr_a), _p(ptr_b), _p(idx_b), num_pairs, C, _p(out), _p(argmax),
              _p(sa), _p(sb), _stream(xa))
        return out, argmax, sa, sb
    rows_a = xa.shape[0] / max(int(num_pairs), 1)
    if (rows_a <= SMALL_DOT_POOL_MAX_ROWS and rows_b <= SMALL_DOT_POOL_MAX_ROWS and xa.is_contiguous() and xb.is_contiguous()
            and _lib.load().glam_pair_dot_pool_small_supported(C)):
        # both sides small (drug-drug): a warp per pair, no block barrier (csrc/dotpool.cu)
        if idx_b is not None:
            assert idx_b.dtype == torch.int32 and idx_b.is_contiguous() and idx_b.numel() == num_pairs
        _call("glam_pair_dot_pool_fwd_small", _p(xa), _p(xb), _p(ptr_a), _p(ptr_b), _p(idx_b), num_pairs, C, _p(out), _p(argmax),
              _p(sa), _p(sb), _stream(xa))
        return out, argmax, sa, sb
    if idx_b is not None:
        assert idx_b.dtype == torch.int32 and idx_b.is_contiguous() and idx_b.numel() == num_pairs
        _call("glam_pair_dot_pool_fwd_idx", _p(xa), _p(xb), _p(ptr_a), _p(ptr_b), _p(idx_b), num_pairs, C, _p(out), _p(argmax),
              _p(sa), _p(sb), _stream(xa))
        return out, argmax, sa, sb
    _call("glam_pair_dot_pool_fwd", _p(xa), _p(xb), _p(ptr_a), _p(ptr_b), num_pairs, C, _p(out), _p(argmax),
                                                  _p(sa), _p(sb), _stream(xa))
    return out, argmax, sa, sb


def pair_dot_pool_bwd(xa, xb, ptr_a, ptr_b, g_out, argmax, sa, sb, num_pairs):
    g_xa = torch.empty_like(xa)
    g_xb = torch.empty_like(xb)
    _call("glam_pair_dot_pool_bwd", _p(xa), _p(xb), _p(ptr_a), _p(ptr_b), _p(g_out), _p(argmax), _p(sa),
                                                  _p(sb), num_pairs, xa.shape[1], _p(g_xa), _p(g_xb), _stream(xa))
    return g_xa, g_xb


# ------------------------------------------------------------------------------------------------ norms
def pair_norm_fwd(x, gptr, num_graphs, eps, out=None):
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty_like(x, memory_format=torch.contiguous_format) if out is None else out
    _call("glam_pair_norm_fwd", _p(x), x.stride(0), _p(gptr), int(num_graphs), x.shape[1], float(eps), _p(out), out.stride(0), _stream(x))
    return out


def pair_norm_bwd(x, g_y, gptr, num_graphs, eps, out=None, accumulate=False):
    assert x.stride(1) == 1 and g_y.stride(1) == 1
    out = torch.empty_like(x, memory_format=torch.contiguous_format) if out is None else out
    _call("glam_pair_norm_bwd", _p(x), x.stride(0), _p(g_y), g_y.stride(0), _p(gptr), int(num_graphs), x.shape[1], float(eps), _p(out),
          out.stride(0), 1 if accumulate else 0, _stream(x))
    return out


# ------------------------------------------------------------------------------------------------ next rows (§8f)
def pool5_fwd(x, gptr, num_graphs):
    N, C = x.shape
    out = torch.empty((num_graphs, 5 * C), dtype=torch.float32, device=x.device)
    top = torch.empty((num_graphs, 3), dtype=torch.int32, device=x.device)
    _call("glam_pool5_fwd", _p(x), x.stride(0), _p(gptr), num_graphs, C, _p(out), _p(top), _stream(x))
    return out, top


def pool5_bwd(g_out, gptr, top, num_graphs, N, C):
    g_x = torch.empty((N, C), dtype=torch.float32, device=g_out.device)
    _call("glam_pool5_bwd", _p(g_out), _p(gptr), _p(top), num_graphs, C, _p(g_x), g_x.stride(0), _stream(g_out))
    return g_x


def csr_aggregate(Y, rowptr, col, edge_w=None, self_w=None, row_scale=None, bias=None, out=None, accumulate=False,
                  num_rows=None):
    """out [num_rows, F]; `col` indexes rows of Y, which may have a different row count (typed gathers of NNConv)."""
    _need_cuda(Y)
    assert Y.dim() == 2 and Y.stride(1) == 1
    N, F = (Y.shape[0] if num_rows is None else int(num_rows)), Y.shape[1]
    assert self_w is None or N == Y.shape[0]
    if out is None:
        out = torch.empty((N, F), dtype=torch.float32, device=Y.device)
    _call("glam_csr_aggregate", _p(Y), Y.stride(0), _p(rowptr), _p(col), _p(edge_w), _p(self_w), _p(row_scale), _p(bias), N, F,
          _p(out), out.stride(0), 1 if accumulate else 0, _stream(Y))
    return out


def gcn_norm(g):
    """(dinv [N], w_dst [E], w_src [E]) of PyG GCNConv's symmetric normalisation for the graph index g."""
    dev = g.dst_rowptr.device
    dinv = torch.empty((g.num_nodes,), dtype=torch.float32, device=dev)
    w_dst = torch.empty((g.num_edges,), dtype=torch.float32, device=dev)
    w_src = torch.empty((g.num_edges,), dtype=torch.float32, device=dev)
    _call("glam_gcn_norm", _p(g.dst_rowptr), _p(g.dst_src), _p(g.src_rowptr), _p(g.src_dst), g.num_nodes, _p(dinv), _p(w_dst),
          _p(w_src), _stream(dinv))
    return dinv, w_dst, w_src


# ------------------------------------------------------------------------------------------------ optimizer
def adam_step(param, grad, exp_avg, exp_avg_sq, lr, state, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              grad_scale=1.0):
    """torch.optim.Adam semantics on flat fp32 buffers; lr [1] and state [3] are device tensors (see the header)."""
    _need_cuda(param, grad)
    n = param.numel()
    assert param.is_contiguous() and grad.is_contiguous() and grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    _call("glam_adam_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, _p(lr), _p(state), float(beta1), float(beta2),
          float(eps), float(weight_decay), float(grad_scale), _stream(param))
