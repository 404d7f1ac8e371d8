"""Debug driver of the one-launch message-stack backward (csrc/mp_fused_bwd.cu): every tensor it writes against the per-op
backward kernels on the same saved activations.   python scripts/dbg_bwd.py [graphs] [C] [De] [res] [act]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import _lib, graph as G, layer, ops, functional as Fn
from glam_b200.synth import make_molecule_batch

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
C = int(sys.argv[2]) if len(sys.argv) > 2 else 36
De = int(sys.argv[3]) if len(sys.argv) > 3 else 3
res = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
actname = sys.argv[5] if len(sys.argv) > 5 else "CELU"
_lib.load()
dev = "cuda"
H, S = 3, 3
b = make_molecule_batch(graphs, seed=5, node_dim=C, edge_dim=De).to(dev)
N, E = b.num_nodes, b.num_edges
torch.manual_seed(0)
blk = layer.MessageBlock(C, C, De, norm="_None", dropout="_None()", conv="_TripletMessage", act=actname, res=res).to(dev).train()
with torch.no_grad():
    for p in blk.parameters():
        if p.dim() == 1:
            p.uniform_(-0.2, 0.2)
inner, gru = blk.conv.conv, blk.gru
act = {"CELU": ops.ACT_CELU, "ReLU": ops.ACT_RELU}[actname]
x0 = torch.randn(N, C, device=dev)
g = G.graph_index(b.edge_index, N)
gptr, B = G.graph_ptr(b.batch, b.num_graphs)
fi = g.fused_index(gptr, B, b.edge_attr)
assert fi is not None
ea = g.sorted_edge_attr(b.edge_attr)
with torch.no_grad():
    w_ext, att_edge = inner.derived()
ld = w_ext.shape[1]
sv = Fn._stack_buffers(x0, S, H, C, ld, E)
w = (w_ext, inner.weight_edge, att_edge, inner.weight_scale)
ops.message_stack_fwd(x0, None, *w, inner.bias, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, g, fi, H, C, S,
                      0.2, act, 1.0, res, save=sv)
cot = [torch.randn(N, C, device=dev) for _ in range(S)]
g_hf = torch.randn(N, C, device=dev)
new = lambda *s: torch.full(s, float("nan"), device=dev)
G_GI, G_GH, G_PRE, G_XPE = new(S, N, 3 * C), new(S, N, 3 * C), new(S, N, C), new(S, N, ld)
gx0, gwe, gae = ops.message_stack_bwd(sv, cot, g_hf, *w, gru.weight_ih_l0, gru.weight_hh_l0, g, fi, H, C, S, 0.2, act, 1.0, res,
                                      G_GI, G_GH, G_PRE, G_XPE)
torch.cuda.synchronize()

# the per-op reverse loop (functional.MessageStackFn.backward)
X, HH, XPE, ALPHA, M, RZN, GH = (sv[k] for k in ("X", "HH", "XPE", "ALPHA", "M", "RZN", "GH"))
R_GI, R_GH, R_PRE, R_XPE = new(S, N, 3 * C), new(S, N, 3 * C), new(S, N, C), new(S, N, ld)
R_LOGIT, R_WE = new(S, E, H), new(S, De, H * C)
g_x, g_h = None, g_hf.clone()
for s in range(S - 1, -1, -1):
    g_x = cot[s].clone() if g_x is None else g_x + cot[s]
    _, _, g_h_prev, g_id = ops.gru_gates_bwd(RZN[s], GH[s], HH[s], X[s + 1], g_x, g_h, act, 1.0, res, g_gi=R_GI[s], g_gh=R_GH[s])
    ops.gemm(R_GI[s], gru.weight_ih_l0, epilogue=ops.EPI_MUL_CELU_GRAD, aux=M[s], out=R_PRE[s])
    ops.gemm(R_GH[s], gru.weight_hh_l0, epilogue=ops.EPI_ACCUM, out=g_h_prev)
    g_agg = ops.gemm(R_PRE[s], inner.weight_scale, transpose_w=True)
    ops.triplet_edge_bwd(XPE[s], ea, inner.weight_edge, att_edge, ALPHA[s], g_agg, g, H, C, 0.2, g_xpe=R_XPE[s], g_logit=R_LOGIT[s], g_we=R_WE[s])
    if res:
        g_x = ops.gemm(R_XPE[s], w_ext, transpose_w=True, epilogue=ops.EPI_ACCUM, out=g_id)
    else:
        g_x = ops.gemm(R_XPE[s], w_ext, transpose_w=True)
    g_h = g_h_prev
r_gx0 = g_x + g_h
r_gae, _ = ops.gemm_tn_ex(ea, R_LOGIT.sum(0))
r_gwe = R_WE.sum(0)
torch.cuda.synchronize()


def rel(a, b):
    a, b = a.double(), b.double()
    bad = ~torch.isfinite(a)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30)), int(bad.sum())


for s in range(S - 1, -1, -1):
    for name, a, r in (("G_GI", G_GI[s], R_GI[s]), ("G_GH", G_GH[s], R_GH[s]), ("G_PRE", G_PRE[s], R_PRE[s]), ("G_XPE", G_XPE[s], R_XPE[s]),
                       ("G_XPE[:, :HC]", G_XPE[s][:, :H * C], R_XPE[s][:, :H * C]), ("G_XPE[:, HC:]", G_XPE[s][:, H * C:], R_XPE[s][:, H * C:])):
        e, nb = rel(a, r)
        print(f"step {s} {name:16s} rel err {e:.3e}  non-finite {nb}")
for name, a, r in (("g_x0", gx0, r_gx0), ("g_w_edge", gwe, r_gwe), ("g_att_edge", gae, r_gae)):
    e, nb = rel(a, r)
    print(f"{name:16s} rel err {e:.3e}  non-finite {nb}")

# the same through the tile-blocked gate save (what training uses)
svt = Fn._stack_buffers(x0, S, H, C, ld, E, tiled_gates=True)
ops.message_stack_fwd(x0, None, *w, inner.bias, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, g, fi, H, C, S,
                      0.2, act, 1.0, res, save=svt)
T_GI, T_GH, T_PRE, T_XPE = new(S, N, 3 * C), new(S, N, 3 * C), new(S, N, C), new(S, N, ld)
tgx0, tgwe, tgae = ops.message_stack_bwd(svt, cot, g_hf, *w, gru.weight_ih_l0, gru.weight_hh_l0, g, fi, H, C, S, 0.2, act, 1.0, res,
                                         T_GI, T_GH, T_PRE, T_XPE)
torch.cuda.synchronize()
print("tile-blocked gate save == row-major save:", all(torch.equal(a, b_) for a, b_ in ((tgx0, gx0), (tgwe, gwe), (tgae, gae), (T_GI, G_GI),
                                                                                      (T_GH, G_GH), (T_PRE, G_PRE), (T_XPE, G_XPE))))
sv = svt

if "time" in sys.argv:
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def f_fused():
        ops.message_stack_bwd(sv, cot, g_hf, *w, gru.weight_ih_l0, gru.weight_hh_l0, g, fi, H, C, S, 0.2, act, 1.0, res, G_GI, G_GH, G_PRE, G_XPE)

    def f_perop():
        g_x, g_h = None, g_hf.clone()
        for s in range(S - 1, -1, -1):
            g_x = cot[s].clone() if g_x is None else g_x + cot[s]
            _, _, g_h_prev, g_id = ops.gru_gates_bwd(RZN[s], GH[s], HH[s], X[s + 1], g_x, g_h, act, 1.0, res, g_gi=R_GI[s], g_gh=R_GH[s])
            ops.gemm(R_GI[s], gru.weight_ih_l0, epilogue=ops.EPI_MUL_CELU_GRAD, aux=M[s], out=R_PRE[s])
            ops.gemm(R_GH[s], gru.weight_hh_l0, epilogue=ops.EPI_ACCUM, out=g_h_prev)
            g_agg = ops.gemm(R_PRE[s], inner.weight_scale, transpose_w=True)
            ops.triplet_edge_bwd(XPE[s], ea, inner.weight_edge, att_edge, ALPHA[s], g_agg, g, H, C, 0.2, g_xpe=R_XPE[s], g_logit=R_LOGIT[s], g_we=R_WE[s])
            g_x = ops.gemm(R_XPE[s], w_ext, transpose_w=True, epilogue=ops.EPI_ACCUM, out=g_id) if res else ops.gemm(R_XPE[s], w_ext, transpose_w=True)
            g_h = g_h_prev

    for name, fn in (("one-launch backward", f_fused), ("per-op reverse loop", f_perop)):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{name:24s} {ts[len(ts)//2]*1e3:8.1f} us  (tiles {int(fi.meta[0])}, N {N}, E {E})")

if "phases" in sys.argv:
    lib = _lib.load()
    clk = torch.zeros(148, 32, dtype=torch.int64, device=dev)
    names = ["index words", "gate bwd", "G copy-out + GRU MMAs", "CELU' epi", "G_PRE copy + scale MMA", "tmem->g_agg", "dst pass", "src pass",
             "G_XPE copy-out", "tile output"]
    lib.glam_message_stack_phase_clock(clk.data_ptr())
    ops.message_stack_bwd(sv, cot, g_hf, *w, gru.weight_ih_l0, gru.weight_hh_l0, g, fi, H, C, S, 0.2, act, 1.0, res, G_GI, G_GH, G_PRE, G_XPE)
    torch.cuda.synchronize()
    lib.glam_message_stack_phase_clock(None)
    c = clk.double().mean(0).cpu()[16:]
    tot = float(c[:10].sum())
    tot += float(c[10] + c[11])
    print(f"backward: {tot/1.965e3:.1f} us of SM cycles per CTA; " + ", ".join(f"{n} {100*float(c[i])/tot:.1f}%" for i, n in enumerate(names))
          + f"; [gate bwd split: wait for g_x MMA + first loads {100*float(c[10])/tot:.1f}%, rounds (thread 0) {100*float(c[11])/tot:.1f}%, rest = barrier]")
