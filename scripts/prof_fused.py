"""Timing / profiling driver of the fused message-stack kernel at the bench shape (4096 graphs; or argv[2] graphs).
    python scripts/prof_fused.py [time|ncu|phases] [graphs]
`time`: CUDA-event medians (L2 flushed between launches) of the eval, save and conv-only modes next to the per-op path.
`ncu` : a few launches of each mode for `ncu -k regex:mp_fused`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import _lib, graph as G, layer, ops, functional as Fn
from glam_b200.synth import make_molecule_batch

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
_lib.load()
dev = "cuda"
C, H, De, S = 36, 3, 3, 3
b = make_molecule_batch(graphs, seed=1234, total_nodes=25 * graphs, total_edges=54 * graphs, node_dim=9, edge_dim=3).to(dev)
N, E = b.num_nodes, b.num_edges
torch.manual_seed(0)
blk = layer.MessageBlock(C, C, De, norm="_None", dropout="_None()", conv="_TripletMessage", act="CELU", res=True).to(dev).eval()
inner, gru = blk.conv.conv, blk.gru
x0 = torch.randn(N, C, device=dev)
g = G.graph_index(b.edge_index, N)
gptr, B = G.graph_ptr(b.batch, b.num_graphs)
fi = g.fused_index(gptr, B, b.edge_attr)
assert fi is not None
ea = g.sorted_edge_attr(b.edge_attr)
print("tiles", int(fi.meta[0]), "N", N, "E", E)
with torch.no_grad():
    w_ext, att_edge = inner.derived()
ld = w_ext.shape[1]
sv = Fn._stack_buffers(x0, S, H, C, ld, E)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
args = (w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0)


def f_eval():
    return ops.message_stack_fwd(x0, None, *args, g, fi, H, C, S, 0.2, ops.ACT_CELU, 1.0, True)


def f_eval1():
    return ops.message_stack_fwd(x0, None, *args, g, fi, H, C, 1, 0.2, ops.ACT_CELU, 1.0, True)


def f_save():
    return ops.message_stack_fwd(x0, None, *args, g, fi, H, C, S, 0.2, ops.ACT_CELU, 1.0, True, save=sv)


def f_conv():
    return ops.message_stack_fwd(x0, None, w_ext, inner.weight_edge, att_edge, inner.weight_scale, inner.bias, None, None, None, None,
                                 g, fi, H, C, 1, 0.2, ops.ACT_NONE, 0.0, False, conv_only=True)


def f_perop():
    layer.USE_FUSED_STACK = False
    try:
        with torch.no_grad():
            return blk.run_steps(x0, b.edge_index, b.edge_attr, S, batch=b.batch, num_graphs=b.num_graphs)
    finally:
        layer.USE_FUSED_STACK = True


def timeit(name, fn, nbytes, reps=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    print(f"{name:40s} {t*1e3:8.1f} us  {nbytes/t/1e6:8.1f} GB/s of {nbytes/1e6:.1f} MB algorithmic; {graphs/t/1e3:.2f} M graphs/s")


n, e = N / graphs, E / graphs
blk_bytes = graphs * (4 * (4 * n * C + e * De) + 8 * e + 4 * (n + 1))          # MessageBlock fwd per step
conv_bytes = graphs * (4 * (2 * n * C + e * De) + 8 * e + 4 * (n + 1))         # TripletMessage fwd
PHASES = ["tile load", "logits", "softmax", "proj wait", "tmem->xp", "aggregate", "agg panels", "scale wait", "celu epi", "gru wait",
          "gates", "outputs", "tile end"]
if mode == "phases":
    lib = _lib.load()
    clk = torch.zeros(148, 32, dtype=torch.int64, device=dev)
    for name, fn in (("eval x3", f_eval), ("save x3", f_save), ("conv only", f_conv)):
        fn(); torch.cuda.synchronize()
        lib.glam_message_stack_phase_clock(clk.data_ptr())
        fn(); torch.cuda.synchronize()
        lib.glam_message_stack_phase_clock(None)
        c = clk.double().mean(0).cpu()
        tot = float(c[:13].sum())
        print(f"{name}: {tot/1.965e3:.1f} us of SM cycles per CTA; " + ", ".join(f"{n} {100*float(c[i])/tot:.1f}%" for i, n in enumerate(PHASES))
              + f"; [issue of the projection + gh MMAs {100*float(c[13])/tot:.1f}%; set-up before the first tile {float(c[14])/1.965e3:.1f} us]")
elif mode == "ncu":
    for fn in (f_eval, f_save, f_conv):
        fn(); fn()
    torch.cuda.synchronize()
else:
    timeit("fused eval, 3 steps", f_eval, 3 * blk_bytes)
    timeit("fused eval, 1 step", f_eval1, blk_bytes)
    timeit("fused save, 3 steps", f_save, 3 * blk_bytes)
    timeit("fused conv only", f_conv, conv_bytes)
    timeit("per-op path, 3 steps (eval)", f_perop, 3 * blk_bytes)
