"""Ad-hoc micro-benchmark of individual library calls at the bench shapes (CUDA events, L2 flushed between reps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import ops, _lib, graph
from glam_b200.synth import make_molecule_batch
_lib.load()
dev = "cuda"
N, C, H, De = 102400, 36, 3, 3
HC, ld = H * C, 116
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def timeit(name, fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    print(f"{name:44s} {t*1e3:8.1f} us  {nbytes/t/1e6:8.1f} GB/s  ({nbytes/1e6:.1f} MB)")

which = sys.argv[1] if len(sys.argv) > 1 else "all"
x = torch.randn(N, C, device=dev); w_ext = torch.randn(C, ld, device=dev); agg = torch.randn(N, HC, device=dev)
w_scale = torch.randn(HC, C, device=dev); bias = torch.randn(C, device=dev); w_ih = torch.randn(3 * C, C, device=dev); b3 = torch.randn(3 * C, device=dev)
g108 = torch.randn(N, HC, device=dev); g116 = torch.randn(N, ld, device=dev)
if which in ("all", "gemm"):
    timeit("gemm node_proj [N,36]x[36,116] exact6", lambda: ops.gemm(x, w_ext, exact_cols=(108, 114)), 4 * N * (C + ld))
    timeit("gemm scale [N,108]x[108,36] celu", lambda: ops.gemm(agg, w_scale, bias=bias, epilogue=1), 4 * N * (HC + C))
    timeit("gemm scale [N,108]x[108,36] bias only", lambda: ops.gemm(agg, w_scale, bias=bias, epilogue=0), 4 * N * (HC + C))
    timeit("gemm scale^T [N,108]x[36,108]^T", lambda: ops.gemm(agg, w_scale.t().contiguous(), transpose_w=True), 4 * N * (HC + C))
    timeit("gemm gru [N,36]x[36,108]^T", lambda: ops.gemm(x, w_ih, transpose_w=True, bias=b3), 4 * N * (C + HC))
    timeit("gemm dgrad [N,116]x[116,36]^T", lambda: ops.gemm(g116, w_ext, transpose_w=True), 4 * N * (ld + C))
if which in ("all", "gemm", "small"):
    B_ = 4096
    q = torch.randn(B_, 3 * C, device=dev); wl = torch.randn(4 * C, 3 * C, device=dev); bl = torch.randn(4 * C, device=dev)
    timeit("gemm lstm gates [4096,108]x[108,144]^T", lambda: ops.gemm(q, wl, transpose_w=True, bias=bl), 4 * B_ * 7 * C)
    qs = torch.randn(65536, 3 * C, device=dev)
    timeit("gemm lstm gates [65536,108]x[108,144]^T", lambda: ops.gemm(qs, wl, transpose_w=True, bias=bl), 4 * 65536 * 7 * C)
    gg = torch.randn(B_, 4 * C, device=dev)
    timeit("gemm lstm dgrad [4096,144]x[144,108]", lambda: ops.gemm(gg, wl), 4 * B_ * 7 * C)
if which in ("all", "tn"):
    timeit("gemm_tn [N,36]^T[N,108] +colsum", lambda: ops.gemm_tn_ex(x, g108, transpose_out=True, want_colsum=True), 4 * N * (C + HC))
    timeit("gemm_tn [N,36]^T[N,116]", lambda: ops.gemm_tn_ex(x, g116), 4 * N * (C + ld))
    timeit("skinny_tn [N,36]^T[N,6]", lambda: ops.gemm_tn_ex(x, g116[:, 108:114]), 4 * N * (C + 6))
    x3 = torch.randn(3 * N, C, device=dev); g3 = torch.randn(3 * N, ld, device=dev); xr = torch.randn(N, 9, device=dev)
    timeit("skinny_tn [3N,36]^T[3N,6] (logit columns)", lambda: ops.gemm_tn_ex(x3, g3[:, 108:114]), 4 * 3 * N * (C + 6))
    timeit("skinny_tn [N,9]^T[N,36] +colsum (input lin)", lambda: ops.gemm_tn_ex(xr, x, want_colsum=True), 4 * N * (C + 9))
if which in ("all", "edge"):
    b = make_molecule_batch(4096, total_nodes=N, total_edges=221184, seed=1).to(dev)
    g = graph.graph_index(b.edge_index, N); ea = g.sorted_edge_attr(b.edge_attr); E = b.num_edges
    xpe = torch.randn(N, ld, device=dev); we = torch.randn(De, HC, device=dev); ae = torch.randn(De, H, device=dev)
    fwd_bytes = 4 * (N * ld + N * HC + E * De + E * H) + 4 * (E + N)
    dst_bytes = 4 * (N * ld + N * HC + E * De + 2 * E * H + N * H) + 4 * (E + N)
    src_bytes = 4 * (N * HC + N * ld + E * De + 2 * E * H) + 4 * (3 * E + N)
    res = {}
    for tiles in (False, True):
        ops.USE_EDGE_TILES = tiles
        tag = "windowed" if tiles else "gather  "
        timeit(f"edge_fwd [{tag}]", lambda: ops.triplet_edge_fwd(xpe, ea, we, ae, g, H, C, 0.2), fwd_bytes)
        aggo, alpha = ops.triplet_edge_fwd(xpe, ea, we, ae, g, H, C, 0.2)
        timeit(f"edge_bwd dst+src [{tag}]", lambda: ops.triplet_edge_bwd(xpe, ea, we, ae, alpha, g108, g, H, C, 0.2), dst_bytes + src_bytes)
        sink = []
        ops.set_profile(sink)
        for _ in range(5):
            flush.zero_()
            ops.triplet_edge_bwd(xpe, ea, we, ae, alpha, g108, g, H, C, 0.2)
        ops.set_profile(None); torch.cuda.synchronize()
        for nm, nb in (("glam_triplet_edge_bwd_dst", dst_bytes), ("glam_triplet_edge_bwd_src", src_bytes)):
            ts = sorted(e0.elapsed_time(e1) for n, e0, e1 in sink if n == nm); t = ts[len(ts) // 2]
            print(f"   {nm:41s} {t*1e3:8.1f} us  {nb/t/1e6:8.1f} GB/s  ({nb/1e6:.1f} MB)")
        res[tiles] = (aggo, alpha) + tuple(ops.triplet_edge_bwd(xpe, ea, we, ae, alpha, g108, g, H, C, 0.2))
    for nm, a_, b_ in zip(("agg", "alpha", "g_xpe", "g_logit", "g_we"), res[False], res[True]):
        d = (a_ - b_).abs().max().item()
        print(f"   windowed vs gather {nm:8s} max|diff| = {d:.3e}  bitwise_equal = {bool(torch.equal(a_, b_))}  (scale {a_.abs().max().item():.3e})")
    gi = torch.randn(N, 3 * C, device=dev); gh = torch.randn(N, 3 * C, device=dev); h = torch.randn(N, C, device=dev)
    timeit("gru_gates_fwd", lambda: ops.gru_gates_fwd(gi.clone(), gh, h, x, 3, 1.0), 4 * N * C * 13)
if which == "edgefwd":
    b = make_molecule_batch(4096, total_nodes=N, total_edges=221184, seed=1).to(dev)
    g = graph.graph_index(b.edge_index, N); ea = g.sorted_edge_attr(b.edge_attr); E = b.num_edges
    xpe = torch.randn(N, ld, device=dev); we = torch.randn(De, HC, device=dev); ae = torch.randn(De, H, device=dev)
    fwd_bytes = 4 * (N * ld + N * HC + E * De + E * H) + 4 * (E + N)
    timeit("edge_fwd windowed dbg=" + os.environ.get("GLAM_B200_EDGE_WIN_DEBUG", "0") + " v=" + os.environ.get("GLAM_B200_EDGE_WIN", "2"),
           lambda: ops.triplet_edge_fwd(xpe, ea, we, ae, g, H, C, 0.2), fwd_bytes, reps=20)
if which in ("all", "s2s"):
    B = 4096
    b = make_molecule_batch(B, total_nodes=N, total_edges=221184, seed=1).to(dev)
    gptr, _ = graph.graph_ptr(b.batch, B)
    U = torch.randn(B, 3 * C, device=dev); wc = torch.randn(4 * C, 3 * C, device=dev); bs = torch.randn(4 * C, device=dev)
    gates = torch.empty(B, 4 * C, device=dev)
    timeit("s2s gates gemm [4096,108]x[144,108]^T", lambda: ops.gemm(U, wc, transpose_w=True, bias=bs, out=gates), 4 * B * 7 * C)
    Gm = torch.randn(B, 4 * C, device=dev)
    timeit("s2s g_u gemm [4096,144]x[144,108]", lambda: ops.gemm(Gm, wc), 4 * B * 7 * C)
    c0 = torch.zeros(B, C, device=dev); c1 = torch.empty(B, C, device=dev); att = torch.empty(N, device=dev); un = torch.empty(B, 3 * C, device=dev)
    pre = torch.randn(B, 4 * C, device=dev)
    def fwd():
        gates.copy_(pre)
        ops.set2set_round_fwd(x, gates, c0, c1, gptr, B, att, un, None)
    timeit("s2s round_fwd (+copy)", fwd, 4 * (N * C + N + B * 12 * C))
    gu = torch.randn(B, 3 * C, device=dev); gc = torch.zeros(B, C, device=dev); gx = torch.zeros(N, C, device=dev); Go = torch.empty(B, 4 * C, device=dev)
    timeit("s2s round_bwd (overwrite)", lambda: ops.set2set_round_bwd(x, gates, c0, c1, att, gptr, B, gu, gc, gx, False, Go), 4 * (2 * N * C + N + B * 13 * C))
    timeit("s2s round_bwd (accumulate)", lambda: ops.set2set_round_bwd(x, gates, c0, c1, att, gptr, B, gu, gc, gx, True, Go), 4 * (3 * N * C + N + B * 13 * C))
    timeit("s2s wgrad tn [12288,108]^T[12288,144]", lambda: ops.gemm_tn_ex(torch.randn(3 * B, 3 * C, device=dev), torch.randn(3 * B, 4 * C, device=dev), transpose_out=True, want_colsum=True), 4 * 3 * B * 7 * C)
if which in ("all", "gru"):
    m = torch.randn(N, C, device=dev); h = torch.randn(N, C, device=dev); w_hh = torch.randn(3 * C, C, device=dev)
    timeit("gru fused fwd (2 GEMMs + gates)", lambda: ops.gru_fused_fwd(m, h, x, w_ih, w_hh, b3, b3, 3, 1.0), 4 * N * C * 9)
    def unf():
        gi = ops.gemm(m, w_ih, transpose_w=True, bias=b3); gh = ops.gemm(h, w_hh, transpose_w=True, bias=b3)
        ops.gru_gates_fwd(gi, gh, h, x, 3, 1.0)
    timeit("gru unfused fwd (3 launches)", unf, 4 * N * C * 9)
