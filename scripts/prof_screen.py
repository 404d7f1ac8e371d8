"""Where a screening batch goes: CUDA events around every library call of one eval-mode forward (64k graphs), plus the
whole-forward time eager and captured."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import _lib, graph as G, model as M, ops
from glam_b200.engine import ScreenStep
from glam_b200.synth import make_molecule_batch
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
_lib.load()
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
kw = dict(hid_dim_alpha=4, e_dim=1024, out_dim=1, mol_block="_TripletMessage", message_steps=3, mol_readout="Set2Set",
          pre_act="ReLU", graph_act="CELU", flat_act="ReLU", graph_do="_None()", flat_do="_None()", end_do="_None()")
torch.manual_seed(0)
net = M.ArchitectureGP(9, 3, **kw).to(dev).eval()
b = make_molecule_batch(graphs, seed=1, total_nodes=25 * graphs, total_edges=54 * graphs).to(dev)
with torch.no_grad():
    for _ in range(2):
        G.clear_caches(); net(b)
    torch.cuda.synchronize()
    sink = []
    for _ in range(3):
        G.clear_caches()
        torch.cuda.synchronize(); torch.cuda._sleep(30_000_000)
        ops.set_profile(sink); net(b); ops.set_profile(None)
    torch.cuda.synchronize()
agg = {}
for name, e0, e1 in sink:
    t, n = agg.get(name, (0.0, 0)); agg[name] = (t + e0.elapsed_time(e1), n + 1)
tot = 0
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:60s} {t/3*1e3:9.1f} us/forward  ({n//3} calls)"); tot += t / 3
print(f"library calls total {tot*1e3:.1f} us")
ss = ScreenStep(net, b, device=dev)
for _ in range(3): ss.run_resident()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): ss.run_resident()
e1.record(); torch.cuda.synchronize()
print(f"captured forward: {e0.elapsed_time(e1)/10*1e3:.1f} us per {graphs}-graph batch = {graphs/(e0.elapsed_time(e1)/10*1e-3)/1e6:.2f} M graphs/s")

# phase clock of the fused kernel inside the captured forward (glam_message_stack_phase_clock)
PH = ["tile load", "logits", "softmax", "proj wait", "tmem->xp", "aggregate", "agg panels", "scale wait", "celu epi", "gru wait", "gates",
      "outputs", "tile end"]
lib = _lib.load()
clk = torch.zeros(148, 32, dtype=torch.int64, device=dev)
with torch.no_grad():
    G.clear_caches(); net(b); torch.cuda.synchronize()
    lib.glam_message_stack_phase_clock(clk.data_ptr())
    G.clear_caches(); net(b); torch.cuda.synchronize()
    lib.glam_message_stack_phase_clock(None)
c = clk.double().mean(0).cpu()
tot = float(c[:13].sum())
print(f"fused kernel in the screening forward: {tot/1.965e3:.1f} us of SM cycles per CTA; " + ", ".join(f"{n} {100*float(c[i])/tot:.1f}%" for i, n in enumerate(PH))
      + f"; set-up {float(c[14])/1.965e3:.1f} us")
