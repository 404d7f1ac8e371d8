import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import test_gpu_parity as T
from glam_b200.engine import ScreenStep
from glam_b200.synth import make_molecule_batch
DEV = "cuda"
from glam_b200 import layer
layer.USE_FUSED_STACK = os.environ.get("FUSED", "1") == "1"
print("fused", layer.USE_FUSED_STACK)
m, o = T._gp_pair(9, 3, "Set2Set", "_TripletMessage")
m = m.to(DEV).eval()
batches = [make_molecule_batch(64, seed=900 + i, total_nodes=64 * 22, total_edges=64 * 46).pin_memory() for i in range(4)]
rel = lambda a, b: float((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max())
with torch.no_grad():
    ref = [o(b) for b in batches]
    want = [m(b.to(DEV)).clone() for b in batches]
print("eager vs oracle", [f"{rel(w, r):.1e}" for w, r in zip(want, ref)])
s1 = ScreenStep(m, batches[0], device=DEV)
s2 = ScreenStep(m, batches[0], device=DEV, double_buffer=True)
for i, b in enumerate(batches):
    o1 = s1.step(b).clone()
    o2 = s2.step(b, prefetch=batches[i + 1] if i + 1 < len(batches) else None).clone()
    print(i, "o1 vs oracle", f"{rel(o1, ref[i]):.1e}", "o2 vs oracle", f"{rel(o2, ref[i]):.1e}", "o1==want", torch.equal(o1, want[i]), "o2==want", torch.equal(o2, want[i]),
          "o1 vs want", f"{rel(o1, want[i]):.1e}")
