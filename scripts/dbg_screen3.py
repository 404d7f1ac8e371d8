import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import test_gpu_parity as T
from glam_b200.engine import ScreenStep
from glam_b200 import graph as G
from glam_b200.synth import make_molecule_batch
DEV = "cuda"
batches = [make_molecule_batch(64, seed=900 + i, total_nodes=64 * 22, total_edges=64 * 46).pin_memory() for i in range(4)]
rel = lambda a, b: float((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max())
def fresh():
    m, o = T._gp_pair(9, 3, "Set2Set", "_TripletMessage")
    with torch.no_grad():
        ref = [o(b) for b in batches]
    return m.to(DEV).eval(), ref
for variant in ("none", "eager_call", "s2_nograph", "s2_single", "s2_double", "s2_double_other_model", "clear_caches_only", "empty_cache"):
    m, ref = fresh()
    s1 = ScreenStep(m, batches[0], device=DEV)
    if variant == "eager_call":
        with torch.no_grad(): m(batches[1].to(DEV))
    elif variant == "s2_nograph":
        s2 = ScreenStep(m, batches[0], device=DEV, use_cuda_graph=False); s2.step(batches[1])
    elif variant == "s2_single":
        s2 = ScreenStep(m, batches[0], device=DEV)
    elif variant == "s2_double":
        s2 = ScreenStep(m, batches[0], device=DEV, double_buffer=True)
    elif variant == "s2_double_other_model":
        m2, _ = fresh(); s2 = ScreenStep(m2, batches[0], device=DEV, double_buffer=True)
    elif variant == "clear_caches_only":
        G.clear_caches()
    elif variant == "empty_cache":
        torch.cuda.empty_cache()
    errs = [rel(s1.step(b).clone(), ref[i]) for i, b in enumerate(batches)]
    print(f"{variant:24s}", ["%.1e" % e for e in errs])
    del s1
