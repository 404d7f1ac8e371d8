import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from glam_b200 import layer, model, graph as G
from glam_b200.engine import ScreenStep
from glam_b200.synth import make_molecule_batch
from oracle import glam_oracle as O
DEV = "cuda"
kw = dict(hid_dim_alpha=4, e_dim=64, out_dim=1, mol_block="_TripletMessage", message_steps=3, mol_readout="Set2Set", pre_act="ReLU", graph_act="CELU", flat_act="ReLU")
torch.manual_seed(0)
o = O.ArchitectureGP(9, 3, **kw).eval()
m = model.ArchitectureGP(9, 3, graph_do="_None()", end_do="_None()", **kw)
m.load_state_dict(o.state_dict()); m = m.to(DEV).eval()
batches = [make_molecule_batch(64, seed=900 + i, total_nodes=64 * 22, total_edges=64 * 46).pin_memory() for i in range(4)]
rel = lambda a, b: float((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max())
with torch.no_grad():
    ref = [o(b) for b in batches]
    for i, b in enumerate(batches):
        e1 = m(b.to(DEV)); 
        layer.USE_FUSED_STACK = False
        e2 = m(b.to(DEV))
        layer.USE_FUSED_STACK = True
        print(i, "eager fused vs oracle", rel(e1, ref[i]), "eager per-op vs oracle", rel(e2, ref[i]))
s1 = ScreenStep(m, batches[0], device=DEV)
for i, b in enumerate(batches):
    o1 = s1.step(b).clone()
    print(i, "captured vs oracle", rel(o1, ref[i]))
s0 = ScreenStep(m, batches[0], device=DEV, use_cuda_graph=False)
for i, b in enumerate(batches):
    print(i, "uncaptured step vs oracle", rel(s0.step(b).clone(), ref[i]))
print("---- flags / double buffer")
from glam_b200 import ops
b = batches[0].to(DEV)
g = G.graph_index(b.edge_index, b.num_nodes)
gptr, B = G.graph_ptr(b.batch, b.num_graphs)
fi = g.fused_index(gptr, B, b.edge_attr)
meta = torch.zeros(4, dtype=torch.int32, device=DEV)
t = ops.build_graph_tiles(gptr, B, g, meta); et = ops.edge_types(g.sorted_edge_attr(b.edge_attr), meta)
print("fi", fi, "meta", meta.tolist(), "tiles", t[:int(meta[0])].tolist()[:4], "gptr tail", gptr[-3:].tolist(), "N", b.num_nodes)
s2 = ScreenStep(m, batches[0], device=DEV, double_buffer=True)
for i, bb in enumerate(batches):
    o2 = s2.step(bb, prefetch=batches[i + 1] if i + 1 < len(batches) else None).clone()
    print(i, "double-buffered vs oracle", rel(o2, ref[i]))
