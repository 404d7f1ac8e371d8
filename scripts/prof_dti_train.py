"""Per-call profile of one GLAM-DTI training step (BASELINE.json configs[3] shape): CUDA events around every library call, eager."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import _lib, graph as G, model as M, ops
from glam_b200.synth import make_molecule_batch, make_protein_batch
_lib.load()
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
P = 256
torch.manual_seed(2)
net = M.ArchitectureDTI(9, 49, 3, 8, hid_dim_alpha=4, e_dim=1024, out_dim=2, mol_block="_TripletMessage", pro_block="_GCNConv",
                        message_steps=3, mol_readout="GlobalPool5", pro_readout="GlobalPool5", pre_act="ReLU", graph_act="CELU",
                        flat_act="ReLU", end_act="ReLU", graph_do="_None()", flat_do="_None()", end_do="_None()").to(dev).train()
lig = make_molecule_batch(P, seed=31, total_nodes=25 * P, total_edges=54 * P, node_dim=9, edge_dim=3).to(dev)
pro = make_protein_batch(P, seed=9600).to(dev)
y = torch.randint(0, 2, (P,), device=dev)
def step():
    for p in net.parameters():
        p.grad = None
    torch.nn.functional.cross_entropy(net(lig, pro), y).backward()
for _ in range(2):
    G.clear_caches(); step()
torch.cuda.synchronize()
sink = []
G.clear_caches(); torch.cuda._sleep(40_000_000)
ops.set_profile(sink); step(); ops.set_profile(None)
torch.cuda.synchronize()
agg = {}
for n, a, b in sink:
    t, c = agg.get(n, (0.0, 0)); agg[n] = (t + a.elapsed_time(b), c + 1)
tot = sum(t for t, _ in agg.values())
print(f"library calls total {tot*1e3:.0f} us ({pro.num_nodes} residues, {pro.num_edges} protein edges)")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"   {k:64s} {t*1e3:9.1f} us  ({c} calls)")
