"""Per-call profile of a DTI screening batch against one target: distinct protein once (pro_index) vs one copy per pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import _lib, graph as G, model as M, ops
from glam_b200.synth import make_molecule_batch, make_protein_batch
_lib.load()
dev = "cuda"
P = 1024
torch.manual_seed(2)
net = M.ArchitectureDTI(9, 49, 3, 8, hid_dim_alpha=4, e_dim=1024, out_dim=2, mol_block="_TripletMessage", pro_block="_GCNConv",
                        message_steps=3, mol_readout="GlobalPool5", pro_readout="GlobalPool5", pre_act="ReLU", graph_act="CELU",
                        flat_act="ReLU", end_act="ReLU", graph_do="_None()", flat_do="_None()", end_do="_None()").to(dev).eval()
lig = make_molecule_batch(P, seed=31, total_nodes=25 * P, total_edges=54 * P, node_dim=9, edge_dim=3).to(dev)
one = make_protein_batch(1, seed=32, min_len=500, max_len=500).to(dev)
rep = make_protein_batch(P, seed=32, min_len=500, max_len=500, same_protein=True).to(dev)
idx = torch.zeros(P, dtype=torch.int32, device=dev)
for name, fn in (("distinct protein once", lambda: net(lig, one, pro_index=idx)), ("one copy per pair", lambda: net(lig, rep))):
    with torch.no_grad():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        sink = []
        torch.cuda._sleep(30_000_000)
        ops.set_profile(sink); fn(); ops.set_profile(None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
    agg = {}
    for n, a, b in sink:
        t, c = agg.get(n, (0.0, 0)); agg[n] = (t + a.elapsed_time(b), c + 1)
    print(f"== {name}: {e0.elapsed_time(e1)/5*1e3:.0f} us per batch; library calls:")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]:
        print(f"   {k:56s} {t*1e3:9.1f} us  ({c} calls)")
