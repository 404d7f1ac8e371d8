"""Generate tests/golden/*.pt by running the REFERENCE's own, unmodified layer.py / model.py.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports torch_geometric==1.7.2 / torch_scatter, which are not installed here; they are
provided by the restated shim under oracle/pyg_shim (third-party semantics only).  Every fixture
stores the seeded inputs, the module state_dict, the forward output and the gradients of
sum(out * cotangent) with respect to the inputs and parameters, all produced by reference code on CPU
in fp32 (and in fp64 for the layer-level cases, as the ground truth for tolerance checks).
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, ROOT)

from glam_b200.synth import make_molecule_batch, make_protein_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_ref(src_dir: str, name: str):
    """Import /root/reference/<src_dir>/<name>.py under a unique module name; `from layer import ...`
    inside model.py resolves to the same directory's layer.py."""
    path = os.path.join(REF, src_dir)
    if name == "model":
        sys.modules["layer"] = load_ref(src_dir, "layer")
    spec = importlib.util.spec_from_file_location(f"ref_{src_dir}_{name}", os.path.join(path, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def grads_of(out, cot, tensors):
    g = torch.autograd.grad((out * cot).sum(), tensors, allow_unused=True)
    return [None if t is None else t.detach().clone() for t in g]


def add_edge_cases(b):
    """Append an isolated node (own graph), and duplicate two edges — the protein featuriser can emit
    duplicate (i,i±1) pairs (src_2gi_dti_scr/dataset.py:76-100)."""
    N = b.x.shape[0]
    x = torch.cat([b.x, b.x[:1] * 0.5 + 0.25])
    batch = torch.cat([b.batch, b.batch[-1:] + 1])
    ei = torch.cat([b.edge_index, b.edge_index[:, 3:5]], dim=1)
    ea = torch.cat([b.edge_attr, b.edge_attr[3:5]])
    return x, ei, ea, batch, N + 1


def case_layer(layer, kind, C, De, seed, dtype, edge_cases=False):
    torch.manual_seed(seed)
    b = make_molecule_batch(4, node_dim=C, edge_dim=De, seed=seed, features="normal")
    x, ei, ea, batch = b.x, b.edge_index, b.edge_attr, b.batch
    if edge_cases:
        x, ei, ea, batch, _ = add_edge_cases(b)
    mod = (layer.TripletMessage(C, De) if kind == "triplet" else layer.TripletMessageLight(C, De))
    with torch.no_grad():
        mod.bias.uniform_(-0.1, 0.1)
    mod = mod.to(dtype)
    x = x.to(dtype).requires_grad_(True)
    ea = ea.to(dtype)
    out = mod(x, ei, ea)
    cot = torch.randn(out.shape).to(dtype)
    params = list(mod.parameters())
    g = grads_of(out, cot, [x] + params)
    return {"x": x.detach(), "edge_index": ei, "edge_attr": ea, "batch": batch, "cot": cot,
            "state": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "out": out.detach(), "grad_x": g[0],
            "grad_params": {n: gg for (n, _), gg in zip(mod.named_parameters(), g[1:])}}


def case_block(layer, conv, norm, act, res, C, De, seed, dtype, steps=3):
    torch.manual_seed(seed)
    b = make_molecule_batch(4, node_dim=C, edge_dim=De, seed=seed, features="normal")
    blk = layer.MessageBlock(C, C, De, norm=norm, dropout="_None()", conv=conv, act=act, res=res).to(dtype)
    x0 = b.x.to(dtype).requires_grad_(True)
    ea = b.edge_attr.to(dtype)
    x, h = x0, None
    for _ in range(steps):
        x, h = blk(x, b.edge_index, ea, h=h, batch=b.batch)
    cot = torch.randn(x.shape).to(dtype)
    coth = torch.randn(h.shape).to(dtype)
    params = list(blk.parameters())
    g = torch.autograd.grad((x * cot).sum() + (h * coth).sum(), [x0] + params)
    return {"x": x0.detach(), "edge_index": b.edge_index, "edge_attr": ea, "batch": b.batch, "cot": cot, "coth": coth,
            "cfg": {"conv": conv, "norm": norm, "act": act, "res": res, "steps": steps, "C": C, "De": De},
            "state": {k: v.detach().clone() for k, v in blk.state_dict().items()},
            "out": x.detach(), "h": h.detach(), "grad_x": g[0].clone(),
            "grad_params": {n: gg.clone() for (n, _), gg in zip(blk.named_parameters(), g[1:])}}


def case_readout(layer, kind, C, seed, dtype):
    from torch_geometric.nn import Set2Set
    torch.manual_seed(seed)
    b = make_molecule_batch(5, node_dim=C, edge_dim=3, seed=seed, features="normal")
    mod = (Set2Set(in_channels=C, processing_steps=3) if kind == "set2set" else layer.GlobalLAPool(C)).to(dtype)
    x = b.x.to(dtype).requires_grad_(True)
    out = mod(x, b.batch)
    cot = torch.randn(out.shape).to(dtype)
    g = grads_of(out, cot, [x] + list(mod.parameters()))
    return {"x": x.detach(), "batch": b.batch, "cot": cot,
            "state": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "out": out.detach(), "grad_x": g[0],
            "grad_params": {n: gg for (n, _), gg in zip(mod.named_parameters(), g[1:])}}


def case_dotpool(layer_ddi, C, seed, dtype, protein=False):
    a = make_molecule_batch(5, node_dim=C, edge_dim=3, seed=seed, features="normal")
    if protein:
        p = make_protein_batch(5, node_dim=C, edge_dim=8, seed=seed + 1, min_len=40, max_len=90)
        pb_x = torch.randn(p.x.shape[0], C, generator=torch.Generator().manual_seed(seed))
        bb = p.batch
    else:
        p = make_molecule_batch(5, node_dim=C, edge_dim=3, seed=seed + 1, features="normal")
        pb_x, bb = p.x, p.batch
    xa = a.x.to(dtype).requires_grad_(True)
    xb = pb_x.to(dtype).requires_grad_(True)
    out = layer_ddi.dot_and_global_pool2(xa, xb, a.batch, bb)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed)).to(dtype)
    g = grads_of(out, cot, [xa, xb])
    return {"xa": xa.detach(), "xb": xb.detach(), "batch_a": a.batch, "batch_b": bb, "cot": cot,
            "out": out.detach(), "grad_xa": g[0], "grad_xb": g[1]}


def case_model_gp(model_mod, readout, block, seed, Din=9, De=3, B=8, graph_norm="_None"):
    torch.manual_seed(seed)
    b = make_molecule_batch(B, node_dim=Din, edge_dim=De, seed=seed, features="chem")
    m = model_mod.Model(Din, De, hid_dim_alpha=4, e_dim=32, out_dim=1, mol_block=block, message_steps=3,
                        mol_readout=readout, graph_norm=graph_norm, graph_do="_None()", end_do="_None()",
                        pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU", graph_res=True)
    m.eval()
    data = types.SimpleNamespace(x=b.x, edge_index=b.edge_index, edge_attr=b.edge_attr, batch=b.batch)
    out = m(data)
    loss = torch.nn.functional.mse_loss(out, b.y)
    g = torch.autograd.grad(loss, list(m.parameters()))
    return {"x": b.x, "edge_index": b.edge_index, "edge_attr": b.edge_attr, "batch": b.batch, "y": b.y,
            "cfg": {"readout": readout, "block": block, "Din": Din, "De": De, "e_dim": 32, "graph_norm": graph_norm},
            "state": {k: v.detach().clone() for k, v in m.state_dict().items()},
            "out": out.detach(), "loss": loss.detach(),
            "grad_params": {n: gg.clone() for (n, _), gg in zip(m.named_parameters(), g)}}


def case_model_ddi(model_mod, seed, Din=9, De=3, B=6):
    torch.manual_seed(seed)
    a = make_molecule_batch(B, node_dim=Din, edge_dim=De, seed=seed, features="chem", targets="binary")
    c = make_molecule_batch(B, node_dim=Din, edge_dim=De, seed=seed + 7, features="chem")
    m = model_mod.Model(Din, De, hid_dim_alpha=4, e_dim=32, out_dim=1, mol_block="_TripletMessage",
                        message_steps=3, mol_readout="Set2Set", graph_do="_None()", end_do="_None()",
                        pre_act="ReLU", graph_act="ReLU", flat_act="CELU", end_act="ReLU")
    m.eval()
    ns = lambda b: types.SimpleNamespace(x=b.x, edge_index=b.edge_index, edge_attr=b.edge_attr, batch=b.batch)
    out = m(ns(a), ns(c))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, a.y)
    g = torch.autograd.grad(loss, list(m.parameters()))
    pack = lambda b, p: {p + "x": b.x, p + "edge_index": b.edge_index, p + "edge_attr": b.edge_attr, p + "batch": b.batch}
    d = {"y": a.y, "cfg": {"Din": Din, "De": De, "e_dim": 32},
         "state": {k: v.detach().clone() for k, v in m.state_dict().items()},
         "out": out.detach(), "loss": loss.detach(),
         "grad_params": {n: gg.clone() for (n, _), gg in zip(m.named_parameters(), g)}}
    d.update(pack(a, "a_")); d.update(pack(c, "b_"))
    return d


def case_pool5(layer, C, seed, dtype, sizes=None):
    """GlobalPool5 incl. graphs with fewer than k=3 nodes and tied keys in the sort channel."""
    torch.manual_seed(seed)
    if sizes is None:
        b = make_molecule_batch(6, node_dim=C, edge_dim=3, seed=seed, features="normal")
        x, batch = b.x, b.batch
    else:
        batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        x = torch.randn(int(sum(sizes)), C)
        x[-4:-2, -1] = 0.5                       # a tie inside the last graph
    mod = layer.GlobalPool5()
    x = x.to(dtype).requires_grad_(True)
    out = mod(x, batch)
    cot = torch.randn(out.shape).to(dtype)
    g = grads_of(out, cot, [x])
    return {"x": x.detach(), "batch": batch, "cot": cot, "out": out.detach(), "grad_x": g[0]}


def case_gcn_layer(layer, C, seed, dtype, protein=False):
    torch.manual_seed(seed)
    if protein:
        b = make_protein_batch(3, node_dim=C, edge_dim=8, seed=seed, min_len=30, max_len=60)
        x = torch.randn(b.x.shape[0], C)
        ei = torch.cat([b.edge_index, torch.tensor([[0, 5, 5], [0, 5, 5]])], dim=1)      # explicit + duplicate self loops
    else:
        b = make_molecule_batch(4, node_dim=C, edge_dim=3, seed=seed, features="normal")
        x, ei = b.x, b.edge_index
    mod = layer._GCNConv(C, C, 3)
    with torch.no_grad():
        mod.conv.bias.uniform_(-0.1, 0.1)
    mod = mod.to(dtype)
    x = x.to(dtype).requires_grad_(True)
    out = mod(x, ei, None)
    cot = torch.randn(out.shape).to(dtype)
    params = list(mod.parameters())
    g = grads_of(out, cot, [x] + params)
    return {"x": x.detach(), "edge_index": ei, "cot": cot,
            "state": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "out": out.detach(), "grad_x": g[0],
            "grad_params": {n: gg for (n, _), gg in zip(mod.named_parameters(), g[1:])}}


def case_model_dti(model_mod, seed, Din=9, De=3, Pin=49, Pe=8, B=4):
    """The drug-target model with the reference's default protein block and readouts (src_2gi_dti_scr/run.py:18-25)."""
    torch.manual_seed(seed)
    a = make_molecule_batch(B, node_dim=Din, edge_dim=De, seed=seed, features="chem", targets="binary")
    p = make_protein_batch(B, node_dim=Pin, edge_dim=Pe, seed=seed + 3, min_len=40, max_len=80)
    m = model_mod.Model(Din, Pin, De, Pe, hid_dim_alpha=4, e_dim=32, out_dim=1, mol_block="_TripletMessage",
                        pro_block="_GCNConv", message_steps=3, mol_readout="GlobalPool5", pro_readout="GlobalPool5",
                        graph_do="_None()", end_do="_None()", pre_act="ReLU", graph_act="LeakyReLU", flat_act="CELU", end_act="ReLU")
    # graph_act must not produce exact ties in the sort-pool key (ReLU zeros): PyG sorts with an unstable torch.sort, so
    # which of several tied rows is kept is not defined by the reference
    m.eval()
    ns = lambda b: types.SimpleNamespace(x=b.x, edge_index=b.edge_index, edge_attr=b.edge_attr, batch=b.batch)
    out = m(ns(a), ns(p))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, a.y)
    g = torch.autograd.grad(loss, list(m.parameters()))
    pack = lambda b, q: {q + "x": b.x, q + "edge_index": b.edge_index, q + "edge_attr": b.edge_attr, q + "batch": b.batch}
    d = {"y": a.y, "cfg": {"Din": Din, "De": De, "Pin": Pin, "Pe": Pe, "e_dim": 32},
         "state": {k: v.detach().clone() for k, v in m.state_dict().items()},
         "out": out.detach(), "loss": loss.detach(),
         "grad_params": {n: gg.clone() for (n, _), gg in zip(m.named_parameters(), g)}}
    d.update(pack(a, "a_")); d.update(pack(p, "b_"))
    return d


def case_nnconv_layer(layer, C, De, seed, dtype):
    torch.manual_seed(seed)
    b = make_molecule_batch(5, node_dim=C, edge_dim=De, seed=seed, features="normal")
    x, ei, ea, batch, N = add_edge_cases(b)                  # isolated node (mean over no edges -> root term only) + duplicates
    mod = layer._NNConv(C, C, De)
    with torch.no_grad():
        mod.conv.bias.uniform_(-0.1, 0.1)
    mod = mod.to(dtype)
    x = x.to(dtype).requires_grad_(True)
    out = mod(x, ei, ea.to(dtype))
    cot = torch.randn(out.shape).to(dtype)
    params = list(mod.parameters())
    g = grads_of(out, cot, [x] + params)
    return {"x": x.detach(), "edge_index": ei, "edge_attr": ea.to(dtype), "cot": cot,
            "state": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "out": out.detach(), "grad_x": g[0],
            "grad_params": {n: gg for (n, _), gg in zip(mod.named_parameters(), g[1:])}}


def main_nnconv():
    """`_NNConv` (src_1gp/layer.py:115-122), the reference's default mol_block, and the GP model with run.py's default
    block + readout (`_NNConv` + `GlobalPool5`, src_1gp/run.py:21,25)."""
    layer = load_ref("src_1gp", "layer")
    fx = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        fx[f"nnconv_C36_{tag}"] = case_nnconv_layer(layer, 36, 3, 1401, dtype)
        fx[f"nnconv_C20_De4_{tag}"] = case_nnconv_layer(layer, 20, 4, 1402, dtype)
        fx[f"block_nnconv_C36_{tag}"] = case_block(layer, "_NNConv", "_None", "ReLU", True, 36, 3, 1403, dtype, steps=2)
    for k in [k for k in fx if k.endswith("_f64")]:
        fx[k].pop("state", None)
        for name in ("edge_index", "batch", "cfg", "x", "cot", "coth", "edge_attr"):
            fx[k].pop(name, None)
    model_gp = load_ref("src_1gp", "model")
    fx["gp_nnconv_pool5"] = case_model_gp(model_gp, "GlobalPool5", "_NNConv", 1404)
    torch.save(fx, os.path.join(OUT, "nnconv.pt"))
    print("nnconv.pt", os.path.getsize(os.path.join(OUT, "nnconv.pt")) // 1024, "KiB")


def main_next():
    """Fixtures for the SURVEY.md §8(f) rows added after the first golden set (kept in their own file)."""
    layer = load_ref("src_1gp", "layer")
    fx = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        fx[f"pool5_C36_{tag}"] = case_pool5(layer, 36, 1301, dtype)
        fx[f"pool5_small_C30_{tag}"] = case_pool5(layer, 30, 1302, dtype, sizes=[1, 2, 3, 7, 40, 6])
        fx[f"gcn_C36_{tag}"] = case_gcn_layer(layer, 36, 1303, dtype)
        fx[f"gcn_protein_C30_{tag}"] = case_gcn_layer(layer, 30, 1304, dtype, protein=True)
        fx[f"block_gcn_C36_{tag}"] = case_block(layer, "_GCNConv", "_None", "ReLU", True, 36, 3, 1305, dtype, steps=2)
    for k in [k for k in fx if k.endswith("_f64")]:
        fx[k].pop("state", None)
        for name in ("edge_index", "batch", "cfg", "x", "cot", "coth", "edge_attr"):
            fx[k].pop(name, None)
    model_dti = load_ref("src_2gi_dti_scr", "model")
    fx["dti_gcn_pool5"] = case_model_dti(model_dti, 1306)
    torch.save(fx, os.path.join(OUT, "next.pt"))
    print("next.pt", os.path.getsize(os.path.join(OUT, "next.pt")) // 1024, "KiB")


def main():
    layer = load_ref("src_1gp", "layer")
    layer_ddi = load_ref("src_2gi_ddi", "layer")
    fx = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        fx[f"triplet_C36_{tag}"] = case_layer(layer, "triplet", 36, 3, 1234, dtype)
        fx[f"triplet_C60_{tag}"] = case_layer(layer, "triplet", 60, 4, 1235, dtype)
        fx[f"triplet_C15_edge_{tag}"] = case_layer(layer, "triplet", 15, 4, 1236, dtype, edge_cases=True)
        fx[f"light_C36_{tag}"] = case_layer(layer, "light", 36, 3, 1237, dtype)
        fx[f"light_C45_edge_{tag}"] = case_layer(layer, "light", 45, 4, 1238, dtype, edge_cases=True)
        fx[f"block_triplet_C36_{tag}"] = case_block(layer, "_TripletMessage", "_None", "CELU", True, 36, 3, 1239, dtype)
        fx[f"block_triplet_pn_C60_{tag}"] = case_block(layer, "_TripletMessage", "_PairNorm", "ReLU", True, 60, 4, 1240, dtype)
        fx[f"block_light_C30_{tag}"] = case_block(layer, "_TripletMessageLight", "_None", "LeakyReLU", False, 30, 4, 1241, dtype)
        fx[f"set2set_C36_{tag}"] = case_readout(layer, "set2set", 36, 1242, dtype)
        fx[f"set2set_C30_{tag}"] = case_readout(layer, "set2set", 30, 1243, dtype)
        fx[f"lapool_C36_{tag}"] = case_readout(layer, "lapool", 36, 1244, dtype)
        fx[f"lapool_C45_{tag}"] = case_readout(layer, "lapool", 45, 1245, dtype)
        fx[f"dotpool_ddi_C36_{tag}"] = case_dotpool(layer_ddi, 36, 1246, dtype)
        fx[f"dotpool_dti_C60_{tag}"] = case_dotpool(layer_ddi, 60, 1247, dtype, protein=True)
    # the f64 cases use the f32 inputs / state / cotangents upcast: store those only once
    for k in [k for k in fx if k.endswith("_f64")]:
        fx[k].pop("state", None)
        for name in ("edge_index", "batch", "batch_a", "batch_b", "cfg", "x", "xa", "xb", "cot", "coth", "edge_attr"):
            fx[k].pop(name, None)
    torch.save(fx, os.path.join(OUT, "layers.pt"))

    model_gp = load_ref("src_1gp", "model")
    mx = {
        "gp_set2set": case_model_gp(model_gp, "Set2Set", "_TripletMessage", 1234),
        "gp_lapool_light": case_model_gp(model_gp, "GlobalLAPool", "_TripletMessageLight", 1250, Din=15, De=4),
        "gp_set2set_pairnorm": case_model_gp(model_gp, "Set2Set", "_TripletMessage", 1251, graph_norm="_PairNorm"),
    }
    model_ddi = load_ref("src_2gi_ddi", "model")
    mx["ddi_set2set"] = case_model_ddi(model_ddi, 1252)
    torch.save(mx, os.path.join(OUT, "models.pt"))
    for f in ("layers.pt", "models.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "next":
        main_next()
    elif len(sys.argv) > 1 and sys.argv[1] == "nnconv":
        main_nnconv()
    else:
        main()
