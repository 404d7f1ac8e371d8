"""CPU: the oracle restatement vs golden vectors produced by the reference's own layer.py/model.py."""
import pytest
import torch

from oracle import glam_oracle as O
from helpers import case, ns

TAGS = ["f32", "f64"]


def _tol(tag):
    return dict(rtol=2e-4, atol=2e-5) if tag == "f32" else dict(rtol=1e-10, atol=1e-11)


def _check_grads(named_params, grads, golden, tag):
    for (n, _), g in zip(named_params, grads):
        torch.testing.assert_close(g, golden[n], msg=lambda m, n=n: f"{n}: {m}", **_tol(tag))


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,C,De", [("triplet_C36", 36, 3), ("triplet_C60", 60, 4), ("triplet_C15_edge", 15, 4)])
def test_triplet_message(golden_layers, name, C, De, tag):
    c = case(golden_layers, f"{name}_{tag}")
    m = O.TripletMessage(C, De).to(c["x"].dtype)
    m.load_state_dict(c["state"])
    x = c["x"].clone().requires_grad_(True)
    out = m(x, c["edge_index"], c["edge_attr"])
    torch.testing.assert_close(out, c["out"], **_tol(tag))
    g = torch.autograd.grad((out * c["cot"]).sum(), [x] + list(m.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(m.named_parameters()), g[1:], c["grad_params"], tag)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,C,De", [("light_C36", 36, 3), ("light_C45_edge", 45, 4)])
def test_triplet_light(golden_layers, name, C, De, tag):
    c = case(golden_layers, f"{name}_{tag}")
    m = O.TripletMessageLight(C, De).to(c["x"].dtype)
    m.load_state_dict(c["state"])
    x = c["x"].clone().requires_grad_(True)
    out = m(x, c["edge_index"], c["edge_attr"])
    torch.testing.assert_close(out, c["out"], **_tol(tag))
    g = torch.autograd.grad((out * c["cot"]).sum(), [x] + list(m.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(m.named_parameters()), g[1:], c["grad_params"], tag)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["block_triplet_C36", "block_triplet_pn_C60", "block_light_C30"])
def test_message_block(golden_layers, name, tag):
    c = case(golden_layers, f"{name}_{tag}")
    cfg = golden_layers[f"{name}_f32"]["cfg"]
    blk = O.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"],
                         act=cfg["act"], res=cfg["res"]).to(c["x"].dtype)
    blk.load_state_dict(c["state"])
    x0 = c["x"].clone().requires_grad_(True)
    x, h = x0, None
    for _ in range(cfg["steps"]):
        x, h = blk(x, c["edge_index"], c["edge_attr"], h=h, batch=c["batch"])
    torch.testing.assert_close(x, c["out"], **_tol(tag))
    torch.testing.assert_close(h, c["h"], **_tol(tag))
    g = torch.autograd.grad((x * c["cot"]).sum() + (h * c["coth"]).sum(), [x0] + list(blk.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(blk.named_parameters()), g[1:], c["grad_params"], tag)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,kind,C", [("set2set_C36", "s2s", 36), ("set2set_C30", "s2s", 30),
                                         ("lapool_C36", "la", 36), ("lapool_C45", "la", 45)])
def test_readouts(golden_layers, name, kind, C, tag):
    c = case(golden_layers, f"{name}_{tag}")
    m = (O.Set2Set(C, 3) if kind == "s2s" else O.GlobalLAPool(C)).to(c["x"].dtype)
    m.load_state_dict(c["state"])
    x = c["x"].clone().requires_grad_(True)
    out = m(x, c["batch"])
    torch.testing.assert_close(out, c["out"], **_tol(tag))
    g = torch.autograd.grad((out * c["cot"]).sum(), [x] + list(m.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(m.named_parameters()), g[1:], c["grad_params"], tag)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name", ["dotpool_ddi_C36", "dotpool_dti_C60"])
def test_dot_pool(golden_layers, name, tag):
    c = case(golden_layers, f"{name}_{tag}")
    xa = c["xa"].clone().requires_grad_(True)
    xb = c["xb"].clone().requires_grad_(True)
    out = O.dot_and_global_pool2(xa, xb, c["batch_a"], c["batch_b"])
    torch.testing.assert_close(out, c["out"], **_tol(tag))
    ga, gb = torch.autograd.grad((out * c["cot"]).sum(), [xa, xb])
    torch.testing.assert_close(ga, c["grad_xa"], **_tol(tag))
    torch.testing.assert_close(gb, c["grad_xb"], **_tol(tag))


@pytest.mark.parametrize("name", ["gp_set2set", "gp_lapool_light", "gp_set2set_pairnorm"])
def test_model_gp(golden_models, name):
    c = golden_models[name]
    cfg = c["cfg"]
    m = O.ArchitectureGP(cfg["Din"], cfg["De"], hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1, mol_block=cfg["block"],
                         message_steps=3, mol_readout=cfg["readout"], graph_norm=cfg["graph_norm"],
                         pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU")
    assert list(m.state_dict().keys()) == list(c["state"].keys())
    m.load_state_dict(c["state"])
    m.eval()
    out = m(ns(c["x"], c["edge_index"], c["edge_attr"], c["batch"]))
    torch.testing.assert_close(out, c["out"], rtol=2e-4, atol=2e-5)
    loss = torch.nn.functional.mse_loss(out, c["y"])
    g = torch.autograd.grad(loss, list(m.parameters()))
    _check_grads(list(m.named_parameters()), g, c["grad_params"], "f32")


def test_model_ddi(golden_models):
    c = golden_models["ddi_set2set"]
    cfg = c["cfg"]
    m = O.ArchitecturePair(cfg["Din"], cfg["Din"], cfg["De"], cfg["De"], prefixes=("mol1", "mol2"), hid_dim_alpha=4,
                           e_dim=cfg["e_dim"], out_dim=1, graph_act="ReLU", pre_act="ReLU", flat_act="CELU",
                           end_act="ReLU")
    assert list(m.state_dict().keys()) == list(c["state"].keys())
    m.load_state_dict(c["state"])
    m.eval()
    out = m(ns(c["a_x"], c["a_edge_index"], c["a_edge_attr"], c["a_batch"]),
            ns(c["b_x"], c["b_edge_index"], c["b_edge_attr"], c["b_batch"]))
    torch.testing.assert_close(out, c["out"], rtol=2e-4, atol=2e-5)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, c["y"])
    g = torch.autograd.grad(loss, list(m.parameters()))
    _check_grads(list(m.named_parameters()), g, c["grad_params"], "f32")


def test_seeded_init_matches_reference(golden_layers):
    """Same parameter registration order + init calls => same seeded init as the reference (make_golden seeds
    torch before constructing the module; bias is then overwritten, so skip it)."""
    c = golden_layers["triplet_C36_f32"]
    from glam_b200.synth import make_molecule_batch
    torch.manual_seed(1234)
    make_molecule_batch(4, node_dim=36, edge_dim=3, seed=1234, features="normal")
    m = O.TripletMessage(36, 3)
    for k in ("weight_node", "weight_edge", "weight_triplet_att", "weight_scale"):
        torch.testing.assert_close(m.state_dict()[k], c["state"][k], rtol=0, atol=0)


# ---------------------------------------------------------------------------------------------- §8(f) rows
@pytest.mark.parametrize("name", ["pool5_C36", "pool5_small_C30"])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_pool5_matches_reference(golden_next, name, tag):
    from oracle import glam_oracle as O
    c = case(golden_next, f"{name}_{tag}")
    x = c["x"].clone().requires_grad_(True)
    out = O.global_pool5(x, c["batch"], int(c["batch"][-1]) + 1)
    tol = dict(rtol=1e-10, atol=1e-12) if tag == "f64" else dict(rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, c["out"], **tol)
    (out * c["cot"]).sum().backward()
    torch.testing.assert_close(x.grad, c["grad_x"], **tol)


@pytest.mark.parametrize("name", ["gcn_C36", "gcn_protein_C30"])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_gcn_matches_reference(golden_next, name, tag):
    from oracle import glam_oracle as O
    c = case(golden_next, f"{name}_{tag}")
    x = c["x"].clone().requires_grad_(True)
    w = c["state"]["conv.weight"].clone().requires_grad_(True)
    b = c["state"]["conv.bias"].clone().requires_grad_(True)
    out = O.gcn_conv(x, c["edge_index"], w, b)
    tol = dict(rtol=1e-10, atol=1e-12) if tag == "f64" else dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out, c["out"], **tol)
    (out * c["cot"]).sum().backward()
    torch.testing.assert_close(x.grad, c["grad_x"], **tol)
    torch.testing.assert_close(w.grad, c["grad_params"]["conv.weight"], **tol)
    torch.testing.assert_close(b.grad, c["grad_params"]["conv.bias"], **tol)


def test_oracle_dti_model_matches_reference(golden_next):
    """Drug-target wiring with the reference's default protein block and readouts (src_2gi_dti_scr/model.py, run.py:18-25)."""
    from oracle import glam_oracle as O
    c = golden_next["dti_gcn_pool5"]
    cfg = c["cfg"]
    m = O.ArchitecturePair(cfg["Din"], cfg["Pin"], cfg["De"], cfg["Pe"], prefixes=("mol", "pro"), hid_dim_alpha=4,
                           e_dim=cfg["e_dim"], out_dim=1, a_block="_TripletMessage", b_block="_GCNConv", message_steps=3,
                           a_readout="GlobalPool5", b_readout="GlobalPool5", graph_act="LeakyReLU", pre_act="ReLU",
                           flat_act="CELU", end_act="ReLU")
    m.load_state_dict(c["state"])
    da = ns(c["a_x"], c["a_edge_index"], c["a_edge_attr"], c["a_batch"])
    db = ns(c["b_x"], c["b_edge_index"], c["b_edge_attr"], c["b_batch"])
    out = m(da, db)
    torch.testing.assert_close(out, c["out"], rtol=1e-4, atol=1e-5)
    torch.nn.functional.binary_cross_entropy_with_logits(out, c["y"]).backward()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.grad, c["grad_params"][n], rtol=2e-3, atol=1e-6, msg=lambda s, n=n: f"{n}: {s}")


# ---------------------------------------------------------------------------------------------- _NNConv (run.py's default block)
@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,C,De", [("nnconv_C36", 36, 3), ("nnconv_C20_De4", 20, 4)])
def test_oracle_nnconv_matches_reference(golden_nnconv, name, C, De, tag):
    c = case(golden_nnconv, f"{name}_{tag}")
    m = O._NNConv(C, C, De).to(c["x"].dtype)
    assert list(m.state_dict().keys()) == list(case(golden_nnconv, f"{name}_f32")["state"].keys())
    m.load_state_dict(c["state"])
    x = c["x"].clone().requires_grad_(True)
    out = m(x, c["edge_index"], c["edge_attr"])
    torch.testing.assert_close(out, c["out"], **_tol(tag))
    g = torch.autograd.grad((out * c["cot"]).sum(), [x] + list(m.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(m.named_parameters()), g[1:], c["grad_params"], tag)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_nnconv_block_matches_reference(golden_nnconv, tag):
    c = case(golden_nnconv, f"block_nnconv_C36_{tag}")
    cfg = c["cfg"]
    blk = O.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"], act=cfg["act"],
                         res=cfg["res"]).to(c["x"].dtype)
    blk.load_state_dict(c["state"])
    x0 = c["x"].clone().requires_grad_(True)
    x, h = x0, None
    for _ in range(cfg["steps"]):
        x, h = blk(x, c["edge_index"], c["edge_attr"], h=h, batch=c["batch"])
    torch.testing.assert_close(x, c["out"], **_tol(tag))
    torch.testing.assert_close(h, c["h"], **_tol(tag))
    g = torch.autograd.grad((x * c["cot"]).sum() + (h * c["coth"]).sum(), [x0] + list(blk.parameters()))
    torch.testing.assert_close(g[0], c["grad_x"], **_tol(tag))
    _check_grads(list(blk.named_parameters()), g[1:], c["grad_params"], tag)


def test_oracle_model_gp_reference_defaults(golden_nnconv):
    """GLAM-GP with run.py's default block and readout (`_NNConv` + `GlobalPool5`, src_1gp/run.py:21,25)."""
    c = golden_nnconv["gp_nnconv_pool5"]
    cfg = c["cfg"]
    m = O.ArchitectureGP(cfg["Din"], cfg["De"], hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1, mol_block=cfg["block"],
                         message_steps=3, mol_readout=cfg["readout"], graph_norm=cfg["graph_norm"],
                         pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU")
    assert list(m.state_dict().keys()) == list(c["state"].keys())
    m.load_state_dict(c["state"])
    m.eval()
    out = m(ns(c["x"], c["edge_index"], c["edge_attr"], c["batch"]))
    torch.testing.assert_close(out, c["out"], rtol=2e-4, atol=2e-5)
    loss = torch.nn.functional.mse_loss(out, c["y"])
    g = torch.autograd.grad(loss, list(m.parameters()))
    _check_grads(list(m.named_parameters()), g, c["grad_params"], "f32")


def test_nnconv_grouped_projection_algebra():
    """The formulation the library uses for `_NNConv` (glam_b200/functional.py::NNConvFn) restated in torch on the CPU:
    with one-hot bond features the per-edge matrix nn(edge_attr) takes edge_dim distinct values Theta_t = nn(e_t), so
    message_e = (x @ [Theta_0 | .. | Theta_{De-1}])[src_e, t_e*C:(t_e+1)*C] and the layer is a typed gather-mean plus the
    root term.  Must equal the reference formulation (oracle.nn_conv, one [C,C] matrix per edge) to fp64 round-off,
    including isolated nodes and duplicate edges."""
    from glam_b200.synth import make_molecule_batch
    torch.manual_seed(3)
    C, De = 12, 4
    b = make_molecule_batch(6, node_dim=C, edge_dim=De, seed=9, features="normal")
    x = torch.cat([b.x, torch.randn(2, C)]).double()                       # two isolated nodes at the end
    ei = torch.cat([b.edge_index, b.edge_index[:, :3]], dim=1)             # duplicate edges
    ea = torch.cat([b.edge_attr, b.edge_attr[:3]]).double()
    m = O._NNConv(C, C, De).double()
    with torch.no_grad():
        m.conv.bias.uniform_(-0.1, 0.1)
    want = m(x, ei, ea)
    N, E = x.shape[0], ei.shape[1]
    theta = m.conv.nn(torch.eye(De, dtype=torch.float64)).view(De, C, C)
    y = x @ theta.permute(1, 0, 2).reshape(C, De * C)                      # grouped projection [N, De*C]
    t = ea.argmax(1)
    rows = y.view(N * De, C)[ei[0] * De + t]                               # typed gather
    deg = torch.zeros(N, dtype=torch.float64).index_add_(0, ei[1], torch.ones(E, dtype=torch.float64)).clamp(min=1)
    got = torch.zeros(N, C, dtype=torch.float64).index_add_(0, ei[1], rows) / deg[:, None] + x @ m.conv.root + m.conv.bias
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(got[-2:], (x @ m.conv.root + m.conv.bias)[-2:], rtol=0, atol=0)


def test_padding_with_dummy_graphs_leaves_the_scores_of_the_real_graphs_unchanged():
    """synth.pad_graph_batch (what engine.ScreenStep uses to replay ONE captured step on batches of varying size): the padded
    batch has exactly the requested shape, keeps the real graphs in front untouched, every padding edge stays inside its dummy
    graph — and the oracle's scores of the real graphs are the same numbers."""
    import pytest
    from glam_b200.synth import make_molecule_batch, pad_graph_batch
    from oracle import glam_oracle as O
    b = make_molecule_batch(23, seed=11)
    N, E, B = b.num_nodes, b.num_edges, b.num_graphs
    p = pad_graph_batch(b, N + 301, E + 500, B + 5)
    assert (p.num_nodes, p.num_edges, p.num_graphs) == (N + 301, E + 500, B + 5)
    assert torch.equal(p.x[:N], b.x) and torch.equal(p.edge_index[:, :E], b.edge_index) and torch.equal(p.edge_attr[:E], b.edge_attr)
    assert bool((p.batch[1:] >= p.batch[:-1]).all()) and int(p.batch[-1]) == B + 4
    assert torch.equal(p.batch[p.edge_index[0, E:]], p.batch[p.edge_index[1, E:]])
    assert bool((p.edge_attr[E:].sum(1) == 1).all()) and int(torch.bincount(p.batch)[B:].max()) <= 128
    torch.manual_seed(0)
    m = O.ArchitectureGP(9, 3, hid_dim_alpha=2, e_dim=16, out_dim=1, message_steps=2, mol_readout="Set2Set", graph_act="CELU",
                         graph_norm="_PairNorm").eval()
    with torch.no_grad():
        torch.testing.assert_close(m(p)[:B], m(b), rtol=1e-6, atol=1e-7)
    assert pad_graph_batch(b, N, E, B) is b
    with pytest.raises(ValueError):
        pad_graph_batch(b, N + 3000, E, B + 2)               # 3000 nodes do not fit two dummy graphs of <= 128
    with pytest.raises(ValueError):
        pad_graph_batch(b, N - 1, E, B)


def test_tile_aware_graph_order_is_a_permutation_that_fills_tiles_and_changes_no_score():
    """synth.tile_order / permute_graphs (collation-time order of the graphs of a batch for the fused kernels' 128-row tiles): a
    true permutation, fewer greedily packed tiles than the random order and close to the minimum, every graph carried over
    intact (features, bonds, bond types, target) — and the oracle's scores are the same numbers, permuted."""
    import numpy as np
    from glam_b200.synth import make_molecule_batch, molecule_sizes, permute_graphs, tile_order
    from oracle import glam_oracle as O

    def greedy_tiles(sz, cap=128, chunk=512):
        t = 0
        for c0 in range(0, len(sz), chunk):
            cur = None
            for s in sz[c0:c0 + chunk]:
                if cur is None or cur + s > cap:
                    t, cur = t + 1, s
                else:
                    cur += s
        return t
    rng = np.random.default_rng(0)
    sizes = molecule_sizes(rng, 4096)
    perm = tile_order(sizes)
    assert sorted(perm.tolist()) == list(range(4096))
    best = -(-int(sizes.sum()) // 128)
    assert greedy_tiles(sizes[perm]) <= 1.03 * best < greedy_tiles(sizes)
    # graphs over the cap and an edge cap that binds
    odd = np.array([200, 5, 128, 1, 127, 64, 64, 300])
    assert sorted(tile_order(odd).tolist()) == list(range(8))
    assert sorted(tile_order(np.full(10, 30), edges=np.full(10, 400)).tolist()) == list(range(10))
    # the library's host routine (glam_tile_order, what impl="auto" runs) and the numpy statement give the same permutation,
    # also when the edge cap binds or a graph exceeds it
    e = (sizes * 2.2).astype(np.int64)
    for kw in (dict(), dict(edges=e), dict(edges=e, cap_edges=150), dict(edges=e, cap_nodes=64, cap_edges=100)):
        assert np.array_equal(tile_order(sizes, impl="native", **kw), tile_order(sizes, impl="numpy", **kw)), kw
    assert np.array_equal(tile_order(odd, impl="native"), tile_order(odd, impl="numpy"))
    b = make_molecule_batch(40, seed=3)
    perm = tile_order(torch.bincount(b.batch).numpy())
    pb = permute_graphs(b, perm)
    assert bool((pb.batch[1:] >= pb.batch[:-1]).all()) and pb.num_graphs == 40 and pb.num_nodes == b.num_nodes
    for gn in range(40):
        go = int(perm[gn])
        assert torch.equal(b.x[b.batch == go], pb.x[pb.batch == gn]) and torch.equal(b.y[go], pb.y[gn])
        mo, mn = b.batch[b.edge_index[0]] == go, pb.batch[pb.edge_index[0]] == gn
        eo = b.edge_index[:, mo] - int(torch.nonzero(b.batch == go)[0])
        en = pb.edge_index[:, mn] - int(torch.nonzero(pb.batch == gn)[0])
        so, io = torch.sort(eo[0] * 1000 + eo[1])
        sn, jn = torch.sort(en[0] * 1000 + en[1])
        assert torch.equal(so, sn) and torch.equal(b.edge_attr[mo][io], pb.edge_attr[mn][jn])
    torch.manual_seed(0)
    m = O.ArchitectureGP(9, 3, hid_dim_alpha=2, e_dim=16, out_dim=1, message_steps=2, mol_readout="Set2Set", graph_act="CELU").eval()
    with torch.no_grad():
        torch.testing.assert_close(m(pb), m(b)[torch.as_tensor(perm)], rtol=1e-5, atol=1e-6)
