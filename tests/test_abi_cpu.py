"""CPU: the C-ABI library builds, loads and exports every symbol include/glam_b200.h declares; host logic."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from glam_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "glam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glam_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/glam_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from glam_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_functions()


def test_arity_matches_header():
    from glam_b200 import _lib
    src = open(os.path.join(ROOT, "include", "glam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\b(glam_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        n = 0 if args.strip() in ("", "void") else args.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), name


def test_abi_version_and_workspace_queries(lib):
    assert lib.glam_abi_version() == 14
    assert lib.glam_csr_workspace_bytes(1000, 2000) > 4 * (2 * 1001 + 4 * 2000)
    assert lib.glam_gemm_tn_workspace_bytes(100000, 36, 116) >= 36 * 116 * 4
    assert lib.glam_colsum_workspace_bytes(100000, 108) >= 108 * 4
    assert lib.glam_triplet_bwd_workspace_bytes(3, 36, 3) >= 3 * 108 * 4


def test_edge_tile_geometry(lib):
    """Host side of the windowed edge kernels: tile rows / counts (no GPU needed)."""
    for n in (1, 15, 16, 1000, 102400, 409600, 10_000_000):
        d, t = lib.glam_edge_tile_rows(n), lib.glam_edge_tile_count(n)
        assert 16 <= d <= 64 and t == (n + d - 1) // d
    # the bench shape: 58-row tiles, a whole number of tiles per resident CTA (444 = 148 SMs x 3, 296 = 148 x 2)
    assert lib.glam_edge_tile_rows(102400) == 58 and lib.glam_edge_tile_count(102400) == 1766
    assert lib.glam_build_edge_tiles(None, None, None, None, 0, 0, None, None, None) == 0          # empty graph: nothing to do
    assert lib.glam_build_edge_tiles(None, None, None, None, 10, 0, None, None, None) < 0           # null pointers are reported


def test_argument_errors_are_reported(lib):
    from glam_b200 import _lib
    rc = lib.glam_triplet_edge_fwd(None, 0, None, None, None, None, None, None, 10, 10, 9, 36, 3, 0.2, None, None, None)
    assert rc < 0 and b"heads" in lib.glam_last_error()
    with pytest.raises(_lib.GlamError):
        _lib.check(rc, "glam_triplet_edge_fwd")


def test_cpu_tensors_raise_instead_of_falling_back():
    from glam_b200 import layer, _lib
    m = layer.TripletMessage(12, 3)
    x = torch.randn(5, 12)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(_lib.GlamError):
        m(x, ei, torch.eye(3)[:2])


def test_state_dict_keys_match_reference(golden_models):
    from glam_b200 import model
    c = golden_models["gp_set2set"]
    m = model.ArchitectureGP(9, 3, e_dim=32, mol_readout="Set2Set", mol_block="_TripletMessage")
    assert list(m.state_dict().keys()) == list(c["state"].keys())
    for k, v in m.state_dict().items():
        assert v.shape == c["state"][k].shape, k
    d = golden_models["ddi_set2set"]
    m2 = model.ArchitectureDDI(9, 3, e_dim=32, mol_block="_TripletMessage", mol_readout="Set2Set")
    assert list(m2.state_dict().keys()) == list(d["state"].keys())
    c2 = golden_models["gp_lapool_light"]
    m3 = model.ArchitectureGP(15, 4, e_dim=32, mol_readout="GlobalLAPool", mol_block="_TripletMessageLight")
    assert list(m3.state_dict().keys()) == list(c2["state"].keys())


def test_seeded_init_matches_reference(golden_layers):
    from glam_b200 import layer
    from glam_b200.synth import make_molecule_batch
    torch.manual_seed(1234)
    make_molecule_batch(4, node_dim=36, edge_dim=3, seed=1234, features="normal")
    m = layer.TripletMessage(36, 3)
    for k in ("weight_node", "weight_edge", "weight_triplet_att", "weight_scale"):
        torch.testing.assert_close(m.state_dict()[k], golden_layers["triplet_C36_f32"]["state"][k], rtol=0, atol=0)


def test_name_string_construction():
    from glam_b200 import layer
    blk = layer.MessageBlock(12, 12, 3, norm="_PairNorm", dropout="Dropout(0.1)", conv="_TripletMessageLight", act="RReLU", res=1)
    assert isinstance(blk.dropout, torch.nn.Dropout) and abs(blk.dropout.p - 0.1) < 1e-9
    assert isinstance(blk.act, torch.nn.RReLU)
    assert isinstance(layer.LinearBlock(4, 4, dropout="_None()", act="_None").act, layer._None)
    nnb = layer.MessageBlock(12, 12, 3, conv="_NNConv")              # the reference's default block builds with the reference's names
    assert list(nnb.conv.state_dict().keys()) == ["conv.root", "conv.bias", "conv.nn.0.weight", "conv.nn.0.bias",
                                                  "conv.nn.2.weight", "conv.nn.2.bias"]
    with pytest.raises(NotImplementedError):
        layer.MessageBlock(12, 12, 3, conv="_GATConv")


def test_synth_batch_follows_reference_edge_order():
    from glam_b200.synth import make_molecule_batch, shard_by_graph
    b = make_molecule_batch(64, seed=7)
    src, dst = b.edge_index
    N = b.num_nodes
    key = src * N + dst
    assert torch.all(key[1:] > key[:-1])                      # (src,dst)-lexicographic, no duplicates
    assert torch.all(b.batch[1:] >= b.batch[:-1])
    assert torch.all(b.batch[src] == b.batch[dst])            # block diagonal
    rev = dst * N + src                                        # every bond in both directions
    assert torch.equal(torch.sort(rev).values, key)
    assert 20 < N / 64 < 30 and 1.9 < b.num_edges / N < 2.4
    parts = [shard_by_graph(b, r, 4) for r in range(4)]
    assert sum(p.num_graphs for p in parts) == 64 and sum(p.num_nodes for p in parts) == N
    assert sum(p.num_edges for p in parts) == b.num_edges
    for p in parts:
        assert int(p.edge_index.max()) < p.num_nodes and int(p.batch.max()) == p.num_graphs - 1
