"""CPU: the literal drop-in claim — the reference's own, unmodified `model.py` files build over `glam_b200.layer` injected as the
module `layer` (the namespace their `from layer import ...` / `exec(...)` resolve names in, src_1gp/model.py:3-5,41), with the
state_dict keys and shapes of the oracle's restatement.  Needs the upstream tree (/root/reference: present in the build
container, absent on the GPU box, where this file skips); the forward of the same wiring on the library is covered on the GPU
by glam_b200.model (`stack_steps=False` = the reference's own per-step loop, tests/test_gpu_fused.py)."""
import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="upstream tree not present")


def _load_reference_model(subdir):
    from glam_b200 import layer
    pyg_nn = types.ModuleType("torch_geometric.nn")
    for name in ("Set2Set", "GCNConv"):
        setattr(pyg_nn, name, getattr(layer, name))
    for name in ("GATConv", "global_max_pool", "global_add_pool", "global_mean_pool", "global_sort_pool"):
        setattr(pyg_nn, name, None)                   # imported by src_2gi_dti_scr/model.py:10-11, never used by Architecture
    pyg = types.ModuleType("torch_geometric")
    pyg.nn = pyg_nn
    saved = {k: sys.modules.get(k) for k in ("layer", "torch_geometric", "torch_geometric.nn")}
    sys.modules.update({"layer": layer, "torch_geometric": pyg, "torch_geometric.nn": pyg_nn})
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_model_{subdir}", os.path.join(REF, subdir, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _shapes(m):
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("block,readout", [("_TripletMessage", "Set2Set"), ("_TripletMessageLight", "GlobalLAPool"),
                                           ("_NNConv", "GlobalPool5")])
def test_reference_gp_model_builds_over_glam_layer(block, readout):
    from glam_b200 import layer, model
    from oracle import glam_oracle as O
    ref = _load_reference_model("src_1gp")
    kw = dict(hid_dim_alpha=4, e_dim=64, out_dim=1, mol_block=block, message_steps=3, mol_readout=readout)
    torch.manual_seed(0)
    m_ref = ref.Architecture(9, 3, **kw)
    assert isinstance(m_ref.mol_conv, layer.MessageBlock) and type(m_ref.mol_readout).__module__ == "glam_b200.layer"
    torch.manual_seed(0)
    m_ours = model.ArchitectureGP(9, 3, **kw)
    assert _shapes(m_ref) == _shapes(m_ours)
    if block != "_NNConv":                                                # the oracle restates the triplet path
        assert _shapes(m_ref) == _shapes(O.ArchitectureGP(9, 3, **kw))
    # seeded init is the same arithmetic in the same order: identical parameters
    for (k, a), (_, b) in zip(m_ref.state_dict().items(), m_ours.state_dict().items()):
        assert torch.equal(a, b), k
    # the reference forward calls mol_conv(x, edge_index, edge_attr, h=hm, batch=...) and mol_readout(x, batch)
    sig = inspect.signature(layer.MessageBlock.forward)
    assert list(sig.parameters)[:6] == ["self", "x", "edge_index", "edge_attr", "h", "batch"]
    assert list(inspect.signature(type(m_ref.mol_readout).forward).parameters)[:3] == ["self", "x", "batch"]


def test_reference_default_constructor_matches():
    """Architecture() with no arguments: the reference's defaults (_NNConv + GlobalPool5, src_1gp/model.py:24-33)."""
    from glam_b200 import model
    ref = _load_reference_model("src_1gp")
    torch.manual_seed(3)
    a = ref.Architecture()
    torch.manual_seed(3)
    b = model.ArchitectureGP()
    assert _shapes(a) == _shapes(b)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k


@pytest.mark.parametrize("subdir,cls", [("src_2gi_ddi", "ArchitectureDDI"), ("src_2gi_dti_scr", "ArchitectureDTI")])
def test_reference_pair_models_build_over_glam_layer(subdir, cls):
    from glam_b200 import model
    ref = _load_reference_model(subdir)
    torch.manual_seed(1)
    a = ref.Architecture()
    torch.manual_seed(1)
    b = getattr(model, cls)()
    assert _shapes(a) == _shapes(b)
