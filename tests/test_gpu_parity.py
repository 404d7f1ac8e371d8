"""GPU parity tests: the sm_100a kernels (through the C ABI / glam_b200.layer) against the CPU oracle and the
golden vectors produced by the reference's own code.

Tolerance (SURVEY.md §8c, fp32 mode): |ours - ref64| <= max(2*|ref32 - ref64|, 1e-5 + 1e-4*scale) on outputs and
gradients; CSR / indices bit-exact; run-to-run bitwise identical.
"""
import copy

import pytest
import torch

from helpers import case, ns, tol_check, tf32_emulated

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _load(mod, state):
    mod.load_state_dict(state)
    return mod.to(DEV)


def _check_param_grads(mod, g32, g64, tag):
    for n, p in mod.named_parameters():
        assert p.grad is not None, f"{tag}: no grad for {n}"
        tol_check(p.grad, g32[n], g64[n], f"{tag}.grad[{n}]")


# ---------------------------------------------------------------------------------------------- CSR
@pytest.mark.parametrize("kind", ["molecules", "random", "hub", "empty", "protein"])
def test_csr_bit_exact(kind):
    from glam_b200 import ops
    from glam_b200.synth import make_molecule_batch, make_protein_batch
    g = torch.Generator().manual_seed(5)
    if kind == "molecules":
        b = make_molecule_batch(300, seed=11)
        ei, N = b.edge_index, b.num_nodes
    elif kind == "protein":
        b = make_protein_batch(3, seed=3)
        ei, N = b.edge_index, b.num_nodes
    elif kind == "random":
        N = 1000
        ei = torch.randint(0, N, (2, 20000), generator=g)          # duplicates, self loops, unsorted
    elif kind == "hub":
        N = 500
        ei = torch.randint(0, N, (2, 5000), generator=g)
        ei[1, :3000] = 7                                            # in-degree 3000 > 32: multi-chunk paths
        ei[0, 1000:2500] = 9
    else:
        N, ei = 17, torch.zeros((2, 0), dtype=torch.int64)
    csr = ops.build_csr(ei.to(DEV), N)
    torch.cuda.synchronize()
    src, dst = ei[0], ei[1]
    perm = torch.argsort(dst, stable=True)
    assert torch.equal(csr["dst_perm"].cpu().long(), perm)
    assert torch.equal(csr["dst_src"].cpu().long(), src[perm])
    assert torch.equal(csr["dst_dst"].cpu().long(), dst[perm])
    rowptr = torch.zeros(N + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    assert torch.equal(csr["dst_rowptr"].cpu().long(), rowptr)
    sperm = torch.argsort(src, stable=True)
    srowptr = torch.zeros(N + 1, dtype=torch.long)
    srowptr[1:] = torch.cumsum(torch.bincount(src, minlength=N), 0)
    assert torch.equal(csr["src_rowptr"].cpu().long(), srowptr)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel())
    assert torch.equal(csr["src_pos"].cpu().long(), inv[sperm])
    assert torch.equal(csr["src_dst"].cpu().long(), dst[sperm])
    if kind == "molecules":
        # reference-ordered batches: stable sort by dst == sort by key dst*N+src (torch_sparse convention)
        assert torch.equal(perm, torch.argsort(dst * N + src))


def test_graph_ptr_and_gather():
    from glam_b200 import ops
    batch = torch.tensor([0, 0, 0, 2, 2, 5, 5, 5, 5], dtype=torch.int64)       # graphs 1, 3, 4, 6 are empty
    ptr = ops.graph_ptr(batch.to(DEV), 7).cpu()
    assert ptr.tolist() == [0, 3, 3, 5, 5, 5, 9, 9]
    x = torch.randn(50, 5)
    perm = torch.randperm(50)
    out = ops.gather_rows(x.to(DEV), perm.to(DEV).int()).cpu()
    assert torch.equal(out, x[perm])


# ---------------------------------------------------------------------------------------------- dense pieces
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (63, 36, 36), (1000, 116, 36), (777, 60, 180), (4096, 188, 60), (130, 270, 90),
                                   (5000, 36, 9), (3000, 60, 15), (2500, 12, 40), (4096, 144, 108), (300, 256, 72), (129, 100, 36), (40000, 144, 108), (30011, 200, 72)])
def test_gemm_variants(M, N, K, math_mode):
    from glam_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g), torch.randn(N, generator=g)
    ref = X.double() @ W.double() + b.double()
    Y = ops.gemm(X.to(DEV), W.to(DEV), bias=b.to(DEV)).cpu()
    tol_check(Y, (X @ W + b), ref, "gemm nn")
    Yt = ops.gemm(X.to(DEV), W.t().contiguous().to(DEV), transpose_w=True, bias=b.to(DEV)).cpu()
    tol_check(Yt, (X @ W + b), ref, "gemm nt")
    Yc = ops.gemm(X.to(DEV), W.to(DEV), bias=b.to(DEV), epilogue=ops.EPI_CELU).cpu()
    tol_check(Yc, torch.celu(X @ W + b), torch.celu(ref), "gemm celu")
    acc = torch.randn(M, N, generator=g)
    out = acc.clone().to(DEV)
    ops.gemm(X.to(DEV), W.to(DEV), epilogue=ops.EPI_ACCUM, out=out)
    tol_check(out.cpu(), acc + X @ W, acc.double() + X.double() @ W.double(), "gemm accum")
    aux = torch.celu(torch.randn(M, N, generator=g))
    Yg = ops.gemm(X.to(DEV), W.to(DEV), epilogue=ops.EPI_MUL_CELU_GRAD, aux=aux.to(DEV)).cpu()
    d = torch.where(aux > 0, torch.ones_like(aux), aux + 1)
    tol_check(Yg, (X @ W) * d, (X.double() @ W.double()) * d.double(), "gemm celu-grad")
    # weight / bias gradients
    A, B = torch.randn(M, K, generator=g), torch.randn(M, N, generator=g)
    tol_check(ops.gemm_tn(A.to(DEV), B.to(DEV)).cpu(), A.t() @ B, A.double().t() @ B.double(), "gemm_tn")
    tol_check(ops.colsum(B.to(DEV)).cpu(), B.sum(0), B.double().sum(0), "colsum")
    for tr in (False, True):
        for cs in (False, True):
            o, c = ops.gemm_tn_ex(A.to(DEV), B.to(DEV), transpose_out=tr, want_colsum=cs)
            r32, r64 = A.t() @ B, A.double().t() @ B.double()
            tol_check(o.cpu(), r32.t() if tr else r32, r64.t() if tr else r64, f"gemm_tn_ex tr={tr} cs={cs}")
            if cs:
                tol_check(c.cpu(), B.sum(0), B.double().sum(0), "gemm_tn_ex colsum")


def test_gemm_tn_long_and_deterministic(math_mode):
    from glam_b200 import ops
    g = torch.Generator().manual_seed(3)
    A, B = torch.randn(100_003, 36, generator=g).to(DEV), torch.randn(100_003, 116, generator=g).to(DEV)
    o1, o2 = ops.gemm_tn(A, B), ops.gemm_tn(A, B)
    assert torch.equal(o1, o2)
    tol_check(o1.cpu(), (A.t() @ B).cpu(), (A.double().t() @ B.double()).cpu(), "gemm_tn long", rtol=1e-4)
    for (ka, kb) in [(36, 116), (108, 36), (180, 60), (60, 188), (3, 3)]:
        A, B = torch.randn(50_001, ka, generator=g).to(DEV), torch.randn(50_001, kb, generator=g).to(DEV)
        (o1, c1), (o2, c2) = ops.gemm_tn_ex(A, B, want_colsum=True), ops.gemm_tn_ex(A, B, want_colsum=True)
        assert torch.equal(o1, o2) and torch.equal(c1, c2)                    # bitwise reproducible
        tol_check(o1.cpu(), (A.t() @ B).cpu(), (A.double().t() @ B.double()).cpu(), f"gemm_tn_ex {ka}x{kb}")
        tol_check(c1.cpu(), B.sum(0).cpu(), B.double().sum(0).cpu(), f"gemm_tn_ex colsum {ka}x{kb}")


def test_skinny_weight_gradients_with_riding_column_sums(math_mode):
    """Exact fp32 A^T B for narrow operands (input LinearBlock 9/15 features, attention-logit columns): every row count around
    the 64-row tile, both operand orders (column sums of the wide and of the narrow operand), strided views."""
    from glam_b200 import ops
    g = torch.Generator().manual_seed(11)
    for M in (1, 63, 64, 65, 1000, 50_001):
        for (ka, kb) in [(9, 36), (36, 9), (15, 36), (36, 6), (63, 16), (64, 8), (3, 3), (64, 16), (16, 64)]:
            A, B = torch.randn(M, ka, generator=g).to(DEV), torch.randn(M, kb, generator=g).to(DEV)
            for tr in (False, True):
                o, c = ops.gemm_tn_ex(A, B, transpose_out=tr, want_colsum=True)
                o2, c2 = ops.gemm_tn_ex(A, B, transpose_out=tr, want_colsum=True)
                assert torch.equal(o, o2) and torch.equal(c, c2)
                r = (A.double().t() @ B.double())
                r = r.t() if tr else r
                assert torch.allclose(o.double(), r, rtol=1e-5, atol=1e-4 * max(1.0, M ** 0.5) * 1e-1), (M, ka, kb, tr)
                assert torch.allclose(c.double(), B.double().sum(0), rtol=1e-5, atol=1e-5 * max(1.0, M ** 0.5)), (M, ka, kb)
    wide = torch.randn(5000, 116, generator=g).to(DEV)
    X = torch.randn(5000, 36, generator=g).to(DEV)
    o, _ = ops.gemm_tn_ex(X, wide[:, 108:114])                                # the logit columns of g_xpe, in place
    assert torch.allclose(o.double(), X.double().t() @ wide[:, 108:114].double(), rtol=1e-5, atol=1e-3)


# ---------------------------------------------------------------------------------------------- golden layer cases
@pytest.mark.parametrize("name,C,De", [("triplet_C36", 36, 3), ("triplet_C60", 60, 4), ("triplet_C15_edge", 15, 4)])
def test_triplet_message_golden(golden_layers, name, C, De, math_mode):
    from glam_b200 import layer
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    m = _load(layer.TripletMessage(C, De), c32["state"])
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(m, c32["grad_params"], c64["grad_params"], name)


@pytest.mark.parametrize("name,C,De", [("light_C36", 36, 3), ("light_C45_edge", 45, 4)])
def test_triplet_light_golden(golden_layers, name, C, De, math_mode):
    from glam_b200 import layer
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    m = _load(layer.TripletMessageLight(C, De), c32["state"])
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(m, c32["grad_params"], c64["grad_params"], name)


@pytest.mark.parametrize("name", ["block_triplet_C36", "block_triplet_pn_C60", "block_light_C30"])
def test_message_block_golden(golden_layers, name, math_mode):
    from glam_b200 import layer
    if math_mode == "tf32" and "_pn_" in name:
        pytest.skip("PairNorm gradients are ill-conditioned w.r.t. TF32 operands: see test_pairnorm_block_tf32_vs_tf32_oracle")
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    cfg = c32["cfg"]
    blk = _load(layer.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"],
                                   act=cfg["act"], res=cfg["res"]), c32["state"])
    x0 = c32["x"].to(DEV).requires_grad_(True)
    ei, ea, batch = c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV), c32["batch"].to(DEV)
    x, h = x0, None
    for _ in range(cfg["steps"]):
        x, h = blk(x, ei, ea, h=h, batch=batch)
    tol_check(x, c32["out"], c64["out"], f"{name}.out")
    tol_check(h, c32["h"], c64["h"], f"{name}.h")
    ((x * c32["cot"].to(DEV)).sum() + (h * c32["coth"].to(DEV)).sum()).backward()
    tol_check(x0.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(blk, c32["grad_params"], c64["grad_params"], name)


@pytest.mark.parametrize("name,kind,C", [("set2set_C36", "s2s", 36), ("set2set_C30", "s2s", 30),
                                         ("lapool_C36", "la", 36), ("lapool_C45", "la", 45)])
def test_readouts_golden(golden_layers, name, kind, C, math_mode):
    from glam_b200 import layer
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    m = _load(layer.Set2Set(C, 3) if kind == "s2s" else layer.GlobalLAPool(C), c32["state"])
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["batch"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(m, c32["grad_params"], c64["grad_params"], name)


@pytest.mark.parametrize("name", ["dotpool_ddi_C36", "dotpool_dti_C60"])
def test_dot_pool_golden(golden_layers, name, math_mode):
    """fp32 mode: exact fp32 products on the CUDA cores; tf32 mode: S = Xa Xb^T on the tcgen05 tensor cores (TF32 operands), the
    mean and the column sums backward needs exact in both."""
    from glam_b200 import layer
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    xa = c32["xa"].to(DEV).requires_grad_(True)
    xb = c32["xb"].to(DEV).requires_grad_(True)
    out = layer.dot_and_global_pool2(xa, xb, c32["batch_a"].to(DEV), c32["batch_b"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(xa.grad, c32["grad_xa"], c64["grad_xa"], f"{name}.grad_xa")
    tol_check(xb.grad, c32["grad_xb"], c64["grad_xb"], f"{name}.grad_xb")


@pytest.mark.parametrize("C", [32, 36, 48, 64])
def test_dot_pool_warp_per_pair_matches_the_cta_per_pair_kernel(C):
    """glam_pair_dot_pool_fwd_small (drug-drug sized pairs: a warp per pair) against glam_pair_dot_pool_fwd on the same pairs —
    small graphs, an empty graph on either side, graphs over 32 rows (the tiled loop), a shared second side (idx_b): max and
    arg-max bitwise (same fmaf chains, same tie rule), column sums bitwise (same row order), mean to fp32 rounding."""
    from glam_b200 import ops
    g = torch.Generator().manual_seed(C)
    na = torch.tensor([25, 1, 32, 0, 7, 33, 70, 12, 31, 5, 29, 40])
    nb = torch.tensor([25, 32, 1, 9, 0, 20, 45, 33, 31, 64, 3, 2])
    P = na.numel()
    ptr_a = torch.zeros(P + 1, dtype=torch.int32); ptr_a[1:] = torch.cumsum(na, 0)
    ptr_b = torch.zeros(P + 1, dtype=torch.int32); ptr_b[1:] = torch.cumsum(nb, 0)
    xa = torch.randn(int(na.sum()), C, generator=g).to(DEV)
    xb = torch.randn(int(nb.sum()), C, generator=g).to(DEV)
    xb[ptr_b[0]:ptr_b[1]] = xa[ptr_a[0]:ptr_a[1]][torch.randperm(25, generator=g)]   # ties in value are still unlikely; equal rows stress the arg-max
    pa, pb = ptr_a.to(DEV), ptr_b.to(DEV)

    def run(entry, idx=None):
        out = torch.full((P, 2), float("nan"), device=DEV); am = torch.full((P, 2), -7, dtype=torch.int32, device=DEV)
        sa = torch.full((P, C), float("nan"), device=DEV); sb = torch.full((P, C), float("nan"), device=DEV)
        args = [ops._p(xa), ops._p(xb), ops._p(pa), ops._p(pb)] + ([ops._p(idx)] if entry != "glam_pair_dot_pool_fwd" else []) + \
               [P, C, ops._p(out), ops._p(am), ops._p(sa), ops._p(sb), ops._stream(xa)]
        ops._call(entry, *args)
        torch.cuda.synchronize()
        return out, am, sa, sb
    o1, a1, s1, t1 = run("glam_pair_dot_pool_fwd")
    o2, a2, s2, t2 = run("glam_pair_dot_pool_fwd_small", None)
    assert torch.equal(a1, a2) and torch.equal(o1[:, 0], o2[:, 0]) and torch.equal(s1, s2) and torch.equal(t1, t2)
    torch.testing.assert_close(o2[:, 1], o1[:, 1], rtol=1e-5, atol=1e-6)
    idx = torch.tensor([3, 3, 0, 1, 5, 5, 6, 2, 8, 9, 0, 11], dtype=torch.int32, device=DEV)
    o3, a3, s3, t3 = run("glam_pair_dot_pool_fwd_idx", idx)
    o4, a4, s4, t4 = run("glam_pair_dot_pool_fwd_small", idx)
    assert torch.equal(a3, a4) and torch.equal(o3[:, 0], o4[:, 0]) and torch.equal(s3, s4) and torch.equal(t3, t4)
    torch.testing.assert_close(o4[:, 1], o3[:, 1], rtol=1e-5, atol=1e-6)


def test_pairnorm_block_tf32_vs_tf32_oracle(golden_layers):
    """With PairNorm in the block the gradients amplify operand rounding ~300x (the fp64 oracle with TF32-truncated
    matmul operands is 3e-1 away from the exact one on grad_x).  The tensor-core path must agree with THAT oracle:
    the deviation is the operand format (what torch 1.10 does by default on Ampere+), not the kernels."""
    from glam_b200 import layer, _lib
    from oracle import glam_oracle as O
    name = "block_triplet_pn_C60"
    c32, c64 = golden_layers[f"{name}_f32"], case(golden_layers, f"{name}_f64")
    cfg = c32["cfg"]
    O.MM_OPERAND_HOOK = O.tf32_truncate
    try:
        ob = O.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"], act=cfg["act"],
                            res=cfg["res"]).double()
        ob.load_state_dict(c64["state"])
        x0 = c64["x"].clone().requires_grad_(True)
        x, h = x0, None
        for _ in range(cfg["steps"]):
            x, h = ob(x, c32["edge_index"], c64["edge_attr"], h=h, batch=c32["batch"])
        g_ref = torch.autograd.grad((x * c64["cot"]).sum() + (h * c64["coth"]).sum(), [x0] + list(ob.parameters()))
    finally:
        O.MM_OPERAND_HOOK = None
    _lib.set_math_mode("tf32")
    blk = _load(layer.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"],
                                   act=cfg["act"], res=cfg["res"]), c32["state"])
    xd = c32["x"].to(DEV).requires_grad_(True)
    y, hh = xd, None
    for _ in range(cfg["steps"]):
        y, hh = blk(y, c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV), h=hh, batch=c32["batch"].to(DEV))
    ((y * c32["cot"].to(DEV)).sum() + (hh * c32["coth"].to(DEV)).sum()).backward()
    rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(y, x.detach()) < 2e-3
    assert rel(xd.grad, g_ref[0]) < 5e-2, rel(xd.grad, g_ref[0])
    for (n, p), g in zip(blk.named_parameters(), g_ref[1:]):
        assert rel(p.grad, g) < 5e-2, (n, rel(p.grad, g))


# ---------------------------------------------------------------------------------------------- golden models
@pytest.mark.parametrize("name", ["gp_set2set", "gp_lapool_light", "gp_set2set_pairnorm"])
def test_model_gp_golden(golden_models, name, math_mode):
    from glam_b200 import model
    if math_mode == "tf32" and "pairnorm" in name:
        pytest.skip("PairNorm gradients are ill-conditioned w.r.t. TF32 operands: see test_pairnorm_block_tf32_vs_tf32_oracle")
    from oracle import glam_oracle as O
    c = golden_models[name]
    cfg = c["cfg"]
    kw = dict(hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1, mol_block=cfg["block"], message_steps=3,
              mol_readout=cfg["readout"], graph_norm=cfg["graph_norm"], pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU")
    m = model.ArchitectureGP(cfg["Din"], cfg["De"], graph_do="_None()", end_do="_None()", **kw)
    m = _load(m, c["state"]).eval()
    o64 = O.ArchitectureGP(cfg["Din"], cfg["De"], **kw)
    o64.load_state_dict(c["state"])
    o64 = o64.double().eval()
    d64 = ns(c["x"].double(), c["edge_index"], c["edge_attr"].double(), c["batch"])
    out64 = o64(d64)
    loss64 = torch.nn.functional.mse_loss(out64, c["y"].double())
    g64 = dict(zip([n for n, _ in o64.named_parameters()], torch.autograd.grad(loss64, list(o64.parameters()))))
    emu = tf32_emulated(lambda: torch.autograd.grad(torch.nn.functional.mse_loss(o64(d64), c["y"].double()), list(o64.parameters())))
    gemu = dict(zip([n for n, _ in o64.named_parameters()], emu))
    data = ns(c["x"].to(DEV), c["edge_index"].to(DEV), c["edge_attr"].to(DEV), c["batch"].to(DEV))
    out = m(data)
    tol_check(out, c["out"], out64, f"{name}.out", rtol=2e-4)
    loss = torch.nn.functional.mse_loss(out, c["y"].to(DEV))
    loss.backward()
    for n, p in m.named_parameters():
        tol_check(p.grad, c["grad_params"][n], g64[n], f"{name}.grad[{n}]", rtol=2e-4, emu64=gemu[n])


def test_model_ddi_golden(golden_models, math_mode):
    from glam_b200 import model
    from oracle import glam_oracle as O
    c = golden_models["ddi_set2set"]
    cfg = c["cfg"]
    m = model.ArchitectureDDI(cfg["Din"], cfg["De"], hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1, graph_do="_None()",
                              end_do="_None()", pre_act="ReLU", graph_act="ReLU", flat_act="CELU", end_act="ReLU",
                              mol_block="_TripletMessage", mol_readout="Set2Set")
    m = _load(m, c["state"]).eval()
    o64 = O.ArchitecturePair(cfg["Din"], cfg["Din"], cfg["De"], cfg["De"], prefixes=("mol1", "mol2"), hid_dim_alpha=4,
                             e_dim=cfg["e_dim"], out_dim=1, graph_act="ReLU", pre_act="ReLU", flat_act="CELU", end_act="ReLU")
    o64.load_state_dict(c["state"])
    o64 = o64.double().eval()
    a64 = ns(c["a_x"].double(), c["a_edge_index"], c["a_edge_attr"].double(), c["a_batch"])
    b64 = ns(c["b_x"].double(), c["b_edge_index"], c["b_edge_attr"].double(), c["b_batch"])
    out64 = o64(a64, b64)
    loss64 = torch.nn.functional.binary_cross_entropy_with_logits(out64, c["y"].double())
    g64 = dict(zip([n for n, _ in o64.named_parameters()], torch.autograd.grad(loss64, list(o64.parameters()))))
    emu = tf32_emulated(lambda: torch.autograd.grad(
        torch.nn.functional.binary_cross_entropy_with_logits(o64(a64, b64), c["y"].double()), list(o64.parameters())))
    gemu = dict(zip([n for n, _ in o64.named_parameters()], emu))
    a = ns(*[c[k].to(DEV) for k in ("a_x", "a_edge_index", "a_edge_attr", "a_batch")])
    b = ns(*[c[k].to(DEV) for k in ("b_x", "b_edge_index", "b_edge_attr", "b_batch")])
    out = m(a, b)
    tol_check(out, c["out"], out64, "ddi.out", rtol=2e-4)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, c["y"].to(DEV))
    loss.backward()
    # TF32 mode: S.max() of the dot-pool routes its gradient to one (a,b) entry; operand rounding can pick another of
    # several near-tied entries, which moves individual gradient tensors by up to ~10 % (a discontinuity of the
    # reference's own formulation, src_2gi_ddi/layer.py:282) -> 2e-1 of scale there, and the direction must agree.
    loose = 2e-1 if math_mode == "tf32" else 2e-4
    for n, p in m.named_parameters():
        tol_check(p.grad, c["grad_params"][n], g64[n], f"ddi.grad[{n}]", rtol=loose, emu64=gemu[n])
    flat = torch.cat([p.grad.flatten().double().cpu() for p in m.parameters()])
    ref = torch.cat([g64[n].flatten() for n, _ in m.named_parameters()])
    assert torch.nn.functional.cosine_similarity(flat, ref, dim=0) > 0.9995


# ---------------------------------------------------------------------------------------------- oracle on seeded inputs
def _gp_pair(Din, De, readout, block, act="CELU"):
    from glam_b200 import model
    from oracle import glam_oracle as O
    kw = dict(hid_dim_alpha=4, e_dim=128, out_dim=1, mol_block=block, message_steps=3, mol_readout=readout,
              pre_act="ReLU", graph_act=act, flat_act="ReLU")
    torch.manual_seed(99)
    o = O.ArchitectureGP(Din, De, **kw).eval()
    m = model.ArchitectureGP(Din, De, graph_do="_None()", end_do="_None()", **kw)
    m.load_state_dict(o.state_dict())
    return m.to(DEV).eval(), o


@pytest.mark.parametrize("Din,De,readout,block,B", [(9, 3, "Set2Set", "_TripletMessage", 128),
                                                   (15, 4, "GlobalLAPool", "_TripletMessage", 96),
                                                   (15, 4, "Set2Set", "_TripletMessageLight", 64)])
def test_gp_training_step_vs_oracle(Din, De, readout, block, B, math_mode):
    """BASELINE config 1 shape (128 ESOL-like molecules): forward + backward against the fp32 and fp64 oracle."""
    from glam_b200.synth import make_molecule_batch
    m, o32 = _gp_pair(Din, De, readout, block)
    o64 = copy.deepcopy(o32).double()
    b = make_molecule_batch(B, node_dim=Din, edge_dim=De, seed=2024)
    d32 = ns(b.x, b.edge_index, b.edge_attr, b.batch)
    d64 = ns(b.x.double(), b.edge_index, b.edge_attr.double(), b.batch)
    out32, out64 = o32(d32), o64(d64)
    l32 = torch.nn.functional.mse_loss(out32, b.y)
    l64 = torch.nn.functional.mse_loss(out64, b.y.double())
    g32 = torch.autograd.grad(l32, list(o32.parameters()))
    g64 = torch.autograd.grad(l64, list(o64.parameters()))
    gemu = tf32_emulated(lambda: torch.autograd.grad(torch.nn.functional.mse_loss(o64(d64), b.y.double()), list(o64.parameters())))
    bd = b.to(DEV)
    out = m(bd)
    tol_check(out, out32, out64, "gp.out", rtol=2e-4)
    torch.nn.functional.mse_loss(out, bd.y).backward()
    for (n, p), a, c, e in zip(m.named_parameters(), g32, g64, gemu):
        tol_check(p.grad, a, c, f"gp.grad[{n}]", rtol=2e-4, emu64=e)


def test_dti_shaped_pair_vs_oracle(math_mode):
    """BASELINE config 4 shape: ligand graph + protein contact-map graph (hundreds of residues, De=8, duplicate edges)."""
    from glam_b200 import model
    from glam_b200.synth import make_molecule_batch, make_protein_batch
    from oracle import glam_oracle as O
    torch.manual_seed(5)
    kw = dict(hid_dim_alpha=2, e_dim=64, out_dim=2, message_steps=2)
    o32 = O.ArchitecturePair(15, 49, 4, 8, prefixes=("mol", "pro"), graph_act="CELU", **kw).eval()
    m = model.ArchitectureDTI(15, 49, 4, 8, graph_do="_None()", end_do="_None()", pre_act="ReLU", graph_act="CELU",
                              flat_act="ReLU", end_act="ReLU", mol_block="_TripletMessage", pro_block="_TripletMessage",
                              mol_readout="Set2Set", pro_readout="Set2Set", **kw)
    m.load_state_dict(o32.state_dict())
    m = m.to(DEV).eval()
    o64 = copy.deepcopy(o32).double()
    a = make_molecule_batch(6, node_dim=15, edge_dim=4, seed=8)
    p = make_protein_batch(6, seed=9, min_len=150, max_len=320)
    y = torch.randint(0, 2, (6,))
    out32 = o32(ns(a.x, a.edge_index, a.edge_attr, a.batch), ns(p.x, p.edge_index, p.edge_attr, p.batch))
    out64 = o64(ns(a.x.double(), a.edge_index, a.edge_attr.double(), a.batch),
                ns(p.x.double(), p.edge_index, p.edge_attr.double(), p.batch))
    g32 = torch.autograd.grad(torch.nn.functional.cross_entropy(out32, y), list(o32.parameters()))
    g64 = torch.autograd.grad(torch.nn.functional.cross_entropy(out64, y), list(o64.parameters()))
    gemu = tf32_emulated(lambda: torch.autograd.grad(torch.nn.functional.cross_entropy(
        o64(ns(a.x.double(), a.edge_index, a.edge_attr.double(), a.batch),
            ns(p.x.double(), p.edge_index, p.edge_attr.double(), p.batch)), y), list(o64.parameters())))
    out = m(a.to(DEV), p.to(DEV))
    tol_check(out, out32, out64, "dti.out", rtol=2e-4)
    torch.nn.functional.cross_entropy(out, y.to(DEV)).backward()
    for (n, prm), g_a, g_c, g_e in zip(m.named_parameters(), g32, g64, gemu):
        tol_check(prm.grad, g_a, g_c, f"dti.grad[{n}]", rtol=2e-4, emu64=g_e)


# ---------------------------------------------------------------------------------------------- properties
def test_hub_node_multichunk_softmax(math_mode):
    """A destination with in-degree > 32 exercises the multi-chunk softmax path; isolated nodes give `bias`."""
    from glam_b200 import layer
    from oracle import glam_oracle as O
    torch.manual_seed(1)
    N, C, De = 300, 36, 3
    g = torch.Generator().manual_seed(2)
    ei = torch.randint(0, N - 10, (2, 1500), generator=g)
    ei[1, :200] = 5
    ei[1, 200:240] = 6
    ea = torch.eye(De)[torch.randint(0, De, (1500,), generator=g)] * torch.rand(1500, 1, generator=g)
    x = torch.randn(N, C, generator=g)
    o = O.TripletMessage(C, De)
    with torch.no_grad():
        o.bias.uniform_(-1, 1)
    m = layer.TripletMessage(C, De)
    m.load_state_dict(o.state_dict())
    m = m.to(DEV)
    x32 = x.clone().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    o64 = copy.deepcopy(o).double()
    out32, out64 = o(x32, ei, ea), o64(x64, ei, ea.double())
    cot = torch.randn(N, C, generator=g)
    g32 = torch.autograd.grad((out32 * cot).sum(), [x32] + list(o.parameters()))
    g64 = torch.autograd.grad((out64 * cot.double()).sum(), [x64] + list(o64.parameters()))
    xd = x.to(DEV).requires_grad_(True)
    out = m(xd, ei.to(DEV), ea.to(DEV))
    tol_check(out, out32, out64, "hub.out")
    assert torch.equal(out[N - 5:].detach().cpu(), o.bias.detach().expand(5, C))      # isolated nodes -> bias exactly
    (out * cot.to(DEV)).sum().backward()
    tol_check(xd.grad, g32[0], g64[0], "hub.grad_x")
    for (n, p), a, c in zip(m.named_parameters(), g32[1:], g64[1:]):
        tol_check(p.grad, a, c, f"hub.grad[{n}]")


def test_edge_permutation_invariance_and_determinism(math_mode):
    from glam_b200 import layer, graph
    from glam_b200.synth import make_molecule_batch
    torch.manual_seed(3)
    b = make_molecule_batch(200, node_dim=36, edge_dim=3, seed=77, features="normal").to(DEV)
    m = layer.TripletMessage(36, 3).to(DEV)
    with torch.no_grad():
        out1 = m(b.x, b.edge_index, b.edge_attr)
        graph.clear_caches()
        out2 = m(b.x, b.edge_index.clone(), b.edge_attr.clone())
        assert torch.equal(out1, out2)                                       # bitwise reproducible
        perm = torch.randperm(b.num_edges, device=DEV)
        out3 = m(b.x, b.edge_index[:, perm].contiguous(), b.edge_attr[perm].contiguous())
    # summation order changes: fp32 round-off, which a TF32 operand of the next projection may amplify to ~1e-3
    tol = dict(rtol=1e-5, atol=1e-6) if math_mode == "fp32" else dict(rtol=5e-3, atol=5e-4)
    torch.testing.assert_close(out3, out1, **tol)


def test_batch_of_one_equals_unbatched():
    from glam_b200 import layer
    from glam_b200.synth import make_molecule_batch, shard_by_graph
    torch.manual_seed(4)
    b = make_molecule_batch(8, node_dim=36, edge_dim=3, seed=5, features="normal")
    blk = layer.MessageBlock(36, 36, 3, norm="_None", dropout="_None()", conv="_TripletMessage", act="ReLU").to(DEV).eval()
    ro = layer.Set2Set(36, 3).to(DEV)
    bd = b.to(DEV)
    with torch.no_grad():
        x, h = blk(bd.x, bd.edge_index, bd.edge_attr, batch=bd.batch)
        full = ro(x, bd.batch)
        for gidx in (0, 3, 7):
            one = shard_by_graph(b, gidx, 8).to(DEV)
            x1, _ = blk(one.x, one.edge_index, one.edge_attr, batch=one.batch)
            r1 = ro(x1, one.batch)
            torch.testing.assert_close(r1[0], full[gidx], rtol=1e-5, atol=1e-6)


def test_full_size_properties():
    """BASELINE config 2 size (4096 graphs): attention rows sum to 1 per (destination, head) and the aggregate
    conserves mass: sum_i agg[i] == sum_e alpha_e * e_ij (.) x_j (checksum of checksums)."""
    from glam_b200 import layer, graph, ops
    from glam_b200.synth import make_molecule_batch
    torch.manual_seed(6)
    b = make_molecule_batch(4096, node_dim=36, edge_dim=3, seed=1234, features="normal").to(DEV)
    m = layer.TripletMessage(36, 3).to(DEV)
    with torch.no_grad():
        g = graph.graph_index(b.edge_index, b.num_nodes)
        ea = g.sorted_edge_attr(b.edge_attr)
        w_ext, att_edge = m.derived()
        xpe = ops.gemm(b.x, w_ext)
        agg, alpha = ops.triplet_edge_fwd(xpe, ea, m.weight_edge, att_edge, g, 3, 36, 0.2)
        seg = torch.zeros(b.num_nodes, 3, device=DEV).index_add_(0, b.edge_index[1][g.dst_perm.long()], alpha)
        has_in = (g.dst_rowptr[1:] > g.dst_rowptr[:-1])
        torch.testing.assert_close(seg[has_in], torch.ones_like(seg[has_in]), rtol=1e-5, atol=1e-5)
        ep = (ea @ m.weight_edge).view(-1, 3, 36)
        xj = xpe[g.dst_src.long(), :108].view(-1, 3, 36)
        total = (alpha.unsqueeze(-1) * ep * xj).double().sum(0).view(-1)
        torch.testing.assert_close(agg.double().sum(0), total, rtol=1e-6, atol=1e-3)
        out = m(b.x, b.edge_index, b.edge_attr)
    assert torch.isfinite(out).all() and out.shape == (b.num_nodes, 36)


# ---------------------------------------------------------------------------------------------- optimizer / engine
def test_adam_step_matches_torch():
    """glam_adam_step == torch.optim.Adam (the reference trainer's optimizer, src_1gp/trainer.py:49-50) on flat buffers."""
    from glam_b200 import ops
    g = torch.Generator().manual_seed(9)
    n = 108_429
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    state = torch.zeros(3, device=DEV)
    lr = torch.full((1,), 1e-3, device=DEV)
    for step in range(5):
        grad = torch.randn(n, generator=g) * (10.0 ** (step - 2))
        ref.grad = grad.double() * 0.5
        opt.step()
        ops.adam_step(p, grad.to(DEV), m, v, lr, state, grad_scale=0.5)
    assert state[0].item() == 5.0
    torch.testing.assert_close(p.cpu(), ref.detach().float(), rtol=2e-6, atol=2e-7)


def test_captured_train_steps_follow_oracle_training():
    """engine.TrainStep (CUDA graph, gathered flat gradients, flat Adam) against the oracle model trained with
    torch.optim.Adam on the same batches: losses and parameters after several steps (fp32 math mode)."""
    from glam_b200._lib import set_math_mode, get_math_mode
    from glam_b200.engine import TrainStep
    from glam_b200.synth import make_molecule_batch
    prev = get_math_mode()
    set_math_mode("fp32")
    try:
        m, o32 = _gp_pair(9, 3, "Set2Set", "_TripletMessage")
        batches = [make_molecule_batch(48, seed=700 + i, total_nodes=48 * 20, total_edges=48 * 42) for i in range(4)]
        opt = torch.optim.Adam(o32.parameters(), lr=1e-3)
        ref_losses = []
        for b in batches:
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(o32(ns(b.x, b.edge_index, b.edge_attr, b.batch)), b.y)
            loss.backward()
            opt.step()
            ref_losses.append(loss.item())
        sd = {k: v.to(DEV) for k, v in _gp_pair(9, 3, "Set2Set", "_TripletMessage")[1].state_dict().items()}
        ts = TrainStep(m, torch.nn.functional.mse_loss, batches[0], lr=1e-3, device=DEV, use_cuda_graph=True, warmup=3)
        # constructing the step (3 warm-up iterations + capture) must not train: parameters and Adam state as built
        for k, p in m.state_dict().items():
            assert torch.equal(p, sd[k]), k
        assert not ts.opt.exp_avg.any() and not ts.opt.exp_avg_sq.any() and not ts.opt.state.any()
        losses = [ts.step(b.pin_memory()).item() for b in batches]
        for a, r in zip(losses, ref_losses):
            assert abs(a - r) <= 2e-4 * max(1.0, abs(r)), (losses, ref_losses)
        for (n, p), q in zip(m.named_parameters(), o32.parameters()):
            torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-3, atol=2e-4, msg=lambda s, n=n: f"{n}: {s}")
        # double-buffered inputs with the next batch's copy in flight during the step: bitwise the same training run
        m2, _ = _gp_pair(9, 3, "Set2Set", "_TripletMessage")
        ts2 = TrainStep(m2, torch.nn.functional.mse_loss, batches[0], lr=1e-3, device=DEV, use_cuda_graph=True, warmup=1,
                        double_buffer=True)
        pinned = [b.pin_memory() for b in batches]
        losses2 = [ts2.step(b, prefetch=pinned[i + 1] if i + 1 < len(pinned) else None).item() for i, b in enumerate(pinned)]
        assert losses2 == losses, (losses2, losses)
        for p, q in zip(m2.parameters(), m.parameters()):
            assert torch.equal(p, q)
    finally:
        set_math_mode(prev)


@pytest.mark.parametrize("N,C,act,with_id", [(1, 36, 3, True), (127, 36, 0, False), (128 * 3 + 5, 32, 1, True), (40_000, 36, 3, True),
                                             (1000, 40, 2, True), (300, 44, 3, False)])
def test_gru_fused_kernel_matches_gate_math(N, C, act, with_id):
    """glam_gru_fused_fwd (two tcgen05 GEMMs + gates in the epilogue) against fp64 torch gate math; TF32 operand rounding
    is the only difference (|W|,|x| ~ 1, K = C: abs error ~ 1e-3 on the pre-activations)."""
    from glam_b200 import ops
    from glam_b200._lib import set_math_mode, get_math_mode
    prev = get_math_mode()
    set_math_mode("tf32")
    try:
        g = torch.Generator().manual_seed(N + C)
        m, h, ident = (torch.randn(N, C, generator=g) for _ in range(3))
        w_ih, w_hh = (torch.randn(3 * C, C, generator=g) / C ** 0.5 for _ in range(2))
        b_ih, b_hh = (torch.randn(3 * C, generator=g) * 0.1 for _ in range(2))
        assert ops.gru_fused_supported(m.to(DEV), h.to(DEV), C)
        d = lambda t: t.to(DEV)
        rzn, gh, h_new, x_out = ops.gru_fused_fwd(d(m), d(h), d(ident) if with_id else None, d(w_ih), d(w_hh), d(b_ih), d(b_hh), act, 0.1)
        torch.cuda.synchronize()
        gi64 = m.double() @ w_ih.double().t() + b_ih.double()
        gh64 = h.double() @ w_hh.double().t() + b_hh.double()
        r = torch.sigmoid(gi64[:, :C] + gh64[:, :C]); z = torch.sigmoid(gi64[:, C:2 * C] + gh64[:, C:2 * C])
        n = torch.tanh(gi64[:, 2 * C:] + r * gh64[:, 2 * C:])
        hn = (1 - z) * n + z * h.double()
        pre = hn + (ident.double() if with_id else 0)
        xo = {0: pre, 1: torch.relu(pre), 2: torch.nn.functional.leaky_relu(pre, 0.1), 3: torch.nn.functional.celu(pre)}[act]
        tol = dict(rtol=0, atol=5e-3)
        torch.testing.assert_close(rzn.cpu().double(), torch.cat([r, z, n], 1), **tol)
        torch.testing.assert_close(gh.cpu().double(), gh64[:, 2 * C:], **tol)
        torch.testing.assert_close(h_new.cpu().double(), hn, **tol)
        torch.testing.assert_close(x_out.cpu().double(), xo, **tol)
        # and bit-identical run to run
        again = ops.gru_fused_fwd(d(m), d(h), d(ident) if with_id else None, d(w_ih), d(w_hh), d(b_ih), d(b_hh), act, 0.1)
        assert torch.equal(again[3], x_out) and torch.equal(again[0], rzn)
    finally:
        set_math_mode(prev)


# ---------------------------------------------------------------------------------------------- §8(f) rows
@pytest.mark.parametrize("name,C", [("pool5_C36", 36), ("pool5_small_C30", 30)])
def test_pool5_golden(golden_next, name, C):
    """GlobalPool5 kernel against the reference's own output (graphs with < 3 nodes, tied sort keys)."""
    from glam_b200 import layer
    c32, c64 = golden_next[f"{name}_f32"], case(golden_next, f"{name}_f64")
    m = layer.GlobalPool5()
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["batch"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    x2 = c32["x"].to(DEV)
    assert torch.allclose(m.composed_forward(x2, c32["batch"].to(DEV)), out.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,C", [("gcn_C36", 36), ("gcn_protein_C30", 30)])
def test_gcn_conv_golden(golden_next, name, C, math_mode):
    from glam_b200 import layer
    c32, c64 = golden_next[f"{name}_f32"], case(golden_next, f"{name}_f64")
    m = _load(layer._GCNConv(C, C, 3), c32["state"])
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["edge_index"].to(DEV), None)
    tol_check(out, c32["out"], c64["out"], f"{name}.out")
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(m, c32["grad_params"], c64["grad_params"], name)


def test_gcn_block_golden(golden_next, math_mode):
    from glam_b200 import layer
    name = "block_gcn_C36"
    c32, c64 = golden_next[f"{name}_f32"], case(golden_next, f"{name}_f64")
    cfg = c32["cfg"]
    blk = _load(layer.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"],
                                   act=cfg["act"], res=cfg["res"]), c32["state"])
    x0 = c32["x"].to(DEV).requires_grad_(True)
    ei, ea, batch = c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV), c32["batch"].to(DEV)
    xs, h = blk.run_steps(x0, ei, ea, cfg["steps"], batch=batch)
    tol_check(xs[-1], c32["out"], c64["out"], f"{name}.out")
    tol_check(h, c32["h"], c64["h"], f"{name}.h")
    ((xs[-1] * c32["cot"].to(DEV)).sum() + (h * c32["coth"].to(DEV)).sum()).backward()
    tol_check(x0.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x")
    _check_param_grads(blk, c32["grad_params"], c64["grad_params"], name)


def test_model_dti_reference_defaults_golden(golden_next, math_mode):
    """Drug-target model with the reference's default protein block (_GCNConv) and readouts (GlobalPool5) against the
    reference's own output and parameter gradients."""
    from glam_b200 import model
    c = golden_next["dti_gcn_pool5"]
    cfg = c["cfg"]
    m = model.ArchitectureDTI(cfg["Din"], cfg["Pin"], cfg["De"], cfg["Pe"], hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1,
                              mol_block="_TripletMessage", pro_block="_GCNConv", message_steps=3, mol_readout="GlobalPool5",
                              pro_readout="GlobalPool5", graph_do="_None()", end_do="_None()", pre_act="ReLU",
                              graph_act="LeakyReLU", flat_act="CELU", end_act="ReLU")
    m = _load(m, c["state"]).eval()
    da = ns(c["a_x"].to(DEV), c["a_edge_index"].to(DEV), c["a_edge_attr"].to(DEV), c["a_batch"].to(DEV))
    db = ns(c["b_x"].to(DEV), c["b_edge_index"].to(DEV), c["b_edge_attr"].to(DEV), c["b_batch"].to(DEV))
    out = m(da, db)
    tf32 = math_mode == "tf32"
    torch.testing.assert_close(out.cpu(), c["out"], rtol=2e-2 if tf32 else 2e-4, atol=2e-2 if tf32 else 2e-5)
    torch.nn.functional.binary_cross_entropy_with_logits(out, c["y"].to(DEV)).backward()
    for n, p in m.named_parameters():
        ref = c["grad_params"][n]
        scale = ref.abs().max().clamp(min=1e-6)
        err = (p.grad.cpu() - ref).abs().max() / scale
        cos = torch.nn.functional.cosine_similarity(p.grad.cpu().flatten(), ref.flatten(), dim=0)
        if tf32:                         # max-routing of the dot-pool can flip under TF32 (see test_model_ddi_golden)
            assert err < 2e-1 and cos > 0.999, f"{n}: rel err {err:.3e}, cos {cos:.6f}"
        else:
            assert err < 2e-3, f"{n}: rel err {err:.3e}"


# ---------------------------------------------------------------------------------------------- _NNConv (run.py's default block)
@pytest.mark.parametrize("name,C,De", [("nnconv_C36", 36, 3), ("nnconv_C20_De4", 20, 4)])
def test_nnconv_golden(golden_nnconv, name, C, De, math_mode):
    """`_NNConv` (typed grouped projection + gather-mean over the CSR) against the reference's own layer, outputs and all
    gradients; includes an isolated node (mean over no edges) and duplicate edges."""
    from glam_b200 import layer
    c32, c64 = case(golden_nnconv, f"{name}_f32"), case(golden_nnconv, f"{name}_f64")
    m = layer._NNConv(C, C, De)
    assert list(m.state_dict().keys()) == list(c32["state"].keys())
    m = _load(m, c32["state"])
    x = c32["x"].to(DEV).requires_grad_(True)
    out = m(x, c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV))
    tol_check(out, c32["out"], c64["out"], f"{name}.out", rtol=2e-4)
    (out * c32["cot"].to(DEV)).sum().backward()
    tol_check(x.grad, c32["grad_x"], c64["grad_x"], f"{name}.grad_x", rtol=2e-4)
    for n, p in m.named_parameters():
        tol_check(p.grad, c32["grad_params"][n], c64["grad_params"][n], f"{name}.grad[{n}]", rtol=2e-4)


def test_nnconv_rejects_general_edge_rows():
    from glam_b200 import layer, _lib
    m = layer._NNConv(36, 36, 3).to(DEV)
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]], device=DEV)
    with pytest.raises(_lib.GlamError):
        m(torch.randn(3, 36, device=DEV), ei, torch.rand(3, 3, device=DEV))


def test_nnconv_block_golden(golden_nnconv, math_mode):
    from glam_b200 import layer
    c32, c64 = case(golden_nnconv, "block_nnconv_C36_f32"), case(golden_nnconv, "block_nnconv_C36_f64")
    cfg = c32["cfg"]
    blk = layer.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"], act=cfg["act"],
                             res=cfg["res"])
    blk = _load(blk, c32["state"])
    x0 = c32["x"].to(DEV).requires_grad_(True)
    x, h = x0, None
    for _ in range(cfg["steps"]):
        x, h = blk(x, c32["edge_index"].to(DEV), c32["edge_attr"].to(DEV), h=h, batch=c32["batch"].to(DEV))
    tol_check(x, c32["out"], c64["out"], "block_nnconv.out", rtol=2e-4)
    tol_check(h, c32["h"], c64["h"], "block_nnconv.h", rtol=2e-4)
    ((x * c32["cot"].to(DEV)).sum() + (h * c32["coth"].to(DEV)).sum()).backward()
    tol_check(x0.grad, c32["grad_x"], c64["grad_x"], "block_nnconv.grad_x", rtol=2e-4)
    for n, p in blk.named_parameters():
        tol_check(p.grad, c32["grad_params"][n], c64["grad_params"][n], f"block_nnconv.grad[{n}]", rtol=2e-4)


def test_model_gp_reference_defaults_golden(golden_nnconv, math_mode):
    """GLAM-GP with the reference's default block and readout (`_NNConv` + `GlobalPool5`, src_1gp/run.py:21,25) against the
    reference's own output and parameter gradients."""
    from glam_b200 import model
    c = golden_nnconv["gp_nnconv_pool5"]
    cfg = c["cfg"]
    m = model.ArchitectureGP(cfg["Din"], cfg["De"], graph_do="_None()", end_do="_None()", hid_dim_alpha=4, e_dim=cfg["e_dim"],
                             out_dim=1, mol_block=cfg["block"], message_steps=3, mol_readout=cfg["readout"],
                             graph_norm=cfg["graph_norm"], pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU")
    assert list(m.state_dict().keys()) == list(c["state"].keys())
    m = _load(m, c["state"]).eval()
    out = m(ns(c["x"].to(DEV), c["edge_index"].to(DEV), c["edge_attr"].to(DEV), c["batch"].to(DEV)))
    tf32 = math_mode == "tf32"
    torch.testing.assert_close(out.cpu(), c["out"], rtol=2e-2 if tf32 else 2e-4, atol=2e-2 if tf32 else 2e-5)
    torch.nn.functional.mse_loss(out, c["y"].to(DEV)).backward()
    for n, p in m.named_parameters():
        ref = c["grad_params"][n]
        scale = ref.abs().max().clamp(min=1e-6)
        err = (p.grad.cpu() - ref).abs().max() / scale
        assert err < (5e-2 if tf32 else 2e-3), f"{n}: rel err {err:.3e}"


# ---------------------------------------------------------------------------------------------- windowed edge kernels
def _edge_phase_ref64(xpe, ea, we, ae, src, dst, N, H, C, slope, g_agg):
    """fp64 restatement of TripletMessage.message + aggregate on the extended projection (SURVEY.md Appendix C;
    src_1gp/layer.py:42-55 with PyG's softmax) and its gradients by autograd."""
    HC = H * C
    xpe = xpe.double().requires_grad_(True)
    ea, ae, g_agg = ea.double(), ae.double(), g_agg.double()
    we = None if we is None else we.double().requires_grad_(True)
    src, dst = src.long(), dst.long()
    E = src.shape[0]
    pre = xpe[:, HC:HC + H][dst] + xpe[:, HC + H:HC + 2 * H][src] + ea @ ae
    pre.retain_grad()
    l = torch.nn.functional.leaky_relu(pre, slope)
    mx = torch.full((N, H), -float("inf"), dtype=torch.float64, device=xpe.device).scatter_reduce(
        0, dst[:, None].expand(E, H), l, "amax", include_self=True)
    ex = torch.exp(l - mx[dst])
    den = torch.zeros((N, H), dtype=torch.float64, device=xpe.device).index_add_(0, dst, ex)
    alpha = ex / (den[dst] + 1e-16)
    msg = alpha[:, :, None] * xpe[:, :HC].view(N, H, C)[src]
    if we is not None:
        msg = msg * (ea @ we).view(E, H, C)
    agg = torch.zeros((N, H, C), dtype=torch.float64, device=xpe.device).index_add_(0, dst, msg).view(N, HC)
    (agg * g_agg).sum().backward()
    return agg.detach(), alpha.detach(), xpe.grad, pre.grad, None if we is None else we.grad


def _edge_case(kind, seed):
    from glam_b200.synth import make_molecule_batch
    g = torch.Generator().manual_seed(seed)
    if kind == "molecules":                       # every tile near: windows staged in shared memory
        b = make_molecule_batch(700, node_dim=9, edge_dim=3, seed=seed)
        return b.edge_index, b.num_nodes
    if kind == "big_molecules":                   # graphs larger than a tile: a share of the windows exceeds the buffer
        import numpy as np
        b = make_molecule_batch(40, node_dim=9, edge_dim=3, seed=seed, sizes=np.full(40, 110))
        return b.edge_index, b.num_nodes
    if kind == "random":                          # no locality at all: every tile is far (rows gathered from global memory)
        N = 6000
        return torch.randint(0, N, (2, 15000), generator=g), N
    if kind == "hub":                             # tiles with more edges than the record buffer + isolated nodes
        N = 2000
        ei = torch.randint(0, N - 50, (2, 9000), generator=g)
        ei[1, :3000] = 17
        ei[1, 3000:3700] = 1203
        ei[0, 4000:5200] = 40                     # a hub SOURCE: overflow in the source pass
        return ei, N
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["molecules", "big_molecules", "random", "hub"])
@pytest.mark.parametrize("H,C,De,onehot,light", [
    (3, 36, 3, True, False),      # BASELINE dims: 3 chunks per item, 9 items
    (3, 60, 4, True, False),      # reference dims: 15 items, two edges per warp pass in the destination pass
    (2, 32, 3, True, False),      # 2 chunks per item
    (1, 20, 5, False, False),     # 1 chunk per item, general edge rows, edge_dim > 4 (destination pass: gather kernels)
    (4, 24, 8, False, False),     # protein-style 8-dim contact features
    (1, 36, 3, True, True),       # TripletMessageLight: no edge projection
])
def test_windowed_edge_kernels(kind, H, C, De, onehot, light):
    """The bulk-copy staged edge kernels (near / far / overflow tiles) and the per-edge gather kernels against an fp64
    restatement, outputs and all three gradients; and the tile descriptors against their definition."""
    from glam_b200 import graph, ops, _lib
    ei, N = _edge_case(kind, 11)
    ei = ei.to(DEV)
    g = graph.GraphIndex(ei, N)
    E = ei.shape[1]
    # ---- descriptors: bit-exact against the definition in include/glam_b200.h
    D = _lib.load().glam_edge_tile_rows(N)
    T = (N + D - 1) // D
    assert g.dst_tiles.shape == (T, 4) and g.src_tiles.shape == (T, 4)
    rp, srcs = g.dst_rowptr.long().cpu(), g.dst_src.long().cpu()
    srp, sdst = g.src_rowptr.long().cpu(), g.src_dst.long().cpu()
    dt, st = g.dst_tiles.cpu(), g.src_tiles.cpu()
    for t in (0, T // 3, T // 2, T - 1):
        t0, t1 = t * D, min(N, (t + 1) * D)
        e0, e1 = int(rp[t0]), int(rp[t1])
        lo = min([t0] + srcs[e0:e1].tolist()); hi = max([t1 - 1] + srcs[e0:e1].tolist()) + 1
        assert dt[t].tolist() == [lo, hi, e0, e1]
        k0, k1 = int(srp[t0]), int(srp[t1])
        want = [int(sdst[k0:k1].min()), int(sdst[k0:k1].max()) + 1, k0, k1] if k1 > k0 else [0, 0, k0, k1]
        assert st[t].tolist() == want
    # ---- kernels
    gen = torch.Generator().manual_seed(5)
    HC = H * C
    ld = (HC + 2 * H + 3) // 4 * 4
    xpe = torch.randn(N, ld, generator=gen).to(DEV)
    if onehot:
        ea = torch.eye(De)[torch.randint(0, De, (E,), generator=gen)] * (0.5 + torch.rand(E, 1, generator=gen))
    else:
        ea = torch.rand(E, De, generator=gen) * (torch.rand(E, De, generator=gen) > 0.3)
    ea = g.sorted_edge_attr(ea.to(DEV))
    we = None if light else (torch.randn(De, HC, generator=gen) * 0.5).to(DEV)
    ae = (torch.randn(De, H, generator=gen) * 0.5).to(DEV)
    g_agg = torch.randn(N, HC, generator=gen).to(DEV)
    ref = _edge_phase_ref64(xpe[:, :HC + 2 * H].contiguous(), ea, we, ae, g.dst_src, g.dst_dst, N, H, C, 0.2, g_agg)
    outs = {}
    try:
        for tiles in (True, False):
            ops.USE_EDGE_TILES = tiles
            agg, alpha = ops.triplet_edge_fwd(xpe, ea, we, ae, g, H, C, 0.2)
            g_xpe, g_logit, g_we = ops.triplet_edge_bwd(xpe, ea, we, ae, alpha, g_agg, g, H, C, 0.2)
            outs[tiles] = (agg, alpha, g_xpe[:, :HC + 2 * H], g_logit, g_we)
            torch.cuda.synchronize()
            assert torch.equal(g_xpe[:, HC + 2 * H:], torch.zeros_like(g_xpe[:, HC + 2 * H:]))     # pad columns zeroed
            for name, ours, want in zip(("agg", "alpha", "g_xpe", "g_logit", "g_w_edge"), outs[tiles], ref):
                if want is None:
                    assert ours is None
                    continue
                scale = want.abs().max().clamp(min=1e-30)
                err = (ours.double() - want).abs().max()
                assert err <= 1e-5 + 1e-4 * scale, f"{kind} tiles={tiles} {name}: err {err:.3e} (scale {scale:.3e})"
            # run-to-run bitwise reproducible
            agg2, alpha2 = ops.triplet_edge_fwd(xpe, ea, we, ae, g, H, C, 0.2)
            b2 = ops.triplet_edge_bwd(xpe, ea, we, ae, alpha, g_agg, g, H, C, 0.2)
            assert torch.equal(agg, agg2) and torch.equal(alpha, alpha2)
            assert torch.equal(g_xpe, b2[0]) and torch.equal(g_logit, b2[1]) and (g_we is None or torch.equal(g_we, b2[2]))
    finally:
        ops.USE_EDGE_TILES = True


def test_screen_step_double_buffer_matches_eager():
    """engine.ScreenStep (captured eval forward) with and without the double-buffered host prefetch against the eager
    module call, bitwise."""
    from glam_b200.engine import ScreenStep
    from glam_b200.synth import make_molecule_batch
    m, _ = _gp_pair(9, 3, "Set2Set", "_TripletMessage")
    m = m.to(DEV).eval()
    batches = [make_molecule_batch(64, seed=900 + i, total_nodes=64 * 22, total_edges=64 * 46).pin_memory() for i in range(4)]
    with torch.no_grad():
        want = [m(b.to(DEV)).clone() for b in batches]
    s1 = ScreenStep(m, batches[0], device=DEV)
    s2 = ScreenStep(m, batches[0], device=DEV, double_buffer=True)
    for i, b in enumerate(batches):
        o1 = s1.step(b).clone()
        o2 = s2.step(b, prefetch=batches[i + 1] if i + 1 < len(batches) else None).clone()
        assert torch.equal(o1, want[i]) and torch.equal(o2, want[i])


# ---------------------------------------------------------------------------------------------- DTI screening: distinct proteins once
def _repeat_graphs(u, index):
    """The collated batch the reference would build (one copy of its protein graph per pair, src_2gi_dti_scr/dataset.py:329-335)
    from a batch `u` of distinct graphs and the pair -> graph index."""
    from glam_b200.synth import GraphBatch
    nptr = torch.zeros(u.num_graphs + 1, dtype=torch.long)
    nptr[1:] = torch.bincount(u.batch, minlength=u.num_graphs).cumsum(0)
    eg = u.batch[u.edge_index[0]]
    xs, eis, eas, bs, off = [], [], [], [], 0
    for p, gidx in enumerate(index):
        n0, n1 = int(nptr[gidx]), int(nptr[gidx + 1])
        sel = eg == gidx
        xs.append(u.x[n0:n1]); eas.append(u.edge_attr[sel]); eis.append(u.edge_index[:, sel] - n0 + off)
        bs.append(torch.full((n1 - n0,), p, dtype=torch.long)); off += n1 - n0
    return GraphBatch(torch.cat(xs), torch.cat(eis, 1), torch.cat(eas), torch.cat(bs), None, len(index))


@pytest.mark.parametrize("keys", [["T1"] * 12, ["a", "b", "a", "c", "b", "a", "c", "c", "a"]])
def test_dti_screening_with_distinct_proteins_once(keys, math_mode):
    """the protein side of a screening batch holds every distinct protein once (`pro_index`); scores are
    bitwise those of the reference-style batch with one protein copy per pair."""
    from glam_b200 import model
    from glam_b200.synth import make_molecule_batch, make_protein_batch
    pos, index = model.dedupe_keys(keys)
    assert [keys[i] for i in pos] == list(dict.fromkeys(keys)) and [list(dict.fromkeys(keys))[i] for i in index.tolist()] == keys
    P = len(keys)
    torch.manual_seed(3)
    m = model.ArchitectureDTI(15, 49, 4, 8, hid_dim_alpha=4, e_dim=64, out_dim=1, mol_block="_TripletMessage", pro_block="_GCNConv",
                              message_steps=3, mol_readout="GlobalPool5", pro_readout="GlobalPool5", graph_do="_None()", end_do="_None()",
                              pre_act="ReLU", graph_act="LeakyReLU", flat_act="CELU", end_act="ReLU").to(DEV).eval()
    lig = make_molecule_batch(P, node_dim=15, edge_dim=4, seed=21)
    uniq = make_protein_batch(len(pos), seed=22, min_len=60, max_len=140)
    full = _repeat_graphs(uniq, index.tolist())
    with torch.no_grad():
        want = m(lig.to(DEV), full.to(DEV))
        got = m(lig.to(DEV), uniq.to(DEV), pro_index=index.to(DEV))
    assert torch.isfinite(got).all() and torch.equal(got, want)
    with pytest.raises(Exception):                       # forward-only: gradients through a shared protein are not defined here
        m.train()
        m(lig.to(DEV), uniq.to(DEV), pro_index=index.to(DEV)).sum().backward()


def test_captured_train_steps_on_batches_of_varying_size():
    """A TrainStep captured on a PADDED example (synth.pad_graph_batch + engine.masked_loss) trains on batches of different sizes
    through one CUDA graph: losses and parameters follow the oracle trained with torch.optim.Adam on the UNPADDED batches (the
    dummy graphs carry zero loss weight, hence no gradient)."""
    from glam_b200._lib import set_math_mode, get_math_mode
    from glam_b200.engine import TrainStep, masked_loss
    from glam_b200.synth import make_molecule_batch, pad_graph_batch
    prev = get_math_mode()
    set_math_mode("fp32")
    try:
        m, o32 = _gp_pair(9, 3, "Set2Set", "_TripletMessage")
        batches = [make_molecule_batch(n, seed=800 + i) for i, n in enumerate((48, 41, 48, 17))]
        opt = torch.optim.Adam(o32.parameters(), lr=1e-3)
        ref_losses = []
        for b in batches:
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(o32(ns(b.x, b.edge_index, b.edge_attr, b.batch)), b.y)
            loss.backward()
            opt.step()
            ref_losses.append(loss.item())
        cap = (max(b.num_nodes for b in batches) + 40, max(b.num_edges for b in batches) + 40, 48 + 6)
        example = pad_graph_batch(batches[0], *cap, with_mask=True)
        assert example.mask is not None and int(example.mask.sum()) == 48
        ts = TrainStep(m, masked_loss(torch.nn.functional.mse_loss), example, lr=1e-3, device=DEV, use_cuda_graph=True, warmup=2,
                       double_buffer=True)
        pinned = [b.pin_memory() for b in batches]
        losses = [ts.step(b, prefetch=pinned[i + 1] if i + 1 < len(pinned) else None).item() for i, b in enumerate(pinned)]
        for a, r in zip(losses, ref_losses):
            assert abs(a - r) <= 2e-4 * max(1.0, abs(r)), (losses, ref_losses)
        for (n, p), q in zip(m.named_parameters(), o32.parameters()):
            torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-3, atol=2e-4, msg=lambda s, n=n: f"{n}: {s}")
    finally:
        set_math_mode(prev)
