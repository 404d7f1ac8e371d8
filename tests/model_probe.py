"""Ad-hoc probe: which TF32 contraction is responsible for the weight_triplet_att gradient error?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from glam_b200 import model, _lib, ops
from oracle import glam_oracle as O
from helpers import ns
mx = torch.load(os.path.join(ROOT, "tests/golden/models.pt"))
c = mx["gp_set2set"]; cfg = c["cfg"]
kw = dict(hid_dim_alpha=4, e_dim=cfg["e_dim"], out_dim=1, mol_block=cfg["block"], message_steps=3, mol_readout=cfg["readout"],
          graph_norm=cfg["graph_norm"], pre_act="ReLU", graph_act="CELU", flat_act="LeakyReLU")
o = O.ArchitectureGP(cfg["Din"], cfg["De"], **kw); o.load_state_dict(c["state"]); o = o.double().eval()
out64 = o(ns(c["x"].double(), c["edge_index"], c["edge_attr"].double(), c["batch"]))
g64 = dict(zip([n for n, _ in o.named_parameters()], torch.autograd.grad(torch.nn.functional.mse_loss(out64, c["y"].double()), list(o.parameters()))))
_lib.load()
orig_gemm, orig_tn = ops.gemm, ops.gemm_tn_ex
seen = set()
def run(exact_pred, tn_exact=False):
    def gemm(X, W, transpose_w=False, bias=None, epilogue=0, aux=None, out=None, ldy=None):
        M, K = X.shape; N = W.shape[0] if transpose_w else W.shape[1]
        sig = (N, K, transpose_w, epilogue); seen.add(sig)
        _lib.set_math_mode("fp32" if exact_pred(sig) else "tf32")
        r = orig_gemm(X, W, transpose_w, bias, epilogue, aux, out, ldy)
        _lib.set_math_mode("tf32"); return r
    def tn(A, B, transpose_out=False, want_colsum=False, out=None):
        _lib.set_math_mode("fp32" if tn_exact else "tf32")
        r = orig_tn(A, B, transpose_out, want_colsum, out)
        _lib.set_math_mode("tf32"); return r
    ops.gemm, ops.gemm_tn_ex = gemm, tn
    m = model.ArchitectureGP(cfg["Din"], cfg["De"], graph_do="_None()", end_do="_None()", **kw); m.load_state_dict(c["state"]); m = m.cuda().eval()
    out = m(ns(c["x"].cuda(), c["edge_index"].cuda(), c["edge_attr"].cuda(), c["batch"].cuda()))
    torch.nn.functional.mse_loss(out, c["y"].cuda()).backward()
    ops.gemm, ops.gemm_tn_ex = orig_gemm, orig_tn
    errs = {n: ((p.grad.double().cpu() - g64[n]).abs().max() / g64[n].abs().max()).item() for n, p in m.named_parameters()}
    att = m.mol_conv.conv.conv.weight_triplet_att.grad.double().cpu().view(3, 3, 36); ref = g64["mol_conv.conv.conv.weight_triplet_att"].view(3, 3, 36)
    parts = [((att[:, i] - ref[:, i]).abs().max() / ref.abs().max()).item() for i in range(3)]
    return errs["mol_conv.conv.conv.weight_triplet_att"], errs["mol_conv.conv.conv.weight_node"], parts
print("all tf32             ", run(lambda s: False))
import glam_b200.functional as F
orig_conv_fwd = F._conv_fwd
def conv_fwd_exact_logits(x, w_ext, w_edge, att_edge, w_scale, bias, ea, g, heads, channels, slope, epilogue):
    xpe = ops.gemm(x, w_ext)
    hc = heads * channels
    _lib.set_math_mode("fp32")
    orig_gemm(x, w_ext[:, hc:hc + 2 * heads].contiguous(), out=xpe[:, hc:hc + 2 * heads])
    _lib.set_math_mode("tf32")
    agg, alpha = ops.triplet_edge_fwd(xpe, ea, w_edge, att_edge, g, heads, channels, slope)
    out = agg if w_scale is None else ops.gemm(agg, w_scale, bias=bias, epilogue=epilogue)
    return xpe, agg, alpha, out
F._conv_fwd = conv_fwd_exact_logits
print("tf32 + exact logit cols", run(lambda s: False))
F._conv_fwd = orig_conv_fwd
