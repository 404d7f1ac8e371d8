"""Ad-hoc probe (not a pytest): TF32 tensor-core GEMM vs fp64 on a few shapes and epilogues."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import ops, _lib
_lib.load(); _lib.set_math_mode("tf32")
torch.manual_seed(0)
shapes = [(128, 16, 8, False), (128, 128, 32, False), (300, 116, 36, False), (1000, 36, 108, False), (5000, 108, 36, True),
          (4096, 188, 60, False), (777, 60, 180, False), (777, 60, 180, True), (500, 180, 60, True), (500, 60, 188, True), (500, 36, 116, True),
          (130, 256, 288, False), (200, 270, 90, True), (200, 90, 270, False)]
for (M, N, K, nt) in shapes:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda"); b = torch.randn(N, device="cuda")
    Wt = W.t().contiguous() if nt else W
    aux = torch.celu(torch.randn(M, N, device="cuda")); acc0 = torch.randn(M, N, device="cuda")
    d = torch.where(aux > 0, torch.ones_like(aux), aux + 1).double()
    base = X.double() @ W.double()
    refs = {0: base + b.double(), 1: torch.celu(base + b.double()), 2: base * d, 3: base + acc0.double()}
    msg = []
    for epi in (0, 1, 2, 3):
        out = acc0.clone() if epi == 3 else None
        Y = ops.gemm(X, Wt, transpose_w=nt, bias=(b if epi < 2 else None), epilogue=epi, aux=(aux if epi == 2 else None), out=out)
        torch.cuda.synchronize()
        err = (Y.double() - refs[epi]).abs().max().item(); scale = refs[epi].abs().max().item()
        msg.append(f"epi{epi} rel {err/scale:.1e}")
    print(f"M={M} N={N} K={K} nt={nt}: " + "  ".join(msg))
