"""Ad-hoc probe: is the tf32-mode gradient error of the PairNorm block explained by operand precision alone?"""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import glam_oracle as O
from helpers import case
fx = torch.load(os.path.join(ROOT, "tests/golden/layers.pt"))
for name in ["block_triplet_C36", "block_triplet_pn_C60"]:
    c32, c64 = fx[name + "_f32"], case(fx, name + "_f64")
    cfg = c32["cfg"]
    for hook in (None, O.tf32_truncate):
        O.MM_OPERAND_HOOK = hook
        blk = O.MessageBlock(cfg["C"], cfg["C"], cfg["De"], norm=cfg["norm"], dropout="_None()", conv=cfg["conv"], act=cfg["act"], res=cfg["res"]).double()
        blk.load_state_dict(c64["state"])
        x0 = c64["x"].clone().requires_grad_(True)
        x, h = x0, None
        for _ in range(cfg["steps"]):
            x, h = blk(x, c32["edge_index"], c64["edge_attr"], h=h, batch=c32["batch"])
        g = torch.autograd.grad((x * c64["cot"]).sum() + (h * c64["coth"]).sum(), [x0] + list(blk.parameters()))
        rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
        print(name, "tf32-emulated oracle" if hook else "exact oracle", f"out {rel(x, c64['out']):.1e} grad_x {rel(g[0], c64['grad_x']):.1e}",
              " ".join(f"{n.split('.')[-1]} {rel(gg, c64['grad_params'][n]):.1e}" for (n, _), gg in zip(blk.named_parameters(), g[1:])))
    O.MM_OPERAND_HOOK = None
