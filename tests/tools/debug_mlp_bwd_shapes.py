"""dW of the dense grouped-MLP backward (fp32 vs tf32x3) against float64 at several shapes / sparsities (GPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gnan_b200 import ops  # noqa: E402
from oracle import gnan_lut  # noqa: E402

DEV = "cuda"


def run(R, G, C, density, bias0, seed=0, scale_x=0.1, ds_rows=None):
    gen = torch.Generator().manual_seed(seed)
    H = 64
    p = dict(w1=torch.randn(G, H, generator=gen) * (2 / 65) ** 0.5, b1=torch.zeros(G, H) if bias0 else torch.randn(G, H, generator=gen) * 0.1,
             wh=torch.randn(1, G, H, H, generator=gen) * (2 / 128) ** 0.5, bh=torch.zeros(1, G, H) if bias0 else torch.randn(1, G, H, generator=gen) * 0.1,
             wo=torch.randn(G, C, H, generator=gen) * (2 / (64 + C)) ** 0.5, bo=torch.zeros(G, C))
    x = (torch.rand(R, G, generator=gen) < density).float() * torch.rand(R, G, generator=gen) * scale_x
    dS = torch.randn(R, C, generator=gen)
    if ds_rows:
        m = torch.zeros(R, 1); m[torch.randperm(R, generator=gen)[:ds_rows]] = 1; dS = dS * m
    q = {k: v.double().requires_grad_(True) for k, v in p.items()}
    S = gnan_lut.feature_sums(q, x.double())
    (S * dS.double()).sum().backward()
    out = {}
    for prec in ("fp32", "tf32x3"):
        d = {k: v.to(DEV).requires_grad_(True) for k, v in p.items()}
        got = ops.mlp(x.to(DEV), d["w1"], d["b1"], d["wh"], d["bh"], d["wo"], d["bo"], 3, precision=prec)
        (got * dS.to(DEV)).sum().backward()
        rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
        out[prec] = {k: f"{rel(d[k].grad, q[k].grad):.1e}" for k in ("w1", "wh", "wo")}
        out[prec]["S"] = f"{rel(got.detach(), S.detach()):.1e}"
    print(f"R={R} G={G} C={C} density={density} bias0={bias0} ds_rows={ds_rows}: {out}", flush=True)


if __name__ == "__main__":
    run(2708, 8, 7, 0.013, True)
    run(19717, 8, 3, 0.1, True)
    run(19717, 8, 3, 0.1, False)
    run(19717, 8, 3, 1.0, True)
    run(19717, 8, 3, 1.0, False)
    run(2000, 8, 3, 1.0, True)
    run(19717, 8, 3, 0.1, True, scale_x=1.0)
    run(19717, 2, 3, 0.1, True)
    run(19717, 8, 3, 0.1, True, ds_rows=None, seed=3)
