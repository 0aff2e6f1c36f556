"""Per-parameter gradient differences of the bench.py parity problem between the aggregation kernel families and against a
two-block row split computed on ONE GPU (emulates the row-sharded path without NCCL)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench  # noqa: E402
from gnan_b200 import ops  # noqa: E402
from gnan_b200.GNAN import TensorGNAN  # noqa: E402
from gnan_b200.preprocess import apsp  # noqa: E402

dev = "cuda"
rng = np.random.default_rng(5)
n, K, C = 4000, 24, 5
ei = bench.random_simple_graph(rng, n, 7000, 30)
x = torch.tensor(rng.normal(size=(n, K))).float().to(dev)
w = torch.tensor(rng.normal(size=(n, C))).float().to(dev)
w[torch.tensor(rng.random(n) > 0.05).to(dev)] = 0.0
torch.manual_seed(1)
m = TensorGNAN(K, C, 3, 64, normalize_rho=True, device=dev).to(dev)
m.fs.xavier_normal_(1.0); m.rho.xavier_normal_(1.0)
hd = apsp(torch.tensor(ei), n, device=dev)
names = [k for k, _ in m.named_parameters()]


def grads(algo, split=None):
    ops.AGG_ALGO = algo
    m.zero_grad(set_to_none=True)
    if split is None:
        out = m(SimpleNamespace(x=x, hop_data=hd))
        (out * w).sum().backward()
    else:
        S = m._feature_sums(x, None)
        outs = []
        for b, e in split:
            h = apsp(torch.tensor(ei), n, device=dev, row_begin=b, row_end=e)
            u = ops.rho_table_inputs(h.nbins, dev, cnt=h.level_counts)
            T = m._row_tables(h, u, h.level_counts).view(h.rows, h.nbins, -1)
            outs.append(ops.aggregate_rows(h.hop, T, S, per_row=True))
        out = torch.cat(outs)
        (out * w).sum().backward()
    ops.AGG_ALGO = "auto"
    return out.detach().double(), [p.grad.detach().double().clone() for p in m.parameters()]


rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-300))
o_c, g_c = grads("cuda")
o_t, g_t = grads("tc")
o_s, g_s = grads("tc", split=[(0, 2000), (2000, 4000)])
print("out tc vs cuda", rel(o_t, o_c), " split vs full (tc)", rel(o_s, o_t))
for nm, a, b, c in zip(names, g_c, g_t, g_s):
    print(f"{nm:10s} tc vs cuda {rel(b, a):.2e}   split vs full (tc) {rel(c, b):.2e}   |g| {float(a.norm()):.3e}")
