"""One forward + backward of the row-block aggregation at a BASELINE shape (for ncu): python tests/tools/run_agg_once.py pubmed tc"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gnan_b200 import ops  # noqa: E402
from bench_agg import SHAPES  # noqa: E402

name, algo = sys.argv[1], sys.argv[2]
R, N, C, nbins = SHAPES[name]
train_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(0)
hop = ops.alloc_hop(R, N, "cuda")
hop[:, :N] = torch.randint(0, nbins - 1, (R, N), device="cuda", dtype=torch.uint8, generator=g)
T = torch.randn(R, nbins, C, device="cuda", requires_grad=True)
S = torch.randn(N, C, device="cuda", requires_grad=True)
gO = torch.randn(R, C, device="cuda")
if train_rows:
    m = torch.zeros(R, 1, device="cuda"); m[torch.randperm(R, device="cuda")[:train_rows]] = 1.0
    gO = gO * m
for _ in range(3):
    out = ops.aggregate_rows(hop, T, S, per_row=True, algo=algo)
    (out * gO).sum().backward()
torch.cuda.synchronize()
