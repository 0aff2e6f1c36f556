"""Micro-benchmark of the row-block aggregation kernels (forward incl. the digit pre-pass; backward) on synthetic hop
matrices of the BASELINE shapes: python tests/tools/bench_agg.py [shape ...]   (shapes: pubmed cora arxiv1 arxiv40 arxiv40b)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gnan_b200 import ops  # noqa: E402

DEV = "cuda"
SHAPES = {"pubmed": (19717, 19717, 3, 14), "cora": (2708, 2708, 7, 15), "arxiv1": (16384, 169343, 1, 12),
          "arxiv40": (4096, 169343, 40, 12), "pubmed8": (19717, 19717, 3, 8), "pubmed24": (19717, 19717, 3, 24),
          "arxiv40_8": (4096, 169343, 40, 8)}


def run(name, R, N, C, nbins, per_row=True, iters=5, train_rows=None):
    g = torch.Generator(device=DEV).manual_seed(0)
    hop = ops.alloc_hop(R, N, DEV)
    hop[:, :N] = torch.randint(0, nbins - 1, (R, N), device=DEV, dtype=torch.uint8, generator=g)
    T = torch.randn((R, nbins, C) if per_row else (nbins, C), device=DEV, requires_grad=True)
    S = torch.randn(N, C, device=DEV, requires_grad=True)
    gO = torch.randn(R, C, device=DEV)
    if train_rows:
        m = torch.zeros(R, 1, device=DEV); m[torch.randperm(R, device=DEV)[:train_rows]] = 1.0
        gO = gO * m
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    rec = {"shape": name, "R": R, "N": N, "C": C, "nbins": nbins, "per_row": per_row, "hop_GB": R * N / 1e9, "train_rows": train_rows}
    ref = None
    for algo in ("cuda", "tc"):
        tf = tb = 0.0
        for it in range(iters + 2):
            flush.fill_(1)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); out = ops.aggregate_rows(hop, T, S, per_row=per_row, algo=algo); e[1].record()
            (out * gO).sum().backward(); e[2].record(); torch.cuda.synchronize()
            if it >= 2:
                tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        tf /= iters; tb /= iters
        rec[algo] = {"fwd_ms": tf, "fwd_GBs": R * N / 1e6 / tf, "bwd_ms": tb, "bwd_GBs": R * N / 1e6 / tb}
        o = out.detach().double()
        if ref is None:
            ref = o
        else:
            rec["tc_vs_cuda_rel"] = float((o - ref).norm() / ref.norm())
        T.grad = None; S.grad = None
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    for s in (sys.argv[1:] or ["pubmed", "cora", "arxiv1", "arxiv40"]):
        run(s, *SHAPES[s])
        torch.cuda.empty_cache()
