"""APSP timings on one B200 (run under gpurun): python scratch/bench_apsp.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import gnan_b200
from gnan_b200.preprocess import apsp, apsp_batched

def graph(n, e, seed=0):
    g = torch.Generator().manual_seed(seed)
    s = torch.randint(0, n, (e,), generator=g); d = torch.randint(0, n, (e,), generator=g)
    keep = s != d
    s, d = s[keep], d[keep]
    key = torch.unique(torch.cat([s * n + d, d * n + s]))
    return torch.stack([key // n, key % n])

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

for name, n, e in [("cora", 2708, 5278), ("pubmed", 19717, 44324), ("arxiv", 169343, 1166243)]:
    ei = graph(n, e).cuda()
    for method in ("warp", "msbfs"):
        if name == "arxiv" and method == "warp" and "--all" not in sys.argv:
            continue
        reps = 1 if name == "arxiv" else 3
        dt = t(lambda: apsp(ei, n, method=method), reps)
        print(f"{name} {method}: {dt*1e3:.2f} ms  {n*n/dt/1e9:.2f} G pairs/s = GB/s of hop bytes", flush=True)
    if name != "arxiv":
        a = apsp(ei, n, method="warp"); b = apsp(ei, n, method="msbfs")
        print("  equal:", torch.equal(a.hop[:, :n], b.hop[:, :n]), torch.equal(a.level_counts, b.level_counts))
    else:
        b = apsp(ei, n, method="msbfs"); a = apsp(ei, n, row_begin=1000, row_end=1512, method="warp")
        print("  rows equal:", torch.equal(a.hop[:, :n], b.hop[1000:1512, :n]), "D =", b.nbins - 2)
        del a, b
