"""torchrun --nproc-per-node 2 tests/tools/diag_parity_n2.py : per-parameter errors of the row-sharded step vs the single-GPU step
(bench.parity_vs_single_gpu's problem), and both against a float64 evaluation of d(bo)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gnan_b200 import dist as gdist  # noqa: E402
from gnan_b200 import ops  # noqa: E402
from gnan_b200.GNAN import TensorGNAN  # noqa: E402
from gnan_b200.preprocess import apsp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
rel = lambda a, c: float((a.double() - c.double()).norm() / c.double().norm().clamp_min(1e-300))
rng = np.random.default_rng(5)
n, K, C = 4000, 24, 5
ei = bench.random_simple_graph(rng, n, 7000, 30)
x = torch.tensor(rng.normal(size=(n, K))).float().to(dev)
w = torch.tensor(rng.normal(size=(n, C))).float().to(dev)
w[torch.tensor(rng.random(n) > 0.05).to(dev)] = 0.0
torch.manual_seed(1)
m = TensorGNAN(K, C, bench.L, bench.H, normalize_rho=True, device=dev).to(dev)
m.fs.xavier_normal_(1.0); m.rho.xavier_normal_(1.0)
gdist.broadcast_parameters(m)
hd_full = apsp(torch.tensor(ei), n, device=dev)
res = {}
for algo in ("auto", "cuda"):
    ops.AGG_ALGO = algo
    m.zero_grad(set_to_none=True)
    full = m(SimpleNamespace(x=x, hop_data=hd_full))
    (full * w).sum().backward()
    g_full = {k: p.grad.clone() for k, p in m.named_parameters()}
    fg = gdist.FlatGradients(m.parameters())
    fg.zero()
    blocks = [gdist.row_block(n, r, world) for r in range(world)]
    b, e = blocks[rank]
    hd = apsp(torch.tensor(ei), n, device=dev, row_begin=b, row_end=e)
    o = gdist.row_sharded_forward(m, x[b:e].contiguous(), hd, [q - p for p, q in blocks])
    (o * w[b:e]).sum().backward()
    fg.all_reduce()
    per = {k: rel(p.grad, g_full[k]) for k, p in m.named_parameters()}
    res[algo] = (per, g_full["fs.bo"].clone() if "fs.bo" in g_full else None, dict(m.named_parameters())["fs.bo"].grad.clone())
    fg = None
    for p in m.parameters():
        p.grad = None
if rank == 0:
    for algo, (per, gf, gs) in res.items():
        print(algo, {k: f"{v:.2e}" for k, v in per.items()})
        print("  d(bo) full   ", gf[0].tolist())
        print("  d(bo) sharded", gs[0].tolist())
    print("cuda-vs-auto full d(bo):", rel(res["auto"][1], res["cuda"][1]), " sharded:", rel(res["auto"][2], res["cuda"][2]))
    for k in res["auto"][0]:
        pass
dist.destroy_process_group()
