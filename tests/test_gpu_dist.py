"""GPU, 2 ranks over NCCL (skipped on a single-GPU box; run with `gpurun --gpus 2`): the row-sharded node-level path and
the data-parallel packed-graph path equal their single-GPU counterparts."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph(rng, n, deg=2.5):
    E = int(n * deg / 2)
    s = rng.integers(0, n, size=E); d = rng.integers(0, n, size=E)
    e = np.unique(np.stack([s[s != d], d[s != d]], 1), axis=0)
    return np.unique(np.concatenate([e, e[:, ::-1]]), axis=0).T.astype(np.int64)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gnan_b200 import dist as gdist
        from gnan_b200.models import GNAN
        from gnan_b200.preprocess import apsp
        rng = np.random.default_rng(5)
        n, K, C = 1500, 12, 3
        ei = _graph(rng, n)
        x = torch.tensor(rng.normal(size=(n, K))).float()
        w = torch.tensor(rng.normal(size=(n, C))).float().to(dev)
        torch.manual_seed(1)
        m = GNAN(K, C, num_layers=3, hidden_channels=64, rho_per_feature=True).to(dev)
        gdist.broadcast_parameters(m)
        # single-GPU result (every rank computes it)
        full = m.forward(SimpleNamespace(x=x, hop_data=apsp(torch.tensor(ei), n, device=dev)))
        (full * w).sum().backward()
        g_full = {k: p.grad.clone() for k, p in m.named_parameters()}
        m.zero_grad()
        blocks = [gdist.row_block(n, r, world) for r in range(world)]
        b, e = blocks[rank]
        hd = apsp(torch.tensor(ei), n, device=dev, row_begin=b, row_end=e)
        out = gdist.row_sharded_forward(m, x[b:e].to(dev), hd, [q - p for p, q in blocks])
        (out * w[b:e]).sum().backward()
        gdist.allreduce_gradients(m.parameters())
        rel = lambda a, c: float((a - c).norm() / c.norm())
        errs = {"out": rel(out, full[b:e].detach())}
        errs.update({k: rel(p.grad, g_full[k]) for k, p in m.named_parameters()})
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


def test_row_sharded_two_gpus_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for r in range(2):
        for k, v in ret[r].items():
            assert v < 2e-5, (r, k, v)
