"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU paths (partitioning, the autograd-aware all-gather /
reduce-scatter of S, the fused gradient all-reduce). The aggregation itself is replaced by its dense definition in torch
here (the CUDA kernels are covered by the -m gpu tests); what is checked is that sharded == unsharded."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_world(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def _problem():
    g = torch.Generator().manual_seed(0)
    N, K, C = 37, 5, 3
    x = torch.randn(N, K, generator=g, dtype=torch.float64)
    Wm = torch.randn(K, C, generator=g, dtype=torch.float64)         # stand-in for the shape functions: S = tanh(x) @ Wm
    W = torch.randn(N, N, C, generator=g, dtype=torch.float64)       # stand-in for the pair weights
    wout = torch.randn(N, C, generator=g, dtype=torch.float64)
    return N, x, Wm, W, wout


def _row_sharded(rank, world):
    from gnan_b200.dist import all_gather_rows, allreduce_gradients, row_block
    N, x, Wm, W, wout = _problem()
    Wm = Wm.clone().requires_grad_(True)
    blocks = [row_block(N, r, world, align=4) for r in range(world)]
    sizes = [e - b for b, e in blocks]
    b, e = blocks[rank]
    s_local = torch.tanh(x[b:e]) @ Wm
    s_full = all_gather_rows(s_local, sizes)
    out = (W[b:e] * s_full.unsqueeze(0)).sum(1)                       # own rows only
    (out * wout[b:e]).sum().backward()
    allreduce_gradients([Wm])
    return out.detach(), Wm.grad.clone(), (b, e)


def test_row_sharded_equals_single_process():
    res = run_world(_row_sharded, 2)
    N, x, Wm, W, wout = _problem()
    Wm = Wm.clone().requires_grad_(True)
    out = (W * (torch.tanh(x) @ Wm).unsqueeze(0)).sum(1)
    (out * wout).sum().backward()
    got = torch.cat([res[r][0] for r in range(2)])
    assert torch.allclose(got, out.detach(), atol=1e-12)
    for r in range(2):
        assert torch.allclose(res[r][1], Wm.grad, atol=1e-12)          # identical summed gradient on every rank
    assert res[0][2][1] == res[1][2][0] and res[1][2][1] == N


def _dp_graphs(rank, world):
    from gnan_b200.dist import allreduce_gradients, balanced_ranges, broadcast_parameters
    g = torch.Generator().manual_seed(1)
    sizes = [int(v) for v in torch.randint(3, 40, (23,), generator=g)]
    ranges = balanced_ranges([n * n + 15 * n for n in sizes], world)
    lin = torch.nn.Linear(4, 2).double()
    torch.manual_seed(rank)                                            # deliberately different init ...
    with torch.no_grad():
        lin.weight.normal_()
    broadcast_parameters(lin)                                          # ... made identical
    b, e = ranges[rank]
    feats = [torch.randn(n, 4, generator=torch.Generator().manual_seed(100 + i), dtype=torch.float64) for i, n in enumerate(sizes)]
    loss = sum(lin(feats[i]).sum() for i in range(b, e))
    loss.backward()
    allreduce_gradients(lin.parameters())
    return lin.weight.grad.clone(), lin.weight.detach().clone(), ranges


def test_data_parallel_gradients_sum_over_ranks():
    res = run_world(_dp_graphs, 2)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    ranges = res[0][2]
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 23
    sizes = [int(v) for v in torch.randint(3, 40, (23,), generator=torch.Generator().manual_seed(1))]
    feats = [torch.randn(n, 4, generator=torch.Generator().manual_seed(100 + i), dtype=torch.float64) for i, n in enumerate(sizes)]
    lin = torch.nn.Linear(4, 2).double()
    with torch.no_grad():
        lin.weight.copy_(res[0][1])
    sum(lin(f).sum() for f in feats).backward()
    assert torch.allclose(res[0][0], lin.weight.grad, atol=1e-10)


def test_balanced_ranges_properties():
    from gnan_b200.dist import balanced_ranges, row_block
    costs = [5, 1, 1, 1, 8, 2, 2, 9, 1]
    for parts in (1, 2, 3, 4, 9):
        r = balanced_ranges(costs, parts)
        assert len(r) == parts and r[0][0] == 0 and r[-1][1] == len(costs)
        assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1)) and all(e > b for b, e in r)
    two = balanced_ranges(costs, 2)
    assert abs(sum(costs[two[0][0]:two[0][1]]) - sum(costs) / 2) <= max(costs)
    blocks = [row_block(169343, r, 8) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 169343 and all(b % 16 == 0 for b, _ in blocks)
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(7))


def _flat_grads_three_ranks(rank, world):
    """world 3, ragged blocks (one rank nearly empty), FlatGradients instead of the concatenating all-reduce"""
    from gnan_b200.dist import FlatGradients, all_gather_rows, row_block
    N, x, Wm, W, wout = _problem()
    Wm = Wm.clone().requires_grad_(True)
    bias = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    fg = FlatGradients([Wm, bias])
    blocks = [row_block(N, r, world, align=16) for r in range(world)]          # 16 + 16 + 5 rows
    sizes = [e - b for b, e in blocks]
    b, e = blocks[rank]
    for _ in range(2):                                                           # the second step must not accumulate
        fg.zero()
        s_local = torch.tanh(x[b:e]) @ Wm + bias
        s_full = all_gather_rows(s_local, sizes)
        out = (W[b:e] * s_full.unsqueeze(0)).sum(1)
        (out * wout[b:e]).sum().backward()
        fg.all_reduce()
    assert Wm.grad.data_ptr() == fg.flat.data_ptr()
    return out.detach(), Wm.grad.clone(), bias.grad.clone(), sizes


def test_flat_gradients_and_ragged_gather_three_ranks():
    res = run_world(_flat_grads_three_ranks, 3)
    N, x, Wm, W, wout = _problem()
    Wm = Wm.clone().requires_grad_(True)
    bias = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    out = (W * (torch.tanh(x) @ Wm + bias).unsqueeze(0)).sum(1)
    (out * wout).sum().backward()
    assert res[0][3] == [16, 16, 5]
    assert torch.allclose(torch.cat([res[r][0] for r in range(3)]), out.detach(), atol=1e-12)
    for r in range(3):
        assert torch.allclose(res[r][1], Wm.grad, atol=1e-12) and torch.allclose(res[r][2], bias.grad, atol=1e-12)
