"""CPU, world_size 2 over gloo: `gnan_b200.dist.row_sharded_forward` and the data-parallel packed-batch step run with the REAL modules
(GNAN.py / models.py flavours, stacked parameters, per-row table dedup, HopData row shards, FlatGradients) — only the CUDA ops are
replaced by their dense definitions in torch (the kernels themselves are covered by the -m gpu tests). Checked: sharded == unsharded ==
the float64 oracle, for outputs and every parameter gradient."""
import numpy as np
import pytest
import torch

from tests.test_dist_cpu import run_world


def _install_torch_ops(setter=setattr):
    """dense torch definitions of the ops the modules call (include/gnan_b200.h semantics); setter: setattr in a spawned worker,
    monkeypatch.setattr inside the pytest process"""
    from gnan_b200 import ops

    def mlp(u, w1, b1, wh, bh, wo, bo, n_layers, dropout_p=0.0, seed=0, precision="fp32", seed_dev=None):
        if n_layers == 1:
            return torch.einsum("rg,gc->rc", u, wo[:, :, 0]) + bo.sum(0)
        h = torch.relu(u[:, :, None] * w1[None] + b1[None])
        for l in range(n_layers - 2):
            h = torch.relu(torch.einsum("rgi,gji->rgj", h, wh[l]) + bh[l][None])
        return torch.einsum("rgh,gch->rc", h, wo) + bo.sum(0)

    def rho_table_inputs(nbins, device, cnt=None, raw=False):
        d = torch.arange(nbins, dtype=torch.float32)
        u = d.clone() if raw else torch.where(d == nbins - 1, torch.zeros(()), 1.0 / (d + 1.0))
        if cnt is None:
            return u
        c = cnt.float()
        return torch.where(c > 0, u[None] / c.clamp(min=1), torch.zeros(()))

    def level_rscale(cnt):
        c = cnt.float()
        return torch.where(c > 0, 1.0 / c.clamp(min=1), torch.zeros(()))

    def bins_of(hop, n, nbins):
        b = hop[:, :n].long()
        return torch.where(b == 255, torch.full_like(b, nbins - 1), b)

    def aggregate_rows(hop, T, S, rscale=None, per_row=False, algo=None):
        nbins, Cr = T.shape[-2], T.shape[-1]
        b = bins_of(hop, S.shape[0], nbins)
        Tg = torch.gather(T, 1, b[:, :, None].expand(-1, -1, Cr)) if per_row else T[b]
        if rscale is not None:
            Tg = Tg * torch.gather(rscale, 1, b)[:, :, None]
        return (Tg * S[None]).sum(1)

    def aggregate_blockdiag(hop, hop_off, node_off, T, S, rscale=None, per_row=False, reduce_graph=False):
        outs = []
        for g in range(node_off.numel() - 1):
            r0, r1 = int(node_off[g]), int(node_off[g + 1])
            n = r1 - r0
            blk = hop[int(hop_off[g]):int(hop_off[g + 1])].view(n, n)
            o = aggregate_rows(blk, T[r0:r1] if per_row else T, S[r0:r1], None if rscale is None else rscale[r0:r1], per_row)
            outs.append(o.sum(0, keepdim=True) if reduce_graph else o)
        return torch.cat(outs)

    for name, fn in (("mlp", mlp), ("rho_table_inputs", rho_table_inputs), ("level_rscale", level_rscale), ("aggregate_rows", aggregate_rows),
                     ("aggregate_blockdiag", aggregate_blockdiag), ("gather_rows", lambda tq, inv, order, seg_ptr: tq[inv])):
        setter(ops, name, fn)


def _node_problem():
    from oracle import apsp as oapsp
    rng = np.random.default_rng(5)
    n, K, C = 45, 6, 3
    src, dst = rng.integers(0, n - 3, size=70), rng.integers(0, n - 3, size=70)      # the last 3 nodes stay isolated
    e = np.stack([src[src != dst], dst[src != dst]])
    e = np.unique(np.concatenate([e, e[::-1]], axis=1), axis=1).astype(np.int64)
    hop = np.asarray(oapsp.apsp(e, n))
    cnt = oapsp.level_counts(hop)
    x = torch.tensor(rng.normal(size=(n, K))).float()
    x[:, -1] = 1.0
    w = torch.tensor(rng.normal(size=(n, C))).float()
    return n, K, C, hop, cnt, x, w


def _build(flavour, K, C):
    from gnan_b200 import GNAN as g, models as mo
    torch.manual_seed(3)
    if flavour == "gnanpy_tensor":
        m = g.TensorGNAN(K, C, 3, 16, normalize_rho=True, is_graph_task=False)           # input-normalised, per-row tables (dedup path)
    elif flavour == "models_tensor":
        m = mo.TensorGNAN(K, C, 3, 16, normalize_rho=True, is_graph_task=False, rho_per_feature=True)
    else:
        m = g.GNAN(K, C, 2, 16, normalize_rho=True)                                      # rho of width 1, output-normalised
    m.fs.xavier_normal_(1.0); m.rho.xavier_normal_(1.0)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() <= 2 and p.numel() and float(p.abs().max()) == 0.0:
                p.normal_(0.0, 0.3)                                                     # biases take part
    return m


def _hop_data(hop, cnt, b, e, n):
    from gnan_b200.ops import hop_ld
    from gnan_b200.preprocess import HopData
    h = torch.full((e - b, hop_ld(n)), 255, dtype=torch.uint8)
    blk = torch.tensor(hop[b:e])
    h[:, :n] = torch.where(blk < 0, torch.full_like(blk, 255), blk).to(torch.uint8)
    return HopData(h, torch.tensor(cnt[b:e]).to(torch.int32), n, row_begin=b)


def _sharded(rank, world, flavour):
    _install_torch_ops()
    from gnan_b200.dist import FlatGradients, broadcast_parameters, row_block, row_sharded_forward
    n, K, C, hop, cnt, x, w = _node_problem()
    model = _build(flavour, K, C)
    broadcast_parameters(model)
    blocks = [row_block(n, r, world, align=16) for r in range(world)]
    b, e = blocks[rank]
    fg = FlatGradients(model.parameters())
    fg.zero()
    out = row_sharded_forward(model, x[b:e], _hop_data(hop, cnt, b, e, n), [q - p for p, q in blocks])
    (out * w[b:e]).sum().backward()
    fg.all_reduce()
    assert all(p.grad.data_ptr() == fg.flat.data_ptr() + 4 * o for p, o in zip(fg.params, fg._offsets()) if p.numel())     # still views of the flat buffer
    return out.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}, (b, e)


FLAVOURS = ["gnanpy_tensor", "models_tensor", "gnan_loop"]
_WORLD_CACHE = {}


def _sharded_all(rank, world):
    return {f: _sharded(rank, world, f) for f in FLAVOURS}


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_row_sharded_modules_equal_single_process_and_oracle(flavour, monkeypatch):
    if "res" not in _WORLD_CACHE:                      # one world-size-2 launch serves the three flavours (spawning costs ~12 s)
        _WORLD_CACHE["res"] = run_world(_sharded_all, 2)
    res = {r: _WORLD_CACHE["res"][r][flavour] for r in range(2)}
    _install_torch_ops(monkeypatch.setattr)
    from types import SimpleNamespace

    from oracle import gnan_lut, gnan_port, params as P
    n, K, C, hop, cnt, x, w = _node_problem()
    model = _build(flavour, K, C)
    out = model.forward(SimpleNamespace(x=x, hop_data=_hop_data(hop, cnt, 0, n, n)))      # the single-process module, same stand-in ops
    (out * w).sum().backward()
    got = torch.cat([res[r][0] for r in range(2)])
    assert res[0][2][1] == res[1][2][0] and res[1][2][1] == n and res[0][2][1] % 16 == 0
    assert float((got - out.detach()).norm() / out.detach().norm()) < 1e-6
    for k, p in model.named_parameters():
        if p.numel() == 0:                                  # the unused hidden-layer stack of a 2-layer MLP
            continue
        for r in range(2):
            assert float((res[r][1][k] - p.grad).norm() / p.grad.norm().clamp(min=1e-30)) < 1e-5, (k, r)
    # and both against the float64 table restatement of the reference (oracle/gnan_lut.py)
    L = model.fs.n_layers
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items() if not k.startswith("rhos.")}
    fs = gnan_port.to_torch(P.stack_mlps(sd, [f"fs.{k}" for k in range(K)], L, 3), torch.float64, True)
    rho = gnan_port.to_torch(P.stack_mlps(sd, ["rho"], L, 2), torch.float64, True)
    mode = "input" if flavour == "gnanpy_tensor" else "output"
    want = gnan_lut.forward_rows(fs, rho, x.double(), torch.tensor(hop).long(), gnan_lut.counts_from_hops(torch.tensor(hop).long()), mode)
    (want * w.double()).sum().backward()
    assert float((got.double() - want.detach()).norm() / want.detach().norm()) < 1e-5
    assert float((res[0][1]["fs.wh" if L > 2 else "fs.wo"].double() - fs["wh" if L > 2 else "wo"].grad).norm() / fs["wh" if L > 2 else "wo"].grad.norm()) < 1e-4
    assert float((res[0][1]["rho.wo"].double() - rho["wo"].grad).norm() / rho["wo"].grad.norm()) < 1e-4


def _graph_problem():
    from oracle import apsp as oapsp
    rng = np.random.default_rng(11)
    graphs = []
    for _ in range(9):
        n = int(rng.integers(3, 14))
        e = np.array([[int(rng.integers(0, v)), v] for v in range(1, n - (1 if n > 8 else 0))]).T.reshape(2, -1)   # tree; larger graphs keep an isolated node
        e = np.unique(np.concatenate([e, e[::-1]], axis=1), axis=1).astype(np.int64)
        hop = np.asarray(oapsp.apsp(e, n))
        x = np.eye(4, dtype=np.float32)[rng.integers(0, 4, size=n)]
        graphs.append((hop, np.concatenate([x, np.ones((n, 1), np.float32)], 1), float(rng.integers(0, 2))))
    return graphs


def _packed(graphs, nbins):
    from gnan_b200.preprocess import PackedBatch
    from oracle import apsp as oapsp
    sizes = [g[0].shape[0] for g in graphs]
    hop = torch.cat([torch.tensor(np.where(g[0] < 0, 255, g[0]).astype(np.uint8)).reshape(-1) for g in graphs])
    cnt = torch.zeros(sum(sizes), nbins, dtype=torch.int32)
    r = 0
    for g in graphs:
        c = torch.tensor(oapsp.level_counts(g[0])).to(torch.int32)
        cnt[r:r + c.shape[0], :c.shape[1] - 1] = c[:, :-1]
        cnt[r:r + c.shape[0], -1] = c[:, -1]
        r += c.shape[0]
    node_off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    hop_off = torch.tensor(np.concatenate([[0], np.cumsum(np.square(sizes))]), dtype=torch.int64)
    return PackedBatch(torch.tensor(np.concatenate([g[1] for g in graphs])), hop, hop_off, node_off, cnt,
                       torch.tensor([g[2] for g in graphs]), max(sizes))


def _dp_step(rank, world):
    _install_torch_ops()
    from gnan_b200 import models as mo
    from gnan_b200.dist import FlatGradients, balanced_ranges, broadcast_parameters
    graphs = _graph_problem()
    torch.manual_seed(100 + rank)                                      # different initial weights per rank ...
    model = mo.TensorGNAN(5, 1, 3, 16, normalize_rho=True, is_graph_task=True, readout_n_layers=0)
    model.fs.xavier_normal_(1.0); model.rho.xavier_normal_(1.0)
    broadcast_parameters(model)                                        # ... made identical
    b, e = balanced_ranges([g[0].shape[0] ** 2 + 5 * g[0].shape[0] for g in graphs], world)[rank]
    pk = _packed(graphs[b:e], 16)
    fg = FlatGradients(model.parameters())
    fg.zero()
    out = model(pk)                                                    # [B_r, 1]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out.flatten(), pk.y, reduction="sum") / len(graphs)
    loss.backward()
    fg.all_reduce()                                                    # sum of the per-rank partial means = the global mean's gradient
    return out.detach(), fg.flat.clone(), {k: v.detach().clone() for k, v in model.state_dict().items()}, (b, e)


def test_data_parallel_packed_batches_equal_single_process(monkeypatch):
    res = run_world(_dp_step, 2)
    _install_torch_ops(monkeypatch.setattr)
    from gnan_b200 import models as mo
    from gnan_b200.dist import FlatGradients
    graphs = _graph_problem()
    assert res[0][3][0] == 0 and res[0][3][1] == res[1][3][0] and res[1][3][1] == len(graphs) and 0 < res[0][3][1] < len(graphs)
    assert torch.equal(res[0][1], res[1][1])                           # identical reduced gradients on both ranks
    model = mo.TensorGNAN(5, 1, 3, 16, normalize_rho=True, is_graph_task=True, readout_n_layers=0)
    model.load_state_dict(res[0][2], strict=True)
    assert all(torch.equal(res[0][2][k], res[1][2][k]) for k in res[0][2])
    fg = FlatGradients(model.parameters())
    fg.zero()
    pk = _packed(graphs, 16)
    out = model(pk)
    (torch.nn.functional.binary_cross_entropy_with_logits(out.flatten(), pk.y, reduction="sum") / len(graphs)).backward()
    got = torch.cat([res[0][0], res[1][0]])
    assert float((got - out.detach()).norm() / out.detach().norm()) < 1e-6
    assert float((res[0][1] - fg.flat).norm() / fg.flat.norm()) < 1e-5
    # per graph against the reference-format call of the same module (one graph per forward, as the reference trains)
    from types import SimpleNamespace
    from gnan_b200.ops import hop_ld
    from gnan_b200.preprocess import HopData
    from oracle import apsp as oapsp
    for i in (0, 4, 8):
        hop, x, _ = graphs[i]
        n = hop.shape[0]
        h = torch.full((n, hop_ld(n)), 255, dtype=torch.uint8)
        h[:, :n] = torch.tensor(np.where(hop < 0, 255, hop).astype(np.uint8))
        one = model(SimpleNamespace(x=torch.tensor(x), hop_data=HopData(h, torch.tensor(oapsp.level_counts(hop)).to(torch.int32), n)))
        assert one.shape == (1, 1) and abs(float(one.detach()) - float(out[i].detach())) < 1e-5 * max(1.0, abs(float(out[i].detach())))
