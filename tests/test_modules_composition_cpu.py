"""CPU: the COMPOSITION logic of the drop-in modules (which table inputs, where the normaliser goes, per-row tables and their dedup,
row subsets, NAM readout, raw-hop batched variant with masking, graph readout shapes) against every golden case of the unmodified
reference — with the CUDA ops replaced by their dense torch definitions (tests/test_dist_modules_cpu.py). The same cases run on the
real kernels in tests/test_gpu_parity.py; this file keeps the host logic pinned where no GPU is present."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import gnan_lut
from tests import _golden as G
from tests._build import build_module
from tests.test_dist_modules_cpu import _install_torch_ops

TOL = 1e-5


def _from_reference_format(nd, nm=None):
    """stand-in for the GPU converter: hops from 1/(1+hop), level counts from the hops (pre_process_datasets.py:112-121)"""
    from gnan_b200.ops import hop_ld
    from gnan_b200.preprocess import HopData
    hop = gnan_lut.hops_from_reference(nd)
    cnt = gnan_lut.counts_from_hops(hop)
    n = hop.shape[1]
    h = torch.full((hop.shape[0], hop_ld(n)), 255, dtype=torch.uint8)
    h[:, :n] = torch.where(hop < 0, torch.full_like(hop, 255), hop).to(torch.uint8)
    return HopData(h, cnt.to(torch.int32), n)


def _grads(stacked):
    return {n: (getattr(stacked, n).grad.numpy() if isinstance(getattr(stacked, n), torch.nn.Parameter) and getattr(stacked, n).grad is not None else None)
            for n in ("w1", "b1", "wh", "bh", "wo", "bo")}


def _check(z, got, want, tag):
    for k in ("w1", "b1", "wh", "bh", "wo", "bo"):
        w = want[k]
        if w is None or w.size == 0 or got[k] is None:
            continue
        if np.linalg.norm(w) == 0:
            assert np.abs(got[k]).max() < 1e-6, (z["name"], tag, k)
        else:
            assert G.rel_err(got[k], w) < 2e-5, (z["name"], tag, k, G.rel_err(got[k], w))


@pytest.mark.parametrize("dedup", [True, False], ids=["dedup", "no_dedup"])
@pytest.mark.parametrize("name", G.MODEL_CASES + G.BATCHED_CASES)
def test_module_composition_matches_reference_golden(name, dedup, monkeypatch):
    from gnan_b200 import _inputs, ops
    _install_torch_ops(monkeypatch.setattr)
    monkeypatch.setattr(_inputs, "from_reference_format", _from_reference_format)
    z = G.load(name)
    m = build_module(z).eval()
    m.dedup = dedup
    if z["variant"] == "batched":
        out = m(torch.tensor(z["x"]), torch.tensor(z["dist_batch"]), torch.tensor(z["batch_vector"]))     # pack_dense is plain torch
    else:
        data = SimpleNamespace(x=torch.tensor(z["x"]), edge_index=torch.tensor(z["edge_index"]), node_distances=torch.tensor(z["node_distances"]),
                               normalization_matrix=torch.tensor(z["normalization_matrix"]))
        out = m.forward(data, z["node_ids"].tolist()) if (z["variant"] == "gnan_loop" and "node_ids" in z) else m.forward(data)
    (out * torch.tensor(z["out_weight"])).sum().backward()
    assert tuple(out.shape) == z["out"].shape
    assert G.rel_err(out.detach().numpy(), z["out"]) < TOL
    _check(z, _grads(m.fs), z["grad_fs"], "fs")
    _check(z, _grads(m.rho), z["grad_rho"], "rho")
    if "grad_readout" in z:
        _check(z, _grads(m.readout_nam.fs), z["grad_readout"], "readout")
    assert ops.mlp.__module__ == __name__.replace("test_modules_composition_cpu", "test_dist_modules_cpu")   # the stand-ins really ran


@pytest.mark.parametrize("name", ["trainer_graph_bce", "trainer_node_ce"])
def test_trainer_epochs_match_reference_trainer_on_cpu(name, monkeypatch):
    """gnan_b200.trainer.train_epoch / test_epoch (label handling, masks, [C,1] -> logits row, per-epoch accumulation, AUC) around the
    real models.TensorGNAN with the stand-in ops, against the UNMODIFIED trainer.py + models.TensorGNAN + Adam runs recorded in
    tests/golden/trainer_*.npz: every epoch's (loss, accuracy, auc) tuple and the final weights. CPU twin of the GPU test of the same name."""
    from gnan_b200 import _inputs, trainer
    from gnan_b200.models import TensorGNAN
    from oracle import apsp as oapsp
    _install_torch_ops(monkeypatch.setattr)
    z = dict(np.load(f"{G.GOLDEN_DIR}/{name}.npz"))
    graph_task, n_items, K, C, H, epochs, compute_auc, _ = [int(t) for t in z["meta"]]
    lr, wd = [float(t) for t in z["hyper"]]
    m = TensorGNAN(K, C, 3, H, is_graph_task=bool(graph_task), readout_n_layers=0)
    m.load_state_dict({k[4:]: torch.tensor(v) for k, v in z.items() if k.startswith("sd0.")}, strict=True)
    items = []
    for i in range(n_items):
        x = torch.tensor(z[f"item{i}.x"])
        ei = z[f"item{i}.edge_index"]
        hop = np.asarray(oapsp.apsp(ei, x.shape[0]))
        nd, _nm = oapsp.reference_format(hop, oapsp.level_counts(hop))
        d = SimpleNamespace(x=x, edge_index=torch.tensor(ei), hop_data=_from_reference_format(torch.tensor(nd)), y=torch.tensor(z[f"item{i}.y"]))
        if not graph_task:
            for mk in ("train_mask", "val_mask", "test_mask"):
                setattr(d, mk, torch.tensor(z[f"item{i}.{mk}"]))
        items.append(d)
    loss_fn = torch.nn.BCEWithLogitsLoss() if graph_task else torch.nn.CrossEntropyLoss()
    opt = torch.optim.Adam(params=m.parameters(), lr=lr, weight_decay=wd)
    hist = []
    for _ in range(epochs):
        tr = trainer.train_epoch(m, dloader=items, loss_fn=loss_fn, optimizer=opt, classify=True, device="cpu", compute_auc=bool(compute_auc),
                                 is_graph_task=bool(graph_task))
        va = trainer.test_epoch(m, dloader=items, loss_fn=loss_fn, classify=True, device="cpu", val_mask=True, compute_auc=bool(compute_auc),
                                is_graph_task=bool(graph_task))
        m.train()
        hist.append(list(tr) + list(va))
    hist, want = np.array(hist, dtype=np.float64), z["history"]
    assert hist.shape == want.shape
    assert np.allclose(hist[:, [0, 3]], want[:, [0, 3]], rtol=1e-4, atol=0), (hist[:, [0, 3]], want[:, [0, 3]])   # losses
    assert np.allclose(hist[:, [1, 4]], want[:, [1, 4]], atol=1e-6)                                              # accuracies
    assert np.allclose(hist[:, [2, 5]], want[:, [2, 5]], atol=1e-6)                                              # auc / -1
    sd = m.state_dict()
    for k, v in z.items():
        if k.startswith("sd1."):
            assert G.rel_err(sd[k[4:]].numpy(), v) < 2e-3, k


@pytest.mark.parametrize("name", ["gnanpy_tensor_graph", "models_tensor_node_shared_rho", "gnan_loop_shared_rho", "batched_graph", "gnanpy_tensor_node_l1"])
def test_interpretability_tables_on_cpu(name, monkeypatch):
    """interpret.shape_function_table / distance_function_table / heatmap (bulk evaluation through the grouped-MLP op with a
    block-diagonal output layer) against the per-point `model.fs[k](t)` / `model.rho(t)` calls of the notebook (cells 4-9), on the
    golden weights of the reference. CPU twin of the GPU test; the accessors themselves are checked against the live reference
    modules in tests/test_modules_cpu.py."""
    from gnan_b200 import interpret
    _install_torch_ops(monkeypatch.setattr)
    z = G.load(name)
    m = build_module(z).eval()
    grid = torch.linspace(-2.0, 2.0, 11)
    got = interpret.shape_function_table(m, grid)
    assert got.shape == (11, z["K"], m.fs.out_channels)
    with torch.no_grad():
        for k in range(z["K"]):
            assert torch.allclose(got[:, k], m.fs[k](grid.view(-1, 1)), atol=1e-5), k
        D = 6
        raw = z["variant"] == "batched"
        r = interpret.distance_function_table(m, D)
        u = torch.arange(D + 1, dtype=torch.float32) if raw else 1.0 / (1.0 + torch.arange(D + 1, dtype=torch.float32))
        assert torch.allclose(r, m.rho(u.view(-1, 1)), atol=1e-5)
        hm = interpret.heatmap(m, D)
        f1 = torch.stack([m.fs[k](torch.ones(1, 1))[0, 0] for k in range(z["K"])])
        assert hm.shape == (z["K"], D + 1) and torch.allclose(hm, torch.outer(f1, r[:, 0]), atol=1e-5)


def _install_torch_entries_ops(setter):
    """dense torch definitions of the compressed-feature ops (gnan_b200.sparse): one evaluation per entry, rows rebuilt as
    baseline sum + exceptions"""
    from gnan_b200 import ops, sparse

    def mlp_entries_fwd(val, grp_ptr, items, w1, b1, wh, bh, wo, bo, n_layers, max_group, precision):
        grp = torch.repeat_interleave(torch.arange(grp_ptr.numel() - 1), grp_ptr[1:] - grp_ptr[:-1])
        if n_layers == 1:
            return val[:, None] * wo[grp, :, 0] + bo[grp]
        h = torch.relu(val[:, None] * w1[grp] + b1[grp])
        for l in range(n_layers - 2):
            h = torch.relu(torch.einsum("ei,eji->ej", h, wh[l][grp]) + bh[l][grp])
        return torch.einsum("eh,ech->ec", h, wo[grp]) + bo[grp]

    def entries_to_rows(Y, grp_ptr, csr_ptr, csr_eid, ent_grp, ent_row):
        base = Y[grp_ptr[:-1]]                                        # [K,C]: every feature's baseline evaluation
        S = base.sum(0, keepdim=True).repeat(csr_ptr.numel() - 1, 1)
        rows = torch.repeat_interleave(torch.arange(csr_ptr.numel() - 1), csr_ptr[1:] - csr_ptr[:-1])
        return S.index_add(0, rows, Y[csr_eid] - base[ent_grp[csr_eid].long()])

    setter(sparse, "mlp_entries_fwd", mlp_entries_fwd)
    setter(sparse, "entries_to_rows", entries_to_rows)
    del ops


@pytest.mark.parametrize("share", [True, False], ids=["shared_values", "per_entry"])
@pytest.mark.parametrize("name", [n for n in G.MODEL_CASES if "readout" not in n])
def test_compressed_feature_path_matches_reference_golden(name, share, monkeypatch):
    """The shared-evaluation path (inputs.x_compressed: baseline value per column + exceptions, optionally one evaluation per distinct
    (feature, value) pair) through the real modules against the golden outputs and gradients of the unmodified reference, which
    evaluates every (node, feature) pair: the algebra of DESIGN.md §2 and the CompressedFeatures structures, on the CPU."""
    from gnan_b200 import _inputs, sparse
    _install_torch_ops(monkeypatch.setattr)
    _install_torch_entries_ops(monkeypatch.setattr)
    monkeypatch.setattr(_inputs, "from_reference_format", _from_reference_format)
    monkeypatch.setattr(sparse, "MAX_SHARED_FRACTION", 1.0)           # share whatever coincides, however little
    z = G.load(name)
    m = build_module(z).eval()
    x = torch.tensor(z["x"])
    cx = sparse.compress_features(x, max_density=None, share_values=share)
    assert torch.equal(cx.to_dense(), x) and (cx.shared is not None) == share
    data = SimpleNamespace(x=None if m.fs.n_layers >= 2 else x, x_compressed=cx,      # a single Linear(1,C) has no shared-evaluation path (dense x needed)
                           edge_index=torch.tensor(z["edge_index"]), node_distances=torch.tensor(z["node_distances"]),
                           normalization_matrix=torch.tensor(z["normalization_matrix"]))
    out = m.forward(data, z["node_ids"].tolist()) if (z["variant"] == "gnan_loop" and "node_ids" in z) else m.forward(data)
    (out * torch.tensor(z["out_weight"])).sum().backward()
    assert tuple(out.shape) == z["out"].shape and G.rel_err(out.detach().numpy(), z["out"]) < TOL
    _check(z, _grads(m.fs), z["grad_fs"], "fs")
    _check(z, _grads(m.rho), z["grad_rho"], "rho")


def test_packed_dataset_round_trips_the_reference_preprocessing(monkeypatch, tmp_path):
    """packed.PackedDataset.from_reference on the outputs of the unmodified pre_process (golden preprocess_* files: fp32 node_distances
    / normalization_matrix per graph) -> packed uint8 blocks + level table widened to a common depth -> save / load -> to_reference():
    bit-identical fp32 matrices, for whole sets and gathered mini-batches (host logic of the format, CPU; the GPU converter is replaced
    by the oracle's)."""
    from gnan_b200 import packed
    monkeypatch.setattr(packed, "from_reference_format", _from_reference_format)
    names = [n for n in G.PREPROCESS_CASES if "node_task" not in n]
    graphs = []
    for i, n in enumerate(names):
        z = G.load(n)
        graphs.append(SimpleNamespace(x=torch.tensor(z["x_out"]), node_distances=torch.tensor(z["node_distances"]),
                                      normalization_matrix=torch.tensor(z["normalization_matrix"]), y=torch.tensor([float(i % 2)])))
    ds = packed.PackedDataset.from_reference(graphs, device="cpu")
    assert len(ds) == len(graphs) and ds.hop.dtype == torch.uint8 and ds.hop.numel() == sum(g.x.shape[0] ** 2 for g in graphs)
    assert ds.nbytes() < sum(2 * 4 * g.x.shape[0] ** 2 for g in graphs)            # 1 byte per pair (+ the level table) instead of 8
    path = tmp_path / "set.gnan_b200.pt"
    ds.save(path)
    back = packed.PackedDataset.load(path, device=None)
    for got, g in zip(back.to_reference(), graphs):
        assert torch.equal(got.node_distances, g.node_distances) and torch.equal(got.normalization_matrix, g.normalization_matrix)
        assert torch.equal(got.x, g.x) and torch.equal(got.y, g.y)
    ids = [len(graphs) - 1, 0, 0]
    pk = back.batch(ids)
    sub = packed.PackedDataset(pk.x, pk.node_off, pk.hop, pk.hop_off, pk.level_counts, pk.y, pk.max_nodes)
    for got, i in zip(sub.to_reference(), ids):
        assert torch.equal(got.node_distances, graphs[i].node_distances) and torch.equal(got.normalization_matrix, graphs[i].normalization_matrix)
