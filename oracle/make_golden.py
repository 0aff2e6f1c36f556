"""TEST INFRASTRUCTURE ONLY (oracle/): generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference has no tests or golden vectors of its own, so these fixtures — outputs
and parameter gradients of the real reference classes on seeded synthetic inputs, and the real
pre_process() on seeded graphs — are what pins the oracle and the CUDA path.

Weights are re-drawn (normal, gain ~1) after construction: the reference's own init
(xavier_normal_ gain=0.01, GNAN.py:49-53) makes outputs ~1e-6 and parity vacuous; biases are made
non-zero so bias paths are exercised.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.pyg_shim import import_reference  # noqa: E402
from oracle import params as P  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def random_graph(rng, n, extra=None, n_isolated=0, directed=False):
    """Random tree on the first n-n_isolated nodes + extra edges; simple graph. returns edge_index [2,E] int64."""
    m = n - n_isolated
    edges = set()
    for v in range(1, m):
        u = int(rng.integers(0, v))
        edges.add((u, v))
    extra = max(1, m // 10) if extra is None else extra
    tries = 0
    while extra > 0 and tries < 1000 and m > 2:
        u, v = int(rng.integers(0, m)), int(rng.integers(0, m))
        tries += 1
        if u != v and (u, v) not in edges and (v, u) not in edges:
            edges.add((u, v)); extra -= 1
    e = sorted(edges)
    if not directed:
        e = e + [(v, u) for (u, v) in e]
    return np.array(e, dtype=np.int64).T.reshape(2, -1)


def reinit(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "weight" in name:
                fan_in = p.shape[1]
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / max(fan_in, 1)) ** 0.5 * 1.2)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)


def preprocess_with_reference(pre_py, x, edge_index, is_graph_task=True):
    """Run the reference pre_process on one graph (graph branch) or as a node task. Returns Data-like namespace."""
    d = types.SimpleNamespace(x=torch.tensor(x, dtype=torch.float32), edge_index=torch.tensor(edge_index))
    import io, contextlib, tempfile
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        if is_graph_task:
            pre_py.pre_process([d], True, "golden", processed_data_dir=tmp)
        else:
            pre_py.pre_process(d, False, "golden", processed_data_dir=tmp)
    return d


def run_and_grads(model, fwd, weight_seed):
    out = fwd()
    g = torch.Generator().manual_seed(weight_seed + 7)
    w = torch.randn(out.shape, generator=g)
    model.zero_grad()
    (out * w).sum().backward()
    return out.detach().numpy(), w.numpy()


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    f32 = lambda v: v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: f32(v) for k, v in arrs.items() if v is not None})
    print("wrote", name, "out", arrs["out"].shape if "out" in arrs else "")


def model_case(name, ref_cls, variant, rng, pre_py, *, N, K_raw, C, H, L, is_graph_task, normalize_rho,
               rho_per_feature=False, readout_n_layers=0, n_isolated=0, node_ids=None, directed=False,
               layers_kw="n_layers", onehot=False, bias=True):
    ei = random_graph(rng, N, n_isolated=n_isolated, directed=directed)
    if onehot:
        x = np.eye(K_raw, dtype=np.float32)[rng.integers(0, K_raw, size=N)]
    else:
        x = (rng.random((N, K_raw)) < 0.5) * rng.normal(size=(N, K_raw))
    data = preprocess_with_reference(pre_py, x.astype(np.float32), ei, is_graph_task=True)
    K = K_raw + 1
    kw = dict(in_channels=K, out_channels=C, hidden_channels=H, bias=bias, dropout=0.0, device="cpu",
              normalize_rho=normalize_rho, rho_per_feature=rho_per_feature)
    kw[layers_kw] = L
    if variant in ("gnanpy_tensor", "models_tensor"):
        kw["is_graph_task"] = is_graph_task
    if variant == "models_tensor":
        kw["readout_n_layers"] = readout_n_layers
    model = ref_cls(**kw)
    seed = int(rng.integers(0, 2 ** 31))
    reinit(model, seed)
    if variant == "gnan_loop":
        fwd = lambda: model.forward(data, node_ids)
    else:
        fwd = lambda: model.forward(data)
    out, w = run_and_grads(model, fwd, seed)
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    fs_pref = [f"fs.{k}" for k in range(K)]
    rho_has_bias = bias if variant == "gnan_loop" else (not is_graph_task)
    arrs = dict(x=data.x.numpy(), edge_index=ei, node_distances=data.node_distances.numpy(),
                normalization_matrix=data.normalization_matrix.numpy(), out=out, out_weight=w,
                meta=np.array([N, K, C, H, L, int(is_graph_task), int(normalize_rho), int(rho_per_feature),
                               readout_n_layers, int(rho_has_bias), int(bias)], dtype=np.int64),
                node_ids=None if node_ids is None else np.array(node_ids, dtype=np.int64))
    arrs.update(P.flatten("grad_fs", P.unstack_grads(model, fs_pref, L, 3, bias)))
    arrs.update(P.flatten("grad_rho", P.unstack_grads(model, ["rho"], L, 2, rho_has_bias)))
    if variant == "models_tensor" and is_graph_task and readout_n_layers > 0:
        rp = [f"readout_nam.fs.{k}" for k in range(K)]
        arrs.update(P.flatten("grad_readout", P.unstack_grads(model, rp, readout_n_layers, 3, bias)))
    # raw state_dict too (tests the product's load_state_dict key compatibility)
    arrs.update({f"sd.{k}": v for k, v in sd.items()})
    save(name, **arrs)


def batched_case(name, batched_cls, rng, *, B, K, C, H, is_graph_task=True):
    import networkx as nx
    xs, dists, bv = [], [], []
    for g in range(B):
        n = int(rng.integers(3, 12))
        ei = random_graph(rng, n, n_isolated=1 if g % 3 == 0 and n > 4 else 0)
        G = nx.Graph(); G.add_nodes_from(range(n)); G.add_edges_from(ei.T.tolist())
        dm = np.full((n, n), -1, dtype=np.float32)                     # batched_pyg_main.py:39-44
        for s in range(n):
            for t, dv in nx.shortest_path_length(G, source=s).items():
                dm[s, t] = dv
        xs.append(rng.normal(size=(n, K)).astype(np.float32)); dists.append(dm); bv += [g] * n
    tot = len(bv)
    dist_batch = np.full((tot, tot), -1, dtype=np.float32)             # :75-80
    o = 0
    for dm in dists:
        n = dm.shape[0]; dist_batch[o:o + n, o:o + n] = dm; o += n
    x_batch = np.concatenate(xs); bv = np.array(bv, dtype=np.int64)
    model = batched_cls(in_channels=K, out_channels=C, n_layers=2, hidden_channels=H, is_graph_task=is_graph_task)
    seed = int(rng.integers(0, 2 ** 31)); reinit(model, seed)
    out, w = run_and_grads(model, lambda: model(torch.tensor(x_batch), torch.tensor(dist_batch), torch.tensor(bv)), seed)
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    fs_pref = [f"fs.{k}" for k in range(K)]
    arrs = dict(x=x_batch, dist_batch=dist_batch, batch_vector=bv, out=out, out_weight=w,
                sizes=np.array([d.shape[0] for d in dists], dtype=np.int64),
                meta=np.array([tot, K, C, H, 2, int(is_graph_task)], dtype=np.int64))
    arrs.update(P.flatten("grad_fs", P.unstack_grads(model, fs_pref, 2, 3)))
    arrs.update(P.flatten("grad_rho", P.unstack_grads(model, ["rho"], 2, 3)))
    arrs.update({f"sd.{k}": v for k, v in sd.items()})
    save(name, **arrs)


def preprocess_case(name, pre_py, rng, n, n_isolated, directed, node_task=False):
    ei = random_graph(rng, n, n_isolated=n_isolated, directed=directed)
    if node_task and n_isolated:                       # keep the last node connected: the node branch infers
        ei = np.concatenate([ei, np.array([[0], [n - 1]])], axis=1)   # num_nodes from max index (:128)
        if not directed:
            ei = np.concatenate([ei, np.array([[n - 1], [0]])], axis=1)
    x = rng.normal(size=(n, 3)).astype(np.float32)
    d = preprocess_with_reference(pre_py, x, ei, is_graph_task=not node_task)
    save(name, x=x, edge_index=ei, x_out=d.x.numpy(), node_distances=d.node_distances.numpy(),
         normalization_matrix=d.normalization_matrix.numpy(), meta=np.array([n, int(directed), int(node_task)]))


class _Data(types.SimpleNamespace):
    """The slice of torch_geometric.data.Data that trainer.py touches (attributes + .to)."""

    def to(self, device):
        return self


def trainer_case(name, models_py, pre_py, trainer_py, rng, *, graph_task, n_items, K_raw, C, H, epochs, lr, wd=0.0,
                 compute_auc=False, pm1_labels=False):
    """Run the UNMODIFIED trainer.train_epoch / test_epoch (trainer.py:23-160) with models.TensorGNAN (what main.py builds,
    main.py:84-90) and Adam (main.py:141) on a tiny seeded dataset; record the per-epoch return tuples and final weights."""
    items = []
    if graph_task:
        for _ in range(n_items):
            n = int(rng.integers(5, 14))
            ei = random_graph(rng, n, n_isolated=1 if n > 9 else 0)
            x = np.eye(K_raw, dtype=np.float32)[rng.integers(0, K_raw, size=n)]
            d = preprocess_with_reference(pre_py, x, ei, is_graph_task=True)
            yv = float(rng.integers(0, 2))
            y = torch.tensor([2 * yv - 1 if pm1_labels else yv])
            items.append(_Data(x=d.x, edge_index=d.edge_index, node_distances=d.node_distances,
                               normalization_matrix=d.normalization_matrix, y=y))
        loss_fn = torch.nn.BCEWithLogitsLoss()
    else:
        n = 60
        ei = random_graph(rng, n, n_isolated=2)
        ei = np.concatenate([ei, np.array([[0, n - 1], [n - 1, 0]])], axis=1)
        x = ((rng.random((n, K_raw)) < 0.5) * rng.normal(size=(n, K_raw))).astype(np.float32)
        d = preprocess_with_reference(pre_py, x, ei, is_graph_task=False)
        split = rng.permutation(n)
        mk = lambda ids: torch.zeros(n, dtype=torch.bool).index_fill_(0, torch.tensor(ids), True)
        items.append(_Data(x=d.x, edge_index=d.edge_index, node_distances=d.node_distances,
                           normalization_matrix=d.normalization_matrix, y=torch.tensor(rng.integers(0, C, size=n)),
                           train_mask=mk(split[:30]), val_mask=mk(split[30:45]), test_mask=mk(split[45:])))
        loss_fn = torch.nn.CrossEntropyLoss()
    K = K_raw + 1
    model = models_py.TensorGNAN(in_channels=K, out_channels=C, n_layers=3, hidden_channels=H, bias=True, dropout=0.0,
                                 device="cpu", normalize_rho=True, is_graph_task=graph_task, readout_n_layers=0)
    reinit(model, int(rng.integers(0, 2 ** 31)))
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(params=model.parameters(), lr=lr, weight_decay=wd)
    hist = []
    import io, contextlib
    for _ in range(epochs):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            tr = trainer_py.train_epoch(model, dloader=items, loss_fn=loss_fn, optimizer=opt, classify=True, device="cpu",
                                        compute_auc=compute_auc, is_graph_task=graph_task)
            va = trainer_py.test_epoch(model, dloader=items, loss_fn=loss_fn, classify=True, device="cpu", val_mask=True,
                                       compute_auc=compute_auc, is_graph_task=graph_task)
        model.train()
        hist.append([float(t) for t in tr] + [float(t) for t in va])
    arrs = dict(history=np.array(hist, dtype=np.float64),
                meta=np.array([int(graph_task), len(items), K, C, H, epochs, int(compute_auc), int(pm1_labels)], dtype=np.int64),
                hyper=np.array([lr, wd], dtype=np.float64))
    for i, it in enumerate(items):
        arrs[f"item{i}.x"] = it.x.numpy()
        arrs[f"item{i}.edge_index"] = it.edge_index.numpy()
        arrs[f"item{i}.y"] = it.y.numpy()
        if not graph_task:
            for m in ("train_mask", "val_mask", "test_mask"):
                arrs[f"item{i}.{m}"] = getattr(it, m).numpy()
    arrs.update({f"sd0.{k}": v for k, v in sd0.items()})
    arrs.update({f"sd1.{k}": v.detach().numpy() for k, v in model.state_dict().items()})
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, "history", np.array(hist)[-1])


def trainer_cases():
    """Separate entry (python -m oracle.make_golden trainer): own seed, leaves the other fixtures untouched."""
    torch.manual_seed(0)
    torch.set_num_threads(4)
    _, models_py, pre_py, _ = import_reference()
    import importlib
    trainer_py = importlib.import_module("trainer")
    rng = np.random.default_rng(20240917)
    trainer_case("trainer_graph_bce", models_py, pre_py, trainer_py, rng, graph_task=True, n_items=10, K_raw=5, C=1, H=16,
                 epochs=3, lr=5e-3, compute_auc=True, pm1_labels=True)
    trainer_case("trainer_node_ce", models_py, pre_py, trainer_py, rng, graph_task=False, n_items=1, K_raw=6, C=3, H=64,
                 epochs=6, lr=5e-3, wd=1e-4)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    gnan_py, models_py, pre_py, batched_cls = import_reference()
    rng = np.random.default_rng(20240601)
    # --- preprocessing
    preprocess_case("preprocess_tree_undirected", pre_py, rng, 23, 0, False)
    preprocess_case("preprocess_isolated_undirected", pre_py, rng, 31, 4, False)
    preprocess_case("preprocess_directed", pre_py, rng, 27, 2, True)
    preprocess_case("preprocess_node_task", pre_py, rng, 40, 3, True, node_task=True)
    preprocess_case("preprocess_single_node", pre_py, rng, 1, 0, False)
    # --- GNAN.py::TensorGNAN (a2-a6), input-normalised rho, C-wide rho
    model_case("gnanpy_tensor_node", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=37, K_raw=5, C=3, H=64, L=3,
               is_graph_task=False, normalize_rho=True, n_isolated=2)
    model_case("gnanpy_tensor_node_nonorm_l2", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=29, K_raw=4, C=7, H=32,
               L=2, is_graph_task=False, normalize_rho=False)
    model_case("gnanpy_tensor_node_l1", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=21, K_raw=3, C=2, H=8, L=1,
               is_graph_task=False, normalize_rho=True)
    model_case("gnanpy_tensor_node_l4", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=26, K_raw=3, C=4, H=16, L=4,
               is_graph_task=False, normalize_rho=True, n_isolated=1)
    model_case("gnanpy_tensor_graph", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=30, K_raw=14, C=1, H=64, L=3,
               is_graph_task=True, normalize_rho=True, onehot=True)
    model_case("gnanpy_tensor_graph_nonorm_disconnected", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=18, K_raw=14,
               C=1, H=64, L=3, is_graph_task=True, normalize_rho=False, onehot=True, n_isolated=3)
    model_case("gnanpy_tensor_node_directed_nobias", gnan_py.TensorGNAN, "gnanpy_tensor", rng, pre_py, N=25, K_raw=4, C=3,
               H=64, L=3, is_graph_task=False, normalize_rho=True, directed=True, bias=False)
    # --- models.py::TensorGNAN (what main.py builds): output-normalised, rho width 1 unless rho_per_feature
    model_case("models_tensor_graph", models_py.TensorGNAN, "models_tensor", rng, pre_py, N=28, K_raw=14, C=1, H=64, L=3,
               is_graph_task=True, normalize_rho=True, readout_n_layers=0, onehot=True)
    model_case("models_tensor_node_shared_rho", models_py.TensorGNAN, "models_tensor", rng, pre_py, N=33, K_raw=6, C=5, H=64,
               L=3, is_graph_task=False, normalize_rho=True, n_isolated=2)
    model_case("models_tensor_node_rho_per_feature", models_py.TensorGNAN, "models_tensor", rng, pre_py, N=31, K_raw=4, C=4,
               H=32, L=3, is_graph_task=False, normalize_rho=True, rho_per_feature=True)
    model_case("models_tensor_graph_readout", models_py.TensorGNAN, "models_tensor", rng, pre_py, N=24, K_raw=6, C=3, H=16,
               L=3, is_graph_task=True, normalize_rho=True, readout_n_layers=2)
    # --- GNAN (row loop), both copies identical; main.py passes num_layers= (models.py:388)
    model_case("gnan_loop_shared_rho", models_py.GNAN, "gnan_loop", rng, pre_py, N=35, K_raw=5, C=4, H=64, L=3,
               is_graph_task=False, normalize_rho=True, layers_kw="num_layers", n_isolated=1)
    model_case("gnan_loop_rho_per_feature_rows", gnan_py.GNAN, "gnan_loop", rng, pre_py, N=32, K_raw=4, C=3, H=64, L=3,
               is_graph_task=False, normalize_rho=True, rho_per_feature=True, node_ids=[3, 0, 17, 31, 8])
    model_case("gnan_loop_nonorm", gnan_py.GNAN, "gnan_loop", rng, pre_py, N=20, K_raw=3, C=2, H=16, L=2,
               is_graph_task=False, normalize_rho=False)
    # --- batched block-diagonal variant
    batched_case("batched_graph", batched_cls, rng, B=6, K=5, C=8, H=16, is_graph_task=True)
    batched_case("batched_node", batched_cls, rng, B=3, K=4, C=2, H=16, is_graph_task=False)


if __name__ == "__main__":
    trainer_cases() if sys.argv[1:] == ["trainer"] else main()
