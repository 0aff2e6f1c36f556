"""CPU: pin the oracle (oracle/gnan_port.py, oracle/gnan_lut.py, oracle/apsp_oracle.c) against the golden
vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import apsp as oapsp
from oracle import gnan_lut, gnan_port
from tests import _golden as G

TOL_PORT = 2e-6   # fp32 port vs fp32 reference (same op order; measured ~1e-7)
TOL_LUT = 5e-6    # fp64 re-ordered restatement vs fp32 reference


def run_port(z, dtype):
    fs = gnan_port.to_torch(z["fs"], dtype, True)
    rho = gnan_port.to_torch(z["rho"], dtype, True)
    x = torch.tensor(z["x"]).to(dtype)
    extra = {}
    if z["variant"] == "batched":
        out = gnan_port.tensor_gnan_batched(fs, rho, x, torch.tensor(z["dist_batch"]).to(dtype),
                                            torch.tensor(z["batch_vector"]), z["is_graph_task"])
    else:
        nd = torch.tensor(z["node_distances"]).to(dtype)
        nm = torch.tensor(z["normalization_matrix"]).to(dtype)
        if z["variant"] == "gnanpy_tensor":
            out = gnan_port.tensor_gnan_gnanpy(fs, rho, x, nd, nm, z["normalize_rho"], z["is_graph_task"])
        elif z["variant"] == "models_tensor":
            ro = gnan_port.to_torch(z["readout"], dtype, True) if "readout" in z else None
            extra["readout"] = ro
            out = gnan_port.tensor_gnan_models(fs, rho, x, nd, nm, z["normalize_rho"], z["is_graph_task"], ro)
        else:
            ids = z["node_ids"].tolist() if "node_ids" in z else None
            out = gnan_port.gnan_rowloop(fs, rho, x, nd, nm, z["normalize_rho"], ids)
    (out * torch.tensor(z["out_weight"]).to(dtype)).sum().backward()
    return out, fs, rho, extra


def check_grads(z, got, key, tol):
    want = z[key]
    has_bias = z["rho_has_bias"] if key == "grad_rho" else z["bias"]
    for k in ("w1", "b1", "wh", "bh", "wo", "bo"):
        if want[k] is None or want[k].size == 0:
            continue
        if k.startswith("b") and not has_bias:      # the reference has no such parameter (GNAN.py:36-37)
            continue
        g = got[k].grad
        g = np.zeros_like(want[k]) if g is None else g.numpy()
        if np.linalg.norm(want[k]) == 0:
            assert np.abs(g).max() < 1e-6, (key, k)
        else:
            assert G.rel_err(g, want[k]) < tol, (z["name"], key, k, G.rel_err(g, want[k]))


@pytest.mark.parametrize("name", G.MODEL_CASES + G.BATCHED_CASES)
def test_port_matches_reference(name):
    z = G.load(name)
    out, fs, rho, extra = run_port(z, torch.float32)
    assert out.shape == z["out"].shape
    assert G.rel_err(out.detach().numpy(), z["out"]) < TOL_PORT
    check_grads(z, fs, "grad_fs", 2e-5)
    check_grads(z, rho, "grad_rho", 2e-5)
    if extra.get("readout") is not None:
        check_grads(z, extra["readout"], "grad_readout", 2e-5)


def lut_mode(z):
    if z["variant"] == "batched":
        return "raw"
    if not z["normalize_rho"]:
        return "none"
    return "input" if z["variant"] == "gnanpy_tensor" else "output"


def run_lut(z):
    """float64 table restatement (oracle/gnan_lut.py) of a loaded case -> (out, fs, rho) after backward of sum(out * out_weight)"""
    dt = torch.float64
    fs = gnan_port.to_torch(z["fs"], dt, True)
    rho = gnan_port.to_torch(z["rho"], dt, True)
    x = torch.tensor(z["x"]).to(dt)
    if z["variant"] == "batched":
        hop = torch.tensor(z["dist_batch"]).long()           # -1 marks cross-graph / unreachable
        cnt = None
    else:
        hop = gnan_lut.hops_from_reference(torch.tensor(z["node_distances"]))
        cnt = gnan_lut.counts_from_hops(hop)
        # the reference's normalisation matrix is exactly the gathered level sizes
        idx = torch.where(hop < 0, torch.full_like(hop, cnt.shape[1] - 1), hop)
        assert torch.equal(torch.gather(cnt, 1, idx).float(), torch.tensor(z["normalization_matrix"]))
    if z["variant"] == "gnan_loop" and "node_ids" in z:
        ids = torch.tensor(z["node_ids"])
        hop, cnt = hop[ids], cnt[ids]
    out = gnan_lut.forward_rows(fs, rho, x, hop, cnt, lut_mode(z))
    if z["variant"] == "batched":
        if z["is_graph_task"]:
            B = int(z["batch_vector"].max()) + 1
            out = torch.zeros(B, out.shape[1], dtype=dt).index_add(0, torch.tensor(z["batch_vector"]), out)
    elif z["is_graph_task"]:
        out = out.sum(0).view(-1, 1)
    (out * torch.tensor(z["out_weight"]).to(dt)).sum().backward()
    return out, fs, rho


@pytest.mark.parametrize("name", [n for n in G.MODEL_CASES if "readout" not in n] + G.BATCHED_CASES)
def test_lut_restatement_matches_reference(name):
    z = G.load(name)
    out, fs, rho = run_lut(z)
    assert out.shape == z["out"].shape
    assert G.rel_err(out.detach().numpy(), z["out"]) < TOL_LUT
    check_grads(z, fs, "grad_fs", 2e-5)
    check_grads(z, rho, "grad_rho", 2e-5)


@pytest.mark.parametrize("name", G.PREPROCESS_CASES)
def test_apsp_oracle_bit_exact(name):
    z = G.load(name)
    n = int(z["meta"][0])
    node_task = bool(z["meta"][2])
    ei = z["edge_index"].reshape(2, -1)
    hop = oapsp.apsp(ei, n)
    cnt = oapsp.level_counts(hop)
    nd, nm = oapsp.reference_format(hop, cnt)
    assert np.array_equal(nd, z["node_distances"])            # bit-exact float32
    assert np.array_equal(nm, z["normalization_matrix"])
    assert np.array_equal(z["x_out"][:, :-1], z["x"]) and np.all(z["x_out"][:, -1] == 1.0)   # :108 / :127
    assert cnt.sum(axis=1).tolist() == [n] * n
    del node_task


@pytest.mark.skipif(not os.path.exists("/root/reference/GNAN.py"), reason="the unmodified reference is only present in the build container")
def test_golden_files_regenerate_bit_exactly_from_the_reference(tmp_path):
    """The committed fixtures ARE the reference's outputs: oracle/make_golden.py (unmodified GNAN.py / models.py / pre_process_datasets.py /
    trainer.py through the torch_geometric stub) rewrites every file of tests/golden/ bit for bit."""
    import contextlib
    import glob
    import io

    import oracle.make_golden as mg
    old = mg.OUT
    mg.OUT = str(tmp_path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mg.main()
            mg.trainer_cases()
    finally:
        mg.OUT = old
    files = sorted(glob.glob(os.path.join(old, "*.npz")))
    assert len(files) >= 23
    for f in files:
        g = os.path.join(str(tmp_path), os.path.basename(f))
        assert os.path.exists(g), f"{os.path.basename(f)} is not produced by make_golden.py"
        a, b = np.load(f, allow_pickle=True), np.load(g, allow_pickle=True)
        assert set(a.files) == set(b.files), os.path.basename(f)
        for k in a.files:
            if a[k].dtype == object:
                assert str(a[k]) == str(b[k]), (os.path.basename(f), k)
            else:
                assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (os.path.basename(f), k)


@pytest.mark.parametrize("seed", range(12))
def test_apsp_oracle_matches_scipy_dijkstra_on_random_graphs(seed):
    """The C oracle against the library call the reference makes (scipy.sparse.csgraph.dijkstra on the unit-weight directed adjacency,
    pre_process_datasets.py:109-110,128-129) and a numpy restatement of its normaliser (:112-121), on random directed / undirected
    graphs with isolated nodes, self loops and several components: hops, level counts and both fp32 matrices bit for bit."""
    import scipy.sparse
    from scipy.sparse.csgraph import dijkstra
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 70))
    m = int(rng.integers(0, 3 * n))
    src, dst = rng.integers(0, n, size=m), rng.integers(0, n, size=m)
    if seed % 2 == 0:                                                     # undirected: both directions
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    ei = np.unique(np.stack([src, dst]), axis=1).astype(np.int64)         # simple (no repeated pairs); self loops stay
    adj = scipy.sparse.coo_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n))
    d = dijkstra(scipy.sparse.lil_matrix(adj))
    want_hop = np.where(np.isinf(d), -1, d).astype(np.int64)
    hop = oapsp.apsp(ei, n)
    assert np.array_equal(np.asarray(hop, dtype=np.int64), want_hop)
    cnt = oapsp.level_counts(hop)
    D = int(want_hop.max())
    assert cnt.shape == (n, D + 2) and cnt.sum(axis=1).tolist() == [n] * n
    for lvl in range(D + 1):
        assert np.array_equal(cnt[:, lvl], (want_hop == lvl).sum(axis=1))
    assert np.array_equal(cnt[:, -1], (want_hop < 0).sum(axis=1))
    nd, nm = oapsp.reference_format(hop, cnt)
    want_nd = (1.0 / (torch.nan_to_num(torch.from_numpy(d).float(), posinf=np.inf) + 1)).numpy()      # :112-115
    assert nd.dtype == np.float32 and np.array_equal(nd, want_nd)
    want_nm = np.stack([(want_nd[i][None, :] == want_nd[i][:, None]).sum(axis=1) for i in range(n)]).astype(np.float32)   # :117-121
    assert np.array_equal(nm, want_nm)
    rows = min(n, 5)
    assert np.array_equal(oapsp.apsp_rows(ei, n, rows), np.asarray(hop)[:rows])


@pytest.mark.skipif(not os.path.exists("/root/reference/GNAN.py"), reason="the unmodified reference is only present in the build container")
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8, 14, 33])     # 14 and 33: ill-conditioned draws (see below)
def test_oracle_matches_the_live_reference_on_fresh_random_cases(tmp_path, monkeypatch, seed):
    """Beyond the committed fixtures: NEW random configurations (sizes, widths, depths, switches drawn from the seed) are run through the
    unmodified reference classes here, and both restatements (fp32 port in the reference's op order, float64 table form) must reproduce
    outputs and all parameter gradients — the oracle is pinned to the reference itself, not only to 23 files."""
    import contextlib
    import io

    import oracle.make_golden as mg
    gnan_py, models_py, pre_py, batched_cls = mg.import_reference()
    rng = np.random.default_rng(9000 + seed)
    monkeypatch.setattr(mg, "OUT", str(tmp_path))
    monkeypatch.setattr(G, "GOLDEN_DIR", str(tmp_path))
    pick = lambda *v: v[int(rng.integers(0, len(v)))]
    names = []
    with contextlib.redirect_stdout(io.StringIO()):
        for variant, cls, kw in (("gnanpy_tensor", gnan_py.TensorGNAN, {}), ("models_tensor", models_py.TensorGNAN, {}),
                                 ("gnan_loop", pick(gnan_py.GNAN, models_py.GNAN), {})):
            graph = bool(rng.integers(0, 2)) and variant != "gnan_loop"
            name = f"{variant if variant != 'gnan_loop' else 'gnan_loop'}_live{seed}"
            if variant == "gnan_loop":
                kw = dict(layers_kw="n_layers" if cls is gnan_py.GNAN else "num_layers", rho_per_feature=bool(rng.integers(0, 2)),
                          node_ids=pick(None, [2, 0, 5]))
            elif variant == "models_tensor":
                kw = dict(rho_per_feature=bool(rng.integers(0, 2)) and not graph)
            mg.model_case(name, cls, variant, rng, pre_py, N=int(rng.integers(8, 30)), K_raw=int(rng.integers(2, 7)),
                          C=1 if graph else int(rng.integers(1, 6)), H=pick(8, 16, 32, 64), L=pick(1, 2, 3, 4) if variant == "gnanpy_tensor" else pick(2, 3),
                          is_graph_task=graph, normalize_rho=bool(rng.integers(0, 2)), n_isolated=int(rng.integers(0, 3)),
                          directed=bool(rng.integers(0, 2)), bias=bool(rng.integers(0, 4)), **kw)
            names.append(name)
        mg.batched_case(f"batched_live{seed}", batched_cls, rng, B=int(rng.integers(2, 6)), K=int(rng.integers(2, 6)), C=int(rng.integers(1, 5)),
                        H=pick(8, 16), is_graph_task=bool(rng.integers(0, 2)))
        names.append(f"batched_live{seed}")
    for name in names:
        test_port_matches_reference(name)                    # fp32 port in the reference's op order == the fp32 reference
        # the table form against the SAME op order carried in float64: a random case may be ill-conditioned (outputs that are small
        # differences of O(1) terms), where the reference's own fp32 rounding exceeds any fixed bound against float64; the algebra
        # is checked where rounding does not enter (left: the fp32 rounding of the 1/(1+d) inputs the port is fed, 6e-8, which such a
        # case amplifies: hence 1e-6 on outputs and 2e-5 on gradients; 40 seeds were tried when the test was written)
        z = G.load(name)
        out_l, fs_l, rho_l = run_lut(z)
        out_p, fs_p, rho_p, _ = run_port(z, torch.float64)
        assert G.rel_err(out_l.detach().numpy(), out_p.detach().numpy()) < 1e-6, name
        for got, want in ((fs_l, fs_p), (rho_l, rho_p)):
            for k in ("w1", "b1", "wh", "bh", "wo", "bo"):
                if want[k] is None or want[k].grad is None or got[k].grad is None or float(want[k].grad.norm()) == 0:
                    continue
                assert G.rel_err(got[k].grad.numpy(), want[k].grad.numpy()) < 2e-5, (name, k)
