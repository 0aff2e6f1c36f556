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
