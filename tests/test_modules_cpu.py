"""CPU: host-side logic of the drop-in modules — reference state_dict keys, fs[k]/rho accessors, ctor signatures,
and that nothing silently computes on the CPU."""
import inspect

import numpy as np
import pytest
import torch

from oracle import gnan_port
from tests import _golden as G
from tests._build import build_module


@pytest.mark.parametrize("name", G.MODEL_CASES + G.BATCHED_CASES)
def test_loads_reference_state_dict(name):
    z = G.load(name)
    m = build_module(z)                      # strict load of the reference's own keys
    sd = m.state_dict()
    ref = {k: v for k, v in z["sd"].items() if not k.startswith("rhos.")}
    assert sorted(sd) == sorted(ref)
    for k, v in ref.items():
        assert np.array_equal(sd[k].numpy(), v), k
    n_ref = sum(v.size for v in ref.values())
    assert sum(p.numel() for p in m.parameters()) == n_ref


@pytest.mark.parametrize("name", ["gnanpy_tensor_node", "gnanpy_tensor_node_l1", "gnanpy_tensor_node_l4", "batched_graph"])
def test_shape_function_accessors_match_oracle(name):
    z = G.load(name)
    m = build_module(z).eval()
    t = torch.linspace(-2, 2, 9).view(-1, 1)
    fs = gnan_port.to_torch(z["fs"]); rho = gnan_port.to_torch(z["rho"])
    for k in (0, z["K"] - 1):
        assert torch.allclose(m.fs[k](t), gnan_port.scalar_mlp(fs, k, t), atol=1e-6)
        assert torch.allclose(m.fs[k].forward(t), gnan_port.scalar_mlp(fs, k, t), atol=1e-6)
    assert torch.allclose(m.rho(t), gnan_port.scalar_mlp(rho, 0, t), atol=1e-6)
    assert len(m.fs) == z["K"]
    names = [n for n, _ in m.rho.named_parameters()]
    assert names[0] == "0.weight"


def test_constructor_signatures_match_reference():
    from gnan_b200 import GNAN as g, batched as b, models as mo
    want = ["in_channels", "out_channels", "n_layers", "hidden_channels", "bias", "dropout", "device", "rho_per_feature",
            "normalize_rho", "is_graph_task", "readout_n_layers"]
    assert list(inspect.signature(g.TensorGNAN.__init__).parameters)[1:] == want      # GNAN.py:10-11
    assert list(inspect.signature(mo.TensorGNAN.__init__).parameters)[1:] == want     # models.py:304-305
    assert list(inspect.signature(mo.GNAN.__init__).parameters)[1:4] == ["in_channels", "out_channels", "num_layers"]
    assert list(inspect.signature(g.GNAN.__init__).parameters)[1:10] == [
        "in_channels", "out_channels", "n_layers", "hidden_channels", "bias", "dropout", "device", "normalize_rho",
        "rho_per_feature"]                                                               # GNAN.py:83-84
    assert list(inspect.signature(b.TensorGNAN.__init__).parameters)[1:] == [
        "in_channels", "out_channels", "n_layers", "hidden_channels", "device", "bias", "dropout", "is_graph_task"]
    assert list(inspect.signature(g.GNAN.forward).parameters)[1:] == ["inputs", "node_ids"]


def test_reference_init_statistics():
    from gnan_b200.GNAN import GNAN, TensorGNAN
    torch.manual_seed(0)
    m = TensorGNAN(40, 7, 3, 64)
    assert float(m.fs.b1.abs().max()) == 0.0                       # GNAN.py:52-53
    assert abs(float(m.fs.wh.std()) - 0.01 * (2 / 128) ** 0.5) < 2e-5   # xavier_normal_(gain=0.01), GNAN.py:51
    g = GNAN(40, 7, 3, 64)
    assert abs(float(g.fs.wh.abs().max()) - 1 / 8) < 1e-3          # nn.Linear default: U(-1/sqrt(64), 1/sqrt(64))
    assert not hasattr(m.rho, "nonexistent")
    assert TensorGNAN(4, 1, 3, 8, is_graph_task=True).rho.has_bias is False     # GNAN.py:36-37


def test_no_cpu_fallback():
    """Forward on CPU tensors must fail loudly, never compute."""
    from types import SimpleNamespace
    from gnan_b200._lib import GnanError
    z = G.load("gnanpy_tensor_node")
    m = build_module(z)
    data = SimpleNamespace(x=torch.tensor(z["x"]), node_distances=torch.tensor(z["node_distances"]),
                           normalization_matrix=torch.tensor(z["normalization_matrix"]))
    with pytest.raises((GnanError, RuntimeError)):
        m.forward(data)
