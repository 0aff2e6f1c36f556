"""Interpretability export: the curves the reference's notebook plots (mutagenicity_visualizations.ipynb cells 4-9,
README.md:29-31), evaluated in bulk on the GPU from the stacked parameters instead of one `model.fs[i].forward(t)` /
`model.m.forward(val)` call per point.

    f   = shape_function_table(model, grid)            # [len(grid), K, C]   f_k(t)            (cell 6 uses t = 1.0)
    rho = distance_function_table(model, max_distance) # [D+1, C_rho]        rho(1/(1+d))      (cell 4)
    z   = heatmap(model, max_distance)                 # [K, D+1]            f_k(1) * rho(d)   (cell 9: np.outer)
"""
import torch

from . import ops


def shape_function_table(model, grid):
    """f_k(t) for every feature k and every t in `grid` (1-D tensor / list) -> [len(grid), K, C]."""
    st = model.fs
    dev = st.wo.device
    t = torch.as_tensor(grid, dtype=torch.float32, device=dev).reshape(-1)
    u = t.view(-1, 1).expand(-1, st.groups).contiguous()
    with torch.no_grad():
        return ops.mlp_per_group(u, *st.kernel_args(), precision="fp32")


def distance_function_table(model, max_distance, raw=None):
    """rho at the hop distances d = 0..max_distance -> [max_distance+1, C_rho]. The distance models feed rho 1/(1+d)
    (GNAN.py:65-67, notebook cell 4); the batched variant feeds the raw hop count d (batched_pyg_main.py:154), which is
    selected automatically for gnan_b200.batched.TensorGNAN or with raw=True."""
    st = model.rho
    dev = st.wo.device
    if raw is None:
        raw = model.__class__.__module__.endswith(".batched")
    u = ops.rho_table_inputs(int(max_distance) + 2, dev, raw=bool(raw))[: int(max_distance) + 1]      # drop the unreachable bin
    with torch.no_grad():
        return ops.mlp(u.reshape(-1, 1).contiguous(), *st.kernel_args(), precision="fp32")


def heatmap(model, max_distance, x_value=1.0, channel=0):
    """z[k,d] = f_k(x_value)[channel] * rho(d)[channel or 0]: the feature x distance contribution map of notebook cell 9."""
    f = shape_function_table(model, [x_value])[0, :, channel]                    # [K]
    r = distance_function_table(model, max_distance)
    r = r[:, channel if r.shape[1] > 1 else 0]                                   # [D+1]
    return torch.outer(f, r)
