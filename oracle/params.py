"""TEST INFRASTRUCTURE ONLY (oracle/). Helpers that move weights between the reference's
`state_dict` naming and the stacked arrays the oracle works on.

Reference naming (GNAN.py:24-47): `fs.{k}` is `Sequential(Linear, ReLU, Dropout, [Linear, ReLU, Dropout]*, Linear)`
=> Linear indices 0,3,6,...; `rho` is `Sequential(Linear, ReLU, [Linear, ReLU]*, Linear)` => 0,2,4,...
The batched variant (batched_pyg_main.py:116-131) is always 2 layers and rho has a Dropout => 0,3 for both.
"""
import numpy as np


def linear_indices(n_layers, stride):
    return [stride * l for l in range(n_layers)]


def stack_mlps(sd, prefixes, n_layers, stride, bias=True):
    """Stack G scalar-input MLPs given by state_dict prefixes into one dict of arrays.

    returns dict(w1[G,H], b1[G,H], wh[NH,G,H,H], bh[NH,G,H], wo[G,C,Hin], bo[G,C]) as float64 numpy;
    for n_layers == 1: w1/b1 are None, wo is [G,C,1].
    """
    idx = linear_indices(n_layers, stride)
    get = lambda p, i, what: np.asarray(sd[f"{p}.{i}.{what}"], dtype=np.float64)
    G = len(prefixes)
    out = {}
    wo = np.stack([get(p, idx[-1], "weight") for p in prefixes])           # [G,C,Hin]
    C = wo.shape[1]
    out["wo"] = wo
    out["bo"] = np.stack([get(p, idx[-1], "bias") for p in prefixes]) if bias else np.zeros((G, C))
    if n_layers == 1:
        out["w1"] = out["b1"] = None
        out["wh"] = np.zeros((0, G, 1, 1)); out["bh"] = np.zeros((0, G, 1))
        return out
    w1 = np.stack([get(p, idx[0], "weight")[:, 0] for p in prefixes])      # [G,H]
    H = w1.shape[1]
    out["w1"] = w1
    out["b1"] = np.stack([get(p, idx[0], "bias") for p in prefixes]) if bias else np.zeros((G, H))
    NH = n_layers - 2
    out["wh"] = np.zeros((NH, G, H, H)); out["bh"] = np.zeros((NH, G, H))
    for l in range(NH):
        out["wh"][l] = np.stack([get(p, idx[1 + l], "weight") for p in prefixes])
        if bias:
            out["bh"][l] = np.stack([get(p, idx[1 + l], "bias") for p in prefixes])
    return out


def unstack_grads(model, prefixes, n_layers, stride, bias=True):
    """Collect .grad of the reference module's Linear layers into the same stacked layout."""
    named = dict(model.named_parameters(remove_duplicate=False))
    sd = {k: (v.grad.detach().cpu().numpy() if v.grad is not None else np.zeros(tuple(v.shape)))
          for k, v in named.items()}
    return stack_mlps(sd, prefixes, n_layers, stride, bias)


def flatten(prefix, d):
    """dict-of-arrays -> flat npz-friendly dict (None skipped)."""
    return {f"{prefix}.{k}": v for k, v in d.items() if v is not None}


def unflatten(prefix, flat):
    keys = ("w1", "b1", "wh", "bh", "wo", "bo")
    return {k: (flat[f"{prefix}.{k}"] if f"{prefix}.{k}" in flat else None) for k in keys}
