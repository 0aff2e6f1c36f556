"""Stacked storage for G scalar-input MLPs behind the reference's `ModuleList` / `Sequential` surface.

The reference builds K `nn.Sequential`s of tiny `nn.Linear`s (GNAN.py:24-34: 3K+3 modules, 4 305 for Cora). Here each
layer's weights of all groups are ONE parameter ([G,H], [n_hidden,G,H,H], [G,C,H]) so the kernels see contiguous
arrays and the optimizer sees 6 tensors; `state_dict()` / `load_state_dict()` are translated to the reference's key
names (`fs.{k}.{0,3,6}.weight`, `rho.{0,2,4}.weight`), and `fs[k]` / `rho` stay indexable, callable sub-modules
(README.md:30; mutagenicity_visualizations.ipynb cells 4,6)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class StackedMLP(nn.Module):
    """G scalar-input MLPs with stacked weights. `seq_stride` is the distance between Linear indices inside the
    reference's Sequential (3 with Dropout, 2 without)."""

    def __init__(self, groups, out_channels, n_layers, hidden_channels, bias, seq_stride, dropout=0.0, single=False):
        super().__init__()
        self._single = bool(single)   # True: one MLP addressed as `rho.{i}.weight`; False: a list `fs.{k}.{i}.weight`
        G, C, L = int(groups), int(out_channels), int(n_layers)
        H = int(hidden_channels) if L >= 2 else 1
        self.groups, self.out_channels, self.n_layers, self.hidden, self.has_bias = G, C, L, H, bool(bias)
        self.seq_stride, self.dropout = seq_stride, float(dropout)
        nh = max(L - 2, 0)
        mk = lambda *s: nn.Parameter(torch.empty(*s))
        if L >= 2:
            self.w1, self.wh, self.wo = mk(G, H), mk(nh, G, H, H), mk(G, C, H)
        else:
            self.w1, self.wh, self.wo = mk(0), mk(0), mk(G, C, 1)
        if bias:
            self.b1, self.bh, self.bo = (mk(G, H), mk(nh, G, H), mk(G, C)) if L >= 2 else (mk(0), mk(0), mk(G, C))
        else:  # kernels take zero biases; they are buffers, not parameters (the reference has no such tensors)
            zb = lambda *s: torch.zeros(*s)
            for n, t in zip(("b1", "bh", "bo"), (zb(G, H), zb(nh, G, H), zb(G, C)) if L >= 2 else (zb(0), zb(0), zb(G, C))):
                self.register_buffer(n, t, persistent=False)
        self.reset_parameters()
        self._register_state_dict_hook(_to_reference_keys)
        self.register_load_state_dict_pre_hook(_from_reference_keys)

    # -- init: nn.Linear's default (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias)
    def reset_parameters(self):
        with torch.no_grad():
            for name, fan_in in (("w1", 1), ("b1", 1), ("wh", self.hidden), ("bh", self.hidden),
                                 ("wo", self.hidden), ("bo", self.hidden)):
                t = getattr(self, name)
                if isinstance(t, nn.Parameter) and t.numel():
                    bound = 1.0 / math.sqrt(fan_in)
                    t.uniform_(-bound, bound)

    def xavier_normal_(self, gain):
        """The reference TensorGNAN init (GNAN.py:49-53): xavier_normal_(gain) on weights, zeros on biases."""
        with torch.no_grad():
            H, C = self.hidden, self.out_channels
            if self.n_layers >= 2:
                self.w1.normal_(0.0, gain * math.sqrt(2.0 / (1 + H)))
                if self.wh.numel():
                    self.wh.normal_(0.0, gain * math.sqrt(2.0 / (H + H)))
                self.wo.normal_(0.0, gain * math.sqrt(2.0 / (H + C)))
            else:
                self.wo.normal_(0.0, gain * math.sqrt(2.0 / (1 + C)))
            for n in ("b1", "bh", "bo"):
                t = getattr(self, n)
                if isinstance(t, nn.Parameter):
                    t.zero_()

    def linear_indices(self):
        return [self.seq_stride * l for l in range(self.n_layers)]

    def layer_tensors(self, g):
        """[(weight [out,in], bias [out] or None)] of group g as views of the stacked storage."""
        bias = self.has_bias
        if self.n_layers == 1:
            return [(self.wo[g], self.bo[g] if bias else None)]
        out = [(self.w1[g].unsqueeze(1), self.b1[g] if bias else None)]
        for l in range(self.n_layers - 2):
            out.append((self.wh[l, g], self.bh[l, g] if bias else None))
        out.append((self.wo[g], self.bo[g] if bias else None))
        return out

    def kernel_args(self):
        return self.w1, self.b1, self.wh, self.bh, self.wo, self.bo, self.n_layers

    # -- the reference's ModuleList / Sequential surface
    def __len__(self):
        return self.groups

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [MLPView(self, i) for i in range(*k.indices(self.groups))]
        if k < 0:
            k += self.groups
        if not 0 <= k < self.groups:
            raise IndexError(k)
        return MLPView(self, k)

    def __iter__(self):
        return (MLPView(self, k) for k in range(self.groups))

    def forward(self, t):
        """`model.rho(t)` (single MLP). For a list, call `model.fs[k](t)`."""
        if not self._single:
            raise TypeError("a list of shape functions is not callable; use fs[k](t)")
        return MLPView(self, 0)(t)

    def reference_named_parameters(self):
        """(name, tensor) pairs under the reference's Sequential key names ("0.weight", "2.bias", ...) for a single MLP:
        views of the stacked storage, for printing / inspection only. `named_parameters()` / `parameters()` keep
        nn.Module semantics (the stacked leaf Parameters), so optimizers, zero_grad and clip_grad_norm_ work on
        `model.rho.parameters()`."""
        return MLPView(self, 0).named_parameters()


class MLPView(nn.Module):
    """`model.fs[k]` / `model.rho`: a callable with the reference Sequential's behaviour, reading the stacked storage.
    Plain torch ops on any device: this is the interpretability accessor (notebook cells 4-6), not the training path."""

    def __init__(self, stacked, g):
        super().__init__()
        object.__setattr__(self, "_stacked", stacked)   # not registered as a sub-module: no parameter duplication
        self._g = g

    def forward(self, t):
        st = self._stacked
        layers = st.layer_tensors(self._g)
        h = t
        for i, (w, b) in enumerate(layers):
            h = F.linear(h, w, b)
            if i < len(layers) - 1:
                h = torch.relu(h)
                if st.dropout > 0 and st.seq_stride == 3:
                    h = F.dropout(h, st.dropout, st.training)
        return h

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        st = self._stacked
        for idx, (w, b) in zip(st.linear_indices(), st.layer_tensors(self._g)):
            yield f"{prefix}{idx}.weight", w
            if b is not None:
                yield f"{prefix}{idx}.bias", b

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p


# ---- state_dict translation ------------------------------------------------------------------------------------------
_STACK_NAMES = ("w1", "b1", "wh", "bh", "wo", "bo")


def _group_prefixes(module, prefix):
    """reference key prefix of each group: 'fs.{k}.' for a list, 'rho.' for a single MLP"""
    base = prefix[:-1] if prefix.endswith(".") else prefix          # e.g. "fs" or "rho" or "readout_nam.fs"
    if getattr(module, "_single", False):
        return [base + "."]
    return [f"{base}.{g}." for g in range(module.groups)]


def _to_reference_keys(module, state_dict, prefix, local_metadata):
    for n in _STACK_NAMES:
        state_dict.pop(prefix + n, None)
    idx = module.linear_indices()
    for g, gp in enumerate(_group_prefixes(module, prefix)):
        for i, (w, b) in zip(idx, module.layer_tensors(g)):
            state_dict[f"{gp}{i}.weight"] = w.detach()
            if b is not None:
                state_dict[f"{gp}{i}.bias"] = b.detach()
    return state_dict


def _from_reference_keys(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    gps = _group_prefixes(module, prefix)
    if not any(k.startswith(gps[0]) for k in state_dict):
        return                                                       # already in stacked form (or absent)
    idx = module.linear_indices()
    new = {n: getattr(module, n).detach().clone() for n in _STACK_NAMES if isinstance(getattr(module, n), nn.Parameter)}
    try:
        for g, gp in enumerate(gps):
            for li, i in enumerate(idx):
                w = state_dict.pop(f"{gp}{i}.weight")
                b = state_dict.pop(f"{gp}{i}.bias", None)
                if module.n_layers == 1 or li == len(idx) - 1:
                    new["wo"][g] = w
                    if b is not None and "bo" in new:
                        new["bo"][g] = b
                elif li == 0:
                    new["w1"][g] = w[:, 0]
                    if b is not None and "b1" in new:
                        new["b1"][g] = b
                else:
                    new["wh"][li - 1, g] = w
                    if b is not None and "bh" in new:
                        new["bh"][li - 1, g] = b
    except KeyError as e:
        missing_keys.append(str(e.args[0]))
        return
    for n, t in new.items():
        state_dict[prefix + n] = t
