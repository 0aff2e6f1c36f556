"""`gnan_b200.optim.Adam`: torch.optim.Adam's update (main.py:141: Adam(model.parameters(), lr, weight_decay)) as ONE kernel launch
over every parameter tensor (gnan_adam_step, csrc/train.cu). Same constructor arguments and semantics (L2 weight decay added to the
gradient, bias correction, no amsgrad); the step counter lives on the device, so the step can be captured in a CUDA graph
(trainer.CapturedStep) and keeps counting on every replay. A learning-rate scheduler may change `param_groups[i]["lr"]` between
steps as usual (a captured graph bakes the value in: re-capture after a change, as trainer.train_epoch does).

State per parameter: `exp_avg`, `exp_avg_sq` (as torch); per group: `state_buf` = device float [3] (step, 1-beta1^step,
sqrt(1-beta2^step))."""
import ctypes

import torch

from ._lib import check, load, stream_handle


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi, group):
        """ctypes pointer tables of the tensors that take part in the step (cached while the same storages are in use)."""
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                raise TypeError("gnan_b200.optim.Adam needs contiguous fp32 CUDA parameters and gradients (no CPU fallback)")
            st = self.state[p]
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if "state_buf" not in group:
            group["state_buf"] = torch.zeros(3, dtype=torch.float32, device=ps[0].device) if ps else None
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        tab = self._tables.get(gi)
        if tab is None or tab[0] != key:
            n = len(ps)
            arr = lambda xs: (ctypes.c_void_p * n)(*xs)
            tab = (key, n, arr([p.data_ptr() for p in ps]), arr([p.grad.data_ptr() for p in ps]),
                   arr([self.state[p]["exp_avg"].data_ptr() for p in ps]), arr([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                   (ctypes.c_int64 * n)(*[p.numel() for p in ps]))
            self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = load()
        for gi, group in enumerate(self.param_groups):
            _, n, P, G, M, V, N = self._table(gi, group)
            if n == 0:
                continue
            b1, b2 = group["betas"]
            check(lib.gnan_adam_step(n, P, G, M, V, N, group["state_buf"].data_ptr(), float(group["lr"]), float(b1), float(b2),
                                     float(group["eps"]), float(group["weight_decay"]), stream_handle()), "gnan_adam_step")
        return loss
