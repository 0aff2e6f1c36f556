"""Drop-in for the reference's trainer.py (`train_epoch` trainer.py:23-86, `test_epoch` :88-160, `get_accuracy` :5-21) around
the gnan_b200 modules, plus `CapturedStep`, the whole training step as one CUDA graph.

Same signatures, same label handling (label_index, {-1,1} -> {0,1}, CrossEntropyLoss -> long, train/val/test masks for node
tasks, `[.,1]` outputs flattened for the loss), same return tuples. What is different is where the work happens:

  * loss, accuracy and sample counts are accumulated in device tensors; the host reads them ONCE per epoch instead of one
    `.item()` per step (trainer.py:72,75), so steps queue back to back on the stream;
  * no `set_detect_anomaly(True)` (trainer.py:24: a debugging aid that doubles the backward cost);
  * a loader item may be the reference's Data-like object (`x`, `edge_index`, `node_distances`, `normalization_matrix`, `y`,
    masks), the same object carrying a `hop_data` attribute (preprocess.HopData, uint8 hops already on the GPU: nothing is
    copied per step, cf. `data.to(device)` at trainer.py:46), or a preprocess.PackedBatch of many graphs (one step per batch);
  * the AUC (tolokers path, trainer.py:68-78) is computed on the device from the concatenated scores.
"""
from types import SimpleNamespace

import torch

from .preprocess import PackedBatch


# ---------------------------------------------------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------------------------------------------------
def _correct(outputs, labels):
    """Number of correct predictions as a 0-d device tensor (trainer.py:5-21 without the .item())."""
    if outputs.dim() == 2 and outputs.shape[-1] > 1:
        return (outputs.argmax(dim=-1) == labels).sum()
    return ((torch.sigmoid(outputs).view(-1) > 0.5) == labels).sum()


def get_accuracy(outputs, labels):
    """trainer.py:5-12: count of correct predictions (python number for the binary case, as upstream)."""
    c = _correct(outputs, labels)
    return c if (outputs.dim() == 2 and outputs.shape[-1] > 1) else c.item()


def get_multiclass_accuracy(outputs, labels):
    assert outputs.size(1) >= labels.max().item() + 1
    return _correct(outputs, labels)


def roc_auc(labels, scores):
    """Area under the ROC curve on the device (what sklearn.metrics.roc_auc_score returns for binary labels): the
    Mann-Whitney statistic with average ranks for tied scores."""
    labels = labels.reshape(-1).to(torch.float64)
    scores = scores.reshape(-1).to(torch.float64)
    n_pos = labels.sum()
    n_neg = labels.numel() - n_pos
    if float(n_pos) == 0 or float(n_neg) == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    order = torch.argsort(scores)
    s = scores[order]
    # average rank of each tie group
    new_group = torch.ones_like(s, dtype=torch.bool)
    new_group[1:] = s[1:] != s[:-1]
    gid = torch.cumsum(new_group, 0) - 1
    pos_idx = torch.arange(1, s.numel() + 1, device=s.device, dtype=torch.float64)
    gsum = torch.zeros(int(gid[-1]) + 1, device=s.device, dtype=torch.float64).index_add_(0, gid, pos_idx)
    gcnt = torch.zeros_like(gsum).index_add_(0, gid, torch.ones_like(pos_idx))
    ranks = (gsum / gcnt)[gid]
    r_pos = (ranks * labels[order]).sum()
    return float((r_pos - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg))


# ---------------------------------------------------------------------------------------------------------------------
# one loader item -> (outputs, labels) exactly as trainer.py prepares them
# ---------------------------------------------------------------------------------------------------------------------
def _labels_of(data, loss_fn, label_index):
    y = data.y
    if y.dim() > 1:
        labels = y[:, label_index].reshape(-1).float()                     # trainer.py:33-35
    else:
        labels = y.flatten()
    if labels.is_cuda:                                                      # trainer.py:38-39 without a host sync
        if not labels.is_floating_point() and loss_fn.__class__.__name__ != 'CrossEntropyLoss':
            labels = labels.float()
        if labels.is_floating_point():
            labels = torch.where((labels == -1).any(), (labels + 1) / 2, labels)
    elif labels.numel() and bool((labels == -1).any()):
        labels = (labels + 1) / 2
    if loss_fn.__class__.__name__ == 'CrossEntropyLoss':
        labels = labels.long()
    return labels


def _forward_item(model, data, labels, device, mask_name, is_graph_task):
    labels = labels.to(device, non_blocking=True)
    if isinstance(data, PackedBatch):
        return model(data), labels
    outputs = model.forward(data)
    if isinstance(outputs, tuple):
        outputs = outputs[0]
    if not is_graph_task:
        mask = getattr(data, mask_name).to(device, non_blocking=True)
        labels, outputs = labels[mask], outputs[mask]                       # trainer.py:53-55
    elif outputs.dim() == 2 and outputs.shape[0] > 1 and outputs.shape[-1] == 1:
        outputs = outputs.T                                                 # [C,1] graph output -> [1,C] logits row
    return outputs, labels


def _fused_loss(loss_fn, outputs, labels):
    """The two losses the reference trains with (main.py: CrossEntropyLoss / BCEWithLogitsLoss with default arguments), as one
    kernel each that returns the value and the gradient (gnan_b200.ops, csrc/train.cu); None when `loss_fn` is anything else."""
    if not (outputs.is_cuda and outputs.dtype == torch.float32 and labels.numel() > 0 and labels.is_cuda):
        return None
    from . import ops
    if (type(loss_fn) is torch.nn.CrossEntropyLoss and loss_fn.weight is None and loss_fn.label_smoothing == 0.0
            and loss_fn.reduction in ("mean", "sum") and outputs.dim() == 2 and labels.dtype == torch.int64 and labels.dim() == 1):
        return ops.cross_entropy_rows(outputs.contiguous(), labels, reduction=loss_fn.reduction)
    if (type(loss_fn) is torch.nn.BCEWithLogitsLoss and loss_fn.weight is None and loss_fn.pos_weight is None
            and loss_fn.reduction in ("mean", "sum") and labels.is_floating_point() and outputs.shape == labels.shape):
        return ops.bce_with_logits(outputs, labels.float(), reduction=loss_fn.reduction)
    return None


def _loss(loss_fn, outputs, labels):
    if outputs.dim() == 2 and outputs.shape[-1] == 1:
        outputs, labels = outputs.flatten(), labels.float()                 # trainer.py:61-62
    fused = _fused_loss(loss_fn, outputs, labels)
    return fused if fused is not None else loss_fn(outputs, labels)


class _Running:
    def __init__(self, device):
        self.loss = torch.zeros((), device=device, dtype=torch.float64)
        self.correct = torch.zeros((), device=device, dtype=torch.float64)
        self.n_samples = 0
        self.steps = 0
        self.scores, self.labels = [], []

    def add(self, loss, outputs, labels, classify, compute_auc):
        self.loss += loss.detach()
        self.steps += 1
        self.n_samples += int(labels.shape[0])
        if classify:
            self.correct += _correct(outputs.detach(), labels)
        if compute_auc:
            self.scores.append(torch.sigmoid(outputs.detach()).view(-1))
            self.labels.append(labels.detach().view(-1))

    def result(self, classify, compute_auc):
        auc = roc_auc(torch.cat(self.labels), torch.cat(self.scores)) if compute_auc else -1
        loss = float(self.loss.item()) / max(self.steps, 1)                 # the single host read of the epoch
        if classify:
            return loss, float(self.correct.item()) / max(self.n_samples, 1), auc
        return loss, -1


def train_epoch(model, dloader, loss_fn, optimizer, device, classify=True, label_index=0, compute_auc=False, is_graph_task=True,
                capture_steps=False):
    """trainer.py:23-86. One optimizer step per loader item; returns (mean loss, accuracy, auc | -1) or (mean loss, -1).
    capture_steps=True (extension, graph tasks with one graph per item): replay one captured CUDA graph per graph size
    (SizeBucketedSteps) instead of launching every step's kernels from Python."""
    run = _Running(device)
    buckets = None
    if capture_steps and is_graph_task and not compute_auc:
        key = (id(optimizer), id(loss_fn), bool(classify))
        cache = model.__dict__.setdefault("_gnan_b200_step_cache", {})
        buckets = cache.get(key)
        if buckets is None:
            buckets = cache[key] = SizeBucketedSteps(model, loss_fn, optimizer, device, bool(classify))
        buckets.begin_epoch()
        run.loss, run.correct = buckets.loss_sum, buckets.correct       # captured steps accumulate into the same tensors
    for data in dloader:
        labels = _labels_of(data, loss_fn, label_index)
        if buckets is not None and buckets.step(data, labels):
            run.steps += 1
            run.n_samples += int(labels.shape[0])
            continue
        optimizer.zero_grad(set_to_none=True)
        outputs, labels = _forward_item(model, data, labels, device, "train_mask", is_graph_task)
        loss = _loss(loss_fn, outputs, labels)
        loss.backward()
        optimizer.step()
        run.add(loss, outputs, labels, classify, compute_auc)
    return run.result(classify, compute_auc)


def test_epoch(model, dloader, loss_fn, device, classify=True, label_index=0, compute_auc=False, val_mask=False, is_graph_task=True):
    """trainer.py:88-160 (calls model.eval(), runs under no_grad, val_mask selects the validation rows of a node task)."""
    run = _Running(device)
    with torch.no_grad():
        model.eval()
        for data in dloader:
            labels = _labels_of(data, loss_fn, label_index)
            outputs, labels = _forward_item(model, data, labels, device, "val_mask" if val_mask else "test_mask", is_graph_task)
            run.add(_loss(loss_fn, outputs, labels), outputs, labels, classify, compute_auc)
    return run.result(classify, compute_auc)


test_epoch.__test__ = False     # not a pytest test


# ---------------------------------------------------------------------------------------------------------------------
# the whole step as one CUDA graph
# ---------------------------------------------------------------------------------------------------------------------
class _no_gc_during_capture:
    """Python's cyclic garbage collector may free an old torch.cuda.CUDAGraph (or any object whose destructor calls into CUDA) at an
    arbitrary allocation; inside a stream capture that call is illegal and invalidates the capture ("operation not permitted when
    stream is capturing"). Collect before, keep the collector off while capturing."""

    def __enter__(self):
        import gc
        gc.collect()
        self._was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()


class CapturedStep:
    """forward + loss + backward + optimizer step of a FIXED-SHAPE input captured into one CUDA graph and replayed.

    A full-graph node-task step is ~40 kernel launches of a few microseconds to a millisecond each; replaying them as one
    graph removes the Python / launch gaps between them (Cora shape: 2.16 -> 1.92 ms per step). Requirements: an optimizer
    that can be captured (`torch.optim.Adam(..., capturable=True)`, ideally `fused=True`), inputs already on the device,
    no collectives and no host synchronisation inside `step_fn`.

        step = CapturedStep(lambda: loss_of(model(static_inputs)), optimizer)
        for epoch in ...: loss = step()                 # a device tensor, rewritten by every replay
    Refresh the static input tensors in place (`copy_`) between replays to train on new values of the same shape.
    Shared evaluations (gnan_b200.sparse) inside a captured step are explicit only: the implicit per-object caches are
    bypassed during capture (the dense kernels run), because a cached compressed form / value->row mapping would be baked
    into the graph and go stale after an in-place refresh. To keep them, pass `x_compressed` (refresh it with
    `CompressedFeatures.copy_tensors_`) and mark hop data whose level counts never change with
    `hop_data.static_level_counts = True`.
    """

    def __init__(self, loss_closure, optimizer, warmup=3, after_backward=None, zero_grad=None):
        """after_backward: optional callable run between backward and the optimizer step INSIDE the captured step, e.g.
        `lambda: flat_grads.all_reduce(average=True)` for data-parallel training: NCCL collectives are captured like kernels
        (every rank must capture the same sequence).
        zero_grad: optional callable replacing `optimizer.zero_grad(set_to_none=True)`, e.g. `dist.FlatGradients.zero` (the
        gradients are views of one persistent buffer that must not be dropped)."""
        self._closure, self._opt, self._after = loss_closure, optimizer, after_backward
        self._zero = zero_grad if zero_grad is not None else (lambda: optimizer.zero_grad(set_to_none=True))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                           # warm-up off the capture stream: allocator + optimizer state
            for _ in range(max(int(warmup), 1)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._zero()
        from ._lib import load
        lib = load()
        n0 = lib.gnan_launch_count()
        # thread_local: the NCCL watchdog thread polls CUDA events while this thread captures
        with _no_gc_during_capture(), torch.cuda.graph(self.graph, capture_error_mode="thread_local" if after_backward is not None else "global"):
            self.loss = self._eager()
        self.kernel_launches = int(lib.gnan_launch_count() - n0)     # gnan_b200 kernels per replay (the counter is host-side)

    def _eager(self):
        self._zero()
        loss = self._closure()
        loss.backward()
        if self._after is not None:
            self._after()
        self._opt.step()
        return loss

    def __call__(self):
        self.graph.replay()
        return self.loss


# ---------------------------------------------------------------------------------------------------------------------
# per-graph steps (the reference's batch_size=1 loaders) replayed from one CUDA graph per graph size
# ---------------------------------------------------------------------------------------------------------------------
class SizeBucketedSteps:
    """`train_epoch(..., capture_steps=True)`: one captured step (forward + loss + backward + optimizer step + metric
    accumulation) per distinct graph size n; an item is copied into that size's static buffers and the graph is replayed.

    The reference trains graph tasks one graph per step (datasets.py:339-341), which on a GPU is ~40 launches of
    microsecond kernels per step: Python / launch bound. Graph sizes repeat (Mutagenicity: ~120 distinct n), so the launch
    sequence is captured once per n. The level table is padded to n + 1 columns (levels 0..n-1 + unreachable; empty levels
    have count 0 and are never indexed), which makes the captured shapes a function of n alone. Dropout stays correct
    through the device-side seed word (GNAN._seed_word). Needs an optimizer that can be captured: torch.optim.Adam is
    switched to capturable=True in place (its step counters move to the device); other optimizers -> eager steps.
    Hyper-parameters are baked into a capture: a change of lr / betas / eps / weight_decay (LR schedulers) drops the cache.
    """

    def __init__(self, model, loss_fn, optimizer, device, classify):
        self.model, self.loss_fn, self.opt, self.device, self.classify = model, loss_fn, optimizer, torch.device(device), classify
        self.entries = {}
        self.loss_sum = torch.zeros((), device=self.device, dtype=torch.float64)
        self.correct = torch.zeros((), device=self.device, dtype=torch.float64)
        self.hyper = self._hyper()
        self.usable = self._make_capturable()

    def _hyper(self):
        return [tuple((k, float(v) if isinstance(v, (int, float)) else str(v)) for k, v in sorted(g.items())
                      if k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize")) for g in self.opt.param_groups]

    def _make_capturable(self):
        if not isinstance(self.opt, torch.optim.Adam) or self.device.type != "cuda":
            return False
        for g in self.opt.param_groups:
            if not g.get("capturable", False):
                g["capturable"] = True
                for p in g["params"]:
                    st = self.opt.state.get(p)
                    if st and "step" in st and torch.is_tensor(st["step"]) and not st["step"].is_cuda:
                        st["step"] = st["step"].to(device=p.device, dtype=torch.float32)
        return True

    def begin_epoch(self):
        if self._hyper() != self.hyper:                      # e.g. ReduceLROnPlateau (main.py:147-153) moved the learning rate
            self.entries.clear()
            self.hyper = self._hyper()
        self.loss_sum.zero_()
        self.correct.zero_()

    def _capture(self, n, K, ld, label_like):
        from .preprocess import HopData
        dev = self.device
        e = SimpleNamespace(x=torch.zeros(n, K, device=dev), hop=torch.full((n, ld), 255, dtype=torch.uint8, device=dev),
                            cnt=torch.zeros(n, n + 1, dtype=torch.int32, device=dev),
                            labels=torch.zeros_like(label_like, device=dev))
        e.cnt[:, 0] = 1                                      # a valid placeholder graph for the warm-up steps: n isolated nodes
        e.cnt[:, -1] = n - 1
        e.hop.fill_(255)
        e.hop[torch.arange(n), torch.arange(n)] = 0
        data = SimpleNamespace(x=e.x, hop_data=HopData(e.hop, e.cnt, n))
        model, loss_fn, opt = self.model, self.loss_fn, self.opt

        def body():
            opt.zero_grad(set_to_none=True)
            outputs, labels = _forward_item(model, data, e.labels, dev, "train_mask", True)
            loss = _loss(loss_fn, outputs, labels)
            loss.backward()
            opt.step()
            self.loss_sum.add_(loss.detach())
            if self.classify:
                self.correct.add_(_correct(outputs.detach(), labels))
            return loss

        # warm-up steps would move the weights: run them on a snapshot and restore it (optimizer state included)
        w0 = [p.detach().clone() for p in model.parameters()]
        s0 = {k: ({kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}) for k, v in opt.state.items()}
        keep_l, keep_c = self.loss_sum.clone(), self.correct.clone()
        old_dedup, model.dedup = model.dedup, False          # static buffers change content: no per-object caches
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            e.graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with _no_gc_during_capture(), torch.cuda.graph(e.graph):
                body()
        finally:
            model.dedup = old_dedup
        with torch.no_grad():                                # capture does not execute; undo the one warm-up step
            for p, w in zip(model.parameters(), w0):
                p.copy_(w)
            for k, v in s0.items():
                for kk, vv in v.items():
                    if torch.is_tensor(vv):
                        opt.state[k][kk].copy_(vv)
            for k in list(opt.state.keys()):
                if k not in s0:                              # state created by the warm-up step of a fresh optimizer
                    for kk, vv in opt.state[k].items():
                        if torch.is_tensor(vv):
                            vv.zero_()
            self.loss_sum.copy_(keep_l)
            self.correct.copy_(keep_c)
        return e

    def step(self, data, labels):
        """True if the item was trained through a captured step; False -> the caller runs it eagerly."""
        from ._inputs import resolve
        if not self.usable or isinstance(data, PackedBatch) or getattr(data, "x", None) is None:
            return False
        x, hd = resolve(data, self.device)
        n, K = x.shape
        if hd.rows != n or hd.nbins > n + 1 or n < 1:
            return False
        key = (n, K, tuple(labels.shape), labels.dtype)
        e = self.entries.get(key)
        if e is None:
            e = self.entries[key] = self._capture(n, K, hd.hop.shape[1], labels)
        pad = getattr(data, "_gnan_b200_cnt_pad", None)
        if pad is None or pad.shape != e.cnt.shape or pad.device != self.device:
            pad = torch.zeros_like(e.cnt)
            pad[:, :hd.nbins - 1] = hd.level_counts[:, :-1]
            pad[:, -1] = hd.level_counts[:, -1]
            try:
                data._gnan_b200_cnt_pad = pad
            except Exception:
                pass
        e.x.copy_(x, non_blocking=True)
        e.hop.copy_(hd.hop, non_blocking=True)
        e.cnt.copy_(pad, non_blocking=True)
        e.labels.copy_(labels.to(self.device, non_blocking=True), non_blocking=True)
        e.graph.replay()
        return True
