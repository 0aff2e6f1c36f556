"""Multi-GPU execution of the GNAN path on one 8xB200 box: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for the plumbing. The reference has no distributed code at all; the two partitionings below are
the ones the path offers naturally:

  * graph batches are data-parallel: each rank runs forward/backward on its own graphs, then ONE all-reduce of the
    flattened gradients (a few hundred KB: latency-bound, so a single fused call);
  * a large node-level graph is row-sharded on the hop matrix: rank r owns node block V_r, computes S[V_r] with the shape
    MLPs, all-gathers S (N x C floats), aggregates its own hop rows, and in backward reduce-scatters dS before the local
    MLP backward; MLP / rho gradients are all-reduced.

Host-side logic only: the collectives are torch.distributed calls on the caller's process group (backend "nccl" on GPUs,
"gloo" in the CPU tests of tests/test_dist_cpu.py).
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------------------
# partitioning
# ---------------------------------------------------------------------------------------------------------------------
def balanced_ranges(costs: Sequence[float], parts: int) -> List[Tuple[int, int]]:
    """Split items 0..n-1 (kept in order) into `parts` contiguous ranges of roughly equal total cost (greedy on the prefix
    sum). Used to balance graphs by n_g^2 + K*n_g and node rows by their count."""
    n = len(costs)
    total = float(sum(costs))
    out, start, acc = [], 0, 0.0
    for p in range(parts):
        target = total * (p + 1) / parts
        end = start
        while end < n and (acc + costs[end] <= target or end == start) and (n - end) > (parts - p - 1):
            acc += costs[end]
            end += 1
        if p == parts - 1:
            end = n
        out.append((start, end))
        start = end
    return out


def row_block(num_nodes: int, rank: int, world: int, align: int = 16) -> Tuple[int, int]:
    """Contiguous node block [begin, end) of `rank`; block starts are multiples of `align` rows."""
    per = (num_nodes + world - 1) // world
    per = (per + align - 1) // align * align
    b = min(rank * per, num_nodes)
    return b, min(b + per, num_nodes)


# ---------------------------------------------------------------------------------------------------------------------
# collectives with autograd
# ---------------------------------------------------------------------------------------------------------------------
def _timed(name):
    from . import ops
    return ops._timed(name)


class _AllGatherRows(torch.autograd.Function):
    """S_full = concat_r S_r along dim 0. Blocks are the `row_block` partition: every rank holds `nmax` rows except the
    trailing ones, so the padded gather buffer [world * nmax, ...] holds the concatenation as a PREFIX: one
    all_gather_into_tensor, no list gather, no concatenation. Backward: every rank holds a partial dL/dS_full (it only
    aggregated its own rows), so the gradient of block r is the SUM over ranks of that slice: one reduce_scatter_tensor on the
    zero-padded gradient."""

    @staticmethod
    def forward(ctx, s_local, sizes, group):
        ctx.sizes, ctx.group = list(sizes), group
        ctx.rank = dist.get_rank(group)
        world, nmax, total = len(ctx.sizes), max(ctx.sizes), sum(ctx.sizes)
        ctx.prefix, short = True, False                            # concatenation == prefix of the padded buffer?
        for n in ctx.sizes:
            if short and n > 0:
                ctx.prefix = False
            short = short or n < nmax
        tail = tuple(s_local.shape[1:])
        mine = s_local.contiguous()
        if mine.shape[0] < nmax:                                   # ragged last block: pad to the common size
            pad = mine.new_zeros((nmax,) + tail)
            pad[:mine.shape[0]] = mine
            mine = pad
        buf = mine.new_empty((world * nmax,) + tail)
        with _timed("allgather_rows"):
            dist.all_gather_into_tensor(buf, mine, group=group)
        if ctx.prefix:
            return buf[:total]
        return torch.cat([buf[r * nmax:r * nmax + n] for r, n in enumerate(ctx.sizes)], dim=0)

    @staticmethod
    def backward(ctx, g_full):
        world, nmax, total = len(ctx.sizes), max(ctx.sizes), sum(ctx.sizes)
        tail = tuple(g_full.shape[1:])
        if ctx.prefix and total == world * nmax:
            padded = g_full.contiguous()
        else:
            padded = g_full.new_zeros((world * nmax,) + tail)
            if ctx.prefix:
                padded[:total] = g_full
            else:
                off = 0
                for r, n in enumerate(ctx.sizes):
                    padded[r * nmax:r * nmax + n] = g_full[off:off + n]
                    off += n
        out = g_full.new_empty((nmax,) + tail)
        with _timed("reduce_scatter_rows"):
            dist.reduce_scatter_tensor(out, padded, op=dist.ReduceOp.SUM, group=ctx.group)
        return out[:ctx.sizes[ctx.rank]], None, None


def all_gather_rows(s_local: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    return _AllGatherRows.apply(s_local, tuple(int(n) for n in sizes), group)


def allreduce_gradients(params, group=None, average: bool = False):
    """One fused all-reduce of all parameter gradients (flat buffer), written back in place. Prefer FlatGradients, which
    needs neither the concatenation nor the copy back."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    with _timed("allreduce_gradients"):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class FlatGradients:
    """All parameter gradients as views of ONE persistent flat buffer: backward accumulates straight into it, the
    data-parallel reduction is a single all_reduce on that buffer and the optimizer reads the views. No per-step
    concatenation or copy back.

        fg = FlatGradients(model.parameters())
        fg.zero(); loss.backward(); fg.all_reduce(average=True); optimizer.step()

    Use `fg.zero()` instead of `optimizer.zero_grad()` (set_to_none would drop the views)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=self.params[0].dtype if self.params else torch.float32, device=dev)
        self.attach()

    def attach(self):
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = self.flat[off:off + k].view_as(p)
            off += k

    def zero(self):
        if any(p.grad is None or (p.numel() and p.grad.data_ptr() != self.flat.data_ptr() + self.flat.element_size() * o)
               for p, o in zip(self.params, self._offsets())):      # (an empty parameter's view has no address to compare)
            self.attach()
        self.flat.zero_()

    def _offsets(self):
        off = 0
        for p in self.params:
            yield off
            off += p.numel()

    def all_reduce(self, group=None, average: bool = False):
        if not self.flat.numel():
            return
        with _timed("allreduce_gradients"):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=group)


# ---------------------------------------------------------------------------------------------------------------------
# row-sharded node-level forward
# ---------------------------------------------------------------------------------------------------------------------
def row_sharded_forward(model, x_local: torch.Tensor, hop_data, sizes: Sequence[int], group=None, x_compressed=None) -> torch.Tensor:
    """out[V_r] for this rank's node block.

    model     a gnan_b200 GNAN / TensorGNAN module (replicated parameters)
    x_local   [|V_r|, K] features of the owned nodes (may be None when x_compressed is given)
    x_compressed  optional sparse.CompressedFeatures of x_local (used when dropout is off, like in the single-GPU modules)
    hop_data  preprocess.HopData holding the owned hop ROWS (all N columns) and their level counts
    sizes     block sizes of all ranks (sum = N)

    Gradients: S is all-gathered with an autograd-aware collective, so `loss.backward()` reduce-scatters dS and runs the
    MLP backward on the local rows only; call allreduce_gradients(model.parameters()) afterwards.
    """
    from . import ops
    dev = hop_data.hop.device
    cx = x_compressed if (x_compressed is not None and model._dedup_ok()) else None     # sparse.compress_features(x_local), built once
    model._seed_salt = int(getattr(hop_data, "row_begin", 0))                   # dropout masks differ between row shards
    try:
        s_local = model._feature_sums(x_local, cx)                              # [|V_r|, C]
    finally:
        model._seed_salt = 0
    s_full = all_gather_rows(s_local, sizes, group)                             # [N, C]
    flavor_input_norm = model.__class__.__module__.endswith(".GNAN") and model.__class__.__name__ == "TensorGNAN"
    if model.normalize_rho and flavor_input_norm:                               # GNAN.py:65-67
        u = ops.rho_table_inputs(hop_data.nbins, dev, cnt=hop_data.level_counts)
        T = model._row_tables(hop_data, u, hop_data.level_counts).view(hop_data.rows, hop_data.nbins, -1)
        return ops.aggregate_rows(hop_data.hop, T, s_full, per_row=True)
    T = model._table(ops.rho_table_inputs(hop_data.nbins, dev))
    rs = ops.level_rscale(hop_data.level_counts) if model.normalize_rho else None
    return ops.aggregate_rows(hop_data.hop, T, s_full, rscale=rs)
