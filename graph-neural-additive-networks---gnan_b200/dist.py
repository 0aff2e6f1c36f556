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
class _AllGatherRows(torch.autograd.Function):
    """S_full = concat_r S_r along dim 0 (blocks may have different lengths). Backward: every rank holds a partial
    dL/dS_full (it only aggregated its own rows), so the gradient of block r is the SUM over ranks of that slice:
    a reduce-scatter (done as all-reduce + slice when block sizes are ragged)."""

    @staticmethod
    def forward(ctx, s_local, sizes, group):
        ctx.sizes, ctx.group = list(sizes), group
        ctx.rank = dist.get_rank(group)
        nmax = max(ctx.sizes)
        tail = tuple(s_local.shape[1:])
        mine = s_local.contiguous()
        if mine.shape[0] < nmax:                                   # ragged last block: pad to the common size
            mine = torch.cat([mine, mine.new_zeros((nmax - mine.shape[0],) + tail)])
        parts = [mine.new_empty((nmax,) + tail) for _ in ctx.sizes]
        dist.all_gather(parts, mine, group=group)
        return torch.cat([p[:n] for p, n in zip(parts, ctx.sizes)], dim=0)

    @staticmethod
    def backward(ctx, g_full):
        g_full = g_full.contiguous()
        if len(set(ctx.sizes)) == 1:
            out = g_full.new_empty((ctx.sizes[0],) + tuple(g_full.shape[1:]))
            dist.reduce_scatter_tensor(out, g_full, op=dist.ReduceOp.SUM, group=ctx.group)
            return out, None, None
        dist.all_reduce(g_full, op=dist.ReduceOp.SUM, group=ctx.group)
        b = sum(ctx.sizes[:ctx.rank])
        return g_full[b:b + ctx.sizes[ctx.rank]].clone(), None, None


def all_gather_rows(s_local: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    return _AllGatherRows.apply(s_local, tuple(int(n) for n in sizes), group)


def allreduce_gradients(params, group=None, average: bool = False):
    """One fused all-reduce of all parameter gradients (flat buffer), written back in place."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


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
    s_local = model._feature_sums(x_local, cx)                                  # [|V_r|, C]
    s_full = all_gather_rows(s_local, sizes, group)                             # [N, C]
    flavor_input_norm = model.__class__.__module__.endswith(".GNAN") and model.__class__.__name__ == "TensorGNAN"
    if model.normalize_rho and flavor_input_norm:                               # GNAN.py:65-67
        u = ops.rho_table_inputs(hop_data.nbins, dev, cnt=hop_data.level_counts)
        T = model._row_tables(hop_data, u, hop_data.level_counts).view(hop_data.rows, hop_data.nbins, -1)
        return ops.aggregate_rows(hop_data.hop, T, s_full, per_row=True)
    T = model._table(ops.rho_table_inputs(hop_data.nbins, dev))
    rs = ops.level_rscale(hop_data.level_counts) if model.normalize_rho else None
    return ops.aggregate_rows(hop_data.hop, T, s_full, rscale=rs)
