"""GPU all-pairs hop-distance preprocessing with the reference's `pre_process` call shape.

Reference: pre_process_datasets.py:104-148 (append a constant-1 feature column, all-pairs Dijkstra on unit weights,
node_distances = 1/(1+d) with inf -> 0, normalization_matrix[i,j] = #{j': d(i,j') = d(i,j)}), and the per-node
networkx BFS of batched_pyg_main.py:36-44. Here the distances stay integers: a uint8 hop matrix (255 = unreachable)
plus int32 BFS level sizes; the reference's two fp32 [N,N] matrices are derived views (HopData.reference_format()).

Graphs must be simple (no duplicate edges): the reference's COO->LIL conversion sums duplicates into weight 2
(pre_process_datasets.py:109), which is not emulated.
"""
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import check, load, ptr, stream_handle
from .ops import hop_ld


class HopData:
    """Compact distance data of one graph (or a row shard [row_begin, row_begin+R) of it).

    hop          uint8 [R, ld]  hop count, 255 = unreachable; columns >= N are padding
    level_counts int32 [R, nbins]  cnt[i,d] = #{j : hop[i,j] = d}, last column = #unreachable; nbins = D+2
    """

    def __init__(self, hop, level_counts, num_nodes, row_begin=0):
        self.hop = hop
        self.level_counts = level_counts
        self.num_nodes = int(num_nodes)
        self.row_begin = int(row_begin)

    @property
    def nbins(self):
        return self.level_counts.shape[1]

    @property
    def rows(self):
        return self.hop.shape[0]

    def to(self, device):
        return HopData(self.hop.to(device, non_blocking=True), self.level_counts.to(device, non_blocking=True),
                       self.num_nodes, self.row_begin)

    def reference_format(self):
        """(node_distances, normalization_matrix) as fp32 [R,N], bit-identical to pre_process_datasets.py:112-121."""
        lib = load()
        R, N = self.rows, self.num_nodes
        nd = torch.empty(R, N, dtype=torch.float32, device=self.hop.device)
        nm = torch.empty(R, N, dtype=torch.float32, device=self.hop.device)
        check(lib.gnan_hops_to_reference(ptr(self.hop), R, N, self.hop.shape[1], ptr(self.level_counts), self.nbins,
                                         ptr(nd), ptr(nm), stream_handle()), "gnan_hops_to_reference")
        return nd, nm


def from_reference_format(node_distances, normalization_matrix=None):
    """fp32 [R,N] reference tensors (on the GPU) -> HopData. One device sync to size the level table."""
    lib = load()
    nd = node_distances.contiguous()
    if nd.dtype != torch.float32:
        raise TypeError("node_distances must be float32")
    R, N = nd.shape
    dev = nd.device
    pos = nd[nd > 0]
    dmax = int(torch.round(1.0 / pos.min() - 1.0).item()) if pos.numel() else 0
    if dmax > 254:
        raise NotImplementedError(f"hop distance {dmax} > 254 does not fit the uint8 hop matrix")
    nbins = dmax + 2
    hop = torch.empty(R, hop_ld(N), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(R, nbins, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    nm = None if normalization_matrix is None else normalization_matrix.contiguous().float()
    check(lib.gnan_hops_from_reference(ptr(nd), ptr(nm), R, N, ptr(hop), hop.shape[1], ptr(cnt), nbins, ptr(flag),
                                       stream_handle()), "gnan_hops_from_reference")
    return HopData(hop, cnt, N)


def build_csr(edge_index, num_nodes, device):
    """Directed CSR (rowptr int32 [N+1], col int32 [E]) of edges followed source -> target, built on the device."""
    ei = torch.as_tensor(edge_index).to(device=device, dtype=torch.int64)
    if ei.numel() == 0:
        return torch.zeros(num_nodes + 1, dtype=torch.int32, device=device), torch.zeros(1, dtype=torch.int32, device=device)
    if int(ei.max().item()) >= num_nodes or int(ei.min().item()) < 0:
        raise ValueError("edge_index out of range")
    order = torch.argsort(ei[0] * num_nodes + ei[1])
    src, dst = ei[0][order], ei[1][order]
    key = src * num_nodes + dst
    if key.numel() > 1 and bool((key[1:] == key[:-1]).any().item()):
        raise ValueError("duplicate edges: the reference sums them into weight 2 (pre_process_datasets.py:109); "
                         "gnan_b200 needs a simple graph")
    deg = torch.bincount(src, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(deg, 0)
    return rowptr.to(torch.int32), dst.to(torch.int32).contiguous()


def _trim_counts(cnt256):
    """[R,256] level histogram -> [R, D+2] (levels 0..D, then the unreachable bin). One device sync."""
    finite = cnt256[:, :255]
    nz = (finite.sum(0) > 0).nonzero()
    D = int(nz.max().item()) if nz.numel() else 0
    return torch.cat([finite[:, :D + 1], cnt256[:, 255:256]], dim=1).contiguous()


MSBFS_MIN_NODES = 4096      # graphs at least this large use the bit-parallel multi-source BFS


def apsp(edge_index, num_nodes, device="cuda", row_begin=0, row_end=None, method="auto"):
    """All-pairs (or rows [row_begin,row_end) of the) hop distances of one graph on the GPU -> HopData.

    method: "auto" | "warp" (one warp per source: small graphs, few rows) | "msbfs" (bit-parallel multi-source BFS over
    1024 columns at a time: large graphs; the cost does not depend on the number of rows kept)."""
    lib = load()
    N = int(num_nodes)
    row_end = N if row_end is None else int(row_end)
    R = row_end - row_begin
    rowptr, col = build_csr(edge_index, N, device)
    hop = torch.empty(R, hop_ld(N), dtype=torch.uint8, device=device)
    cnt = torch.empty(R, 256, dtype=torch.int32, device=device)
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    if method == "auto":
        method = "msbfs" if N >= MSBFS_MIN_NODES and R * 8 >= N else "warp"
    if method == "msbfs":
        ws = torch.empty(max(lib.gnan_apsp_msbfs_workspace_bytes(N), 1), dtype=torch.uint8, device=device)
        check(lib.gnan_apsp_msbfs(ptr(rowptr), ptr(col), N, row_begin, row_end, ptr(hop), hop.shape[1], ptr(cnt), 256, ptr(flag),
                                  ptr(ws), ws.numel(), stream_handle()), "gnan_apsp_msbfs")
    else:
        ws = torch.empty(max(lib.gnan_apsp_bfs_workspace_bytes(N, R), 1), dtype=torch.uint8, device=device)
        check(lib.gnan_apsp_bfs(ptr(rowptr), ptr(col), N, row_begin, row_end, ptr(hop), hop.shape[1], ptr(cnt), 256, ptr(flag),
                                ptr(ws), ws.numel(), stream_handle()), "gnan_apsp_bfs")
    if R and int(flag.item()):
        raise NotImplementedError("a hop distance > 254 does not fit the uint8 hop matrix")
    return HopData(hop, _trim_counts(cnt) if R else cnt[:, :2], N, row_begin)


class PackedBatch:
    """A batch of small graphs in packed block-diagonal form (replaces the dense (sum N)^2 collate of
    batched_pyg_main.py:54-91).

    x [sumN,K] fp32, hop uint8 [sum n_b^2] (graph b's n_b x n_b block at hop_off[b]), node_off int32 [B+1],
    hop_off int64 [B+1], level_counts int32 [sumN,nbins], batch_vector int64 [sumN] (kept for API parity), y.
    """

    def __init__(self, x, hop, hop_off, node_off, level_counts, y=None, max_nodes=None):
        self.x, self.hop, self.hop_off, self.node_off, self.level_counts, self.y = x, hop, hop_off, node_off, level_counts, y
        self.max_nodes = max_nodes

    @property
    def num_graphs(self):
        return self.node_off.numel() - 1

    @property
    def nbins(self):
        return self.level_counts.shape[1]

    @property
    def batch_vector(self):
        sizes = (self.node_off[1:] - self.node_off[:-1]).long()
        return torch.repeat_interleave(torch.arange(self.num_graphs, device=sizes.device), sizes)

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device, non_blocking=True)
        return PackedBatch(mv(self.x), mv(self.hop), mv(self.hop_off), mv(self.node_off), mv(self.level_counts), mv(self.y),
                           self.max_nodes)

    def pin_memory(self):
        pm = lambda t: None if t is None else t.pin_memory()
        return PackedBatch(pm(self.x), pm(self.hop), pm(self.hop_off), pm(self.node_off), pm(self.level_counts), pm(self.y),
                           self.max_nodes)


def apsp_batched(edge_index, node_off, device="cuda", x=None, y=None):
    """Hop blocks of B small graphs (<= 256 nodes each) in one launch. edge_index uses GLOBAL node ids of the
    concatenated node set; node_off [B+1] are the graph boundaries."""
    lib = load()
    node_off = torch.as_tensor(node_off).to(device=device, dtype=torch.int32)
    B = node_off.numel() - 1
    sumN = int(node_off[-1].item())
    sizes = (node_off[1:] - node_off[:-1]).long()
    max_n = int(sizes.max().item()) if B else 1
    hop_off = torch.zeros(B + 1, dtype=torch.int64, device=device)
    hop_off[1:] = torch.cumsum(sizes * sizes, 0)
    total = int(hop_off[-1].item())
    rowptr, col = build_csr(edge_index, sumN, device)
    hop = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
    nb = min(256, max_n + 1)                     # a finite hop inside a graph is at most max_n - 1; last column = unreachable
    cnt = torch.empty(sumN, nb, dtype=torch.int32, device=device)
    flag = torch.zeros(2, dtype=torch.int32, device=device)           # [overflow, largest finite hop]
    check(lib.gnan_apsp_bfs_batched_n(ptr(rowptr), ptr(col), ptr(node_off), ptr(hop_off), B, max_n, sumN, total, ptr(hop), ptr(cnt), nb,
                                      ptr(flag), flag.data_ptr() + 4, stream_handle()), "gnan_apsp_bfs_batched")
    over, D = (int(v) for v in flag.tolist()) if B else (0, 0)
    if over:
        raise NotImplementedError("a hop distance > 254 does not fit the uint8 hop matrix")
    if sumN == 0:
        return PackedBatch(x, hop, hop_off, node_off, cnt[:, :2], y, max_n)
    lc = torch.cat([cnt[:, :D + 1], cnt[:, nb - 1:nb]], dim=1).contiguous()
    return PackedBatch(x, hop, hop_off, node_off, lc, y, max_n)


def pre_process(data, is_graph_task, data_name=None, processed_data_dir=None, device="cuda", reference_format=False):
    """Drop-in for pre_process_datasets.pre_process (:104-148): same call shape, same mutation of the inputs.

    Appends the constant-1 column to x (:108/:127) and attaches `.hop_data` (HopData, on `device`). With
    reference_format=True also attaches fp32 `.node_distances` / `.normalization_matrix` like the reference.
    Graph task: `data` is a list of graphs; node task: one graph (num_nodes inferred from edge_index like :128).
    With `processed_data_dir` (and `data_name`) the result is written once, like :144-148, but in the packed format of
    gnan_b200.packed ({dir}/{name}.gnan_b200.pt: 1 byte per node pair instead of 8); read it back with
    packed.PackedDataset.load (graph task) or packed.load_node (node task).
    """
    import os
    graphs = list(data) if is_graph_task else [data]
    for g in graphs:
        g.x = torch.cat((g.x, torch.ones(g.x.size(0), 1, dtype=g.x.dtype, device=g.x.device)), dim=-1)
        ei = torch.as_tensor(g.edge_index)
        n = g.x.size(0) if is_graph_task else (int(ei.max().item()) + 1 if ei.numel() else g.x.size(0))
        hd = apsp(ei, n, device=device)
        g.hop_data = hd
        if reference_format:
            g.node_distances, g.normalization_matrix = hd.reference_format()
    if processed_data_dir is not None and data_name is not None:
        path = os.path.join(processed_data_dir, f"{data_name}.gnan_b200.pt")
        if not os.path.exists(path):
            os.makedirs(processed_data_dir, exist_ok=True)
            from . import packed
            if is_graph_task:
                packed.PackedDataset.from_graphs(graphs, device=device, add_constant_column=False).save(path)
            else:
                packed.save_node(data, path)
    return data
