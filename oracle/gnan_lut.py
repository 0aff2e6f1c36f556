"""TEST INFRASTRUCTURE ONLY (oracle/): the reference forward restated in the re-ordered form the
reference README itself licenses (README.md:24-27, implementation.png):

    out[i,c] = sum_j W[i,j,c'] * S[j,c],   S[j,:] = sum_k f_k(x[j,k])           (GNAN.py:157,169)

with the pair weight W taken from a table over integer hop distances instead of an MLP evaluation per
pair (the N*N inputs of rho take only D+2 distinct values per row: GNAN.py:65-67, models.py:366-370).
It exists so parity can be checked at shapes where the shipped forward cannot run (PubMed/arxiv:
the [N*N,H] activation of GNAN.py:67 is 99.5 GB / 7.3 TB). Runs in float64 on CPU, gradients by autograd.
Checked against oracle/gnan_port.py and the golden fixtures in tests/test_oracle_golden.py.

hop: int64 [R,N], -1 = unreachable (reference: dijkstra inf -> 1/(inf+1) = 0, pre_process_datasets.py:112-114)
cnt: int64 [R,D+2], cnt[i,d] = #{j : hop[i,j]=d}, last column = number of unreachable j
     (reference: normalization_matrix[i,j] = cnt[i,hop[i,j]], pre_process_datasets.py:117-121)
"""
import torch


def feature_sums(p_fs, x):
    """S[j,:] = sum_k f_k(x[j,k]), all K groups at once (same math as gnan_port.shape_functions + sum)."""
    if p_fs["w1"] is None:
        return torch.einsum("nk,kc->nc", x, p_fs["wo"][:, :, 0]) + p_fs["bo"].sum(0)
    h = torch.relu(x.unsqueeze(-1) * p_fs["w1"] + p_fs["b1"])                  # [N,K,H]
    for l in range(p_fs["wh"].shape[0]):
        h = torch.relu(torch.einsum("nki,kji->nkj", h, p_fs["wh"][l]) + p_fs["bh"][l])
    return torch.einsum("nkj,kcj->nc", h, p_fs["wo"]) + p_fs["bo"].sum(0)


def table_mlp(p_rho, u):
    """rho on a flat vector of scalar inputs u [M] -> [M,Cr]."""
    t = u.reshape(-1, 1)
    if p_rho["w1"] is None:
        return t @ p_rho["wo"][0].T + p_rho["bo"][0]
    h = torch.relu(t * p_rho["w1"][0] + p_rho["b1"][0])
    for l in range(p_rho["wh"].shape[0]):
        h = torch.relu(h @ p_rho["wh"][l, 0].T + p_rho["bh"][l, 0])
    return h @ p_rho["wo"][0].T + p_rho["bo"][0]


def pair_weights(p_rho, hop, cnt, mode):
    """W[i,j,:] from tables. mode: 'none' | 'input' | 'output' | 'raw' (see module docstring / gnan_port)."""
    R, N = hop.shape
    nb = cnt.shape[1] if cnt is not None else int(hop.max().item()) + 2
    dt = p_rho["wo"].dtype
    d = torch.arange(nb, dtype=dt)
    idx = torch.where(hop < 0, torch.full_like(hop, nb - 1), hop)              # unreachable -> last bin
    if mode == "raw":                                                          # batched_pyg_main.py:154-159
        T = table_mlp(p_rho, d)
        T = torch.cat([T[:-1], torch.zeros_like(T[-1:])])                      # masked pairs contribute 0
        return T[idx]
    u = 1.0 / (1.0 + d)
    u[-1] = 0.0                                                                # 1/(inf+1)
    if mode == "input":                                                        # GNAN.py:65-67
        ui = u.unsqueeze(0) / cnt.to(dt).clamp(min=1)                          # [R,nb]; empty bins never gathered
        Ti = table_mlp(p_rho, ui.reshape(-1)).view(R, nb, -1)
        return torch.gather(Ti, 1, idx.unsqueeze(-1).expand(-1, -1, Ti.shape[-1]))
    T = table_mlp(p_rho, u)                                                    # [nb,Cr]
    W = T[idx]
    if mode == "output":                                                       # models.py:368-370, GNAN.py:163-168
        W = W / torch.gather(cnt.to(dt), 1, idx).unsqueeze(-1)
    return W


def forward_rows(p_fs, p_rho, x, hop, cnt, mode):
    """out[i,:] for the R rows of `hop` (a row block of the full matrix). [R,C]."""
    S = feature_sums(p_fs, x)                                                  # [N,C]
    W = pair_weights(p_rho, hop, cnt, mode)                                    # [R,N,Cr]
    return (W * S.unsqueeze(0)).sum(dim=1)


def forward_graph(p_fs, p_rho, x, hop, cnt, mode):
    """Graph-level readout of GNAN.py:76-79: sum over nodes, returned as [C,1]."""
    return forward_rows(p_fs, p_rho, x, hop, cnt, mode).sum(dim=0).view(-1, 1)


def hops_from_reference(node_distances):
    """Invert node_distances = 1/(1+hop) (0 for unreachable) -> int64 hops with -1 unreachable."""
    nd = node_distances.double()
    hop = torch.round(1.0 / nd.clamp(min=1e-30) - 1.0).long()
    return torch.where(nd > 0, hop, torch.full_like(hop, -1))


def counts_from_hops(hop):
    R = hop.shape[0]
    D = int(hop.max().item())
    nb = D + 2
    idx = torch.where(hop < 0, torch.full_like(hop, nb - 1), hop)
    cnt = torch.zeros(R, nb, dtype=torch.long)
    cnt.scatter_add_(1, idx, torch.ones_like(idx))
    return cnt
