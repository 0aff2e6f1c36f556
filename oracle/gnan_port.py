"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement ("port") of the reference GNAN forward passes.

This file restates, as plain functions over stacked weight tensors, what the reference `nn.Module`s
compute, following the reference's own ORDER of operations (K sequential shape-function evaluations
written into a pre-allocated `fx`, the distance MLP evaluated on all N*N pairs, the K-resolved
batched matmul, then the sums), so that it can stand in for the reference on a box where
`/root/reference` is absent: as the parity checker in tests/ and smoke(), and as the timed
`cpu_baseline` / `--impl reference` arm of bench.py (kind = "port"). Backward is torch autograd on
CPU, exactly as in the reference (trainer.py:66).

It is pinned against outputs of the real reference (tests/golden/*.npz, written by
oracle/make_golden.py which imports /root/reference unmodified). The product package never imports it.

Reference lines restated:
  GNAN.py:55-79    TensorGNAN.forward            -> tensor_gnan_gnanpy
  models.py:358-384 TensorGNAN.forward (main.py's) -> tensor_gnan_models
  GNAN.py:146-172  GNAN.forward                   -> gnan_rowloop
  models.py:291-300 NAM.forward                   -> nam
  batched_pyg_main.py:133-184                     -> tensor_gnan_batched
"""
import torch


def to_torch(p, dtype=torch.float32, requires_grad=False):
    """numpy stacked dict -> torch stacked dict."""
    out = {}
    for k, v in p.items():
        if v is None:
            out[k] = None
        else:
            t = torch.as_tensor(v).to(dtype).clone()
            out[k] = t.requires_grad_(requires_grad) if t.numel() > 0 else t
    return out


def scalar_mlp(p, g, t):
    """Group g's MLP on a column t [M,1] -> [M,C]. GNAN.py:24-34 (fs) / :38-47 (rho); dropout inactive."""
    lin = torch.nn.functional.linear                      # the op nn.Linear runs (one addmm), as in the reference modules
    if p["w1"] is None:                                   # n_layers == 1: Linear(1, C)
        return lin(t, p["wo"][g], p["bo"][g])
    h = torch.relu(lin(t, p["w1"][g].view(-1, 1), p["b1"][g]))
    for l in range(p["wh"].shape[0]):
        h = torch.relu(lin(h, p["wh"][l, g], p["bh"][l, g]))
    return lin(h, p["wo"][g], p["bo"][g])


def shape_functions(p_fs, x):
    """fx[j,k,:] = f_k(x[j,k]) with the reference's K-loop and slice assignment. GNAN.py:57-62."""
    N, K = x.shape
    C = p_fs["wo"].shape[1]
    fx = torch.empty(N, K, C, dtype=x.dtype)
    for k in range(K):
        fx[:, k] = scalar_mlp(p_fs, k, x[:, k].view(-1, 1))
    return fx


def tensor_gnan_gnanpy(p_fs, p_rho, x, node_distances, normalization_matrix, normalize_rho=True,
                       is_graph_task=False):
    """GNAN.py:55-79. rho is fed node_distances / normalization_matrix (input normalised)."""
    N = x.shape[0]
    C = p_fs["wo"].shape[1]
    fx = shape_functions(p_fs, x)
    fx_perm = fx.permute(2, 0, 1)                                   # [C,N,K]
    nd = node_distances
    if normalize_rho:
        nd = torch.div(nd, normalization_matrix)                    # :65-66
    m = scalar_mlp(p_rho, 0, nd.flatten().view(-1, 1)).view(N, N, C)  # :67
    mf = torch.matmul(m.permute(2, 0, 1), fx_perm)                  # :70  [C,N,K]
    if not is_graph_task:
        out = mf.sum(dim=2)                                         # :73  [C,N]
    else:
        out = mf.sum(dim=1).sum(dim=1).view(1, -1)                  # :76-78 [1,C]
    return out.T                                                    # [N,C] | [C,1]


def nam(p_read, x):
    """models.py:291-300 (readout NAM): sum over features of per-feature MLPs."""
    return shape_functions(p_read, x).sum(dim=1)


def tensor_gnan_models(p_fs, p_rho, x, node_distances, normalization_matrix, normalize_rho=True,
                       is_graph_task=False, p_readout=None):
    """models.py:358-384. rho's OUTPUT is divided by the normalisation matrix; rho width may be 1."""
    N = x.shape[0]
    Cr = p_rho["wo"].shape[1]
    fx = shape_functions(p_fs, x)
    fx_perm = fx.permute(2, 0, 1)                                   # [Cf,N,K]
    m = scalar_mlp(p_rho, 0, node_distances.flatten().view(-1, 1)).view(N, N, Cr)
    if normalize_rho:
        m = torch.div(m, normalization_matrix.unsqueeze(-1))        # :368-370
    mf = torch.matmul(m.permute(2, 0, 1), fx_perm)                  # broadcast [Cr|Cf, N, K]
    if not is_graph_task:
        out = mf.sum(dim=2)
    else:
        hidden = mf.sum(dim=1)                                      # [Cf,K]
        if p_readout is not None:
            out = nam(p_readout, hidden)                            # :380-381
        else:
            out = hidden.sum(dim=1).view(1, -1)
    return out.T


def gnan_rowloop(p_fs, p_rho, x, node_distances, normalization_matrix, normalize_rho=True, node_ids=None):
    """GNAN.py:146-172: feature sums first, then one rho evaluation per target row."""
    N = x.shape[0]
    C = p_fs["wo"].shape[1]
    if node_ids is None:
        node_ids = range(N)
    f_sums = shape_functions(p_fs, x).sum(dim=1)                    # :157
    rows = []
    for node in node_ids:
        rho_dist = scalar_mlp(p_rho, 0, node_distances[node].view(-1, 1))   # [N,1|C]
        if normalize_rho:
            rho_dist = rho_dist / normalization_matrix[node].view(-1, 1)    # :163-168
        rows.append((rho_dist * f_sums).sum(dim=0))                 # :169
    return torch.stack(rows) if rows else torch.empty(0, C, dtype=x.dtype)


def tensor_gnan_batched(p_fs, p_rho, x_batch, dist_batch, batch_vector, is_graph_task=True):
    """batched_pyg_main.py:133-184: raw hop counts into rho, -1 entries zeroed AFTER rho, scatter per graph."""
    N = x_batch.shape[0]
    C = p_fs["wo"].shape[1]
    fx = shape_functions(p_fs, x_batch)
    emb = scalar_mlp(p_rho, 0, dist_batch.flatten().view(-1, 1)).view(N, N, C)
    emb = emb * (dist_batch >= 0).unsqueeze(-1).to(emb.dtype)       # :158-159
    mf = torch.matmul(emb.permute(2, 0, 1), fx.permute(2, 0, 1)).sum(dim=2).permute(1, 0)   # [N,C]
    if not is_graph_task:
        return mf
    B = int(batch_vector.max().item()) + 1
    out = torch.zeros(B, C, dtype=mf.dtype)
    return out.index_add(0, batch_vector, mf)                       # :176-181
