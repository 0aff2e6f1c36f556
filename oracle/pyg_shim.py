"""TEST INFRASTRUCTURE ONLY (oracle/): a stand-in for the `torch_geometric` names the reference
imports at module top (GNAN.py:2-5, models.py:2-5, pre_process_datasets.py:4), so that the
UNMODIFIED reference files can be imported in the build container to generate golden vectors.
Nothing in the product package imports this file.
"""
import sys
import types

import numpy as np
import scipy.sparse


def _to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    # same contract as torch_geometric.utils.to_scipy_sparse_matrix: COO with unit data,
    # num_nodes inferred as max index + 1 when omitted (pre_process_datasets.py:128 omits it).
    ei = edge_index.cpu().numpy() if hasattr(edge_index, "cpu") else np.asarray(edge_index)
    row, col = ei[0], ei[1]
    n = int(ei.max()) + 1 if num_nodes is None else int(num_nodes)
    data = np.ones(row.shape[0]) if edge_attr is None else np.asarray(edge_attr)
    return scipy.sparse.coo_matrix((data, (row, col)), shape=(n, n))


def install():
    """Insert stub modules into sys.modules (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_gnan_stub", False):
        return
    pyg = types.ModuleType("torch_geometric")
    pyg._gnan_stub = True
    nn_mod = types.ModuleType("torch_geometric.nn")
    for name in ("GraphConv", "GINConv", "GATv2Conv", "GraphSAGE", "TransformerConv", "global_mean_pool"):
        setattr(nn_mod, name, type(name, (), {}))
    utils = types.ModuleType("torch_geometric.utils")
    utils.to_scipy_sparse_matrix = _to_scipy_sparse_matrix
    pyg.nn = nn_mod
    pyg.utils = utils
    sys.modules["torch_geometric"] = pyg
    sys.modules["torch_geometric.nn"] = nn_mod
    sys.modules["torch_geometric.utils"] = utils


def import_reference(ref_dir="/root/reference"):
    """Import the reference modules unmodified. Returns (GNAN_py, models_py, pre_process_py, BatchedTensorGNAN)."""
    install()
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import importlib
    gnan_py = importlib.import_module("GNAN")
    models_py = importlib.import_module("models")
    pre_py = importlib.import_module("pre_process_datasets")
    # batched_pyg_main.py loads a dataset at import (lines 192-193): exec only the class (lines 94-184)
    src = open(f"{ref_dir}/batched_pyg_main.py").read().splitlines()
    ns = {}
    exec(compile("\n".join(src[93:184]), f"{ref_dir}/batched_pyg_main.py[94:184]", "exec"), ns)
    return gnan_py, models_py, pre_py, ns["TensorGNAN"]
