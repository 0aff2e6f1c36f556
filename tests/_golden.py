"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np

from oracle import params as P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                     if os.path.basename(p).split("_")[0] in ("gnanpy", "models", "gnan"))
BATCHED_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "batched_*.npz")))
PREPROCESS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "preprocess_*.npz")))


def variant_of(name):
    if name.startswith("gnanpy_tensor"):
        return "gnanpy_tensor"
    if name.startswith("models_tensor"):
        return "models_tensor"
    if name.startswith("gnan_loop"):
        return "gnan_loop"
    if name.startswith("batched"):
        return "batched"
    raise ValueError(name)


def load(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    z["name"] = name
    z["sd"] = {k[3:]: v for k, v in z.items() if k.startswith("sd.")}
    if name.startswith("preprocess"):
        return z
    v = variant_of(name)
    z["variant"] = v
    m = z["meta"]
    if v == "batched":
        N, K, C, H, L, graph = [int(t) for t in m]
        z.update(N=N, K=K, C=C, H=H, L=L, is_graph_task=bool(graph), bias=True, rho_has_bias=True)
        z["fs"] = P.stack_mlps(z["sd"], [f"fs.{k}" for k in range(K)], 2, 3)
        z["rho"] = P.stack_mlps(z["sd"], ["rho"], 2, 3)
    else:
        N, K, C, H, L, graph, norm, rpf, ro, rb, b = [int(t) for t in m]
        z.update(N=N, K=K, C=C, H=H, L=L, is_graph_task=bool(graph), normalize_rho=bool(norm),
                 rho_per_feature=bool(rpf), readout_n_layers=ro, rho_has_bias=bool(rb), bias=bool(b))
        z["fs"] = P.stack_mlps(z["sd"], [f"fs.{k}" for k in range(K)], L, 3, bool(b))
        z["rho"] = P.stack_mlps(z["sd"], ["rho"], L, 2, bool(rb))
        if ro > 0 and graph and v == "models_tensor":
            z["readout"] = P.stack_mlps(z["sd"], [f"readout_nam.fs.{k}" for k in range(K)], ro, 3, bool(b))
            z["grad_readout"] = P.unflatten("grad_readout", z)
    z["grad_fs"] = P.unflatten("grad_fs", z)
    z["grad_rho"] = P.unflatten("grad_rho", z)
    return z


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / max(||b||, tiny) (the achievable criterion)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
