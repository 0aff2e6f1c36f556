"""The in-step preprocessing of the molecule workload once (CSR build + batched BFS), for ncu:
python tests/tools/run_apsp_batched_once.py [n_graphs]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gnan_b200.preprocess import apsp_batched  # noqa: E402

wl = bench.make_mol_workload(seed=0, n_graphs=int(sys.argv[1]) if len(sys.argv) > 1 else 32768)
ei = wl.edge_index.cuda()
for _ in range(3):
    pk = apsp_batched(ei, wl.node_off.numpy(), device="cuda", nbins=bench.MOL_NBINS)
torch.cuda.synchronize()
print(pk.status.tolist())
