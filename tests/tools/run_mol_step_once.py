"""Three molecule-workload steps (CSR build + batched BFS + model forward/backward), eagerly, for ncu:
python tests/tools/run_mol_step_once.py [n_graphs]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gnan_b200.models import TensorGNAN  # noqa: E402
from gnan_b200.preprocess import apsp_batched  # noqa: E402
from gnan_b200.sparse import compress_features  # noqa: E402

wl = bench.make_mol_workload(seed=0, n_graphs=int(sys.argv[1]) if len(sys.argv) > 1 else 32768)
dev = "cuda"
torch.manual_seed(0)
model = TensorGNAN(wl.K, wl.C, bench.L, bench.H, normalize_rho=True, is_graph_task=True, readout_n_layers=0, device=dev).to(dev)
model.fs.xavier_normal_(1.0); model.rho.xavier_normal_(1.0)
model.precision = "tf32x3"
ei, x, y = wl.edge_index.to(dev), wl.x.to(dev), wl.y.to(dev)
cx = compress_features(x)
loss_fn = torch.nn.BCEWithLogitsLoss()
for _ in range(3):
    pk = apsp_batched(ei, wl.node_off.numpy(), device=dev, x=None, y=y, nbins=bench.MOL_NBINS, rscale=True)
    pk.x_compressed = cx
    model.zero_grad(set_to_none=True)
    loss = loss_fn(model(pk).flatten(), y)
    loss.backward()
torch.cuda.synchronize()
print(float(loss), pk.status.tolist())
