"""torch profiler over a few eager Cora-shape steps (run under gpurun)."""
import sys, torch
from types import SimpleNamespace
sys.path.insert(0, "/root/repo")
import bench
from gnan_b200.GNAN import TensorGNAN
from gnan_b200.preprocess import apsp
from gnan_b200.sparse import compress_features
wl = bench.make_node_workload("cora")
dev = "cuda"
torch.manual_seed(0)
m = TensorGNAN(wl.K, wl.C, 3, 64, normalize_rho=True).to(dev)
m.fs.xavier_normal_(1.0); m.rho.xavier_normal_(1.0); m.precision = "tf32x3"
hd = apsp(wl.edge_index, wl.n, device=dev)
x = wl.x.to(dev)
data = SimpleNamespace(x=x, hop_data=hd, x_compressed=compress_features(x))
opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
y = wl.y.to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(m.forward(data), y); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
