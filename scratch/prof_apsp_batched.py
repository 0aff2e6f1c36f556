"""Where does apsp_batched spend its time? (run under gpurun)"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from gnan_b200 import preprocess as P
wl = bench.make_mol_workload(seed=0)
dev = "cuda"
ei, no = wl.edge_index.to(dev), wl.node_off.to(dev)
def T(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, out
print("apsp_batched total %.2f ms" % T(lambda: P.apsp_batched(ei, no, device=dev))[0])
sumN = int(no[-1])
print("build_csr %.2f ms" % T(lambda: P.build_csr(ei, sumN, dev))[0])
cnt = torch.zeros(sumN, 256, dtype=torch.int32, device=dev)
print("trim_counts %.2f ms" % T(lambda: P._trim_counts(cnt))[0])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): P.apsp_batched(ei, no, device=dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
