"""Measurements for the SURVEY §8f rows on one B200 (run under gpurun): python scratch/bench_next_rows.py
Mutagenicity-shape dataset (4337 graphs): dataset build + file size, one training epoch through gnan_b200.trainer at the
reference's batch size 1 and at mini-batch sizes, test epoch, interpretability export. Writes profiles/next_rows_r01.json."""
import json, os, sys, tempfile, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gnan_b200
from gnan_b200 import interpret, trainer
from gnan_b200.models import TensorGNAN
from gnan_b200.packed import PackedDataset

dev = "cuda"
wl = bench.make_graph_workload(seed=0)
no = wl.node_off.tolist()
ei = wl.edge_index
graphs = []
order = torch.argsort(ei[0])
src = ei[0][order]
bounds = torch.searchsorted(src, wl.node_off)
for b in range(len(no) - 1):
    e = ei[:, order[bounds[b]:bounds[b + 1]]] - no[b]
    graphs.append(SimpleNamespace(x=wl.x[no[b]:no[b + 1], :14], edge_index=e, y=wl.y[b:b + 1]))
res = {}

def sync_time(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps, out

PackedDataset.from_graphs(graphs[:64], device=dev)                     # warm-up
t, ds = sync_time(lambda: PackedDataset.from_graphs(graphs, device=dev))
res["from_graphs_s"] = t; res["graphs"] = len(ds); res["graphs_per_s_preprocess"] = len(ds) / t
with tempfile.TemporaryDirectory() as tmp:
    p = os.path.join(tmp, "mutag.gnan_b200.pt")
    t, _ = sync_time(lambda: ds.save(p)); res["save_s"] = t; res["file_bytes"] = os.path.getsize(p)
    t, _ = sync_time(lambda: PackedDataset.load(p, device=dev)); res["load_s"] = t
pairs = int((ds.sizes ** 2).sum())
res["reference_pt_payload_bytes"] = pairs * 8 + int(ds.x.numel()) * 4    # two fp32 [n,n] per graph + x (pickle overhead not counted)

torch.manual_seed(0)
model = TensorGNAN(15, 1, 3, 64, is_graph_task=True, readout_n_layers=0).to(dev)
model.fs.xavier_normal_(1.0); model.rho.xavier_normal_(1.0)
loss_fn = torch.nn.BCEWithLogitsLoss()
for prec in ("fp32", "tf32x3"):
    model.precision = prec
    for bs in (1, 32, 256, 4337):
        n_graphs = 512 if bs == 1 else len(ds)
        sub = ds if bs > 1 else PackedDataset(*[getattr(ds.batch(list(range(n_graphs))), k) for k in ("x", "node_off", "hop", "hop_off", "level_counts", "y")])
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        trainer.train_epoch(model, sub.loader(bs), loss_fn, opt, dev, is_graph_task=True)          # warm-up epoch
        t, out = sync_time(lambda: trainer.train_epoch(model, sub.loader(bs, shuffle=True), loss_fn, opt, dev, is_graph_task=True))
        res[f"train_epoch[{prec},batch={bs}]"] = {"graphs": n_graphs, "s": t, "graphs_per_s": n_graphs / t, "loss": out[0], "acc": out[1]}
        print(prec, bs, res[f"train_epoch[{prec},batch={bs}]"], flush=True)
    t, out = sync_time(lambda: trainer.test_epoch(model, ds.loader(4337), loss_fn, dev, is_graph_task=True))
    res[f"test_epoch[{prec},batch=4337]"] = {"s": t, "graphs_per_s": len(ds) / t}
# the reference's mode (one graph per step) replayed from one captured CUDA graph per graph size, without and with dropout
items = [SimpleNamespace(x=g.x, hop_data=None, y=g.y) for g in []]
from gnan_b200.preprocess import apsp
sub_items = []
for b in range(1024):
    g = graphs[b]
    sub_items.append(SimpleNamespace(x=torch.cat([g.x, torch.ones(g.x.shape[0], 1)], 1).to(dev), hop_data=apsp(g.edge_index, g.x.shape[0], device=dev), y=g.y))
for p_drop in (0.0, 0.6):
    torch.manual_seed(0)
    mm = TensorGNAN(15, 1, 3, 64, dropout=p_drop, is_graph_task=True, readout_n_layers=0).to(dev)
    mm.fs.xavier_normal_(1.0); mm.rho.xavier_normal_(1.0); mm.train()
    for cap in (False, True):
        opt = torch.optim.Adam(mm.parameters(), lr=1e-3)
        trainer.train_epoch(mm, sub_items, loss_fn, opt, dev, is_graph_task=True, capture_steps=cap)      # warm-up / capture epoch
        t, out = sync_time(lambda: trainer.train_epoch(mm, sub_items, loss_fn, opt, dev, is_graph_task=True, capture_steps=cap))
        res[f"train_epoch[batch=1,dropout={p_drop},captured={cap}]"] = {"graphs": len(sub_items), "s": t, "graphs_per_s": len(sub_items) / t, "loss": out[0]}
        print("batch1", p_drop, cap, res[f"train_epoch[batch=1,dropout={p_drop},captured={cap}]"], flush=True)
model.train()
t, f = sync_time(lambda: interpret.shape_function_table(model, torch.linspace(-1, 1, 1001)), reps=5)
res["shape_function_table_1001x15_s"] = t
t, z = sync_time(lambda: interpret.heatmap(model, 30), reps=5)
res["heatmap_15x31_s"] = t
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "next_rows_r01.json"), "w"), indent=1)
