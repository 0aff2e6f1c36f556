#!/usr/bin/env python
"""Benchmark of the GNAN hot path (BASELINE.json metric: fwd+bwd nodes/s on node tasks, graphs/s on graph tasks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cora|pubmed|mutag] [--impl reference]

Default workload = BASELINE.json configs[1]: TensorGNAN (GNAN.py:9-79) node classification on a Cora-shaped synthetic
graph (2708 nodes, 1433 features + the constant column, 7 classes, hidden 64, 3 layers), dense all-pairs hop distances.
A step is forward + cross-entropy on the 140 train-mask rows + backward + Adam step (trainer.py:48-67).

One JSON line on stdout (rank 0). `value` times the step with inputs resident in HBM (CUDA events per step, L2 flushed
between steps); `e2e` times the same step through the module API from pinned HOST buffers (x, hop bytes, level counts
copied every step, loss read back every step). `roofline` describes the dominant kernel (the grouped shape-MLP
backward), timed live with CUDA events on the launching stream. `cpu_baseline` / `--impl reference` time the oracle's
port of the reference's own CPU path (oracle/gnan_port.py: the reference is pure Python and cannot travel to the GPU
box) with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, L = 64, 3   # run.sh:10-11


# ---------------------------------------------------------------------------------------------------------------------
# synthetic workloads
# ---------------------------------------------------------------------------------------------------------------------
def random_simple_graph(rng, n, n_undirected_edges, n_isolated):
    m = n - n_isolated
    need = n_undirected_edges
    got = np.zeros((0, 2), dtype=np.int64)
    while got.shape[0] < need:
        s = rng.integers(0, m, size=2 * (need - got.shape[0]) + 16)
        d = rng.integers(0, m, size=s.shape[0])
        e = np.stack([np.minimum(s, d), np.maximum(s, d)], 1)[s != d]
        got = np.unique(np.concatenate([got, e]), axis=0)
    got = got[rng.permutation(got.shape[0])[:need]]
    return np.concatenate([got, got[:, ::-1]]).T.copy()


def make_node_workload(name, seed=0):
    rng = np.random.default_rng(seed)
    if name == "cora":
        n, k_raw, c, e_und, iso, dens = 2708, 1433, 7, 5278, 54, 0.0127
        x = (rng.random((n, k_raw)) < dens).astype(np.float32)
        x /= np.maximum(x.sum(1, keepdims=True), 1.0)                       # row-normalised bag of words (datasets.py:94)
    elif name == "pubmed":
        n, k_raw, c, e_und, iso, dens = 19717, 500, 3, 44338, 0, 0.10
        x = ((rng.random((n, k_raw)) < dens) * rng.random((n, k_raw)) * 0.1).astype(np.float32)
    else:
        raise ValueError(name)
    x = np.concatenate([x, np.ones((n, 1), np.float32)], 1)                 # pre_process_datasets.py:127
    ei = random_simple_graph(rng, n, e_und, iso)
    y = rng.integers(0, c, size=n)
    mask = np.zeros(n, bool)
    mask[rng.permutation(n)[:140]] = True
    return SimpleNamespace(name=name, n=n, K=k_raw + 1, C=c, x=torch.from_numpy(x), edge_index=torch.from_numpy(ei),
                           y=torch.from_numpy(y), train_mask=torch.from_numpy(mask), unit="nodes/s", units_per_step=n)


def flops_per_eval(C):
    return 2 * H + (L - 2) * 2 * H * H + 2 * H * C      # forward FLOPs of one shape-function evaluation


# ---------------------------------------------------------------------------------------------------------------------
def clocks_sampler():
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                 "-i", os.environ.get("LOCAL_RANK", "0")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def clocks_summary(proc):
    if proc is None:
        return None
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except Exception:
        return None
    sm, mx, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in out.strip().splitlines():
        f = [t.strip() for t in line.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1])); mx = max(mx, float(f[2]))
        except ValueError:
            continue
        for nme, v in zip(names, f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(nme)
    if not sm:
        return None
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's port of the reference CPU path
# ---------------------------------------------------------------------------------------------------------------------
def reference_step_fn(wl, seed=0):
    """Returns (step_fn, n_threads). One step == the GPU arm's step, run by oracle.gnan_port on the host cores."""
    from oracle import apsp as oapsp
    from oracle import gnan_port
    from oracle import params as P
    from gnan_b200.GNAN import TensorGNAN
    torch.manual_seed(seed)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = TensorGNAN(wl.K, wl.C, L, H, normalize_rho=True, is_graph_task=False)
    m.fs.xavier_normal_(1.0); m.rho.xavier_normal_(1.0)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    fs = gnan_port.to_torch(P.stack_mlps(sd, [f"fs.{k}" for k in range(wl.K)], L, 3), torch.float32, True)
    rho = gnan_port.to_torch(P.stack_mlps(sd, ["rho"], L, 2), torch.float32, True)
    hop = oapsp.apsp(wl.edge_index.numpy(), wl.n)
    nd, nm = oapsp.reference_format(hop, oapsp.level_counts(hop))
    nd, nm = torch.from_numpy(nd), torch.from_numpy(nm)
    params = [t for d in (fs, rho) for t in d.values() if t is not None and t.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    loss_fn = torch.nn.CrossEntropyLoss()

    def step():
        opt.zero_grad()
        out = gnan_port.tensor_gnan_gnanpy(fs, rho, wl.x, nd, nm, True, False)
        loss = loss_fn(out[wl.train_mask], wl.y[wl.train_mask])
        loss.backward()
        opt.step()
        return float(loss.item())

    return step, threads


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, threads = reference_step_fn(wl)
    budget_s = 300.0
    t_begin = time.perf_counter()
    w_done = 0
    for _ in range(args.warmup):
        if time.perf_counter() - t_begin > budget_s / 4:
            break
        step(); w_done += 1
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    val = wl.units_per_step / (ms / 1e3)
    sample = f"{len(times)} full {wl.name}-shape steps (fwd+CE+bwd+Adam) after {w_done} warm-up, oracle/gnan_port.py on torch CPU"
    print(json.dumps({
        "impl": "reference", "metric": f"GNAN fwd+bwd {wl.unit} ({wl.name}-shape TensorGNAN)", "value": val, "unit": wl.unit,
        "n_gpus": args.gpus, "steps": len(times), "warmup": w_done, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, "cpu"),
        "cpu_baseline": {"value": val, "unit": wl.unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_config(wl, where):
    return {"workload": f"{wl.name}-shape TensorGNAN node classification (BASELINE.json configs[1])" if wl.name == "cora"
            else f"{wl.name}-shape TensorGNAN", "nodes": wl.n, "features": wl.K, "classes": wl.C, "hidden": H, "n_layers": L,
            "step": "forward + CE loss on 140 train rows + backward + Adam", "normalize_rho": True,
            "timing": "CUDA events per step; 256 MiB L2 flush between timed steps" if where == "gpu" else "perf_counter",
            "parallelism": "replicas only (SURVEY.md §8e: small node-level graph)"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gnan_b200", choices=["gnan_b200", "reference"])
    ap.add_argument("--workload", default="cora", choices=["cora", "pubmed"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32x3", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gnan_b200" else args.warmup
    wl = make_node_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args, wl)

    from gnan_b200 import _lib, ops
    from gnan_b200.GNAN import TensorGNAN
    from gnan_b200.preprocess import HopData, apsp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: torch.distributed.barrier()) if world > 1 else (lambda: None)

    torch.manual_seed(0)
    model = TensorGNAN(wl.K, wl.C, L, H, normalize_rho=True, is_graph_task=False, device=dev).to(dev)
    model.fs.xavier_normal_(1.0); model.rho.xavier_normal_(1.0)
    model.precision = args.precision
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    loss_fn = torch.nn.CrossEntropyLoss()
    hd = apsp(wl.edge_index, wl.n, device=dev)                      # GPU preprocessing (not timed here)
    x_d, y_d, mask_d = wl.x.to(dev), wl.y.to(dev), wl.train_mask.to(dev)
    idx_d = mask_d.nonzero().flatten()
    yl_d = y_d[idx_d]
    data_d = SimpleNamespace(x=x_d, hop_data=hd)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(data):
        opt.zero_grad(set_to_none=True)
        out = model.forward(data)
        loss = loss_fn(out.index_select(0, idx_d), yl_d)
        loss.backward()
        opt.step()
        return loss

    lib = _lib.load()
    for _ in range(args.warmup):
        step(data_d)
    torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------------------------
    ops.enable_timing(True)
    clk = clocks_sampler() if rank == 0 else None
    launches0 = lib.gnan_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier(); torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1)                                              # evict L2 (126 MB) between timed steps
        a.record(); step(data_d); b.record()
    torch.cuda.synchronize(); barrier()
    launches = lib.gnan_launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kt = ops.timing_results()
    ops.enable_timing(False)
    clocks = clocks_summary(clk)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = wl.units_per_step * world / (ms_per_step / 1e3)

    # ---- end to end from pinned host buffers --------------------------------------------------------------------
    x_h = wl.x.pin_memory()
    hop_h = hd.hop.cpu().pin_memory()
    cnt_h = hd.level_counts.cpu().pin_memory()
    h2d = x_h.numel() * 4 + hop_h.numel() + cnt_h.numel() * 4

    def e2e_step():
        data = SimpleNamespace(x=x_h.to(dev, non_blocking=True),
                               hop_data=HopData(hop_h.to(dev, non_blocking=True), cnt_h.to(dev, non_blocking=True), wl.n))
        return float(step(data).item())                              # loss read back every step (trainer.py:72)

    for _ in range(3):
        e2e_step()
    barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        e2e_step()
    b.record(); torch.cuda.synchronize(); barrier()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_val = wl.units_per_step * world * args.steps / (float(t.item()) / 1e3)

    if rank != 0:
        return
    hbm, tflops, peak_src = measured_peaks()
    calls, kms = kt.get("mlp_bwd", (0, 0.0))
    fs_calls = args.steps                                              # one fs-backward per step; rho's is tiny (G=1)
    alg_flops = 2.0 * flops_per_eval(wl.C) * wl.n * wl.K              # backward = 2x forward FLOPs; recompute not counted
    # the op is called twice per step (fs and rho table); the rho call is ~1e-4 of the work, so per-launch time of the
    # dominant (fs) launch ~= total / steps
    dur_ms = kms / max(fs_calls, 1)
    achieved = alg_flops / (dur_ms / 1e3) / 1e12 if dur_ms > 0 else 0.0
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    traffic = json.load(open(prof)).get(f"{wl.name}:mlp_bwd:{args.precision}") if os.path.exists(prof) else None
    line = {
        "metric": f"GNAN fwd+bwd {wl.unit} ({wl.name}-shape TensorGNAN)", "value": value, "unit": wl.unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split, fp32 accumulate)", "tf32": "tf32"}[args.precision],
        "data": "synthetic", "config": workload_config(wl, "gpu"),
        "e2e": {"value": e2e_val, "unit": wl.unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "mlp_bwd_kernel (grouped shape-MLP backward incl. partial-gradient reduce)",
                     "bound": "tensor", "achieved": achieved, "peak": tflops, "unit": "TFLOP/s", "frac": achieved / tflops,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_flops_per_launch": alg_flops,
                     "avg_launch_ms": dur_ms, "pipe": "fp32 FFMA (CUDA cores)" if args.precision == "fp32" else "tcgen05 tf32",
                     "kernel_ms_per_step": {k: v[1] / args.steps for k, v in kt.items()}},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        stepf, threads = reference_step_fn(wl)
        t0 = time.perf_counter(); stepf(); dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": wl.units_per_step / dt, "unit": wl.unit, "cores": threads, "kind": "port",
                                "sample": f"1 full {wl.name}-shape step (fwd+CE+bwd+Adam), no warm-up, oracle/gnan_port.py on torch CPU, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
