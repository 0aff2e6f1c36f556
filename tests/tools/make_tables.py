"""Markdown tables of DESIGN.md §6 / §7 from the committed bench lines: python tests/tools/make_tables.py [dir=profiles]"""
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "..", "profiles")


def load(n):
    p = os.path.join(d, f"r02_bench_n{n}.json")
    if not os.path.exists(p):
        return None
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


lines = {n: load(n) for n in (1, 2, 4, 8)}
print("| N | molecules (DP, weak): step, graphs/s, x | all-reduce | ogbn-arxiv C=40 (row-sharded, strong): step, nodes/s, x | all-gather S / reduce-scatter dS / all-reduce | parity (row-sharded / DP) |")
print("|---|---|---|---|---|---|")
base = lines[1]
for n, r in lines.items():
    if r is None:
        continue
    ax = (r.get("sub_records") or {}).get("arxiv") or {}
    bx = ((base or {}).get("sub_records") or {}).get("arxiv") or {}
    col = r.get("collectives_ms_per_step") or {}
    acol = ax.get("collectives_ms_per_step") or {}
    par = r.get("parity_vs_single_gpu") or {}
    ps = "—" if not par else f"{par['row_sharded']['max_param_grad_rel_err']:.1e} / {par['data_parallel']['max_param_grad_rel_err']:.1e}"
    print(f"| {n} | {r['ms_per_step']:.3f} ms, {r['value'] / 1e6:.1f} M, x{r['value'] / base['value']:.2f} | "
          f"{col.get('allreduce_gradients', 0):.3f} ms | "
          + (f"{ax['ms_per_step']:.2f} ms, {ax['value'] / 1e6:.2f} M, x{ax['value'] / bx['value']:.2f} | " if ax and 'value' in ax else "— | ")
          + (f"{acol.get('allgather_rows', 0):.2f} / {acol.get('reduce_scatter_rows', 0):.2f} / {acol.get('allreduce_gradients', 0):.2f} ms | " if acol else "— | ")
          + ps + " |")
print()
r = lines[1]
if r:
    print("| workload (BASELINE.json config) | step | value | e2e (h2d per step) | reference CPU | largest kernels (ms, eager attribution) |")
    print("|---|---|---|---|---|---|")
    recs = [("molecules, 32 768 graphs of 10-100 nodes incl. hop preprocessing ([4]; headline)", r)] + [
        ({"cora": "Cora shape, 2 708 x 1 434, C = 7 ([1])", "mutag": "Mutagenicity shape, 4 337 graphs per step ([0])",
          "pubmed": "PubMed shape, 19 717 x 501, C = 3 ([2])", "arxiv": "ogbn-arxiv shape, 169 343 x 129, C = 40, 28.7 GB of hops ([3])"}[k], v)
        for k, v in (r.get("sub_records") or {}).items()]
    for name, v in recs:
        if "value" not in v:
            continue
        e = v.get("e2e") or {}
        cb = v.get("cpu_baseline") or {}
        ks = (v.get("roofline") or {}).get("kernel_ms_per_step") or {}
        top = sorted(ks.items(), key=lambda kv: -kv[1])[:4]
        u = v["unit"].split("/")[0]
        print(f"| {name} | **{v['ms_per_step']:.3f} ms** | {v['value'] / 1e6:.2f} M {v['unit']} | {e.get('value', 0) / 1e6:.2f} M ({e.get('h2d_bytes_per_step', 0) / 1e6:.0f} MB) | "
              f"{cb.get('value', 0):.3g} {cb.get('unit', '')} ({cb.get('kind', '')}, {cb.get('cores', '')} cores) | "
              + ", ".join(f"{k} {t:.3f}" for k, t in top) + " |")
