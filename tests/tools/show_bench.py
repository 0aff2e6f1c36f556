"""print the per-workload summary of a bench.py JSON line: python tests/tools/show_bench.py <file>"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
def show(r, name):
    if "error" in r:
        print("==", name, "ERROR", r["error"][:200]); return
    print(f"== {name}: {r.get('ms_per_step'):.4f} ms  value {r.get('value'):.4g} {r.get('unit')}  e2e {(r.get('e2e') or {}).get('value'):.4g}  graph={r.get('cuda_graph')} launches={r.get('gpu_launches')} scaling={r.get('scaling')} n={r.get('n_gpus')}")
    rf = r.get("roofline") or {}
    print(f"   roofline {rf.get('kernel')}: {rf.get('achieved'):.4g} {rf.get('unit')} frac {rf.get('frac'):.4f}")
    print("   kernels", {k: round(v, 4) for k, v in (rf.get("kernel_ms_per_step") or {}).items()})
    ag = r.get("aggregation")
    if ag:
        for k, v in ag.items():
            if isinstance(v, dict):
                print(f"   agg {k}: {v.get('achieved'):.4g} GB/s frac {v.get('frac'):.4f} ({v.get('avg_launch_ms'):.4f} ms)")
    for k in ("strict_fp32", "dense_kernels"):
        if r.get(k):
            print("  ", k, r[k].get("ms_per_step"), {kk: round(vv, 4) for kk, vv in r[k]["kernel_ms_per_step"].items()}, r[k].get("dominant_kernel_tflops"))
    if r.get("collectives_ms_per_step"): print("   collectives", r["collectives_ms_per_step"])
    cb = r.get("cpu_baseline")
    if cb: print(f"   cpu {cb['value']:.4g} {cb['unit']} kind={cb['kind']} cores={cb['cores']}")
show(d, "head")
for k, v in (d.get("sub_records") or {}).items():
    show(v, k)
print("clocks", d.get("clocks"))
if d.get("parity_vs_single_gpu"): print("parity", d["parity_vs_single_gpu"])
