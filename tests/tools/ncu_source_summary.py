"""Summarise an `ncu --page source --csv` export: instructions executed per opcode, and the instructions with most stall
samples (with their dominant stall reasons). usage: ncu -i x.ncu-rep --page source --csv --kernel-name regex:K > s.csv; python ncu_source_summary.py s.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ops, total_inst, samples = Counter(), 0, Counter()
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[col["Instructions Executed"]].isdigit():
        if r and r[0] in ("Address", "Kernel Name") and recs:
            break                       # next kernel / next view of the same kernel
        continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.rstrip(";")
    n = int(r[col["Instructions Executed"]] or 0)
    s = int(r[col["# Samples"]] or 0)
    ops[op.split(".")[0]] += n
    total_inst += n
    st = {c: int(r[col[c]] or 0) for c in stall_cols}
    for c, v in st.items():
        samples[c] += v
    recs.append((s, n, src, st))
print("instructions executed (warp-level):", total_inst)
for op, n in ops.most_common(25):
    print(f"  {op:12s} {n:12d} {100.0 * n / total_inst:5.1f}%")
tot_s = sum(samples.values())
print("stall samples by reason:")
for c, v in samples.most_common(12):
    print(f"  {c:28s} {v:8d} {100.0 * v / max(tot_s, 1):5.1f}%")
print("top instructions by samples:")
for s, n, src, st in sorted(recs, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {s:7d} smp {n:10d} exec  {src[:70]:70s} {[(k[6:], v) for k, v in top if v]}")
