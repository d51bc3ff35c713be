"""Summarise an ncu source-page CSV: hottest source lines / SASS instructions by stall samples.
Usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K > f.csv; python tools/ncu_hot.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line") or (len(r) > 3 and "Source" in r))
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
src_c = col.get("Source")
smp_c = col.get("# Samples") or col.get("Warp Stall Sampling (All Samples)")
stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
data = []
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= smp_c:
        continue
    try:
        s = float(r[smp_c] or 0)
    except ValueError:
        continue
    total += s
    data.append((s, r))
data.sort(key=lambda t: -t[0])
print(f"total samples {total:.0f}")
for s, r in data[:topn]:
    stalls = sorted(((float(r[i] or 0), h) for h, i in stall_cols if i < len(r) and (r[i] or "0").replace('.', '', 1).isdigit()), reverse=True)[:3]
    st = " ".join(f"{h[6:]}={v:.0f}" for v, h in stalls if v > 0)
    thr = r[col["Avg. Threads Executed"]] if "Avg. Threads Executed" in col else ""
    print(f"{100*s/max(total,1):5.1f}%  thr={thr:>5}  {r[src_c][:110]:110s} | {st}")
