"""Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, per launch) of one `ncu --set full` capture -> JSON.
bench.py reads profiles/ncu_traffic.json for its roofline.traffic field.
Usage: python tools/ncu_traffic.py capture.ncu-rep [more.ncu-rep ...] > profiles/ncu_traffic.json"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"source": [], "kernels": {}}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out["source"].append(rep.split("/")[-1])
    acc = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("vpc::", "").replace("void ", "").replace("<0>", "<false>").replace("<1>", "<true>")
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
        t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}[units[col["gpu__time_duration.sum"]]]
        acc.setdefault(name, []).append((b, t))
    for name, v in acc.items():
        if name in out["kernels"]:
            continue                       # first capture wins (the DBSCAN step before the ICP one)
        out["kernels"][name] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "ncu_time_us": sum(x[1] for x in v) / len(v), "launches": len(v)}
print(json.dumps(out, indent=1))
