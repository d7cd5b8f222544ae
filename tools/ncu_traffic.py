#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from an `ncu --set full` report of bench.py: DRAM bytes (read + write) per launch of
the kernels bench.py reports a roofline for.  usage: tools/ncu_traffic.py <report.ncu-rep> <pairs> <kpts> <hyps>"""
import csv, json, os, subprocess, sys
rep, pairs, kpts, hyps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = "k_count" if "k_count" in name else "k_score" if "k_score" in name else "k_knn2_tc4" if "k_knn2_tc4" in name else None
    if not key or key in res:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]]
    res[key] = {"dram_bytes_per_launch": tot, "pairs": pairs, "kpts": kpts, "hyps": hyps,
                "duration_us_under_ncu": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) *
                {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[ix["gpu__time_duration.sum"]]],
                "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "ncu_traffic.json")
try:
    old = json.load(open(path))   # keep the entries of kernels this report does not contain
except Exception:
    old = {}
old.update(res)
res = old
json.dump(res, open(path, "w"), indent=1)
print(json.dumps(res, indent=1))
