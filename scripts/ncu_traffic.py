#!/usr/bin/env python
"""Extract the per-launch DRAM traffic of the radix pass from an `ncu --page raw --csv` export and record it in
profiles/radix_traffic.json (read by bench.py for roofline.traffic).
usage: ncu_traffic.py raw.csv <reads> <k> <source-description>"""
import csv
import json
import os
import sys

raw, reads, k, src = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(raw)))
h, units = rows[0], rows[1]
ir, iw, it, ik = (h.index(x) for x in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = []
for r in rows[2:]:
    if "k_radix_pass" in r[ik]:
        tot.append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
if not tot:
    sys.exit("no radix pass launch in " + raw)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "radix_traffic.json")
j = json.load(open(out)) if os.path.exists(out) else {"captures": []}
j["captures"] = [e for e in j["captures"] if not (e["reads"] == reads and e["k"] == k)]
j["captures"].append({"reads": reads, "k": k, "dram_bytes_per_launch": sum(tot) / len(tot), "launches": len(tot),
                      "kernel": rows[2][ik].split("(")[0], "source": src})
json.dump(j, open(out, "w"), indent=1)
print(j["captures"][-1])
