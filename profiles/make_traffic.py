#!/usr/bin/env python
"""ncu_traffic.json from an `ncu --set full` report: DRAM bytes read + written per launch and the
profiled duration, first captured launch of every kernel.  usage: make_traffic.py REPORT.ncu-rep NOTE"""
import csv
import io
import json
import subprocess
import sys

rep, note = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[idx[name]].replace(",", ""))
    u = units[idx[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1)
    return v * scale


kern = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("h2::", "")
    if name in kern:
        continue
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    kern[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic": rd + wr,
                  "ncu_duration_us": val(r, "gpu__time_duration.sum")}
json.dump({"source": f"ncu --set full --clock-control none, {note}", "kernels": kern}, sys.stdout, indent=1)
print()
