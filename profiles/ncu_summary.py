#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (raw page) + stall/opcode breakdown (source page).
usage: ncu_summary.py REPORT.ncu-rep [kernel-regex-for-source-page]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]][:70])
        for w in WANT:
            if w in idx:
                print(f"  {w:66s} {r[idx[w]]:>18s} {units[idx[w]]}")


def source(rep, regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, nsamp, op = collections.Counter(), 0, collections.Counter()
    for r in data:
        nsamp += int(r[idx["# Samples"]] or 0)
        for h in stalls:
            tot[h] += int(r[idx[h]] or 0)
        parts = r[idx["Source"]].split()
        if parts:
            o = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            op[o] += int(r[idx["Instructions Executed"]] or 0)
    print(f"==== source page, kernels matching {regex}: {nsamp} samples")
    for h, v in tot.most_common(8):
        print(f"  {h:28s} {100 * v / max(nsamp, 1):5.1f}%")
    tt = sum(op.values())
    print(f"  warp instructions: {tt}")
    for o, v in op.most_common(12):
        print(f"  {o:28s} {100 * v / tt:5.1f}%")
    top = sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:14]
    print("  hottest instructions:")
    for r in top:
        st = {h: int(r[idx[h]] or 0) for h in stalls}
        print(f"    {r[idx['# Samples']]:>7s}  {r[idx['Source']][:70]:70s} {max(st, key=st.get)}")


if __name__ == "__main__":
    raw(sys.argv[1])
    for rx in sys.argv[2:]:
        source(sys.argv[1], rx)
