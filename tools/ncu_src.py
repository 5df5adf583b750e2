#!/usr/bin/env python
"""Source-level digest of an ncu report (needs -lineinfo + --import-source on): stall mix, hottest SASS lines, and per-opcode
runs of the hot loop.   python tools/ncu_src.py gpurun_out/x.ncu-rep [min_exec]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("kernel:", rows[0][1][:120])
print("total samples", tot, " warp-instructions", sum(int(r[ix["Instructions Executed"]]) for r in data))
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print("hottest lines:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:28]:
    s = int(r[ix["# Samples"]])
    dom = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print("  %5d %5.1f%% exec=%9s  %-58s %s" % (s, 100.0 * s / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:58], dom))
if min_exec:
    print("opcode runs (exec >= %d):" % min_exec)
    run = None
    for i, r in enumerate(data):
        ex, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        if ex < min_exec:
            continue
        op = r[ix["Source"]].strip().split()
        op = op[1] if op[0].startswith("@") else op[0]
        if run and run[2] == op:
            run = (run[0], i, op, run[3] + s, run[4] + 1)
        else:
            if run:
                print("  %4d-%4d n=%3d samples=%5d %s" % (run[0], run[1], run[4], run[3], run[2]))
            run = (i, i, op, s, 1)
    if run:
        print("  %4d-%4d n=%3d samples=%5d %s" % (run[0], run[1], run[4], run[3], run[2]))
