#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/ (round-tagged).
  python tools/summarize_profiles.py r1
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]


def launches():
    path = os.path.join(G, "launches.csv")
    if not os.path.isfile(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    per, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("sb::(anonymous namespace)::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        key = (name, row["Grid Size"])
        per.setdefault(key, [0, 0.0])
        per[key][0] += 1
        per[key][1] += v
        tot += v
        n += 1
    out = ["# ncu launch list of bench.py steps (%s)" % tag, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sb:: -s <warm-up> -c %d python bench.py ...`" % n,
           "Cold-cache, serialised launch times: compare SHARES.  %d launches, %.1f us total (= %.2f forward steps)." % (n, tot, n / 159.0),
           "", "| kernel | grid | launches | total us | avg us | share |", "|---|---|---|---|---|---|"]
    for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %s | %d | %.1f | %.1f | %.1f%% |" % (k[0], k[1], c, t, t / c, 100 * t / tot))
    open(os.path.join(P, "launches_%s.md" % tag), "w").write("\n".join(out) + "\n")


def full(rep, title):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.isfile(path):
        return None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    sel = {k: d[k] for k in KEYS if k in d}
    sel["kernel"] = d.get("Kernel Name", ("", ""))[1]
    out = ["# ncu --set full: %s (%s)" % (title, tag), "", "kernel: `%s`" % sel["kernel"], "", "| metric | unit | value |", "|---|---|---|"]
    for k in KEYS:
        if k in sel:
            out.append("| %s | %s | %s |" % (k, sel[k][0], sel[k][1]))
    open(os.path.join(P, "ncu_%s_%s.md" % (rep, tag)), "w").write("\n".join(out) + "\n")
    return sel


launches()
ffn1 = full("ffn1", "FFN Conv1d k=9 256->1024, B=64 T=1024 bf16 (dominant kernel)")
full("attn", "attention, B=64 T=1024 4 heads bf16")
if ffn1 is not None:
    def num(k):
        u, v = ffn1[k]
        v = float(v.replace(",", ""))
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    Bk = 128   # bench.py runs the clean and noisy decodes as one batched pass: 2 x 64 utterances per FFN launch
    json.dump({"kernel": "conv1d_tc_kernel<bf16,relu,FAST> FFN Conv1d k=9 256->1024 (B=%d,T=1024), tools/prof_kernels.py --only ffn1 --B %d" % (Bk, Bk),
               "round": tag, "dram_bytes_per_launch": traffic,
               "source": "profiles/ncu_ffn1_%s.md (ncu --set full, one launch; part of the 268 MB output is still in L2 when the kernel ends)" % tag,
               "algorithmic_bytes_per_launch": Bk * 1024 * (256 + 1024) * 2 + 9 * 1024 * 256 * 2,
               "algorithmic_flops_per_launch": 2.0 * Bk * 1024 * 1024 * 9 * 256},
              open(os.path.join(P, "dominant_kernel.json"), "w"), indent=1)
for f in ("prof_kernels_bf16.json",):
    src = os.path.join(G, f)
    if os.path.isfile(src):
        json.dump(json.load(open(src)), open(os.path.join(P, f.replace(".json", "_%s.json" % tag)), "w"), indent=1)
print(os.listdir(P))
