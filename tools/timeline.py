#!/usr/bin/env python
"""Per-stream device timeline of one eager forward at the bench shape (no profiler needed).

Every `styler_*_fwd` entry of the C ABI is bracketed by CUDA events on the stream it is enqueued on; a device-side sleep in
front of the forward lets the host run ahead, so the gaps seen are dependencies, not launch overhead.  Output: one row per
library call (stream, start, end, duration) + busy time per stream + the chain that ends last at each phase boundary.
  python tools/timeline.py [--B 64] [--precision bf16] [--csv gpurun_out/timeline.csv]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styler_b200 import STYLER, _lib, synthetic as so  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--csv", default=None)
args = ap.parse_args()

dev = torch.device("cuda:0")
m = STYLER(precision=args.precision)
m.load_state_dict(so.make_state_dict(0))
m = m.to(dev).eval()
B, L, T = args.B, 128, 1024
b = so.make_inputs(B=B, L=L, seed=1234, d_mode="const", frames=8)
t = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
a = (t["src_seq"], t["mel_target"], t["mel_aug"], t["p_norm"], t["e_input"], t["src_len"], t["mel_len"])
kw = dict(d_target=t["d_target"], p_target=t["p_target"], e_target=t["e_target"], max_src_len=L, max_mel_len=T,
          speaker_embed=t["speaker_embed"])
for _ in range(3):
    m(*a, **kw)
torch.cuda.synchronize()

h = _lib.lib()
import ctypes

base = torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(12e6))          # ~6 ms at 1.9 GHz: the host enqueues the whole forward behind it
_lib.check(h.styler_debug_trace(1), "trace")   # native: an event pair around every LEAF entry, also inside the composite calls
base.record()
m(*a, **kw)
end = torch.cuda.Event(enable_timing=True)
end.record()
torch.cuda.synchronize()
_lib.check(h.styler_debug_trace(0), "trace")
need = int(h.styler_debug_trace_dump(None, 0))
buf = ctypes.create_string_buffer(need + 16)
h.styler_debug_trace_dump(buf, need + 16)
ids = {}
rows = []
for line in buf.value.decode().splitlines():
    f = line.split()
    name, (a_, b_, c_, d_), sid, t0, t1 = f[0], map(int, f[1:5]), int(f[5]), float(f[6]), float(f[7])
    s_ = ids.setdefault(sid, len(ids))
    extra = ""
    if name == "conv1d":
        extra = "B%d T%d Cin%d N%d k%d" % (a_, b_, c_, d_ // 100, d_ % 100)
    elif name in ("attention", "bilstm"):
        extra = "%d %d %d" % (a_, b_, c_)
    rows.append((t0, t1, s_, name, extra))
rows.sort()
print("total %.3f ms (sleep excluded), %d leaf calls, %d streams" % (rows[-1][1] - rows[0][0] if rows else 0.0, len(rows), len(ids)))
print("%8s %8s %7s  s  call" % ("start", "end", "dur"))
for s0, s1, s_, name, extra in rows:
    print("%8.3f %8.3f %7.3f  %d  %s %s" % (s0, s1, s1 - s0, s_, name, extra))
busy = {}
agg = {}
for s0, s1, s_, name, extra in rows:
    busy[s_] = busy.get(s_, 0.0) + (s1 - s0)
    k = name + " " + extra
    agg[k] = (agg.get(k, (0, 0.0))[0] + 1, agg.get(k, (0, 0.0))[1] + (s1 - s0))
print("busy per stream (ms):", {k: round(v, 3) for k, v in sorted(busy.items())})
print("by call (count, total ms):")
for k, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  %-40s %3d  %.3f" % (k, n, tt))
if args.csv:
    with open(args.csv, "w") as f:
        f.write("start_ms,end_ms,stream,call,extra\n")
        for r in rows:
            f.write("%.4f,%.4f,%d,%s,%s\n" % r)
