#!/usr/bin/env python
"""Per-CTA phase breakdown (clock64 stamps) of the tcgen05 conv kernel on the config-3 shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styler_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
B, T = 64, 1024
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev, dt)
fp = lambda *s: torch.randn(*s, generator=g).to(dev)
x256 = rnd(B, T, 256)
lens = torch.full((B,), T, dtype=torch.int64, device=dev)
ln = (fp(256) * 0.1 + 1, fp(256) * 0.1)
w1, b1 = rnd(9, 1024, 256, scale=0.02), fp(1024)
h1024 = torch.empty(B, T, 1024, device=dev, dtype=dt)
w2, b2 = rnd(1, 256, 1024, scale=0.03), fp(256)
y256 = torch.empty(B, T, 256, device=dev, dtype=dt)
wqkv, bqkv = rnd(1, 768, 256, scale=0.06), fp(768)
qk = torch.empty(B, T, 512, device=dev, dtype=dt)
vt = torch.empty(B, 256, T, device=dev, dtype=dt)
qkv_plain = torch.empty(B, T, 768, device=dev, dtype=dt)
wfc, bfc = rnd(1, 256, 256, scale=0.06), fp(256)
cases = {
    "ffn1": lambda: ops.conv1d(x256, w1, b1, pad=4, act=ops.ACT_RELU, out=h1024, impl=ops.IMPL_TC),
    "ffn2_ln": lambda: ops.conv1d(h1024, w2, b2, residual=x256, ln=ln, lens=lens, out=y256, impl=ops.IMPL_TC),
    "qkv_vt": lambda: ops.conv1d(x256, wqkv, bqkv, out=qk, vt=vt, vt_col0=512, impl=ops.IMPL_TC),
    "qkv_plain": lambda: ops.conv1d(x256, wqkv, bqkv, out=qkv_plain, impl=ops.IMPL_TC),
    "fc_ln": lambda: ops.conv1d(x256, wfc, bfc, residual=x256, ln=ln, lens=lens, out=y256, impl=ops.IMPL_TC),
    "fc_plain": lambda: ops.conv1d(x256, wfc, bfc, out=y256, impl=ops.IMPL_TC),
}
cap = 4096
buf = torch.zeros(cap * 8, dtype=torch.int64, device=dev)
names = ["setup(alloc+sync)", "first stage full", "mainloop issue", "mma drain -> epi start", "LN pass1", "epi main pass", "teardown"]
for name, fn in cases.items():
    fn(); fn()
    torch.cuda.synchronize()
    buf.zero_()
    _lib.check(_lib.lib().styler_debug_set_phase_buffer(_lib.ptr(buf), cap))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().styler_debug_set_phase_buffer(None, 0))
    st = buf.view(cap, 8).cpu()
    st = st[st[:, 0] != 0]
    d = (st[:, 1:] - st[:, :-1]).double()
    life = (st[:, 7] - st[:, 0]).double()
    print("%-10s %7.1f us  ctas=%d  lifetime med %.0f clk | " % (name, e0.elapsed_time(e1) * 1e3, st.shape[0], life.median()) +
          "  ".join("%s %.0f" % (n, d[:, i].median()) for i, n in enumerate(names)), flush=True)
    if "ln" not in name:   # non-LN tiles: slot5-slot4 = time in TMEM ld+wait, slot6-slot3 = TMA-store tail (slot3 reused)
        print("           epilogue detail: tmem ld+wait %.0f clk, compute+smem store %.0f clk, tma store tail %.0f clk" % (
            (st[:, 5] - st[:, 4]).double().median(), ((st[:, 3] - st[:, 4]) - (st[:, 5] - st[:, 4])).double().median(),
            (st[:, 6] - st[:, 3]).double().median()), flush=True)
