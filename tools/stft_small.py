#!/usr/bin/env python
"""Small STFT call for compute-sanitizer (racecheck / memcheck): every occupancy shape of the kernel, plain and extended entries,
more frame blocks than persistent CTAs would exist on a tiny grid is not needed -- the item loop runs whenever blocks > 3 x SMs,
so one long row batch is included.   compute-sanitizer --tool racecheck python tools/stft_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styler_b200 import _lib, ops  # noqa: E402
from styler_b200.stft import mel_filterbank  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
basis = torch.from_numpy(mel_filterbank(22050, 1024, 80, 0.0, 8000.0)).to(dev)
y = ((torch.rand(4, 20000, generator=g) * 2 - 1) * 0.9).to(dev)
ns = torch.tensor([20000, 9000, 600, 19999], dtype=torch.int64, device=dev)
big = ((torch.rand(48, 44100, generator=g) * 2 - 1) * 0.5).to(dev)      # 48 x 11 blocks of 16 frames > 3 x 148 CTAs: item loop
ref = None
for occ in (0, 1, 2, 3):
    _lib.set_tuning("STFT_OCC", occ)
    a = ops.stft_mel(y, basis)
    b = ops.stft_mel_ex(y * 1.2, basis, clamp=True, frame_major=True, energy_range=(0.1, 525.43), n_samples=ns)
    c = ops.stft_mel(big, basis)
    torch.cuda.synchronize()
    outs = [t.clone() for t in list(a) + [b[0], b[1], b[3]] + list(c)]
    if ref is None:
        ref = outs
    else:
        assert all(torch.equal(p, q) for p, q in zip(ref, outs)), occ
_lib.set_tuning("STFT_OCC", -1)
print("stft_small ok", [tuple(t.shape) for t in ref])
