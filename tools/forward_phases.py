#!/usr/bin/env python
"""Where a forward step goes: CUDA-event timing of the three phases of Engine.forward at the bench shape (eager launches;
the audio-encoder side streams are joined into the main stream at the phase boundaries, as in the real forward)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styler_b200 import STYLER, synthetic as so  # noqa: E402

dev = torch.device("cuda:0")
m = STYLER(precision="bf16")
m.load_state_dict(so.make_state_dict(0))
m = m.to(dev).eval()
eng = m._engine_for()
B, L, T = 64, 128, 1024
b = so.make_inputs(B=B, L=L, seed=1234, d_mode="const", frames=8)
t = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
names = ["encode (text enc || 4 audio branches, MLPs, duration predictor)", "variance adaptor (LR, pitch/energy predictors, embed sum)",
         "decode (4 FFT blocks at 2B, mel_linear, PostNet)"]
acc = [0.0, 0.0, 0.0]
N = 10
with torch.no_grad():
    for it in range(N + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        enc, log_d, post = eng.encode(t["src_seq"], t["speaker_embed"], t["mel_target"], t["mel_aug"], t["p_norm"], t["e_input"],
                                      t["src_len"], t["mel_len"])
        ev[1].record()
        x, xn, _, p_pred, e_pred, _ = eng.variance_adapt(enc, log_d, T, t["mel_len"], t["d_target"], t["p_target"], t["e_target"])
        ev[2].record()
        eng.decode(eng._xx, t["mel_len"].repeat(2), has_pos=True)
        eng.join_audio_streams()
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
for n, a in zip(names, acc):
    print("%-75s %.3f ms" % (n, a / N))
print("%-75s %.3f ms" % ("sum", sum(acc) / N))
