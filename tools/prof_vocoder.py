#!/usr/bin/env python
"""HiFi-GAN vocoder row: waveform samples/s of styler_b200.vocoder.Generator on one B200 for a
batch of mel spectrograms already resident in HBM (CUDA events, 3 warm-ups), beside the oracle port on the host cores on a
bounded sample.  Algorithmic work: 2 * MACs of every convolution (ConvTranspose1d counted at its true k taps).
  python tools/prof_vocoder.py [--B 8] [--T 1024] [--precision bf16]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops_per_utt(T, h):
    fl, ch, t = 2.0 * T * 80 * h["upsample_initial_channel"] * 7, h["upsample_initial_channel"], T
    for u, k in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
        fl += 2.0 * t * ch * (ch // 2) * k          # every input frame touches all k taps of every (ci, co) pair
        ch, t = ch // 2, t * u
        for ks in h["resblock_kernel_sizes"]:
            fl += 6 * 2.0 * t * ch * ch * ks
    return fl + 2.0 * t * ch * 7


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-frames", type=int, default=64)
    args = ap.parse_args()
    from styler_b200.vocoder import CONFIG_V1, Generator
    from styler_b200 import _lib
    from oracle import hifigan_oracle as ho
    dev = torch.device("cuda:0")
    sd = ho.make_state_dict(seed=0)
    voc = Generator(precision=args.precision)
    voc.load_state_dict(sd)
    voc = voc.to(dev).eval()
    mel = ho.make_mel(args.B, args.T, seed=0).to(dev)
    for _ in range(3):
        voc(mel)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        wav = voc(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    launches = (_lib.launch_count() - n0) // args.iters
    fl = flops_per_utt(args.T, CONFIG_V1) * args.B
    peak = 1368.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    if args.precision != "bf16":
        peak /= 2
    # oracle port on the host cores, bounded sample
    melc = ho.make_mel(1, args.cpu_frames, seed=0)
    with torch.no_grad():
        ho.generator_forward(sd, melc)
        t0 = time.perf_counter()
        ho.generator_forward(sd, melc)
        cpu_s = time.perf_counter() - t0
    print(json.dumps({"metric": "waveform samples/sec (HiFi-GAN V1 generator forward)", "value": wav.numel() / (ms * 1e-3),
                      "unit": "samples/s", "ms_per_batch": ms, "B": args.B, "mel_frames": args.T, "dtype": args.precision,
                      "gpu_launches_per_forward": launches, "tflops": fl / (ms * 1e-3) / 1e12,
                      "frac_of_measured_sustained_tensor_peak": fl / (ms * 1e-3) / 1e12 / peak,
                      "realtime_factor_22050Hz": wav.numel() / (ms * 1e-3) / 22050.0,
                      "cpu_baseline": {"value": args.cpu_frames * 256 / cpu_s, "unit": "samples/s", "kind": "port",
                                       "cores": torch.get_num_threads(), "sample": "1 utterance x %d frames" % args.cpu_frames}}))


if __name__ == "__main__":
    main()
