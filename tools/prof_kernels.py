#!/usr/bin/env python
"""Isolated timings (CUDA events, L2 flushed between iterations) of the hot kernels at BASELINE config-3/4 shapes,
with algorithmic FLOPs/bytes and the fraction of the measured peaks.  Also the target of the ncu captures:
   ncu --set full -k regex:conv1d_tc_kernel -s 3 -c 1 python tools/prof_kernels.py --only ffn1
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styler_b200 import ops  # noqa: E402


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops"]), float(p["hbm_gbs"])
    except Exception:
        return 1590.0, 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--B", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    es = 2 if dt == torch.bfloat16 else 4
    B, T, L = args.B, 1024, 128
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    tf_peak, hbm_peak = peaks()
    if dt == torch.float32:
        tf_peak /= 2

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev, dt)

    def fp(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev)

    cases = {}
    x256 = rnd(B, T, 256)
    lens = torch.full((B,), T, dtype=torch.int64, device=dev)
    ln = (fp(256) * 0.1 + 1, fp(256) * 0.1)

    w1, b1 = rnd(9, 1024, 256, scale=0.02), fp(1024)
    h1024 = torch.empty(B, T, 1024, device=dev, dtype=dt)
    cases["ffn1"] = (lambda: ops.conv1d(x256, w1, b1, pad=4, act=ops.ACT_RELU, out=h1024, impl=ops.IMPL_TC),
                     2.0 * B * T * 1024 * 9 * 256, B * T * (256 + 1024) * es + w1.numel() * es, "tensor")
    w2, b2 = rnd(1, 256, 1024, scale=0.03), fp(256)
    y256 = torch.empty(B, T, 256, device=dev, dtype=dt)
    cases["ffn2_ln"] = (lambda: ops.conv1d(h1024, w2, b2, residual=x256, ln=ln, lens=lens, out=y256, impl=ops.IMPL_TC),
                        2.0 * B * T * 256 * 1024, B * T * (1024 + 256 + 256) * es, "tensor")
    wqkv, bqkv = rnd(1, 768, 256, scale=0.06), fp(768)
    qk = torch.empty(B, T, 512, device=dev, dtype=dt)
    vt = torch.empty(B, 256, T, device=dev, dtype=dt)
    cases["qkv"] = (lambda: ops.conv1d(x256, wqkv, bqkv, out=qk, vt=vt, vt_col0=512, impl=ops.IMPL_TC),
                    2.0 * B * T * 768 * 256, B * T * (256 + 768) * es, "tensor")
    wfc, bfc = rnd(1, 256, 256, scale=0.06), fp(256)
    cases["fc_ln"] = (lambda: ops.conv1d(x256, wfc, bfc, residual=x256, ln=ln, lens=lens, out=y256, impl=ops.IMPL_TC),
                      2.0 * B * T * 256 * 256, B * T * 256 * 3 * es, "hbm")
    ops.conv1d(x256, wqkv, bqkv, out=qk, vt=vt, vt_col0=512, impl=ops.IMPL_TC)
    ctx = torch.empty(B, T, 256, device=dev, dtype=dt)
    cases["attention"] = (lambda: ops.attention(qk, vt, lens, 4, out=ctx, impl=ops.IMPL_TC),
                          4.0 * B * 4 * T * T * 64, B * T * 1024 * es, "tensor")
    if dt == torch.bfloat16:   # the layout the engine uses in bf16: fused [B,T,768] buffer, V read row-major
        qkv = torch.empty(B, T, 768, device=dev, dtype=dt)
        ops.conv1d(x256, wqkv, bqkv, out=qkv, impl=ops.IMPL_TC)
        cases["qkv_fused"] = (lambda: ops.conv1d(x256, wqkv, bqkv, out=qkv, impl=ops.IMPL_TC),
                              2.0 * B * T * 768 * 256, B * T * (256 + 768) * es, "tensor")
        cases["attention_qkv"] = (lambda: ops.attention(qkv, None, lens, 4, out=ctx, impl=ops.IMPL_TC),
                                  4.0 * B * 4 * T * T * 64, B * T * 1024 * es, "tensor")
        # random-init-like score statistics (std ~0.4, as in the bench model): the running max is raised on the first key tile only
        qkv_lo = torch.cat([rnd(B, T, 256, scale=0.075), rnd(B, T, 256, scale=0.6), rnd(B, T, 256)], dim=-1).contiguous()
        cases["attention_lowvar"] = (lambda: ops.attention(qkv_lo, None, lens, 4, out=ctx, impl=ops.IMPL_TC),
                                     4.0 * B * 4 * T * T * 64, B * T * 1024 * es, "tensor")
    mel80 = rnd(B, T, 80)
    wp0, bp0 = rnd(5, 512, 80, scale=0.05), fp(512)
    h512 = torch.empty(B, T, 512, device=dev, dtype=dt)
    cases["postnet0"] = (lambda: ops.conv1d(mel80, wp0, bp0, pad=2, act=ops.ACT_TANH, out=h512, impl=ops.IMPL_TC),
                         2.0 * B * T * 512 * 5 * 80, B * T * (80 + 512) * es, "tensor")
    wp1, bp1 = rnd(5, 512, 512, scale=0.02), fp(512)
    h512b = torch.empty(B, T, 512, device=dev, dtype=dt)
    cases["postnet1"] = (lambda: ops.conv1d(h512, wp1, bp1, pad=2, act=ops.ACT_TANH, out=h512b, impl=ops.IMPL_TC),
                         2.0 * B * T * 512 * 5 * 512, B * T * 1024 * es, "tensor")
    xa320 = rnd(B, T, 320)
    wa, ba = rnd(5, 320, 320, scale=0.03), fp(320)
    cases["audio_c320_k5"] = (lambda: ops.conv1d(xa320, wa, ba, pad=2, impl=ops.IMPL_TC),
                              2.0 * B * T * 320 * 5 * 320, B * T * 640 * es, "tensor")
    wpl, bpl = rnd(5, 80, 512, scale=0.02), fp(80)
    melf = torch.zeros(B, T, 80, device=dev)
    postf = torch.empty(B, T, 80, device=dev)
    cases["postnet_last"] = (lambda: ops.conv1d(h512, wpl, bpl, pad=2, residual_f32=melf, out_f32=postf, want_out=False, impl=ops.IMPL_TC),
                             2.0 * B * T * 80 * 5 * 512, B * T * (512 * es + 80 * 8), "tensor")
    wc3, bc3 = rnd(3, 256, 256, scale=0.04), fp(256)
    cases["pred_conv_ln"] = (lambda: ops.conv1d(x256, wc3, bc3, pad=1, act=ops.ACT_RELU, ln=ln, out=y256, impl=ops.IMPL_TC),
                             2.0 * B * T * 256 * 3 * 256, B * T * 512 * es, "tensor")
    enc = rnd(B, L, 1280)
    dur = torch.full((B, L), 8, dtype=torch.int64, device=dev)
    encT = torch.empty(B, T, 1280, device=dev, dtype=dt)
    cases["length_regulator"] = (lambda: ops.length_regulator(enc, dur, T, out=encT), 0.0,
                                 (B * L * 1280 + B * T * 1280) * es + B * L * 8 + B * 8, "hbm")
    xa = rnd(B, T, 320)
    gam, bet = fp(320) * 0.1 + 1, fp(320) * 0.1
    cases["groupnorm_relu"] = (lambda: ops.groupnorm_relu_(xa, gam, bet), 0.0, 3 * B * T * 320 * es, "hbm")
    gx = fp(B, L, 640)
    whh = fp(2, 320, 80) * 0.1
    cases["bilstm_h80"] = (lambda: ops.bilstm_layer(gx, whh, dt), 0.0, B * L * 640 * 4, "latency")
    y = ((torch.rand(256, 88200, generator=g) * 2 - 1) * 0.5).to(dev)
    from styler_b200.stft import mel_filterbank
    basis = torch.from_numpy(mel_filterbank(22050, 1024, 80, 0.0, 8000.0)).to(dev)
    cases["stft_mel_c4"] = (lambda: ops.stft_mel(y, basis), 0.0, 256 * (88200 * 4 + 345 * 80 * 4 + 345 * 4), "hbm")

    rows = []
    for name, (fn, flops, nbytes, bound) in cases.items():
        if args.only and name not in args.only.split(","):
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tfl = flops / (ms * 1e-3) / 1e12
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append(dict(kernel=name, dtype=args.dtype, ms=ms, tflops=tfl, frac_tensor=tfl / tf_peak, gbs=gbs,
                         frac_hbm=gbs / hbm_peak, bound=bound, flops=flops, bytes=nbytes))
        print("%-18s %8.3f ms  %8.1f TFLOP/s (%.2f of %.0f)  %8.1f GB/s (%.2f of %.0f)  [%s]" % (
            name, ms, tfl, tfl / tf_peak, tf_peak, gbs, gbs / hbm_peak, hbm_peak, bound), flush=True)
    out = os.path.join(ROOT, "gpurun_out", "prof_kernels_%s.json" % args.dtype)
    if not args.only:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
