#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the batched non-AR STYLER forward on N B200s (driver contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): per GPU a batch of 64 utterances,
128 phonemes -> 1024 mel frames (teacher-forced 8 frames/phoneme), Tr = 1024 reference frames, 80-bin mels, full
STYLER forward (style encoders, variance adaptor, LengthRegulator, clean + noisy decode, PostNet), bf16 compute.
A "step" = one forward over one resident batch; mel-frames/s = sum(mel_len) / time.  Weak scaling: 64 utterances per
rank; the step includes the NCCL gather of the four mel tensors to rank 0.

`value`  : device-timed (CUDA events, max over ranks), inputs already resident in HBM, rotating over 4 distinct
           batches so a step's inputs are never L2-resident from their previous use (4 x 46 MB > 126 MB L2; the
           per-step activation traffic is itself >> L2).
`e2e`    : same metric through the public API (styler_b200.STYLER.forward) with HOST (pinned) inputs: H2D of the
           step's inputs and D2H of the post-net mels + lengths inside the timed region.
`roofline`: dominant kernel = FFN Conv1d k=9 256->1024 implicit GEMM (conv1d_tc), CUDA events around each of its
           launches inside the timed steps; FLOPs/launch = 2*B*T*1024*9*256.
`cpu_baseline`: the oracle port (torch CPU fp32, all host threads) on a bounded sample (B=8 of the same workload).
`--impl reference`: times that CPU path alone (the reference is pure PyTorch; it cannot travel to the GPU box, so the
           oracle restatement -- pinned against the reference in tests/golden -- stands in, kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, L, T, FRAMES = 64, 128, 1024, 8
CPU_SAMPLE_B = 8
NBUF = 4


def load_peaks():
    """MEASURED_PEAKS.json (driver-written): cuBLAS bf16 burst and sustained TFLOP/s, STREAM-copy GB/s; else the
    profiling recipe's fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained, 6.65 TB/s)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]),
                    sm_max_mhz=float(p.get("sm_max_mhz", 1965.0)), src="measured")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, sm_max_mhz=1965.0, src="fallback")


def tensor_roofline(ach_tflops, peaks, clocks, scale=1.0):
    """Fractions of BOTH cuBLAS peaks; `frac`/`peak` is the one that matches the SM clock sampled during the timed region:
    the burst figure when the median clock stayed within 10 % of the maximum (a short run that never settles under the
    power cap), the sustained figure when the clock had dropped (a long, thermally settled step).  scale: 0.5 for tf32."""
    burst, sust = peaks["burst"] * scale, peaks["sustained"] * scale
    mhz = (clocks or {}).get("sm_mhz")
    mx = (clocks or {}).get("sm_max_mhz") or peaks["sm_max_mhz"]
    use_burst = mhz is None or mhz >= 0.9 * mx
    peak = burst if use_burst else sust
    return {"peak": peak, "frac": ach_tflops / peak, "frac_burst": ach_tflops / burst, "frac_sustained": ach_tflops / sust,
            "peak_source": "%s %s (median SM clock %s MHz of %s during the timed region)" % (
                peaks["src"], "bf16_tflops (burst)" if use_burst else "bf16_tflops_sustained", mhz, mx) +
            ("" if scale == 1.0 else " x %.1f (tf32)" % scale)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def mark(self):
        """Samples before this call (nvidia-smi start-up, warm-up) are not part of the timed region."""
        self.first = len(self.rows)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = self.rows[self.first:] if len(self.rows) > self.first else self.rows
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), start=3):
            if any(len(r) > i and r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist) BEFORE it allocates pinned host memory, so the step's
    H2D / D2H DMA does not cross the socket interconnect (8 ranks x 127 MB per step otherwise converge on one socket)."""
    try:
        prop = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)[:2] + ["..."] + [len(allowed)]
    except Exception:
        pass
    return None


def host_threads():
    """Threads the CPU arm may really use: affinity mask and cgroup CPU quota (a 128-CPU box with a small quota
    collapses under 128 torch threads), capped at 32 (the forward stops scaling beyond that)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, min(n, 32))


def cpu_forward_fps(max_b, budget_s=20.0):
    """Oracle port on the host cores: mel-frames/s of the same workload on a bounded sample (~budget_s of CPU work):
    one B=1 forward sizes the sample, then the best of up to 3 forwards at B<=max_b."""
    from oracle import styler_oracle as so
    from oracle import make_golden as mg
    torch.set_num_threads(host_threads())
    sd = so.make_state_dict(0)

    def once(b):
        batch = so.make_inputs(B=b, L=L, seed=1234, d_mode="const", frames=FRAMES)
        a, kw = mg.call_kwargs(batch)
        with torch.no_grad():
            t0 = time.perf_counter()
            so.styler_forward(sd, *a, **kw)
            return time.perf_counter() - t0

    once(1)                                           # warm-up (thread pools, allocator)
    t1 = once(1)
    b = max(1, min(max_b, int(budget_s / 3.0 / max(t1, 1e-3))))
    best, spent, reps = None, 0.0, 0
    while reps < 3 and (reps == 0 or spent + (best or 0) < budget_s):
        dt = once(b)
        spent += dt
        reps += 1
        best = dt if best is None else min(best, dt)
    return b * T / best, best, b, reps


def run_reference(args, rank, world):
    """CPU arm: the reference's own (PyTorch CPU) algorithm for the path, via the oracle port, all usable host threads.
    Each step is a bounded sample (B <= 8 utterances of the workload) sized from a B=1 probe so that K steps end within
    a few minutes."""
    if rank != 0:
        return
    t_all = time.perf_counter()
    from oracle import styler_oracle as so
    from oracle import make_golden as mg
    cores = host_threads()
    torch.set_num_threads(cores)
    sd = so.make_state_dict(0)

    def make(b):
        return mg.call_kwargs(so.make_inputs(B=b, L=L, seed=1234, d_mode="const", frames=FRAMES))

    with torch.no_grad():
        a, kw = make(1)
        so.styler_forward(sd, *a, **kw)
        t0 = time.perf_counter()
        so.styler_forward(sd, *a, **kw)
        t1 = time.perf_counter() - t0
        cap = int(os.environ.get("STYLER_BENCH_CPU_SAMPLE_B", CPU_SAMPLE_B))
        sample_b = max(1, min(cap, int(150.0 / (max(args.steps + args.warmup, 1) * max(t1, 1e-3)))))
        a, kw = make(sample_b)
        for _ in range(args.warmup):
            so.styler_forward(sd, *a, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            so.styler_forward(sd, *a, **kw)
        dt = time.perf_counter() - t0
    fps = args.steps * sample_b * T / dt
    sample = "B=%d utterances per step of the same workload (L=%d, T=%d, teacher-forced), fp32, %d threads" % (
        sample_b, L, T, cores)
    print(json.dumps({
        "impl": "reference", "metric": "mel-frames/sec (batched non-AR forward)", "value": fps, "unit": "mel-frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2] full STYLER forward, phoneme_len=%d -> mel_len=%d, 80-bin; CPU sample "
                               "of B=%d utterances per step" % (L, T, sample_b)},
        "cpu_baseline": {"value": fps, "unit": "mel-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}))


def vocoder_flops_per_utt(T, h):
    """2 * MACs of every convolution of the HiFi-GAN generator (ConvTranspose1d at its true k taps)."""
    fl, ch, t = 2.0 * T * 80 * h["upsample_initial_channel"] * 7, h["upsample_initial_channel"], T
    for u, k in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
        fl += 2.0 * t * ch * (ch // 2) * k
        ch, t = ch // 2, t * u
        for ks in h["resblock_kernel_sizes"]:
            fl += 6 * 2.0 * t * ch * ch * ks
    return fl + 2.0 * t * ch * 7


def run_vocoder(args):
    """Secondary workload: HiFi-GAN V1 generator forward, B=8 mel spectrograms of 1024 frames resident
    in HBM -> waveform samples/s on one B200 (CUDA events), with the oracle port timed on the host cores on a bounded sample."""
    from styler_b200 import _lib
    from styler_b200 import synthetic as syn
    from styler_b200.vocoder import CONFIG_V1, Generator
    B, T = 8, 1024
    dev = torch.device("cuda", 0)
    voc = Generator(precision=args.precision)
    voc.load_state_dict(syn.make_vocoder_state_dict(0))
    voc = voc.to(dev).eval()
    mel = syn.make_mel(B, T, seed=0).to(dev)
    for _ in range(max(args.warmup, 3)):
        voc(mel)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wav = voc(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.launch_count() - n0
    fl = vocoder_flops_per_utt(T, CONFIG_V1) * B
    peaks = load_peaks()
    roofv = tensor_roofline(fl / (ms * 1e-3) / 1e12, peaks, None, 1.0 if args.precision == "bf16" else 0.5)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import hifigan_oracle as ho          # CPU-baseline leg only
        torch.set_num_threads(host_threads())
        frames = 64
        sd, melc = ho.make_state_dict(0), ho.make_mel(1, frames, seed=0)
        with torch.no_grad():
            ho.generator_forward(sd, melc)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                ho.generator_forward(sd, melc)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        cpu = {"value": frames * 256 / best, "unit": "samples/s", "cores": host_threads(), "kind": "port",
               "sample": "oracle port (torch CPU fp32), 1 utterance x %d mel frames, best of 3 (%.2f s each)" % (frames, best)}
    print(json.dumps({
        "metric": "waveform samples/sec (HiFi-GAN V1 generator forward)", "value": wav.numel() / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "HiFi-GAN V1 generator (hifigan/config.json), B=%d mel spectrograms x %d frames -> %d samples each; "
                               "seeded synthetic weights" % (B, T, T * 256), "global_batch": B,
                   "l2": "per-step activations (134 MB per 32-channel tensor) >> L2"},
        "gpu_launches": launches,
        "roofline": dict({"bound": "tensor", "kernel": "whole forward (72 resblock convs + 4 transposed convs)",
                          "achieved": fl / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "flops_per_launch": fl, "traffic": None}, **roofv),
        "realtime_factor_22050Hz": wav.numel() / (ms * 1e-3) / 22050.0, "cpu_baseline": cpu}))


def _flush_buffer(dev):
    return torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > the 126 MB L2


def cpu_fftblock(B=16, Ln=128):
    """Oracle port (torch CPU fp32) of the configs[1] FFT-block call on the host threads."""
    from oracle import styler_oracle as so       # CPU-baseline leg only
    from styler_b200 import synthetic as syn
    torch.set_num_threads(host_threads())
    sd = syn.make_state_dict(2)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, Ln, 256, generator=g)
    lens = torch.randint(64, Ln + 1, (B,), generator=g)
    lens[0] = Ln
    mask = so.mask_from_lengths(lens, Ln)
    x = x.masked_fill(mask.unsqueeze(-1), 0)
    with torch.no_grad():
        so.fft_block(sd, "decoder.layer_stack.2.", x, mask)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            so.fft_block(sd, "decoder.layer_stack.2.", x, mask)
        dt = (time.perf_counter() - t0) / reps
    return {"value": B * Ln / dt, "unit": "tokens/s", "cores": host_threads(), "kind": "port",
            "sample": "oracle port (torch CPU fp32) of the same FFT block call, mean of %d (%.1f ms each)" % (reps, dt * 1e3)}, dt


def cpu_stft(y_host, F, cb=32):
    """Oracle port of the reference's dense conv-DFT mel path (audio/stft.py:51-79) on a bounded sample of the same utterances."""
    from oracle import stft_oracle               # CPU-baseline leg only
    torch.set_num_threads(host_threads())
    yc = y_host[:cb].clone()
    with torch.no_grad():
        stft_oracle.mel_spectrogram(yc, dense=True)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            stft_oracle.mel_spectrogram(yc, dense=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return {"value": cb * F / best, "unit": "STFT frames/s", "cores": host_threads(), "kind": "port",
            "sample": "oracle port of the reference's dense conv-DFT path (audio/stft.py:51-79), B=%d of the same 4 s utterances, "
                      "best of 3 (%.2f s each)" % (cb, best)}, best


def run_reference_side(args):
    """`--impl reference` for the secondary workloads: the CPU arm alone, same JSON keys."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "fftblock":
        cpu, dt = cpu_fftblock()
        metric, unit, wl = "tokens/sec (one FFTBlock: attention + Conv1d FFN)", "tokens/s", "BASELINE configs[1]: one FFTBlock, B=16, L=128, fp32 (CPU oracle port)"
    elif args.workload == "stft":
        g = torch.Generator().manual_seed(5)
        y = (torch.rand(32, 88200, generator=g) * 2 - 1) * 0.5
        cpu, dt = cpu_stft(y, 1 + 88200 // 256)
        metric, unit, wl = "STFT frames/sec (TacotronSTFT mel extraction)", "STFT frames/s", "BASELINE configs[3] sample: B=32 x 4 s utterances (CPU oracle port)"
    else:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU arm for workload %s" % args.workload}))
        return
    print(json.dumps({"impl": "reference", "metric": metric, "value": cpu["value"], "unit": unit, "n_gpus": args.gpus, "steps": 1,
                      "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": {"workload": wl}, "cpu_baseline": cpu,
                      "e2e": {"value": cpu["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_fftblock(args):
    """BASELINE configs[1]: ONE FFTBlock (attention + Conv1d FFN, transformer/Layers.py:26-34) at B=16, L=128, fp32 I/O --
    the fp32-parity tensor-core mode (tcgen05 kind::tf32, fp32 storage) unless --precision bf16.  tokens/s; L2 flushed
    between timed iterations (the whole problem, 2 MB of activations + 5 MB of weights, would otherwise sit in L2)."""
    from styler_b200 import _lib
    from styler_b200 import synthetic as syn
    from styler_b200.engine import Engine
    precision = args.precision if args.precision_given else "tf32"
    B, Ln = 16, 128
    dev = torch.device("cuda", 0)
    sd = syn.make_state_dict(2)
    eng = Engine(sd, dev, precision)
    W = eng.w.dec_layers[2]
    g = torch.Generator().manual_seed(9)
    x_host = torch.randn(B, Ln, 256, generator=g).pin_memory()
    lens = torch.randint(64, Ln + 1, (B,), generator=g)
    lens[0] = Ln
    x_host.masked_fill_(syn.mask_from_lengths(lens, Ln).unsqueeze(-1), 0)
    lens_d = lens.to(dev)
    x_dev = x_host.to(dev).to(eng.dt)
    out = torch.empty(B, Ln, 256, device=dev, dtype=eng.dt)
    flush = _flush_buffer(dev)
    steps, warm = args.steps, max(args.warmup, 3)
    for _ in range(warm):
        eng.fft_block(x_dev, lens_d, W, out=out)
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.__enter__()
    time.sleep(0.3)
    clocks.mark()
    n0 = _lib.launch_count()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.fft_block(x_dev, lens_d, W, out=out)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    ms = tot / steps
    # end to end: fp32 host buffer -> H2D -> (cast) -> block -> (cast) -> D2H of the fp32 result
    y_host = torch.empty(B, Ln, 256, dtype=torch.float32).pin_memory()

    def e2e_once():
        xd = x_host.to(dev, non_blocking=True)
        y = eng.fft_block(eng._act(xd), lens_d, W)
        y_host.copy_(y if y.dtype == torch.float32 else y.float(), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(warm):
        e2e_once()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_once()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    clocks.__exit__()
    flops = (5898240.0) * B * Ln                  # 5,767,168 + 1024*L per token at L=128 -> 12.08 GFLOP
    peaks = load_peaks()
    ck = clocks.summary()
    ach = flops / (ms * 1e-3) / 1e12
    roof = dict({"bound": "tensor", "kernel": "one FFT block = 5 launches (QKV, attention, out-proj+LN, Conv1d k9+ReLU, Conv1d k1+LN)",
                 "achieved": ach, "unit": "TFLOP/s", "flops_per_launch": flops, "traffic": None,
                 "note": "2048 tokens = 16 M-tiles of 128: far too small to fill 148 SMs; launch latency and tile quantisation bound it"},
                **tensor_roofline(ach, peaks, ck, 0.5 if precision == "tf32" else 1.0))
    cpu = None if args.no_cpu_baseline else cpu_fftblock(B, Ln)[0]
    print(json.dumps({
        "metric": "tokens/sec (one FFTBlock: attention + Conv1d FFN)", "value": B * Ln / (ms * 1e-3), "unit": "tokens/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if precision == "tf32" else precision, "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: one FFTBlock(256, 1024, 4 heads), B=%d, L=%d, ragged key mask, fp32 I/O, %s compute"
                               % (B, Ln, precision), "global_batch": B, "l2": "256 MB flush write between timed iterations"},
        "e2e": {"value": B * Ln / (e2e_ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": y_host.numel() * 4, "ms_per_step": e2e_ms},
        "gpu_launches": launches, "clocks": ck, "roofline": roof, "cpu_baseline": cpu}))


def run_stft(args):
    """BASELINE configs[3]: TacotronSTFT mel extraction, B=256 x 22050 Hz x 4 s, 1024-pt FFT, hop 256, 80 mels -> STFT frames/s."""
    from styler_b200 import TacotronSTFT, _lib
    B, N = 256, 88200
    F = 1 + N // 256
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    y_host = ((torch.rand(B, N, generator=g) * 2 - 1) * 0.5).pin_memory()
    ys = [y_host.to(dev), (y_host * 0.5).to(dev)]              # 2 x 90 MB inputs; L2 flushed anyway
    stft = TacotronSTFT().to(dev)
    flush = _flush_buffer(dev)
    steps, warm = args.steps, max(args.warmup, 3)
    for i in range(warm):
        stft.mel_spectrogram(ys[i % 2])
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    stft.mel_spectrogram(ys[0])
    per_call = _lib.launch_count() - n0            # library launches of one call (band scan + transform), counted eagerly
    # The two launches of a call are replayed as a CUDA graph (one per input buffer), like the headline workload: enqueued
    # eagerly from Python the second launch reaches the GPU up to ~0.1 ms after the first on a busy host, and that idle gap
    # is as long as a third of the 0.24 ms kernel.  --no-graph times the eager calls.
    graphs, launch_mode = None, "eager"
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            graphs = []
            for w in range(2):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    out = stft.mel_spectrogram(ys[w])
                graphs.append((gph, out))
            for gph, _ in graphs:
                gph.replay()
            torch.cuda.synchronize()
            launch_mode = "CUDA-graph replay of the call's two kernels"
        except Exception as ex:                    # capture is a convenience of the measurement, not of the product path
            graphs, launch_mode = None, "eager (graph capture failed: %s)" % str(ex)[:80]
            torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.__enter__()
    time.sleep(0.3)
    clocks.mark()
    tot = 0.0
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graphs is not None:
            graphs[i % 2][0].replay()
        else:
            mel, energy = stft.mel_spectrogram(ys[i % 2])
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    launches = per_call * steps
    ms = tot / steps
    mel_host = torch.empty(B, 80, F, dtype=torch.float32).pin_memory()
    en_host = torch.empty(B, F, dtype=torch.float32).pin_memory()

    def e2e_once():
        yd = y_host.to(dev, non_blocking=True)
        m, e = stft.mel_spectrogram(yd)
        mel_host.copy_(m, non_blocking=True)
        en_host.copy_(e, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(warm):
        e2e_once()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_once()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    clocks.__exit__()
    peaks = load_peaks()
    nbytes = B * (N * 4 + F * 80 * 4 + F * 4)      # 464,580 B per utterance -> 118.9 MB
    ach = nbytes / (ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "stft_kernel.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": "stft_mel_kernel (reflect pad + Hann + 1024-pt real FFT + magnitude + mel + log + energy, one launch)",
            "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "peak_source": peaks["src"] + " hbm_gbs",
            "bytes_per_launch": nbytes, "traffic": traffic,
            "note": "HBM is the contractual bound (SURVEY.md 8(d)); the FFT stage is ALU/issue work, see profiles/"}
    cpu = None if args.no_cpu_baseline else cpu_stft(y_host, F)[0]
    print(json.dumps({
        "metric": "STFT frames/sec (TacotronSTFT mel extraction)", "value": B * F / (ms * 1e-3), "unit": "STFT frames/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[3]: TacotronSTFT mel extraction, B=%d x 22050 Hz x 4 s (%d samples), n_fft=1024, hop=256, "
                               "80 mels -> %d frames each" % (B, N, F), "global_batch": B, "launch": launch_mode,
                   "l2": "256 MB flush write between timed iterations"},
        "utterances_per_s": B / (ms * 1e-3),
        "e2e": {"value": B * F / (e2e_ms * 1e-3), "unit": "STFT frames/s", "h2d_bytes_per_step": y_host.numel() * 4,
                "d2h_bytes_per_step": (mel_host.numel() + en_host.numel()) * 4, "ms_per_step": e2e_ms},
        "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roof, "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16", "tf32"])
    ap.add_argument("--pipeline", type=int, default=0, choices=[0, 1],
                    help="1: two-stage software pipeline over consecutive batches (model.PipelinedSTYLER): stage one (style encoders + "
                         "variance adaptor) of batch i+1 runs on a low-priority stream under stage two (decoder + PostNet) of batch i")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every forward kernel by kernel instead of replaying CUDA graphs")
    ap.add_argument("--no-extras", action="store_true", help="skip the tf32 and B=1 latency side measurements")
    ap.add_argument("--extras-pipeline", action="store_true",
                    help="also time the two-stage batch pipeline (PipelinedSTYLER) as a side leg (opt-in: one of eight pipelined runs at the "
                         "bench shape stopped making progress on a fresh box and was killed by its timeout; not reproduced, cause unknown)")
    ap.add_argument("--gather", default="push", choices=["push", "peer", "nccl"],
                    help="N > 1, how the mels reach rank 0: push = DMA copy of the packed results into rank 0's IPC-mapped region on a "
                         "side stream + our flag protocol (overlaps the next step, no SM time); peer = FUSED, the producing kernels store "
                         "into that region from their epilogues; nccl = one packed NCCL gather per step")
    ap.add_argument("--workload", default="styler", choices=["styler", "vocoder", "fftblock", "stft"],
                    help="styler = BASELINE.json's headline metric, configs[2] (default); fftblock = configs[1]; stft = configs[3]; "
                         "vocoder = secondary HiFi-GAN line (all but styler: single GPU)")
    args = ap.parse_args()
    args.precision_given = args.precision is not None
    args.precision = args.precision or "bf16"
    if args.impl == "reference" and args.workload != "styler":
        run_reference_side(args)
        return
    if args.workload == "vocoder":
        args.steps = min(args.steps, 10)
        run_vocoder(args)
        return
    if args.workload == "fftblock":
        run_fftblock(args)
        return
    if args.workload == "stft":
        run_stft(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    from styler_b200 import dist as sdist
    rank, world, local = sdist.init_from_env("gloo" if args.impl == "reference" else None)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from styler_b200 import synthetic as so      # seeded synthetic weights/inputs (no oracle import on the GPU arm)
    from styler_b200 import STYLER, GraphedSTYLER, PipelinedSTYLER, _lib
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    model = STYLER(precision=args.precision)
    model.load_state_dict(so.make_state_dict(0))
    model = model.to(dev).eval()

    keys = ("src_seq", "mel_target", "mel_aug", "p_norm", "e_input", "src_len", "mel_len", "d_target", "p_target",
            "e_target", "speaker_embed")
    host, resident = [], []
    for i in range(NBUF):
        b = so.make_inputs(B=B_PER_GPU, L=L, seed=1234 + 16 * rank + i, d_mode="const", frames=FRAMES)
        host.append({k: b[k].pin_memory() for k in keys})
        resident.append({k: b[k].to(dev) for k in keys})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    frames_per_step = int(host[0]["mel_len"].sum().item())

    def split(bt):
        a = (bt["src_seq"], bt["mel_target"], bt["mel_aug"], bt["p_norm"], bt["e_input"], bt["src_len"], bt["mel_len"])
        kw = dict(d_target=bt["d_target"], p_target=bt["p_target"], e_target=bt["e_target"], max_src_len=L, max_mel_len=T,
                  speaker_embed=bt["speaker_embed"])
        return a, kw

    # Fixed geometry -> the forward is replayed as a CUDA graph (GraphedSTYLER: ~130 launches + the side-stream fork/join of
    # the audio-encoder branches become one graph launch).  TWO graphs with their own static input/output buffers alternate,
    # so the gather (N > 1) or the D2H copies (e2e) of step i can still be reading graph k's outputs while step i+1 runs in
    # the other graph.
    use_graph = not args.no_graph
    dbg = (lambda m: print("[bench %.1f] %s" % (time.perf_counter(), m), file=sys.stderr, flush=True)) if os.environ.get("STYLER_BENCH_DEBUG") else (lambda m: None)
    LAUNCH_DESC = ("two-stage batch pipeline (PipelinedSTYLER): per slot one CUDA graph for style encoders + variance adaptor on a "
                   "low-priority stream and one for decoder + PostNet on a high-priority stream; stage one of batch i+1 overlaps stage "
                   "two of batch i; 2 slots" if args.pipeline else "CUDA-graph replay of the forward (2 alternating graphs)")
    from styler_b200.engine import packed_nbytes
    peer = world > 1 and args.gather == "peer"          # fused: the forward itself writes rank 0's memory
    push = world > 1 and args.gather == "push"
    gatherer = None
    if world > 1:
        gatherer = sdist.AsyncPeerGather(dev, packed_nbytes(B_PER_GPU, T), push=push) if (peer or push) else sdist.AsyncGather(dev)
    for i in range(args.warmup):                    # eager warm-up: engine build, position tables, allocator pools
        model(*split(resident[i % NBUF])[0], **split(resident[i % NBUF])[1])
    torch.cuda.synchronize()
    graphs = []
    pipe = None
    d2h_done = [None, None]                          # e2e: event behind the D2H copy that still reads slot k's results
    if use_graph and args.pipeline:
        a0, k0 = split(resident[0])
        pipe = PipelinedSTYLER(model, a0, k0, slots=2, result_mirrors=[gatherer.buffer(k) for k in range(2)] if peer else None,
                               back_priority=int(os.environ.get("STYLER_PIPE_PRIO", "-1")))
        dbg("pipeline captured")
    elif use_graph:
        a0, k0 = split(resident[0])
        # peer gather: graph k is captured with this rank's slice (slot k) of rank 0's receive region as the second destination
        # of mel_linear / the last PostNet conv -- the NVLink stores are part of the captured kernels
        graphs = [GraphedSTYLER(model, a0, k0, warmup=1, result_mirror=gatherer.buffer(k) if peer else None) for k in range(2)]

    def step(bt, slot):
        a, kw = split(bt) if bt is not None else ((), {})      # bt None: the inputs already sit in graph `slot`'s static buffers
        if pipe is not None:
            if bt is not None:
                pipe.load_inputs(slot, *a, **kw)

            def pre():                                # (stage-two stream current) flow control of the gather
                if peer or push:
                    gatherer.begin(slot)
                elif gatherer is not None:
                    gatherer.before_reuse(slot)

            def post():
                if gatherer is not None:
                    gatherer.launch_packed(pipe.packed[slot], slot)

            pipe.run(slot, after=(d2h_done[slot],), pre_back=pre, post_back=post)
            return pipe.outputs(slot)
        if peer or push:
            gatherer.begin(slot)                     # flow control: rank 0 has consumed the previous contents of this slot
        elif gatherer is not None and use_graph:
            gatherer.before_reuse(slot)              # the previous gather out of this graph's static outputs has drained
        if use_graph:
            out = graphs[slot](*a, **kw) if bt is not None else graphs[slot].replay()
            packed = graphs[slot].packed
        else:
            eng = model._engine_for()
            eng.result_mirror = gatherer.buffer(slot) if peer else None
            out = model(*a, **kw)
            eng.result_mirror = None
            packed = eng.last_packed
        if gatherer is not None:   # nccl: ONE gather (4 mels + lengths, one byte buffer) on the comm stream; peer: publish the
            gatherer.launch_packed(packed, slot)     # completion counter (rank 0: wait for all ranks, then release the slot)
        return out

    def barrier():
        if gatherer is not None:
            gatherer.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    enqueue = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
            if i == min(3, steps) - 1:               # host time per step while the launch queue is still empty: later steps
                enqueue[0] = (time.perf_counter() - t0) / (i + 1)   # are throttled by queue back-pressure to the device pace
        if pipe is not None:
            pipe.join()                              # both pipeline streams are joined into the timed stream before e1
        if gatherer is not None:
            gatherer.wait()                          # the timed events cover the gathers: join the comm stream before e1
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3, wall

    # nvidia-smi starts (and takes its driver locks) during the warm-up, not inside the timed region
    clocks = ClockSampler(local)
    clocks.__enter__()
    for i in range(args.warmup):
        step(resident[i % NBUF], i % 2)
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while not clocks.rows and clocks.proc is not None and time.perf_counter() - t_w < 3.0:
        time.sleep(0.05)                              # first sample seen: the sampler is up

    # ---- device-resident timed region (value) + per-launch events on the dominant kernel ----------------------
    launches0 = _lib.launch_count()
    clocks.mark()
    dbg("warm-up done, timing")
    secs, wall = timed(lambda i: step(resident[i % NBUF], i % 2), args.steps)
    dbg("timed region done")
    clocks.__exit__()
    launches_counted = _lib.launch_count() - launches0
    value = world * frames_per_step * args.steps / secs
    ck = clocks.summary()

    # Roofline leg: the dominant kernel (FFN Conv1d k9) timed live with CUDA events around each of its launches, recorded
    # inside the native FFT-block call while the same forward runs eagerly (events cannot be read back out of a graph
    # replay) -- same kernels, same shapes, same clocks; `launches_per_step` is counted in this pass too.
    import ctypes
    dbg("roofline leg")
    roof_steps = min(args.steps, 10)
    _lib.check(_lib.lib().styler_debug_ffn1_timing(1, T), "ffn1_timing")
    n0 = _lib.launch_count()
    for i in range(roof_steps):
        model(*split(resident[i % NBUF])[0], **split(resident[i % NBUF])[1])
    torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - n0) // max(roof_steps, 1)
    t_ms, t_n, t_b, t_t = ctypes.c_float(0), ctypes.c_int32(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(_lib.lib().styler_debug_ffn1_timing_read(ctypes.byref(t_ms), ctypes.byref(t_n), ctypes.byref(t_b), ctypes.byref(t_t)),
               "ffn1_timing_read")
    _lib.check(_lib.lib().styler_debug_ffn1_timing(0, 0), "ffn1_timing")
    torch.cuda.synchronize()
    launches = launches_per_step * args.steps if use_graph else launches_counted

    peaks = load_peaks()
    roof = None
    if t_n.value > 0:
        avg_ms = t_ms.value / t_n.value
        flops = 2.0 * int(t_b.value) * int(t_t.value) * 1024 * 9 * 256
        ach = flops / (avg_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = dict({"bound": "tensor", "kernel": "conv1d_tc_kernel<bf16> FFN Conv1d k=9 256->1024 (+bias+ReLU), B=%d T=%d" % (t_b.value, t_t.value),
                     "achieved": ach, "unit": "TFLOP/s", "avg_launch_ms": avg_ms, "launches_timed": int(t_n.value),
                     "flops_per_launch": flops, "traffic": traffic},
                    **tensor_roofline(ach, peaks, ck, 1.0 if args.precision == "bf16" else 0.5))
        roof["whole_step"] = {"flops": 5.53e12 * B_PER_GPU / 64, "achieved": 5.53e12 * B_PER_GPU / 64 / (secs / args.steps) / 1e12,
                              "unit": "TFLOP/s", "note": "SURVEY.md 8(d): 86.35 GFLOP per utterance x 64; step time from `value`"}

    # ---- end to end through the public API with host buffers ---------------------------------------------------
    # Two user streams alternate (each with its own graph / pinned result buffers): step i's H2D / D2H copies overlap step
    # i+-1's kernels, as a serving loop would drive the public API.  Every step does its own H2D of all inputs and the D2H of
    # ALL FOUR mel tensors + lengths (one packed buffer) inside the timed region.
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    d2h = [torch.empty(packed_nbytes(B_PER_GPU, T), dtype=torch.uint8).pin_memory() for _ in range(2)]

    compute_done = [None]                              # event after the previous step's kernels
    diag = os.environ.get("STYLER_BENCH_E2E_DIAG", "")  # diagnostic runs only: "noh2d" / "nod2h" drop one copy direction

    def e2e_step_pipe(i):
        k = i % 2
        with torch.cuda.stream(e2e_streams[k]):       # copy stream of slot k: H2D -> (pipeline streams: both stages) -> D2H
            hb = host[i % NBUF]
            pipe.load_inputs(k, *split(hb)[0], **split(hb)[1])
            step(None, k)
            e2e_streams[k].wait_event(pipe.done(k))
            d2h[k].copy_(pipe.packed[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(e2e_streams[k])
            d2h_done[k] = ev

    def e2e_step(i):
        if pipe is not None:
            return e2e_step_pipe(i)
        k = i % 2
        with torch.cuda.stream(e2e_streams[k]):
            hb = host[i % NBUF]
            if "noh2d" in diag:
                hb = resident[i % NBUF]
            if use_graph:                              # H2D of the (pinned host) inputs straight into graph k's static buffers
                graphs[k].load_inputs(*split(hb)[0], **split(hb)[1])
            else:
                bt = {kk: v.to(dev, non_blocking=True) for kk, v in hb.items()}
            # the copies of step i overlap the kernels of step i-1 (other stream); the KERNELS of consecutive steps run back to
            # back rather than interleaved (two forwards sharing the SMs thrash each other's L2 working set)
            if compute_done[0] is not None:
                e2e_streams[k].wait_event(compute_done[0])
            if use_graph:
                step(None, k)
                packed = graphs[k].packed
            else:
                step(bt, k)
                packed = model._engine.last_packed
            ev = torch.cuda.Event()
            ev.record(e2e_streams[k])
            compute_done[0] = ev
            if "nod2h" not in diag:
                d2h[k].copy_(packed, non_blocking=True)

    def e2e_timed(steps):
        barrier()
        for st in e2e_streams:
            st.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            e2e_step(i)
        for st in e2e_streams:
            st.synchronize()
        if pipe is not None:
            pipe.s_back.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    for i in range(max(2 * NBUF, args.warmup)):     # every host batch through both streams once: allocator pools, NCCL
        e2e_step(i)                                 # channels and receive buffers of the e2e streams exist before timing
    dbg("e2e warm-up done")
    e2e_secs = e2e_timed(args.steps)
    dbg("e2e done")
    e2e_value = world * frames_per_step * args.steps / e2e_secs
    d2h_bytes = d2h[0].numel()

    # ---- side measurements (rank 0, N = 1): the fp32-parity tf32 mode on the same workload, and BASELINE configs[0] latency ----
    extras = {}
    if world == 1 and not args.no_extras:
        torch.cuda.synchronize()
        del graphs[:]
        # every side measurement starts after a short idle period: the main legs leave the board at its power cap, and a leg that
        # starts there runs at the sustained clock (MEASURED_PEAKS clocks_under_load ~1.3 GHz), not at the clock of the headline
        time.sleep(2.0)
        xclk = ClockSampler(local)                   # SM clock under each side leg, next to its number
        xclk.__enter__()

        def leg_clock():
            time.sleep(0.12)
            c = xclk.summary()
            return {"sm_mhz": c.get("sm_mhz"), "reasons": c.get("reasons")}
        if pipe is None and use_graph and args.extras_pipeline:
            # the two-stage batch pipeline (model.PipelinedSTYLER) on the same workload, device-resident, same timing rules
            pp = PipelinedSTYLER(model, *split(resident[0]), slots=2)
            for i in range(4):
                pp.load_inputs(i % 2, *split(resident[i % NBUF])[0], **split(resident[i % NBUF])[1])
                pp.run(i % 2)
            pp.join()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xclk.mark()
            e0.record()
            for i in range(args.steps):
                pp.load_inputs(i % 2, *split(resident[i % NBUF])[0], **split(resident[i % NBUF])[1])
                pp.run(i % 2)
            pp.join()
            e1.record()
            torch.cuda.synchronize()
            msp = e0.elapsed_time(e1) / args.steps
            extras["pipelined_two_stage"] = {"ms_per_step": msp, "value": frames_per_step / (msp * 1e-3), "unit": "mel-frames/s",
                                             "steps": args.steps, "clocks": leg_clock(),
                                             "note": "PipelinedSTYLER: style encoders + variance adaptor of batch i+1 on a low-priority stream "
                                                     "under decoder + PostNet of batch i (two batches in flight); results bitwise those of forward()"}
            del pp
            torch.cuda.empty_cache()
        pipe = None
        for other in [m_ for m_ in ("fp16", "tf32", "bf16") if m_ != args.precision]:
            m2 = STYLER(precision=other)
            m2.load_state_dict(so.make_state_dict(0))
            m2 = m2.to(dev).eval()
            a0, k0 = split(resident[0])
            for _ in range(3):
                m2(*a0, **k0)
            torch.cuda.synchronize()
            g2 = GraphedSTYLER(m2, a0, k0, warmup=1)   # graph replay like the headline: an eager leg measured the host (7-10 ms of
            for i in range(3):                         # Python per forward when the sampler threads are busy), not the device
                g2(*split(resident[i % NBUF])[0], **split(resident[i % NBUF])[1])
            torch.cuda.synchronize()
            time.sleep(2.0)
            n2 = min(args.steps, 20)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xclk.mark()
            e0.record()
            for i in range(n2):
                a, kw = split(resident[i % NBUF])
                g2(*a, **kw)
            e1.record()
            torch.cuda.synchronize()
            del g2
            ms2 = e0.elapsed_time(e1) / n2
            extras[other + "_same_workload"] = {
                "ms_per_step": ms2, "value": frames_per_step / (ms2 * 1e-3), "unit": "mel-frames/s", "steps": n2, "clocks": leg_clock(),
                "note": "device-resident, CUDA-graph replay as the headline; SM clock of the leg in `clocks` (a leg that starts on a board "
                        "already at its power cap runs at the sustained clock); " +
                        {"tf32": "fp32 storage + tcgen05 kind::tf32; meets the 1e-3 fp32 tolerance on the mels",
                         "fp16": "IEEE-half storage + tcgen05 kind::f16 (the bf16 kernels, 11-bit significand); meets the 1e-3 fp32 "
                                 "tolerance on the mels (tests/test_forward_gpu.py, same gates as tf32)",
                         "bf16": "bf16 storage + tcgen05 kind::f16"}[other]}
            del m2
            torch.cuda.empty_cache()
        # BASELINE configs[0]: single utterance, 50 phonemes, free-running durations (8 frames/phoneme via the duration bias)
        sd1 = so.set_duration_bias(so.make_state_dict(0), FRAMES)
        m1 = STYLER(precision=args.precision)
        m1.load_state_dict(sd1)
        m1 = m1.to(dev).eval()
        b1 = so.make_inputs(B=1, L=50, Tr=400, seed=16, d_mode=None)
        a1 = tuple(b1[k].to(dev) for k in ("src_seq", "mel_target", "mel_aug", "p_norm", "e_input", "src_len", "mel_len"))
        k1 = dict(max_src_len=50, speaker_embed=b1["speaker_embed"].to(dev))

        def lat(fn, n=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
                torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n * 1e3

        eager_ms = lat(lambda: m1(*a1, **k1))
        g1 = GraphedSTYLER(m1, a1, dict(k1, max_mel_len=50 * FRAMES))
        graph_ms = lat(lambda: g1(*a1, **k1))
        extras["latency_config0_b1_l50"] = {"eager_ms": eager_ms, "graph_ms": graph_ms, "mel_frames": 50 * FRAMES,
                                            "note": "BASELINE configs[0] on the GPU: one utterance, 50 phonemes -> 400 frames, free-running; "
                                                    "wall clock per call incl. synchronize; eager has one host read of max(mel_len)"}
        del g1, m1
        xclk.__exit__()

    if peer or push:
        torch.cuda.synchronize()
        gatherer.close()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        fps, best, cb, reps = cpu_forward_fps(CPU_SAMPLE_B)
        cpu = {"value": fps, "unit": "mel-frames/s", "cores": host_threads(), "kind": "port",
               "sample": "oracle port (torch CPU fp32, position table cached), B=%d utterances of the same workload "
                         "(L=%d, T=%d), best of %d (%.2f s each)" % (cb, L, T, reps, best)}
    print(json.dumps({
        "metric": "mel-frames/sec (batched non-AR forward)", "value": value, "unit": "mel-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: full STYLER forward, B=%d/GPU, phoneme_len=%d -> mel_len=%d "
                               "(teacher-forced %d frames/phoneme), ref mel %d frames, 80-bin, %s compute; random-init weights"
                               % (B_PER_GPU, L, T, FRAMES, T, args.precision),
                   "global_batch": world * B_PER_GPU, "parallelism": "dp%d" % world,
                   "launch": LAUNCH_DESC if use_graph else "eager kernel launches",
                   "gather": ("none (N=1)" if world == 1 else
                              ("fused compute+gather: mel_linear / last PostNet conv store into rank 0's IPC-mapped buffer over NVLink "
                               "(%d bytes/rank/step), flag protocol; inside the timed events" if peer else
                               "DMA push of the packed results (%d bytes/rank/step) into rank 0's IPC-mapped buffer on a side stream + flag "
                               "protocol; inside the timed events" if push else
                               "one packed NCCL gather per step (4 fp32 mels + lengths, %d bytes/rank), inside the timed events")
                              % packed_nbytes(B_PER_GPU, T)),
                   "l2": "inputs rotate over %d resident batches (> L2); per-step activations >> L2; no explicit flush" % NBUF},
        "e2e": {"value": e2e_value, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_secs / args.steps,
                "what": "pinned host inputs -> H2D -> forward -> D2H of all four mel tensors + lengths, every step" +
                        (" [DIAGNOSTIC RUN %s: not an e2e number]" % diag if diag else "")},
        "gpu_launches": launches, "launches_per_step": launches_per_step, "clocks": ck, "roofline": roof, "cpu_baseline": cpu,
        "wall_s_timed_region": wall, "host_enqueue_ms_per_step": 1e3 * enqueue[0], "cpu_binding": numa, "extras": extras or None}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
