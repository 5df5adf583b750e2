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
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tensor=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="measured")
    except Exception:
        return dict(tensor=1400.0, hbm=6650.0, src="fallback")


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
    peak = peaks["tensor"] if args.precision == "bf16" else peaks["tensor"] / 2
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
        "roofline": {"bound": "tensor", "kernel": "whole forward (72 resblock convs + 4 transposed convs)", "achieved": fl / (ms * 1e-3) / 1e12,
                     "peak": peak, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak,
                     "peak_source": peaks["src"] + " bf16_tflops_sustained" + ("" if args.precision == "bf16" else " / 2 (tf32)"),
                     "flops_per_launch": fl, "traffic": None},
        "realtime_factor_22050Hz": wav.numel() / (ms * 1e-3) / 22050.0, "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the forward as one CUDA graph per resident batch (experiment)")
    ap.add_argument("--workload", default="styler", choices=["styler", "vocoder"],
                    help="styler = BASELINE.json's headline metric (default); vocoder = secondary HiFi-GAN line (single GPU)")
    args = ap.parse_args()
    if args.workload == "vocoder":
        args.steps = min(args.steps, 10)
        run_vocoder(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    from styler_b200 import dist as sdist
    rank, world, local = sdist.init_from_env("gloo" if args.impl == "reference" else None)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from styler_b200 import synthetic as so      # seeded synthetic weights/inputs (no oracle import on the GPU arm)
    from styler_b200 import STYLER, _lib
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
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

    gatherer = sdist.AsyncGather(dev) if world > 1 else None
    graphs = {}

    def step(bt):
        a = (bt["src_seq"], bt["mel_target"], bt["mel_aug"], bt["p_norm"], bt["e_input"], bt["src_len"], bt["mel_len"])
        kw = dict(d_target=bt["d_target"], p_target=bt["p_target"], e_target=bt["e_target"], max_src_len=L, max_mel_len=T,
                  speaker_embed=bt["speaker_embed"])
        if args.graph and id(bt) in graphs:
            out = graphs[id(bt)](*a, **kw)
        else:
            out = model(*a, **kw)
        if gatherer is not None:   # NCCL gather of the 4 mels + lengths to rank 0 on the comm stream (overlaps the next step)
            gatherer.launch([out[0][0], out[0][1], out[1][0], out[1][1], out[7]])
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
        step(resident[i % NBUF])
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while not clocks.rows and clocks.proc is not None and time.perf_counter() - t_w < 3.0:
        time.sleep(0.05)                              # first sample seen: the sampler is up
    if args.graph:
        from styler_b200 import GraphedSTYLER
        g0 = GraphedSTYLER(model, (resident[0]["src_seq"], resident[0]["mel_target"], resident[0]["mel_aug"], resident[0]["p_norm"],
                                   resident[0]["e_input"], resident[0]["src_len"], resident[0]["mel_len"]),
                           dict(d_target=resident[0]["d_target"], p_target=resident[0]["p_target"], e_target=resident[0]["e_target"],
                                max_src_len=L, max_mel_len=T, speaker_embed=resident[0]["speaker_embed"]))
        for bt in resident:
            graphs[id(bt)] = g0          # one graph; inputs are copied into its static buffers on every replay

    # ---- device-resident timed region (value) + per-launch events on the dominant kernel ----------------------
    # per-launch CUDA events around the dominant kernel (FFN Conv1d k9) are recorded inside the native FFT-block call
    _lib.check(_lib.lib().styler_debug_ffn1_timing(1, T), "ffn1_timing")
    launches0 = _lib.launch_count()
    clocks.mark()
    secs, wall = timed(lambda i: step(resident[i % NBUF]), args.steps)
    clocks.__exit__()
    launches = _lib.launch_count() - launches0
    import ctypes
    t_ms, t_n, t_b, t_t = ctypes.c_float(0), ctypes.c_int32(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(_lib.lib().styler_debug_ffn1_timing_read(ctypes.byref(t_ms), ctypes.byref(t_n), ctypes.byref(t_b), ctypes.byref(t_t)),
               "ffn1_timing_read")
    _lib.check(_lib.lib().styler_debug_ffn1_timing(0, 0), "ffn1_timing")
    torch.cuda.synchronize()
    value = world * frames_per_step * args.steps / secs

    dec = [(t_ms.value / t_n.value, int(t_b.value), int(t_t.value))] * int(t_n.value) if t_n.value > 0 else []   # decoder-level launches (T >= 1024)
    peaks = load_peaks()
    roof = None
    if dec:
        avg_ms = sum(d[0] for d in dec) / len(dec)
        flops = 2.0 * dec[0][1] * dec[0][2] * 1024 * 9 * 256
        ach = flops / (avg_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "conv1d_tc_kernel<bf16> FFN Conv1d k=9 256->1024 (+bias+ReLU)",
                "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": ach / peaks["tensor"],
                "peak_source": peaks["src"] + " bf16_tflops_sustained", "avg_launch_ms": avg_ms, "launches_timed": len(dec),
                "flops_per_launch": flops, "traffic": traffic}

    # ---- end to end through the public API with host buffers ---------------------------------------------------
    # Two user streams alternate (each with its own pinned result buffers): step i's H2D / D2H copies overlap step
    # i+-1's kernels, as a serving loop would drive the public API.  Every step still does its own H2D of all inputs and
    # D2H of its results inside the timed region.
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    d2h = [[torch.empty(B_PER_GPU, T, 80, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
    len_buf = [torch.empty(B_PER_GPU, dtype=torch.int64).pin_memory() for _ in range(2)]

    def e2e_step(i):
        k = i % 2
        with torch.cuda.stream(e2e_streams[k]):
            hb = host[i % NBUF]
            bt = {kk: v.to(dev, non_blocking=True) for kk, v in hb.items()}
            out = step(bt)
            d2h[k][0].copy_(out[1][0], non_blocking=True)
            d2h[k][1].copy_(out[1][1], non_blocking=True)
            len_buf[k].copy_(out[7], non_blocking=True)

    def e2e_timed(steps):
        barrier()
        for st in e2e_streams:
            st.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            e2e_step(i)
        for st in e2e_streams:
            st.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    for i in range(max(2 * NBUF, args.warmup)):     # every host batch through both streams once: allocator pools, NCCL
        e2e_step(i)                                 # channels and receive buffers of the e2e streams exist before timing
    e2e_secs = e2e_timed(args.steps)
    e2e_value = world * frames_per_step * args.steps / e2e_secs
    d2h_bytes = 2 * d2h[0][0].numel() * 4 + len_buf[0].numel() * 8

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
                   "l2": "inputs rotate over %d resident batches (> L2); per-step activations >> L2; no explicit flush" % NBUF},
        "e2e": {"value": e2e_value, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_secs / args.steps},
        "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roof, "cpu_baseline": cpu,
        "wall_s_timed_region": wall, "host_enqueue_ms_per_step": 1e3 * enqueue[0]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
