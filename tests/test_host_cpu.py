"""CPU tests (-m "not gpu") of the host side: the C-ABI library loads and exports every declared symbol, the drop-in
module keeps the reference's state_dict surface and signatures, the product path refuses to run without CUDA, and the
data-parallel sharding/gather logic works at world_size 2 over gloo."""
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from styler_b200 import _lib
    h = _lib.lib()
    header = open(os.path.join(ROOT, "include", "styler_b200.h")).read()
    declared = set(re.findall(r"\b(styler_[a-z0-9_]+)\s*\(", header))
    declared.discard("styler_conv1d_args")
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(h, name), name
    assert h.styler_version() >= 100
    assert _lib.launch_count() == 0          # nothing launched on a CPU-only box


def test_conv_args_struct_matches_header_order():
    from styler_b200 import _lib
    header = open(os.path.join(ROOT, "include", "styler_b200.h")).read()
    body = header[header.index("typedef struct {"):header.index("} styler_conv1d_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.replace("typedef struct {", "").strip()
        if not stmt:
            continue
        decl = stmt.split(",")
        names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", decl[0].strip())[0])
        names += [d.strip() for d in decl[1:]]
    assert names == [f[0] for f in _lib.Conv1dArgs._fields_], names


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of the two ABI structs as gcc lays them out from include/styler_b200.h == the ctypes mirrors."""
    import ctypes
    from styler_b200 import _lib
    structs = {"styler_conv1d_args": _lib.Conv1dArgs, "styler_fft_weights": _lib.FftWeights,
               "styler_predictor_weights": _lib.PredictorWeights, "styler_postnet_weights": _lib.PostnetWeights,
               "styler_decoder_weights": _lib.DecoderWeights}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "styler_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_dropin_surface():
    from oracle import styler_oracle as so
    from styler_b200 import STYLER
    m = STYLER()
    sd = so.make_state_dict(0)
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=False)   # DataParallel-prefixed checkpoints
    params = list(inspect.signature(m.forward).parameters)
    assert params == ["src_seq", "mel_target", "mel_aug", "p_norm", "e_input", "src_len", "mel_len", "d_target", "p_target",
                      "e_target", "max_src_len", "max_mel_len", "speaker_embed", "d_control", "p_control", "e_control"]
    assert list(inspect.signature(m.decode).parameters) == ["style_modeling_output", "mel_mask"]
    sm = m.style_modeling
    for attr in ("style_encoder", "augmentation_classifier_d", "pitch_linear", "predict_inference", "duration_predictor",
                 "pitch_predictor", "energy_predictor", "length_regulator"):
        assert hasattr(sm, attr)
    for attr in ("encoder_input_cat", "audio_encoder", "speaker_linear", "speaker_linear_p", "text_encoder"):
        assert hasattr(sm.style_encoder, attr)


def test_product_path_has_no_cpu_fallback():
    from styler_b200 import STYLER, TacotronSTFT
    m = STYLER().eval()
    z = torch.zeros(1, 4, dtype=torch.long)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(z, z, z, z, z, z, z)
    with pytest.raises(RuntimeError, match="CUDA"):
        TacotronSTFT().mel_spectrogram(torch.zeros(1, 4000))
    with pytest.raises(RuntimeError, match="eval"):
        STYLER().train()._engine_for()  # training mode is rejected before any device work on a CUDA box too
    import styler_b200.engine as engine
    src = open(engine.__file__).read() + open(os.path.join(ROOT, "styler_b200", "ops.py")).read()
    assert "oracle" not in src, "the product path must never import the oracle"


def test_shard_range():
    from styler_b200 import dist as sdist
    spans = [sdist.shard_range(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [sdist.shard_range(512, r, 8) for r in range(8)][3] == (192, 256)


WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
from styler_b200 import dist as sdist
rank, world, local = sdist.init_from_env("gloo")
assert world == 2
batch = {"src_seq": torch.arange(12).view(6, 2), "mel_len": torch.arange(6), "max_mel_len": 7}
mine = sdist.shard_batch(batch, rank, world)
assert mine["src_seq"].shape[0] == 3 and mine["max_mel_len"] == 7
mel = torch.full((3, 4, 80), float(rank))
outs = sdist.gather_to_rank0([mel, mine["mel_len"]])
if rank == 0:
    assert outs[0].shape == (6, 4, 80) and outs[0][:3].eq(0).all() and outs[0][3:].eq(1).all()
    assert outs[1].tolist() == [0, 1, 2, 3, 4, 5]
else:
    assert outs[0] is None
# one packed collective per step (engine.packed_views / dist.gather_packed): four mels + lengths in one byte buffer
from styler_b200 import engine
B, T = 3, 4
(buf, lens), mel2, post2 = engine.packed_views(B, T, "cpu")
assert buf.numel() == engine.packed_nbytes(B, T)
mel2.fill_(float(rank)); post2.fill_(float(rank) + 0.5); mel2[B:] += 10; post2[B:] += 10
lens.copy_(torch.arange(B) + 100 * rank)
got = sdist.gather_packed(buf)
if rank == 0:
    for r in range(2):
        mel, mel_n, post, post_n, ln = engine.unpack_results(got[r], B, T)
        assert mel.shape == (B, T, 80) and mel.eq(r).all() and mel_n.eq(r + 10).all()
        assert post.eq(r + 0.5).all() and post_n.eq(r + 10.5).all() and ln.tolist() == [100 * r, 100 * r + 1, 100 * r + 2]
    print("GATHER_OK")
else:
    assert got is None
"""


def test_gloo_world_size_2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm) emits the JSON contract line; tiny run."""
    env = dict(os.environ, STYLER_BENCH_CPU_SAMPLE_B="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["value"] == line["value"]
