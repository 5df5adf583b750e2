"""SURVEY.md 8(f) rank 3 -- the inspection / controllability API as synthesize.py drives it (synthesize.py:114-144,170-205):
after a forward the caller reads the encodings left on `style_modeling`, pushes some of them through the model's own
sub-modules (`pitch_linear`, `style_encoder.speaker_linear_p`, `style_encoder.speaker_linear`), swaps encodings between two
references, then `predict_inference` -> `decode`.  Here those sub-modules are kernel-backed (`model.KernelMLP`); the tests
replay `get_encodings` / `infer_comb` on the drop-in and compare every step with the CPU oracle.  Also: `use_postnet=False`
(styler.py:16-37), the control factors, and the one-process-many-GPUs use of nn.DataParallel (train.py:33)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import make_golden as mg
from oracle import styler_oracle as so

pytestmark = pytest.mark.gpu
P = "style_modeling."
SE = "style_modeling.style_encoder."


def rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _model(sd, precision, cuda, **kw):
    from styler_b200 import STYLER
    m = STYLER(precision=precision, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(cuda).eval()


def _run(model, batch, cuda, **extra):
    args, kw = mg.call_kwargs(batch)
    kw.update(extra)
    out = model(*[a.to(cuda) for a in args], **{k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in kw.items()})
    torch.cuda.synchronize()
    return mg.flatten_outputs(out)


def _oracle_encodings(sd, batch):
    """What modules.py:328-348 leaves on `self` after a forward, from the oracle (fp32 CPU)."""
    src_mask = so.mask_from_lengths(batch["src_len"], batch["max_src_len"])
    text = so.text_encoder(sd, SE + "text_encoder.", batch["src_seq"], src_mask)
    neck = F.relu(F.linear(F.relu(F.linear(text, sd[SE + "text_linear_down.0.weight"], sd[SE + "text_linear_down.0.bias"])),
                           sd[P + "text_linear_up.0.weight"], sd[P + "text_linear_up.0.bias"]))
    cat = so.encoder_input_cat(batch["mel_target"], batch["p_norm"], batch["e_input"], batch["mel_aug"])
    d_enc, p_enc, e_enc, n_enc = so.audio_encoder(sd, SE + "audio_encoder.", cat, batch["mel_len"], batch["src_len"])
    L = batch["max_src_len"]
    spk = F.relu(F.linear(batch["speaker_embed"], sd[SE + "speaker_linear.0.weight"], sd[SE + "speaker_linear.0.bias"]))
    spk_p = F.relu(F.linear(batch["speaker_embed"], sd[SE + "speaker_linear_p.0.weight"], sd[SE + "speaker_linear_p.0.bias"]))
    return dict(t=text, t_neck=neck, p_down=p_enc, s_down=spk_p.unsqueeze(1).repeat(1, L, 1), s=spk.unsqueeze(1).repeat(1, L, 1),
                d=so._mlp2(sd, P + "duration_linear.", d_enc), e=so._mlp2(sd, P + "energy_linear.", e_enc),
                n=so._mlp2(sd, P + "residual_linear.", n_enc), src_mask=src_mask)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tf32", 5e-3), ("bf16", 4e-2)])
def test_kernel_backed_submodules(cuda, precision, tol):
    """pitch_linear / speaker_linear / speaker_linear_p / ... called directly (synthesize.py:118-119,194-196) run the
    library's GEMM kernels and agree with the reference arithmetic; keys and shapes of their state_dict are unchanged."""
    from styler_b200 import _lib
    sd = so.make_state_dict(0)
    model = _model(sd, precision, cuda)
    sm = model.style_modeling
    g = torch.Generator().manual_seed(3)
    x128 = torch.randn(2, 19, 128, generator=g)
    spk = torch.randn(3, 512, generator=g)
    cases = [(sm.pitch_linear, x128, lambda v: so._mlp2(sd, P + "pitch_linear.", v)),
             (sm.energy_linear, x128, lambda v: so._mlp2(sd, P + "energy_linear.", v)),
             (sm.residual_linear, x128, lambda v: so._mlp2(sd, P + "residual_linear.", v)),
             (sm.pitch_norm_linear, x128, lambda v: so._mlp2(sd, P + "pitch_norm_linear.", v)),
             (sm.duration_linear, torch.randn(2, 19, 160, generator=g), lambda v: so._mlp2(sd, P + "duration_linear.", v)),
             (sm.style_encoder.speaker_linear, spk,
              lambda v: F.relu(F.linear(v, sd[SE + "speaker_linear.0.weight"], sd[SE + "speaker_linear.0.bias"]))),
             (sm.style_encoder.speaker_linear_p, spk,
              lambda v: F.relu(F.linear(v, sd[SE + "speaker_linear_p.0.weight"], sd[SE + "speaker_linear_p.0.bias"]))),
             (sm.style_encoder.text_linear_down, torch.randn(2, 19, 256, generator=g),
              lambda v: F.relu(F.linear(v, sd[SE + "text_linear_down.0.weight"], sd[SE + "text_linear_down.0.bias"]))),
             (sm.text_linear_up, torch.rand(2, 19, 4, generator=g),
              lambda v: F.relu(F.linear(v, sd[P + "text_linear_up.0.weight"], sd[P + "text_linear_up.0.bias"])))]
    for mod, x, ref_fn in cases:
        n0 = _lib.launch_count()
        got = mod(x.to(cuda))
        assert _lib.launch_count() > n0, "sub-module must run library kernels, not torch's"
        ref = ref_fn(x)
        assert got.dtype == torch.float32 and got.shape == ref.shape
        assert rel(got, ref) < tol, (mod._key, rel(got, ref))


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tf32", 5e-3)])
def test_infer_comb_replay(cuda, precision, tol):
    """Replay of synthesize.py:114-144 (`get_encodings`) and :180-205 (`infer_comb` -> `infer`): text / duration / energy /
    noise encodings of reference A, pitch re-projected with the TARGET speaker B's down-projection, predict_inference with
    speaker_normalized=False, decode of t + p + s + e.  Every intermediate is compared with the oracle."""
    sd = so.make_state_dict(0)
    so.set_duration_bias(sd, 3)
    model = _model(sd, precision, cuda)
    sm = model.style_modeling
    bA = so.make_inputs(B=1, L=30, Tr=80, seed=61, d_mode=None)
    bB = so.make_inputs(B=1, L=30, Tr=64, seed=62, d_mode=None)

    def get_encodings(batch):                      # synthesize.py:114-144 on the drop-in
        _run(model, batch, cuda)
        p_down, s_down = sm.pitch_encoding, sm.speaker_encoding_p
        return dict(max_mel_len=sm.max_len, src_mask=sm.src_mask, t=sm.text_encoding, t_neck=sm.text_encoding_neck, p_down=p_down,
                    s_down=s_down, p_norm=sm.pitch_linear(p_down), p=sm.pitch_linear(p_down + s_down), d=sm.duration_encoding,
                    s=sm.speaker_encoding, e=sm.energy_encoding, n=sm.noise_encoding)

    encA = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in get_encodings(bA).items()}
    get_encodings(bB)
    refA = _oracle_encodings(sd, bA)
    for k in ("t", "t_neck", "p_down", "s_down", "d", "s", "e", "n"):
        assert rel(encA[k], refA[k]) < tol, (k, rel(encA[k], refA[k]))
    assert rel(encA["p"], so._mlp2(sd, P + "pitch_linear.", refA["p_down"] + refA["s_down"])) < tol
    assert rel(encA["p_norm"], so._mlp2(sd, P + "pitch_linear.", refA["p_down"])) < tol

    # infer_comb (synthesize.py:180-205): target speaker = B's embedding
    L = bA["max_src_len"]
    spk_B = bB["speaker_embed"]
    s_down_tgt = sm.style_encoder.speaker_linear_p(spk_B.to(cuda)).unsqueeze(1).repeat(1, L, 1)
    s_tgt = sm.style_encoder.speaker_linear(spk_B.to(cuda)).unsqueeze(1).repeat(1, L, 1)
    p_tgt = sm.pitch_linear(encA["p_down"] + s_down_tgt)
    sm.speaker_encoding = s_tgt
    t, t_neck, d, e, s, n = (encA[k] for k in ("t", "t_neck", "d", "e", "s", "n"))
    got = sm.predict_inference(t, t_neck + p_tgt, t_neck + e, t_neck + d, s, n, encA["src_mask"], encA["max_mel_len"], False)
    gt, gp, gs, ge, gn, _, f0_out, en_out, mel_mask = got
    mel, mel_post = model.decode(gt.float() + gp.float() + gs.float() + ge.float(), mel_mask)
    torch.cuda.synchronize()

    r_s_down_tgt = F.relu(F.linear(spk_B, sd[SE + "speaker_linear_p.0.weight"], sd[SE + "speaker_linear_p.0.bias"])).unsqueeze(1).repeat(1, L, 1)
    r_p_tgt = so._mlp2(sd, P + "pitch_linear.", refA["p_down"] + r_s_down_tgt)
    assert rel(p_tgt, r_p_tgt) < tol
    with torch.no_grad():
        ref = so.predict_inference(sd, refA["t"], refA["t_neck"] + r_p_tgt, refA["t_neck"] + refA["e"], refA["t_neck"] + refA["d"],
                                   refA["s"], refA["n"], refA["src_mask"], None, False)
        rt, rp, rs, re_, rn, _, rf0, ren, rmask = ref
        r_mel, r_post = so.decode(sd, rt + rp + rs + re_, rmask)
    assert torch.equal(mel_mask.cpu(), rmask)
    for a, b_ in ((gt, rt), (gs, rs), (gn, rn), (f0_out, rf0), (en_out, ren)):
        assert rel(a, b_) < tol
    assert rel(gp, rp) < 1e-6 and rel(ge, re_) < 1e-6, "embedding rows are exact copies when the bucket indices agree"
    assert rel(mel, r_mel) < tol and rel(mel_post, r_post) < tol


def test_predict_inference_controls(cuda):
    """d/p/e_control of predict_inference (modules.py:290-305): predictions come back SCALED, the duration control stretches
    the utterance, and the caller's input tensors are left untouched."""
    sd = so.make_state_dict(0)
    so.set_duration_bias(sd, 4)
    model = _model(sd, "fp32", cuda)
    sm = model.style_modeling
    b = so.make_inputs(B=2, L=16, Tr=40, seed=71, ragged=True, d_mode=None)
    _run(model, b, cuda)
    refe = _oracle_encodings(sd, b)
    p_up = so._mlp2(sd, P + "pitch_linear.", refe["p_down"] + refe["s_down"])
    a = (refe["t"], refe["t_neck"] + p_up, refe["t_neck"] + refe["e"], refe["t_neck"] + refe["d"], refe["s"], refe["n"], refe["src_mask"], None)
    dev_a = tuple(x.to(cuda) if torch.is_tensor(x) else x for x in a)
    keep = [x.clone() if torch.is_tensor(x) else x for x in dev_a]
    with torch.no_grad():
        ref = so.predict_inference(sd, *a, speaker_normalized=False, d_control=1.5, p_control=1.2, e_control=0.8)
    got = sm.predict_inference(*dev_a, speaker_normalized=False, d_control=1.5, p_control=1.2, e_control=0.8)
    assert torch.equal(got[8].cpu(), ref[8]) and got[8].shape[1] == 6 * 16
    for i in (0, 2, 4, 5, 6, 7):
        assert rel(got[i], ref[i]) < 2e-4, i
    for x, y in zip(dev_a, keep):
        if torch.is_tensor(x):
            assert torch.equal(x, y)


def test_forward_controls_return_scaled_predictions(cuda):
    """Free-running forward with p_control / e_control != 1 (modules.py:370,380): p/e predictions in the 9-tuple are the scaled ones."""
    sd = so.make_state_dict(0)
    so.set_duration_bias(sd, 3)
    model = _model(sd, "fp32", cuda)
    b = so.make_inputs(B=2, L=12, Tr=30, seed=72, d_mode=None)
    args, kw = mg.call_kwargs(b)
    with torch.no_grad():
        ref = mg.flatten_outputs(so.styler_forward(sd, *args, **dict(kw, p_control=1.3, e_control=0.6, d_control=1.0)))
    got = _run(model, b, cuda, p_control=1.3, e_control=0.6)
    got1 = _run(model, b, cuda)
    assert torch.equal(got["mel_len"].cpu(), ref["mel_len"])
    for k in ("p_pred", "e_pred", "mel", "mel_postnet"):
        assert rel(got[k], ref[k]) < 2e-4, (k, rel(got[k], ref[k]))
    assert rel(got["p_pred"], got1["p_pred"] * 1.3) < 1e-6


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_use_postnet_false(cuda, precision, tol):
    """styler.py:16-37: without the PostNet there is no `postnet` sub-module (293 state_dict keys) and decode() returns the
    mel twice."""
    sd = {k: v for k, v in so.make_state_dict(0).items() if not k.startswith("postnet.")}
    model = _model(sd, precision, cuda, use_postnet=False)
    assert not hasattr(model, "postnet") and len(model.state_dict()) == len(sd)
    b = so.make_inputs(B=2, L=16, seed=73, ragged=True, d_mode="ragged")
    got = _run(model, b, cuda)
    args, kw = mg.call_kwargs(b)
    full = so.make_state_dict(0)
    with torch.no_grad():
        ref = mg.flatten_outputs(so.styler_forward(full, *args, **kw))
    assert rel(got["mel"], ref["mel"]) < tol and rel(got["mel_noisy"], ref["mel_noisy"]) < tol
    assert torch.equal(got["mel_postnet"], got["mel"]) and torch.equal(got["mel_postnet_noisy"], got["mel_noisy"])


def test_model_on_non_current_device_and_data_parallel(cuda):
    """One process, several GPUs (nn.DataParallel: train.py:33, synthesize.py:62): kernels must launch on the device the
    tensors live on, the >48 KB shared-memory opt-in must exist per device, and replicas must use their own device's engine."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    sd = so.make_state_dict(0)
    # L = 32: a shard of 2 utterances still has B*L >= 64 rows, so shards and the full batch take the same (tensor-core) kernels
    # and the comparison can be bitwise
    b = so.make_inputs(B=4, L=32, seed=74, d_mode="const", frames=4)
    m0 = _model(sd, "bf16", torch.device("cuda:0"))
    ref = _run(m0, b, torch.device("cuda:0"))
    m1 = _model(sd, "bf16", torch.device("cuda:1"))
    assert torch.cuda.current_device() == 0
    out1 = _run(m1, b, torch.device("cuda:1"))          # cuda:0 is current, the model lives on cuda:1
    for k in ("mel", "mel_postnet_noisy", "p_pred"):
        assert out1[k].device.index == 1 and torch.equal(out1[k].cpu(), ref[k].cpu()), k
    dp = torch.nn.DataParallel(m0, device_ids=[0, 1])
    args, kw = mg.call_kwargs(b)
    out = dp(*[a.to("cuda:0") for a in args], **{k: (v.to("cuda:0") if torch.is_tensor(v) else v) for k, v in kw.items()})
    torch.cuda.synchronize()
    got = mg.flatten_outputs(out)
    for k in ("mel", "mel_postnet", "mel_postnet_noisy", "log_d"):
        assert torch.equal(got[k].cpu(), ref[k].cpu()), k
