"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED reference.

Run in the dev container (needs /root/reference):   python -m oracle.make_golden
For every case the reference `styler.STYLER` (eval, no_grad, CPU fp32) is loaded (strict) with the
seeded `make_state_dict` weights and fed the seeded `make_inputs` batch; its 9-tuple output (plus a
few StyleModeling intermediates the reference itself stores on `self`, modules.py:328-348) is saved as
float32/int64 tensors.  Fixtures hold only OUTPUTS (+ a weight checksum); inputs/weights are
regenerated from seeds at test time, which also pins the generators across machines.
"""
import hashlib
import os
import sys

import torch

from . import ref_shim, styler_oracle as so, stft_oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (weights seed, make_inputs kwargs, duration-bias frames for free-running or None)
CASES = {
    "tf_const_b2_l16": (0, dict(B=2, L=16, seed=11, d_mode="const", frames=4), None),
    "tf_ragged_b3_l24": (0, dict(B=3, L=24, seed=12, ragged=True, d_mode="ragged"), None),
    "tf_ragged_b2_l40_long": (1, dict(B=2, L=40, seed=13, ragged=True, d_mode="ragged"), None),
    "free_b2_l12_tr50": (0, dict(B=2, L=12, Tr=50, seed=14, ragged=False, d_mode=None), 5),
    "free_ragged_b3_l20_tr90": (1, dict(B=3, L=20, Tr=90, seed=15, ragged=True, d_mode=None), 3),
    "free_single_l50_tr400": (0, dict(B=1, L=50, Tr=400, seed=16, d_mode=None), 8),   # BASELINE configs[0]
    # BASELINE configs[2] geometry (the bench shape): L=128 -> T=1024 > hp.max_seq_len, so the reference rebuilds the
    # sinusoid table on the fly (transformer/Models.py:120-122); ragged text and reference-mel lengths
    "tf_headline_b2_l128_t1024": (0, dict(B=2, L=128, seed=17, ragged=True, d_mode="const", frames=8), None),
}


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def call_kwargs(batch):
    kw = dict(d_target=batch.get("d_target"), p_target=batch.get("p_target"), e_target=batch.get("e_target"),
              max_src_len=batch["max_src_len"], max_mel_len=batch.get("max_mel_len"),
              speaker_embed=batch["speaker_embed"])
    args = (batch["src_seq"], batch["mel_target"], batch["mel_aug"], batch["p_norm"], batch["e_input"],
            batch["src_len"], batch["mel_len"])
    return args, kw


def flatten_outputs(out):
    (mel, mel_n), (post, post_n), log_d, p_pred, e_pred, src_mask, mel_mask, mel_len, (pd, pp, pe) = out
    return dict(mel=mel, mel_noisy=mel_n, mel_postnet=post, mel_postnet_noisy=post_n, log_d=log_d,
                p_pred=p_pred, e_pred=e_pred, src_mask=src_mask, mel_mask=mel_mask, mel_len=mel_len,
                aug_d=pd, aug_p=pp, aug_e=pe)


def build_case(name):
    wseed, in_kw, frames = CASES[name]
    sd = so.make_state_dict(wseed)
    if frames is not None:
        so.set_duration_bias(sd, frames)
    return sd, so.make_inputs(**in_kw)


def main(only=None):
    """only: optional list of case names to (re)generate; default all (+ the STFT golden)."""
    if not ref_shim.available():
        sys.exit("reference tree not available; goldens can only be generated in the dev container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    STYLER = ref_shim.load_reference_styler()
    torch.manual_seed(0)
    ref = STYLER().eval()
    ref_keys = sorted(ref.state_dict().keys())
    for name in CASES:
        if only and name not in only:
            continue
        sd, batch = build_case(name)
        assert sorted(sd.keys()) == ref_keys, "state_dict surface mismatch"
        ref.load_state_dict(sd, strict=True)
        args, kw = call_kwargs(batch)
        with torch.no_grad():
            out = ref(*args, **kw)
            flat = flatten_outputs(out)
            sm = ref.style_modeling
            flat["i_text_encoding"] = sm.text_encoding.clone()
            flat["i_duration_encoding"] = sm.duration_encoding.clone()
            flat["i_energy_encoding"] = sm.energy_encoding.clone()
            flat["i_noise_encoding"] = sm.noise_encoding.clone()
            flat["i_pitch_encoding_raw"] = sm.pitch_encoding.clone()
            flat["i_text_encoding_neck"] = sm.text_encoding_neck.clone()
            mine = flatten_outputs(so.styler_forward(sd, *args, **kw))
        for k, v in mine.items():
            r = flat[k]
            if r.dtype in (torch.bool, torch.int64):
                assert torch.equal(r, v), (name, k)
            else:
                err = (r - v).abs().max().item() / max(r.abs().max().item(), 1e-12)
                assert err < 2e-5, (name, k, err)
                print("  %-28s %-20s oracle-vs-reference rel err %.2e" % (name, k, err))
        flat = {k: v.clone() for k, v in flat.items()}
        flat["_weights_sha256"] = sd_checksum(sd)
        torch.save(flat, os.path.join(GOLDEN_DIR, name + ".pt"))
        print("wrote", name, {k: tuple(v.shape) for k, v in flat.items() if hasattr(v, "shape")})

    if only:
        return
    # ---- TacotronSTFT golden: reference conv-DFT path on CPU (audio/stft.py) -------------------
    Taco = ref_shim.load_reference_tacotron_stft()
    taco = Taco(1024, 256, 1024, 80, 22050, 0.0, 8000.0)
    g = torch.Generator().manual_seed(21)
    y = (torch.rand(3, 6000, generator=g) * 2 - 1) * 0.5
    with torch.no_grad():
        mel, energy = taco.mel_spectrogram(y)
    m2, e2 = stft_oracle.mel_spectrogram(y, dense=True)
    m3, e3 = stft_oracle.mel_spectrogram(y, dense=False)
    print("stft oracle(dense) vs reference: mel %.2e energy %.2e" % ((mel - m2).abs().max(), ((energy - e2).abs() / energy).max()))
    print("stft oracle(rfft)  vs reference: mel %.2e energy %.2e" % ((mel - m3).abs().max(), ((energy - e3).abs() / energy).max()))
    assert (mel - m2).abs().max() < 1e-4 and (mel - m3).abs().max() < 1e-4
    torch.save(dict(seed=21, shape=(3, 6000), mel=mel.clone(), energy=energy.clone(),
                    mel_basis_row0=taco.mel_basis[0, :8].clone(), mel_basis_sum=taco.mel_basis.sum().clone()),
               os.path.join(GOLDEN_DIR, "stft_b3_n6000.pt"))
    print("wrote stft_b3_n6000")


if __name__ == "__main__":
    main(sys.argv[1:] or None)
