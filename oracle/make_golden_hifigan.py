"""Pins oracle/hifigan_oracle.py against the UNMODIFIED reference generator (hifigan/models.py, imported from
/root/reference in this container only) and writes tests/golden/hifigan_b2_t24.pt.
  python -m oracle.make_golden_hifigan
The reference module runs with weight_norm active and the seeded checkpoint-form weights (weight_g / weight_v); the oracle
folds them (fold_weight_norm) and must agree to 1e-5; a second pass after the reference's own remove_weight_norm() must
give the same waveform."""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as ho  # noqa: E402


def main():
    sys.path.insert(0, "/root/reference")
    import hifigan  # the reference package: torch only
    h = hifigan.AttrDict(ho.CONFIG_V1)
    torch.manual_seed(0)
    ref = hifigan.Generator(h).eval()
    sd_wn = ho.make_state_dict(seed=7, weight_norm=True)
    missing = ref.load_state_dict(sd_wn, strict=True)
    B, T = 2, 24
    mel = ho.make_mel(B, T, seed=7)
    with torch.no_grad():
        y_ref = ref(mel)
        y_orc = ho.generator_forward(sd_wn, mel)
        with contextlib.redirect_stdout(io.StringIO()):
            ref.remove_weight_norm()
        y_ref2 = ref(mel)
        sd_plain = {k: v.clone() for k, v in ref.state_dict().items()}
        y_orc2 = ho.generator_forward(sd_plain, mel)
    e1 = (y_ref - y_orc).abs().max().item()
    e2 = (y_ref2 - y_ref).abs().max().item()
    e3 = (y_ref2 - y_orc2).abs().max().item()
    print("reference vs oracle (weight-norm form) %.2e | remove_weight_norm drift %.2e | plain form %.2e | wav absmax %.3f std %.3f"
          % (e1, e2, e3, y_ref.abs().max().item(), y_ref.std().item()))
    assert y_ref.shape == (B, 1, T * 256)
    assert e1 < 1e-5 and e2 < 1e-5 and e3 < 1e-5
    assert 0.05 < y_ref.std().item() < 0.9, "golden waveform saturated or vanishing: retune make_state_dict"
    out = os.path.join(ROOT, "tests", "golden", "hifigan_b2_t24.pt")
    torch.save({"seed": 7, "B": B, "T": T, "wav": y_ref.clone(), "n_state_tensors": len(sd_wn)}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
