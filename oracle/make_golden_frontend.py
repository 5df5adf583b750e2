"""Pins oracle/frontend_oracle.py against the UNMODIFIED reference functions (audio/tools.py:get_mel_from_wav,
utils.py:f0_normalization / energy_rescaling; imported from /root/reference in this container only, with the
non-arithmetic stubs of oracle/ref_shim.py) and writes tests/golden/frontend_b3.pt.
  python -m oracle.make_golden_frontend"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_oracle as fo  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    ref_shim.load_reference_tacotron_stft()          # installs the librosa stub + CPU .cuda() before audio.tools is imported
    _, _, _, _, ref_utils = ref_shim.load_reference_modules()
    import audio.tools as ref_tools                  # the reference module, unmodified
    wav, n_samples, f0, frames = fo.make_case(seed=0)
    gold = {"seed": 0, "mel": [], "energy": [], "e_input": [], "f0_norm": [], "clipt": []}
    worst = 0.0
    for b in range(wav.shape[0]):
        n, fr = int(n_samples[b]), int(frames[b])
        for norm in (True, False):
            x = wav[b, :n] if norm else wav[b, :n] / 16384.0          # norm=False input overshoots [-1,1] -> clamp path
            m_ref, e_ref, c_ref = ref_tools.get_mel_from_wav(x.clone(), norm=norm)
            m, e, c = fo.get_mel_from_wav(x.clone(), norm=norm)
            worst = max(worst, (m - m_ref).abs().max().item(), ((e - e_ref).abs() / e_ref.abs().clamp_min(1e-6)).max().item())
            assert c == c_ref, (b, norm, c, c_ref)
            if norm:
                gold["mel"].append(m_ref.clone()); gold["energy"].append(e_ref.clone())
                er = ref_utils.energy_rescaling(e_ref.numpy().astype(np.float32))
                assert np.array_equal(er, fo.energy_rescaling(e.numpy().astype(np.float32))) or np.allclose(er, fo.energy_rescaling(e.numpy().astype(np.float32)), atol=1e-6)
                gold["e_input"].append(torch.from_numpy(er))
            else:
                gold["clipt"].append(bool(c_ref))
        f = f0[b, :fr].numpy().astype(np.float32)
        fn_ref = ref_utils.f0_normalization(f.copy())
        fn = fo.f0_normalization(f.copy())
        assert np.allclose(fn_ref, fn, atol=0, rtol=0, equal_nan=True), b
        gold["f0_norm"].append(torch.from_numpy(np.asarray(fn_ref, dtype=np.float64)))
    print("oracle vs reference: worst mel abs / energy rel error %.2e; clipt %s; all-unvoiced row -> zeros: %s"
          % (worst, gold["clipt"], bool((gold["f0_norm"][2] == 0).all())))
    assert worst < 2e-5
    out = os.path.join(ROOT, "tests", "golden", "frontend_b3.pt")
    torch.save(gold, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
