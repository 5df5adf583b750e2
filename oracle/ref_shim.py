"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference (dev container only).

The GPU box has no /root/reference; nothing that runs there may import this module.  It is used by
``oracle/make_golden.py`` (fixture generation) and by the dev-container-only tests that compare the
oracle restatement to the live reference (they skip when the directory is absent).

Only non-arithmetic third-party modules are stubbed (matplotlib, unidecode, inflect, and -- for the
STFT path -- librosa, whose ``filters.mel`` is replaced by the Slaney restatement in stft_oracle.py and
whose hard-coded ``.cuda()`` hop in audio/stft.py:66-67 is neutralised).  No reference file is edited
or copied.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("STYLER_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "styler.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_done = {"model": False, "audio": False}


def _prepare_model_imports():
    if _done["model"]:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib", use=lambda *a, **k: None)
            mpl.pyplot = _stub("matplotlib.pyplot")
    for name, attrs in (("unidecode", dict(unidecode=lambda s: s)), ("inflect", dict(engine=lambda: None))):
        try:
            __import__(name)
        except Exception:
            _stub(name, **attrs)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    _done["model"] = True


def load_reference_styler():
    """Return the reference ``styler.STYLER`` class (unmodified)."""
    _prepare_model_imports()
    import styler as ref_styler  # noqa: E402  (reference module)
    return ref_styler.STYLER


def load_reference_modules():
    """Return (styler, modules, transformer.Models/Layers, utils) reference modules."""
    _prepare_model_imports()
    import styler as ref_styler
    import modules as ref_modules
    import utils as ref_utils
    import transformer.Models as ref_models
    import transformer.Layers as ref_layers
    return ref_styler, ref_modules, ref_models, ref_layers, ref_utils


def load_reference_tacotron_stft():
    """Return the reference ``audio.stft.TacotronSTFT`` class running on CPU."""
    _prepare_model_imports()
    if not _done["audio"]:
        import numpy as np
        import torch
        from . import stft_oracle

        def pad_center(data, size, axis=-1, **kw):
            n = data.shape[axis]
            lpad = int((size - n) // 2)
            lengths = [(0, 0)] * data.ndim
            lengths[axis] = (lpad, int(size - n - lpad))
            return np.pad(data, lengths, mode="constant")

        def tiny(x):
            x = np.asarray(x)
            dt = x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32
            return np.finfo(dt).tiny

        def normalize(S, **kw):
            return S

        util = _stub("librosa.util", pad_center=pad_center, tiny=tiny, normalize=normalize)
        filters = _stub("librosa.filters",
                        mel=lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw:
                        stft_oracle.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax))
        lib = _stub("librosa", util=util, filters=filters)
        lib.util, lib.filters = util, filters
        torch.Tensor.cuda = lambda self, *a, **k: self  # audio/stft.py:66-67 hard-codes .cuda()
        _done["audio"] = True
    import audio.stft as ref_stft
    return ref_stft.TacotronSTFT
