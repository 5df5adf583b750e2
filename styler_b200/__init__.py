"""styler_b200 -- B200-native (sm_100a) implementation of the STYLER non-autoregressive mel-synthesis forward
and the TacotronSTFT mel front end, behind the reference's Python surface.  See DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401

__all__ = ["STYLER", "GraphedSTYLER", "PipelinedSTYLER", "TacotronSTFT", "Generator", "ReferenceFrontEnd", "STYLERLoss", "DomainAdversarialTrainingLoss", "ops", "hparams"]


def __getattr__(name):
    if name == "STYLER":
        from .model import STYLER
        return STYLER
    if name == "GraphedSTYLER":
        from .model import GraphedSTYLER
        return GraphedSTYLER
    if name == "PipelinedSTYLER":
        from .model import PipelinedSTYLER
        return PipelinedSTYLER
    if name == "TacotronSTFT":
        from .stft import TacotronSTFT
        return TacotronSTFT
    if name == "Generator":                      # HiFi-GAN vocoder (hifigan.Generator drop-in)
        from .vocoder import Generator
        return Generator
    if name == "ReferenceFrontEnd":
        from .frontend import ReferenceFrontEnd
        return ReferenceFrontEnd
    if name in ("STYLERLoss", "DomainAdversarialTrainingLoss"):    # loss.py drop-ins (forward values; evaluate.py:88-104)
        from . import loss as _loss
        return getattr(_loss, name)
    if name in ("ops", "hparams", "model", "stft", "engine", "dist", "vocoder", "frontend", "loss"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
