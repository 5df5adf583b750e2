"""Drop-in for the reference's loss.py (STYLERLoss, DomainAdversarialTrainingLoss) as used by evaluate.py:88-104 and
train.py:139-153 -- forward values only (no autograd graph: this package implements the eval-mode path).

The reference materialises seven `masked_select` copies and runs nine reductions per call; here one library call
(`styler_loss_fwd`, csrc/loss.cu) reads every tensor once in place and writes the six scalars.  Masks follow the callers'
convention: True = KEEP (evaluate.py passes `~src_mask`, `~mel_mask`)."""
import torch
import torch.nn as nn

from . import _lib as L


def _f32(t, dev):
    return None if t is None else t.detach().to(dev, torch.float32).contiguous()


def _loss_call(dev, B, T, Ln, n_mel, mel=None, post=None, target=None, mel_keep=None, d_pred=None, d_tgt=None, src_keep=None,
               p_pred=None, p_tgt=None, e_pred=None, e_tgt=None, post_d=None, post_p=None, post_e=None, label=None):
    if dev.type != "cuda":
        raise RuntimeError("styler_b200.loss: expected CUDA tensors (the product path has no CPU implementation)")
    with torch.cuda.device(dev):
        ws = torch.empty(int(L.lib().styler_loss_workspace_bytes()), device=dev, dtype=torch.uint8)
        out = torch.empty(8, device=dev, dtype=torch.float32)
        keep = [t for t in (mel, post, target, mel_keep, d_pred, d_tgt, src_keep, p_pred, p_tgt, e_pred, e_tgt, post_d, post_p, post_e, label)]
        L.check(L.lib().styler_loss_fwd(*[L.ptr(t) for t in keep], B, T, Ln, n_mel, L.ptr(ws), ws.numel(), L.ptr(out), L.stream_ptr()),
                "loss")
    return out


def _keep(mask, dev):
    return mask.detach().to(dev).to(torch.uint8).contiguous()      # bool -> one byte per position, 1 = keep


class STYLERLoss(nn.Module):
    """loss.py:7-50.  forward(...) -> (mel_loss, mel_postnet_loss, d_loss, p_loss, e_loss, classifier_loss_a), 0-dim fp32 tensors."""

    def cal_mel_loss(self, mel, mel_postnet, mel_target, mel_mask):
        """loss.py:16-24 (also called on its own for the noisy decode, evaluate.py:92-93)."""
        dev = mel.device
        B, T, n_mel = mel.shape
        out = _loss_call(dev, B, T, 0, n_mel, mel=_f32(mel, dev), post=_f32(mel_postnet, dev), target=_f32(mel_target, dev),
                         mel_keep=_keep(mel_mask, dev))
        return out[0], out[1]

    def forward(self, log_d_predicted, log_d_target, p_predicted, p_target, e_predicted, e_target, mel, mel_postnet, mel_target,
                src_mask, mel_mask, src_len, mel_len, aug_posteriors, aug_label):
        dev = mel.device
        B, T, n_mel = mel.shape
        Ln = log_d_predicted.shape[1]
        pd, pp, pe = aug_posteriors
        out = _loss_call(dev, B, T, Ln, n_mel, mel=_f32(mel, dev), post=_f32(mel_postnet, dev), target=_f32(mel_target, dev),
                         mel_keep=_keep(mel_mask, dev), d_pred=_f32(log_d_predicted, dev), d_tgt=_f32(log_d_target, dev),
                         src_keep=_keep(src_mask, dev), p_pred=_f32(p_predicted, dev), p_tgt=_f32(p_target, dev),
                         e_pred=_f32(e_predicted, dev), e_tgt=_f32(e_target, dev), post_d=_f32(pd, dev), post_p=_f32(pp, dev),
                         post_e=_f32(pe, dev), label=aug_label.detach().to(dev, torch.int64).contiguous())
        return out[0], out[1], out[2], out[3], out[4], out[5]


class DomainAdversarialTrainingLoss(nn.Module):
    """loss.py:53-66: sum of the three NLL means of the augmentation posteriors against `aug_label`."""

    def forward(self, augmentation_posterior, aug_label):
        pd, pp, pe = augmentation_posterior
        dev = pd.device
        out = _loss_call(dev, pd.shape[0], 0, 0, 1, post_d=_f32(pd, dev), post_p=_f32(pp, dev), post_e=_f32(pe, dev),
                         label=aug_label.detach().to(dev, torch.int64).contiguous())
        return out[5]
