"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the reference's loss.py.

`styler_loss` follows STYLERLoss.forward (loss.py:26-50) and cal_mel_loss (loss.py:16-24); `dat_loss` follows
DomainAdversarialTrainingLoss.forward (loss.py:60-66).  Masks are True = KEEP, as the callers pass them
(evaluate.py:88-90: `~src_mask`, `~mel_mask`).  Pinned against the unmodified reference by oracle/make_golden_loss.py."""
import torch
import torch.nn.functional as F


def cal_mel_loss(mel, mel_postnet, mel_target, mel_keep):
    """loss.py:16-24."""
    sel = mel_keep.unsqueeze(-1)
    t = mel_target.masked_select(sel)
    return F.mse_loss(mel.masked_select(sel), t), F.mse_loss(mel_postnet.masked_select(sel), t)


def styler_loss(log_d_pred, log_d_target, p_pred, p_target, e_pred, e_target, mel, mel_postnet, mel_target, src_keep, mel_keep,
                aug_posteriors, aug_label):
    """loss.py:26-50 -> (mel, mel_postnet, d, p, e, classifier) losses."""
    mel_loss, post_loss = cal_mel_loss(mel, mel_postnet, mel_target, mel_keep)
    d_loss = F.l1_loss(log_d_pred.masked_select(src_keep), log_d_target.masked_select(src_keep))
    p_loss = F.l1_loss(p_pred.masked_select(mel_keep), p_target.masked_select(mel_keep))
    e_loss = F.l1_loss(e_pred.masked_select(mel_keep), e_target.masked_select(mel_keep))
    return mel_loss, post_loss, d_loss, p_loss, e_loss, dat_loss(aug_posteriors, aug_label)


def dat_loss(aug_posteriors, aug_label):
    """loss.py:60-66 (and :45-47): sum of three NLL means over log-probabilities [B,2]."""
    return sum(F.nll_loss(p, aug_label) for p in aug_posteriors)


def make_case(seed=0, B=3, L=11, T=37, n_mel=80):
    """Seeded inputs with ragged lengths (one utterance full length)."""
    g = torch.Generator().manual_seed(seed)
    src_len = torch.tensor([L] + [int(x) for x in torch.randint(3, L, (B - 1,), generator=g)])
    mel_len = torch.tensor([T] + [int(x) for x in torch.randint(5, T, (B - 1,), generator=g)])
    src_keep = torch.arange(L).unsqueeze(0) < src_len.unsqueeze(1)
    mel_keep = torch.arange(T).unsqueeze(0) < mel_len.unsqueeze(1)
    r = lambda *s: torch.randn(*s, generator=g)
    post = [F.log_softmax(r(B, 2), -1) for _ in range(3)]
    return dict(log_d_pred=r(B, L), log_d_target=r(B, L).abs(), p_pred=r(B, T) * 100, p_target=r(B, T) * 100 + 200, e_pred=r(B, T) * 30,
                e_target=r(B, T).abs() * 50, mel=r(B, T, n_mel), mel_postnet=r(B, T, n_mel), mel_target=r(B, T, n_mel) - 3.0,
                src_keep=src_keep, mel_keep=mel_keep, src_len=src_len, mel_len=mel_len, post=post,
                label=torch.randint(0, 2, (B,), generator=g))
