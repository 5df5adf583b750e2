"""GPU parity of the loss drop-ins (styler_b200.loss, csrc/loss.cu) through the C ABI: against the reference-generated golden
values, against the oracle at the bench shape, and the evaluate.py:81-104 call pattern on a real forward."""
import os

import pytest
import torch

from oracle import loss_oracle as lo
from oracle import make_golden as mg
from oracle import styler_oracle as so

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_b3.pt")


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)


def _call(L, c, dev):
    g = lambda k: c[k].to(dev)
    return L(g("log_d_pred"), g("log_d_target"), g("p_pred"), g("p_target"), g("e_pred"), g("e_target"), g("mel"), g("mel_postnet"),
             g("mel_target"), g("src_keep"), g("mel_keep"), g("src_len"), g("mel_len"), tuple(p.to(dev) for p in c["post"]), g("label"))


def test_loss_vs_reference_golden(cuda):
    from styler_b200 import STYLERLoss, DomainAdversarialTrainingLoss
    gold = torch.load(GOLD)
    L, D = STYLERLoss(), DomainAdversarialTrainingLoss()
    for i, seed in enumerate(gold["seeds"]):
        c = lo.make_case(seed)
        got = _call(L, c, cuda)
        for j in range(6):
            assert got[j].dim() == 0 and _rel(got[j], gold["values"][i][j]) < 2e-6, (seed, j, float(got[j]), float(gold["values"][i][j]))
        n = L.cal_mel_loss(c["mel_postnet"].to(cuda), c["mel"].to(cuda), c["mel_target"].to(cuda), c["mel_keep"].to(cuda))
        assert _rel(n[0], gold["noisy"][i][0]) < 2e-6 and _rel(n[1], gold["noisy"][i][1]) < 2e-6
        d = D(tuple(p.to(cuda) for p in c["post"]), (1 - c["label"]).to(cuda))
        assert _rel(d, gold["dat"][i]) < 2e-6
    again = _call(L, lo.make_case(1), cuda)
    assert all(torch.equal(a, b) for a, b in zip(again, got)), "two launches must agree bitwise (fixed reduction order)"


def test_loss_bench_shape_and_edge_masks(cuda):
    """B = 64 x T = 1024 (the bench geometry) against the oracle, plus edge masks: one utterance with a single kept frame, and an
    entirely masked duration row."""
    from styler_b200 import STYLERLoss
    c = lo.make_case(seed=5, B=64, L=128, T=1024)
    c["mel_keep"][3] = False
    c["mel_keep"][3, 0] = True
    c["src_keep"][7] = False
    ref = lo.styler_loss(c["log_d_pred"], c["log_d_target"], c["p_pred"], c["p_target"], c["e_pred"], c["e_target"], c["mel"],
                         c["mel_postnet"], c["mel_target"], c["src_keep"], c["mel_keep"], c["post"], c["label"])
    got = _call(STYLERLoss(), c, cuda)
    for j in range(6):
        assert _rel(got[j], ref[j]) < 1e-5, (j, float(got[j]), float(ref[j]))


def test_loss_on_a_forward_like_evaluate(cuda):
    """evaluate.py:81-104: forward -> STYLERLoss on the clean outputs, cal_mel_loss on the noisy ones, DAT loss on the posteriors;
    fp32 mode against the oracle forward + oracle loss."""
    from styler_b200 import STYLER, STYLERLoss, DomainAdversarialTrainingLoss
    sd = so.make_state_dict(0)
    batch = so.make_inputs(B=3, L=20, seed=71, ragged=True, d_mode="ragged")
    model = STYLER(precision="fp32")
    model.load_state_dict(sd)
    model = model.to(cuda).eval()
    args, kw = mg.call_kwargs(batch)
    out = model(*[a.to(cuda) for a in args], **{k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in kw.items()})
    (mel, mel_n), (post, post_n), log_d, p_pred, e_pred, src_mask, mel_mask, _, aug = out
    log_D = torch.log(batch["d_target"].float() + 1.0)
    L, D = STYLERLoss(), DomainAdversarialTrainingLoss()
    zeros = torch.zeros(3, dtype=torch.long, device=cuda)
    got = L(log_d, log_D.to(cuda), p_pred, batch["p_target"].to(cuda), e_pred, batch["e_target"].to(cuda), mel, post,
            batch["mel_target"].to(cuda), ~src_mask, ~mel_mask, batch["src_len"].to(cuda), batch["mel_len"].to(cuda), aug, zeros)
    got_n = L.cal_mel_loss(mel_n, post_n, batch["mel_aug"].to(cuda), ~mel_mask)
    got_d = D(aug, torch.ones(3, dtype=torch.long, device=cuda))
    with torch.no_grad():
        r = mg.flatten_outputs(so.styler_forward(sd, *args, **kw))
        ref = lo.styler_loss(r["log_d"], log_D, r["p_pred"], batch["p_target"], r["e_pred"], batch["e_target"], r["mel"], r["mel_postnet"],
                             batch["mel_target"], ~r["src_mask"], ~r["mel_mask"], [r["aug_d"], r["aug_p"], r["aug_e"]], torch.zeros(3, dtype=torch.long))
        ref_n = lo.cal_mel_loss(r["mel_noisy"], r["mel_postnet_noisy"], batch["mel_aug"], ~r["mel_mask"])
        ref_d = lo.dat_loss([r["aug_d"], r["aug_p"], r["aug_e"]], torch.ones(3, dtype=torch.long))
    for j in range(6):
        assert _rel(got[j], ref[j]) < 2e-4, (j, float(got[j]), float(ref[j]))
    assert _rel(got_n[0], ref_n[0]) < 2e-4 and _rel(got_n[1], ref_n[1]) < 2e-4 and _rel(got_d, ref_d) < 2e-4
