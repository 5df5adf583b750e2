"""CPU test of the loss row (forward values): the oracle restatement of loss.py against the golden values
produced by the unmodified reference (oracle/make_golden_loss.py)."""
import os

import torch

from oracle import loss_oracle as lo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_b3.pt")


def test_loss_oracle_matches_reference_golden():
    gold = torch.load(GOLD)
    for i, seed in enumerate(gold["seeds"]):
        c = lo.make_case(seed)
        o = lo.styler_loss(c["log_d_pred"], c["log_d_target"], c["p_pred"], c["p_target"], c["e_pred"], c["e_target"], c["mel"],
                           c["mel_postnet"], c["mel_target"], c["src_keep"], c["mel_keep"], c["post"], c["label"])
        assert torch.equal(torch.stack(list(o)), gold["values"][i])
        assert torch.equal(torch.stack(list(lo.cal_mel_loss(c["mel_postnet"], c["mel"], c["mel_target"], c["mel_keep"]))), gold["noisy"][i])
        assert torch.equal(lo.dat_loss(c["post"], 1 - c["label"]), gold["dat"][i])
