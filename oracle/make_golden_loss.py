"""Pins oracle/loss_oracle.py against the UNMODIFIED reference loss.py (imported from /root/reference in this container only)
and writes tests/golden/loss_b3.pt.
  python -m oracle.make_golden_loss"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle as lo  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(ref_shim.REFERENCE_DIR, "loss.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                       # the reference module, unmodified
    gold = {"seeds": [0, 1], "values": [], "noisy": [], "dat": []}
    for seed in gold["seeds"]:
        c = lo.make_case(seed)
        L, D = ref.STYLERLoss(), ref.DomainAdversarialTrainingLoss()
        with torch.no_grad():
            r = L(c["log_d_pred"], c["log_d_target"].clone(), c["p_pred"], c["p_target"].clone(), c["e_pred"], c["e_target"].clone(),
                  c["mel"], c["mel_postnet"], c["mel_target"].clone(), c["src_keep"], c["mel_keep"], c["src_len"], c["mel_len"],
                  tuple(c["post"]), c["label"].clone().float().long())
            rn = L.cal_mel_loss(c["mel_postnet"], c["mel"], c["mel_target"].clone(), c["mel_keep"])
            rd = D(tuple(c["post"]), 1 - c["label"])
            o = lo.styler_loss(c["log_d_pred"], c["log_d_target"], c["p_pred"], c["p_target"], c["e_pred"], c["e_target"], c["mel"],
                               c["mel_postnet"], c["mel_target"], c["src_keep"], c["mel_keep"], c["post"], c["label"])
            on = lo.cal_mel_loss(c["mel_postnet"], c["mel"], c["mel_target"], c["mel_keep"])
            od = lo.dat_loss(c["post"], 1 - c["label"])
        for a, b in list(zip(r, o)) + list(zip(rn, on)) + [(rd, od)]:
            assert torch.equal(a, b), (seed, float(a), float(b))        # the same torch ops on the same data: bitwise
        gold["values"].append(torch.stack(list(r)))
        gold["noisy"].append(torch.stack(list(rn)))
        gold["dat"].append(rd.clone())
    out = os.path.join(ROOT, "tests", "golden", "loss_b3.pt")
    torch.save(gold, out)
    print("oracle == reference (bitwise) on seeds", gold["seeds"], "->", out, os.path.getsize(out), "bytes")
    print(gold["values"][0])


if __name__ == "__main__":
    main()
