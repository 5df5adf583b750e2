"""Child-process body of tests/test_forward_gpu.py::test_pipelined_batches_match_eager (run as `python tests/pipeline_case.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import make_golden as mg  # noqa: E402
from oracle import styler_oracle as so  # noqa: E402


def main():
    from styler_b200 import STYLER, PipelinedSTYLER
    cuda = torch.device("cuda:0")
    sd = so.make_state_dict(0)
    model = STYLER(precision="bf16")
    model.load_state_dict(sd)
    model = model.to(cuda).eval()
    batches = [so.make_inputs(B=3, L=48, seed=60 + i, d_mode="const", frames=6) for i in range(5)]
    to = lambda a, k: ([x.to(cuda) for x in a], {n: (v.to(cuda) if torch.is_tensor(v) else v) for n, v in k.items()})
    calls = [to(*mg.call_kwargs(dict(b, max_mel_len=48 * 6))) for b in batches]
    eager = []
    for a, k in calls:
        o = mg.flatten_outputs(model(*a, **k))
        eager.append({n: v.clone() for n, v in o.items()})
    pipe = PipelinedSTYLER(model, *calls[0])
    keys = ("mel", "mel_noisy", "mel_postnet", "mel_postnet_noisy", "log_d", "p_pred", "e_pred", "aug_d", "mel_len")
    got = []
    slot = 0
    for i, (a, k) in enumerate(calls):
        slot = pipe.submit(*a, **k)
        if i >= 1:                                   # read batch i-1 while batch i is in flight (its slot is not reused yet)
            pslot = (i - 1) % pipe.slots
            pipe.done(pslot).synchronize()
            o = mg.flatten_outputs(pipe.outputs(pslot))
            got.append({n: o[n].clone() for n in keys})
    pipe.done(slot).synchronize()
    o = mg.flatten_outputs(pipe.outputs(slot))
    got.append({n: o[n].clone() for n in keys})
    torch.cuda.synchronize()
    for i in range(len(calls)):
        for n in keys:
            assert torch.equal(got[i][n], eager[i][n]), (i, n)
    print("PIPELINE_CASE_OK")


if __name__ == "__main__":
    main()
