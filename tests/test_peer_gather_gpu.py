"""2-GPU test of the fused compute + gather (dist.PeerGather): every rank's forward stores its mels straight into rank 0's
IPC-mapped receive region from the epilogues of mel_linear / the last PostNet convolution; rank 0 must see, bitwise, what
an NCCL gather of the ranks' local results delivers -- over several steps and both slots, eagerly and through CUDA graphs
(the NVLink stores are then part of the captured kernels).  Needs >= 2 GPUs in one box (skipped otherwise); run by
`gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
import torch.distributed as dist
from styler_b200 import STYLER, GraphedSTYLER
from styler_b200 import dist as sdist, synthetic as syn
from styler_b200.engine import packed_nbytes, unpack_results
rank, world, local = sdist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
B, L, F = 3, 24, 4
T = L * F
model = STYLER(precision="bf16"); model.load_state_dict(syn.make_state_dict(0)); model = model.to(dev).eval()
def batch(seed):
    b = syn.make_inputs(B=B, L=L, seed=seed, d_mode="const", frames=F)
    a = tuple(b[k].to(dev) for k in ("src_seq", "mel_target", "mel_aug", "p_norm", "e_input", "src_len", "mel_len"))
    kw = dict(d_target=b["d_target"].to(dev), p_target=b["p_target"].to(dev), e_target=b["e_target"].to(dev), max_src_len=L,
              max_mel_len=T, speaker_embed=b["speaker_embed"].to(dev))
    return a, kw
nb = packed_nbytes(B, T)
g = sdist.AsyncPeerGather(dev, nb, slots=2)
eng = model._engine_for()
a0, k0 = batch(100 + rank)
graphs = [GraphedSTYLER(model, a0, k0, warmup=1, result_mirror=g.buffer(k)) for k in range(2)]
ok = True
for step in range(7):
    slot = step %% 2
    a, kw = batch(1000 * step + rank)
    g.begin(slot)
    if step < 3:                                   # eager forwards with the mirror set by hand, then graph replays
        eng.result_mirror = g.buffer(slot)
        model(*a, **kw)
        eng.result_mirror = None
        local_packed = eng.last_packed.clone()
    else:
        graphs[slot](*a, **kw)
        local_packed = graphs[slot].packed.clone()
    g.launch_packed(None, slot)
    # reference: plain NCCL gather of the ranks' LOCAL packed results
    bufs = [torch.empty_like(local_packed) for _ in range(world)] if rank == 0 else None
    dist.gather(local_packed, bufs, dst=0)
    if rank == 0:
        got = g.wait()
        torch.cuda.synchronize()
        for r in range(world):
            same = torch.equal(got[r], bufs[r])
            mel, mel_n, post, post_n, ln = unpack_results(got[r], B, T)
            ok = ok and same and bool(torch.isfinite(post).all()) and ln.tolist() == [T] * B
            if not same:
                print("MISMATCH step", step, "rank", r, (got[r] != bufs[r]).sum().item())
    dist.barrier()
g.close()
# push mode: the forward writes only its local packed buffer; a DMA copy on the side stream carries it into rank 0's region
g2 = sdist.AsyncPeerGather(dev, nb, slots=2, push=True)
for step in range(5):
    slot = step %% 2
    a, kw = batch(5000 * step + rank)
    g2.begin(slot)
    model(*a, **kw)
    local_packed = eng.last_packed
    g2.launch_packed(local_packed, slot)
    bufs = [torch.empty_like(local_packed) for _ in range(world)] if rank == 0 else None
    dist.gather(local_packed, bufs, dst=0)
    if rank == 0:
        got = g2.wait()
        torch.cuda.synchronize()
        for r in range(world):
            if not torch.equal(got[r], bufs[r]):
                ok = False
                print("PUSH MISMATCH step", step, "rank", r)
    dist.barrier()
g2.close()
if rank == 0:
    print("PEER_GATHER_OK" if ok else "PEER_GATHER_BAD")
dist.destroy_process_group()
"""


def test_peer_gather_matches_nccl_gather(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one box")
    script = tmp_path / "peer_worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, TORCHELASTIC_ERROR_FILE=str(tmp_path / "err.json"))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=600, env=env)
    tail = (out.stdout + out.stderr)[-4000:]
    assert out.returncode == 0, tail
    assert "PEER_GATHER_OK" in out.stdout, tail
