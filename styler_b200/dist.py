"""Data-parallel plumbing: one process per GPU (torchrun), utterances sharded across ranks with no data-path
collective; the only exchange is the final gather of the mel tensors (and lengths) to rank 0 over NCCL/NVLink,
replacing nn.DataParallel's gather (train.py:33, synthesize.py:62 of the reference).  Every utterance is independent
in eval mode; to stay bitwise identical to a single-GPU run all ranks must pad to the same
max_src_len / max_mel_len, which the caller passes explicitly (both are forward() arguments)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items utterances for `rank` (remainder spread over the first ranks)."""
    q, r = divmod(n_items, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_batch(batch, rank, world):
    """Slice every per-utterance tensor of a forward() kwargs/args dict along dim 0."""
    n = batch["src_seq"].shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def gather_to_rank0(tensors, dst=0):
    """Gather same-shaped per-rank tensors to `dst` (concatenated along dim 0 there, None elsewhere).
    All ranks must hold equal shapes (equal shard sizes and the global padded T)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(tensors)
    world, rank = dist.get_world_size(), dist.get_rank()
    outs = []
    for t in tensors:
        t = t.contiguous()
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        outs.append(torch.cat(bufs, dim=0) if rank == dst else None)
    return outs


def gather_packed(packed, dst=0, bufs=None):
    """One collective for a step's results: `packed` is the byte buffer of engine.packed_views (four fp32 mel tensors + int64
    lengths back to back).  Returns the list of per-rank buffers on `dst` (engine.unpack_results reads them), None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [packed]
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst and bufs is None:
        bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.gather(packed, bufs if rank == dst else None, dst=dst)
    return bufs if rank == dst else None


class AsyncGather:
    """Gather of the per-rank mel tensors to rank 0 on a dedicated communication stream, so the NVLink transfer of
    step i overlaps the compute of step i+1 (the transfer is 84 MB per rank per step at config 3).  Receive buffers
    on rank 0 are allocated once and reused.  `wait()` joins the communication stream into the current stream."""

    def __init__(self, device, dst=0):
        self.device, self.dst = device, dst
        self.stream = torch.cuda.Stream(device=device)
        self._bufs = {}
        self._done = {}
        self.result = None

    def launch_packed(self, packed, slot=0):
        """ONE gather per step: `packed` is Engine.last_packed (uint8: the four fp32 mel tensors and the int64 lengths back
        to back, engine.packed_views).  `slot` names the (reused) buffer `packed` lives in -- e.g. the static outputs of CUDA
        graph `slot`: call `before_reuse(slot)` before that buffer is written again.  On rank 0 `result` becomes the list
        of per-rank byte buffers (engine.unpack_results turns each back into tensors)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            self.result = [packed]
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            packed.record_stream(self.stream)
            bufs = None
            if rank == self.dst:
                key = ("packed", slot, packed.numel())
                if key not in self._bufs:
                    self._bufs[key] = [torch.empty_like(packed) for _ in range(world)]
                bufs = self._bufs[key]
            gather_packed(packed, self.dst, bufs)
            ev = self._done.get(slot)
            if ev is None:
                ev = self._done[slot] = torch.cuda.Event()
            ev.record(self.stream)
        self.result = bufs

    def before_reuse(self, slot=0):
        """Make the current stream wait for the last gather that read buffer `slot`."""
        ev = self._done.get(slot)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def launch(self, tensors):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            self.result = list(tensors)
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)
        outs = []
        with torch.cuda.stream(self.stream):
            for i, t in enumerate(tensors):
                t = t.contiguous()
                t.record_stream(self.stream)
                bufs = None
                if rank == self.dst:
                    key = (i, tuple(t.shape), t.dtype)
                    if key not in self._bufs:
                        self._bufs[key] = [torch.empty_like(t) for _ in range(world)]
                    bufs = self._bufs[key]
                dist.gather(t, bufs, dst=self.dst)
                outs.append(bufs)
        self.result = outs

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.result
