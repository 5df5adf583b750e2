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


class _RawDeviceMemory:
    """A device allocation that did not come from torch (styler_peer_alloc / styler_peer_open) exposed through the CUDA array
    interface, so `torch.as_tensor` can wrap it without copying."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = int(ptr), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}


class PeerGather:
    """Fused compute + gather over NVLink peer memory (SURVEY.md 8(e) phase 2), the replacement for nn.DataParallel's gather
    (train.py:33, synthesize.py:62 of the reference) when every GPU has its own process.

    Rank 0 owns a receive region [world][slots][packed_nbytes]; every other rank maps it (CUDA IPC) and hands ITS slice to the
    engine as `result_mirror`: mel_linear and the last PostNet convolution then store their fp32 results straight into rank
    0's memory from their epilogues -- the mel tensors cross NVLink while the tensor-core kernel that computes them is still
    running, there is no staging copy and no collective.  Completion and buffer reuse are two counters per (rank, slot):
      ready[r][slot]  (in rank 0's memory, written by rank r after its forward, system-scope release)
      ack[slot]       (in rank r's memory, written by rank 0 once it has consumed that slot)
    `slots` = 2 lets step i+1 run while rank 0 still reads step i.

    Per step, on every rank:      buf = pg.begin(slot)    # waits (on the stream) until rank 0 has released the slot
                                  ... forward with engine.result_mirror = pg.buffer(slot) ...
                                  pg.commit(slot)         # publishes ready[rank][slot]
    and on rank 0 additionally:   bufs = pg.collect(slot) # stream-waits for every rank; list of per-rank byte buffers
                                  ... consume ...         # engine.unpack_results(bufs[r], B, T)
                                  pg.release(slot)        # lets the ranks overwrite the slot
    """

    def __init__(self, device, nbytes, slots=2, dst=0):
        import ctypes
        from . import _lib
        assert dist.is_initialized(), "PeerGather needs an initialised process group (handle exchange)"
        self._lib, self._ct = _lib, ctypes
        self.device, self.nbytes, self.slots, self.dst = torch.device(device), int(nbytes), int(slots), dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.stride = (self.nbytes + 255) // 256 * 256
        self.uses = [0] * self.slots
        self._own, self._mapped, self._keep = [], [], []
        with torch.cuda.device(self.device):
            # ack flags live on every rank (rank 0 writes them remotely); the big region and the ready flags on rank 0
            ack_ptr, ack_h = self._alloc(256)
            if self.rank == dst:
                recv_ptr, recv_h = self._alloc(self.world * self.slots * self.stride)
                ready_ptr, ready_h = self._alloc(256 * self.world)          # ready[r][slot] at byte 256 r + 8 slot
            else:
                recv_h = ready_h = None
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (recv_h, ready_h, ack_h))
            if self.rank == dst:
                self.recv_ptr, self.ready_ptr = recv_ptr, ready_ptr
                self.ack_remote = [None if r == dst else self._open(gathered[r][2]) for r in range(self.world)]
            else:
                self.recv_ptr, self.ready_ptr = self._open(gathered[dst][0]), self._open(gathered[dst][1])
            self.ack_ptr = ack_ptr
            dist.barrier()                                                   # everybody has mapped everything

    # -- raw memory helpers --------------------------------------------------------------------------------------------
    def _alloc(self, nbytes):
        ct = self._ct
        p, h = ct.c_void_p(), (ct.c_ubyte * 64)()
        self._lib.check(self._lib.lib().styler_peer_alloc(int(nbytes), ct.byref(p), h), "peer_alloc")
        self._own.append(p.value)
        return p.value, bytes(h)

    def _open(self, handle):
        ct = self._ct
        p = ct.c_void_p()
        buf = (ct.c_ubyte * 64).from_buffer_copy(handle)
        self._lib.check(self._lib.lib().styler_peer_open(buf, ct.byref(p)), "peer_open")
        self._mapped.append(p.value)
        return p.value

    def _tensor(self, ptr, nbytes):
        raw = _RawDeviceMemory(ptr, nbytes)
        self._keep.append(raw)
        return torch.as_tensor(raw, device=self.device)

    def _stream(self):
        return self._ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def slice_ptr(self, rank, slot):
        return self.recv_ptr + (rank * self.slots + slot) * self.stride

    # -- protocol ----------------------------------------------------------------------------------------------------------
    def buffer(self, slot, rank=None):
        """uint8 view [nbytes] of (rank, slot) of the receive region (peer memory on every rank but dst)."""
        return self._tensor(self.slice_ptr(self.rank if rank is None else rank, slot), self.nbytes)

    @property
    def remote(self):
        return self.rank != self.dst

    def begin(self, slot):
        """Before the forward that writes slot `slot`: wait (stream order) until rank 0 has consumed its previous contents."""
        self.uses[slot] += 1
        if self.remote and self.uses[slot] > 1:
            with torch.cuda.device(self.device):
                self._lib.check(self._lib.lib().styler_peer_wait(self._ct.c_void_p(self.ack_ptr + 8 * slot), 1, 1,
                                                               self.uses[slot] - 1, self._stream()), "peer_wait(ack)")

    def commit(self, slot):
        """After the forward: publish `ready` for this rank and slot (everything enqueued before it on the stream is visible)."""
        if self.remote:
            with torch.cuda.device(self.device):
                self._lib.check(self._lib.lib().styler_peer_signal(self._ct.c_void_p(self.ready_ptr + 256 * self.rank + 8 * slot),
                                                                 self.uses[slot], self._stream()), "peer_signal(ready)")

    def collect(self, slot):
        """Rank 0: stream-wait until every other rank has committed this use of `slot`; returns the per-rank byte buffers."""
        assert not self.remote
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r != self.dst:
                    self._lib.check(self._lib.lib().styler_peer_wait(self._ct.c_void_p(self.ready_ptr + 256 * r + 8 * slot), 1, 1,
                                                                   self.uses[slot], self._stream()), "peer_wait(ready)")
        return [self.buffer(slot, r) for r in range(self.world)]

    def release(self, slot):
        """Rank 0: the consumer is done with `slot` (stream order): let the other ranks overwrite it."""
        assert not self.remote
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r != self.dst:
                    self._lib.check(self._lib.lib().styler_peer_signal(self._ct.c_void_p(self.ack_remote[r] + 8 * slot),
                                                                     self.uses[slot], self._stream()), "peer_signal(ack)")

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier()
        for p in self._mapped:
            self._lib.lib().styler_peer_close(self._ct.c_void_p(p))
        self._mapped = []
        if dist.is_initialized():
            dist.barrier()
        for p in self._own:
            self._lib.lib().styler_peer_free(self._ct.c_void_p(p))
        self._own = []


class AsyncPeerGather:
    """Drives PeerGather with the interface of AsyncGather (bench.py / serving loops): the ranks' forwards write their packed
    results directly into rank 0's receive region; rank 0 collects and releases on a side stream so that its own next
    forward is not held up by the slowest rank.

        buf = g.begin(slot)            # every rank, before the forward (stream-waits for the slot to be free)
        ... forward with engine.result_mirror = g.buffer(slot) (or a CUDA graph captured with it) ...
        g.launch_packed(None, slot)    # every rank, after the forward: commit (+ rank 0: collect + release on the side stream)
        g.wait()                       # join the side stream (end of a timed region / before reading `result`)
    """

    def __init__(self, device, nbytes, slots=2, dst=0, push=False):
        """push=False: FUSED -- the producing kernels store into the peer region (engine.result_mirror = buffer(slot)).
        push=True: the forward writes its local packed buffer only; `launch_packed(packed, slot)` then pushes it into the peer
        region on the side stream with one DMA copy (copy engine over NVLink, no SM time) and publishes the flag behind it, so
        the transfer overlaps the next step.  (Measured at N = 8: the fused stores of all ranks converge on rank 0's NVLink
        ingress at the END of every step -- 588 MB, ~0.8 ms, on the critical path; pushed or gathered asynchronously the
        same bytes hide under the next step's 7 ms of compute.)"""
        self.pg = PeerGather(device, nbytes, slots, dst)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.push = bool(push)
        self._done = {}
        self._bufs = {}
        self.result = None

    def buffer(self, slot):
        return self.pg.buffer(slot)

    @property
    def remote(self):
        return self.pg.remote

    def begin(self, slot):
        self.before_reuse(slot)
        if self.push:            # the ack wait belongs in front of the COPY (side stream), not in front of the forward
            with torch.cuda.stream(self.stream):
                self.pg.begin(slot)
        else:
            self.pg.begin(slot)

    def launch_packed(self, packed, slot=0):
        main = torch.cuda.current_stream(self.device)
        if self.push:
            # side stream: [wait for the forward] -> DMA copy local packed -> this rank's slice of rank 0's region -> flag
            ready = torch.cuda.Event()
            ready.record(main)
            self.stream.wait_event(ready)
            with torch.cuda.stream(self.stream):
                packed.record_stream(self.stream)
                if slot not in self._bufs:
                    self._bufs[slot] = self.pg.buffer(slot)
                self._bufs[slot].copy_(packed, non_blocking=True)
                self.pg.commit(slot)
                if not self.pg.remote:
                    self.result = self.pg.collect(slot)
                    self.pg.release(slot)
                ev = self._done.get(slot)
                if ev is None:
                    ev = self._done[slot] = torch.cuda.Event()
                ev.record(self.stream)
            return
        self.pg.commit(slot)
        if not self.pg.remote:
            ready = torch.cuda.Event()
            ready.record(main)
            self.stream.wait_event(ready)
            with torch.cuda.stream(self.stream):
                self.result = self.pg.collect(slot)
                self.pg.release(slot)
                ev = self._done.get(slot)
                if ev is None:
                    ev = self._done[slot] = torch.cuda.Event()
                ev.record(self.stream)

    def before_reuse(self, slot=0):
        ev = self._done.get(slot)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.result

    def close(self):
        self.pg.close()
