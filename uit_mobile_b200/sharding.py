"""Batch sharding of the hot path across the GPUs of one box (one process per GPU, torch.distributed).

The path is embarrassingly clip-parallel: rank r takes a contiguous slice of the clips, weights are
replicated.  Two tiny exchanges keep the result identical to a single-GPU call:
  1. all-reduce(MAX) of ONE 32-bit word -- the bit pattern of the largest mel power -- so that the reference's
     batch-global top-dB cutoff (Q2) spans the global batch (done inside ``UITBase.forward`` when the model has a
     ``process_group``);
  2. all-gather of the ``[B/G, outputdim]`` scores (this module).
Backend: NCCL over NVLink on the GPUs; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``total`` clips owned by ``rank``; the first ranks get the remainder.

    ``align`` > 1 makes every boundary (except the end) a multiple of ``align`` clips: with ``align =
    model.tile_clips(T)`` every clip keeps the position inside its 128-row encoder tile that it has in a single-GPU
    run, which makes the sharded scores bit-identical to the single-GPU ones."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    units = -(-total // align)                       # ceil: the last unit may be ragged
    base, rem = divmod(units, world)
    ub = rank * base + min(rank, rem)
    ue = ub + base + (1 if rank < rem else 0)
    return min(ub * align, total), min(ue * align, total)


def window_shard_bounds(n_windows: int, hop: int, win: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """Sliding-window streaming (BASELINE config 5): rank owns windows [w0, w1) and needs the samples
    [w0*hop, (w1-1)*hop + win) of the stream, i.e. its time range plus a (win - hop) halo."""
    w0, w1 = shard_bounds(n_windows, rank, world)
    s0 = w0 * hop
    s1 = (w1 - 1) * hop + win if w1 > w0 else s0
    return w0, w1, s0, s1


def allreduce_max_word(word: torch.Tensor, group=None) -> torch.Tensor:
    """In-place MAX of the int32 max-power word.  Non-negative floats order like their bit patterns."""
    if word.dtype != torch.int32:
        raise TypeError("the max-power word is the int32 bit pattern of a non-negative float")
    dist.all_reduce(word, op=dist.ReduceOp.MAX, group=group)
    return word


def gather_scores(local: torch.Tensor, total: int, group=None, align: int = 1) -> torch.Tensor:
    """All-gather per-rank score slices ``[n_r, C]`` (n_r from ``shard_bounds``) into ``[total, C]`` on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, r, world, align)[1] - shard_bounds(total, r, world, align)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    C = local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty((total, C), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, C), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


def sharded_forward(model, wav_local: torch.Tensor, total: int, group=None, gather: bool = True, align: int = 1) -> torch.Tensor:
    """Run the model on this rank's slice and (optionally) gather all scores.  ``model.process_group`` is set so
    that the top-dB scope is the global batch.  ``align`` must be the value used for ``shard_bounds``."""
    model.process_group = group if group is not None else dist.group.WORLD
    local = model(wav_local)
    return gather_scores(local, total, group, align) if gather else local


class PeerWords:
    """The batch-maximum word of every rank, exchanged through NVLink peer memory instead of ``all_reduce(MAX)``.

    Every rank owns a 64-byte slot in a symmetric-memory allocation (mapped by all peers once, at construction); per step the
    rank PUBLISHES its word there (one 1-thread kernel: word, then an epoch counter with release semantics) and later COLLECTS
    the maximum over all ranks (one 1-warp kernel that reads the peers' slots with system-scope loads, waiting for their epoch).
    No NCCL kernel, no collective launch on the host: an NCCL kernel needs a CTA slot the persistent encoder does not leave
    until its last tile wave, and costs ~40 us of host time per enqueue.

        model.peer_words = PeerWords(group, device)      # next to model.process_group; UITBase._finish uses it when present

    Every rank must run the same sequence of steps (publish and collect are collective in the same sense all_reduce is)."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = int(N.lib().uitk_peer_words_slot_bytes()) // 4
        self.slot = symm.empty((n,), dtype=torch.int32, device=self.device)
        self.slot.zero_()
        self.hdl = symm.rendezvous(self.slot, group)
        self.peers = [self.hdl.get_buffer(r, (n,), torch.int32) for r in range(self.world)]
        self.ptrs = torch.tensor([p.data_ptr() for p in self.peers], dtype=torch.int64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group)                      # every slot is zeroed before anyone publishes or polls
        self.epoch = 0

    def publish(self, word: torch.Tensor) -> int:
        """Publish this rank's int32 word (device tensor [1]) on the current stream; returns the epoch to collect with."""
        from . import _native as N
        self.epoch += 1
        if self.epoch >= 1 << 32:
            self.epoch = 1
        N.check(N.lib().uitk_peer_words_publish(self.slot.data_ptr(), word.data_ptr(), self.epoch,
                                                torch.cuda.current_stream(self.device).cuda_stream), "uitk_peer_words_publish")
        return self.epoch

    def collect(self, epoch: int, out: torch.Tensor) -> torch.Tensor:
        """MAX of all ranks' words of ``epoch`` into ``out`` (int32 device tensor [1]) on the current stream."""
        from . import _native as N
        N.check(N.lib().uitk_peer_words_collect(self.ptrs.data_ptr(), self.world, epoch, out.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream), "uitk_peer_words_collect")
        return out


class PeerGather:
    """All-gather of the per-rank score blocks through NVLink peer memory, moved by the COPY ENGINES.

    NCCL's all-gather runs CTAs on the SMs and was measured to slow the log-mel kernel of the next step by 9 % at 8 GPUs; here
    every rank publishes its block in a symmetric-memory buffer (torch.distributed._symmetric_memory: CUDA IPC / fabric handles
    exchanged once) and PULLS the peers' blocks with plain device-to-device copies on a side stream - no SM is taken from the
    kernels of the following batch.  Two tiny barrier kernels (signal pads in the same symmetric allocation) frame the copies:
    "all blocks published" and "all ranks have read" (the slot may be overwritten again).

        g = PeerGather(sizes, cols, group, device)       # sizes[r] = rows of rank r (sharding.shard_bounds)
        out, done = g(local)                              # out [sum(sizes), cols]; `done` is recorded when it is complete
        torch.cuda.current_stream().wait_event(done)      # out stays valid until `depth` more gathers

    Raises if the group cannot map peer memory (no NVLink / IPC): callers fall back to ``gather_scores`` (NCCL)."""

    def __init__(self, sizes: List[int], cols: int, group=None, device=None, depth: int = 2):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if len(sizes) != self.world:
            raise ValueError("one size per rank")
        self.sizes, self.cols, self.depth = list(sizes), cols, depth
        self.offsets = [sum(sizes[:r]) for r in range(self.world)]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        cap = max(max(sizes), 1)
        self.src = symm.empty((depth, cap, cols), dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.src, group)
        self.peer_src = [self.hdl.get_buffer(r, (depth, cap, cols), torch.float32) for r in range(self.world)]
        self.out = [torch.empty((sum(sizes), cols), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.stream = torch.cuda.Stream(self.device)
        self.n = 0

    @torch.no_grad()
    def __call__(self, local: torch.Tensor):
        if tuple(local.shape) != (self.sizes[self.rank], self.cols) or local.dtype != torch.float32:
            raise ValueError(f"rank {self.rank} must pass a float32 [{self.sizes[self.rank]}, {self.cols}] block")
        slot = self.n % self.depth
        self.n += 1
        self.stream.wait_stream(torch.cuda.current_stream(self.device))       # `local` was produced on the caller's stream
        local.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            self.src[slot, : local.shape[0]].copy_(local, non_blocking=True)
            self.hdl.barrier(channel=0)                                       # every rank's block is published
            out = self.out[slot]
            for step in range(self.world):
                r = (self.rank - step) % self.world                           # every rank starts at a different peer
                n = self.sizes[r]
                if n:
                    out[self.offsets[r]: self.offsets[r] + n].copy_(self.peer_src[r][slot, :n], non_blocking=True)
            self.hdl.barrier(channel=1)                                       # every rank has read: the slot may be reused
            done = torch.cuda.Event()
            done.record(self.stream)
        return out, done
