"""Host-buffer entry point of the hot path: pinned host waveforms in, host scores out.

This is what a caller sitting where the reference's DataLoader/`inference.py` sits uses when the audio is in host
memory (reference: evaluate.py:53-66 moves every batch host->device and reads the scores back).  The batch is cut
into chunks; the H2D copy of chunk i+1 (second stream) overlaps the log-mel AND encoder kernels of chunk i, and the
scores of every chunk travel back on a third stream under the following uploads.

The reference's top-dB cutoff is batch-global (Q2: max over the whole batch - 120 dB), which would force the encoder
to wait for the last chunk's front-end.  Instead every chunk is encoded *speculatively* with the running maximum of
the chunks seen so far, while the kernels also track the batch-wide MIN mel power.  A running cutoff is never above
the final one, so if `min_db >= final_max_db - 120` no value of any chunk was (or would have been) clamped and the
speculative result is exactly the reference's; otherwise (a >120 dB dynamic range inside one batch, e.g. digital
silence next to a loud clip) the encoder is simply re-run for the whole batch with the final maximum.  Both words
come back with the scores in the same D2H, so the check costs nothing.
"""
from __future__ import annotations

import math
import struct
from typing import Optional

import torch

from . import _native as N

_INF_BITS = 0x7F800000


def _bits_to_db(bits: int) -> float:
    p = struct.unpack("<f", struct.pack("<I", bits & 0xFFFFFFFF))[0]
    return 10.0 * math.log10(max(p, 1e-10))


class HostPipeline:
    """``pipe(wav_host)`` is the synchronous call (host scores back when it returns).  ``submit`` / ``result`` keep up to
    ``depth`` batches in flight: the uploads of batch i+1 queue right behind those of batch i, so the encoder tail, the score
    download and the host-side check of batch i run under the H2D of batch i+1 and the PCIe link never idles:

        t = pipe.submit(wav_host)    # pinned host [B, L]; must stay unchanged until result(t)
        ...                          # submit the next batch before asking for this one
        y = pipe.result(t)           # pinned host [B, outputdim]; valid until `depth` more batches were submitted
    """

    def __init__(self, model, max_batch: int, L: int = 16000, chunk: int = 1024, device: Optional[torch.device] = None,
                 speculative: bool = True, dtype: torch.dtype = torch.float32, depth: int = 2):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.UitkError("HostPipeline needs the model on a CUDA device (no CPU fallback)")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        T0 = int(N.lib().uitk_num_frames(L))
        tile = model.tile_clips(T0)                      # tile-aligned chunks keep the result bit-identical to one launch
        self.max_batch, self.L, self.depth = max_batch, L, depth
        self.chunk = max(tile, min(chunk, max_batch) // tile * tile)
        self.speculative = speculative
        T = int(N.lib().uitk_num_frames(L))
        dev = self.device
        if dtype not in (torch.float32, torch.int16):
            raise ValueError("dtype must be float32 (reference contract) or int16 (PCM ingest, x = pcm / 32768)")
        self.dtype = dtype
        self.stage = [torch.empty((self.chunk, L), dtype=dtype, device=dev) for _ in range(2)]
        self._free = [None, None]                      # compute-done events per staging buffer (carried across batches)
        self._n_chunks = 0
        self._slots = [None] * depth                   # per-batch buffers, allocated on first use
        self._T = T
        self.words_init = torch.tensor([0, _INF_BITS], dtype=torch.int32, device=dev)
        self.copy_stream = torch.cuda.Stream(dev)
        self.d2h_stream = torch.cuda.Stream(dev)       # PCIe is full duplex: scores of chunk i go back under the H2D of chunk i+2
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.respeculated = 0          # how many calls needed the exact re-run
        self._next = 0
        for i in range(depth):
            self._slot(i)
        torch.cuda.current_stream(dev).synchronize()   # buffers were allocated on the caller's stream: order the side streams once

    def _slot(self, i: int):
        if self._slots[i] is None:
            dev, m = self.device, self.model
            self._slots[i] = {
                "db": torch.empty((self.max_batch, 64, self._T), dtype=torch.float32, device=dev),
                "words": torch.zeros(2, dtype=torch.int32, device=dev),            # [max power bits, min power bits]
                "probs": torch.empty((self.max_batch, m.outputdim), dtype=torch.float32, device=dev),
                "out_host": torch.empty((self.max_batch, m.outputdim), dtype=torch.float32).pin_memory(),
                "words_host": torch.zeros(2, dtype=torch.int32).pin_memory(),
                "ticket": None,
            }
        return self._slots[i]

    @torch.no_grad()
    def submit(self, wav_host: torch.Tensor) -> int:
        """Queue one batch (pinned host [B, L]); returns a ticket for ``result``.  No host synchronisation."""
        if wav_host.is_cuda or wav_host.dtype != self.dtype or wav_host.dim() != 2 or wav_host.shape[1] != self.L:
            raise ValueError(f"expected a host {self.dtype} [B, {self.L}] tensor")
        B = wav_host.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the pipeline capacity {self.max_batch}")
        m = self.model
        if m.training:
            raise NotImplementedError("inference only: call model.eval()")
        ticket = self._next
        self._next += 1
        S = self._slot(ticket % self.depth)
        if S["ticket"] is not None:
            raise RuntimeError(f"batch {S['ticket']} is still pending: call result() before submitting {self.depth} more batches")
        main = torch.cuda.current_stream(self.device)
        words = S["words"]
        words.copy_(self.words_init)
        max_w, min_w = words[0:1], words[1:2]
        sharded = m.process_group is not None
        # (no stream-wide waits here: the staging buffers are guarded by their own events and the slot's previous download was
        # consumed by result(), so the uploads of this batch queue right behind those of the previous one)
        # Sharded: the top-dB scope is the GLOBAL batch.  Speculation stays on (tensor-core configuration): every chunk is encoded
        # with this rank's running maximum, ONE all-reduce(MAX) of the word follows the last chunk, and the host check in
        # result() compares this rank's minimum with the global cutoff.
        spec = self.speculative and (not sharded or m._cfg().tensor_core)
        free = self._free
        db, probs, out_host = S["db"], S["probs"], S["out_host"]
        for b0 in range(0, B, self.chunk):
            i = self._n_chunks
            self._n_chunks += 1
            nb = min(self.chunk, B - b0)
            buf = self.stage[i & 1][:nb]
            with torch.cuda.stream(self.copy_stream):
                if free[i & 1] is not None:
                    self.copy_stream.wait_event(free[i & 1])
                buf.copy_(wav_host[b0:b0 + nb], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            main.wait_event(ready)
            db_i = db[b0:b0 + nb]
            m.front_end.logmel_unclamped(buf, out=db_i, max_pow=max_w, min_pow=min_w)
            done = torch.cuda.Event()
            done.record(main)
            free[i & 1] = done
            if spec:
                m.encode(db_i, max_w, out=probs[b0:b0 + nb])
                scored = torch.cuda.Event()
                scored.record(main)
                with torch.cuda.stream(self.d2h_stream):
                    self.d2h_stream.wait_event(scored)
                    out_host[b0:b0 + nb].copy_(probs[b0:b0 + nb], non_blocking=True)
        final = words                                  # [final max, this rank's min]
        if sharded:
            final = words.clone()
            torch.distributed.all_reduce(final[0:1], op=torch.distributed.ReduceOp.MAX, group=m.process_group)
        if not spec:
            m.encode(db[:B], final[0:1], out=probs[:B])
        tail = torch.cuda.Event()
        tail.record(main)
        with torch.cuda.stream(self.d2h_stream):       # the last bytes of the batch: (exact-path scores,) the two words
            self.d2h_stream.wait_event(tail)
            if not spec:
                out_host[:B].copy_(probs[:B], non_blocking=True)
            S["words_host"].copy_(final, non_blocking=True)
            finished = torch.cuda.Event()
            finished.record(self.d2h_stream)
        S.update(ticket=ticket, B=B, spec=spec, final=final, finished=finished)
        self.h2d_bytes = B * self.L * wav_host.element_size()
        self.d2h_bytes = B * m.outputdim * 4 + 8
        return ticket

    @torch.no_grad()
    def result(self, ticket: int) -> torch.Tensor:
        """Host scores [B, outputdim] (pinned view) of a submitted batch; blocks until its last download has landed."""
        S = self._slots[ticket % self.depth]
        if S is None or S["ticket"] != ticket:
            raise RuntimeError(f"unknown or already collected ticket {ticket}")
        S["finished"].synchronize()
        B, m = S["B"], self.model
        out = S["out_host"][:B]
        if S["spec"]:
            mx, mn = int(S["words_host"][0]), int(S["words_host"][1])
            if B and _bits_to_db(mn) < _bits_to_db(mx) - 120.0 + 1e-3:      # conservative margin vs the device's log2-based dB
                # some value lies below the final cutoff: redo the encoder with the final batch maximum (exact path)
                self.respeculated += 1
                main = torch.cuda.current_stream(self.device)
                m.encode(S["db"][:B], S["final"][0:1], out=S["probs"][:B])
                out.copy_(S["probs"][:B], non_blocking=True)
                main.synchronize()
        S["ticket"] = None
        return out

    def __call__(self, wav_host: torch.Tensor) -> torch.Tensor:
        """wav_host: pinned float32 [B, L] host tensor.  Returns a pinned host view [B, outputdim]."""
        return self.result(self.submit(wav_host))


class FrontEndHostPipeline:
    """``model.front_end`` for host buffers (BASELINE config 4: the log-mel front-end alone): pinned host waveforms [B, L] in,
    pinned host log-mel dB [B, 64, T] out.  Chunked like ``HostPipeline``: the H2D copy of chunk i+1 overlaps the log-mel kernel
    of chunk i, and the UN-clamped dB of every chunk goes back at once on a third stream.  The reference's top-dB clamp is
    batch-global (Q2); the kernels track the batch minimum, and ``min_dB >= max_dB - 120`` proves that the clamp would not have
    changed a single value.  Otherwise the clamp pass runs on the device and the result is copied again (exact either way)."""

    def __init__(self, model, max_batch: int, L: int, chunk: int = 256, device: Optional[torch.device] = None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.UitkError("FrontEndHostPipeline needs the model on a CUDA device (no CPU fallback)")
        if model.process_group is not None:
            raise NotImplementedError("FrontEndHostPipeline is per rank; shard the clips and all-reduce the words yourself")
        self.max_batch, self.L, self.chunk = max_batch, L, max(1, min(chunk, max_batch))
        T = int(N.lib().uitk_num_frames(L))
        dev = self.device
        self.stage = [torch.empty((self.chunk, L), dtype=torch.float32, device=dev) for _ in range(2)]
        self.db = torch.empty((max_batch, 64, T), dtype=torch.float32, device=dev)
        self.out_host = torch.empty((max_batch, 64, T), dtype=torch.float32).pin_memory()
        self.words = torch.zeros(2, dtype=torch.int32, device=dev)
        self.words_init = torch.tensor([0, _INF_BITS], dtype=torch.int32, device=dev)
        self.words_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.copy_stream, self.d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.h2d_bytes = self.d2h_bytes = 0
        self.reclamped = 0

    @torch.no_grad()
    def __call__(self, wav_host: torch.Tensor) -> torch.Tensor:
        if wav_host.is_cuda or wav_host.dtype != torch.float32 or wav_host.dim() != 2 or wav_host.shape[1] != self.L:
            raise ValueError(f"expected a host float32 [B, {self.L}] tensor")
        B = wav_host.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the pipeline capacity {self.max_batch}")
        fe = self.model.front_end
        main = torch.cuda.current_stream(self.device)
        self.words.copy_(self.words_init)
        max_w, min_w = self.words[0:1], self.words[1:2]
        self.copy_stream.wait_stream(main)
        self.d2h_stream.wait_stream(main)
        free = [None, None]
        for i, b0 in enumerate(range(0, B, self.chunk)):
            nb = min(self.chunk, B - b0)
            buf = self.stage[i & 1][:nb]
            with torch.cuda.stream(self.copy_stream):
                if free[i & 1] is not None:
                    self.copy_stream.wait_event(free[i & 1])
                buf.copy_(wav_host[b0:b0 + nb], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            main.wait_event(ready)
            fe.logmel_unclamped(buf, out=self.db[b0:b0 + nb], max_pow=max_w, min_pow=min_w)
            done = torch.cuda.Event()
            done.record(main)
            free[i & 1] = done
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(done)
                self.out_host[b0:b0 + nb].copy_(self.db[b0:b0 + nb], non_blocking=True)
        self.words_host.copy_(self.words, non_blocking=True)
        main.synchronize()
        self.d2h_stream.synchronize()
        mx, mn = int(self.words_host[0]), int(self.words_host[1])
        if B and _bits_to_db(mn) < _bits_to_db(mx) - 120.0 + 1e-3:
            self.reclamped += 1
            fe.clamp_(self.db[:B], max_w)
            self.out_host[:B].copy_(self.db[:B], non_blocking=True)
            main.synchronize()
        self.h2d_bytes = B * self.L * 4
        self.d2h_bytes = self.out_host[:B].numel() * 4 + 8
        return self.out_host[:B]


class BatchPipeline:
    """Device-resident batches, two in flight: the front-end kernel (K1) of batch i+1 runs on its own stream while the encoder
    (K2 + head) of batch i drains on another, so the SMs that the persistent encoder CTAs leave idle in their last tile wave
    (820 tiles over 296 CTAs at 4096 clips) take log-mel CTAs of the next batch, and the encoder CTAs of batch i+1 start on
    SMs as the log-mel grid runs out.  Every batch is computed by exactly the kernels (and in the order) ``model.forward``
    uses - the scores are bit-identical - only the launches of consecutive batches are no longer serialised.

        t = pipe.submit(x)           # x: CUDA [B, L]; must stay unchanged until result(t)
        ...                          # submit the next batch before asking for this one
        y = pipe.result(t)           # CUDA [B, outputdim]; valid until `depth` more batches were submitted

    ``result`` orders the caller's current stream behind the batch (no host synchronisation)."""

    def __init__(self, model, depth: int = 2, device: Optional[torch.device] = None, timing: bool = False):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.UitkError("BatchPipeline needs the model on a CUDA device (no CPU fallback)")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch in the front-end, one in the encoder)")
        self.depth = depth
        self.fe_stream = torch.cuda.Stream(self.device)
        self.enc_stream = torch.cuda.Stream(self.device)
        ev = lambda: torch.cuda.Event(enable_timing=timing)
        self.slots = [dict(db=None, words=None, probs=None, fe_start=ev(), fe_done=ev(), enc_done=ev(), ticket=-1) for _ in range(depth)]
        self.n = 0

    @torch.no_grad()
    def submit(self, x: torch.Tensor, wait_current: bool = True) -> int:
        m = self.model
        m._check_ready(x, "BatchPipeline.submit")
        if x.dim() != 2:
            raise ValueError(f"expected a [B, L] waveform batch, got shape {tuple(x.shape)}")
        slot = self.slots[self.n % self.depth]
        B, L = x.shape
        T = int(N.lib().uitk_num_frames(L))
        if slot["db"] is None or tuple(slot["db"].shape) != (B, 64, T):
            # (re)allocated on the front-end stream, which is also the stream that first writes them
            for old in (slot["db"], slot["probs"]):
                if old is not None:
                    old.record_stream(self.enc_stream)
            with torch.cuda.stream(self.fe_stream):
                slot["db"] = torch.empty((B, 64, T), dtype=torch.float32, device=self.device)
                slot["probs"] = torch.empty((B, m.outputdim), dtype=torch.float32, device=self.device)
        if wait_current:
            self.fe_stream.wait_stream(torch.cuda.current_stream(self.device))      # x was produced on the caller's stream
        with torch.cuda.stream(self.fe_stream):
            if slot["ticket"] >= 0:
                self.fe_stream.wait_event(slot["enc_done"])                          # the slot's previous batch has left the encoder
            slot["fe_start"].record(self.fe_stream)
            words = m._new_words(self.device)
            if B:
                m.front_end.logmel_unclamped(x, out=slot["db"], max_pow=words[0:1], min_pow=words[1:2])
            slot["fe_done"].record(self.fe_stream)
        with torch.cuda.stream(self.enc_stream):
            self.enc_stream.wait_event(slot["fe_done"])
            words.record_stream(self.enc_stream)
            if B:
                m._finish(slot["db"], words, out=slot["probs"])
            elif m.process_group is not None:      # an empty shard still joins the collective of its group
                m._finish(torch.empty((0, 64, 1), dtype=torch.float32, device=self.device), words)
            slot["enc_done"].record(self.enc_stream)
        slot["ticket"] = self.n
        self.n += 1
        return slot["ticket"]

    def result(self, ticket: int) -> torch.Tensor:
        slot = self.slots[ticket % self.depth]
        if slot["ticket"] != ticket:
            raise ValueError(f"batch {ticket} is no longer held (only the last {self.depth} submitted batches are)")
        torch.cuda.current_stream(self.device).wait_event(slot["enc_done"])
        return slot["probs"]

    def events(self, ticket: int):
        """(front-end start, front-end done, encoder done) events of a held batch (timing=True to read durations)."""
        slot = self.slots[ticket % self.depth]
        return slot["fe_start"], slot["fe_done"], slot["enc_done"]
