"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partition, max-word all-reduce, score gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uit_mobile_b200 import sharding as S


def test_shard_bounds_partition():
    for total in (0, 1, 5, 4096, 65536, 35991):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.shard_bounds(10, 2, 2)
    for total in (7, 131, 4096, 65536):                      # tile-aligned shards (5 clips per 128-row tile)
        for world in (2, 3, 8):
            spans = [S.shard_bounds(total, r, world, align=5) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(b % 5 == 0 or b == total for b, _ in spans)


def test_window_shards_cover_stream_with_halo():
    n, hop, win = 35991, 1600, 16000           # 1 h of audio, BASELINE config 5
    assert (57_600_000 - win) // hop + 1 == n
    prev_w1 = 0
    for r in range(8):
        w0, w1, s0, s1 = S.window_shard_bounds(n, hop, win, r, 8)
        assert w0 == prev_w1 and s0 == w0 * hop and s1 - s0 == (w1 - w0 - 1) * hop + win
        prev_w1 = w1
    assert prev_w1 == n and s1 == 57_600_000


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # max word: the larger POWER must win through the int32 view
        powers = torch.tensor([3.5e-7 if rank == 0 else 12.25], dtype=torch.float32)
        word = powers.view(torch.int32).clone()
        S.allreduce_max_word(word)
        assert word.view(torch.float32).item() == 12.25
        with pytest.raises(TypeError):
            S.allreduce_max_word(powers)
        # gather (even and ragged)
        b, e = S.shard_bounds(total, rank, world)
        full = torch.arange(total * 3, dtype=torch.float32).view(total, 3)
        got = S.gather_scores(full[b:e].clone(), total)
        assert torch.equal(got, full)
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gloo_world2_allreduce_and_gather(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
