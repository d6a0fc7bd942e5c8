"""Multi-GPU invariance: scores of a batch sharded over G GPUs == the single-GPU scores, bit for bit,
including the batch-global top-dB scope (a silent clip on one rank, the loud one on another)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import uit_mobile_b200 as U
from uit_mobile_b200 import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = U.models.uit_xxs(outputdim=537, target_length=102).to(dev).eval()
g = torch.Generator().manual_seed(5)
total = 64 * world + 3                                   # ragged shards
x = (0.1 * torch.randn(total, 16000, generator=g)).clamp_(-1, 1)
x[0] = 0.0                                               # silent clip on rank 0 ...
x[-1] = (0.9 * torch.randn(16000, generator=g)).clamp_(-1, 1)   # ... the loudest one on the last rank
with torch.no_grad():
    ref = model(x.to(dev))                               # single-GPU, whole batch
    align = model.tile_clips(101)                        # tile-aligned shards: bit-identical to the single-GPU run
    b, e = sharding.shard_bounds(total, rank, world, align)
    got = sharding.sharded_forward(model, x[b:e].to(dev), total, align=align)
ok = torch.equal(ref, got)
# the same gather with the copy engines over NVLink peer memory (sharding.PeerGather), twice (slot reuse), ragged shards
peer_note = "PeerGather"
try:
    sizes = [sharding.shard_bounds(total, r, world, align)[1] - sharding.shard_bounds(total, r, world, align)[0] for r in range(world)]
    pg = sharding.PeerGather(sizes, 537, dist.group.WORLD, dev)
    model.process_group = dist.group.WORLD
    for rep in range(3):
        with torch.no_grad():
            out, done = pg(model(x[b:e].to(dev)))
        torch.cuda.current_stream().wait_event(done)
        ok = ok and torch.equal(ref, out)
    # BatchPipeline under a process group (async max all-reduce + conditional re-run on the encoder stream)
    from uit_mobile_b200.pipeline import BatchPipeline
    bp = BatchPipeline(model, depth=2)
    xl = x[b:e].to(dev)
    t0 = bp.submit(xl); t1 = bp.submit(xl)
    for t in (t0, t1):
        out, done = pg(bp.result(t))
        torch.cuda.current_stream().wait_event(done)
        ok = ok and torch.equal(ref, out)
    # the batch-maximum word through NVLink peer memory (sharding.PeerWords) instead of the NCCL all-reduce: plain forward,
    # several steps (epoch ring), and through BatchPipeline.  The silent clip sits on rank 0, the loudest on the last rank, so
    # the conditional exact re-run really fires on rank 0.
    model.peer_words = sharding.PeerWords(dist.group.WORLD, dev)
    for rep in range(6):
        with torch.no_grad():
            out, done = pg(model(xl))
        torch.cuda.current_stream().wait_event(done)
        ok = ok and torch.equal(ref, out)
    t0 = bp.submit(xl); t1 = bp.submit(xl)
    for t in (t0, t1):
        out, done = pg(bp.result(t))
        torch.cuda.current_stream().wait_event(done)
        ok = ok and torch.equal(ref, out)
    model.peer_words = None
    peer_note += " + PeerWords"
except Exception as exc:                                   # no peer mapping on this box
    peer_note = f"PeerGather unavailable ({type(exc).__name__}: {str(exc)[:100]})"
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"shard invariance over {world} GPUs ({total} clips, ragged; NCCL gather, {peer_note}, BatchPipeline): "
          f"{'BIT-EXACT' if t.item() else 'MISMATCH'}; max|d| = {(ref - got).abs().max().item():.3e}")
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
