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
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"shard invariance over {world} GPUs ({total} clips, ragged): {'BIT-EXACT' if t.item() else 'MISMATCH'}; "
          f"max|d| = {(ref - got).abs().max().item():.3e}")
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
