"""Host-side cost of ISSUING one sharded step (world-size-1 NCCL group: same code path as N GPUs, one GPU).
Prints a cProfile of 300 steps through model._finish with a process group, and the per-call wall time of each piece."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import uit_mobile_b200 as U
from uit_mobile_b200 import sharding

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to(dev).eval()
model.process_group = dist.group.WORLD
x = (0.1 * torch.randn(4096, 16000, device=dev)).clamp_(-1, 1)
try:
    peer = sharding.PeerGather([4096], 537, dist.group.WORLD, dev, depth=2)
except Exception as e:
    print("PeerGather unavailable:", e); peer = None

def step():
    with torch.no_grad():
        words = model._new_words(dev)
        db, _ = model.front_end.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
        probs = model._finish(db, words)
        if peer is not None:
            out, done = peer(probs)
            torch.cuda.current_stream().wait_event(done)
    return probs

for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3 * (t1 - t0) / n:.3f} ms/step; with drain {1e3 * (t2 - t0) / n:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
dist.destroy_process_group()
