"""Where the sharded step's extra time goes, on ONE GPU (world-size-1 NCCL group = the N-GPU code path): ms/step through
BatchPipeline for  (a) no process group, (b) group + NCCL all-reduce, (c) group + PeerWords, (d) = (c) + PeerGather.
Measured (one B200): 0.669 / 0.701 / 0.689 / 0.691 ms per step: the sharded step costs ~3 % on one GPU (publish, collect and the
two conditional re-run launches on the encoder stream); the rest of the 0.719 ms seen on 2 - 8 GPUs is the slowest rank setting
the pace."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import uit_mobile_b200 as U
from uit_mobile_b200 import sharding
from uit_mobile_b200.pipeline import BatchPipeline

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to(dev).eval()
x = (0.1 * torch.randn(4096, 16000, device=dev)).clamp_(-1, 1)
pg = sharding.PeerGather([4096], 537, dist.group.WORLD, dev, depth=2)
pw = sharding.PeerWords(dist.group.WORLD, dev)

def run(n, gather, depth):
    bp = BatchPipeline(model, depth=depth)
    def loop(k):
        prev = None
        for _ in range(k):
            t = bp.submit(x)
            if prev is not None:
                y = bp.result(prev)
                if gather:
                    out, done = pg(y); torch.cuda.current_stream().wait_event(done)
            prev = t
        y = bp.result(prev)
        if gather:
            out, done = pg(y); torch.cuda.current_stream().wait_event(done)
    loop(10); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    for name, group, words, gather, depth in (("(a) no group, depth 2", None, None, False, 2), ("(a') no group, depth 3", None, None, False, 3),
                                              ("(b) NCCL all-reduce", dist.group.WORLD, None, False, 3),
                                              ("(c) PeerWords", dist.group.WORLD, pw, False, 3), ("(d) PeerWords + PeerGather", dist.group.WORLD, pw, True, 3)):
        model.process_group, model.peer_words = group, words
        print(f"{name:62s} {run(300, gather, depth):.4f} ms/step")
dist.destroy_process_group()
