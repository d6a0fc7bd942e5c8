"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): both precisions, ragged tile, 2-crop branch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U
torch.manual_seed(0)
for prec in ("bf16", "fp32"):
    m = U.models.uit_xxxs(outputdim=537, target_length=102, precision=prec).to("cuda:0").eval()
    for B, L in ((7, 16000), (3, 16384), (2, 2400)):
        x = (0.1 * torch.randn(B, L)).clamp_(-1, 1).to("cuda:0")
        y = m(x)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all() and y.shape == (B, 537)
    db = m.front_end(x)
    torch.cuda.synchronize()
# round 2: sliding-window front-end (shared STFT), 16-bit PCM ingest, the two-batch pipeline, MobileNetV2
m = U.models.uit_xxxs(outputdim=537, target_length=102).to("cuda:0").eval()
stream = (0.1 * torch.randn(16000 * 3 + 777)).clamp_(-1, 1).to("cuda:0")
y = m.forward_sliding(stream, hop=1600)
pcm = (torch.randn(4, 16000) * 3000).to(torch.int16).to("cuda:0")
y = m(pcm)
from uit_mobile_b200.pipeline import BatchPipeline
bp = BatchPipeline(m)
xs = (0.1 * torch.randn(11, 16000)).clamp_(-1, 1).to("cuda:0")
t0 = bp.submit(xs); t1 = bp.submit(xs)
assert torch.equal(bp.result(t0), bp.result(t1))
mn = U.models.MobileNetV2(outputdim=537).to("cuda:0").eval()
y = mn(xs[:3, :4800])
torch.cuda.synchronize()
assert torch.isfinite(y).all() and y.shape == (3, 537)
print("sanitize_small ok")
