"""Raw pinned H2D bandwidth and the host pipeline's end-to-end rate vs chunk size (4096 x 1 s clips, UiT-XS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U
from uit_mobile_b200.pipeline import HostPipeline

dev = "cuda:0"
torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to(dev).eval()
B = 4096
x_host = (0.1 * torch.randn(B, 16000)).clamp_(-1, 1).pin_memory()
x_dev = torch.empty(B, 16000, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(2):
    x_dev.copy_(x_host, non_blocking=True)
torch.cuda.synchronize()
a, b = ev(), ev()
a.record()
for _ in range(10):
    x_dev.copy_(x_host, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"raw H2D 262 MB: {ms:.3f} ms = {B * 64000 / ms / 1e6:.1f} GB/s  (PCIe floor for fp32 input: {B / ms * 1e3 / 1e6:.3f} M clips/s)", flush=True)
for chunk in (255, 510, 1020, 2040, 4096):
    pipe = HostPipeline(model, B, 16000, chunk=chunk)
    for _ in range(3):
        pipe(x_host)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(20):
        pipe(x_host)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"chunk {pipe.chunk:5d}: {ms:.3f} ms/step  {B / ms * 1e3 / 1e6:.3f} M clips/s", flush=True)
