"""Time the log-mel front-end alone (CUDA events): 4096 x 1 s clips and 1024 x 10 s clips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U

torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to("cuda:0").eval()
for B, L in ((4096, 16000), (1024, 160000)):
    x = (0.1 * torch.randn(B, L, device="cuda:0")).clamp_(-1, 1)
    T = 1 + L // 160
    out = torch.empty(B, 64, T, device="cuda:0")
    with torch.no_grad():
        for _ in range(3):
            model.front_end.logmel_unclamped(x, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            model.front_end.logmel_unclamped(x, out=out)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = B * (4 * L + 4 * 64 * T) / ms / 1e6
    print(f"logmel {B} x {L}: {ms:.4f} ms  {gbs:.0f} GB/s algorithmic", flush=True)
