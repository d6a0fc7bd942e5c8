"""Time log-mel kernel variants (experiment builds libuitk_x*.so, -DK1X=n): 4096 x 1 s clips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uit_mobile_b200 import _native as N
variant = sys.argv[1]
if variant != "0":
    N.LIB_PATH = os.path.join(N.HERE, f"libuitk_x{variant}.so")
import uit_mobile_b200 as U
torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to("cuda:0").eval()
x = (0.1 * torch.randn(4096, 16000, device="cuda:0")).clamp_(-1, 1)
out = torch.empty(4096, 64, 101, device="cuda:0")
with torch.no_grad():
    for _ in range(3):
        model.front_end.logmel_unclamped(x, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        model.front_end.logmel_unclamped(x, out=out)
    e1.record()
    torch.cuda.synchronize()
print(f"variant {variant}: {e0.elapsed_time(e1) / 20:.4f} ms", flush=True)
