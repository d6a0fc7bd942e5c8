"""Sweep the start offset of the second resident CTA of the tensor-core encoder (debug knob) and time the encoder alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U
from uit_mobile_b200 import _native as N

lib = N.lib()
torch.manual_seed(0)
arch = os.environ.get("SWEEP_ARCH", "uit_xs")
batch = int(os.environ.get("SWEEP_BATCH", 4096))
model = getattr(U.models, arch)(outputdim=537, target_length=102).to("cuda:0").eval()
x = (0.1 * torch.randn(batch, 16000, device="cuda:0")).clamp_(-1, 1)
with torch.no_grad():
    db, mp = model.front_end.logmel_unclamped(x)[:2]
    out = torch.empty(batch, 537, device="cuda:0")
    for cyc in [0, 2048, 4096, 6144, 8192, 10240, 12288, 16384, 24576]:
        lib.uitk_debug_taps(((cyc // 256) + 1) << 8)
        for _ in range(3):
            model.encode(db, mp, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            model.encode(db, mp, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"stagger {cyc:6d} cycles: encoder {e0.elapsed_time(e1) / 20:.4f} ms", flush=True)
    lib.uitk_debug_taps(0)
