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
print("sanitize_small ok")
