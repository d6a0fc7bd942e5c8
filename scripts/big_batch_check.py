"""BASELINE config 3 shape on one GPU: UiT-XXS, 65 536 synthetic 1 s clips in one forward call (chunked launches inside)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U
torch.manual_seed(0)
dev = "cuda:0"
m = U.models.uit_xxs(outputdim=537, target_length=102).to(dev).eval()
B = 65536
g = torch.Generator(device=dev).manual_seed(1)
x = (0.1 * torch.randn(B, 16000, generator=g, device=dev)).clamp_(-1, 1)
with torch.no_grad():
    y = m(x); torch.cuda.synchronize()
    t0 = time.perf_counter(); y = m(x); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    small = m(x[:16380].contiguous())            # first launch chunk on its own: same tile positions
assert torch.isfinite(y).all() and y.shape == (B, 537)
print(f"uit_xxs B={B}: {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M clips/s; first chunk equal to standalone run: {torch.equal(y[:16380], small)}")
# sliding windows over a 10-minute stream (config 5 shape, scaled): windows read in place, hop 1600
stream = (0.1 * torch.randn(16000 * 600, generator=g, device=dev)).clamp_(-1, 1)
hop = 1600; W = (stream.numel() - 16000) // hop + 1
with torch.no_grad():
    db, mp = m.front_end.logmel_unclamped(stream, ld=hop, B=W, L=16000)
    p = m.encode(db, mp); torch.cuda.synchronize()
    t0 = time.perf_counter()
    db, mp = m.front_end.logmel_unclamped(stream, ld=hop, B=W, L=16000); p = m.encode(db, mp); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"sliding windows: {W} windows (hop {hop}) over {stream.numel()/16000:.0f} s of audio in {dt*1e3:.2f} ms -> {W/dt/1e6:.2f} M windows/s, "
      f"{stream.numel()/16000/dt:.0f}x real time; finite={bool(torch.isfinite(p).all())}")
