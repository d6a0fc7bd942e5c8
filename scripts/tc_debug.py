"""Bring-up diagnostics for the tensor-core encoder (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import helpers as H
from tests.test_gpu_tensorcore import pack_kmajor, _model
from uit_mobile_b200 import _native as N
DEV = "cuda:0"
lib = N.lib()
for (n, k, init) in [(128, 128, False), (96, 128, False), (128, 32, True), (128, 256, False)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(128, k, generator=g); b = torch.randn(n, k, generator=g); c0 = torch.randn(128, n, generator=g)
    ref = a.to(torch.bfloat16).double() @ b.to(torch.bfloat16).double().T + (c0.double() if init else 0)
    out = torch.full((128, n), float("nan"), device=DEV)
    ad, bd, cd = a.to(DEV), pack_kmajor(b).to(DEV), c0.to(DEV)
    rc = lib.uitk_selftest_umma(ad.data_ptr(), bd.data_ptr(), cd.data_ptr() if init else None, out.data_ptr(), n, k, None)
    torch.cuda.synchronize()
    d = (out.cpu().double() - ref).abs()
    print(f"selftest N={n} K={k} init={init}: rc={rc} max err {d.max().item():.3e} nan={torch.isnan(out).sum().item()}", flush=True)
z = H.load_golden("trace_xxxs.npz")
x = torch.from_numpy(H.noise_clips(32)[:2]).to(DEV)
for depth in (1, 2, 4):
    m = _model("uit_xxxs", "trained", "bf16", depth)
    lib.uitk_debug_taps(1)
    y = m(x); torch.cuda.synchronize()
    lib.uitk_debug_taps(0)
    tok = m._last_workspace[: 2 * 24 * 128 * 4].view(torch.float32).view(2, 24, 128).cpu().numpy()
    ref = z["blocks"][depth - 1]
    print(f"depth {depth}: x max|d| {np.abs(tok-ref).max():.4f} (ref max {np.abs(ref).max():.2f}); per-row err {np.abs(tok-ref).max(-1)[0][:6]}", flush=True)
g = H.load_golden("probs.npz")
for arch in H.ARCHS:
    for kind in ("init", "trained"):
        m = _model(arch, kind, "bf16")
        y = m(torch.from_numpy(H.noise_clips(32)).to(DEV)).cpu().numpy()
        ref = g[f"{arch}/{kind}/noise"]
        print(f"{arch}/{kind}: max|d prob| {np.abs(y-ref).max():.3e}  top5 tie-aware {H.tie_aware_topk_equal(ref, y, 5, 1e-3)}", flush=True)
