import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import helpers as H
from tests.test_gpu_tensorcore import _model
from uit_mobile_b200 import _native as N
lib = N.lib(); DEV = "cuda:0"
z = H.load_golden("trace_xxxs.npz")
x = torch.from_numpy(H.noise_clips(32)[:7]).to(DEV)
m = _model("uit_xxxs", "trained", "bf16", 1)
outs = {}
for name, flag in (("tc", 1), ("legacy", 3)):
    lib.uitk_debug_taps(flag)
    m(x); torch.cuda.synchronize()
    outs[name] = m._last_workspace[: 7 * 24 * 128 * 4].view(torch.float32).view(7 * 24, 128).cpu().numpy().copy()
lib.uitk_debug_taps(0)
d = np.abs(outs["tc"] - outs["legacy"])
print("max diff tc vs legacy", d.max())
rows = d.max(1)
for r0 in range(0, 168, 24):
    print("clip", r0 // 24, np.array2string(rows[r0:r0 + 24], precision=3, max_line_width=250))
cols = d.max(0); print("worst cols", np.argsort(-cols)[:10], cols.max())
