"""Stage timeline of the tensor-core encoder (needs a -DUITK_TRACE build: UITK_TRACE=1 python -m uit_mobile_b200.build --force).

Prints, for CTA 0, the mean cycles between consecutive stage stamps of compute thread 0 over all blocks of each tile, and
the same for the MMA-issuer thread.  Output also lands in gpurun_out/tc_trace.txt.
"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import uit_mobile_b200 as U
from uit_mobile_b200 import _native as N

NAMES = {1: "tile start", 2: "patch gather+signal", 3: "patch MMA wait", 10: "pos add / block start (param wait)",
         11: "LN1+signal", 12: "qkv MMA wait", 13: "qkv epilogue+signal", 14: "S MMA wait", 15: "softmax+P+signal",
         16: "PV MMA wait", 17: "O read(+signal)", 18: "A_o signal", 19: "proj MMA wait", 20: "LN2+signal",
         21: "fc1/H wait", 22: "relu epilogue+signal", 23: "final fc2 waits", 30: "final LN + pool"}
INAMES = {1: "ready signal received", 2: "weights in ring", 3: "MMAs issued+commit"}


def decode(buf):
    ids = (buf >> 44).astype(np.int64)
    clk = (buf & ((1 << 44) - 1)).astype(np.int64)
    n = int(np.argmax(ids == 0)) if (ids == 0).any() else len(ids)
    return ids[:n], clk[:n]


def main():
    batch = int(os.environ.get("TRACE_BATCH", 4096))
    arch = os.environ.get("TRACE_ARCH", "uit_xs")
    lib = N.lib()
    torch.manual_seed(0)
    model = getattr(U.models, arch)(outputdim=537, target_length=102).to("cuda:0").eval()
    x = (0.1 * torch.randn(batch, 16000, device="cuda:0")).clamp_(-1, 1)
    with torch.no_grad():
        for _ in range(int(os.environ.get("TRACE_WARM", 3))):
            model(x)
    torch.cuda.synchronize()
    out = []
    # event-timed duration of the two halves of forward() in this (trace) build, for comparison with the cycle stamps
    fe = model.front_end
    with torch.no_grad():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_fe = t_enc = 0.0
        for _ in range(10):
            words = fe.new_words(x.device)
            e[0].record()
            db, _ = fe.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
            e[1].record()
            model.encode(db, words[0:1])
            e[2].record()
            torch.cuda.synchronize()
            t_fe += e[0].elapsed_time(e[1]); t_enc += e[1].elapsed_time(e[2])
        out.append(f"== event-timed (trace build): front-end {t_fe / 10:.4f} ms, encoder+head {t_enc / 10:.4f} ms")
        for _ in range(int(os.environ.get("TRACE_WARM", 3))):
            model(x)
        torch.cuda.synchronize()
    cta = np.zeros(4096, dtype=np.int64)
    N.check(lib.uitk_debug_read_trace(cta.ctypes.data, 2, 4096), "read_trace")
    cta = cta.reshape(1024, 4)
    cta = cta[cta[:, 0] > 0]
    t0 = cta[:, 0].min()
    st, en = (cta[:, 0] - t0) / 1e3, (cta[:, 1] - t0) / 1e3
    out.append(f"== per-CTA lifetimes (globaltimer, us): {len(cta)} CTAs; start min/median/max {st.min():.1f}/{np.median(st):.1f}/{st.max():.1f}; "
               f"end min/median/max {en.min():.1f}/{np.median(en):.1f}/{en.max():.1f}; CTA 0: {st[0]:.1f} .. {en[0]:.1f}")
    mhz = cta[:, 3] / (cta[:, 1] - cta[:, 0]) * 1e3
    out.append(f"   effective SM clock over a CTA's life (cycles / globaltimer): median {np.median(mhz):.0f} MHz, min {mhz.min():.0f}, max {mhz.max():.0f}")
    order = np.argsort(en)
    out.append("   latest CTAs (idx, sm, start, end): " + "  ".join(f"({i},{cta[i, 2]},{st[i]:.0f},{en[i]:.0f})" for i in order[-6:]))
    sm_count = np.bincount(cta[:, 2].astype(int))
    out.append(f"   CTAs per SM: min {sm_count[sm_count > 0].min()} max {sm_count.max()} (SMs used {np.count_nonzero(sm_count)})")
    for which, names in ((0, NAMES), (1, INAMES)):
        buf = np.zeros(4096, dtype=np.int64)
        N.check(lib.uitk_debug_read_trace(buf.ctypes.data, which, 4096), "read_trace")
        ids, clk = decode(buf)
        if len(ids) < 2:
            out.append(f"== {'compute thread 0' if which == 0 else 'MMA issuer'}: no stamps (UITK_TRACE=2 build)")
            continue
        out.append(f"== {'compute thread 0' if which == 0 else 'MMA issuer'}: {len(ids)} stamps, total {clk[-1] - clk[0]} cycles")
        d = np.diff(clk)
        agg = defaultdict(list)
        if which == 0:
            # key = (previous id, id, occurrence index within the block for repeated ids)
            occ = defaultdict(int)
            for i in range(1, len(ids)):
                if ids[i] == 10:
                    occ.clear()
                k = (int(ids[i - 1]), int(ids[i]))
                occ[k] += 1
                agg[(k, occ[k])].append(int(d[i - 1]))
            block_total = [int(clk[j] - clk[i]) for i, j in zip(np.where(ids == 10)[0][:-1], np.where(ids == 10)[0][1:]) if j - i < 40]
            out.append(f"   cycles per block (stamp 10 -> next 10): mean {np.mean(block_total):.0f}  min {np.min(block_total)}  max {np.max(block_total)}")
            order = sorted(agg.keys(), key=lambda k: min(i for i in range(1, len(ids)) if (int(ids[i - 1]), int(ids[i])) == k[0]) * 100 + k[1])
            for k in order:
                v = agg[k]
                out.append(f"   {names.get(k[0][0], k[0][0])!s:>36} -> {names.get(k[0][1], k[0][1])!s:<36} #{k[1]:<2d} n={len(v):3d} mean {np.mean(v):8.0f}  min {np.min(v):6d}  max {np.max(v):6d}")
        else:
            for i in range(1, len(ids)):
                agg[(int(ids[i - 1]), int(ids[i]))].append(int(d[i - 1]))
            for k, v in sorted(agg.items()):
                out.append(f"   {names.get(k[0], k[0])!s:>24} -> {names.get(k[1], k[1])!s:<24} n={len(v):4d} mean {np.mean(v):8.0f}  min {np.min(v):6d}  max {np.max(v):6d}  sum {np.sum(v):9d}")
    # merged absolute timeline of one block (compute thread 0 + issuer), cycles relative to the block start
    bufs = []
    for which in (0, 1):
        buf = np.zeros(4096, dtype=np.int64)
        N.check(lib.uitk_debug_read_trace(buf.ctypes.data, which, 4096), "read_trace")
        bufs.append(decode(buf))
    ids0, clk0 = bufs[0]
    starts = np.where(ids0 == 10)[0]
    if len(starts) > 7:
        t_a, t_b = clk0[starts[5]], clk0[starts[6]]
        ev = [(int(c - t_a), "C", NAMES.get(int(i), str(int(i)))) for i, c in zip(ids0, clk0) if t_a <= c <= t_b]
        ids1, clk1 = bufs[1]
        ev += [(int(c - t_a), "  I", INAMES.get(int(i), str(int(i)))) for i, c in zip(ids1, clk1) if t_a - 500 <= c <= t_b]
        out.append("== merged timeline of block 5 (cycles from block start; C = compute thread 0, I = MMA issuer)")
        for t, w, n in sorted(ev):
            out.append(f"   {t:7d} {w} {n}")
    txt = "\n".join(out)
    print(txt)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_trace.txt", "w") as f:
        f.write(txt + "\n")


if __name__ == "__main__":
    main()
