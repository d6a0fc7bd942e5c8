"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "Warp Stall Sampling (All Samples)"
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
sa = H.index(key); ie = H.index("Instructions Executed")
wf = H.index("L1 Wavefronts Shared"); wx = H.index("L1 Wavefronts Shared Excessive")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= sa or r[2] != "-":
        continue
    try:
        agg[(r[0], r[1])] = (int(r[sa] or 0), int(r[ie] or 0), int(r[wf] or 0), int(r[wx] or 0))
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()); totw = sum(v[2] for v in agg.values())
print("total", key, tot, " total smem wavefronts", totw)
for (ln, src), (s, n, w, x) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot,1):5.1f}%  inst={n:9d} smem_wf={100*w/max(totw,1):5.1f}% (excess {100*x/max(totw,1):4.1f}%)  L{ln:>4s}  {src.strip()[:100]}")
