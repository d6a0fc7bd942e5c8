"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line (stall samples)."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
sa = H.index("Warp Stall Sampling (All Samples)"); ie = H.index("Instructions Executed")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= sa or r[2] != "-":      # keep the per-line summary rows (Address == "-")
        continue
    try:
        agg[(r[0], r[1])] = (int(r[sa] or 0), int(r[ie] or 0))
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for (ln, src), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}%  inst={n:9d}  L{ln:>4s}  {src.strip()[:120]}")
