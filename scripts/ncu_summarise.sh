#!/bin/bash
# Turn gpurun_out/${TAG}_prof_{encoder,logmel,head}.ncu-rep (scripts/gpu_final.sh) into the committed summaries under profiles/:
# raw metric CSVs, per-source-line tables (stalls for the encoder, instructions / smem wavefronts for log-mel), traffic.json.
# Runs here (no GPU needed): ncu -i only reads the report.
set -u
TAG=${TAG:-r2}
cd "$(dirname "$0")/.."
for k in encoder logmel head; do
  rep=gpurun_out/${TAG}_prof_$k.ncu-rep
  [ -f "$rep" ] || { echo "missing $rep"; continue; }
  ncu -i "$rep" --page raw --csv > profiles/${TAG}_${k}_ncu_raw.csv 2>/dev/null
done
ncu -i gpurun_out/${TAG}_prof_encoder.ncu-rep --page source --print-source cuda,sass --csv > /tmp/enc_src.csv 2>/dev/null \
  && python scripts/ncu_lines.py /tmp/enc_src.csv 45 > profiles/${TAG}_encoder_stall_by_line.txt
ncu -i gpurun_out/${TAG}_prof_logmel.ncu-rep --page source --print-source cuda,sass --csv > /tmp/lm_src.csv 2>/dev/null \
  && python scripts/ncu_lines.py /tmp/lm_src.csv 45 "Instructions Executed" > profiles/${TAG}_logmel_inst_by_line.txt
python - <<'PY'
import csv, json, os
TAG = os.environ.get("TAG", "r2")
out = {}
for k, name in (("encoder", "encoder_tc_kernel"), ("logmel", "logmel_kernel"), ("head", "head_tc_kernel")):
    p = f"profiles/{TAG}_{k}_ncu_raw.csv"
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(open(p)))
    hdr, vals = rows[0], rows[-1]
    m = {h: v for h, v in zip(hdr, vals)}
    f = lambda key: float(m[key].replace(",", "")) if key in m and m[key] not in ("", "n/a") else None
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    unit = {r[0]: r for r in rows[:3]}
    # ncu prints units in the second row; dram bytes may be in Mbyte / Gbyte
    units = dict(zip(hdr, rows[1])) if len(rows) > 2 else {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = rd * scale.get(units.get("dram__bytes_read.sum", "byte"), 1) if rd is not None else None
    wr = wr * scale.get(units.get("dram__bytes_write.sum", "byte"), 1) if wr is not None else None
    dur = f("gpu__time_duration.sum")
    if dur is not None:
        dur *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units.get("gpu__time_duration.sum", "us"), 1)
    out[name] = {"traffic_bytes": (rd or 0) + (wr or 0), "dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us_under_ncu": dur,
                 "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "inst_executed": f("smsp__inst_executed.sum"),
                 "smem_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                 "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 "registers_per_thread": f("launch__registers_per_thread")}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
