#!/bin/bash
# per-kernel durations of one B=8 graph replay with WARM caches (ncu --cache-control none): the real cost of each launch in the step
mkdir -p gpurun_out
KPS=1430
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s $((KPS * 2)) -c $KPS --csv \
    --log-file gpurun_out/r02v_b8_warm_launches.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu --no-parity --eager-gpu 0 > gpurun_out/r02v_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02v_b8_warm_launches.csv gpurun_out/r02v_b8_warm_launches_summary.csv | head -14
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open("gpurun_out/r02v_b8_warm_launches.csv", errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[0]; ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = rows[1:]
idx = [i for i, r in enumerate(data) if "token_taps" in r[ik]]
s = idx[16]
tot = 0
for r in data[s:s + 42]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    tot += float(r[iv].replace(",", ""))
    print(f"{name:34s} {r[ig]:>14s} {r[iv]:>8s} ns")
print("step total us", tot / 1e3)
PY
