#!/bin/bash
# Warm-cache per-kernel durations (ncu gpu__time_duration, caches not flushed) of one kernel_probe run.
# usage (GPU box): scripts/ncu_times.sh [kernel_probe args]
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:'march|guidance|llg|update|euler|init|fused' --csv --log-file gpurun_out/ncuw.csv python scripts/kernel_probe.py "$@" > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/ncuw.csv")) if len(r)>10]
i={h:k for k,h in enumerate(rows[0])}
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[i["Kernel Name"]].split("(")[0][-45:]].append(float(r[i["Metric Value"]]))
for k,v in d.items(): print(f"{k:48s} n={len(v)} min={min(v)/1e3:.2f} med={sorted(v)[len(v)//2]/1e3:.2f} us")
PY
