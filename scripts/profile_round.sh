#!/bin/bash
# Evidence capture for profiles/ (run on the GPU box through gpurun).  Only text summaries come back: the
# .ncu-rep files are summarised on the box (scripts/ncu_summary.py) and deleted, gpurun_out/ must stay < 64 MiB.
# usage: scripts/profile_round.sh <tag>      e.g. r1j
tag=${1:-r1x}
OURS='regex:march|guidance|llg_|heun_update|euler_|init_kernel|finalize|halo'
BENCH="python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-large --skip-gpu-ref --skip-extras"
O=gpurun_out
# 1. the bench command, plain (its output is the only number that counts)
$BENCH > $O/${tag}_bench_plain.json 2> $O/${tag}_bench_plain.err || { tail -5 $O/${tag}_bench_plain.err; exit 1; }
# 2. launch list of the timed region of the same command: every kernel of ONE step, serialised, cold cache
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${tag}_bench_launches.csv $BENCH > $O/${tag}_ncu_launches.log 2>&1
TAG=$tag OUT=$O BENCH_CMD="$BENCH" python - <<'PY'
import csv, collections, gzip, os, shutil
tag, O, bench = os.environ["TAG"], os.environ["OUT"], os.environ["BENCH_CMD"]
rows = [r for r in csv.reader(open(f"{O}/{tag}_bench_launches.csv")) if len(r) > 10]
i = {h: k for k, h in enumerate(rows[0])}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    name = r[i["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1][:60]
    tot[short] += float(r[i["Metric Value"]]); cnt[short] += 1
allt = sum(tot.values())
with open(f"{O}/{tag}_bench_launch_shares.txt", "w") as f:
    f.write(f"# one timed step of [{bench}] under ncu (gpu__time_duration.sum, ns; serialised, cold cache): {len(rows)-1} launches, {allt/1e6:.3f} ms\n")
    for k, v in tot.most_common(40):
        f.write(f"{k:62s} n={cnt[k]:5d} total={v/1e3:12.1f} us share={v/allt:8.5f}\n")
    ours = [k for k in tot if any(s in k for s in ("march", "guidance", "llg_", "heun_update", "euler_", "init_kernel"))]
    f.write("# our kernels\n")
    for k in ours:
        f.write(f"{k:62s} n={cnt[k]:5d} total={tot[k]/1e3:12.1f} us share={tot[k]/allt:8.5f}\n")
with open(f"{O}/{tag}_bench_launches.csv", "rb") as a, gzip.open(f"{O}/{tag}_bench_launch_list.csv.gz", "wb") as b:
    shutil.copyfileobj(a, b)
PY
rm -f $O/${tag}_bench_launches.csv
cat $O/${tag}_bench_launch_shares.txt | tail -12
[ "$2" = "launches-only" ] && exit 0
# 3. full sections for our kernels inside the timed region
timeout 600 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k "$OURS" -c 5 -o $O/${tag}_bench_kernels $BENCH > $O/${tag}_ncu_full.log 2>&1
python scripts/ncu_summary.py $O/${tag}_bench_kernels.ncu-rep > $O/${tag}_bench_kernels_ncu_full_summary.txt 2>&1
# dram bytes per launch of each C-ABI kernel of the bench step (bench.py reports it as roofline.traffic)
TAG=$tag OUT=$O python - <<'PY'
import csv, io, json, os, subprocess
tag, O = os.environ["TAG"], os.environ["OUT"]
raw = subprocess.run(["ncu", "-i", f"{O}/{tag}_bench_kernels.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
i = {h: k for k, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"march_reduce": "dpde_guidance_reduce", "march_vjp": "dpde_guidance_vjp", "euler_predict_kernel": "dpde_euler_predict",
         "euler_bwd": "dpde_euler_predict_bwd", "heun_update": "dpde_heun_guided_update", "init_kernel": "dpde_sampler_init"}
out = {}
for r in rows[2:]:
    k = next((v for s_, v in names.items() if s_ in r[i["Kernel Name"]]), None)
    if k and k not in out:
        out[k] = int(sum(float(r[i[m]]) * scale.get(units[i[m]], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
json.dump({"captured": f"{tag}: ncu --set full of the bench command's timed region, dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "kernels": out}, open(f"{O}/{tag}_traffic.json", "w"), indent=1)
print(out)
PY
rm -f $O/${tag}_bench_kernels.ncu-rep
# 4. the same kernels on config 5's shape (8 x 2 x 4096^2) and the LLG kernels on 8 x 6 x 2048^2
timeout 600 ncu --set full --clock-control none -k "$OURS" -c 12 -o $O/${tag}_large_heat python scripts/kernel_probe.py --reps=1 > $O/${tag}_large_heat.log 2>&1
python scripts/ncu_summary.py $O/${tag}_large_heat.ncu-rep > $O/${tag}_large_heat_ncu_full_summary.txt 2>&1; rm -f $O/${tag}_large_heat.ncu-rep
timeout 600 ncu --set full --clock-control none -k "$OURS" -c 8 -o $O/${tag}_large_llg python scripts/kernel_probe.py 8 2048 2048 --llg --reps=1 > $O/${tag}_large_llg.log 2>&1
python scripts/ncu_summary.py $O/${tag}_large_llg.ncu-rep > $O/${tag}_large_llg_ncu_full_summary.txt 2>&1; rm -f $O/${tag}_large_llg.ncu-rep
python scripts/kernel_probe.py > $O/${tag}_probe_heat.log 2>&1
python scripts/kernel_probe.py 8 2048 2048 --llg > $O/${tag}_probe_llg.log 2>&1
du -sh $O; ls -la $O | tail -16
