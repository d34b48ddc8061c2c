#!/bin/bash
# Quick per-launch counters of the marching kernels (kernel_probe workload); prints a compact table.
# usage (on the GPU box): scripts/ncu_quick.sh <tag> [kernel_probe args]
tag=$1; shift
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
python scripts/kernel_probe.py "$@" > gpurun_out/probe_$tag.log 2>&1 || { tail -5 gpurun_out/probe_$tag.log; exit 1; }
ncu --metrics $M --clock-control none -k regex:'march|guidance|llg' -s 2 -c 4 --csv --log-file gpurun_out/ncuq_$tag.csv python scripts/kernel_probe.py "$@" > /dev/null 2>&1
cat gpurun_out/probe_$tag.log
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncuq_$tag.csv")) if len(r)>10]
hdr=rows[0]; i={h:k for k,h in enumerate(hdr)}
seen={}
for r in rows[1:]:
    key=(r[i["ID"]], r[i["Kernel Name"]][:60]); seen.setdefault(key,{})[r[i["Metric Name"]]]=r[i["Metric Value"]]
for (id_,name),m in seen.items():
    print(id_, name)
    print("   ", ", ".join(f"{k.split('__')[-1][:38]}={v}" for k,v in m.items()))
PY
