python scripts/kernel_probe.py --guidance-only
ncu --set full --import-source on --clock-control none -k regex:'heat_march_reduce' -s 1 -c 1 -o gpurun_out/prof_r1i_reduce python scripts/kernel_probe.py --guidance-only > gpurun_out/ncu_r1i.log 2>&1
tail -2 gpurun_out/ncu_r1i.log
