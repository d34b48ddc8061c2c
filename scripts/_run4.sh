ncu --set full --import-source on --clock-control none -k regex:'llg_tile_vjp' -s 1 -c 1 -o gpurun_out/prof_r1g_llg_vjp python scripts/kernel_probe.py 8 2048 2048 --llg > gpurun_out/ncu_r1g.log 2>&1
tail -3 gpurun_out/ncu_r1g.log
