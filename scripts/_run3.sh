timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "llg or slab" 2>&1 | tail -15
timeout 300 bash scripts/ncu_quick.sh llg_c3 32 128 128 --llg 2>&1 | tail -14 | cut -c1-400
timeout 300 bash scripts/ncu_quick.sh llg_big 8 2048 2048 --llg 2>&1 | tail -14 | cut -c1-400
