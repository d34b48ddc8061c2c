python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "llg or slab or tuning" 2>&1 | tail -3
python scripts/kernel_probe.py 8 2048 2048 --llg
scripts/ncu_times.sh 32 128 128 --llg
