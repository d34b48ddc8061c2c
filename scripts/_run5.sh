python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "llg_norm" 2>&1 | tail -2
python scripts/kernel_probe.py 8 2048 2048 --llg
