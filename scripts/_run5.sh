python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "heat or slab or empty" 2>&1 | tail -2
python scripts/kernel_probe.py --guidance-only
python scripts/kernel_probe.py --guidance-only
