python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k heat 2>&1 | tail -3
for t in "2:0" "2:64" "2:128" "2:256"; do echo "== tune $t"; python scripts/kernel_probe.py --guidance-only --tune=$t; done
