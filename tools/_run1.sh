python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -5
