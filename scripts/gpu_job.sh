timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -12
