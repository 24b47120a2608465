#!/bin/bash
cd /root/repo
python -m pytest tests/test_prep_gpu.py -x -q 2>&1 | tail -25
python -m pytest tests -m gpu -x -q --deselect tests/test_prep_gpu.py 2>&1 | tail -5
