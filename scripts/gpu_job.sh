#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_constraint_gpu.py -x -q 2>&1 | tail -25
