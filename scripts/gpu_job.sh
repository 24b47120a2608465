#!/bin/bash
cd /root/repo
timeout 200 python -m pytest tests/test_map_gpu.py -x -q 2>&1 | tail -25
