#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_pose_gpu.py -x -q 2>&1 | tail -30
