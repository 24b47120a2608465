#!/bin/bash
cd /root/repo
SVS_HOST_TIMING=1 python scripts/dev_e2e.py 2>&1 | tail -10
