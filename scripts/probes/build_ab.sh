#!/bin/bash
# A/B of the build kernel's scheduling on a B200 box (gpurun -- 'bash scripts/probes/build_ab.sh')
cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_map_gpu.py tests/test_cpp_shim.py tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/probes/build_ab.py
SVS_BUILD_NO_PAD=1 python scripts/probes/build_ab.py C2d
for c in 12 16 20 24 32; do SVS_BUILD_CHUNK=$c python scripts/probes/build_ab.py C2 C2d C2l; done
SVS_BUILD_CHUNK=48 python scripts/probes/build_ab.py C5
SVS_HOST_TIMING=1 python scripts/probes/e2e_host_phases.py 2 2>&1 | tail -60
} 2>&1 | grep -v "^$" | tee gpurun_out/build_ab2.txt
