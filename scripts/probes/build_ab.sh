#!/bin/bash
# A/B of the build kernel's scheduling on a B200 box (gpurun -- 'bash scripts/probes/build_ab.sh')
cd /root/repo
mkdir -p gpurun_out
{
python scripts/probes/build_ab.py
SVS_BUILD_STATIC=1 python scripts/probes/build_ab.py
SVS_BUILD_STATIC=1 SVS_BUILD_NO_LPT=1 python scripts/probes/build_ab.py
SVS_BUILD_NO_LPT=1 python scripts/probes/build_ab.py C2 C5
for c in 4 6 12; do SVS_BUILD_CHUNK=$c python scripts/probes/build_ab.py C2 C2d C5; done
} 2>&1 | grep -v "^$" | tee gpurun_out/build_ab.txt
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_dt_gpu.py tests/test_cpp_shim.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/build_ab.txt
