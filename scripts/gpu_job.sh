#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_dist_gpu.py tests/test_cpp_shim.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/bench_f.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'], d['frontend']['fps_e2e'], d['frontend']['fps_resident'])
PY
