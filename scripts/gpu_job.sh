#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/bench_e.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'], d['frontend']['fps_e2e'], d['frontend']['fps_resident'])
PY
