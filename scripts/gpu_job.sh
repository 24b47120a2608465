timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_c.json').read()); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms_per_step','gpu_launches')}); print(d['roofline']); print(d['frontend'])"
