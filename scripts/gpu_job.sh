#!/bin/bash
cd /root/repo
python bench.py --steps 5 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sampler', d['frontend']['fps_e2e'], d['frontend']['fps_resident'], d['e2e']['value'])"
SVS_BENCH_NO_SAMPLER=1 python bench.py --steps 5 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nosampler', d['frontend']['fps_e2e'], d['frontend']['fps_resident'], d['e2e']['value'])"
