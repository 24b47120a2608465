python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','kernel_ms_per_step')})"
