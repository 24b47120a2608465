python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python scripts/dev_ba.py 2>&1 | grep -A1 "^C[25]"
