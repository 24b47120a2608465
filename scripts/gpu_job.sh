SVS_SOLVE_TIMING=1 python - <<'PY'
import sys; sys.path.insert(0,'.')
from scavislam_b200 import capi, synth
pb = synth.make_config("C2"); ba = capi.BundleAdjuster(); ba.set_problem(pb)
for i in range(2): ba.reset_state(); it, st = ba.optimize(2); print(st["ms_solve"]/st["trials_total"])
PY
python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
