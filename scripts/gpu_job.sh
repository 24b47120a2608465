timeout 200 python -m pytest tests/test_ba_gpu.py tests/test_dist_gpu.py tests/test_cpp_shim.py -x -q -m gpu 2>&1 | tail -3
SVS_HOST_TIMING=1 timeout 100 python - 2>&1 <<'PY' | tail -12
import sys, time; sys.path.insert(0,'.')
from scavislam_b200 import capi, synth
pb = synth.make_config("C2"); ba = capi.BundleAdjuster()
for i in range(4):
    t=time.perf_counter(); it,p,s,st = ba.optimise_inner_and_outer_window(pb, 10); t2=time.perf_counter()
    print("one-call", (t2-t)*1e3, "ms; device", st["ms_total"])
PY
nproc
