SVS_HOST_TIMING=1 python - <<'PY'
import sys, time; sys.path.insert(0,'.')
from scavislam_b200 import capi, synth
pb = synth.make_config("C2"); ba = capi.BundleAdjuster()
for i in range(3):
    t=time.perf_counter(); ba.set_problem(pb); t1=time.perf_counter(); it,p,s,st = ba.optimise_inner_and_outer_window(pb, 10); t2=time.perf_counter()
    print("set_problem", (t1-t)*1e3, "ms; one-call", (t2-t1)*1e3, "ms; device", st["ms_total"])
PY
