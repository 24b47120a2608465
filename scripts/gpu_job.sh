for c in 4 6 8 10 12 16 24; do
SVS_BUILD_CHUNK=$c timeout 60 python - <<PY
import sys, os; sys.path.insert(0,'.')
from scavislam_b200 import capi, synth
pb = synth.make_config("C2"); ba = capi.BundleAdjuster(); ba.set_problem(pb)
for i in range(3): ba.reset_state(); it, st = ba.optimize(10)
print("chunk", os.environ["SVS_BUILD_CHUNK"], "build us", round(100*st["ms_build"],1), "solve", round(100*st["ms_solve"],1), "update", round(100*st["ms_update"],1))
PY
done
