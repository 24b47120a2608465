import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scavislam_b200 import capi, synth
mode = sys.argv[1]
if mode != "none":
    ba = capi.BundleAdjuster()
    pb = synth.make_config("C2")
    if mode == "e2e":
        ba.optimise_inner_and_outer_window(pb, 10)
    else:
        ba.set_problem(pb); ba.optimize(10)
r = bench.frontend_bench(0)
print(mode, os.environ.get("OMP_WAIT_POLICY"), r["fps_e2e"], r["fps_resident"])
