"""Developer probe (GPU box): host-side phases of one end-to-end call on C2 (SVS_HOST_TIMING=1)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scavislam_b200 import synth, capi

ba = capi.BundleAdjuster()
pb = synth.make_config("C2")
for rep in range(4):
    t = time.perf_counter()
    it, _, _, st = ba.optimise_inner_and_outer_window(pb, 10)
    dt = time.perf_counter() - t
    sys.stderr.write("== call %d: %.3f ms wall, kernels %.3f ms\n" % (rep, dt * 1e3, st["ms_total"] if st else -1))
