"""Developer probe (GPU box): host-side phases of end-to-end calls on C2 (SVS_HOST_TIMING=1), alternating between
two windows with different edge lists (full structure analysis every call), then the same window again."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import synth, capi

ba = capi.BundleAdjuster()
pb = synth.make_config("C2")
pb2 = synth.with_dropouts(pb, 0.02, seed=5)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cold = len(sys.argv) > 2 and sys.argv[2] == "cold"   # pause between calls, like a back-end tick (helper threads asleep)
if cold:
    import torch   # (bench.py runs with torch and its OpenMP runtime loaded)
for rep in range(8):
    w = (pb, pb2)[rep & 1] if rep < 6 else pb
    if cold:
        time.sleep(0.02)
    t = time.perf_counter()
    it, _, _, st = ba.optimise_inner_and_outer_window(w, iters)
    dt = time.perf_counter() - t
    sys.stderr.write("== call %d (%s): %.3f ms wall, kernels %.3f ms\n" % (rep, w.name, dt * 1e3, st["ms_total"] if st else -1))
