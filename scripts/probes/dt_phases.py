"""Developer probe (GPU box): per-stage cycle counters of the dense tracker's LM passes (SVS_DT_TIMING=1).
Tracks the same frame pair N times from the identity; the library prints cumulative cycles per level and stage."""
import os
import sys
os.environ["SVS_DT_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import capi, frontend_inputs as fi, synth_images as si

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seq = si.sequence(2)
cams = fi.level_cams(nlevels=3)
prev_p, cur_p = fi.float_pyramid(seq[0]["img"], 3), fi.float_pyramid(seq[1]["img"], 3)
dt = capi.DenseTracker(640, 480, 3, 0)
dt.set_disparity(seq[0]["disp"])
for l in range(3):
    dx, dy = fi.gradients(cur_p[l])
    dt.set_intrinsics(l, cams[l][0], cams[l][1], cams[l][2])
    dt.set_images(l, prev_p[l], cur_p[l], dx, dy)
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])
dt.compute_point_cloud(I7, cams)
tot_pass = np.zeros(3)
ms = []
devnull = os.open(os.devnull, os.O_WRONLY)
saved = os.dup(2)
for i in range(N):
    if i < N - 1:
        os.dup2(devnull, 2)        # only the last call's (cumulative) counters are shown
    else:
        os.dup2(saved, 2)
    T, st = dt.track(I7)
    tot_pass += np.array(st["passes"])
    ms.append(st["ms_total"])
print("frames", N, "passes per level (0,1,2) total", tot_pass, "ms per frame median", float(np.median(ms)), "us per pass", 1e3 * float(np.sum(ms)) / tot_pass.sum())
