"""Developer probe (GPU box): device time of calcFastMotionOnly for n observations, cluster kernel vs single CTA
(SVS_POSE_SINGLE_CTA=1)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import capi, synth_pose as sp

tag = "single CTA" if os.environ.get("SVS_POSE_SINGLE_CTA") else "cluster of 8"
po = capi.PoseOptimizer()
for n in (600, 1800, 5000):
    tr = sp.make_track(n, seed=3, outlier_frac=0.15)
    ms = []
    for rep in range(8):
        T, st = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], True, 2.0, 15, -1.0)
        ms.append(st["ms"])
    print(f"[{tag}] n={n}: {np.median(ms[2:]) * 1e3:.1f} us, iterations {st['iterations']} trials {st['trials']} chi2 {st['chi2']:.9g}")
po.close()
