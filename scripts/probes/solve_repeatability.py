"""Developer probe (GPU box): solve_reduced vs numpy on a mid-size banded window, several repetitions (race hunting)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import synth, capi
P = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
pb = synth.make_window(P, L, seed=1235)
ba = capi.BundleAdjuster()
ba.set_problem(pb)
S, bs, _ = ba.reduced_system(True, 1.0, 50.0)
xr = np.linalg.solve(S, bs)
for rep in range(6):
    x, failed = ba.solve_reduced(True, 1.0, 50.0)
    err = np.abs(x - xr).max() / np.abs(xr).max()
    bad = np.nonzero(np.abs(x - xr).reshape(-1, 6).max(1) > 1e-9 * np.abs(xr).max())[0]
    print("rep", rep, "failed", failed, "rel err %.3e" % err, "bad poses", bad[:12], "..." if len(bad) > 12 else "", len(bad))
