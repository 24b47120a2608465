"""Developer probe (GPU box): C2/C5 timing breakdown of the BA kernels + k_solve phase cycles."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import synth, capi
from oracle import pyoracle as po

os.environ.setdefault("SVS_SOLVE_TIMING", "1")
ba = capi.BundleAdjuster()
print(capi.device_info())
def rel(a, b): return np.abs(a - b).max() / np.abs(b).max()
for name in sys.argv[1:] or ("C2", "C5"):
    pb = synth.make_config(name)
    t = time.time(); ba.set_problem(pb); t_set = time.time() - t
    for rep in range(3):
        ba.reset_state()
        t = time.time(); it, st = ba.optimize(10); dt = time.time() - t
    print(name, "P L E C", pb.P, pb.L, pb.E, pb.C, "set_problem s", t_set, "optimize wall s", dt, "iters", it)
    print("   ", {k: st[k] for k in ("ms_total", "ms_build", "ms_solve", "ms_update", "ms_control", "launches", "nnzb_S", "nnzb_L", "max_track", "trials_total")})
    t = time.time(); p_o, s_o, sto = po.optimize(pb, 10); dto = time.time() - t
    print("    oracle s", dto, "it/s", sto["iterations"] / dto, "pose rel", rel(ba.poses(), p_o), "psi rel", rel(ba.points(), s_o))
    print("    chi gpu", st["chi2_iter"][-1], "cpu", sto["chi2_iter"][-1])
