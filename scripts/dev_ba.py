"""Developer probe (GPU box): rejection-path cases, C2/C5 timing breakdown."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scavislam_b200 import synth, capi
from oracle import pyoracle as po

ba = capi.BundleAdjuster()
print(capi.device_info())
def rel(a, b): return np.abs(a - b).max() / np.abs(b).max()
cases = [(100, (0.8, 0.25), 0.8, True, 1e-3, 10), (101, (0.8, 0.25), 0.8, True, 1e-3, 10),
         (100, (1.0, 0.3), 0.8, False, 10., 10), (101, (0.8, 0.25), 0.8, False, 1.0, 10),
         (100, (0.5, 0.15), 0.6, False, 1e-2, 10), (101, (1.0, 0.3), 0.8, False, 10., 10)]
for seed, pn, dn, rob, lam, mt in cases:
    pb = synth.make_window(12, 400, seed=seed, pose_noise=pn, depth_noise=dn)
    ba.set_problem(pb)
    it, st = ba.optimize(8, rob, 1.0, lam, mt)
    p_o, s_o, sto = po.optimize(pb, 8, rob, 1.0, lam, mt)
    print("case", seed, pn, dn, rob, lam, "gpu", st["trials_iter"], "cpu", sto["trials_iter"],
          "pose rel", rel(ba.poses(), p_o))
    print("   chi gpu", ["%.6g" % c for c in st["chi2_iter"]])
    print("   chi cpu", ["%.6g" % c for c in sto["chi2_iter"]])
for name in ("C2", "C5"):
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
