"""Generates tests/golden/ba_golden.npz from the CPU oracle on the seeded C1 window.
(The reference has no golden vectors and cannot be built here; see DESIGN.md.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scavislam_b200 import synth
from oracle import pyoracle as po

pb = synth.make_config("C1")
p1, s1, st1 = po.optimize(pb, 1)
p5, s5, st5 = po.optimize(pb, 5)
S, bs, chi = po.reduced_system(pb, True, 1.0, 50.0)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ba_golden.npz")
np.savez_compressed(out, E=pb.E, C=pb.C, e_obs=pb.e_obs, poses_it1=p1, psi_it1=s1, poses_it5=p5, psi_it5=s5,
                    chi2_init=st5["chi2_init"], chi2_iter5=np.array(st5["chi2_iter"]),
                    lambda_iter5=np.array(st5["lambda_iter"]), bs_lambda50=bs, S_diag_lambda50=np.diag(S).copy())
print("wrote", out, os.path.getsize(out), "bytes")
