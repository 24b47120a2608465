"""Generates tests/golden/next_rows_golden.npz from the CPU oracles of the rows added after the BA core:
motion-only LM (pose_oracle.c), the non-CUDA-build dense tracker (dt_oracle.c: odtc_*), computeConstraint
(constraint_oracle.c).  The reference has no golden vectors and cannot be built here; these
files pin the oracles against silent drift, and the GPU parity tests compare against the same oracles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph, synth_pose
from test_dtc_oracle import levels

I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])
tr = synth_pose.make_track(300, seed=21, outlier_frac=0.1)
T_pose, st_pose = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], True, 2.0, 15)
lv, seq, cams = levels(po, 3)
T_dtc, st_dtc = po.dtc_track(lv, I7)
chi_dtc, H_dtc, b_dtc, n_dtc = po.dtc_pass(lv[1], I7)
pb = synth.make_window(14, 600, seed=3)
g = synth_graph.graph_tables(pb)
T12, Lam, ns = po.compute_constraints(g["poses"], g["feat_ptr"], g["feat_point"], g["point_anchor"], g["xyz_anchor"],
                                      pb.c_i[:40], pb.c_j[:40])
out = os.path.join(ROOT, "tests", "golden", "next_rows_golden.npz")
np.savez_compressed(out, pose_T=T_pose, pose_chi2=st_pose["chi2"], pose_initial_chi2=st_pose["initial_chi2"],
                    pose_counts=np.array([st_pose["iterations"], st_pose["trials"], st_pose["num_obs"]]),
                    dtc_T=T_dtc, dtc_chi2=np.array(st_dtc["chi2"]), dtc_passes=np.array(st_dtc["passes"]),
                    dtc_pass_chi2=chi_dtc, dtc_pass_H=H_dtc, dtc_pass_b=b_dtc, dtc_pass_n=n_dtc,
                    dtc_cloud_l2=lv[2]["cloud"], con_T=T12, con_Lambda_diag=np.array([np.diag(L) for L in Lam]), con_n=ns)
print("wrote", out, os.path.getsize(out), "bytes")
