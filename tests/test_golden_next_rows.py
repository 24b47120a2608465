"""The oracles of the rows added after the BA core against their committed outputs
(tests/golden/next_rows_golden.npz, scripts/make_golden_next.py).  These are oracle outputs, not reference
outputs -- the reference has no golden vectors (PARITY UNPINNED) -- so the file guards against drift only."""
import os

import numpy as np

from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph, synth_pose
from test_dtc_oracle import levels

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows_golden.npz"))
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


def test_motion_only_lm():
    tr = synth_pose.make_track(300, seed=21, outlier_frac=0.1)
    T, st = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], True, 2.0, 15)
    np.testing.assert_allclose(T, G["pose_T"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose([st["chi2"], st["initial_chi2"]], [G["pose_chi2"], G["pose_initial_chi2"]], rtol=1e-10)
    assert st["num_obs"] == G["pose_counts"][2]


def test_dense_tracker_non_cuda_build():
    lv, _, _ = levels(po, 3)
    np.testing.assert_array_equal(lv[2]["cloud"], G["dtc_cloud_l2"])
    chi, H, b, n = po.dtc_pass(lv[1], I7)
    assert n == int(G["dtc_pass_n"])
    np.testing.assert_allclose(chi, G["dtc_pass_chi2"], rtol=1e-10)
    np.testing.assert_allclose(H, G["dtc_pass_H"], rtol=1e-9, atol=1e-9 * np.abs(G["dtc_pass_H"]).max())
    np.testing.assert_allclose(b, G["dtc_pass_b"], rtol=1e-9, atol=1e-9 * np.abs(G["dtc_pass_b"]).max())
    T, st = po.dtc_track(lv, I7)
    np.testing.assert_allclose(T, G["dtc_T"], rtol=1e-8, atol=1e-10)
    assert st["passes"] == list(G["dtc_passes"])


def test_compute_constraint():
    pb = synth.make_window(14, 600, seed=3)
    g = synth_graph.graph_tables(pb)
    T12, Lam, ns = po.compute_constraints(g["poses"], g["feat_ptr"], g["feat_point"], g["point_anchor"], g["xyz_anchor"],
                                          pb.c_i[:40], pb.c_j[:40])
    np.testing.assert_array_equal(ns, G["con_n"])
    np.testing.assert_allclose(T12, G["con_T"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(np.array([np.diag(L) for L in Lam]), G["con_Lambda_diag"], rtol=1e-12)
