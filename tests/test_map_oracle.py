"""CPU test of the window-assembly restatement (oracle.pyoracle.copy_data_to_g2o): assembling the window out of a
larger map gives back the window the map was made from (same edges, up to the reference's edge order)."""
import numpy as np

from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph


def test_assembly_returns_the_window_the_map_was_built_from():
    pb = synth.make_window(8, 200, seed=5)
    m, win, act = synth_graph.make_map(pb)
    g = po.copy_data_to_g2o(m, win, act)
    assert len(g["e_point"]) == pb.E                     # observations from outside frames / inactive points dropped
    np.testing.assert_array_equal(g["pose_qt"], pb.pose_qt)
    np.testing.assert_allclose(g["psi"], pb.psi, rtol=1e-14)
    key = lambda ep, es: np.lexsort((es, ep))
    o1, o2 = key(g["e_point"], g["e_pose"]), key(pb.e_point, pb.e_pose)
    np.testing.assert_array_equal(g["e_pose"][o1], pb.e_pose[o2])
    np.testing.assert_array_equal(g["e_anchor"][o1], pb.e_anchor[o2])
    np.testing.assert_array_equal(g["e_obs"][o1], pb.e_obs[o2])
    np.testing.assert_allclose(g["e_info"][o1], pb.e_info[o2], rtol=1e-12)
    # inside a point the edges follow ascending vertex ids (vis_set order), not ascending BA pose index
    first = g["e_pose"][g["e_point"] == 0]
    assert list(win[first]) == sorted(win[first])
