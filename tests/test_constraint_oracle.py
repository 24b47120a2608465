"""CPU tests pinning oracle/constraint_oracle.c (SlamGraph::computeConstraint, reference slam_graph.cpp:785-846)
with an independent numpy computation; the reference ships no test for it (PARITY UNPINNED)."""
import numpy as np

from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph


def numpy_constraint(g, i, j):
    def Rt(T):
        x, y, z, w = T[:4]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return R, T[4:]
    R1, t1 = Rt(g["poses"][i]); R2, t2 = Rt(g["poses"][j])
    R12, t12 = R1 @ R2.T, t1 - R1 @ R2.T @ t2
    a = g["feat_point"][g["feat_ptr"][i]:g["feat_ptr"][i + 1]]
    b = g["feat_point"][g["feat_ptr"][j]:g["feat_ptr"][j + 1]]
    shared = np.intersect1d(a, b)
    d = []
    for p in shared:
        Ra, ta = Rt(g["poses"][g["point_anchor"][p]])
        xw = Ra.T @ (g["xyz_anchor"][p] - ta)
        d.append(np.linalg.norm(R1 @ xw + t1))
    n = len(d)
    lam = np.zeros((6, 6))
    if n:
        med = np.median(d)     # numpy's median = the multiset median of maths_utils.h:113-136
        lam[:3, :3] = np.eye(3) * n * (350 * np.linalg.norm(t12) / med) ** 2
        lam[3:, 3:] = np.eye(3) * n * 100.0 ** 2
    return R12, t12, lam, n


def test_matches_numpy_on_a_window():
    pb = synth.make_window(14, 600, seed=3)
    g = synth_graph.graph_tables(pb)
    v1, v2 = pb.c_i[:40], pb.c_j[:40]
    T, Lam, ns = po.compute_constraints(g["poses"], g["feat_ptr"], g["feat_point"], g["point_anchor"], g["xyz_anchor"], v1, v2)
    for k in range(len(v1)):
        R12, t12, lam, n = numpy_constraint(g, v1[k], v2[k])
        assert ns[k] == n and n > 0
        np.testing.assert_allclose(T[k, 4:], t12, atol=1e-12)
        np.testing.assert_allclose(po.se3_act(T[k], np.array([0.3, -0.2, 1.0])), R12 @ np.array([0.3, -0.2, 1.0]) + t12, atol=1e-12)
        np.testing.assert_allclose(Lam[k], lam, rtol=1e-10)


def test_even_odd_counts_and_disjoint_pairs():
    # hand-made tables: pose 0 and 1 share 4 points (even: mean of the two middle depths), 0 and 2 share 3, 1 and 3 none
    poses = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (4, 1))
    poses[1, 4] = 0.5; poses[2, 5] = -0.25; poses[3, 6] = 1.0
    xyz = np.array([[0, 0, 1.0], [0, 0, 2.0], [0, 0, 4.0], [0, 0, 8.0], [1, 1, 1.0]])
    anchor = np.zeros(5, np.int32)
    feat_ptr = np.array([0, 5, 9, 12, 12], np.int32)
    feat_point = np.array([0, 1, 2, 3, 4, 0, 1, 2, 3, 1, 2, 4], np.int32)
    T, Lam, ns = po.compute_constraints(poses, feat_ptr, feat_point, anchor, xyz, [0, 0, 1, 1], [1, 2, 3, 0])
    assert list(ns) == [4, 3, 0, 4]
    assert np.isclose(Lam[0][0, 0], 4 * (350 * 0.5 / 3.0) ** 2)          # median of {1, 2, 4, 8} = 3
    d = sorted([2.0, 4.0, np.sqrt(3.0)])                                   # points 1, 2, 4 seen from pose 0
    assert np.isclose(Lam[1][0, 0], 3 * (350 * 0.25 / d[1]) ** 2) and np.isclose(Lam[1][5, 5], 3 * 100.0 ** 2)
    assert not Lam[2].any()                                                # no shared points
    # frame 1 sees the same points half a metre to the side: depths change, the count does not
    dd = np.sort(np.linalg.norm(xyz[:4] + np.array([0.5, 0, 0]), axis=1))
    assert np.isclose(Lam[3][0, 0], 4 * (350 * 0.5 / (0.5 * (dd[1] + dd[2]))) ** 2)
