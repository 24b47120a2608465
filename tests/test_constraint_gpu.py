"""GPU parity of svs_computeConstraint_batch against oracle/constraint_oracle.c through the C ABI:
visibility counts identical, relative poses and Lambda to 1e-12 relative."""
import numpy as np
import pytest

from scavislam_b200 import synth, synth_graph

pytestmark = pytest.mark.gpu


def _check(svs, oracle, g, v1, v2):
    cb = svs.ConstraintBuilder()
    T_g, L_g, n_g = cb.compute(g["poses"], g["feat_ptr"], g["feat_point"], g["point_anchor"], g["xyz_anchor"], v1, v2)
    cb.close()
    T_o, L_o, n_o = oracle.compute_constraints(g["poses"], g["feat_ptr"], g["feat_point"], g["point_anchor"], g["xyz_anchor"], v1, v2)
    np.testing.assert_array_equal(n_g, n_o)
    np.testing.assert_allclose(T_g, T_o, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(L_g, L_o, rtol=1e-12)
    return L_g, n_g


def test_window_constraints_match_oracle(svs, oracle):
    pb = synth.make_config("C2")
    g = synth_graph.graph_tables(pb)
    L, n = _check(svs, oracle, g, pb.c_i, pb.c_j)          # all 1 020 constraint pairs of the 200 KF window
    assert n.min() > 0
    # the window generator wrote the same weights from ground truth; here they come from the noisy estimate
    assert np.all(np.abs(np.log(L[:, 5, 5] / pb.c_Lambda.reshape(-1, 6, 6)[:, 5, 5])) < 1e-9)


def test_pairs_sharing_thousands_of_points_and_none(svs, oracle):
    # two frames that see the same 3 000 points (> the shared-memory list), one that sees nothing in common
    rng = np.random.default_rng(5)
    L = 3000
    xyz = np.stack([rng.uniform(-2, 2, L), rng.uniform(-1, 1, L), rng.uniform(2, 20, L)], 1)
    xyz[:50, 2] = 5.0; xyz[:50, :2] = 0.0                  # ties in the depth multiset
    poses = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (3, 1))
    poses[1] = oracle.se3_exp(np.array([0.3, 0.0, 0.1, 0.0, 0.05, 0.0]))
    poses[2, 4] = 9.0
    feat_ptr = np.array([0, L, 2 * L - 1, 2 * L - 1], np.int32)
    feat_point = np.concatenate([np.arange(L), np.arange(L - 1)]).astype(np.int32)   # pose 1 misses the last point (even/odd)
    g = dict(poses=poses, feat_ptr=feat_ptr, feat_point=feat_point, point_anchor=np.zeros(L, np.int32), xyz_anchor=xyz)
    L_g, n = _check(svs, oracle, g, [0, 1, 0, 2], [1, 0, 2, 1])
    assert list(n) == [L - 1, L - 1, 0, 0] and not L_g[2].any()


def test_bad_tables_are_rejected(svs):
    cb = svs.ConstraintBuilder()
    poses = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (2, 1))
    with pytest.raises(svs.SvsError):      # feature table not ascending
        cb.compute(poses, [0, 2, 2], [1, 0], [0, 0], np.zeros((2, 3)), [0], [1])
    with pytest.raises(svs.SvsError):      # pair names a pose that does not exist
        cb.compute(poses, [0, 1, 2], [0, 0], [0], np.ones((1, 3)), [0], [5])
    cb.close()
