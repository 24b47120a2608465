"""GPU parity tests of the BA path: CUDA (through the C ABI) vs the CPU oracle.

Tolerances: north_star asks 1e-6 relative on pose parameters; intermediate
quantities (chi2, reduced system, solve) are checked much tighter because both
sides are FP64 and differ only in summation order.
"""
import numpy as np
import pytest

from scavislam_b200 import synth

pytestmark = pytest.mark.gpu

POSE_RTOL = 1e-6


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def c1():
    return synth.make_config("C1")


@pytest.fixture(scope="module")
def ba(svs):
    b = svs.BundleAdjuster()
    yield b
    b.close()


def test_device_is_blackwell(svs):
    info = svs.device_info()
    assert "sm_100" in info, info


def test_chi2_matches_oracle(ba, oracle, c1):
    ba.set_problem(c1)
    for robust in (True, False):
        g = ba.chi2(robust, 1.0)
        o = oracle.chi2(c1, robust, 1.0)
        assert abs(g - o) <= 1e-11 * abs(o), (robust, g, o)


@pytest.mark.parametrize("lam", [50.0, 1e-3])
def test_reduced_system_matches_oracle(ba, oracle, c1, lam):
    ba.set_problem(c1)
    S, bs, chi = ba.reduced_system(True, 1.0, lam)
    So, bso, chio = oracle.reduced_system(c1, True, 1.0, lam)
    assert abs(chi - chio) <= 1e-11 * abs(chio)
    assert _rel(S, So) < 1e-11
    assert _rel(bs, bso) < 1e-10
    assert np.abs(S - S.T).max() <= 1e-9 * np.abs(S).max()


def test_reduced_solve_matches_numpy(ba, c1):
    ba.set_problem(c1)
    S, bs, _ = ba.reduced_system(True, 1.0, 50.0)
    x, failed = ba.solve_reduced(True, 1.0, 50.0)
    assert failed == 0
    xr = np.linalg.solve(S, bs)
    assert _rel(x, xr) < 1e-9


def test_one_iteration_c1(ba, oracle, c1):
    """BASELINE config C1: 10 KF / 500 pt, single GN iteration."""
    ba.set_problem(c1)
    it, st = ba.optimize(1)
    po, pso, sto = oracle.optimize(c1, 1)
    assert it == sto["iterations"] == 1
    assert st["trials_iter"] == sto["trials_iter"]
    assert abs(st["chi2_init"] - sto["chi2_init"]) <= 1e-10 * sto["chi2_init"]
    assert abs(st["chi2_final"] - sto["chi2_final"]) <= 1e-8 * sto["chi2_final"]
    assert _rel(ba.poses(), po) < POSE_RTOL
    assert _rel(ba.points(), pso) < POSE_RTOL


@pytest.mark.parametrize("iters,lam0", [(5, 50.0), (10, 1e-4), (6, 1e5)])
def test_multi_iteration_c1(ba, oracle, c1, iters, lam0):
    ba.set_problem(c1)
    it, st = ba.optimize(iters, True, 1.0, lam0, 5)
    po, pso, sto = oracle.optimize(c1, iters, True, 1.0, lam0, 5)
    assert it == sto["iterations"]
    assert st["trials_iter"] == sto["trials_iter"]
    np.testing.assert_allclose(st["chi2_iter"], sto["chi2_iter"], rtol=1e-7)
    np.testing.assert_allclose(st["lambda_iter"], sto["lambda_iter"], rtol=1e-6)
    assert _rel(ba.poses(), po) < POSE_RTOL
    assert _rel(ba.points(), pso) < POSE_RTOL


def test_rejected_trials_path(ba, oracle):
    """Large perturbation + small lambda0: an iteration with 6 rejected Levenberg trials."""
    pb = synth.make_window(12, 400, seed=100, pose_noise=(0.5, 0.15), depth_noise=0.6)
    ba.set_problem(pb)
    it, st = ba.optimize(8, False, 1.0, 1e-2, 10)
    po, pso, sto = oracle.optimize(pb, 8, False, 1.0, 1e-2, 10)
    assert sum(sto["trials_iter"]) > sto["iterations"], "case does not exercise rejections"
    assert it == sto["iterations"]
    assert st["trials_iter"] == sto["trials_iter"]
    np.testing.assert_allclose(st["chi2_iter"], sto["chi2_iter"], rtol=1e-6)
    np.testing.assert_allclose(st["lambda_iter"], sto["lambda_iter"], rtol=1e-6)
    assert _rel(ba.poses(), po) < POSE_RTOL


def test_non_robust_and_fixed_pose(ba, oracle, c1):
    pb = c1.copy()
    pb.fixed[0] = 1
    pb.fixed[4] = 1
    ba.set_problem(pb)
    it, st = ba.optimize(3, False, 1.0, 50.0, 5)
    po, pso, sto = oracle.optimize(pb, 3, False, 1.0, 50.0, 5)
    assert it == sto["iterations"]
    np.testing.assert_allclose(st["chi2_iter"], sto["chi2_iter"], rtol=1e-7)
    g = ba.poses()
    assert np.array_equal(g[0], pb.pose_qt[0]) and np.array_equal(g[4], pb.pose_qt[4])
    assert _rel(g, po) < POSE_RTOL


def test_one_call_api_and_reset(ba, oracle, c1):
    it, poses, psi, st = ba.optimise_inner_and_outer_window(c1, 2)
    po, pso, sto = oracle.optimize(c1, 2)
    assert it == 2
    assert _rel(poses, po) < POSE_RTOL and _rel(psi, pso) < POSE_RTOL
    ba.reset_state()
    assert np.array_equal(ba.poses(), c1.pose_qt)
    assert np.array_equal(ba.points(), c1.psi)
    it2, st2 = ba.optimize(2)
    assert _rel(ba.poses(), po) < POSE_RTOL


def test_empty_and_degenerate(ba, svs):
    pb = synth.make_window(4, 0, seed=5)   # no landmarks, no edges
    ba.set_problem(pb)
    it, st = ba.optimize(2)
    assert it == 1          # rho == 0 -> Terminate after the first iteration, like g2o
    assert np.allclose(ba.poses(), pb.pose_qt)
    pb0 = synth.make_window(0, 0, seed=5)
    ba.set_problem(pb0)
    it, _ = ba.optimize(2)
    assert it == -1   # g2o: 0 vertices to optimize


def test_invalid_inputs_rejected(ba, svs, c1):
    pb = c1.copy()
    pb.e_pose[3] = 10_000
    with pytest.raises(svs.SvsError):
        ba.set_problem(pb)
    pb = c1.copy()
    pb.e_anchor[0] = (pb.e_anchor[0] + 1) % pb.P   # two anchors for one point
    with pytest.raises(svs.SvsError):
        ba.set_problem(pb)


def test_c2_full_size_parity(ba, oracle):
    """BASELINE config C2: 200 KF / 20k pt, 10 iterations (oracle takes ~0.5 s)."""
    pb = synth.make_config("C2")
    ba.set_problem(pb)
    it, st = ba.optimize(10)
    po, pso, sto = oracle.optimize(pb, 10)
    assert it == sto["iterations"] == 10
    assert st["trials_iter"] == sto["trials_iter"]
    np.testing.assert_allclose(st["chi2_iter"], sto["chi2_iter"], rtol=1e-7)
    assert _rel(ba.poses(), po) < POSE_RTOL
    assert _rel(ba.points(), pso) < POSE_RTOL
    # size-independent property: chi2 decreases monotonically over accepted iterations
    assert all(a >= b for a, b in zip([st["chi2_init"]] + st["chi2_iter"][:-1], st["chi2_iter"]))
