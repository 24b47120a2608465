"""CPU tests pinning oracle/pose_oracle.c (motion-only LM).  The reference has no test or golden vector
for calcFastMotionOnly (PARITY UNPINNED); the restatement is pinned by finite differences of its own
projection, an independent numpy Levenberg-Marquardt, and recovery of the generating pose."""
import numpy as np

from oracle import pyoracle as po
from scavislam_b200 import synth_pose as sp


def test_frame_jacobian_is_minus_the_derivative_of_map_under_left_update():
    rng = np.random.default_rng(0)
    cam = np.array(sp.CAM)
    for _ in range(5):
        T = po.se3_exp(rng.normal(0, 0.2, 6))
        X = np.array([rng.normal(0, 1), rng.normal(0, 1), rng.uniform(2, 8)])
        J = po.pose_frame_jac(cam, T, X)
        num = np.zeros((3, 6))
        h = 1e-6
        for k in range(6):
            d = np.zeros(6); d[k] = h
            num[:, k] = (po.pose_map(cam, po.se3_mul(po.se3_exp(d), T), X) - po.pose_map(cam, po.se3_mul(po.se3_exp(-d), T), X)) / (2 * h)
        np.testing.assert_allclose(J, -num, rtol=2e-6, atol=2e-6)     # frameJac = d(obs - map)/d(delta)


def _numpy_lm(tr, robust, kparam, num_iter):
    """Independent restatement of the published loop (pose_optimizer.h:135-298) on numpy arrays."""
    cam, T = tr["cam"], tr["T_init"].copy()
    def resid(T):
        f = np.array([tr["obs"][i] - po.pose_map(cam, T, tr["xyz"][tr["pid"][i]]) for i in range(len(tr["pid"]))])
        if robust:
            nrm = np.maximum(1e-10, np.linalg.norm(f, axis=1))
            k = np.where(nrm < kparam, nrm * nrm, 2 * kparam * nrm - kparam * kparam)
            f = f * (np.sqrt(k) / nrm)[:, None]
        return f
    def jac(T):
        return np.array([po.pose_frame_jac(cam, T, tr["xyz"][p]) for p in tr["pid"]])
    f = resid(T); J = jac(T)
    chi2 = (f * f).sum()
    mu = 1e-5 * max((J[:, :, c] ** 2).sum(1).max() for c in range(6))
    nu, stop, trial = 2.0, False, 0
    for _ in range(num_iter):
        while True:
            f = resid(T); J = jac(T)
            A = mu * np.eye(6) + np.einsum("nij,nik->jk", J, J)
            B = -np.einsum("nij,ni->j", J, f)
            Tn = po.se3_mul(po.se3_exp(np.linalg.solve(A, B)), T)
            fn = resid(Tn)
            rho = chi2 - (fn * fn).sum()
            if rho > 0:
                T, chi2 = Tn, (fn * fn).sum()
                stop = np.abs(B).max() <= 1e-10
                mu *= max(1 / 3, 1 - (2 * rho - 1) ** 3); nu = 2.0; trial = 0
            else:
                mu *= nu; nu *= 2; trial += 1
                stop = stop or trial == 5
            if rho > 0 or stop:
                break
        if stop:
            break
    return T, chi2


def test_oracle_matches_an_independent_numpy_lm():
    for seed, robust, out in ((1, True, 0.1), (2, False, 0.0), (3, True, 0.0)):
        tr = sp.make_track(150, seed=seed, outlier_frac=out, shared_points=(seed == 3))
        T_o, st = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], robust, 2.0, 15)
        T_n, chi_n = _numpy_lm(tr, robust, 2.0, 15)
        np.testing.assert_allclose(T_o, T_n, rtol=0, atol=1e-9)
        assert abs(st["chi2"] - chi_n) <= 1e-9 * chi_n
        assert st["num_obs"] == 150 and st["chi2"] < st["initial_chi2"]


def test_oracle_recovers_the_generating_pose_despite_outliers():
    tr = sp.make_track(600, seed=7, pixel_noise=0.2, outlier_frac=0.15)
    T, st = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], True, 2.0, 25)
    assert np.abs(T[4:] - tr["T_true"][4:]).max() < 5e-3
    assert np.abs(T[:4] - tr["T_true"][:4]).max() < 2e-3
    # without the kernel the outliers pull the estimate away
    T2, _ = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], False, 2.0, 25)
    assert np.abs(T2[4:] - tr["T_true"][4:]).max() > np.abs(T[4:] - tr["T_true"][4:]).max()


def test_zero_iterations_only_reports_the_initial_cost():
    tr = sp.make_track(50, seed=9)
    T, st = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], True, 2.0, 0)
    assert np.array_equal(T, tr["T_init"]) and st["chi2"] == st["initial_chi2"] and st["trials"] == 0
