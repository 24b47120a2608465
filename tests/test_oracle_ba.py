"""CPU tests pinning the BA oracle (oracle/ba_oracle.c) with independent checks:
finite differences, a numpy Schur complement, and committed golden vectors.

The reference ships no tests or golden vectors for this path, so these
are the secondary pins: invariants derived from the reference code itself.
"""
import os

import numpy as np
import pytest

from scavislam_b200 import synth

CAM = np.array([synth.CAM_F, synth.CAM_PX, synth.CAM_PY, synth.CAM_B])
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ba_golden.npz")


def rand_pose(rng, scale=0.3):
    from oracle import pyoracle as po
    return po.se3_exp(rng.normal(0, scale, 6))


def test_se3_exp_log_roundtrip(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        d = rng.normal(0, 0.5, 6)
        T = oracle.se3_exp(d)
        assert abs(np.linalg.norm(T[:4]) - 1) < 1e-14
        np.testing.assert_allclose(oracle.se3_log(T), d, atol=1e-12)
    # tiny rotation branch (theta < 1e-10)
    d = np.array([0.1, -0.2, 0.3, 1e-12, 0, 0])
    np.testing.assert_allclose(oracle.se3_log(oracle.se3_exp(d)), d, atol=1e-12)


def test_se3_group_laws(oracle):
    rng = np.random.default_rng(1)
    A, B = rand_pose(rng), rand_pose(rng)
    x = rng.normal(0, 1, 3)
    np.testing.assert_allclose(oracle.se3_act(oracle.se3_mul(A, B), x), oracle.se3_act(A, oracle.se3_act(B, x)), atol=1e-13)
    I = oracle.se3_mul(A, oracle.se3_inv(A))
    np.testing.assert_allclose(I, [0, 0, 0, 1, 0, 0, 0], atol=1e-14)
    # Adj: A exp(d) A^-1 = exp(Adj d)
    d = rng.normal(0, 0.1, 6)
    lhs = oracle.se3_mul(oracle.se3_mul(A, oracle.se3_exp(d)), oracle.se3_inv(A))
    rhs = oracle.se3_exp(oracle.se3_adj(A) @ d)
    np.testing.assert_allclose(lhs, rhs, atol=1e-12)


def test_invert_depth_involution(oracle):
    # maths_utils.h:66: invert_depth(invert_depth(x)) == x
    from oracle import pyoracle as po
    import ctypes as C
    x = np.array([0.3, -0.2, 4.0])
    y, z = np.zeros(3), np.zeros(3)
    po.lib().oba_invert_depth(x.ctypes.data_as(po.c_dp), y.ctypes.data_as(po.c_dp))
    po.lib().oba_invert_depth(y.ctypes.data_as(po.c_dp), z.ctypes.data_as(po.c_dp))
    np.testing.assert_allclose(z, x, rtol=1e-15)


def _fd_edge(oracle, Tp, Ta, psi, obs, h=1e-6):
    """Numeric Jacobians of the error wrt psi (additive) and both poses (T <- exp(d) T)."""
    def err(Tp_, Ta_, psi_):
        return oracle.edge_error(CAM, Tp_, Ta_, psi_, obs)
    Jpsi, Jp, Ja = np.zeros((3, 3)), np.zeros((3, 6)), np.zeros((3, 6))
    for k in range(3):
        d = np.zeros(3); d[k] = h
        Jpsi[:, k] = (err(Tp, Ta, psi + d) - err(Tp, Ta, psi - d)) / (2 * h)
    for k in range(6):
        d = np.zeros(6); d[k] = h
        Jp[:, k] = (err(oracle.se3_mul(oracle.se3_exp(d), Tp), Ta, psi) -
                    err(oracle.se3_mul(oracle.se3_exp(-d), Tp), Ta, psi)) / (2 * h)
        Ja[:, k] = (err(Tp, oracle.se3_mul(oracle.se3_exp(d), Ta), psi) -
                    err(Tp, oracle.se3_mul(oracle.se3_exp(-d), Ta), psi)) / (2 * h)
    return Jpsi, Jp, Ja


def test_edge_jacobians_vs_finite_differences(oracle):
    """linearizeOplus (anchored_points.cpp:168-189) against numeric differentiation of
    computeError (:148-166) -- the check the reference's numeric-Jacobian defaults
    (transformations.h:320-383) make implicitly."""
    rng = np.random.default_rng(2)
    for _ in range(20):
        Tp, Ta = rand_pose(rng), rand_pose(rng)
        z = rng.uniform(2, 15)
        psi = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), 1 / z])
        obs = rng.uniform(0, 400, 3)
        Jpsi, Jp, Ja = oracle.edge_jacobians(CAM, Tp, Ta, psi)
        Fpsi, Fp, Fa = _fd_edge(oracle, Tp, Ta, psi, obs)
        for A, F in ((Jpsi, Fpsi), (Jp, Fp), (Ja, Fa)):
            assert np.abs(A - F).max() <= 1e-5 * max(1.0, np.abs(F).max())


def test_self_anchor_edge_jacobians_cancel(oracle):
    # pose == anchor  =>  J_pose = -J_anchor (anchored_points.cpp:187-188 with T_ca = I)
    rng = np.random.default_rng(3)
    T = rand_pose(rng)
    psi = np.array([0.1, -0.05, 0.2])
    _, Jp, Ja = oracle.edge_jacobians(CAM, T, T, psi)
    np.testing.assert_allclose(Jp, -Ja, atol=1e-9 * np.abs(Jp).max())


def test_posepose_jacobians_vs_finite_differences(oracle):
    rng = np.random.default_rng(4)
    T1, T2 = rand_pose(rng), rand_pose(rng)
    # measurement close to the true relative pose so that the BCH truncation in third() is accurate
    T21 = oracle.se3_mul(oracle.se3_exp(rng.normal(0, 0.01, 6)), oracle.se3_mul(T2, oracle.se3_inv(T1)))
    e = oracle.posepose_error(T21, T1, T2)
    Ji, Jj = oracle.posepose_jacobians(T21, e)
    h = 1e-6
    Fi, Fj = np.zeros((6, 6)), np.zeros((6, 6))
    for k in range(6):
        d = np.zeros(6); d[k] = h
        Fi[:, k] = (oracle.posepose_error(T21, oracle.se3_mul(oracle.se3_exp(d), T1), T2) -
                    oracle.posepose_error(T21, oracle.se3_mul(oracle.se3_exp(-d), T1), T2)) / (2 * h)
        Fj[:, k] = (oracle.posepose_error(T21, T1, oracle.se3_mul(oracle.se3_exp(d), T2)) -
                    oracle.posepose_error(T21, T1, oracle.se3_mul(oracle.se3_exp(-d), T2))) / (2 * h)
    # third() is a 2nd-order BCH approximation: agreement to O(|e|^3)
    assert np.abs(Ji - Fi).max() < 5e-3 and np.abs(Jj - Fj).max() < 5e-3


def _small_problem(seed=7, self_edges=True):
    pb = synth.make_window(6, 40, seed=seed)
    if not self_edges:
        keep = pb.e_pose != pb.e_anchor
        for k in ("e_point", "e_pose", "e_anchor", "e_obs", "e_info"):
            setattr(pb, k, np.ascontiguousarray(getattr(pb, k)[keep]))
        pb.E = int(keep.sum())
    return pb


def test_full_system_is_gauss_newton_of_fd_residuals(oracle):
    """H = J^T W J and b = -J^T W e with J from finite differences of the stacked residual
    (no self-anchor edges, so g2o's duplicate-vertex quirk does not enter)."""
    pb = _small_problem(self_edges=False)
    pb.C = 0
    pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    H, b, chi = oracle.full_system(pb, robust=False)
    P, L, E = pb.P, pb.L, pb.E
    n = 6 * P + 3 * L

    def residual(dx):
        r = np.zeros(3 * E)
        poses = np.array([oracle.se3_mul(oracle.se3_exp(dx[6 * i:6 * i + 6]), pb.pose_qt[i]) for i in range(P)])
        psi = pb.psi + dx[6 * P:].reshape(L, 3)
        for e in range(E):
            r[3 * e:3 * e + 3] = oracle.edge_error(CAM, poses[pb.e_pose[e]], poses[pb.e_anchor[e]], psi[pb.e_point[e]],
                                                   pb.e_obs[e])
        return r
    r0 = residual(np.zeros(n))
    J = np.zeros((3 * E, n))
    h = 1e-6
    for k in range(n):
        d = np.zeros(n); d[k] = h
        J[:, k] = (residual(d) - residual(-d)) / (2 * h)
    W = pb.e_info.reshape(-1)
    Hfd = J.T @ (W[:, None] * J)
    bfd = -J.T @ (W * r0)
    assert abs(chi - r0 @ (W * r0)) <= 1e-10 * chi
    assert np.abs(H - Hfd).max() <= 2e-5 * np.abs(Hfd).max()
    assert np.abs(b - bfd).max() <= 2e-5 * np.abs(bfd).max()


def test_self_anchor_quirk_adds_J1WJ1_to_anchor_pose(oracle):
    """an observation made in the anchor frame itself leaves +J1^T W J1 on the
    anchor pose's diagonal block and nothing in Hpl / b for that pose."""
    pb = _small_problem(self_edges=True)
    pb.C = 0
    pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    H, b, _ = oracle.full_system(pb, robust=False)
    pb2 = _small_problem(self_edges=False)
    pb2.C = 0
    pb2.c_i = pb2.c_i[:0]; pb2.c_j = pb2.c_j[:0]; pb2.c_T = pb2.c_T[:0]; pb2.c_Lambda = pb2.c_Lambda[:0]
    H2, b2, _ = oracle.full_system(pb2, robust=False)
    dH = H - H2
    P = pb.P
    expect = np.zeros_like(dH)
    bexp = np.zeros_like(b)
    for e in np.nonzero(pb.e_pose == pb.e_anchor)[0]:
        a, l = pb.e_pose[e], pb.e_point[e]
        Jpsi, Jp, _ = oracle.edge_jacobians(CAM, pb.pose_qt[a], pb.pose_qt[a], pb.psi[l])
        W = np.diag(pb.e_info[e])
        err = oracle.edge_error(CAM, pb.pose_qt[a], pb.pose_qt[a], pb.psi[l], pb.e_obs[e])
        expect[6 * a:6 * a + 6, 6 * a:6 * a + 6] += Jp.T @ W @ Jp
        sl = slice(6 * P + 3 * l, 6 * P + 3 * l + 3)
        expect[sl, sl] += Jpsi.T @ W @ Jpsi
        bexp[sl] += -Jpsi.T @ W @ err
    assert np.abs(dH - expect).max() <= 1e-9 * np.abs(H).max()
    assert np.abs((b - b2) - bexp).max() <= 1e-9 * np.abs(b).max()


@pytest.mark.parametrize("lam", [50.0, 0.01])
def test_reduced_system_is_schur_complement(oracle, lam):
    pb = synth.make_config("C1")
    H, b, chi = oracle.full_system(pb, robust=True)
    S, bs, chi2 = oracle.reduced_system(pb, True, 1.0, lam)
    assert chi == chi2
    n = 6 * pb.P
    Hl = H + lam * np.eye(H.shape[0])
    Hpp, Hpl, Hll = Hl[:n, :n], Hl[:n, n:], Hl[n:, n:]
    X = np.linalg.solve(Hll, np.concatenate([Hpl.T, b[n:, None]], 1))
    Sref = Hpp - Hpl @ X[:, :-1]
    bref = b[:n] - Hpl @ X[:, -1]
    assert np.abs(S - Sref).max() <= 1e-10 * np.abs(Sref).max()
    assert np.abs(bs - bref).max() <= 1e-9 * np.abs(bref).max()


def test_lm_converges_to_truth_without_noise(oracle):
    pb = synth.make_window(8, 300, seed=11, obs_sigma=0.0, outlier_frac=0.0)
    pb.C = 0   # the synthetic pose-pose measurements carry noise of their own
    pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    poses, psi, st = oracle.optimize(pb, 30, robust=True)
    assert st["chi2_final"] < 1e-6 * st["chi2_init"]
    # gauge is free (no fixed pose): compare relative poses
    from oracle import pyoracle as po
    for i in range(1, pb.P):
        rel = po.se3_mul(poses[i], po.se3_inv(poses[0]))
        rel_t = po.se3_mul(pb.truth_pose_qt[i], po.se3_inv(pb.truth_pose_qt[0]))
        assert np.abs(po.se3_log(po.se3_mul(rel, po.se3_inv(rel_t)))).max() < 2e-3


def test_lm_bookkeeping(oracle):
    pb = synth.make_config("C1")
    poses, psi, st = oracle.optimize(pb, 5)
    assert st["iterations"] == 5 and st["trials_total"] == sum(st["trials_iter"])
    assert all(a > b for a, b in zip([st["chi2_init"]] + st["chi2_iter"][:-1], st["chi2_iter"]))
    # accepted first trial with rho ~ 1: lambda shrinks by 1/3 (goodStepLowerScale)
    np.testing.assert_allclose(st["lambda_iter"][0], 50.0 / 3)
    # empty problem: g2o returns -1
    assert oracle.optimize(synth.make_window(0, 0, seed=1), 3)[2]["iterations"] == -1


def test_golden_vectors(oracle):
    """Committed outputs of the oracle (scripts/make_golden.py) on the seeded C1 window: pins the
    oracle against silent drift and is what the GPU tests compare against on the box."""
    g = np.load(GOLDEN)
    pb = synth.make_config("C1")
    assert pb.E == int(g["E"]) and pb.C == int(g["C"])
    np.testing.assert_allclose(pb.e_obs, g["e_obs"], rtol=0, atol=1e-9)
    p1, s1, st1 = oracle.optimize(pb, 1)
    p5, s5, st5 = oracle.optimize(pb, 5)
    np.testing.assert_allclose(p1, g["poses_it1"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(p5, g["poses_it5"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(s5, g["psi_it5"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(st5["chi2_iter"], g["chi2_iter5"], rtol=1e-10)


def test_threaded_timing_variant_agrees_with_the_sequential_restatement(oracle):
    """oba_set_threads(n > 1) only reorders sums (per-thread copies of the pose blocks); bench.py reports it beside
    the single-thread figure.  Every parity test runs with one thread."""
    pb = synth.make_window(12, 400, seed=8)
    p1, s1, st1 = oracle.optimize(pb, 4)
    try:
        oracle.set_threads(4)
        p4, s4, st4 = oracle.optimize(pb, 4)
    finally:
        oracle.set_threads(1)
    assert st1["trials_iter"] == st4["trials_iter"]
    np.testing.assert_allclose(p4, p1, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(s4, s1, rtol=1e-9, atol=1e-12)
    p1b, s1b, _ = oracle.optimize(pb, 4)
    assert np.array_equal(p1, p1b) and np.array_equal(s1, s1b)


def test_se3_exp_log_match_the_matrix_exponential(oracle):
    """Independent pin of the Sophus restatement: SE3::exp([upsilon; omega]) is the matrix exponential
    of the twist [[hat(omega), upsilon], [0, 0]] (scipy.linalg.expm), log its inverse, including tiny angles."""
    from scipy.linalg import expm, logm
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-3, 1e-9, 2.5):
        d = rng.normal(0, 1, 6) * scale
        d[3:] *= min(1.0, 3.0 / max(np.linalg.norm(d[3:]), 1e-30))        # keep the angle below pi
        w = d[3:]
        X = np.zeros((4, 4))
        X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        X[:3, 3] = d[:3]
        M = expm(X)
        T = oracle.se3_exp(d)
        R = np.array([oracle.se3_act(np.concatenate([T[:4], [0, 0, 0]]), e) for e in np.eye(3)]).T
        np.testing.assert_allclose(R, M[:3, :3], atol=1e-12)
        np.testing.assert_allclose(T[4:], M[:3, 3], atol=1e-12 * max(1.0, np.abs(M[:3, 3]).max()))
        np.testing.assert_allclose(oracle.se3_log(T), d, atol=1e-9 * max(1.0, scale), rtol=1e-9)
        if scale >= 1e-3:
            Lg = np.real(logm(M))
            np.testing.assert_allclose(oracle.se3_log(T)[:3], Lg[:3, 3], atol=1e-9)


def test_zero_information_edges_change_nothing(oracle):
    """The basis of the library's track padding (svs_ba_set_problem completes a track with visibility drop-outs
    with ZERO-WEIGHT edges to the frames it skips): in the reference's arithmetic an edge with Omega = 0 contributes
    exactly nothing -- J^T 0 J = 0, J^T 0 e = 0, chi2 += 0 -- so the reduced system, chi2 and the whole Levenberg
    trajectory of the padded edge list are those of the caller's list."""
    pb = synth.with_dropouts(synth.make_window(14, 500, seed=51), 0.25, seed=7)
    ep, ef, ea, eo, ei = [list(getattr(pb, k)) for k in ("e_point", "e_pose", "e_anchor", "e_obs", "e_info")]
    added = 0
    for l in range(pb.L):
        sel = pb.e_point == l
        if not sel.any():
            continue
        a = int(pb.e_anchor[sel][0])
        obs = sorted(int(f) for f in pb.e_pose[sel] if f != a)
        if len(obs) < 2:
            continue
        for f in range(obs[0], obs[-1] + 1):
            if f != a and f not in obs:       # a frame the track skips: observation arbitrary, information zero
                ep.append(l); ef.append(f); ea.append(a)
                eo.append(np.array([123.0, 45.0, 100.0])); ei.append(np.zeros(3))
                added += 1
    assert added > 50
    pad = pb.copy()
    pad.e_point = np.asarray(ep, np.int32); pad.e_pose = np.asarray(ef, np.int32); pad.e_anchor = np.asarray(ea, np.int32)
    pad.e_obs = np.asarray(eo, np.float64).reshape(-1, 3); pad.e_info = np.asarray(ei, np.float64).reshape(-1, 3)
    pad.E = len(ep)
    for robust in (True, False):
        S0, b0, c0 = oracle.reduced_system(pb, robust, 1.0, 50.0)
        S1, b1, c1 = oracle.reduced_system(pad, robust, 1.0, 50.0)
        assert c0 == c1 or abs(c0 - c1) <= 1e-14 * abs(c0)
        np.testing.assert_allclose(S1, S0, rtol=0, atol=1e-12 * np.abs(S0).max())
        np.testing.assert_allclose(b1, b0, rtol=0, atol=1e-12 * np.abs(b0).max())
    p0, s0, st0 = oracle.optimize(pb, 5)
    p1, s1, st1 = oracle.optimize(pad, 5)
    assert st0["trials_iter"] == st1["trials_iter"]
    np.testing.assert_allclose(st1["chi2_iter"], st0["chi2_iter"], rtol=1e-12)
    np.testing.assert_allclose(p1, p0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(s1, s0, rtol=0, atol=1e-11)
