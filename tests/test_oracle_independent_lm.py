"""Independent end-to-end pin of the BA oracle (the reference itself cannot be built here).

A second implementation of SlamGraph::optimize's arithmetic that shares NO code with oracle/ba_oracle.c and takes a
different route at every step:
  * poses are 4x4 homogeneous matrices, the update is T <- expm(hat(delta)) T with scipy.linalg.expm, the pose-pose
    residual is vee(logm(T21 T1 T2^-1)) with scipy.linalg.logm (the oracle: unit quaternions, closed-form Sophus exp/log);
  * every Jacobian is a central finite difference of the residuals (the oracle: the analytic blocks of
    anchored_points.cpp:168-189 and the truncated series `third`, :207-235);
  * the FULL system over poses and landmarks is assembled as a scipy.sparse matrix and solved directly
    (the oracle: Schur complement on the landmarks + block Cholesky of the reduced camera system);
  * the Levenberg schedule is written from SURVEY.md Appendix A.2 (g2o OptimizationAlgorithmLevenberg).
Both must produce the same chi2 / lambda / trial trajectory and the same poses and points for several iterations.

The three g2o-semantics ASSUMPTIONS that no in-container source can confirm are isolated at the bottom as explicit,
falsifiable statements (a future baseline/_ref g2o checkout can be run against them)."""
import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from scavislam_b200 import synth


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def _hat6(d):
    M = np.zeros((4, 4))
    M[:3, :3] = _hat(d[3:])
    M[:3, 3] = d[:3]
    return M


def _qt_to_T(qt):
    T = np.tile(np.eye(4), (len(qt), 1, 1))
    T[:, :3, :3] = synth.quat_to_R(qt[:, :4])
    T[:, :3, 3] = qt[:, 4:]
    return T


def _exp_batch(d):
    """expm(hat6(d)) for a batch of 6-vectors (scipy's expm on each 4x4)."""
    return np.stack([scipy.linalg.expm(_hat6(x)) for x in d])


def _obs_residuals(cam, Tp, Ta, psi, obs):
    """e = z - pi_stereo(Tp Ta^-1 invert_depth(psi)), batched over edges (anchored_points.cpp:148-166)."""
    f, px, py, b = cam
    xa = np.stack([psi[:, 0] / psi[:, 2], psi[:, 1] / psi[:, 2], 1.0 / psi[:, 2], np.ones(len(psi))], -1)
    y = np.einsum("eij,ej->ei", Tp @ np.linalg.inv(Ta), xa)
    u = f * y[:, 0] / y[:, 2] + px
    v = f * y[:, 1] / y[:, 2] + py
    ur = f * (y[:, 0] - b) / y[:, 2] + px
    return obs - np.stack([u, v, ur], -1)


def _pp_residuals(T21, T1, T2):
    out = np.zeros((len(T21), 6))
    for c in range(len(T21)):
        M = np.real(scipy.linalg.logm(T21[c] @ T1[c] @ np.linalg.inv(T2[c])))
        out[c, :3] = M[:3, 3]
        out[c, 3:] = [M[2, 1], M[0, 2], M[1, 0]]
    return out


def _huber(e2, delta):
    """(rho, rho') of g2o's RobustKernelHuber on the squared error."""
    small = e2 <= delta * delta
    se = np.sqrt(np.maximum(e2, 1e-300))
    return np.where(small, e2, 2 * delta * se - delta * delta), np.where(small, 1.0, delta / se)


class IndependentLM:
    def __init__(self, pb, robust=True, delta=1.0):
        self.pb, self.robust, self.delta = pb, robust, delta
        self.T = _qt_to_T(pb.pose_qt)
        self.psi = pb.psi.copy()
        self.T21 = _qt_to_T(pb.c_T) if pb.C else np.zeros((0, 4, 4))
        self.n = 6 * pb.P + 3 * pb.L

    def chi2(self, T, psi):
        pb = self.pb
        e = _obs_residuals(pb.cam, T[pb.e_pose], T[pb.e_anchor], psi[pb.e_point], pb.e_obs)
        e2 = np.einsum("ei,ei,ei->e", e, pb.e_info, e)
        chi = (_huber(e2, self.delta)[0] if self.robust else e2).sum()
        if pb.C:
            r = _pp_residuals(self.T21, T[pb.c_i], T[pb.c_j])
            chi += np.einsum("ci,cij,cj->", r, pb.c_Lambda.reshape(-1, 6, 6), r)
        return chi

    def linearise(self):
        """Full Gauss-Newton system (H, b) at the current state from finite-difference Jacobians."""
        pb, T, psi, h = self.pb, self.T, self.psi, 1e-6
        E, P, L = pb.E, pb.P, pb.L
        Tp, Ta, ps = T[pb.e_pose], T[pb.e_anchor], psi[pb.e_point]
        e0 = _obs_residuals(pb.cam, Tp, Ta, ps, pb.e_obs)
        Jp, Ja, Js = np.zeros((E, 3, 6)), np.zeros((E, 3, 6)), np.zeros((E, 3, 3))
        for k in range(6):
            d = np.zeros(6); d[k] = h
            Dp, Dm = scipy.linalg.expm(_hat6(d)), scipy.linalg.expm(_hat6(-d))
            Jp[:, :, k] = (_obs_residuals(pb.cam, Dp @ Tp, Ta, ps, pb.e_obs) - _obs_residuals(pb.cam, Dm @ Tp, Ta, ps, pb.e_obs)) / (2 * h)
            Ja[:, :, k] = (_obs_residuals(pb.cam, Tp, Dp @ Ta, ps, pb.e_obs) - _obs_residuals(pb.cam, Tp, Dm @ Ta, ps, pb.e_obs)) / (2 * h)
        for k in range(3):
            d = np.zeros(3); d[k] = h * 0.01      # psi_z is O(0.1): a smaller step keeps the quotient accurate
            Js[:, :, k] = (_obs_residuals(pb.cam, Tp, Ta, ps + d, pb.e_obs) - _obs_residuals(pb.cam, Tp, Ta, ps - d, pb.e_obs)) / (2 * d[k])
        e2 = np.einsum("ei,ei,ei->e", e0, pb.e_info, e0)
        rho, w = _huber(e2, self.delta) if self.robust else (e2, np.ones(E))
        W = pb.e_info * w[:, None]                       # rho' Omega (diagonal)
        fx = pb.fixed.astype(bool)
        Jp = Jp * (~fx[pb.e_pose])[:, None, None]
        Ja = Ja * (~fx[pb.e_anchor])[:, None, None]
        H = sp.lil_matrix((self.n, self.n))
        b = np.zeros(self.n)
        chi = rho.sum()
        for e in range(E):
            ip, ia, il = 6 * pb.e_pose[e], 6 * pb.e_anchor[e], 6 * P + 3 * pb.e_point[e]
            J = [(ip, Jp[e]), (ia, Ja[e]), (il, Js[e])]
            We = np.diag(W[e])
            for x, (ox, Jx) in enumerate(J):
                b[ox:ox + Jx.shape[1]] -= Jx.T @ We @ e0[e]
                for y, (oy, Jy) in enumerate(J):
                    if y < x:
                        continue
                    blk = Jx.T @ We @ Jy
                    if ox == oy and x != y:
                        # ASSUMPTION A1: two vertices of one multi-edge that are the SAME vertex
                        # (observation in the landmark's own anchor frame): g2o adds the cross term J1^T W J2 to the
                        # diagonal block ONCE (its upper-triangular storage has one block for the pair), not twice.
                        H[ox:ox + 6, oy:oy + 6] += blk
                    elif x == y:
                        H[ox:ox + Jx.shape[1], ox:ox + Jx.shape[1]] += blk
                    else:
                        H[ox:ox + Jx.shape[1], oy:oy + Jy.shape[1]] += blk
                        H[oy:oy + Jy.shape[1], ox:ox + Jx.shape[1]] += blk.T
        if pb.C:
            T1, T2 = T[pb.c_i], T[pb.c_j]
            r0 = _pp_residuals(self.T21, T1, T2)
            J1, J2 = np.zeros((pb.C, 6, 6)), np.zeros((pb.C, 6, 6))
            hh = 1e-5
            for k in range(6):
                d = np.zeros(6); d[k] = hh
                Dp, Dm = scipy.linalg.expm(_hat6(d)), scipy.linalg.expm(_hat6(-d))
                J1[:, :, k] = (_pp_residuals(self.T21, Dp @ T1, T2) - _pp_residuals(self.T21, Dm @ T1, T2)) / (2 * hh)
                J2[:, :, k] = (_pp_residuals(self.T21, T1, Dp @ T2) - _pp_residuals(self.T21, T1, Dm @ T2)) / (2 * hh)
            Lam = pb.c_Lambda.reshape(-1, 6, 6)
            for c in range(pb.C):
                i, j = 6 * pb.c_i[c], 6 * pb.c_j[c]
                A, B = J1[c] * (not fx[pb.c_i[c]]), J2[c] * (not fx[pb.c_j[c]])
                chi += r0[c] @ Lam[c] @ r0[c]
                b[i:i + 6] -= A.T @ Lam[c] @ r0[c]
                b[j:j + 6] -= B.T @ Lam[c] @ r0[c]
                H[i:i + 6, i:i + 6] += A.T @ Lam[c] @ A
                H[j:j + 6, j:j + 6] += B.T @ Lam[c] @ B
                H[i:i + 6, j:j + 6] += A.T @ Lam[c] @ B
                H[j:j + 6, i:i + 6] += B.T @ Lam[c] @ A
        return H.tocsc(), b, chi

    def step(self, dx):
        pb = self.pb
        T = np.stack([scipy.linalg.expm(_hat6(dx[6 * i:6 * i + 6])) @ self.T[i] for i in range(pb.P)])
        return T, self.psi + dx[6 * pb.P:].reshape(pb.L, 3)

    def optimize(self, num_iters, lambda_init=50.0, max_trials=5):
        """g2o OptimizationAlgorithmLevenberg::solve as SURVEY.md A.2 states it."""
        lam, ni = lambda_init, 2.0
        chis, lams, trials = [], [], []
        fixed_diag = np.zeros(self.n)
        fixed_diag[:6 * self.pb.P] = np.repeat(self.pb.fixed.astype(float), 6)   # a fixed vertex is not in the system: identity row
        for it in range(num_iters):
            H, b, chi = self.linearise()
            q, rho = 0, -1.0
            while True:
                Hl = H + sp.diags(np.full(self.n, lam) + fixed_diag)
                dx = spla.spsolve(Hl.tocsc(), b)
                Tn, psin = self.step(dx)
                chin = self.chi2(Tn, psin)
                rho = (chi - chin) / (dx @ (lam * dx + b) + 1e-3)
                if rho > 0 and np.isfinite(chin):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    ni = 2.0
                    self.T, self.psi, chi = Tn, psin, chin
                else:
                    lam *= ni
                    ni *= 2
                q += 1
                if not (rho < 0 and q < max_trials):
                    break
            chis.append(chi); lams.append(lam); trials.append(q)
            if q == max_trials or rho == 0:
                break
        return chis, lams, trials


def _T_to_qt_rel(T, qt):
    """Largest difference between the 4x4 poses and the quaternion poses, as matrices."""
    return np.abs(T - _qt_to_T(qt)).max()


@pytest.mark.parametrize("window", ["C1", "K30", "rejections"])
def test_independent_lm_reproduces_the_oracle_trajectory(oracle, window):
    if window == "C1":
        pb, iters, lam0, mt, robust = synth.make_config("C1"), 3, 50.0, 5, True
    elif window == "K30":
        pb, iters, lam0, mt, robust = synth.make_window(30, 450, seed=41), 2, 50.0, 5, True
    else:   # large perturbation + small lambda0: iterations with rejected Levenberg trials.  No pose-pose edges here: the
            # reference linearises them with the truncated series `third` (anchored_points.cpp:207-215), which is not the
            # exact derivative once the relative-pose error is as large as this case makes it
        pb, iters, lam0, mt, robust = synth.make_window(8, 150, seed=100, pose_noise=(0.8, 0.25), depth_noise=0.6), 5, 1e-2, 10, False
        pb.C = 0
        pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    ref = IndependentLM(pb, robust=robust)
    chis, lams, trials = ref.optimize(iters, lam0, mt)
    poses, psi, st = oracle.optimize(pb, iters, robust, 1.0, lam0, mt)
    assert st["trials_iter"] == trials, (st["trials_iter"], trials)
    np.testing.assert_allclose(st["chi2_iter"], chis, rtol=2e-6)
    np.testing.assert_allclose(st["lambda_iter"], lams, rtol=2e-5)
    if window == "rejections":
        assert sum(trials) > len(trials), "case does not exercise rejected trials"
    assert _T_to_qt_rel(ref.T, poses) < 2e-6 * np.abs(pb.pose_qt[:, 4:]).max()
    assert np.abs(ref.psi - psi).max() < 2e-6 * np.abs(psi).max()


# ------------------------------------------------------------------ explicit g2o-semantics assumptions

def test_assumption_A1_self_anchor_cross_term_counted_once(oracle):
    """With the cross term counted TWICE (what a symmetric accumulation would do) the self-anchor observation would
    leave the anchor pose's block untouched; the oracle (and the CUDA path) keep +J1^T W J1.  A g2o checkout decides."""
    pb = synth.make_window(6, 40, seed=7)
    pb.C = 0
    pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    H, b, _ = oracle.full_system(pb, robust=False)
    Hi, bi, _ = IndependentLM(pb, robust=False).linearise()
    assert np.abs(H - Hi.toarray()).max() <= 5e-5 * np.abs(H).max()
    assert np.abs(b - bi).max() <= 5e-5 * np.abs(b).max()
    self_edges = np.nonzero(pb.e_pose == pb.e_anchor)[0]
    assert len(self_edges) > 0


def test_assumption_A2_gain_ratio_denominator(oracle):
    """rho = (chi - chi_new) / (dx . (lambda dx + b) + 1e-3) with b = -J^T W e of the FULL system (poses and points):
    after one accepted iteration from lambda0 = 50 the oracle's lambda equals the value this formula gives."""
    pb = synth.make_config("C1")
    ref = IndependentLM(pb)
    chis, lams, trials = ref.optimize(1)
    _, _, st = oracle.optimize(pb, 1)
    assert trials == [1] and st["trials_iter"] == [1]
    np.testing.assert_allclose(st["lambda_iter"][0], lams[0], rtol=1e-6)


def test_assumption_A3_terminate_rule(oracle):
    """optimize() stops early when a solve() ends with qmax == max_trials or rho == 0 (Terminate), and one outer
    iteration counts however many trials it took."""
    pb = synth.make_window(4, 0, seed=5)            # no landmarks, no edges: chi2 = 0, rho = 0 at once
    _, _, st = oracle.optimize(pb, 3)
    assert st["iterations"] == 1
    pb = synth.make_window(8, 150, seed=100, pose_noise=(0.8, 0.25), depth_noise=0.6)
    pb.C = 0
    pb.c_i = pb.c_i[:0]; pb.c_j = pb.c_j[:0]; pb.c_T = pb.c_T[:0]; pb.c_Lambda = pb.c_Lambda[:0]
    _, _, st = oracle.optimize(pb, 6, False, 1.0, 1e-2, 2)   # two trials allowed: the first iteration exhausts them
    ref = IndependentLM(pb, robust=False)
    chis, lams, trials = ref.optimize(6, 1e-2, 2)
    assert st["trials_iter"] == trials and st["iterations"] == len(trials)
