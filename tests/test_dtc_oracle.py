"""CPU tests pinning the restatement of the reference's non-CUDA dense tracker (SURVEY.md 8 row a18,
oracle/dt_oracle.c: odtc_*).  No reference test exists for it (PARITY UNPINNED): pinned by the stereo
back-projection formula, an independent vectorised numpy sweep, the derivative of chi2 and the
recovery of the rendered camera motion."""
import numpy as np

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si

I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


def levels(oracle, nlevels=3, T_cloud=I7):
    seq = si.sequence(2)
    cams = fi.level_cams(nlevels=nlevels)
    p8 = fi.uint8_pyramid(seq[0]["img"], nlevels)
    cur = fi.float_pyramid(seq[1]["img"], nlevels)
    out = []
    for l in range(nlevels):
        dx, dy = fi.gradients(cur[l])
        h, w = cur[l].shape
        out.append(dict(prev_u8=p8[l], cur=cur[l], dx=dx, dy=dy, cam=cams[l],
                        cloud=oracle.dtc_point_cloud(T_cloud, cams[l], seq[0]["disp"], l, w, h)))
    return out, seq, cams


def test_point_cloud_is_every_fourth_pixel_back_projected(oracle):
    lv, seq, cams = levels(oracle)
    for l in range(3):
        f, px, py, b = cams[l]
        c = lv[l]["cloud"]
        assert c.shape == ((480 >> l) // 4, (640 >> l) // 4, 4)
        v, u = c.shape[0] // 2, c.shape[1] // 3
        d = seq[0]["disp"][(v * 4) << l, (u * 4) << l] / (1 << l)      # disparity scaled by 2^-level
        assert d > 0 and c[v, u, 3] == 1.0
        z = f * b / d
        np.testing.assert_allclose(c[v, u, :3], [(u * 4 - px) * z / f, (v * 4 - py) * z / f, z], rtol=2e-6)
    # an invalid disparity gives w = -1, and the cloud is expressed in the active keyframe (T^-1 applied)
    disp = seq[0]["disp"].copy(); disp[0, 0] = 0
    c = oracle.dtc_point_cloud(I7, cams[0], disp, 0, 640, 480)
    assert tuple(c[0, 0]) == (0, 0, 0, -1)
    T = oracle.se3_exp(np.array([0.1, -0.05, 0.02, 0.01, 0.02, -0.03]))
    cT = oracle.dtc_point_cloud(T, cams[0], seq[0]["disp"], 0, 640, 480)
    back = np.array([oracle.se3_act(T, cT[60, 80, :3].astype(np.float64))])[0]
    np.testing.assert_allclose(back, lv[0]["cloud"][60, 80, :3], rtol=1e-5, atol=1e-5)


def _numpy_pass(lv, T, oracle):
    """Vectorised restatement of one sweep (dense_tracking.cpp:276-331)."""
    f, px, py, b = lv["cam"]
    h, w = lv["cur"].shape
    c = lv["cloud"].reshape(-1, 4)
    vv, uu = np.divmod(np.arange(len(c)), w // 4)
    ok = c[:, 3] > 0
    R = np.array([oracle.se3_act(np.concatenate([T[:4], [0, 0, 0]]), e) for e in np.eye(3)]).T
    X = c[:, :3].astype(np.float64) @ R.T + T[4:]
    with np.errstate(all="ignore"):
        uc = (f * (X[:, 0] / X[:, 2]) + px).astype(np.float32)
        vc = (f * (X[:, 1] / X[:, 2]) + py).astype(np.float32)
    ui, vi = np.trunc(uc).astype(np.int64), np.trunc(vc).astype(np.int64)
    ok &= (ui >= 2) & (ui < w - 2) & (vi >= 2) & (vi < h - 2)
    uc, vc, X, uu, vv = uc[ok], vc[ok], X[ok], uu[ok], vv[ok]

    def interp(img):
        x, y = np.floor(uc), np.floor(vc)
        sx, sy = uc - x, vc - y
        xi, yi = x.astype(int), y.astype(int)
        one = np.float32(1)
        return ((one - sx) * (one - sy)) * img[yi, xi] + ((one - sx) * sy) * img[yi + 1, xi] + \
               (sx * (one - sy)) * img[yi, xi + 1] + (sx * sy) * img[yi + 1, xi + 1]

    ip = ((1.0 / 255.0) * lv["prev_u8"][vv * 4, uu * 4].astype(np.float64)).astype(np.float32)
    res = np.clip(ip - interp(lv["cur"]), np.float32(-0.1), np.float32(0.1)).astype(np.float32)
    dx = (0.5 * interp(lv["dx"]).astype(np.float64)).astype(np.float32).astype(np.float64)
    dy = (0.5 * interp(lv["dy"]).astype(np.float64)).astype(np.float32).astype(np.float64)
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    z2 = z * z
    r0 = np.stack([-1 / z * f, 0 * z, x / z2 * f, x * y / z2 * f, -(1 + x * x / z2) * f, y / z * f], 1)
    r1 = np.stack([0 * z, -1 / z * f, y / z2 * f, (1 + y * y / z2) * f, -x * y / z2 * f, -x / z * f], 1)
    J = dx[:, None] * r0 + dy[:, None] * r1
    H = J.T @ J
    return float((res.astype(np.float64) ** 2).sum()), H[np.triu_indices(6)], J.T @ res.astype(np.float64), int(ok.sum())


def test_sweep_matches_an_independent_numpy_restatement(oracle):
    lv, _, _ = levels(oracle)
    for l, T in ((2, I7), (1, oracle.se3_exp(np.array([0.01, -0.004, 0.02, 0.002, -0.003, 0.001]))), (0, I7)):
        chi, H, b, n = oracle.dtc_pass(lv[l], T)
        chi_n, H_n, b_n, n_n = _numpy_pass(lv[l], np.asarray(T), oracle)
        assert n == n_n and n > 0.8 * lv[l]["cloud"].shape[0] * lv[l]["cloud"].shape[1] * 0.5
        # (the float32 roundings of u, v and of the taps see a rotation matrix built another way)
        np.testing.assert_allclose(chi, chi_n, rtol=1e-6)
        np.testing.assert_allclose(H, H_n, rtol=1e-6, atol=1e-6 * np.abs(H_n).max())
        np.testing.assert_allclose(b, b_n, rtol=1e-5, atol=1e-6 * np.abs(b_n).max())


def test_jres_is_half_the_derivative_of_chi2(oracle):
    """With no residual clamped (the frame tracked against itself from a slightly wrong pose), Jres = J^T r is
    half the derivative of chi2 along exp(eps e_k) T.  (Clamped residuals keep their full Jacobian in the
    reference, so the identity does not hold on real frame pairs.)"""
    import cv2
    lv, seq, cams = levels(oracle, 1)
    # a smooth image: on pixel-scale noise the central-difference gradient is not the slope of the bilinear
    # interpolant, and the identity would be off by up to a factor of two
    img = cv2.GaussianBlur(seq[0]["img"], (0, 0), 3.0)
    cur = fi.float_pyramid(img, 1)[0]
    dx, dy = fi.gradients(cur)
    lv0 = dict(lv[0], prev_u8=img, cur=cur, dx=dx, dy=dy)
    T0 = oracle.se3_exp(np.array([0.0004, -0.0003, 0.0005, 0.0001, -0.0001, 0.0002]))
    chi, H, b, n = oracle.dtc_pass(lv0, T0)
    eps = 2e-5
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        cp = oracle.dtc_pass(lv0, oracle.se3_mul(oracle.se3_exp(d), T0))[0]
        cm = oracle.dtc_pass(lv0, oracle.se3_mul(oracle.se3_exp(-d), T0))[0]
        g = (cp - cm) / (2 * eps)
        assert abs(g - 2 * b[k]) < 0.1 * abs(g) + 0.02 * np.abs(b).max(), (k, g, 2 * b[k])


def test_track_recovers_the_rendered_motion(oracle):
    lv, _, _ = levels(oracle)
    c0 = oracle.dtc_pass(lv[0], I7)[0]
    T, st = oracle.dtc_track(lv, I7)
    assert st["chi2"][0] < 0.5 * c0
    assert abs(T[6] + 0.02) < 0.01                          # synth_images.sequence: 2 cm forward per frame
    assert all(2 <= p <= 17 for p in st["passes"])
