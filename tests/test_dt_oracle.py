"""CPU sanity tests of the dense-tracker oracle (oracle/dt_oracle.c)."""
import numpy as np

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si

I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


def _levels(oracle, nlevels=2):
    seq = si.sequence(2)
    cams = fi.level_cams(nlevels=nlevels)
    prev_p, cur_p = fi.float_pyramid(seq[0]["img"], nlevels), fi.float_pyramid(seq[1]["img"], nlevels)
    out = []
    for l in range(nlevels):
        dx, dy = fi.gradients(cur_p[l])
        h, w = prev_p[l].shape
        out.append(dict(prev=prev_p[l], cur=cur_p[l], dx=dx, dy=dy, f=cams[l][0], px=cams[l][1], py=cams[l][2],
                        cloud=oracle.dt_point_cloud(I7, cams[l], seq[0]["disp"], l, w, h)))
    return out, seq


def test_point_cloud_level0_is_stereo_backprojection(oracle):
    lv, seq = _levels(oracle, 1)
    c = lv[0]["cloud"]
    d = seq[0]["disp"]
    v, u = 300, 200
    z = fi.level_cams()[0][0] * fi.level_cams()[0][3] / d[v, u]
    np.testing.assert_allclose(c[v, u, 2], z, rtol=1e-5)
    np.testing.assert_allclose(c[v, u, 0], (u - 320.0) * z / fi.level_cams()[0][0], rtol=1e-4, atol=1e-4)
    assert c[v, u, 3] == 1.0


def test_gradient_of_chi2_matches_jacobian_sum(oracle):
    """b = J^T r is half the derivative of chi2 along exp(eps e_k) T (exact-bilinear mode): the LM
    step solves H x = -b (dense_tracking.cpp:127-135)."""
    lv, _ = _levels(oracle, 1)
    chi, H, b, n = oracle.dt_pass(lv[0], I7, exact=True)
    eps = 1e-3
    for k in (0, 3, 4, 5):
        d = np.zeros(6); d[k] = eps
        cp = oracle.dt_pass(lv[0], oracle.se3_exp(d), exact=True, want_jac=False)[0]
        cm = oracle.dt_pass(lv[0], oracle.se3_exp(-d), exact=True, want_jac=False)[0]
        g = (cp - cm) / (2 * eps)
        # d chi2 / d x_k = 2 b_k up to the discretisation of the image gradients
        assert abs(g - 2 * b[k]) < 0.15 * abs(g)


def test_track_reduces_chi2_and_recovers_motion(oracle):
    lv, _ = _levels(oracle, 2)
    c0 = oracle.dt_pass(lv[0], I7)[0]
    T, st = oracle.dt_track(lv, I7)
    assert st["chi2"][0] < 0.2 * c0
    assert abs(T[6] + 0.02) < 0.01


def _numpy_pass(lv, T, exact=True):
    """Vectorised float32 restatement of one sweep of the CUDA-build tracker (gpu/dense_tracking.cu:24-80,
    172-263): independent of oracle/dt_oracle.c except for the point cloud it is given."""
    f32 = np.float32
    x, y, z, w = T[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]).astype(f32)
    t = np.asarray(T[4:], np.float64).astype(f32)
    h, wd = lv["prev"].shape
    c = lv["cloud"].reshape(-1, 4).astype(f32)
    vv, uu = np.divmod(np.arange(h * wd), wd)
    ok = c[:, 3] > 0
    P = c[:, :3] @ R.T + c[:, 3:4] * t
    fl, px, py = f32(lv["f"]), f32(lv["px"]), f32(lv["py"])
    with np.errstate(all="ignore"):
        uc = fl * P[:, 0] / P[:, 2] + px
        vc = fl * P[:, 1] / P[:, 2] + py
    ok &= (uc >= 1) & (vc >= 1) & (uc <= wd - 2) & (vc <= h - 2)
    uc, vc, P, uu, vv = uc[ok], vc[ok], P[ok], uu[ok], vv[ok]

    def interp(img):
        x0, y0 = np.floor(uc), np.floor(vc)
        a, b = uc - x0, vc - y0
        if not exact:
            a = np.floor(a * f32(256) + f32(0.5)) / f32(256)
            b = np.floor(b * f32(256) + f32(0.5)) / f32(256)
        xi, yi = x0.astype(int), y0.astype(int)
        one = f32(1)
        return ((one - a) * (one - b)) * img[yi, xi] + (a * (one - b)) * img[yi, xi + 1] + \
               ((one - a) * b) * img[yi + 1, xi] + (a * b) * img[yi + 1, xi + 1]

    res = lv["prev"][vv, uu] - interp(lv["cur"])
    dx = f32(0.5) * interp(lv["dx"]) * fl
    dy = f32(0.5) * interp(lv["dy"]) * fl
    cx, cy, cz = P[:, 0], P[:, 1], P[:, 2]
    zs = cz * cz
    J = np.stack([-dx / cz, -dy / cz, dx * cx / zs + dy * cy / zs, dx * (cx * cy) / zs + dy * (1 + cy * cy / zs),
                  -dx * (1 + cx * cx / zs) - dy * (cx * cy) / zs, dx * cy / cz - dy * cx / cz], 1).astype(np.float64)
    H = J.T @ J
    return float((res.astype(np.float64) ** 2).sum()), H[np.tril_indices(6)], J.T @ res.astype(np.float64), int(ok.sum())


def test_sweep_matches_an_independent_numpy_restatement(oracle):
    lv, _ = _levels(oracle, 2)
    for l, T, exact in ((1, I7, True), (0, I7, False), (0, oracle.se3_exp(np.array([0.01, -0.004, 0.02, 0.002, -0.003, 0.001])), True)):
        chi, H, b, n = oracle.dt_pass(lv[l], T, exact=exact)
        chi_n, H_n, b_n, n_n = _numpy_pass(lv[l], np.asarray(T), exact)
        assert abs(n - n_n) <= 2                      # pixels whose projection rounds onto the frame border
        np.testing.assert_allclose(chi, chi_n, rtol=2e-4)
        np.testing.assert_allclose(H, H_n, rtol=2e-4, atol=2e-4 * np.abs(H_n).max())
        np.testing.assert_allclose(b, b_n, rtol=2e-3, atol=2e-4 * np.abs(b_n).max())


def test_residual_image_is_consistent_with_the_chi2_sweep(oracle):
    """residualImage_kernel classifies exactly the pixels chi2_kernel sums: grey pixels = contributing pixels, and the
    grey value inverts to the squared residual wherever it is not clamped."""
    lv, _ = _levels(oracle, 2)
    T = oracle.se3_exp(np.array([0.03, 0.01, -0.05, 0.002, -0.02, 0.004]))
    for l in (0, 1):
        chi, _, _, n = oracle.dt_pass(lv[l], T, exact=False, want_jac=False)
        img = oracle.dt_residual_image(lv[l], T, exact=False)
        assert img.shape == lv[l]["prev"].shape + (4,) and (img[..., 3] == 1).all()
        grey = (img[..., 0] == img[..., 1]) & (img[..., 1] == img[..., 2])
        red = (img[..., 0] == 1) & (img[..., 1] == 0) & (img[..., 2] == 0)
        green = (img[..., 0] == 0) & (img[..., 1] == 1) & (img[..., 2] == 0)
        assert int(grey.sum()) == n and int((grey | red | green).sum()) == grey.size
        assert int(green.sum()) == int((np.asarray(lv[l]["cloud"])[..., 3] <= 0).sum())
        g = img[..., 0][grey].astype(np.float64)
        unclamped = g > 0
        r2 = (1 - g[unclamped]) / 50
        assert r2.sum() <= chi * (1 + 1e-3) and (g >= 0).all() and (g <= 1).all()
