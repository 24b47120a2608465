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
