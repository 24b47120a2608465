"""GPU parity of the motion-only LM kernel (svs_calcFastMotionOnly) against oracle/pose_oracle.c,
through the C ABI.  Tolerance: 1e-6 relative on the pose parameters (north_star); the trial sequence
(accepted steps, solves) must be identical."""
import numpy as np
import pytest

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si
from scavislam_b200 import synth_pose as sp

pytestmark = pytest.mark.gpu
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


def _check(svs, oracle, tr, robust, kparam, iters, mu=-1.0, strict_trials=True):
    po = svs.PoseOptimizer()
    T_g, sg = po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], robust, kparam, iters, mu)
    T_o, so = oracle.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"], robust, kparam, iters, mu)
    po.close()
    assert sg["num_obs"] == so["num_obs"]
    if strict_trials:   # far from the fixed point the accept/reject sequence is identical
        assert (sg["iterations"], sg["trials"]) == (so["iterations"], so["trials"])
    np.testing.assert_allclose(T_g, T_o, rtol=1e-6, atol=1e-9)
    for k in ("initial_chi2", "chi2", "max_err"):
        assert abs(sg[k] - so[k]) <= 1e-9 * max(1.0, abs(so[k])), k
    return T_g, sg


@pytest.mark.parametrize("n,seed,robust,out,shared", [(20, 1, True, 0.0, False), (333, 2, True, 0.1, False),
                                                     (1800, 3, True, 0.15, True), (513, 4, False, 0.0, False),
                                                     (5000, 5, True, 0.05, False), (1, 6, True, 0.0, False)])
def test_matches_oracle(svs, oracle, n, seed, robust, out, shared):
    tr = sp.make_track(n, seed=seed, outlier_frac=out, shared_points=shared)
    _check(svs, oracle, tr, robust, 2.0, 4)
    # The reference's two settings (stereo_frontend.cpp:1061, backend.cpp:758).  Once converged the test
    # `chi2 - new_chi2 > 0` (pose_optimizer.h:269-270) is decided by the rounding of a sum over n terms, so
    # the count of further (numerically void) trials depends on summation order; pose and chi2 do not.
    for iters in (15, 25):
        _check(svs, oracle, tr, robust, 2.0, iters, strict_trials=False)


def test_rejected_steps_follow_the_same_mu_schedule(svs, oracle):
    # a far-off start and a tiny fixed initial mu force rejected trials (mu *= nu, nu *= 2, 5 in a row = stop)
    tr = sp.make_track(200, seed=11, start_noise=(3.0, 1.2))
    T, st = _check(svs, oracle, tr, False, 1.0, 8, mu=1e-9)
    assert st["trials"] > st["iterations"] + 1
    tr = sp.make_track(200, seed=13, start_noise=(1.5, 0.8))       # ends on five rejections in a row (:289-290)
    T, st = _check(svs, oracle, tr, False, 1.0, 8, mu=1e-9)
    assert st["trials"] == st["iterations"] + 5
    tr2 = sp.make_track(64, seed=12, pixel_noise=0.0, start_noise=(0.0, 0.0))   # starts at the optimum
    _check(svs, oracle, tr2, False, 1.0, 10, strict_trials=False)


def test_zero_iterations_and_errors(svs, oracle):
    tr = sp.make_track(100, seed=13)
    T, st = _check(svs, oracle, tr, True, 2.0, 0)
    assert np.array_equal(T, tr["T_init"]) and st["trials"] == 0
    po = svs.PoseOptimizer(max_obs=64)
    with pytest.raises(svs.SvsError):                       # over capacity
        po.calc_fast_motion_only(tr["pid"], tr["obs"], tr["xyz"], tr["cam"], tr["T_init"])
    bad = tr["pid"][:10].copy(); bad[3] = 1000
    with pytest.raises(svs.SvsError):                       # point_id outside point_list
        po.calc_fast_motion_only(bad, tr["obs"][:10], tr["xyz"][:10], tr["cam"], tr["T_init"])
    obs = tr["obs"][:10].copy(); obs[2, 1] = np.nan
    with pytest.raises(svs.SvsError) as e:                  # the reference throws "Res is NaN!"
        po.calc_fast_motion_only(tr["pid"][:10], obs, tr["xyz"][:10], tr["cam"], tr["T_init"])
    assert e.value.rc == -6 and "NaN" in str(e.value)
    po.close()


def test_refines_the_pose_straight_from_the_matcher_results(svs, oracle):
    """match -> calcFastMotionOnly without the TrackData leaving the device
    (stereo_frontend.cpp:1035-1063)."""
    seq = si.sequence(2)
    cams = fi.level_cams()
    cam = (cams[0][0], cams[0][1], cams[0][2], cams[0][3])
    lv2 = [(640 >> l, 480 >> l, cams[l][0], cams[l][1], cams[l][2]) for l in range(2)]
    fg = svs.FastGrid(640, 480, 222, 74, 25, 3, 3)
    fg.set_image(seq[0]["img"])
    kxy, _ = fg.detect_adaptively(5)
    fg.set_image(seq[1]["img"])
    xy, off = fg.detect_adaptively(5)
    m = svs.GuidedMatcher(lv2)
    m.set_keyframe(0, I7, fi.uint8_pyramid(seq[0]["img"], 2))
    m.set_current(fi.uint8_pyramid(seq[1]["img"], 2), seq[1]["disp"])
    m.set_features(0, xy, np.concatenate([np.arange(off[c + 1] - off[c]) for c in range(9)]).astype(np.int32))
    m.set_features(1, np.zeros((0, 2), np.int32), np.zeros(0, np.int32))
    d = seq[0]["disp"][kxy[:, 1], kxy[:, 0]]
    kxy, d = kxy[d > 0], d[d > 0]
    z = cam[0] * cam[3] / d
    pts = np.zeros(len(kxy), svs.MATCH_POINT_DTYPE)
    pts["xyz_anchor"] = np.stack([(kxy[:, 0] - cam[1]) / cam[0] * z, (kxy[:, 1] - cam[2]) / cam[0] * z, z], 1)
    pts["anchor_obs_pyr"] = kxy
    res = m.match(I7, I7, pts, 4, 22, 10)
    ok = res["matched"] == 1
    assert ok.sum() >= 20                                    # the reference gives up below 20 (stereo_frontend.cpp:1053)
    po = svs.PoseOptimizer()
    T_dev, sd = po.calc_fast_motion_only_matched(m, cam, I7, True, 2.0, 15)
    # the same TrackData gathered on the host, as the reference would build it (obs_list / point_list)
    obs, xyz = np.ascontiguousarray(res["obs"][ok]), np.ascontiguousarray(res["xyz_actkey"][ok])
    T_host, sh = po.calc_fast_motion_only(np.arange(ok.sum()), obs, xyz, cam, I7, True, 2.0, 15)
    T_o, so = oracle.calc_fast_motion_only(np.arange(ok.sum()), obs, xyz, np.array(cam), I7, True, 2.0, 15)
    assert sd["num_obs"] == int(ok.sum()) == sh["num_obs"]
    np.testing.assert_allclose(T_dev, T_host, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(T_dev, T_o, rtol=1e-6, atol=1e-9)
    assert -0.03 < T_dev[6] < -0.005                         # towards the rendered motion (t_z = -0.02; radius-4 integer matches)
    for h in (po, m, fg):
        h.close()
