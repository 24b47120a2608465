"""GPU parity tests of the guided matcher: CUDA (through the C ABI) vs oracle/match_oracle.c.
Bar: every field bit-exact (integer scores, chosen corner incl. tie-break, truncated uint8 warp),
observations and 3-D points equal to the last bit (same FP64 operations, no FMA contraction)."""
import numpy as np
import pytest

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si

pytestmark = pytest.mark.gpu

I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])
NLV = 2


def _levels():
    cams = fi.level_cams(nlevels=NLV)
    return [(640 >> l, 480 >> l, cams[l][0], cams[l][1], cams[l][2]) for l in range(NLV)], cams


def _features(oracle, pyr):
    feats = []
    for l in range(NLV):
        g = oracle.fast_grid(640 >> l, 480 >> l, 222 if l == 0 else 55, 74 if l == 0 else 18, 25, 3, 3)
        xy, off = oracle.fast_detect_adaptively(pyr[l], g, 5)
        content = np.concatenate([np.arange(off[c + 1] - off[c]) for c in range(9)]).astype(np.int32)
        feats.append((xy, content))
    return feats


def _points(oracle, kf_pyr, disp, cams, keyframe=0):
    pts = []
    for l in range(NLV):
        g = oracle.fast_grid(640 >> l, 480 >> l, 222 if l == 0 else 55, 74 if l == 0 else 18, 25, 3, 3)
        kxy, _ = oracle.fast_detect_adaptively(kf_pyr[l], g, 5)
        d = disp[kxy[:, 1] << l, kxy[:, 0] << l] / (1 << l)
        ok = d > 0
        kxy, d = kxy[ok], d[ok]
        z = cams[l][0] * cams[l][3] / d
        p = np.zeros(len(kxy), oracle.MATCH_POINT_DTYPE)
        p["keyframe"] = keyframe
        p["anchor_level"] = l
        p["xyz_anchor"] = np.stack([(kxy[:, 0] - cams[l][1]) / cams[l][0] * z, (kxy[:, 1] - cams[l][2]) / cams[l][0] * z, z], 1)
        p["anchor_obs_pyr"] = kxy
        pts.append(p)
    return np.concatenate(pts)


def _assert_same(res, ref):
    for f in ("predicted", "textured", "matched", "n_candidates", "index", "min_dist", "uv_pyr"):
        np.testing.assert_array_equal(res[f], ref[f], err_msg=f)
    np.testing.assert_array_equal(res["obs"], ref["obs"])
    np.testing.assert_array_equal(res["xyz_actkey"], ref["xyz_actkey"])


@pytest.mark.parametrize("radius", [4, 10])
def test_match_frame_to_keyframe(svs, oracle, radius):
    """Front-end call shape: anchor keyframe = frame 0, current = frame 1 or 2, radius 4 (GPU build)
    and 10 (back-end re-registration, backend.cpp:746)."""
    seq = si.sequence(3)
    levels, cams = _levels()
    kf_pyr = fi.uint8_pyramid(seq[0]["img"], NLV)
    for i_cur in (1, 2):
        cur_pyr = fi.uint8_pyramid(seq[i_cur]["img"], NLV)
        feats = _features(oracle, cur_pyr)
        trees = [oracle.QuadTree(640 >> l, 480 >> l, *feats[l]) for l in range(NLV)]
        pts = _points(oracle, kf_pyr, seq[0]["disp"], cams)
        T_cur = oracle.se3_exp(np.array([0.001, 0.0, -0.02 * i_cur, 0.0, -0.0035 * i_cur, 0.0]))
        T_key_w = oracle.se3_exp(np.array([0.3, -0.1, 0.2, 0.01, 0.02, -0.01]))
        kf_T = T_key_w                                         # anchor keyframe == active keyframe
        m = svs.GuidedMatcher(levels)
        m.set_keyframe(0, kf_T, kf_pyr)
        m.set_current(cur_pyr, seq[i_cur]["disp"])
        for l in range(NLV):
            m.set_features(l, *feats[l])
        res = m.match(T_cur, T_key_w, pts, radius, 22, 10)
        ref = oracle.match(levels, cur_pyr, seq[i_cur]["disp"], trees, [(kf_T, kf_pyr)], T_cur, T_key_w, pts, radius, 22, 10)
        assert ref["matched"].sum() > 0.5 * len(pts)
        _assert_same(res, ref)
        m.close()


def test_two_keyframes_rejections_and_missing_anchor(svs, oracle):
    seq = si.sequence(3)
    levels, cams = _levels()
    kf0, kf1 = fi.uint8_pyramid(seq[0]["img"], NLV), fi.uint8_pyramid(seq[1]["img"], NLV)
    cur_pyr = fi.uint8_pyramid(seq[2]["img"], NLV)
    feats = _features(oracle, cur_pyr)
    trees = [oracle.QuadTree(640 >> l, 480 >> l, *feats[l]) for l in range(NLV)]
    T0 = I7
    T1 = oracle.se3_exp(np.array([0.0, 0.0, -0.02, 0.0, -0.0035, 0.0]))
    pts = np.concatenate([_points(oracle, kf0, seq[0]["disp"], cams, 0), _points(oracle, kf1, seq[1]["disp"], cams, 1)])
    pts["keyframe"][::17] = -1                                 # anchor not in vertex_map
    pts["xyz_anchor"][5::23] *= 6.0                            # depth-ratio rejection
    pts["anchor_obs_pyr"][7::29] = [2.0, 3.0]                  # key pixel too close to the border
    pts["anchor_obs_pyr"][11::31] += 0.37                      # non-integer anchor observations
    T_cur = oracle.se3_exp(np.array([0.0, 0.0, -0.04, 0.0, -0.007, 0.0]))
    m = svs.GuidedMatcher(levels)
    m.set_keyframe(0, T0, kf0)
    m.set_keyframe(1, T1, kf1)
    m.set_current(cur_pyr, seq[2]["disp"])
    for l in range(NLV):
        m.set_features(l, *feats[l])
    # high thr_std rejects low-texture patches; low thr_mean rejects weak matches
    for thr_mean, thr_std in ((22, 10), (8, 60)):
        res = m.match(T_cur, T0, pts, 8, thr_mean, thr_std)
        ref = oracle.match(levels, cur_pyr, seq[2]["disp"], trees, [(T0, kf0), (T1, kf1)], T_cur, T0, pts, 8, thr_mean, thr_std)
        assert 0 < ref["matched"].sum() < len(pts) and (ref["predicted"] == 0).sum() > 0
        _assert_same(res, ref)
    m.close()


def test_score_ties_follow_quadtree_order(svs, oracle):
    """Periodic texture => identical 8x8 patches at several corners => equal scores; the reference
    keeps the first one its quadtree enumerates."""
    levels, cams = _levels()
    yy, xx = np.mgrid[0:480, 0:640]
    img = ((xx % 4) * 40 + (yy % 4) * 17 + 30).astype(np.uint8)
    pyr = [img, img[::2, ::2].copy()]
    disp = np.full((480, 640), 6.0, np.float32)
    # corners on the 4-px lattice: all patches around them are identical
    gx, gy = np.meshgrid(np.arange(40, 600, 4), np.arange(40, 440, 4))
    xy0 = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.int32)
    rng = np.random.default_rng(3)
    xy0 = xy0[rng.permutation(len(xy0))[:3000]]
    feats = [(xy0, np.arange(len(xy0), dtype=np.int32)), (np.zeros((0, 2), np.int32), np.zeros(0, np.int32))]
    trees = [oracle.QuadTree(640 >> l, 480 >> l, *feats[l]) for l in range(NLV)]
    pts = np.zeros(400, oracle.MATCH_POINT_DTYPE)
    uv = np.stack([rng.integers(60, 580, 400) // 4 * 4, rng.integers(60, 420, 400) // 4 * 4], 1).astype(np.float64)
    z = cams[0][0] * cams[0][3] / 6.0
    pts["xyz_anchor"] = np.stack([(uv[:, 0] - cams[0][1]) / cams[0][0] * z, (uv[:, 1] - cams[0][2]) / cams[0][0] * z,
                                  np.full(400, z)], 1)
    pts["anchor_obs_pyr"] = uv
    m = svs.GuidedMatcher(levels)
    m.set_keyframe(0, I7, pyr)
    m.set_current(pyr, disp)
    for l in range(NLV):
        m.set_features(l, *feats[l])
    res = m.match(I7, I7, pts, 10, 22, 1)
    ref = oracle.match(levels, pyr, disp, trees, [(I7, pyr)], I7, I7, pts, 10, 22, 1)
    assert (ref["n_candidates"] > 3).mean() > 0.5 and ref["matched"].sum() > 100
    _assert_same(res, ref)
    m.close()


def test_empty_inputs(svs, oracle):
    levels, cams = _levels()
    m = svs.GuidedMatcher(levels)
    img = np.zeros((480, 640), np.uint8)
    pyr = [img, img[::2, ::2].copy()]
    m.set_keyframe(0, I7, pyr)
    m.set_current(pyr, np.zeros((480, 640), np.float32))
    for l in range(NLV):
        m.set_features(l, np.zeros((0, 2), np.int32), np.zeros(0, np.int32))
    assert len(m.match(I7, I7, np.zeros(0, oracle.MATCH_POINT_DTYPE), 4, 22, 10)) == 0
    pts = np.zeros(3, oracle.MATCH_POINT_DTYPE)
    pts["xyz_anchor"] = [0, 0, 5.0]
    pts["anchor_obs_pyr"] = [320, 240]
    res = m.match(I7, I7, pts, 4, 22, 10)
    assert res["matched"].sum() == 0 and (res["index"] == -1).all()
    with pytest.raises(svs.SvsError):
        m.set_features(0, np.array([[700, 10]], np.int32), np.array([0], np.int32))
    m.close()


def test_features_handed_over_on_the_device(svs, oracle):
    """FAST corners taken from the detector's device buffers (svs_matcher_set_features_from_fast) give the same
    matches, bit for bit, as the corners copied through the host with their per-cell ordinals."""
    seq = si.sequence(2)
    levels, cams = _levels()
    kf_pyr = fi.uint8_pyramid(seq[0]["img"], NLV)
    cur_pyr = fi.uint8_pyramid(seq[1]["img"], NLV)
    pts = _points(oracle, kf_pyr, seq[0]["disp"], cams)
    T_cur = oracle.se3_exp(np.array([0.001, 0.0, -0.02, 0.0, -0.0035, 0.0]))
    res = []
    for on_device in (False, True):
        m = svs.GuidedMatcher(levels)
        m.set_keyframe(0, I7, kf_pyr)
        m.set_current(cur_pyr, seq[1]["disp"])
        for l in range(NLV):
            g = svs.FastGrid(640 >> l, 480 >> l, 222 if l == 0 else 55, 74 if l == 0 else 18, 25, 3, 3)
            g.set_image(cur_pyr[l])
            xy, off = g.detect_adaptively(5)
            if on_device:
                m.set_features_from_fast(l, g)
            else:
                m.set_features(l, xy, np.concatenate([np.arange(off[c + 1] - off[c]) for c in range(9)]).astype(np.int32))
            g.close()
        res.append(m.match(T_cur, I7, pts, 4, 22, 10))
        m.close()
    assert res[0]["matched"].sum() > 0.5 * len(pts)
    _assert_same(res[1], res[0])
