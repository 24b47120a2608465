"""CPU tests: the FAST oracle (oracle/fast_oracle.c) against OpenCV's FastFeatureDetector
(cv2 4.13: same 9/16 segment test and raster order as the 2.4.2 detector the reference calls at
fast_grid.cpp:72,104) and the adaptive threshold walk against hand-computed cases."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from scavislam_b200 import synth_images as si


@pytest.fixture(scope="module")
def frame():
    return si.render_frame(np.zeros(3), 0.0)[0]


def _cv_detect(img, u0, u1, v0, v1, thr):
    det = cv2.FastFeatureDetector_create(int(thr), False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(np.ascontiguousarray(img[v0:v1, u0:u1]))
    return np.array([[int(k.pt[0]) + u0, int(k.pt[1]) + v0] for k in kps], np.int32).reshape(-1, 2)


@pytest.mark.parametrize("thr", [10, 17, 25, 40])
@pytest.mark.parametrize("roi", [(0, 640, 0, 480), (213, 426, 160, 320), (426, 639, 0, 160), (5, 30, 7, 40)])
def test_roi_detection_equals_opencv(oracle, frame, thr, roi):
    u0, u1, v0, v1 = roi
    np.testing.assert_array_equal(oracle.fast_detect_roi(frame, u0, u1, v0, v1, thr), _cv_detect(frame, u0, u1, v0, v1, thr))


def test_random_noise_image_equals_opencv(oracle):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    for thr in (5, 20, 60):
        np.testing.assert_array_equal(oracle.fast_detect_roi(img, 0, 131, 0, 97, thr), _cv_detect(img, 0, 131, 0, 97, thr))


def test_tiny_and_degenerate_rois(oracle, frame):
    assert len(oracle.fast_detect_roi(frame, 10, 16, 10, 60, 10)) == 0    # 6 px wide: no interior
    assert len(oracle.fast_detect_roi(frame, 10, 17, 10, 17, 0)) <= 1     # 7x7: exactly one candidate pixel
    assert len(oracle.fast_detect_roi(frame, 10, 10, 10, 10, 10)) == 0


def test_score_is_max_threshold(oracle, frame):
    from oracle import pyoracle as po
    img, p, pitch = po._u8(frame)
    L = po.lib()
    rng = np.random.default_rng(1)
    for _ in range(300):
        x, y = int(rng.integers(3, 637)), int(rng.integers(3, 477))
        s = L.ofast_score(p, pitch, x, y)
        if s >= 0:
            assert L.ofast_is_corner(p, pitch, x, y, s) and not L.ofast_is_corner(p, pitch, x, y, s + 1)
        else:
            assert not L.ofast_is_corner(p, pitch, x, y, 0)


def test_grid_layout_matches_reference_constructor(oracle):
    # fast_grid.cpp:23-58 with the front-end's level-0 parameters (stereo_frontend.cpp:71-88)
    g = oracle.fast_grid(640, 480, 222, 74, 25, 3, 3)
    assert (g.min_inner, g.min_outer, g.max_inner, g.max_outer) == (197, 148, 246, 296)
    c = g.cells[5]
    assert (c.u0, c.u1, c.v0, c.v1, c.thr) == (426, 639, 160, 320, 25)


def test_detect_groups_cells_in_row_major_order(oracle, frame):
    g = oracle.fast_grid(640, 480, 222, 74, 25, 3, 3)
    cells = [(c.u0, c.u1, c.v0, c.v1, c.thr) for c in list(g.cells)[:9]]
    xy, off = oracle.fast_detect(frame, cells)
    assert off[-1] == len(xy)
    for k, (u0, u1, v0, v1, thr) in enumerate(cells):
        np.testing.assert_array_equal(xy[off[k]:off[k + 1]], _cv_detect(frame, u0, u1, v0, v1, thr))


def test_adaptive_walk_literal_semantics(oracle, frame):
    """detectAdaptively emits the keypoints of the LAST detect() call while the stored threshold
    may already have moved on; thresholds stay inside [fast_min, fast_max]."""
    g = oracle.fast_grid(640, 480, 222, 74, 25, 3, 3)
    thr_hist = []
    for it in range(4):
        xy, off = oracle.fast_detect_adaptively(frame, g, 6 if it else 5)
        thr_hist.append([g.cells[k].thr for k in range(9)])
        assert off[-1] == len(xy)
        assert all(10 <= t <= 40 for t in thr_hist[-1])
    assert thr_hist[0] != [25] * 9          # the walk moved
    # trials = 0: no detection, no keypoints, thresholds untouched
    g2 = oracle.fast_grid(640, 480, 222, 74, 25, 3, 3)
    xy, off = oracle.fast_detect_adaptively(frame, g2, 0)
    assert len(xy) == 0 and [g2.cells[k].thr for k in range(9)] == [25] * 9
