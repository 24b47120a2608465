"""GPU parity tests of the grid FAST detector: CUDA (through the C ABI) vs the CPU oracle and
vs OpenCV.  Bar: bit-exact keypoint coordinates, order and per-cell ordinals (north_star)."""
import numpy as np
import pytest

from scavislam_b200 import synth_images as si

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames():
    return [f["img"] for f in si.sequence(3)]


def _level1(img):
    import cv2
    return cv2.pyrDown(img)


def _front_end_grids(svs, level):
    # stereo_frontend.cpp:71-88
    dim = max(3 - int(level * 0.5), 1)
    inv = 0.5 ** level
    total = int(2000 * inv * inv)
    per_cell = total // (dim * dim)
    bound = max(per_cell // 3, 10)
    w, h = 640 >> level, 480 >> level
    return (w, h, per_cell, bound, 25, dim, dim)


@pytest.mark.parametrize("level", [0, 1])
def test_detect_static_bit_exact(svs, oracle, frames, level):
    img = frames[0] if level == 0 else _level1(frames[0])
    args = _front_end_grids(svs, level)
    fg = svs.FastGrid(*args)
    fg.set_image(img)
    for thr in (10, 25, 40):
        cells = [(u0, u1, v0, v1, thr + (k % 3)) for k, (u0, u1, v0, v1, _) in enumerate(fg.cell_list())]
        xy, off = fg.detect(cells)
        xo, oo = oracle.fast_detect(img, cells)
        np.testing.assert_array_equal(off, oo)
        np.testing.assert_array_equal(xy, xo)
    fg.close()


def test_detect_equals_opencv_whole_image(svs, frames):
    import cv2
    img = frames[1]
    fg = svs.FastGrid(640, 480, 100, 30, 20, 1, 1)
    fg.set_image(img)
    xy, off = fg.detect([(0, 640, 0, 480, 20)])
    det = cv2.FastFeatureDetector_create(20, False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = np.array([[int(k.pt[0]), int(k.pt[1])] for k in det.detect(img)], np.int32).reshape(-1, 2)
    np.testing.assert_array_equal(xy, ref)
    fg.close()


@pytest.mark.parametrize("level", [0, 1])
def test_adaptive_threshold_trajectory_bit_exact(svs, oracle, frames, level):
    """Same keypoints, same per-cell ordinals and same threshold trajectory over a frame sequence
    (5 trials on the first frame, 6 afterwards: stereo_frontend.cpp computeFastCorners callers)."""
    args = _front_end_grids(svs, level)
    fg = svs.FastGrid(*args)
    og = oracle.fast_grid(*args)
    for it, f in enumerate(frames + frames[::-1]):
        img = f if level == 0 else _level1(f)
        trials = 5 if it == 0 else 6
        fg.set_image(img)
        xy, off = fg.detect_adaptively(trials)
        xo, oo = oracle.fast_detect_adaptively(img, og, trials)
        np.testing.assert_array_equal(off, oo)
        np.testing.assert_array_equal(xy, xo)
        assert [c[4] for c in fg.cell_list()] == [og.cells[k].thr for k in range(fg.ncells)]
    fg.close()


def test_edge_cases(svs, oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (61, 77), dtype=np.uint8)     # ragged size, pure noise: many corners
    fg = svs.FastGrid(77, 61, 50, 10, 15, 2, 2, max_keypoints=50)   # output smaller than the result
    fg.set_image(img)
    cells = fg.cell_list()
    xy, off = fg.detect(cells)
    xo, oo = oracle.fast_detect(img, cells, max_out=100000)
    np.testing.assert_array_equal(off, oo)                   # counts are exact even when truncated
    assert len(xy) == 50
    np.testing.assert_array_equal(xy, xo[:50])
    # cells too small to hold an interior, and a constant image
    xy, off = fg.detect([(0, 6, 0, 6, 10), (10, 17, 10, 17, 0)])
    xo, oo = oracle.fast_detect(img, [(0, 6, 0, 6, 10), (10, 17, 10, 17, 0)])
    np.testing.assert_array_equal(off, oo)
    fg.set_image(np.full((61, 77), 128, np.uint8))
    xy, off = fg.detect(cells)
    assert off[-1] == 0 and len(xy) == 0
    # trials = 0 emits nothing and keeps thresholds
    xy, off = fg.detect_adaptively(0)
    assert off[-1] == 0 and [c[4] for c in fg.cell_list()] == [15] * 4
    with pytest.raises(svs.SvsError):
        fg.detect([(0, 700, 0, 61, 10)])                     # cell outside the image
    fg.close()
