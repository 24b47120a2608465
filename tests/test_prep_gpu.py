"""GPU tests of the frame preprocessing kernels ("next" row) against OpenCV, which is
what the reference calls (cv::buildPyramid, convertTo, gpu::pyrDown, ksize-1 derivative filters), and
of the device-to-device hand-over into the dense tracker / FAST / matcher handles."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si

pytestmark = pytest.mark.gpu
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


@pytest.mark.parametrize("shape", [(480, 640), (97, 131)])
def test_pyramids_and_gradients_match_opencv(svs, shape):
    if shape == (480, 640):
        img = si.render_frame(np.zeros(3), 0.0)[0]
    else:
        img = np.random.default_rng(4).integers(0, 256, shape, dtype=np.uint8)   # odd sizes
    pp = svs.FramePreprocessor(shape[1], shape[0], 3)
    pp.process(img)
    u8, f32 = img, img.astype(np.float32) * np.float32(1.0 / 255.0)
    for l in range(3):
        if l > 0:
            u8, f32 = cv2.pyrDown(u8), cv2.pyrDown(f32)
        np.testing.assert_array_equal(pp.get_u8(l), u8)                       # integer path: bit-exact
        np.testing.assert_allclose(pp.get_f32(l, 0), f32, rtol=0, atol=2e-7)  # float taps: same order, <= 1 ulp
        # derivatives of OUR level image must equal OpenCV's filter on the same input exactly
        mine = pp.get_f32(l, 0)
        dx = cv2.Sobel(mine, cv2.CV_32F, 1, 0, ksize=1, borderType=cv2.BORDER_REPLICATE)
        dy = cv2.Sobel(mine, cv2.CV_32F, 0, 1, ksize=1, borderType=cv2.BORDER_REPLICATE)
        np.testing.assert_array_equal(pp.get_f32(l, 1), dx)
        np.testing.assert_array_equal(pp.get_f32(l, 2), dy)
    pp.close()


def test_device_handover_feeds_tracker_fast_and_matcher(svs, oracle):
    """Upload two raw frames only; pyramids/gradients stay on the device and are handed to the
    consumers by pointer.  Results must equal the host-fed path on the preprocessor's own outputs."""
    seq = si.sequence(2)
    cams = fi.level_cams()
    pa, pb = svs.FramePreprocessor(640, 480, 3), svs.FramePreprocessor(640, 480, 3)
    pa.process(seq[0]["img"])
    pb.process(seq[1]["img"])
    # dense tracker fed from device pointers
    dt = svs.DenseTracker(640, 480, 3)
    dt.set_disparity(seq[0]["disp"])
    for l in range(3):
        la, lb = pa.level(l), pb.level(l)
        dt.set_intrinsics(l, cams[l][0], cams[l][1], cams[l][2])
        dt.set_images_device(l, la["f32"], lb["f32"], lb["dx"], lb["dy"], lb["stride_f32"])
    dt.compute_point_cloud(I7, cams)
    T_dev, st_dev = dt.track(I7)
    # same tracker fed from host copies of the same images
    dt2 = svs.DenseTracker(640, 480, 3)
    dt2.set_disparity(seq[0]["disp"])
    for l in range(3):
        dt2.set_intrinsics(l, cams[l][0], cams[l][1], cams[l][2])
        dt2.set_images(l, pa.get_f32(l, 0), pb.get_f32(l, 0), pb.get_f32(l, 1), pb.get_f32(l, 2))
    dt2.compute_point_cloud(I7, cams)
    T_host, st_host = dt2.track(I7)
    assert st_dev["passes"] == st_host["passes"] and np.array_equal(T_dev, T_host)
    assert abs(T_dev[6] + 0.02) < 0.01
    # FAST from the device-resident uint8 level
    fg = svs.FastGrid(640, 480, 222, 74, 25, 3, 3)
    l0 = pb.level(0)
    fg.set_image_device(l0["u8"], l0["pitch_u8"], 640, 480)
    xy, off = fg.detect_adaptively(5)
    og = oracle.fast_grid(640, 480, 222, 74, 25, 3, 3)
    xo, oo = oracle.fast_detect_adaptively(seq[1]["img"], og, 5)
    np.testing.assert_array_equal(xy, xo)
    # matcher pyramids by pointer
    lv2 = [(640 >> l, 480 >> l, cams[l][0], cams[l][1], cams[l][2]) for l in range(2)]
    m = svs.GuidedMatcher(lv2)
    m.set_pyramid_device(0, [pa.level(l)["u8"] for l in range(2)], [pa.level(l)["pitch_u8"] for l in range(2)], I7)
    m.set_pyramid_device(-1, [pb.level(l)["u8"] for l in range(2)], [pb.level(l)["pitch_u8"] for l in range(2)])
    m.set_current_disparity(seq[1]["disp"])   # disparity still comes from the host
    content = np.concatenate([np.arange(off[c + 1] - off[c]) for c in range(9)]).astype(np.int32)
    m.set_features(0, xy, content)
    m.set_features(1, np.zeros((0, 2), np.int32), np.zeros(0, np.int32))
    pts = np.zeros(len(xy), svs.MATCH_POINT_DTYPE)
    d = seq[1]["disp"][xy[:, 1], xy[:, 0]]
    z = cams[0][0] * cams[0][3] / np.maximum(d, 1e-3)
    pts["xyz_anchor"] = np.stack([(xy[:, 0] - 320.0) / cams[0][0] * z, (xy[:, 1] - 240.0) / cams[0][0] * z, z], 1)
    pts["anchor_obs_pyr"] = xy
    res = m.match(I7, I7, pts, 4, 22, 10)
    assert res["predicted"].sum() > 0.9 * len(pts)
    for h in (pa, pb, dt, dt2, fg, m):
        h.close()
