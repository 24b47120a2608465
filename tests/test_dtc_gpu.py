"""GPU parity of the non-CUDA-build dense tracker semantics (svs_denseTrackingCpu,
svs_computeDensePointCloudCpu) against oracle/dt_oracle.c (odtc_*), through the C ABI.  Point cloud:
bit-exact.  Tracking: identical pass counts, chi2 to 1e-9 relative (fixed-order FP64 sums vs the oracle's
sequential FP64 sums), pose to 1e-9."""
import numpy as np
import pytest

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si
from test_dtc_oracle import levels

pytestmark = pytest.mark.gpu
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])


def _tracker(svs, lv, seq, T_cloud, cams):
    t = svs.DenseTrackerCpuVariant(640, 480, len(lv))
    t.set_disparity(seq[0]["disp"])
    for l, L in enumerate(lv):
        t.set_prev_u8(l, L["prev_u8"])
        t.set_cur(l, L["cur"], L["dx"], L["dy"])
    t.compute_point_cloud(T_cloud, cams)
    return t


def test_point_cloud_is_bit_exact(svs, oracle):
    T = oracle.se3_exp(np.array([0.05, -0.02, 0.03, 0.01, 0.02, -0.015]))
    for Tc in (I7, T):
        lv, seq, cams = levels(oracle, 3, Tc)
        t = _tracker(svs, lv, seq, Tc, cams)
        for l in range(3):
            np.testing.assert_array_equal(t.point_cloud(l), lv[l]["cloud"])
        t.close()


@pytest.mark.parametrize("start", [np.zeros(6), np.array([0.01, -0.005, 0.015, 0.002, -0.003, 0.001])])
def test_tracking_matches_oracle(svs, oracle, start):
    lv, seq, cams = levels(oracle, 3)
    T0 = oracle.se3_exp(start)
    t = _tracker(svs, lv, seq, I7, cams)
    T_g, sg = t.track(T0, cams)
    T_o, so = oracle.dtc_track(lv, T0)
    t.close()
    assert sg["passes"] == so["passes"]
    np.testing.assert_allclose(sg["chi2"], so["chi2"], rtol=1e-9)
    np.testing.assert_allclose(T_g, T_o, rtol=1e-9, atol=1e-11)
    assert abs(T_g[6] + 0.02) < 0.01


def test_device_inputs_and_uploaded_cloud(svs, oracle):
    """Planes handed over on the device (svs_prep_level) and a point cloud set by the caller."""
    lv, seq, cams = levels(oracle, 3)
    pa, pb = svs.FramePreprocessor(640, 480, 3), svs.FramePreprocessor(640, 480, 3)
    pa.process(seq[0]["img"]); pb.process(seq[1]["img"])
    t = svs.DenseTrackerCpuVariant(640, 480, 3)
    for l in range(3):
        a, b = pa.level(l), pb.level(l)
        t.set_prev_u8_device(l, a["u8"], a["pitch_u8"])
        t.set_cur_device(l, b["f32"], b["dx"], b["dy"], b["stride_f32"])
        t.set_point_cloud(l, lv[l]["cloud"])
    T_g, sg = t.track(I7, cams)
    # the oracle on the preprocessor's own planes (its float pyramid may differ from OpenCV's in the last bit)
    lv2 = [dict(L, prev_u8=pa.get_u8(l), cur=pb.get_f32(l, 0), dx=pb.get_f32(l, 1), dy=pb.get_f32(l, 2)) for l, L in enumerate(lv)]
    T_o, so = oracle.dtc_track(lv2, I7)
    assert sg["passes"] == so["passes"]
    np.testing.assert_allclose(T_g, T_o, rtol=1e-9, atol=1e-11)
    for h in (pa, pb, t):
        h.close()


def test_level_sizes_must_be_multiples_of_four(svs):
    with pytest.raises(svs.SvsError):            # 100 >> 1 = 50: the reference asserts (dense_tracking.cpp:42-43)
        svs.DenseTrackerCpuVariant(100, 96, 2)
