"""Developer probe (GPU box): per-call wall time of the device-chained front-end frame."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scavislam_b200 import capi, frontend_inputs as fi, synth_images as si

seq = si.sequence(4)
cams = fi.level_cams()
I7 = np.array([0, 0, 0, 1, 0, 0, 0.0])
pps = [capi.FramePreprocessor(640, 480, 3) for _ in range(2)]
grids = [capi.FastGrid(640, 480, 222, 74, 25, 3, 3), capi.FastGrid(320, 240, 55, 18, 25, 3, 3)]
dt = capi.DenseTracker(640, 480, 3)
for l in range(3):
    dt.set_intrinsics(l, cams[l][0], cams[l][1], cams[l][2])
lv2 = [(640 >> l, 480 >> l, cams[l][0], cams[l][1], cams[l][2]) for l in range(2)]
mt = capi.GuidedMatcher(lv2)
T = {}
def lap(name, t0):
    t = time.perf_counter(); T[name] = T.get(name, 0) + (t - t0); return t
for rep in range(6):
    if rep == 2: T.clear()
    a, b = seq[rep % 3], seq[rep % 3 + 1]
    pp, pq = pps[rep & 1], pps[(rep & 1) ^ 1]
    t = time.perf_counter()
    pp.process(b["img"]); t = lap("prep.process", t)
    lv = [pp.level(l) for l in range(3)]; t = lap("prep.level", t)
    for l in range(2):
        grids[l].set_image_device(lv[l]["u8"], lv[l]["pitch_u8"], lv[l]["w"], lv[l]["h"]); t = lap("fast.set", t)
        xy, off = grids[l].detect_adaptively(6); t = lap("fast.detect", t)
    dt.set_disparity(a["disp"]); t = lap("dt.disp", t)
    dt.swap_prev_cur()
    for l in range(3):
        dt.set_images_device(l, None, lv[l]["f32"], lv[l]["dx"], lv[l]["dy"], lv[l]["stride_f32"])
    t = lap("dt.set_dev", t)
    dt.compute_point_cloud(I7, cams); t = lap("dt.cloud", t)
    Tt, st = dt.track(I7); t = lap("dt.track", t)
    lq = [pq.level(l) for l in range(2)]
    mt.set_pyramid_device(0, [x["u8"] for x in lq], [x["pitch_u8"] for x in lq], I7)
    mt.set_pyramid_device(-1, [x["u8"] for x in lv[:2]], [x["pitch_u8"] for x in lv[:2]]); t = lap("mt.pyr", t)
    mt.set_current_disparity(b["disp"]); t = lap("mt.disp", t)
print({k: round(v / 4 * 1e3, 3) for k, v in T.items()}, "ms per frame; passes", st["passes"])
