"""Builds the per-level inputs of the dense tracker / matcher from raw uint8 frames the way the
reference's FrameGrabber::preprocessing does (frame_grabber.cpp:287-336): float image / 255,
5-tap pyrDown pyramid, 3-tap [-1 0 1] derivatives with replicated border; per-level cameras
(frame_grabber-impl.cpp:50-59).  Test/bench input preparation only (these producers are a
"next" row of SURVEY.md section 8f and use OpenCV here)."""
from __future__ import annotations

import numpy as np

from .synth import CAM_B, CAM_F, CAM_PX, CAM_PY

NUM_PYR_LEVELS = 3


def level_cams(f=CAM_F, px=CAM_PX, py=CAM_PY, b=CAM_B, nlevels=NUM_PYR_LEVELS):
    return [(f / (1 << l), px / (1 << l), py / (1 << l), b * (1 << l)) for l in range(nlevels)]


def float_pyramid(img8, nlevels=NUM_PYR_LEVELS):
    import cv2
    pyr = [img8.astype(np.float32) * np.float32(1.0 / 255.0)]
    for _ in range(1, nlevels):
        pyr.append(cv2.pyrDown(pyr[-1]))
    return pyr


def uint8_pyramid(img8, nlevels=NUM_PYR_LEVELS):
    import cv2
    pyr = [img8]
    for _ in range(1, nlevels):
        pyr.append(cv2.pyrDown(pyr[-1]))
    return pyr


def gradients(img32):
    import cv2
    dx = cv2.Sobel(img32, cv2.CV_32F, 1, 0, ksize=1, borderType=cv2.BORDER_REPLICATE)
    dy = cv2.Sobel(img32, cv2.CV_32F, 0, 1, ksize=1, borderType=cv2.BORDER_REPLICATE)
    return dx, dy
