"""CPU tests of the matcher oracle (oracle/match_oracle.c): quadtree query semantics and order
(quadtree.h), the literal integer score (matcher.cpp:73) and the affine warp."""
import numpy as np

from scavislam_b200 import frontend_inputs as fi
from scavislam_b200 import synth_images as si


def test_quadtree_query_is_window_filter_in_z_order(oracle):
    rng = np.random.default_rng(0)
    pts = np.unique(rng.integers(0, [640, 480], (3000, 2)), axis=0).astype(np.int32)
    rng.shuffle(pts)                      # insertion order must not matter
    qt = oracle.QuadTree(640, 480, pts, np.arange(len(pts), dtype=np.int32))
    for _ in range(50):
        u, v, r = int(rng.integers(-5, 645)), int(rng.integers(-5, 485)), int(rng.integers(1, 15))
        got = qt.query(u - r, v - r, 2 * r + 1, 2 * r + 1)
        inside = pts[(np.abs(pts[:, 0] - u) <= r) & (np.abs(pts[:, 1] - v) <= r)]
        assert sorted(map(tuple, got[:, :2])) == sorted(map(tuple, inside))
        # depth-first order xy, xY, Xy, XY == Z-order of recursive midpoint subdivision, x before y
        def zkey(p):
            tx, ty = (int(p[0]) << 12) // 640, (int(p[1]) << 12) // 480
            k = 0
            for b in range(12):
                k |= ((tx >> b) & 1) << (2 * b + 1) | ((ty >> b) & 1) << (2 * b)
            return k
        keys = [zkey(p) for p in got[:, :2]]
        assert keys == sorted(keys)


def test_znssd_literal_formula(oracle):
    from oracle import pyoracle as po
    import ctypes as C
    rng = np.random.default_rng(1)
    key = rng.integers(0, 256, 64, dtype=np.uint8)
    cur = rng.integers(0, 256, (8, 8), dtype=np.uint8)
    sa, saa = int(key.astype(np.int64).sum()), int((key.astype(np.int64) ** 2).sum())
    L = po.lib()
    L.omatch_znssd.argtypes = [po.c_up, po.c_up, C.c_int, C.c_int, C.c_int]
    z = L.omatch_znssd(key.ctypes.data_as(po.c_up), cur.ctypes.data_as(po.c_up), 8, sa, saa)
    b = cur.astype(np.int64).reshape(-1)
    sb, sbb, sab = int(b.sum()), int((b * b).sum()), int((b * key.astype(np.int64)).sum())
    num = sa * sa - 2 * sa * sb - sb * sb
    trunc = abs(num) // 64 * (1 if num >= 0 else -1)       # C integer division truncates toward zero
    assert z == saa - 2 * sab - sbb - trunc


def test_identity_warp_reproduces_the_patch(oracle):
    from oracle import pyoracle as po
    import ctypes as C
    img = si.render_frame(np.zeros(3), 0.0)[0]
    cams = fi.level_cams()
    L = po.OMatchLevel(640, 480, cams[0][0], cams[0][1], cams[0][2])
    I = np.array([0, 0, 0, 1, 0, 0, 0.0])
    out = np.zeros(100, np.uint8)
    key = np.array([200.0, 150.0])
    po.lib().omatch_warp_affine(img.ctypes.data_as(po.c_up), img.strides[0], C.byref(L), po._dp(I), 7.5, po._dp(key), 5,
                                out.ctypes.data_as(po.c_up))
    ref = img[145:155, 195:205].astype(np.int32)
    got = out.reshape(10, 10).astype(np.int32)
    # exact up to the reference's own truncation when the affine map lands a hair below an integer
    assert np.abs(got - ref).max() <= np.abs(np.diff(img[144:156, 194:206].astype(np.int32), axis=1)).max()
    assert (got == ref).mean() > 0.5
