"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
scavislam_b200/ never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_ubyte)


c_fp = C.POINTER(C.c_float)
ODT_MAX_LEVELS = 8


class ODtLevel(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("stride", C.c_int), ("cloud_stride", C.c_int),
                ("f", C.c_float), ("px", C.c_float), ("py", C.c_float),
                ("prev", c_fp), ("cur", c_fp), ("dx", c_fp), ("dy", c_fp), ("cloud", c_fp)]


class ODtStats(C.Structure):
    _fields_ = [("chi2", C.c_double * ODT_MAX_LEVELS), ("passes", C.c_int * ODT_MAX_LEVELS)]


OMATCH_MAX_LEVELS = 4


class OMatchLevel(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("f", C.c_double), ("px", C.c_double), ("py", C.c_double)]


class OMatchFrame(C.Structure):
    _fields_ = [("levels", OMatchLevel * OMATCH_MAX_LEVELS), ("pyr", c_up * OMATCH_MAX_LEVELS),
                ("pitch", C.c_int * OMATCH_MAX_LEVELS), ("disp", c_fp), ("disp_pitch", C.c_int),
                ("trees", C.c_void_p * OMATCH_MAX_LEVELS)]


class OMatchKeyframe(C.Structure):
    _fields_ = [("T_me_from_w", C.c_double * 7), ("pyr", c_up * OMATCH_MAX_LEVELS), ("pitch", C.c_int * OMATCH_MAX_LEVELS)]


class OBAProblem(C.Structure):
    _fields_ = [("P", C.c_int), ("L", C.c_int), ("E", C.c_int), ("C", C.c_int),
                ("pose_qt", c_dp), ("fixed", c_up), ("psi", c_dp),
                ("e_point", c_ip), ("e_pose", c_ip), ("e_anchor", c_ip),
                ("e_obs", c_dp), ("e_info", c_dp),
                ("c_i", c_ip), ("c_j", c_ip), ("c_T", c_dp), ("c_Lambda", c_dp),
                ("f", C.c_double), ("px", C.c_double), ("py", C.c_double), ("b", C.c_double)]


OBA_MAX_ITERS = 64


class OBAStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials_total", C.c_int),
                ("chi2_init", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double),
                ("chi2_iter", C.c_double * OBA_MAX_ITERS), ("lambda_iter", C.c_double * OBA_MAX_ITERS),
                ("trials_iter", C.c_int * OBA_MAX_ITERS), ("nnzb_S", C.c_int), ("nnzb_L", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        for name, nd in (("oba_se3_exp", 2), ("oba_se3_log", 2), ("oba_se3_inv", 2), ("oba_se3_adj", 2),
                         ("oba_invert_depth", 2), ("oba_se3_mul", 3), ("oba_se3_act", 3)):
            getattr(L, name).argtypes = [c_dp] * nd
            getattr(L, name).restype = None
        L.oba_edge_error.argtypes = [c_dp] * 6
        L.oba_edge_jacobians.argtypes = [c_dp] * 7
        L.oba_posepose_error.argtypes = [c_dp] * 4
        L.oba_posepose_jacobians.argtypes = [c_dp] * 4
        L.oba_chi2.argtypes = [C.POINTER(OBAProblem), C.c_int, C.c_double]
        L.oba_chi2.restype = C.c_double
        L.oba_reduced_system.argtypes = [C.POINTER(OBAProblem), C.c_int, C.c_double, C.c_double, c_dp, c_dp]
        L.oba_reduced_system.restype = C.c_double
        L.oba_full_system.argtypes = [C.POINTER(OBAProblem), C.c_int, C.c_double, c_dp, c_dp]
        L.oba_full_system.restype = C.c_double
        L.oba_optimize.argtypes = [C.POINTER(OBAProblem), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                   c_dp, c_dp, C.POINTER(OBAStats)]
        L.oba_optimize.restype = C.c_int
        L.ofast_detect_roi.argtypes = [c_up, C.c_int] + [C.c_int] * 5 + [c_ip, C.c_int]
        L.ofast_detect.argtypes = [c_up, C.c_int, C.c_void_p, C.c_int, c_ip, C.c_int, c_ip]
        L.ofast_detect_adaptively.argtypes = [c_up, C.c_int, C.c_void_p, C.c_int, c_ip, C.c_int, c_ip]
        L.ofast_score.argtypes = [c_up, C.c_int, C.c_int, C.c_int]
        L.ofast_is_corner.argtypes = [c_up, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ofast_grid_init.argtypes = [C.c_void_p] + [C.c_int] * 9
        L.ofast_grid_init.restype = None
        c_fp = C.POINTER(C.c_float)
        L.odt_pass.argtypes = [C.POINTER(ODtLevel), c_dp, C.c_int, c_dp, c_dp, c_dp, c_ip]
        L.odt_pass.restype = None
        L.odt_residual_image.argtypes = [C.POINTER(ODtLevel), c_dp, C.c_int, c_fp]
        L.odt_residual_image.restype = None
        L.odt_track.argtypes = [C.POINTER(ODtLevel), C.c_int, c_dp, C.c_int, C.POINTER(ODtStats)]
        L.odt_track.restype = None
        L.odt_point_cloud.argtypes = [c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]
        L.odt_point_cloud.restype = None
        L.odt_make_TQ.argtypes = [c_dp, C.c_double, C.c_double, C.c_double, C.c_double, c_fp]
        L.odt_make_TQ.restype = None
        L.omatch_tree_build.argtypes = [C.c_int, C.c_int, c_ip, c_ip, C.c_int]
        L.omatch_tree_build.restype = C.c_void_p
        L.omatch_tree_free.argtypes = [C.c_void_p]
        L.omatch_tree_free.restype = None
        L.omatch_tree_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, C.c_int]
        L.omatch_warp_affine.argtypes = [c_up, C.c_int, C.POINTER(OMatchLevel), c_dp, C.c_double, c_dp, C.c_int, c_up]
        L.omatch_warp_affine.restype = None
        L.omatch_match.argtypes = [C.POINTER(OMatchFrame), C.POINTER(OMatchKeyframe), C.c_int, c_dp, c_dp, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def as_c_problem(pb):
    """pb: scavislam_b200.synth.BAProblem-like (numpy arrays).  Returns (struct, keepalive)."""
    keep = dict(
        pose_qt=np.ascontiguousarray(pb.pose_qt, np.float64), fixed=np.ascontiguousarray(pb.fixed, np.uint8),
        psi=np.ascontiguousarray(pb.psi, np.float64),
        e_point=np.ascontiguousarray(pb.e_point, np.int32), e_pose=np.ascontiguousarray(pb.e_pose, np.int32),
        e_anchor=np.ascontiguousarray(pb.e_anchor, np.int32),
        e_obs=np.ascontiguousarray(pb.e_obs, np.float64), e_info=np.ascontiguousarray(pb.e_info, np.float64),
        c_i=np.ascontiguousarray(pb.c_i, np.int32), c_j=np.ascontiguousarray(pb.c_j, np.int32),
        c_T=np.ascontiguousarray(pb.c_T, np.float64), c_Lambda=np.ascontiguousarray(pb.c_Lambda, np.float64))
    k = keep
    s = OBAProblem(pb.P, pb.L, pb.E, pb.C, _dp(k["pose_qt"]), k["fixed"].ctypes.data_as(c_up), _dp(k["psi"]),
                   _ip(k["e_point"]), _ip(k["e_pose"]), _ip(k["e_anchor"]), _dp(k["e_obs"]), _dp(k["e_info"]),
                   _ip(k["c_i"]), _ip(k["c_j"]), _dp(k["c_T"]), _dp(k["c_Lambda"]),
                   float(pb.cam[0]), float(pb.cam[1]), float(pb.cam[2]), float(pb.cam[3]))
    return s, keep


def f64(*shape):
    return np.zeros(shape, np.float64)


def se3_exp(d):
    T = f64(7); lib().oba_se3_exp(_dp(np.ascontiguousarray(d, np.float64)), _dp(T)); return T


def se3_log(T):
    d = f64(6); lib().oba_se3_log(_dp(np.ascontiguousarray(T, np.float64)), _dp(d)); return d


def se3_mul(A, B):
    o = f64(7)
    lib().oba_se3_mul(_dp(np.ascontiguousarray(A, np.float64)), _dp(np.ascontiguousarray(B, np.float64)), _dp(o))
    return o


def se3_inv(A):
    o = f64(7); lib().oba_se3_inv(_dp(np.ascontiguousarray(A, np.float64)), _dp(o)); return o


def se3_act(A, x):
    o = f64(3)
    lib().oba_se3_act(_dp(np.ascontiguousarray(A, np.float64)), _dp(np.ascontiguousarray(x, np.float64)), _dp(o))
    return o


def se3_adj(A):
    o = f64(6, 6); lib().oba_se3_adj(_dp(np.ascontiguousarray(A, np.float64)), _dp(o)); return o


def edge_error(cam, Tp, Ta, psi, obs):
    a = [np.ascontiguousarray(x, np.float64) for x in (cam, Tp, Ta, psi, obs)]
    e = f64(3); lib().oba_edge_error(*[_dp(x) for x in a], _dp(e)); return e


def edge_jacobians(cam, Tp, Ta, psi):
    a = [np.ascontiguousarray(x, np.float64) for x in (cam, Tp, Ta, psi)]
    Jpsi, Jp, Ja = f64(3, 3), f64(3, 6), f64(3, 6)
    lib().oba_edge_jacobians(*[_dp(x) for x in a], _dp(Jpsi), _dp(Jp), _dp(Ja))
    return Jpsi, Jp, Ja


def posepose_error(T21, T1, T2):
    a = [np.ascontiguousarray(x, np.float64) for x in (T21, T1, T2)]
    e = f64(6); lib().oba_posepose_error(*[_dp(x) for x in a], _dp(e)); return e


def posepose_jacobians(T21, err):
    a = [np.ascontiguousarray(x, np.float64) for x in (T21, err)]
    Ji, Jj = f64(6, 6), f64(6, 6)
    lib().oba_posepose_jacobians(*[_dp(x) for x in a], _dp(Ji), _dp(Jj)); return Ji, Jj


def chi2(pb, robust=True, delta=1.0):
    s, keep = as_c_problem(pb)
    return lib().oba_chi2(C.byref(s), int(robust), float(delta))


def reduced_system(pb, robust=True, delta=1.0, lam=50.0):
    s, keep = as_c_problem(pb)
    n = 6 * pb.P
    S, bs = f64(n, n), f64(n)
    chi = lib().oba_reduced_system(C.byref(s), int(robust), float(delta), float(lam), _dp(S), _dp(bs))
    return S, bs, chi


def full_system(pb, robust=True, delta=1.0):
    s, keep = as_c_problem(pb)
    n = 6 * pb.P + 3 * pb.L
    H, b = f64(n, n), f64(n)
    chi = lib().oba_full_system(C.byref(s), int(robust), float(delta), _dp(H), _dp(b))
    return H, b, chi


def set_threads(n):
    """Timing variant only: OpenMP threads of the oracle's build and Schur loops (1 = sequential restatement)."""
    L = lib(); L.oba_set_threads.argtypes = [C.c_int]; L.oba_set_threads.restype = None
    L.oba_set_threads(int(n))


def optimize(pb, num_iters, robust=True, delta=1.0, lambda_init=50.0, max_trials=5):
    s, keep = as_c_problem(pb)
    poses, psi = f64(pb.P, 7), f64(pb.L, 3)
    st = OBAStats()
    it = lib().oba_optimize(C.byref(s), int(num_iters), int(robust), float(delta), float(lambda_init),
                            int(max_trials), _dp(poses), _dp(psi), C.byref(st))
    stats = dict(iterations=it, trials_total=st.trials_total, chi2_init=st.chi2_init, chi2_final=st.chi2_final,
                 lambda_final=st.lambda_final, chi2_iter=list(st.chi2_iter[:max(it, 0)]),
                 lambda_iter=list(st.lambda_iter[:max(it, 0)]), trials_iter=list(st.trials_iter[:max(it, 0)]),
                 nnzb_S=st.nnzb_S, nnzb_L=st.nnzb_L)
    return poses, psi, stats


# ------------------------------------------------------------------ FAST grid (fast_oracle.c)

class OFastCell(C.Structure):
    _fields_ = [("u0", C.c_int), ("u1", C.c_int), ("v0", C.c_int), ("v1", C.c_int), ("thr", C.c_int)]


OFAST_MAX_CELLS = 64


class OFastGrid(C.Structure):
    _fields_ = [("grid_w", C.c_int), ("grid_h", C.c_int), ("fast_min", C.c_int), ("fast_max", C.c_int),
                ("min_inner", C.c_int), ("min_outer", C.c_int), ("max_inner", C.c_int), ("max_outer", C.c_int),
                ("cells", OFastCell * OFAST_MAX_CELLS)]


def fast_grid(img_w, img_h, num_features_per_cell, boundary_per_cell, fast_thr, grid_w, grid_h,
              fast_min=10, fast_max=40):
    g = OFastGrid()
    lib().ofast_grid_init(C.byref(g), img_w, img_h, num_features_per_cell, boundary_per_cell, fast_thr,
                          grid_w, grid_h, fast_min, fast_max)
    return g


def _u8(img):
    img = np.ascontiguousarray(img, np.uint8)
    return img, img.ctypes.data_as(c_up), img.strides[0]


def fast_detect_roi(img, u0, u1, v0, v1, thr, max_out=200000):
    img, p, pitch = _u8(img)
    out = np.zeros((max_out, 2), np.int32)
    n = lib().ofast_detect_roi(p, pitch, u0, u1, v0, v1, thr, _ip(out), max_out)
    return out[:min(n, max_out)].copy()


def fast_detect(img, cells, max_out=200000):
    """cells: list of (u0,u1,v0,v1,thr).  Returns (xy[n,2], cell_off[ncells+1])."""
    img, p, pitch = _u8(img)
    arr = (OFastCell * len(cells))(*[OFastCell(*c) for c in cells])
    out = np.zeros((max_out, 2), np.int32)
    off = np.zeros(len(cells) + 1, np.int32)
    n = lib().ofast_detect(p, pitch, arr, len(cells), _ip(out), max_out, _ip(off))
    return out[:min(n, max_out)].copy(), off


def fast_detect_adaptively(img, grid, trials, max_out=200000):
    """Mutates grid (cell thresholds), like FastGrid::detectAdaptively."""
    img, p, pitch = _u8(img)
    out = np.zeros((max_out, 2), np.int32)
    off = np.zeros(grid.grid_w * grid.grid_h + 1, np.int32)
    n = lib().ofast_detect_adaptively(p, pitch, C.byref(grid), trials, _ip(out), max_out, _ip(off))
    return out[:min(n, max_out)].copy(), off


def fast_score_map(img):
    img, p, pitch = _u8(img)
    h, w = img.shape
    s = np.full((h, w), -1, np.int32)
    L = lib()
    for y in range(3, h - 3):
        for x in range(3, w - 3):
            s[y, x] = L.ofast_score(p, pitch, x, y)
    return s


# ------------------------------------------------------------------ dense tracker (dt_oracle.c)

def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def dt_levels(levels):
    """levels: list of dicts(prev, cur, dx, dy, cloud[h,w,4], f, px, py).  Returns (array, keepalive)."""
    arr = (ODtLevel * len(levels))()
    keep = []
    for i, lv in enumerate(levels):
        ims = {k: _f32(lv[k]) for k in ("prev", "cur", "dx", "dy", "cloud")}
        keep.append(ims)
        h, w = ims["prev"].shape
        arr[i] = ODtLevel(w, h, w, w, float(lv["f"]), float(lv["px"]), float(lv["py"]),
                          *[ims[k].ctypes.data_as(c_fp) for k in ("prev", "cur", "dx", "dy", "cloud")])
    return arr, keep


def dt_pass(level, T, exact=False, want_jac=True):
    arr, keep = dt_levels([level])
    T = np.ascontiguousarray(T, np.float64)
    chi = C.c_double()
    n = C.c_int()
    H, b = np.zeros(21), np.zeros(6)
    lib().odt_pass(arr, _dp(T), int(exact), C.byref(chi), _dp(H) if want_jac else None, _dp(b) if want_jac else None,
                   C.byref(n))
    return chi.value, H, b, n.value


def dt_residual_image(level, T, exact=False):
    arr, keep = dt_levels([level])
    T = np.ascontiguousarray(T, np.float64)
    h, w = keep[0]["prev"].shape
    out = np.zeros((h, w, 4), np.float32)
    lib().odt_residual_image(arr, _dp(T), int(exact), out.ctypes.data_as(c_fp))
    return out


def dt_track(levels, T, exact=False):
    arr, keep = dt_levels(levels)
    T = np.ascontiguousarray(T, np.float64).copy()
    st = ODtStats()
    lib().odt_track(arr, len(levels), _dp(T), int(exact), C.byref(st))
    return T, dict(chi2=list(st.chi2[:len(levels)]), passes=list(st.passes[:len(levels)]))


def dt_point_cloud(T, cam, disp, level, w, h):
    """cam = (f, px, py, b) of the level camera; disp = level-0 float disparity."""
    TQ = np.zeros(16, np.float32)
    T = np.ascontiguousarray(T, np.float64)
    lib().odt_make_TQ(_dp(T), float(cam[0]), float(cam[1]), float(cam[2]), float(cam[3]), TQ.ctypes.data_as(c_fp))
    d = _f32(disp)
    out = np.zeros((h, w, 4), np.float32)
    lib().odt_point_cloud(TQ.ctypes.data_as(c_fp), d.ctypes.data_as(c_fp), w, h, d.shape[1], w, 1 << level,
                          out.ctypes.data_as(c_fp))
    return out


# ------------------------------------------------------------------ guided matcher (match_oracle.c)

MATCH_RESULT_DTYPE = np.dtype([("predicted", "i4"), ("textured", "i4"), ("matched", "i4"), ("n_candidates", "i4"),
                               ("index", "i4"), ("min_dist", "i4"), ("uv_pyr", "i4", 2), ("obs", "f8", 3),
                               ("xyz_actkey", "f8", 3)])
MATCH_POINT_DTYPE = np.dtype([("keyframe", "i4"), ("anchor_level", "i4"), ("xyz_anchor", "f8", 3),
                              ("anchor_obs_pyr", "f8", 2)])


class QuadTree:
    """QuadTree<int>(Rectangle(0, 0, w, h), 1) filled with (x, y) -> content (quadtree.h)."""

    def __init__(self, w, h, xy, content=None):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        c = None if content is None else np.ascontiguousarray(content, np.int32)
        self.ptr = lib().omatch_tree_build(w, h, _ip(xy), None if c is None else _ip(c), len(xy))

    def query(self, x, y, w, h, max_out=4096):
        out = np.zeros((max_out, 3), np.int32)
        n = lib().omatch_tree_query(self.ptr, x, y, w, h, _ip(out), max_out)
        return out[:min(n, max_out)].copy()

    def __del__(self):
        try:
            lib().omatch_tree_free(self.ptr)
        except Exception:
            pass


def match(levels, cur_pyr, disp, trees, keyframes, T_cur_from_actkey, T_actkey_from_w, points, search_radius,
          thr_mean, thr_std):
    """levels: [(w,h,f,px,py)], cur_pyr: uint8 images, trees: [QuadTree], keyframes: [(T, pyr)]."""
    fr = OMatchFrame()
    keep = []
    for l, (w, h, f, px, py) in enumerate(levels):
        fr.levels[l] = OMatchLevel(int(w), int(h), float(f), float(px), float(py))
        im = np.ascontiguousarray(cur_pyr[l], np.uint8); keep.append(im)
        fr.pyr[l] = im.ctypes.data_as(c_up); fr.pitch[l] = im.strides[0]
        fr.trees[l] = trees[l].ptr
    d = _f32(disp); keep.append(d)
    fr.disp = d.ctypes.data_as(c_fp); fr.disp_pitch = d.shape[1]
    kfs = (OMatchKeyframe * len(keyframes))()
    for k, (T, pyr) in enumerate(keyframes):
        for i in range(7):
            kfs[k].T_me_from_w[i] = float(T[i])
        for l in range(len(levels)):
            im = np.ascontiguousarray(pyr[l], np.uint8); keep.append(im)
            kfs[k].pyr[l] = im.ctypes.data_as(c_up); kfs[k].pitch[l] = im.strides[0]
    pts = np.ascontiguousarray(points, MATCH_POINT_DTYPE)
    out = np.zeros(len(pts), MATCH_RESULT_DTYPE)
    Ta = np.ascontiguousarray(T_cur_from_actkey, np.float64)
    Tb = np.ascontiguousarray(T_actkey_from_w, np.float64)
    lib().omatch_match(C.byref(fr), kfs, len(keyframes), _dp(Ta), _dp(Tb), pts.ctypes.data, len(pts),
                       int(search_radius), int(thr_mean), int(thr_std), out.ctypes.data)
    return out


# ---------------------------------------------------------------- motion-only LM (pose_oracle.c)
class OPOStats(C.Structure):
    _fields_ = [("initial_chi2", C.c_double), ("chi2", C.c_double), ("max_err", C.c_double), ("num_obs", C.c_int),
                ("iterations", C.c_int), ("trials", C.c_int), ("nan_error", C.c_int)]


def pose_map(cam, T, xyz):
    o = f64(3)
    L = lib(); L.opo_map.argtypes = [c_dp] * 4; L.opo_map.restype = None
    L.opo_map(_dp(np.ascontiguousarray(cam, np.float64)), _dp(np.ascontiguousarray(T, np.float64)),
              _dp(np.ascontiguousarray(xyz, np.float64)), _dp(o))
    return o


def pose_frame_jac(cam, T, xyz):
    o = f64(3, 6)
    L = lib(); L.opo_frame_jac.argtypes = [c_dp] * 4; L.opo_frame_jac.restype = None
    L.opo_frame_jac(_dp(np.ascontiguousarray(cam, np.float64)), _dp(np.ascontiguousarray(T, np.float64)),
                    _dp(np.ascontiguousarray(xyz, np.float64)), _dp(o))
    return o


def calc_fast_motion_only(obs_point_id, obs_uvu, point_xyz, cam, T_frame, robust_kernel=True, kernel_param=1.0,
                          num_iter=50, initial_mu=-1.0, tau=0.00001):
    L = lib()
    L.opo_calc_fast_motion_only.argtypes = [C.c_int, C.POINTER(C.c_int), c_dp, c_dp, c_dp, C.c_int, C.c_double, C.c_int,
                                            C.c_double, C.c_double, c_dp, C.POINTER(OPOStats)]
    L.opo_calc_fast_motion_only.restype = None
    pid = np.ascontiguousarray(obs_point_id, np.int32)
    obs = np.ascontiguousarray(obs_uvu, np.float64)
    xyz = np.ascontiguousarray(point_xyz, np.float64)
    T = np.array(T_frame, np.float64).copy()
    st = OPOStats()
    L.opo_calc_fast_motion_only(len(pid), pid.ctypes.data_as(C.POINTER(C.c_int)), _dp(obs), _dp(xyz),
                                _dp(np.ascontiguousarray(cam, np.float64)), int(robust_kernel), float(kernel_param),
                                int(num_iter), float(initial_mu), float(tau), _dp(T), C.byref(st))
    return T, dict(initial_chi2=st.initial_chi2, chi2=st.chi2, max_err=st.max_err, num_obs=st.num_obs,
                   iterations=st.iterations, trials=st.trials, nan_error=st.nan_error)


# ---------------------------------------------------------------- dense tracker, non-CUDA build of the reference (a18)
class ODtcLevel(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("stride", C.c_int), ("pitch_u8", C.c_int),
                ("f", C.c_double), ("px", C.c_double), ("py", C.c_double), ("b", C.c_double),
                ("prev_u8", C.c_void_p), ("cur", c_fp), ("dx", c_fp), ("dy", c_fp), ("cloud", c_fp)]


def dtc_point_cloud(T, cam, disp, level, w, h):
    """cam = (f, px, py, b) of the level camera, (w, h) = level size; returns (h/4, w/4, 4) float32."""
    L = lib()
    L.odtc_point_cloud.argtypes = [c_dp, C.c_double, C.c_double, C.c_double, C.c_double, c_fp, C.c_int, C.c_int, C.c_int,
                                   C.c_int, c_fp]
    L.odtc_point_cloud.restype = None
    d = _f32(disp)
    out = np.zeros((h // 4, w // 4, 4), np.float32)
    L.odtc_point_cloud(_dp(np.ascontiguousarray(T, np.float64)), float(cam[0]), float(cam[1]), float(cam[2]), float(cam[3]),
                       d.ctypes.data_as(c_fp), d.shape[1], int(level), int(w), int(h), out.ctypes.data_as(c_fp))
    return out


def dtc_levels(levels):
    """levels: list of dicts(prev_u8, cur, dx, dy, cloud[h/4, w/4, 4], cam=(f, px, py, b))."""
    arr = (ODtcLevel * len(levels))()
    keep = []
    for i, lv in enumerate(levels):
        p8 = np.ascontiguousarray(lv["prev_u8"], np.uint8)
        ims = {k: _f32(lv[k]) for k in ("cur", "dx", "dy", "cloud")}
        keep.append((p8, ims))
        h, w = ims["cur"].shape
        f, px, py, b = [float(x) for x in lv["cam"][:4]]
        arr[i] = ODtcLevel(w, h, w, p8.strides[0], f, px, py, b, p8.ctypes.data,
                           *[ims[k].ctypes.data_as(c_fp) for k in ("cur", "dx", "dy", "cloud")])
    return arr, keep


def dtc_pass(level, T):
    L = lib()
    L.odtc_pass.argtypes = [C.POINTER(ODtcLevel), c_dp, c_dp, c_dp, c_dp, C.POINTER(C.c_int)]
    L.odtc_pass.restype = None
    arr, keep = dtc_levels([level])
    chi, n = np.zeros(1), C.c_int()
    H, b = np.zeros(21), np.zeros(6)
    L.odtc_pass(arr, _dp(np.ascontiguousarray(T, np.float64)), _dp(chi), _dp(H), _dp(b), C.byref(n))
    return float(chi[0]), H, b, n.value


def dtc_track(levels, T):
    L = lib()
    L.odtc_track.argtypes = [C.POINTER(ODtcLevel), C.c_int, c_dp, C.POINTER(ODtStats)]
    L.odtc_track.restype = None
    arr, keep = dtc_levels(levels)
    T = np.ascontiguousarray(T, np.float64).copy()
    st = ODtStats()
    L.odtc_track(arr, len(levels), _dp(T), C.byref(st))
    return T, dict(chi2=list(st.chi2[:len(levels)]), passes=list(st.passes[:len(levels)]))


# ---------------------------------------------------------------- pose-pose constraint weights (constraint_oracle.c)
def compute_constraints(poses, feat_ptr, feat_point, point_anchor, xyz_anchor, v1, v2):
    L = lib()
    c_ip_ = C.POINTER(C.c_int)
    L.occ_compute_constraints.argtypes = [C.c_int, c_dp, c_ip_, c_ip_, C.c_int, c_ip_, c_dp, C.c_int, c_ip_, c_ip_, c_dp, c_dp,
                                          c_ip_]
    L.occ_compute_constraints.restype = None
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
    fp, fpt = np.ascontiguousarray(feat_ptr, np.int32), np.ascontiguousarray(feat_point, np.int32)
    pa = np.ascontiguousarray(point_anchor, np.int32)
    xyz = np.ascontiguousarray(xyz_anchor, np.float64).reshape(-1, 3)
    v1, v2 = np.ascontiguousarray(v1, np.int32), np.ascontiguousarray(v2, np.int32)
    n = len(v1)
    T, Lam, ns = np.zeros((n, 7)), np.zeros((n, 36)), np.zeros(n, np.int32)
    ip = lambda a: a.ctypes.data_as(c_ip_)
    L.occ_compute_constraints(len(poses), _dp(poses), ip(fp), ip(fpt), len(pa), ip(pa), _dp(xyz), n, ip(v1), ip(v2), _dp(T),
                              _dp(Lam), ip(ns))
    return T, Lam.reshape(n, 6, 6), ns


# ---------------------------------------------------------------- window assembly (plain Python restatement)
def copy_data_to_g2o(m, window_vertex, active_point):
    """SlamGraph::copyDataToG2o / copyPosesToG2o / addPointToG2o / addObsToG2o (reference slam_graph.cpp:907-1032,
    slam_graph-impl.cpp:29-126) on the tables of scavislam_b200.synth_graph.make_map: returns dict(pose_qt, psi,
    e_point, e_pose, e_anchor, e_obs, e_info).  TEST INFRASTRUCTURE ONLY."""
    win = {int(v): i for i, v in enumerate(window_vertex)}
    pose_qt = np.array([m["poses"][v] for v in window_vertex], np.float64)
    psi, ep, es, ea, obs, info = [], [], [], [], [], []
    for l, p in enumerate(active_point):
        x, y, z = m["xyz_anchor"][p]
        psi.append([x / z, y / z, 1.0 / z])                       # invert_depth (maths_utils.h:66-69)
        a = win[int(m["point_anchor"][p])]
        for i in range(m["vis_ptr"][p], m["vis_ptr"][p + 1]):     # for pose_id in p.vis_set (ascending ids)
            v = int(m["vis_pose"][i])
            if v not in win:                                       # :1004-1005
                continue
            s = (1.0 / (1 << int(m["feat_level"][i]))) ** 2        # Po2(pyrFromZero_d(1., level)), :1010-1015
            ep.append(l); es.append(win[v]); ea.append(a)
            obs.append(m["feat_center"][i]); info.append([s, s, 0.333 * 0.333])
    return dict(pose_qt=pose_qt, psi=np.array(psi, np.float64).reshape(-1, 3), e_point=np.array(ep, np.int32),
                e_pose=np.array(es, np.int32), e_anchor=np.array(ea, np.int32),
                e_obs=np.array(obs, np.float64).reshape(-1, 3), e_info=np.array(info, np.float64).reshape(-1, 3))


# ---------------------------------------------------------------- window selection and map growth (plain Python)
def compute_double_window(nbr_ptr, nbr_id, root, inner_window_size, double_window_size):
    """SlamGraph::computeInitialDoubleWin (reference slam_graph.cpp:556-598): breadth-first from `root`; a vertex that
    is popped and not yet in the window joins it (INNER while fewer than inner_window_size vertices are in, OUTER
    afterwards) and pushes its neighbours, strongest first (nbr_id[nbr_ptr[v]:nbr_ptr[v+1]] is that order).
    Returns {vertex: 1 (INNER) | 2 (OUTER)}.  TEST INFRASTRUCTURE ONLY."""
    from collections import deque
    win = {}
    q = deque([int(root)])
    while len(win) < double_window_size and q:
        v = q.popleft()
        if v in win:                                              # "Avoid cycles!"
            continue
        win[v] = 1 if len(win) < inner_window_size else 2
        for i in range(nbr_ptr[v], nbr_ptr[v + 1]):
            q.append(int(nbr_id[i]))
    return win


def compute_active_points(m, nbr_ptr, nbr_id, win):
    """SlamGraph::computeActivePointsAndExtendOuterWindow (slam_graph.cpp:600-663) on the tables of synth_graph:
    a point seen by an INNER frame is active when its anchor frame is in the window, or when that inner frame has a
    direct edge to the anchor frame -- the anchor then joins the outer window.  Returns (sorted active point ids,
    window dict including the extension)."""
    edges = set()
    for v in range(len(nbr_ptr) - 1):
        for i in range(nbr_ptr[v], nbr_ptr[v + 1]):
            edges.add((v, int(nbr_id[i]))); edges.add((int(nbr_id[i]), v))
    feature_table = {}
    for p in range(len(m["point_anchor"])):
        for i in range(m["vis_ptr"][p], m["vis_ptr"][p + 1]):
            feature_table.setdefault(int(m["vis_pose"][i]), []).append(p)
    active, extend = set(), {}
    for f in sorted(win):                                        # WindowTable is a std::map: ascending frame ids
        if win[f] != 1:
            continue
        for p in feature_table.get(f, []):
            if p in active:
                continue
            a = int(m["point_anchor"][p])
            if a in win:
                active.add(p)
            elif (f, a) in edges:
                active.add(p)
                extend[a] = 2
    out = dict(win)
    out.update(extend)
    return sorted(active), out


def select_constraints(nbr_ptr, nbr_id, nbr_T, nbr_Lambda, win):
    """The pair loop of SlamGraph::copyContraintsToG2o (slam_graph.cpp:938-981): every ordered pair (id1, id2) of window
    frames with an edge and at least one OUTER frame gives a constraint T_2_from_1; pairs in ascending (id1, id2) order.
    nbr_T[e] / nbr_Lambda[e] belong to the directed entry e = (v -> nbr_id[e]).  Returns (c_i, c_j, c_T, c_Lambda) with
    c_i / c_j as positions in the ascending window list."""
    order = sorted(win)
    pos = {v: i for i, v in enumerate(order)}
    ci, cj, cT, cL = [], [], [], []
    for a in order:
        ents = sorted((int(nbr_id[e]), e) for e in range(nbr_ptr[a], nbr_ptr[a + 1]))
        for b, e in ents:
            if b == a or b not in win:
                continue
            if win[a] == 2 or win[b] == 2:
                ci.append(pos[a]); cj.append(pos[b]); cT.append(nbr_T[e]); cL.append(nbr_Lambda[e])
    return (np.array(ci, np.int32), np.array(cj, np.int32), np.array(cT, np.float64).reshape(-1, 7),
            np.array(cL, np.float64).reshape(-1, 36))


def add_keyframe(m, oldkey, T_newkey_from_oldkey, new_anchor, new_xyz, new_anchor_center, new_anchor_level, new_center,
                 new_level, track_point, track_center, track_level):
    """SlamGraph::addKeyframe without the strength bookkeeping (slam_graph.cpp:144-186, 359-421): the new vertex gets
    T_newkey_from_oldkey * T_oldkey_from_world; every new point is anchored in an existing frame and is seen by that
    frame and by the new keyframe; tracked points gain an observation by the new keyframe.  Returns the grown tables
    (observations of a point by ascending vertex id, as synth_graph.make_map lists them)."""
    V, Np = len(m["poses"]), len(m["point_anchor"])
    poses = np.vstack([m["poses"], se3_mul(T_newkey_from_oldkey, m["poses"][oldkey])[None]])
    rows = []
    for p in range(Np):
        for i in range(m["vis_ptr"][p], m["vis_ptr"][p + 1]):
            rows.append((p, int(m["vis_pose"][i]), np.asarray(m["feat_center"][i], np.float64), int(m["feat_level"][i])))
    for t, p in enumerate(track_point):
        rows.append((int(p), V, np.asarray(track_center[t], np.float64), int(track_level[t])))
    for q in range(len(new_anchor)):
        rows.append((Np + q, int(new_anchor[q]), np.asarray(new_anchor_center[q], np.float64), int(new_anchor_level[q])))
        rows.append((Np + q, V, np.asarray(new_center[q], np.float64), int(new_level[q])))
    rows.sort(key=lambda r: (r[0], r[1]))
    vis_point = np.array([r[0] for r in rows])
    Np2 = Np + len(new_anchor)
    return dict(poses=poses, point_anchor=np.concatenate([m["point_anchor"], np.asarray(new_anchor, np.int32)]).astype(np.int32),
                xyz_anchor=np.vstack([m["xyz_anchor"], np.asarray(new_xyz, np.float64).reshape(-1, 3)]),
                vis_ptr=np.searchsorted(vis_point, np.arange(Np2 + 1)).astype(np.int32),
                vis_pose=np.array([r[1] for r in rows], np.int32),
                feat_center=np.array([r[2] for r in rows], np.float64).reshape(-1, 3),
                feat_level=np.array([r[3] for r in rows], np.int32))
