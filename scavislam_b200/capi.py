"""ctypes binding of libsvsb200.so (the C ABI declared in include/svs_b200.h).

This is how the tests and bench.py reach the product: through the same C ABI a
maintainer of the reference would bind from C++.  There is no
CPU fallback here: if the library is missing or no CUDA device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvsb200.so")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_ubyte)

SVS_BA_MAX_ITERS = 64
SVS_BA_SKIP_SELF_ANCHOR_HESSIAN = 1
SVS_BA_NATURAL_ORDER = 2


class SvsCam(C.Structure):
    _fields_ = [("f", C.c_double), ("px", C.c_double), ("py", C.c_double), ("b", C.c_double)]


class SvsBaOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("flags", C.c_int), ("reserved", C.c_int * 6)]


class SvsBaStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials_total", C.c_int),
                ("chi2_init", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double),
                ("chi2_iter", C.c_double * SVS_BA_MAX_ITERS), ("lambda_iter", C.c_double * SVS_BA_MAX_ITERS),
                ("trials_iter", C.c_int * SVS_BA_MAX_ITERS),
                ("num_frames", C.c_int), ("num_points", C.c_int),
                ("num_point_edges", C.c_int), ("num_frame_edges", C.c_int),
                ("nnzb_S", C.c_int), ("nnzb_L", C.c_int), ("max_track", C.c_int),
                ("ms_total", C.c_float), ("ms_build", C.c_float), ("ms_solve", C.c_float),
                ("ms_update", C.c_float), ("ms_control", C.c_float), ("launches", C.c_int)]

    def as_dict(self):
        n = max(self.iterations, 0)
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("chi2_iter", "lambda_iter", "trials_iter")}
        d["chi2_iter"] = list(self.chi2_iter[:n])
        d["lambda_iter"] = list(self.lambda_iter[:n])
        d["trials_iter"] = list(self.trials_iter[:n])
        return d


class SvsFastCell(C.Structure):
    _fields_ = [("u0", C.c_int), ("u1", C.c_int), ("v0", C.c_int), ("v1", C.c_int), ("thr", C.c_int)]


class SvsFastGridParams(C.Structure):
    _fields_ = [("grid_w", C.c_int), ("grid_h", C.c_int), ("fast_min", C.c_int), ("fast_max", C.c_int),
                ("min_inner", C.c_int), ("min_outer", C.c_int), ("max_inner", C.c_int), ("max_outer", C.c_int)]


SVS_DT_MAX_LEVELS = 8
SVS_DT_EXACT_BILINEAR = 1


class SvsDtStats(C.Structure):
    _fields_ = [("chi2", C.c_double * SVS_DT_MAX_LEVELS), ("passes", C.c_int * SVS_DT_MAX_LEVELS),
                ("launches", C.c_int), ("ms_total", C.c_float)]


class SvsPoseParams(C.Structure):
    _fields_ = [("robust_kernel", C.c_int), ("kernel_param", C.c_double), ("num_iter", C.c_int),
                ("initial_mu", C.c_double), ("tau", C.c_double)]


class SvsPoseStats(C.Structure):
    _fields_ = [("initial_chi2", C.c_double), ("chi2", C.c_double), ("max_err", C.c_double), ("num_obs", C.c_int),
                ("iterations", C.c_int), ("trials", C.c_int), ("ms", C.c_float)]


class SvsMatchLevel(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("f", C.c_double), ("px", C.c_double), ("py", C.c_double)]


class SvsMatchPoint(C.Structure):
    _fields_ = [("keyframe", C.c_int), ("anchor_level", C.c_int), ("xyz_anchor", C.c_double * 3),
                ("anchor_obs_pyr", C.c_double * 2)]


class SvsMatchResult(C.Structure):
    _fields_ = [("predicted", C.c_int), ("textured", C.c_int), ("matched", C.c_int), ("n_candidates", C.c_int),
                ("index", C.c_int), ("min_dist", C.c_int), ("uv_pyr", C.c_int * 2), ("obs", C.c_double * 3),
                ("xyz_actkey", C.c_double * 3)]


MATCH_RESULT_DTYPE = np.dtype([("predicted", "i4"), ("textured", "i4"), ("matched", "i4"), ("n_candidates", "i4"),
                               ("index", "i4"), ("min_dist", "i4"), ("uv_pyr", "i4", 2), ("obs", "f8", 3),
                               ("xyz_actkey", "f8", 3)])
MATCH_POINT_DTYPE = np.dtype([("keyframe", "i4"), ("anchor_level", "i4"), ("xyz_anchor", "f8", 3),
                              ("anchor_obs_pyr", "f8", 2)])


EXPORTS = [
    "svs_ba_create", "svs_ba_destroy", "svs_last_error", "svs_ba_set_problem", "svs_ba_optimize",
    "svs_ba_get_poses", "svs_ba_get_points", "svs_ba_reset_state", "svs_optimiseInnerAndOuterWindow",
    "svs_ba_chi2", "svs_ba_reduced_system", "svs_ba_solve_reduced", "svs_device_info",
    "svs_ba_set_structure", "svs_ba_lm_begin", "svs_ba_trial_build", "svs_ba_system_buffers", "svs_ba_trial_solve",
    "svs_ba_trial_decide", "svs_ba_lm_stats",
    "svs_comm_unique_id", "svs_ba_comm_init", "svs_ba_set_problem_sharded", "svs_ba_get_points_all",
    "svs_fast_create", "svs_fast_destroy", "svs_fast_last_error", "svs_fast_grid_init", "svs_fast_set_image",
    "svs_fast_set_image_device", "svs_fast_detect", "svs_fast_detect_adaptively",
    "svs_dt_create", "svs_dt_destroy", "svs_dt_last_error", "svs_dt_set_intrinsics", "svs_dt_set_images",
    "svs_dt_set_disparity", "svs_dt_compute_point_cloud", "svs_dt_set_point_cloud", "svs_dt_get_point_cloud",
    "svs_dt_chi2", "svs_dt_jacobian_reduction", "svs_dt_track", "svs_dt_residual_image",
    "svs_matcher_create", "svs_matcher_destroy", "svs_matcher_last_error", "svs_matcher_set_keyframe",
    "svs_matcher_set_current", "svs_matcher_set_features", "svs_matcher_set_features_from_fast", "svs_match",
    "svs_prep_create", "svs_prep_destroy", "svs_prep_last_error", "svs_prep_process", "svs_prep_level",
    "svs_prep_get_u8", "svs_prep_get_f32", "svs_dt_set_images_device", "svs_dt_swap_prev_cur",
    "svs_matcher_set_pyramid_device",
    "svs_pose_create", "svs_pose_destroy", "svs_pose_last_error", "svs_calcFastMotionOnly",
    "svs_calcFastMotionOnly_matched",
    "svs_dtc_create", "svs_dtc_destroy", "svs_dtc_last_error", "svs_dtc_set_prev_u8", "svs_dtc_set_cur",
    "svs_dtc_set_disparity", "svs_computeDensePointCloudCpu", "svs_dtc_get_point_cloud", "svs_dtc_set_point_cloud",
    "svs_denseTrackingCpu",
    "svs_constraints_create", "svs_constraints_destroy", "svs_constraints_last_error", "svs_computeConstraint_batch",
    "svs_map_create", "svs_map_destroy", "svs_map_last_error", "svs_map_set", "svs_map_update_poses",
    "svs_map_update_points", "svs_map_get", "svs_map_absorb", "svs_map_set_graph", "svs_map_select_window",
    "svs_map_add_keyframe",
    "svs_ba_set_problem_from_map", "svs_map_last_edges",
]


def comm_unique_id():
    """128-byte NCCL rendezvous id (rank 0 creates it, the caller broadcasts it)."""
    buf = C.create_string_buffer(128)
    rc = lib().svs_comm_unique_id(buf)
    if rc != 0:
        raise SvsError(rc, "svs_comm_unique_id: NCCL not loadable")
    return buf.raw


def lib():
    """Load libsvsb200.so; raises if it was not built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.svs_ba_create.argtypes = [C.POINTER(SvsBaOpts), C.POINTER(vp)]
    L.svs_ba_destroy.argtypes = [vp]
    L.svs_ba_destroy.restype = None
    L.svs_last_error.argtypes = [vp]
    L.svs_last_error.restype = C.c_char_p
    prob = [C.c_int, c_dp, c_up, C.c_int, c_dp, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp,
            C.c_int, c_ip, c_ip, c_dp, c_dp, C.POINTER(SvsCam)]
    L.svs_ba_set_problem.argtypes = [vp] + prob
    L.svs_ba_optimize.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(SvsBaStats)]
    L.svs_ba_get_poses.argtypes = [vp, c_dp]
    L.svs_ba_get_points.argtypes = [vp, c_dp]
    L.svs_ba_reset_state.argtypes = [vp]
    L.svs_optimiseInnerAndOuterWindow.argtypes = [vp] + prob + [C.c_int, C.c_int, C.c_double, C.POINTER(SvsBaStats)]
    L.svs_ba_chi2.argtypes = [vp, C.c_int, C.c_double, c_dp]
    L.svs_ba_reduced_system.argtypes = [vp, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp]
    L.svs_ba_solve_reduced.argtypes = [vp, C.c_int, C.c_double, C.c_double, c_dp]
    L.svs_device_info.argtypes = [C.c_char_p, C.c_int]
    L.svs_ba_set_structure.argtypes = [vp, C.c_int, c_ip, c_ip]
    L.svs_ba_lm_begin.argtypes = [vp, C.c_double, C.c_int]
    L.svs_ba_trial_build.argtypes = [vp, C.c_int, C.c_double]
    pp = C.POINTER(C.c_void_p)
    pl = C.POINTER(C.c_longlong)
    L.svs_ba_system_buffers.argtypes = [vp, pp, pl, pp, pp, pl, pp]
    L.svs_ba_trial_solve.argtypes = [vp, C.c_int, C.c_double]
    L.svs_ba_trial_decide.argtypes = [vp, c_ip, c_ip, c_ip]
    L.svs_ba_lm_stats.argtypes = [vp, C.POINTER(SvsBaStats)]
    L.svs_comm_unique_id.argtypes = [C.c_char_p]
    L.svs_ba_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.svs_ba_set_problem_sharded.argtypes = [vp] + prob
    L.svs_ba_get_points_all.argtypes = [vp, c_dp]
    L.svs_fast_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.svs_fast_destroy.argtypes = [vp]
    L.svs_fast_destroy.restype = None
    L.svs_fast_last_error.argtypes = [vp]
    L.svs_fast_last_error.restype = C.c_char_p
    L.svs_fast_grid_init.argtypes = [C.c_int] * 9 + [C.POINTER(SvsFastGridParams), C.POINTER(SvsFastCell)]
    L.svs_fast_set_image.argtypes = [vp, c_up, C.c_int, C.c_int, C.c_int]
    L.svs_fast_set_image_device.argtypes = [vp, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.svs_fast_detect.argtypes = [vp, C.POINTER(SvsFastCell), C.c_int, c_ip, C.c_int, c_ip]
    L.svs_fast_detect_adaptively.argtypes = [vp, C.POINTER(SvsFastGridParams), C.POINTER(SvsFastCell), C.c_int,
                                             c_ip, C.c_int, c_ip]
    c_fp = C.POINTER(C.c_float)
    L.svs_dt_create.argtypes = [C.c_int] * 5 + [C.POINTER(vp)]
    L.svs_dt_destroy.argtypes = [vp]
    L.svs_dt_destroy.restype = None
    L.svs_dt_last_error.argtypes = [vp]
    L.svs_dt_last_error.restype = C.c_char_p
    L.svs_dt_set_intrinsics.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float]
    L.svs_dt_set_images.argtypes = [vp, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_int]
    L.svs_dt_set_disparity.argtypes = [vp, c_fp, C.c_int, C.c_int, C.c_int]
    L.svs_dt_compute_point_cloud.argtypes = [vp, c_dp, C.POINTER(SvsCam)]
    L.svs_dt_set_point_cloud.argtypes = [vp, C.c_int, c_fp]
    L.svs_dt_get_point_cloud.argtypes = [vp, C.c_int, c_fp]
    L.svs_dt_chi2.argtypes = [vp, C.c_int, c_dp, c_dp]
    L.svs_dt_jacobian_reduction.argtypes = [vp, C.c_int, c_dp, c_dp, c_dp, c_dp]
    L.svs_dt_track.argtypes = [vp, c_dp, C.POINTER(SvsDtStats)]
    L.svs_dt_residual_image.argtypes = [vp, C.c_int, c_dp, c_fp]
    L.svs_prep_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.svs_prep_destroy.argtypes = [vp]
    L.svs_prep_destroy.restype = None
    L.svs_prep_last_error.argtypes = [vp]
    L.svs_prep_last_error.restype = C.c_char_p
    L.svs_prep_process.argtypes = [vp, c_up, C.c_int]
    pvp = C.POINTER(C.c_void_p)
    L.svs_prep_level.argtypes = [vp, C.c_int, c_ip, c_ip, pvp, c_ip, pvp, pvp, pvp, c_ip]
    L.svs_prep_get_u8.argtypes = [vp, C.c_int, c_up]
    L.svs_prep_get_f32.argtypes = [vp, C.c_int, C.c_int, c_fp]
    L.svs_dt_set_images_device.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.svs_dt_swap_prev_cur.argtypes = [vp]
    L.svs_matcher_set_pyramid_device.argtypes = [vp, C.c_int, c_dp, C.POINTER(C.c_void_p), c_ip]
    L.svs_map_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.svs_map_destroy.argtypes = [vp]
    L.svs_map_destroy.restype = None
    L.svs_map_last_error.argtypes = [vp]
    L.svs_map_last_error.restype = C.c_char_p
    L.svs_map_set.argtypes = [vp, C.c_int, c_dp, C.c_int, c_ip, c_dp, c_ip, c_ip, c_dp, c_ip]
    L.svs_map_update_poses.argtypes = [vp, C.c_int, c_ip, c_dp]
    L.svs_map_update_points.argtypes = [vp, C.c_int, c_ip, c_dp]
    L.svs_map_get.argtypes = [vp, c_dp, c_dp]
    L.svs_map_absorb.argtypes = [vp, vp]
    L.svs_map_set_graph.argtypes = [vp, c_ip, c_ip, c_dp, c_dp]
    L.svs_map_select_window.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_up, C.c_int, c_ip, c_ip, C.c_int, c_ip,
                                        c_ip, c_ip, c_dp, c_dp]
    L.svs_map_add_keyframe.argtypes = [vp, C.c_int, c_dp, C.c_int, c_ip, c_dp, c_dp, c_ip, c_dp, c_ip, C.c_int, c_ip, c_dp, c_ip,
                                       c_ip, c_ip]
    L.svs_ba_set_problem_from_map.argtypes = [vp, vp, C.c_int, c_ip, c_up, C.c_int, c_ip, C.c_int, c_ip, c_ip, c_dp, c_dp,
                                              C.POINTER(SvsCam), c_ip]
    L.svs_map_last_edges.argtypes = [vp, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp]
    L.svs_constraints_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.svs_constraints_destroy.argtypes = [vp]
    L.svs_constraints_destroy.restype = None
    L.svs_constraints_last_error.argtypes = [vp]
    L.svs_constraints_last_error.restype = C.c_char_p
    L.svs_computeConstraint_batch.argtypes = [vp, C.c_int, c_dp, c_ip, c_ip, C.c_int, c_ip, c_dp, C.c_int, c_ip, c_ip,
                                              c_dp, c_dp, c_ip]
    L.svs_dtc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.svs_dtc_destroy.argtypes = [vp]
    L.svs_dtc_destroy.restype = None
    L.svs_dtc_last_error.argtypes = [vp]
    L.svs_dtc_last_error.restype = C.c_char_p
    L.svs_dtc_set_prev_u8.argtypes = [vp, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.svs_dtc_set_cur.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.svs_dtc_set_disparity.argtypes = [vp, c_fp, C.c_int]
    L.svs_computeDensePointCloudCpu.argtypes = [vp, c_dp, C.POINTER(SvsCam)]
    L.svs_dtc_get_point_cloud.argtypes = [vp, C.c_int, c_fp]
    L.svs_dtc_set_point_cloud.argtypes = [vp, C.c_int, c_fp]
    L.svs_denseTrackingCpu.argtypes = [vp, C.POINTER(SvsCam), c_dp, C.POINTER(SvsDtStats)]
    L.svs_pose_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.svs_pose_destroy.argtypes = [vp]
    L.svs_pose_destroy.restype = None
    L.svs_pose_last_error.argtypes = [vp]
    L.svs_pose_last_error.restype = C.c_char_p
    L.svs_calcFastMotionOnly.argtypes = [vp, C.c_int, c_ip, c_dp, C.c_int, c_dp, C.POINTER(SvsCam),
                                         C.POINTER(SvsPoseParams), c_dp, C.POINTER(SvsPoseStats)]
    L.svs_calcFastMotionOnly_matched.argtypes = [vp, vp, C.POINTER(SvsCam), C.POINTER(SvsPoseParams), c_dp,
                                                 C.POINTER(SvsPoseStats)]
    ucpp = C.POINTER(C.POINTER(C.c_ubyte))
    L.svs_matcher_create.argtypes = [C.c_int, C.c_int, C.POINTER(SvsMatchLevel), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.svs_matcher_destroy.argtypes = [vp]
    L.svs_matcher_destroy.restype = None
    L.svs_matcher_last_error.argtypes = [vp]
    L.svs_matcher_last_error.restype = C.c_char_p
    L.svs_matcher_set_keyframe.argtypes = [vp, C.c_int, c_dp, ucpp, c_ip]
    L.svs_matcher_set_current.argtypes = [vp, ucpp, c_ip, c_fp, C.c_int]
    L.svs_matcher_set_features.argtypes = [vp, C.c_int, c_ip, c_ip, C.c_int]
    L.svs_matcher_set_features_from_fast.argtypes = [vp, C.c_int, vp]
    L.svs_match.argtypes = [vp, c_dp, c_dp, C.POINTER(SvsMatchPoint), C.c_int, C.c_int, C.c_int, C.c_int,
                            C.POINTER(SvsMatchResult)]
    _LIB = L
    return L


def device_info() -> str:
    buf = C.create_string_buffer(256)
    rc = lib().svs_device_info(buf, 256)
    if rc != 0:
        raise RuntimeError(f"svs_device_info: {buf.value.decode()} (rc={rc})")
    return buf.value.decode()


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


class SvsError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"svs error {rc}: {msg}")
        self.rc = rc


class BundleAdjuster:
    """Thin host-side mirror of SlamGraph::optimize (reference slam_graph.cpp:319-355)."""

    def __init__(self, device: int = -1, flags: int = 0):
        self._h = C.c_void_p()
        o = SvsBaOpts(device, flags)
        rc = lib().svs_ba_create(C.byref(o), C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_ba_create failed (no CUDA device? there is no CPU fallback)")
        self._keep = None
        self.P = self.L = 0

    def close(self):
        if self._h:
            lib().svs_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_last_error(self._h).decode())

    @staticmethod
    def _arrays(pb):
        return dict(
            pose_qt=np.ascontiguousarray(pb.pose_qt, np.float64), fixed=np.ascontiguousarray(pb.fixed, np.uint8),
            psi=np.ascontiguousarray(pb.psi, np.float64),
            e_point=np.ascontiguousarray(pb.e_point, np.int32), e_pose=np.ascontiguousarray(pb.e_pose, np.int32),
            e_anchor=np.ascontiguousarray(pb.e_anchor, np.int32),
            e_obs=np.ascontiguousarray(pb.e_obs, np.float64), e_info=np.ascontiguousarray(pb.e_info, np.float64),
            c_i=np.ascontiguousarray(pb.c_i, np.int32), c_j=np.ascontiguousarray(pb.c_j, np.int32),
            c_T=np.ascontiguousarray(pb.c_T, np.float64), c_Lambda=np.ascontiguousarray(pb.c_Lambda, np.float64))

    @staticmethod
    def _prob_args(pb, k):
        cam = SvsCam(float(pb.cam[0]), float(pb.cam[1]), float(pb.cam[2]), float(pb.cam[3]))
        return [pb.P, _dp(k["pose_qt"]), k["fixed"].ctypes.data_as(c_up), pb.L, _dp(k["psi"]),
                pb.E, _ip(k["e_point"]), _ip(k["e_pose"]), _ip(k["e_anchor"]), _dp(k["e_obs"]), _dp(k["e_info"]),
                pb.C, _ip(k["c_i"]), _ip(k["c_j"]), _dp(k["c_T"]), _dp(k["c_Lambda"]), C.byref(cam)], cam

    def set_problem(self, pb):
        k = self._arrays(pb)
        args, cam = self._prob_args(pb, k)
        self._check(lib().svs_ba_set_problem(self._h, *args))
        self.P, self.L = pb.P, pb.L

    def optimize(self, num_iters, robust=True, huber_delta=1.0, lambda_init=50.0, max_trials=5):
        st = SvsBaStats()
        it = lib().svs_ba_optimize(self._h, int(num_iters), int(robust), float(huber_delta), float(lambda_init),
                                   int(max_trials), C.byref(st))
        if it <= -100:
            raise SvsError(it + 100, lib().svs_last_error(self._h).decode())
        return it, st.as_dict()

    def poses(self):
        out = np.zeros((self.P, 7))
        self._check(lib().svs_ba_get_poses(self._h, _dp(out)))
        return out

    def points(self):
        out = np.zeros((self.L, 3))
        self._check(lib().svs_ba_get_points(self._h, _dp(out)))
        return out

    def reset_state(self):
        self._check(lib().svs_ba_reset_state(self._h))

    def chi2(self, robust=True, huber_delta=1.0):
        v = C.c_double()
        self._check(lib().svs_ba_chi2(self._h, int(robust), float(huber_delta), C.byref(v)))
        return v.value

    def reduced_system(self, robust=True, huber_delta=1.0, lam=50.0):
        n = 6 * self.P
        S, bs = np.zeros((n, n)), np.zeros(n)
        chi = C.c_double()
        self._check(lib().svs_ba_reduced_system(self._h, int(robust), float(huber_delta), float(lam), _dp(S), _dp(bs),
                                                C.byref(chi)))
        return S, bs, chi.value

    def solve_reduced(self, robust=True, huber_delta=1.0, lam=50.0):
        x = np.zeros(6 * self.P)
        rc = lib().svs_ba_solve_reduced(self._h, int(robust), float(huber_delta), float(lam), _dp(x))
        if rc < 0:
            self._check(rc)
        return x, rc

    # ---- stepwise trial API (window split by landmarks across ranks)
    def set_structure(self, pairs):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        pi, pj = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        self._check(lib().svs_ba_set_structure(self._h, len(pairs), _ip(pi), _ip(pj)))

    def lm_begin(self, lambda_init=50.0, max_trials=5):
        self._check(lib().svs_ba_lm_begin(self._h, float(lambda_init), int(max_trials)))

    def trial_build(self, robust=True, huber_delta=1.0):
        self._check(lib().svs_ba_trial_build(self._h, int(robust), float(huber_delta)))

    def system_buffers(self):
        """(ptr_S, nS, ptr_bp, ptr_bc, nb, ptr_totals): raw device pointers for the caller's collective."""
        S, bp, bc, tot = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nS, nb = C.c_longlong(), C.c_longlong()
        self._check(lib().svs_ba_system_buffers(self._h, C.byref(S), C.byref(nS), C.byref(bp), C.byref(bc), C.byref(nb),
                                                C.byref(tot)))
        return S.value, nS.value, bp.value, bc.value, nb.value, tot.value

    def trial_solve(self, robust=True, huber_delta=1.0):
        self._check(lib().svs_ba_trial_solve(self._h, int(robust), float(huber_delta)))

    def trial_decide(self):
        a, s, i = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().svs_ba_trial_decide(self._h, C.byref(a), C.byref(s), C.byref(i)))
        return a.value, s.value, i.value

    # ---- one window sharded by landmarks across GPUs, driven inside the library (NCCL on the handle's stream)
    def comm_init(self, nranks, rank, unique_id):
        """unique_id: the 128 bytes rank 0 got from `comm_unique_id()`, broadcast by the caller."""
        self._check(lib().svs_ba_comm_init(self._h, int(nranks), int(rank), bytes(unique_id)))

    def set_problem_sharded(self, pb):
        """Every rank passes the WHOLE window; the library keeps landmarks l % nranks == rank."""
        k = self._arrays(pb)
        args, cam = self._prob_args(pb, k)
        self._check(lib().svs_ba_set_problem_sharded(self._h, *args))
        self.P, self.L = pb.P, pb.L

    def points_all(self):
        out = np.zeros((self.L, 3))
        self._check(lib().svs_ba_get_points_all(self._h, _dp(out)))
        return out

    def lm_stats(self):
        st = SvsBaStats()
        self._check(lib().svs_ba_lm_stats(self._h, C.byref(st)))
        return st.as_dict()

    def optimise_inner_and_outer_window(self, pb, num_iters, robust=True, huber_delta=1.0):
        """SlamGraph::optimize in one call from host buffers; returns (iters, poses, psi, stats)."""
        k = self._arrays(pb)
        k["pose_qt"] = k["pose_qt"].copy()
        k["psi"] = k["psi"].copy()
        args, cam = self._prob_args(pb, k)
        st = SvsBaStats()
        it = lib().svs_optimiseInnerAndOuterWindow(self._h, *args, int(num_iters), int(robust), float(huber_delta),
                                                   C.byref(st))
        if it <= -100:
            raise SvsError(it + 100, lib().svs_last_error(self._h).decode())
        self.P, self.L = pb.P, pb.L
        return it, k["pose_qt"], k["psi"], st.as_dict()


class FastGrid:
    """Host-side mirror of ScaViSLAM's FastGrid (reference fast_grid.h:30-64): same constructor
    arguments, detect / detectAdaptively on the GPU through the C ABI.  Keypoints come back as
    (xy[n,2] int32, cell_off[ncells+1]); the quadtree content of keypoint i in cell c is
    i - cell_off[c]."""

    def __init__(self, img_w, img_h, num_features_per_cell, boundary_per_cell, fast_thr, grid_w, grid_h,
                 fast_min=10, fast_max=40, device=-1, max_keypoints=200000):
        self._h = C.c_void_p()
        rc = lib().svs_fast_create(device, img_w, img_h, max_keypoints, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_fast_create failed (no CUDA device? there is no CPU fallback)")
        self.params = SvsFastGridParams()
        self.cells = (SvsFastCell * (grid_w * grid_h))()
        rc = lib().svs_fast_grid_init(img_w, img_h, num_features_per_cell, boundary_per_cell, fast_thr, grid_w,
                                      grid_h, fast_min, fast_max, C.byref(self.params), self.cells)
        if rc != 0:
            raise SvsError(rc, "svs_fast_grid_init")
        self.max_kp = max_keypoints
        self.ncells = grid_w * grid_h

    def close(self):
        if self._h:
            lib().svs_fast_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        raise SvsError(rc, lib().svs_fast_last_error(self._h).decode())

    def set_image(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        rc = lib().svs_fast_set_image(self._h, img.ctypes.data_as(c_up), img.strides[0], img.shape[1], img.shape[0])
        if rc != 0:
            self._err(rc)

    def set_image_device(self, ptr, pitch, w, h):
        rc = lib().svs_fast_set_image_device(self._h, C.c_void_p(ptr), pitch, w, h)
        if rc != 0:
            self._err(rc)

    def cell_list(self):
        return [(c.u0, c.u1, c.v0, c.v1, c.thr) for c in self.cells]

    def detect(self, cells=None):
        """FastGrid::detect(img, cell_grid2d, qt) with static per-cell thresholds."""
        if cells is None:
            arr, n = self.cells, self.ncells
        else:
            n = len(cells)
            arr = (SvsFastCell * n)(*[SvsFastCell(*c) for c in cells])
        out = np.zeros((self.max_kp, 2), np.int32)
        off = np.zeros(n + 1, np.int32)
        tot = lib().svs_fast_detect(self._h, arr, n, _ip(out), self.max_kp, _ip(off))
        if tot < 0:
            self._err(tot)
        return out[:min(tot, self.max_kp)].copy(), off

    def detect_adaptively(self, trials):
        """FastGrid::detectAdaptively(img, trials, qt); updates the per-cell thresholds in place."""
        out = np.zeros((self.max_kp, 2), np.int32)
        off = np.zeros(self.ncells + 1, np.int32)
        tot = lib().svs_fast_detect_adaptively(self._h, C.byref(self.params), self.cells, int(trials), _ip(out),
                                               self.max_kp, _ip(off))
        if tot < 0:
            self._err(tot)
        return out[:min(tot, self.max_kp)].copy(), off


class DenseTracker:
    """Host-side mirror of DenseTracker / GpuTracker (reference dense_tracking.h:40-96,
    gpu/dense_tracking.cuh:276-342) on top of the C ABI."""

    def __init__(self, w0, h0, nlevels=3, flags=0, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_dt_create(device, w0, h0, nlevels, flags, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_dt_create failed (no CUDA device? there is no CPU fallback)")
        self.w0, self.h0, self.nlevels = w0, h0, nlevels

    def close(self):
        if self._h:
            lib().svs_dt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_dt_last_error(self._h).decode())

    @staticmethod
    def _fp(a):
        return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))

    def set_intrinsics(self, level, f, px, py):
        self._ck(lib().svs_dt_set_intrinsics(self._h, level, f, px, py))

    def set_images(self, level, prev=None, cur=None, dx=None, dy=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (prev, cur, dx, dy)]
        w = self.w0 >> level
        self._ck(lib().svs_dt_set_images(self._h, level, *[self._fp(a) for a in arrs], w))

    def set_images_device(self, level, prev=None, cur=None, dx=None, dy=None, stride=0):
        self._ck(lib().svs_dt_set_images_device(self._h, level, prev, cur, dx, dy, stride))

    def swap_prev_cur(self):
        self._ck(lib().svs_dt_swap_prev_cur(self._h))

    def set_disparity(self, disp):
        d = np.ascontiguousarray(disp, np.float32)
        self._ck(lib().svs_dt_set_disparity(self._h, self._fp(d), d.shape[1], d.shape[1], d.shape[0]))

    def compute_point_cloud(self, T, cams):
        arr = (SvsCam * len(cams))(*[SvsCam(*map(float, c)) for c in cams])
        T = np.ascontiguousarray(T, np.float64)
        self._ck(lib().svs_dt_compute_point_cloud(self._h, _dp(T), arr))

    def set_point_cloud(self, level, cloud):
        c = np.ascontiguousarray(cloud, np.float32)
        self._ck(lib().svs_dt_set_point_cloud(self._h, level, self._fp(c)))

    def get_point_cloud(self, level):
        out = np.zeros((self.h0 >> level, self.w0 >> level, 4), np.float32)
        self._ck(lib().svs_dt_get_point_cloud(self._h, level, self._fp(out)))
        return out

    def chi2(self, level, T):
        T = np.ascontiguousarray(T, np.float64)
        v = C.c_double()
        self._ck(lib().svs_dt_chi2(self._h, level, _dp(T), C.byref(v)))
        return v.value

    def jacobian_reduction(self, level, T):
        T = np.ascontiguousarray(T, np.float64)
        H, b, v = np.zeros(21), np.zeros(6), C.c_double()
        self._ck(lib().svs_dt_jacobian_reduction(self._h, level, _dp(T), _dp(H), _dp(b), C.byref(v)))
        return H, b, v.value

    def residual_image(self, level, T):
        """GpuTracker::residualImage: (h, w, 4) float32."""
        T = np.ascontiguousarray(T, np.float64)
        out = np.zeros((self.h0 >> level, self.w0 >> level, 4), np.float32)
        self._ck(lib().svs_dt_residual_image(self._h, level, _dp(T), self._fp(out)))
        return out

    def track(self, T):
        T = np.ascontiguousarray(T, np.float64).copy()
        st = SvsDtStats()
        self._ck(lib().svs_dt_track(self._h, _dp(T), C.byref(st)))
        return T, dict(chi2=list(st.chi2[:self.nlevels]), passes=list(st.passes[:self.nlevels]),
                       launches=st.launches, ms_total=st.ms_total)


class GuidedMatcher:
    """Host-side mirror of GuidedMatcher<StereoCamera> (reference matcher.hpp:62-186)."""

    def __init__(self, levels, max_keyframes=8, max_points=8192, max_keypoints=65536, device=-1):
        """levels: list of (w, h, f, px, py) per pyramid level (cam_vec)."""
        self._h = C.c_void_p()
        arr = (SvsMatchLevel * len(levels))(*[SvsMatchLevel(int(w), int(h), float(f), float(px), float(py))
                                              for (w, h, f, px, py) in levels])
        rc = lib().svs_matcher_create(device, len(levels), arr, max_keyframes, max_points, max_keypoints,
                                      C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_matcher_create failed (no CUDA device? there is no CPU fallback)")
        self.nlevels = len(levels)

    def close(self):
        if self._h:
            lib().svs_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_matcher_last_error(self._h).decode())

    def _pyr_args(self, pyr):
        ims = [np.ascontiguousarray(p, np.uint8) for p in pyr]
        ptrs = (C.POINTER(C.c_ubyte) * len(ims))(*[im.ctypes.data_as(c_up) for im in ims])
        pitch = np.array([im.strides[0] for im in ims], np.int32)
        return ims, ptrs, pitch

    def set_keyframe(self, slot, T_me_from_w, pyr):
        ims, ptrs, pitch = self._pyr_args(pyr)
        T = np.ascontiguousarray(T_me_from_w, np.float64)
        self._ck(lib().svs_matcher_set_keyframe(self._h, slot, _dp(T), ptrs, _ip(pitch)))

    def set_current(self, pyr, disp=None):
        ims, ptrs, pitch = self._pyr_args(pyr)
        d = None if disp is None else np.ascontiguousarray(disp, np.float32)
        self._ck(lib().svs_matcher_set_current(self._h, ptrs, _ip(pitch),
                                               None if d is None else d.ctypes.data_as(C.POINTER(C.c_float)),
                                               0 if d is None else d.shape[1]))

    def set_current_disparity(self, disp):
        d = np.ascontiguousarray(disp, np.float32)
        self._ck(lib().svs_matcher_set_current(self._h, None, None, d.ctypes.data_as(C.POINTER(C.c_float)), d.shape[1]))

    def set_pyramid_device(self, which, ptrs, pitches, T_me_from_w=None):
        """which = -1: current frame, >= 0: keyframe slot (needs T_me_from_w); device pointers per level."""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        pitch = np.asarray(pitches, np.int32)
        T = None if T_me_from_w is None else np.ascontiguousarray(T_me_from_w, np.float64)
        self._ck(lib().svs_matcher_set_pyramid_device(self._h, which, None if T is None else _dp(T), arr, _ip(pitch)))

    def set_features(self, level, xy, content):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        content = np.ascontiguousarray(content, np.int32)
        self._ck(lib().svs_matcher_set_features(self._h, level, _ip(xy), _ip(content), len(xy)))

    def set_features_from_fast(self, level, fast_grid):
        """FAST corners of `fast_grid`'s last detect call, taken where they lie on the device."""
        self._ck(lib().svs_matcher_set_features_from_fast(self._h, level, fast_grid._h))

    def match(self, T_cur_from_actkey, T_actkey_from_w, points, search_radius, thr_mean, thr_std):
        """points: structured array with MATCH_POINT_DTYPE.  Returns a MATCH_RESULT_DTYPE array."""
        pts = np.ascontiguousarray(points, MATCH_POINT_DTYPE)
        out = np.zeros(len(pts), MATCH_RESULT_DTYPE)
        Ta = np.ascontiguousarray(T_cur_from_actkey, np.float64)
        Tb = np.ascontiguousarray(T_actkey_from_w, np.float64)
        rc = lib().svs_match(self._h, _dp(Ta), _dp(Tb), pts.ctypes.data_as(C.POINTER(SvsMatchPoint)), len(pts),
                             int(search_radius), int(thr_mean), int(thr_std),
                             out.ctypes.data_as(C.POINTER(SvsMatchResult)))
        if rc < 0:
            self._ck(rc)
        return out


class FramePreprocessor:
    """FrameGrabber::preprocessing (reference frame_grabber.cpp:287-336) on the device: uint8 and
    float pyramids and the x/y derivatives of one frame; outputs stay on the GPU."""

    def __init__(self, w, h, nlevels=3, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_prep_create(device, w, h, nlevels, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_prep_create failed (no CUDA device? there is no CPU fallback)")
        self.nlevels = nlevels

    def close(self):
        if self._h:
            lib().svs_prep_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_prep_last_error(self._h).decode())

    def process(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        self._ck(lib().svs_prep_process(self._h, img.ctypes.data_as(c_up), img.strides[0]))

    def level(self, l):
        """dict(w, h, u8, pitch_u8, f32, dx, dy, stride_f32) with raw device pointers."""
        w, h, p8, s32 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        u8, f32, dx, dy = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(lib().svs_prep_level(self._h, l, C.byref(w), C.byref(h), C.byref(u8), C.byref(p8), C.byref(f32),
                                      C.byref(dx), C.byref(dy), C.byref(s32)))
        return dict(w=w.value, h=h.value, u8=u8.value, pitch_u8=p8.value, f32=f32.value, dx=dx.value, dy=dy.value,
                    stride_f32=s32.value)

    def get_u8(self, l):
        lv = self.level(l)
        out = np.zeros((lv["h"], lv["w"]), np.uint8)
        self._ck(lib().svs_prep_get_u8(self._h, l, out.ctypes.data_as(c_up)))
        return out

    def get_f32(self, l, which=0):
        lv = self.level(l)
        out = np.zeros((lv["h"], lv["w"]), np.float32)
        self._ck(lib().svs_prep_get_f32(self._h, l, which, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out


class PoseOptimizer:
    """BA_SE3_XYZ_STEREO (reference pose_optimizer.h:495): motion-only LM, whole loop in one kernel."""

    def __init__(self, max_obs=16384, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_pose_create(device, max_obs, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_pose_create failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().svs_pose_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_pose_last_error(self._h).decode())

    @staticmethod
    def _params(robust_kernel, kernel_param, num_iter, initial_mu):
        return SvsPoseParams(int(robust_kernel), float(kernel_param), int(num_iter), float(initial_mu), 0.00001)

    @staticmethod
    def _stats(st):
        return dict(initial_chi2=st.initial_chi2, chi2=st.chi2, max_err=st.max_err, num_obs=st.num_obs,
                    iterations=st.iterations, trials=st.trials, ms=st.ms)

    def calc_fast_motion_only(self, obs_point_id, obs_uvu, point_xyz, cam, T_frame, robust_kernel=True,
                              kernel_param=1.0, num_iter=50, initial_mu=-1.0):
        """Returns (T_frame_new, stats); cam = (f, px, py, baseline)."""
        pid = np.ascontiguousarray(obs_point_id, np.int32)
        obs = np.ascontiguousarray(obs_uvu, np.float64).reshape(-1, 3)
        xyz = np.ascontiguousarray(point_xyz, np.float64).reshape(-1, 3)
        T = np.array(T_frame, np.float64).copy()
        c = SvsCam(*[float(x) for x in cam])
        p = self._params(robust_kernel, kernel_param, num_iter, initial_mu)
        st = SvsPoseStats()
        self._ck(lib().svs_calcFastMotionOnly(self._h, len(pid), _ip(pid), _dp(obs), len(xyz), _dp(xyz), C.byref(c),
                                              C.byref(p), _dp(T), C.byref(st)))
        return T, self._stats(st)

    def calc_fast_motion_only_matched(self, matcher, cam, T_frame, robust_kernel=True, kernel_param=1.0, num_iter=50,
                                      initial_mu=-1.0):
        T = np.array(T_frame, np.float64).copy()
        c = SvsCam(*[float(x) for x in cam])
        p = self._params(robust_kernel, kernel_param, num_iter, initial_mu)
        st = SvsPoseStats()
        self._ck(lib().svs_calcFastMotionOnly_matched(self._h, matcher._h, C.byref(c), C.byref(p), _dp(T), C.byref(st)))
        return T, self._stats(st)


class DenseTrackerCpuVariant:
    """DenseTracker as the reference builds it without SCAVISLAM_CUDA_SUPPORT (dense_tracking.cpp:222-423):
    denseTrackingCpu / computeDensePointCloudCpu semantics, executed on the GPU."""

    def __init__(self, w, h, nlevels=3, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_dtc_create(device, w, h, nlevels, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_dtc_create failed (no CUDA device, or a level size is not a multiple of 4)")
        self.w, self.h, self.nlevels = w, h, nlevels

    def close(self):
        if self._h:
            lib().svs_dtc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_dtc_last_error(self._h).decode())

    @staticmethod
    def _cams(cams):
        arr = (SvsCam * len(cams))()
        for i, c in enumerate(cams):
            arr[i] = SvsCam(*[float(x) for x in c[:4]])
        return arr

    def set_prev_u8(self, level, img):
        img = np.ascontiguousarray(img, np.uint8)
        self._ck(lib().svs_dtc_set_prev_u8(self._h, level, img.ctypes.data, img.strides[0], 0))

    def set_prev_u8_device(self, level, ptr, pitch):
        self._ck(lib().svs_dtc_set_prev_u8(self._h, level, ptr, pitch, 1))

    def set_cur(self, level, cur, dx, dy):
        ims = [np.ascontiguousarray(x, np.float32) for x in (cur, dx, dy)]
        self._ck(lib().svs_dtc_set_cur(self._h, level, ims[0].ctypes.data, ims[1].ctypes.data, ims[2].ctypes.data,
                                       ims[0].shape[1], 0))

    def set_cur_device(self, level, cur, dx, dy, stride):
        self._ck(lib().svs_dtc_set_cur(self._h, level, cur, dx, dy, stride, 1))

    def set_disparity(self, disp):
        d = np.ascontiguousarray(disp, np.float32)
        self._ck(lib().svs_dtc_set_disparity(self._h, d.ctypes.data_as(C.POINTER(C.c_float)), d.shape[1]))

    def compute_point_cloud(self, T, cams):
        T = np.ascontiguousarray(T, np.float64)
        self._ck(lib().svs_computeDensePointCloudCpu(self._h, _dp(T), self._cams(cams)))

    def point_cloud(self, level):
        out = np.zeros(((self.h >> level) // 4, (self.w >> level) // 4, 4), np.float32)
        self._ck(lib().svs_dtc_get_point_cloud(self._h, level, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def set_point_cloud(self, level, cloud):
        c = np.ascontiguousarray(cloud, np.float32)
        self._ck(lib().svs_dtc_set_point_cloud(self._h, level, c.ctypes.data_as(C.POINTER(C.c_float))))

    def track(self, T, cams):
        T = np.array(T, np.float64).copy()
        st = SvsDtStats()
        self._ck(lib().svs_denseTrackingCpu(self._h, self._cams(cams), _dp(T), C.byref(st)))
        return T, dict(chi2=list(st.chi2[:self.nlevels]), passes=list(st.passes[:self.nlevels]), ms_total=st.ms_total)


class ConstraintBuilder:
    """SlamGraph::computeConstraint (reference slam_graph.cpp:785-846) for a batch of pose pairs."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_constraints_create(device, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_constraints_create failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().svs_constraints_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, poses, feat_ptr, feat_point, point_anchor, xyz_anchor, v1, v2):
        """Returns (T_1_from_2[n,7], Lambda[n,6,6], visibility_strength[n])."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        fp, fpt = np.ascontiguousarray(feat_ptr, np.int32), np.ascontiguousarray(feat_point, np.int32)
        pa = np.ascontiguousarray(point_anchor, np.int32)
        xyz = np.ascontiguousarray(xyz_anchor, np.float64).reshape(-1, 3)
        v1, v2 = np.ascontiguousarray(v1, np.int32), np.ascontiguousarray(v2, np.int32)
        n = len(v1)
        T, Lam, ns = np.zeros((n, 7)), np.zeros((n, 36)), np.zeros(n, np.int32)
        rc = lib().svs_computeConstraint_batch(self._h, len(poses), _dp(poses), _ip(fp), _ip(fpt), len(pa), _ip(pa), _dp(xyz),
                                               n, _ip(v1), _ip(v2), _dp(T), _dp(Lam), _ip(ns))
        if rc != 0:
            raise SvsError(rc, lib().svs_constraints_last_error(self._h).decode())
        return T, Lam.reshape(n, 6, 6), ns


class DeviceMap:
    """The part of SlamGraph the optimiser reads, kept on the device (reference slam_graph.hpp:65-137), and
    copyDataToG2o (slam_graph.cpp:985-1032) as kernels feeding a BundleAdjuster."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        rc = lib().svs_map_create(device, C.byref(self._h))
        if rc != 0:
            raise SvsError(rc, "svs_map_create failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().svs_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SvsError(rc, lib().svs_map_last_error(self._h).decode())

    def set(self, poses, point_anchor, xyz_anchor, vis_ptr, vis_pose, feat_center, feat_level):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        pa = np.ascontiguousarray(point_anchor, np.int32)
        xyz = np.ascontiguousarray(xyz_anchor, np.float64).reshape(-1, 3)
        vp_, vs = np.ascontiguousarray(vis_ptr, np.int32), np.ascontiguousarray(vis_pose, np.int32)
        cen = np.ascontiguousarray(feat_center, np.float64).reshape(-1, 3)
        lvl = np.ascontiguousarray(feat_level, np.int32)
        self._ck(lib().svs_map_set(self._h, len(poses), _dp(poses), len(pa), _ip(pa), _dp(xyz), _ip(vp_), _ip(vs), _dp(cen),
                                   _ip(lvl)))
        self.V, self.Np = len(poses), len(pa)

    def update_poses(self, vertex, poses):
        v = np.ascontiguousarray(vertex, np.int32)
        T = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        self._ck(lib().svs_map_update_poses(self._h, len(v), _ip(v), _dp(T)))

    def update_points(self, point, xyz_anchor):
        v = np.ascontiguousarray(point, np.int32)
        x = np.ascontiguousarray(xyz_anchor, np.float64).reshape(-1, 3)
        self._ck(lib().svs_map_update_points(self._h, len(v), _ip(v), _dp(x)))

    def get(self):
        """(poses [V,7], xyz_anchor [Np,3]) as they lie on the device."""
        T, x = np.zeros((self.V, 7)), np.zeros((max(self.Np, 1), 3))
        self._ck(lib().svs_map_get(self._h, _dp(T), _dp(x)))
        return T, x[:self.Np]

    def absorb(self, ba):
        """SlamGraph::restoreDataFromG2o, device to device: the optimised window of `ba` goes back into the map."""
        self._ck(lib().svs_map_absorb(self._h, ba._h))

    def set_graph(self, nbr_ptr, nbr_id, nbr_T=None, nbr_Lambda=None):
        """The pose graph: per vertex its neighbours, strongest first; optionally the marginalised constraint of every
        directed entry (T_nbr_from_me [7], Lambda [36])."""
        ptr, ids = np.ascontiguousarray(nbr_ptr, np.int32), np.ascontiguousarray(nbr_id, np.int32)
        T = None if nbr_T is None else np.ascontiguousarray(nbr_T, np.float64).reshape(-1, 7)
        Lm = None if nbr_Lambda is None else np.ascontiguousarray(nbr_Lambda, np.float64).reshape(-1, 36)
        self._nn = len(ids)
        self._ck(lib().svs_map_set_graph(self._h, _ip(ptr), _ip(ids), None if T is None else _dp(T), None if Lm is None else _dp(Lm)))

    def select_window(self, root, inner_window_size, double_window_size):
        """computeInitialDoubleWin + computeActivePointsAndExtendOuterWindow + the pair selection of copyContraintsToG2o
        on the device.  Returns dict(window_vertex, inner, active_point, c_i, c_j, c_T, c_Lambda)."""
        capP, capL, capC = self.V, max(self.Np, 1), max(getattr(self, "_nn", 0), 1)
        win, inner, act = np.zeros(capP, np.int32), np.zeros(capP, np.uint8), np.zeros(capL, np.int32)
        ci, cj, cT, cL = np.zeros(capC, np.int32), np.zeros(capC, np.int32), np.zeros((capC, 7)), np.zeros((capC, 36))
        P, L, Cn = C.c_int(), C.c_int(), C.c_int()
        self._ck(lib().svs_map_select_window(self._h, int(root), int(inner_window_size), int(double_window_size), capP, C.byref(P),
                                             _ip(win), inner.ctypes.data_as(c_up), capL, C.byref(L), _ip(act), capC, C.byref(Cn),
                                             _ip(ci), _ip(cj), _dp(cT), _dp(cL)))
        P, L, Cn = P.value, L.value, Cn.value
        return dict(window_vertex=win[:P].copy(), inner=inner[:P].copy(), active_point=act[:L].copy(), c_i=ci[:Cn].copy(),
                    c_j=cj[:Cn].copy(), c_T=cT[:Cn].copy(), c_Lambda=cL[:Cn].copy())

    def add_keyframe(self, oldkey, T_newkey_from_oldkey, new_anchor=(), new_xyz=None, new_anchor_center=None,
                     new_anchor_level=(), new_center=None, new_level=(), track_point=(), track_center=None, track_level=()):
        """SlamGraph::addKeyframe on the device tables.  Returns (index of the new vertex, index of the first new point)."""
        T = np.ascontiguousarray(T_newkey_from_oldkey, np.float64).reshape(7)
        na = np.ascontiguousarray(new_anchor, np.int32)
        d3 = lambda a, n: np.zeros((n, 3)) if a is None else np.ascontiguousarray(a, np.float64).reshape(n, 3)
        nx, nac, nc = d3(new_xyz, len(na)), d3(new_anchor_center, len(na)), d3(new_center, len(na))
        nal, nl = np.ascontiguousarray(new_anchor_level, np.int32), np.ascontiguousarray(new_level, np.int32)
        tp = np.ascontiguousarray(track_point, np.int32)
        tc, tl = d3(track_center, len(tp)), np.ascontiguousarray(track_level, np.int32)
        v, q = C.c_int(), C.c_int()
        self._ck(lib().svs_map_add_keyframe(self._h, int(oldkey), _dp(T), len(na), _ip(na), _dp(nx), _dp(nac), _ip(nal), _dp(nc),
                                            _ip(nl), len(tp), _ip(tp), _dp(tc), _ip(tl), C.byref(v), C.byref(q)))
        self.V += 1
        self.Np += len(na)
        return v.value, q.value

    def set_problem(self, ba, window_vertex, active_point, cam, fixed=None, c_i=(), c_j=(), c_T=None, c_Lambda=None):
        """Assembles the window on the device and loads it into `ba` (a BundleAdjuster).  Returns E."""
        win = np.ascontiguousarray(window_vertex, np.int32)
        act = np.ascontiguousarray(active_point, np.int32)
        fx = None if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        ci, cj = np.ascontiguousarray(c_i, np.int32), np.ascontiguousarray(c_j, np.int32)
        cT = np.zeros((0, 7)) if c_T is None else np.ascontiguousarray(c_T, np.float64)
        cL = np.zeros((0, 36)) if c_Lambda is None else np.ascontiguousarray(c_Lambda, np.float64)
        cm = SvsCam(*[float(x) for x in cam])
        E = C.c_int()
        self._ck(lib().svs_ba_set_problem_from_map(ba._h, self._h, len(win), _ip(win),
                                                   None if fx is None else fx.ctypes.data_as(c_up), len(act), _ip(act),
                                                   len(ci), _ip(ci), _ip(cj), _dp(cT), _dp(cL), C.byref(cm), C.byref(E)))
        ba.P, ba.L = len(win), len(act)
        return E.value

    def last_edges(self, E):
        ep, es, ea = np.zeros(E, np.int32), np.zeros(E, np.int32), np.zeros(E, np.int32)
        obs, info = np.zeros((E, 3)), np.zeros((E, 3))
        self._ck(lib().svs_map_last_edges(self._h, E, _ip(ep), _ip(es), _ip(ea), _dp(obs), _dp(info)))
        return ep, es, ea, obs, info
