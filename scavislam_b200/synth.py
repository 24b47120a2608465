"""Seeded synthetic stereo double-window generator.

Produces flat arrays in exactly the layout of the C ABI (include/svs_b200.h,
``svs_ba_set_problem``), i.e. what ``SlamGraph::copyDataToG2o``
(reference scavislam/slam_graph.cpp:985-1032, slam_graph-impl.cpp:29-126)
hands to g2o: SE3 poses (all non-fixed), inverse-depth anchored points,
ternary (point, pose, anchor) stereo observations with Lambda =
diag(s, s, 0.333^2), and duplicated ordered pose-pose constraints with the
``computeConstraint`` weighting (slam_graph.cpp:785-846).

Input generation only: no optimisation arithmetic lives here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# stereo_slam.cpp:655-660 defaults
CAM_W, CAM_H = 640, 480
CAM_F, CAM_PX, CAM_PY, CAM_B = 570.342, 320.0, 240.0, 0.075

CONFIGS = {
    # name: (P, L, seed index)
    "C1": (10, 500, 0),
    "C2": (200, 20000, 1),
    "C4": (200, 20000, 3),
    "C5": (1000, 100000, 4),
}


@dataclass
class BAProblem:
    P: int
    L: int
    E: int
    C: int
    pose_qt: np.ndarray      # [P,7] qx qy qz qw tx ty tz (T_me_from_world)
    fixed: np.ndarray        # [P] uint8
    psi: np.ndarray          # [L,3]
    e_point: np.ndarray      # [E] int32
    e_pose: np.ndarray       # [E] int32
    e_anchor: np.ndarray     # [E] int32
    e_obs: np.ndarray        # [E,3]
    e_info: np.ndarray       # [E,3]
    c_i: np.ndarray          # [C] int32
    c_j: np.ndarray          # [C] int32
    c_T: np.ndarray          # [C,7]  T_2_from_1
    c_Lambda: np.ndarray     # [C,36]
    cam: np.ndarray          # f px py b
    truth_pose_qt: np.ndarray = field(default=None, repr=False)
    truth_psi: np.ndarray = field(default=None, repr=False)
    name: str = ""

    def copy(self) -> "BAProblem":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.copy() if isinstance(v, np.ndarray) else v
        return BAProblem(**kw)


# ---------------------------------------------------------------- SE3 (batched, generator only)

def _hat(v):
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def _exp_so3(om):
    th = np.linalg.norm(om, axis=-1)[..., None, None]
    th = np.maximum(th, 1e-12)
    K = _hat(om)
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def _exp_se3(d):
    ups, om = d[..., :3], d[..., 3:]
    th = np.maximum(np.linalg.norm(om, axis=-1)[..., None, None], 1e-12)
    K = _hat(om)
    R = _exp_so3(om)
    V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * (K @ K)
    return R, (V @ ups[..., None])[..., 0]


def _R_to_quat(R):
    """Batched rotation matrix -> (x y z w), w >= 0."""
    R = np.asarray(R)
    q = np.empty(R.shape[:-2] + (4,))
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    # robust branchless-ish: use the largest diagonal route per element
    for idx in np.ndindex(R.shape[:-2]):
        m = R[idx]
        t = tr[idx]
        if t > 0:
            s = math.sqrt(t + 1.0) * 2
            w = 0.25 * s
            x = (m[2, 1] - m[1, 2]) / s
            y = (m[0, 2] - m[2, 0]) / s
            z = (m[1, 0] - m[0, 1]) / s
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            w = (m[2, 1] - m[1, 2]) / s
            x = 0.25 * s
            y = (m[0, 1] + m[1, 0]) / s
            z = (m[0, 2] + m[2, 0]) / s
        elif m[1, 1] > m[2, 2]:
            s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            w = (m[0, 2] - m[2, 0]) / s
            x = (m[0, 1] + m[1, 0]) / s
            y = 0.25 * s
            z = (m[1, 2] + m[2, 1]) / s
        else:
            s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            w = (m[1, 0] - m[0, 1]) / s
            x = (m[0, 2] + m[2, 0]) / s
            y = (m[1, 2] + m[2, 1]) / s
            z = 0.25 * s
        v = np.array([x, y, z, w])
        if w < 0:
            v = -v
        q[idx] = v / np.linalg.norm(v)
    return q


def _to_qt(R, t):
    return np.concatenate([_R_to_quat(R), t], -1)


def quat_to_R(q):
    """Batched (x y z w) -> rotation matrix (Eigen toRotationMatrix)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


# ---------------------------------------------------------------- generator

def make_window(P: int, L: int, seed: int, T: int = 8, name: str = "",
                pose_noise=(0.02, 0.005), depth_noise=0.05, obs_sigma=0.5,
                outlier_frac=0.02, noise_seed=None) -> BAProblem:
    """`seed` fixes the window (trajectory, landmarks, visibility); measurement / pose / depth noise come
    from the same stream, or from `noise_seed` when given (same window, another noise realisation)."""
    rng = np.random.default_rng(seed)
    f, px, py, b = CAM_F, CAM_PX, CAM_PY, CAM_B

    # keyframes on a smooth planar curve, 0.3 m spacing, yaw following the tangent
    s = 0.3 * np.arange(P)
    yaw = 0.35 * np.sin(2 * np.pi * s / 45.0)
    pos = np.zeros((P, 3))
    step = 0.3 * np.stack([np.sin(yaw), np.zeros(P), np.cos(yaw)], -1)
    pos[1:] = np.cumsum(step[:-1], 0)
    c, sn = np.cos(yaw), np.sin(yaw)
    R_wc = np.zeros((P, 3, 3))
    R_wc[:, 0, 0] = c; R_wc[:, 0, 2] = sn; R_wc[:, 1, 1] = 1; R_wc[:, 2, 0] = -sn; R_wc[:, 2, 2] = c
    R_cw = np.transpose(R_wc, (0, 2, 1))
    t_cw = -(R_cw @ pos[..., None])[..., 0]

    # landmarks: pixel uniform in the anchor image, depth U[2,20]
    anchor = (np.arange(L, dtype=np.int64) * P // L).astype(np.int32)
    u = rng.uniform(0, CAM_W, L)
    v = rng.uniform(0, CAM_H, L)
    z = rng.uniform(2.0, 20.0, L)
    xa = np.stack([(u - px) / f * z, (v - py) / f * z, z], -1)
    psi_true = np.stack([xa[:, 0] / z, xa[:, 1] / z, 1.0 / z], -1)
    xw = (R_wc[anchor] @ (xa - 0)[..., None])[..., 0] + pos[anchor]

    # observations in frames a..a+T-1 where the point is inside the image with positive disparity
    ep, eq, eo = [], [], []
    for d in range(T):
        j = anchor + d
        ok = j < P
        jj = np.minimum(j, P - 1)
        y = (R_cw[jj] @ xw[..., None])[..., 0] + t_cw[jj]
        zz = np.where(y[:, 2] > 1e-6, y[:, 2], 1.0)
        uu = f * y[:, 0] / zz + px
        vv = f * y[:, 1] / zz + py
        ur = f * (y[:, 0] - b) / zz + px
        ok &= (y[:, 2] > 0.3) & (uu >= 0) & (uu < CAM_W) & (vv >= 0) & (vv < CAM_H) & (uu - ur > 0)
        idx = np.nonzero(ok)[0]
        ep.append(idx.astype(np.int32))
        eq.append(jj[idx].astype(np.int32))
        eo.append(np.stack([uu[idx], vv[idx], ur[idx]], -1))
    e_point = np.concatenate(ep)
    e_pose = np.concatenate(eq)
    e_obs = np.concatenate(eo)
    # order edges by (point, pose) like a walk over point_table_/vis_set would group them
    order = np.lexsort((e_pose, e_point))
    e_point, e_pose, e_obs = e_point[order], e_pose[order], e_obs[order]
    E = e_point.shape[0]
    e_anchor = anchor[e_point]
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)
    e_obs = e_obs + rng.normal(0, obs_sigma, (E, 3))
    outl = rng.uniform(size=E) < outlier_frac
    e_obs[outl] += rng.uniform(-20, 20, (int(outl.sum()), 3))
    level = (rng.uniform(size=E) < 0.25).astype(np.int32)
    sfac = np.where(level == 1, 0.25, 1.0)
    e_info = np.stack([sfac, sfac, np.full(E, 0.333 ** 2)], -1)

    # truth + perturbed initial state
    truth_qt = _to_qt(R_cw, t_cw)
    d = np.concatenate([rng.normal(0, pose_noise[0], (P, 3)), rng.normal(0, pose_noise[1], (P, 3))], -1)
    dR, dt = _exp_se3(d)
    R0 = dR @ R_cw
    t0 = (dR @ t_cw[..., None])[..., 0] + dt
    pose_qt = _to_qt(R0, t0)
    psi0 = psi_true.copy()
    psi0[:, 2] *= 1 + rng.normal(0, depth_noise, L)

    # pose-pose constraints: i, i+1..i+3 (both orders) when >= 1 is OUTER
    n_inner = int(math.ceil(0.15 * P))
    obs_mat = None
    # per-(pose) sorted landmark lists for shared counts
    o2 = np.argsort(e_pose, kind="stable")
    cuts = np.searchsorted(e_pose[o2], np.arange(P + 1))
    by_pose = [np.sort(e_point[o2[cuts[k]:cuts[k + 1]]].astype(np.int64)) for k in range(P)]
    ci, cj, cT, cL = [], [], [], []
    for i in range(P):
        for dd in range(1, 4):
            j = i + dd
            if j >= P:
                continue
            if i < n_inner and j < n_inner:
                continue
            shared = np.intersect1d(by_pose[i], by_pose[j], assume_unique=True)
            n = shared.shape[0]
            if n == 0:
                continue
            for (a, bb) in ((i, j), (j, i)):
                # measurement T_b_from_a with the reference's weighting computed in frame b ("v1")
                nz = rng.normal(0, 1, 6) * np.array([0.005] * 3 + [0.001] * 3)
                nR, nt = _exp_se3(nz[None])
                Rba = nR[0] @ (R_cw[bb] @ R_wc[a])
                tba = nR[0] @ (t_cw[bb] - R_cw[bb] @ R_wc[a] @ t_cw[a]) + nt[0]
                xv = (R_cw[bb] @ xw[shared].T).T + t_cw[bb]
                depth = np.sort(np.linalg.norm(xv, axis=1))
                med = depth[(n - 1) // 2] if n % 2 else 0.5 * (depth[n // 2 - 1] + depth[n // 2])
                norm_dist = np.linalg.norm(tba) / med
                lam = np.zeros((6, 6))
                lam[:3, :3] = np.eye(3) * n * (350 * norm_dist) ** 2
                lam[3:, 3:] = np.eye(3) * n * 100.0 ** 2
                ci.append(a); cj.append(bb)
                cT.append(_to_qt(Rba[None], tba[None])[0])
                cL.append(lam.reshape(36))
    C = len(ci)
    _ = obs_mat
    return BAProblem(
        P=P, L=L, E=E, C=C,
        pose_qt=np.ascontiguousarray(pose_qt), fixed=np.zeros(P, np.uint8),
        psi=np.ascontiguousarray(psi0),
        e_point=np.ascontiguousarray(e_point, np.int32), e_pose=np.ascontiguousarray(e_pose, np.int32),
        e_anchor=np.ascontiguousarray(e_anchor, np.int32),
        e_obs=np.ascontiguousarray(e_obs), e_info=np.ascontiguousarray(e_info),
        c_i=np.asarray(ci, np.int32), c_j=np.asarray(cj, np.int32),
        c_T=np.asarray(cT, np.float64).reshape(C, 7), c_Lambda=np.asarray(cL, np.float64).reshape(C, 36),
        cam=np.array([f, px, py, b]), truth_pose_qt=truth_qt, truth_psi=psi_true, name=name)


def make_config(name: str) -> BAProblem:
    P, L, idx = CONFIGS[name]
    return make_window(P, L, 1234 + idx, name=name)


# ---------------------------------------------------------------- structure variants of a window (bench / parity extras)

def with_dropouts(pb: BAProblem, frac: float = 0.2, seed: int = 0) -> BAProblem:
    """The same window with `frac` of the non-anchor observations missing at random (matching failures): tracks
    get holes, so neighbouring landmarks stop sharing one slot list."""
    rng = np.random.default_rng(seed)
    keep = (pb.e_pose == pb.e_anchor) | (rng.uniform(size=pb.E) >= frac)
    out = pb.copy()
    for k in ("e_point", "e_pose", "e_anchor", "e_obs", "e_info"):
        setattr(out, k, np.ascontiguousarray(getattr(pb, k)[keep]))
    out.E = int(keep.sum())
    out.name = f"{pb.name}+dropouts{int(round(100 * frac))}"
    return out


def with_loop_closures(pb: BAProblem, n: int = 10, seed: int = 0, min_gap=None) -> BAProblem:
    """The same window plus `n` loop-closure constraints between far-apart keyframes (both orders, as
    copyContraintsToG2o adds them: slam_graph.cpp:941-980), measurements taken from the truth with a little noise.
    They break the band of the reduced camera system."""
    rng = np.random.default_rng(seed)
    P = pb.P
    gap = min_gap if min_gap is not None else P // 3
    R = quat_to_R(pb.truth_pose_qt[:, :4])
    t = pb.truth_pose_qt[:, 4:]
    ci, cj, cT, cL = list(pb.c_i), list(pb.c_j), list(pb.c_T), list(pb.c_Lambda)
    seen = set()
    while len(seen) < n:
        i = int(rng.integers(0, P - gap))
        j = int(rng.integers(i + gap, P))
        if (i, j) in seen:
            continue
        seen.add((i, j))
        for (a, b) in ((i, j), (j, i)):
            nR, nt = _exp_se3((rng.normal(0, 1, 6) * np.array([0.005] * 3 + [0.001] * 3))[None])
            Rba = nR[0] @ (R[b] @ R[a].T)
            tba = nR[0] @ (t[b] - R[b] @ R[a].T @ t[a]) + nt[0]
            lam = np.zeros((6, 6))
            lam[:3, :3] = np.eye(3) * 4e4
            lam[3:, 3:] = np.eye(3) * 1e5
            ci.append(a); cj.append(b)
            cT.append(_to_qt(Rba[None], tba[None])[0])
            cL.append(lam.reshape(36))
    out = pb.copy()
    out.C = len(ci)
    out.c_i = np.asarray(ci, np.int32); out.c_j = np.asarray(cj, np.int32)
    out.c_T = np.asarray(cT, np.float64).reshape(out.C, 7)
    out.c_Lambda = np.asarray(cL, np.float64).reshape(out.C, 36)
    out.name = f"{pb.name}+loops{n}"
    return out
