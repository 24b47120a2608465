"""Synthetic TrackData for the motion-only pose refinement (obs_list / point_list of
PoseOptimizer::calcFastMotionOnly, reference pose_optimizer.h:135-140)."""
import numpy as np

CAM = (480.0, 320.0, 240.0, 0.1)   # f, px, py, baseline


def _quat_from_rotvec(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([0, 0, 0, 1.0])
    return np.concatenate([np.sin(th / 2) * w / th, [np.cos(th / 2)]])


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_track(n, seed=0, pixel_noise=0.3, outlier_frac=0.0, motion=(0.05, 0.02), start_noise=(0.03, 0.01),
               shared_points=False, cam=CAM):
    """Returns dict(pid, obs, xyz, T_true, T_init): n stereo observations (u, v, u_right) of points
    seen from T_true, and a perturbed start T_init.  shared_points: several obs per point (the reference's
    obs_list indexes point_list through point_id)."""
    rng = np.random.default_rng(seed)
    f, px, py, b = cam
    npts = n if not shared_points else max(1, n // 2)
    z = rng.uniform(1.5, 12.0, npts)
    u = rng.uniform(20, 620, npts)
    v = rng.uniform(20, 460, npts)
    xyz = np.stack([(u - px) / f * z, (v - py) / f * z, z], 1)
    T_true = np.concatenate([_quat_from_rotvec(rng.normal(0, motion[1], 3)), rng.normal(0, motion[0], 3)])
    pid = np.arange(n) % npts if shared_points else np.arange(n)
    pid = pid.astype(np.int32)
    P = xyz[pid] @ _rot(T_true[:4]).T + T_true[4:]
    obs = np.stack([f * P[:, 0] / P[:, 2] + px, f * P[:, 1] / P[:, 2] + py, f * (P[:, 0] - b) / P[:, 2] + px], 1)
    obs += rng.normal(0, pixel_noise, obs.shape)
    nout = int(outlier_frac * n)
    if nout:
        idx = rng.choice(n, nout, replace=False)
        obs[idx] += rng.normal(0, 25.0, (nout, 3))
    dq = _quat_from_rotvec(rng.normal(0, start_noise[1], 3))
    qx, qy, qz, qw = T_true[:4]
    ax, ay, az, aw = dq
    q = np.array([aw * qx + ax * qw + ay * qz - az * qy, aw * qy + ay * qw + az * qx - ax * qz,
                  aw * qz + az * qw + ax * qy - ay * qx, aw * qw - ax * qx - ay * qy - az * qz])
    T_init = np.concatenate([q / np.linalg.norm(q), T_true[4:] + rng.normal(0, start_noise[0], 3)])
    return dict(pid=pid, obs=obs, xyz=xyz, T_true=T_true, T_init=T_init, cam=np.array(cam))
