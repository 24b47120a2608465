"""Generates tests/golden/round2_golden.npz from the CPU oracles of the rows added in round 2: GpuTracker::residualImage
(dt_oracle.c: odt_residual_image), computeInitialDoubleWin + active points + constraint pairs and addKeyframe
(pyoracle.py).  Oracle outputs, not reference outputs (the reference has no golden vectors and cannot be built here):
the file guards the restatements against silent drift; the GPU parity tests compare against the same oracles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph


def inputs():
    from test_dt_oracle import _levels
    lv, _ = _levels(po, 3)
    T = po.se3_exp(np.array([0.03, 0.01, -0.05, 0.002, -0.02, 0.004]))
    pb = synth.make_window(24, 900, seed=17)
    m, _, _ = synth_graph.make_map(pb, extra_vertices=5, extra_points=12, seed=2)
    return lv, T, m


def compute():
    lv, T, m = inputs()
    img = po.dt_residual_image(lv[2], T, exact=False)
    grey = (img[..., 0] == img[..., 1]) & (img[..., 1] == img[..., 2])
    out = dict(res_shape=np.array(img.shape), res_grey=int(grey.sum()),
               res_red=int(((img[..., 0] == 1) & (img[..., 1] == 0)).sum()), res_green=int(((img[..., 0] == 0) & (img[..., 1] == 1)).sum()),
               res_sample=img[::7, ::5, 0].copy(), res_sum=float(img[..., 0].astype(np.float64).sum()))
    ptr, ids, Tn, Lm = synth_graph.make_pose_graph(m)
    win = po.compute_double_window(ptr, ids, root=9, inner_window_size=4, double_window_size=12)
    active, win2 = po.compute_active_points(m, ptr, ids, win)
    ci, cj, cT, cL = po.select_constraints(ptr, ids, Tn, Lm, win2)
    out.update(win_v=np.array(sorted(win2)), win_type=np.array([win2[v] for v in sorted(win2)]), active=np.array(sorted(active)),
               con_i=np.array(ci), con_j=np.array(cj), con_T=np.array(cT))
    V, Np = len(m["poses"]), len(m["point_anchor"])
    g = po.add_keyframe(m, oldkey=V - 1, T_newkey_from_oldkey=np.array([0, 0, 0, 1, 0.1, 0, -0.05]), new_anchor=[V - 1, V - 3],
                        new_xyz=[[0, 0, 3], [1, 0, 4]], new_anchor_center=[[1, 2, 3], [4, 5, 6]], new_anchor_level=[0, 1],
                        new_center=[[7, 8, 9], [1, 1, 1]], new_level=[1, 0], track_point=[Np - 1, 3], track_center=[[9, 9, 9], [2, 2, 2]],
                        track_level=[0, 2])
    out.update(kf_pose=g["poses"][V], kf_vis_ptr=np.asarray(g["vis_ptr"]), kf_vis_pose=np.asarray(g["vis_pose"]))
    return out


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "round2_golden.npz")
    np.savez_compressed(out, **compute())
    print("wrote", out, os.path.getsize(out), "bytes")
