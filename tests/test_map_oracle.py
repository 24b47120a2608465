"""CPU test of the window-assembly restatement (oracle.pyoracle.copy_data_to_g2o): assembling the window out of a
larger map gives back the window the map was made from (same edges, up to the reference's edge order)."""
import numpy as np

from oracle import pyoracle as po
from scavislam_b200 import synth, synth_graph


def test_assembly_returns_the_window_the_map_was_built_from():
    pb = synth.make_window(8, 200, seed=5)
    m, win, act = synth_graph.make_map(pb)
    g = po.copy_data_to_g2o(m, win, act)
    assert len(g["e_point"]) == pb.E                     # observations from outside frames / inactive points dropped
    np.testing.assert_array_equal(g["pose_qt"], pb.pose_qt)
    np.testing.assert_allclose(g["psi"], pb.psi, rtol=1e-14)
    key = lambda ep, es: np.lexsort((es, ep))
    o1, o2 = key(g["e_point"], g["e_pose"]), key(pb.e_point, pb.e_pose)
    np.testing.assert_array_equal(g["e_pose"][o1], pb.e_pose[o2])
    np.testing.assert_array_equal(g["e_anchor"][o1], pb.e_anchor[o2])
    np.testing.assert_array_equal(g["e_obs"][o1], pb.e_obs[o2])
    np.testing.assert_allclose(g["e_info"][o1], pb.e_info[o2], rtol=1e-12)
    # inside a point the edges follow ascending vertex ids (vis_set order), not ascending BA pose index
    first = g["e_pose"][g["e_point"] == 0]
    assert list(win[first]) == sorted(win[first])


def _chain_map(V=12, per_frame=6, seed=0):
    """Frames on a line; every point is anchored in one frame and seen by it and the next two."""
    rng = np.random.default_rng(seed)
    anchors, rows = [], []
    for a in range(V):
        for k in range(per_frame):
            p = len(anchors)
            anchors.append(a)
            for v in range(a, min(a + 3, V)):
                rows.append((p, v, rng.uniform(0, 600, 3), int(rng.integers(0, 3))))
    rows.sort(key=lambda r: (r[0], r[1]))
    Np = len(anchors)
    vis_point = np.array([r[0] for r in rows])
    poses = np.zeros((V, 7)); poses[:, 3] = 1; poses[:, 4] = np.arange(V)
    return dict(poses=poses, point_anchor=np.array(anchors, np.int32),
                xyz_anchor=np.stack([rng.uniform(-1, 1, Np), rng.uniform(-1, 1, Np), rng.uniform(2, 9, Np)], 1),
                vis_ptr=np.searchsorted(vis_point, np.arange(Np + 1)).astype(np.int32),
                vis_pose=np.array([r[1] for r in rows], np.int32), feat_center=np.array([r[2] for r in rows]),
                feat_level=np.array([r[3] for r in rows], np.int32))


def test_double_window_is_breadth_first_and_capped():
    m = _chain_map()
    ptr, ids, _, _ = synth_graph.make_pose_graph(m, with_constraints=False)
    win = po.compute_double_window(ptr, ids, root=5, inner_window_size=3, double_window_size=7)
    assert len(win) == 7 and win[5] == 1 and sum(1 for t in win.values() if t == 1) == 3
    assert set(win) == {2, 3, 4, 5, 6, 7, 8}                      # breadth-first on a line: the root's neighbourhood
    # the whole component when the cap is larger than it
    assert len(po.compute_double_window(ptr, ids, 0, 2, 100)) == len(m["poses"])


def test_active_points_need_an_inner_observer_and_extend_the_outer_window():
    m = _chain_map()
    ptr, ids, _, _ = synth_graph.make_pose_graph(m, with_constraints=False)
    win = po.compute_double_window(ptr, ids, root=6, inner_window_size=2, double_window_size=3)
    active, ext = po.compute_active_points(m, ptr, ids, win)
    inner = {v for v, t in win.items() if t == 1}
    for p in active:
        seen_by = set(m["vis_pose"][m["vis_ptr"][p]:m["vis_ptr"][p + 1]].tolist())
        assert seen_by & inner and int(m["point_anchor"][p]) in ext
    added = set(ext) - set(win)
    assert added and all(ext[v] == 2 for v in added)               # anchors outside the window joined as OUTER frames
    # a point seen only by outer frames is not active
    outer_only = [p for p in range(len(m["point_anchor"]))
                  if not set(m["vis_pose"][m["vis_ptr"][p]:m["vis_ptr"][p + 1]].tolist()) & inner]
    assert not set(outer_only) & set(active)


def test_constraint_pairs_touch_an_outer_frame_and_come_in_both_directions():
    m = _chain_map()
    ptr, ids, T, Lm = synth_graph.make_pose_graph(m)
    win = po.compute_double_window(ptr, ids, root=6, inner_window_size=3, double_window_size=6)
    _, win = po.compute_active_points(m, ptr, ids, win)
    ci, cj, cT, cL = po.select_constraints(ptr, ids, T, Lm, win)
    order = sorted(win)
    pairs = {(order[i], order[j]) for i, j in zip(ci, cj)}
    assert pairs and all((b, a) in pairs for a, b in pairs)
    assert all(win[a] == 2 or win[b] == 2 for a, b in pairs)
    assert list(zip(ci, cj)) == sorted(zip(ci, cj))


def test_add_keyframe_appends_a_vertex_points_and_observations():
    m = _chain_map(V=6)
    V, Np = len(m["poses"]), len(m["point_anchor"])
    T = np.array([0, 0, 0, 1, 0.5, 0, 0.0])
    g = po.add_keyframe(m, oldkey=V - 1, T_newkey_from_oldkey=T, new_anchor=[V - 1, V - 2], new_xyz=[[0, 0, 3], [1, 0, 4]],
                        new_anchor_center=[[1, 2, 3], [4, 5, 6]], new_anchor_level=[0, 1], new_center=[[7, 8, 9], [1, 1, 1]],
                        new_level=[1, 0], track_point=[Np - 1, 3], track_center=[[9, 9, 9], [2, 2, 2]], track_level=[0, 2])
    assert len(g["poses"]) == V + 1 and len(g["point_anchor"]) == Np + 2
    np.testing.assert_allclose(g["poses"][V][4:], m["poses"][V - 1][4:] + [0.5, 0, 0])
    assert g["vis_ptr"][-1] == m["vis_ptr"][-1] + 2 + 4
    seg = lambda p: g["vis_pose"][g["vis_ptr"][p]:g["vis_ptr"][p + 1]].tolist()
    assert seg(Np) == [V - 1, V] and seg(Np + 1) == [V - 2, V] and seg(3)[-1] == V and seg(Np - 1)[-1] == V
    assert all(seg(p) == sorted(seg(p)) for p in range(Np + 2))
