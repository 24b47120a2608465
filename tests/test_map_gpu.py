"""GPU tests of the device-resident map and window assembly (svs_map_*,
svs_ba_set_problem_from_map) against the plain-Python restatement of copyDataToG2o, through the C ABI.
The assembled edge list is bit-exact; optimising the assembled window gives the result of loading the same window
through svs_ba_set_problem (to the last bits: FP64 atomics) and the oracle's result to 1e-6."""
import dataclasses

import numpy as np
import pytest

from scavislam_b200 import synth, synth_graph

pytestmark = pytest.mark.gpu


def _assembled_problem(pb, g):
    return dataclasses.replace(pb, E=len(g["e_point"]), pose_qt=g["pose_qt"], psi=g["psi"], e_point=g["e_point"],
                               e_pose=g["e_pose"], e_anchor=g["e_anchor"], e_obs=g["e_obs"], e_info=g["e_info"])


@pytest.mark.parametrize("P,L,seed", [(8, 200, 5), (30, 3000, 6)])
def test_assembled_window_equals_the_restatement_and_optimises_identically(svs, oracle, P, L, seed):
    pb = synth.make_window(P, L, seed=seed)
    m, win, act = synth_graph.make_map(pb, seed=seed)
    g = oracle.copy_data_to_g2o(m, win, act)
    dm, ba, ba2 = svs.DeviceMap(), svs.BundleAdjuster(), svs.BundleAdjuster()
    dm.set(m["poses"], m["point_anchor"], m["xyz_anchor"], m["vis_ptr"], m["vis_pose"], m["feat_center"], m["feat_level"])
    E = dm.set_problem(ba, win, act, pb.cam, c_i=pb.c_i, c_j=pb.c_j, c_T=pb.c_T, c_Lambda=pb.c_Lambda)
    assert E == len(g["e_point"]) == pb.E
    ep, es, ea, obs, info = dm.last_edges(E)
    np.testing.assert_array_equal(ep, g["e_point"]); np.testing.assert_array_equal(es, g["e_pose"])
    np.testing.assert_array_equal(ea, g["e_anchor"])
    np.testing.assert_array_equal(obs, g["e_obs"]); np.testing.assert_array_equal(info, g["e_info"])
    np.testing.assert_array_equal(ba.poses(), g["pose_qt"])
    np.testing.assert_array_equal(ba.points(), g["psi"])
    pa = _assembled_problem(pb, g)
    it, st = ba.optimize(3)
    ba2.set_problem(pa)
    it2, st2 = ba2.optimize(3)
    assert it == it2 == 3
    # same edge order -> same internal order; the Schur scatter uses FP64 atomics, so the last bits may differ
    np.testing.assert_allclose(st["chi2_iter"], st2["chi2_iter"], rtol=1e-12)
    np.testing.assert_allclose(ba.poses(), ba2.poses(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ba.points(), ba2.points(), rtol=1e-10, atol=1e-12)
    p_o, s_o, _ = oracle.optimize(pa, 3)
    assert np.abs(ba.poses() - p_o).max() <= 1e-6 * np.abs(p_o).max()
    assert np.abs(ba.points() - s_o).max() <= 1e-6 * np.abs(s_o).max()
    for h in (dm, ba, ba2):
        h.close()


def test_pose_updates_and_invalid_windows(svs, oracle):
    pb = synth.make_window(8, 200, seed=5)
    m, win, act = synth_graph.make_map(pb)
    dm, ba = svs.DeviceMap(), svs.BundleAdjuster()
    dm.set(m["poses"], m["point_anchor"], m["xyz_anchor"], m["vis_ptr"], m["vis_pose"], m["feat_center"], m["feat_level"])
    newT = oracle.se3_exp(np.array([0.1, 0.2, -0.1, 0.01, 0.0, 0.02]))
    dm.update_poses([win[3]], [newT])                             # restoreDataFromG2o's direction
    dm.set_problem(ba, win, act, pb.cam)
    np.testing.assert_array_equal(ba.poses()[3], newT)
    with pytest.raises(svs.SvsError):                             # a window without the anchor frame of an active point
        dm.set_problem(ba, win[1:], act, pb.cam)
    with pytest.raises(svs.SvsError):                             # the same vertex twice
        dm.set_problem(ba, np.concatenate([win, win[:1]]), act, pb.cam)
    dm.close(); ba.close()


def test_optimised_window_is_absorbed_on_the_device(svs, oracle):
    """restoreDataFromG2o (slam_graph.cpp:1037-1058) device to device: after optimize() the window's vertex poses and
    xyz_anchor = invert_depth(psi) of its points are in the map, everything outside the window is untouched; the
    next assembly starts from the optimised state.  svs_map_update_points is the host-side way in."""
    pb = synth.make_window(30, 3000, seed=6)
    m, win, act = synth_graph.make_map(pb, seed=6)
    dm, ba = svs.DeviceMap(), svs.BundleAdjuster()
    dm.set(m["poses"], m["point_anchor"], m["xyz_anchor"], m["vis_ptr"], m["vis_pose"], m["feat_center"], m["feat_level"])
    T0, x0 = dm.get()
    np.testing.assert_array_equal(T0, m["poses"]); np.testing.assert_array_equal(x0, m["xyz_anchor"])
    dm.set_problem(ba, win, act, pb.cam, c_i=pb.c_i, c_j=pb.c_j, c_T=pb.c_T, c_Lambda=pb.c_Lambda)
    ba.optimize(3)
    poses, psi = ba.poses(), ba.points()
    dm.absorb(ba)
    T1, x1 = dm.get()
    np.testing.assert_array_equal(T1[win], poses)
    xyz = np.stack([psi[:, 0] / psi[:, 2], psi[:, 1] / psi[:, 2], 1.0 / psi[:, 2]], -1)
    np.testing.assert_array_equal(x1[act], xyz)
    out_v = np.setdiff1d(np.arange(len(T0)), win)
    out_p = np.setdiff1d(np.arange(len(x0)), act)
    np.testing.assert_array_equal(T1[out_v], T0[out_v]); np.testing.assert_array_equal(x1[out_p], x0[out_p])
    assert np.abs(T1[win] - T0[win]).max() > 0
    # the next assembly of the same window starts where the optimiser stopped
    dm.set_problem(ba, win, act, pb.cam, c_i=pb.c_i, c_j=pb.c_j, c_T=pb.c_T, c_Lambda=pb.c_Lambda)
    np.testing.assert_array_equal(ba.poses(), poses)
    np.testing.assert_allclose(ba.points(), psi, rtol=1e-14)
    # host-side point write-back
    some = np.asarray(act[:5])
    dm.update_points(some, m["xyz_anchor"][some] + 0.25)
    np.testing.assert_array_equal(dm.get()[1][some], m["xyz_anchor"][some] + 0.25)
    dm.close(); ba.close()


def _load(dm, m):
    dm.set(m["poses"], m["point_anchor"], m["xyz_anchor"], m["vis_ptr"], m["vis_pose"], m["feat_center"], m["feat_level"])


def _all_observations(svs, dm, m):
    """Every observation of the device map, read through an assembly of the whole map (all vertices, all points)."""
    ba = svs.BundleAdjuster()
    V, Np = len(m["poses"]), len(m["point_anchor"])
    E = dm.set_problem(ba, np.arange(V), np.arange(Np), (500.0, 320.0, 240.0, 0.1))
    out = dm.last_edges(E)
    ba.close()
    return out


@pytest.mark.parametrize("P,L,seed,root_i,inner,double", [(12, 300, 3, 0, 3, 6), (40, 4000, 8, 17, 8, 20), (40, 4000, 9, 5, 15, 40)])
def test_window_selection_equals_the_restatement(svs, oracle, P, L, seed, root_i, inner, double):
    """computeInitialDoubleWin + computeActivePointsAndExtendOuterWindow + the pair loop of copyContraintsToG2o
    (slam_graph.cpp:556-663, 938-981) on the device against the plain-Python restatement: window (with INNER/OUTER
    types and the outer-window extension), active points and constraints are identical, bit for bit."""
    pb = synth.make_window(P, L, seed=seed)
    m, win0, _ = synth_graph.make_map(pb, extra_vertices=6, extra_points=40, seed=seed)
    ptr, ids, T, Lm = synth_graph.make_pose_graph(m, seed=seed)
    root = int(win0[root_i])
    dm = svs.DeviceMap()
    _load(dm, m)
    with pytest.raises(svs.SvsError):
        dm.select_window(root, inner, double)                     # no pose graph yet
    dm.set_graph(ptr, ids, T, Lm)
    got = dm.select_window(root, inner, double)
    w = oracle.compute_double_window(ptr, ids, root, inner, double)
    active, w = oracle.compute_active_points(m, ptr, ids, w)
    np.testing.assert_array_equal(got["window_vertex"], sorted(w))
    np.testing.assert_array_equal(got["inner"], [w[v] == 1 for v in sorted(w)])
    np.testing.assert_array_equal(got["active_point"], active)
    ci, cj, cT, cL = oracle.select_constraints(ptr, ids, T, Lm, w)
    np.testing.assert_array_equal(got["c_i"], ci); np.testing.assert_array_equal(got["c_j"], cj)
    np.testing.assert_array_equal(got["c_T"], cT); np.testing.assert_array_equal(got["c_Lambda"], cL)
    assert len(active) > 0 and len(ci) > 0
    with pytest.raises(svs.SvsError):
        dm.select_window(root, double, double)                    # assert(inner_window_size < double_window_size)
    # the selection feeds the assembly unchanged, and the assembled window optimises
    ba = svs.BundleAdjuster()
    fixed = (got["inner"] == 0).astype(np.uint8)
    fixed[0] = 1
    E = dm.set_problem(ba, got["window_vertex"], got["active_point"], pb.cam, fixed=fixed, c_i=got["c_i"], c_j=got["c_j"],
                       c_T=got["c_T"], c_Lambda=got["c_Lambda"])
    g = oracle.copy_data_to_g2o(m, got["window_vertex"], got["active_point"])
    assert E == len(g["e_point"])
    ep, es, ea, obs, info = dm.last_edges(E)
    np.testing.assert_array_equal(ep, g["e_point"]); np.testing.assert_array_equal(es, g["e_pose"])
    np.testing.assert_array_equal(obs, g["e_obs"])
    dm.close(); ba.close()


def test_map_grows_by_a_keyframe_on_the_device(svs, oracle):
    """SlamGraph::addKeyframe (slam_graph.cpp:144-186, 359-421) on the device tables against the restatement: the new
    vertex's pose is composed on the device, new points and the new keyframe's observations are in the rebuilt
    observation lists, everything else is unchanged; twice in a row."""
    pb = synth.make_window(10, 400, seed=4)
    m, win, act = synth_graph.make_map(pb, seed=4)
    dm = svs.DeviceMap()
    _load(dm, m)
    rng = np.random.default_rng(1)
    for step in range(2):
        V, Np = len(m["poses"]), len(m["point_anchor"])
        oldkey = int(rng.integers(0, V))
        dq = rng.normal(0, 0.05, 3)
        T = np.concatenate([dq, [np.sqrt(1 - dq @ dq)], rng.normal(0, 0.3, 3)])
        n_new, n_track = 25, 60
        kw = dict(new_anchor=rng.integers(0, V, n_new).astype(np.int32), new_xyz=np.stack([rng.uniform(-1, 1, n_new), rng.uniform(-1, 1, n_new), rng.uniform(2, 9, n_new)], 1),
                  new_anchor_center=rng.uniform(0, 600, (n_new, 3)), new_anchor_level=rng.integers(0, 3, n_new).astype(np.int32),
                  new_center=rng.uniform(0, 600, (n_new, 3)), new_level=rng.integers(0, 3, n_new).astype(np.int32),
                  track_point=rng.choice(Np, n_track, replace=False).astype(np.int32), track_center=rng.uniform(0, 600, (n_track, 3)),
                  track_level=rng.integers(0, 3, n_track).astype(np.int32))
        v, q = dm.add_keyframe(oldkey, T, **kw)
        assert (v, q) == (V, Np)
        m = oracle.add_keyframe(m, oldkey, T, **kw)
        Tm, xm = dm.get()
        np.testing.assert_array_equal(Tm[:V], m["poses"][:V])
        np.testing.assert_allclose(Tm[V], m["poses"][V], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(xm, m["xyz_anchor"])
        mm = dict(m); mm["poses"] = Tm
        ep, es, ea, obs, info = _all_observations(svs, dm, mm)
        g = oracle.copy_data_to_g2o(mm, np.arange(V + 1), np.arange(Np + n_new))
        np.testing.assert_array_equal(ep, g["e_point"]); np.testing.assert_array_equal(es, g["e_pose"])
        np.testing.assert_array_equal(ea, g["e_anchor"]); np.testing.assert_array_equal(obs, g["e_obs"])
        np.testing.assert_array_equal(info, g["e_info"])
        m = mm
    with pytest.raises(svs.SvsError):
        dm.add_keyframe(0, T, track_point=[3, 3], track_center=np.zeros((2, 3)), track_level=[0, 0])   # listed twice
    with pytest.raises(svs.SvsError):
        dm.add_keyframe(len(m["poses"]) + 5, T)                                                        # unknown old keyframe
    dm.close()
