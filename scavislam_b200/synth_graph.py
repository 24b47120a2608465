"""Feature tables and anchored points of a synthetic window, in the form SlamGraph::computeConstraint reads
them (reference slam_graph.hpp:65-137: Vertex::feature_table, Point::anchorframe_id / xyz_anchor)."""
import numpy as np


def graph_tables(pb):
    """pb: synth.BAProblem.  Returns dict(poses, feat_ptr, feat_point, point_anchor, xyz_anchor)."""
    P, L = pb.P, pb.L
    order = np.lexsort((pb.e_point, pb.e_pose))
    e_pose, e_point = pb.e_pose[order], pb.e_point[order]
    keep = np.ones(len(e_pose), bool)
    keep[1:] = (e_pose[1:] != e_pose[:-1]) | (e_point[1:] != e_point[:-1])
    e_pose, e_point = e_pose[keep], e_point[keep]
    feat_ptr = np.searchsorted(e_pose, np.arange(P + 1)).astype(np.int32)
    anchor = np.zeros(L, np.int32)
    anchor[pb.e_point] = pb.e_anchor
    psi = pb.psi
    xyz = np.stack([psi[:, 0] / psi[:, 2], psi[:, 1] / psi[:, 2], 1.0 / psi[:, 2]], 1)
    return dict(poses=pb.pose_qt.copy(), feat_ptr=feat_ptr, feat_point=e_point.astype(np.int32), point_anchor=anchor,
                xyz_anchor=xyz)


def make_map(pb, extra_vertices=4, extra_points=10, seed=0):
    """A map that contains the window `pb` plus vertices and points outside it: returns (map tables, window_vertex,
    active_point).  Vertex ids are a permutation of the BA pose indices; every point's observations are listed by
    ascending vertex id (Point::vis_set is a std::set<int>, reference slam_graph.hpp:102-137); pyramid levels are
    recovered from the window's information values."""
    rng = np.random.default_rng(seed)
    P, L = pb.P, pb.L
    V, Np = P + extra_vertices, L + extra_points
    perm = rng.permutation(V)
    window_vertex = perm[:P].astype(np.int32)                     # BA pose i lives in vertex window_vertex[i]
    outside = perm[P:]
    point_perm = rng.permutation(Np)
    active_point = point_perm[:L].astype(np.int32)                # BA point l is map point active_point[l]
    poses = np.zeros((V, 7)); poses[:, 3] = 1.0
    poses[window_vertex] = pb.pose_qt
    poses[outside, 4:] = rng.normal(0, 1, (len(outside), 3))
    xyz = np.stack([rng.uniform(-1, 1, Np), rng.uniform(-1, 1, Np), rng.uniform(2, 9, Np)], 1)
    psi = pb.psi
    xyz[active_point] = np.stack([psi[:, 0] / psi[:, 2], psi[:, 1] / psi[:, 2], 1.0 / psi[:, 2]], 1)
    anchor = np.full(Np, window_vertex[0], np.int32)
    a_of_l = np.zeros(L, np.int64); a_of_l[pb.e_point] = pb.e_anchor
    anchor[active_point] = window_vertex[a_of_l]
    level = np.where(pb.e_info[:, 0] > 0.5, 0, 1).astype(np.int32)   # s = (2^-level)^2 in {1, 1/4}
    rows = [(int(active_point[pb.e_point[e]]), int(window_vertex[pb.e_pose[e]]), pb.e_obs[e], int(level[e])) for e in range(pb.E)]
    for p in active_point[: max(1, L // 3)]:                       # observations from frames outside the window
        for v in outside[: 1 + int(p) % len(outside)]:
            rows.append((int(p), int(v), rng.uniform(0, 600, 3), int(rng.integers(0, 3))))
    for p in point_perm[L:]:                                       # points that are not active
        rows.append((int(p), int(window_vertex[0]), rng.uniform(0, 600, 3), 0))
    rows.sort(key=lambda r: (r[0], r[1]))
    vis_point = np.array([r[0] for r in rows])
    vis_ptr = np.searchsorted(vis_point, np.arange(Np + 1)).astype(np.int32)
    m = dict(poses=poses, point_anchor=anchor, xyz_anchor=xyz, vis_ptr=vis_ptr,
             vis_pose=np.array([r[1] for r in rows], np.int32), feat_center=np.array([r[2] for r in rows], np.float64),
             feat_level=np.array([r[3] for r in rows], np.int32))
    return m, window_vertex, active_point


def make_pose_graph(m, max_neighbours=6, seed=0, with_constraints=True):
    """A pose graph over the vertices of `m` in the form SlamGraph keeps it: two frames are neighbours when they share
    points; Vertex::neighbor_ids_ordered_by_strength holds ONE neighbour per strength value (it is a std::map keyed by
    the strength), visited from the strongest (reference slam_graph.hpp:88-99, slam_graph.cpp:584-590).  Returns
    (nbr_ptr, nbr_id, nbr_T, nbr_Lambda): symmetric lists, per directed entry a relative pose and an information matrix."""
    rng = np.random.default_rng(seed)
    V, Np = len(m["poses"]), len(m["point_anchor"])
    strength = {}
    for p in range(Np):
        obs = m["vis_pose"][m["vis_ptr"][p]:m["vis_ptr"][p + 1]]
        for i in range(len(obs)):
            for j in range(i + 1, len(obs)):
                a, b = int(obs[i]), int(obs[j])
                if a != b:
                    strength[(min(a, b), max(a, b))] = strength.get((min(a, b), max(a, b)), 0) + 1
    cand = [dict() for _ in range(V)]
    for (a, b), s in strength.items():
        cand[a][b] = s; cand[b][a] = s
    keep = set()
    for v in range(V):
        by_strength = {}
        for b in sorted(cand[v]):                                  # one neighbour per strength value (std::map<int,int>)
            by_strength[cand[v][b]] = b
        for s in sorted(by_strength, reverse=True)[:max_neighbours]:
            keep.add((min(v, by_strength[s]), max(v, by_strength[s])))
    nbrs = [[] for _ in range(V)]
    for a, b in keep:
        nbrs[a].append((strength[(a, b)], b)); nbrs[b].append((strength[(a, b)], a))
    nbr_ptr, nbr_id = [0], []
    for v in range(V):
        for s, b in sorted(nbrs[v], key=lambda t: (-t[0], t[1])):  # strongest first
            nbr_id.append(b)
        nbr_ptr.append(len(nbr_id))
    nbr_ptr, nbr_id = np.array(nbr_ptr, np.int32), np.array(nbr_id, np.int32)
    if not with_constraints:
        return nbr_ptr, nbr_id, None, None
    n = len(nbr_id)
    q = rng.normal(0, 1, (n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    nbr_T = np.concatenate([q, rng.normal(0, 1, (n, 3))], 1)
    A = rng.normal(0, 1, (n, 6, 6))
    nbr_L = (A @ A.transpose(0, 2, 1) + 6 * np.eye(6)).reshape(n, 36)
    return nbr_ptr, nbr_id, nbr_T, nbr_L
