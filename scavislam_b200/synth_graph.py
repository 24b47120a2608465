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
