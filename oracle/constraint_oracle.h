/* constraint_oracle.h -- CPU restatement of SlamGraph::computeConstraint (slam_graph.cpp:785-846) for a list
 * of pose pairs.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference has no test for it); pinned by an
 * independent numpy computation (tests/test_constraint_oracle.py). */
#ifndef SVS_CONSTRAINT_ORACLE_H
#define SVS_CONSTRAINT_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif
/* poses: T_me_from_world[P][7]; feat_ptr[P+1] / feat_point: feature_table keys of every pose, ascending;
 * point_anchor[L]: index of the anchor pose of each point (into poses), xyz_anchor[L][3];
 * pairs (v1[k], v2[k]) -> T_1_from_2[k][7], Lambda[k][36] (row-major), visibility_strength[k].
 * A pair without shared points gets Lambda = 0 (the reference would take the median of an empty set). */
void occ_compute_constraints(int P, const double *poses, const int *feat_ptr, const int *feat_point, int L,
                             const int *point_anchor, const double *xyz_anchor, int npairs, const int *v1,
                             const int *v2, double *T_1_from_2, double *Lambda, int *visibility_strength);
#ifdef __cplusplus
}
#endif
#endif
