// internal.cuh -- cross-module hooks inside libsvsb200.so (not part of the C ABI)
#pragma once
#include "../../include/svs_b200.h"

namespace svs {
// device-resident results of the last svs_match on this handle (n = number of candidate points)
__attribute__((visibility("hidden"))) void matcher_device_results(svs_matcher* m, const svs_match_result** d_res, int* n,
                                                                   int* device);
// svs_ba_set_problem with the observations [E][3] and weights [E][3] (one buffer, in this order) already on the
// BA handle's device in the caller's edge order; returns after the handle has consumed the buffer
__attribute__((visibility("hidden"))) int ba_set_problem_device_obs(
    svs_ba* h, int P, const double* T_qt, const unsigned char* fixed, int L, const double* psi, int E, const int* e_point,
    const int* e_pose, const int* e_anchor, const double* d_obs_info, int C, const int* c_i, const int* c_j, const double* c_T,
    const double* c_Lambda, const svs_cam* cam);
__attribute__((visibility("hidden"))) int ba_device(const svs_ba* h);
// the accepted state where it lies on the BA handle's device: pose[2][P][7], psi[2][L][3] (internal landmark order),
// lm_user[L] (internal -> caller's landmark), *cur = index of the accepted buffers, the handle's stream
__attribute__((visibility("hidden"))) int ba_state_on_device(svs_ba* h, const double* const** pose, const double* const** psi,
                                                             const int** lm_user, const int** cur, cudaStream_t* stream, int* P, int* L);
// keypoints of the last svs_fast_detect* call on this handle, on its device: xy [n][2] in cell order, cell_off [ncells + 1]
__attribute__((visibility("hidden"))) void fast_device_results(svs_fast* f, const int** d_xy, const int** d_cell_off, int* ncells,
                                                                int* n, int* device);
}  // namespace svs
