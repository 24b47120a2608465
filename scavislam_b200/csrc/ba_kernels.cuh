// ba_kernels.cuh -- launchers of the BA kernels (ba_kernels.cu)
#pragma once
#include "ba_types.cuh"

namespace svs {
// true when kernel `slot` has not yet been given `bytes` of dynamic shared memory on the CURRENT device
// (cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute); thread-safe
bool device_needs_smem_optin(int slot, size_t bytes);
size_t build_smem_bytes(int warps, int Kmax);
void launch_prep(const BaDev& d, int buf, cudaStream_t st);
void launch_regroup(const BaDev& d, const double* raw, cudaStream_t st);
void launch_build(const BaDev& d, int Kmax, int robust, double delta, cudaStream_t st);
void launch_solve(const BaDev& d, int max_col_branch, int max_col_sep, int nsep, cudaStream_t st);
int solve_ring_capacity(int P, int nblk, int nsep);
void launch_solve_general(const BaDev& d, cudaStream_t st);
int update_grid_blocks(int L, int C);
void launch_update(const BaDev& d, int robust, double delta, int defer_decision, cudaStream_t st);
void launch_decide_deferred(const BaDev& d, cudaStream_t st);
void launch_chi2(const BaDev& d, int robust, double delta, cudaStream_t st);
void launch_export(const BaDev& d, double* out, cudaStream_t st);
}  // namespace svs
