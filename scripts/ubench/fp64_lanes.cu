// Does a DFMA with fewer active lanes occupy the FP64 pipe for less time?  4 warps on one
// sub-partition (warps 0,4,8,12 of a 512-thread CTA) issue independent DFMAs with `active` lanes on.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, int active, int nw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((warp & 3) != 0 || (warp >> 2) >= nw) return;
  double acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  const long long t0 = clock64();
  if (lane < active)
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fma(acc[q], 1.0000001, 0.5);
  const long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 8; ++q) s += acc[q];
  if (s == 1.2345) out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  const int act[5] = {32, 17, 16, 8, 1};
  for (int nw = 1; nw <= 4; nw *= 2)
    for (int a = 0; a < 5; ++a) {
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 512>>>(out, cyc, 2000, act[a], nw); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps on the sub-partition %d  active lanes %2d  cycles per DFMA (this warp) %.2f\n", nw, act[a], h / 16000.0);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
