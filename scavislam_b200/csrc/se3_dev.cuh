// se3_dev.cuh -- FP64 device SE3/SO3 arithmetic with Sophus a621ff semantics
// (unit quaternion + translation, tangent = (upsilon, omega)).  Call sites in the
// reference: g2o_types/anchored_points.cpp:57,164-165,178-188,209-223.
#pragma once
#include <cuda_runtime.h>

namespace svs {

constexpr double kSmallEps = 1e-10;

// q = x y z w  ->  row-major R (Eigen::Quaterniond::toRotationMatrix)
__device__ __forceinline__ void quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double c[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  c[3] = aw * bw - ax * bx - ay * by - az * bz;
  c[0] = aw * bx + ax * bw + ay * bz - az * by;
  c[1] = aw * by + ay * bw + az * bx - ax * bz;
  c[2] = aw * bz + az * bw + ax * by - ay * bx;
}

__device__ __forceinline__ void mat3_vec(const double R[9], const double x[3], double y[3]) {
  y[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  y[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  y[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
}

__device__ __forceinline__ void hat3(const double v[3], double M[9]) {
  M[0] = 0;     M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2];  M[4] = 0;     M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0];  M[8] = 0;
}

__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// SE3::exp (Sophus): T = (SO3::exp(omega), V * upsilon)
__device__ inline void se3_exp(const double d[6], double T[7]) {
  const double* om = d + 3;
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double half = 0.5 * theta;
  double imag;
  const double real = cos(half);
  if (theta < kSmallEps) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = sin(half) / theta;
  }
  const double q[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  double Om[9], Om2[9], V[9];
  hat3(om, Om);
  mat3_mul(Om, Om, Om2);
  if (theta < kSmallEps) {
    quat_to_R(q, V);
  } else {
    const double t2 = theta * theta;
    const double a = (1 - cos(theta)) / t2, b = (theta - sin(theta)) / (t2 * theta);
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = a * Om[i] + b * Om2[i];
    V[0] += 1; V[4] += 1; V[8] += 1;
  }
  double t[3];
  mat3_vec(V, d, t);
  T[0] = q[0]; T[1] = q[1]; T[2] = q[2]; T[3] = q[3];
  T[4] = t[0]; T[5] = t[1]; T[6] = t[2];
}

// A * B with the quaternion renormalised (Sophus SO3::operator*=)
__device__ inline void se3_mul(const double A[7], const double B[7], double AB[7]) {
  double R[9], t[3], q[4];
  quat_to_R(A, R);
  mat3_vec(R, B + 4, t);
  quat_mul(A, B, q);
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  AB[0] = q[0] / n; AB[1] = q[1] / n; AB[2] = q[2] / n; AB[3] = q[3] / n;
  AB[4] = A[4] + t[0]; AB[5] = A[5] + t[1]; AB[6] = A[6] + t[2];
}

__device__ inline void se3_inv(const double A[7], double Ai[7]) {
  const double q[4] = {-A[0], -A[1], -A[2], A[3]};
  const double mt[3] = {-A[4], -A[5], -A[6]};
  double R[9], t[3];
  quat_to_R(q, R);
  mat3_vec(R, mt, t);
  Ai[0] = q[0]; Ai[1] = q[1]; Ai[2] = q[2]; Ai[3] = q[3];
  Ai[4] = t[0]; Ai[5] = t[1]; Ai[6] = t[2];
}

// SE3::log (Sophus, atan-based SO3::log)
__device__ inline void se3_log(const double T[7], double d[6]) {
  const double n = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
  const double w = T[3];
  double k;
  if (n < kSmallEps) {
    k = 2. / w - 2. * (n * n) / (w * w * w);
  } else if (fabs(w) < kSmallEps) {
    k = (w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
  } else {
    k = 2 * atan(n / w) / n;
  }
  const double theta = k * n;
  const double om[3] = {k * T[0], k * T[1], k * T[2]};
  double Om[9], Om2[9], Vi[9];
  hat3(om, Om);
  mat3_mul(Om, Om, Om2);
  double c;
  if (theta < kSmallEps) c = 1. / 12.;
  else c = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
#pragma unroll
  for (int i = 0; i < 9; ++i) Vi[i] = -0.5 * Om[i] + c * Om2[i];
  Vi[0] += 1; Vi[4] += 1; Vi[8] += 1;
  mat3_vec(Vi, T + 4, d);
  d[3] = om[0]; d[4] = om[1]; d[5] = om[2];
}

// SE3::Adj = [[R, hat(t) R], [0, R]]
__device__ inline void se3_adj(const double A[7], double Adj[36]) {
  double R[9], tx[9], tR[9];
  quat_to_R(A, R);
  hat3(A + 4, tx);
  mat3_mul(tx, R, tR);
#pragma unroll
  for (int i = 0; i < 36; ++i) Adj[i] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Adj[i * 6 + j] = R[i * 3 + j];
      Adj[(i + 3) * 6 + (j + 3)] = R[i * 3 + j];
      Adj[i * 6 + (j + 3)] = tR[i * 3 + j];
    }
}

// g2o RobustKernelHuber::robustify: rho0 (cost) and rho1 (weight)
__device__ __forceinline__ void huber(double e2, double delta, double& rho0, double& rho1) {
  const double dsqr = delta * delta;
  if (e2 <= dsqr) {
    rho0 = e2; rho1 = 1.;
  } else {
    const double sq = sqrt(e2);
    rho0 = 2 * sq * delta - dsqr;
    rho1 = delta / sq;
  }
}

}  // namespace svs
