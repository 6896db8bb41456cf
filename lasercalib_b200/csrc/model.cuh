// Camera model of the reference (lasercalib/pySBA.py:61-101) and its analytic
// derivatives, FP64, device side.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lcba {

constexpr int NCP = 11;       // rotvec(3) t(3) f k1 k2 cx cy  (pySBA.py:31-35)
// Per-camera table: [0..10] parameters, [11..19] R row-major, [20..28] right Jacobian J_r of
// SO(3) row-major:  d(R X)/dr = -R [X]x J_r.  Stride 29 doubles (odd) => conflict-free
// shared-memory reads when the lanes of a half-warp address different cameras; the table is
// read once per observation by every streaming pass, so it is kept small (shared-memory
// bandwidth, not FP64, was the limiter with the 27 doubles of explicit dR/dr_k).
constexpr int CT_R = 11;
constexpr int CT_JR = 20;
constexpr int CAMTAB = 29;

// R(r) = I + a K + b K^2,  J_r(r) = I - b K + c K^2,  K = [r]x,
// a = sin t / t, b = (1 - cos t)/t^2, c = (t - sin t)/t^3; Taylor series below t = 0.1
// (t = 0 => identity, like pySBA.py:66-68).
__device__ __forceinline__ void rodrigues_coeffs(double t2, double& a, double& b, double& c) {
  if (t2 < 0.01) {
    a = 1.0 - t2 / 6.0 * (1.0 - t2 / 20.0 * (1.0 - t2 / 42.0 * (1.0 - t2 / 72.0 * (1.0 - t2 / 110.0))));
    b = 0.5 * (1.0 - t2 / 12.0 * (1.0 - t2 / 30.0 * (1.0 - t2 / 56.0 * (1.0 - t2 / 90.0 * (1.0 - t2 / 132.0)))));
    c = 1.0 / 6.0 * (1.0 - t2 / 20.0 * (1.0 - t2 / 42.0 * (1.0 - t2 / 72.0 * (1.0 - t2 / 110.0 * (1.0 - t2 / 156.0)))));
  } else {
    double t = sqrt(t2), sn, cs;
    sincos(t, &sn, &cs);
    a = sn / t;
    b = (1.0 - cs) / t2;
    c = (t - sn) / (t * t2);
  }
}

// Build one camera table from its 11-vector. One thread per camera.
__device__ inline void cam_table_build(const double* __restrict__ cam, double* __restrict__ T) {
  for (int i = 0; i < NCP; ++i) T[i] = cam[i];
  const double r0 = cam[0], r1 = cam[1], r2 = cam[2];
  const double t2 = r0 * r0 + r1 * r1 + r2 * r2;
  double a, b, c;
  rodrigues_coeffs(t2, a, b, c);
  double K[9] = {0, -r2, r1, r2, 0, -r0, -r1, r0, 0};
  double K2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
  for (int i = 0; i < 9; ++i) {
    const double eye = (i % 4 == 0) ? 1.0 : 0.0;
    T[CT_R + i] = eye + a * K[i] + b * K2[i];
    T[CT_JR + i] = eye - b * K[i] + c * K2[i];
  }
}

// Residual and Jacobian blocks of one observation.
//   Jc columns 0..8 = (r0 r1 r2 t0 t1 t2 f k1 k2); columns 9,10 (cx, cy) are w*[1,0], w*[0,1].
struct ObsLin {
  double ru, rv;
  double Jc[2][9];
  double Jp[2][3];
};

// Pixel projection only (camera table T; point X).  Returns (pu, pv).
__device__ __forceinline__ void project_tab(const double* __restrict__ T, double X, double Y,
                                            double Z, double& pu, double& pv) {
  const double* R = T + CT_R;
  const double xc = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, T[3])));
  const double yc = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, T[4])));
  const double zc = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, T[5])));
  const double iz = 1.0 / zc;
  const double x = xc * iz, y = yc * iz;
  const double n = x * x + y * y;
  const double fd = T[6] * (1.0 + n * (T[7] + T[8] * n));
  pu = fma(fd, x, T[9]);
  pv = fma(fd, y, T[10]);
}

// `live` = false turns the observation into an exact zero (w must be 0 too): the perspective
// divide is masked so that no Inf/NaN can appear for a camera that does not see the point.
template <bool WITH_RES>
__device__ __forceinline__ void obs_linearize(const double* __restrict__ T, double X, double Y,
                                              double Z, double uo, double vo, double w,
                                              ObsLin& o, bool live = true) {
  const double* R = T + CT_R;
  const double xc = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, T[3])));
  const double yc = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, T[4])));
  const double zc = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, T[5])));
  const double iz = live ? 1.0 / zc : 0.0;
  const double x = xc * iz, y = yc * iz;
  const double n = x * x + y * y;
  const double f = T[6], k1 = T[7], k2 = T[8];
  const double d = 1.0 + n * (k1 + k2 * n);
  if (WITH_RES) {
    const double fd = f * d;
    o.ru = w * (fma(fd, x, T[9]) - uo);
    o.rv = w * (fma(fd, y, T[10]) - vo);
  }
  const double dp = k1 + 2.0 * k2 * n;
  const double wf = w * f;
  const double e00 = wf * (d + 2.0 * x * x * dp);
  const double e01 = wf * (2.0 * x * y * dp);
  const double e11 = wf * (d + 2.0 * y * y * dp);
  // G = w * d(u,v)/dXc
  const double g00 = e00 * iz, g01 = e01 * iz, g02 = -(e00 * x + e01 * y) * iz;
  const double g10 = e01 * iz, g11 = e11 * iz, g12 = -(e01 * x + e11 * y) * iz;
  // d(u,v)/dr_k = G * (-R (X x J_r[:,k]));  GR = G R is needed for Jp anyway
  double GR[2][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    GR[0][j] = fma(g00, R[j], fma(g01, R[3 + j], g02 * R[6 + j]));
    GR[1][j] = fma(g10, R[j], fma(g11, R[3 + j], g12 * R[6 + j]));
  }
  const double* Jr = T + CT_JR;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double c0 = Jr[k], c1 = Jr[3 + k], c2 = Jr[6 + k];
    const double m0 = fma(Y, c2, -Z * c1), m1 = fma(Z, c0, -X * c2), m2 = fma(X, c1, -Y * c0);
    o.Jc[0][k] = -fma(GR[0][0], m0, fma(GR[0][1], m1, GR[0][2] * m2));
    o.Jc[1][k] = -fma(GR[1][0], m0, fma(GR[1][1], m1, GR[1][2] * m2));
  }
  o.Jc[0][3] = g00; o.Jc[0][4] = g01; o.Jc[0][5] = g02;
  o.Jc[1][3] = g10; o.Jc[1][4] = g11; o.Jc[1][5] = g12;
  const double wd = w * d, wfn = wf * n, wfn2 = wfn * n;
  o.Jc[0][6] = wd * x;   o.Jc[1][6] = wd * y;
  o.Jc[0][7] = wfn * x;  o.Jc[1][7] = wfn * y;
  o.Jc[0][8] = wfn2 * x; o.Jc[1][8] = wfn2 * y;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o.Jp[0][j] = GR[0][j];
    o.Jp[1][j] = GR[1][j];
  }
}

// Row-wise Rodrigues rotation exactly as the reference writes it (pySBA.py:61-73):
// used by lcba_rotate / lcba_project where every ROW carries its own rotation vector.
__device__ __forceinline__ void rotate_row(double r0, double r1, double r2, double X, double Y,
                                           double Z, double& ox, double& oy, double& oz) {
  const double th = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  if (th > 0.0) { v0 = r0 / th; v1 = r1 / th; v2 = r2 / th; }
  double s, c;
  sincos(th, &s, &c);
  const double dot = X * v0 + Y * v1 + Z * v2;
  const double cx = v1 * Z - v2 * Y, cy = v2 * X - v0 * Z, cz = v0 * Y - v1 * X;
  const double k = dot * (1.0 - c);
  ox = c * X + s * cx + k * v0;
  oy = c * Y + s * cy + k * v1;
  oz = c * Z + s * cz + k * v2;
}

}  // namespace lcba
