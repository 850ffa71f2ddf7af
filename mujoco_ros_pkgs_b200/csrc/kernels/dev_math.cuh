// dev_math.cuh — FP64 device helpers for the batched step kernels (sm_100a).
// Conventions as MuJoCo 2.3.7 (the engine behind the reference's mj_step call, mujoco_env.cpp:498):
// quaternions (w,x,y,z), row-major 3x3, spatial vectors [angular; linear], 10-number spatial inertia.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "dev_model.h"

namespace b2k {

#define B2K_MINVAL 1E-15
#define B2K_DI __device__ __forceinline__

B2K_DI void copy3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
B2K_DI void copy4(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = a[3]; }
B2K_DI void zero3(double* r) { r[0] = 0; r[1] = 0; r[2] = 0; }
B2K_DI void scl3(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
B2K_DI void add3(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
B2K_DI void sub3(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
B2K_DI void addTo3(double* r, const double* a) { r[0] += a[0]; r[1] += a[1]; r[2] += a[2]; }
B2K_DI void addToScl3(double* r, const double* a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
B2K_DI double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
B2K_DI double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
B2K_DI void cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
B2K_DI double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
B2K_DI double normalize3(double* a) {
  double n = norm3(a);
  if (n < B2K_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { double s = 1 / n; a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
// unit quaternions are left untouched so rest states stay bitwise fixed points
B2K_DI double normalize4(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < B2K_MINVAL) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
  else if (fabs(n - 1) > B2K_MINVAL) { double s = 1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
  return n;
}
B2K_DI void mulQuat(double* r, const double* a, const double* b) {
  double t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
B2K_DI void mulQuatAxis(double* r, const double* q, const double* ax) {
  double t0 = -q[1] * ax[0] - q[2] * ax[1] - q[3] * ax[2];
  double t1 = q[0] * ax[0] + q[2] * ax[2] - q[3] * ax[1];
  double t2 = q[0] * ax[1] + q[3] * ax[0] - q[1] * ax[2];
  double t3 = q[0] * ax[2] + q[1] * ax[1] - q[2] * ax[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
B2K_DI void negQuat(double* r, const double* q) { r[0] = q[0]; r[1] = -q[1]; r[2] = -q[2]; r[3] = -q[3]; }
B2K_DI void quat2Mat(double* m, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
B2K_DI void rotVecMat(double* r, const double* v, const double* m) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
B2K_DI void rotVecMatT(double* r, const double* v, const double* m) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
B2K_DI void rotVecQuat(double* r, const double* v, const double* q) {
  if (v[0] == 0 && v[1] == 0 && v[2] == 0) { zero3(r); return; }
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) { copy3(r, v); return; }
  double m[9];
  quat2Mat(m, q);
  rotVecMat(r, v, m);
}
B2K_DI void axisAngle2Quat(double* q, const double* axis, double angle) {
  if (angle == 0) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; return; }
  double s, c;
  sincos(angle * 0.5, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
B2K_DI void quat2Vel(double* res, const double* q, double dt) {
  double axis[3] = {q[1], q[2], q[3]};
  double sin_a_2 = normalize3(axis);
  double speed = 2 * atan2(sin_a_2, q[0]);
  if (speed > 3.14159265358979323846) speed -= 2 * 3.14159265358979323846;
  speed /= dt;
  scl3(res, axis, speed);
}
B2K_DI void subQuat(double* res, const double* qa, const double* qb) {
  double qneg[4], qdif[4];
  negQuat(qneg, qb);
  mulQuat(qdif, qneg, qa);
  quat2Vel(res, qdif, 1);
}
B2K_DI void quatIntegrate(double* quat, const double* vel, double scale) {
  double tmp[3], qrot[4], qres[4];
  copy3(tmp, vel);
  double angle = scale * normalize3(tmp);
  axisAngle2Quat(qrot, tmp, angle);
  normalize4(quat);
  mulQuat(qres, quat, qrot);
  copy4(quat, qres);
}
B2K_DI void inertCom(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  double t0 = mat[0] * inert[0], t3 = mat[1] * inert[1], t6 = mat[2] * inert[2];
  double t1 = mat[3] * inert[0], t4 = mat[4] * inert[1], t7 = mat[5] * inert[2];
  double t2 = mat[6] * inert[0], t5 = mat[7] * inert[1], t8 = mat[8] * inert[2];
  res[0] = mat[0] * t0 + mat[1] * t3 + mat[2] * t6;
  res[1] = mat[3] * t1 + mat[4] * t4 + mat[5] * t7;
  res[2] = mat[6] * t2 + mat[7] * t5 + mat[8] * t8;
  res[3] = mat[0] * t1 + mat[1] * t4 + mat[2] * t7;
  res[4] = mat[0] * t2 + mat[1] * t5 + mat[2] * t8;
  res[5] = mat[3] * t2 + mat[4] * t5 + mat[5] * t8;
  res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] -= mass * dif[0] * dif[1];
  res[4] -= mass * dif[0] * dif[2];
  res[5] -= mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
B2K_DI void mulInertVec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
B2K_DI void crossMotion(double* r, const double* vel, const double* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5];
  r[4] = vel[2] * v[3] - vel[0] * v[5];
  r[5] = -vel[1] * v[3] + vel[0] * v[4];
  r[3] += -vel[5] * v[1] + vel[4] * v[2];
  r[4] += vel[5] * v[0] - vel[3] * v[2];
  r[5] += -vel[4] * v[0] + vel[3] * v[1];
}
B2K_DI void crossForce(double* r, const double* vel, const double* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2];
  r[1] = vel[2] * f[0] - vel[0] * f[2];
  r[2] = -vel[1] * f[0] + vel[0] * f[1];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
  r[0] += -vel[5] * f[4] + vel[4] * f[5];
  r[1] += vel[5] * f[3] - vel[3] * f[5];
  r[2] += -vel[4] * f[3] + vel[3] * f[4];
}
B2K_DI void makeFrame(double* frame) {
  if (normalize3(frame) < 0.5) { frame[0] = 1; frame[1] = 0; frame[2] = 0; }
  double* y = frame + 3;
  if (norm3(y) < 0.5) {
    zero3(y);
    if (frame[1] < 0.5 && frame[1] > -0.5) y[1] = 1;
    else y[2] = 1;
  }
  double d = dot3(frame, y);
  addToScl3(y, frame, -d);
  normalize3(y);
  cross(frame + 6, frame, y);
}
B2K_DI double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// spatial transform of a [angular; linear] motion / force vector between reference points
B2K_DI void transformSpatial(double* res, const double* vec, int flg_force, const double* newpos, const double* oldpos,
                             const double* rotnew2old) {
  double dif[3], cr[3], tran[6];
  for (int k = 0; k < 6; k++) tran[k] = vec[k];
  sub3(dif, newpos, oldpos);
  if (flg_force) { cross(cr, dif, vec + 3); sub3(tran, vec, cr); }
  else { cross(cr, dif, vec); sub3(tran + 3, vec + 3, cr); }
  if (rotnew2old) { rotVecMatT(res, tran, rotnew2old); rotVecMatT(res + 3, tran + 3, rotnew2old); }
  else { for (int k = 0; k < 6; k++) res[k] = tran[k]; }
}

// ---- env-group collectives.  An env is served by a group of B2K_G consecutive lanes (32: one env per warp; 16: two envs per
// warp, each SIMT instruction doing the work of two envs -- see dev_model.h for the measured trade-off).
// `mask` names the lanes of the caller's group; the halves of a warp may diverge and reconverge freely.
B2K_DI double warpSum(unsigned mask, double v) {
#pragma unroll
  for (int o = B2K_G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, B2K_G);
  return v;
}
B2K_DI int warpMaxInt(unsigned mask, int v) {
#pragma unroll
  for (int o = B2K_G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(mask, v, o, B2K_G));
  return v;
}
B2K_DI int warpInclusiveScan(unsigned mask, int v, int lane) {
#pragma unroll
  for (int o = 1; o < B2K_G; o <<= 1) {
    int t = __shfl_up_sync(mask, v, o, B2K_G);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace b2k
